"""Short GPU workload for ncu captures (manual tool under gpurun): C3 frame twice, then a 2 spp / depth-4 frame."""
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import build, layouts as L, native, scenes  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

build.build_core()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 708
sc = oracle_scene_from_spec(scenes.terrain_scene(n_quads=nq, n_spheres=256))
ctx = native.Context(0)
ctx.scene_upload(sc.arrays())
W, H = 3840, 2160
cam = oracle_camera("C3", W, H)
for spp, depth in ((1, 0), (1, 0), (2, 4)):
    ctx.render(cam, L.make_render_config(W, H, spp=spp, max_depth=depth)); ctx.sync()
    s = ctx.stats(); print(spp, depth, s["lastRenderMs"], s["raysPrimary"], s["raysBounce"], s["raysShadow"])
