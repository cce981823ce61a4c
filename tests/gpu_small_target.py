"""Small GPU workload for compute-sanitizer / ncu (manual tool under gpurun): terrain(64 quads)+spheres, 320x180, 2 spp, depth 4, plus a reuse frame and a present."""
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import layouts as L, native, scenes  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

sc = oracle_scene_from_spec(scenes.terrain_scene(n_quads=64, n_spheres=16))
ctx = native.Context(0)
ctx.scene_upload(sc.arrays())
W, H = 320, 180
cam = oracle_camera("C3", W, H)
for frame in range(2):
    ctx.render(cam, L.make_render_config(W, H, spp=2, max_depth=4, frame=frame, temporal=1, spatial=1, flags=L.RT_FLAG_PATH_AOVS), prev_cam=cam); ctx.sync()
ctx.present(480, 270, taau=True); ctx.sync()
s = ctx.stats(); print(s["lastRenderMs"], s["raysPrimary"], s["raysBounce"], s["raysShadow"])
ctx.close()
