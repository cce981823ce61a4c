"""ncu target (manual tool under gpurun): rank r of N's share of a C4 frame on one GPU.  python tests/gpu_frame_rank.py N rank spp frames"""
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import layouts as L, native, scenes  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

N, rank, spp, frames = (int(v) for v in (sys.argv[1:5] + ["8", "0", "64", "2"][len(sys.argv) - 1:]))
W, H = 3840, 2160
sc = oracle_scene_from_spec(scenes.terrain_scene(n_quads=708, n_spheres=256))
ctx = native.Context(0)
ctx.scene_upload(sc.arrays())
cam = oracle_camera("C3", W, H)
cfg = L.make_render_config(W, H, spp=spp, max_depth=8, rank=rank, world_size=N)
for _ in range(frames):
    ctx.render(cam, cfg); ctx.sync(); s = ctx.stats()
    print(s["lastRenderMs"], s["kernelLaunches"])
