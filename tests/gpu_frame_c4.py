"""ncu target (manual tool under gpurun): the bench workload C4 (1M triangles + 256 spheres, 4K, depth 8) for N frames at a given spp."""
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import engine, layouts as L  # noqa: E402
import bench  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flags = int(sys.argv[3], 0) if len(sys.argv) > 3 else 0   # e.g. 0x100 = RT_FLAG_FAST_SHADING
W, H = 3840, 2160
rdr = engine.RTRenderer(0, W, H)
rdr.scene.load_spec(bench.make_spec("terrain+spheres"))
rdr.Commit()
cam = engine.config_camera("C3", W, H)
ctx = rdr.native
cfg = L.make_render_config(W, H, spp=spp, max_depth=8, flags=flags)
for _ in range(frames):
    ctx.render(cam, cfg); ctx.sync(); s = ctx.stats()
    print(s["lastRenderMs"], s["kernelLaunches"], s["raysPrimary"], s["raysBounce"], s["raysShadow"])
