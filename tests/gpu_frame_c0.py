"""ncu target (manual tool under gpurun): workload C0 - the reference's operating point (1280x720 window traced at 858x482, 2 spp,
depth 3, ReSTIR reuse on, TAAU) through RTRenderer.RenderDirectToPbo for N frames.  argv: frames [graph 0|1]."""
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
from ilgpu_raytracing_b200 import engine, layouts as L  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
graph = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W, H = 1280, 720
rdr = engine.RTRenderer(0, W, H)
rdr.configure(renderScale=0.67, enableTAAU=1, enableTemporalReuse=1, enableSpatialReuse=1, spp=2, maxDepth=3, rngLockNoise=1, fixedSeed=1,
              flags=L.RT_FLAG_FRAME_GRAPH if graph else 0)
pbo = torch.zeros(W * H, dtype=torch.int32, device="cuda")
for f in range(frames):
    cam = engine.config_camera("C1B", W, H)
    rdr.camera = engine.camera_translate(cam, 0.004 * f, 0.001 * f, -0.003 * f)
    rdr.RenderDirectToPbo(pbo.data_ptr(), W, H, f, 0.016)
    s = rdr.native.stats()
    print(f, s["lastRenderMs"], s["kernelLaunches"], s["raysPrimary"], s["raysBounce"], s["raysShadow"])
rdr.close()
