import sys, time
sys.path.insert(0, ".")
import torch
from ilgpu_raytracing_b200 import engine, layouts as L
W, H = 1280, 720
for variant in ("own_stream", "own_stream_graph", "torch_stream", "torch_stream_graph"):
    rdr = engine.RTRenderer(0, W, H)
    rdr.configure(renderScale=0.67, enableTAAU=1, enableTemporalReuse=1, enableSpatialReuse=1, spp=2, maxDepth=3, rngLockNoise=1, fixedSeed=1, asyncSubmit=1, flags=(L.RT_FLAG_FRAME_GRAPH if variant.endswith('graph') else 0))
    ctx = rdr.native
    stream = torch.cuda.Stream()
    if variant.startswith("torch_stream"):
        ctx.set_stream(stream.cuda_stream)
    pbo = torch.zeros(W * H, dtype=torch.int32, device="cuda")
    cams = [engine.camera_translate(engine.config_camera("C1B", W, H), 0.004 * f, 0.001 * f, -0.003 * f) for f in range(2300)]
    ptr = pbo.data_ptr()
    def step(f):
        rdr.camera = cams[f]
        rdr.RenderDirectToPbo(ptr, W, H, f, 0.016)
    for f in range(100): step(f)
    rdr.Synchronize(); torch.cuda.synchronize()
    for n in (200, 2000):
        t0 = time.perf_counter()
        for f in range(100, 100 + n): step(f)
        t_submit = time.perf_counter() - t0
        rdr.Synchronize(); torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print(variant, n, "submit us/frame", 1e6 * t_submit / n, "total us/frame", 1e6 * t_all / n, flush=True)
    # sync per frame
    rdr.configure(asyncSubmit=0)
    t0 = time.perf_counter()
    for f in range(100, 600): step(f)
    print(variant, "sync us/frame", 1e6 * (time.perf_counter() - t0) / 500, "device ms", ctx.stats()["lastRenderMs"], flush=True)
    rdr.close()
