"""-m gpu: the multi-GPU path behind the C ABI (rt_comm_init / rt_gather_frame, include/rtcore_b200.h) - VERDICT r1 "next" 3.

On one GPU: a world-of-1 communicator exercises the run-time NCCL binding, the scatter of the root's own payload, the gathered
downloads and the present of a gathered frame.  With >= 2 GPUs visible (gpurun --gpus N): tests/multirank_worker.py under
torchrun - real NCCL send / recv between processes, the gathered image against a single-context render on rank 0."""
import os
import subprocess
import sys

import numpy as np
import pytest

from ilgpu_raytracing_b200 import layouts as L
from ilgpu_raytracing_b200 import native
from oracle import orc
from tests.util import oracle_camera

pytestmark = pytest.mark.gpu


def test_world_of_one_communicator_and_gathered_frame(gpu_ctx):
    W, H = 200, 120
    sc = orc.Scene()
    sc.build_default()
    ctx = native.Context(0)
    try:
        ctx.scene_upload(sc.arrays())
        cam = oracle_camera("C1B", W, H)
        ctx.render(cam, L.make_render_config(W, H, spp=2, max_depth=3))
        with pytest.raises(native.RtError) as e:
            ctx.gather_frame(0)                                  # no communicator yet
        assert e.value.status == L.RT_ERR_INVALID_STATE
        ctx.comm_init(native.comm_unique_id(), 0, 1)
        with pytest.raises(native.RtError) as e:
            ctx.comm_init(native.comm_unique_id(), 0, 1)         # already has one
        assert e.value.status == L.RT_ERR_INVALID_STATE
        with pytest.raises(native.RtError) as e:
            ctx.gather_frame(0, L.RT_GATHER_DEPTH_OBJID)         # neither colour format named
        assert e.value.status == L.RT_ERR_INVALID_ARGUMENT
        with pytest.raises(native.RtError) as e:
            ctx.download(L.RT_BUF_GATHERED_RGBA8)                # nothing gathered yet
        assert e.value.status == L.RT_ERR_INVALID_STATE
        ctx.gather_frame(0, L.RT_GATHER_RADIANCE | L.RT_GATHER_DEPTH_OBJID)
        ctx.sync()
        assert np.array_equal(ctx.download(L.RT_BUF_GATHERED_RGBA8), ctx.download(L.RT_BUF_RGBA8))
        assert np.array_equal(ctx.download(L.RT_BUF_GATHERED_DEPTH), ctx.download(L.RT_BUF_DEPTH))
        assert np.array_equal(ctx.download(L.RT_BUF_GATHERED_OBJID), ctx.download(L.RT_BUF_OBJID))
        assert np.array_equal(ctx.download(L.RT_BUF_GATHERED_RADIANCE), ctx.download(L.RT_BUF_RADIANCE))
        ref = orc.render(sc, cam, orc.make_config(W, H, spp=2, max_depth=3), aovs=False)
        assert np.array_equal(ctx.download(L.RT_BUF_GATHERED_RGBA8), ref.rgba8)
        ctx.gather_frame(0, L.RT_GATHER_RGBA8)                   # display-only gather: depth / objId are not part of it
        with pytest.raises(native.RtError) as e:
            ctx.download(L.RT_BUF_GATHERED_DEPTH)
        assert e.value.status == L.RT_ERR_INVALID_STATE
        assert np.array_equal(ctx.download(L.RT_BUF_GATHERED_RGBA8), ref.rgba8)
        with pytest.raises(native.RtError) as e:                 # a frame rendered as another partition than the communicator's
            ctx.render(cam, L.make_render_config(W, H, spp=2, max_depth=3, rank=1, world_size=2))
            ctx.gather_frame(0)
        assert e.value.status == L.RT_ERR_INVALID_STATE
        ctx.comm_destroy()
    finally:
        ctx.close()


def test_engine_mirror_multi_gpu_flow_world_of_one(gpu_ctx):
    """RTRenderer.InitMultiGpu + RenderDirectToPbo: render -> rt_gather_frame -> present on the root -> Framebuffer.DownloadToCpu
    reads the GATHERED colour / depth / objectId (same values as the single-GPU flow)."""
    from ilgpu_raytracing_b200 import engine
    W, H = 240, 136
    imgs = []
    for multi in (False, True):
        rdr = engine.RTRenderer(0, W, H)
        rdr.camera = engine.config_camera("C1B", W, H)
        rdr.configure(renderScale=0.67, enableTAAU=1, enableTemporalReuse=0, enableSpatialReuse=0, spp=2, maxDepth=3, rngLockNoise=1, fixedSeed=3)
        if multi:
            rdr.InitMultiGpu(engine.RTRenderer.NewCommunicatorId(), 0, 1)
        for frame in range(2):
            rdr.RenderDirectToPbo(None, W, H, frame, 0.016)
        color, depth, objid = rdr.DownloadToCpu()
        imgs.append((color.copy(), depth.copy(), objid.copy(), rdr.native.download(L.RT_BUF_PRESENT).copy()))
        rdr.close()
    for a, b in zip(*imgs):
        assert np.array_equal(a, b)


def test_native_gather_across_processes():
    """Real NCCL between processes (needs >= 2 GPUs: `gpurun --gpus 2`): three gather modes, overlapped back-to-back frames,
    gathered colour / depth / objId / radiance and the TAAU present equal to a single-context render."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(root, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_gl_interop_entry_points_without_a_gl_context(gpu_ctx):
    """rt_gl_register_buffer / rt_gl_map / rt_gl_unmap / rt_gl_unregister replace the reference's driver-API binding
    (Engine/CudaGlInteropIndexBuffer.cs:18-34,44-103).  There is no GL context on the test box, so only the error contract can be
    exercised here: the driver's refusal comes back as a status + message, nothing throws or crashes, null arguments are caught."""
    import ctypes as C
    l = gpu_ctx._l
    res = C.c_void_p()
    rc = l.rt_gl_register_buffer(gpu_ctx.h, 12345, C.byref(res))
    assert rc in (L.RT_ERR_CUDA, L.RT_ERR_UNSUPPORTED) and not res.value
    msg = (l.rt_last_error() or b"").decode()
    assert "cuGraphicsGLRegisterBuffer" in msg or "not available" in msg
    assert l.rt_gl_register_buffer(gpu_ctx.h, 1, None) == L.RT_ERR_INVALID_ARGUMENT
    p, n = C.c_void_p(), C.c_size_t()
    assert l.rt_gl_map(gpu_ctx.h, None, C.byref(p), C.byref(n)) == L.RT_ERR_INVALID_ARGUMENT
    assert l.rt_gl_unmap(gpu_ctx.h, None) == L.RT_ERR_INVALID_ARGUMENT
    assert l.rt_gl_unregister(gpu_ctx.h, None) == L.RT_OK          # nothing to unregister
