"""ctypes binding of the C ABI (include/rtcore_b200.h) — the same entry points the C# P/Invoke layer binds.

The library is loaded from the package directory.  There is no fallback of any kind: a missing
library or a machine without an sm_100 CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import layouts as L

_PKG = os.path.dirname(os.path.abspath(__file__))
# RTCORE_B200_LIB: developer override used by the tuning sweeps under tests/ (an alternative BUILD of the same CUDA library)
_LIB_PATH = os.environ.get("RTCORE_B200_LIB") or os.path.join(_PKG, "librtcore_b200.so")

EXPORTS = ["rt_abi_version", "rt_last_error", "rt_create", "rt_destroy", "rt_set_stream", "rt_scene_upload", "rt_render", "rt_sync",
           "rt_download", "rt_buffer_bytes", "rt_get_device_buffer", "rt_map_external_color", "rt_tiles_owned_pixels",
           "rt_deinterleave_tiles", "rt_get_stats", "rt_present", "rt_scene_refit", "rt_scene_upload_ex", "rt_download_async",
           "rt_comm_get_unique_id", "rt_comm_init", "rt_comm_destroy", "rt_gather_frame",
           "rt_gl_register_buffer", "rt_gl_map", "rt_gl_unmap", "rt_gl_unregister", "rt_get_stream", "rt_bind_readback"]


class RtError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"rtcore_b200 status {status}: {message}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). The renderer core has no CPU or pure-Python fallback.")
    l = C.CDLL(_LIB_PATH)
    if os.environ.get("RTCORE_B200_LIB"):   # A/B builds of older sources (tests/gpu_variants.py) may lack the newest entry points: give them inert stand-ins
        for name in EXPORTS:
            if not hasattr(l, name):
                setattr(l, name, C.CFUNCTYPE(C.c_int)(lambda *a: L.RT_ERR_UNSUPPORTED))
    l.rt_abi_version.restype = C.c_int
    l.rt_last_error.restype = C.c_char_p
    l.rt_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
    l.rt_destroy.argtypes = [C.c_void_p]
    l.rt_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    l.rt_get_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    l.rt_scene_upload.argtypes = [C.c_void_p, C.POINTER(L.RtSceneDesc)]
    l.rt_download_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    l.rt_scene_refit.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    l.rt_scene_upload_ex.argtypes = [C.c_void_p, C.POINTER(L.RtSceneDesc), C.c_uint32]
    l.rt_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(L.RtRenderConfig)]
    l.rt_sync.argtypes = [C.c_void_p]
    l.rt_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    l.rt_buffer_bytes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
    l.rt_get_device_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    l.rt_map_external_color.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    l.rt_bind_readback.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    l.rt_tiles_owned_pixels.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    l.rt_deinterleave_tiles.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    l.rt_get_stats.argtypes = [C.c_void_p, C.POINTER(L.RtStats)]
    l.rt_present.argtypes = [C.c_void_p, C.POINTER(L.RtPresentConfig), C.c_void_p, C.c_size_t]
    l.rt_comm_get_unique_id.argtypes = [C.c_void_p, C.c_size_t]
    l.rt_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    l.rt_comm_destroy.argtypes = [C.c_void_p]
    l.rt_gather_frame.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    l.rt_gl_register_buffer.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_void_p)]
    l.rt_gl_map.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    l.rt_gl_unmap.argtypes = [C.c_void_p, C.c_void_p]
    l.rt_gl_unregister.argtypes = [C.c_void_p, C.c_void_p]
    for name in EXPORTS:
        if name not in ("rt_last_error",):
            getattr(l, name).restype = C.c_int
    l.rt_last_error.restype = C.c_char_p
    _lib = l
    return l


def check(status: int) -> None:
    if status != L.RT_OK:
        raise RtError(status, (lib().rt_last_error() or b"").decode(errors="replace"))


_BUF_DTYPES = {L.RT_BUF_RGBA8: (np.int32, 1), L.RT_BUF_DEPTH: (np.float32, 1), L.RT_BUF_OBJID: (np.int32, 1), L.RT_BUF_RADIANCE: (np.float32, 4),
               L.RT_BUF_ACCUM: (np.float32, 4), L.RT_BUF_PRIM_ID: (np.int32, 1), L.RT_BUF_INST_ID: (np.int32, 1), L.RT_BUF_PRIMARY_T: (np.float32, 1),
               L.RT_BUF_SEG_COUNT: (np.uint8, 1), L.RT_BUF_TERM_CODE: (np.uint8, 1), L.RT_BUF_PATH_HASH: (np.uint32, 1),
               L.RT_BUF_GB_WORLDPOS: (np.float32, 3), L.RT_BUF_GB_NORMAL: (np.float32, 3), L.RT_BUF_GB_BASECOLOR: (np.float32, 3),
               L.RT_BUF_GB_MATID: (np.int32, 1), L.RT_BUF_TILE_RADIANCE: (np.float32, 4), L.RT_BUF_RESERVOIR: (L.RESERVOIR, 1), L.RT_BUF_PRESENT: (np.int32, 1),
               L.RT_BUF_GATHERED_RGBA8: (np.int32, 1), L.RT_BUF_GATHERED_DEPTH: (np.float32, 1), L.RT_BUF_GATHERED_OBJID: (np.int32, 1), L.RT_BUF_GATHERED_RADIANCE: (np.float32, 4)}


class Context:
    """Owns one rt_ctx (one GPU)."""

    def __init__(self, device: int = 0):
        self._l = lib()
        self.h = C.c_void_p()
        dev = (C.c_int * 1)(device)
        check(self._l.rt_create(dev, 1, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._l.rt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int | None):
        check(self._l.rt_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def stream_handle(self) -> int:
        """The cudaStream_t the context renders on (rt_get_stream), e.g. for torch.cuda.ExternalStream."""
        s = C.c_void_p()
        check(self._l.rt_get_stream(self.h, C.byref(s)))
        return s.value or 0

    def scene_upload(self, arrays: dict, device_build: bool = False):
        """rt_scene_upload; device_build=True builds the wide BVH on the GPU (rt_scene_upload_ex, RT_BUILD_DEVICE_LBVH)."""
        desc, keep = L.scene_desc_from_arrays(arrays)
        if device_build:
            check(self._l.rt_scene_upload_ex(self.h, C.byref(desc), L.RT_BUILD_DEVICE_LBVH))
        else:
            check(self._l.rt_scene_upload(self.h, C.byref(desc)))
        del keep

    def scene_refit(self, positions: np.ndarray):
        """New vertex positions for the uploaded mesh topology (BvhManager.BuildOrRefit(ForceRefit)): device-side refit of the wide BVH."""
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        check(self._l.rt_scene_refit(self.h, p.ctypes.data, len(p)))

    def scene_upload_desc(self, desc: L.RtSceneDesc):
        check(self._l.rt_scene_upload(self.h, C.byref(desc)))

    def render(self, cam: np.ndarray, cfg: L.RtRenderConfig, prev_cam: np.ndarray | None = None):
        cam = np.ascontiguousarray(cam, dtype=L.CAMERA)
        pc = None if prev_cam is None else np.ascontiguousarray(prev_cam, dtype=L.CAMERA)
        check(self._l.rt_render(self.h, cam.ctypes.data, None if pc is None else pc.ctypes.data, C.byref(cfg)))

    def sync(self):
        check(self._l.rt_sync(self.h))

    def buffer_bytes(self, which: int) -> int:
        n = C.c_size_t()
        check(self._l.rt_buffer_bytes(self.h, which, C.byref(n)))
        return n.value

    def download(self, which: int, out: np.ndarray | None = None) -> np.ndarray:
        dt, comps = _BUF_DTYPES[which]
        nbytes = self.buffer_bytes(which)
        n = nbytes // (np.dtype(dt).itemsize * comps)
        if out is None:
            out = np.empty((n, comps) if comps > 1 else (n,), dtype=dt)
        assert out.nbytes == nbytes and out.flags.c_contiguous
        check(self._l.rt_download(self.h, which, out.ctypes.data, nbytes))
        return out

    def device_buffer(self, which: int) -> tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        check(self._l.rt_get_device_buffer(self.h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def bind_readback(self, which: int, host: np.ndarray | None):
        """rt_bind_readback: every frame copies RGBA8 / depth / objectId into `host` (page-locked, image-sized) as soon as it is final."""
        if host is None:
            check(self._l.rt_bind_readback(self.h, which, None, 0))
        else:
            assert host.flags.c_contiguous
            check(self._l.rt_bind_readback(self.h, which, host.ctypes.data, host.nbytes))

    def map_external_color(self, dev_ptr: int | None, nbytes: int = 0):
        check(self._l.rt_map_external_color(self.h, C.c_void_p(dev_ptr or 0), nbytes))

    def present(self, out_width: int, out_height: int, taau: bool = True, dst_ptr: int | None = None, dst_bytes: int = 0, reset_history: bool = False,
                feedback: float = 0.075, sharpness: float = 0.10, clamp_k: float = 1.25):
        """The tail of RenderDirectToPbo (RTRenderer.cs:208-231): TAAU resolve, or blit / bilinear upsample."""
        pc = L.RtPresentConfig(L.RT_PRESENT_TAAU if taau else L.RT_PRESENT_COPY, out_width, out_height, feedback, sharpness, clamp_k, int(reset_history), (C.c_int32 * 4)(0, 0, 0, 0))
        check(self._l.rt_present(self.h, C.byref(pc), C.c_void_p(dst_ptr or 0), dst_bytes))

    def deinterleave_tiles(self, gathered_ptr: int, rank_offsets_px, world_size, width, height, tile_size, out_radiance_ptr=None, out_rgba8_ptr=None):
        offs = (C.c_int64 * world_size)(*[int(o) for o in rank_offsets_px])
        check(self._l.rt_deinterleave_tiles(self.h, C.c_void_p(gathered_ptr), offs, world_size, width, height, tile_size,
                                            C.c_void_p(out_radiance_ptr or 0), C.c_void_p(out_rgba8_ptr or 0)))

    # ---- multi-GPU (one process + one context per GPU; the communicator lives in the library) ----
    def comm_init(self, unique_id: bytes, rank: int, world_size: int):
        assert len(unique_id) == L.RT_COMM_ID_BYTES
        preload_host_nccl()
        buf = C.create_string_buffer(unique_id, L.RT_COMM_ID_BYTES)
        check(self._l.rt_comm_init(self.h, buf, L.RT_COMM_ID_BYTES, rank, world_size))

    def comm_destroy(self):
        check(self._l.rt_comm_destroy(self.h))

    def gather_frame(self, root: int = 0, what: int = L.RT_GATHER_RGBA8 | L.RT_GATHER_DEPTH_OBJID):
        """Tile payloads of the last frame -> the gathered image on `root` (NCCL send / recv + fused de-interleave / PackRGBA8), async."""
        check(self._l.rt_gather_frame(self.h, root, what))

    def stats(self) -> dict:
        s = L.RtStats()
        check(self._l.rt_get_stats(self.h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in L.RtStats._fields_ if k != "reserved"}
        d["extendLaunchesTimed"] = int(s.reserved[0])
        d["raysAnyHitTraced"] = int(s.reserved[1])   # shadow rays traced individually + shared sun probes (raysShadow counts the reference's ShadowOcclusion calls)
        d["raysSunProbe"] = int(s.reserved[2])
        d["lastGatherMs"] = int(s.reserved[3]) / 1000.0
        return d


_nccl_preloaded = False


def preload_host_nccl() -> None:
    """A Python host usually carries its own NCCL (the nvidia-nccl wheel PyTorch links against).  The library binds whichever
    libnccl.so.2 is already in the process, so load the wheel's copy BEFORE the first rt_comm_* call: otherwise the system copy gets
    in first and a later `import torch` finds that one under the same soname (and may miss symbols of its newer version)."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass   # no wheel copy: the library falls back to the system's libnccl.so.2


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 makes it; the host distributes the 128 bytes to every rank)."""
    preload_host_nccl()
    buf = C.create_string_buffer(L.RT_COMM_ID_BYTES)
    check(lib().rt_comm_get_unique_id(buf, L.RT_COMM_ID_BYTES))
    return buf.raw


def tiles_owned_pixels(width, height, tile_size, rank, world_size) -> int:
    n = C.c_int64()
    check(lib().rt_tiles_owned_pixels(width, height, tile_size, rank, world_size, C.byref(n)))
    return n.value
