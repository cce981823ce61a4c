"""ctypes binding of tests/hostsim (CPU build of the core's stage bodies; test harness only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ilgpu_raytracing_b200 import layouts as L

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")


class HsOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rgba8", "depth", "objId", "radiance4", "accum4", "primId", "instId", "primaryT",
                                          "gbPos", "gbNrm", "gbAlb", "gbMat", "segCount", "termCode", "pathHash")] + [("counters", C.c_uint64 * 8)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
        _lib = C.CDLL(os.path.join(_HERE, "libhostsim.so"))
        _lib.hs_scene_create.restype = C.c_void_p
        _lib.hs_scene_create.argtypes = [C.POINTER(L.RtSceneDesc)]
        _lib.hs_scene_create_ex.restype = C.c_void_p
        _lib.hs_scene_create_ex.argtypes = [C.POINTER(L.RtSceneDesc), C.c_int]
        _lib.hs_scene_error.restype = C.c_char_p
        _lib.hs_scene_error.argtypes = [C.c_void_p]
        _lib.hs_scene_destroy.argtypes = [C.c_void_p]
        _lib.hs_scene_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib.hs_scene_hash.argtypes = [C.c_void_p]
        _lib.hs_scene_hash.restype = C.c_uint64
        _lib.hs_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint, C.c_void_p, C.c_void_p]
        _lib.hs_render.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(L.RtRenderConfig), C.POINTER(HsOutputs)]
        _lib.hs_bilinear_upsample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _lib.hs_taa_resolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        _lib.hs_set_lane_schedule.argtypes = [C.c_int]
        _lib.hs_set_plane_pad.argtypes = [C.c_float]
        _lib.hs_capture.argtypes = [C.c_int]
        _lib.hs_capture_count.restype = C.c_longlong
        _lib.hs_simulate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
        _lib.hs_pow.argtypes = [C.c_float, C.c_float]
        _lib.hs_pow.restype = C.c_float
        _lib.hs_render_reuse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(L.RtRenderConfig), C.POINTER(HsOutputs), C.c_void_p, C.c_void_p]
    return _lib


def set_lane_schedule(node_steps: int):
    """0: a node step then all of its primitives; n > 0: k_extend's per-lane schedule (n node steps, one primitive step, queued groups)."""
    lib().hs_set_lane_schedule(int(node_steps))


def capture_rays(on: bool):
    """Analysis tool: record every ray the next renders trace (wave by wave) for simulate()."""
    lib().hs_capture(1 if on else 0)


def simulate(scene, any_hit: bool, policy=0, node_steps=2, prim_vote=1, warps=256, c_refill=54.0, c_node=250.0, c_prim=200.0, c_loop=45.0, pre_cull=0.0) -> dict:
    """The captured waves through k_extend's warp loop on simulated 32-lane warps (tests/hostsim/hostsim.cpp: hs_simulate)."""
    out = np.zeros(11, np.float64)
    lib().hs_simulate(scene.h, int(any_hit), policy, node_steps, prim_vote, warps, c_refill, c_node, c_prim, c_loop, pre_cull, out.ctypes.data)
    keys = ["rays", "iterations", "warp_instr", "node_phases", "node_lanes", "prim_phases", "prim_lanes", "node_steps", "prim_steps", "pre_culled", "accepted"]
    d = dict(zip(keys, out.tolist()))
    d["warp_instr_per_ray"] = d["warp_instr"] / max(1.0, d["rays"])
    d["lanes_per_node_phase"] = d["node_lanes"] / max(1.0, d["node_phases"])
    d["lanes_per_prim_phase"] = d["prim_lanes"] / max(1.0, d["prim_phases"])
    return d


def set_plane_pad(quanta: float):
    """Analysis knob: widen every child box of the wide BVH by that many quanta per side (0 = the shipped test)."""
    lib().hs_set_plane_pad(float(quanta))


class HostSimScene:
    def __init__(self, arrays: dict, max_depth: int = 0):
        """max_depth: depth limit handed to the wide-BVH builder (0 = the traversal stack's; small values force the depth-bounded rebuild)."""
        self.desc, self._keep = L.scene_desc_from_arrays(arrays)
        self.h = C.c_void_p(lib().hs_scene_create_ex(C.byref(self.desc), int(max_depth)))
        err = lib().hs_scene_error(self.h).decode()
        if err:
            raise ValueError(err)

    def __del__(self):
        try:
            lib().hs_scene_destroy(self.h)
        except Exception:
            pass

    def stats(self):
        s = np.zeros(6, np.int64)
        lib().hs_scene_stats(self.h, s.ctypes.data)
        return dict(nPrims=int(s[0]), nTris=int(s[1]), nSpheres=int(s[2]), nWideNodes=int(s[3]), maxDepth=int(s[4]), depthBounded=int(s[5]))

    def bvh_hash(self) -> int:
        """FNV-1a of the wide nodes + primitive records the host builder produced."""
        return int(lib().hs_scene_hash(self.h))

    def trace(self, o, d, any_hit=False, t_max=1e30, flags=0):
        o = np.ascontiguousarray(o, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        out = np.zeros(5, np.float32)
        cnt = np.zeros(3, np.uint32)
        hit = lib().hs_trace(self.h, o.ctypes.data, d.ctypes.data, int(any_hit), float(t_max), flags, out.ctypes.data, cnt.ctypes.data)
        return bool(hit), float(out[0]), int(out[2]), int(out[1]), cnt

    def render(self, cam: np.ndarray, cfg: L.RtRenderConfig, aovs=True, prev_cam=None, res_prev=None, res_cur=None):
        W, H, spp = cfg.width, cfg.height, max(1, cfg.spp)
        n = W * H
        r = dict(rgba8=np.zeros(n, np.int32), depth=np.zeros(n, np.float32), objId=np.zeros(n, np.int32), radiance4=np.zeros((n, 4), np.float32),
                 accum4=np.zeros((n, 4), np.float32), primId=np.full(n, -2, np.int32), instId=np.full(n, -2, np.int32), primaryT=np.zeros(n, np.float32),
                 gbPos=np.zeros((n, 3), np.float32), gbNrm=np.zeros((n, 3), np.float32), gbAlb=np.zeros((n, 3), np.float32), gbMat=np.zeros(n, np.int32),
                 segCount=np.zeros((spp, n), np.uint8), termCode=np.zeros((spp, n), np.uint8), pathHash=np.zeros((spp, n), np.uint32))
        o = HsOutputs()
        for k, v in r.items():
            setattr(o, k, v.ctypes.data)
        if aovs:
            cfg.flags |= L.RT_FLAG_PATH_AOVS
        cam = np.ascontiguousarray(cam, dtype=L.CAMERA)
        if res_prev is not None:
            pc = np.ascontiguousarray(prev_cam if prev_cam is not None else cam, dtype=L.CAMERA)
            assert res_prev.dtype == L.RESERVOIR and res_cur.dtype == L.RESERVOIR and res_cur.flags.c_contiguous
            rc = lib().hs_render_reuse(self.h, cam.ctypes.data, pc.ctypes.data, C.byref(cfg), C.byref(o), res_prev.ctypes.data, res_cur.ctypes.data)
        else:
            rc = lib().hs_render(self.h, cam.ctypes.data, C.byref(cfg), C.byref(o))
        if rc != 0:
            raise RuntimeError(f"hs_render: {rc}")
        names = ["raysPrimary", "raysBounce", "raysShadow", "nodes", "tris", "spheres"]
        r["counters"] = {k: int(o.counters[i]) for i, k in enumerate(names)}
        r["radiance"] = r["radiance4"][:, :3]
        return r
