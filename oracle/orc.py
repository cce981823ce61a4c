"""ctypes binding of the CPU oracle (oracle/rt_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under ilgpu_raytracing_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# ---- element layouts (same bytes as the reference's device structs; cf. include/rtcore_b200.h) ----
F3 = np.dtype([("X", "<f4"), ("Y", "<f4"), ("Z", "<f4")])
F2 = np.dtype([("X", "<f4"), ("Y", "<f4")])
AFFINE = np.dtype([(f"m{r}{c}", "<f4") for r in range(3) for c in range(4)])
MATERIAL = np.dtype([("Kd", F3), ("HasDiffuseMap", "<i4"), ("DiffuseTexIndex", "<i4"), ("Shading", "<i4"), ("IOR", "<f4"),
                     ("HasAlphaMap", "<i4"), ("AlphaTexIndex", "<i4"), ("TwoSided", "<i4"), ("AlphaCutoff", "<f4")])
SPHERE = np.dtype([("center", F3), ("radius", "<f4"), ("albedo", F3), ("material", MATERIAL), ("shading", "<i4"), ("ior", "<f4")])
BVHNODE = np.dtype([("boundsMin", F3), ("boundsMax", F3), ("left", "<i4"), ("right", "<i4"), ("first", "<i4"), ("count", "<i4"), ("skipIndex", "<i4")])
INSTANCE = np.dtype([("type", "<i4"), ("blasRoot", "<i4"), ("blasNodeCount", "<i4"), ("primIndexFirst", "<i4"), ("primIndexCount", "<i4"),
                     ("objectToWorld", AFFINE), ("worldToObject", AFFINE), ("uniformScale", "<f4"), ("worldBoundsMin", F3), ("worldBoundsMax", F3)])
MESHTRI = np.dtype([("i0", "<i4"), ("i1", "<i4"), ("i2", "<i4")])
RGBA32 = np.dtype([("R", "u1"), ("G", "u1"), ("B", "u1"), ("A", "u1")])
TEXINFO = np.dtype([("Offset", "<i4"), ("Width", "<i4"), ("Height", "<i4")])
CAMERA = np.dtype([("origin", F3), ("lowerLeft", F3), ("horizontal", F3), ("vertical", F3), ("forward", F3), ("right", F3), ("up", F3),
                   ("aspect", "<f4"), ("fovYRadians", "<f4")])
RESERVOIR = np.dtype([("L", F3), ("wi", F3), ("pdf", "<f4"), ("w", "<f4"), ("wSum", "<f4"), ("m", "<i4"), ("lightId", "<i4")])
assert (F3.itemsize, AFFINE.itemsize, MATERIAL.itemsize, SPHERE.itemsize, BVHNODE.itemsize, INSTANCE.itemsize, CAMERA.itemsize, RESERVOIR.itemsize) == (12, 48, 44, 80, 44, 144, 92, 44)

ARRAY_DTYPES = [BVHNODE, np.dtype("<i4"), INSTANCE, BVHNODE, np.dtype("<i4"), SPHERE, np.dtype("<i4"), F3, MESHTRI, F2, MESHTRI,
                np.dtype("<i4"), MATERIAL, RGBA32, TEXINFO]
ARRAY_NAMES = ["tlasNodes", "tlasInstanceIndices", "instances", "blasNodes", "spherePrimIdx", "spheres", "triPrimIdx", "meshPositions",
               "meshTris", "meshTexcoords", "meshTriUVs", "triMatIndex", "materials", "texels", "texInfos"]


class _F3(C.Structure):
    _fields_ = [("X", C.c_float), ("Y", C.c_float), ("Z", C.c_float)]


class OrcConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("frame", C.c_int32), ("spp", C.c_int32), ("maxDepth", C.c_int32),
                ("rngLockNoise", C.c_int32), ("enableTemporalReuse", C.c_int32), ("enableSpatialReuse", C.c_int32),
                ("dirLightDir", _F3), ("dirLightRadiance", _F3), ("skyTintTop", _F3), ("skyTintBottom", _F3),
                ("flags", C.c_uint32), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("threads", C.c_int32), ("noCull", C.c_int32)]


class OrcOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rgba8", "depth", "objId", "radiance", "primId", "instId", "primaryT", "hitMask",
                                          "gbPos", "gbNrm", "gbAlb", "gbMat", "segCount", "termCode", "pathHash", "resPrev", "resCur")] + \
               [("counters", C.c_uint64 * 8), ("seconds", C.c_double * 2)]


def build(force: bool = False) -> None:
    """Compile the oracle with gcc (Makefile next to this file)."""
    so = os.path.join(_HERE, "librt_oracle.so")
    src = os.path.join(_HERE, "rt_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True)


_libs: dict[str, C.CDLL] = {}


def lib(variant: str = "") -> C.CDLL:
    """variant '' = portable transcendentals (the pinned oracle); 'libm' = libm sinf/cosf/atan2f/acosf."""
    if variant in _libs:
        return _libs[variant]
    build()
    name = "librt_oracle.so" if not variant else f"librt_oracle_{variant}.so"
    L = C.CDLL(os.path.join(_HERE, name))
    L.orc_scene_new.restype = C.c_void_p
    L.orc_scene_free.argtypes = [C.c_void_p]
    L.orc_scene_clear.argtypes = [C.c_void_p]
    L.orc_scene_build_default.argtypes = [C.c_void_p]
    L.orc_scene_add_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_scene_add_sphere.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_scene_add_sphere_instance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_scene_add_mesh_instance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_scene_rebuild_tlas.argtypes = [C.c_void_p]
    L.orc_scene_sort_ties.argtypes = [C.c_void_p]
    L.orc_scene_sort_ties.restype = C.c_long
    L.orc_scene_array.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.orc_scene_array.restype = C.c_int64
    L.orc_camera_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_camera_translate.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.orc_camera_bake.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_rng_seed.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int]
    L.orc_rng_seed.restype = C.c_uint32
    L.orc_rng_stream.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_pack_rgba8.argtypes = [C.c_float, C.c_float, C.c_float]
    L.orc_pack_rgba8.restype = C.c_int32
    for fn in ("orc_intersect_triangle",):
        getattr(L, fn).argtypes = [C.c_void_p] * 6
    L.orc_intersect_sphere.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    L.orc_intersect_aabb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
    L.orc_math_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.orc_math_atan2.argtypes = [C.c_float, C.c_float]
    L.orc_math_atan2.restype = C.c_float
    L.orc_math_acos.argtypes = [C.c_float]
    L.orc_math_acos.restype = C.c_float
    L.orc_sample_hemisphere.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.orc_math_pow.argtypes = [C.c_float, C.c_float]
    L.orc_math_pow.restype = C.c_float
    L.orc_bilinear_upsample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.orc_taa_resolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
    L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OrcConfig), C.POINTER(OrcOutputs)]
    _libs[variant] = L
    return L


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f3(x, y, z) -> _F3:
    return _F3(float(x), float(y), float(z))


def affine_identity() -> np.ndarray:
    a = np.zeros((), dtype=AFFINE)
    a["m00"] = a["m11"] = a["m22"] = 1.0
    return a


class Scene:
    """Host side of Engine/Scene.cs as restated by the oracle."""

    def __init__(self, variant: str = ""):
        self.L = lib(variant)
        self.h = C.c_void_p(self.L.orc_scene_new())

    def __del__(self):
        try:
            self.L.orc_scene_free(self.h)
        except Exception:
            pass

    def build_default(self):
        self.L.orc_scene_build_default(self.h)

    def add_texture(self, texels: np.ndarray) -> int:
        t = np.ascontiguousarray(texels, dtype=np.uint8)
        assert t.ndim == 3 and t.shape[2] == 4
        return self.L.orc_scene_add_texture(self.h, t.shape[1], t.shape[0], _p(t))

    def add_sphere(self, sphere: np.ndarray) -> int:
        s = np.ascontiguousarray(sphere, dtype=SPHERE)
        return self.L.orc_scene_add_sphere(self.h, _p(s))

    def add_sphere_instance(self, ids, o2w: np.ndarray | None = None):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        m = affine_identity() if o2w is None else np.ascontiguousarray(o2w, dtype=AFFINE)
        self.L.orc_scene_add_sphere_instance(self.h, _p(ids), len(ids), _p(m))

    def add_mesh_instance(self, positions, tris, texcoords, tri_uvs, tri_mat, materials, o2w: np.ndarray | None = None):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        tr = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        uv = np.ascontiguousarray(texcoords, dtype=np.float32).reshape(-1, 2)
        tuv = np.ascontiguousarray(tri_uvs, dtype=np.int32).reshape(-1, 3)
        tm = np.ascontiguousarray(tri_mat, dtype=np.int32)
        mats = np.ascontiguousarray(materials, dtype=MATERIAL).reshape(-1)
        m = affine_identity() if o2w is None else np.ascontiguousarray(o2w, dtype=AFFINE)
        assert len(tuv) == len(tr) == len(tm)
        self.L.orc_scene_add_mesh_instance(self.h, _p(pos), len(pos), _p(tr), len(tr), _p(uv), len(uv), _p(tuv), _p(tm), _p(mats), len(mats), _p(m))

    def rebuild_tlas(self):
        self.L.orc_scene_rebuild_tlas(self.h)

    def sort_ties(self) -> int:
        return int(self.L.orc_scene_sort_ties(self.h))

    def arrays(self) -> dict[str, np.ndarray]:
        """Copies of the 15 SceneDeviceViews arrays (Engine/SceneDeviceViews.cs:11-27)."""
        out = {}
        for i, (name, dt) in enumerate(zip(ARRAY_NAMES, ARRAY_DTYPES)):
            ptr = C.c_void_p()
            n = self.L.orc_scene_array(self.h, i, C.byref(ptr))
            if n <= 0 or not ptr.value:
                out[name] = np.zeros(0, dtype=dt)
            else:
                buf = (C.c_char * (n * dt.itemsize)).from_address(ptr.value)
                out[name] = np.frombuffer(buf, dtype=dt).copy()
        return out

    def trace_closest(self, o, d, cull=True, flags=0):
        o = np.ascontiguousarray(o, dtype=np.float32)
        d = np.ascontiguousarray(d, dtype=np.float32)
        t, inst, prim = C.c_float(), C.c_int(), C.c_int()
        hit = self.L.orc_trace_closest(self.h, _p(o), _p(d), 1 if cull else 0, flags, C.byref(t), C.byref(inst), C.byref(prim))
        return bool(hit), t.value, inst.value, prim.value


def camera_create(width, height, fov_deg, origin=(0.0, 1.0, 3.0), look_at=(0.0, 0.5, 0.0), variant="") -> np.ndarray:
    """Camera.CreateCamera (Engine/Camera.cs:19-47) with the origin / lookAt made parameters."""
    cam = np.zeros((), dtype=CAMERA)
    o = np.asarray(origin, dtype=np.float32)
    l = np.asarray(look_at, dtype=np.float32)
    lib(variant).orc_camera_create(width, height, float(fov_deg), _p(o), _p(l), _p(cam))
    return cam


def camera_bake(cam: np.ndarray, width: int, height: int, variant="") -> np.ndarray:
    """RTRenderer.BakeCameraDerived (Engine/RTRenderer.cs:241-263): forward / right / up / fovY / aspect, as the renderer does before every launch."""
    lib(variant).orc_camera_bake(_p(cam), int(width), int(height))
    return cam


def camera_translate(cam: np.ndarray, dx, dy, dz, variant=""):
    lib(variant).orc_camera_translate(_p(cam), float(dx), float(dy), float(dz))
    return cam


@dataclass
class RenderResult:
    width: int
    height: int
    spp: int
    rgba8: np.ndarray
    depth: np.ndarray
    objId: np.ndarray
    radiance: np.ndarray
    primId: np.ndarray
    instId: np.ndarray
    primaryT: np.ndarray
    hitMask: np.ndarray
    gbPos: np.ndarray
    gbNrm: np.ndarray
    gbAlb: np.ndarray
    gbMat: np.ndarray
    segCount: np.ndarray
    termCode: np.ndarray
    pathHash: np.ndarray
    counters: dict
    seconds: tuple


def bilinear_upsample(src: np.ndarray, src_w: int, src_h: int, dst_w: int, dst_h: int, variant="") -> np.ndarray:
    """RTRenderer.BilinearUpsampleKernel (Engine/RTRenderer.cs:287-320) over the whole output image."""
    src = np.ascontiguousarray(src, np.int32)
    dst = np.zeros(dst_w * dst_h, np.int32)
    lib(variant).orc_bilinear_upsample(src.ctypes.data, src_w, src_h, dst.ctypes.data, dst_w, dst_h)
    return dst


class TaaState:
    """RTTaa (Engine/RTTaa.cs): history colour / object id and the _historyValid flag; resolve() = ResolveUpsample with the reference's tunables."""

    def __init__(self, out_w: int, out_h: int, variant=""):
        self.w, self.h, self.valid, self.variant = out_w, out_h, False, variant
        self.hist_color = np.zeros(out_w * out_h, np.int32)
        self.hist_obj = np.zeros(out_w * out_h, np.int32)

    def resolve(self, low_color: np.ndarray, low_obj: np.ndarray, in_w: int, in_h: int, feedback=0.075, sharpness=0.10, clamp_k=1.25) -> np.ndarray:
        out = np.zeros(self.w * self.h, np.int32)
        lc, lo = np.ascontiguousarray(low_color, np.int32), np.ascontiguousarray(low_obj, np.int32)
        lib(self.variant).orc_taa_resolve(out.ctypes.data, lc.ctypes.data, lo.ctypes.data, in_w, in_h, self.w, self.h, self.hist_color.ctypes.data,
                                          self.hist_obj.ctypes.data, 0 if self.valid else 1, feedback, sharpness, clamp_k)
        self.valid = True
        return out


def default_sun_dir(azimuth=0.0, elevation=0.9) -> np.ndarray:
    """Engine/RTRenderer.cs:174-178 with the defaults of :59-60 (host float math)."""
    az, el = np.float32(azimuth), np.float32(elevation)
    v = np.array([np.cos(az) * np.cos(el), np.sin(el), np.sin(az) * np.cos(el)], dtype=np.float32)
    inv = np.float32(1.0) / np.sqrt(np.maximum(np.float32(1e-20), v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), dtype=np.float32)
    return (v * inv).astype(np.float32)


def make_config(width, height, spp=1, max_depth=1, frame=0, rng_lock_noise=1, flags=0, crop=None, threads=0,
                temporal=0, spatial=0, sun_dir=None, no_cull=0) -> OrcConfig:
    sd = default_sun_dir() if sun_dir is None else np.asarray(sun_dir, dtype=np.float32)
    x0, y0, x1, y1 = crop if crop is not None else (0, 0, width, height)
    return OrcConfig(width, height, frame, spp, max_depth, rng_lock_noise, temporal, spatial,
                     f3(*sd), f3(10, 10, 10), f3(0.5, 0.7, 1.0), f3(1.0, 1.0, 1.0),   # Engine/RTRenderer.cs:191-194
                     flags, x0, y0, x1, y1, threads, no_cull)


def render(scene: Scene, cam: np.ndarray, cfg: OrcConfig, prev_cam: np.ndarray | None = None, aovs: bool = True,
           res_prev: np.ndarray | None = None, res_cur: np.ndarray | None = None) -> RenderResult:
    cw, ch = cfg.x1 - cfg.x0, cfg.y1 - cfg.y0
    n = cw * ch
    spp = max(1, cfg.spp)
    r = RenderResult(cw, ch, spp,
                     np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros((n, 3), np.float32),
                     np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32),
                     np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.int32),
                     np.zeros((spp, n) if aovs else (0, 0), np.uint8), np.zeros((spp, n) if aovs else (0, 0), np.uint8),
                     np.zeros((spp, n) if aovs else (0, 0), np.uint32), {}, (0.0, 0.0))
    o = OrcOutputs()
    for name in ("rgba8", "depth", "objId", "radiance", "primId", "instId", "primaryT", "hitMask", "gbPos", "gbNrm", "gbAlb", "gbMat"):
        setattr(o, name, getattr(r, name).ctypes.data)
    if aovs:
        o.segCount, o.termCode, o.pathHash = r.segCount.ctypes.data, r.termCode.ctypes.data, r.pathHash.ctypes.data
    if res_prev is not None and res_cur is not None:
        o.resPrev, o.resCur = res_prev.ctypes.data, res_cur.ctypes.data
    rc = scene.L.orc_render(scene.h, _p(cam), _p(prev_cam) if prev_cam is not None else None, C.byref(cfg), C.byref(o))
    if rc != 0:
        raise RuntimeError(f"orc_render failed: {rc}")
    names = ["raysPrimary", "raysBounce", "raysShadow", "nodes", "tris", "spheres"]
    r.counters = {k: int(o.counters[i]) for i, k in enumerate(names)}
    r.seconds = (o.seconds[0], o.seconds[1])
    return r
