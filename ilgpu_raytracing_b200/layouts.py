"""Blittable layouts of the C ABI (include/rtcore_b200.h) as numpy dtypes and ctypes structures.

Element layouts are byte-identical to the reference's device structs (file:line under
/root/reference/ILGPU_Raytracing/Engine): Float3.cs:6-10, Affine3x4.cs:3-7, Scene.cs:703-745,
MeshLoaderOBJ.cs:33-63, Sphere.cs:3-15, Camera.cs:5-17.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

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

assert (F3.itemsize, F2.itemsize, AFFINE.itemsize, MATERIAL.itemsize, SPHERE.itemsize, BVHNODE.itemsize, INSTANCE.itemsize,
        MESHTRI.itemsize, RGBA32.itemsize, TEXINFO.itemsize, CAMERA.itemsize) == (12, 8, 48, 44, 80, 44, 144, 12, 4, 12, 92)

SHADING_LAMBERT, SHADING_MIRROR, SHADING_GLASS = 0, 1, 2
BLAS_SPHERESET, BLAS_TRIMESH = 1, 2

# SceneDeviceViews order (Engine/SceneDeviceViews.cs:11-27)
SCENE_ARRAYS = [("tlasNodes", BVHNODE), ("tlasInstanceIndices", np.dtype("<i4")), ("instances", INSTANCE), ("blasNodes", BVHNODE),
                ("spherePrimIdx", np.dtype("<i4")), ("spheres", SPHERE), ("triPrimIdx", np.dtype("<i4")), ("meshPositions", F3),
                ("meshTris", MESHTRI), ("meshTexcoords", F2), ("meshTriUVs", MESHTRI), ("triMatIndex", np.dtype("<i4")),
                ("materials", MATERIAL), ("texels", RGBA32), ("texInfos", TEXINFO)]

RT_BUILD_DEVICE_LBVH = 1
RT_FLAG_TRI_MATERIALS = 1 << 0
RT_FLAG_ACCUMULATE = 1 << 1
RT_FLAG_RESET_ACCUM = 1 << 2
RT_FLAG_PATH_AOVS = 1 << 3
RT_FLAG_COUNTERS = 1 << 4
RT_FLAG_KERNEL_TIMING = 1 << 5
RT_FLAG_RESET_RESERVOIRS = 1 << 6
RT_FLAG_PUBLISH_RESERVOIRS = 1 << 7
RT_FLAG_FAST_SHADING = 1 << 8
RT_FLAG_FRAME_GRAPH = 1 << 9

(RT_BUF_RGBA8, RT_BUF_DEPTH, RT_BUF_OBJID, RT_BUF_RADIANCE, RT_BUF_ACCUM, RT_BUF_PRIM_ID, RT_BUF_INST_ID, RT_BUF_PRIMARY_T,
 RT_BUF_SEG_COUNT, RT_BUF_TERM_CODE, RT_BUF_PATH_HASH, RT_BUF_GB_WORLDPOS, RT_BUF_GB_NORMAL, RT_BUF_GB_BASECOLOR, RT_BUF_GB_MATID,
 RT_BUF_TILE_RADIANCE, RT_BUF_RESERVOIR, RT_BUF_PRESENT) = range(18)
RT_BUF_GATHERED_RGBA8, RT_BUF_GATHERED_DEPTH, RT_BUF_GATHERED_OBJID, RT_BUF_GATHERED_RADIANCE = 18, 19, 20, 21   # the image rt_gather_frame assembled on its root
RT_GATHER_RGBA8, RT_GATHER_RADIANCE, RT_GATHER_DEPTH_OBJID = 1, 2, 4
RT_COMM_ID_BYTES = 128
RT_PRESENT_TAAU, RT_PRESENT_COPY = 0, 1
# Reservoir (Engine/RTRay.cs:171-179), the element of RT_BUF_RESERVOIR
RESERVOIR = np.dtype([("L", F3), ("wi", F3), ("pdf", "<f4"), ("w", "<f4"), ("wSum", "<f4"), ("m", "<i4"), ("lightId", "<i4")])

RT_OK, RT_ERR_INVALID_ARGUMENT, RT_ERR_NO_DEVICE, RT_ERR_CUDA, RT_ERR_INVALID_STATE, RT_ERR_UNSUPPORTED, RT_ERR_OUT_OF_MEMORY = 0, -1, -2, -3, -4, -5, -6
RT_ERR_NCCL = -7


class CFloat3(C.Structure):
    _fields_ = [("X", C.c_float), ("Y", C.c_float), ("Z", C.c_float)]


class RtSceneDesc(C.Structure):
    _fields_ = [f for name, _ in SCENE_ARRAYS for f in ((name, C.c_void_p), ("n" + name[0].upper() + name[1:], C.c_int64))]


class RtRenderConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("frame", C.c_int32), ("spp", C.c_int32), ("maxDepth", C.c_int32),
                ("rngLockNoise", C.c_int32), ("enableTemporalReuse", C.c_int32), ("enableSpatialReuse", C.c_int32),
                ("dirLightDir", CFloat3), ("dirLightRadiance", CFloat3), ("skyTintTop", CFloat3), ("skyTintBottom", CFloat3),
                ("flags", C.c_uint32), ("tileSize", C.c_int32), ("rank", C.c_int32), ("worldSize", C.c_int32),
                ("samplesPerPass", C.c_int32), ("reserved", C.c_int32 * 3)]


class RtPresentConfig(C.Structure):
    _fields_ = [("mode", C.c_int32), ("outWidth", C.c_int32), ("outHeight", C.c_int32), ("feedback", C.c_float), ("sharpness", C.c_float),
                ("clampK", C.c_float), ("resetHistory", C.c_int32), ("reserved", C.c_int32 * 4)]


class RtStats(C.Structure):
    _fields_ = [("raysPrimary", C.c_uint64), ("raysBounce", C.c_uint64), ("raysShadow", C.c_uint64), ("wideNodes", C.c_uint64),
                ("trisTested", C.c_uint64), ("spheresTested", C.c_uint64), ("kernelLaunches", C.c_uint64),
                ("lastRenderMs", C.c_float), ("lastTraceMs", C.c_float),
                ("bvhWideNodeCount", C.c_uint64), ("bvhPrimCount", C.c_uint64), ("bvhBytes", C.c_uint64), ("reserved", C.c_uint64 * 4)]


def scene_desc_from_arrays(arrays: dict) -> tuple[RtSceneDesc, list]:
    """Build an RtSceneDesc over numpy arrays (kept alive by the returned list)."""
    d = RtSceneDesc()
    keep = []
    for name, dt in SCENE_ARRAYS:
        a = np.ascontiguousarray(arrays.get(name, np.zeros(0, dt)), dtype=dt).reshape(-1)
        keep.append(a)
        setattr(d, name, a.ctypes.data if len(a) else None)
        setattr(d, "n" + name[0].upper() + name[1:], len(a))
    return d, keep


def affine_identity() -> np.ndarray:
    a = np.zeros((), dtype=AFFINE)
    a["m00"] = a["m11"] = a["m22"] = 1.0
    return a


def affine_trs(translate=(0.0, 0.0, 0.0), rot_y_deg=0.0, scale=1.0) -> np.ndarray:
    """Rigid + uniform-scale objectToWorld (the only kind Scene.InvertRigidOrUniform inverts exactly, Scene.cs:616-638)."""
    a = np.zeros((), dtype=AFFINE)
    c, s = np.cos(np.deg2rad(rot_y_deg)), np.sin(np.deg2rad(rot_y_deg))
    a["m00"], a["m02"] = c * scale, s * scale
    a["m11"] = scale
    a["m20"], a["m22"] = -s * scale, c * scale
    a["m03"], a["m13"], a["m23"] = translate
    return a


def default_sun_dir(azimuth=0.0, elevation=0.9) -> np.ndarray:
    """RTRenderer.cs:174-178 with the defaults of :59-60 (host float math)."""
    az, el = np.float32(azimuth), np.float32(elevation)
    v = np.array([np.cos(az) * np.cos(el), np.sin(el), np.sin(az) * np.cos(el)], dtype=np.float32)
    inv = np.float32(1.0) / np.sqrt(np.maximum(np.float32(1e-20), v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), dtype=np.float32)
    return (v * inv).astype(np.float32)


def make_render_config(width, height, spp=1, max_depth=1, frame=0, rng_lock_noise=1, flags=0, tile_size=32, rank=0, world_size=1,
                       samples_per_pass=0, sun_dir=None, temporal=0, spatial=0) -> RtRenderConfig:
    """Light / sky constants are the reference's (RTRenderer.cs:191-194)."""
    sd = default_sun_dir() if sun_dir is None else np.asarray(sun_dir, dtype=np.float32)
    return RtRenderConfig(width, height, frame, spp, max_depth, rng_lock_noise, temporal, spatial,
                          CFloat3(*[float(v) for v in sd]), CFloat3(10, 10, 10), CFloat3(0.5, 0.7, 1.0), CFloat3(1.0, 1.0, 1.0),
                          flags, tile_size, rank, world_size, samples_per_pass, (C.c_int32 * 3)(0, 0, 0))
