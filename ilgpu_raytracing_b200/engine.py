"""Python face of the C++ Engine mirror (csrc/host/engine.*): the reference's host API for the hot path.

Names follow the reference (Engine/Scene.cs, SceneManager.cs, Camera.cs, RTRenderer.cs, Framebuffer.cs).
Everything heavy happens in native code: librtengine_host.so (scene lists, BVH2 builders, camera,
renderer orchestration) on top of librtcore_b200.so (the CUDA core behind the C ABI).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import layouts as L
from . import native
from .scenes import CAMERAS, SceneSpec

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_PKG, "librtengine_host.so")


class EngineError(RuntimeError):
    """Carries the name of the reference exception the C++ mirror raised (ArgumentNullException, ...)."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class EngKnobs(C.Structure):
    _fields_ = [("renderScale", C.c_float), ("enableTemporalReuse", C.c_int), ("enableSpatialReuse", C.c_int), ("rngLockNoise", C.c_int),
                ("fixedSeed", C.c_int), ("spp", C.c_int), ("maxDepth", C.c_int), ("flags", C.c_uint), ("tileSize", C.c_int), ("rank", C.c_int),
                ("worldSize", C.c_int), ("samplesPerPass", C.c_int), ("enableTAAU", C.c_int), ("asyncSubmit", C.c_int)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    native.lib()   # librtcore_b200.so first (the engine links against it)
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: run __graft_entry__.build()")
    l = C.CDLL(_LIB_PATH)
    l.eng_last_error.restype = C.c_char_p
    l.eng_scene_new_hostonly.restype = C.c_void_p
    l.eng_scene_free_hostonly.argtypes = [C.c_void_p]
    for fn in ("eng_scene_build_default", "eng_scene_rebuild_tlas", "eng_scene_upload_all"):
        getattr(l, fn).argtypes = [C.c_void_p]
    l.eng_scene_reset.argtypes = [C.c_void_p]
    l.eng_scene_reset.restype = None
    l.eng_scene_add_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    l.eng_scene_add_sphere.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    l.eng_scene_add_sphere_instance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    l.eng_scene_load_mesh_instance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_int, C.c_void_p]
    l.eng_scene_load_obj_instance.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_float]
    l.eng_scene_set_mesh_positions.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    l.eng_scene_can_refit.argtypes = [C.c_void_p]
    l.eng_scene_set_device_build.argtypes = [C.c_void_p, C.c_int]
    l.eng_scene_set_device_build.restype = None
    l.eng_renderer_commit_policy.argtypes = [C.c_void_p, C.c_int]
    l.eng_scene_sort_ties.argtypes = [C.c_void_p]
    l.eng_scene_sort_ties.restype = C.c_long
    l.eng_scene_fill_desc.argtypes = [C.c_void_p, C.POINTER(L.RtSceneDesc)]
    l.eng_scene_fill_desc.restype = None
    l.eng_camera_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_void_p]
    l.eng_camera_create_at.argtypes = [C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    l.eng_camera_look_at.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
    l.eng_camera_translate.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    l.eng_camera_set_fov.argtypes = [C.c_void_p, C.c_float, C.c_float]
    l.eng_camera_rotate_yaw_pitch.argtypes = [C.c_void_p, C.c_float, C.c_float]
    l.eng_camera_bake.argtypes = [C.c_void_p, C.c_int, C.c_int]
    l.eng_renderer_new.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    l.eng_renderer_free.argtypes = [C.c_void_p]
    l.eng_renderer_native.argtypes = [C.c_void_p]
    l.eng_renderer_native.restype = C.c_void_p
    l.eng_renderer_scene.argtypes = [C.c_void_p]
    l.eng_renderer_scene.restype = C.c_void_p
    l.eng_renderer_commit.argtypes = [C.c_void_p]
    l.eng_renderer_get_camera.argtypes = [C.c_void_p, C.c_void_p]
    l.eng_renderer_set_camera.argtypes = [C.c_void_p, C.c_void_p]
    l.eng_renderer_set_sun_params.argtypes = [C.c_void_p, C.c_float, C.c_float]
    l.eng_renderer_set_knobs.argtypes = [C.c_void_p, C.POINTER(EngKnobs)]
    l.eng_renderer_render_direct_to_pbo.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
    l.eng_renderer_last_config.argtypes = [C.c_void_p, C.POINTER(L.RtRenderConfig)]
    l.eng_renderer_new_communicator_id.argtypes = [C.c_void_p]
    l.eng_renderer_init_multi_gpu.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    l.eng_framebuffer_download_to_cpu.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    l.eng_framebuffer_bind_cpu_targets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    _lib = l
    return l


def _check(rc: int):
    if rc != 0:
        raise EngineError(rc, (lib().eng_last_error() or b"").decode(errors="replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------------------------- Camera
def create_camera(width: int, height: int, fov_degrees: float = 60.0) -> np.ndarray:
    """Camera.CreateCamera (Engine/Camera.cs:19-47)."""
    cam = np.zeros((), L.CAMERA)
    lib().eng_camera_create(width, height, fov_degrees, _p(cam))
    return cam


def create_camera_at(width, height, fov_degrees, origin, look_at) -> np.ndarray:
    cam = np.zeros((), L.CAMERA)
    o, l = np.asarray(origin, np.float32), np.asarray(look_at, np.float32)
    lib().eng_camera_create_at(width, height, fov_degrees, _p(o), _p(l), _p(cam))
    return cam


def camera_translate(cam: np.ndarray, dx, dy, dz) -> np.ndarray:
    """Camera.Translate (Engine/Camera.cs:121-126)."""
    lib().eng_camera_translate(_p(cam), dx, dy, dz)
    return cam


def camera_set_fov(cam, fov_degrees, aspect):
    lib().eng_camera_set_fov(_p(cam), fov_degrees, aspect)
    return cam


def camera_rotate_yaw_pitch(cam, yaw_deg, pitch_deg):
    lib().eng_camera_rotate_yaw_pitch(_p(cam), yaw_deg, pitch_deg)
    return cam


def config_camera(name: str, width: int, height: int) -> np.ndarray:
    """Cameras of the benchmark configs."""
    c = CAMERAS[name]
    cam = create_camera_at(width, height, c["fov"], c["origin"], c["look_at"])
    if c["translate"] is not None:
        camera_translate(cam, *c["translate"])
    return cam


# ---------------------------------------------------------------------------------------------------------------- Scene
class Scene:
    """Engine/Scene.cs host part.  Owned by an RTRenderer (device-backed) or host-only (builders only)."""

    def __init__(self, handle=None, owner=None):
        self._l = lib()
        self._own = handle is None
        self.h = C.c_void_p(self._l.eng_scene_new_hostonly()) if handle is None else C.c_void_p(handle)
        self._owner = owner

    def __del__(self):
        try:
            if self._own and self.h:
                self._l.eng_scene_free_hostonly(self.h)
        except Exception:
            pass

    def BuildDefaultScene(self):
        _check(self._l.eng_scene_build_default(self.h))

    def Reset(self):
        self._l.eng_scene_reset(self.h)

    def AddTexture(self, texels: np.ndarray) -> int:
        t = np.ascontiguousarray(texels, np.uint8)
        if t.ndim != 3 or t.shape[2] != 4:
            raise ValueError("texture must be (h, w, 4) uint8 RGBA")
        out = C.c_int()
        _check(self._l.eng_scene_add_texture(self.h, t.shape[1], t.shape[0], _p(t), C.byref(out)))
        return out.value

    def AddSphere(self, sphere: np.ndarray) -> int:
        s = np.ascontiguousarray(sphere, L.SPHERE)
        out = C.c_int()
        _check(self._l.eng_scene_add_sphere(self.h, _p(s), C.byref(out)))
        return out.value

    def AddSphereInstance(self, sphere_ids, object_to_world=None):
        ids = np.ascontiguousarray(sphere_ids, np.int32)
        m = L.affine_identity() if object_to_world is None else np.ascontiguousarray(object_to_world, L.AFFINE)
        _check(self._l.eng_scene_add_sphere_instance(self.h, _p(ids), len(ids), _p(m)))

    def LoadMeshInstance(self, positions, tris, texcoords, tri_uvs, tri_mat, materials, object_to_world=None):
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        tr = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
        uv = np.ascontiguousarray(texcoords, np.float32).reshape(-1, 2)
        tuv = np.ascontiguousarray(tri_uvs, np.int32).reshape(-1, 3)
        tm = np.ascontiguousarray(tri_mat, np.int32).reshape(-1)
        mats = np.ascontiguousarray(materials, L.MATERIAL).reshape(-1)
        if not (len(tr) == len(tuv) == len(tm)):
            raise ValueError("tris, tri_uvs and tri_mat must have the same length")
        m = L.affine_identity() if object_to_world is None else np.ascontiguousarray(object_to_world, L.AFFINE)
        _check(self._l.eng_scene_load_mesh_instance(self.h, _p(pos), len(pos), _p(tr), len(tr), _p(uv), len(uv), _p(tuv), _p(tm), _p(mats), len(mats), _p(m)))

    def LoadObjInstance(self, obj_path: str, object_to_world=None, uniform_scale: float = 1.0):
        """Scene.LoadObjInstance (Engine/Scene.cs:144-256): OBJ + MTL + TGA / BMP / PNG textures from disk (MeshLoaderOBJ.cs)."""
        m = L.affine_identity() if object_to_world is None else np.ascontiguousarray(object_to_world, L.AFFINE)
        _check(self._l.eng_scene_load_obj_instance(self.h, str(obj_path).encode(), _p(m), float(uniform_scale)))

    def SetMeshPositions(self, positions):
        """Moved vertices for the loaded mesh (same count, same triangles); Commit(FORCE_REFIT) then refits instead of rebuilding."""
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        _check(self._l.eng_scene_set_mesh_positions(self.h, _p(p), len(p)))

    def SetDeviceBuild(self, on: bool):
        """UploadAll builds the wide BVH on the GPU (Morton order, radix tree / PLOC, SAH-optimal 8-wide collapse) instead of the host's SAH build."""
        self._l.eng_scene_set_device_build(self.h, 1 if on else 0)

    def CanRefit(self) -> bool:
        return bool(self._l.eng_scene_can_refit(self.h))

    def RebuildTLAS(self):
        _check(self._l.eng_scene_rebuild_tlas(self.h))

    def UploadAll(self):
        _check(self._l.eng_scene_upload_all(self.h))

    def sort_ties(self) -> int:
        return int(self._l.eng_scene_sort_ties(self.h))

    def arrays(self) -> dict:
        """Copies of the 15 host arrays behind SceneDeviceViews (Engine/SceneDeviceViews.cs:11-27)."""
        d = L.RtSceneDesc()
        self._l.eng_scene_fill_desc(self.h, C.byref(d))
        out = {}
        for name, dt in L.SCENE_ARRAYS:
            ptr, n = getattr(d, name), getattr(d, "n" + name[0].upper() + name[1:])
            if not ptr or n <= 0:
                out[name] = np.zeros(0, dt)
            else:
                out[name] = np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(ptr), dtype=dt).copy()
        return out

    def load_spec(self, spec: SceneSpec):
        """Replace the scene by a SceneSpec: textures, all spheres, then the instances in the spec's order, then the TLAS."""
        self.Reset()
        for t in spec.textures:
            self.AddTexture(t)
        for s in spec.spheres:
            self.AddSphere(s)

        def mesh():
            m = spec.mesh
            self.LoadMeshInstance(m.positions, m.tris, m.texcoords, m.tri_uvs, m.tri_mat, m.materials, m.object_to_world)

        if spec.mesh is not None and spec.mesh_first:
            mesh()
        for ids, xf in spec.sphere_instances:
            self.AddSphereInstance(ids, xf)
        if spec.mesh is not None and not spec.mesh_first:
            mesh()
        self.RebuildTLAS()
        return self


# ---------------------------------------------------------------------------------------------------------------- RTRenderer
class RTRenderer:
    """Engine/RTRenderer.cs: owns the native context, the SceneManager/Scene, the camera and the Framebuffer."""

    def __init__(self, device_index: int = 0, width: int = 1280, height: int = 720):
        self._l = lib()
        self.h = C.c_void_p()
        _check(self._l.eng_renderer_new(device_index, width, height, C.byref(self.h)))
        self.scene = Scene(self._l.eng_renderer_scene(self.h), owner=self)
        self.knobs = EngKnobs(1.0, 0, 0, 1, 1, 2, 3, 0, 16, 0, 1, 0, 0)   # benchmark-style defaults: full-resolution trace, reuse and TAAU off
        self._native_view = None

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self._l.eng_renderer_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def native(self) -> "native.Context":
        """The rt_ctx behind this renderer, wrapped for downloads / stats (not owned)."""
        if self._native_view is None:
            v = native.Context.__new__(native.Context)
            v._l = native.lib()
            v.h = C.c_void_p(self._l.eng_renderer_native(self.h))
            v.close = lambda: None
            self._native_view = v
        return self._native_view

    AUTO, FORCE_REFIT, FORCE_REBUILD = 0, 1, 2   # RebuildPolicy, Engine/BvhManager.cs:13-18

    def Commit(self, policy: int = 0):
        """SceneManager.Commit(policy) -> BvhManager.BuildOrRefit: Scene.UploadAll, or the device-side refit for FORCE_REFIT when only
        vertex positions changed since the last upload."""
        _check(self._l.eng_renderer_commit_policy(self.h, int(policy)))

    def SetSunParams(self, speed_rad_per_sec: float, elevation_rad: float):
        self._l.eng_renderer_set_sun_params(self.h, speed_rad_per_sec, elevation_rad)

    @property
    def camera(self) -> np.ndarray:
        cam = np.zeros((), L.CAMERA)
        self._l.eng_renderer_get_camera(self.h, _p(cam))
        return cam

    @camera.setter
    def camera(self, cam: np.ndarray):
        c = np.ascontiguousarray(cam, L.CAMERA)
        self._l.eng_renderer_set_camera(self.h, _p(c))

    def configure(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.knobs, k):
                raise AttributeError(k)
            setattr(self.knobs, k, v)
        self._l.eng_renderer_set_knobs(self.h, C.byref(self.knobs))

    def RenderDirectToPbo(self, pbo_device_ptr: int | None, width: int, height: int, frame: int = 0, dt: float = 0.0):
        self._l.eng_renderer_set_knobs(self.h, C.byref(self.knobs))
        _check(self._l.eng_renderer_render_direct_to_pbo(self.h, C.c_void_p(pbo_device_ptr or 0), width, height, frame, dt))

    @staticmethod
    def NewCommunicatorId() -> bytes:
        """Rank 0: the 128-byte id every rank's InitMultiGpu needs (ncclGetUniqueId behind rt_comm_get_unique_id)."""
        native.preload_host_nccl()
        buf = C.create_string_buffer(L.RT_COMM_ID_BYTES)
        _check(lib().eng_renderer_new_communicator_id(buf))
        return buf.raw

    def InitMultiGpu(self, unique_id: bytes, rank: int, world_size: int):
        """This renderer = rank `rank` of `world_size` processes (one per GPU): RenderDirectToPbo then renders this rank's screen tiles,
        gathers colour + depth + objectId on rank 0 (NCCL inside the library) and presents there."""
        native.preload_host_nccl()
        buf = C.create_string_buffer(unique_id, L.RT_COMM_ID_BYTES)
        _check(self._l.eng_renderer_init_multi_gpu(self.h, buf, rank, world_size))
        self.knobs.rank, self.knobs.worldSize = rank, world_size

    def Synchronize(self):
        """_cuda.Synchronize() (Engine/RTRenderer.cs:233): wait for everything queued on the renderer's context."""
        self.native.sync()

    def last_config(self) -> L.RtRenderConfig:
        cfg = L.RtRenderConfig()
        self._l.eng_renderer_last_config(self.h, C.byref(cfg))
        return cfg

    def BindCpuTargets(self, color, depth, objid):
        """Framebuffer.BindCpuTargets: every frame lands in these page-locked arrays (int32 / float32 / int32, one entry per traced pixel;
        all None unbinds); DownloadToCpu(the same arrays) then only waits for the copies."""
        n = 0 if color is None else color.size
        _check(self._l.eng_framebuffer_bind_cpu_targets(self.h, _p(color), _p(depth), _p(objid), n))

    def DownloadToCpu(self, out_color=None, out_depth=None, out_objid=None):
        """Framebuffer.DownloadToCpu(0) + CpuColor / CpuDepth / CpuObjectId (Engine/Framebuffer.cs:148-160)."""
        cfg = self.last_config()
        n = cfg.width * cfg.height
        color = np.empty(n, np.int32) if out_color is None else out_color
        depth = np.empty(n, np.float32) if out_depth is None else out_depth
        objid = np.empty(n, np.int32) if out_objid is None else out_objid
        _check(self._l.eng_framebuffer_download_to_cpu(self.h, 0, _p(color), _p(depth), _p(objid), n))
        return color, depth, objid
