// Engine.Patched.cs - the members of the reference's Engine classes that change when the ILGPU kernels are replaced by
// librtcore_b200 (include/rtcore_b200.h).  Each member below REPLACES the member of the same name in the cited reference file;
// everything not listed (Program.cs, RTWindow, Camera, CameraController, FlyCameraController, MeshLoaderOBJ, the host lists and
// builders of Scene, Float3 / Affine3x4 / Sphere ...) stays as it is.  Line numbers are the reference's.
//
//   RTRenderer   (Engine/RTRenderer.cs)  fields :25-41,53-56 -> one RtDevice; ctor :63-92; Accelerator :94; RenderDirectToPbo :105-237;
//                                        Dispose :347-375.  New: InitMultiGpu.
//   SceneManager (Engine/SceneManager.cs) ctors :20-21 take an RtDevice.  Commit :23 is UNCHANGED (`_bvh.BuildOrRefit(_scene, policy)`);
//                                        GetDeviceViews :25-28 is removed (no caller is left: the kernels read the scene natively).
//   BvhManager   (Engine/BvhManager.cs)   ctor :25 takes an RtDevice; BuildOrRefit(Scene, RebuildPolicy) :27 keeps its signature;
//                                        GetDeviceViews :29-48 is removed.
//   Scene        (Engine/Scene.cs)        ctor :60-64 takes an RtDevice; UploadAll :258-279; the 15 MemoryBuffer1D fields :41-58 and
//                                        the *View properties :66-81 are removed.
//   Framebuffer  (Engine/Framebuffer.cs)  ctor :53; DownloadToCpu / CpuColor / CpuDepth / CpuObjectId :148-160; the device buffers live
//                                        in the native context (EnsureLength / GetGpu / GetReservoirPair have no caller left).
//   CudaGlInteropIndexBuffer (Engine/CudaGlInteropIndexBuffer.cs) no longer derives from ILGPU's MemoryBuffer: same constructor shape,
//                                        MapCuda / UnmapCuda / IsValid / Dispose, `DevicePointer` instead of GetCudaArrayView.
//
// NOTE: shipped as source; this image has no .NET toolchain, so it has not been compiled here.
using System;
using OpenTK.Graphics.OpenGL4;

namespace ILGPU_Raytracing.Engine
{
    // ================================================================================================ RTRenderer
    public sealed unsafe class RTRenderer : IDisposable
    {
        private readonly RTWindow _window;
        private readonly RtDevice _device;                      // replaces _context / _cuda / _stream (:25-27)
        private readonly SceneManager _sceneManager;
        private readonly CameraController _cameraController;
        private Camera _camera, _prevCamera;
        private readonly Framebuffer _framebuffer;

        private float _renderScale = 0.67f;                     // :43-49, unchanged defaults
        private bool _enableTAAU = true;
        private int _enableTemporalReuse = 1, _enableSpatialReuse = 1, _rngLockNoise = 1, _spp = 2;
        private float _sunAzimuth = 0.0f, _sunElevation = 0.9f, _sunSpeedRadPerSec = 0.0f;   // :59-61

        private int _rank = 0, _worldSize = 1;                  // multi-GPU (new): one process per GPU
        private bool _multiGpu;

        public RTRenderer(RTWindow window, int deviceIndex = 0)  // :63-92
        {
            _window = window ?? throw new ArgumentNullException(nameof(window));
            _device = new RtDevice(deviceIndex);                 // Context.Create(...Cuda()...) + CreateCudaAccelerator(deviceIndex) :66-68
            _sceneManager = new SceneManager(_device);           // :70
            _sceneManager.BuildDefaultScene();                   // :71
            _sceneManager.Commit(RebuildPolicy.Auto);            // :72
            _cameraController = new FlyCameraController(_window);
            int w = Math.Max(1, _window.Size.X), h = Math.Max(1, _window.Size.Y);
            _camera = Camera.CreateCamera(w, h, 60f);            // :78
            _camera.Translate(new Float3(1, 0, -4));             // :79
            _prevCamera = _camera;
            _framebuffer = new Framebuffer(_device);
            // no kernels to load (:85-89) and no GBuffer / RTTaa objects (:83,91): G-buffer, reservoirs and TAA history live in the native context
        }

        public RtDevice Accelerator => _device;                  // :94 - same name, so RTWindow.CreatePbos (:320) is untouched
        public void UpdateCamera(float dtSeconds) => _cameraController.Update(ref _camera, dtSeconds);                       // :96
        public void SetSunParams(float speedRadPerSec, float elevationRad) { _sunSpeedRadPerSec = speedRadPerSec; _sunElevation = elevationRad; }   // :99-103

        // Multi-GPU (new).  Rank 0 creates the id, the host hands the 128 bytes to the other processes (file, pipe, MPI ...):
        //     byte[] id = RTRenderer.NewCommunicatorId();                  // rank 0
        //     renderer.InitMultiGpu(id, rank, worldSize);                  // every rank, same id
        // From then on RenderDirectToPbo renders this rank's interleaved screen tiles, gathers colour + depth + objectId on rank 0
        // (NCCL inside the library) and presents there; the other ranks pass pbo = null.
        public static byte[] NewCommunicatorId()
        {
            var id = new byte[RtNative.CommIdBytes];
            fixed (byte* p = id) RtNative.ThrowIfFailed(RtNative.rt_comm_get_unique_id(p, (UIntPtr)id.Length));
            return id;
        }
        public void InitMultiGpu(byte[] uniqueId, int rank, int worldSize)
        {
            if (uniqueId is null) throw new ArgumentNullException(nameof(uniqueId));
            if (uniqueId.Length != RtNative.CommIdBytes) throw new ArgumentOutOfRangeException(nameof(uniqueId));
            fixed (byte* p = uniqueId) RtNative.ThrowIfFailed(RtNative.rt_comm_init(_device.Handle, p, (UIntPtr)uniqueId.Length, rank, worldSize));
            _rank = rank; _worldSize = worldSize; _multiGpu = true;
            _framebuffer.Gathered = worldSize > 1;
        }

        public void RenderDirectToPbo(CudaGlInteropIndexBuffer pbo, int width, int height, int frame, float dt)   // :105-237
        {
            bool presents = _worldSize <= 1 || _rank == 0;
            if (pbo is null && presents) throw new ArgumentNullException(nameof(pbo));
            int outW = Math.Max(1, width), outH = Math.Max(1, height);
            int inW = Math.Max(1, (int)MathF.Round(outW * _renderScale, MidpointRounding.ToEven));   // :113-116 (XMath.Round)
            int inH = Math.Max(1, (int)MathF.Round(outH * _renderScale, MidpointRounding.ToEven));

            BakeCameraDerived(ref _camera, inW, inH);                                                  // :123-124 (method :241-263 unchanged)
            BakeCameraDerived(ref _prevCamera, inW, inH);

            int temporalSeed = (_rngLockNoise == 0) ? 0 : Random.Shared.Next(int.MinValue, int.MaxValue);   // :166
            float dtClamped = Math.Clamp(dt, 0f, 0.1f);                                                // :169-172
            _sunAzimuth += _sunSpeedRadPerSec * dtClamped;
            const float TwoPi = 6.28318530717958647692f;
            if (_sunAzimuth >= TwoPi) _sunAzimuth -= TwoPi; else if (_sunAzimuth < 0f) _sunAzimuth += TwoPi;
            Float3 sunDir = Float3.Normalize(new Float3(MathF.Cos(_sunAzimuth) * MathF.Cos(_sunElevation), MathF.Sin(_sunElevation), MathF.Sin(_sunAzimuth) * MathF.Cos(_sunElevation)));   // :174-178

            var cfg = new RtRenderConfig
            {
                width = inW, height = inH, frame = frame, spp = _spp, maxDepth = 3,                    // :181-205 (maxDepth was SpecializedValue.New(3), :204)
                rngLockNoise = temporalSeed,
                enableTemporalReuse = _enableTemporalReuse, enableSpatialReuse = _enableSpatialReuse,
                dirLightDir = sunDir, dirLightRadiance = new Float3(10, 10, 10),                       // :191-192
                skyTintTop = new Float3(0.5f, 0.7f, 1.0f), skyTintBottom = new Float3(1.0f, 1.0f, 1.0f),   // :193-194
                flags = (uint)RtFlags.None, tileSize = 16, rank = _rank, worldSize = _worldSize, samplesPerPass = 0
            };
            Camera cam = _camera, prev = _prevCamera;
            IntPtr rt = _device.Handle;
            RtNative.ThrowIfFailed(RtNative.rt_render(rt, &cam, &prev, &cfg));                         // replaces _primaryKernel :152-153 and _integratorKernel :205

            if (_multiGpu && _worldSize > 1)                                                           // every rank: its tiles to rank 0
                RtNative.ThrowIfFailed(RtNative.rt_gather_frame(rt, 0, (uint)(RtGather.Rgba8 | RtGather.DepthObjId)));

            if (presents)
            {
                pbo.MapCuda();                                                                         // :208-209
                var pc = new RtPresentConfig
                {
                    mode = (int)(_enableTAAU ? RtPresentMode.Taau : RtPresentMode.Copy),               // :211-231: RTTaa.ResolveUpsample, or blit / bilinear upsample
                    outWidth = outW, outHeight = outH, feedback = 0.075f, sharpness = 0.10f, clampK = 1.25f   // RTTaa.cs:80-82
                };
                RtNative.ThrowIfFailed(RtNative.rt_present(rt, &pc, pbo.DevicePointer, (UIntPtr)((long)outW * outH * 4)));
            }
            _device.Synchronize();                                                                     // _cuda.Synchronize() :233
            if (presents) pbo.UnmapCuda();
            _prevCamera = _camera;                                                                     // :236
        }

        public void Dispose()                                                                          // :347-375
        {
            try { _device.Synchronize(); } catch { }
            if (_cameraController is IDisposable d) d.Dispose();
            _sceneManager?.Dispose();
            _framebuffer?.Dispose();
            _device.Dispose();                                                                         // rt_destroy: frees every device buffer the context owns
        }
    }

    // ================================================================================================ SceneManager / BvhManager
    public sealed partial class SceneManager
    {
        // ctors :20-21 - `RtDevice device` where the reference has `CudaAccelerator cuda`; Scene / BuildDefaultScene / LoadObjInstance /
        // Commit (:22-23: `_bvh.BuildOrRefit(_scene, policy)`), ReplaceScene (:30-36) and Dispose (:38) are unchanged
        public SceneManager(RtDevice device) { _device = device ?? throw new ArgumentNullException(nameof(device)); _scene = new Scene(_device); _bvh = new BvhManager(_device, _scene); }
        public SceneManager(RtDevice device, Scene existingScene) { _device = device ?? throw new ArgumentNullException(nameof(device)); _scene = existingScene ?? throw new ArgumentNullException(nameof(existingScene)); _bvh = new BvhManager(_device, _scene); }
        private readonly RtDevice _device;
    }

    public sealed partial class BvhManager
    {
        private readonly RtDevice _device;
        public BvhManager(RtDevice device, Scene scene) { _device = device ?? throw new ArgumentNullException(nameof(device)); _scene = scene ?? throw new ArgumentNullException(nameof(scene)); }   // :25
        // :27 - the signature SceneManager.Commit (:23) and ReplaceScene (:35) call.  The reference ignores the policy; here ForceRefit
        // refits the uploaded tree on the device when only vertex positions moved, everything else is the full commit.
        public void BuildOrRefit(Scene scene, RebuildPolicy policy)
        {
            if (scene == null) throw new ArgumentNullException(nameof(scene));
            _scene = scene;
            if (policy == RebuildPolicy.ForceRefit && scene.CanRefit) scene.RefitUpload(); else scene.UploadAll();
        }
    }

    // ================================================================================================ Scene
    public sealed unsafe partial class Scene
    {
        private readonly RtDevice _device;                                                             // replaces `_cuda` (:39)
        public Scene(RtDevice device) { _device = device ?? throw new ArgumentNullException(nameof(device)); UploadAll(); }   // :60-64

        // :258-279 - the same 15 host lists with the same element layouts, pinned for the call; the native side copies them, derives the
        // reference's visiting order from the BVH2 arrays (tie-break ranks) and builds its compressed 8-wide BVH
        public void UploadAll()
        {
            var tlas = _hTLASNodes ?? Array.Empty<TLASNode>(); var tlasIdx = _hTLASInstanceIndices ?? Array.Empty<int>(); var inst = _hInstances ?? Array.Empty<InstanceRecord>();
            var blas = _hBLASNodes.ToArray(); var sIdx = _hSpherePrimIndices.ToArray(); var sph = _hSpheres.ToArray(); var tIdx = _hTriPrimIndices.ToArray();
            var pos = _hMeshPositions.ToArray(); var tris = _hMeshTris.ToArray(); var uvs = _hMeshTexcoords.ToArray(); var tuv = _hMeshTriUVs.ToArray();
            var tmat = _hTriMaterialIndex.ToArray(); var mats = _hMaterials.ToArray(); var tex = _hTexels.ToArray(); var ti = _hTexInfos.ToArray();
            fixed (TLASNode* p0 = tlas) fixed (int* p1 = tlasIdx) fixed (InstanceRecord* p2 = inst) fixed (BLASNode* p3 = blas) fixed (int* p4 = sIdx)
            fixed (Sphere* p5 = sph) fixed (int* p6 = tIdx) fixed (Float3* p7 = pos) fixed (MeshTri* p8 = tris) fixed (Float2* p9 = uvs)
            fixed (MeshTriUV* p10 = tuv) fixed (int* p11 = tmat) fixed (MaterialRecord* p12 = mats) fixed (RGBA32* p13 = tex) fixed (TexInfo* p14 = ti)
            {
                var d = new RtSceneDesc
                {
                    tlasNodes = p0, nTlasNodes = tlas.Length, tlasInstanceIndices = p1, nTlasInstanceIndices = tlasIdx.Length, instances = p2, nInstances = inst.Length,
                    blasNodes = p3, nBlasNodes = blas.Length, spherePrimIdx = p4, nSpherePrimIdx = sIdx.Length, spheres = p5, nSpheres = sph.Length,
                    triPrimIdx = p6, nTriPrimIdx = tIdx.Length, meshPositions = p7, nMeshPositions = pos.Length, meshTris = p8, nMeshTris = tris.Length,
                    meshTexcoords = p9, nMeshTexcoords = uvs.Length, meshTriUVs = p10, nMeshTriUVs = tuv.Length, triMatIndex = p11, nTriMatIndex = tmat.Length,
                    materials = p12, nMaterials = mats.Length, texels = p13, nTexels = tex.Length, texInfos = p14, nTexInfos = ti.Length
                };
                // DeviceBuild: the wide BVH is built on the GPU (Morton order, radix tree / PLOC, SAH-optimal 8-wide collapse): ~15x faster commit, traversal within ~4 %
                RtNative.ThrowIfFailed(DeviceBuild ? RtNative.rt_scene_upload_ex(_device.Handle, &d, 1u /* RT_BUILD_DEVICE_LBVH */) : RtNative.rt_scene_upload(_device.Handle, &d));
                _uploadedVersion = _topologyVersion;
            }
        }

        // Optional additions that give RebuildPolicy (BvhManager.cs:13-18) a meaning; the reference ignores the policy (:27).
        // _topologyVersion is bumped by every method that changes more than vertex positions (AddSphere, LoadObjInstance, ...).
        public bool DeviceBuild;
        private long _topologyVersion, _uploadedVersion = -1;
        public bool CanRefit => _uploadedVersion == _topologyVersion && _hMeshPositions.Count > 0;
        public void SetMeshPositions(ReadOnlySpan<Float3> positions)
        {
            if (positions.Length != _hMeshPositions.Count) throw new ArgumentOutOfRangeException(nameof(positions));
            for (int i = 0; i < positions.Length; i++) _hMeshPositions[i] = positions[i];
        }
        public void RefitUpload()
        {
            var pos = _hMeshPositions.ToArray();
            fixed (Float3* p = pos) RtNative.ThrowIfFailed(RtNative.rt_scene_refit(_device.Handle, p, pos.Length));
        }
        public void Dispose() { }   // :281-300 disposed 15 device buffers; they belong to the native context now
    }

    // ================================================================================================ Framebuffer
    public sealed unsafe class Framebuffer : IDisposable
    {
        private readonly RtDevice _device;
        private int[] _cpuColor = Array.Empty<int>(); private float[] _cpuDepth = Array.Empty<float>(); private int[] _cpuObjectId = Array.Empty<int>();
        internal bool Gathered;   // multi-GPU: the frame is the image rt_gather_frame assembled on rank 0

        public Framebuffer(RtDevice device) { _device = device ?? throw new ArgumentNullException(nameof(device)); }   // :53-57

        public void DownloadToCpu(int slot)                                                            // :148-156 (the reference allocates 3 slots and only ever uses slot 0, RTRenderer.cs:164)
        {
            if (slot != 0) throw new ArgumentOutOfRangeException(nameof(slot));
            IntPtr rt = _device.Handle;
            int bc = (int)(Gathered ? RtBuffer.GatheredRgba8 : RtBuffer.Rgba8), bd = (int)(Gathered ? RtBuffer.GatheredDepth : RtBuffer.Depth), bo = (int)(Gathered ? RtBuffer.GatheredObjId : RtBuffer.ObjId);
            RtNative.ThrowIfFailed(RtNative.rt_buffer_bytes(rt, bc, out UIntPtr bytes));
            int n = (int)((ulong)bytes / 4);
            if (_cpuColor.Length != n) { _cpuColor = new int[n]; _cpuDepth = new float[n]; _cpuObjectId = new int[n]; }
            fixed (int* c = _cpuColor) fixed (float* z = _cpuDepth) fixed (int* o = _cpuObjectId)
            {
                RtNative.ThrowIfFailed(RtNative.rt_download(rt, bc, c, bytes));
                RtNative.ThrowIfFailed(RtNative.rt_download(rt, bd, z, bytes));
                RtNative.ThrowIfFailed(RtNative.rt_download(rt, bo, o, bytes));
            }
        }
        public int[] CpuColor(int slot) { if (slot != 0) throw new ArgumentOutOfRangeException(nameof(slot)); return _cpuColor; }       // :158
        public float[] CpuDepth(int slot) { if (slot != 0) throw new ArgumentOutOfRangeException(nameof(slot)); return _cpuDepth; }     // :159
        public int[] CpuObjectId(int slot) { if (slot != 0) throw new ArgumentOutOfRangeException(nameof(slot)); return _cpuObjectId; } // :160
        public void Dispose() { }
    }

    // ================================================================================================ CudaGlInteropIndexBuffer
    // GL PixelUnpackBuffer registered with CUDA (Engine/CudaGlInteropIndexBuffer.cs:38-176).  The GL side is the reference's own; the
    // CUDA side goes through rt_gl_* (the driver-API calls of :18-34, issued by the native library on the context's stream), so the
    // class needs neither ILGPU's MemoryBuffer base (:38,44-45) nor a CudaStream (:62,90).
    public sealed unsafe class CudaGlInteropIndexBuffer : IDisposable
    {
        private readonly RtDevice _device;
        private IntPtr _cudaResource;
        public int glBufferHandle;
        private readonly int _elementCount;
        private bool _mapped;
        public IntPtr DevicePointer { get; private set; }      // valid while mapped (replaces GetCudaArrayView :76-88)

        public CudaGlInteropIndexBuffer(int elementCount, RtDevice accelerator)                        // :44-60
        {
            _device = accelerator ?? throw new ArgumentNullException(nameof(accelerator));
            _elementCount = elementCount;
            glBufferHandle = GL.GenBuffer();
            GL.BindBuffer(BufferTarget.PixelUnpackBuffer, glBufferHandle);
            GL.BufferData(BufferTarget.PixelUnpackBuffer, elementCount * sizeof(int), IntPtr.Zero, BufferUsageHint.StreamDraw);
            GL.BindBuffer(BufferTarget.PixelUnpackBuffer, 0);
            RtNative.ThrowIfFailed(RtNative.rt_gl_register_buffer(_device.Handle, (uint)glBufferHandle, out _cudaResource));   // cuGraphicsGLRegisterBuffer, WriteDiscard :55-56
        }

        public void MapCuda()                                                                          // :62-74
        {
            if (_mapped) return;
            RtNative.ThrowIfFailed(RtNative.rt_gl_map(_device.Handle, _cudaResource, out IntPtr p, out UIntPtr bytes));
            System.Diagnostics.Trace.Assert((ulong)bytes == (ulong)_elementCount * sizeof(int));       // :84
            DevicePointer = p; _mapped = true;
        }
        public void UnmapCuda()                                                                        // :90-103
        {
            if (!_mapped) return;
            RtNative.ThrowIfFailed(RtNative.rt_gl_unmap(_device.Handle, _cudaResource));
            DevicePointer = IntPtr.Zero; _mapped = false;
        }
        public bool IsValid() => glBufferHandle != 0 && GL.IsBuffer(glBufferHandle);                   // :105

        public void Dispose()                                                                          // :140-174
        {
            try { if (_mapped) UnmapCuda(); } catch { /* best effort */ }
            try { if (_cudaResource != IntPtr.Zero) { RtNative.rt_gl_unregister(_device.Handle, _cudaResource); _cudaResource = IntPtr.Zero; } } catch { /* best effort */ }
            if (glBufferHandle != 0) { GL.BindBuffer(BufferTarget.PixelUnpackBuffer, 0); GL.DeleteBuffer(glBufferHandle); glBufferHandle = 0; }
        }
    }
}
