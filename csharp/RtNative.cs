// RtNative.cs — P/Invoke binding of librtcore_b200 (include/rtcore_b200.h) for the reference's C# Engine.
//
// Drop into ILGPU_Raytracing/Engine/.  Follows the precedent the reference already has for native calls:
// the `DllImport("nvcuda", EntryPoint = ...)` static class in Engine/CudaGlInteropIndexBuffer.cs:18-34.
// The element structs the reference already defines (Float3, Affine3x4, TLASNode/BLASNode, InstanceRecord, Sphere,
// MaterialRecord, MeshTri, MeshTriUV, Float2, RGBA32, TexInfo, Camera) are blittable and byte-identical to the Rt*
// structs of the header, so they are passed as they are.
//
// NOTE: this image has no .NET toolchain; this file is shipped as source and has not been compiled here.  Its layout
// contract is checked from the native side (static_asserts at the end of csrc/rtcore.cu, tests/test_host.py).
using System;
using System.Runtime.InteropServices;

namespace ILGPU_Raytracing.Engine
{
    public sealed class RtNativeException : Exception
    {
        public int Status { get; }
        public RtNativeException(int status, string message) : base(message) { Status = status; }
    }

    [Flags]
    public enum RtFlags : uint
    {
        None = 0, TriMaterials = 1u << 0, Accumulate = 1u << 1, ResetAccum = 1u << 2, PathAovs = 1u << 3, Counters = 1u << 4, KernelTiming = 1u << 5,
        ResetReservoirs = 1u << 6, PublishReservoirs = 1u << 7, FastShading = 1u << 8, FrameGraph = 1u << 9
    }

    [Flags]
    public enum RtGather : uint { Rgba8 = 1u << 0, Radiance = 1u << 1, DepthObjId = 1u << 2 }

    public enum RtBuffer
    {
        Rgba8 = 0, Depth = 1, ObjId = 2, Radiance = 3, Accum = 4, PrimId = 5, InstId = 6, PrimaryT = 7, SegCount = 8, TermCode = 9, PathHash = 10,
        GbWorldPos = 11, GbNormal = 12, GbBaseColor = 13, GbMatId = 14, TileRadiance = 15, Reservoir = 16, Present = 17,   // Reservoir: Engine/RTRay.cs:171-179 records; Present: what rt_present wrote
        GatheredRgba8 = 18, GatheredDepth = 19, GatheredObjId = 20, GatheredRadiance = 21                                   // the image rt_gather_frame assembled on its root
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct RtSceneDesc   // 15 x (pointer, int64 count), SceneDeviceViews order (Engine/SceneDeviceViews.cs:11-27)
    {
        public TLASNode* tlasNodes; public long nTlasNodes;
        public int* tlasInstanceIndices; public long nTlasInstanceIndices;
        public InstanceRecord* instances; public long nInstances;
        public BLASNode* blasNodes; public long nBlasNodes;
        public int* spherePrimIdx; public long nSpherePrimIdx;
        public Sphere* spheres; public long nSpheres;
        public int* triPrimIdx; public long nTriPrimIdx;
        public Float3* meshPositions; public long nMeshPositions;
        public MeshTri* meshTris; public long nMeshTris;
        public Float2* meshTexcoords; public long nMeshTexcoords;
        public MeshTriUV* meshTriUVs; public long nMeshTriUVs;
        public int* triMatIndex; public long nTriMatIndex;
        public MaterialRecord* materials; public long nMaterials;
        public RGBA32* texels; public long nTexels;
        public TexInfo* texInfos; public long nTexInfos;
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct RtRenderConfig   // 112 bytes
    {
        public int width, height, frame, spp, maxDepth, rngLockNoise, enableTemporalReuse, enableSpatialReuse;
        public Float3 dirLightDir, dirLightRadiance, skyTintTop, skyTintBottom;
        public uint flags;
        public int tileSize, rank, worldSize, samplesPerPass;
        public fixed int reserved[3];
    }

    public enum RtPresentMode { Taau = 0, Copy = 1 }   // Copy = BlitKernel when the sizes match, BilinearUpsampleKernel otherwise

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct RtPresentConfig   // 44 bytes
    {
        public int mode, outWidth, outHeight;
        public float feedback, sharpness, clampK;   // RTTaa.cs:80-82
        public int resetHistory;
        public fixed int reserved[4];
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct RtStats
    {
        public ulong raysPrimary, raysBounce, raysShadow, wideNodes, trisTested, spheresTested, kernelLaunches;
        public float lastRenderMs, lastTraceMs;
        public ulong bvhWideNodeCount, bvhPrimCount, bvhBytes;
        public fixed ulong reserved[4];
    }

    public static unsafe class RtNative
    {
        private const string Lib = "rtcore_b200";   // librtcore_b200.so / rtcore_b200.dll next to the executable

        [DllImport(Lib)] public static extern int rt_abi_version();
        [DllImport(Lib)] public static extern IntPtr rt_last_error();
        [DllImport(Lib)] public static extern int rt_create(int* deviceIds, int nDev, out IntPtr ctx);
        [DllImport(Lib)] public static extern int rt_destroy(IntPtr ctx);
        [DllImport(Lib)] public static extern int rt_set_stream(IntPtr ctx, IntPtr cudaStream);
        [DllImport(Lib)] public static extern int rt_get_stream(IntPtr ctx, out IntPtr cudaStream);
        [DllImport(Lib)] public static extern int rt_scene_upload(IntPtr ctx, RtSceneDesc* scene);
        [DllImport(Lib)] public static extern int rt_scene_upload_ex(IntPtr ctx, RtSceneDesc* scene, uint buildFlags);   // 1 = RT_BUILD_DEVICE_LBVH
        // BvhManager.BuildOrRefit(RebuildPolicy.ForceRefit): moved vertices for the uploaded topology, refitted on the device
        [DllImport(Lib)] public static extern int rt_scene_refit(IntPtr ctx, Float3* meshPositions, long nMeshPositions);
        [DllImport(Lib)] public static extern int rt_render(IntPtr ctx, Camera* cam, Camera* prevCam, RtRenderConfig* cfg);
        [DllImport(Lib)] public static extern int rt_sync(IntPtr ctx);
        [DllImport(Lib)] public static extern int rt_download(IntPtr ctx, int which, void* dstHost, UIntPtr bytes);
        [DllImport(Lib)] public static extern int rt_download_async(IntPtr ctx, int which, void* dstHost, UIntPtr bytes);   // queued; valid after rt_sync
        [DllImport(Lib)] public static extern int rt_buffer_bytes(IntPtr ctx, int which, out UIntPtr bytes);
        [DllImport(Lib)] public static extern int rt_get_device_buffer(IntPtr ctx, int which, out IntPtr devPtr, out UIntPtr bytes);
        [DllImport(Lib)] public static extern int rt_map_external_color(IntPtr ctx, IntPtr devPtr, UIntPtr bytes);
        [DllImport(Lib)] public static extern int rt_bind_readback(IntPtr ctx, int which, void* hostPinned, UIntPtr bytes);   // every frame -> page-locked host arrays, overlapped
        [DllImport(Lib)] public static extern int rt_present(IntPtr ctx, RtPresentConfig* cfg, IntPtr dstDevRgba8, UIntPtr dstBytes);
        [DllImport(Lib)] public static extern int rt_tiles_owned_pixels(int width, int height, int tileSize, int rank, int worldSize, out long nPixels);
        [DllImport(Lib)] public static extern int rt_deinterleave_tiles(IntPtr ctx, IntPtr gatheredDev, long* rankOffsetsPx, int worldSize, int width, int height, int tileSize, IntPtr outRadianceDev, IntPtr outRgba8Dev);
        [DllImport(Lib)] public static extern int rt_get_stats(IntPtr ctx, out RtStats stats);
        // multi-GPU: one process + one context per GPU, the NCCL communicator lives in the library
        public const int CommIdBytes = 128;
        [DllImport(Lib)] public static extern int rt_comm_get_unique_id(void* id, UIntPtr bytes);
        [DllImport(Lib)] public static extern int rt_comm_init(IntPtr ctx, void* id, UIntPtr bytes, int rank, int worldSize);
        [DllImport(Lib)] public static extern int rt_comm_destroy(IntPtr ctx);
        [DllImport(Lib)] public static extern int rt_gather_frame(IntPtr ctx, int root, uint what);
        // CUDA-GL interop of the present target (replaces the DllImport("nvcuda") block of Engine/CudaGlInteropIndexBuffer.cs:18-34)
        [DllImport(Lib)] public static extern int rt_gl_register_buffer(IntPtr ctx, uint glBuffer, out IntPtr resource);
        [DllImport(Lib)] public static extern int rt_gl_map(IntPtr ctx, IntPtr resource, out IntPtr devPtr, out UIntPtr bytes);
        [DllImport(Lib)] public static extern int rt_gl_unmap(IntPtr ctx, IntPtr resource);
        [DllImport(Lib)] public static extern int rt_gl_unregister(IntPtr ctx, IntPtr resource);

        // CudaException.ThrowIfFailed analogue (Engine/CudaGlInteropIndexBuffer.cs:56)
        public static void ThrowIfFailed(int status)
        {
            if (status != 0) throw new RtNativeException(status, Marshal.PtrToStringAnsi(rt_last_error()) ?? "rtcore_b200 error");
        }

        public static IntPtr Create(int deviceIndex)
        {
            if (rt_abi_version() != 1) throw new InvalidOperationException("rtcore_b200 ABI version mismatch");
            ThrowIfFailed(rt_create(&deviceIndex, 1, out IntPtr ctx));
            return ctx;
        }
    }
}
