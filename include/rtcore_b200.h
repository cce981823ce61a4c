/*
 * rtcore_b200.h — C ABI of the B200-native (sm_100a) renderer core.
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * NullandKale/ILGPU_Raytracing (camera ray generation -> BVH traversal with
 * ray/sphere + ray/triangle intersection -> material shading with bounce and
 * RNG -> framebuffer accumulation and tone-map).  Everything ILGPU did for that
 * path (device buffers, kernel load, kernel launch) is behind these entry
 * points; the C# Engine (Scene / SceneManager / Camera / RTRenderer /
 * Framebuffer) keeps its public surface and P/Invokes into this library
 * (binding shown in INTEGRATION.md, C# source under csharp/).
 *
 * All citations are file:line under /root/reference/ILGPU_Raytracing/.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns RT_OK (0) or a negative RtStatus; rt_last_error()
 *     returns a thread-local human readable message for the last failure.
 *     Nothing throws across the ABI (the C# wrapper turns a negative status
 *     into an exception the way CudaException.ThrowIfFailed does,
 *     Engine/CudaGlInteropIndexBuffer.cs:56).
 *   - an rt_ctx is NOT thread safe (the reference is single threaded on the GL
 *     thread: Engine/RTWindow.cs:148-205).  One rt_ctx drives one GPU; multi-GPU
 *     is one process (and one rt_ctx) per GPU with screen-tile partitioning and an
 *     NCCL communicator inside the library (rt_comm_init / rt_gather_frame).
 *   - host arrays passed in are borrowed for the duration of the call only.
 *   - rt_render() is asynchronous on the context's stream; rt_sync() /
 *     rt_download() synchronise (the reference does one Synchronize() per frame:
 *     Engine/RTRenderer.cs:233).
 *   - There is NO CPU fallback: rt_create() fails with RT_ERR_NO_DEVICE when no
 *     sm_100 CUDA device is usable.
 */
#ifndef RTCORE_B200_H
#define RTCORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RT_API __declspec(dllexport)
#else
#define RT_API __attribute__((visibility("default")))
#endif

#define RT_ABI_VERSION 1

/* ------------------------------------------------------------------------- */
/* Blittable element layouts == the reference's device element layouts.       */
/* (sizes are static-asserted at the end of csrc/rtcore.cu and mirrored in C#)       */
/* ------------------------------------------------------------------------- */

/* Engine/Float3.cs:6-10 (12 B) */
typedef struct RtFloat3 { float X, Y, Z; } RtFloat3;
/* Engine/MeshLoaderOBJ.cs:33 (8 B) */
typedef struct RtFloat2 { float X, Y; } RtFloat2;
/* Engine/Affine3x4.cs:3-7 (48 B, row major 3x4) */
typedef struct RtAffine3x4 {
    float m00, m01, m02, m03;
    float m10, m11, m12, m13;
    float m20, m21, m22, m23;
} RtAffine3x4;

/* TLASNode / BLASNode, Engine/Scene.cs:705-739 (44 B each, identical layout) */
typedef struct RtBvhNode {
    RtFloat3 boundsMin;
    RtFloat3 boundsMax;
    int32_t left, right, first, count, skipIndex;
} RtBvhNode;

/* BlasType, Engine/Scene.cs:703 */
enum { RT_BLAS_SPHERESET = 1, RT_BLAS_TRIMESH = 2 };

/* InstanceRecord, Engine/Scene.cs:716-728 (144 B) */
typedef struct RtInstanceRecord {
    int32_t type;            /* RT_BLAS_* */
    int32_t blasRoot;
    int32_t blasNodeCount;
    int32_t primIndexFirst;
    int32_t primIndexCount;
    RtAffine3x4 objectToWorld;
    RtAffine3x4 worldToObject;
    float uniformScale;
    RtFloat3 worldBoundsMin;
    RtFloat3 worldBoundsMax;
} RtInstanceRecord;

/* Shading codes, Engine/Sphere.cs:5-7, Engine/MeshLoaderOBJ.cs:46-48 */
enum { RT_SHADING_LAMBERT = 0, RT_SHADING_MIRROR = 1, RT_SHADING_GLASS = 2 };

/* MaterialRecord, Engine/MeshLoaderOBJ.cs:44-63 (44 B; the "48 B" comment there is wrong) */
typedef struct RtMaterialRecord {
    RtFloat3 Kd;
    int32_t HasDiffuseMap;
    int32_t DiffuseTexIndex;
    int32_t Shading;
    float IOR;
    int32_t HasAlphaMap;
    int32_t AlphaTexIndex;
    int32_t TwoSided;
    float AlphaCutoff;
} RtMaterialRecord;

/* Sphere, Engine/Sphere.cs:3-15 (80 B) */
typedef struct RtSphere {
    RtFloat3 center;
    float radius;
    RtFloat3 albedo;
    RtMaterialRecord material;
    int32_t shading;
    float ior;
} RtSphere;

/* MeshTri / MeshTriUV, Engine/Scene.cs:741, Engine/MeshLoaderOBJ.cs:34 (12 B) */
typedef struct RtMeshTri { int32_t i0, i1, i2; } RtMeshTri;
typedef struct RtMeshTriUV { int32_t t0, t1, t2; } RtMeshTriUV;
/* RGBA32 / TexInfo, Engine/Scene.cs:743-745 */
typedef struct RtRGBA32 { uint8_t R, G, B, A; } RtRGBA32;
typedef struct RtTexInfo { int32_t Offset, Width, Height; } RtTexInfo;

/* Camera, Engine/Camera.cs:5-17 (92 B).  The kernels read origin/lowerLeft/
 * horizontal/vertical (Engine/RTUtils.cs:13-17); forward/right/up/aspect/
 * fovYRadians only feed temporal reprojection (Engine/RTRay.cs:339-360). */
typedef struct RtCamera {
    RtFloat3 origin, lowerLeft, horizontal, vertical;
    RtFloat3 forward, right, up;
    float aspect, fovYRadians;
} RtCamera;

/* The 15 arrays of SceneDeviceViews (Engine/SceneDeviceViews.cs:11-27) as host
 * pointer + element count, exactly what Scene.UploadAll copies to the device
 * (Engine/Scene.cs:258-279).  Empty arrays may be passed as (NULL, 0); the
 * library applies the reference's AllocateOrEmpty rule (one zeroed element,
 * Engine/Scene.cs:370-377). */
typedef struct RtSceneDesc {
    const RtBvhNode*        tlasNodes;           int64_t nTlasNodes;
    const int32_t*          tlasInstanceIndices; int64_t nTlasInstanceIndices;
    const RtInstanceRecord* instances;           int64_t nInstances;
    const RtBvhNode*        blasNodes;           int64_t nBlasNodes;
    const int32_t*          spherePrimIdx;       int64_t nSpherePrimIdx;
    const RtSphere*         spheres;             int64_t nSpheres;
    const int32_t*          triPrimIdx;          int64_t nTriPrimIdx;
    const RtFloat3*         meshPositions;       int64_t nMeshPositions;
    const RtMeshTri*        meshTris;            int64_t nMeshTris;
    const RtFloat2*         meshTexcoords;       int64_t nMeshTexcoords;
    const RtMeshTriUV*      meshTriUVs;          int64_t nMeshTriUVs;
    const int32_t*          triMatIndex;         int64_t nTriMatIndex;
    const RtMaterialRecord* materials;           int64_t nMaterials;
    const RtRGBA32*         texels;              int64_t nTexels;
    const RtTexInfo*        texInfos;            int64_t nTexInfos;
} RtSceneDesc;

/* rt_render flags */
enum {
    /* Extension (off = reference-faithful): triangles take shade/ior from their
     * MaterialRecord instead of the forced Lambert of SceneDeviceViews.cs:61. */
    RT_FLAG_TRI_MATERIALS = 1u << 0,
    /* Progressive accumulation (new functionality):
     * accum.rgb += Lout, accum.w += 1; RGBA8 = PackRGBA8(accum.rgb / accum.w). */
    RT_FLAG_ACCUMULATE    = 1u << 1,
    RT_FLAG_RESET_ACCUM   = 1u << 2,
    /* Keep per-(pixel,sample) parity AOVs: segment count, terminator, path hash. */
    RT_FLAG_PATH_AOVS     = 1u << 3,
    /* Count rays / nodes / primitives on the device (slower; for roofline). */
    RT_FLAG_COUNTERS      = 1u << 4,
    /* Bracket every extend (traversal) launch with CUDA events; rt_get_stats reports their sum. */
    RT_FLAG_KERNEL_TIMING = 1u << 5,
    /* Zero both ReSTIR reservoir buffers before this frame (start of a sequence).  The reference never clears them:
     * frame 0 with reuse on reads uninitialised memory there (Engine/Framebuffer.cs:85-97); here they start at zero. */
    RT_FLAG_RESET_RESERVOIRS = 1u << 6,
    /* Publish resCur (the reservoir of every sample's first Lambert vertex, Engine/RTRay.cs:289-296) on a frame that imports
     * nothing, so that switching reuse on at the NEXT frame finds this frame's reservoirs like the reference does.  Frames with a
     * reuse flag set always publish.  Without it a reuse frame whose predecessor (frame - 1) did not publish imports zeros
     * (never stale reservoirs of some older frame). */
    RT_FLAG_PUBLISH_RESERVOIRS = 1u << 7,
    /* Tolerance mode (off = bit-exact): the eight sky candidates of ReSTIR_Direct (Engine/RTRay.cs:452-462) are scored with fused
     * multiply-adds and hardware sin / cos / sqrt / reciprocal.  Nothing in that loop decides the path (every candidate draws exactly
     * three random numbers; bounce directions, Russian roulette and intersections stay exact): hit ids, bounce counts and terminators
     * are unchanged, radiance moves by ~1e-6 relative RMS (north_star allows 1e-4). */
    RT_FLAG_FAST_SHADING = 1u << 8,
    /* Replay the frame's kernel sequence from a CUDA graph (captured once per configuration, node parameters refreshed per frame):
     * small, launch-bound frames (the reference's interactive regime, 858x482 x 2 spp) lose their per-launch gaps.  The image is
     * the same with or without it. */
    RT_FLAG_FRAME_GRAPH = 1u << 9
};

/* Everything the reference passes in GBufferParams / IntegratorParams
 * (Engine/RTRay.cs:112-169) that is not a buffer, plus the JIT-specialised
 * MaxDepth (Engine/RTRenderer.cs:204-205) and the tile partition. */
typedef struct RtRenderConfig {
    int32_t width, height, frame;
    int32_t spp;                 /* IntegratorParams.spp (max(1,spp) samples) */
    int32_t maxDepth;            /* SpecializedValue<int> MaxDepth; 0 = primary only */
    int32_t rngLockNoise;        /* IntegratorParams.rngLockNoise (Engine/RTUtils.cs:116-137) */
    int32_t enableTemporalReuse; /* Engine/RTRay.cs:476 */
    int32_t enableSpatialReuse;  /* Engine/RTRay.cs:486 */
    RtFloat3 dirLightDir, dirLightRadiance;   /* Engine/RTRenderer.cs:174-178,191-192 */
    RtFloat3 skyTintTop, skyTintBottom;       /* Engine/RTRenderer.cs:193-194 */
    uint32_t flags;              /* RT_FLAG_* */
    /* Screen-tile partition for multi-GPU: this context renders only tiles with
     * (tx + 3*ty) % worldSize == rank.  worldSize <= 1 renders the whole image. */
    int32_t tileSize, rank, worldSize;
    int32_t samplesPerPass;      /* wavefront batch of samples; 0 = auto */
    int32_t reserved[3];
} RtRenderConfig;

/* rt_download / rt_get_device_buffer selectors.  Per-pixel buffers are in the
 * reference's index order (index = y*width + x, row 0 = bottom scanline,
 * Engine/RTRay.cs:122-125); with a tile partition the non-owned pixels are
 * left untouched. */
enum {
    RT_BUF_RGBA8        = 0,  /* int32  per px : PackRGBA8 (Engine/RTRay.cs:66-76) */
    RT_BUF_DEPTH        = 1,  /* float  per px : GpuFramebuffer.depth */
    RT_BUF_OBJID        = 2,  /* int32  per px : GpuFramebuffer.objectId (tri id, -1 for spheres/miss) */
    RT_BUF_RADIANCE     = 3,  /* float4 per px : Lout pre-pack (Engine/RTRay.cs:323), w = 1 */
    RT_BUF_ACCUM        = 4,  /* float4 per px : progressive accumulator */
    RT_BUF_PRIM_ID      = 5,  /* int32  per px : primary hit sphere index or global tri index, -1 miss */
    RT_BUF_INST_ID      = 6,  /* int32  per px : primary hit instance index, -1 miss */
    RT_BUF_PRIMARY_T    = 7,  /* float  per px : primary closestT (1e30 on miss) */
    RT_BUF_SEG_COUNT    = 8,  /* uint8  per (sample,px): number of TraceNext calls   (RT_FLAG_PATH_AOVS) */
    RT_BUF_TERM_CODE    = 9,  /* uint8  per (sample,px): RT_TERM_*                   (RT_FLAG_PATH_AOVS) */
    RT_BUF_PATH_HASH    = 10, /* uint32 per (sample,px): fold of hit prim ids        (RT_FLAG_PATH_AOVS) */
    RT_BUF_GB_WORLDPOS  = 11, /* float3 per px : GpuGBuffer.worldPos  (Engine/RTRay.cs:80-109) */
    RT_BUF_GB_NORMAL    = 12, /* float3 per px : GpuGBuffer.normalWS */
    RT_BUF_GB_BASECOLOR = 13, /* float3 per px : GpuGBuffer.baseColor */
    RT_BUF_GB_MATID     = 14, /* int32  per px : GpuGBuffer.matId */
    RT_BUF_TILE_RADIANCE = 15,/* float4 per OWNED px, tile-compacted (multi-GPU gather payload) */
    RT_BUF_PRESENT      = 17, /* int32 per OUTPUT px : what the last rt_present wrote (its own buffer or the mapped PBO) */
    /* the image rt_gather_frame assembled on its root (whole frame, all ranks' tiles) */
    RT_BUF_GATHERED_RGBA8    = 18, /* int32  per px */
    RT_BUF_GATHERED_DEPTH    = 19, /* float  per px (RT_GATHER_DEPTH_OBJID) */
    RT_BUF_GATHERED_OBJID    = 20, /* int32  per px (RT_GATHER_DEPTH_OBJID) */
    RT_BUF_GATHERED_RADIANCE = 21, /* float4 per px (RT_GATHER_RADIANCE): what each pixel shows (Lout, or the progressive mean), w = frames accumulated */
    RT_BUF_RESERVOIR    = 16  /* RtReservoir per px : the reservoir buffer the last frame wrote ("resCur", Engine/RTRay.cs:23-48,294);
                                 only frames rendered with a reuse flag set write reservoirs */
};

/* Reservoir (Engine/RTRay.cs:171-179), the element of RT_BUF_RESERVOIR. */
typedef struct RtReservoir { RtFloat3 L, wi; float pdf, w, wSum; int32_t m, lightId; } RtReservoir;

/* Path terminators reported in RT_BUF_TERM_CODE */
enum {
    RT_TERM_PRIMARY_MISS = 0, /* hitMask == 0 (Engine/RTRay.cs:214-219) */
    RT_TERM_MISS         = 1, /* TraceNext missed -> sky (Engine/RTRay.cs:241-242,272-273,314-315) */
    RT_TERM_MAXDEPTH     = 2, /* depth loop ran out (Engine/RTRay.cs:233) */
    RT_TERM_ROULETTE     = 3  /* Russian roulette kill (Engine/RTRay.cs:306-312) */
};

/* Counters (valid after rt_sync; ray counts always, the rest with RT_FLAG_COUNTERS) */
typedef struct RtStats {
    uint64_t raysPrimary;   /* primary TraceClosest calls (= pixels rendered) */
    uint64_t raysBounce;    /* TraceNext calls (Engine/RTRay.cs:659-671) */
    uint64_t raysShadow;    /* ShadowOcclusion calls (Engine/RTRay.cs:618-624) */
    uint64_t wideNodes;     /* 80-byte wide nodes fetched */
    uint64_t trisTested;    /* 48-byte triangle records tested */
    uint64_t spheresTested; /* 48-byte sphere records tested */
    uint64_t kernelLaunches;/* kernels launched by the last rt_render */
    float    lastRenderMs;  /* CUDA-event time of the last rt_render on its stream */
    float    lastTraceMs;   /* of which: extend (closest + shadow) kernels */
    uint64_t bvhWideNodeCount, bvhPrimCount, bvhBytes;
    uint64_t reserved[4];   /* [0] extend launches timed, [1] any-hit rays traced, [2] shared sun probes, [3] last rt_gather_frame in microseconds (on its stream) */
} RtStats;

typedef struct rt_ctx rt_ctx;

typedef enum RtStatus {
    RT_OK = 0,
    RT_ERR_INVALID_ARGUMENT = -1, /* ArgumentNull/ArgumentOutOfRange analogue */
    RT_ERR_NO_DEVICE        = -2, /* no usable sm_100 CUDA device: there is no CPU fallback */
    RT_ERR_CUDA             = -3, /* a CUDA runtime call failed (message has the cudaError) */
    RT_ERR_INVALID_STATE    = -4, /* InvalidOperationException analogue (e.g. render before scene upload) */
    RT_ERR_UNSUPPORTED      = -5, /* feature outside the hot-path scope built so far */
    RT_ERR_OUT_OF_MEMORY    = -6,
    RT_ERR_NCCL             = -7  /* an NCCL call failed (message has ncclGetErrorString) */
} RtStatus;

/* ---- lifetime: replaces Context.Create(...Cuda()...) + CreateCudaAccelerator(deviceIndex)
 *      (Engine/RTRenderer.cs:66-68) and Dispose (Engine/RTRenderer.cs:347-375). ---- */
RT_API int rt_abi_version(void);
RT_API const char* rt_last_error(void);
RT_API int rt_create(const int* deviceIds, int nDev, rt_ctx** out); /* nDev must be 1 */
RT_API int rt_destroy(rt_ctx* ctx);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the context's own stream. */
RT_API int rt_set_stream(rt_ctx* ctx, void* cudaStream);
/* The cudaStream_t the context currently renders on (its own unless rt_set_stream replaced it): for callers that order their own
 * work or events against the frame (the reference shares one stream between its kernels and the PBO map, Engine/RTRenderer.cs:68,208). */
RT_API int rt_get_stream(rt_ctx* ctx, void** cudaStream);

/* ---- scene commit: replaces Scene.UploadAll (Engine/Scene.cs:258-279) reached through
 *      SceneManager.Commit -> BvhManager.BuildOrRefit (Engine/BvhManager.cs:27).
 *      Copies the arrays, derives the reference's traversal order from the BVH2
 *      arrays (tie-break ranks) and builds the compressed 8-wide BVH. ---- */
RT_API int rt_scene_upload(rt_ctx* ctx, const RtSceneDesc* scene);
/* The same commit with a choice of builder (what RebuildPolicy, Engine/BvhManager.cs:13-18, is there to express):
 * 0 = the default - binned-SAH binary tree + SAH-optimal 8-wide collapse on the host (best traversal, ~1 s per million
 * triangles); RT_BUILD_DEVICE_LBVH = the tree built on the device - Morton order, then the radix tree and PLOC side by side,
 * the SAH-optimal 8-wide collapse of whichever costs less (commit of a million triangles in ~0.02 s, traversal within ~4 % of
 * the host tree's; the scratch stays allocated between device commits).  The images are the same either way. */
enum { RT_BUILD_DEVICE_LBVH = 1u };
RT_API int rt_scene_upload_ex(rt_ctx* ctx, const RtSceneDesc* scene, uint32_t buildFlags);
/* BvhManager.BuildOrRefit(RebuildPolicy.ForceRefit) (Engine/BvhManager.cs:13-27; the reference accepts the policy and
 * ignores it): new mesh vertex positions for the topology of the last rt_scene_upload (same count, same triangles).
 * The triangle records and the wide BVH are refitted on the device, bottom-up; tie-break ranks, materials, spheres and
 * instances are kept.  RT_ERR_INVALID_ARGUMENT when the count differs, RT_ERR_INVALID_STATE without a scene. */
RT_API int rt_scene_refit(rt_ctx* ctx, const RtFloat3* meshPositions, int64_t nMeshPositions);

/* ---- per-frame hot path: replaces the two kernel launches of
 *      RTRenderer.RenderDirectToPbo (Engine/RTRenderer.cs:152-153 PrimaryVisibilityKernel,
 *      :181-205 PathTraceKernel).  prevCam may be NULL when temporal reuse is off. ---- */
RT_API int rt_render(rt_ctx* ctx, const RtCamera* cam, const RtCamera* prevCam,
                     const RtRenderConfig* cfg);
RT_API int rt_sync(rt_ctx* ctx); /* _cuda.Synchronize(), Engine/RTRenderer.cs:233 */

/* ---- read-back: replaces Framebuffer.DownloadToCpu / CpuColor / CpuDepth / CpuObjectId
 *      (Engine/Framebuffer.cs:148-160).  bytes must equal the buffer's size. ---- */
RT_API int rt_download(rt_ctx* ctx, int which, void* dstHost, size_t bytes);
/* The same copy queued on the context's stream without waiting (dstHost should be page-locked and must stay valid until
 * rt_sync): several buffers, one wait.  Framebuffer outputs only (RGBA8, depth, objId, radiance, accumulator, tile payload,
 * presented image); RT_ERR_UNSUPPORTED for the buffers rt_download gathers through a staging buffer. */
RT_API int rt_download_async(rt_ctx* ctx, int which, void* dstHost, size_t bytes);
RT_API int rt_buffer_bytes(rt_ctx* ctx, int which, size_t* bytes);
/* Device pointer of an output buffer (for NCCL / torch plumbing, no copy). */
RT_API int rt_get_device_buffer(rt_ctx* ctx, int which, void** devPtr, size_t* bytes);
/* Persistent read-back targets for hosts that want EVERY frame on the CPU (the pattern "RenderDirectToPbo; Framebuffer.DownloadToCpu",
 * Engine/Framebuffer.cs:148-156, once per frame): `which` = RT_BUF_RGBA8, RT_BUF_DEPTH or RT_BUF_OBJID, hostPinned = a page-locked
 * buffer of the image's size (NULL unbinds).  Every rt_render then copies the output as soon as it is final - depth and objectId right
 * after the primary pass, on a copy stream, overlapping the rest of the frame; RGBA8 behind the frame - and rt_sync waits for the
 * copies.  On the root of a tile partition rt_gather_frame fills the targets with the gathered image instead. */
RT_API int rt_bind_readback(rt_ctx* ctx, int which, void* hostPinned, size_t bytes);
/* Write packed RGBA8 into a caller-owned device buffer (CUDA-mapped PBO):
 * replaces Framebuffer.GetGpuWithExternalColor (Engine/Framebuffer.cs:112-124). NULL unmaps. */
RT_API int rt_map_external_color(rt_ctx* ctx, void* devPtr, size_t bytes);

/* ---- present: replaces the tail of RTRenderer.RenderDirectToPbo (Engine/RTRenderer.cs:208-231): RTTaa.ResolveUpsample
 *      (Engine/RTTaa.cs:49-179: two-tap "Catmull-Rom" upsample in linearised sRGB, 3x3 neighbourhood clamp of the history,
 *      objId disocclusion, temporal blend, light sharpening) when TAAU is on, else BlitKernel / BilinearUpsampleKernel
 *      (Engine/RTRenderer.cs:281-320).  Input = RGBA8 + objId of the last rt_render (width x height of that frame, whole
 *      image: no tile partition); output = outWidth x outHeight RGBA8 into dstDevRgba8 (a CUDA-mapped PBO) or, when that
 *      is NULL, into a buffer of the context (RT_BUF_PRESENT).  The TAA history lives in the context (RTTaa._historyColor /
 *      _historyObjId) and is dropped when the output size changes or resetHistory != 0. ---- */
enum { RT_PRESENT_TAAU = 0, RT_PRESENT_COPY = 1 /* blit when the sizes match, bilinear upsample otherwise */ };
typedef struct RtPresentConfig {
    int32_t mode;                        /* RT_PRESENT_* (RTRenderer._enableTAAU) */
    int32_t outWidth, outHeight;
    float   feedback, sharpness, clampK; /* RTTaa.cs:80-82: 0.075, 0.10, 1.25 */
    int32_t resetHistory;
    int32_t reserved[4];
} RtPresentConfig;
RT_API int rt_present(rt_ctx* ctx, const RtPresentConfig* cfg, void* dstDevRgba8, size_t dstBytes);

/* ---- multi-GPU finish on the gathering rank: scatter the tile-compacted float4
 *      payloads of all ranks (concatenated rank-major, as NCCL gather delivers them)
 *      into the full image, and write Lout float4 + RGBA8 for every pixel. ---- */
RT_API int rt_tiles_owned_pixels(int width, int height, int tileSize, int rank, int worldSize,
                                 int64_t* nPixels);
RT_API int rt_deinterleave_tiles(rt_ctx* ctx, const void* gatheredDev, const int64_t* rankOffsetsPx,
                                 int worldSize, int width, int height, int tileSize,
                                 void* outRadianceDev /* float4*w*h or NULL */,
                                 void* outRgba8Dev   /* int32*w*h or NULL */);

/* ---- CUDA-GL interop for the present target: replaces the reference's driver-API binding and PBO class
 *      (Engine/CudaGlInteropIndexBuffer.cs:18-34 DllImport "nvcuda" cuGraphicsGLRegisterBuffer / MapResources /
 *      ResourceGetMappedPointer_v2 / UnmapResources / UnregisterResource; :44-60 register with WriteDiscard; :62-103 MapCuda /
 *      GetCudaArrayView / UnmapCuda; :150-160 unregister).  The host creates the GL PixelUnpackBuffer and passes its name; map and
 *      unmap run on the context's stream, so a frame is: rt_gl_map -> rt_render -> rt_present(mapped pointer) -> rt_gl_unmap ->
 *      glTexSubImage2D from the PBO (Engine/RTWindow.cs:160-168).  Needs a GL context current on the calling thread (the reference's
 *      GL thread); RT_ERR_CUDA with the driver's error name otherwise. ---- */
RT_API int rt_gl_register_buffer(rt_ctx* ctx, unsigned int glBuffer, void** resource);
RT_API int rt_gl_map(rt_ctx* ctx, void* resource, void** devPtr, size_t* bytes);
RT_API int rt_gl_unmap(rt_ctx* ctx, void* resource);
RT_API int rt_gl_unregister(rt_ctx* ctx, void* resource);

/* ---- multi-GPU behind the ABI.  The reference drives ONE device (ctor `RTRenderer(RTWindow, int deviceIndex = 0)`,
 *      Engine/RTRenderer.cs:63,67) and its framebuffer is colour + depth + objectId (Engine/RTRay.cs:59-64,
 *      Engine/Framebuffer.cs:112-160).  Here one process (and one rt_ctx) drives one GPU of the box; the contexts of a job share an
 *      NCCL communicator owned by the library: rank 0 calls rt_comm_get_unique_id, the host hands the 128 bytes to every rank over
 *      any channel it has (the C# host: a file, a pipe, MPI ...), every rank calls rt_comm_init.  Per frame every rank calls
 *      rt_render with RtRenderConfig.rank / worldSize equal to the communicator's, then rt_gather_frame: grouped ncclSend / ncclRecv
 *      of the tile payloads straight from the buffers the frame's kernels wrote (exact counts, no staging copy on the senders), and
 *      on the root a fused de-interleave + PackRGBA8 into the gathered image: colour, and with RT_GATHER_DEPTH_OBJID depth and
 *      objectId too, so that rt_present (TAAU needs objectId, Engine/RTTaa.cs:117-171) and Framebuffer.DownloadToCpu work on the root
 *      exactly as on one GPU (RT_BUF_GATHERED_*).  The gather runs on the communicator's own stream: the next rt_render overlaps it,
 *      and its depth / objectId phase (final since the primary pass; sent first, in the same order on every rank) overlaps the frame
 *      it belongs to when rt_gather_frame is called right after rt_render, which only queues work.
 *      NCCL is loaded at run time (libnccl.so.2); RT_ERR_UNSUPPORTED when it is absent. ---- */
#define RT_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */
enum {
    RT_GATHER_RGBA8       = 1u << 0, /* 4 B/px: packed colour only (display) */
    RT_GATHER_RADIANCE    = 1u << 1, /* 16 B/px: float4 radiance (exact parity); the root packs RGBA8 from it */
    RT_GATHER_DEPTH_OBJID = 1u << 2  /* + 8 B/px: depth and objectId */
};
RT_API int rt_comm_get_unique_id(void* id, size_t bytes /* RT_COMM_ID_BYTES */);
RT_API int rt_comm_init(rt_ctx* ctx, const void* id, size_t bytes, int rank, int worldSize);
RT_API int rt_comm_destroy(rt_ctx* ctx);
RT_API int rt_gather_frame(rt_ctx* ctx, int root, uint32_t what /* RT_GATHER_* */);

RT_API int rt_get_stats(rt_ctx* ctx, RtStats* out);

#ifdef __cplusplus
}
#endif
#endif /* RTCORE_B200_H */
