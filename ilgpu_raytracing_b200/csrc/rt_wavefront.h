// rt_wavefront.h — the wavefront split of the reference's two megakernels.
//
//   PrimaryVisibilityKernel (RTRay.cs:188-201)  ->  generate_primary + [extend] + primary_finish
//   PathTraceKernel         (RTRay.cs:203-325)  ->  per sample batch:
//        shade_first, then for every depth: [extend closest + shadow rays, one launch] + shade_next (rays that HIT only),
//        then accumulate (settles what is still pending per path - the last shadow ray's contribution, the sky of a bounce ray
//        that left the scene -, then the per-pixel sum over samples in sample order, SafeColor, mean, PackRGBA8).
//
// Deferred settlement (round 2): the extend kernel never reads path state.  A shadow ray's visibility goes to st[slot].c.w and is
// added to Li by whoever touches the path next (shade_next of the bounce ray's hit, or accumulate); a bounce ray that misses only
// leaves its direction in st[slot].miss, and accumulate adds throughput * sky(direction).  Per path the additions happen in the
// reference's order (direct light of vertex d, then whatever vertex d + 1 adds), so Li is bit-identical.
//
// Each function below is the body of one kernel for one work item; rt_kernels.cu wraps them in
// grid-stride / persistent __global__ kernels, tests/hostsim wraps them in plain loops.
// Per-path arithmetic order is exactly the reference's, so Li and the RNG stream are bit-identical
// to a per-pixel loop no matter how the wavefront schedules the paths.
#pragma once
#include "rt_core.h"
#include "rt_traverse.h"

namespace rtx {

struct FrameConst {
    int width, height, frame, spp, maxDepth, rngLockNoise;
    uint32_t flags;
    f3 camOrigin, camLowerLeft, camHorizontal, camVertical;   // the four Camera fields the kernels read (RTUtils.cs:13-17)
    LightEnv env;
    int npx;                 // pixels owned by this context (whole image or its screen tiles)
    const int* pixelMap;     // owned index -> global pixel index y*width+x (8x4 micro-tile order for ray coherence)
    // ReSTIR reuse of the previous frame's reservoirs (RTRay.cs:475-516); everything below is unused when both flags are 0
    int enableTemporal, enableSpatial;
    f3 prevOrigin, prevRight, prevUp, prevForward; float prevFovY, prevAspect;   // the prevCam fields ReprojectToPrevPixel reads (:342-351)
    // G-buffer lookups at OTHER pixels (SpatialCompatible, :365-372): global pixel index -> index into the look* arrays.  One GPU:
    // the context's own G-buffer, lookBase = 0.  Tile partition: the G-buffers of ALL ranks, exchanged after the primary pass and
    // concatenated in rank order; this rank's own pixels start at lookBase.
    const int* invPixelMap;
    const float4 *lookPosHit, *lookNrmMat, *lookAlbObj; int lookBase;
    const float4 *resPrev0, *resPrev1, *resPrev2; // previous frame's reservoirs per global pixel: L|pdf, wi|w, wSum|m|lightId
};

struct alignas(16) PathState {
    float4 thr;    // throughput.xyz | rng state
    float4 li;     // Li.xyz | segCount (bits 0-7), terminator (bits 8-15), path flags (bits 16+)
    float4 c;      // pending direct-light term of the path's latest Lambert vertex: throughput * f/p * W | visibility (written by the any-hit extend)
    float4 miss;   // direction of the path's bounce ray when it left the scene (written by the closest-hit extend) | unused
};

struct WaveBuffers {
    // per owned pixel
    float4 *gbPosHit, *gbNrmMat, *gbAlbObj;   // GpuGBuffer (RTRay.cs:80-109): worldPos|hitMask, normalWS|matId, baseColor|objId
    int *primId, *instId; float* primaryT;    // parity taps of the primary hit
    float4* lframe;                           // Lframe (RTRay.cs:208) across sample batches
    float4* tileRadiance;                     // Lout per owned pixel, tile-compacted (multi-GPU gather payload)
    uint2* tileAux;                           // depth (float bits) | objId per owned pixel (gather payload; may be null)
    int* tileRgba;                            // PackRGBA8 of what the pixel shows, per owned pixel (display-only gather payload; may be null)
    // per global pixel
    int* rgba8; float* depth; int* objId; float4* radiance; float4* accum;
    // per path slot (path j = sampleInBatch * npx + ownedPixel): ONE 64-byte record, so that the scattered accesses of a path's next
    // touch (shade_next of a hit, the extend kernel's visibility / miss stores, accumulate) hit one line instead of four arrays
    PathState* st;
    uint32_t* pathHash;
    // reservoirs (null unless a reuse flag is set): per path slot, the reservoir of the path's first Lambert vertex ("outRes",
    // RTRay.cs:289-296), and per global pixel the frame's resCur, which accumulate() fills from the LAST sample that wrote one
    float4 *resPath0, *resPath1, *resPath2;
    float4 *resCur0, *resCur1, *resCur2;
    // parity AOVs per (sample, global pixel); null unless RT_FLAG_PATH_AOVS
    uint8_t *segCountOut, *termCodeOut; uint32_t* pathHashOut;
};

// Ray queues (SoA of float4): o.w = d.w = path slot (int bits): a consumer that needs only one of the two vectors still gets the
// slot.  32 B per ray: the box-test reciprocal of d is derived by the extend kernel (three divisions per ray against 32 B of
// DRAM traffic per ray it cost to store and re-read it; round 1 stored it).
struct RayQueue { float4* o; float4* d; };
// What the closest-hit extend writes per ray: the primitive index for EVERY ray (4 B; -1 = miss: all the scan of shade_next reads),
// and t | bu | bv only for rays that hit (16 B).
struct HitQueue { int* prim; float4* tuv; };
RT_HD HitRec load_hit(const HitQueue& h, int k) { const float4 v = h.tuv[k]; HitRec r; r.t = v.x; r.prim = h.prim[k]; r.bu = v.y; r.bv = v.z; return r; }
struct ShadowQueue { float4* o; float4* d; };   // the pending contribution and the visibility live per PATH (PathState::c), not per queue entry

RT_HD uint32_t fnv_fold(uint32_t h, uint32_t v) { return (h ^ v) * 16777619u; }

// queue slot allocation: warp-aggregated atomic on the device, plain increment in the host simulator
RT_HD int queue_alloc(int* counter) {
#if defined(__CUDA_ARCH__)
    const unsigned m = __activemask();
    const int lane = (int)(threadIdx.x & 31u);
    const int leader = __ffs((int)m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
#else
    return (*counter)++;
#endif
}

RT_HD void write_ray(float4* qo, float4* qd, int k, const RayOD& r, int path) {
    qo[k] = make_float4(r.o.x, r.o.y, r.o.z, u2f((uint32_t)path));
    qd[k] = make_float4(r.d.x, r.d.y, r.d.z, u2f((uint32_t)path));
}

RT_HD void pixel_xy(const FrameConst& fc, int pix, int* x, int* y) { *x = pix % fc.width; *y = pix / fc.width; }   // RTRay.cs:122
RT_HD f3 primary_dir(const FrameConst& fc, int x, int y) {   // GBufferParams.PrimaryRay RTRay.cs:120-126 + Ray.GenerateRay RTUtils.cs:13-17
    float u = ((float)x + 0.5f) / (float)max(1, fc.width);
    float v = ((float)y + 0.5f) / (float)max(1, fc.height);
    return normalize(fc.camLowerLeft + fc.camHorizontal * u + fc.camVertical * v - fc.camOrigin);
}

// ------------------------------------------------------------------------------------------------ generate
RT_HD void generate_primary(const FrameConst& fc, const RayQueue& q, int i) {
    int x, y; pixel_xy(fc, fc.pixelMap[i], &x, &y);
    f3 d = primary_dir(fc, x, y);
    q.o[i] = make_float4(fc.camOrigin.x, fc.camOrigin.y, fc.camOrigin.z, u2f((uint32_t)i));
    q.d[i] = make_float4(d.x, d.y, d.z, u2f((uint32_t)i));
}

// ------------------------------------------------------------------------------------------------ primary finish
// PrimaryVisibilityKernel after TraceClosest (RTRay.cs:197-200) + GpuGBuffer.StoreHit/StoreMiss (:90-108);
// also the depth / objectId halves of GpuFramebuffer.Store (:59-64, :324), which depend on the G-buffer only.
RT_HD void primary_finish(const FrameConst& fc, const DeviceScene& sc, const WaveBuffers& wb, const RayQueue& q, const HitQueue& hits, int i) {
    const float4 ro = q.o[i], rd = q.d[i];
    const f3 o = mk3(ro.x, ro.y, ro.z), d = mk3(rd.x, rd.y, rd.z);
    HitRec h; h.prim = hits.prim[i]; h.t = 1e30f; h.bu = 0.0f; h.bv = 0.0f;
    if (h.prim >= 0) h = load_hit(hits, i);
    const int pix = fc.pixelMap[i];
    f3 pos; int oid;
    if (!(h.t < 1e29f)) {
        pos = o + d * 1e6f;
        wb.gbPosHit[i] = make_float4(pos.x, pos.y, pos.z, u2f(0u));
        wb.gbNrmMat[i] = make_float4(0.0f, 1.0f, 0.0f, u2f(0xFFFFFFFFu));
        wb.gbAlbObj[i] = make_float4(0.0f, 0.0f, 0.0f, u2f(0xFFFFFFFFu));
        wb.primId[i] = -1; wb.instId[i] = -1; wb.primaryT[i] = h.t;
        oid = -1;
    } else {
        const Surface s = eval_surface(sc, o, d, h);
        pos = o + d * h.t;
        const int packedMat = (s.shade & 0xFFFF) | (float_to_i16(s.ior) << 16);
        wb.gbPosHit[i] = make_float4(pos.x, pos.y, pos.z, u2f(1u));
        wb.gbNrmMat[i] = make_float4(s.normal.x, s.normal.y, s.normal.z, u2f((uint32_t)packedMat));
        wb.gbAlbObj[i] = make_float4(s.albedo.x, s.albedo.y, s.albedo.z, u2f((uint32_t)s.objId));
        wb.primId[i] = s.primId; wb.instId[i] = s.instId; wb.primaryT[i] = h.t;
        oid = s.objId;
    }
    const f3 dc = pos - fc.camOrigin;   // IntegratorParams.DistanceFromCamera RTRay.cs:158-162
    const float dist = sqrtf(dc.x * dc.x + dc.y * dc.y + dc.z * dc.z);
    wb.depth[pix] = dist;
    wb.objId[pix] = oid;
    if (wb.tileAux) wb.tileAux[i] = make_uint2(f2u(dist), (uint32_t)oid);
}

// ------------------------------------------------------------------------------------------------ shared sun probe
// The reference has no sub-pixel jitter: all spp of a pixel start from ONE primary vertex.  When ReSTIR
// selects the directional ("sun") candidate there, Visible() (RTRay.cs:618-624) traces the very same ray - origin
// pos + n * EPS_N, direction normalize(dirLightDir) - for every such sample of the pixel.  The device traces that ray once
// per pixel (sun_probe_generate -> any-hit extend -> sun_probe_store) and the first-vertex shade reads the answer: same
// bits, ~spp x fewer shadow rays at depth 0.  Flags live in gbPosHit.w next to the hit mask (bit 0).
enum : uint32_t { GB_HIT = 1u, GB_SUN_KNOWN = 2u, GB_SUN_VISIBLE = 4u };
RT_HD void sun_probe_generate(const FrameConst& fc, const WaveBuffers& wb, int i, const ShadowQueue& shQ, int* shCount) {
    const float4 ph = wb.gbPosHit[i];
    if ((f2u(ph.w) & GB_HIT) == 0u) return;
    const float4 nm = wb.gbNrmMat[i];
    if (((int)f2u(nm.w) & 0xFFFF) != RT_SHADING_LAMBERT) return;
    const f3 n = normalize(mk3(nm.x, nm.y, nm.z));        // v.nrm of shade_first (:222)
    const f3 wi = normalize(fc.env.dirLightDir);          // the delta candidate's direction (:467)
    if (!(fmaxf(0.0f, dot(n, wi)) > 0.0f) || !(dot(n, wi) > 0.0f)) return;   // restir_finalize would not trace (:524, :620-621)
    const RayOD s = make_ray_normal_offset(mk3(ph.x, ph.y, ph.z), n, wi);   // Visible() :622
    write_ray(shQ.o, shQ.d, queue_alloc(shCount), s, i);
}
// the probe's visibility arrives where every shadow ray's does: st[slot].c.w, slot = the owned pixel (shade_first, which runs after
// this, overwrites the record of the path slots it uses)
RT_HD void sun_probe_store(const WaveBuffers& wb, const ShadowQueue& shQ, int k) {
    const int i = (int)f2u(shQ.o[k].w);
    float4 ph = wb.gbPosHit[i];
    ph.w = u2f(f2u(ph.w) | GB_SUN_KNOWN | (wb.st[i].c.w != 0.0f ? GB_SUN_VISIBLE : 0u));
    wb.gbPosHit[i] = ph;
}

// ------------------------------------------------------------------------------------------------ shade
struct PathVertex { f3 pos, nrm, alb, I; int shade; float ior; };
// The rays a shaded vertex wants queued.  The device kernels collect them per thread and allocate the queue slots once per BLOCK
// (one atomic per queue per 256 paths: with one per warp the two queue counters were the hottest addresses of the frame);
// the host simulator and single-vertex callers push them right away (push_vertex_out).
struct VertexOut { int pushNext, pushShadow; RayOD next, shadow; };

// path flags kept in PathState::li.w above the parity taps (bits 0-7 segment count, 8-15 terminator)
enum : uint32_t {
    PATH_WROTE_RESERVOIR = 1u << 16,
    PATH_PENDING_SHADOW  = 1u << 17,   // st[path].c holds a contribution whose shadow ray has been (or is being) traced and not yet added to Li
    PATH_RAY_IN_FLIGHT   = 1u << 18    // a bounce ray was pushed and no hit of it has been shaded: at accumulate time that means it missed (st[path].miss)
};

// ReprojectToPrevPixel (RTRay.cs:339-360)
RT_HD int reproject_to_prev_pixel(const FrameConst& fc, f3 posWS) {
    f3 p = posWS - fc.prevOrigin;
    float x = dot(p, fc.prevRight), y = dot(p, fc.prevUp), z = dot(p, fc.prevForward);
    if (z <= 1e-4f) return -1;
    float tanHalfFov = tan_p(0.5f * fc.prevFovY);
    float ndcX = x / (z * tanHalfFov * fc.prevAspect);
    float ndcY = y / (z * tanHalfFov);
    float fx = 0.5f * (ndcX + 1.0f) * (float)fc.width;
    float fy = 0.5f * (ndcY + 1.0f) * (float)fc.height;
    int px = (int)fx, py = (int)fy;
    if ((uint32_t)px >= (uint32_t)fc.width || (uint32_t)py >= (uint32_t)fc.height) return -1;
    return py * fc.width + px;
}
RT_HD float distance_from_camera(const FrameConst& fc, float4 pos) {   // IntegratorParams.DistanceFromCamera RTRay.cs:158-162
    const f3 d = mk3(pos.x, pos.y, pos.z) - fc.camOrigin;
    return sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
}
// SpatialCompatible (RTRay.cs:363-374): compares the CURRENT frame's G-buffer at the two pixel indices; iA / iB are owned indices
RT_HD bool spatial_compatible(const FrameConst& fc, const WaveBuffers&, int iA, int iB, f3 nA) {
    const int objA = (int)f2u(fc.lookAlbObj[iA].w), objB = (int)f2u(fc.lookAlbObj[iB].w);
    if (objA == objB) return true;
    const float4 nb4 = fc.lookNrmMat[iB];
    const f3 nB = normalize(mk3(nb4.x, nb4.y, nb4.z));
    const float ndot = dot(nA, nB);
    if (ndot < 0.85f) return false;
    const float zA = distance_from_camera(fc, fc.lookPosHit[iA]);
    const float zB = distance_from_camera(fc, fc.lookPosHit[iB]);
    const float rel = fabsf(zA - zB) / fmaxf(1e-3f, zA);
    return rel < 0.05f;
}
// ImportFromPrevReservoir (RTRay.cs:408-435); curOwned = owned index of the path's pixel, prevIdx = global pixel index
RT_HD void import_from_prev_reservoir(const FrameConst& fc, const WaveBuffers& wb, int prevIdx, int curOwned, f3 n, f3 albedo, uint32_t& rng, Reservoir& r) {
    if (prevIdx < 0 || fc.width * fc.height <= prevIdx) return;
    if (!spatial_compatible(fc, wb, curOwned + fc.lookBase, fc.invPixelMap[prevIdx], n)) return;
    const float4 p1 = fc.resPrev1[prevIdx], p2 = fc.resPrev2[prevIdx];   // (plane 0 = L | pdf is not read by the import)
    const int prM = (int)f2u(p2.y), prLight = (int)f2u(p2.z);
    const float prW = p1.w, prWSum = p2.x;
    if (!(prM > 0 && prW > 0.0f && prWSum > 0.0f)) return;
    const f3 wi = mk3(p1.x, p1.y, p1.z);
    const int lid = prLight == 2 ? 2 : 1;
    const f3 LiImp = (lid == 2) ? fc.env.dirLightRadiance : sky_weighted(fc.env, wi);
    const float nl = fmaxf(0.0f, dot(n, wi));
    const float pdfHere = (lid == 2) ? fmaxf(RTX_EPS_MIN, RTX_MIX_DELTA) : fmaxf(RTX_EPS_MIN, cos_hemisphere_pdf(n, wi) * RTX_MIX_LOCAL);
    const f3 f_over_p = albedo * LiImp * ((nl / pdfHere) * RTX_INV_PI);
    const float sHere = luminance(f_over_p);
    const float Wsrc = prWSum / ((float)max(1, prM) * fmaxf(RTX_EPS_MIN, prW));
    const float eff = sHere * Wsrc;
    reservoir_update(r, wi, pdfHere, LiImp, eff, 1, lid, rng);
}
// steps (3) and (4) of ReSTIR_Direct (RTRay.cs:475-516): temporal import through reprojection, then the 8 rotated neighbours of
// the PREVIOUS frame.  "index" is the path's global pixel index.
RT_HD void restir_imports(const FrameConst& fc, const WaveBuffers& wb, int path, f3 pos, f3 n, f3 albedo, uint32_t& rng, Reservoir& r) {
    const int curOwned = path % fc.npx;
    const int index = fc.pixelMap[curOwned];
    if (fc.enableTemporal != 0) {
        const int prevIdx = reproject_to_prev_pixel(fc, pos);
        if (prevIdx >= 0) import_from_prev_reservoir(fc, wb, prevIdx, curOwned, n, albedo, rng, r);
    }
    if (fc.enableSpatial != 0) {
        const uint32_t h = hash3((uint32_t)index, (uint32_t)fc.frame, 0xB31F5AB1u);
        const int rot = (int)(h & 3u);
        const int rad = 1 + (int)((h >> 2) & 1u);
        const int x0 = index % fc.width, y0 = index / fc.width;
        for (int i = 0; i < 8; i++) {   // Neighbor8 (:377-391): (-r,0) (r,0) (0,-r) (0,r) (-r,-r) (r,-r) (-r,r) (r,r), rotated by rot quarter turns
            const int sx = (i == 0 || i == 4 || i == 6) ? -rad : ((i == 2 || i == 3) ? 0 : rad);
            const int sy = (i < 2) ? 0 : ((i == 2 || i == 4 || i == 5) ? -rad : rad);
            const int dx = rot == 0 ? sx : (rot == 1 ? -sy : (rot == 2 ? -sx : sy));
            const int dy = rot == 0 ? sy : (rot == 1 ? sx : (rot == 2 ? -sy : -sx));
            const int ni = ((uint32_t)(x0 + dx) < (uint32_t)fc.width && (uint32_t)(y0 + dy) < (uint32_t)fc.height) ? (y0 + dy) * fc.width + (x0 + dx) : -1;
            import_from_prev_reservoir(fc, wb, ni, curOwned, n, albedo, rng, r);
        }
    }
}

// One iteration of the depth loop of PathTraceKernel up to (not including) TraceNext (RTRay.cs:233-317).
// Pushes the continuation ray and, for a Lambert vertex, the ReSTIR-selected shadow ray.
// Returns false when Russian roulette killed the path.
// REUSE = a reuse flag is set (compile-time, so the common path does not carry the import code's registers)
template <bool REUSE, bool FAST = false>
// sunFlags: the pixel's GB_SUN_* bits when this is the first vertex of the path and the shared probe applies, else 0;
// *direct receives "throughput * direct" when the probe answered instead of a shadow ray (else stays 0), *probed counts it.
RT_HD bool shade_vertex(const FrameConst& fc, const WaveBuffers& wb, const PathVertex& v, int depth, int path, f3& thr, uint32_t& rng, uint32_t& pflags,
                        VertexOut& vo, uint32_t sunFlags = 0u, f3* direct = nullptr, int* probed = nullptr) {
    RayOD ray;
    if (v.shade == RT_SHADING_MIRROR) {   // :235-244
        f3 dirR = reflect3(v.I, v.nrm);
        ray = make_ray_normal_offset(v.pos, v.nrm, dirR);
        thr = thr * v.alb;
    } else if (v.shade == RT_SHADING_GLASS) {   // :246-275
        f3 Nuse = v.nrm;
        bool outside = dot(v.I, v.nrm) < 0.0f;
        if (!outside) Nuse = Nuse * -1.0f;
        float etaI = outside ? 1.0f : (v.ior > 0.0f ? v.ior : 1.5f);
        float etaT = outside ? (v.ior > 0.0f ? v.ior : 1.5f) : 1.0f;
        f3 dirR = reflect3(v.I, Nuse);
        f3 dirT;
        bool refrOk = refract3(v.I, Nuse, etaI, etaT, &dirT);
        float cosI = fabsf(dot(v.I, Nuse));
        float Fr = schlick_fresnel(cosI, etaI, etaT);
        float xi = rng_next_f(rng);
        ray = (!refrOk || xi < Fr) ? make_ray_normal_offset(v.pos, Nuse, dirR) : make_ray_normal_offset(v.pos, neg(Nuse), dirT);
        if (refrOk && xi >= Fr) {
            f3 transTint = (v.alb.x == 0.0f && v.alb.y == 0.0f && v.alb.z == 0.0f) ? mk3(1.0f, 1.0f, 1.0f) : v.alb;
            float etaScale = (etaI * etaI) / (etaT * etaT);
            thr = thr * transTint * etaScale;
        }
    } else {   // Lambert: ReSTIR-DI + cosine bounce (:277-317)
        f3 wiSel, contrib;
        const Basis B = orthonormal_basis(v.nrm);
        Reservoir r;
#if defined(__CUDA_ARCH__)
        if (FAST) restir_new_candidates_fast(fc.env, v.nrm, B, v.alb, rng, r);   // RT_FLAG_FAST_SHADING (device only)
        else
#endif
        restir_new_candidates(fc.env, v.nrm, B, v.alb, rng, r);
        // only the FIRST Lambert vertex of a sample imports from the previous frame and publishes its reservoir ("wroteReservoir", :280-297)
        const bool first = (pflags & PATH_WROTE_RESERVOIR) == 0u;
        if (REUSE && first) restir_imports(fc, wb, path, v.pos, v.nrm, v.alb, rng, r);
        if (REUSE && first && wb.resPath0) {
            wb.resPath0[path] = make_float4(r.L.x, r.L.y, r.L.z, r.pdf);
            wb.resPath1[path] = make_float4(r.wi.x, r.wi.y, r.wi.z, r.w);
            wb.resPath2[path] = make_float4(r.wSum, u2f((uint32_t)r.m), u2f((uint32_t)r.lightId), 0.0f);
            pflags |= PATH_WROTE_RESERVOIR;
        }
        const bool wantShadow = restir_finalize(fc.env, v.nrm, v.alb, r, &wiSel, &contrib);
        if (wantShadow && (sunFlags & GB_SUN_KNOWN) != 0u && r.lightId == 2) {
            // the selected sample is the sun: its shadow ray is the pixel's shared probe
            if (sunFlags & GB_SUN_VISIBLE) *direct = thr * contrib;
            *probed = (sunFlags & GB_SUN_VISIBLE) ? 2 : 1;
        } else if (wantShadow) {
            RayOD s = make_ray_normal_offset(v.pos, v.nrm, wiSel);   // Visible() :622
            f3 c = thr * contrib;                                    // "Li += throughput * direct" :286,291
            vo.pushShadow = 1; vo.shadow = s;
            wb.st[path].c = make_float4(c.x, c.y, c.z, 0.0f);
            pflags |= PATH_PENDING_SHADOW;
        }
        f3 wi = sample_hemisphere_cosine(v.nrm, B, rng);   // :302
        ray = make_ray_normal_offset(v.pos, v.nrm, wi);
        thr = thr * v.alb;
        if (depth >= 3) {   // :306-312
            float maxC = fmaxf(thr.x, fmaxf(thr.y, thr.z));
            maxC = fmaxf(fminf(maxC, 0.98f), 0.05f);   // XMath.Clamp
            if (rng_next_f(rng) > maxC) { thr = mk3(0.0f, 0.0f, 0.0f); return false; }
            thr = thr * (1.0f / maxC);
        }
    }
    vo.pushNext = 1; vo.next = ray;
    pflags |= PATH_RAY_IN_FLIGHT;
    return true;
}
// immediate push (one slot allocation per call): host simulator, and the fallback of shade_first / shade_next without a VertexOut
RT_HD void push_vertex_out(const VertexOut& vo, int path, const RayQueue& nextQ, int* nextCount, const ShadowQueue& shQ, int* shCount) {
    if (vo.pushShadow) write_ray(shQ.o, shQ.d, queue_alloc(shCount), vo.shadow, path);
    if (vo.pushNext) write_ray(nextQ.o, nextQ.d, queue_alloc(nextCount), vo.next, path);
}

// "Li += throughput * direct" when Visible() (RTRay.cs:286,291,526-537) for the shadow ray queued at the path's latest Lambert
// vertex: runs at the path's next touch, i.e. before anything of the next vertex is added, like the reference's statement order.
RT_HD void settle_pending_shadow(const WaveBuffers& wb, int j, uint32_t& pflags, f3& Li) {
    if ((pflags & PATH_PENDING_SHADOW) == 0u) return;
    pflags &= ~PATH_PENDING_SHADOW;
    const float4 c = wb.st[j].c;
    const bool visible = c.w != 0.0f;
    if (wb.pathHash) wb.pathHash[j] = fnv_fold(wb.pathHash[j], 0x100u | (visible ? 1u : 0u));
    if (visible) { Li.x = Li.x + c.x; Li.y = Li.y + c.y; Li.z = Li.z + c.z; }
}

RT_HD uint32_t pack_aov(int seg, int term) { return (uint32_t)(seg & 0xFF) | ((uint32_t)(term & 0xFF) << 8); }

// depth 0: start every path of the batch from the G-buffer (RTRay.cs:210-232)
template <bool REUSE = false, bool FAST = false>
RT_HD void shade_first(const FrameConst& fc, const WaveBuffers& wb, int sampleBase, int j,
                       const RayQueue& nextQ, int* nextCount, const ShadowQueue& shQ, int* shCount, unsigned* probedCount = nullptr, VertexOut* defer = nullptr) {
    if (defer) { defer->pushNext = 0; defer->pushShadow = 0; }
    const int i = j % fc.npx;
    const int s = sampleBase + j / fc.npx;
    int x, y; pixel_xy(fc, fc.pixelMap[i], &x, &y);
    uint32_t rng = rng_seed_pixel((uint32_t)x, (uint32_t)y, fc.frame, (uint32_t)s, 0xC0FFEEu, fc.rngLockNoise);   // :212
    const float4 ph = wb.gbPosHit[i];
    if ((f2u(ph.w) & GB_HIT) == 0u) return;   // :214-219 - a primary miss has no path state at all: accumulate() adds the sky per sample from the G-buffer flag alone
    if (wb.pathHash) wb.pathHash[j] = 0x811C9DC5u;
    f3 thr = mk3(1.0f, 1.0f, 1.0f);
    if (fc.maxDepth <= 0) {
        wb.st[j].li = make_float4(0.0f, 0.0f, 0.0f, u2f(pack_aov(0, RT_TERM_MAXDEPTH)));
        return;
    }
    const float4 nm = wb.gbNrmMat[i], ao = wb.gbAlbObj[i];
    PathVertex v;
    v.pos = mk3(ph.x, ph.y, ph.z);
    v.nrm = normalize(mk3(nm.x, nm.y, nm.z));          // :222
    v.alb = mk3(ao.x, ao.y, ao.z);
    const int packedMat = (int)f2u(nm.w);
    v.shade = packedMat & 0xFFFF;                      // :225
    v.ior = i16_to_float((packedMat >> 16) & 0xFFFF);  // :226
    v.I = normalize(v.pos - fc.camOrigin);             // ViewDirFromCam :156,230
    uint32_t pflags = 0u;
    // with reuse on, an imported reservoir may carry another frame's sun direction: every sample traces its own ray then
    const uint32_t sunFlags = REUSE ? 0u : (f2u(ph.w) & (GB_SUN_KNOWN | GB_SUN_VISIBLE));
    f3 direct = mk3(0.0f, 0.0f, 0.0f); int probed = 0;
    VertexOut vo; vo.pushNext = 0; vo.pushShadow = 0;
    bool alive = shade_vertex<REUSE, FAST>(fc, wb, v, 0, j, thr, rng, pflags, vo, sunFlags, &direct, &probed);
    if (defer) *defer = vo; else push_vertex_out(vo, j, nextQ, nextCount, shQ, shCount);
    wb.st[j].thr = make_float4(thr.x, thr.y, thr.z, u2f(rng));
    // a probed shadow ray is settled on the spot, before anything else touches Li: Li = 0 + throughput * direct, and the visibility fold of the path hash
    if (probed != 0 && wb.pathHash) wb.pathHash[j] = fnv_fold(wb.pathHash[j], 0x100u | (probed == 2 ? 1u : 0u));
    wb.st[j].li = make_float4(0.0f + direct.x, 0.0f + direct.y, 0.0f + direct.z, u2f(pflags | pack_aov(0, alive ? RT_TERM_MAXDEPTH : RT_TERM_ROULETTE)));
    if (probedCount && probed != 0) (*probedCount)++;
}

// What the extend kernels leave behind when a ray is done (the device kernels write the same words with streaming stores):
//   closest hit: the primitive index for every ray, t | bu | bv for hits, and for a bounce ray that left the scene its direction
//                under the path's slot (settled by accumulate); any hit: the visibility under the path's slot.
RT_HD void store_closest_result(const HitQueue& hq, PathState* st, int k, int slot, const HitRec& h, f3 d) {   // st = null: primary rays (no path yet)
    const bool hit = h.t < 1e29f;   // SceneDeviceViews.cs:85
    hq.prim[k] = hit ? h.prim : -1;
    if (hit) hq.tuv[k] = make_float4(h.t, h.bu, h.bv, 0.0f);
    else if (st) st[slot].miss = make_float4(d.x, d.y, d.z, 0.0f);
}
RT_HD void store_anyhit_result(PathState* st, int slot, bool occluded) { st[slot].c.w = occluded ? 0.0f : 1.0f; }

// depth >= 1: consume the closest-hit result of a ray traced at depth-1 that HIT something (TraceNext, RTRay.cs:659-671) and shade
// the new vertex.  Rays that missed never come here: accumulate() settles them (settle_miss).
template <bool REUSE = false, bool FAST = false>
RT_HD void shade_next(const FrameConst& fc, const DeviceScene& sc, const WaveBuffers& wb, int depth,
                      const RayQueue& curQ, const HitQueue& hits, int k,
                      const RayQueue& nextQ, int* nextCount, const ShadowQueue& shQ, int* shCount, VertexOut* defer = nullptr, int* pathOut = nullptr) {
    if (defer) { defer->pushNext = 0; defer->pushShadow = 0; }
    const float4 ro = curQ.o[k], rd = curQ.d[k];
    const int j = (int)f2u(ro.w);
    if (pathOut) *pathOut = j;
    const f3 o = mk3(ro.x, ro.y, ro.z), d = mk3(rd.x, rd.y, rd.z);
    const HitRec h = load_hit(hits, k);
    const float4 st = wb.st[j].thr;
    float4 li4 = wb.st[j].li;
    f3 thr = mk3(st.x, st.y, st.z), Li = mk3(li4.x, li4.y, li4.z);
    uint32_t rng = f2u(st.w);
    int seg = (int)(f2u(li4.w) & 0xFFu) + 1;
    uint32_t pflags = f2u(li4.w) & 0xFFFF0000u & ~PATH_RAY_IN_FLIGHT;   // the ray arrived
    settle_pending_shadow(wb, j, pflags, Li);                            // direct light of the previous vertex first (:286,291)
    if (wb.pathHash || depth < fc.maxDepth) {
        const Surface s = eval_surface(sc, o, d, h);
        if (wb.pathHash) wb.pathHash[j] = fnv_fold(fnv_fold(wb.pathHash[j], (uint32_t)s.instId), (uint32_t)s.primId);
        if (depth < fc.maxDepth) {
            PathVertex v;
            v.pos = o + d * h.t;            // :665
            v.nrm = normalize(s.normal);   // :666
            v.alb = s.albedo; v.shade = s.shade; v.ior = s.ior;
            v.I = d;                        // "I = ray.dir" :243,274,316
            VertexOut vo; vo.pushNext = 0; vo.pushShadow = 0;
            bool alive = shade_vertex<REUSE, FAST>(fc, wb, v, depth, j, thr, rng, pflags, vo);
            if (defer) *defer = vo; else push_vertex_out(vo, j, nextQ, nextCount, shQ, shCount);
            wb.st[j].thr = make_float4(thr.x, thr.y, thr.z, u2f(rng));
            wb.st[j].li = make_float4(Li.x, Li.y, Li.z, u2f(pflags | pack_aov(seg, alive ? RT_TERM_MAXDEPTH : RT_TERM_ROULETTE)));
            return;
        }
    }
    // depth == maxDepth: the depth loop has run out (:233); nothing more is added
    wb.st[j].li = make_float4(Li.x, Li.y, Li.z, u2f(pflags | pack_aov(seg, RT_TERM_MAXDEPTH)));
}

// ------------------------------------------------------------------------------------------------ accumulate
// "Lframe += SafeColor(Li)" over the samples of this batch in sample order (RTRay.cs:217,320); on the last
// batch: Lout = Lframe / max(1,spp), float4 radiance, progressive accumulator and PackRGBA8 (:323-324, :66-76).
RT_HD void accumulate(const FrameConst& fc, const WaveBuffers& wb, int sampleBase, int nSamples, bool last, int i) {
    const int pix = fc.pixelMap[i];
    f3 L = mk3(0.0f, 0.0f, 0.0f);
    if (sampleBase > 0) { const float4 l4 = wb.lframe[i]; L = mk3(l4.x, l4.y, l4.z); }
    f3 skyMiss = mk3(0.0f, 0.0f, 0.0f);
    const bool primaryMiss = (f2u(wb.gbPosHit[i].w) & GB_HIT) == 0u;
    if (primaryMiss) { int x, y; pixel_xy(fc, pix, &x, &y); skyMiss = safe_color(sky_weighted(fc.env, primary_dir(fc, x, y))); }   // :216-217
    const size_t plane = (size_t)fc.width * (size_t)fc.height;
    int resOwner = -1;   // the reference's samples run in order and each overwrites resCur[index] (:294): the last writer's reservoir stays
    for (int s = 0; s < nSamples; s++) {
        const int j = s * fc.npx + i;
        if (primaryMiss) {   // no path was started (shade_first): "Lframe += SafeColor(sky)" per sample (:214-219), nothing to read
            L = L + skyMiss;
            if (wb.segCountOut) {
                const size_t oi = (size_t)(sampleBase + s) * plane + (size_t)pix;
                wb.segCountOut[oi] = 0; wb.termCodeOut[oi] = (uint8_t)RT_TERM_PRIMARY_MISS; wb.pathHashOut[oi] = 0x811C9DC5u;
            }
            continue;
        }
        float4 li4 = wb.st[j].li;
        if (f2u(li4.w) & PATH_WROTE_RESERVOIR) resOwner = j;
        if ((f2u(li4.w) & (PATH_PENDING_SHADOW | PATH_RAY_IN_FLIGHT)) != 0u) {
            // what the wavefront left pending for this path, in the reference's order: the direct light of its last Lambert vertex
            // (RTRay.cs:286,291), then the sky seen by a bounce ray that left the scene: "Li += throughput * SkyWeighted(ray.dir)" (:242,273,315)
            uint32_t pf = f2u(li4.w);
            f3 Li = mk3(li4.x, li4.y, li4.z);
            settle_pending_shadow(wb, j, pf, Li);
            if (pf & PATH_RAY_IN_FLIGHT) {
                const float4 st = wb.st[j].thr, md = wb.st[j].miss;
                Li = Li + mk3(st.x, st.y, st.z) * sky_weighted(fc.env, mk3(md.x, md.y, md.z));
                if (wb.pathHash) wb.pathHash[j] = fnv_fold(wb.pathHash[j], 0xFFFFFFFFu);
                pf = (pf & 0xFFFF0000u & ~PATH_RAY_IN_FLIGHT) | pack_aov((int)(pf & 0xFFu) + 1, RT_TERM_MISS);
            }
            li4 = make_float4(Li.x, Li.y, Li.z, u2f(pf));
        }
        L = L + safe_color(mk3(li4.x, li4.y, li4.z));
        if (wb.segCountOut) {
            const uint32_t a = f2u(li4.w);
            const size_t oi = (size_t)(sampleBase + s) * plane + (size_t)pix;
            wb.segCountOut[oi] = (uint8_t)(a & 0xFFu);
            wb.termCodeOut[oi] = (uint8_t)((a >> 8) & 0xFFu);
            wb.pathHashOut[oi] = wb.pathHash[j];
        }
    }
    if (wb.resCur0 && resOwner >= 0) { wb.resCur0[pix] = wb.resPath0[resOwner]; wb.resCur1[pix] = wb.resPath1[resOwner]; wb.resCur2[pix] = wb.resPath2[resOwner]; }
    if (!last) { wb.lframe[i] = make_float4(L.x, L.y, L.z, 0.0f); return; }
    const f3 Lout = L * (1.0f / (float)max(1, fc.spp));   // :323
    wb.radiance[pix] = make_float4(Lout.x, Lout.y, Lout.z, 1.0f);
    f3 shown = Lout;
    float nAccum = 1.0f;
    if (fc.flags & RT_FLAG_ACCUMULATE) {
        float4 a = (fc.flags & RT_FLAG_RESET_ACCUM) ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : wb.accum[pix];
        a.x = a.x + Lout.x; a.y = a.y + Lout.y; a.z = a.z + Lout.z; a.w = a.w + 1.0f;
        wb.accum[pix] = a;
        const float inv = 1.0f / a.w;
        shown = mk3(a.x * inv, a.y * inv, a.z * inv);
        nAccum = a.w;
    }
    // multi-GPU gather payload: what this pixel shows (the progressive mean when accumulating), so that rank 0's
    // de-interleaved image packs to the same RGBA8 as a single-GPU run
    wb.tileRadiance[i] = make_float4(shown.x, shown.y, shown.z, nAccum);
    const int packed = pack_rgba8(shown);
    wb.rgba8[pix] = packed;
    if (wb.tileRgba) wb.tileRgba[i] = packed;
}

}   // namespace rtx
