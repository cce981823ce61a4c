// hostsim.cpp — TEST HARNESS ONLY (never a product path, never loaded by ilgpu_raytracing_b200).
//
// Compiles the renderer core's __host__ __device__ stage bodies (rt_core.h / rt_traverse.h /
// rt_wavefront.h) and the host BVH builder for the CPU and runs the wavefront as plain loops, so
// that `pytest -m "not gpu"` can check the wide-BVH build, the traversal, the tie-break rule and the
// wavefront state machine against the oracle in a container without a GPU.  The warp-level
// machinery (persistent fetch, ballot/shuffle refill, atomics, shared-memory stacks) exists only in
// rt_kernels.cu and is covered by the -m gpu tests.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../ilgpu_raytracing_b200/csrc/rt_bvh.h"
#include "../../ilgpu_raytracing_b200/csrc/rt_tiles.h"
#include "../../ilgpu_raytracing_b200/csrc/rt_wavefront.h"

using namespace rtx;

#define HS_API extern "C" __attribute__((visibility("default")))

struct HsScene {
    HostBvh bvh;
    std::vector<RtInstanceRecord> instances; std::vector<RtSphere> spheres; std::vector<RtFloat2> texcoords; std::vector<RtMeshTriUV> triUVs;
    std::vector<int32_t> triMat; std::vector<RtMaterialRecord> materials; std::vector<RtRGBA32> texels; std::vector<RtTexInfo> texInfos;
    DeviceScene ds;
    std::string err;
};

template <class T> static void copy_or_one(std::vector<T>& dst, const T* src, int64_t n) {   // AllocateOrEmpty, Scene.cs:370-377
    if (src && n > 0) dst.assign(src, src + n); else { dst.assign(1, T()); memset(dst.data(), 0, sizeof(T)); }
}

HS_API HsScene* hs_scene_create_ex(const RtSceneDesc* d, int maxDepth);
HS_API HsScene* hs_scene_create(const RtSceneDesc* d) { return hs_scene_create_ex(d, 0); }
// maxDepth: the depth limit handed to the builder (0 = the traversal stack's); a small value forces the depth-bounded rebuild
HS_API HsScene* hs_scene_create_ex(const RtSceneDesc* d, int maxDepth) {
    HsScene* s = new HsScene();
    if (!build_wide_bvh(*d, s->bvh, s->err, nullptr, maxDepth)) return s;
    copy_or_one(s->instances, d->instances, d->nInstances); copy_or_one(s->spheres, d->spheres, d->nSpheres);
    copy_or_one(s->texcoords, d->meshTexcoords, d->nMeshTexcoords); copy_or_one(s->triUVs, d->meshTriUVs, d->nMeshTriUVs);
    copy_or_one(s->triMat, d->triMatIndex, d->nTriMatIndex); copy_or_one(s->materials, d->materials, d->nMaterials);
    copy_or_one(s->texels, d->texels, d->nTexels); copy_or_one(s->texInfos, d->texInfos, d->nTexInfos);
    DeviceScene& ds = s->ds;
    ds.nodes = s->bvh.nodes.data(); ds.nNodes = (int)s->bvh.nodes.size(); ds.prims = s->bvh.prims.data(); ds.nPrims = (int)s->bvh.prims.size();
    ds.instances = s->instances.data(); ds.nInstances = (int)s->instances.size(); ds.spheres = s->spheres.data(); ds.nSpheres = (int)s->spheres.size();
    ds.texcoords = s->texcoords.data(); ds.triUVs = s->triUVs.data(); ds.triMatIndex = s->triMat.data();
    ds.materials = s->materials.data(); ds.nMaterials = (int)s->materials.size(); ds.texels = s->texels.data();
    ds.texInfos = s->texInfos.data(); ds.nTexInfos = (int)s->texInfos.size(); ds.triMaterials = 0; ds.tFarScale = s->bvh.stats.maxInstanceScale;
    return s;
}
HS_API const char* hs_scene_error(HsScene* s) { return s->err.c_str(); }
HS_API void hs_scene_destroy(HsScene* s) { delete s; }
// FNV-1a over the wide nodes and primitive records: builder regression checks compare it across builder versions
HS_API uint64_t hs_scene_hash(HsScene* s) {
    uint64_t h = 1469598103934665603ull;
    auto eat = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
    eat(s->bvh.nodes.data(), s->bvh.nodes.size() * sizeof(WideNode)); eat(s->bvh.prims.data(), s->bvh.prims.size() * sizeof(PrimRec));
    return h;
}
HS_API void hs_scene_stats(HsScene* s, int64_t* out6) {
    out6[0] = s->bvh.stats.nPrims; out6[1] = s->bvh.stats.nTris; out6[2] = s->bvh.stats.nSpheres; out6[3] = s->bvh.stats.nWideNodes; out6[4] = s->bvh.stats.maxDepth; out6[5] = s->bvh.stats.depthBounded;
}
// 0 = node step + all of its primitives; n > 0 = k_extend's lane schedule (n node steps, one primitive step) with queued primitive groups
HS_API void hs_set_lane_schedule(int nodeSteps) { host_lane_schedule() = nodeSteps < 0 ? 0 : nodeSteps; }
HS_API void hs_set_plane_pad(float quanta) { host_plane_pad() = quanta < 0.0f ? 0.0f : quanta; }   // analysis knob, see rt_traverse.h
// ---- ray capture + SIMT schedule simulator (analysis tool: tests/cpu_simt_schedule.py) ----------------------------------------
struct CapRay { f3 o, d; int any; int wave; };
static std::vector<CapRay> g_cap; static bool g_capOn = false; static int g_capWave = 0;
static inline void cap_ray(f3 o, f3 d, int any) { if (g_capOn) g_cap.push_back({o, d, any, g_capWave}); }
HS_API void hs_capture(int on) { g_capOn = on != 0; if (on) { g_cap.clear(); g_capWave = 0; } }
HS_API long long hs_capture_count() { return (long long)g_cap.size(); }
// Runs the captured waves through k_extend's warp loop on 32-lane warps (rays handed out in batches of 96 from a shared cursor,
// `warps` warps taking turns one iteration at a time) and charges every executed phase its SASS length.
//   policy 0: the shipped loop - refill, `nodeSteps` node steps, one primitive step when >= `primVote` lanes hold one (or no lane can step)
//   policy 1: one phase per iteration, whichever has more ready lanes (node lanes weighted by cN / cP)
//   policy 2: two-stage primitive tests - after the node steps every lane holding a candidate runs a cheap conservative test on ONE
//             candidate (cost cCheap = -preCull, modelled as exact: it passes exactly the candidates the exact test accepts); a
//             survivor waits in its lane (which keeps taking node steps) until >= primVote lanes hold one, or nothing else can run,
//             and only then the exact test (cPrim) is issued
// out: [0] rays, [1] iterations, [2] warp instructions, [3] node phases, [4] node lanes, [5] prim phases, [6] prim lanes, [7] node steps, [8] prim steps
//   preCull: probability that a candidate whose exact test would FAIL is removed for free before the primitive phase (what a cheap
//   conservative pre-test in the node step would do); out[9] = candidates removed that way, out[10] = exact tests that accepted a hit
HS_API void hs_simulate(HsScene* s, int anyHit, int policy, int nodeSteps, int primVote, int warps, double cRefill, double cNode, double cPrim, double cLoop, double preCull, double* out) {
    for (int i = 0; i < 11; i++) out[i] = 0.0;
    uint32_t lcg = 12345u;
    std::vector<size_t> waveStart;
    for (size_t i = 0; i < g_cap.size(); i++) if (i == 0 || g_cap[i].wave != g_cap[i - 1].wave) waveStart.push_back(i);
    waveStart.push_back(g_cap.size());
    struct Lane { Traversal<false, true> c; Traversal<true, true> a; LaneStack st; bool active = false; bool survivor = false; };
    struct Warp { std::vector<Lane> lanes; size_t poolNext = 0, poolEnd = 0; bool exhausted = false; };
    TraceCounters tc = {0, 0, 0};
    for (size_t w = 0; w + 1 < waveStart.size(); w++) {
        const size_t b = waveStart[w], e = waveStart[w + 1];
        if (e == b || (g_cap[b].any != 0) != (anyHit != 0)) continue;
        size_t cursor = b;
        std::vector<Warp> ws((size_t)std::max(1, warps));
        for (auto& wp : ws) wp.lanes.resize(32);
        size_t live = ws.size();
        out[0] += (double)(e - b);
        while (live > 0) {
            live = 0;
            for (auto& wp : ws) {
                // refill
                int idle = 0; for (auto& l : wp.lanes) idle += l.active ? 0 : 1;
                bool refilled = false;
                while (idle > 0 && !wp.exhausted) {
                    if (wp.poolNext >= wp.poolEnd) { if (cursor >= e) { wp.exhausted = true; break; } wp.poolNext = cursor; wp.poolEnd = std::min(e, cursor + 96); cursor = wp.poolEnd; }
                    for (auto& l : wp.lanes) {
                        if (l.active || wp.poolNext >= wp.poolEnd) continue;
                        const CapRay& r = g_cap[wp.poolNext++];
                        if (anyHit) l.a.init(r.o, r.d, box_idir(r.d), 1e29f, l.st); else l.c.init(r.o, r.d, box_idir(r.d), 1e30f, l.st);
                        l.active = true; l.survivor = false; idle--; refilled = true;
                    }
                }
                int nActive = 0; for (auto& l : wp.lanes) nActive += l.active ? 1 : 0;
                if (nActive == 0) continue;
                live++;
                out[1] += 1.0; out[2] += cLoop + (refilled ? cRefill : 0.0);
                auto canN = [&](Lane& l) { return l.active && (anyHit ? (!l.a.done && l.a.can_node_step(l.st)) : (!l.c.done && l.c.can_node_step(l.st))); };
                auto hasP = [&](Lane& l) { return l.active && (anyHit ? (!l.a.done && l.a.has_prims()) : (!l.c.done && l.c.has_prims())); };
                auto doN = [&]() { int n = 0; for (auto& l : wp.lanes) if (canN(l)) { if (anyHit) l.a.node_step(s->ds, l.st, &tc); else l.c.node_step(s->ds, l.st, &tc); n++; }
                                   if (n) { out[2] += cNode; out[3] += 1.0; out[4] += n; out[7] += n; } return n; };
                auto accepts = [&](Lane& l) {   // would the next exact test of this lane accept a hit?  (run on a copy)
                    Lane c = l; const int before = anyHit ? (c.a.occluded ? 1 : 0) : c.c.best.prim;
                    if (anyHit) c.a.prim_step(s->ds, c.st, &tc); else c.c.prim_step(s->ds, c.st, &tc);
                    return (anyHit ? (c.a.occluded ? 1 : 0) : c.c.best.prim) != before; };
                auto ownBoxMiss = [&](Lane& l) {   // preCull < 0: the candidate's OWN box (plain triangles only) against the ray, culled at the current closest t
                    const uint32_t tg = anyHit ? l.a.tgroup.y : l.c.tgroup.y, tgx = anyHit ? l.a.tgroup.x : l.c.tgroup.x, tv = anyHit ? l.a.tvalid : l.c.tvalid;
                    const int bit = rt_bfind(tg);
                    const PrimRec& pr = s->ds.prims[(int)tgx + rt_popc(tv & ~(0xFFFFFFFFu << bit))];
                    if (f2u(pr.q2.w) & (PRIM_SPHERE | PRIM_XFORM)) return false;
                    const f3 o = anyHit ? l.a.o : l.c.o, id = anyHit ? l.a.idir : l.c.idir;
                    const float tmax = anyHit ? l.a.tMax : l.c.best.t;
                    const float lo[3] = {fminf(pr.q0.x, fminf(pr.q1.x, pr.q2.x)), fminf(pr.q0.y, fminf(pr.q1.y, pr.q2.y)), fminf(pr.q0.z, fminf(pr.q1.z, pr.q2.z))};
                    const float hi[3] = {fmaxf(pr.q0.x, fmaxf(pr.q1.x, pr.q2.x)), fmaxf(pr.q0.y, fmaxf(pr.q1.y, pr.q2.y)), fmaxf(pr.q0.z, fmaxf(pr.q1.z, pr.q2.z))};
                    const float oo[3] = {o.x, o.y, o.z}, ii[3] = {id.x, id.y, id.z};
                    float tn = 0.001f, tf = tmax;
                    for (int a = 0; a < 3; a++) { const float pad = 1e-4f * (fabsf(lo[a]) + fabsf(hi[a]) + 1.0f); const float t1 = (lo[a] - pad - oo[a]) * ii[a], t2 = (hi[a] + pad - oo[a]) * ii[a];
                                                  tn = fmaxf(tn, fminf(t1, t2)); tf = fminf(tf, fmaxf(t1, t2)); }
                    return tn > tf * 1.0000007f; };
                auto cull = [&]() { if (preCull == 0.0 || policy == 2) return; for (auto& l : wp.lanes) while (hasP(l)) {
                                        if (preCull < 0.0) { if (!ownBoxMiss(l)) break; if (anyHit) l.a.prim_step(s->ds, l.st, &tc); else l.c.prim_step(s->ds, l.st, &tc); out[9] += 1.0; continue; }
                                        if (accepts(l)) break;
                                        lcg = lcg * 1664525u + 1013904223u;
                                        if ((double)(lcg >> 8) / 16777216.0 >= preCull) break;
                                        if (anyHit) l.a.prim_step(s->ds, l.st, &tc); else l.c.prim_step(s->ds, l.st, &tc); out[9] += 1.0; } };
                auto doP = [&]() { cull(); int n = 0; for (auto& l : wp.lanes) if (hasP(l)) { if (accepts(l)) out[10] += 1.0; if (anyHit) l.a.prim_step(s->ds, l.st, &tc); else l.c.prim_step(s->ds, l.st, &tc); n++; }
                                   if (n) { out[2] += cPrim; out[5] += 1.0; out[6] += n; out[8] += n; } return n; };
                if (policy == 0) {
                    for (int k = 0; k < nodeSteps; k++) { doN(); cull(); }
                    int pm = 0, nm = 0; for (auto& l : wp.lanes) { pm += hasP(l) ? 1 : 0; nm += canN(l) ? 1 : 0; }
                    if (pm > 0 && (pm >= primVote || nm == 0)) doP();
                } else if (policy == 1) {
                    int pm = 0, nm = 0; for (auto& l : wp.lanes) { pm += hasP(l) ? 1 : 0; nm += canN(l) ? 1 : 0; }
                    if ((double)nm * cPrim >= (double)pm * cNode && nm > 0) doN(); else if (pm > 0) doP(); else doN();
                } else {
                    const double cCheap = -preCull;
                    for (int k = 0; k < nodeSteps; k++) doN();
                    int nc = 0;   // cheap stage: one candidate per lane that holds one and has no survivor waiting
                    for (auto& l : wp.lanes) if (hasP(l) && !l.survivor) {
                        nc++;
                        if (accepts(l)) l.survivor = true;
                        else { if (anyHit) l.a.prim_step(s->ds, l.st, &tc); else l.c.prim_step(s->ds, l.st, &tc); out[9] += 1.0; }
                    }
                    if (nc) out[2] += cCheap;
                    int sv = 0, other = 0;
                    for (auto& l : wp.lanes) { sv += (l.active && l.survivor) ? 1 : 0; other += (canN(l) || (hasP(l) && !l.survivor)) ? 1 : 0; }
                    if (sv > 0 && (sv >= primVote || other == 0)) {
                        for (auto& l : wp.lanes) if (l.active && l.survivor) { if (anyHit) l.a.prim_step(s->ds, l.st, &tc); else l.c.prim_step(s->ds, l.st, &tc); l.survivor = false; out[10] += 1.0; }
                        out[2] += cPrim; out[5] += 1.0; out[6] += sv; out[8] += sv;
                    }
                }
                for (auto& l : wp.lanes) if (l.active && (anyHit ? l.a.done : l.c.done)) l.active = false;
            }
        }
    }
}

// one ray; returns hit flag; out = {t, primId, instId, bu, bv}, counters = {nodes, tris, spheres}
HS_API int hs_trace(HsScene* s, const float* o, const float* d, int anyHit, float tMax, unsigned flags, float* out5, uint32_t* counters3) {
    s->ds.triMaterials = (flags & RT_FLAG_TRI_MATERIALS) ? 1 : 0;
    LaneStack st; TraceCounters c = {0, 0, 0}; HitRec h; h.t = 1e30f; h.prim = -1; h.bu = h.bv = 0;
    bool r;
    f3 ro = mk3(o[0], o[1], o[2]), rd = mk3(d[0], d[1], d[2]);
    if (anyHit) r = trace_wide<true, true>(s->ds, ro, rd, tMax, st, &h, &c);
    else r = trace_wide<false, true>(s->ds, ro, rd, tMax, st, &h, &c);
    out5[0] = h.t; out5[1] = -1; out5[2] = -1; out5[3] = h.bu; out5[4] = h.bv;
    if (!anyHit && r) { Surface sf = eval_surface(s->ds, ro, rd, h); out5[1] = (float)sf.primId; out5[2] = (float)sf.instId; }
    if (counters3) { counters3[0] = c.nodes; counters3[1] = c.tris; counters3[2] = c.spheres; }
    return r ? 1 : 0;
}

struct HsOutputs {
    int32_t* rgba8; float* depth; int32_t* objId; float* radiance4; float* accum4;
    int32_t* primId; int32_t* instId; float* primaryT;           // per GLOBAL pixel (scattered through the pixel map)
    float* gbPos; float* gbNrm; float* gbAlb; int32_t* gbMat;   // per global pixel, 3 floats each
    uint8_t* segCount; uint8_t* termCode; uint32_t* pathHash;   // [spp][W*H]
    uint64_t counters[8];                                        // raysPrimary, raysBounce, raysShadow, nodes, tris, spheres
};

// The whole frame as plain loops; mirrors the launch sequence of rt_render() in rt_api.cu.
// resPrev / resCur: the reference's Reservoir records per global pixel (ReSTIR reuse frames only; both null otherwise)
HS_API int hs_render_reuse(HsScene* s, const RtCamera* cam, const RtCamera* prevCam, const RtRenderConfig* cfg, HsOutputs* out, const RtReservoir* resPrev, RtReservoir* resCur) {
    const bool reuse = cfg->enableTemporalReuse != 0 || cfg->enableSpatialReuse != 0;
    if (reuse && (!resPrev || !resCur || cfg->worldSize > 1)) return RT_ERR_INVALID_ARGUMENT;
    const int W = cfg->width, H = cfg->height;
    std::vector<int> pmap; build_pixel_map(W, H, cfg->tileSize, cfg->rank, cfg->worldSize, pmap);
    const int npx = (int)pmap.size();
    const int spp = cfg->spp > 1 ? cfg->spp : 1;
    int S = cfg->samplesPerPass > 0 ? cfg->samplesPerPass : 2;
    if (S > spp) S = spp;
    const size_t P = (size_t)npx * S;
    s->ds.triMaterials = (cfg->flags & RT_FLAG_TRI_MATERIALS) ? 1 : 0;

    FrameConst fc; memset(&fc, 0, sizeof(fc));
    fc.width = W; fc.height = H; fc.frame = cfg->frame; fc.spp = cfg->spp; fc.maxDepth = cfg->maxDepth; fc.rngLockNoise = cfg->rngLockNoise; fc.flags = cfg->flags;
    fc.camOrigin = mk3(cam->origin); fc.camLowerLeft = mk3(cam->lowerLeft); fc.camHorizontal = mk3(cam->horizontal); fc.camVertical = mk3(cam->vertical);
    fc.env.dirLightDir = mk3(cfg->dirLightDir); fc.env.dirLightRadiance = mk3(cfg->dirLightRadiance); fc.env.skyTop = mk3(cfg->skyTintTop); fc.env.skyBottom = mk3(cfg->skyTintBottom);
    fc.npx = npx; fc.pixelMap = pmap.data();
    const size_t G = (size_t)W * H;
    std::vector<int> invMap; std::vector<float4> rp0, rp1, rp2, rc0, rc1, rc2, rq0, rq1, rq2;
    if (reuse) {
        const RtCamera* pc = prevCam ? prevCam : cam;
        fc.enableTemporal = cfg->enableTemporalReuse; fc.enableSpatial = cfg->enableSpatialReuse;
        fc.prevOrigin = mk3(pc->origin); fc.prevRight = mk3(pc->right); fc.prevUp = mk3(pc->up); fc.prevForward = mk3(pc->forward); fc.prevFovY = pc->fovYRadians; fc.prevAspect = pc->aspect;
        invMap.assign(G, 0); for (int i = 0; i < npx; i++) invMap[(size_t)pmap[i]] = i;
        fc.invPixelMap = invMap.data(); fc.lookBase = 0;   // look* pointers: set below, once the G-buffer vectors exist
        rp0.resize(G); rp1.resize(G); rp2.resize(G); rc0.resize(G); rc1.resize(G); rc2.resize(G);
        auto split = [](const RtReservoir& r, float4& a, float4& b, float4& c) {
            a = make_float4(r.L.X, r.L.Y, r.L.Z, r.pdf); b = make_float4(r.wi.X, r.wi.Y, r.wi.Z, r.w); c = make_float4(r.wSum, u2f((uint32_t)r.m), u2f((uint32_t)r.lightId), 0.0f);
        };
        for (size_t p = 0; p < G; p++) { split(resPrev[p], rp0[p], rp1[p], rp2[p]); split(resCur[p], rc0[p], rc1[p], rc2[p]); }
        fc.resPrev0 = rp0.data(); fc.resPrev1 = rp1.data(); fc.resPrev2 = rp2.data();
        rq0.resize(P); rq1.resize(P); rq2.resize(P);
    }

    std::vector<float4> gbPosHit(npx), gbNrmMat(npx), gbAlbObj(npx), lframe(npx), tileRad(npx);
    std::vector<int> primId(npx), instId(npx); std::vector<float> primaryT(npx);
    std::vector<uint32_t> pathHash(P);
    std::vector<float4> radiance((size_t)W * H), accum((size_t)W * H);
    std::vector<float4> qo[2], qd[2], so(P), sd(P), hitTuv(P);
    std::vector<PathState> pathState(P);
    for (int b = 0; b < 2; b++) { qo[b].resize(P); qd[b].resize(P); }
    std::vector<int> hitPrim(P);
    const HitQueue hq = {hitPrim.data(), hitTuv.data()};
    std::vector<int32_t> dummyI((size_t)W * H); std::vector<float> dummyF((size_t)W * H);

    WaveBuffers wb; memset(&wb, 0, sizeof(wb));
    wb.gbPosHit = gbPosHit.data(); wb.gbNrmMat = gbNrmMat.data(); wb.gbAlbObj = gbAlbObj.data();
    wb.primId = primId.data(); wb.instId = instId.data(); wb.primaryT = primaryT.data(); wb.lframe = lframe.data(); wb.tileRadiance = tileRad.data();
    wb.rgba8 = out->rgba8 ? out->rgba8 : dummyI.data(); wb.depth = out->depth ? out->depth : dummyF.data(); wb.objId = out->objId ? out->objId : dummyI.data();
    wb.radiance = radiance.data(); wb.accum = out->accum4 ? (float4*)out->accum4 : accum.data();
    wb.st = pathState.data();
    fc.lookPosHit = gbPosHit.data(); fc.lookNrmMat = gbNrmMat.data(); fc.lookAlbObj = gbAlbObj.data();
    const bool aov = (cfg->flags & RT_FLAG_PATH_AOVS) && out->segCount && out->termCode && out->pathHash;
    if (reuse) { wb.resPath0 = rq0.data(); wb.resPath1 = rq1.data(); wb.resPath2 = rq2.data(); wb.resCur0 = rc0.data(); wb.resCur1 = rc1.data(); wb.resCur2 = rc2.data(); }
    if (aov) { wb.pathHash = pathHash.data(); wb.segCountOut = out->segCount; wb.termCodeOut = out->termCode; wb.pathHashOut = out->pathHash; }

    TraceCounters tc = {0, 0, 0}; uint64_t nodes = 0, tris = 0, sph = 0, raysB = 0, raysS = 0;
    auto flushCnt = [&]() { nodes += tc.nodes; tris += tc.tris; sph += tc.spheres; tc.nodes = tc.tris = tc.spheres = 0; };
    LaneStack st;

    // primary visibility
    RayQueue q0 = {qo[0].data(), qd[0].data()};
    for (int i = 0; i < npx; i++) generate_primary(fc, q0, i);
    for (int i = 0; i < npx; i++) { f3 o = mk3(q0.o[i].x, q0.o[i].y, q0.o[i].z), d = mk3(q0.d[i].x, q0.d[i].y, q0.d[i].z); cap_ray(o, d, 0); HitRec h; trace_wide<false, true>(s->ds, o, d, 1e30f, st, &h, &tc); flushCnt(); store_closest_result(hq, nullptr, i, i, h, d); }
    g_capWave++;
    for (int i = 0; i < npx; i++) primary_finish(fc, s->ds, wb, q0, hq, i);

    // integrator, one batch of S samples at a time
    ShadowQueue shq = {so.data(), sd.data()};
    for (int s0 = 0; s0 < spp; s0 += S) {
        const int ns = (s0 + S <= spp) ? S : (spp - s0);
        int cur = 0; int nNext = 0, nSh = 0;
        RayQueue nq = {qo[cur].data(), qd[cur].data()};
        for (int j = 0; j < npx * ns; j++) { if (reuse) shade_first<true>(fc, wb, s0, j, nq, &nNext, shq, &nSh); else shade_first<false>(fc, wb, s0, j, nq, &nNext, shq, &nSh); }
        for (int depth = 1; depth <= fc.maxDepth; depth++) {
            for (int k = 0; k < nSh; k++) {
                f3 o = mk3(shq.o[k].x, shq.o[k].y, shq.o[k].z), d = mk3(shq.d[k].x, shq.d[k].y, shq.d[k].z);
                cap_ray(o, d, 1);
                bool occ = trace_wide<true, true>(s->ds, o, d, 1e29f, st, nullptr, &tc); flushCnt();
                store_anyhit_result(wb.st, (int)f2u(shq.o[k].w), occ);   // settled at the path's next touch (shade_next / accumulate)
            }
            raysS += (uint64_t)nSh; g_capWave++;
            RayQueue cq = {qo[cur].data(), qd[cur].data()};
            for (int k = 0; k < nNext; k++) { f3 o = mk3(cq.o[k].x, cq.o[k].y, cq.o[k].z), d = mk3(cq.d[k].x, cq.d[k].y, cq.d[k].z); cap_ray(o, d, 0); HitRec h; trace_wide<false, true>(s->ds, o, d, 1e30f, st, &h, &tc); flushCnt(); store_closest_result(hq, wb.st, k, (int)f2u(cq.o[k].w), h, d); }
            g_capWave++;
            raysB += (uint64_t)nNext;
            const int nCur = nNext; nNext = 0; nSh = 0;
            RayQueue nq2 = {qo[cur ^ 1].data(), qd[cur ^ 1].data()};
            for (int k = 0; k < nCur; k++) {   // only rays that hit are shaded; the others are settled by accumulate
                if (hitPrim[k] < 0) continue;
                if (reuse) shade_next<true>(fc, s->ds, wb, depth, cq, hq, k, nq2, &nNext, shq, &nSh); else shade_next<false>(fc, s->ds, wb, depth, cq, hq, k, nq2, &nNext, shq, &nSh);
            }
            cur ^= 1;
        }
        const bool last = (s0 + ns >= spp);
        for (int i = 0; i < npx; i++) accumulate(fc, wb, s0, ns, last, i);
    }

    for (int i = 0; i < npx; i++) {
        const int pix = pmap[i];
        if (out->primId) out->primId[pix] = primId[i];
        if (out->instId) out->instId[pix] = instId[i];
        if (out->primaryT) out->primaryT[pix] = primaryT[i];
        if (out->gbPos) { out->gbPos[3 * pix] = gbPosHit[i].x; out->gbPos[3 * pix + 1] = gbPosHit[i].y; out->gbPos[3 * pix + 2] = gbPosHit[i].z; }
        if (out->gbNrm) { out->gbNrm[3 * pix] = gbNrmMat[i].x; out->gbNrm[3 * pix + 1] = gbNrmMat[i].y; out->gbNrm[3 * pix + 2] = gbNrmMat[i].z; }
        if (out->gbAlb) { out->gbAlb[3 * pix] = gbAlbObj[i].x; out->gbAlb[3 * pix + 1] = gbAlbObj[i].y; out->gbAlb[3 * pix + 2] = gbAlbObj[i].z; }
        if (out->gbMat) out->gbMat[pix] = (int32_t)f2u(gbNrmMat[i].w);
        if (out->radiance4) memcpy(out->radiance4 + 4 * (size_t)pix, &radiance[pix], 16);
    }
    if (reuse) for (size_t p = 0; p < G; p++) {
        RtReservoir& r = resCur[p];
        r.L.X = rc0[p].x; r.L.Y = rc0[p].y; r.L.Z = rc0[p].z; r.pdf = rc0[p].w; r.wi.X = rc1[p].x; r.wi.Y = rc1[p].y; r.wi.Z = rc1[p].z; r.w = rc1[p].w;
        r.wSum = rc2[p].x; r.m = (int32_t)f2u(rc2[p].y); r.lightId = (int32_t)f2u(rc2[p].z);
    }
    out->counters[0] = (uint64_t)npx; out->counters[1] = raysB; out->counters[2] = raysS; out->counters[3] = nodes; out->counters[4] = tris; out->counters[5] = sph;
    return 0;
}

HS_API int hs_render(HsScene* s, const RtCamera* cam, const RtRenderConfig* cfg, HsOutputs* out) {
    if (cfg->enableTemporalReuse || cfg->enableSpatialReuse) return RT_ERR_INVALID_ARGUMENT;   // use hs_render_reuse
    return hs_render_reuse(s, cam, nullptr, cfg, out, nullptr, nullptr);
}

// ---- present chain bodies (rt_core.h) as plain loops ----
HS_API void hs_bilinear_upsample(const int* src, int srcW, int srcH, int* dst, int dstW, int dstH) {
    for (int i = 0; i < dstW * dstH; i++) dst[i] = bilinear_upsample_pixel(src, srcW, srcH, dstW, dstH, i);
}
HS_API void hs_taa_resolve(int* out, const int* lowColor, const int* lowObj, int inW, int inH, int outW, int outH, int* histColor, int* histObj,
                           int isFirstFrame, float feedback, float sharpness, float clampK) {
    float lut[256];
    for (int v = 0; v < 256; v++) lut[v] = srgb_to_linear_u8(v);
    TaaConst p; p.outW = outW; p.outH = outH; p.inW = inW; p.inH = inH; p.feedback = feedback; p.sharpness = sharpness; p.clampK = clampK; p.isFirstFrame = isFirstFrame;
    for (int i = 0; i < outW * outH; i++) { int obj; const int c = taa_resolve_pixel(p, lut, lowColor, lowObj, histColor[i], histObj[i], i, &obj); out[i] = c; histColor[i] = c; histObj[i] = obj; }
}
HS_API float hs_pow(float x, float y) { return pow_p(x, y); }

