// rt_traverse.h — closest-hit / any-hit traversal of the compressed 8-wide BVH.
//
// Replaces the reference's two-level stackless BVH2 walks: TraceClosest / TraverseBLAS_* and
// ShadowOcclusion / AnyHit_* (SceneDeviceViews.cs:30-327).  The BOX tests here are ours (quantised
// child boxes, fused multiply-adds, octant-ordered visiting) and only have to be conservative; the
// PRIMITIVE tests and the hit-acceptance rules are the reference's, bit for bit (rt_core.h).
//
// Ordering rule that makes the result independent of visiting order: the reference keeps the FIRST
// hit it meets among equal t ("t < tClosest", :142,:199; "tWorld < closestT", :68).  Every primitive
// record carries its rank in the reference's own visiting order, and we keep the lexicographic
// minimum of (tWorld, [same instance: tObj], rank).
#pragma once
#include "rt_core.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace rtx {

struct TraceCounters { uint32_t nodes, tris, spheres; };

// Per-lane traversal stack: RT_STACK_ENTRIES (node group, hit mask) pairs.  Device: shared memory, strided by the
// block size so the lanes of a warp hit distinct banks; no local-memory spill path, so push / pop are branch-free
// (rt_bvh.cpp refuses trees deeper than the stack).  Host (tests/hostsim): plain array.
#ifndef RT_STACK_ENTRIES
#define RT_STACK_ENTRIES 20
#endif
struct LaneStack {
#if defined(__CUDACC__)
    uint2* smem;      // &shared[threadIdx.x]
    int stride;       // blockDim.x
    int sp;
    RT_HD void store_top(uint2 v) { smem[sp * stride] = v; }
    RT_HD uint2 load_below() const { return smem[(sp > 0 ? sp - 1 : 0) * stride]; }
#else
    uint2 all[RT_STACK_ENTRIES + 1];
    int sp;
    RT_HD void store_top(uint2 v) { all[sp] = v; }
    RT_HD uint2 load_below() const { return all[sp > 0 ? sp - 1 : 0]; }
#endif
};

// box-test reciprocal of a ray direction: ours (never inf); the reference's 1e-8 substitution is kept for d == 0 (RTRay.cs:548-549)
RT_HD f3 box_idir(f3 d) {
    f3 i;
    i.x = 1.0f / (fabsf(d.x) > 1e-20f ? d.x : (d.x < 0.0f ? -1e-20f : (d.x == 0.0f ? 1e-8f : 1e-20f)));
    i.y = 1.0f / (fabsf(d.y) > 1e-20f ? d.y : (d.y < 0.0f ? -1e-20f : (d.y == 0.0f ? 1e-8f : 1e-20f)));
    i.z = 1.0f / (fabsf(d.z) > 1e-20f ? d.z : (d.z < 0.0f ? -1e-20f : (d.z == 0.0f ? 1e-8f : 1e-20f)));
    return i;
}

template <typename T> RT_HD T rt_ldg(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// state of the best hit so far
struct BestHit {
    float t;        // world t (1e30 = none)
    float tObj;     // object-space t of the same hit (== t for identity instances)
    uint32_t rank;
    int inst;
    int prim;       // index into DeviceScene::prims
    float bu, bv;
};

RT_HD bool better_hit(const BestHit& b, float tW, float tO, uint32_t rank, int inst) {
    if (tW < b.t) return true;
    if (tW > b.t) return false;
    if (inst == b.inst) return tO < b.tObj || (tO == b.tObj && rank < b.rank);
    return rank < b.rank;
}

// Slow path: spheres, transformed instances, alpha-tested or closest-invisible triangles.
// Everything goes in and out by value so the caller's traversal state stays in registers.
struct GenResult { bool accept; float tW, tO, bu, bv; };
template <bool ANY_HIT>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
GenResult test_prim_general(const DeviceScene& sc, f3 o, f3 d, float4 q0, float4 q1, float4 q2, uint32_t meta, float tMaxWorld,
                            float bestT, float bestTObj, uint32_t bestRank, int bestInst) {
    GenResult res; res.accept = false; res.tW = 0.0f; res.tO = 0.0f; res.bu = 0.0f; res.bv = 0.0f;
    int inst = (int)(meta & PRIM_INST_MASK);
    f3 oo = o, od = d;
    float scale = 1.0f;
    if (meta & PRIM_XFORM) {   // TransformRay (SceneDeviceViews.cs:46-47,105-107)
        const RtInstanceRecord& ir = sc.instances[inst];
        oo = xf_point(ir.worldToObject, o);
        od = xf_vector(ir.worldToObject, d);
        scale = ir.uniformScale > 0.0f ? ir.uniformScale : 1.0f;
    }
    float tO, bu = 0.0f, bv = 0.0f;
    if (meta & PRIM_SPHERE) {
        if (!intersect_sphere(oo, od, mk3(q0.x, q0.y, q0.z), q1.x, &tO)) return res;
    } else {
        if (!intersect_tri(oo, od, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), &tO, &bu, &bv)) return res;
    }
    if (ANY_HIT) {
        float tMaxObj = tMaxWorld * scale;                       // :107
        if (!(tO > 0.001f && tO < tMaxObj)) return res;          // :257, :292
        res.accept = (meta & PRIM_ALPHA) ? tri_alpha_pass_anyhit(sc, (int)f2u(q0.w), bu, bv) : true;
        return res;
    }
    if (!(tO > 0.001f)) return res;                              // :142, :199
    if (meta & PRIM_NO_CLOSEST) return res;                      // alpha(=1) < AlphaCutoff -> continue (:209,218)
    float tW = tO / scale;                                       // :67
    BestHit b; b.t = bestT; b.tObj = bestTObj; b.rank = bestRank; b.inst = bestInst; b.prim = -1; b.bu = 0.0f; b.bv = 0.0f;
    if (!better_hit(b, tW, tO, f2u(q1.w), inst)) return res;
    if ((meta & PRIM_ALPHA) && !tri_alpha_pass_closest(sc, (int)f2u(q0.w), bu, bv)) return res;   // :215-218
    res.accept = true; res.tW = tW; res.tO = tO; res.bu = bu; res.bv = bv;
    return res;
}

// Bytes J and K of w as the floats 1024 + b: ONE byte permute builds the half2 {0x64bb, 0x64bb} (fp16 1024 + b, exact),
// two HADD2.F32 widen it.  No int->float conversion (slow XU pipe) and half the permutes of a per-byte decode, so the
// box test leans on the FMA pipe instead of the saturated ALU pipe.  With O = o - 1024 a:  fma(m, a, O) = b * a + o.
template <int J, int K> RT_HD void byte_pair_1024(uint32_t w, float& fj, float& fk) {
#if defined(__CUDA_ARCH__)
    const uint32_t h2 = __byte_perm(w, 0x64646464u, 0x4040u | J | (K << 8));
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2));
    fj = f.x; fk = f.y;
#else
    fj = 1024.0f + (float)((w >> (8 * J)) & 0xFFu);
    fk = 1024.0f + (float)((w >> (8 * K)) & 0xFFu);
#endif
}

// One ray's traversal state.  init(), then node_step() / prim_step() until done:
//   node_step : fetch one wide node, test its 8 quantised child boxes, queue the primitives of the leaf children hit
//   prim_step : test ONE queued primitive (the reference's exact intersectors and acceptance rules)
// The kernel votes warp-wide on when to run prim_step so that the exact (expensive, divergent) primitive tests
// execute with many lanes; step() = node_step + all its prim_steps, for single-ray use.
// Closest hit: hit()/result() afterwards.  Any hit: occluded afterwards.
template <bool ANY_HIT, bool COUNT>
struct Traversal {
    f3 o, d, idir;
    float tMax;
    uint32_t octinv;
    uint2 ngroup, tgroup;
    BestHit best;
    bool occluded, done;

    RT_HD void init(f3 o_, f3 d_, f3 idir_, float tMax_, LaneStack& stack) {
        o = o_; d = d_; idir = idir_; tMax = tMax_;
        best.t = ANY_HIT ? tMax_ : 1e30f; best.tObj = best.t; best.rank = 0xFFFFFFFFu; best.inst = -1; best.prim = -1; best.bu = 0.0f; best.bv = 0.0f;
        occluded = false; done = false;
        octinv = 7u - ((idir.x < 0.0f ? 1u : 0u) | (idir.y < 0.0f ? 2u : 0u) | (idir.z < 0.0f ? 4u : 0u));
        stack.sp = 0;
        ngroup = make_uint2(0u, 0x80000000u);   // "child 7^octinv of a virtual parent whose child block starts at node 0" = the root
        tgroup = make_uint2(0u, 0u);
    }

    RT_HD bool has_prims() const { return tgroup.y != 0u; }

    // next node group from the stack, or done (selects only: every lane of the warp runs this together)
    RT_HD void advance(LaneStack& stack) {
        const uint2 top = stack.load_below();
        const bool need = ngroup.y <= 0x00FFFFFFu;
        const bool empty = stack.sp == 0;
        done = done || (need && empty);
        if (need && !empty) { ngroup = top; stack.sp--; }
    }

    // one child: slab test on the decoded quantised planes (m = 1024 + q), then its bits into the hit mask
    template <int J>
    RT_HD void child_test(float nxq, float fxq, float nyq, float fyq, float nzq, float fzq, float Ax, float Ay, float Az,
                          float Ox, float Oy, float Oz, float tFar, uint32_t childBits4, uint32_t bitIndex4, uint32_t& hitmask) const {
        const float t0x = rt_fma(nxq, Ax, Ox), t1x = rt_fma(fxq, Ax, Ox);
        const float t0y = rt_fma(nyq, Ay, Oy), t1y = rt_fma(fyq, Ay, Oy);
        const float t0z = rt_fma(nzq, Az, Oz), t1z = rt_fma(fzq, Az, Oz);
        const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.001f));   // box tMin: SceneDeviceViews.cs:37,131
        const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tFar));
        // conservative: allow for the rounding of either side (quantised planes carry >= 0.01 quantum of slack, rt_bvh.cpp)
        if (tn <= tf * 1.0000007f) {
            const uint32_t cb = (childBits4 >> (8 * J)) & 0xFFu;
            const uint32_t bi = (bitIndex4 >> (8 * J)) & 0xFFu;
            hitmask |= cb << bi;
        }
    }
    // four children whose plane bytes sit in the same six words
    RT_HD void quad_test(uint32_t wnx, uint32_t wfx, uint32_t wny, uint32_t wfy, uint32_t wnz, uint32_t wfz, uint32_t meta4, uint32_t octinv4,
                         float Ax, float Ay, float Az, float Ox, float Oy, float Oz, float tFar, uint32_t& hitmask) const {
        const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t innerMask4 = (isInner4 >> 4) * 0xFFu;
        const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1F1F1F1Fu;
        const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
        float nx0, nx1, nx2, nx3, fx0, fx1, fx2, fx3, ny0, ny1, ny2, ny3, fy0, fy1, fy2, fy3, nz0, nz1, nz2, nz3, fz0, fz1, fz2, fz3;
        byte_pair_1024<0, 1>(wnx, nx0, nx1); byte_pair_1024<2, 3>(wnx, nx2, nx3); byte_pair_1024<0, 1>(wfx, fx0, fx1); byte_pair_1024<2, 3>(wfx, fx2, fx3);
        byte_pair_1024<0, 1>(wny, ny0, ny1); byte_pair_1024<2, 3>(wny, ny2, ny3); byte_pair_1024<0, 1>(wfy, fy0, fy1); byte_pair_1024<2, 3>(wfy, fy2, fy3);
        byte_pair_1024<0, 1>(wnz, nz0, nz1); byte_pair_1024<2, 3>(wnz, nz2, nz3); byte_pair_1024<0, 1>(wfz, fz0, fz1); byte_pair_1024<2, 3>(wfz, fz2, fz3);
        child_test<0>(nx0, fx0, ny0, fy0, nz0, fz0, Ax, Ay, Az, Ox, Oy, Oz, tFar, childBits4, bitIndex4, hitmask);
        child_test<1>(nx1, fx1, ny1, fy1, nz1, fz1, Ax, Ay, Az, Ox, Oy, Oz, tFar, childBits4, bitIndex4, hitmask);
        child_test<2>(nx2, fx2, ny2, fy2, nz2, fz2, Ax, Ay, Az, Ox, Oy, Oz, tFar, childBits4, bitIndex4, hitmask);
        child_test<3>(nx3, fx3, ny3, fy3, nz3, fz3, Ax, Ay, Az, Ox, Oy, Oz, tFar, childBits4, bitIndex4, hitmask);
    }

    // precondition: !done, tgroup.y == 0 (so ngroup.y > 0x00FFFFFF)
    RT_HD void node_step(const DeviceScene& sc, LaneStack& stack, TraceCounters* cnt) {
        const uint32_t hits = ngroup.y;
        const int bit = rt_bfind(hits);
        const uint32_t base = ngroup.x;
        ngroup.y &= ~(1u << bit);
        stack.store_top(ngroup);                          // unconditional store; it only counts if sp moves
        stack.sp += (ngroup.y > 0x00FFFFFFu) ? 1 : 0;
        const uint32_t slot = (uint32_t)(bit - 24) ^ octinv;
        const uint32_t imaskP = hits & 0xFFu;
        const uint32_t rel = (uint32_t)rt_popc(imaskP & ~(0xFFFFFFFFu << slot));
        const WideNode* np = sc.nodes + (base + rel);
        if (COUNT) cnt->nodes++;
        const uint4 n0 = rt_ldg(&np->n0), n1 = rt_ldg(&np->n1), n2 = rt_ldg(&np->n2), n3 = rt_ldg(&np->n3), n4 = rt_ldg(&np->n4);

        const bool nx = idir.x < 0.0f, ny = idir.y < 0.0f, nz = idir.z < 0.0f;
        const uint32_t octinv4 = octinv * 0x01010101u;
        // a = 2^e / d,  O = (p - o) / d - 1024 a   (see byte_pair_1024)
        const float Ax = u2f((n0.w & 0xFFu) << 23) * idir.x, Ay = u2f(((n0.w >> 8) & 0xFFu) << 23) * idir.y, Az = u2f(((n0.w >> 16) & 0xFFu) << 23) * idir.z;
        const float Ox = rt_fma(u2f(n0.x) - o.x, idir.x, -1024.0f * Ax), Oy = rt_fma(u2f(n0.y) - o.y, idir.y, -1024.0f * Ay), Oz = rt_fma(u2f(n0.z) - o.z, idir.z, -1024.0f * Az);
        // near/far plane words chosen by the ray's sign, once per node
        const uint32_t nearx0 = nx ? n3.z : n2.x, nearx1 = nx ? n3.w : n2.y, farx0 = nx ? n2.x : n3.z, farx1 = nx ? n2.y : n3.w;
        const uint32_t neary0 = ny ? n4.x : n2.z, neary1 = ny ? n4.y : n2.w, fary0 = ny ? n2.z : n4.x, fary1 = ny ? n2.w : n4.y;
        const uint32_t nearz0 = nz ? n4.z : n3.x, nearz1 = nz ? n4.w : n3.y, farz0 = nz ? n3.x : n4.z, farz1 = nz ? n3.y : n4.w;
        const float tFar = best.t * sc.tFarScale;
        uint32_t hitmask = 0;
        quad_test(nearx0, farx0, neary0, fary0, nearz0, farz0, n1.z, octinv4, Ax, Ay, Az, Ox, Oy, Oz, tFar, hitmask);
        quad_test(nearx1, farx1, neary1, fary1, nearz1, farz1, n1.w, octinv4, Ax, Ay, Az, Ox, Oy, Oz, tFar, hitmask);
        ngroup.x = n1.x;
        ngroup.y = (hitmask & 0xFF000000u) | (n0.w >> 24);
        tgroup.x = n1.y;
        tgroup.y = hitmask & 0x00FFFFFFu;
        if (tgroup.y == 0u) advance(stack);
    }

    // precondition: !done, tgroup.y != 0
    RT_HD void prim_step(const DeviceScene& sc, LaneStack& stack, TraceCounters* cnt) {
        const int bit = rt_bfind(tgroup.y);
        tgroup.y &= ~(1u << bit);
        const int pi = (int)tgroup.x + bit;
        const PrimRec* pp = sc.prims + pi;
        const float4 q0 = rt_ldg(&pp->q0), q1 = rt_ldg(&pp->q1), q2 = rt_ldg(&pp->q2);
        const uint32_t meta = f2u(q2.w);
        if ((meta & (PRIM_SPHERE | PRIM_XFORM | PRIM_ALPHA | PRIM_NO_CLOSEST)) == 0u) {
            // fast path: plain triangle of an identity instance (object ray == world ray, bit for bit)
            if (COUNT) cnt->tris++;
            float t, bu, bv;
            if (intersect_tri(o, d, mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), &t, &bu, &bv)) {
                if (ANY_HIT) {
                    if (t > 0.001f && t < tMax) { occluded = true; done = true; return; }   // SceneDeviceViews.cs:292,317
                } else if (t > 0.001f) {                                                    // :199
                    const uint32_t rank = f2u(q1.w);
                    const int inst = (int)(meta & PRIM_INST_MASK);
                    if (better_hit(best, t, t, rank, inst)) {
                        best.t = t; best.tObj = t; best.rank = rank; best.inst = inst; best.prim = pi; best.bu = bu; best.bv = bv;
                    }
                }
            }
        } else {
            if (COUNT) { if (meta & PRIM_SPHERE) cnt->spheres++; else cnt->tris++; }
            const GenResult r = test_prim_general<ANY_HIT>(sc, o, d, q0, q1, q2, meta, tMax, best.t, best.tObj, best.rank, best.inst);
            if (r.accept) {
                if (ANY_HIT) { occluded = true; done = true; return; }
                best.t = r.tW; best.tObj = r.tO; best.rank = f2u(q1.w); best.inst = (int)(meta & PRIM_INST_MASK); best.prim = pi; best.bu = r.bu; best.bv = r.bv;
            }
        }
        if (tgroup.y == 0u) advance(stack);
    }

    // one node and all of its primitives; returns true when the traversal is finished
    RT_HD bool step(const DeviceScene& sc, LaneStack& stack, TraceCounters* cnt) {
        node_step(sc, stack, cnt);
        while (!done && tgroup.y != 0u) prim_step(sc, stack, cnt);
        return done;
    }

    RT_HD bool hit() const { return best.t < 1e29f; }   // SceneDeviceViews.cs:85
    RT_HD HitRec result() const { HitRec h; h.t = best.t; h.prim = best.prim; h.bu = best.bu; h.bv = best.bv; return h; }
};

// Convenience wrapper: run one ray to completion.
// Closest hit: returns true and fills *out when something is hit.  Any hit: returns true when occluded.
template <bool ANY_HIT, bool COUNT>
RT_HD bool trace_wide(const DeviceScene& sc, f3 o, f3 d, float tMax, LaneStack& stack, HitRec* out, TraceCounters* cnt) {
    if (sc.nNodes <= 0) { if (!ANY_HIT) { out->t = 1e30f; out->prim = -1; out->bu = 0.0f; out->bv = 0.0f; } return false; }
    Traversal<ANY_HIT, COUNT> tr;
    tr.init(o, d, box_idir(d), tMax, stack);
    while (!tr.step(sc, stack, cnt)) {}
    if (ANY_HIT) return tr.occluded;
    *out = tr.result();
    return tr.hit();
}

}   // namespace rtx
