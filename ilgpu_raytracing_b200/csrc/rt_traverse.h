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

namespace rtx {

struct TraceCounters { uint32_t nodes, tris, spheres; };

// Per-lane traversal stack.  Device: first RT_SMEM_STACK entries live in shared memory (strided by
// the block size so lanes hit distinct banks), the rest spill to local memory.  Host: plain array.
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 12
#endif
#define RT_STACK_TOTAL 40
struct LaneStack {
#if defined(__CUDACC__)
    uint2* smem;      // &shared[threadIdx.x], stride = blockDim.x
    int stride;
    uint2 spill[RT_STACK_TOTAL - RT_SMEM_STACK];
    int sp;
    RT_HD void push(uint2 v) {
        if (sp < RT_SMEM_STACK) smem[sp * stride] = v; else spill[sp - RT_SMEM_STACK] = v;
        sp++;
    }
    RT_HD uint2 pop() {
        sp--;
        return (sp < RT_SMEM_STACK) ? smem[sp * stride] : spill[sp - RT_SMEM_STACK];
    }
#else
    uint2 all[RT_STACK_TOTAL];
    int sp;
    RT_HD void push(uint2 v) { all[sp++] = v; }
    RT_HD uint2 pop() { return all[--sp]; }
#endif
};

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

RT_HD float byte_to_float(uint32_t w, int j) { return (float)((w >> (8 * j)) & 0xFFu); }

// One ray's traversal state.  init() then step() until it returns true.  A step = fetch one wide node,
// test its 8 quantised child boxes, test the primitives of the leaf children that were hit, pop.
// Closest hit: hit()/result() afterwards.  Any hit: occluded afterwards.
template <bool ANY_HIT, bool COUNT>
struct Traversal {
    f3 o, d, idir;
    float tMax;
    uint32_t octinv;
    uint2 ngroup, tgroup;
    BestHit best;
    bool occluded;

    RT_HD void init(f3 o_, f3 d_, float tMax_, LaneStack& stack) {
        o = o_; d = d_; tMax = tMax_;
        best.t = ANY_HIT ? tMax_ : 1e30f; best.tObj = best.t; best.rank = 0xFFFFFFFFu; best.inst = -1; best.prim = -1; best.bu = 0.0f; best.bv = 0.0f;
        occluded = false;
        // box-test reciprocal: ours (never inf); the reference's 1e-8 substitution is kept for d == 0 (RTRay.cs:548-549)
        idir.x = 1.0f / (fabsf(d.x) > 1e-20f ? d.x : (d.x < 0.0f ? -1e-20f : (d.x == 0.0f ? 1e-8f : 1e-20f)));
        idir.y = 1.0f / (fabsf(d.y) > 1e-20f ? d.y : (d.y < 0.0f ? -1e-20f : (d.y == 0.0f ? 1e-8f : 1e-20f)));
        idir.z = 1.0f / (fabsf(d.z) > 1e-20f ? d.z : (d.z < 0.0f ? -1e-20f : (d.z == 0.0f ? 1e-8f : 1e-20f)));
        octinv = 7u - ((idir.x < 0.0f ? 1u : 0u) | (idir.y < 0.0f ? 2u : 0u) | (idir.z < 0.0f ? 4u : 0u));
        stack.sp = 0;
        ngroup = make_uint2(0u, 0x80000000u);   // "child 7^octinv of a virtual parent whose child block starts at node 0" = the root
        tgroup = make_uint2(0u, 0u);
    }

    RT_HD bool step(const DeviceScene& sc, LaneStack& stack, TraceCounters* cnt) {
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t hits = ngroup.y;
            const int bit = rt_bfind(hits);
            const uint32_t base = ngroup.x;
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00FFFFFFu) stack.push(ngroup);
            const uint32_t slot = (uint32_t)(bit - 24) ^ octinv;
            const uint32_t imaskP = hits & 0xFFu;
            const uint32_t rel = (uint32_t)rt_popc(imaskP & ~(0xFFFFFFFFu << slot));
            const WideNode* np = sc.nodes + (base + rel);
            if (COUNT) cnt->nodes++;
            const uint4 n0 = rt_ldg(&np->n0), n1 = rt_ldg(&np->n1), n2 = rt_ldg(&np->n2), n3 = rt_ldg(&np->n3), n4 = rt_ldg(&np->n4);

            const bool nx = idir.x < 0.0f, ny = idir.y < 0.0f, nz = idir.z < 0.0f;
            const uint32_t octinv4 = octinv * 0x01010101u;
            const float sx = u2f((n0.w & 0xFFu) << 23), sy = u2f(((n0.w >> 8) & 0xFFu) << 23), sz = u2f(((n0.w >> 16) & 0xFFu) << 23);
            const float aix = sx * idir.x, aiy = sy * idir.y, aiz = sz * idir.z;
            const float ox = (u2f(n0.x) - o.x) * idir.x, oy = (u2f(n0.y) - o.y) * idir.y, oz = (u2f(n0.z) - o.z) * idir.z;
            // near/far plane words chosen by the ray's sign, once per node
            const uint32_t nearx0 = nx ? n3.z : n2.x, nearx1 = nx ? n3.w : n2.y, farx0 = nx ? n2.x : n3.z, farx1 = nx ? n2.y : n3.w;
            const uint32_t neary0 = ny ? n4.x : n2.z, neary1 = ny ? n4.y : n2.w, fary0 = ny ? n2.z : n4.x, fary1 = ny ? n2.w : n4.y;
            const uint32_t nearz0 = nz ? n4.z : n3.x, nearz1 = nz ? n4.w : n3.y, farz0 = nz ? n3.x : n4.z, farz1 = nz ? n3.y : n4.w;
            const float tFar = best.t;
            const float tMinRay = 0.001f;   // SceneDeviceViews.cs:37,131 (box tMin)
            uint32_t hitmask = 0;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const uint32_t meta4 = half ? n1.w : n1.z;
                const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                const uint32_t innerMask4 = (isInner4 >> 4) * 0xFFu;
                const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1F1F1F1Fu;
                const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
                const uint32_t wnx = half ? nearx1 : nearx0, wfx = half ? farx1 : farx0;
                const uint32_t wny = half ? neary1 : neary0, wfy = half ? fary1 : fary0;
                const uint32_t wnz = half ? nearz1 : nearz0, wfz = half ? farz1 : farz0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float t0x = rt_fma(byte_to_float(wnx, j), aix, ox), t1x = rt_fma(byte_to_float(wfx, j), aix, ox);
                    const float t0y = rt_fma(byte_to_float(wny, j), aiy, oy), t1y = rt_fma(byte_to_float(wfy, j), aiy, oy);
                    const float t0z = rt_fma(byte_to_float(wnz, j), aiz, oz), t1z = rt_fma(byte_to_float(wfz, j), aiz, oz);
                    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tMinRay));
                    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tFar));
                    // conservative: allow for the rounding of either side
                    if (tn <= tf * 1.0000007f) {
                        const uint32_t cb = (childBits4 >> (8 * j)) & 0xFFu;
                        const uint32_t bi = (bitIndex4 >> (8 * j)) & 0xFFu;
                        hitmask |= cb << bi;
                    }
                }
            }
            ngroup.x = n1.x;
            ngroup.y = (hitmask & 0xFF000000u) | (n0.w >> 24);
            tgroup.x = n1.y;
            tgroup.y = hitmask & 0x00FFFFFFu;
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }

        while (tgroup.y != 0u) {
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
                        if (t > 0.001f && t < tMax) { occluded = true; return true; }   // SceneDeviceViews.cs:292,317
                    } else if (t > 0.001f) {                                            // :199
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
                    if (ANY_HIT) { occluded = true; return true; }
                    best.t = r.tW; best.tObj = r.tO; best.rank = f2u(q1.w); best.inst = (int)(meta & PRIM_INST_MASK); best.prim = pi; best.bu = r.bu; best.bv = r.bv;
                }
            }
        }

        if (ngroup.y <= 0x00FFFFFFu) {
            if (stack.sp == 0) return true;
            ngroup = stack.pop();
        }
        return false;
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
    tr.init(o, d, tMax, stack);
    while (!tr.step(sc, stack, cnt)) {}
    if (ANY_HIT) return tr.occluded;
    *out = tr.result();
    return tr.hit();
}

}   // namespace rtx
