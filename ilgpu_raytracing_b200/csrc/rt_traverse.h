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
#include <stdio.h>
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
    uint2* smem;           // &shared[threadIdx.x]
    int stride;            // blockDim.x
    int sp;
    const uint32_t* lut;   // hit table (RT_HIT_TABLE_WORDS words of shared memory, filled by fill_hit_table)
#if RT_DEBUG_BOUNDS
    int entries;
    RT_HD void store_top(uint2 v) {
#if defined(__CUDA_ARCH__)
        if (sp < 0 || sp >= entries) { printf("stack overflow: sp %d entries %d\n", sp, entries); __trap(); }
#endif
        smem[sp * stride] = v;
    }
#else
    RT_HD void store_top(uint2 v) { smem[sp * stride] = v; }
#endif
    RT_HD uint2 load_below() const { return smem[(sp > 0 ? sp - 1 : 0) * stride]; }
#else
    uint2 all[RT_STACK_ENTRIES + 1];
    int sp;
    const uint32_t* lut;
    RT_HD void store_top(uint2 v) { all[sp] = v; }
    RT_HD uint2 load_below() const { return all[sp > 0 ? sp - 1 : 0]; }
#endif
};

// Hit table: turns the 8 box-test results of a node into the traversal's hit mask with one shared-memory load.
// Index = octinv * 256 + mr, where bit (7 - s) of mr is set when the child in slot s was MISSED (the order in which
// node_step shifts the sign bits of the eight slab differences together).  Entry:
//   bits 24..31 : slot s hit -> bit 24 + (s ^ octinv)   (octant-ordered visiting: the highest bit is visited first)
//   bits  0..23 : slot s hit -> the 3-bit field 3s..3s+2 (masked with the node's valid24 to give the queued primitives)
#define RT_HIT_TABLE_WORDS 2048
RT_HD uint32_t hit_table_entry(uint32_t octinv, uint32_t mr) {
    uint32_t e = 0;
    for (uint32_t s = 0; s < 8; s++)
        if (!((mr >> (7u - s)) & 1u)) e |= (1u << (24u + (s ^ octinv))) | (7u << (3u * s));
    return e;
}
#if !defined(__CUDACC__)
inline const uint32_t* host_hit_table() {
    static const struct Tab { uint32_t w[RT_HIT_TABLE_WORDS]; Tab() { for (uint32_t i = 0; i < RT_HIT_TABLE_WORDS; i++) w[i] = hit_table_entry(i >> 8, i & 255u); } } tab;
    return tab.w;
}
#endif

// box-test reciprocal of a ray direction: ours (never inf); the reference's 1e-8 substitution is kept for d == 0 (RTRay.cs:548-549)
RT_HD f3 box_idir(f3 d) {
    f3 i;
    i.x = 1.0f / (fabsf(d.x) > 1e-20f ? d.x : (d.x < 0.0f ? -1e-20f : (d.x == 0.0f ? 1e-8f : 1e-20f)));
    i.y = 1.0f / (fabsf(d.y) > 1e-20f ? d.y : (d.y < 0.0f ? -1e-20f : (d.y == 0.0f ? 1e-8f : 1e-20f)));
    i.z = 1.0f / (fabsf(d.z) > 1e-20f ? d.z : (d.z < 0.0f ? -1e-20f : (d.z == 0.0f ? 1e-8f : 1e-20f)));
    return i;
}

#if defined(__CUDACC__)
// The extend kernel derives the reciprocal itself (32 B less DRAM traffic per ray than storing it).  MUFU.RCP + one Newton step
// (error < 1 ulp) instead of the IEEE division: the reciprocal only feeds the conservative box test, whose 6-ulp margin on
// tfar covers 2.5 ulp per side (difference, product, reciprocal at 1 ulp, fma) - never an exact quantity.
__device__ __forceinline__ float rt_rcp_nr(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    const float e = __fmaf_rn(-x, r, 1.0f);
    return __fmaf_rn(r, e, r);
}
__device__ __forceinline__ f3 box_idir_device(f3 d) {
    f3 i;
    i.x = rt_rcp_nr(fabsf(d.x) > 1e-20f ? d.x : (d.x < 0.0f ? -1e-20f : (d.x == 0.0f ? 1e-8f : 1e-20f)));
    i.y = rt_rcp_nr(fabsf(d.y) > 1e-20f ? d.y : (d.y < 0.0f ? -1e-20f : (d.y == 0.0f ? 1e-8f : 1e-20f)));
    i.z = rt_rcp_nr(fabsf(d.z) > 1e-20f ? d.z : (d.z < 0.0f ? -1e-20f : (d.z == 0.0f ? 1e-8f : 1e-20f)));
    return i;
}
#endif

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

// Two plane bytes of a node word as the floats 1024 + b.  sel picks the bytes (PRMT selector: result bytes 0 and 2 from
// the word, bytes 1 and 3 = 0x64 from the constant), so ONE byte permute builds the half2 {0x64bb, 0x64bb} (fp16 1024 + b,
// exact) whichever of the word's lo / hi planes the ray's direction sign makes "near"; two HADD2.F32 widen it.  No
// int->float conversion (slow XU pipe), no per-node near/far selects, and the widening runs on the FMA pipe instead of the
// saturated ALU pipe.  With O = o - 1024 a:  fma(m, a, O) = b * a + o.
//   selector 0x4240 = { byte 0, byte 2 } = the qlo planes of the word's two children, 0x4341 = { byte 1, byte 3 } = the qhi planes
#if !defined(__CUDACC__)
inline float& host_plane_pad() { static float pad = 0.0f; return pad; }
#endif
RT_HD void plane_pair_1024(uint32_t w, uint32_t sel, float& fa, float& fb) {
#if defined(__CUDA_ARCH__)
    const uint32_t h2 = __byte_perm(w, 0x64646464u, sel);
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2));
    fa = f.x; fb = f.y;
#else
    const uint32_t ia = sel & 7u, ib = (sel >> 8) & 7u;
    // host_plane_pad(): analysis knob of the host simulator (tests/hostsim) - child boxes widened by that many quanta on every
    // side, to measure what a coarser (e.g. half-precision) slab test would cost in extra nodes and primitive tests; 0 in every test
#if !defined(__CUDACC__)
    const float pad = host_plane_pad();
#else
    const float pad = 0.0f;   // nvcc's host pass: the shipped test
#endif
    fa = 1024.0f + (float)((w >> (8u * ia)) & 0xFFu) + ((ia & 1u) ? pad : -pad);   // odd bytes are the hi planes
    fb = 1024.0f + (float)((w >> (8u * ib)) & 0xFFu) + ((ib & 1u) ? pad : -pad);
#endif
}
#if defined(__CUDA_ARCH__)
// (a0, a1) * (b, b) + (c, c) as one packed fma.rn.f32x2 (FFMA2 on sm_100)
__device__ __forceinline__ void rt_fma2(float a0, float a1, float b, float c, float& d0, float& d1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
#endif
#ifndef RT_DEBUG_BOUNDS
#define RT_DEBUG_BOUNDS 0   // 1: device-side bounds checks of stack / node / primitive indices (debug builds; compute-sanitizer is not available on the pool)
#endif
#ifndef RT_USE_FFMA2
#define RT_USE_FFMA2 1
#endif
#ifndef RT_PRIM_SLOTS
#define RT_PRIM_SLOTS 2   // queued primitive groups per lane (1 or 2), see Traversal::tgroupB
#endif
RT_HD uint32_t rt_funnel_l1(uint32_t lo, uint32_t hi) {   // (hi << 1) | (lo >> 31)
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, 1);
#else
    return (hi << 1) | (lo >> 31);
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
    uint32_t selNx, selNy, selNz;   // PRMT selectors of the near planes per axis (plane_pair_1024); far = near ^ 0x0101
    const uint32_t* lut;            // this ray's octant slice of the hit table
    uint2 ngroup, tgroup;
    uint32_t tvalid;                // valid24 of the node whose primitives are queued in tgroup
#if RT_PRIM_SLOTS == 2
    uint2 tgroupB;                  // a second queued primitive group: a lane that found primitives can take ONE more node step
    uint32_t tvalidB;               // before its primitive tests run (the closest t it culls with is at most one group stale)
#endif
    BestHit best;
    bool occluded, done;

    RT_HD void init(f3 o_, f3 d_, f3 idir_, float tMax_, LaneStack& stack) {
        o = o_; d = d_; idir = idir_; tMax = tMax_;
        best.t = ANY_HIT ? tMax_ : 1e30f; best.tObj = best.t; best.rank = 0xFFFFFFFFu; best.inst = -1; best.prim = -1; best.bu = 0.0f; best.bv = 0.0f;
        occluded = false; done = false;
        octinv = 7u - ((idir.x < 0.0f ? 1u : 0u) | (idir.y < 0.0f ? 2u : 0u) | (idir.z < 0.0f ? 4u : 0u));
        selNx = idir.x < 0.0f ? 0x4341u : 0x4240u; selNy = idir.y < 0.0f ? 0x4341u : 0x4240u; selNz = idir.z < 0.0f ? 0x4341u : 0x4240u;
#if defined(__CUDACC__)
        lut = stack.lut + octinv * 256u;
#else
        lut = host_hit_table() + octinv * 256u;
#endif
        tvalid = 0u;
#if RT_PRIM_SLOTS == 2
        tgroupB = make_uint2(0u, 0u); tvalidB = 0u;
#endif
        stack.sp = 0;
        ngroup = make_uint2(0u, 0x80000000u);   // "child 7^octinv of a virtual parent whose child block starts at node 0" = the root
        tgroup = make_uint2(0u, 0u);
    }

    RT_HD bool has_prims() const { return tgroup.y != 0u; }
    RT_HD bool has_node_work() const { return ngroup.y > 0x00FFFFFFu; }
#if RT_PRIM_SLOTS == 2
    RT_HD bool can_node_step(const LaneStack&) const { return has_node_work() && tgroupB.y == 0u; }
    RT_HD bool prims_pending() const { return (tgroup.y | tgroupB.y) != 0u; }
#else
    RT_HD bool can_node_step(const LaneStack&) const { return has_node_work() && tgroup.y == 0u; }
    RT_HD bool prims_pending() const { return tgroup.y != 0u; }
#endif

    // next node group from the stack; done when there is neither node work nor a queued primitive left
    // (selects only: every lane of the warp runs this together)
    RT_HD void advance(const DeviceScene&, LaneStack& stack) {
        const uint2 top = stack.load_below();
        const bool need = ngroup.y <= 0x00FFFFFFu;
        const bool empty = stack.sp == 0;
        done = done || (need && empty && !prims_pending());
        if (need && !empty) { ngroup = top; stack.sp--; }
    }

    // one child: slab test on the decoded quantised planes (m = 1024 + q); the sign bit of the slab difference is shifted
    // into mr (set = missed).  No NaN can occur: box_idir keeps |1/d| <= 1e20, so every t is finite.
    RT_HD void child_test(float nxq, float fxq, float nyq, float fyq, float nzq, float fzq, float Ax, float Ay, float Az,
                          float Ox, float Oy, float Oz, float tFar, uint32_t& mr) const {
        const float t0x = rt_fma(nxq, Ax, Ox), t1x = rt_fma(fxq, Ax, Ox);
        const float t0y = rt_fma(nyq, Ay, Oy), t1y = rt_fma(fyq, Ay, Oy);
        const float t0z = rt_fma(nzq, Az, Oz), t1z = rt_fma(fzq, Az, Oz);
        const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.001f));   // box tMin: SceneDeviceViews.cs:37,131
        const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tFar));
        // conservative: hit when tn <= tf * (1 + 6 ulp), allowing for the rounding of either side (the quantised planes carry
        // >= 0.01 quantum of slack on top, rt_bvh.cpp)
        const float diff = rt_fma(tf, 1.0000007f, -tn);
        mr = rt_funnel_l1(f2u(diff), mr);
    }
    // the two children whose plane bytes share one word per axis
    RT_HD void pair_test(uint32_t wx, uint32_t wy, uint32_t wz, uint32_t sFx, uint32_t sFy, uint32_t sFz,
                         float Ax, float Ay, float Az, float Ox, float Oy, float Oz, float tFar, uint32_t& mr) const {
        float nxa, nxb, fxa, fxb, nya, nyb, fya, fyb, nza, nzb, fza, fzb;
        plane_pair_1024(wx, selNx, nxa, nxb); plane_pair_1024(wx, sFx, fxa, fxb);
        plane_pair_1024(wy, selNy, nya, nyb); plane_pair_1024(wy, sFy, fya, fyb);
        plane_pair_1024(wz, selNz, nza, nzb); plane_pair_1024(wz, sFz, fza, fzb);
#if defined(__CUDA_ARCH__) && RT_USE_FFMA2
        // sm_100 packed fp32: one FFMA2 evaluates the same plane of both children (each half is an ordinary IEEE fma)
        float t0xa, t0xb, t1xa, t1xb, t0ya, t0yb, t1ya, t1yb, t0za, t0zb, t1za, t1zb;
        rt_fma2(nxa, nxb, Ax, Ox, t0xa, t0xb); rt_fma2(fxa, fxb, Ax, Ox, t1xa, t1xb);
        rt_fma2(nya, nyb, Ay, Oy, t0ya, t0yb); rt_fma2(fya, fyb, Ay, Oy, t1ya, t1yb);
        rt_fma2(nza, nzb, Az, Oz, t0za, t0zb); rt_fma2(fza, fzb, Az, Oz, t1za, t1zb);
        const float tna = fmaxf(fmaxf(t0xa, t0ya), fmaxf(t0za, 0.001f)), tnb = fmaxf(fmaxf(t0xb, t0yb), fmaxf(t0zb, 0.001f));
        const float tfa = fminf(fminf(t1xa, t1ya), fminf(t1za, tFar)), tfb = fminf(fminf(t1xb, t1yb), fminf(t1zb, tFar));
        mr = rt_funnel_l1(f2u(rt_fma(tfa, 1.0000007f, -tna)), mr);
        mr = rt_funnel_l1(f2u(rt_fma(tfb, 1.0000007f, -tnb)), mr);
#else
        child_test(nxa, fxa, nya, fya, nza, fza, Ax, Ay, Az, Ox, Oy, Oz, tFar, mr);
        child_test(nxb, fxb, nyb, fyb, nzb, fzb, Ax, Ay, Az, Ox, Oy, Oz, tFar, mr);
#endif
    }

    // precondition: !done, can_node_step()
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
#if defined(__CUDA_ARCH__) && RT_DEBUG_BOUNDS
        if ((int)(base + rel) < 0 || (int)(base + rel) >= sc.nNodes) { printf("node index %u + %u outside %d (hits %08x bit %d slot %u)\n", base, rel, sc.nNodes, hits, bit, slot); __trap(); }
#endif
        if (COUNT) cnt->nodes++;
        const uint4 n0 = rt_ldg(&np->n0), n1 = rt_ldg(&np->n1), n2 = rt_ldg(&np->n2), n3 = rt_ldg(&np->n3), n4 = rt_ldg(&np->n4);

        const uint32_t pmask = lut[n1.w >> 24];           // octant-ordered bits of the node's internal children (independent of the box tests)
        // a = 2^e / d,  O = (p - o) / d - 1024 a   (see plane_pair_1024)
        const float Ax = u2f((n0.w & 0xFFu) << 23) * idir.x, Ay = u2f(((n0.w >> 8) & 0xFFu) << 23) * idir.y, Az = u2f(((n0.w >> 16) & 0xFFu) << 23) * idir.z;
        const float Ox = rt_fma(u2f(n0.x) - o.x, idir.x, -1024.0f * Ax), Oy = rt_fma(u2f(n0.y) - o.y, idir.y, -1024.0f * Ay), Oz = rt_fma(u2f(n0.z) - o.z, idir.z, -1024.0f * Az);
        const uint32_t sFx = selNx ^ 0x0101u, sFy = selNy ^ 0x0101u, sFz = selNz ^ 0x0101u;
        const float tFar = best.t * sc.tFarScale;
        uint32_t mr = 0;
        pair_test(n2.x, n3.x, n4.x, sFx, sFy, sFz, Ax, Ay, Az, Ox, Oy, Oz, tFar, mr);
        pair_test(n2.y, n3.y, n4.y, sFx, sFy, sFz, Ax, Ay, Az, Ox, Oy, Oz, tFar, mr);
        pair_test(n2.z, n3.z, n4.z, sFx, sFy, sFz, Ax, Ay, Az, Ox, Oy, Oz, tFar, mr);
        pair_test(n2.w, n3.w, n4.w, sFx, sFy, sFz, Ax, Ay, Az, Ox, Oy, Oz, tFar, mr);
        const uint32_t hitmask = lut[mr];
        ngroup.x = n1.x;
        ngroup.y = (hitmask & pmask & 0xFF000000u) | (n0.w >> 24);
#if RT_PRIM_SLOTS == 2
        const uint32_t found = hitmask & n1.z & 0x00FFFFFFu;
        const bool toA = tgroup.y == 0u;           // precondition: slot B is free
        tgroupB.x = n1.y; tvalidB = n1.z; tgroupB.y = toA ? 0u : found;
        if (toA) { tgroup.x = n1.y; tvalid = n1.z; tgroup.y = found; }
        advance(sc, stack);
#else
        tgroup.x = n1.y;
        tvalid = n1.z;
        tgroup.y = hitmask & n1.z & 0x00FFFFFFu;
        if (tgroup.y == 0u) advance(sc, stack);
#endif
    }

    // precondition: !done, tgroup.y != 0
    RT_HD void prim_step(const DeviceScene& sc, LaneStack& stack, TraceCounters* cnt) {
        const int bit = rt_bfind(tgroup.y);
        tgroup.y &= ~(1u << bit);
        const int pi = (int)tgroup.x + rt_popc(tvalid & ~(0xFFFFFFFFu << bit));   // records are compacted: count the valid bits below
#if defined(__CUDA_ARCH__) && RT_DEBUG_BOUNDS
        if (pi < 0 || pi >= sc.nPrims) { printf("prim index %d outside %d (tgroup %u %08x valid %08x)\n", pi, sc.nPrims, tgroup.x, tgroup.y, tvalid); __trap(); }
#endif
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
        } else if ((meta & (PRIM_SPHERE | PRIM_XFORM | PRIM_ALPHA | PRIM_NO_CLOSEST)) == PRIM_SPHERE) {
            // fast path: sphere of an identity instance (scale 1: tWorld = tObj / 1, tMaxObj = tMaxWorld * 1, SceneDeviceViews.cs:67,107)
            if (COUNT) cnt->spheres++;
            float t;
            if (intersect_sphere(o, d, mk3(q0.x, q0.y, q0.z), q1.x, &t)) {
                if (ANY_HIT) {
                    if (t > 0.001f && t < tMax) { occluded = true; done = true; return; }   // :257
                } else if (t > 0.001f) {                                                    // :142
                    const uint32_t rank = f2u(q1.w);
                    const int inst = (int)(meta & PRIM_INST_MASK);
                    if (better_hit(best, t, t, rank, inst)) {
                        best.t = t; best.tObj = t; best.rank = rank; best.inst = inst; best.prim = pi; best.bu = 0.0f; best.bv = 0.0f;
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
#if RT_PRIM_SLOTS == 2
        if (tgroup.y == 0u) { tgroup = tgroupB; tvalid = tvalidB; tgroupB.y = 0u; }
        advance(sc, stack);
#else
        if (tgroup.y == 0u) advance(sc, stack);
#endif
    }

    // one node and all of its primitives; returns true when the traversal is finished
    RT_HD bool step(const DeviceScene& sc, LaneStack& stack, TraceCounters* cnt) {
        if (has_node_work()) node_step(sc, stack, cnt);
        while (!done && tgroup.y != 0u) prim_step(sc, stack, cnt);
        return done;
    }

    RT_HD bool hit() const { return best.t < 1e29f; }   // SceneDeviceViews.cs:85
    RT_HD HitRec result() const { HitRec h; h.t = best.t; h.prim = best.prim; h.bu = best.bu; h.bv = best.bv; return h; }
};

#if !defined(__CUDACC__)
// Host simulator only: 0 = step() (a node step, then all of its primitives); n > 0 = the per-lane schedule of k_extend - up to n
// node steps while the primitive queue has room, then ONE primitive step - which exercises the queued primitive groups.
inline int& host_lane_schedule() { static int nodeSteps = 0; return nodeSteps; }
#endif

// Convenience wrapper: run one ray to completion.
// Closest hit: returns true and fills *out when something is hit.  Any hit: returns true when occluded.
template <bool ANY_HIT, bool COUNT>
RT_HD bool trace_wide(const DeviceScene& sc, f3 o, f3 d, float tMax, LaneStack& stack, HitRec* out, TraceCounters* cnt) {
    if (sc.nNodes <= 0) { if (!ANY_HIT) { out->t = 1e30f; out->prim = -1; out->bu = 0.0f; out->bv = 0.0f; } return false; }
    Traversal<ANY_HIT, COUNT> tr;
    tr.init(o, d, box_idir(d), tMax, stack);
#if !defined(__CUDACC__)
    if (const int ns = host_lane_schedule()) {
        while (!tr.done) {
            for (int k = 0; k < ns; k++) if (!tr.done && tr.can_node_step(stack)) tr.node_step(sc, stack, cnt);
            if (!tr.done && tr.has_prims()) tr.prim_step(sc, stack, cnt);
        }
    } else
#endif
    while (!tr.step(sc, stack, cnt)) {}
    if (ANY_HIT) return tr.occluded;
    *out = tr.result();
    return tr.hit();
}

}   // namespace rtx
