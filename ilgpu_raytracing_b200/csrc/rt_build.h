// rt_build.h — device-side refit and build of the compressed 8-wide BVH ("GPU-side wide-BVH
// build / refit replacing host Scene.cs:381-510", which makes BvhManager.BuildOrRefit's RebuildPolicy, BvhManager.cs:13-27, real).
//
//   refit : new vertex positions for the SAME topology.  The tree keeps its shape; k_refit_prims rewrites the triangle
//           records and recomputes every primitive box, k_refit_nodes recomputes child boxes / quantisation frames level
//           by level from the leaves up.
//   build : the tree itself on the device - Morton codes of the primitive centroids, radix sort (CUB), the binary radix
//           tree of Karras (HPG 2012), bottom-up boxes, then a level-synchronous greedy collapse to 8-wide nodes (largest
//           surface area first, subtrees of <= 3 primitives become leaf children) with the octant slot assignment of the host
//           builder.  ~10x faster than the host's binned-SAH + optimal-collapse build and a worse tree (measured: DESIGN.md);
//           the primitive stage (validation of every index, the reference's visiting-order ranks, records, padded boxes)
//           stays on the host (rt_bvh.cpp sections 1-2) because it walks the reference's skip-link arrays sequentially.
// Both write nodes with quantize_node(): the formulas of rt_bvh.cpp section 4, so a device-written node is exactly as
// conservative as a host-written one.  Included by rtcore.cu only.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "rt_core.h"

namespace rtx {

// child boxes (world space, float) -> quantisation frame n0 and plane words n2..n4 of a wide node; n1 is the caller's
__device__ inline void quantize_node(WideNode& wn, uint32_t imask, const float (*clo)[3], const float (*chi)[3], const bool* used, const float* nlo, const float* nhi) {
    uint32_t eb[3]; double scale[3]; float pf[3];
    for (int a = 0; a < 3; a++) {
        const double ext = (double)nhi[a] - (double)nlo[a];
        int e = -126;
        if (ext > 0.0) { e = ilogb(ext / 252.0); if (ldexp(1.0, e) < ext / 252.0) e++; while (ext / ldexp(1.0, e) > 252.0) e++; }
        e = max(-126, min(100, e));
        eb[a] = (uint32_t)(e + 127); scale[a] = ldexp(1.0, e);
        pf[a] = (float)((double)nlo[a] - scale[a]);
        if ((double)pf[a] > (double)nlo[a]) pf[a] = nextafterf(pf[a], -INFINITY);
    }
    uint32_t pw[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int s = 0; s < 8; s++)
        for (int a = 0; a < 3; a++) {
            uint32_t ql = 255u, qh = 0u;   // empty slots carry inverted planes and can never be hit
            if (used[s]) {
                const double l = floor(((double)clo[s][a] - (double)pf[a]) / scale[a] - 0.01), h = ceil(((double)chi[s][a] - (double)pf[a]) / scale[a] + 0.01);
                ql = (uint32_t)fmax(0.0, fmin(255.0, l)); qh = (uint32_t)fmax(0.0, fmin(255.0, h));
            }
            pw[a][s >> 1] |= (ql | (qh << 8)) << (16 * (s & 1));   // word k of an axis = { qlo[2k], qhi[2k], qlo[2k+1], qhi[2k+1] }
        }
    wn.n0 = make_uint4(__float_as_uint(pf[0]), __float_as_uint(pf[1]), __float_as_uint(pf[2]), eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24));
    wn.n2 = make_uint4(pw[0][0], pw[0][1], pw[0][2], pw[0][3]);
    wn.n3 = make_uint4(pw[1][0], pw[1][1], pw[1][2], pw[1][3]);
    wn.n4 = make_uint4(pw[2][0], pw[2][1], pw[2][2], pw[2][3]);
}

// ------------------------------------------------------------------------------------------------ refit
__global__ void k_refit_prims(PrimRec* prims, int nPrims, const RtFloat3* pos, const RtMeshTri* tris, const double* instXf, float4* primBox, unsigned* sceneAbsBits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (i < nPrims) {
        PrimRec r = prims[i];
        const uint32_t meta = __float_as_uint(r.q2.w);
        float lo[3], hi[3];
        if (meta & PRIM_SPHERE) {
            const float rad = fabsf(r.q1.x);
            lo[0] = r.q0.x - rad; lo[1] = r.q0.y - rad; lo[2] = r.q0.z - rad; hi[0] = r.q0.x + rad; hi[1] = r.q0.y + rad; hi[2] = r.q0.z + rad;
        } else {
            const RtMeshTri t = tris[(int)__float_as_uint(r.q0.w)];
            const RtFloat3 v0 = pos[t.i0], v1 = pos[t.i1], v2 = pos[t.i2];
            r.q0.x = v0.X; r.q0.y = v0.Y; r.q0.z = v0.Z; r.q1.x = v1.X; r.q1.y = v1.Y; r.q1.z = v1.Z; r.q2.x = v2.X; r.q2.y = v2.Y; r.q2.z = v2.Z;
            prims[i] = r;
            lo[0] = fminf(v0.X, fminf(v1.X, v2.X)); lo[1] = fminf(v0.Y, fminf(v1.Y, v2.Y)); lo[2] = fminf(v0.Z, fminf(v1.Z, v2.Z));
            hi[0] = fmaxf(v0.X, fmaxf(v1.X, v2.X)); hi[1] = fmaxf(v0.Y, fmaxf(v1.Y, v2.Y)); hi[2] = fmaxf(v0.Z, fmaxf(v1.Z, v2.Z));
        }
        if (meta & PRIM_XFORM) {   // world box of the object-space box, as the builder takes it (8 corners, double)
            const double* x = instXf + (size_t)(meta & PRIM_INST_MASK) * 12;
            float wl[3] = {3.4e38f, 3.4e38f, 3.4e38f}, wh[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
            for (int c = 0; c < 8; c++) {
                const double p0 = (c & 1) ? hi[0] : lo[0], p1 = (c & 2) ? hi[1] : lo[1], p2 = (c & 4) ? hi[2] : lo[2];
                for (int a = 0; a < 3; a++) { const float w = (float)(x[a * 4] * p0 + x[a * 4 + 1] * p1 + x[a * 4 + 2] * p2 + x[a * 4 + 3]); wl[a] = fminf(wl[a], w); wh[a] = fmaxf(wh[a], w); }
            }
            for (int a = 0; a < 3; a++) { lo[a] = wl[a]; hi[a] = wh[a]; }
        }
        primBox[2 * i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        primBox[2 * i + 1] = make_float4(hi[0], hi[1], hi[2], (meta & PRIM_XFORM) ? 1.0f : 0.0f);
        for (int a = 0; a < 3; a++) m = fmaxf(m, fmaxf(fabsf(lo[a]), fabsf(hi[a])));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31u) == 0u) atomicMax(sceneAbsBits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}
__global__ void k_refit_nodes(WideNode* nodes, int first, int last, const float4* primBox, float4* nodeBox, const unsigned* sceneAbsBits) {
    const int n = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= last) return;
    const float sceneAbs = __uint_as_float(*sceneAbsBits);
    WideNode wn = nodes[n];
    const uint32_t imask = wn.n0.w >> 24, valid24 = wn.n1.z;
    float clo[8][3], chi[8][3]; bool used[8];
    float nlo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, nhi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int s = 0; s < 8; s++) {
        used[s] = false;
        for (int a = 0; a < 3; a++) { clo[s][a] = 3.4e38f; chi[s][a] = -3.4e38f; }
        if ((imask >> s) & 1u) {
            const int child = (int)wn.n1.x + __popc(imask & ((1u << s) - 1u));
            const float4 l = nodeBox[2 * child], h = nodeBox[2 * child + 1];
            clo[s][0] = l.x; clo[s][1] = l.y; clo[s][2] = l.z; chi[s][0] = h.x; chi[s][1] = h.y; chi[s][2] = h.z;
            used[s] = true;
        } else {
            const uint32_t field = (valid24 >> (3 * s)) & 7u;
            if (field == 0u) continue;
            const int cnt = __popc(field), start = (int)wn.n1.y + __popc(valid24 & ((1u << (3 * s)) - 1u));
            for (int k = 0; k < cnt; k++) {
                const float4 l = primBox[2 * (start + k)], h = primBox[2 * (start + k) + 1];
                const float lo3[3] = {l.x, l.y, l.z}, hi3[3] = {h.x, h.y, h.z};
                for (int a = 0; a < 3; a++) {   // conservative padding, rt_bvh.cpp section 2
                    const float mag = fmaxf(fabsf(lo3[a]), fabsf(hi3[a]));
                    const float pad = 2e-6f * sceneAbs + (h.w != 0.0f ? 2e-5f : 2e-6f) * mag + 1e-30f;
                    clo[s][a] = fminf(clo[s][a], lo3[a] - pad); chi[s][a] = fmaxf(chi[s][a], hi3[a] + pad);
                }
            }
            used[s] = true;
        }
        for (int a = 0; a < 3; a++) { nlo[a] = fminf(nlo[a], clo[s][a]); nhi[a] = fmaxf(nhi[a], chi[s][a]); }
    }
    nodeBox[2 * n] = make_float4(nlo[0], nlo[1], nlo[2], 0.0f);
    nodeBox[2 * n + 1] = make_float4(nhi[0], nhi[1], nhi[2], 0.0f);
    quantize_node(wn, imask, clo, chi, used, nlo, nhi);
    nodes[n] = wn;
}

// ------------------------------------------------------------------------------------------------ build
// Binary radix tree over the Morton-sorted primitives: internal nodes 0 .. n-2, leaves 0 .. n-1 (sorted order).
// A child reference is >= 0 for an internal node and ~leaf for a leaf.
struct LbvhTree {
    int* left; int* right; int* parent;      // parent[] has 2n - 1 entries: internal nodes first, then the leaves
    int* first; int* last;                   // sorted-order range of every internal node
    float4* box;                             // 2 float4 per internal node
    int* visits;                             // bottom-up arrival counters
};

__device__ __forceinline__ uint64_t morton_spread21(uint64_t v) {   // 21 bits -> every third bit
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
__global__ void k_lbvh_morton(const float4* primBox, int n, float3 lo, float3 inv, uint64_t* keys, int* vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 l = primBox[2 * i], h = primBox[2 * i + 1];
    const float cx = (0.5f * (l.x + h.x) - lo.x) * inv.x, cy = (0.5f * (l.y + h.y) - lo.y) * inv.y, cz = (0.5f * (l.z + h.z) - lo.z) * inv.z;
    const uint64_t qx = (uint64_t)fminf(fmaxf(cx * 2097152.0f, 0.0f), 2097151.0f), qy = (uint64_t)fminf(fmaxf(cy * 2097152.0f, 0.0f), 2097151.0f),
                   qz = (uint64_t)fminf(fmaxf(cz * 2097152.0f, 0.0f), 2097151.0f);
    keys[i] = (morton_spread21(qx) << 2) | (morton_spread21(qy) << 1) | morton_spread21(qz);
    vals[i] = i;
}
// common-prefix length of sorted keys i and j (index bits break ties between equal codes); -1 outside the array
__device__ __forceinline__ int lbvh_delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(i ^ j) : __clzll((long long)(a ^ b));
}
__global__ void k_lbvh_tree(const uint64_t* keys, int n, LbvhTree t) {   // Karras 2012, section 3
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dMin = lbvh_delta(keys, n, i, i - d);
    int lMax = 2;
    while (lbvh_delta(keys, n, i, i + lMax * d) > dMin) lMax <<= 1;
    int l = 0;
    for (int s = lMax >> 1; s >= 1; s >>= 1) if (lbvh_delta(keys, n, i, i + (l + s) * d) > dMin) l += s;
    const int j = i + l * d;
    const int dNode = lbvh_delta(keys, n, i, j);
    int sp = 0;
    for (int div = 2, s = (l + 1) >> 1;; div <<= 1, s = (l + div - 1) / div) {
        if (lbvh_delta(keys, n, i, i + (sp + s) * d) > dNode) sp += s;
        if (s <= 1) break;
    }
    const int gamma = i + sp * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = lo == gamma ? ~gamma : gamma, rc = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    t.left[i] = lc; t.right[i] = rc; t.first[i] = lo; t.last[i] = hi;
    t.parent[lc >= 0 ? lc : (n - 1) + ~lc] = i;
    t.parent[rc >= 0 ? rc : (n - 1) + ~rc] = i;
    if (i == 0) t.parent[0] = -1;
}
__global__ void k_lbvh_boxes(const int* vals, const float4* primBox, int n, LbvhTree t) {   // bottom-up: the second arrival at a node unions its children
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cur = t.parent[(n - 1) + i];
    while (cur >= 0) {
        if (atomicAdd(&t.visits[cur], 1) == 0) return;
        __threadfence();
        float4 lo = make_float4(3.4e38f, 3.4e38f, 3.4e38f, 0.0f), hi = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, 0.0f);
        const int ch[2] = {t.left[cur], t.right[cur]};
        for (int k = 0; k < 2; k++) {
            float4 l, h;
            if (ch[k] >= 0) { l = __ldcg(&t.box[2 * ch[k]]); h = __ldcg(&t.box[2 * ch[k] + 1]); }
            else { const int p = vals[~ch[k]]; l = primBox[2 * p]; h = primBox[2 * p + 1]; }
            lo.x = fminf(lo.x, l.x); lo.y = fminf(lo.y, l.y); lo.z = fminf(lo.z, l.z); hi.x = fmaxf(hi.x, h.x); hi.y = fmaxf(hi.y, h.y); hi.z = fmaxf(hi.z, h.z);
        }
        __stcg(&t.box[2 * cur], lo); __stcg(&t.box[2 * cur + 1], hi);
        __threadfence();
        cur = t.parent[cur];
    }
}
__device__ __forceinline__ float box_area(const float* lo, const float* hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return (dx < 0.0f || dy < 0.0f || dz < 0.0f) ? 0.0f : 2.0f * (dx * dy + dy * dz + dz * dx);
}
// One level of the collapse: wide nodes [first, last) each own the binary subtree workB2[w]; their internal children get the
// next contiguous block of wide nodes (atomic counter: level order = breadth-first order), their leaf children the next block
// of primitive records.
__device__ __forceinline__ void lbvh_collapse_node(int w, const LbvhTree& t, const int* vals, const float4* primBox, const PrimRec* primsIn,
                                                   WideNode* nodes, PrimRec* primsOut, int* workB2, int* counters /* [0] nodes, [1] prims */, int leafMax /* 1..3 primitives per leaf child */) {
    const int root = workB2[w];
    int ch[8]; int nch = 2;
    ch[0] = t.left[root]; ch[1] = t.right[root];
    // greedy: open the child of largest area while there is room - first the subtrees that must become internal children
    // (> 3 primitives), then, with the slots that are left, the small subtrees (2-3 primitives): tighter leaf boxes for free
    for (int pass = 0; pass < 2; pass++)
        while (nch < 8) {
            int best = -1; float bestA = -1.0f;
            for (int c = 0; c < nch; c++) {
                if (ch[c] < 0) continue;   // a single primitive
                const int cnt = t.last[ch[c]] - t.first[ch[c]] + 1;
                if (pass == 0 ? cnt <= leafMax : cnt > leafMax) continue;
                const float4 l = t.box[2 * ch[c]], h = t.box[2 * ch[c] + 1];
                const float lo3[3] = {l.x, l.y, l.z}, hi3[3] = {h.x, h.y, h.z};
                const float a = box_area(lo3, hi3);
                if (a > bestA) { bestA = a; best = c; }
            }
            if (best < 0) break;
            const int b = ch[best];
            ch[best] = t.left[b]; ch[nch++] = t.right[b];
        }
    float clo[8][3], chi[8][3], cbl[8][3], cbh[8][3]; bool used[8], leaf[8];
    float nlo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, nhi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int c = 0; c < nch; c++) {
        float4 l, h;
        if (ch[c] >= 0) { l = t.box[2 * ch[c]]; h = t.box[2 * ch[c] + 1]; leaf[c] = t.last[ch[c]] - t.first[ch[c]] + 1 <= leafMax; }
        else { const int p = vals[~ch[c]]; l = primBox[2 * p]; h = primBox[2 * p + 1]; leaf[c] = true; }
        cbl[c][0] = l.x; cbl[c][1] = l.y; cbl[c][2] = l.z; cbh[c][0] = h.x; cbh[c][1] = h.y; cbh[c][2] = h.z;
        for (int a = 0; a < 3; a++) { nlo[a] = fminf(nlo[a], cbl[c][a]); nhi[a] = fmaxf(nhi[a], cbh[c][a]); }
    }
    // slot assignment (rt_bvh.cpp section 4): slot s prefers the child lying furthest against the direction (sx, sy, sz), bit a of s set = negative axis a
    float cost[8][8]; int slotOf[8]; bool slotUsed[8], chDone[8];
    for (int s = 0; s < 8; s++) { slotUsed[s] = false; chDone[s] = false; used[s] = false; for (int a = 0; a < 3; a++) { clo[s][a] = 3.4e38f; chi[s][a] = -3.4e38f; } }
    for (int c = 0; c < nch; c++) {
        float cc[3];
        for (int a = 0; a < 3; a++) cc[a] = 0.5f * (cbl[c][a] + cbh[c][a]) - 0.5f * (nlo[a] + nhi[a]);
        for (int s = 0; s < 8; s++) cost[c][s] = cc[0] * ((s & 1) ? -1.0f : 1.0f) + cc[1] * ((s & 2) ? -1.0f : 1.0f) + cc[2] * ((s & 4) ? -1.0f : 1.0f);
    }
    for (int k = 0; k < nch; k++) {
        int bc = -1, bs = -1; float bv = 3.4e38f;
        for (int c = 0; c < nch; c++) if (!chDone[c]) for (int s = 0; s < 8; s++) if (!slotUsed[s] && cost[c][s] < bv) { bv = cost[c][s]; bc = c; bs = s; }
        chDone[bc] = true; slotUsed[bs] = true; slotOf[bc] = bs;
    }
    int childAt[8];
    for (int s = 0; s < 8; s++) childAt[s] = -1;
    for (int c = 0; c < nch; c++) childAt[slotOf[c]] = c;
    uint32_t imask = 0, valid24 = 0; int nInternal = 0, nLeafPrims = 0;
    for (int s = 0; s < 8; s++) {
        const int c = childAt[s];
        if (c < 0) continue;
        used[s] = true;
        for (int a = 0; a < 3; a++) { clo[s][a] = cbl[c][a]; chi[s][a] = cbh[c][a]; }
        if (!leaf[c]) { imask |= 1u << s; nInternal++; }
        else { const int cnt = ch[c] < 0 ? 1 : t.last[ch[c]] - t.first[ch[c]] + 1; valid24 |= (cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u)) << (3 * s); nLeafPrims += cnt; }
    }
    const int childBase = nInternal ? atomicAdd(&counters[0], nInternal) : 0;
    const int primBase = nLeafPrims ? atomicAdd(&counters[1], nLeafPrims) : 0;
    int ci = childBase, pi = primBase;
    for (int s = 0; s < 8; s++) {
        const int c = childAt[s];
        if (c < 0) continue;
        if (!leaf[c]) { workB2[ci++] = ch[c]; continue; }
        const int f = ch[c] < 0 ? ~ch[c] : t.first[ch[c]], l = ch[c] < 0 ? ~ch[c] : t.last[ch[c]];
        for (int k = f; k <= l; k++, pi++) primsOut[pi] = primsIn[vals[k]];
    }
    uint32_t imr = 0;
    for (int s = 0; s < 8; s++) if (!((imask >> s) & 1u)) imr |= 1u << (7 - s);
    WideNode wn;
    wn.n1 = make_uint4((uint32_t)childBase, (uint32_t)primBase, valid24, imr << 24);
    quantize_node(wn, imask, clo, chi, used, nlo, nhi);
    nodes[w] = wn;
}
// Level `level` of the collapse, with its range read from the DEVICE: levelStart[level] .. levelStart[level + 1] were written by the
// previous level's launch; the last block of this launch to finish (ticket) publishes levelStart[level + 2] = nodes made so far.
// The host queues one launch per possible level back to back and reads the level table once at the end: no host
// synchronisation per tree level.  An empty level's launch exits at once.
__global__ void __launch_bounds__(128) k_lbvh_collapse_level(int level, int* levelStart, unsigned* tickets, int n, LbvhTree t, const int* vals, const float4* primBox,
                                                             const PrimRec* primsIn, WideNode* nodes, PrimRec* primsOut, int* workB2, int* counters, int leafMax) {
    const int first = levelStart[level], last = levelStart[level + 1];
    for (int w = first + blockIdx.x * blockDim.x + threadIdx.x; w < last; w += gridDim.x * blockDim.x)
        lbvh_collapse_node(w, t, vals, primBox, primsIn, nodes, primsOut, workB2, counters, leafMax);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&tickets[level], 1u) == gridDim.x - 1u) {
        __threadfence();
        levelStart[level + 2] = atomicAdd(&counters[0], 0);
    }
}

}   // namespace rtx
