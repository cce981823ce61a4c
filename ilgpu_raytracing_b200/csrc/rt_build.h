// rt_build.h — device-side refit and build of the compressed 8-wide BVH ("GPU-side wide-BVH
// build / refit replacing host Scene.cs:381-510", which makes BvhManager.BuildOrRefit's RebuildPolicy, BvhManager.cs:13-27, real).
//
//   refit : new vertex positions for the SAME topology.  The tree keeps its shape; k_refit_prims rewrites the triangle
//           records and recomputes every primitive box, k_refit_nodes recomputes child boxes / quantisation frames level
//           by level from the leaves up.
//   build : the tree itself on the device - Morton codes of the primitive centroids (cubic cells), radix sort (CUB), then TWO
//           binary trees over that order: the radix tree of Karras (HPG 2012) with bottom-up boxes, and PLOC (Meister & Bittner
//           2018).  Every binary node carries the table of the SAH-optimal 8-wide collapse (the dynamic program of the host
//           builder, evaluated as the node is made), the tree whose collapsed cost is lower is kept, and a level-synchronous
//           pass writes the wide nodes the tables chose (leaf children of <= 3 primitives) with the octant slot assignment of
//           the host builder.  The primitive stage (validation of every index, the reference's visiting-order ranks, records,
//           padded boxes) stays on the host (rt_bvh.cpp sections 1-2, all threads) because it starts from a sequential walk of
//           the reference's skip-link arrays.  Measured against the host's binned-SAH tree: DESIGN.md.
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
    int* left; int* right;                   // children of every internal node
    int* count;                              // primitives below every internal node
    float4* box;                             // 2 float4 per internal node
    int* parent;                             // radix tree only: 2n - 1 entries, internal nodes first, then the leaves
    int* visits;                             // radix tree only: bottom-up arrival counters
    float4* dp;                              // per internal node 2 float4: C(n, 1..7) and the packed decisions (lbvh_dp_node)
    float cPrim;                             // cost of one exact primitive test relative to one wide-node step (rt_bvh.cpp section 4)
};
__device__ __forceinline__ int lbvh_count(const LbvhTree& t, int ref) { return ref < 0 ? 1 : t.count[ref]; }
// the (<= 3) leaves below ref, left to right
__device__ __forceinline__ int lbvh_leaves(const LbvhTree& t, int ref, int* out) {
    int n = 0, st[4], sp = 0;
    st[sp++] = ref;
    while (sp) { const int r = st[--sp]; if (r < 0) out[n++] = ~r; else { st[sp++] = t.right[r]; st[sp++] = t.left[r]; } }
    return n;
}

__device__ __forceinline__ float box_area4(float4 lo, float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return (dx < 0.0f || dy < 0.0f || dz < 0.0f) ? 0.0f : 2.0f * (dx * dy + dy * dz + dz * dx);
}
// SAH-optimal collapse to 8-wide, the dynamic program of Ylitie, Karras & Laine (HPG 2017, section 4.1) as the host builder runs
// it (rt_bvh.cpp section 4), evaluated the moment a binary node is made - its children's tables are complete by then:
//   C(n,1) = min(C_leaf(n), C_internal(n)),  C_internal(n) = A_n + min_k C(left,k) + C(right,8-k),  C_leaf(n) = A_n P_n c_prim if P_n <= 3
//   C(n,i) = min(C(n,i-1), min_k C(left,k) + C(right,i-k)),  i = 2..7;   a single primitive: C(n,i) = A_n c_prim
// Decisions: 4 bits per i (0 = "leaf child" for i = 1 / "as with i - 1 slots" for i > 1, else the left subtree's share k);
// bits 28-30 keep the best k of C_internal even when the leaf won (a root of <= 3 primitives is opened all the same).
__device__ __forceinline__ void lbvh_dp_load(const LbvhTree& t, int ref, float area, float* C) {
    if (ref < 0) { for (int i = 0; i < 7; i++) C[i] = area * t.cPrim; return; }
    const float4 a = __ldcg(&t.dp[2 * ref]), b = __ldcg(&t.dp[2 * ref + 1]);   // L2: the radix tree's tables are written by other blocks of the SAME launch
    C[0] = a.x; C[1] = a.y; C[2] = a.z; C[3] = a.w; C[4] = b.x; C[5] = b.y; C[6] = b.z;
}
__device__ __forceinline__ uint32_t lbvh_dp_decisions(const LbvhTree& t, int ref) { return __float_as_uint(__ldcg(&t.dp[2 * ref + 1]).w); }
__device__ __forceinline__ void lbvh_dp_node(const LbvhTree& t, int id, int refL, float areaL, int refR, float areaR, float areaN, int count) {
    float CL[7], CR[7], Cn[7];
    lbvh_dp_load(t, refL, areaL, CL); lbvh_dp_load(t, refR, areaR, CR);
    float best = 3.4e38f; int bk = 1;
    for (int k = 1; k <= 7; k++) { const float c = CL[k - 1] + CR[7 - k]; if (c < best) { best = c; bk = k; } }
    const float cInt = best + areaN, cLeaf = count <= 3 ? areaN * (float)count * t.cPrim : 3.4e38f;
    uint32_t D = (uint32_t)bk << 28;
    if (cLeaf <= cInt) Cn[0] = cLeaf; else { Cn[0] = cInt; D |= (uint32_t)bk; }
    for (int i = 2; i <= 7; i++) {
        float bi = Cn[i - 2]; int d = 0;
        for (int k = 1; k < i; k++) { const float c = CL[k - 1] + CR[i - k - 1]; if (c < bi) { bi = c; d = k; } }
        Cn[i - 1] = bi; D |= (uint32_t)d << (4 * (i - 1));
    }
    __stcg(&t.dp[2 * id], make_float4(Cn[0], Cn[1], Cn[2], Cn[3]));
    __stcg(&t.dp[2 * id + 1], make_float4(Cn[4], Cn[5], Cn[6], __uint_as_float(D)));
}

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
    t.left[i] = lc; t.right[i] = rc; t.count[i] = hi - lo + 1;
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
        {
            float4 bl[2], bh[2];
            for (int k = 0; k < 2; k++) {
                if (ch[k] >= 0) { bl[k] = __ldcg(&t.box[2 * ch[k]]); bh[k] = __ldcg(&t.box[2 * ch[k] + 1]); }
                else { const int p = vals[~ch[k]]; bl[k] = primBox[2 * p]; bh[k] = primBox[2 * p + 1]; }
            }
            lbvh_dp_node(t, cur, ch[0], box_area4(bl[0], bh[0]), ch[1], box_area4(bl[1], bh[1]), box_area4(lo, hi), t.count[cur]);
        }
        __threadfence();
        cur = t.parent[cur];
    }
}
// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner, TVCG 2018) over the Morton-sorted primitives ----
// Clusters live in Morton order.  One iteration: every cluster looks RT_PLOC_RADIUS places to either side for the neighbour whose
// union with it has the smallest surface area; two clusters that chose EACH OTHER merge into a new binary node, which takes the
// place of the left one; the survivors are compacted in order.  The merge key is (area, i xor j): symmetric in the pair and
// unique per neighbour, so the globally smallest key is always a mutual choice (every iteration merges at least one pair) and
// runs of identical boxes pair up (i, i ^ 1) instead of forming a chain that merges one pair per iteration.
// No host round trip per iteration: the cluster count lives on the device (nc[iteration parity]), every launch is sized for the
// first iteration and idles through what is no longer there; the host reads the count once per batch of iterations.
#define RT_PLOC_TILE 256
#define RT_PLOC_RADIUS 8   // measured 8 / 16 / 32 / 64 on the C4 height field and on a debris field: the collapsed SAH cost is lowest at 8 on both
#define RT_PLOC_GONE ((int)0x80000000)
struct PlocState {
    int* cid; float4* cbox;        // clusters of this iteration: node reference (>= 0 internal, ~k = the k-th sorted primitive) and box
    int* tcid; float4* tbox;       // the same after the merge step, in place (RT_PLOC_GONE = merged into its left partner)
    int* tileCount; int* tileOffset;
    int* nc;                       // nc[2]: cluster count by iteration parity
    int* nodeCounter;
};
__global__ void k_ploc_init(const int* vals, const float4* primBox, int n, PlocState S) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int p = vals[k];
    S.cid[k] = ~k; S.cbox[2 * k] = primBox[2 * p]; S.cbox[2 * k + 1] = primBox[2 * p + 1];
}
__global__ void __launch_bounds__(RT_PLOC_TILE) k_ploc_merge(int parity, PlocState S, LbvhTree t) {
    constexpr int R = RT_PLOC_RADIUS, T = RT_PLOC_TILE;
    __shared__ float4 slo[T + 4 * R], shi[T + 4 * R];
    __shared__ int snn[T + 2 * R];
    const int nc = S.nc[parity];
    const int tiles = (nc + T - 1) / T;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int t0 = tile * T;
        for (int k = threadIdx.x; k < T + 4 * R; k += T) {
            const int g = t0 - 2 * R + k;
            if (g >= 0 && g < nc) { slo[k] = S.cbox[2 * g]; shi[k] = S.cbox[2 * g + 1]; }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < T + 2 * R; k += T) {   // nearest neighbour of the tile's clusters and of R more on either side
            const int g = t0 - R + k;
            int nn = -1;
            if (g >= 0 && g < nc) {
                const float4 al = slo[k + R], ah = shi[k + R];
                unsigned long long best = ~0ull;
                for (int dj = -R; dj <= R; dj++) {
                    const int j = g + dj;
                    if (dj == 0 || j < 0 || j >= nc) continue;
                    const float4 bl = slo[k + R + dj], bh = shi[k + R + dj];
                    const float dx = fmaxf(ah.x, bh.x) - fminf(al.x, bl.x), dy = fmaxf(ah.y, bh.y) - fminf(al.y, bl.y), dz = fmaxf(ah.z, bh.z) - fminf(al.z, bl.z);
                    const float area = dx * dy + dy * dz + dz * dx;   // >= 0: orders like its bit pattern
                    const unsigned long long key = ((unsigned long long)__float_as_uint(area) << 32) | (unsigned)(g ^ j);
                    if (key < best) { best = key; nn = j; }
                }
            }
            snn[k] = nn;
        }
        __syncthreads();
        const int i = t0 + (int)threadIdx.x;
        bool keep = false;
        if (i < nc) {
            const int j = snn[threadIdx.x + R];
            const bool mutual = j >= 0 && snn[j - t0 + R] == i;
            int ref = S.cid[i];
            float4 lo = slo[threadIdx.x + 2 * R], hi = shi[threadIdx.x + 2 * R];
            keep = !mutual || i < j;
            if (mutual && i < j) {
                const int other = S.cid[j];
                const float4 bl = slo[j - t0 + 2 * R], bh = shi[j - t0 + 2 * R];
                lo = make_float4(fminf(lo.x, bl.x), fminf(lo.y, bl.y), fminf(lo.z, bl.z), 0.0f);
                hi = make_float4(fmaxf(hi.x, bh.x), fmaxf(hi.y, bh.y), fmaxf(hi.z, bh.z), 0.0f);
                const int id = atomicAdd(S.nodeCounter, 1);
                t.left[id] = ref; t.right[id] = other; t.count[id] = lbvh_count(t, ref) + lbvh_count(t, other);
                t.box[2 * id] = lo; t.box[2 * id + 1] = hi;
                lbvh_dp_node(t, id, ref, box_area4(slo[threadIdx.x + 2 * R], shi[threadIdx.x + 2 * R]), other, box_area4(bl, bh), box_area4(lo, hi), t.count[id]);
                ref = id;
            }
            S.tcid[i] = keep ? ref : RT_PLOC_GONE;
            if (keep) { S.tbox[2 * i] = lo; S.tbox[2 * i + 1] = hi; }
        }
        const int cnt = __syncthreads_count(keep);   // also fences the shared arrays before the next tile
        if (threadIdx.x == 0) S.tileCount[tile] = cnt;
    }
}
__global__ void __launch_bounds__(1024) k_ploc_scan(int parity, PlocState S) {   // one block: tile offsets and the next cluster count
    __shared__ int warpSum[32];
    __shared__ int carry;
    const int nc = S.nc[parity];
    const int tiles = (nc + RT_PLOC_TILE - 1) / RT_PLOC_TILE;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int k = base + (int)threadIdx.x;
        const int v = k < tiles ? S.tileCount[k] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((threadIdx.x & 31) >= (unsigned)o) incl += u; }
        if ((threadIdx.x & 31) == 31) warpSum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warpSum[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (threadIdx.x >= (unsigned)o) wi += u; }
            warpSum[threadIdx.x] = wi - w;   // exclusive
        }
        __syncthreads();
        const int excl = carry + warpSum[threadIdx.x >> 5] + incl - v;
        if (k < tiles) S.tileOffset[k] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) S.nc[parity ^ 1] = tiles ? carry : nc;
}
__global__ void __launch_bounds__(RT_PLOC_TILE) k_ploc_scatter(int parity, PlocState S) {
    __shared__ int warpBase[RT_PLOC_TILE / 32];
    const int nc = S.nc[parity];
    const int tiles = (nc + RT_PLOC_TILE - 1) / RT_PLOC_TILE;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int i = tile * RT_PLOC_TILE + (int)threadIdx.x;
        const int ref = i < nc ? S.tcid[i] : RT_PLOC_GONE;
        const bool keep = ref != RT_PLOC_GONE;
        const unsigned b = __ballot_sync(0xFFFFFFFFu, keep);
        if ((threadIdx.x & 31) == 0) warpBase[threadIdx.x >> 5] = __popc(b);
        __syncthreads();
        int base = S.tileOffset[tile];
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) base += warpBase[w];
        if (keep) {
            const int pos = base + __popc(b & ((1u << (threadIdx.x & 31)) - 1u));
            S.cid[pos] = ref; S.cbox[2 * pos] = S.tbox[2 * i]; S.cbox[2 * pos + 1] = S.tbox[2 * i + 1];
        }
        __syncthreads();
    }
}

// One level of the collapse: wide nodes [first, last) each own the binary subtree workB2[w]; their internal children get the
// next contiguous block of wide nodes (atomic counter: level order = breadth-first order), their leaf children the next block
// of primitive records.
__device__ __forceinline__ void lbvh_collapse_node(int w, const LbvhTree& t, const int* vals, const float4* primBox, const PrimRec* primsIn,
                                                   WideNode* nodes, PrimRec* primsOut, int* workB2, int* counters /* [0] nodes, [1] prims */) {
    const int root = workB2[w];
    int ch[8]; bool leafDp[8]; int nch = 0;
    {
        // the children the dynamic program chose: the root's best split k, then each side's decisions down to single slots
        const uint32_t Dr = lbvh_dp_decisions(t, root);
        int stRef[16], stI[16], sp = 0;
        const int k0 = (int)(Dr >> 28) & 7;
        stRef[sp] = t.right[root]; stI[sp++] = 8 - k0;
        stRef[sp] = t.left[root]; stI[sp++] = k0;
        while (sp) {
            const int ref = stRef[--sp]; int i = stI[sp];
            if (ref < 0) { leafDp[nch] = true; ch[nch++] = ref; continue; }
            const uint32_t D = lbvh_dp_decisions(t, ref);
            int d = 0;
            while (i > 1 && (d = (int)(D >> (4 * (i - 1))) & 15) == 0) i--;   // "no better than with one slot fewer"
            if (i == 1) { leafDp[nch] = (D & 15u) == 0u; ch[nch++] = ref; continue; }
            stRef[sp] = t.right[ref]; stI[sp++] = i - d;
            stRef[sp] = t.left[ref]; stI[sp++] = d;
        }
    }
    float clo[8][3], chi[8][3], cbl[8][3], cbh[8][3]; bool used[8], leaf[8];
    float nlo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, nhi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int c = 0; c < nch; c++) {
        float4 l, h;
        if (ch[c] >= 0) { l = t.box[2 * ch[c]]; h = t.box[2 * ch[c] + 1]; leaf[c] = leafDp[c]; }
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
        else { const int cnt = lbvh_count(t, ch[c]); valid24 |= (cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u)) << (3 * s); nLeafPrims += cnt; }
    }
    const int childBase = nInternal ? atomicAdd(&counters[0], nInternal) : 0;
    const int primBase = nLeafPrims ? atomicAdd(&counters[1], nLeafPrims) : 0;
    int ci = childBase, pi = primBase;
    for (int s = 0; s < 8; s++) {
        const int c = childAt[s];
        if (c < 0) continue;
        if (!leaf[c]) { workB2[ci++] = ch[c]; continue; }
        int lv[3];
        const int nl = lbvh_leaves(t, ch[c], lv);
        for (int k = 0; k < nl; k++, pi++) primsOut[pi] = primsIn[vals[lv[k]]];
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
                                                             const PrimRec* primsIn, WideNode* nodes, PrimRec* primsOut, int* workB2, int* counters) {
    const int first = levelStart[level], last = levelStart[level + 1];
    for (int w = first + blockIdx.x * blockDim.x + threadIdx.x; w < last; w += gridDim.x * blockDim.x)
        lbvh_collapse_node(w, t, vals, primBox, primsIn, nodes, primsOut, workB2, counters);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&tickets[level], 1u) == gridDim.x - 1u) {
        __threadfence();
        levelStart[level + 2] = atomicAdd(&counters[0], 0);
    }
}

}   // namespace rtx
