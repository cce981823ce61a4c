// rt_bvh.cpp — host-side scene flattening: reference BVH2 arrays -> ranked primitive list ->
// binned-SAH binary BVH -> compressed 8-wide BVH (80-byte nodes) + 48-byte primitive records.
//
// Replaces, as the acceleration structure the device walks, the TLAS/BLAS pair built by
// Scene.BuildTLASNodeRecursive / BuildBLASNodeRecursive (Engine/Scene.cs:405-510).  The reference
// arrays are still the INPUT (rt_scene_upload receives what Scene.UploadAll uploads, Scene.cs:258-279):
// they define which (instance, primitive) pairs exist and the order in which the reference visits
// them, which is the tie-break rank for equal-t hits.
#include "rt_bvh.h"

#include <algorithm>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>

namespace rtx {

namespace {

struct Aabb {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = std::numeric_limits<float>::max(); hi[a] = -std::numeric_limits<float>::max(); } }
    void grow(const Aabb& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float* p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct BuildPrim { Aabb box; float c[3]; PrimRec rec; };

struct B2Node { Aabb box; int left, right, first, count; };

bool is_identity(const RtInstanceRecord& ir) {
    const RtAffine3x4& m = ir.worldToObject;
    const RtAffine3x4& n = ir.objectToWorld;
    auto id = [](const RtAffine3x4& a) {
        return a.m00 == 1.0f && a.m01 == 0.0f && a.m02 == 0.0f && a.m03 == 0.0f && a.m10 == 0.0f && a.m11 == 1.0f && a.m12 == 0.0f && a.m13 == 0.0f &&
               a.m20 == 0.0f && a.m21 == 0.0f && a.m22 == 1.0f && a.m23 == 0.0f;
    };
    return id(m) && id(n) && ir.uniformScale == 1.0f;
}

// inverse of the 3x4 affine worldToObject in double precision: where does an object-space point live in the world?
bool invert_affine(const RtAffine3x4& m, double inv[12]) {
    double a = m.m00, b = m.m01, c = m.m02, d = m.m10, e = m.m11, f = m.m12, g = m.m20, h = m.m21, i = m.m22;
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (!(std::fabs(det) > 1e-30)) return false;
    double id = 1.0 / det;
    double r[9] = {(e * i - f * h) * id, (c * h - b * i) * id, (b * f - c * e) * id, (f * g - d * i) * id, (a * i - c * g) * id, (c * d - a * f) * id,
                   (d * h - e * g) * id, (b * g - a * h) * id, (a * e - b * d) * id};
    double t[3] = {m.m03, m.m13, m.m23};
    for (int rr = 0; rr < 3; rr++) {
        inv[rr * 4 + 0] = r[rr * 3 + 0]; inv[rr * 4 + 1] = r[rr * 3 + 1]; inv[rr * 4 + 2] = r[rr * 3 + 2];
        inv[rr * 4 + 3] = -(r[rr * 3 + 0] * t[0] + r[rr * 3 + 1] * t[1] + r[rr * 3 + 2] * t[2]);
    }
    return true;
}

struct Builder {
    std::vector<BuildPrim> prims;
    std::vector<int> idx;
    std::vector<B2Node> b2;
    bool median = false;   // depth-bounded rebuild: object-median splits along the longest centroid axis (depth = ceil(log2 n))

    // The binary tree goes down to single primitives, so a subtree over n primitives has exactly 2n - 1 nodes: node indices are
    // known before the subtree is built (depth-first order: node, left subtree, right subtree) and big subtrees can be built by
    // their own threads into the pre-sized array - the result is the array the sequential recursion would push, byte for byte.
    void build_all(int n) {
        b2.assign((size_t)2 * (size_t)n - 1, B2Node());
        build_at(0, 0, n, 0);
    }
    void build_at(int ni, int first, int count, int forks) {
        Aabb box, cb; box.reset(); cb.reset();
        for (int i = first; i < first + count; i++) { box.grow(prims[idx[i]].box); cb.grow(prims[idx[i]].c); }
        b2[ni].box = box; b2[ni].left = b2[ni].right = -1; b2[ni].first = first; b2[ni].count = count;
        if (count == 1) return;

        // binned SAH over the centroid bounds
        const int NB = 16;
        int bestAxis = -1, bestSplit = -1; float bestCost = std::numeric_limits<float>::max();
        for (int a = 0; a < 3 && !median; a++) {
            float ext = cb.hi[a] - cb.lo[a];
            if (!(ext > 0.0f)) continue;
            Aabb bb[NB]; int bc[NB];
            for (int b = 0; b < NB; b++) { bb[b].reset(); bc[b] = 0; }
            float k = (float)NB / ext;
            for (int i = first; i < first + count; i++) {
                const BuildPrim& p = prims[idx[i]];
                int b = std::min(NB - 1, std::max(0, (int)((p.c[a] - cb.lo[a]) * k)));
                bb[b].grow(p.box); bc[b]++;
            }
            float rightArea[NB]; Aabb acc; acc.reset(); int n = 0;
            for (int b = NB - 1; b > 0; b--) { acc.grow(bb[b]); rightArea[b] = acc.area(); }
            acc.reset(); int rn = count;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bb[b]); n += bc[b]; rn = count - n;
                if (n == 0 || rn == 0) continue;
                float cost = acc.area() * (float)n + rightArea[b + 1] * (float)rn;
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = b; }
            }
        }
        // the binary tree goes down to single primitives; the wide collapse below decides where leaves (<= 3 primitives) form
        int mid;
        if (median) {
            int a = 0;
            for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[a] - cb.lo[a]) a = k;
            mid = first + count / 2;
            std::nth_element(idx.begin() + first, idx.begin() + mid, idx.begin() + first + count, [&](int x, int y) {
                const float cx = prims[x].c[a], cy = prims[y].c[a];
                return cx < cy || (cx == cy && x < y);
            });
        } else if (bestAxis >= 0) {
            float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
            float k = 16.0f / ext; float lo = cb.lo[bestAxis]; int a = bestAxis, sp = bestSplit;
            auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](int pi) {
                int b = std::min(15, std::max(0, (int)((prims[pi].c[a] - lo) * k)));
                return b <= sp;
            });
            mid = (int)(it - idx.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        } else {
            mid = first + count / 2;   // all centroids coincide
        }
        const int nl = mid - first, l = ni + 1, r = ni + 2 * nl;
        b2[ni].left = l; b2[ni].right = r;
        if (count >= (1 << 15) && forks < 4) {   // up to 16 threads at the top of a big tree
            std::thread t; bool forked = true;
            try { t = std::thread([=] { build_at(l, first, nl, forks + 1); }); } catch (...) { forked = false; }   // no thread to be had: build it here
            build_at(r, mid, count - nl, forks + 1);
            if (forked) t.join(); else build_at(l, first, nl, forks + 1);
        } else {
            build_at(l, first, nl, forks);
            build_at(r, mid, count - nl, forks);
        }
    }
};

inline uint32_t pack4(const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); }
inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

}   // namespace

bool build_wide_bvh(const RtSceneDesc& d, HostBvh& out, std::string& err, PrimSink* sink, int maxDepth) {
    if (maxDepth <= 0 || maxDepth > RT_STACK_ENTRIES - 2) maxDepth = RT_STACK_ENTRIES - 2;
    out.nodes.clear(); out.prims.clear(); out.levelStart.clear(); out.stats = HostBvhStats();
    out.instBoxXf.assign((size_t)std::max<int64_t>(0, d.nInstances) * 12, 0.0);
    for (int64_t i = 0; i < d.nInstances; i++) { out.instBoxXf[(size_t)i * 12 + 0] = 1.0; out.instBoxXf[(size_t)i * 12 + 5] = 1.0; out.instBoxXf[(size_t)i * 12 + 10] = 1.0; }
    Builder B;
    const bool timing = getenv("RT_BUILD_TIMING") && atoi(getenv("RT_BUILD_TIMING")) != 0;   // tuning runs: phases of the primitive stage on stderr
    auto tPhase = std::chrono::steady_clock::now();
    auto phase = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "rtcore_b200 host stage: %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tPhase).count());
        tPhase = now;
    };

    // ---- 1. enumerate (instance, primitive) pairs in the reference's visiting order --------------------------------
    // TLAS walk with every box test taken (SceneDeviceViews.cs:33-84): leaves in left-first skip-link order.
    std::vector<int> instOrder;
    if (d.nTlasNodes > 0) {
        int64_t guard = 0; int cur = 0;
        while (cur != -1) {
            if (cur < 0 || cur >= d.nTlasNodes || ++guard > 4 * d.nTlasNodes + 8) { err = "tlasNodes: bad link"; return false; }
            const RtBvhNode& n = d.tlasNodes[cur];
            if (n.count > 0) {
                for (int i = n.first; i < n.first + n.count; i++) {
                    if (i < 0 || i >= d.nTlasInstanceIndices) { err = "tlasNodes: leaf range outside tlasInstanceIndices"; return false; }
                    int ii = d.tlasInstanceIndices[i];
                    if (ii < 0 || ii >= d.nInstances) { err = "tlasInstanceIndices: bad instance index"; return false; }
                    instOrder.push_back(ii);
                }
                cur = n.skipIndex;
            } else cur = n.left;
        }
    }
    if (instOrder.size() > (size_t)PRIM_INST_MASK) { err = "too many instances"; return false; }

    // (a) sequential and cheap: the BLAS walks themselves (SceneDeviceViews.cs:127-168 / :176-235, every box test taken) - one hop
    // per node, no per-primitive work - yield the leaves in visiting order with the rank of their first primitive; (b) the
    // per-primitive work (index validation, records, boxes) then runs on all host threads into pre-computed slots: the same
    // records and the same ranks as a sequential pass, whatever the thread count.
    struct InstInfo { int ii; bool ident, sph; double w2oInv[12]; };
    struct LeafItem { int inst; int first, count; uint32_t rank; };
    std::vector<InstInfo> infos(instOrder.size());
    std::vector<LeafItem> leaves;
    uint32_t rank = 0;
    for (size_t io = 0; io < instOrder.size(); io++) {
        int ii = instOrder[io];
        const RtInstanceRecord& ir = d.instances[ii];
        InstInfo& info = infos[io];
        info.ii = ii; info.ident = is_identity(ir);
        if (ir.uniformScale > out.stats.maxInstanceScale) out.stats.maxInstanceScale = ir.uniformScale;
        if (!info.ident && !invert_affine(ir.worldToObject, info.w2oInv)) { err = "instance worldToObject is singular"; return false; }
        if (!info.ident) for (int k = 0; k < 12; k++) out.instBoxXf[(size_t)ii * 12 + k] = info.w2oInv[k];
        info.sph = ir.type == RT_BLAS_SPHERESET;
        if (!info.sph && ir.type != RT_BLAS_TRIMESH) { err = "instance: unknown BlasType"; return false; }
        int blasStart = ir.blasRoot, blasEnd = ir.blasRoot + ir.blasNodeCount;
        if (blasStart < 0 || blasEnd > d.nBlasNodes) { err = "instance: BLAS range outside blasNodes"; return false; }
        // Big BLAS: the walk below, cut into segments that all host threads follow at once.  Cut points are subtree roots taken from
        // the left / right pointers of the top levels, in left-first order; segment j is the SAME deterministic walk (left child at
        // an internal node, skip link at a leaf) started at cut point j and followed until it arrives at cut point j + 1 (the last
        // one: until it leaves the BLAS).  A terminating walk visits no node twice, so if every segment arrives, their concatenation
        // IS the sequential walk - whatever the pointers used to choose the cut points are worth.  A segment that does not arrive
        // (pointers inconsistent with the links, a bad link, a bad leaf range) sends the whole BLAS to the sequential walk below,
        // which is also what reports errors.
        if (ir.blasNodeCount >= (1 << 15)) {
            const unsigned hwT = std::thread::hardware_concurrency();
            const size_t nT = std::max<size_t>(1, std::min<size_t>(hwT ? hwT : 4, 32));
            const int64_t limit = info.sph ? d.nSpherePrimIdx : d.nTriPrimIdx;
            std::vector<int> cuts(1, blasStart);
            bool cutsOk = true;
            while (cutsOk && cuts.size() < 8 * nT) {   // expand every internal cut point into its children, left first
                std::vector<int> next; next.reserve(cuts.size() * 2);
                bool grew = false;
                for (int r : cuts) {
                    const RtBvhNode& n = d.blasNodes[r];
                    if (n.count > 0) { next.push_back(r); continue; }
                    if (n.left < blasStart || n.left >= blasEnd || n.right < blasStart || n.right >= blasEnd) { cutsOk = false; break; }
                    next.push_back(n.left); next.push_back(n.right); grew = true;
                }
                if (!grew) break;
                cuts.swap(next);
            }
            if (cutsOk && cuts.size() > 1) {
                const size_t nSeg = cuts.size();
                std::vector<uint64_t> primsOf(nSeg, 0), leavesOf(nSeg, 0); std::vector<char> okOf(nSeg, 0);
                auto follow = [&](size_t j, LeafItem* out, uint32_t rk) -> bool {   // out == null: count only
                    const int stop = j + 1 < nSeg ? cuts[j + 1] : -2;
                    int cur = cuts[j]; int64_t steps = 0; uint64_t np = 0, nl = 0;
                    for (;;) {
                        if (cur == stop) break;
                        if (cur == -1 || cur >= blasEnd) { if (stop != -2) return false; break; }   // left the BLAS: only the last segment may
                        if (cur < blasStart || ++steps > (int64_t)ir.blasNodeCount + 8) return false;
                        const RtBvhNode& n = d.blasNodes[cur];
                        if (n.count > 0) {
                            if (n.first < 0 || (int64_t)n.first + n.count > limit) return false;
                            if (out) { out[nl] = {(int)io, n.first, n.count, rk}; rk += (uint32_t)n.count; }
                            np += (uint64_t)n.count; nl++;
                            cur = n.skipIndex;
                        } else cur = n.left;
                    }
                    primsOf[j] = np; leavesOf[j] = nl;
                    return true;
                };
                auto run = [&](const std::function<void(size_t)>& f) {   // segments dealt out round-robin
                    std::vector<std::thread> pool;
                    for (size_t t = 1; t < nT; t++) { try { pool.emplace_back(f, t); } catch (...) { f(t); } }
                    f(0);
                    for (auto& th : pool) th.join();
                };
                run([&](size_t t) { for (size_t j = t; j < nSeg; j += nT) okOf[j] = follow(j, nullptr, 0u) ? 1 : 0; });
                bool ok = true; uint64_t totalPrims = 0, totalLeaves = 0;
                for (size_t j = 0; j < nSeg; j++) { ok = ok && okOf[j]; totalPrims += primsOf[j]; totalLeaves += leavesOf[j]; }
                if (ok && (uint64_t)rank + totalPrims <= 0x7FFFFFFFull) {
                    const size_t base = leaves.size();
                    leaves.resize(base + (size_t)totalLeaves);
                    std::vector<uint64_t> leafAt(nSeg, 0), rankAt(nSeg, 0);
                    uint64_t la = 0, ra = rank;
                    for (size_t j = 0; j < nSeg; j++) { leafAt[j] = la; rankAt[j] = ra; la += leavesOf[j]; ra += primsOf[j]; }
                    run([&](size_t t) { for (size_t j = t; j < nSeg; j += nT) follow(j, leaves.data() + base + (size_t)leafAt[j], (uint32_t)rankAt[j]); });
                    rank += (uint32_t)totalPrims;
                    continue;
                }
            }
        }
        int cur = blasStart; int64_t guard = 0;
        while (cur != -1 && cur < blasEnd) {
            if (cur < blasStart || ++guard > 4 * (int64_t)ir.blasNodeCount + 8) { err = "blasNodes: bad link"; return false; }
            const RtBvhNode& n = d.blasNodes[cur];
            if (n.count > 0) {
                const int64_t limit = info.sph ? d.nSpherePrimIdx : d.nTriPrimIdx;
                if (n.first < 0 || (int64_t)n.first + n.count > limit) { err = info.sph ? "BLAS leaf range outside spherePrimIdx" : "BLAS leaf range outside triPrimIdx"; return false; }
                if ((uint64_t)rank + (uint64_t)n.count > 0x7FFFFFFFull) { err = "too many primitives"; return false; }
                leaves.push_back({(int)io, n.first, n.count, rank});
                rank += (uint32_t)n.count;
                cur = n.skipIndex;
            } else cur = n.left;
        }
    }
    phase("walk of the reference's BVH2 arrays");
    if (sink) { if (rank && (!sink->reserve || !sink->reserve(sink, (size_t)rank))) { err = "out of memory for the primitive staging arrays"; return false; } }
    else B.prims.resize((size_t)rank);
    unsigned hw = std::thread::hardware_concurrency();
    const size_t nThreads = (rank < (1u << 15)) ? 1 : std::max<size_t>(1, std::min<size_t>(hw ? hw : 4, 32));
    auto run_threads = [&](const std::function<void(size_t)>& work) {
        if (nThreads == 1) { work(0); return; }
        std::vector<std::thread> pool;
        for (size_t t = 1; t < nThreads; t++) { try { pool.emplace_back(work, t); } catch (...) { work(t); } }
        work(0);
        for (auto& th : pool) th.join();
    };
    std::vector<float> absMax(nThreads, 0.0f);
    {
        const size_t nLeaves = leaves.size();
        std::vector<std::string> errs(nThreads); std::vector<size_t> errAt(nThreads, (size_t)-1);
        std::vector<int64_t> nSph(nThreads, 0), nTri(nThreads, 0);
        auto work = [&](size_t t) {
            const size_t lo = nLeaves * t / nThreads, hi = nLeaves * (t + 1) / nThreads;
            bool failed = false;
            auto bad = [&](size_t li, const char* m) { if (!failed) { failed = true; errAt[t] = li; errs[t] = m; } };
            float myAbs = 0.0f; int64_t mySph = 0, myTri = 0;   // locals: the per-thread slots share cache lines
            for (size_t li = lo; li < hi && !failed; li++) {
                const LeafItem& L = leaves[li];
                const InstInfo& info = infos[(size_t)L.inst];
                const int ii = info.ii;
                for (int i = L.first; i < L.first + L.count; i++) {
                    const uint32_t rk = L.rank + (uint32_t)(i - L.first);
                    BuildPrim bp; memset(&bp, 0, sizeof(bp));
                    uint32_t meta = (uint32_t)ii;
                    if (!info.ident) meta |= PRIM_XFORM;
                    float obb[2][3];
                    if (info.sph) {
                        int prim = d.spherePrimIdx[i];
                        if (prim < 0 || prim >= d.nSpheres) { bad(li, "spherePrimIdx: bad sphere index"); break; }
                        const RtSphere& s = d.spheres[prim];
                        meta |= PRIM_SPHERE;
                        bp.rec.q0 = make_float4(s.center.X, s.center.Y, s.center.Z, bitsf((uint32_t)prim));
                        bp.rec.q1 = make_float4(s.radius, 0.0f, 0.0f, bitsf(rk));
                        bp.rec.q2 = make_float4(0.0f, 0.0f, 0.0f, bitsf(meta));
                        float r = std::fabs(s.radius);
                        obb[0][0] = s.center.X - r; obb[0][1] = s.center.Y - r; obb[0][2] = s.center.Z - r;
                        obb[1][0] = s.center.X + r; obb[1][1] = s.center.Y + r; obb[1][2] = s.center.Z + r;
                        mySph++;
                    } else {
                        int tri = d.triPrimIdx[i];
                        if (tri < 0 || tri >= d.nMeshTris) { bad(li, "triPrimIdx: bad triangle index"); break; }
                        const RtMeshTri& tr = d.meshTris[tri];
                        if (tr.i0 < 0 || tr.i1 < 0 || tr.i2 < 0 || tr.i0 >= d.nMeshPositions || tr.i1 >= d.nMeshPositions || tr.i2 >= d.nMeshPositions) { bad(li, "meshTris: bad vertex index"); break; }
                        if (tri >= d.nTriMatIndex || tri >= d.nMeshTriUVs) { bad(li, "triMatIndex/meshTriUVs shorter than meshTris"); break; }
                        int mi = d.triMatIndex[tri];
                        if (mi < 0 || mi >= d.nMaterials) { bad(li, "triMatIndex: bad material index"); break; }
                        const RtMeshTriUV& tuv = d.meshTriUVs[tri];
                        if (tuv.t0 < 0 || tuv.t1 < 0 || tuv.t2 < 0 || tuv.t0 >= d.nMeshTexcoords || tuv.t1 >= d.nMeshTexcoords || tuv.t2 >= d.nMeshTexcoords) { bad(li, "meshTriUVs: bad texcoord index"); break; }
                        const RtMaterialRecord& m = d.materials[mi];
                        int64_t nTexLen = d.nTexInfos > 0 ? d.nTexInfos : 1;   // AllocateOrEmpty (Scene.cs:370-377)
                        bool alphaMap = m.HasAlphaMap != 0 && m.AlphaTexIndex >= 0 && m.AlphaTexIndex < nTexLen;
                        if (alphaMap) meta |= PRIM_ALPHA;
                        else if (1.0f < m.AlphaCutoff) meta |= PRIM_NO_CLOSEST;
                        const RtFloat3 &v0 = d.meshPositions[tr.i0], &v1 = d.meshPositions[tr.i1], &v2 = d.meshPositions[tr.i2];
                        bp.rec.q0 = make_float4(v0.X, v0.Y, v0.Z, bitsf((uint32_t)tri));
                        bp.rec.q1 = make_float4(v1.X, v1.Y, v1.Z, bitsf(rk));
                        bp.rec.q2 = make_float4(v2.X, v2.Y, v2.Z, bitsf(meta));
                        obb[0][0] = std::min(v0.X, std::min(v1.X, v2.X)); obb[0][1] = std::min(v0.Y, std::min(v1.Y, v2.Y)); obb[0][2] = std::min(v0.Z, std::min(v1.Z, v2.Z));
                        obb[1][0] = std::max(v0.X, std::max(v1.X, v2.X)); obb[1][1] = std::max(v0.Y, std::max(v1.Y, v2.Y)); obb[1][2] = std::max(v0.Z, std::max(v1.Z, v2.Z));
                        myTri++;
                    }
                    bool finite = true;
                    for (int a = 0; a < 3; a++) if (!std::isfinite(obb[0][a]) || !std::isfinite(obb[1][a])) finite = false;
                    if (!finite) { bad(li, "non-finite primitive bounds"); break; }
                    bp.box.reset();
                    if (info.ident) { bp.box.grow(obb[0]); bp.box.grow(obb[1]); }
                    else {
                        for (int c = 0; c < 8; c++) {
                            double p[3] = {obb[c & 1][0], obb[(c >> 1) & 1][1], obb[(c >> 2) & 1][2]};
                            float w[3];
                            for (int r = 0; r < 3; r++) w[r] = (float)(info.w2oInv[r * 4] * p[0] + info.w2oInv[r * 4 + 1] * p[1] + info.w2oInv[r * 4 + 2] * p[2] + info.w2oInv[r * 4 + 3]);
                            bp.box.grow(w);
                        }
                    }
                    if (sink) {
                        sink->prims[rk] = bp.rec;
                        sink->boxes[2 * (size_t)rk] = make_float4(bp.box.lo[0], bp.box.lo[1], bp.box.lo[2], 0.0f);
                        sink->boxes[2 * (size_t)rk + 1] = make_float4(bp.box.hi[0], bp.box.hi[1], bp.box.hi[2], 0.0f);
                        for (int a = 0; a < 3; a++) myAbs = std::max(myAbs, std::max(std::fabs(bp.box.lo[a]), std::fabs(bp.box.hi[a])));
                    } else B.prims[rk] = bp;
                }
            }
            absMax[t] = myAbs; nSph[t] = mySph; nTri[t] = myTri;
        };
        run_threads(work);
        size_t firstBad = (size_t)-1, who = 0;
        for (size_t t = 0; t < nThreads; t++) { out.stats.nSpheres += nSph[t]; out.stats.nTris += nTri[t]; if (errAt[t] < firstBad) { firstBad = errAt[t]; who = t; } }
        if (firstBad != (size_t)-1) { err = errs[who]; return false; }   // the first offending leaf in visiting order, like a sequential pass
    }
    phase("records (all threads)");
    const int N = (int)rank;
    out.stats.nPrims = N;
    if (N == 0) return true;   // empty scene: everything misses

    if (sink) {   // section 2 below, on all threads and in place: the same padding arithmetic, the same boxes
        if (sink->recordsDone) sink->recordsDone(sink, (size_t)N);
        float sceneAbs = 0.0f;
        for (float m : absMax) sceneAbs = std::max(sceneAbs, m);
        std::vector<Aabb> sbs(nThreads);
        run_threads([&](size_t t) {
            Aabb sb; sb.reset();
            const size_t lo = (size_t)N * t / nThreads, hi = (size_t)N * (t + 1) / nThreads;
            for (size_t i = lo; i < hi; i++) {
                const bool xf = (fbits(sink->prims[i].q2.w) & PRIM_XFORM) != 0;
                float4& bl = sink->boxes[2 * i]; float4& bh = sink->boxes[2 * i + 1];
                float* l3[3] = {&bl.x, &bl.y, &bl.z}; float* h3[3] = {&bh.x, &bh.y, &bh.z};
                for (int a = 0; a < 3; a++) {
                    float mag = std::max(std::fabs(*l3[a]), std::fabs(*h3[a]));
                    float pad = 2e-6f * sceneAbs + (xf ? 2e-5f : 2e-6f) * mag + 1e-30f;
                    *l3[a] -= pad; *h3[a] += pad;
                    sb.lo[a] = std::min(sb.lo[a], *l3[a]); sb.hi[a] = std::max(sb.hi[a], *h3[a]);
                }
            }
            sbs[t] = sb;
        });
        Aabb sb; sb.reset();
        for (const Aabb& b : sbs) if (b.lo[0] <= b.hi[0]) sb.grow(b);
        for (int a = 0; a < 3; a++) { out.stats.sceneLo[a] = sb.lo[a]; out.stats.sceneHi[a] = sb.hi[a]; }
        phase("padding (all threads)");
        return true;
    }

    // ---- 2. conservative padding (rounding of the exact primitive tests and of the quantised slab test) ----------
    float sceneAbs = 0.0f;
    for (auto& p : B.prims) for (int a = 0; a < 3; a++) sceneAbs = std::max(sceneAbs, std::max(std::fabs(p.box.lo[a]), std::fabs(p.box.hi[a])));
    for (auto& p : B.prims) {
        bool xf = (fbits(p.rec.q2.w) & PRIM_XFORM) != 0;
        for (int a = 0; a < 3; a++) {
            float mag = std::max(std::fabs(p.box.lo[a]), std::fabs(p.box.hi[a]));
            float pad = 2e-6f * sceneAbs + (xf ? 2e-5f : 2e-6f) * mag + 1e-30f;
            p.box.lo[a] -= pad; p.box.hi[a] += pad;
            p.c[a] = 0.5f * (p.box.lo[a] + p.box.hi[a]);
        }
    }

    // The SAH tree minimises cost, not depth: a scene with a huge dynamic range (a small detailed mesh on a giant ground plane)
    // can come out deeper than the traversal stack.  The reference's skip-link walk has no stack and accepts any scene, so a
    // too-deep tree is rebuilt depth-bounded instead of refusing the commit: object-median splits (binary depth ceil(log2 n))
    // collapsed three binary levels per wide node (wide depth <= ceil(31 / 3) + 1 for any 32-bit primitive count).
    for (int attempt = 0; attempt < 2; attempt++) {
    const bool bounded = attempt == 1;
    out.nodes.clear(); out.prims.clear(); out.levelStart.clear();
    // ---- 3. binary BVH (binned SAH, leaves of <= 3 primitives) ---------------------------------------------------
    B.median = bounded;
    B.idx.resize((size_t)N);
    for (int i = 0; i < N; i++) B.idx[i] = i;
    B.build_all(N);

    // ---- 4. collapse to 8-wide with the SAH-optimal dynamic program of Ylitie, Karras & Laine (HPG 2017, section 4.1):
    //   C(n,1) = min(C_leaf(n), C_internal(n)),  C_internal(n) = A_n c_node + min_k C(left,k) + C(right,8-k),
    //   C(n,i) = min(C(n,i-1), min_k C(left,k) + C(right,i-k)),  i = 2..7,   C_leaf(n) = A_n P_n c_prim if P_n <= 3.
    const int nb2 = (int)B.b2.size();
    float cPrimTune = 1.6f;   // cost of one exact primitive test relative to one wide-node step (tuning sweeps: RT_BVH_CPRIM)
    if (const char* e = getenv("RT_BVH_CPRIM")) { float v = (float)atof(e); if (v > 0.0f) cPrimTune = v; }
    const float cNode = 1.0f, cPrim = cPrimTune, INF = std::numeric_limits<float>::max();
    std::vector<float> C((size_t)nb2 * 7); std::vector<uint8_t> D((size_t)nb2 * 7);
    for (int n = nb2 - 1; n >= 0; n--) {   // children have larger indices than their parent
        const B2Node& bn = B.b2[n];
        const float A = bn.box.area();
        float* Cn = &C[(size_t)n * 7]; uint8_t* Dn = &D[(size_t)n * 7];
        if (bn.left < 0) { for (int i = 0; i < 7; i++) { Cn[i] = A * cPrim; Dn[i] = 0; } continue; }
        const float* CL = &C[(size_t)bn.left * 7]; const float* CR = &C[(size_t)bn.right * 7];
        float best = INF; int bk = 1;
        for (int k = 1; k <= 7; k++) { float c = CL[k - 1] + CR[7 - k]; if (c < best) { best = c; bk = k; } }
        const float cInt = best + A * cNode;
        const float cLeaf = bn.count <= 3 ? A * (float)bn.count * cPrim : INF;
        if (cLeaf <= cInt) { Cn[0] = cLeaf; Dn[0] = 0; } else { Cn[0] = cInt; Dn[0] = (uint8_t)bk; }
        for (int i = 2; i <= 7; i++) {
            float bi = Cn[i - 2]; int d = 0;
            for (int k = 1; k < i; k++) { float c = CL[k - 1] + CR[i - k - 1]; if (c < bi) { bi = c; d = k; } }
            Cn[i - 1] = bi; Dn[i - 1] = (uint8_t)d;
        }
    }
    struct Child { int b2; bool leaf; };
    struct Collector {
        const std::vector<B2Node>& b2; const std::vector<uint8_t>& D;
        void collect(int n, int i, Child* out, int& cnt) const {
            const B2Node& bn = b2[n];
            if (bn.left < 0) { out[cnt++] = {n, true}; return; }
            if (i == 1) { out[cnt++] = {n, D[(size_t)n * 7] == 0}; return; }
            int d = D[(size_t)n * 7 + i - 1];
            if (d == 0) collect(n, i - 1, out, cnt);
            else { collect(bn.left, d, out, cnt); collect(bn.right, i - d, out, cnt); }
        }
        // depth-bounded variant: the subtrees `levels` binary levels below n (a subtree of <= 3 primitives becomes a leaf child)
        void collect_fixed(int n, int levels, Child* out, int& cnt) const {
            const B2Node& bn = b2[n];
            if (bn.count <= 3) { out[cnt++] = {n, true}; return; }
            if (levels == 0) { out[cnt++] = {n, false}; return; }
            collect_fixed(bn.left, levels - 1, out, cnt); collect_fixed(bn.right, levels - 1, out, cnt);
        }
    } collector{B.b2, D};

    struct Work { int b2; int wide; };
    std::vector<Work> queue;
    out.nodes.reserve((size_t)N / 6 + 16);
    out.prims.reserve((size_t)N);
    out.nodes.push_back(WideNode());
    queue.push_back({0, 0});
    int maxDepthSeen = 0;
    std::vector<int> depthOf; depthOf.push_back(1);
    for (size_t qi = 0; qi < queue.size(); qi++) {
        const Work w = queue[qi];
        const B2Node& root = B.b2[w.b2];
        Child ch[8]; int nch = 0;
        if (root.left < 0 || (w.wide == 0 && root.count <= 3 && N <= 3)) { ch[nch++] = {w.b2, true}; }   // a tree of <= 3 primitives: one leaf child
        else if (bounded) { collector.collect_fixed(root.left, 2, ch, nch); collector.collect_fixed(root.right, 2, ch, nch); }
        else {
            int k = D[(size_t)w.b2 * 7];
            if (k == 0) {   // only the root can get here with a "leaf" decision: force an internal node
                const float* CL = &C[(size_t)root.left * 7]; const float* CR = &C[(size_t)root.right * 7];
                float best = INF; k = 1;
                for (int kk = 1; kk <= 7; kk++) { float c = CL[kk - 1] + CR[7 - kk]; if (c < best) { best = c; k = kk; } }
            }
            collector.collect(root.left, k, ch, nch);
            collector.collect(root.right, 8 - k, ch, nch);
        }
        // slot assignment: slot s prefers the child lying furthest "against" the direction (sx,sy,sz), bit a of s set = negative axis a
        Aabb nb = root.box;
        float ncx[3] = {0.5f * (nb.lo[0] + nb.hi[0]), 0.5f * (nb.lo[1] + nb.hi[1]), 0.5f * (nb.lo[2] + nb.hi[2])};
        float cost[8][8]; int slotOf[8]; bool slotUsed[8] = {false, false, false, false, false, false, false, false}; bool chDone[8] = {false, false, false, false, false, false, false, false};
        for (int c = 0; c < nch; c++) {
            const Aabb& cb = B.b2[ch[c].b2].box;
            float cc[3] = {0.5f * (cb.lo[0] + cb.hi[0]) - ncx[0], 0.5f * (cb.lo[1] + cb.hi[1]) - ncx[1], 0.5f * (cb.lo[2] + cb.hi[2]) - ncx[2]};
            for (int s = 0; s < 8; s++) cost[c][s] = cc[0] * ((s & 1) ? -1.0f : 1.0f) + cc[1] * ((s & 2) ? -1.0f : 1.0f) + cc[2] * ((s & 4) ? -1.0f : 1.0f);
        }
        for (int k = 0; k < nch; k++) {
            int bc = -1, bs = -1; float bv = std::numeric_limits<float>::max();
            for (int c = 0; c < nch; c++) if (!chDone[c]) for (int s = 0; s < 8; s++) if (!slotUsed[s] && cost[c][s] < bv) { bv = cost[c][s]; bc = c; bs = s; }
            chDone[bc] = true; slotUsed[bs] = true; slotOf[bc] = bs;
        }
        int childAt[8]; for (int s = 0; s < 8; s++) childAt[s] = -1;
        for (int c = 0; c < nch; c++) childAt[slotOf[c]] = c;

        // quantisation frame: origin one quantum below the node's min, scale so that the node spans <= 252 quanta.
        // Every child plane then quantises to 1..254 with room for the 0.01-quantum outward slack that covers the
        // traversal's 2^-9-quantum decode error (byte_unit_float) without ever clamping at 0 / 255.
        uint8_t eb[3]; double scale[3]; float pf[3];
        for (int a = 0; a < 3; a++) {
            double ext = (double)nb.hi[a] - (double)nb.lo[a];
            int e = -126;
            if (ext > 0.0) { e = (int)std::ceil(std::log2(ext / 252.0)); while (ext / std::ldexp(1.0, e) > 252.0) e++; }
            e = std::max(-126, std::min(100, e));
            eb[a] = (uint8_t)(e + 127); scale[a] = std::ldexp(1.0, e);
            pf[a] = (float)((double)nb.lo[a] - scale[a]);
            if ((double)pf[a] > (double)nb.lo[a]) pf[a] = std::nextafterf(pf[a], -std::numeric_limits<float>::infinity());
        }
        uint8_t qlo[3][8], qhi[3][8]; uint32_t imask = 0, valid24 = 0;
        uint32_t childBase = (uint32_t)out.nodes.size(), primBase = (uint32_t)out.prims.size();
        for (int s = 0; s < 8; s++) {
            for (int a = 0; a < 3; a++) { qlo[a][s] = 255; qhi[a][s] = 0; }
            if (childAt[s] < 0) continue;
            const int c = ch[childAt[s]].b2;
            const bool leafChild = ch[childAt[s]].leaf;
            const B2Node& cn = B.b2[c];
            for (int a = 0; a < 3; a++) {
                double l = std::floor(((double)cn.box.lo[a] - (double)pf[a]) / scale[a] - 0.01);
                double h = std::ceil(((double)cn.box.hi[a] - (double)pf[a]) / scale[a] + 0.01);
                qlo[a][s] = (uint8_t)std::max(0.0, std::min(255.0, l));
                qhi[a][s] = (uint8_t)std::max(0.0, std::min(255.0, h));
            }
            if (!leafChild) {
                imask |= 1u << s;
                int wi = (int)out.nodes.size();
                out.nodes.push_back(WideNode());
                depthOf.push_back(depthOf[w.wide] + 1);
                maxDepthSeen = std::max(maxDepthSeen, depthOf[w.wide] + 1);
                queue.push_back({c, wi});
            } else {
                uint32_t unary = cn.count == 1 ? 1u : (cn.count == 2 ? 3u : 7u);
                valid24 |= unary << (3 * s);
                for (int i = 0; i < cn.count; i++) out.prims.push_back(B.prims[B.idx[cn.first + i]].rec);
            }
        }
        uint32_t imr = 0;
        for (int s = 0; s < 8; s++) if (!((imask >> s) & 1u)) imr |= 1u << (7 - s);
        uint32_t pw[3][4];   // word k of axis a = { qlo[2k], qhi[2k], qlo[2k+1], qhi[2k+1] }
        for (int a = 0; a < 3; a++) for (int k = 0; k < 4; k++) { const uint8_t b[4] = {qlo[a][2 * k], qhi[a][2 * k], qlo[a][2 * k + 1], qhi[a][2 * k + 1]}; pw[a][k] = pack4(b); }
        WideNode& wn = out.nodes[w.wide];
        wn.n0 = make_uint4(fbits(pf[0]), fbits(pf[1]), fbits(pf[2]), (uint32_t)eb[0] | ((uint32_t)eb[1] << 8) | ((uint32_t)eb[2] << 16) | (imask << 24));
        wn.n1 = make_uint4(childBase, primBase, valid24, imr << 24);
        wn.n2 = make_uint4(pw[0][0], pw[0][1], pw[0][2], pw[0][3]);
        wn.n3 = make_uint4(pw[1][0], pw[1][1], pw[1][2], pw[1][3]);
        wn.n4 = make_uint4(pw[2][0], pw[2][1], pw[2][2], pw[2][3]);
    }
    for (size_t i = 0; i < depthOf.size(); i++) if (i == 0 || depthOf[i] != depthOf[i - 1]) out.levelStart.push_back((int)i);   // breadth-first: depths never decrease
    out.levelStart.push_back((int)out.nodes.size());
    out.stats.nWideNodes = (int64_t)out.nodes.size();
    out.stats.maxDepth = std::max(1, maxDepthSeen);
    for (int a = 0; a < 3; a++) { out.stats.sceneLo[a] = B.b2[0].box.lo[a]; out.stats.sceneHi[a] = B.b2[0].box.hi[a]; }
    out.stats.depthBounded = bounded ? 1 : 0;
    if (out.stats.maxDepth <= maxDepth) break;
    if (bounded) { err = "wide BVH deeper than the traversal stack"; return false; }   // unreachable for 32-bit primitive counts
    }
    if ((int64_t)out.prims.size() != (int64_t)N) { err = "internal: primitive count mismatch"; return false; }
    return true;
}

}   // namespace rtx
