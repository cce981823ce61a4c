// rtcore.cu — sm_100a kernels and the C-ABI implementation (include/rtcore_b200.h).
//
// Kernels (one per wavefront stage; bodies in rt_wavefront.h, traversal in rt_traverse.h):
//   k_generate_primary   camera rays                         (RTUtils.cs:13-17, RTRay.cs:120-126)
//   k_extend<ANY>        persistent-thread wide-BVH traversal (replaces SceneDeviceViews.cs:30-327)
//   k_primary_finish     G-buffer + depth/objId              (RTRay.cs:90-108,188-201)
//   k_shade_first/next   material switch, ReSTIR-DI candidates, bounce, RNG (RTRay.cs:203-317,438-543)
//   k_accumulate         per-pixel sample sum, mean, float4 radiance, accumulator, PackRGBA8 (RTRay.cs:320-324,66-76)
//   k_deinterleave       multi-GPU: scatter gathered tile payloads into the full image
//
// Compiled with --fmad=false: every mul/add below is a separate IEEE operation unless written rt_fma().
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is loaded at run time (nccl_api), nothing links against it

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <tuple>
#include <utility>
#include <chrono>
#include <vector>

#include "rt_bvh.h"
#include "rt_tiles.h"
#include "rt_wavefront.h"

using namespace rtx;

// ------------------------------------------------------------------------------------------------ kernels
#ifndef RT_EXTEND_THREADS
#define RT_EXTEND_THREADS 128
#endif
#ifndef RT_EXTEND_MIN_BLOCKS
#define RT_EXTEND_MIN_BLOCKS 6   // 80 registers per lane: 24 warps per SM (sweep: 6 x 128 and 8 x 96 threads tie, 5 x 128 is 4-5 % slower)
#endif
#ifndef RT_EXTEND_BATCH
#define RT_EXTEND_BATCH 96   // ray indices a warp takes from the global queue per atomic
#endif
#ifndef RT_EXTEND_SMALL_BATCH
#define RT_EXTEND_SMALL_BATCH 32   // ... when the queue holds fewer than RT_EXTEND_SMALL_QUEUE batches per warp of the grid: the deep depths of a frame,
#endif                             // an eighth of a frame per GPU, interactive frame sizes - where whole batches per warp leave the tail of the launch to a few warps
#ifndef RT_EXTEND_SMALL_QUEUE
#define RT_EXTEND_SMALL_QUEUE 4
#endif
#ifndef RT_NODE_STEPS
#define RT_NODE_STEPS 2       // node steps per loop iteration (amortises the refill / vote / finalise overhead)
#endif
#ifndef RT_PRIM_VOTE
#define RT_PRIM_VOTE 1       // lanes that must have a primitive queued before the warp runs a primitive phase (sweep: 1 is best)
#endif

#define RT_PATH_BYTES 180   // device memory per path of a wavefront pass: state record 64 (throughput | rng, Li | flags, pending direct light, miss direction), two ray queues 2 x 32, shadow queue 32, hit record 4 + 16

struct DeviceStats {   // zeroed at the start of every rt_render
    unsigned long long raysPrimary, raysBounce, raysShadow, wideNodes, tris, spheres;
    unsigned long long raysSunProbe;    // shared sun-visibility probes traced (one per Lambert primary vertex facing the sun)
    unsigned long long shadowProbed;    // first-vertex shadow rays answered by a probe instead of their own trace
};

struct ExtendArgs {
    DeviceScene sc;
    const float4* rayO; const float4* rayD;   // origin | path slot, direction | path slot (the box-test reciprocal is derived here: 3 divisions against 32 B of DRAM traffic per ray)
    const int* count;          // number of rays in the queue (device memory: written by the producing kernel)
    int* work;                 // global fetch cursor for this launch (zeroed per frame)
    HitQueue hits;             // closest: primitive index per ray, t | bu | bv per hit
    PathState* missSt;         // closest: a bounce ray that left the scene leaves its direction in st[path slot].miss (null: primary rays)
    PathState* visSt;          // any-hit: the visibility goes to st[path slot].c.w, next to the pending contribution
    DeviceStats* stats;
    int statSlot;              // 0 primary, 1 bounce, 2 shadow, 3 sun probe
    int stackEntries;          // traversal stack entries per lane in shared memory (debug bounds checks)
};

__global__ void k_generate_primary(FrameConst fc, RayQueue q, int* countOut) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fc.npx; i += stride) generate_primary(fc, q, i);
    if (blockIdx.x == 0 && threadIdx.x == 0) *countOut = fc.npx;
}

// Persistent threads: the grid is sized to the machine, every warp pulls RT_EXTEND_BATCH consecutive ray
// indices per atomic and hands them to lanes as they go idle (ballot + popc rank), so a finished ray is
// replaced at the next step instead of idling until the slowest lane of its batch is done.
#ifndef RT_PHASE_STATS
#define RT_PHASE_STATS 0   // 1 (tuning builds, tests/gpu_phase_stats.py): with RT_FLAG_COUNTERS, lane participation of the extend kernel's phases, printed to stderr per frame
#endif
#if RT_PHASE_STATS
__device__ unsigned long long g_phase[32];
#define PHASE_ADD(i, v) do { if (COUNT && lane == 0) ph[i] += (unsigned)(v); } while (0)   // v must not contain warp-synchronous calls
#else
#define PHASE_ADD(i, v) do { } while (0)
#endif
// One queue's worth of the persistent loop.  `stack` / the hit table are set up by the kernel; a warp returns when the queue is
// exhausted and its own lanes are done - it does NOT wait for the other warps, so in the fused kernel below a warp that runs out
// of closest-hit rays goes straight on to the any-hit queue while others still finish theirs (no serialised launch tail).
template <bool ANY_HIT, bool COUNT>
__device__ __forceinline__ void extend_queue(const ExtendArgs& a, LaneStack& stack) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = (int)(threadIdx.x & 31u);
    const unsigned ltMask = (1u << lane) - 1u;
    const int n = *a.count;
    const int batch = (long long)n < (long long)gridDim.x * (RT_EXTEND_THREADS / 32) * RT_EXTEND_SMALL_QUEUE * RT_EXTEND_BATCH ? RT_EXTEND_SMALL_BATCH : RT_EXTEND_BATCH;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long* slot = a.statSlot == 0 ? &a.stats->raysPrimary : (a.statSlot == 1 ? &a.stats->raysBounce : (a.statSlot == 2 ? &a.stats->raysShadow : &a.stats->raysSunProbe));
        atomicAdd(slot, (unsigned long long)n);
    }
    if (n <= 0 || a.sc.nNodes <= 0) {
        if (!ANY_HIT) {   // empty scene: everything misses
            const int stride = gridDim.x * blockDim.x;
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
                const float4 ro = a.rayO[i], rd = a.rayD[i];
                HitRec h; h.t = 1e30f; h.prim = -1; h.bu = 0.0f; h.bv = 0.0f;
                store_closest_result(a.hits, a.missSt, i, (int)f2u(ro.w), h, mk3(rd.x, rd.y, rd.z));
            }
        } else {
            const int stride = gridDim.x * blockDim.x;
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) store_anyhit_result(a.visSt, (int)f2u(a.rayO[i].w), false);   // nothing can occlude
        }
        return;
    }
    Traversal<ANY_HIT, COUNT> tr;
    TraceCounters cnt; cnt.nodes = 0; cnt.tris = 0; cnt.spheres = 0;
    bool active = false, exhausted = false;
    int myRay = -1;
    uint32_t mySlot = 0u;            // the path slot the ray belongs to (where a miss / the visibility is settled)
    int poolNext = 0, poolEnd = 0;   // warp-uniform
#if RT_PHASE_STATS
    unsigned long long ph[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif

    for (;;) {
        // ---- refill idle lanes straight from the queue (warp-uniform control flow) ----
        for (;;) {
            const unsigned idle = __ballot_sync(FULL, !active);
            if (idle == 0u || exhausted) break;
            if (poolNext >= poolEnd) {
                int base = 0;
                if (lane == 0) base = atomicAdd(a.work, batch);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; break; }
                poolNext = base;
                poolEnd = min(base + batch, n);
            }
            const int nIdle = __popc(idle);
            const int take = min(nIdle, poolEnd - poolNext);
            if (!active) {
                const int r = __popc(idle & ltMask);
                if (r < take) {
                    myRay = poolNext + r;
                    const float4 ro = __ldcs(a.rayO + myRay), rd = __ldcs(a.rayD + myRay);
                    const f3 d = mk3(rd.x, rd.y, rd.z);
                    mySlot = f2u(ro.w);
                    tr.init(mk3(ro.x, ro.y, ro.z), d, box_idir_device(d), ANY_HIT ? 1e29f : 1e30f, stack);   // shadow tMax: RTRay.cs:623
                    active = true;
                }
            }
            poolNext += take;
            if (take == nIdle) break;
        }
        if (__ballot_sync(FULL, active) == 0u) break;
#if RT_PHASE_STATS
        if (COUNT) { const int held = __popc(__ballot_sync(FULL, active)); PHASE_ADD(0, 1); PHASE_ADD(7, held); }
#endif

        // ---- node phase: lanes without queued primitives advance by up to RT_NODE_STEPS wide nodes ----
#pragma unroll
        for (int ns = 0; ns < RT_NODE_STEPS; ns++) {
#if RT_PHASE_STATS
            if (COUNT) { const int nl = __popc(__ballot_sync(FULL, active && !tr.done && tr.can_node_step(stack))); PHASE_ADD(1 + (ns > 0), nl); PHASE_ADD(3 + (ns > 0), nl > 0); }
#endif
            if (active && !tr.done && tr.can_node_step(stack)) tr.node_step(a.sc, stack, &cnt);
        }
        // ---- primitive phase, voted warp-wide: the exact intersectors are long and divergent, so run them only when
        //      enough lanes have a primitive queued (or nobody can do node work); lanes holding primitives wait ----
        const bool wantPrim = active && !tr.done && tr.has_prims();
        const unsigned pm = __ballot_sync(FULL, wantPrim);
        if (pm != 0u) {
            const unsigned nm = __ballot_sync(FULL, active && !tr.done && tr.can_node_step(stack));
            if (__popc(pm) >= RT_PRIM_VOTE || nm == 0u) {
#if RT_PHASE_STATS
                if (COUNT) { unsigned q = wantPrim ? (unsigned)__popc(tr.tgroup.y) : 0u; for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(FULL, q, o); const int np = __popc(pm); PHASE_ADD(5, 1); PHASE_ADD(6, np); PHASE_ADD(8, q); }
#endif
                if (wantPrim) tr.prim_step(a.sc, stack, &cnt);
            }
        }
        if (active && tr.done) {
            // fire-and-forget stores, no path state is read here (same words as store_anyhit_result / store_closest_result)
            if (ANY_HIT) __stcs(&a.visSt[mySlot].c.w, tr.occluded ? 0.0f : 1.0f);   // visibility of the pending contribution: added to Li at the path's next touch
            else {
                const HitRec h = tr.result();
                const bool hit = h.t < 1e29f;
                __stcs(a.hits.prim + myRay, hit ? h.prim : -1);
                if (hit) __stcs(a.hits.tuv + myRay, make_float4(h.t, h.bu, h.bv, 0.0f));
                else if (a.missSt) __stcs(&a.missSt[mySlot].miss, make_float4(tr.d.x, tr.d.y, tr.d.z, 0.0f));   // accumulate adds throughput * sky(d)
            }
            active = false;
        }
    }
    if (COUNT) {
        unsigned nn = cnt.nodes, tt = cnt.tris, ss = cnt.spheres;
        for (int o = 16; o > 0; o >>= 1) { nn += __shfl_xor_sync(FULL, nn, o); tt += __shfl_xor_sync(FULL, tt, o); ss += __shfl_xor_sync(FULL, ss, o); }
        if (lane == 0) { atomicAdd(&a.stats->wideNodes, (unsigned long long)nn); atomicAdd(&a.stats->tris, (unsigned long long)tt); atomicAdd(&a.stats->spheres, (unsigned long long)ss); }
#if RT_PHASE_STATS
        if (lane == 0) for (int i = 0; i < 9; i++) atomicAdd(&g_phase[(ANY_HIT ? 16 : 0) + i], ph[i]);
#endif
    }
}

// dynamic shared memory of the extend kernels: hit table | traversal stacks [entries][RT_EXTEND_THREADS], entries = depth of this
// scene's wide BVH + 1 (sized by the host, so a shallow tree leaves more of the SM's 228 KB to the L1 cache that serves the node fetches)
__device__ __forceinline__ void extend_setup(LaneStack& stack, int stackEntries) {
    extern __shared__ __align__(16) uint32_t smemRaw[];
    uint32_t* hitTable = smemRaw;
    uint2* smemStack = reinterpret_cast<uint2*>(hitTable + RT_HIT_TABLE_WORDS);
    stack.smem = smemStack + threadIdx.x;
    stack.stride = RT_EXTEND_THREADS;
    stack.sp = 0;
    stack.lut = hitTable;
#if RT_DEBUG_BOUNDS
    stack.entries = stackEntries;
#endif
    for (uint32_t i = threadIdx.x; i < RT_HIT_TABLE_WORDS; i += RT_EXTEND_THREADS) hitTable[i] = hit_table_entry(i >> 8, i & 255u);
    __syncthreads();
}

template <bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(RT_EXTEND_THREADS, RT_EXTEND_MIN_BLOCKS) k_extend(const __grid_constant__ ExtendArgs a) {
    LaneStack stack;
    extend_setup(stack, a.stackEntries);
    extend_queue<ANY_HIT, COUNT>(a, stack);
}

// The closest-hit rays and the shadow rays of one depth are independent (their results meet again in the next shade / accumulate): ONE persistent launch walks both queues - the long closest-hit rays first, the any-hit rays fill the tail.
struct ExtendPairArgs { ExtendArgs closest, anyhit; };
template <bool COUNT>
__global__ void __launch_bounds__(RT_EXTEND_THREADS, RT_EXTEND_MIN_BLOCKS) k_extend_pair(const __grid_constant__ ExtendPairArgs a) {
    LaneStack stack;
    extend_setup(stack, a.closest.stackEntries);
    // (measured and rejected, round 2: for queues too small to give every lane two rays, every second warp starting with the any-hit
    // queue so that the two drain side by side - either order in the code costs the hot loop registers: spills, +4 % per frame)
    extend_queue<false, COUNT>(a.closest, stack);
    extend_queue<true, COUNT>(a.anyhit, stack);
}

__global__ void k_primary_finish(FrameConst fc, DeviceScene sc, WaveBuffers wb, RayQueue q, HitQueue hits) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fc.npx; i += stride) primary_finish(fc, sc, wb, q, hits, i);
}

__global__ void k_sun_generate(FrameConst fc, WaveBuffers wb, ShadowQueue shq, int* shCount) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fc.npx; i += stride) sun_probe_generate(fc, wb, i, shq, shCount);
}
__global__ void k_sun_store(WaveBuffers wb, ShadowQueue shq, const int* count) {
    const int n = *count, stride = gridDim.x * blockDim.x;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) sun_probe_store(wb, shq, k);
}

// Queue slots for the rays of a block's 256 vertices: ballot per warp, an 8-entry scan in shared memory, ONE atomic per queue per
// block (warp-level aggregation left the two queue counters as the hottest addresses of the frame: 33 M same-address atomics).
// Called by every thread of the block; two barriers: `sm` is one of two 16-int halves the callers alternate between, so the next
// call's counts never overwrite offsets a slow warp has not read yet.
__device__ __forceinline__ void block_push(const VertexOut& vo, int path, const RayQueue& nextQ, int* nextCount, const ShadowQueue& shq, int* shCount, int* sm /* 16 ints */) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
    const unsigned ltMask = (1u << lane) - 1u;
    const unsigned mN = __ballot_sync(FULL, vo.pushNext != 0), mS = __ballot_sync(FULL, vo.pushShadow != 0);
    if (lane == 0) { sm[warp] = __popc(mN); sm[8 + warp] = __popc(mS); }
    __syncthreads();
    if (warp == 0) {
        const int cN = lane < 8 ? sm[lane] : 0, cS = lane < 8 ? sm[8 + lane] : 0;
        int pN = cN, pS = cS;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const int a = __shfl_up_sync(FULL, pN, o), b = __shfl_up_sync(FULL, pS, o); if (lane >= o) { pN += a; pS += b; } }
        int baseN = 0, baseS = 0;
        if (lane == 7) { if (pN > 0) baseN = atomicAdd(nextCount, pN); if (pS > 0) baseS = atomicAdd(shCount, pS); }
        baseN = __shfl_sync(FULL, baseN, 7); baseS = __shfl_sync(FULL, baseS, 7);
        if (lane < 8) { sm[lane] = baseN + pN - cN; sm[8 + lane] = baseS + pS - cS; }
    }
    __syncthreads();
    if (vo.pushNext) write_ray(nextQ.o, nextQ.d, sm[warp] + __popc(mN & ltMask), vo.next, path);
    if (vo.pushShadow) write_ray(shq.o, shq.d, sm[8 + warp] + __popc(mS & ltMask), vo.shadow, path);
}

template <bool REUSE, bool FAST>
__global__ void __launch_bounds__(256) k_shade_first(FrameConst fc, WaveBuffers wb, int sampleBase, int nPaths, RayQueue nextQ, int* nextCount, ShadowQueue shq, int* shCount, DeviceStats* stats) {
    __shared__ int smPush[32];
    unsigned probed = 0, flip = 0;
    for (long long base = (long long)blockIdx.x * 256; base < nPaths; base += (long long)gridDim.x * 256) {
        const int j = (int)(base + threadIdx.x);
        VertexOut vo; vo.pushNext = 0; vo.pushShadow = 0;
        if (j < nPaths) shade_first<REUSE, FAST>(fc, wb, sampleBase, j, nextQ, nextCount, shq, shCount, &probed, &vo);
        block_push(vo, j, nextQ, nextCount, shq, shCount, smPush + 16 * (flip++ & 1u));
    }
    for (int o = 16; o > 0; o >>= 1) probed += __shfl_xor_sync(0xFFFFFFFFu, probed, o);
    if ((threadIdx.x & 31u) == 0u && probed != 0u) atomicAdd(&stats->shadowProbed, (unsigned long long)probed);
}

// Most bounce rays of an open scene leave it; those never reach this kernel's shading (the extend kernel leaves their direction
// with the path and accumulate adds the sky).  Each block takes chunks of RT_SHADE_CHUNK consecutive rays and makes two passes:
// (1) a scan of the 4-byte primitive indices (sixteen per thread as four independent 128-bit loads): a hit is classified by the
// material of what it hit and appended to a list in shared memory - Lambert vertices (long: nine ReSTIR candidates) from the
// front, mirror / glass vertices (short) from the back - with ballot + one shared-memory atomic per warp: no global atomics;
// (2) the two ends of the list are shaded one after the other, each with every lane busy on the same kind of vertex.
#ifndef RT_SHADE_CHUNK
#define RT_SHADE_CHUNK 4096
#endif
__device__ __forceinline__ bool hit_is_specular(const DeviceScene& sc, int prim) {   // the "shade" TraceClosest will report (SceneDeviceViews.cs:61,158)
    const uint32_t meta = __float_as_uint(__ldg(&sc.prims[prim].q2.w));
    const int primId = (int)__float_as_uint(__ldg(&sc.prims[prim].q0.w));
    int shade = RT_SHADING_LAMBERT;
    if (meta & PRIM_SPHERE) shade = sc.spheres[primId].shading;
    else if (sc.triMaterials) shade = sc.materials[sc.triMatIndex[primId]].Shading;
    return shade == RT_SHADING_MIRROR || shade == RT_SHADING_GLASS;
}
template <bool REUSE, bool FAST, int CHUNK>
__device__ __forceinline__ void shade_next_chunks(const FrameConst& fc, const DeviceScene& sc, const WaveBuffers& wb, int depth, const RayQueue& curQ, const HitQueue& hits, int n,
                                                  const RayQueue& nextQ, int* nextCount, const ShadowQueue& shq, int* shCount, int* hitCount, int* list, int* nFrontP, int* nBackP, int* smPush) {
    constexpr int SUB = CHUNK;
    unsigned flip = 0;
    const int nChunks = (n + SUB - 1) / SUB;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = (int)(threadIdx.x & 31u);
    const unsigned ltMask = (1u << lane) - 1u;
    static_assert(CHUNK % 1024 == 0 && RT_SHADE_CHUNK % CHUNK == 0, "scan: CHUNK / 1024 int4 loads per thread; the index array is padded to RT_SHADE_CHUNK");
    constexpr int LOADS = SUB / 1024;
    for (int ch = blockIdx.x; ch < nChunks; ch += gridDim.x) {
        if (threadIdx.x == 0) { *nFrontP = 0; *nBackP = 0; }
        __syncthreads();
        const int base = ch * SUB;
        // the primitive-index array is padded to a multiple of the chunk (ensure_frame_buffers), so whole-int4 loads stay in bounds
        const int4* src = reinterpret_cast<const int4*>(hits.prim + base);
        int4 pv[LOADS];
#pragma unroll
        for (int it = 0; it < LOADS; it++) pv[it] = __ldcs(src + it * 256 + threadIdx.x);
#pragma unroll
        for (int it = 0; it < LOADS; it++) {
            const int k0 = base + (it * 256 + (int)threadIdx.x) * 4;
            const int pr[4] = {pv[it].x, pv[it].y, pv[it].z, pv[it].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = k0 + e;
                const bool hit = k < n && pr[e] >= 0;
                const bool spec = hit && depth < fc.maxDepth && hit_is_specular(sc, pr[e]);
                const unsigned mF = __ballot_sync(FULL, hit && !spec), mB = __ballot_sync(FULL, spec);
                if (mF != 0u) {
                    int b = 0;
                    if (lane == 0) b = atomicAdd(nFrontP, __popc(mF));
                    b = __shfl_sync(FULL, b, 0);
                    if (hit && !spec) list[b + __popc(mF & ltMask)] = k;
                }
                if (mB != 0u) {
                    int b = 0;
                    if (lane == 0) b = atomicAdd(nBackP, __popc(mB));
                    b = __shfl_sync(FULL, b, 0);
                    if (spec) list[SUB - 1 - (b + __popc(mB & ltMask))] = k;
                }
            }
        }
        __syncthreads();
        const int nf = *nFrontP, nb = *nBackP;
        if (hitCount && threadIdx.x == 0) atomicAdd(hitCount, nf + nb);
        const int rf = (nf + 31) & ~31;   // a warp never mixes the two kinds: the specular part starts on a warp boundary
        // (measured and rejected, round 2: pulling the records of the hits one / two iterations ahead towards L2 with prefetch.global.L2,
        // here and for the G-buffer lines of k_shade_first: +8 ms per C4 frame)
        auto entry = [&](int i) -> int { return i < rf ? (i < nf ? list[i] : -1) : (i < rf + nb ? list[SUB - 1 - (i - rf)] : -1); };
        for (int i0 = 0; i0 < rf + nb; i0 += 256) {   // uniform trip count: block_push has barriers
            const int k = entry(i0 + (int)threadIdx.x);
            VertexOut vo; vo.pushNext = 0; vo.pushShadow = 0;
            int path = 0;
            if (k >= 0) shade_next<REUSE, FAST>(fc, sc, wb, depth, curQ, hits, k, nextQ, nextCount, shq, shCount, &vo, &path);
            if (depth < fc.maxDepth) block_push(vo, path, nextQ, nextCount, shq, shCount, smPush + 16 * (flip++ & 1u));   // the last depth queues nothing
        }
        __syncthreads();
    }
}
template <bool REUSE, bool FAST, int CHUNK>
__global__ void __launch_bounds__(256, REUSE ? 2 : 4) k_shade_next(FrameConst fc, DeviceScene sc, WaveBuffers wb, int depth, RayQueue curQ, HitQueue hits, const int* curCount,
                                                   RayQueue nextQ, int* nextCount, ShadowQueue shq, int* shCount, int* hitCount /* null unless RT_DUMP_COUNTERS */, int fineBelow, int fineMode) {
    __shared__ int list[CHUNK];
    __shared__ int nFront, nBack;
    __shared__ int smPush[32];
    const int n = *curCount;
    // The chunk size a queue is scanned with is chosen on the DEVICE, by its length: big wavefronts launch the 4096-ray and the
    // 1024-ray instantiation back to back and exactly one of them runs (fineMode 1: only queues of at least fineBelow rays, 2: only
    // shorter ones, 0: always).  At 4096 rays per chunk the 500 K-ray queue of a deep depth is 122 blocks' worth on a machine that
    // holds 592; one kernel with both chunk sizes costs the common case registers (spills), an idle launch costs ~3 us.
    if ((fineMode == 1 && n < fineBelow) || (fineMode == 2 && n >= fineBelow)) return;
    shade_next_chunks<REUSE, FAST, CHUNK>(fc, sc, wb, depth, curQ, hits, n, nextQ, nextCount, shq, shCount, hitCount, list, &nFront, &nBack, smPush);
}

__global__ void k_accumulate(FrameConst fc, WaveBuffers wb, int sampleBase, int nSamples, int last) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fc.npx; i += stride) accumulate(fc, wb, sampleBase, nSamples, last != 0, i);
}

__global__ void k_copy_color(const int* src, int* dst, const int* pixelMap, int npx) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) { const int p = pixelMap[i]; dst[p] = src[p]; }
}

// scatter a per-owned-pixel array to global pixel order (read-back of G-buffer / AOV taps)
template <typename T> __global__ void k_scatter(const T* src, T* dst, const int* pixelMap, int npx) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) dst[pixelMap[i]] = src[i];
}
__global__ void k_scatter_f4_to_f3(const float4* src, float* dst, const int* pixelMap, int npx) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) { const float4 v = src[i]; const size_t p = (size_t)pixelMap[i] * 3; dst[p] = v.x; dst[p + 1] = v.y; dst[p + 2] = v.z; }
}
__global__ void k_scatter_f4_w(const float4* src, int* dst, const int* pixelMap, int npx) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) dst[pixelMap[i]] = __float_as_int(src[i].w);
}

// ---- present chain (RTRenderer.cs:281-320, RTTaa.cs:117-179): one thread per OUTPUT pixel, 4 B written (+ 8 B of history) ----
__global__ void k_bilinear_upsample(const int* src, int srcW, int srcH, int* dst, int dstW, int dstH) {
    const int n = dstW * dstH, stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = bilinear_upsample_pixel(src, srcW, srcH, dstW, dstH, i);
}
__global__ void __launch_bounds__(256) k_taa_resolve(TaaConst p, const int* lowColor, const int* lowObj, int* histColor, int* histObj, int* out) {
    __shared__ float lut[256];   // the sRGB decode has 256 possible inputs per channel: tabulate it once per block (same bits as evaluating it in place)
    lut[threadIdx.x] = srgb_to_linear_u8((int)threadIdx.x);
    __syncthreads();
    const int n = p.outW * p.outH, stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int obj;
        const int c = taa_resolve_pixel(p, lut, lowColor, lowObj, histColor[i], histObj[i], i, &obj);
        out[i] = c; histColor[i] = c; histObj[i] = obj;
    }
}

// reservoir planes (L|pdf, wi|w, wSum|m|lightId) -> the reference's 44-byte Reservoir records (read-back for parity)
__global__ void k_pack_reservoirs(const float4* res, int g, float* out) {
    const int stride = gridDim.x * blockDim.x;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g; p += stride) {
        const float4 a = res[p], b = res[(size_t)g + p], c = res[2 * (size_t)g + p];
        float* o = out + (size_t)p * 11;
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = a.w; o[7] = b.w; o[8] = c.x; o[9] = c.y; o[10] = c.z;
    }
}

// multi-GPU finish: payload[k] (float4 Lout of the k-th owned pixel of some rank) -> full image
__global__ void k_deinterleave(const float4* payload, const int* pixelMap, int npx, float4* outRadiance, int* outRgba8) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) {
        const float4 v = __ldcs(payload + i);
        const int p = pixelMap[i];
        if (outRadiance) outRadiance[p] = v;
        if (outRgba8) outRgba8[p] = pack_rgba8(mk3(v.x, v.y, v.z));
    }
}

// rt_gather_frame, receive side: the tile-compacted payloads of one or more ranks (concatenated in rank order, the way their
// owned-pixel lists are concatenated in `map`) -> the gathered full image.  De-interleave and PackRGBA8 in one pass.
__global__ void __launch_bounds__(256) k_gather_scatter(const float4* rad, const int* rgba, const uint2* aux, const int* map, int n,
                                                        float4* outRadiance, int* outRgba8, float* outDepth, int* outObjId) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int p = map[i];
        if (rad) { const float4 v = __ldcs(rad + i); outRadiance[p] = v; outRgba8[p] = pack_rgba8(mk3(v.x, v.y, v.z)); }
        else if (rgba) outRgba8[p] = __ldcs(rgba + i);
        if (aux) { const uint2 a = __ldcs(aux + i); outDepth[p] = __uint_as_float(a.x); outObjId[p] = (int)a.y; }
    }
}

// ReSTIR reuse across a tile partition: this rank's reservoirs (global pixel order) -> its segment of the exchange buffer
// (owned-pixel lists of all ranks concatenated in rank order), and the other ranks' segments back into global pixel order.
__global__ void k_res_pack(const float4* cur, size_t g, const int* pixelMap, int npx, size_t start, float4* pack) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) {
        const size_t p = (size_t)pixelMap[i], q = start + (size_t)i;
        pack[q] = cur[p]; pack[g + q] = cur[g + p]; pack[2 * g + q] = cur[2 * g + p];
    }
}
__global__ void k_res_unpack(const float4* pack, size_t g, const int* allMap, size_t ownStart, size_t ownEnd, float4* cur) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < g; q += stride) {
        if (q >= ownStart && q < ownEnd) continue;
        const size_t p = (size_t)allMap[q];
        cur[p] = pack[q]; cur[g + p] = pack[g + q]; cur[2 * g + p] = pack[2 * g + q];
    }
}

// ------------------------------------------------------------------------------------------------ host side
#include "rt_build.h"   // device-side refit and build of the wide BVH

static thread_local std::string g_lastError;
static int fail(int code, const std::string& msg) { g_lastError = msg; return code; }
#define CUDA_TRY(expr)                                                                                                 \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess) return fail(_e == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_CUDA,       \
                                           std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
    } while (0)

// ------------------------------------------------------------------------------------------------ frame recorder (RT_FLAG_FRAME_GRAPH)
// Every launch of a frame goes through a FrameRecorder.  Direct: cudaLaunchKernel on the stream (the default).  Build: the same
// sequence becomes an explicit CUDA graph (a chain of kernel / memset / memcpy nodes) that is instantiated once per configuration.
// Update: later frames of that configuration only refresh the kernel nodes' parameters (camera, frame index, buffer parity ...)
// with cudaGraphExecKernelNodeSetParams and replay the graph with ONE launch: ~1 us of host work per kernel instead of a launch,
// and no launch gaps on the device - what the reference's interactive regime (858x482, 2 spp) is bound by.
struct FrameRecorder {
    enum Mode { Direct, Build, Update };
    Mode mode = Direct;
    cudaStream_t st = nullptr;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    std::vector<cudaGraphNode_t>* kernelNodes = nullptr; std::vector<const void*>* kernelFuncs = nullptr;
    size_t idx = 0; cudaGraphNode_t last = nullptr; bool haveLast = false, mismatch = false;
    cudaError_t err = cudaSuccess; const char* where = "";
    const void* l2Base = nullptr; size_t l2Bytes = 0; float l2Ratio = 1.0f;   // the persisting window of the BVH: graph kernel nodes do not inherit the stream's

    void chain(cudaGraphNode_t n) { last = n; haveLast = true; }
    template <typename... KArgs, typename... Args>
    void launch(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
        static_assert(sizeof...(KArgs) == sizeof...(Args), "kernel argument count");
        if (err != cudaSuccess || mismatch) return;
        std::tuple<typename std::decay<KArgs>::type...> vals{static_cast<typename std::decay<KArgs>::type>(args)...};
        void* ptrs[sizeof...(KArgs) > 0 ? sizeof...(KArgs) : 1];
        fill(ptrs, vals, std::index_sequence_for<KArgs...>{});
        if (mode == Direct) { err = cudaLaunchKernel((const void*)k, grid, block, ptrs, smem, st); where = "cudaLaunchKernel"; return; }
        cudaKernelNodeParams p; memset(&p, 0, sizeof(p));
        p.func = (void*)k; p.gridDim = grid; p.blockDim = block; p.sharedMemBytes = (unsigned)smem; p.kernelParams = ptrs; p.extra = nullptr;
        if (mode == Build) {
            cudaGraphNode_t n;
            err = cudaGraphAddKernelNode(&n, graph, haveLast ? &last : nullptr, haveLast ? 1 : 0, &p); where = "cudaGraphAddKernelNode";
            if (err == cudaSuccess) {
                kernelNodes->push_back(n); kernelFuncs->push_back((const void*)k); chain(n);
                if (l2Bytes > 0) {
                    cudaKernelNodeAttrValue av; memset(&av, 0, sizeof(av));
                    av.accessPolicyWindow.base_ptr = const_cast<void*>(l2Base); av.accessPolicyWindow.num_bytes = l2Bytes; av.accessPolicyWindow.hitRatio = l2Ratio;
                    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    if (cudaGraphKernelNodeSetAttribute(n, cudaKernelNodeAttributeAccessPolicyWindow, &av) != cudaSuccess) (void)cudaGetLastError();
                }
            }
        } else {
            if (idx >= kernelNodes->size() || (*kernelFuncs)[idx] != (const void*)k) { mismatch = true; return; }   // another sequence: rebuild
            err = cudaGraphExecKernelNodeSetParams(exec, (*kernelNodes)[idx++], &p); where = "cudaGraphExecKernelNodeSetParams";
        }
    }
    template <typename Tuple, size_t... I> static void fill(void** ptrs, Tuple& t, std::index_sequence<I...>) { (void)ptrs; (void)t; int dummy[] = {0, ((ptrs[I] = (void*)&std::get<I>(t)), 0)...}; (void)dummy; }
    void memset32(void* dst, size_t bytes) {   // zero `bytes` (a multiple of 4); static across the frames of a configuration (same pointer, same size)
        if (err != cudaSuccess || mismatch) return;
        if (mode == Direct) { err = cudaMemsetAsync(dst, 0, bytes, st); return; }
        if (mode == Update) return;
        cudaMemsetParams mp; memset(&mp, 0, sizeof(mp));
        mp.dst = dst; mp.value = 0; mp.elementSize = 4; mp.width = bytes / 4; mp.height = 1; mp.pitch = 0;
        cudaGraphNode_t n;
        err = cudaGraphAddMemsetNode(&n, graph, haveLast ? &last : nullptr, haveLast ? 1 : 0, &mp); where = "cudaGraphAddMemsetNode";
        if (err == cudaSuccess) chain(n);
    }
    void memcpy_d2h(void* dstHost, const void* srcDev, size_t bytes) {
        if (err != cudaSuccess || mismatch) return;
        if (mode == Direct) { err = cudaMemcpyAsync(dstHost, srcDev, bytes, cudaMemcpyDeviceToHost, st); return; }
        if (mode == Update) return;
        cudaGraphNode_t n;
        err = cudaGraphAddMemcpyNode1D(&n, graph, haveLast ? &last : nullptr, haveLast ? 1 : 0, dstHost, srcDev, bytes, cudaMemcpyDeviceToHost); where = "cudaGraphAddMemcpyNode1D";
        if (err == cudaSuccess) chain(n);
    }
};
// what makes two frames "the same configuration" for graph replay: everything that decides WHICH kernels run, their grids' upper
// bounds and the buffers the static nodes (memsets, the statistics copy) touch
struct FrameKey {
    int width, height, spp, S, maxDepth, worldSize, rank, tileSize, npx; uint32_t structuralFlags; int reuse, sunProbe, extColor, pad;
    const void *counters, *dstats, *hstats, *bvh, *stream; size_t nCounters, pathCap; unsigned long long sceneVersion;
    bool operator==(const FrameKey& o) const { return memcmp(this, &o, sizeof(FrameKey)) == 0; }
};

template <typename T> struct DevBuf {   // owning device allocation (freed with its owner: the context, or a scope)
    T* p = nullptr; size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t ensure(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; n = 0; }
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// Device builder (rt_build.h): scratch that survives between commits - a scene that is rebuilt every few frames should not pay
// for device allocations and page-locking each time - and the page-locked staging arrays the host primitive stage writes into.
struct BuildScratch {
    DevBuf<PrimRec> primsIn, prims; DevBuf<WideNode> nodes; DevBuf<float4> primBox, boxes; DevBuf<uint64_t> keys; DevBuf<int> vals, ints; DevBuf<unsigned char> sortTmp;
    PrimRec* stagePrims = nullptr; float4* stageBoxes = nullptr; size_t stageCap = 0;
    int* hostInts = nullptr;   // page-locked: small read-backs
    PrimSink sink;
    cudaStream_t stream = nullptr; bool recordsQueued = false;   // the records' upload starts while the host still pads the boxes
    bool reserve(size_t n) {
        if (!hostInts && cudaHostAlloc(reinterpret_cast<void**>(&hostInts), 16 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) { hostInts = nullptr; return false; }
        if (primsIn.ensure(n) != cudaSuccess) { cudaGetLastError(); return false; }
        if (n <= stageCap) return true;
        release_stage();
        const size_t cap = n + n / 8;
        if (cudaHostAlloc(reinterpret_cast<void**>(&stagePrims), cap * sizeof(PrimRec), cudaHostAllocDefault) != cudaSuccess) { stagePrims = nullptr; cudaGetLastError(); return false; }
        if (cudaHostAlloc(reinterpret_cast<void**>(&stageBoxes), 2 * cap * sizeof(float4), cudaHostAllocDefault) != cudaSuccess) { stageBoxes = nullptr; release_stage(); cudaGetLastError(); return false; }
        stageCap = cap; sink.prims = stagePrims; sink.boxes = stageBoxes;
        return true;
    }
    void release_stage() { if (stagePrims) cudaFreeHost(stagePrims); if (stageBoxes) cudaFreeHost(stageBoxes); stagePrims = nullptr; stageBoxes = nullptr; stageCap = 0; sink.prims = nullptr; sink.boxes = nullptr; }
    void release() {
        release_stage(); if (hostInts) cudaFreeHost(hostInts); hostInts = nullptr;
        primsIn.release(); prims.release(); nodes.release(); primBox.release(); boxes.release(); keys.release(); vals.release(); ints.release(); sortTmp.release();
    }
    ~BuildScratch() { release(); }
};

struct rt_ctx {
    BuildScratch build;
    int device = 0;
    int smCount = 148;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    bool rendered = false, hasScene = false;

    // scene
    DevBuf<unsigned char> bvhBlob;   // wide nodes then primitive records in ONE allocation: a single L2 persisting window covers both
    DevBuf<RtInstanceRecord> instances; DevBuf<RtSphere> spheres;
    DevBuf<RtFloat2> texcoords; DevBuf<RtMeshTriUV> triUVs; DevBuf<int32_t> triMat; DevBuf<RtMaterialRecord> materials;
    DevBuf<RtRGBA32> texels; DevBuf<RtTexInfo> texInfos;
    // refit (rt_scene_refit): the mesh topology, the level ranges of the breadth-first wide BVH, per-instance box transforms, scratch
    DevBuf<RtMeshTri> meshTris; int64_t nMeshPositions = 0, nMeshTris = 0;
    std::vector<int> levelStart; DevBuf<double> instBoxXf;
    DevBuf<RtFloat3> refitPos; DevBuf<float4> refitPrimBox, refitNodeBox; DevBuf<unsigned> refitAbs;
    DeviceScene ds;
    HostBvhStats bvhStats; size_t bvhBytes = 0;

    // frame geometry
    int width = 0, height = 0, tileSize = 0, rank = 0, worldSize = 1, npx = 0, spp = 1;
    bool aovs = false;
    DevBuf<int> pixelMap, invPixelMap;
    // ReSTIR reservoirs: A/B per global pixel (3 float4 planes each, zero-initialised: frame 0 imports nothing), per-path staging
    DevBuf<float4> resAB[2]; DevBuf<float4> resPath; size_t resPathCap = 0; int resLastWritten = -1; bool resValid[2] = {false, false}; int resFrame[2] = {0, 0};
    // per owned pixel
    DevBuf<float4> gbPosHit, gbNrmMat, gbAlbObj, lframe; DevBuf<int> primId, instId; DevBuf<float> primaryT;
    // gather payloads per owned pixel, double-buffered: frame k + 1 renders into one set while the gather of frame k still reads the other
    DevBuf<float4> tileRadiance[2]; DevBuf<uint2> tileAux[2]; DevBuf<int> tileRgba[2]; int tileBuf = 0;
    // per global pixel
    DevBuf<int> rgba8, objId; DevBuf<float> depth; DevBuf<float4> radiance, accum;
    // per path
    size_t pathCap = 0;
    DevBuf<PathState> pathState; DevBuf<float4> qO[2], qD[2], shO, shD, hitTuv; DevBuf<int> hitPrim; DevBuf<uint32_t> pathHash;
    // AOV outputs
    DevBuf<uint8_t> segOut, termOut; DevBuf<uint32_t> hashOut;
    // scratch for scattered read-backs
    DevBuf<float> scratch;
    // counters
    DevBuf<int> counters; DevBuf<DeviceStats> dstats;
    DeviceStats* hstats = nullptr;   // page-locked: the end-of-frame copy of the device counters must not make rt_render wait for the frame
    unsigned long long launches = 0;
    // kernel timing (extend kernels)
    std::vector<cudaEvent_t> traceEvents; size_t traceEventsUsed = 0; bool timeKernels = false;
    int* extColor = nullptr; size_t extColorBytes = 0;
    // present chain: own output buffer, TAA history (RTTaa._historyColor / _historyObjId), where the last present went
    DevBuf<int> presentBuf, taaHistColor, taaHistObj; int presentW = 0, presentH = 0; bool taaHistoryValid = false; const int* presentPtr = nullptr; bool presentOnComm = false;
    // multi-GPU finish: cached owned-pixel lists of every rank
    DevBuf<int> deintMap; std::vector<int64_t> deintStart; int deintW = 0, deintH = 0, deintT = 0, deintWorld = 0;
    int extendBlocks = 0;
    size_t memTotal = 0;
    size_t l2PersistMax = 0, l2WindowMax = 0, l2Persist = 0, l2Window = 0; cudaStream_t l2WindowStream = nullptr;
    size_t extendSmem = 0; int stackEntries = 0;
    // multi-GPU (rt_comm_init / rt_gather_frame): one NCCL communicator, its own stream, the gathered image on the root
    ncclComm_t comm = nullptr; int commRank = 0, commWorld = 1;
    cudaStream_t commStream = nullptr; cudaEvent_t evTileReady[2] = {nullptr, nullptr}, evGatherDone[2] = {nullptr, nullptr}, evAuxReady[2] = {nullptr, nullptr}, evGatherStart = nullptr, evGatherStop = nullptr;
    bool auxReady[2] = {false, false};   // evAuxReady[b] was recorded behind the primary pass of the frame in payload buffer b: its depth | objectId payload can leave before the frame ends
    bool gatherPending[2] = {false, false}, gatherTimed = false;
    DevBuf<unsigned char> gatherStage; DevBuf<int> gRgba8, gObjId; DevBuf<float> gDepth; DevBuf<float4> gRadiance;
    bool gatheredValid = false; uint32_t gatheredWhat = 0; int gatheredW = 0, gatheredH = 0;
    // ReSTIR reuse across the partition: every rank's G-buffer (exchanged after the primary pass) and the reservoir exchange buffer
    DevBuf<float4> gbAll, resPack; std::vector<int> deintHost;
    const float4 *gbPosPtr = nullptr, *gbNrmPtr = nullptr, *gbAlbPtr = nullptr;   // where the last frame's (own) G-buffer lives
    // rt_bind_readback: page-locked host targets for RGBA8 / depth / objectId, filled by every frame on a copy stream
    void* rbHost[3] = {nullptr, nullptr, nullptr}; size_t rbBytes[3] = {0, 0, 0}; cudaStream_t copyStream = nullptr; cudaEvent_t evPrimaryDone = nullptr, evCopyDone = nullptr; bool copyPending = false;
    cudaGraph_t frameGraph = nullptr;   // kept alive: the node handles used for per-frame parameter updates belong to it
    cudaGraphExec_t frameGraphExec = nullptr; unsigned long long frameGraphBuilds = 0, sceneVersion = 0;   // RT_FLAG_FRAME_GRAPH
    FrameKey frameKey; std::vector<cudaGraphNode_t> frameNodes; std::vector<const void*> frameFuncs;
    bool envNoL2Persist = false, envNoSunProbe = false; long long envPathsPerPass = 0; int envDeviceTree = 0 /* 0 = the better of the two, 1 = radix, 2 = ploc */; bool envBuildTiming = false, envDumpCounters = false; float envCPrim = 0.0f; int ctrStride = 0, ctrHeader = 0, ctrDepths = 0, ctrPasses = 0;   // developer knobs, read once in rt_create
};

template <typename T> static cudaError_t upload_or_one(DevBuf<T>& dst, const T* src, int64_t n, cudaStream_t st, int* lenOut) {
    // AllocateOrEmpty (Scene.cs:370-377): an empty array becomes one zeroed element
    size_t cnt = (src && n > 0) ? (size_t)n : 1;
    cudaError_t e = dst.ensure(cnt);
    if (e != cudaSuccess) return e;
    if (src && n > 0) e = cudaMemcpyAsync(dst.p, src, cnt * sizeof(T), cudaMemcpyHostToDevice, st);
    else e = cudaMemsetAsync(dst.p, 0, sizeof(T), st);
    if (lenOut) *lenOut = (int)cnt;
    return e;
}

static cudaError_t trace_event(rt_ctx* c) {
    if (!c->timeKernels) return cudaSuccess;
    if (c->traceEventsUsed == c->traceEvents.size()) { cudaEvent_t ev; cudaError_t e = cudaEventCreate(&ev); if (e != cudaSuccess) return e; c->traceEvents.push_back(ev); }
    return cudaEventRecord(c->traceEvents[c->traceEventsUsed++], c->stream);
}

template <bool ANY> static cudaError_t launch_extend(rt_ctx* c, FrameRecorder& rec, const ExtendArgs& a0, bool count) {
    ExtendArgs a = a0; a.stackEntries = c->stackEntries;
    cudaError_t e = trace_event(c);
    if (e != cudaSuccess) return e;
    if (count) rec.launch(k_extend<ANY, true>, dim3(c->extendBlocks), dim3(RT_EXTEND_THREADS), c->extendSmem, a);
    else rec.launch(k_extend<ANY, false>, dim3(c->extendBlocks), dim3(RT_EXTEND_THREADS), c->extendSmem, a);
    c->launches++;
    if (rec.err != cudaSuccess) { fprintf(stderr, "rtcore_b200: %s failed for an extend kernel (grid %d, smem %zu): %s\n", rec.where, c->extendBlocks, c->extendSmem, cudaGetErrorString(rec.err)); return rec.err; }
    return trace_event(c);
}

static cudaError_t launch_extend_pair(rt_ctx* c, FrameRecorder& rec, const ExtendArgs& closest, const ExtendArgs& anyhit, bool count) {
    ExtendPairArgs a; a.closest = closest; a.anyhit = anyhit; a.closest.stackEntries = a.anyhit.stackEntries = c->stackEntries;
    cudaError_t e = trace_event(c);
    if (e != cudaSuccess) return e;
    if (count) rec.launch(k_extend_pair<true>, dim3(c->extendBlocks), dim3(RT_EXTEND_THREADS), c->extendSmem, a);
    else rec.launch(k_extend_pair<false>, dim3(c->extendBlocks), dim3(RT_EXTEND_THREADS), c->extendSmem, a);
    c->launches++;
    if (rec.err != cudaSuccess) { fprintf(stderr, "rtcore_b200: %s failed for an extend kernel (grid %d, smem %zu): %s\n", rec.where, c->extendBlocks, c->extendSmem, cudaGetErrorString(rec.err)); return rec.err; }
    return trace_event(c);
}

// persistent grid of the extend kernels: shared memory for a traversal stack of `entries` per lane, blocks = occupancy x SMs
static int size_extend_launch(rt_ctx* c, int entries) {
    entries = std::max(2, std::min(entries, RT_STACK_ENTRIES));
    c->stackEntries = entries;
    c->extendSmem = (size_t)RT_HIT_TABLE_WORDS * sizeof(uint32_t) + (size_t)entries * RT_EXTEND_THREADS * sizeof(uint2);
    if (c->extendSmem > 48u * 1024u) {   // big blocks (tuning builds) need the opt-in
        CUDA_TRY(cudaFuncSetAttribute(k_extend<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->extendSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_extend<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->extendSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_extend<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->extendSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_extend<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->extendSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_extend_pair<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->extendSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_extend_pair<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->extendSmem));
    }
    int perSm = 0, perSmPair = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_extend<false, false>, RT_EXTEND_THREADS, c->extendSmem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmPair, k_extend_pair<false>, RT_EXTEND_THREADS, c->extendSmem));
    perSm = std::min(perSm, perSmPair);
    if (perSm < 1) perSm = 1;
    c->extendBlocks = perSm * c->smCount;
    return RT_OK;
}

// RT_BUILD_TIMING=1: phases of a scene commit on stderr (each mark synchronises the stream, so only the tuning runs pay for it)
struct BuildTimer {
    bool on; cudaStream_t st; std::chrono::steady_clock::time_point t0; std::string line;
    BuildTimer(bool on_, cudaStream_t st_) : on(on_), st(st_), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        char buf[96]; snprintf(buf, sizeof(buf), " %s %.2f ms |", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        line += buf; t0 = t1;
    }
    void print(const char* head) { if (on) fprintf(stderr, "rtcore_b200 %s:%s\n", head, line.c_str()); }
};

// Device-side build (rt_build.h) of the tree over the primitive records the host stage left in the staging arrays: fills
// bs.nodes / bs.prims, the level ranges and the depth.  Returns a cudaError_t; *tooDeep when the tree would not fit the traversal
// stack.  Two binary trees are built over the same Morton order - the radix tree (spatial-median splits: the better one for
// regular geometry such as a height field) and PLOC (agglomerative: the better one for uneven density) - and the one whose
// SAH-optimal collapse costs less, C(root, 1) of the dynamic program both carry, is collapsed (RT_DEVICE_TREE=radix|ploc forces one).
static cudaError_t build_on_device(rt_ctx* c, const HostBvh& hb, std::vector<int>& levelStart, int* nNodesOut, bool* tooDeep) {
    const int n = (int)hb.stats.nPrims;
    BuildScratch& bs = c->build;
    cudaStream_t st = c->stream;
    cudaError_t e;
#define BTRY(x) do { e = (x); if (e != cudaSuccess) return e; } while (0)
    BuildTimer tm(c->envBuildTiming, st);
    const bool wantRadix = c->envDeviceTree != 2, wantPloc = c->envDeviceTree != 1;
    const int tiles = (n + RT_PLOC_TILE - 1) / RT_PLOC_TILE;
    const int maxLevels = RT_STACK_ENTRIES - 2;
    BTRY(bs.primsIn.ensure(n)); BTRY(bs.primBox.ensure(2 * (size_t)n)); BTRY(bs.keys.ensure(2 * (size_t)n)); BTRY(bs.vals.ensure(2 * (size_t)n));
    BTRY(bs.ints.ensure(10 * (size_t)n + 2 * (size_t)tiles + 16 + 2 * (size_t)maxLevels)); BTRY(bs.boxes.ensure(12 * (size_t)n)); BTRY(bs.nodes.ensure(n)); BTRY(bs.prims.ensure(n));
    tm.mark("scratch");
    if (!bs.recordsQueued) BTRY(cudaMemcpyAsync(bs.primsIn.p, bs.stagePrims, (size_t)n * sizeof(PrimRec), cudaMemcpyHostToDevice, st));
    BTRY(cudaMemcpyAsync(bs.primBox.p, bs.stageBoxes, 2 * (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
    tm.mark("h2d records + boxes");
    const float3 lo = make_float3(hb.stats.sceneLo[0], hb.stats.sceneLo[1], hb.stats.sceneLo[2]);
    // cubic Morton cells (one scale for the three axes): a flat scene spends no high bits on its thin axis
    const float ext = fmaxf(fmaxf(hb.stats.sceneHi[0] - lo.x, hb.stats.sceneHi[1] - lo.y), fmaxf(hb.stats.sceneHi[2] - lo.z, 1e-30f));
    const float3 inv = make_float3(1.0f / ext, 1.0f / ext, 1.0f / ext);
    uint64_t* keys = bs.keys.p; uint64_t* keys2 = bs.keys.p + n; int* vals = bs.vals.p; int* vals2 = bs.vals.p + n;
    k_lbvh_morton<<<(n + 255) / 256, 256, 0, st>>>(bs.primBox.p, n, lo, inv, keys, vals);
    size_t tmpBytes = 0;
    BTRY(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keys2, vals, vals2, n, 0, 63, st));
    BTRY(bs.sortTmp.ensure(tmpBytes + 16));
    BTRY(cub::DeviceRadixSort::SortPairs(bs.sortTmp.p, tmpBytes, keys, keys2, vals, vals2, n, 0, 63, st));
    tm.mark("morton + sort");
    // ints: [radix: left right count visits parent(2n)] [ploc: left right count] [workB2 n] [ploc state] ; boxes: [radix box, dp] [ploc box, dp] [ploc cluster boxes 4n]
    int* ip = bs.ints.p; float4* bp = bs.boxes.p;
    LbvhTree tR, tP;
    tR.left = ip; tR.right = ip + n; tR.count = ip + 2 * (size_t)n; tR.visits = ip + 3 * (size_t)n; tR.parent = ip + 4 * (size_t)n;
    tP.left = ip + 6 * (size_t)n; tP.right = ip + 7 * (size_t)n; tP.count = ip + 8 * (size_t)n; tP.visits = nullptr; tP.parent = nullptr;
    int* workB2 = ip + 9 * (size_t)n;
    tR.box = bp; tR.dp = bp + 2 * (size_t)n; tP.box = bp + 4 * (size_t)n; tP.dp = bp + 6 * (size_t)n;
    tR.cPrim = tP.cPrim = c->envCPrim > 0.0f ? c->envCPrim : 1.6f;
    PlocState S;
    int* ps = ip + 10 * (size_t)n;   // tileCount, tileOffset, nc[2], nodeCounter, then the collapse's level table and tickets
    S.tileCount = ps; S.tileOffset = ps + tiles; S.nc = ps + 2 * (size_t)tiles; S.nodeCounter = S.nc + 2;
    int* dLevel = S.nodeCounter + 2; unsigned* dTickets = reinterpret_cast<unsigned*>(dLevel + maxLevels + 2); int* counters = dLevel + 2 * (size_t)maxLevels + 4;
    S.cbox = bp + 8 * (size_t)n; S.tbox = bp + 10 * (size_t)n;
    S.cid = reinterpret_cast<int*>(keys); S.tcid = S.cid + n;   // cluster ids: the unsorted key array is dead after the sort (n 64-bit keys = 2n ints)
    if (wantRadix) {
        BTRY(cudaMemsetAsync(tR.visits, 0, (size_t)n * sizeof(int), st));
        k_lbvh_tree<<<(n + 255) / 256, 256, 0, st>>>(keys2, n, tR);
        k_lbvh_boxes<<<(n + 255) / 256, 256, 0, st>>>(vals2, bs.primBox.p, n, tR);
    }
    int rootP = 0, it = 0;
    if (wantPloc) {
        const int initNc[3] = {n, n, 0};
        BTRY(cudaMemcpyAsync(S.nc, initNc, sizeof(initNc), cudaMemcpyHostToDevice, st));
        k_ploc_init<<<(n + 255) / 256, 256, 0, st>>>(vals2, bs.primBox.p, n, S);
        const int plocGrid = std::max(1, std::min(tiles, c->smCount * 8));
        int nc = n;
        while (nc > 1) {
            for (int b = 0; b < 24; b++, it++) {
                k_ploc_merge<<<plocGrid, RT_PLOC_TILE, 0, st>>>(it & 1, S, tP);
                k_ploc_scan<<<1, 1024, 0, st>>>(it & 1, S);
                k_ploc_scatter<<<plocGrid, RT_PLOC_TILE, 0, st>>>(it & 1, S);
            }
            BTRY(cudaGetLastError());
            BTRY(cudaMemcpyAsync(&bs.hostInts[0], S.nc + (it & 1), sizeof(int), cudaMemcpyDeviceToHost, st));
            BTRY(cudaMemcpyAsync(&bs.hostInts[1], S.cid, sizeof(int), cudaMemcpyDeviceToHost, st));
            BTRY(cudaStreamSynchronize(st));
            nc = bs.hostInts[0];
            if (it > 64 * 24) return cudaErrorUnknown;   // cannot happen: every iteration merges at least one pair
        }
        rootP = bs.hostInts[1];   // the last cluster
    }
    tm.mark("binary trees");
    bool usePloc = wantPloc;
    float costR = 0.0f, costP = 0.0f;
    if (wantRadix && wantPloc) {
        float4 cr, cp;
        BTRY(cudaMemcpyAsync(&cr, tR.dp, sizeof(float4), cudaMemcpyDeviceToHost, st));   // pageable targets: these copies return when done
        BTRY(cudaMemcpyAsync(&cp, tP.dp + 2 * (size_t)rootP, sizeof(float4), cudaMemcpyDeviceToHost, st));
        BTRY(cudaStreamSynchronize(st));
        costR = cr.x; costP = cp.x;
        usePloc = costP < costR;
    }
    const LbvhTree& t = usePloc ? tP : tR;
    const int init[4] = {1, 0, usePloc ? rootP : 0, 0};   // counters: wide nodes made, primitive records placed; workB2[0]: wide node 0 owns the binary root
    BTRY(cudaMemcpyAsync(counters, init, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
    BTRY(cudaMemcpyAsync(workB2, init + 2, sizeof(int), cudaMemcpyHostToDevice, st));
    // one launch per POSSIBLE level, queued back to back: every launch reads its range from the device-side level table the previous
    // one completed (k_lbvh_collapse_level), so the host waits once, at the end, for the whole tree
    std::vector<int> hLevel((size_t)maxLevels + 2, 0); hLevel[1] = 1;
    BTRY(cudaMemcpyAsync(dLevel, hLevel.data(), hLevel.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    BTRY(cudaMemsetAsync(dTickets, 0, (size_t)maxLevels * sizeof(unsigned), st));
    const int collapseGrid = std::max(1, std::min((n + 127) / 128, c->smCount * 16));
    for (int level = 0; level < maxLevels; level++)
        k_lbvh_collapse_level<<<collapseGrid, 128, 0, st>>>(level, dLevel, dTickets, n, t, vals2, bs.primBox.p, bs.primsIn.p, bs.nodes.p, bs.prims.p, workB2, counters);
    BTRY(cudaGetLastError());
    BTRY(cudaMemcpyAsync(hLevel.data(), dLevel, hLevel.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    BTRY(cudaStreamSynchronize(st));
    levelStart.assign(1, 0);
    int last = 1;
    *tooDeep = false;
    for (int level = 0; level < maxLevels; level++) {
        if (hLevel[(size_t)level + 1] <= hLevel[(size_t)level]) break;   // an empty level: the tree ended above it
        levelStart.push_back(hLevel[(size_t)level + 1]);
        last = hLevel[(size_t)level + 1];
    }
    if (hLevel[(size_t)maxLevels + 1] > hLevel[(size_t)maxLevels]) *tooDeep = true;   // the last possible level still made children
    *nNodesOut = last;
    tm.mark("collapse");
    if (c->envBuildTiming) fprintf(stderr, "rtcore_b200 device build: %s tree (collapsed SAH cost: radix %.4g, ploc %.4g after %d iterations), %d wide nodes\n", usePloc ? "ploc" : "radix", costR, costP, it, last);
    tm.print("device build");
#undef BTRY
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------ multi-GPU behind the ABI
// NCCL is loaded at run time ("libnccl.so.2": the copy already in the process when the host brought one - e.g. PyTorch's - else
// the system's), so the library has no link-time dependency and a single-GPU host never touches it.
struct NcclApi {
    void* handle = nullptr; std::string err;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr; decltype(&ncclCommInitRank) CommInitRank = nullptr; decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr; decltype(&ncclRecv) Recv = nullptr; decltype(&ncclGroupStart) GroupStart = nullptr; decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr; decltype(&ncclGetErrorString) GetErrorString = nullptr; decltype(&ncclGetVersion) GetVersion = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api; static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    // 1. RT_NCCL_LIBRARY: an explicit path; 2. the copy ALREADY in the process (a host that ships its own NCCL - PyTorch's wheel does -
    // must load it before the first rt_comm_* call, or a later load of that host library would find this one under the same soname);
    // 3. the system's libnccl.so.2
    void* h = nullptr;
    if (const char* path = getenv("RT_NCCL_LIBRARY")) h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { api.err = std::string("NCCL is not available: ") + dlerror(); return nullptr; }
    bool ok = true;
    auto sym = [&](const char* name) -> void* { void* p = dlsym(h, name); if (!p) { ok = false; api.err = std::string("NCCL symbol missing: ") + name; } return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId"); api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy"); api.Send = (decltype(api.Send))sym("ncclSend"); api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart"); api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather"); api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    if (!ok) return nullptr;
    api.handle = h;
    return &api;
}
static const char* nccl_err() { static NcclApi dummy; NcclApi* a = nccl_api(); (void)dummy; return a ? "" : "NCCL is not available (libnccl.so.2 could not be loaded)"; }
#define NCCL_TRY(expr)                                                                                                  \
    do {                                                                                                                \
        ncclResult_t _r = (expr);                                                                                       \
        if (_r != ncclSuccess) return fail(RT_ERR_NCCL, std::string(#expr) + ": " + nc->GetErrorString(_r));            \
    } while (0)

// "all-gather" of arrays laid out like the concatenated owned-pixel lists (rank r owns [deintStart[r], deintStart[r + 1])): every
// rank sends its own segment of every array to every other rank and receives theirs in place (exact counts; grouped send / recv,
// because the segments differ in length by a tile or two).  On the context's render stream: the kernels that follow need the data.
static int exchange_segments(rt_ctx* c, float4* const* arrays, int nArrays) {
    NcclApi* nc = nccl_api();
    if (!nc) return fail(RT_ERR_UNSUPPORTED, nccl_err());
    const int world = c->commWorld, me = c->commRank;
    const size_t myStart = (size_t)c->deintStart[(size_t)me], myN = (size_t)(c->deintStart[(size_t)me + 1] - c->deintStart[(size_t)me]);
    NCCL_TRY(nc->GroupStart());
    for (int r = 0; r < world; r++) {
        if (r == me) continue;
        const size_t rStart = (size_t)c->deintStart[(size_t)r], rN = (size_t)(c->deintStart[(size_t)r + 1] - c->deintStart[(size_t)r]);
        for (int a = 0; a < nArrays; a++) {
            if (myN) NCCL_TRY(nc->Send(arrays[a] + myStart, myN * 4, ncclFloat, r, c->comm, c->stream));
            if (rN) NCCL_TRY(nc->Recv(arrays[a] + rStart, rN * 4, ncclFloat, r, c->comm, c->stream));
        }
    }
    NCCL_TRY(nc->GroupEnd());
    return RT_OK;
}

extern "C" {
static int ensure_deint_map(rt_ctx* c, int width, int height, int T, int worldSize);   // defined further down, inside the extern "C" block

RT_API int rt_abi_version(void) { return RT_ABI_VERSION; }
RT_API const char* rt_last_error(void) { return g_lastError.c_str(); }

RT_API int rt_destroy(rt_ctx* c);
RT_API int rt_comm_destroy(rt_ctx* c);

RT_API int rt_create(const int* deviceIds, int nDev, rt_ctx** out) {
    if (!out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_create: out is null");
    *out = nullptr;
    if (nDev != 1 && !(nDev == 0 && deviceIds == nullptr))
        return fail(RT_ERR_INVALID_ARGUMENT, "rt_create: one context drives one GPU (nDev must be 1); use one process per GPU for multi-GPU");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return fail(RT_ERR_NO_DEVICE, std::string("rt_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    int dev = (deviceIds && nDev == 1) ? deviceIds[0] : 0;
    if (dev < 0 || dev >= count) return fail(RT_ERR_INVALID_ARGUMENT, "rt_create: device index out of range");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail(RT_ERR_NO_DEVICE, std::string("rt_create: device '") + prop.name + "' is not sm_100 class; the kernels are built for sm_100a only");
    CUDA_TRY(cudaSetDevice(dev));
    rt_ctx* c = new rt_ctx();
    c->device = dev; c->smCount = prop.multiProcessorCount;
    c->memTotal = (size_t)prop.totalGlobalMem;
    c->l2PersistMax = (size_t)prop.persistingL2CacheMaxSize; c->l2WindowMax = (size_t)prop.accessPolicyMaxWindowSize;
    // developer knobs (tuning sweeps under tests/), read ONCE here: nothing on the per-frame path calls getenv
    c->envNoL2Persist = getenv("RT_NO_L2_PERSIST") != nullptr;
    c->envNoSunProbe = getenv("RT_NO_SUN_PROBE") != nullptr;
    if (const char* e = getenv("RT_PATHS_PER_PASS")) { const long long v = atoll(e); if (v > 0) c->envPathsPerPass = v; }
    if (const char* e = getenv("RT_DEVICE_TREE")) c->envDeviceTree = strcmp(e, "radix") == 0 ? 1 : (strcmp(e, "ploc") == 0 ? 2 : 0);
    if (const char* e = getenv("RT_BVH_CPRIM")) c->envCPrim = (float)atof(e);
    if (const char* e = getenv("RT_DUMP_COUNTERS")) c->envDumpCounters = atoi(e) != 0;
    if (const char* e = getenv("RT_BUILD_TIMING")) c->envBuildTiming = atoi(e) != 0;
    const int rc = [&]() -> int {   // any failure below releases what the context already holds (rt_destroy)
        CUDA_TRY(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
        c->stream = c->ownStream;
        CUDA_TRY(cudaEventCreate(&c->evStart)); CUDA_TRY(cudaEventCreate(&c->evStop));
        const int r = size_extend_launch(c, RT_STACK_ENTRIES);
        if (r != RT_OK) return r;
        memset(&c->ds, 0, sizeof(c->ds));
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->hstats), sizeof(DeviceStats), cudaHostAllocDefault));
        memset(c->hstats, 0, sizeof(DeviceStats));
        return RT_OK;
    }();
    if (rc != RT_OK) { const std::string msg = g_lastError; rt_destroy(c); g_lastError = msg; return rc; }
    *out = c;
    return RT_OK;
}

RT_API int rt_destroy(rt_ctx* c) {
    if (!c) return RT_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    rt_comm_destroy(c);
    if (c->frameGraphExec) { cudaGraphExecDestroy(c->frameGraphExec); c->frameGraphExec = nullptr; }
    if (c->frameGraph) { cudaGraphDestroy(c->frameGraph); c->frameGraph = nullptr; }
    if (c->copyStream) { cudaStreamSynchronize(c->copyStream); cudaStreamDestroy(c->copyStream); c->copyStream = nullptr; }
    if (c->evPrimaryDone) cudaEventDestroy(c->evPrimaryDone);
    if (c->evCopyDone) cudaEventDestroy(c->evCopyDone);
    c->bvhBlob.release(); c->instances.release(); c->spheres.release(); c->texcoords.release(); c->triUVs.release(); c->triMat.release();
    c->materials.release(); c->texels.release(); c->texInfos.release(); c->presentBuf.release(); c->taaHistColor.release(); c->taaHistObj.release(); c->pixelMap.release(); c->invPixelMap.release(); c->resAB[0].release(); c->resAB[1].release(); c->resPath.release();
    c->gbPosHit.release(); c->gbNrmMat.release(); c->gbAlbObj.release(); c->lframe.release(); for (int b = 0; b < 2; b++) { c->tileRadiance[b].release(); c->tileAux[b].release(); c->tileRgba[b].release(); } c->primId.release(); c->instId.release(); c->primaryT.release();
    c->rgba8.release(); c->objId.release(); c->depth.release(); c->radiance.release(); c->accum.release();
    c->pathState.release(); for (int b = 0; b < 2; b++) { c->qO[b].release(); c->qD[b].release(); } c->shO.release(); c->shD.release(); c->hitTuv.release();
    c->hitPrim.release(); c->pathHash.release(); c->segOut.release(); c->termOut.release(); c->hashOut.release(); c->scratch.release(); c->counters.release(); c->dstats.release(); c->deintMap.release();
    for (auto ev : c->traceEvents) cudaEventDestroy(ev);
    if (c->evStart) cudaEventDestroy(c->evStart);
    if (c->evStop) cudaEventDestroy(c->evStop);
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    if (c->hstats) cudaFreeHost(c->hstats);
    delete c;
    return RT_OK;
}

RT_API int rt_set_stream(rt_ctx* c, void* s) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_set_stream: ctx is null");
    c->stream = s ? (cudaStream_t)s : c->ownStream;
    return RT_OK;
}

RT_API int rt_get_stream(rt_ctx* c, void** s) {
    if (!c || !s) return fail(RT_ERR_INVALID_ARGUMENT, "rt_get_stream: null argument");
    *s = (void*)c->stream;
    return RT_OK;
}

RT_API int rt_scene_upload(rt_ctx* c, const RtSceneDesc* d) { return rt_scene_upload_ex(c, d, 0u); }

RT_API int rt_scene_upload_ex(rt_ctx* c, const RtSceneDesc* d, uint32_t buildFlags) {
    if (!c || !d) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: null argument");
    if (buildFlags & ~(uint32_t)RT_BUILD_DEVICE_LBVH) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload_ex: unknown build flag");
    const void* ptrs[15] = {d->tlasNodes, d->tlasInstanceIndices, d->instances, d->blasNodes, d->spherePrimIdx, d->spheres, d->triPrimIdx, d->meshPositions,
                            d->meshTris, d->meshTexcoords, d->meshTriUVs, d->triMatIndex, d->materials, d->texels, d->texInfos};
    const int64_t cnts[15] = {d->nTlasNodes, d->nTlasInstanceIndices, d->nInstances, d->nBlasNodes, d->nSpherePrimIdx, d->nSpheres, d->nTriPrimIdx, d->nMeshPositions,
                              d->nMeshTris, d->nMeshTexcoords, d->nMeshTriUVs, d->nTriMatIndex, d->nMaterials, d->nTexels, d->nTexInfos};
    for (int i = 0; i < 15; i++) {
        if (cnts[i] < 0 || cnts[i] > 0x7FFFFFFF) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: array length out of range");
        if (cnts[i] > 0 && !ptrs[i]) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: null array with non-zero length");
    }
    for (int64_t i = 0; i < d->nTexInfos; i++) {
        const RtTexInfo& ti = d->texInfos[i];
        if (ti.Width > 0 && ti.Height > 0 && (ti.Offset < 0 || (int64_t)ti.Offset + (int64_t)ti.Width * ti.Height > d->nTexels))
            return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: texInfos entry addresses texels out of range");
    }
    HostBvh bvh; std::string err;
    BuildTimer tmU(c->envBuildTiming, c->stream);
    CUDA_TRY(cudaSetDevice(c->device));
    // the device builder needs a few primitives to make a tree of; tiny scenes take the host builder either way
    bool onDevice = (buildFlags & RT_BUILD_DEVICE_LBVH) != 0;
    if (!onDevice) c->build.release();   // the device builder keeps its scratch between ITS commits only
    if (onDevice) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));   // an earlier commit's copies out of the staging arrays are done
        c->build.sink.user = &c->build; c->build.stream = c->stream; c->build.recordsQueued = false;
        c->build.sink.reserve = [](PrimSink* self, size_t n) { return static_cast<BuildScratch*>(self->user)->reserve(n); };
        c->build.sink.recordsDone = [](PrimSink* self, size_t n) {
            BuildScratch* b = static_cast<BuildScratch*>(self->user);
            b->recordsQueued = cudaMemcpyAsync(b->primsIn.p, b->stagePrims, n * sizeof(PrimRec), cudaMemcpyHostToDevice, b->stream) == cudaSuccess;
        };
    }
    if (!build_wide_bvh(*d, bvh, err, onDevice ? &c->build.sink : nullptr)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: " + err);
    if (onDevice && bvh.stats.nPrims < 64) { onDevice = false; if (!build_wide_bvh(*d, bvh, err)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: " + err); }
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // nothing in flight may still read the old scene
    tmU.mark("host stage");
    cudaStream_t st = c->stream;
    // From here on the old scene is torn down buffer by buffer (DevBuf::ensure frees before it allocates): the context has NO
    // scene until every array of the new one is in place, so a failure half way (out of memory, a CUDA error) leaves a context
    // that refuses to render (RT_ERR_INVALID_STATE) instead of one that traverses freed or null pointers.
    c->hasScene = false; c->ds.nNodes = 0; c->ds.nPrims = 0;
    DeviceScene& ds = c->ds;
    int builtNodeCount = 0;
    if (onDevice) {
        bool tooDeep = false;
        CUDA_TRY(build_on_device(c, bvh, bvh.levelStart, &builtNodeCount, &tooDeep));
        if (tooDeep) {   // a degenerate Morton order (very uneven extents): the host builder bounds the depth
            onDevice = false;
            if (!build_wide_bvh(*d, bvh, err)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_upload: " + err);
        } else {
            bvh.stats.nWideNodes = builtNodeCount; bvh.stats.maxDepth = std::max(1, (int)bvh.levelStart.size() - 1);
        }
    }
    tmU.mark("device build");
    const size_t nNodes = onDevice ? (size_t)builtNodeCount : bvh.nodes.size(), nPrims = (size_t)bvh.stats.nPrims;
    const size_t nodeBytes = (std::max<size_t>(1, nNodes) * sizeof(WideNode) + 255) / 256 * 256;
    const size_t primBytes = std::max<size_t>(1, nPrims) * sizeof(PrimRec);
    CUDA_TRY(c->bvhBlob.ensure(nodeBytes + primBytes));
    WideNode* dNodes = reinterpret_cast<WideNode*>(c->bvhBlob.p);
    PrimRec* dPrims = reinterpret_cast<PrimRec*>(c->bvhBlob.p + nodeBytes);
    if (onDevice) {
        CUDA_TRY(cudaMemcpyAsync(dNodes, c->build.nodes.p, nNodes * sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(dPrims, c->build.prims.p, nPrims * sizeof(PrimRec), cudaMemcpyDeviceToDevice, st));
    } else {
        if (nNodes) CUDA_TRY(cudaMemcpyAsync(dNodes, bvh.nodes.data(), nNodes * sizeof(WideNode), cudaMemcpyHostToDevice, st));
        if (nPrims) CUDA_TRY(cudaMemcpyAsync(dPrims, bvh.prims.data(), nPrims * sizeof(PrimRec), cudaMemcpyHostToDevice, st));
    }
    ds.nodes = dNodes; ds.nNodes = nPrims == 0 ? 0 : (int)nNodes; ds.prims = dPrims; ds.nPrims = (int)nPrims;
    // keep the BVH resident in L2 while GBs of path state stream past it: persisting carve-out + access-policy window (applied per stream in rt_render)
    c->l2Window = 0;
    if (nPrims > 0 && c->l2PersistMax > 0 && !c->envNoL2Persist) {
        const size_t want = std::min(nodeBytes + primBytes, c->l2PersistMax);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) c->l2Window = std::min(nodeBytes + primBytes, c->l2WindowMax);
        c->l2Persist = want;
    }
    c->l2WindowStream = nullptr;
    tmU.mark("bvh blob");
    CUDA_TRY(upload_or_one(c->instances, d->instances, d->nInstances, st, &ds.nInstances)); ds.instances = c->instances.p;
    CUDA_TRY(upload_or_one(c->spheres, d->spheres, d->nSpheres, st, &ds.nSpheres)); ds.spheres = c->spheres.p;
    CUDA_TRY(upload_or_one(c->texcoords, d->meshTexcoords, d->nMeshTexcoords, st, nullptr)); ds.texcoords = c->texcoords.p;
    CUDA_TRY(upload_or_one(c->triUVs, d->meshTriUVs, d->nMeshTriUVs, st, nullptr)); ds.triUVs = c->triUVs.p;
    CUDA_TRY(upload_or_one(c->triMat, d->triMatIndex, d->nTriMatIndex, st, nullptr)); ds.triMatIndex = c->triMat.p;
    CUDA_TRY(upload_or_one(c->materials, d->materials, d->nMaterials, st, &ds.nMaterials)); ds.materials = c->materials.p;
    CUDA_TRY(upload_or_one(c->texels, d->texels, d->nTexels, st, nullptr)); ds.texels = c->texels.p;
    CUDA_TRY(upload_or_one(c->texInfos, d->texInfos, d->nTexInfos, st, &ds.nTexInfos)); ds.texInfos = c->texInfos.p;
    ds.triMaterials = 0;
    ds.tFarScale = bvh.stats.maxInstanceScale;
    CUDA_TRY(upload_or_one(c->meshTris, d->meshTris, d->nMeshTris, st, nullptr));
    CUDA_TRY(upload_or_one(c->instBoxXf, bvh.instBoxXf.data(), (int64_t)bvh.instBoxXf.size(), st, nullptr));
    c->nMeshPositions = d->nMeshPositions; c->nMeshTris = d->nMeshTris; c->levelStart = bvh.levelStart;
    CUDA_TRY(cudaStreamSynchronize(st));   // host arrays are only borrowed for the duration of the call
    tmU.mark("scene arrays");
    const int rcSize = size_extend_launch(c, bvh.stats.maxDepth + 1);   // a node step pushes at most one entry per level
    if (rcSize != RT_OK) return rcSize;
    c->bvhStats = bvh.stats;
    c->bvhBytes = nNodes * sizeof(WideNode) + nPrims * sizeof(PrimRec);
    c->hasScene = true; c->sceneVersion++;
    tmU.mark("launch sizing");
    tmU.print("scene commit");
    return RT_OK;
}

RT_API int rt_scene_refit(rt_ctx* c, const RtFloat3* meshPositions, int64_t nMeshPositions) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_refit: ctx is null");
    if (!c->hasScene) return fail(RT_ERR_INVALID_STATE, "rt_scene_refit: no scene uploaded yet (rt_scene_upload)");
    if (nMeshPositions != c->nMeshPositions) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_refit: the vertex count differs from the uploaded mesh (a refit keeps the topology; use rt_scene_upload)");
    if (nMeshPositions > 0 && !meshPositions) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_refit: null array with non-zero length");
    for (int64_t i = 0; i < nMeshPositions; i++)
        if (!std::isfinite(meshPositions[i].X) || !std::isfinite(meshPositions[i].Y) || !std::isfinite(meshPositions[i].Z)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_scene_refit: non-finite vertex position");
    const int nPrims = c->ds.nPrims, nNodes = c->ds.nNodes;
    if (nPrims <= 0 || nNodes <= 0) return RT_OK;   // empty scene: nothing to refit
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;   // stream order: frames already queued still see the old geometry
    CUDA_TRY(upload_or_one(c->refitPos, meshPositions, nMeshPositions, st, nullptr));
    CUDA_TRY(c->refitPrimBox.ensure(2 * (size_t)nPrims)); CUDA_TRY(c->refitNodeBox.ensure(2 * (size_t)nNodes)); CUDA_TRY(c->refitAbs.ensure(1));
    CUDA_TRY(cudaMemsetAsync(c->refitAbs.p, 0, sizeof(unsigned), st));
    WideNode* nodes = const_cast<WideNode*>(c->ds.nodes); PrimRec* prims = const_cast<PrimRec*>(c->ds.prims);
    k_refit_prims<<<(nPrims + 255) / 256, 256, 0, st>>>(prims, nPrims, c->refitPos.p, c->meshTris.p, c->instBoxXf.p, c->refitPrimBox.p, c->refitAbs.p);
    for (int l = (int)c->levelStart.size() - 2; l >= 0; l--) {   // leaves first
        const int first = c->levelStart[l], last = c->levelStart[l + 1];
        if (last > first) k_refit_nodes<<<(last - first + 127) / 128, 128, 0, st>>>(nodes, first, last, c->refitPrimBox.p, c->refitNodeBox.p, c->refitAbs.p);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));   // the host array is only borrowed for the duration of the call
    return RT_OK;
}

static inline int grid_for(const rt_ctx* c, size_t n, int threads) {
    size_t blocks = (n + threads - 1) / threads;
    size_t cap = (size_t)c->smCount * 16;
    return (int)std::max<size_t>(1, std::min(blocks, cap));
}

static int ensure_frame_buffers(rt_ctx* c, const RtRenderConfig* cfg, int S) {
    const int W = cfg->width, H = cfg->height;
    const int T = effective_tile_size(cfg->tileSize);
    const int world = cfg->worldSize > 1 ? cfg->worldSize : 1;
    const int rank = world > 1 ? cfg->rank : 0;
    if (c->width != W || c->height != H || c->tileSize != T || c->rank != rank || c->worldSize != world || !c->pixelMap.p) {
        if (c->commStream) CUDA_TRY(cudaStreamSynchronize(c->commStream));   // a gather in flight still reads the payload buffers about to be re-sized
        c->gatherPending[0] = c->gatherPending[1] = false;
        std::vector<int> pm; build_pixel_map(W, H, T, rank, world, pm);
        CUDA_TRY(c->pixelMap.ensure(std::max<size_t>(1, pm.size())));
        if (!pm.empty()) CUDA_TRY(cudaMemcpyAsync(c->pixelMap.p, pm.data(), pm.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->width = W; c->height = H; c->tileSize = T; c->rank = rank; c->worldSize = world; c->npx = (int)pm.size();
        c->invPixelMap.release(); c->resAB[0].release(); c->resAB[1].release(); c->resLastWritten = -1; c->resValid[0] = c->resValid[1] = false;   // reservoirs belong to one image size (Framebuffer.EnsureLength, Framebuffer.cs:60-83)
        const size_t n = std::max<size_t>(1, (size_t)c->npx), g = (size_t)W * H;
        CUDA_TRY(c->gbPosHit.ensure(n)); CUDA_TRY(c->gbNrmMat.ensure(n)); CUDA_TRY(c->gbAlbObj.ensure(n)); CUDA_TRY(c->lframe.ensure(n));
        for (int b = 0; b < (world > 1 ? 2 : 1); b++) { CUDA_TRY(c->tileRadiance[b].ensure(n)); CUDA_TRY(c->tileAux[b].ensure(n)); CUDA_TRY(c->tileRgba[b].ensure(n)); }
        CUDA_TRY(c->primId.ensure(n)); CUDA_TRY(c->instId.ensure(n)); CUDA_TRY(c->primaryT.ensure(n));
        CUDA_TRY(c->rgba8.ensure(g)); CUDA_TRY(c->objId.ensure(g)); CUDA_TRY(c->depth.ensure(g)); CUDA_TRY(c->radiance.ensure(g)); CUDA_TRY(c->accum.ensure(g));
        CUDA_TRY(cudaMemsetAsync(c->rgba8.p, 0, g * sizeof(int), c->stream)); CUDA_TRY(cudaMemsetAsync(c->objId.p, 0, g * sizeof(int), c->stream));
        CUDA_TRY(cudaMemsetAsync(c->depth.p, 0, g * sizeof(float), c->stream)); CUDA_TRY(cudaMemsetAsync(c->radiance.p, 0, g * sizeof(float4), c->stream));
        CUDA_TRY(cudaMemsetAsync(c->accum.p, 0, g * sizeof(float4), c->stream));
    }
    const size_t P = std::max<size_t>(1, (size_t)c->npx * S);
    if (P > c->pathCap) {
        CUDA_TRY(c->pathState.ensure(P));
        for (int b = 0; b < 2; b++) { CUDA_TRY(c->qO[b].ensure(P)); CUDA_TRY(c->qD[b].ensure(P)); }
        CUDA_TRY(c->shO.ensure(P)); CUDA_TRY(c->shD.ensure(P)); CUDA_TRY(c->hitTuv.ensure(P));
        CUDA_TRY(c->hitPrim.ensure((P + RT_SHADE_CHUNK - 1) / RT_SHADE_CHUNK * RT_SHADE_CHUNK));   // padded: k_shade_next scans whole chunks with 128-bit loads
        c->pathCap = P;
    }
    c->aovs = (cfg->flags & RT_FLAG_PATH_AOVS) != 0;
    c->spp = cfg->spp > 1 ? cfg->spp : 1;
    if (c->aovs) {
        const size_t g = (size_t)W * H * c->spp;
        CUDA_TRY(c->pathHash.ensure(P)); CUDA_TRY(c->segOut.ensure(g)); CUDA_TRY(c->termOut.ensure(g)); CUDA_TRY(c->hashOut.ensure(g));
        CUDA_TRY(cudaMemsetAsync(c->segOut.p, 0, g, c->stream)); CUDA_TRY(cudaMemsetAsync(c->termOut.p, 0, g, c->stream)); CUDA_TRY(cudaMemsetAsync(c->hashOut.p, 0, g * 4, c->stream));
    }
    return RT_OK;
}

RT_API int rt_render(rt_ctx* c, const RtCamera* cam, const RtCamera* prevCam, const RtRenderConfig* cfg) {
    if (!c || !cam || !cfg) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: null argument");
    if (!c->hasScene) return fail(RT_ERR_INVALID_STATE, "rt_render: no scene uploaded (call rt_scene_upload first)");
    if (cfg->width <= 0 || cfg->height <= 0 || (int64_t)cfg->width * cfg->height > 0x7FFFFFFF) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: bad image size");
    if (cfg->maxDepth < 0 || cfg->maxDepth > 255) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: maxDepth must be in [0, 255]");
    if (cfg->worldSize > 1 && (cfg->rank < 0 || cfg->rank >= cfg->worldSize)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: rank outside [0, worldSize)");
    // "reuse" frames run the <REUSE> shade kernels: they publish resCur (RTRay.cs:289-296) and import when a reuse flag is set.
    // The reference publishes on EVERY frame; here a frame without imports publishes only on request (RT_FLAG_PUBLISH_RESERVOIRS:
    // 48 B per path of extra traffic), and a reuse frame whose previous-frame buffer was not written by frame - 1 imports zeros.
    const bool reuse = cfg->enableTemporalReuse != 0 || cfg->enableSpatialReuse != 0 || (cfg->flags & RT_FLAG_PUBLISH_RESERVOIRS) != 0;
    // Reuse reads the previous frame's reservoirs and the current G-buffer of neighbouring / reprojected pixels (RTRay.cs:363-374,
    // 408-435, 476-516), which a screen-tile partition does not hold: with the library's communicator the ranks exchange both
    // (G-buffer after the primary pass, reservoirs after accumulate); without one the frame cannot be rendered.
    const bool reuseDist = reuse && cfg->worldSize > 1;
    if (reuseDist && !(c->comm && c->commWorld == cfg->worldSize && c->commRank == cfg->rank))
        return fail(RT_ERR_UNSUPPORTED, "rt_render: ReSTIR temporal/spatial reuse across a screen-tile partition needs the previous frame's reservoirs and the current "
                                        "G-buffer of other ranks' pixels: call rt_comm_init with this rank / worldSize first (or render reuse frames with worldSize = 1)");
    if (cfg->enableTemporalReuse != 0 && !prevCam) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: enableTemporalReuse needs prevCam");
    if (c->extColor && c->extColorBytes < (size_t)cfg->width * cfg->height * sizeof(int))   // Framebuffer.GetGpuWithExternalColor's guard (Framebuffer.cs:117), before anything is queued
        return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: mapped external colour buffer is smaller than the image");
    for (int b = 0; b < 3; b++)
        if (c->rbHost[b] && c->rbBytes[b] != (size_t)cfg->width * cfg->height * 4)
            return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: a bound read-back target (rt_bind_readback) does not have the size of this frame's image");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->copyPending) { CUDA_TRY(cudaStreamWaitEvent(c->stream, c->evCopyDone, 0)); c->copyPending = false; }   // the previous frame's bound read-backs still read depth / objectId

    const int spp = cfg->spp > 1 ? cfg->spp : 1;
    const int64_t npxOwned = count_owned_pixels(cfg->width, cfg->height, cfg->tileSize, cfg->worldSize > 1 ? cfg->rank : 0, cfg->worldSize > 1 ? cfg->worldSize : 1);
    int S = cfg->samplesPerPass;
    if (S <= 0) {
        // samples per wavefront pass: as many paths in flight as ~60 % of the device memory holds (180 B of path state and queue
        // slots each), capped at 768 Mi (path indices are 32-bit).  Bigger passes mean fewer, longer launches: C4 on a 180 GB B200
        // runs its 64 spp (531 M paths, 96 GB) in ONE pass - 191.2 ms against 193.1 ms in two passes of 32 spp, 197.8 ms in four,
        // 230 ms with 16 Mi-path passes (launch tails of the deep, nearly empty wavefronts)
        int64_t target = std::min<int64_t>(768ll << 20, (int64_t)(c->memTotal / 10 * 6 / RT_PATH_BYTES));
        // only a frame whose paths do not fit the buffers already there asks the driver how much is free (cudaMemGetInfo costs ~1 ms:
        // more than a whole interactive frame): the buffers then grow to at most 85 % of what is free right now (plus what they hold)
        if ((size_t)std::min<int64_t>(target, npxOwned * (int64_t)spp) > c->pathCap) {
            size_t freeB = 0, totalB = 0;
            if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) target = std::min<int64_t>(target, (int64_t)((freeB / 100 * 85 + c->pathCap * RT_PATH_BYTES) / RT_PATH_BYTES));
            else (void)cudaGetLastError();
        }
        target = std::max<int64_t>(target, 1ll << 20);
        if (c->envPathsPerPass > 0) target = c->envPathsPerPass;
        S = (int)std::max<int64_t>(1, std::min<int64_t>(spp, target / std::max<int64_t>(1, npxOwned)));
    }
    S = std::min(S, spp);
    if ((int64_t)npxOwned * S > 0x7FFFFFFF) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: samplesPerPass * pixels exceeds 2^31 paths");
    int rc = ensure_frame_buffers(c, cfg, S);
    if (rc != RT_OK) return rc;
    const int npx = c->npx;
    const int nPasses = (spp + S - 1) / S;
    if (c->worldSize > 1) {   // the other set of gather payload buffers; the gather that read it two frames ago must be through
        c->tileBuf ^= 1;
        if (c->gatherPending[c->tileBuf]) { CUDA_TRY(cudaStreamWaitEvent(c->stream, c->evGatherDone[c->tileBuf], 0)); c->gatherPending[c->tileBuf] = false; }
    } else c->tileBuf = 0;
    if (reuse) {
        const size_t g = (size_t)cfg->width * cfg->height;
        if (reuseDist) {
            const int rcMap = ensure_deint_map(c, cfg->width, cfg->height, c->tileSize, cfg->worldSize);
            if (rcMap != RT_OK) return rcMap;
            CUDA_TRY(c->gbAll.ensure(3 * g)); CUDA_TRY(c->resPack.ensure(3 * g));
        }
        if (!c->invPixelMap.p) {
            std::vector<int> pm, inv(g, 0);
            if (reuseDist) pm = c->deintHost;   // all ranks' owned-pixel lists concatenated: index = position in the exchanged arrays
            else build_pixel_map(cfg->width, cfg->height, c->tileSize, 0, 1, pm);
            for (size_t i = 0; i < pm.size(); i++) inv[(size_t)pm[i]] = (int)i;
            CUDA_TRY(c->invPixelMap.ensure(g));
            CUDA_TRY(cudaMemcpyAsync(c->invPixelMap.p, inv.data(), g * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
        const int prevB = (cfg->frame & 1) ^ 1;
        for (int b = 0; b < 2; b++) {
            // zero: a new buffer, an explicit restart, or the "previous frame" buffer when frame - 1 did not write it (reuse was off,
            // or the frame index jumped): stale reservoirs of some older frame must not be imported as if they were last frame's
            const bool stale = b == prevB && !(c->resValid[b] && c->resFrame[b] == cfg->frame - 1);
            if (!c->resAB[b].p || (cfg->flags & RT_FLAG_RESET_RESERVOIRS) || stale) {
                CUDA_TRY(c->resAB[b].ensure(3 * g)); CUDA_TRY(cudaMemsetAsync(c->resAB[b].p, 0, 3 * g * sizeof(float4), c->stream));
                c->resValid[b] = false;
            }
        }
        const size_t P = std::max<size_t>(1, (size_t)npx * S);
        if (P > c->resPathCap) { CUDA_TRY(c->resPath.ensure(3 * P)); c->resPathCap = P; }
    }
    const bool count = (cfg->flags & RT_FLAG_COUNTERS) != 0;
    const bool fast = (cfg->flags & RT_FLAG_FAST_SHADING) != 0;
    c->timeKernels = (cfg->flags & RT_FLAG_KERNEL_TIMING) != 0;   // per-launch event pairs around the extend kernels (roofline measurements)
    c->traceEventsUsed = 0;
    c->launches = 0;
    c->ds.triMaterials = (cfg->flags & RT_FLAG_TRI_MATERIALS) ? 1 : 0;

    // device counters: [0] primary ray count, [1] its work cursor, [2] sun-probe count, [3] its work cursor, then per (pass, depth):
    // nextCount, shCount, workClosest, workShadow, hits shaded at this depth (RT_DUMP_COUNTERS), 3 spare
    const int CS = 8, CH = 4;   // ints per (pass, depth); header ints
    c->ctrStride = CS; c->ctrHeader = CH; c->ctrDepths = cfg->maxDepth + 1; c->ctrPasses = nPasses;
    const size_t nCounters = CH + (size_t)nPasses * (cfg->maxDepth + 1) * CS;
    CUDA_TRY(c->counters.ensure(nCounters));
    CUDA_TRY(c->dstats.ensure(1));
    cudaStream_t st = c->stream;
    if (c->l2Window > 0 && c->l2WindowStream != st) {   // (re)attach the persisting window to whichever stream renders
        cudaStreamAttrValue av; memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.base_ptr = c->bvhBlob.p; av.accessPolicyWindow.num_bytes = c->l2Window;
        av.accessPolicyWindow.hitRatio = c->l2Persist >= c->l2Window ? 1.0f : (float)c->l2Persist / (float)c->l2Window;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) (void)cudaGetLastError();
        c->l2WindowStream = st;
    }
    // RT_FLAG_FRAME_GRAPH: the frame's launch sequence is static for a configuration (queue sizes live on the device), so it is
    // captured into a CUDA graph; the first frame instantiates it, later frames refresh the node parameters (camera, frame index,
    // buffer parity) with cudaGraphExecUpdate and replay it with ONE launch - the per-launch gaps of small, launch-bound frames go.
    // Not with per-launch timing, the phase statistics build or NCCL exchanges inside the frame.
    const bool useGraph = (cfg->flags & RT_FLAG_FRAME_GRAPH) != 0 && !c->timeKernels && !reuseDist && !RT_PHASE_STATS;
    const bool sunProbe = spp >= 2 && cfg->maxDepth >= 1 && !reuse && !c->envNoSunProbe;
    FrameKey key; memset(&key, 0, sizeof(key));
    key.width = cfg->width; key.height = cfg->height; key.spp = spp; key.S = S; key.maxDepth = cfg->maxDepth; key.worldSize = c->worldSize; key.rank = c->rank; key.tileSize = c->tileSize; key.npx = npx;
    key.structuralFlags = cfg->flags & (RT_FLAG_COUNTERS | RT_FLAG_FAST_SHADING | RT_FLAG_PATH_AOVS); key.reuse = reuse ? 1 : 0; key.sunProbe = sunProbe ? 1 : 0; key.extColor = c->extColor ? 1 : 0;
    key.counters = c->counters.p; key.dstats = c->dstats.p; key.hstats = c->hstats; key.bvh = c->bvhBlob.p; key.stream = st; key.nCounters = nCounters; key.pathCap = c->pathCap; key.sceneVersion = c->sceneVersion;
    // bound read-backs: with plain launches on one GPU depth / objectId leave as soon as the primary pass is done; otherwise (frame graph,
    // tile partition: the root's gathered image is what the host wants, see rt_gather_frame) everything is copied behind the frame
    const bool anyReadback = c->rbHost[0] || c->rbHost[1] || c->rbHost[2];
    const bool earlyReadback = anyReadback && !useGraph && c->worldSize <= 1 && (c->rbHost[1] || c->rbHost[2]);
    FrameRecorder rec; rec.st = st;
    if (c->l2Window > 0) { rec.l2Base = c->bvhBlob.p; rec.l2Bytes = c->l2Window; rec.l2Ratio = c->l2Persist >= c->l2Window ? 1.0f : (float)c->l2Persist / (float)c->l2Window; }
    if (useGraph) {
        if (c->frameGraphExec && key == c->frameKey) { rec.mode = FrameRecorder::Update; rec.exec = c->frameGraphExec; }
        else {
            if (c->frameGraphExec) { cudaGraphExecDestroy(c->frameGraphExec); c->frameGraphExec = nullptr; }
            if (c->frameGraph) { cudaGraphDestroy(c->frameGraph); c->frameGraph = nullptr; }
            c->frameNodes.clear(); c->frameFuncs.clear();
            rec.mode = FrameRecorder::Build;
            CUDA_TRY(cudaGraphCreate(&rec.graph, 0));
        }
        rec.kernelNodes = &c->frameNodes; rec.kernelFuncs = &c->frameFuncs;
    }
    CUDA_TRY(cudaEventRecord(c->evStart, st));
    const int rcFrame = [&]() -> int {
    rec.memset32(c->counters.p, nCounters * sizeof(int));
    rec.memset32(c->dstats.p, sizeof(DeviceStats));

    FrameConst fc; memset(&fc, 0, sizeof(fc));
    fc.width = cfg->width; fc.height = cfg->height; fc.frame = cfg->frame; fc.spp = cfg->spp; fc.maxDepth = cfg->maxDepth; fc.rngLockNoise = cfg->rngLockNoise; fc.flags = cfg->flags;
    fc.camOrigin = mk3(cam->origin); fc.camLowerLeft = mk3(cam->lowerLeft); fc.camHorizontal = mk3(cam->horizontal); fc.camVertical = mk3(cam->vertical);
    fc.env.dirLightDir = mk3(cfg->dirLightDir); fc.env.dirLightRadiance = mk3(cfg->dirLightRadiance); fc.env.skyTop = mk3(cfg->skyTintTop); fc.env.skyBottom = mk3(cfg->skyTintBottom);
    fc.npx = npx; fc.pixelMap = c->pixelMap.p;
    if (reuse) {
        const size_t g = (size_t)cfg->width * cfg->height;
        const RtCamera* pc = prevCam ? prevCam : cam;
        fc.enableTemporal = cfg->enableTemporalReuse; fc.enableSpatial = cfg->enableSpatialReuse;
        fc.prevOrigin = mk3(pc->origin); fc.prevRight = mk3(pc->right); fc.prevUp = mk3(pc->up); fc.prevForward = mk3(pc->forward); fc.prevFovY = pc->fovYRadians; fc.prevAspect = pc->aspect;
        fc.invPixelMap = c->invPixelMap.p;
        const float4* prev = c->resAB[(cfg->frame & 1) ^ 1].p;   // GetReservoirPair (Framebuffer.cs:127-146): even frame: prev = B, cur = A
        fc.resPrev0 = prev; fc.resPrev1 = prev + g; fc.resPrev2 = prev + 2 * g;
    }

    WaveBuffers wb; memset(&wb, 0, sizeof(wb));
    wb.gbPosHit = c->gbPosHit.p; wb.gbNrmMat = c->gbNrmMat.p; wb.gbAlbObj = c->gbAlbObj.p; wb.primId = c->primId.p; wb.instId = c->instId.p; wb.primaryT = c->primaryT.p;
    wb.lframe = c->lframe.p; wb.tileRadiance = c->tileRadiance[c->tileBuf].p; wb.tileAux = c->tileAux[c->tileBuf].p; wb.tileRgba = c->tileRgba[c->tileBuf].p; wb.rgba8 = c->rgba8.p; wb.depth = c->depth.p; wb.objId = c->objId.p; wb.radiance = c->radiance.p; wb.accum = c->accum.p;
    wb.st = c->pathState.p;
    if (reuse) {
        const size_t g = (size_t)cfg->width * cfg->height, P = c->resPathCap;
        float4* cur = c->resAB[cfg->frame & 1].p;
        wb.resCur0 = cur; wb.resCur1 = cur + g; wb.resCur2 = cur + 2 * g;
        wb.resPath0 = c->resPath.p; wb.resPath1 = c->resPath.p + P; wb.resPath2 = c->resPath.p + 2 * P;
        c->resLastWritten = cfg->frame & 1;
        c->resValid[cfg->frame & 1] = true; c->resFrame[cfg->frame & 1] = cfg->frame;
    }
    if (c->aovs) { wb.pathHash = c->pathHash.p; wb.segCountOut = c->segOut.p; wb.termCodeOut = c->termOut.p; wb.pathHashOut = c->hashOut.p; }
    size_t ownStart = 0;
    if (reuseDist) {   // the own G-buffer IS this rank's segment of the exchanged arrays: no copy on either side of the exchange
        const size_t g = (size_t)cfg->width * cfg->height;
        ownStart = (size_t)c->deintStart[(size_t)cfg->rank];
        wb.gbPosHit = c->gbAll.p + ownStart; wb.gbNrmMat = c->gbAll.p + g + ownStart; wb.gbAlbObj = c->gbAll.p + 2 * g + ownStart;
        fc.lookPosHit = c->gbAll.p; fc.lookNrmMat = c->gbAll.p + g; fc.lookAlbObj = c->gbAll.p + 2 * g; fc.lookBase = (int)ownStart;
    } else { fc.lookPosHit = wb.gbPosHit; fc.lookNrmMat = wb.gbNrmMat; fc.lookAlbObj = wb.gbAlbObj; fc.lookBase = 0; }
    c->gbPosPtr = wb.gbPosHit; c->gbNrmPtr = wb.gbNrmMat; c->gbAlbPtr = wb.gbAlbObj;

    if (npx > 0 || reuseDist) {
        // ---- primary visibility -------------------------------------------------------------------------------------
        RayQueue q0 = {c->qO[0].p, c->qD[0].p};
        int* primaryCount = c->counters.p + 0;
        int* primaryWork = c->counters.p + 1;
        rec.launch(k_generate_primary, dim3(grid_for(c, npx, 256)), dim3(256), 0, fc, q0, primaryCount); c->launches++;
        const HitQueue hq = {c->hitPrim.p, c->hitTuv.p};
        ExtendArgs ea; memset(&ea, 0, sizeof(ea));
        ea.sc = c->ds; ea.rayO = q0.o; ea.rayD = q0.d; ea.count = primaryCount; ea.work = primaryWork; ea.hits = hq; ea.missSt = nullptr; ea.stats = c->dstats.p; ea.statSlot = 0;
        CUDA_TRY(launch_extend<false>(c, rec, ea, count));
        rec.launch(k_primary_finish, dim3(grid_for(c, npx, 256)), dim3(256), 0, fc, c->ds, wb, q0, hq); c->launches++;
        if (c->comm && c->worldSize > 1) {   // the depth | objectId gather payload is final: rt_gather_frame may send it while the frame still renders
            c->auxReady[c->tileBuf] = false;
            if (rec.mode == FrameRecorder::Direct && rec.err == cudaSuccess) { CUDA_TRY(cudaEventRecord(c->evAuxReady[c->tileBuf], st)); c->auxReady[c->tileBuf] = true; }
        }
        if (earlyReadback) {   // depth and objectId are final now: their read-back overlaps the rest of the frame (rt_bind_readback)
            if (rec.err != cudaSuccess) return fail(RT_ERR_CUDA, std::string("frame launch: ") + cudaGetErrorString(rec.err));
            CUDA_TRY(cudaEventRecord(c->evPrimaryDone, st));
            CUDA_TRY(cudaStreamWaitEvent(c->copyStream, c->evPrimaryDone, 0));
            if (c->rbHost[1]) CUDA_TRY(cudaMemcpyAsync(c->rbHost[1], c->depth.p, c->rbBytes[1], cudaMemcpyDeviceToHost, c->copyStream));
            if (c->rbHost[2]) CUDA_TRY(cudaMemcpyAsync(c->rbHost[2], c->objId.p, c->rbBytes[2], cudaMemcpyDeviceToHost, c->copyStream));
            CUDA_TRY(cudaEventRecord(c->evCopyDone, c->copyStream));
            c->copyPending = true;
        }
        if (reuseDist) {   // every rank's G-buffer segment to every other rank (SpatialCompatible compares the CURRENT frame's G-buffer at both pixels)
            const size_t g = (size_t)cfg->width * cfg->height;
            float4* arrays[3] = {c->gbAll.p, c->gbAll.p + g, c->gbAll.p + 2 * g};
            const int rcx = exchange_segments(c, arrays, 3);
            if (rcx != RT_OK) return rcx;
        }

        ShadowQueue shq = {c->shO.p, c->shD.p};
        // ---- shared sun probe: one any-hit ray per Lambert primary vertex facing the sun, instead of one per sample that selects it
        if (sunProbe) {
            int* sunCount = c->counters.p + 2;
            rec.launch(k_sun_generate, dim3(grid_for(c, npx, 256)), dim3(256), 0, fc, wb, shq, sunCount); c->launches++;
            ExtendArgs pa; memset(&pa, 0, sizeof(pa));
            pa.sc = c->ds; pa.rayO = shq.o; pa.rayD = shq.d; pa.count = sunCount; pa.work = sunCount + 1; pa.visSt = c->pathState.p; pa.stats = c->dstats.p; pa.statSlot = 3;
            CUDA_TRY(launch_extend<true>(c, rec, pa, count));
            rec.launch(k_sun_store, dim3(grid_for(c, npx, 256)), dim3(256), 0, wb, shq, (const int*)sunCount); c->launches++;
        }

        // ---- integrator: batches of S samples, one wavefront iteration per depth ---------------------------------
        for (int pass = 0; pass < nPasses; pass++) {
            const int s0 = pass * S, ns = std::min(S, spp - s0);
            const size_t nPaths = (size_t)npx * ns;
            int* ctr = c->counters.p + CH + (size_t)pass * (cfg->maxDepth + 1) * CS;
            int cur = 0;
            RayQueue nq = {c->qO[cur].p, c->qD[cur].p};
            {
                const int g1 = grid_for(c, nPaths, 256);
                if (reuse && fast) rec.launch(k_shade_first<true, true>, dim3(g1), dim3(256), 0, fc, wb, s0, (int)nPaths, nq, ctr + 0, shq, ctr + 1, c->dstats.p);
                else if (reuse) rec.launch(k_shade_first<true, false>, dim3(g1), dim3(256), 0, fc, wb, s0, (int)nPaths, nq, ctr + 0, shq, ctr + 1, c->dstats.p);
                else if (fast) rec.launch(k_shade_first<false, true>, dim3(g1), dim3(256), 0, fc, wb, s0, (int)nPaths, nq, ctr + 0, shq, ctr + 1, c->dstats.p);
                else rec.launch(k_shade_first<false, false>, dim3(g1), dim3(256), 0, fc, wb, s0, (int)nPaths, nq, ctr + 0, shq, ctr + 1, c->dstats.p);
            }
            c->launches++;
            for (int depth = 1; depth <= cfg->maxDepth; depth++) {
                int* prev = ctr + (size_t)(depth - 1) * CS;   // counts produced by the shade of depth-1
                int* mine = ctr + (size_t)depth * CS;
                RayQueue cq = {c->qO[cur].p, c->qD[cur].p};
                ExtendArgs sa; memset(&sa, 0, sizeof(sa));
                sa.sc = c->ds; sa.rayO = shq.o; sa.rayD = shq.d; sa.count = prev + 1; sa.work = prev + 3; sa.visSt = c->pathState.p; sa.stats = c->dstats.p; sa.statSlot = 2;
                ExtendArgs ca; memset(&ca, 0, sizeof(ca));
                ca.sc = c->ds; ca.rayO = cq.o; ca.rayD = cq.d; ca.count = prev + 0; ca.work = prev + 2; ca.hits = hq; ca.missSt = c->pathState.p; ca.stats = c->dstats.p; ca.statSlot = 1;
                CUDA_TRY(launch_extend_pair(c, rec, ca, sa, count));   // closest-hit and shadow rays of this depth in ONE persistent launch
                RayQueue nq2 = {c->qO[cur ^ 1].p, c->qD[cur ^ 1].p};
                // chunks of 4096 rays per block for big wavefronts; 1024 when that would leave SMs without a block (interactive frame sizes:
                // 827 K paths are 202 chunks of 4096 - one 8-warp block per SM - but 808 of 1024)
                const bool smallChunks = nPaths < (size_t)c->smCount * 8 * RT_SHADE_CHUNK;
                const int fineBelow = c->smCount * 8 * RT_SHADE_CHUNK;   // queues shorter than two chunks per resident block are scanned 1024 rays at a time
                const int shadeGrid = grid_for(c, (nPaths + (smallChunks ? 1024 : RT_SHADE_CHUNK) - 1) / (smallChunks ? 1024 : RT_SHADE_CHUNK), 1);
#define RT_LAUNCH_SHADE_NEXT(R, F, C, M) do { rec.launch(k_shade_next<R, F, C>, dim3(shadeGrid), dim3(256), 0, fc, c->ds, wb, depth, cq, hq, (const int*)(prev + 0), nq2, mine + 0, shq, mine + 1, c->envDumpCounters ? mine + 4 : (int*)nullptr, fineBelow, M); c->launches++; } while (0)
#define RT_LAUNCH_SHADE_NEXT_RF(C, M) do { if (reuse && fast) RT_LAUNCH_SHADE_NEXT(true, true, C, M); else if (reuse) RT_LAUNCH_SHADE_NEXT(true, false, C, M); else if (fast) RT_LAUNCH_SHADE_NEXT(false, true, C, M); else RT_LAUNCH_SHADE_NEXT(false, false, C, M); } while (0)
                if (smallChunks) RT_LAUNCH_SHADE_NEXT_RF(1024, 0);
                else { RT_LAUNCH_SHADE_NEXT_RF(RT_SHADE_CHUNK, 1); RT_LAUNCH_SHADE_NEXT_RF(1024, 2); }
#undef RT_LAUNCH_SHADE_NEXT_RF
#undef RT_LAUNCH_SHADE_NEXT
                cur ^= 1;
            }
            rec.launch(k_accumulate, dim3(grid_for(c, npx, 256)), dim3(256), 0, fc, wb, s0, ns, pass == nPasses - 1 ? 1 : 0); c->launches++;
        }
        if (reuseDist) {   // this frame's reservoirs of every rank's pixels to every rank: next frame's imports read any pixel's
            const size_t g = (size_t)cfg->width * cfg->height;
            float4* cur = c->resAB[cfg->frame & 1].p;
            k_res_pack<<<grid_for(c, npx, 256), 256, 0, st>>>(cur, g, c->pixelMap.p, npx, ownStart, c->resPack.p); c->launches++;
            float4* arrays[3] = {c->resPack.p, c->resPack.p + g, c->resPack.p + 2 * g};
            const int rcx = exchange_segments(c, arrays, 3);
            if (rcx != RT_OK) return rcx;
            k_res_unpack<<<grid_for(c, g, 256), 256, 0, st>>>(c->resPack.p, g, c->deintMap.p, ownStart, ownStart + (size_t)npx, cur); c->launches++;
        }
        if (c->extColor) {
            rec.launch(k_copy_color, dim3(grid_for(c, npx, 256)), dim3(256), 0, (const int*)c->rgba8.p, c->extColor, (const int*)c->pixelMap.p, npx); c->launches++;
        }
    }
    rec.memcpy_d2h(c->hstats, c->dstats.p, sizeof(DeviceStats));
    if (rec.err != cudaSuccess) return fail(rec.err == cudaErrorMemoryAllocation ? RT_ERR_OUT_OF_MEMORY : RT_ERR_CUDA, std::string("frame launch: ") + cudaGetErrorString(rec.err));
    CUDA_TRY(cudaGetLastError());
    return RT_OK;
    }();
    if (useGraph) {
        const bool ok = rcFrame == RT_OK && !rec.mismatch && (rec.mode != FrameRecorder::Update || rec.idx == c->frameNodes.size());
        if (rec.mode == FrameRecorder::Build) {
            if (ok) {
                const cudaError_t ei = cudaGraphInstantiate(&c->frameGraphExec, rec.graph, 0);
                if (ei != cudaSuccess) { cudaGraphDestroy(rec.graph); c->frameGraphExec = nullptr; return fail(RT_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei)); }
                c->frameGraph = rec.graph; c->frameKey = key; c->frameGraphBuilds++;
            } else { cudaGraphDestroy(rec.graph); c->frameNodes.clear(); c->frameFuncs.clear(); }
        } else if (!ok && c->frameGraphExec) {   // the sequence did not match the recorded one after all: drop the graph (the next frame rebuilds it)
            cudaGraphExecDestroy(c->frameGraphExec); c->frameGraphExec = nullptr;
            if (c->frameGraph) { cudaGraphDestroy(c->frameGraph); c->frameGraph = nullptr; }
        }
        if (rcFrame != RT_OK) return rcFrame;
        if (!ok) return fail(RT_ERR_INVALID_STATE, "rt_render: the frame's launch sequence changed under a recorded frame graph");
        CUDA_TRY(cudaGraphLaunch(c->frameGraphExec, st));
    } else if (rcFrame != RT_OK) return rcFrame;
    CUDA_TRY(cudaEventRecord(c->evStop, st));
    if (anyReadback && c->worldSize <= 1) {   // what is left of the bound read-backs, behind the frame on its own stream
        if (c->rbHost[0]) CUDA_TRY(cudaMemcpyAsync(c->rbHost[0], c->rgba8.p, c->rbBytes[0], cudaMemcpyDeviceToHost, st));
        if (!earlyReadback) {
            if (c->rbHost[1]) CUDA_TRY(cudaMemcpyAsync(c->rbHost[1], c->depth.p, c->rbBytes[1], cudaMemcpyDeviceToHost, st));
            if (c->rbHost[2]) CUDA_TRY(cudaMemcpyAsync(c->rbHost[2], c->objId.p, c->rbBytes[2], cudaMemcpyDeviceToHost, st));
        }
    }
#if RT_PHASE_STATS
    if (count) {
        unsigned long long h[32], z[32] = {0};
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpyFromSymbol(h, g_phase, sizeof(h)));
        CUDA_TRY(cudaMemcpyToSymbol(g_phase, z, sizeof(z)));
        for (int k = 0; k < 2; k++) {
            const unsigned long long* p = h + 16 * k;
            fprintf(stderr, "[phase %s] iters %llu  lanes holding a ray %.2f | node step 1: run in %.3f of iters, %.2f lanes; step 2: %.3f, %.2f lanes | prim phase: %.3f of iters, %.2f lanes, %.2f queued prims\n",
                    k ? "anyhit " : "closest", p[0], (double)p[7] / p[0], (double)p[3] / p[0], (double)p[1] / (p[3] ? p[3] : 1), (double)p[4] / p[0], (double)p[2] / (p[4] ? p[4] : 1),
                    (double)p[5] / p[0], (double)p[6] / (p[5] ? p[5] : 1), (double)p[8] / (p[5] ? p[5] : 1));
        }
    }
#endif
    c->rendered = true;
    return RT_OK;
}

RT_API int rt_sync(rt_ctx* c) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_sync: ctx is null");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->commStream) CUDA_TRY(cudaStreamSynchronize(c->commStream));   // a gather (and a present of the gathered image) behind the frame
    if (c->copyStream) CUDA_TRY(cudaStreamSynchronize(c->copyStream));   // bound read-backs (rt_bind_readback)
    return RT_OK;
}

static int buffer_info(rt_ctx* c, int which, size_t* bytes) {
    const size_t g = (size_t)c->width * c->height;
    switch (which) {
        case RT_BUF_RGBA8: case RT_BUF_OBJID: case RT_BUF_PRIM_ID: case RT_BUF_INST_ID: case RT_BUF_GB_MATID: *bytes = g * 4; return RT_OK;
        case RT_BUF_DEPTH: case RT_BUF_PRIMARY_T: *bytes = g * 4; return RT_OK;
        case RT_BUF_RADIANCE: case RT_BUF_ACCUM: *bytes = g * 16; return RT_OK;
        case RT_BUF_GB_WORLDPOS: case RT_BUF_GB_NORMAL: case RT_BUF_GB_BASECOLOR: *bytes = g * 12; return RT_OK;
        case RT_BUF_SEG_COUNT: case RT_BUF_TERM_CODE: *bytes = g * c->spp; return RT_OK;
        case RT_BUF_PATH_HASH: *bytes = g * c->spp * 4; return RT_OK;
        case RT_BUF_TILE_RADIANCE: *bytes = (size_t)c->npx * 16; return RT_OK;
        case RT_BUF_RESERVOIR: *bytes = g * sizeof(RtReservoir); return RT_OK;
        case RT_BUF_PRESENT: *bytes = (size_t)c->presentW * c->presentH * 4; return RT_OK;
        case RT_BUF_GATHERED_RGBA8: case RT_BUF_GATHERED_DEPTH: case RT_BUF_GATHERED_OBJID: *bytes = (size_t)c->gatheredW * c->gatheredH * 4; return RT_OK;
        case RT_BUF_GATHERED_RADIANCE: *bytes = (size_t)c->gatheredW * c->gatheredH * 16; return RT_OK;
    }
    return fail(RT_ERR_INVALID_ARGUMENT, "unknown buffer selector");
}

RT_API int rt_buffer_bytes(rt_ctx* c, int which, size_t* bytes) {
    if (!c || !bytes) return fail(RT_ERR_INVALID_ARGUMENT, "rt_buffer_bytes: null argument");
    if (!c->rendered) return fail(RT_ERR_INVALID_STATE, "rt_buffer_bytes: nothing rendered yet");
    return buffer_info(c, which, bytes);
}

static int gathered_ptr(rt_ctx* c, int which, const void** p);
RT_API int rt_get_device_buffer(rt_ctx* c, int which, void** devPtr, size_t* bytes) {
    if (!c || !devPtr || !bytes) return fail(RT_ERR_INVALID_ARGUMENT, "rt_get_device_buffer: null argument");
    if (!c->rendered) return fail(RT_ERR_INVALID_STATE, "rt_get_device_buffer: nothing rendered yet");
    int rc = buffer_info(c, which, bytes);
    if (rc != RT_OK) return rc;
    switch (which) {
        case RT_BUF_RGBA8: *devPtr = c->rgba8.p; return RT_OK;
        case RT_BUF_DEPTH: *devPtr = c->depth.p; return RT_OK;
        case RT_BUF_OBJID: *devPtr = c->objId.p; return RT_OK;
        case RT_BUF_RADIANCE: *devPtr = c->radiance.p; return RT_OK;
        case RT_BUF_ACCUM: *devPtr = c->accum.p; return RT_OK;
        case RT_BUF_TILE_RADIANCE: *devPtr = c->tileRadiance[c->tileBuf].p; return RT_OK;
        case RT_BUF_GATHERED_RGBA8: case RT_BUF_GATHERED_DEPTH: case RT_BUF_GATHERED_OBJID: case RT_BUF_GATHERED_RADIANCE: {
            const int rcg = gathered_ptr(c, which, const_cast<const void**>(devPtr));
            return rcg;
        }
        case RT_BUF_SEG_COUNT: if (!c->aovs) break; *devPtr = c->segOut.p; return RT_OK;
        case RT_BUF_TERM_CODE: if (!c->aovs) break; *devPtr = c->termOut.p; return RT_OK;
        case RT_BUF_PATH_HASH: if (!c->aovs) break; *devPtr = c->hashOut.p; return RT_OK;
        default: return fail(RT_ERR_UNSUPPORTED, "rt_get_device_buffer: this buffer is stored per owned pixel; use rt_download");
    }
    return fail(RT_ERR_INVALID_STATE, "rt_get_device_buffer: path AOVs were not requested (RT_FLAG_PATH_AOVS)");
}

// the gathered image on the root of the last rt_gather_frame (valid on its stream, commStream)
static int gathered_ptr(rt_ctx* c, int which, const void** p) {
    if (!c->gatheredValid) return fail(RT_ERR_INVALID_STATE, "no gathered image: this context was not the root of an rt_gather_frame yet");
    const bool aux = (c->gatheredWhat & RT_GATHER_DEPTH_OBJID) != 0, rad = (c->gatheredWhat & RT_GATHER_RADIANCE) != 0;
    switch (which) {
        case RT_BUF_GATHERED_RGBA8: *p = c->gRgba8.p; return RT_OK;
        case RT_BUF_GATHERED_DEPTH: if (!aux) break; *p = c->gDepth.p; return RT_OK;
        case RT_BUF_GATHERED_OBJID: if (!aux) break; *p = c->gObjId.p; return RT_OK;
        case RT_BUF_GATHERED_RADIANCE: if (!rad) break; *p = c->gRadiance.p; return RT_OK;
    }
    return fail(RT_ERR_INVALID_STATE, "the last rt_gather_frame did not gather this buffer (RT_GATHER_DEPTH_OBJID / RT_GATHER_RADIANCE)");
}
static int download_impl(rt_ctx* c, int which, void* dst, size_t bytes, bool wait);
RT_API int rt_download(rt_ctx* c, int which, void* dst, size_t bytes) { return download_impl(c, which, dst, bytes, true); }
RT_API int rt_download_async(rt_ctx* c, int which, void* dst, size_t bytes) { return download_impl(c, which, dst, bytes, false); }
static int download_impl(rt_ctx* c, int which, void* dst, size_t bytes, bool wait) {
    if (!c || !dst) return fail(RT_ERR_INVALID_ARGUMENT, "rt_download: null argument");
    if (!c->rendered) return fail(RT_ERR_INVALID_STATE, "rt_download: nothing rendered yet");
    size_t need = 0;
    int rc = buffer_info(c, which, &need);
    if (rc != RT_OK) return rc;
    if (bytes != need) return fail(RT_ERR_INVALID_ARGUMENT, "rt_download: bytes does not match the buffer size");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const void* src = nullptr;
    const int npx = c->npx;
    const size_t g = (size_t)c->width * c->height;
    if (which >= RT_BUF_GATHERED_RGBA8 && which <= RT_BUF_GATHERED_RADIANCE) {   // stream-ordered behind the gather that made the image, before the next one
        rc = gathered_ptr(c, which, &src);
        if (rc != RT_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->commStream));
        if (wait) CUDA_TRY(cudaStreamSynchronize(c->commStream));
        return RT_OK;
    }
    if (!wait && !(which == RT_BUF_RGBA8 || which == RT_BUF_DEPTH || which == RT_BUF_OBJID || which == RT_BUF_RADIANCE || which == RT_BUF_ACCUM ||
                   which == RT_BUF_TILE_RADIANCE || which == RT_BUF_PRESENT))
        return fail(RT_ERR_UNSUPPORTED, "rt_download_async: this buffer is gathered through a shared staging buffer; use rt_download");
    auto scatterPrep = [&](size_t floats) -> cudaError_t {
        cudaError_t e = c->scratch.ensure(floats);
        if (e != cudaSuccess) return e;
        return cudaMemsetAsync(c->scratch.p, 0, floats * sizeof(float), st);
    };
    switch (which) {
        case RT_BUF_RGBA8: src = c->rgba8.p; break;
        case RT_BUF_DEPTH: src = c->depth.p; break;
        case RT_BUF_OBJID: src = c->objId.p; break;
        case RT_BUF_RADIANCE: src = c->radiance.p; break;
        case RT_BUF_ACCUM: src = c->accum.p; break;
        case RT_BUF_TILE_RADIANCE: src = c->tileRadiance[c->tileBuf].p; break;
        case RT_BUF_SEG_COUNT: case RT_BUF_TERM_CODE: case RT_BUF_PATH_HASH:
            if (!c->aovs) return fail(RT_ERR_INVALID_STATE, "rt_download: path AOVs were not requested (RT_FLAG_PATH_AOVS)");
            src = which == RT_BUF_SEG_COUNT ? (const void*)c->segOut.p : (which == RT_BUF_TERM_CODE ? (const void*)c->termOut.p : (const void*)c->hashOut.p);
            break;
        case RT_BUF_PRESENT:
            if (!c->presentPtr) return fail(RT_ERR_INVALID_STATE, "rt_download: rt_present has not run yet");
            src = c->presentPtr;
            if (c->presentOnComm) st = c->commStream;   // the present of a gathered frame ran behind the gather
            break;
        case RT_BUF_RESERVOIR:
            if (c->resLastWritten < 0) return fail(RT_ERR_INVALID_STATE, "rt_download: no frame with a reuse flag set has written reservoirs yet");
            CUDA_TRY(scatterPrep(g * 11));
            k_pack_reservoirs<<<grid_for(c, g, 256), 256, 0, st>>>(c->resAB[c->resLastWritten].p, (int)g, c->scratch.p);
            src = c->scratch.p; break;
        case RT_BUF_PRIM_ID: CUDA_TRY(scatterPrep(g)); if (npx) k_scatter<int><<<grid_for(c, npx, 256), 256, 0, st>>>(c->primId.p, (int*)c->scratch.p, c->pixelMap.p, npx); src = c->scratch.p; break;
        case RT_BUF_INST_ID: CUDA_TRY(scatterPrep(g)); if (npx) k_scatter<int><<<grid_for(c, npx, 256), 256, 0, st>>>(c->instId.p, (int*)c->scratch.p, c->pixelMap.p, npx); src = c->scratch.p; break;
        case RT_BUF_PRIMARY_T: CUDA_TRY(scatterPrep(g)); if (npx) k_scatter<float><<<grid_for(c, npx, 256), 256, 0, st>>>(c->primaryT.p, c->scratch.p, c->pixelMap.p, npx); src = c->scratch.p; break;
        case RT_BUF_GB_WORLDPOS: CUDA_TRY(scatterPrep(g * 3)); if (npx) k_scatter_f4_to_f3<<<grid_for(c, npx, 256), 256, 0, st>>>(const_cast<float4*>(c->gbPosPtr), c->scratch.p, c->pixelMap.p, npx); src = c->scratch.p; break;
        case RT_BUF_GB_NORMAL: CUDA_TRY(scatterPrep(g * 3)); if (npx) k_scatter_f4_to_f3<<<grid_for(c, npx, 256), 256, 0, st>>>(const_cast<float4*>(c->gbNrmPtr), c->scratch.p, c->pixelMap.p, npx); src = c->scratch.p; break;
        case RT_BUF_GB_BASECOLOR: CUDA_TRY(scatterPrep(g * 3)); if (npx) k_scatter_f4_to_f3<<<grid_for(c, npx, 256), 256, 0, st>>>(const_cast<float4*>(c->gbAlbPtr), c->scratch.p, c->pixelMap.p, npx); src = c->scratch.p; break;
        case RT_BUF_GB_MATID: CUDA_TRY(scatterPrep(g)); if (npx) k_scatter_f4_w<<<grid_for(c, npx, 256), 256, 0, st>>>(const_cast<float4*>(c->gbNrmPtr), (int*)c->scratch.p, c->pixelMap.p, npx); src = c->scratch.p; break;
        default: return fail(RT_ERR_INVALID_ARGUMENT, "rt_download: unknown buffer selector");
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    if (wait) CUDA_TRY(cudaStreamSynchronize(st));
    return RT_OK;
}

RT_API int rt_bind_readback(rt_ctx* c, int which, void* hostPinned, size_t bytes) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_bind_readback: ctx is null");
    if (which != RT_BUF_RGBA8 && which != RT_BUF_DEPTH && which != RT_BUF_OBJID) return fail(RT_ERR_INVALID_ARGUMENT, "rt_bind_readback: only RT_BUF_RGBA8, RT_BUF_DEPTH and RT_BUF_OBJID can be bound");
    if (hostPinned && bytes == 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_bind_readback: zero-sized target");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // no frame in flight may still write the old target
    if (c->copyStream) CUDA_TRY(cudaStreamSynchronize(c->copyStream));
    if (c->commStream) CUDA_TRY(cudaStreamSynchronize(c->commStream));
    if (hostPinned && !c->copyStream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c->evPrimaryDone, cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&c->evCopyDone, cudaEventDisableTiming));
    }
    const int slot = which == RT_BUF_RGBA8 ? 0 : (which == RT_BUF_DEPTH ? 1 : 2);
    c->rbHost[slot] = hostPinned; c->rbBytes[slot] = hostPinned ? bytes : 0;
    return RT_OK;
}

RT_API int rt_map_external_color(rt_ctx* c, void* devPtr, size_t bytes) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_map_external_color: ctx is null");
    c->extColor = (int*)devPtr; c->extColorBytes = devPtr ? bytes : 0;
    return RT_OK;
}

RT_API int rt_present(rt_ctx* c, const RtPresentConfig* pc, void* dstDevRgba8, size_t dstBytes) {
    if (!c || !pc) return fail(RT_ERR_INVALID_ARGUMENT, "rt_present: null argument");
    if (!c->rendered) return fail(RT_ERR_INVALID_STATE, "rt_present: nothing rendered yet");
    // a frame rendered as one rank of a tile partition is presented from the image rt_gather_frame assembled on its root
    const bool fromGather = c->worldSize > 1;
    if (fromGather && !c->gatheredValid) return fail(RT_ERR_INVALID_STATE, "rt_present: the last frame was rendered as one rank of a tile partition; call rt_gather_frame and present on its root");
    if (fromGather && pc->mode == RT_PRESENT_TAAU && !(c->gatheredWhat & RT_GATHER_DEPTH_OBJID))
        return fail(RT_ERR_INVALID_STATE, "rt_present: the TAAU resolve needs objectId (Engine/RTTaa.cs:117-171): gather with RT_GATHER_DEPTH_OBJID");
    if (pc->outWidth <= 0 || pc->outHeight <= 0 || (int64_t)pc->outWidth * pc->outHeight > 0x7FFFFFFF) return fail(RT_ERR_INVALID_ARGUMENT, "rt_present: bad output size");
    if (pc->mode != RT_PRESENT_TAAU && pc->mode != RT_PRESENT_COPY) return fail(RT_ERR_INVALID_ARGUMENT, "rt_present: unknown mode");
    const size_t outLen = (size_t)pc->outWidth * pc->outHeight;
    if (dstDevRgba8 && dstBytes < outLen * sizeof(int)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_present: destination is smaller than the output image");   // Framebuffer.cs:117
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = fromGather ? c->commStream : c->stream;   // behind the gather that made the image
    int* dst = (int*)dstDevRgba8;
    if (!dst) { CUDA_TRY(c->presentBuf.ensure(outLen)); dst = c->presentBuf.p; }
    const int inW = c->width, inH = c->height;
    const int* srcRgba8 = fromGather ? c->gRgba8.p : c->rgba8.p;
    const int* srcObjId = fromGather ? c->gObjId.p : c->objId.p;
    if (pc->mode == RT_PRESENT_TAAU) {
        if (c->taaHistColor.n != outLen || !c->taaHistColor.p) {   // RTTaa.Ensure (RTTaa.cs:34-47): a new size drops the history
            c->taaHistColor.release(); c->taaHistObj.release();
            CUDA_TRY(c->taaHistColor.ensure(outLen)); CUDA_TRY(c->taaHistObj.ensure(outLen));
            CUDA_TRY(cudaMemsetAsync(c->taaHistColor.p, 0, outLen * sizeof(int), st)); CUDA_TRY(cudaMemsetAsync(c->taaHistObj.p, 0, outLen * sizeof(int), st));
            c->taaHistoryValid = false;
        }
        if (pc->resetHistory) c->taaHistoryValid = false;
        TaaConst tc; tc.outW = pc->outWidth; tc.outH = pc->outHeight; tc.inW = inW; tc.inH = inH;
        tc.feedback = pc->feedback; tc.sharpness = pc->sharpness; tc.clampK = pc->clampK; tc.isFirstFrame = c->taaHistoryValid ? 0 : 1;
        k_taa_resolve<<<grid_for(c, outLen, 256), 256, 0, st>>>(tc, srcRgba8, srcObjId, c->taaHistColor.p, c->taaHistObj.p, dst);
        c->taaHistoryValid = true;
    } else if (inW == pc->outWidth && inH == pc->outHeight) {
        CUDA_TRY(cudaMemcpyAsync(dst, srcRgba8, outLen * sizeof(int), cudaMemcpyDeviceToDevice, st));   // BlitKernel
    } else {
        k_bilinear_upsample<<<grid_for(c, outLen, 256), 256, 0, st>>>(srcRgba8, inW, inH, dst, pc->outWidth, pc->outHeight);
    }
    CUDA_TRY(cudaGetLastError());
    c->presentW = pc->outWidth; c->presentH = pc->outHeight; c->presentPtr = dst; c->presentOnComm = fromGather;
    return RT_OK;
}

RT_API int rt_tiles_owned_pixels(int width, int height, int tileSize, int rank, int worldSize, int64_t* nPixels) {
    if (!nPixels || width <= 0 || height <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_tiles_owned_pixels: bad argument");
    if (worldSize > 1 && (rank < 0 || rank >= worldSize)) return fail(RT_ERR_INVALID_ARGUMENT, "rt_tiles_owned_pixels: rank outside [0, worldSize)");
    *nPixels = count_owned_pixels(width, height, tileSize, worldSize > 1 ? rank : 0, worldSize > 1 ? worldSize : 1);
    return RT_OK;
}

// every rank's owned-pixel list, concatenated in rank order; built once per (image, tile, world) and kept on the device
static int ensure_deint_map(rt_ctx* c, int width, int height, int T, int worldSize) {
    if (c->deintW != width || c->deintH != height || c->deintT != T || c->deintWorld != worldSize) {
        std::vector<int> all; all.reserve((size_t)width * height);
        c->deintStart.assign((size_t)worldSize + 1, 0);
        for (int r = 0; r < worldSize; r++) {
            std::vector<int> m; build_pixel_map(width, height, T, r, worldSize, m);
            all.insert(all.end(), m.begin(), m.end());
            c->deintStart[(size_t)r + 1] = (int64_t)all.size();
        }
        CUDA_TRY(c->deintMap.ensure(std::max<size_t>(1, all.size())));
        CUDA_TRY(cudaMemcpyAsync(c->deintMap.p, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->deintW = width; c->deintH = height; c->deintT = T; c->deintWorld = worldSize;
        c->deintHost.swap(all);
    }
    return RT_OK;
}

RT_API int rt_deinterleave_tiles(rt_ctx* c, const void* gatheredDev, const int64_t* rankOffsetsPx, int worldSize, int width, int height, int tileSize,
                                 void* outRadianceDev, void* outRgba8Dev) {
    if (!c || !gatheredDev || !rankOffsetsPx || worldSize < 1 || width <= 0 || height <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_deinterleave_tiles: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const int T = effective_tile_size(tileSize);
    const int rcMap = ensure_deint_map(c, width, height, T, worldSize);
    if (rcMap != RT_OK) return rcMap;
    for (int r = 0; r < worldSize; r++) {
        const int64_t n = c->deintStart[(size_t)r + 1] - c->deintStart[(size_t)r];
        if (n <= 0) continue;
        k_deinterleave<<<grid_for(c, (size_t)n, 256), 256, 0, c->stream>>>((const float4*)gatheredDev + rankOffsetsPx[r], c->deintMap.p + c->deintStart[(size_t)r], (int)n,
                                                                          (float4*)outRadianceDev, (int*)outRgba8Dev);
    }
    CUDA_TRY(cudaGetLastError());
    return RT_OK;
}


}   // extern "C"

// ------------------------------------------------------------------------------------------------ multi-GPU behind the ABI (entry points)
extern "C" {

RT_API int rt_comm_get_unique_id(void* id, size_t bytes) {
    if (!id || bytes != RT_COMM_ID_BYTES) return fail(RT_ERR_INVALID_ARGUMENT, "rt_comm_get_unique_id: id must point at RT_COMM_ID_BYTES (128) bytes");
    static_assert(sizeof(ncclUniqueId) == RT_COMM_ID_BYTES, "ncclUniqueId size");
    NcclApi* nc = nccl_api();
    if (!nc) return fail(RT_ERR_UNSUPPORTED, nccl_err());
    ncclUniqueId u;
    NCCL_TRY(nc->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return RT_OK;
}

RT_API int rt_comm_destroy(rt_ctx* c) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_comm_destroy: ctx is null");
    cudaSetDevice(c->device);
    if (c->commStream) cudaStreamSynchronize(c->commStream);
    if (c->comm) { NcclApi* nc = nccl_api(); if (nc) nc->CommDestroy(c->comm); c->comm = nullptr; }
    for (int b = 0; b < 2; b++) {
        if (c->evTileReady[b]) cudaEventDestroy(c->evTileReady[b]);
        if (c->evGatherDone[b]) cudaEventDestroy(c->evGatherDone[b]);
        if (c->evAuxReady[b]) cudaEventDestroy(c->evAuxReady[b]);
        c->evTileReady[b] = c->evGatherDone[b] = c->evAuxReady[b] = nullptr; c->gatherPending[b] = false; c->auxReady[b] = false;
    }
    if (c->evGatherStart) cudaEventDestroy(c->evGatherStart);
    if (c->evGatherStop) cudaEventDestroy(c->evGatherStop);
    c->evGatherStart = c->evGatherStop = nullptr;
    if (c->commStream) cudaStreamDestroy(c->commStream);
    c->commStream = nullptr; c->commRank = 0; c->commWorld = 1; c->gatheredValid = false; c->gatherTimed = false;
    c->gatherStage.release(); c->gRgba8.release(); c->gObjId.release(); c->gDepth.release(); c->gRadiance.release();
    return RT_OK;
}

RT_API int rt_comm_init(rt_ctx* c, const void* id, size_t bytes, int rank, int worldSize) {
    if (!c || !id || bytes != RT_COMM_ID_BYTES) return fail(RT_ERR_INVALID_ARGUMENT, "rt_comm_init: null context / id, or id is not RT_COMM_ID_BYTES (128) bytes");
    if (worldSize < 1 || rank < 0 || rank >= worldSize) return fail(RT_ERR_INVALID_ARGUMENT, "rt_comm_init: rank outside [0, worldSize)");
    if (c->comm) return fail(RT_ERR_INVALID_STATE, "rt_comm_init: the context already has a communicator (rt_comm_destroy first)");
    NcclApi* nc = nccl_api();
    if (!nc) return fail(RT_ERR_UNSUPPORTED, nccl_err());
    CUDA_TRY(cudaSetDevice(c->device));
    ncclUniqueId u; memcpy(&u, id, sizeof(u));
    NCCL_TRY(nc->CommInitRank(&c->comm, worldSize, u, rank));
    c->commRank = rank; c->commWorld = worldSize;
    const int rc = [&]() -> int {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->commStream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) { CUDA_TRY(cudaEventCreateWithFlags(&c->evTileReady[b], cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&c->evGatherDone[b], cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&c->evAuxReady[b], cudaEventDisableTiming)); }
        CUDA_TRY(cudaEventCreate(&c->evGatherStart)); CUDA_TRY(cudaEventCreate(&c->evGatherStop));
        return RT_OK;
    }();
    if (rc != RT_OK) { const std::string msg = g_lastError; rt_comm_destroy(c); g_lastError = msg; }
    return rc;
}

// One exchange per displayed frame: every rank sends the payloads of ITS tiles (what the last rt_render wrote, straight from the
// buffers the accumulate / primary-finish kernels filled, exact counts, no staging copy) to `root`; the root receives them into a
// staging buffer and scatters all of them - its own included - into the gathered image (de-interleave + PackRGBA8 fused).  Runs
// on the communicator's own stream behind the frame, so the next rt_render overlaps it.
RT_API int rt_gather_frame(rt_ctx* c, int root, uint32_t what) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_gather_frame: ctx is null");
    if (!c->comm) return fail(RT_ERR_INVALID_STATE, "rt_gather_frame: no communicator (rt_comm_init first)");
    if (!c->rendered) return fail(RT_ERR_INVALID_STATE, "rt_gather_frame: nothing rendered yet");
    if (root < 0 || root >= c->commWorld) return fail(RT_ERR_INVALID_ARGUMENT, "rt_gather_frame: root outside [0, worldSize)");
    if ((what & ~(uint32_t)(RT_GATHER_RGBA8 | RT_GATHER_RADIANCE | RT_GATHER_DEPTH_OBJID)) != 0u || (what & (RT_GATHER_RGBA8 | RT_GATHER_RADIANCE)) == 0u)
        return fail(RT_ERR_INVALID_ARGUMENT, "rt_gather_frame: `what` must name RT_GATHER_RGBA8 or RT_GATHER_RADIANCE (optionally | RT_GATHER_DEPTH_OBJID)");
    if (c->worldSize != c->commWorld || c->rank != c->commRank)
        return fail(RT_ERR_INVALID_STATE, "rt_gather_frame: the last frame was not rendered as this communicator's rank (RtRenderConfig.rank / worldSize must equal rt_comm_init's)");
    NcclApi* nc = nccl_api();
    if (!nc) return fail(RT_ERR_UNSUPPORTED, nccl_err());
    CUDA_TRY(cudaSetDevice(c->device));
    const int W = c->width, H = c->height, world = c->commWorld, b = c->tileBuf;
    const size_t g = (size_t)W * H;
    const bool sendRad = (what & RT_GATHER_RADIANCE) != 0, sendRgba = !sendRad, sendAux = (what & RT_GATHER_DEPTH_OBJID) != 0;
    const bool isRoot = c->commRank == root;
    cudaStream_t cs = c->commStream;
    CUDA_TRY(cudaEventRecord(c->evTileReady[b], c->stream));
    // staging on the root: three regions (radiance | rgba | aux), each indexed like the concatenated owned-pixel lists
    unsigned char* stage = nullptr; size_t offRad = 0, offRgba = 0, offAux = 0;
    if (isRoot) {
        const int rcMap = ensure_deint_map(c, W, H, c->tileSize, world);
        if (rcMap != RT_OK) return rcMap;
        size_t bytes = 0;
        offRad = bytes; if (sendRad) bytes += g * sizeof(float4);
        offRgba = bytes; if (sendRgba) bytes += g * sizeof(int);
        offAux = bytes; if (sendAux) bytes += g * sizeof(uint2);
        if (world > 1) CUDA_TRY(c->gatherStage.ensure(bytes));
        stage = c->gatherStage.p;
        if (c->gatheredW != W || c->gatheredH != H || !c->gRgba8.p) {
            CUDA_TRY(c->gRgba8.ensure(g)); CUDA_TRY(c->gObjId.ensure(g)); CUDA_TRY(c->gDepth.ensure(g));
            CUDA_TRY(cudaMemsetAsync(c->gRgba8.p, 0, g * 4, cs)); CUDA_TRY(cudaMemsetAsync(c->gObjId.p, 0, g * 4, cs)); CUDA_TRY(cudaMemsetAsync(c->gDepth.p, 0, g * 4, cs));
            c->gatheredW = W; c->gatheredH = H;
        }
        if (sendRad && c->gRadiance.n < g) { CUDA_TRY(c->gRadiance.ensure(g)); CUDA_TRY(cudaMemsetAsync(c->gRadiance.p, 0, g * sizeof(float4), cs)); }
    }
    // Two phases, in the SAME order on every rank whatever each of them could overlap: (1) depth | objectId - final since the
    // primary pass, so when the frame recorded evAuxReady (plain launches) this phase, its scatter and the root's bound depth /
    // objectId read-backs run while the frame still renders; (2) colour, behind the end of the frame.
    struct Seg { size_t start, n; bool own; };
    const size_t s0 = isRoot ? (size_t)c->deintStart[(size_t)root] : 0, s1 = isRoot ? (size_t)c->deintStart[(size_t)root + 1] : 0;
    const Seg segs[3] = {{0, s0, false}, {s0, s1 - s0, true}, {s1, g - s1, false}};   // root: the ranks below it and those above it are contiguous in the staging regions; its own payload is read in place
    for (int phase = 0; phase < 2; phase++) {
        const bool aux = phase == 0;
        if (aux && !sendAux) continue;
        CUDA_TRY(cudaStreamWaitEvent(cs, (aux && c->auxReady[b]) ? c->evAuxReady[b] : c->evTileReady[b], 0));
        if (!aux) CUDA_TRY(cudaEventRecord(c->evGatherStart, cs));   // timed: what the gather adds behind the frame
        if (world > 1) {
            NCCL_TRY(nc->GroupStart());
            if (!isRoot) {
                const size_t n = (size_t)c->npx;
                if (n > 0) {
                    if (aux) NCCL_TRY(nc->Send(c->tileAux[b].p, n * 2, ncclInt32, root, c->comm, cs));
                    else if (sendRad) NCCL_TRY(nc->Send(c->tileRadiance[b].p, n * 4, ncclFloat, root, c->comm, cs));
                    else NCCL_TRY(nc->Send(c->tileRgba[b].p, n, ncclInt32, root, c->comm, cs));
                }
            } else {
                for (int r = 0; r < world; r++) {
                    if (r == root) continue;
                    const size_t start = (size_t)c->deintStart[(size_t)r], n = (size_t)(c->deintStart[(size_t)r + 1] - c->deintStart[(size_t)r]);
                    if (n == 0) continue;
                    if (aux) NCCL_TRY(nc->Recv(reinterpret_cast<uint2*>(stage + offAux) + start, n * 2, ncclInt32, r, c->comm, cs));
                    else if (sendRad) NCCL_TRY(nc->Recv(reinterpret_cast<float4*>(stage + offRad) + start, n * 4, ncclFloat, r, c->comm, cs));
                    else NCCL_TRY(nc->Recv(reinterpret_cast<int*>(stage + offRgba) + start, n, ncclInt32, r, c->comm, cs));
                }
            }
            NCCL_TRY(nc->GroupEnd());
        }
        if (isRoot) {
            for (const Seg& sg : segs) {
                if (sg.n == 0) continue;
                const float4* rad = (aux || !sendRad) ? nullptr : (sg.own ? c->tileRadiance[b].p : reinterpret_cast<const float4*>(stage + offRad) + sg.start);
                const int* rgba = (aux || !sendRgba) ? nullptr : (sg.own ? c->tileRgba[b].p : reinterpret_cast<const int*>(stage + offRgba) + sg.start);
                const uint2* ax = !aux ? nullptr : (sg.own ? c->tileAux[b].p : reinterpret_cast<const uint2*>(stage + offAux) + sg.start);
                k_gather_scatter<<<grid_for(c, sg.n, 256), 256, 0, cs>>>(rad, rgba, ax, c->deintMap.p + sg.start, (int)sg.n, c->gRadiance.p, c->gRgba8.p, c->gDepth.p, c->gObjId.p);
            }
            CUDA_TRY(cudaGetLastError());
            // bound read-backs on the root: the GATHERED colour / depth / objectId, behind their scatter on the communicator's stream
            if (aux && c->rbHost[1] && c->rbBytes[1] == g * 4) CUDA_TRY(cudaMemcpyAsync(c->rbHost[1], c->gDepth.p, g * 4, cudaMemcpyDeviceToHost, cs));
            if (aux && c->rbHost[2] && c->rbBytes[2] == g * 4) CUDA_TRY(cudaMemcpyAsync(c->rbHost[2], c->gObjId.p, g * 4, cudaMemcpyDeviceToHost, cs));
            if (!aux && c->rbHost[0] && c->rbBytes[0] == g * 4) CUDA_TRY(cudaMemcpyAsync(c->rbHost[0], c->gRgba8.p, g * 4, cudaMemcpyDeviceToHost, cs));
        }
    }
    if (isRoot) { c->gatheredValid = true; c->gatheredWhat = what; }
    CUDA_TRY(cudaEventRecord(c->evGatherStop, cs));
    CUDA_TRY(cudaEventRecord(c->evGatherDone[b], cs));
    c->gatherPending[b] = true; c->gatherTimed = true;
    return RT_OK;
}

RT_API int rt_get_stats(rt_ctx* c, RtStats* out) {
    if (!c || !out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_get_stats: null argument");
    memset(out, 0, sizeof(*out));
    out->bvhWideNodeCount = (uint64_t)c->bvhStats.nWideNodes; out->bvhPrimCount = (uint64_t)c->bvhStats.nPrims; out->bvhBytes = (uint64_t)c->bvhBytes;
    if (!c->rendered) return RT_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    out->raysPrimary = c->hstats->raysPrimary; out->raysBounce = c->hstats->raysBounce;
    out->raysShadow = c->hstats->raysShadow + c->hstats->shadowProbed;   // ShadowOcclusion calls of the reference: own traces + those answered by a shared sun probe
    out->reserved[1] = c->hstats->raysShadow + c->hstats->raysSunProbe;   // any-hit rays actually traced
    out->reserved[2] = c->hstats->raysSunProbe;
    out->wideNodes = c->hstats->wideNodes; out->trisTested = c->hstats->tris; out->spheresTested = c->hstats->spheres;
    out->kernelLaunches = c->launches;
    if (c->envDumpCounters && c->ctrStride) {   // tuning: queue sizes by depth of the frame just rendered
        std::vector<int> h((size_t)c->ctrHeader + (size_t)c->ctrPasses * c->ctrDepths * c->ctrStride);
        if (h.size() <= c->counters.n && cudaMemcpy(h.data(), c->counters.p, h.size() * sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int p = 0; p < c->ctrPasses; p++) for (int d = 0; d < c->ctrDepths; d++) {
                const int* q = h.data() + c->ctrHeader + ((size_t)p * c->ctrDepths + d) * c->ctrStride;
                fprintf(stderr, "rtcore_b200 counters: pass %d after depth %d: closest rays %d shadow rays %d | hits shaded at this depth %d\n", p, d, q[0], q[1], q[4]);
            }
    }
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, c->evStart, c->evStop) == cudaSuccess) out->lastRenderMs = ms;
    float tr = 0.0f;
    for (size_t i = 0; i + 1 < c->traceEventsUsed; i += 2) { float m = 0.0f; if (cudaEventElapsedTime(&m, c->traceEvents[i], c->traceEvents[i + 1]) == cudaSuccess) tr += m; }
    out->lastTraceMs = tr;
    out->reserved[0] = (uint64_t)(c->traceEventsUsed / 2);   // number of extend launches timed
    if (c->gatherTimed && c->commStream) {
        CUDA_TRY(cudaStreamSynchronize(c->commStream));
        float gm = 0.0f;
        if (cudaEventElapsedTime(&gm, c->evGatherStart, c->evGatherStop) == cudaSuccess) out->reserved[3] = (uint64_t)(gm * 1000.0f);
    }
    return RT_OK;
}

}   // extern "C"

// ------------------------------------------------------------------------------------------------ CUDA-GL interop (present target)
// Replaces the reference's own driver-API binding (Engine/CudaGlInteropIndexBuffer.cs:18-34: DllImport "nvcuda" of
// cuGraphicsGLRegisterBuffer / MapResources / GetMappedPointer_v2 / UnmapResources / UnregisterResource) and the register / map /
// unmap logic of the PBO class (:44-60, :62-99): the host keeps creating the GL PixelUnpackBuffer (GL.GenBuffer / BufferData stay in
// C#) and hands its name over; map / unmap run on the CONTEXT's stream, so rt_present into the mapped pointer is ordered between them.
// The five entry points come from the driver library already in the process (libcuda.so.1), resolved at first use.
struct GlInteropApi {
    bool ok = false; std::string err;
    int (*Register)(void**, unsigned, unsigned) = nullptr; int (*Unregister)(void*) = nullptr;
    int (*Map)(unsigned, void**, void*) = nullptr; int (*Unmap)(unsigned, void**, void*) = nullptr;
    int (*GetPtr)(unsigned long long*, size_t*, void*) = nullptr; int (*ErrName)(int, const char**) = nullptr;
};
static GlInteropApi* gl_api() {
    static GlInteropApi api; static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) { api.err = std::string("CUDA driver library not found: ") + dlerror(); return nullptr; }
    bool ok = true;
    auto sym = [&](const char* n) -> void* { void* p = dlsym(h, n); if (!p) { ok = false; api.err = std::string("driver symbol missing: ") + n; } return p; };
    api.Register = (decltype(api.Register))sym("cuGraphicsGLRegisterBuffer"); api.Unregister = (decltype(api.Unregister))sym("cuGraphicsUnregisterResource");
    api.Map = (decltype(api.Map))sym("cuGraphicsMapResources"); api.Unmap = (decltype(api.Unmap))sym("cuGraphicsUnmapResources");
    api.GetPtr = (decltype(api.GetPtr))sym("cuGraphicsResourceGetMappedPointer_v2"); api.ErrName = (decltype(api.ErrName))sym("cuGetErrorName");
    api.ok = ok;
    return ok ? &api : nullptr;
}
static int gl_fail(GlInteropApi* g, const char* what, int cuErr) {
    const char* name = nullptr;
    if (g->ErrName) g->ErrName(cuErr, &name);
    return fail(RT_ERR_CUDA, std::string(what) + ": " + (name ? name : "CUDA driver error") + " (" + std::to_string(cuErr) + ")");
}

extern "C" {

RT_API int rt_gl_register_buffer(rt_ctx* c, unsigned int glBuffer, void** resource) {
    if (!c || !resource) return fail(RT_ERR_INVALID_ARGUMENT, "rt_gl_register_buffer: null argument");
    *resource = nullptr;
    GlInteropApi* g = gl_api();
    if (!g) return fail(RT_ERR_UNSUPPORTED, "rt_gl_register_buffer: the CUDA driver's GL interop entry points are not available");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaFree(nullptr));   // the driver calls below need this device's primary context current on the calling (GL) thread
    const int e = g->Register(resource, glBuffer, 2u /* CU_GRAPHICS_REGISTER_FLAGS_WRITE_DISCARD, CudaGlInteropIndexBuffer.cs:55-56 */);
    if (e != 0) { *resource = nullptr; return gl_fail(g, "cuGraphicsGLRegisterBuffer (is a GL context current on this thread, on this GPU?)", e); }
    return RT_OK;
}
RT_API int rt_gl_map(rt_ctx* c, void* resource, void** devPtr, size_t* bytes) {
    if (!c || !resource || !devPtr || !bytes) return fail(RT_ERR_INVALID_ARGUMENT, "rt_gl_map: null argument");
    GlInteropApi* g = gl_api();
    if (!g) return fail(RT_ERR_UNSUPPORTED, "rt_gl_map: GL interop is not available");
    CUDA_TRY(cudaSetDevice(c->device));
    void* res = resource;
    int e = g->Map(1u, &res, (void*)c->stream);            // MapCuda, CudaGlInteropIndexBuffer.cs:62-74 (on the context's stream)
    if (e != 0) return gl_fail(g, "cuGraphicsMapResources", e);
    unsigned long long p = 0; size_t n = 0;
    e = g->GetPtr(&p, &n, resource);                        // GetCudaArrayView, :76-88
    if (e != 0) { g->Unmap(1u, &res, (void*)c->stream); return gl_fail(g, "cuGraphicsResourceGetMappedPointer", e); }
    *devPtr = (void*)(uintptr_t)p; *bytes = n;
    return RT_OK;
}
RT_API int rt_gl_unmap(rt_ctx* c, void* resource) {
    if (!c || !resource) return fail(RT_ERR_INVALID_ARGUMENT, "rt_gl_unmap: null argument");
    GlInteropApi* g = gl_api();
    if (!g) return fail(RT_ERR_UNSUPPORTED, "rt_gl_unmap: GL interop is not available");
    CUDA_TRY(cudaSetDevice(c->device));
    void* res = resource;
    const int e = g->Unmap(1u, &res, (void*)c->stream);    // UnmapCuda, :90-103
    if (e != 0) return gl_fail(g, "cuGraphicsUnmapResources", e);
    return RT_OK;
}
RT_API int rt_gl_unregister(rt_ctx* c, void* resource) {
    if (!c) return fail(RT_ERR_INVALID_ARGUMENT, "rt_gl_unregister: ctx is null");
    if (!resource) return RT_OK;
    GlInteropApi* g = gl_api();
    if (!g) return fail(RT_ERR_UNSUPPORTED, "rt_gl_unregister: GL interop is not available");
    CUDA_TRY(cudaSetDevice(c->device));
    const int e = g->Unregister(resource);                  // DisposeAcceleratorObject, :150-160
    if (e != 0) return gl_fail(g, "cuGraphicsUnregisterResource", e);
    return RT_OK;
}

}   // extern "C"

// ------------------------------------------------------------------------------------------------ ABI layout checks
// The element layouts are the reference's device layouts; C# mirrors them with
// [StructLayout(LayoutKind.Sequential)] (csharp/RtNative.cs).
static_assert(sizeof(RtFloat3) == 12 && sizeof(RtFloat2) == 8 && sizeof(RtAffine3x4) == 48, "ABI layout");
static_assert(sizeof(RtBvhNode) == 44 && sizeof(RtInstanceRecord) == 144 && sizeof(RtMaterialRecord) == 44 && sizeof(RtSphere) == 80, "ABI layout");
static_assert(sizeof(RtMeshTri) == 12 && sizeof(RtMeshTriUV) == 12 && sizeof(RtRGBA32) == 4 && sizeof(RtTexInfo) == 12 && sizeof(RtCamera) == 92, "ABI layout");
static_assert(sizeof(RtSceneDesc) == 15 * 16 && sizeof(RtRenderConfig) == 32 + 48 + 4 + 12 + 4 + 12, "ABI layout");
static_assert(sizeof(RtPresentConfig) == 44, "ABI layout");
static_assert(sizeof(RtReservoir) == 44, "ABI layout");
static_assert(sizeof(WideNode) == 80 && sizeof(PrimRec) == 48 && sizeof(HitRec) == 16, "device layout");
