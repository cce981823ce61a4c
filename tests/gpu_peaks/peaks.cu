// peaks.cu - TEST / MEASUREMENT TOOL (not a product path): micro-benchmarks for the roofline denominators SURVEY.md section 8d
// asks to MEASURE on the box instead of deriving them: FP32 issue (FFMA, packed FFMA2, separate FMUL + FADD as the exact
// arithmetic of the primitive tests and the shading runs), plain ALU issue, and L2 read bandwidth at the working-set size of
// the 1M-triangle wide BVH (58 MB, L2 resident on B200's 126 MB).  Driven by tests/gpu_peaks.py -> profiles/rNN_gpu_peaks.json.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define PK_API extern "C" __attribute__((visibility("default")))

// 8 independent chains per thread, ITER x 8 x UNROLL operations
template <int MODE> __global__ void __launch_bounds__(256) k_fp32(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) {          // FFMA
                x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
                x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
            } else if (MODE == 1) {   // FFMA2 (fma.rn.f32x2): two lanes' worth per instruction
                asm volatile("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%8, %8};\n\tmov.b64 rc, {%9, %9};\n\tfma.rn.f32x2 ra, ra, rb, rc;\n\tmov.b64 {%0, %1}, ra;\n\t"
                             "mov.b64 ra, {%2, %3};\n\tfma.rn.f32x2 ra, ra, rb, rc;\n\tmov.b64 {%2, %3}, ra;\n\t"
                             "mov.b64 ra, {%4, %5};\n\tfma.rn.f32x2 ra, ra, rb, rc;\n\tmov.b64 {%4, %5}, ra;\n\t"
                             "mov.b64 ra, {%6, %7};\n\tfma.rn.f32x2 ra, ra, rb, rc;\n\tmov.b64 {%6, %7}, ra;\n\t}"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(a), "f"(b));
            } else if (MODE == 2) {   // FMUL then FADD, not contracted (the exact arithmetic of the intersectors / shading)
                x0 = __fadd_rn(__fmul_rn(x0, a), b); x1 = __fadd_rn(__fmul_rn(x1, a), b); x2 = __fadd_rn(__fmul_rn(x2, a), b); x3 = __fadd_rn(__fmul_rn(x3, a), b);
                x4 = __fadd_rn(__fmul_rn(x4, a), b); x5 = __fadd_rn(__fmul_rn(x5, a), b); x6 = __fadd_rn(__fmul_rn(x6, a), b); x7 = __fadd_rn(__fmul_rn(x7, a), b);
            } else {                  // integer ALU issue: LOP3 / IADD3 chains
                uint32_t* p = reinterpret_cast<uint32_t*>(&x0);
                (void)p;
                x0 = __uint_as_float((__float_as_uint(x0) ^ 0x9E3779B9u) + 0x7F4A7C15u); x1 = __uint_as_float((__float_as_uint(x1) ^ 0x9E3779B9u) + 0x7F4A7C15u);
                x2 = __uint_as_float((__float_as_uint(x2) ^ 0x9E3779B9u) + 0x7F4A7C15u); x3 = __uint_as_float((__float_as_uint(x3) ^ 0x9E3779B9u) + 0x7F4A7C15u);
                x4 = __uint_as_float((__float_as_uint(x4) ^ 0x9E3779B9u) + 0x7F4A7C15u); x5 = __uint_as_float((__float_as_uint(x5) ^ 0x9E3779B9u) + 0x7F4A7C15u);
                x6 = __uint_as_float((__float_as_uint(x6) ^ 0x9E3779B9u) + 0x7F4A7C15u); x7 = __uint_as_float((__float_as_uint(x7) ^ 0x9E3779B9u) + 0x7F4A7C15u);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// every thread streams uint4 loads over a buffer of `n16` 16-byte words, `reps` times; .cg = cache in L2 only (the L1s of 148 SMs
// together hold 30+ MB and would serve part of a 58 MB set)
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ buf, size_t n16, int reps, uint32_t* out) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)r * 977;   // shift the phase per repetition
        for (; i + 3 * stride < n16; i += 4 * stride) {
            const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
            acc += a.x ^ b.y ^ c.z ^ d.w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// the traversal's access pattern: each lane fetches a random 80-byte record (five 16-byte loads) - what a node fetch looks like to L2
__global__ void __launch_bounds__(256) k_l2_gather80(const uint4* __restrict__ buf, uint32_t nRec, int fetches, uint32_t* out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int f = 0; f < fetches; f++) {
        s ^= s << 13; s ^= s >> 17; s ^= s << 5;
        const uint4* p = buf + (size_t)(s % nRec) * 5;
        const uint4 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2), d = __ldcg(p + 3), e = __ldcg(p + 4);
        acc += a.x ^ b.y ^ c.z ^ d.w ^ e.x;
        s += acc & 1u;   // the next address depends on the data, like a child pointer
    }
    if (acc == 0x12345678u) out[0] = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

// out[0..3] = T op/s of FFMA, FFMA2 (counted as 2 FMA per instruction), FMUL+FADD pairs, ALU (LOP3+IADD3 pairs)  [1e12 lane-operations / s]
// out[4] = L2 streaming read GB/s at wsBytes, out[5] = L2 random 80-byte-record gather GB/s at wsBytes, out[6] = DRAM-sized (4 GB) streaming read GB/s
PK_API int pk_measure(int device, double wsBytes, double* out, char* err, int errLen) {
#define TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(err, errLen, "%s: %s", #x, cudaGetErrorString(e_)); return -1; } } while (0)
    TRY(cudaSetDevice(device));
    cudaDeviceProp prop; TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8;
    float* dout; TRY(cudaMalloc(&dout, (size_t)blocks * 256 * sizeof(float)));
    cudaEvent_t e0, e1; TRY(cudaEventCreate(&e0)); TRY(cudaEventCreate(&e1));
    const int iters = 4096;
    for (int mode = 0; mode < 4; mode++) {
        double best = 0;
        for (int rep = 0; rep < 5; rep++) {
            TRY(cudaEventRecord(e0));
            if (mode == 0) k_fp32<0><<<blocks, 256>>>(dout, iters, 1.0000001f, 1e-9f);
            else if (mode == 1) k_fp32<1><<<blocks, 256>>>(dout, iters, 1.0000001f, 1e-9f);
            else if (mode == 2) k_fp32<2><<<blocks, 256>>>(dout, iters, 1.0000001f, 1e-9f);
            else k_fp32<3><<<blocks, 256>>>(dout, iters, 1.0000001f, 1e-9f);
            TRY(cudaEventRecord(e1)); TRY(cudaEventSynchronize(e1)); TRY(cudaGetLastError());
            const double ops = (double)blocks * 256 * iters * 16 * 8;   // FMAs / mul+add pairs / xor+add pairs
            const double r = ops / (time_ms(e0, e1) * 1e-3) / 1e12;
            if (rep > 0 && r > best) best = r;
        }
        out[mode] = best;
    }
    // L2: working set resident (read it once to warm), then timed repetitions
    const size_t n16 = (size_t)(wsBytes / 16);
    uint4* buf; TRY(cudaMalloc(&buf, n16 * 16)); TRY(cudaMemset(buf, 1, n16 * 16));
    uint32_t* flag; TRY(cudaMalloc(&flag, 4));
    const int lblocks = prop.multiProcessorCount * 8;
    k_l2_read<<<lblocks, 256>>>(buf, n16, 2, flag);
    double best = 0;
    const int reps = 40;
    for (int rep = 0; rep < 4; rep++) {
        TRY(cudaEventRecord(e0));
        k_l2_read<<<lblocks, 256>>>(buf, n16, reps, flag);
        TRY(cudaEventRecord(e1)); TRY(cudaEventSynchronize(e1)); TRY(cudaGetLastError());
        const size_t stride = (size_t)lblocks * 256; const size_t per = (n16 / (4 * stride)) * 4 * stride;
        const double r = (double)per * 16 * reps / (time_ms(e0, e1) * 1e-3) / 1e9;
        if (r > best) best = r;
    }
    out[4] = best;
    best = 0;
    const uint32_t nRec = (uint32_t)(n16 / 5);
    const int fetches = 2048;
    for (int rep = 0; rep < 4; rep++) {
        TRY(cudaEventRecord(e0));
        k_l2_gather80<<<lblocks, 256>>>(buf, nRec, fetches, flag);
        TRY(cudaEventRecord(e1)); TRY(cudaEventSynchronize(e1)); TRY(cudaGetLastError());
        const double r = (double)lblocks * 256 * fetches * 80 / (time_ms(e0, e1) * 1e-3) / 1e9;
        if (r > best) best = r;
    }
    out[5] = best;
    TRY(cudaFree(buf));
    // DRAM-sized streaming read for reference (4 GB >> L2)
    const size_t big16 = ((size_t)4 << 30) / 16;
    TRY(cudaMalloc(&buf, big16 * 16)); TRY(cudaMemset(buf, 1, big16 * 16));
    best = 0;
    for (int rep = 0; rep < 4; rep++) {
        TRY(cudaEventRecord(e0));
        k_l2_read<<<lblocks, 256>>>(buf, big16, 1, flag);
        TRY(cudaEventRecord(e1)); TRY(cudaEventSynchronize(e1)); TRY(cudaGetLastError());
        const size_t stride = (size_t)lblocks * 256; const size_t per = (big16 / (4 * stride)) * 4 * stride;
        const double r = (double)per * 16 / (time_ms(e0, e1) * 1e-3) / 1e9;
        if (rep > 0 && r > best) best = r;
    }
    out[6] = best;
    out[7] = prop.multiProcessorCount; out[8] = prop.clockRate * 1e-3; out[9] = (double)prop.l2CacheSize;
    cudaFree(buf); cudaFree(flag); cudaFree(dout); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
#undef TRY
}
