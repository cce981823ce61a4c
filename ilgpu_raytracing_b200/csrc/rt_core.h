// rt_core.h — device data layouts and the "exact" arithmetic layer of the renderer core.
//
// Everything in namespace rtx that decides control flow or produces radiance (intersectors,
// RNG, sampling, Fresnel, texture filtering, pack) is written in the reference's C# evaluation
// order with IEEE-754 binary32 round-to-nearest + - * / sqrt and NO fused multiply-add, so the
// results are bit-identical to the CPU semantics of the reference kernels.  The .cu files are
// compiled with --fmad=false --prec-div=true --prec-sqrt=true --ftz=false; fused multiply-adds
// appear only where they are spelled rt_fma() (the wide-BVH box test, which only has to be
// conservative, never exact).
//
// The functions are __host__ __device__ so that tests/hostsim can drive the very same code on the
// CPU (a test harness, never a product path: the library refuses to run without a CUDA device).
//
// Citations "File.cs:line" are relative to /root/reference/ILGPU_Raytracing/Engine/.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rtcore_b200.h"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__
#else
// host-only translation units (BVH builder, tests/hostsim): minimal stand-ins for the CUDA vector types
#include <algorithm>
#define RT_HD inline
struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
using std::max;
using std::min;
#endif

namespace rtx {

// ----------------------------------------------------------------------------- small vector type
struct f3 {
    float x, y, z;
};
RT_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD f3 mk3(const RtFloat3& v) { return mk3(v.X, v.Y, v.Z); }
// Float3.cs:17-64 operators (component-wise, this exact order)
RT_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD f3 neg(f3 v) { return mk3(-v.x, -v.y, -v.z); }
RT_HD f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }   // Float3.cs:79-82
RT_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                                        // Float3.cs:85-88
RT_HD f3 normalize(f3 v) {                                                                                       // Float3.cs:91-95
    float inv = 1.0f / sqrtf(fmaxf(1e-20f, v.x * v.x + v.y * v.y + v.z * v.z));   // XMath.Rsqrt = 1/sqrt
    return mk3(v.x * inv, v.y * inv, v.z * inv);
}
RT_HD f3 inv_dir(f3 d) {   // RTRay.cs:548-549
    return mk3(1.0f / (d.x != 0.0f ? d.x : 1e-8f), 1.0f / (d.y != 0.0f ? d.y : 1e-8f), 1.0f / (d.z != 0.0f ? d.z : 1e-8f));
}
RT_HD float rt_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
RT_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RT_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
RT_HD int rt_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
RT_HD int rt_bfind(uint32_t x) {   // index of highest set bit, x != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}

// ----------------------------------------------------------------------------- transcendentals
// Stand-ins for XMath.Sin/Cos/Atan2/Acos/Tan (RTRay.cs:592-593,349; SceneDeviceViews.cs:152-153):
// Cephes-style binary32 kernels in + - * / sqrt only.  The oracle carries an independent copy of the
// same operation sequence (oracle/rt_oracle.cpp orc_sincos/orc_atan2/orc_acos); DESIGN.md documents it.
RT_HD void sincos_pi2(float x, float* s, float* c) {
    float ax = fabsf(x);
    int k = (int)(ax * 0.63661977236758134308f + 0.5f);
    float fk = (float)k;
    float r = ((ax - fk * 1.5703125f) - fk * 4.837512969970703125e-4f) - fk * 7.54978995489188216e-8f;
    float z = r * r;
    float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    float ss, cc;
    switch (k & 3) {
        case 0: ss = sp; cc = cp; break;
        case 1: ss = cp; cc = -sp; break;
        case 2: ss = -sp; cc = -cp; break;
        default: ss = -cp; cc = sp; break;
    }
    if (x < 0.0f) ss = -ss;
    *s = ss; *c = cc;
}
RT_HD float tan_p(float x) { float s, c; sincos_pi2(x, &s, &c); return s / c; }   // XMath.Tan stand-in (RTRay.cs:349)
RT_HD float atan_pos(float a) {
    float y0;
    if (a > 2.414213562373095f) { y0 = 1.5707963267948966f; a = -(1.0f / a); }
    else if (a > 0.4142135623730950f) { y0 = 0.7853981633974483f; a = (a - 1.0f) / (a + 1.0f); }
    else y0 = 0.0f;
    float z = a * a;
    float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * a + a;
    return y0 + p;
}
RT_HD float atan2_p(float y, float x) {
    const float PI_F = 3.14159265358979323846f;
    if (x == 0.0f) {
        if (y > 0.0f) return 1.5707963267948966f;
        if (y < 0.0f) return -1.5707963267948966f;
        return 0.0f;
    }
    float q = y / x;
    float a = atan_pos(fabsf(q));
    if (q < 0.0f) a = -a;
    if (x < 0.0f) a = (y >= 0.0f) ? (a + PI_F) : (a - PI_F);
    return a;
}
RT_HD float acos_p(float x) {
    const float PI_F = 3.14159265358979323846f;
    float a = fabsf(x);
    float r;
    if (a > 0.5f) {
        float z = 0.5f * (1.0f - a);
        float s = sqrtf(z);
        float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * s + s;
        r = 2.0f * p;
        if (x < 0.0f) r = PI_F - r;
    } else {
        float z = x * x;
        float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * x + x;
        r = 1.5707963267948966f - p;
    }
    return r;
}

// Stand-in for XMath.Pow (RTTaa.cs:238-253, the sRGB transfer curves of the TAAU resolve), x > 0: exp2(y * log2(x)) with
//   log2: x = 2^e * m, m in [sqrt(1/2), sqrt(2)], ln m = 2 s (1 + z/3 + z^2/5 + z^3/7 + z^4/9), s = (m-1)/(m+1), z = s^2
//   exp2: t = k + r, |r| <= 1/2, e^(r ln 2) by its degree-7 Taylor polynomial in Horner form, scaled by 2^k through the exponent bits
// in + - * / only (relative error ~1e-6; the consumers round to 8 bits).  The oracle carries an independent copy (orc_pow).
RT_HD float log2_p(float x) {
    uint32_t b = f2u(x);
    int e = (int)((b >> 23) & 0xFFu) - 127;
    float m = u2f((b & 0x007FFFFFu) | 0x3F800000u);
    if (m > 1.41421356f) { m = m * 0.5f; e = e + 1; }
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float p = (((0.1111111111f * z + 0.1428571429f) * z + 0.2f) * z + 0.3333333333f) * z + 1.0f;
    float ln = 2.0f * s * p;
    return (float)e + ln * 1.4426950408889634f;
}
RT_HD float exp2_p(float t) {
    if (t < -126.0f) return 0.0f;
    if (t > 127.0f) t = 127.0f;
    float kf = floorf(t + 0.5f);
    float r = t - kf;
    float u = r * 0.6931471805599453f;
    float p = ((((((u * (1.0f / 7.0f) + 1.0f) * u * (1.0f / 6.0f) + 1.0f) * u * 0.2f + 1.0f) * u * 0.25f + 1.0f) * u * (1.0f / 3.0f) + 1.0f) * u * 0.5f + 1.0f) * u + 1.0f;
    return p * u2f((uint32_t)((int)kf + 127) << 23);
}
RT_HD float pow_p(float x, float y) { return x > 0.0f ? exp2_p(y * log2_p(x)) : 0.0f; }

// ----------------------------------------------------------------------------- RNG (RTUtils.cs:20-138)
RT_HD uint32_t rotl32(uint32_t v, int r) { return (v << (r & 31)) | (v >> ((32 - r) & 31)); }   // :100-103
RT_HD uint32_t splitmix32(uint64_t x) {                                                         // :54-62
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    x ^= (x >> 31);
    return (uint32_t)(x ^ (x >> 32));
}
RT_HD uint32_t pcg_permute(uint32_t x) { x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x; }   // :65-74
RT_HD uint32_t hash32(uint32_t x) {   // :77-84 (same mixer as RTRay.cs:637-641)
    x ^= x >> 17; x *= 0xED5AD4BBu; x ^= x >> 11; x *= 0xAC4C1B51u; x ^= x >> 15; x *= 0x31848BABu; x ^= x >> 14; return x;
}
RT_HD uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) { return hash32(a ^ hash32(b ^ hash32(c))); }   // RTRay.cs:643
RT_HD uint32_t rng_seed_pixel(uint32_t px, uint32_t py, int frame, uint32_t sample, uint32_t salt, int lockNoise) {   // :116-137 + :87-97
    uint32_t f = (lockNoise != 0) ? 0u : (uint32_t)frame;
    uint32_t ln = (uint32_t)lockNoise;
    uint32_t lnMix0 = (lockNoise != 0) ? (hash32(ln) ^ (ln * 0x1B873593u)) : 0u;
    uint32_t lnMix1 = (lockNoise != 0) ? (rotl32(ln, 7) * 0x85EBCA6Bu) : 0u;
    uint32_t a = px ^ 0xB5297A4Du;
    uint32_t b = (py * 0x68E31DA4u) ^ (f * 0x9E3779B1u + 0x85EBCA6Bu) ^ lnMix0;
    uint32_t c = (sample ^ 0xC2B2AE35u) + rotl32(px, 16);
    uint32_t d = ((salt ^ 0x27D4EB2Fu) + rotl32(py, 8)) ^ lnMix1;
    uint64_t lane0 = ((uint64_t)a << 32) | b;
    uint64_t lane1 = ((uint64_t)c << 32) | d;
    uint32_t s0 = splitmix32(lane0 ^ 0xD1B54A32D192ED03ULL);
    uint32_t s1 = splitmix32(lane1 ^ 0x94D049BB133111EBULL);
    uint32_t s = pcg_permute(s0 ^ (rotl32(s1, 13) + 0x9E3779B1u));
    s |= 1u;
    return (s == 0u) ? 1u : s;   // RNG.Create :25-30
}
RT_HD uint32_t rng_next_u(uint32_t& state) {   // :33-42
    uint32_t x = state; x ^= x << 13; x ^= x >> 17; x ^= x << 5; state = (x != 0u) ? x : 1u; return state;
}
RT_HD float rng_next_f(uint32_t& state) { return (float)(rng_next_u(state) & 0x00FFFFFFu) * (1.0f / 16777216.0f); }   // :45-49

// ----------------------------------------------------------------------------- device scene
// Primitive record, 48 B = 3 x float4, stored in wide-BVH leaf order (one record per (instance, prim)).
//   triangle: q0 = v0.xyz | primId      q1 = v1.xyz | rank       q2 = v2.xyz | meta
//   sphere  : q0 = c.xyz  | primId      q1 = r,0,0  | rank       q2 = 0,0,0  | meta
// primId = global triangle index / sphere index (what the reference reports), rank = position in the
// reference's own traversal order (tie-break for equal t, SceneDeviceViews.cs:142,199,68), meta bits below.
struct PrimRec { float4 q0, q1, q2; };
enum : uint32_t {
    PRIM_INST_MASK   = 0x00FFFFFFu,   // instance index
    PRIM_SPHERE      = 1u << 31,
    PRIM_XFORM       = 1u << 30,      // instance transform is not the identity: transform the ray per test
    PRIM_ALPHA       = 1u << 29,      // triangle material has a valid alpha map (SceneDeviceViews.cs:215,297)
    PRIM_NO_CLOSEST  = 1u << 28       // no alpha map and 1 < AlphaCutoff: invisible to closest-hit (:209,218), still occludes
};

// Compressed 8-wide node, 80 B = 5 x uint4 (after Ylitie/Karras/Laine, HPG 2017; plane bytes re-ordered for one-PRMT decode).
//   n0 = px, py, pz (quantisation origin, float bits), ex | ey<<8 | ez<<16 | imask<<24
//   n1 = childBase, primBase, valid24, imr<<24
//   n2 = x planes, n3 = y planes, n4 = z planes; word k of an axis = { qlo[2k], qhi[2k], qlo[2k+1], qhi[2k+1] } (children 2k, 2k+1)
// imask   : bit s set = slot s holds an internal child (children are stored compacted in slot order from childBase).
// valid24 : 3-bit field per slot (bits 3s..3s+2) = unary primitive count (1/3/7) of a leaf child, 0 otherwise; the node's
//           primitives are stored compacted in (slot, k) order from primBase, so bit b is record primBase + popc(valid24 & below(b)).
// imr     : the hit-table index (rt_traverse.h: hit_table_entry) whose hit set is exactly imask, i.e. sum over slots of (!imask_s) << (7-s).
// Empty slots carry inverted planes (qlo = 255, qhi = 0) and can never be hit.
struct WideNode { uint4 n0, n1, n2, n3, n4; };

struct DeviceScene {
    const WideNode* nodes;        int nNodes;
    const PrimRec* prims;         int nPrims;
    const RtInstanceRecord* instances; int nInstances;
    const RtSphere* spheres;      int nSpheres;
    const RtFloat2* texcoords;
    const RtMeshTriUV* triUVs;
    const int32_t* triMatIndex;
    const RtMaterialRecord* materials; int nMaterials;
    const RtRGBA32* texels;
    const RtTexInfo* texInfos;    int nTexInfos;   // "Length" after AllocateOrEmpty (>= 1)
    int triMaterials;             // RT_FLAG_TRI_MATERIALS
    // max(1, largest instance uniformScale): the reference reports tWorld = tObj / scale (SceneDeviceViews.cs:67) although the
    // un-normalised object ray already runs in world parameter, so a hit of an instance scaled by s carries t = distance / s;
    // boxes (true distances) are therefore culled against best.t * tFarScale.  Exactly 1.0 (a no-op) for rigid instances.
    float tFarScale;
};

// ----------------------------------------------------------------------------- instance transforms (SceneDeviceViews.cs:475-493)
RT_HD f3 xf_point(const RtAffine3x4& m, f3 p) {
    return mk3(m.m00 * p.x + m.m01 * p.y + m.m02 * p.z + m.m03, m.m10 * p.x + m.m11 * p.y + m.m12 * p.z + m.m13, m.m20 * p.x + m.m21 * p.y + m.m22 * p.z + m.m23);
}
RT_HD f3 xf_vector(const RtAffine3x4& m, f3 v) {
    return mk3(m.m00 * v.x + m.m01 * v.y + m.m02 * v.z, m.m10 * v.x + m.m11 * v.y + m.m12 * v.z, m.m20 * v.x + m.m21 * v.y + m.m22 * v.z);
}

// ----------------------------------------------------------------------------- intersectors
// IntersectTriangleMT_Bary, SceneDeviceViews.cs:540-558 (the normal is evaluated later, at the final hit only)
RT_HD bool intersect_tri(f3 o, f3 d, f3 v0, f3 v1, f3 v2, float* t, float* bu, float* bv) {
    f3 e1 = v1 - v0, e2 = v2 - v0;
    f3 p = cross(d, e2);
    float det = dot(e1, p);
    if (fabsf(det) < 1e-8f) return false;
    float invDet = 1.0f / det;
    f3 tv = o - v0;
    float u = dot(tv, p) * invDet;
    if (u < 0.0f || u > 1.0f) return false;
    f3 q = cross(tv, e1);
    float v = dot(d, q) * invDet;
    if (v < 0.0f || u + v > 1.0f) return false;
    float tt = dot(e2, q) * invDet;
    if (tt <= 0.0f) return false;
    *t = tt; *bu = u; *bv = v;
    return true;
}
// IntersectSphere, SceneDeviceViews.cs:517-537 (normal evaluated later)
RT_HD bool intersect_sphere(f3 o, f3 d, f3 c, float radius, float* t) {
    f3 oc = o - c;
    float a = dot(d, d);
    float b = 2.0f * dot(oc, d);
    float cc = dot(oc, oc) - radius * radius;
    float disc = b * b - 4.0f * a * cc;
    if (disc < 0.0f) return false;
    float sqrtD = sqrtf(disc);
    float t0 = (-b - sqrtD) / (2.0f * a);
    float t1 = (-b + sqrtD) / (2.0f * a);
    float tt = t0;
    if (tt < 0.001f) { tt = t1; if (tt < 0.001f) return false; }
    *t = tt;
    return true;
}

// ----------------------------------------------------------------------------- textures (SceneDeviceViews.cs:330-472)
RT_HD RtRGBA32 texel_raw(const DeviceScene& sc, RtTexInfo info, int x, int y) {   // :330-339
    int w = info.Width, h = info.Height;
    RtRGBA32 z; z.R = z.G = z.B = z.A = 0;
    if (w <= 0 || h <= 0) return z;
    int sx = max(0, min(w - 1, x)), sy = max(0, min(h - 1, y));
    return sc.texels[info.Offset + sy * w + sx];
}
RT_HD float luma01(RtRGBA32 p) {   // :342-348
    float r = p.R * (1.0f / 255.0f), g = p.G * (1.0f / 255.0f), b = p.B * (1.0f / 255.0f);
    return 0.2126f * r + 0.7152f * g + 0.0722f * b;
}
RT_HD f3 texel_rgb(const DeviceScene& sc, RtTexInfo info, int x, int y) {   // :351-355
    RtRGBA32 p = texel_raw(sc, info, x, y);
    return mk3(p.R * (1.0f / 255.0f), p.G * (1.0f / 255.0f), p.B * (1.0f / 255.0f));
}
struct BilinearTap { int x0, y0, x1, y1; float tx, ty; };
RT_HD BilinearTap bilinear_tap(RtTexInfo info, float u, float v) {   // shared prologue of :358-375, :388-405, :431-448
    int w = info.Width, h = info.Height;
    float fu = u - floorf(u);
    float fv = 1.0f - (v - floorf(v));
    float x = fu * (float)(w - 1), y = fv * (float)(h - 1);
    BilinearTap b;
    b.x0 = (int)floorf(x); b.y0 = (int)floorf(y);
    b.x1 = min(w - 1, b.x0 + 1); b.y1 = min(h - 1, b.y0 + 1);
    b.tx = x - (float)b.x0; b.ty = y - (float)b.y0;
    return b;
}
RT_HD f3 sample_texture_linear(const DeviceScene& sc, RtTexInfo info, float u, float v) {   // :358-385 and the RGB part of :431-472
    if (info.Width <= 0 || info.Height <= 0) return mk3(1.0f, 1.0f, 1.0f);
    BilinearTap b = bilinear_tap(info, u, v);
    f3 c00 = texel_rgb(sc, info, b.x0, b.y0), c10 = texel_rgb(sc, info, b.x1, b.y0), c01 = texel_rgb(sc, info, b.x0, b.y1), c11 = texel_rgb(sc, info, b.x1, b.y1);
    f3 cx0 = c00 * (1.0f - b.tx) + c10 * b.tx;
    f3 cx1 = c01 * (1.0f - b.tx) + c11 * b.tx;
    return cx0 * (1.0f - b.ty) + cx1 * b.ty;
}
RT_HD float sample_mask_linear(const DeviceScene& sc, RtTexInfo info, float u, float v) {   // :388-415
    if (info.Width <= 0 || info.Height <= 0) return 1.0f;
    BilinearTap b = bilinear_tap(info, u, v);
    float a00 = luma01(texel_raw(sc, info, b.x0, b.y0)), a10 = luma01(texel_raw(sc, info, b.x1, b.y0));
    float a01 = luma01(texel_raw(sc, info, b.x0, b.y1)), a11 = luma01(texel_raw(sc, info, b.x1, b.y1));
    float ax0 = a00 * (1.0f - b.tx) + a10 * b.tx;
    float ax1 = a01 * (1.0f - b.tx) + a11 * b.tx;
    return ax0 * (1.0f - b.ty) + ax1 * b.ty;
}
RT_HD float sample_mask_point(const DeviceScene& sc, RtTexInfo info, float u, float v) {   // :418-428 (XMath.Round = half-to-even)
    int w = info.Width, h = info.Height;
    if (w <= 0 || h <= 0) return 1.0f;
    float fu = u - floorf(u);
    float fv = 1.0f - (v - floorf(v));
    int x = (int)rintf(fu * (float)(w - 1));
    int y = (int)rintf(fv * (float)(h - 1));
    return luma01(texel_raw(sc, info, x, y));
}
RT_HD bool tex_valid(const DeviceScene& sc, int has, int idx) { return has != 0 && idx >= 0 && idx < sc.nTexInfos; }

// interpolated UV of a triangle hit, SceneDeviceViews.cs:201-207 / :299-305
RT_HD void tri_uv(const DeviceScene& sc, int tri, float bu, float bv, float* uu, float* vv) {
    RtMeshTriUV tuv = sc.triUVs[tri];
    RtFloat2 t0 = sc.texcoords[tuv.t0], t1 = sc.texcoords[tuv.t1], t2 = sc.texcoords[tuv.t2];
    float w = 1.0f - bu - bv;
    *uu = t0.X * w + t1.X * bu + t2.X * bv;
    *vv = t0.Y * w + t1.Y * bu + t2.Y * bv;
}
// closest-hit alpha rule: "if (alpha < mat.AlphaCutoff) continue" (SceneDeviceViews.cs:209-218); true = hit survives
RT_HD bool tri_alpha_pass_closest(const DeviceScene& sc, int tri, float bu, float bv) {
    const RtMaterialRecord& mat = sc.materials[sc.triMatIndex[tri]];
    float alpha = 1.0f;
    if (tex_valid(sc, mat.HasAlphaMap, mat.AlphaTexIndex)) {
        float uu, vv; tri_uv(sc, tri, bu, bv, &uu, &vv);
        alpha = sample_mask_linear(sc, sc.texInfos[mat.AlphaTexIndex], uu, vv);
    }
    return !(alpha < mat.AlphaCutoff);
}
// any-hit alpha rule (SceneDeviceViews.cs:297-315); true = occludes
RT_HD bool tri_alpha_pass_anyhit(const DeviceScene& sc, int tri, float bu, float bv) {
    const RtMaterialRecord& mat = sc.materials[sc.triMatIndex[tri]];
    if (tex_valid(sc, mat.HasAlphaMap, mat.AlphaTexIndex)) {
        float uu, vv; tri_uv(sc, tri, bu, bv, &uu, &vv);
        RtTexInfo ti = sc.texInfos[mat.AlphaTexIndex];
        float aPoint = sample_mask_point(sc, ti, uu, vv);
        float cutoff = mat.AlphaCutoff;
        const float Band = 0.10f;
        if (aPoint < cutoff - Band) return false;
        if (aPoint >= cutoff + Band) return true;
        float aLin = sample_mask_linear(sc, ti, uu, vv);
        if (aLin < cutoff) return false;
    }
    return true;
}

// ----------------------------------------------------------------------------- hit record + surface evaluation
// What extend writes per ray (16 B): world t (1e30 = miss), index into DeviceScene::prims, barycentrics.
struct alignas(16) HitRec { float t; int prim; float bu, bv; };

// Everything TraceClosest returns for the winning hit (SceneDeviceViews.cs:30-86), evaluated once, after traversal.
struct Surface { f3 normal; f3 albedo; int objId, shade; float ior; int primId, instId; };

RT_HD Surface eval_surface(const DeviceScene& sc, f3 wo, f3 wd, const HitRec& h) {
    Surface s;
    const PrimRec& pr = sc.prims[h.prim];
    uint32_t meta = f2u(pr.q2.w);
    int inst = (int)(meta & PRIM_INST_MASK);
    const RtInstanceRecord& ir = sc.instances[inst];
    s.instId = inst;
    s.primId = (int)f2u(pr.q0.w);
    // object-space ray (TransformRay :475-481); for the identity the products with 1/0 are exact
    f3 oo = wo, od = wd;
    if (meta & PRIM_XFORM) { oo = xf_point(ir.worldToObject, wo); od = xf_vector(ir.worldToObject, wd); }
    f3 nObj;
    if (meta & PRIM_SPHERE) {
        const RtSphere& sp = sc.spheres[s.primId];
        float tObj = h.t;
        if (meta & PRIM_XFORM) intersect_sphere(oo, od, mk3(sp.center), sp.radius, &tObj);   // same arithmetic as in traversal
        f3 p = oo + od * tObj;
        nObj = normalize(p - mk3(sp.center));                                                 // :534-535
        f3 kd = mk3(sp.material.Kd);                                                         // :146-157
        f3 col = (kd.x == 0.0f && kd.y == 0.0f && kd.z == 0.0f) ? mk3(sp.albedo) : kd;
        if (tex_valid(sc, sp.material.HasDiffuseMap, sp.material.DiffuseTexIndex)) {
            const float PI = 3.14159265358979323846f;
            float u = 0.5f + atan2_p(nObj.z, nObj.x) / (2.0f * PI);
            float v = acos_p(fminf(1.0f, fmaxf(-1.0f, nObj.y))) / PI;
            col = sample_texture_linear(sc, sc.texInfos[sp.material.DiffuseTexIndex], u, v);
        }
        s.albedo = col;
        s.shade = sp.shading;                   // :158
        s.ior = sp.ior > 0.0f ? sp.ior : 1.0f;  // :159
        s.objId = -1;                           // triLocal = -1 for spheres (:56,73)
    } else {
        f3 v0 = mk3(pr.q0.x, pr.q0.y, pr.q0.z), v1 = mk3(pr.q1.x, pr.q1.y, pr.q1.z), v2 = mk3(pr.q2.x, pr.q2.y, pr.q2.z);
        nObj = normalize(cross(v1 - v0, v2 - v0));   // :556
        const RtMaterialRecord& mat = sc.materials[sc.triMatIndex[s.primId]];
        f3 kdCol = mk3(mat.Kd);                      // :210-213
        if (tex_valid(sc, mat.HasDiffuseMap, mat.DiffuseTexIndex)) {
            float uu, vv; tri_uv(sc, s.primId, h.bu, h.bv, &uu, &vv);
            kdCol = sample_texture_linear(sc, sc.texInfos[mat.DiffuseTexIndex], uu, vv);
        }
        if (mat.TwoSided != 0 && dot(nObj, od) > 0.0f) nObj = nObj * -1.0f;   // :222
        s.albedo = kdCol;
        s.shade = 0; s.ior = 1.0f;                   // :61 — triangles are always Lambert in the reference
        if (sc.triMaterials) { s.shade = mat.Shading; s.ior = mat.IOR > 0.0f ? mat.IOR : 1.0f; }   // RT_FLAG_TRI_MATERIALS extension
        s.objId = s.primId;
    }
    s.normal = normalize(xf_vector(ir.objectToWorld, nObj));   // :71
    return s;
}

// ----------------------------------------------------------------------------- shading math (RTRay.cs)
#define RTX_PI 3.14159265358979323846f      // RTRay.cs:183
#define RTX_INV_PI 0.31830988618379067154f  // :184
#define RTX_EPS_N 0.0025f                   // :185
#define RTX_EPS_MIN 1e-6f                   // :186

struct RayOD { f3 o, d; };
RT_HD RayOD make_ray_normal_offset(f3 origin, f3 n, f3 dir) {   // :552-558 (invDir is derived by extend)
    f3 d = normalize(dir);
    float s = dot(n, d) >= 0.0f ? 1.0f : -1.0f;
    RayOD r; r.o = origin + n * (RTX_EPS_N * s); r.d = d; return r;
}
RT_HD f3 reflect3(f3 I, f3 N) { return I - N * (2.0f * dot(I, N)); }   // :561
RT_HD bool refract3(f3 I, f3 N, float etaI, float etaT, f3* T) {       // :564-572
    float eta = etaI / etaT;
    float cosI = -dot(I, N);
    float k = 1.0f - eta * eta * (1.0f - cosI * cosI);
    if (k < 0.0f) { *T = mk3(0, 0, 0); return false; }
    *T = normalize(I * eta + N * (eta * cosI - sqrtf(k)));
    return true;
}
RT_HD float schlick_fresnel(float c, float etaI, float etaT) {   // :575-583
    float r0 = (etaI - etaT) / (etaI + etaT);
    r0 = r0 * r0;
    float m = 1.0f - c;
    float m2 = m * m;
    float m5 = m2 * m2 * m;
    return r0 + (1.0f - r0) * m5;
}
// OrthonormalBasis (:601-606) is a pure function of the normal: a vertex evaluates it once for its nine calls of
// SampleHemisphereCosine (eight ReSTIR candidates + the bounce) instead of once per call; same operations, same bits.
struct Basis { f3 t, b; };
RT_HD Basis orthonormal_basis(f3 n) {
    f3 up = fabsf(n.y) < 0.999f ? mk3(0, 1, 0) : mk3(1, 0, 0);
    Basis B;
    B.t = normalize(cross(up, n));
    B.b = cross(n, B.t);
    return B;
}
RT_HD f3 hemisphere_cosine_dir(f3 n, const Basis& B, float r1, float r2);
RT_HD f3 sample_hemisphere_cosine(f3 n, const Basis& B, uint32_t& rng) {   // :586-598
    float r1 = rng_next_f(rng), r2 = rng_next_f(rng);
    return hemisphere_cosine_dir(n, B, r1, r2);
}
RT_HD f3 hemisphere_cosine_dir(f3 n, const Basis& B, float r1, float r2) {   // the arithmetic of :588-597 for given random numbers
    float phi = 2.0f * RTX_PI * r1;
    float cosTheta = sqrtf(1.0f - r2);
    float sinTheta = sqrtf(r2);
    float sphi, cphi; sincos_pi2(phi, &sphi, &cphi);
    float x = cphi * sinTheta, y = sphi * sinTheta, z = cosTheta;
    f3 v = B.t * x + B.b * y + n * z;
    return normalize(v);
}
RT_HD float luminance(f3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }              // :627
RT_HD float cos_hemisphere_pdf(f3 n, f3 wi) { float nl = fmaxf(0.0f, dot(n, wi)); return nl * RTX_INV_PI; }   // :630-634
RT_HD f3 safe_color(f3 c) {   // :646-655
    float x = isfinite(c.x) ? c.x : 0.0f, y = isfinite(c.y) ? c.y : 0.0f, z = isfinite(c.z) ? c.z : 0.0f;
    return mk3(fminf(1e6f, fmaxf(-1e6f, x)), fminf(1e6f, fmaxf(-1e6f, y)), fminf(1e6f, fmaxf(-1e6f, z)));
}
RT_HD int pack_rgba8(f3 c) {   // :66-76
    int R = (int)(255.99f * fminf(1.0f, fmaxf(0.0f, c.x)));
    int G = (int)(255.99f * fminf(1.0f, fmaxf(0.0f, c.y)));
    int B = (int)(255.99f * fminf(1.0f, fmaxf(0.0f, c.z)));
    return (int)((255u << 24) | ((uint32_t)R << 16) | ((uint32_t)G << 8) | (uint32_t)B);
}
RT_HD int float_to_i16(float x) { float cl = fmaxf(0.0f, fminf(65535.0f, x * 1000.0f)); return (int)cl & 0xFFFF; }   // :609-613
RT_HD float i16_to_float(int v) { return (float)v / 1000.0f; }                                                         // :615

struct LightEnv { f3 dirLightDir, dirLightRadiance, skyTop, skyBottom; };
RT_HD f3 sky_weighted(const LightEnv& e, f3 dir) {   // :164-168
    float tbg = 0.5f * (dir.y + 1.0f);
    return e.skyBottom * (1.0f - tbg) + e.skyTop * tbg;
}

// Reservoir (RTRay.cs:171-179) and its update (:394-405)
struct Reservoir { f3 L, wi; float pdf, w, wSum; int m, lightId; };
RT_HD void reservoir_update(Reservoir& r, f3 wi, float pdfSel, f3 Li, float scoreS, int multiplicity, int lightId, uint32_t& rng) {
    float add = scoreS;
    float newSum = r.wSum + add;
    float acceptP = (newSum > 0.0f) ? add / newSum : 0.0f;
    if (rng_next_f(rng) < acceptP) { r.wi = wi; r.pdf = pdfSel; r.L = Li; r.w = scoreS; r.lightId = lightId; }
    r.wSum = newSum;
    r.m = r.m + max(1, multiplicity);
}

// ReSTIR_Direct (RTRay.cs:438-543) in three pieces, so the optional prev-frame imports (:475-516, rt_wavefront.h) can sit
// between candidate generation and the final selection without touching the common path:
//   restir_new_candidates : (1) eight cosine-hemisphere sky candidates + (2) the directional delta candidate
//   restir_finalize       : (5) up to the visibility test: true when a shadow ray must be traced; *wiSel is its direction and
//                           *contrib the value "f_over_p * W" (:535-537) the caller adds to Li (times throughput) if unoccluded
#define RTX_MIX_LOCAL (8.0f / 9.0f)   // (float)LocalCandidates / (float)TotalNew, :446
#define RTX_MIX_DELTA (1.0f / 9.0f)   // :447
RT_HD void restir_new_candidates(const LightEnv& env, f3 n, const Basis& B, f3 albedo, uint32_t& rng, Reservoir& r) {
    const int LocalCandidates = 8, DeltaCandidates = 1, TotalNew = LocalCandidates + DeltaCandidates;
    float mixLocal = (float)LocalCandidates / (float)TotalNew;
    float mixDelta = (float)DeltaCandidates / (float)TotalNew;
    r.L = mk3(0, 0, 0); r.wi = mk3(0, 0, 0); r.pdf = 0; r.w = 0; r.wSum = 0; r.m = 0; r.lightId = 0;   // :330-335
    for (int i = 0; i < LocalCandidates; i++) {   // (1) :452-462
        f3 wi = sample_hemisphere_cosine(n, B, rng);
        float nl = fmaxf(0.0f, dot(n, wi));
        float pdfLocal = fmaxf(RTX_EPS_MIN, cos_hemisphere_pdf(n, wi));
        float pdfSel = fmaxf(RTX_EPS_MIN, pdfLocal * mixLocal);
        f3 LiLoc = sky_weighted(env, wi);
        f3 f_over_p = albedo * LiLoc * ((nl / pdfSel) * RTX_INV_PI);
        float s = luminance(f_over_p);
        reservoir_update(r, wi, pdfSel, LiLoc, s, 1, 1, rng);
    }
    {   // (2) :465-473
        f3 wi = normalize(env.dirLightDir);
        float nl = fmaxf(0.0f, dot(n, wi));
        float pdfSel = fmaxf(RTX_EPS_MIN, mixDelta);
        f3 LiDir = env.dirLightRadiance;
        f3 f_over_p = albedo * LiDir * ((nl / pdfSel) * RTX_INV_PI);
        float s = luminance(f_over_p);
        reservoir_update(r, wi, pdfSel, LiDir, s, 1, 2, rng);
    }
}
#if defined(__CUDACC__)
// RT_FLAG_FAST_SHADING (opt-in; north_star asks radiance only to 1e-4 relative RMS): the eight sky candidates of (1) scored with
// fused multiply-adds and the special-function unit (sin / cos, sqrt, rsqrt, reciprocal to ~1e-7 relative) instead of ~235
// instructions of exact arithmetic each.  Nothing here decides the path: every candidate still draws exactly three random
// numbers, and the bounce direction, Russian roulette and all intersections stay exact - so hit ids, bounce counts and the RNG
// streams are unchanged; only which sky direction wins a near-tie and the last bits of the direct-light term can differ.
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void restir_new_candidates_fast(const LightEnv& env, f3 n, const Basis& B, f3 albedo, uint32_t& rng, Reservoir& r) {
    const float mixLocal = 8.0f / 9.0f, mixDelta = 1.0f / 9.0f;
    r.L = mk3(0, 0, 0); r.wi = mk3(0, 0, 0); r.pdf = 0; r.w = 0; r.wSum = 0; r.m = 0; r.lightId = 0;
    // luminance(albedo * Li * k) = k * dot(albedo * lumaWeights, Li), Li = skyBottom + (skyTop - skyBottom) * t: two constants per vertex
    const f3 aw = mk3(0.2126f * albedo.x, 0.7152f * albedo.y, 0.0722f * albedo.z);
    const f3 dsky = mk3(env.skyTop.x - env.skyBottom.x, env.skyTop.y - env.skyBottom.y, env.skyTop.z - env.skyBottom.z);
    const float lumB = __fmaf_rn(aw.x, env.skyBottom.x, __fmaf_rn(aw.y, env.skyBottom.y, aw.z * env.skyBottom.z));
    const float lumD = __fmaf_rn(aw.x, dsky.x, __fmaf_rn(aw.y, dsky.y, aw.z * dsky.z));
    float selR1 = 0.0f, selR2 = 0.0f; bool selected = false;
#pragma unroll 2
    for (int i = 0; i < 8; i++) {
        const float r1 = rng_next_f(rng), r2 = rng_next_f(rng);
        float sphi, cphi;
        __sincosf(2.0f * RTX_PI * r1, &sphi, &cphi);
        const float sinT = fast_sqrt(r2), cosT = fast_sqrt(1.0f - r2);
        const float x = cphi * sinT, y = sphi * sinT;
        f3 v = mk3(__fmaf_rn(B.t.x, x, __fmaf_rn(B.b.x, y, n.x * cosT)), __fmaf_rn(B.t.y, x, __fmaf_rn(B.b.y, y, n.y * cosT)), __fmaf_rn(B.t.z, x, __fmaf_rn(B.b.z, y, n.z * cosT)));
        const float inv = fast_rsqrt(fmaxf(1e-20f, __fmaf_rn(v.x, v.x, __fmaf_rn(v.y, v.y, v.z * v.z))));
        const f3 wi = mk3(v.x * inv, v.y * inv, v.z * inv);
        const float nl = fmaxf(0.0f, __fmaf_rn(n.x, wi.x, __fmaf_rn(n.y, wi.y, n.z * wi.z)));
        const float pdfSel = fmaxf(RTX_EPS_MIN, fmaxf(RTX_EPS_MIN, nl * RTX_INV_PI) * mixLocal);
        const float tbg = __fmaf_rn(0.5f, wi.y, 0.5f);
        const float s = (nl * fast_rcp(pdfSel) * RTX_INV_PI) * __fmaf_rn(lumD, tbg, lumB);
        // ReservoirUpdate (:394-405)
        const float newSum = r.wSum + s;
        const float acceptP = (newSum > 0.0f) ? s * fast_rcp(newSum) : 0.0f;
        if (rng_next_f(rng) < acceptP) { selR1 = r1; selR2 = r2; selected = true; }
        r.wSum = newSum;
        r.m = r.m + 1;
    }
    if (selected) {
        // the ONE candidate that won is re-evaluated with the exact arithmetic of restir_new_candidates: its direction (the shadow ray),
        // pdf, radiance and score are bit-identical to the exact mode's; only wSum (and, at a ~1e-7 chance, which candidate won) is approximate
        const f3 wi = hemisphere_cosine_dir(n, B, selR1, selR2);
        const float nl = fmaxf(0.0f, dot(n, wi));
        const float pdfLocal = fmaxf(RTX_EPS_MIN, cos_hemisphere_pdf(n, wi));
        const float pdfSel = fmaxf(RTX_EPS_MIN, pdfLocal * mixLocal);
        const f3 LiLoc = sky_weighted(env, wi);
        const f3 f_over_p = albedo * LiLoc * ((nl / pdfSel) * RTX_INV_PI);
        r.wi = wi; r.pdf = pdfSel; r.L = LiLoc; r.w = luminance(f_over_p); r.lightId = 1;
    }
    {   // (2) the directional candidate, exact as in restir_new_candidates (once per vertex)
        f3 wi = normalize(env.dirLightDir);
        float nl = fmaxf(0.0f, dot(n, wi));
        float pdfSel = fmaxf(RTX_EPS_MIN, mixDelta);
        f3 LiDir = env.dirLightRadiance;
        f3 f_over_p = albedo * LiDir * ((nl / pdfSel) * RTX_INV_PI);
        float s = luminance(f_over_p);
        reservoir_update(r, wi, pdfSel, LiDir, s, 1, 2, rng);
    }
}
#endif
RT_HD bool restir_finalize(const LightEnv& env, f3 n, f3 albedo, const Reservoir& r, f3* wiSel, f3* contrib) {
    const float mixLocal = 8.0f / 9.0f, mixDelta = 1.0f / 9.0f;   // mixLocal2 / mixDelta2, :528-529
    // (5) :519-539 up to the visibility test
    if (!(r.m > 0 && r.wSum > 0.0f && r.w > 0.0f)) return false;
    f3 wsel = r.wi;
    int lidSel = r.lightId == 2 ? 2 : 1;
    float nlSel = fmaxf(0.0f, dot(n, wsel));
    if (!(nlSel > 0.0f)) return false;
    if (!(dot(n, wsel) > 0.0f)) return false;   // Visible(): "if (nl <= 0f) return false" :620-621
    float pdfSel = (lidSel == 2) ? fmaxf(RTX_EPS_MIN, mixDelta) : fmaxf(RTX_EPS_MIN, cos_hemisphere_pdf(n, wsel) * mixLocal);
    f3 LiSel = (lidSel == 2) ? env.dirLightRadiance : sky_weighted(env, wsel);
    f3 f_over_p = albedo * LiSel * ((nlSel / pdfSel) * RTX_INV_PI);
    float W = r.wSum / (float)max(1, r.m) / fmaxf(RTX_EPS_MIN, r.w);
    *contrib = f_over_p * W;
    *wiSel = wsel;
    return true;
}

// ----------------------------------------------------------------------------- present chain (RTRenderer.cs:281-346, RTTaa.cs:117-258)
RT_HD f3 unpack_rgb(int rgba8) {   // RTRenderer.cs:322-328
    float r = (float)((rgba8 >> 16) & 255) * (1.0f / 255.0f), g = (float)((rgba8 >> 8) & 255) * (1.0f / 255.0f), b = (float)(rgba8 & 255) * (1.0f / 255.0f);
    return mk3(r, g, b);
}
RT_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }   // XMath.Clamp(int)
// BilinearUpsampleKernel (RTRenderer.cs:287-320); PackRGBA8 / ToByte there are the same as RTRay.cs:66-76
RT_HD int bilinear_upsample_pixel(const int* src, int srcW, int srcH, int dstW, int dstH, int index) {
    int x = index % dstW, y = index / dstW;
    float u = (((float)x + 0.5f) * (float)srcW / (float)dstW) - 0.5f;
    float v = (((float)y + 0.5f) * (float)srcH / (float)dstH) - 0.5f;
    int x0 = clampi((int)floorf(u), 0, srcW - 1), y0 = clampi((int)floorf(v), 0, srcH - 1);
    int x1 = clampi(x0 + 1, 0, srcW - 1), y1 = clampi(y0 + 1, 0, srcH - 1);
    float tx = fminf(1.0f, fmaxf(0.0f, u - (float)x0)), ty = fminf(1.0f, fmaxf(0.0f, v - (float)y0));
    f3 c00 = unpack_rgb(src[y0 * srcW + x0]), c10 = unpack_rgb(src[y0 * srcW + x1]), c01 = unpack_rgb(src[y1 * srcW + x0]), c11 = unpack_rgb(src[y1 * srcW + x1]);
    f3 cx0 = c00 * (1.0f - tx) + c10 * tx;
    f3 cx1 = c01 * (1.0f - tx) + c11 * tx;
    return pack_rgba8(cx0 * (1.0f - ty) + cx1 * ty);
}
// sRGB decode of one 8-bit channel value (RTTaa.cs:236-246).  Only 256 inputs exist: callers may tabulate it.
RT_HD float srgb_to_linear_u8(int v) {
    float c = (float)v / 255.0f;
    return (c <= 0.04045f) ? (c / 12.92f) : pow_p((c + 0.055f) / 1.055f, 2.4f);
}
RT_HD int pack_srgb(f3 c) {   // RTTaa.cs:248-262 (XMath.Round = half-to-even)
    float rL = fmaxf(0.0f, fminf(1.0f, c.x)), gL = fmaxf(0.0f, fminf(1.0f, c.y)), bL = fmaxf(0.0f, fminf(1.0f, c.z));
    float r = (rL <= 0.0031308f) ? 12.92f * rL : 1.055f * pow_p(rL, 1.0f / 2.4f) - 0.055f;
    float g = (gL <= 0.0031308f) ? 12.92f * gL : 1.055f * pow_p(gL, 1.0f / 2.4f) - 0.055f;
    float b = (bL <= 0.0031308f) ? 12.92f * bL : 1.055f * pow_p(bL, 1.0f / 2.4f) - 0.055f;
    int R = (int)rintf(fmaxf(0.0f, fminf(1.0f, r)) * 255.0f), G = (int)rintf(fmaxf(0.0f, fminf(1.0f, g)) * 255.0f), B = (int)rintf(fmaxf(0.0f, fminf(1.0f, b)) * 255.0f);
    return (int)((255u << 24) | ((uint32_t)R << 16) | ((uint32_t)G << 8) | (uint32_t)B);
}
struct TaaConst { int outW, outH, inW, inH; float feedback, sharpness, clampK; int isFirstFrame; };
// lut[v] = srgb_to_linear_u8(v), v = 0..255
RT_HD f3 unpack_srgb_lut(const float* lut, int rgba) { return mk3(lut[(rgba >> 16) & 255], lut[(rgba >> 8) & 255], lut[rgba & 255]); }
RT_HD f3 catrom2(f3 a, f3 b, float t) { float tt = t * (2.0f - t); return a * (1.0f - tt) + b * tt; }   // RTTaa.cs:228-233
RT_HD f3 sample_catrom_srgb(const float* lut, const int* a, int w, int h, float x, float y) {   // RTTaa.cs:209-226
    int x1 = clampi((int)floorf(x), 0, w - 1), y1 = clampi((int)floorf(y), 0, h - 1);
    float fx = x - (float)x1, fy = y - (float)y1;
    int xr = min(x1 + 1, w - 1), yr = min(y1 + 1, h - 1);
    f3 c00 = unpack_srgb_lut(lut, a[y1 * w + x1]), c10 = unpack_srgb_lut(lut, a[y1 * w + xr]);
    f3 c01 = unpack_srgb_lut(lut, a[yr * w + x1]), c11 = unpack_srgb_lut(lut, a[yr * w + xr]);
    return catrom2(catrom2(c00, c10, fx), catrom2(c01, c11, fx), fy);
}
// TaaResolveKernel (RTTaa.cs:117-179) for one output pixel; returns the packed colour (also the new history) and the object id
RT_HD int taa_resolve_pixel(const TaaConst& p, const float* lut, const int* inColorLow, const int* inObjIdLow, int histColor, int histObj, int idx, int* objOut) {
    int px = idx % p.outW, py = idx / p.outW;
    float sx = ((float)px + 0.5f) * ((float)p.inW / (float)p.outW) - 0.5f;
    float sy = ((float)py + 0.5f) * ((float)p.inH / (float)p.outH) - 0.5f;
    f3 cur = sample_catrom_srgb(lut, inColorLow, p.inW, p.inH, sx, sy);
    f3 nmin = cur, nmax = cur;
    for (int oy = -1; oy <= 1; oy++)
        for (int ox = -1; ox <= 1; ox++) {
            if (ox == 0 && oy == 0) continue;
            f3 c = sample_catrom_srgb(lut, inColorLow, p.inW, p.inH, sx + (float)ox * 0.5f, sy + (float)oy * 0.5f);
            nmin = mk3(fminf(nmin.x, c.x), fminf(nmin.y, c.y), fminf(nmin.z, c.z));
            nmax = mk3(fmaxf(nmax.x, c.x), fmaxf(nmax.y, c.y), fmaxf(nmax.z, c.z));
        }
    int ix = clampi((int)rintf(sx), 0, p.inW - 1), iy = clampi((int)rintf(sy), 0, p.inH - 1);   // SampleNearestObj :200-205
    int objId = inObjIdLow[iy * p.inW + ix];
    f3 hist = unpack_srgb_lut(lut, histColor);
    bool reset = (p.isFirstFrame != 0) || (histObj != objId);
    f3 cmin = mk3(nmin.x - p.clampK * 0.0f, nmin.y - p.clampK * 0.0f, nmin.z - p.clampK * 0.0f);   // Clamp :191-198
    f3 cmax = mk3(nmax.x + p.clampK * 0.0f, nmax.y + p.clampK * 0.0f, nmax.z + p.clampK * 0.0f);
    f3 hc = mk3(fminf(cmax.x, fmaxf(cmin.x, hist.x)), fminf(cmax.y, fmaxf(cmin.y, hist.y)), fminf(cmax.z, fmaxf(cmin.z, hist.z)));
    float a = reset ? 1.0f : p.feedback;
    f3 accum = hc * (1.0f - a) + cur * a;
    f3 sharpen = accum * (1.0f + 2.0f * p.sharpness) - (nmin + nmax) * (0.5f * p.sharpness);
    accum = accum * (1.0f - p.sharpness) + sharpen * p.sharpness;
    *objOut = objId;
    return pack_srgb(accum);
}

}   // namespace rtx
