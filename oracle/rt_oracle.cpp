// =====================================================================================
// rt_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A line-by-line C++ restatement of the hot path of NullandKale/ILGPU_Raytracing
// (the C# kernels ILGPU JIT-compiles to PTX) plus the host scene/BVH builders that feed
// it.  It exists to CHECK the CUDA product path and to time a CPU baseline.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
// nothing under ilgpu_raytracing_b200/ links, imports or executes it.
//
// PARITY UNPINNED: the reference has no tests, golden vectors or fixtures, cannot be built
// or run here (no .NET, needs a CUDA GPU + an OpenGL window), and its transcendental /
// rounding semantics live in ILGPU 1.5.3 + ILGPU.Algorithms 1.5.3 (NuGet, not vendored:
// ILGPU_Raytracing.csproj:12-13).  What is pinned here, as OUR reading of the CPU
// semantics of that code: IEEE-754 binary32 for + - * / sqrt in C# evaluation order, no
// FMA contraction (build with -ffp-contract=off -fno-fast-math), XMath.Rsqrt(x)=1/sqrt(x),
// XMath.Min/Max = fminf/fmaxf, and documented polynomial sin/cos/atan2/acos (orc_sincos,
// orc_atan2, orc_acos below; <= 2 ulp from libm; build with -DORC_LIBM to use libm
// instead — tests/test_oracle.py bounds the difference).  The RNG known-answer vectors of
// SURVEY.md §8c are checked in tests/test_oracle.py.
//
// All "file:line" citations are relative to /root/reference/ILGPU_Raytracing/.
// Every function names the lines it follows.  Names follow the reference.
// =====================================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace orc {

// ------------------------------------------------------------------ Engine/Float3.cs:6-114
struct Float3 {
    float X, Y, Z;
    Float3() : X(0), Y(0), Z(0) {}
    Float3(float x, float y, float z) : X(x), Y(y), Z(z) {}
};
static inline Float3 operator+(Float3 a, Float3 b) { return Float3(a.X + b.X, a.Y + b.Y, a.Z + b.Z); }   // :18-21
static inline Float3 operator-(Float3 a, Float3 b) { return Float3(a.X - b.X, a.Y - b.Y, a.Z - b.Z); }   // :24-27
static inline Float3 operator*(Float3 a, float s) { return Float3(a.X * s, a.Y * s, a.Z * s); }          // :30-33
static inline Float3 operator*(float s, Float3 a) { return Float3(a.X * s, a.Y * s, a.Z * s); }          // :36-39
static inline Float3 operator*(Float3 a, Float3 b) { return Float3(a.X * b.X, a.Y * b.Y, a.Z * b.Z); }   // :42-45
static inline Float3 operator/(Float3 a, float s) { float inv = 1.0f / s; return Float3(a.X * inv, a.Y * inv, a.Z * inv); } // :48-52
static inline Float3 operator-(Float3 v) { return Float3(-v.X, -v.Y, -v.Z); }                            // :61-64
static inline float XMin(float a, float b) { return fminf(a, b); }
static inline float XMax(float a, float b) { return fmaxf(a, b); }
static inline int XMin(int a, int b) { return a < b ? a : b; }
static inline int XMax(int a, int b) { return a > b ? a : b; }
static inline float XClamp(float v, float lo, float hi) { return XMax(XMin(v, hi), lo); }
static inline float XRsqrt(float x) { return 1.0f / sqrtf(x); }   // XMath.Rsqrt = Rcp(Sqrt(x)) on the CPU
static inline Float3 F3Min(Float3 a, Float3 b) { return Float3(XMin(a.X, b.X), XMin(a.Y, b.Y), XMin(a.Z, b.Z)); } // :67-70
static inline Float3 F3Max(Float3 a, Float3 b) { return Float3(XMax(a.X, b.X), XMax(a.Y, b.Y), XMax(a.Z, b.Z)); } // :73-76
static inline Float3 Cross(Float3 a, Float3 b) {                                                         // :79-82
    return Float3(a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X);
}
static inline float Dot(Float3 a, Float3 b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }                // :85-88
static inline Float3 Normalize(Float3 v) {                                                               // :91-95
    float inv = XRsqrt(XMax(1e-20f, v.X * v.X + v.Y * v.Y + v.Z * v.Z));
    return Float3(v.X * inv, v.Y * inv, v.Z * inv);
}
static inline float Length(Float3 v) { return sqrtf(v.X * v.X + v.Y * v.Y + v.Z * v.Z); }                // :104-107
static inline Float3 Center(Float3 a, Float3 b) { return Float3(0.5f * (a.X + b.X), 0.5f * (a.Y + b.Y), 0.5f * (a.Z + b.Z)); } // :110-113

// ------------------------------------------------------------------ portable transcendentals
// Stand-ins for XMath.Sin/Cos/Atan2/Acos (call sites: Engine/RTRay.cs:592-593,
// Engine/SceneDeviceViews.cs:152-153).  Cephes-style single precision kernels written with
// + - * / sqrt only, so the CUDA side can evaluate the identical operation sequence with
// round-to-nearest intrinsics and get the same bits.
static const float ORC_PI_F = 3.14159265358979323846f;
static inline void orc_sincos(float x, float* s, float* c) {
#ifdef ORC_LIBM
    *s = sinf(x); *c = cosf(x);
#else
    // argument reduction to [-pi/4, pi/4] by quadrants of pi/2 (Cody-Waite, 3 constants);
    // callers pass x in [0, 2*pi] so k is 0..4 and k*DP1 is exact.
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
#endif
}
static inline float orc_atan_pos(float a) {   // atan for a >= 0
    float y0;
    if (a > 2.414213562373095f) { y0 = 1.5707963267948966f; a = -(1.0f / a); }
    else if (a > 0.4142135623730950f) { y0 = 0.7853981633974483f; a = (a - 1.0f) / (a + 1.0f); }
    else y0 = 0.0f;
    float z = a * a;
    float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * a + a;
    return y0 + p;
}
static inline float orc_atan2(float y, float x) {
#ifdef ORC_LIBM
    return atan2f(y, x);
#else
    if (x == 0.0f) {
        if (y > 0.0f) return 1.5707963267948966f;
        if (y < 0.0f) return -1.5707963267948966f;
        return 0.0f;
    }
    float q = y / x;
    float a = orc_atan_pos(fabsf(q));
    if (q < 0.0f) a = -a;
    if (x < 0.0f) a = (y >= 0.0f) ? (a + ORC_PI_F) : (a - ORC_PI_F);
    return a;
#endif
}
static inline float orc_acos(float x) {   // x already clamped to [-1,1] by the caller
#ifdef ORC_LIBM
    return acosf(x);
#else
    float a = fabsf(x);
    float r;
    if (a > 0.5f) {
        float z = 0.5f * (1.0f - a);
        float s = sqrtf(z);
        float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * s + s;
        r = 2.0f * p;                       // acos(|x|)
        if (x < 0.0f) r = ORC_PI_F - r;
    } else {
        float z = x * x;
        float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * x + x;
        r = 1.5707963267948966f - p;
    }
    return r;
#endif
}
static inline float orc_tan(float x) { float s, c; orc_sincos(x, &s, &c); return s / c; }
// Stand-in for XMath.Pow (call sites: Engine/RTTaa.cs:241-243,254-256), x > 0: exp2(y * log2 x).
//   log2: x = 2^e m with m in [sqrt(1/2), sqrt(2)]; ln m = 2s(1 + z/3 + z^2/5 + z^3/7 + z^4/9), s = (m-1)/(m+1), z = s^2.
//   exp2: t = k + r, |r| <= 1/2; e^(r ln2) from its degree-7 Taylor polynomial (Horner), times 2^k through the exponent field.
// Only + - * / (documented pinning like orc_sincos; build with -DORC_LIBM to use powf).
static inline float orc_log2(float x) {
    uint32_t bits; memcpy(&bits, &x, 4);
    int e = (int)((bits >> 23) & 0xFFu) - 127;
    uint32_t mb = (bits & 0x007FFFFFu) | 0x3F800000u;
    float m; memcpy(&m, &mb, 4);
    if (m > 1.41421356f) { m = m * 0.5f; e = e + 1; }
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float p = (((0.1111111111f * z + 0.1428571429f) * z + 0.2f) * z + 0.3333333333f) * z + 1.0f;
    float ln = 2.0f * s * p;
    return (float)e + ln * 1.4426950408889634f;
}
static inline float orc_exp2(float t) {
    if (t < -126.0f) return 0.0f;
    if (t > 127.0f) t = 127.0f;
    float kf = floorf(t + 0.5f);
    float r = t - kf;
    float u = r * 0.6931471805599453f;
    float p = ((((((u * (1.0f / 7.0f) + 1.0f) * u * (1.0f / 6.0f) + 1.0f) * u * 0.2f + 1.0f) * u * 0.25f + 1.0f) * u * (1.0f / 3.0f) + 1.0f) * u * 0.5f + 1.0f) * u + 1.0f;
    uint32_t sb = (uint32_t)((int)kf + 127) << 23;
    float sc; memcpy(&sc, &sb, 4);
    return p * sc;
}
static inline float orc_pow(float x, float y) {
#ifdef ORC_LIBM
    return powf(x, y);
#else
    return x > 0.0f ? orc_exp2(y * orc_log2(x)) : 0.0f;
#endif
}

// ------------------------------------------------------------------ POD layouts
struct Float2 { float X, Y; };                                        // Engine/MeshLoaderOBJ.cs:33
struct Affine3x4 {                                                    // Engine/Affine3x4.cs:3-7
    float m00, m01, m02, m03, m10, m11, m12, m13, m20, m21, m22, m23;
};
static inline Affine3x4 AffineIdentity() { Affine3x4 a; memset(&a, 0, sizeof(a)); a.m00 = 1; a.m11 = 1; a.m22 = 1; return a; } // :9-14
struct MaterialRecord {                                               // Engine/MeshLoaderOBJ.cs:44-63
    Float3 Kd; int HasDiffuseMap; int DiffuseTexIndex; int Shading; float IOR;
    int HasAlphaMap; int AlphaTexIndex; int TwoSided; float AlphaCutoff;
};
struct Sphere {                                                       // Engine/Sphere.cs:3-15
    Float3 center; float radius; Float3 albedo; MaterialRecord material; int shading; float ior;
};
enum { SHADING_LAMBERT = 0, SHADING_MIRROR = 1, SHADING_GLASS = 2 };
struct BvhNode {                                                      // TLASNode/BLASNode Engine/Scene.cs:705-739
    Float3 boundsMin, boundsMax; int left, right, first, count, skipIndex;
};
enum { BLAS_SPHERESET = 1, BLAS_TRIMESH = 2 };                        // Engine/Scene.cs:703
struct InstanceRecord {                                               // Engine/Scene.cs:716-728
    int type, blasRoot, blasNodeCount, primIndexFirst, primIndexCount;
    Affine3x4 objectToWorld, worldToObject; float uniformScale; Float3 worldBoundsMin, worldBoundsMax;
};
struct MeshTri { int i0, i1, i2; };
struct MeshTriUV { int t0, t1, t2; };
struct RGBA32 { uint8_t R, G, B, A; };
struct TexInfo { int Offset, Width, Height; };
struct Camera {                                                       // Engine/Camera.cs:5-17
    Float3 origin, lowerLeft, horizontal, vertical, forward, right, up; float aspect, fovYRadians;
};
static_assert(sizeof(Float3) == 12 && sizeof(BvhNode) == 44 && sizeof(InstanceRecord) == 144, "layout");
static_assert(sizeof(MaterialRecord) == 44 && sizeof(Sphere) == 80 && sizeof(Camera) == 92, "layout");

// ------------------------------------------------------------------ Engine/RTUtils.cs
struct Ray { Float3 origin, dir, invDir; };                          // :6-10
static inline Float3 InvDir(Float3 d) {                               // Engine/RTRay.cs:548-549
    return Float3(1.0f / (d.X != 0.0f ? d.X : 1e-8f), 1.0f / (d.Y != 0.0f ? d.Y : 1e-8f), 1.0f / (d.Z != 0.0f ? d.Z : 1e-8f));
}
static inline Ray GenerateRay(const Camera& cam, float u, float v) { // :13-17
    Float3 dir = Normalize(cam.lowerLeft + cam.horizontal * u + cam.vertical * v - cam.origin);
    Ray r; r.origin = cam.origin; r.dir = dir;
    r.invDir = Float3(1.0f / (dir.X != 0.0f ? dir.X : 1e-8f), 1.0f / (dir.Y != 0.0f ? dir.Y : 1e-8f), 1.0f / (dir.Z != 0.0f ? dir.Z : 1e-8f));
    return r;
}
static inline uint32_t RotateLeft(uint32_t v, int r) { return (v << (r & 31)) | (v >> ((32 - r) & 31)); } // :100-103
struct RNG {                                                          // :20-138
    uint32_t state;
    static RNG Create(uint32_t seed) { RNG r; r.state = (seed == 0u) ? 1u : seed; return r; }        // :25-30
    uint32_t NextUInt() {                                                                             // :33-42
        uint32_t x = state; x ^= x << 13; x ^= x >> 17; x ^= x << 5; state = (x != 0u) ? x : 1u; return state;
    }
    float NextFloat() { uint32_t u = NextUInt(); return (float)(u & 0x00FFFFFFu) * (1.0f / 16777216.0f); } // :45-49
    static uint32_t SplitMix32(uint64_t x) {                                                          // :54-62
        x += 0x9E3779B97F4A7C15ULL;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
        x ^= (x >> 31);
        return (uint32_t)(x ^ (x >> 32));
    }
    static uint32_t PcgPermute(uint32_t x) {                                                          // :65-74
        x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x;
    }
    static uint32_t Hash32(uint32_t x) {                                                              // :77-84
        x ^= x >> 17; x *= 0xED5AD4BBu; x ^= x >> 11; x *= 0xAC4C1B51u; x ^= x >> 15; x *= 0x31848BABu; x ^= x >> 14; return x;
    }
    static uint32_t MakeSeed32(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {                      // :87-97
        uint64_t lane0 = ((uint64_t)a << 32) | b;
        uint64_t lane1 = ((uint64_t)c << 32) | d;
        uint32_t s0 = SplitMix32(lane0 ^ 0xD1B54A32D192ED03ULL);
        uint32_t s1 = SplitMix32(lane1 ^ 0x94D049BB133111EBULL);
        uint32_t s = PcgPermute(s0 ^ (RotateLeft(s1, 13) + 0x9E3779B1u));
        s |= 1u;
        return s;
    }
    static RNG CreateFromPixel(int px_, int py_, int frame, uint32_t sample, uint32_t salt, int lockNoise) { // :116-137
        uint32_t px = (uint32_t)px_, py = (uint32_t)py_;
        uint32_t f = (lockNoise != 0) ? 0u : (uint32_t)frame;
        uint32_t ln = (uint32_t)lockNoise;
        uint32_t lnMix0 = (lockNoise != 0) ? (Hash32(ln) ^ (ln * 0x1B873593u)) : 0u;
        uint32_t lnMix1 = (lockNoise != 0) ? (RotateLeft(ln, 7) * 0x85EBCA6Bu) : 0u;
        uint32_t lane0a = px ^ 0xB5297A4Du;
        uint32_t lane0b = (py * 0x68E31DA4u) ^ (f * 0x9E3779B1u + 0x85EBCA6Bu) ^ lnMix0;
        uint32_t lane1a = (sample ^ 0xC2B2AE35u) + RotateLeft(px, 16);
        uint32_t lane1b = ((salt ^ 0x27D4EB2Fu) + RotateLeft(py, 8)) ^ lnMix1;   // C#: '+' binds tighter than '^'
        return Create(MakeSeed32(lane0a, lane0b, lane1a, lane1b));
    }
    static RNG CreateFromIndex1D(int index, int width, int height, int frame, uint32_t sample, uint32_t salt, int lockNoise) { // :108-113
        (void)height;
        uint32_t x = (uint32_t)(index % std::max(1, width));
        uint32_t y = (uint32_t)(index / std::max(1, width));
        return CreateFromPixel((int)x, (int)y, frame, sample, salt, lockNoise);
    }
};

// ------------------------------------------------------------------ counters (per thread, summed)
struct Counters {
    uint64_t raysPrimary = 0, raysBounce = 0, raysShadow = 0, nodes = 0, tris = 0, spheres = 0;
};
static thread_local Counters* g_cnt = nullptr;
#define CNT(field) do { if (g_cnt) g_cnt->field++; } while (0)

// ------------------------------------------------------------------ scene (host side of Engine/Scene.cs)
struct Scene {
    std::vector<BvhNode> tlasNodes; std::vector<int> tlasInstanceIndices; std::vector<InstanceRecord> instances;
    std::vector<BvhNode> blasNodes; std::vector<int> spherePrimIdx; std::vector<Sphere> spheres;
    std::vector<int> triPrimIdx; std::vector<Float3> meshPositions; std::vector<MeshTri> meshTris;
    std::vector<Float2> meshTexcoords; std::vector<MeshTriUV> meshTriUVs; std::vector<int> triMatIndex;
    std::vector<MaterialRecord> materials; std::vector<TexInfo> texInfos; std::vector<RGBA32> texels;
    long sortTies = 0;   // number of equal-key neighbours met by the builders' sorts (Array.Sort is unstable)
    bool noCull = false; // debugging aid: treat every IntersectAABB as true (culling-free reference answer)
    bool triMaterials = false; // RT_FLAG_TRI_MATERIALS extension

    // "device views" after UploadAll: AllocateOrEmpty gives empty arrays one zeroed element (Engine/Scene.cs:370-377)
    int texInfosLength() const { return texInfos.empty() ? 1 : (int)texInfos.size(); }
    TexInfo texInfoAt(int i) const { if (texInfos.empty()) { TexInfo z = {0, 0, 0}; return z; } return texInfos[i]; }

    void Clear() {   // Engine/Scene.cs:85-96
        tlasNodes.clear(); tlasInstanceIndices.clear(); instances.clear(); blasNodes.clear();
        spherePrimIdx.clear(); spheres.clear(); triPrimIdx.clear(); meshPositions.clear(); meshTris.clear();
        meshTexcoords.clear(); meshTriUVs.clear(); triMatIndex.clear(); materials.clear(); texInfos.clear(); texels.clear();
        sortTies = 0;
    }
    int AddTexture(int w, int h, const RGBA32* px) {
        int offset = (int)texels.size();
        texels.insert(texels.end(), px, px + (size_t)w * h);
        TexInfo ti = {offset, w, h}; texInfos.push_back(ti);
        return (int)texInfos.size() - 1;
    }
    int AddCheckerTexture(int w, int h, int step, RGBA32 c0, RGBA32 c1) {   // Engine/Scene.cs:98-109
        int offset = (int)texels.size();
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
            bool a = (((x / step) + (y / step)) & 1) == 0;
            texels.push_back(a ? c0 : c1);
        }
        TexInfo ti = {offset, w, h}; texInfos.push_back(ti);
        return (int)texInfos.size() - 1;
    }
    int AddSphere(const Sphere& s) {   // Engine/Scene.cs:315-321
        int id = (int)spheres.size(); spheres.push_back(s); spherePrimIdx.push_back(id); return id;
    }

    static Float3 TransformPoint(const Affine3x4& m, Float3 p) {   // Engine/Scene.cs:640-645
        return Float3(m.m00 * p.X + m.m01 * p.Y + m.m02 * p.Z + m.m03, m.m10 * p.X + m.m11 * p.Y + m.m12 * p.Z + m.m13, m.m20 * p.X + m.m21 * p.Y + m.m22 * p.Z + m.m23);
    }
    static Float3 TransformVector(const Affine3x4& m, Float3 v) {  // Engine/Scene.cs:647-652
        return Float3(m.m00 * v.X + m.m01 * v.Y + m.m02 * v.Z, m.m10 * v.X + m.m11 * v.Y + m.m12 * v.Z, m.m20 * v.X + m.m21 * v.Y + m.m22 * v.Z);
    }
    static void TransformAABB(const Affine3x4& m, Float3 bmin, Float3 bmax, Float3* outMin, Float3* outMax) { // :560-580
        Float3 c[8] = {Float3(bmin.X, bmin.Y, bmin.Z), Float3(bmax.X, bmin.Y, bmin.Z), Float3(bmin.X, bmax.Y, bmin.Z), Float3(bmin.X, bmin.Y, bmax.Z),
                       Float3(bmax.X, bmax.Y, bmin.Z), Float3(bmin.X, bmax.Y, bmax.Z), Float3(bmax.X, bmin.Y, bmax.Z), Float3(bmax.X, bmax.Y, bmax.Z)};
        Float3 mn(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f), mx(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
        for (int i = 0; i < 8; i++) { Float3 w = TransformPoint(m, c[i]); mn = F3Min(mn, w); mx = F3Max(mx, w); }
        *outMin = mn; *outMax = mx;
    }
    static Affine3x4 InvertRigidOrUniform(const Affine3x4& m, float* uniformScale) {   // Engine/Scene.cs:616-638
        float sx = Length(Float3(m.m00, m.m10, m.m20));
        float sy = Length(Float3(m.m01, m.m11, m.m21));
        float sz = Length(Float3(m.m02, m.m12, m.m22));
        *uniformScale = (sx + sy + sz) / 3.0f;
        float inv = *uniformScale > 0.0f ? 1.0f / *uniformScale : 1.0f;
        Float3 r0 = Normalize(Float3(m.m00, m.m10, m.m20));
        Float3 r1 = Normalize(Float3(m.m01, m.m11, m.m21));
        Float3 r2 = Normalize(Float3(m.m02, m.m12, m.m22));
        Affine3x4 invM; memset(&invM, 0, sizeof(invM));
        invM.m00 = r0.X * inv; invM.m01 = r1.X * inv; invM.m02 = r2.X * inv; invM.m03 = 0.0f;
        invM.m10 = r0.Y * inv; invM.m11 = r1.Y * inv; invM.m12 = r2.Y * inv; invM.m13 = 0.0f;
        invM.m20 = r0.Z * inv; invM.m21 = r1.Z * inv; invM.m22 = r2.Z * inv; invM.m23 = 0.0f;
        Float3 t(m.m03, m.m13, m.m23);
        Float3 it = TransformVector(invM, t) * -1.0f;
        invM.m03 = it.X; invM.m13 = it.Y; invM.m23 = it.Z;
        return invM;
    }
    void BoundsOfTriangle(int triIndex, Float3* mn, Float3* mx) const {   // Engine/Scene.cs:597-605
        MeshTri tri = meshTris[triIndex];
        Float3 v0 = meshPositions[tri.i0], v1 = meshPositions[tri.i1], v2 = meshPositions[tri.i2];
        *mn = F3Min(v0, F3Min(v1, v2)); *mx = F3Max(v0, F3Max(v1, v2));
    }
    Float3 CenterOfTriangle(int triIndex) const {                         // Engine/Scene.cs:607-614
        MeshTri tri = meshTris[triIndex];
        Float3 v0 = meshPositions[tri.i0], v1 = meshPositions[tri.i1], v2 = meshPositions[tri.i2];
        return Float3((v0.X + v1.X + v2.X) / 3.0f, (v0.Y + v1.Y + v2.Y) / 3.0f, (v0.Z + v1.Z + v2.Z) / 3.0f);
    }

    // Array.Sort(idx, start, count, comparer) is .NET's UNSTABLE introsort (ArraySortHelper<T>.IntrospectiveSort:
    // median-of-three quicksort, insertion sort for partitions <= 16, heapsort at depth 2*(log2(n)+1)).  Restated here
    // (from the published .NET runtime sources, not vendored under /root/reference: UNPINNED) so equal keys land where
    // the reference's sort would put them; sortTies counts equal neighbours so callers can see whether that matters.
    template <class KeyFn> struct NetSort {
        KeyFn key;
        int cmp(int a, int b) { float va = key(a), vb = key(b); if (va < vb) return -1; if (va > vb) return 1; return 0; }
        void SwapIfGreater(int* k, int i, int j) { if (cmp(k[i], k[j]) > 0) { int t = k[i]; k[i] = k[j]; k[j] = t; } }
        void InsertionSort(int* k, int n) {
            for (int i = 0; i < n - 1; i++) { int t = k[i + 1]; int j = i; while (j >= 0 && cmp(t, k[j]) < 0) { k[j + 1] = k[j]; j--; } k[j + 1] = t; }
        }
        void DownHeap(int* k, int i, int n) {
            int d = k[i - 1];
            while (i <= n >> 1) {
                int child = 2 * i;
                if (child < n && cmp(k[child - 1], k[child]) < 0) child++;
                if (!(cmp(d, k[child - 1]) < 0)) break;
                k[i - 1] = k[child - 1]; i = child;
            }
            k[i - 1] = d;
        }
        void HeapSort(int* k, int n) {
            for (int i = n >> 1; i >= 1; i--) DownHeap(k, i, n);
            for (int i = n; i > 1; i--) { int t = k[0]; k[0] = k[i - 1]; k[i - 1] = t; DownHeap(k, 1, i - 1); }
        }
        int PickPivotAndPartition(int* k, int n) {
            int hi = n - 1, middle = hi >> 1;
            SwapIfGreater(k, 0, middle); SwapIfGreater(k, 0, hi); SwapIfGreater(k, middle, hi);
            int pivot = k[middle];
            { int t = k[middle]; k[middle] = k[hi - 1]; k[hi - 1] = t; }
            int left = 0, right = hi - 1;
            while (left < right) {
                while (cmp(k[++left], pivot) < 0) {}
                while (cmp(pivot, k[--right]) < 0) {}
                if (left >= right) break;
                int t = k[left]; k[left] = k[right]; k[right] = t;
            }
            if (left != hi - 1) { int t = k[left]; k[left] = k[hi - 1]; k[hi - 1] = t; }
            return left;
        }
        void IntroSort(int* k, int n, int depthLimit) {
            int partitionSize = n;
            while (partitionSize > 1) {
                if (partitionSize <= 16) {
                    if (partitionSize == 2) { SwapIfGreater(k, 0, 1); return; }
                    if (partitionSize == 3) { SwapIfGreater(k, 0, 1); SwapIfGreater(k, 0, 2); SwapIfGreater(k, 1, 2); return; }
                    InsertionSort(k, partitionSize); return;
                }
                if (depthLimit == 0) { HeapSort(k, partitionSize); return; }
                depthLimit--;
                int p = PickPivotAndPartition(k, partitionSize);
                IntroSort(k + p + 1, partitionSize - (p + 1), depthLimit);
                partitionSize = p;
            }
        }
    };
    template <class KeyFn> void SortRange(int* idx, int start, int count, KeyFn key) {
        if (count > 1) {
            int lg = 0; for (unsigned v = (unsigned)count; v >>= 1;) lg++;
            NetSort<KeyFn> ns{key};
            ns.IntroSort(idx + start, count, 2 * (lg + 1));
        }
        for (int i = 1; i < count; i++) if (key(idx[start + i]) == key(idx[start + i - 1])) sortTies++;
    }

    // Engine/Scene.cs:405-467.  bminPre/bmaxPre are indexed by POSITION i (quirk 1 of SURVEY §8a) exactly as the reference does.
    int BuildBLASNodeRecursive(int* idx, int start, int count, const Float3* bminPre, const Float3* bmaxPre, int parentSkip, bool spheresMode) {
        int nodeIndex = (int)blasNodes.size();
        BvhNode node; node.first = -1; node.count = 0; node.left = -1; node.right = -1; node.skipIndex = parentSkip;
        Float3 nbMin(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f), nbMax(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
        std::vector<int>& primIdx = spheresMode ? spherePrimIdx : triPrimIdx;
        if (bminPre != nullptr) {
            for (int i = start; i < start + count; i++) { nbMin = F3Min(nbMin, bminPre[i]); nbMax = F3Max(nbMax, bmaxPre[i]); }
        } else {
            for (int i = start; i < start + count; i++) {
                int triIndex = primIdx[idx[i]];
                Float3 mn, mx; BoundsOfTriangle(triIndex, &mn, &mx);
                nbMin = F3Min(nbMin, mn); nbMax = F3Max(nbMax, mx);
            }
        }
        node.boundsMin = nbMin; node.boundsMax = nbMax;
        blasNodes.push_back(node);

        const int LeafThreshold = 4;
        if (count <= LeafThreshold) {
            int leafStart = (int)primIdx.size();
            for (int i = start; i < start + count; i++) primIdx.push_back(primIdx[idx[i]]);
            blasNodes[nodeIndex].first = leafStart; blasNodes[nodeIndex].count = count; blasNodes[nodeIndex].skipIndex = parentSkip;
            return nodeIndex;
        }
        Float3 extent = nbMax - nbMin;
        int axis = 0;
        if (extent.Y > extent.X && extent.Y >= extent.Z) axis = 1;
        else if (extent.Z > extent.X && extent.Z >= extent.Y) axis = 2;

        if (spheresMode) {   // BLASPrimComparatorSpheres :512-526
            SortRange(idx, start, count, [&](int a) { const Sphere& s = spheres[primIdx[a]]; return axis == 0 ? s.center.X : (axis == 1 ? s.center.Y : s.center.Z); });
        } else {             // BLASPrimComparatorTris :528-543
            SortRange(idx, start, count, [&](int a) { Float3 c = CenterOfTriangle(primIdx[a]); return axis == 0 ? c.X : (axis == 1 ? c.Y : c.Z); });
        }
        int mid = start + (count >> 1);
        int rightRoot = BuildBLASNodeRecursive(idx, mid, count - (mid - start), bminPre, bmaxPre, parentSkip, spheresMode);
        int leftRoot = BuildBLASNodeRecursive(idx, start, mid - start, bminPre, bmaxPre, rightRoot, spheresMode);
        blasNodes[nodeIndex].left = leftRoot; blasNodes[nodeIndex].right = rightRoot; blasNodes[nodeIndex].skipIndex = parentSkip;
        return nodeIndex;
    }
    void BuildBLAS_Spheres(int primStart, int primCount) {   // Engine/Scene.cs:381-396
        std::vector<int> idx((size_t)primCount);
        for (int i = 0; i < primCount; i++) idx[i] = primStart + i;
        std::vector<Float3> bmin((size_t)primCount), bmax((size_t)primCount);
        for (int i = 0; i < primCount; i++) {
            const Sphere& s = spheres[spherePrimIdx[primStart + i]];
            bmin[i] = Float3(s.center.X - s.radius, s.center.Y - s.radius, s.center.Z - s.radius);
            bmax[i] = Float3(s.center.X + s.radius, s.center.Y + s.radius, s.center.Z + s.radius);
        }
        BuildBLASNodeRecursive(idx.data(), 0, primCount, bmin.data(), bmax.data(), -1, true);
    }
    void BuildBLAS_Triangles(int primStart, int primCount) { // Engine/Scene.cs:398-403
        std::vector<int> idx((size_t)primCount);
        for (int i = 0; i < primCount; i++) idx[i] = primStart + i;
        BuildBLASNodeRecursive(idx.data(), 0, primCount, nullptr, nullptr, -1, false);
    }
    InstanceRecord BuildSphereInstance(const int* sphereIds, int n, const Affine3x4& objectToWorld) {   // Engine/Scene.cs:323-356
        Float3 bmin(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f), bmax(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
        for (int i = 0; i < n; i++) {
            const Sphere& s = spheres[sphereIds[i]];
            bmin = F3Min(bmin, Float3(s.center.X - s.radius, s.center.Y - s.radius, s.center.Z - s.radius));
            bmax = F3Max(bmax, Float3(s.center.X + s.radius, s.center.Y + s.radius, s.center.Z + s.radius));
        }
        int primStart = sphereIds[0], primCount = n;
        int blasStart = (int)blasNodes.size();
        BuildBLAS_Spheres(primStart, primCount);
        int blasCount = (int)blasNodes.size() - blasStart;
        Float3 wmin, wmax; TransformAABB(objectToWorld, bmin, bmax, &wmin, &wmax);
        float uniScale; Affine3x4 worldToObject = InvertRigidOrUniform(objectToWorld, &uniScale);
        InstanceRecord inst; memset(&inst, 0, sizeof(inst));
        inst.type = BLAS_SPHERESET; inst.blasRoot = blasStart; inst.blasNodeCount = blasCount;
        inst.primIndexFirst = primStart; inst.primIndexCount = primCount;
        inst.objectToWorld = objectToWorld; inst.worldToObject = worldToObject; inst.uniformScale = uniScale;
        inst.worldBoundsMin = wmin; inst.worldBoundsMax = wmax;
        return inst;
    }
    int BuildTLASNodeRecursive(int* idx, int start, int count, int parentSkip) {   // Engine/Scene.cs:469-510
        int nodeIndex = (int)tlasNodes.size();
        BvhNode node; node.first = -1; node.count = 0; node.left = -1; node.right = -1; node.skipIndex = parentSkip;
        Float3 nbMin(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f), nbMax(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
        for (int i = start; i < start + count; i++) {
            const InstanceRecord& r = instances[idx[i]];
            nbMin = F3Min(nbMin, r.worldBoundsMin); nbMax = F3Max(nbMax, r.worldBoundsMax);
        }
        node.boundsMin = nbMin; node.boundsMax = nbMax;
        tlasNodes.push_back(node);
        const int LeafThreshold = 2;
        if (count <= LeafThreshold) {
            tlasNodes[nodeIndex].first = start; tlasNodes[nodeIndex].count = count; tlasNodes[nodeIndex].skipIndex = parentSkip;
            return nodeIndex;
        }
        Float3 extent = nbMax - nbMin;
        int axis = 0;
        if (extent.Y > extent.X && extent.Y >= extent.Z) axis = 1;
        else if (extent.Z > extent.X && extent.Z >= extent.Y) axis = 2;
        SortRange(idx, start, count, [&](int a) {   // TLASInstComparator :545-558
            Float3 c = Center(instances[a].worldBoundsMin, instances[a].worldBoundsMax);
            return axis == 0 ? c.X : (axis == 1 ? c.Y : c.Z);
        });
        int mid = start + (count >> 1);
        int rightRoot = BuildTLASNodeRecursive(idx, mid, count - (mid - start), parentSkip);
        int leftRoot = BuildTLASNodeRecursive(idx, start, mid - start, rightRoot);
        tlasNodes[nodeIndex].left = leftRoot; tlasNodes[nodeIndex].right = rightRoot; tlasNodes[nodeIndex].skipIndex = parentSkip;
        return nodeIndex;
    }
    void RebuildTLAS() {   // Engine/Scene.cs:358-368
        int n = (int)instances.size();
        tlasInstanceIndices.resize((size_t)n);
        for (int i = 0; i < n; i++) tlasInstanceIndices[i] = i;
        tlasNodes.clear();
        if (n > 0) BuildTLASNodeRecursive(tlasInstanceIndices.data(), 0, n, -1);
    }
    void ComputeMeshBounds(int baseTri, int nTris, Float3* bmin, Float3* bmax) const {   // Engine/Scene.cs:582-595
        *bmin = Float3(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f); *bmax = Float3(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
        for (int i = 0; i < nTris; i++) {
            MeshTri t = meshTris[baseTri + i];
            Float3 v0 = meshPositions[t.i0], v1 = meshPositions[t.i1], v2 = meshPositions[t.i2];
            *bmin = F3Min(*bmin, F3Min(v0, F3Min(v1, v2))); *bmax = F3Max(*bmax, F3Max(v0, F3Max(v1, v2)));
        }
    }
    // Engine/Scene.cs:144-256 after MeshLoaderOBJ.Load: mesh arrays arrive already decoded; material texture
    // indices are GLOBAL texInfos indices (the per-mesh texture remap :180-227 is asset-path logic, out of scope).
    void AddMeshInstance(const Float3* pos, int nPos, const MeshTri* tris, int nTris, const Float2* uv, int nUV,
                         const MeshTriUV* triUVs, const int* triMat, const MaterialRecord* mats, int nMats, const Affine3x4& objectToWorld) {
        int baseVertex = (int)meshPositions.size(), baseTri = (int)meshTris.size(), baseUV = (int)meshTexcoords.size(), baseMat = (int)materials.size();
        meshPositions.insert(meshPositions.end(), pos, pos + nPos);
        meshTexcoords.insert(meshTexcoords.end(), uv, uv + nUV);
        for (int i = 0; i < nTris; i++) {
            MeshTri t = tris[i]; t.i0 += baseVertex; t.i1 += baseVertex; t.i2 += baseVertex; meshTris.push_back(t);
            MeshTriUV tuv = triUVs[i]; tuv.t0 += baseUV; tuv.t1 += baseUV; tuv.t2 += baseUV; meshTriUVs.push_back(tuv);
            triMatIndex.push_back(baseMat + (triMat ? triMat[i] : 0));
            triPrimIdx.push_back(baseTri + i);
        }
        materials.insert(materials.end(), mats, mats + nMats);
        int blasStart = (int)blasNodes.size();
        BuildBLAS_Triangles(baseTri, nTris);
        int blasCount = (int)blasNodes.size() - blasStart;
        Float3 bmin, bmax; ComputeMeshBounds(baseTri, nTris, &bmin, &bmax);
        Float3 wmin, wmax; TransformAABB(objectToWorld, bmin, bmax, &wmin, &wmax);
        float uniScale; Affine3x4 worldToObject = InvertRigidOrUniform(objectToWorld, &uniScale);
        InstanceRecord r; memset(&r, 0, sizeof(r));
        r.type = BLAS_TRIMESH; r.blasRoot = blasStart; r.blasNodeCount = blasCount; r.primIndexFirst = baseTri; r.primIndexCount = nTris;
        r.objectToWorld = objectToWorld; r.worldToObject = worldToObject; r.uniformScale = uniScale; r.worldBoundsMin = wmin; r.worldBoundsMax = wmax;
        instances.push_back(r);
        RebuildTLAS();
    }
    void BuildDefaultScene() {   // Engine/Scene.cs:83-142 (TryAddSponzaFromKnownLocations finds nothing: no .obj is shipped)
        Clear();
        RGBA32 w255 = {255, 255, 255, 255}, g20 = {20, 20, 20, 255}, b0 = {40, 40, 200, 255}, y0 = {200, 200, 40, 255};
        int checker0 = AddCheckerTexture(256, 256, 16, w255, g20);
        int checker1 = AddCheckerTexture(256, 256, 8, b0, y0);
        auto mk = [](Float3 kd, int hasMap, int tex) { MaterialRecord m; m.Kd = kd; m.HasDiffuseMap = hasMap; m.DiffuseTexIndex = tex; m.Shading = SHADING_LAMBERT; m.IOR = 1.0f; m.HasAlphaMap = 0; m.AlphaTexIndex = -1; m.AlphaCutoff = 0.5f; m.TwoSided = 0; return m; };
        MaterialRecord matGround = mk(Float3(1, 1, 1), 1, checker0), matRed = mk(Float3(0.8f, 0.3f, 0.3f), 0, -1), matGreen = mk(Float3(0.3f, 0.8f, 0.3f), 0, -1);
        MaterialRecord matTex = mk(Float3(1, 1, 1), 1, checker1), matWhite = mk(Float3(1, 1, 1), 0, -1);
        auto sp = [](Float3 c, float r, Float3 alb, MaterialRecord m, int shading, float ior) { Sphere s; s.center = c; s.radius = r; s.albedo = alb; s.material = m; s.shading = shading; s.ior = ior; return s; };
        int ground = AddSphere(sp(Float3(0.0f, -1000.5f, 0.0f), 1000.0f, Float3(1, 1, 1), matGround, SHADING_LAMBERT, 1.0f));
        int s0 = AddSphere(sp(Float3(-0.9f, 0.5f, -0.2f), 0.5f, Float3(0.8f, 0.3f, 0.3f), matRed, SHADING_LAMBERT, 1.0f));
        int s1 = AddSphere(sp(Float3(0.9f, 0.35f, 0.2f), 0.35f, Float3(0.3f, 0.8f, 0.3f), matGreen, SHADING_LAMBERT, 1.0f));
        int s2 = AddSphere(sp(Float3(0.0f, 0.75f, 0.6f), 0.75f, Float3(1, 1, 1), matTex, SHADING_LAMBERT, 1.0f));
        int sMirror = AddSphere(sp(Float3(-1.8f, 0.5f, 0.8f), 0.5f, Float3(1, 1, 1), matWhite, SHADING_MIRROR, 1.0f));
        int sGlass = AddSphere(sp(Float3(1.8f, 0.5f, -0.8f), 0.5f, Float3(1, 1, 1), matWhite, SHADING_GLASS, 1.5f));
        int ids[6] = {ground, s0, s1, s2, sMirror, sGlass};
        Affine3x4 I = AffineIdentity();
        for (int i = 0; i < 6; i++) instances.push_back(BuildSphereInstance(&ids[i], 1, I));
        RebuildTLAS();
    }
};

// ------------------------------------------------------------------ Engine/SceneDeviceViews.cs
struct HitInfo {     // the out-parameters of TraceClosest (+ parity extras that do not alter semantics)
    float closestT; Float3 bestNormal, bestAlbedo; int bestObjId, bestShade; float bestIor;
    int instIdx, primIdx;   // extras: instance index and sphere index / global tri index of the winning hit
};

static inline Float3 DV_TransformPoint(const Affine3x4& m, Float3 p) {   // :484-487
    return Float3(m.m00 * p.X + m.m01 * p.Y + m.m02 * p.Z + m.m03, m.m10 * p.X + m.m11 * p.Y + m.m12 * p.Z + m.m13, m.m20 * p.X + m.m21 * p.Y + m.m22 * p.Z + m.m23);
}
static inline Float3 DV_TransformVector(const Affine3x4& m, Float3 v) {  // :490-493
    return Float3(m.m00 * v.X + m.m01 * v.Y + m.m02 * v.Z, m.m10 * v.X + m.m11 * v.Y + m.m12 * v.Z, m.m20 * v.X + m.m21 * v.Y + m.m22 * v.Z);
}
static inline Ray TransformRay(const Affine3x4& m, const Ray& w) {       // :475-481
    Ray r; r.origin = DV_TransformPoint(m, w.origin); r.dir = DV_TransformVector(m, w.dir); r.invDir = InvDir(r.dir); return r;
}
static inline bool IntersectAABB(const Ray& ray, Float3 bmin, Float3 bmax, float tMin, float tMax) {   // :496-514
    float t1 = (bmin.X - ray.origin.X) * ray.invDir.X;
    float t2 = (bmax.X - ray.origin.X) * ray.invDir.X;
    float tmin = XMin(t1, t2), tmax = XMax(t1, t2);
    t1 = (bmin.Y - ray.origin.Y) * ray.invDir.Y; t2 = (bmax.Y - ray.origin.Y) * ray.invDir.Y;
    tmin = XMax(tmin, XMin(t1, t2)); tmax = XMin(tmax, XMax(t1, t2));
    t1 = (bmin.Z - ray.origin.Z) * ray.invDir.Z; t2 = (bmax.Z - ray.origin.Z) * ray.invDir.Z;
    tmin = XMax(tmin, XMin(t1, t2)); tmax = XMin(tmax, XMax(t1, t2));
    return tmax >= XMax(tmin, tMin) && tmin <= tMax;
}
static inline bool IntersectSphere(const Ray& ray, const Sphere& s, float* t, Float3* n) {   // :517-537
    Float3 oc = ray.origin - s.center;
    float a = Dot(ray.dir, ray.dir);
    float b = 2.0f * Dot(oc, ray.dir);
    float c = Dot(oc, oc) - s.radius * s.radius;
    float disc = b * b - 4.0f * a * c;
    if (disc < 0.0f) { *t = 0.0f; *n = Float3(); return false; }
    float sqrtD = sqrtf(disc);
    float t0 = (-b - sqrtD) / (2.0f * a);
    float t1 = (-b + sqrtD) / (2.0f * a);
    *t = t0;
    if (*t < 0.001f) { *t = t1; if (*t < 0.001f) { *n = Float3(); return false; } }
    Float3 p = ray.origin + ray.dir * (*t);
    *n = Normalize(p - s.center);
    return true;
}
static inline bool IntersectTriangleMT_Bary(const Ray& ray, Float3 v0, Float3 v1, Float3 v2, float* t, Float3* n, float* bu, float* bv) { // :540-558
    Float3 e1 = v1 - v0, e2 = v2 - v0;
    Float3 p = Cross(ray.dir, e2);
    float det = Dot(e1, p);
    if (fabsf(det) < 1e-8f) { *t = 0; *n = Float3(); *bu = 0; *bv = 0; return false; }
    float invDet = 1.0f / det;
    Float3 tv = ray.origin - v0;
    *bu = Dot(tv, p) * invDet;
    if (*bu < 0.0f || *bu > 1.0f) { *t = 0; *n = Float3(); *bv = 0; return false; }
    Float3 q = Cross(tv, e1);
    *bv = Dot(ray.dir, q) * invDet;
    if (*bv < 0.0f || *bu + *bv > 1.0f) { *t = 0; *n = Float3(); return false; }
    *t = Dot(e2, q) * invDet;
    if (*t <= 0.0f) { *n = Float3(); return false; }
    *n = Normalize(Cross(e1, e2));
    return true;
}

struct Views {   // SceneDeviceViews :11-27 bound to a host Scene
    const Scene* sc;

    RGBA32 TexelRaw(TexInfo info, int x, int y) const {   // :330-339
        int w = info.Width, h = info.Height;
        RGBA32 z = {0, 0, 0, 0};
        if (w <= 0 || h <= 0) return z;
        int sx = XMax(0, XMin(w - 1, x)), sy = XMax(0, XMin(h - 1, y));
        int idx = info.Offset + sy * w + sx;
        return sc->texels[(size_t)idx];
    }
    static float Luma01(RGBA32 p) {                        // :342-348
        float r = p.R * (1.0f / 255.0f), g = p.G * (1.0f / 255.0f), b = p.B * (1.0f / 255.0f);
        return 0.2126f * r + 0.7152f * g + 0.0722f * b;
    }
    Float3 TexelRGB(TexInfo info, int x, int y) const {    // :351-355
        RGBA32 p = TexelRaw(info, x, y);
        return Float3(p.R * (1.0f / 255.0f), p.G * (1.0f / 255.0f), p.B * (1.0f / 255.0f));
    }
    Float3 SampleTextureLinear(TexInfo info, float u, float v) const {   // :358-385
        int w = info.Width, h = info.Height;
        if (w <= 0 || h <= 0) return Float3(1, 1, 1);
        float fu = u - floorf(u);
        float fv = 1.0f - (v - floorf(v));
        float x = fu * (float)(w - 1), y = fv * (float)(h - 1);
        int x0 = (int)floorf(x), y0 = (int)floorf(y);
        int x1 = XMin(w - 1, x0 + 1), y1 = XMin(h - 1, y0 + 1);
        float tx = x - (float)x0, ty = y - (float)y0;
        Float3 c00 = TexelRGB(info, x0, y0), c10 = TexelRGB(info, x1, y0), c01 = TexelRGB(info, x0, y1), c11 = TexelRGB(info, x1, y1);
        Float3 cx0 = c00 * (1.0f - tx) + c10 * tx;
        Float3 cx1 = c01 * (1.0f - tx) + c11 * tx;
        return cx0 * (1.0f - ty) + cx1 * ty;
    }
    float SampleMaskLinear(TexInfo info, float u, float v) const {       // :388-415
        int w = info.Width, h = info.Height;
        if (w <= 0 || h <= 0) return 1.0f;
        float fu = u - floorf(u);
        float fv = 1.0f - (v - floorf(v));
        float x = fu * (float)(w - 1), y = fv * (float)(h - 1);
        int x0 = (int)floorf(x), y0 = (int)floorf(y);
        int x1 = XMin(w - 1, x0 + 1), y1 = XMin(h - 1, y0 + 1);
        float tx = x - (float)x0, ty = y - (float)y0;
        float a00 = Luma01(TexelRaw(info, x0, y0)), a10 = Luma01(TexelRaw(info, x1, y0)), a01 = Luma01(TexelRaw(info, x0, y1)), a11 = Luma01(TexelRaw(info, x1, y1));
        float ax0 = a00 * (1.0f - tx) + a10 * tx;
        float ax1 = a01 * (1.0f - tx) + a11 * tx;
        return ax0 * (1.0f - ty) + ax1 * ty;
    }
    float SampleMaskPoint(TexInfo info, float u, float v) const {        // :418-428 (XMath.Round = round-half-to-even)
        int w = info.Width, h = info.Height;
        if (w <= 0 || h <= 0) return 1.0f;
        float fu = u - floorf(u);
        float fv = 1.0f - (v - floorf(v));
        int x = (int)rintf(fu * (float)(w - 1));
        int y = (int)rintf(fv * (float)(h - 1));
        return Luma01(TexelRaw(info, x, y));
    }
    Float3 SampleTextureLinearRGB_A(TexInfo info, float u, float v, float* a) const {   // :431-472
        int w = info.Width, h = info.Height;
        if (w <= 0 || h <= 0) { *a = 1.0f; return Float3(1, 1, 1); }
        float fu = u - floorf(u);
        float fv = 1.0f - (v - floorf(v));
        float x = fu * (float)(w - 1), y = fv * (float)(h - 1);
        int x0 = (int)floorf(x), y0 = (int)floorf(y);
        int x1 = XMin(w - 1, x0 + 1), y1 = XMin(h - 1, y0 + 1);
        float tx = x - (float)x0, ty = y - (float)y0;
        RGBA32 p00 = TexelRaw(info, x0, y0), p10 = TexelRaw(info, x1, y0), p01 = TexelRaw(info, x0, y1), p11 = TexelRaw(info, x1, y1);
        Float3 c00(p00.R * (1.0f / 255.0f), p00.G * (1.0f / 255.0f), p00.B * (1.0f / 255.0f));
        Float3 c10(p10.R * (1.0f / 255.0f), p10.G * (1.0f / 255.0f), p10.B * (1.0f / 255.0f));
        Float3 c01(p01.R * (1.0f / 255.0f), p01.G * (1.0f / 255.0f), p01.B * (1.0f / 255.0f));
        Float3 c11(p11.R * (1.0f / 255.0f), p11.G * (1.0f / 255.0f), p11.B * (1.0f / 255.0f));
        float a00 = p00.A * (1.0f / 255.0f), a10 = p10.A * (1.0f / 255.0f), a01 = p01.A * (1.0f / 255.0f), a11 = p11.A * (1.0f / 255.0f);
        Float3 cx0 = c00 * (1.0f - tx) + c10 * tx;
        Float3 cx1 = c01 * (1.0f - tx) + c11 * tx;
        float ax0 = a00 * (1.0f - tx) + a10 * tx;
        float ax1 = a01 * (1.0f - tx) + a11 * tx;
        *a = ax0 * (1.0f - ty) + ax1 * ty;
        return cx0 * (1.0f - ty) + cx1 * ty;
    }

    bool box(const Ray& r, const BvhNode& n, float tMin, float tMax) const {
        CNT(nodes);
        return IntersectAABB(r, n.boundsMin, n.boundsMax, tMin, tMax) || sc->noCull;
    }

    // :124-170
    bool TraverseBLAS_Sphere(const Ray& rayObj, int blasStart, int blasEnd, float* tClosest, Float3* nObj, Float3* albedo, int* shading, float* ior, int* primOut) const {
        *tClosest = 1e30f; *nObj = Float3(); *albedo = Float3(1, 1, 1); *shading = 0; *ior = 1.0f; *primOut = -1;
        int cur = blasStart;
        while (cur != -1 && cur < blasEnd) {
            const BvhNode& n = sc->blasNodes[(size_t)cur];
            if (box(rayObj, n, 0.001f, *tClosest)) {
                if (n.count > 0) {
                    int end = n.first + n.count;
                    for (int i = n.first; i < end; i++) {
                        int prim = sc->spherePrimIdx[(size_t)i];
                        float t; Float3 nn;
                        CNT(spheres);
                        if (IntersectSphere(rayObj, sc->spheres[(size_t)prim], &t, &nn)) {
                            if (t > 0.001f && t < *tClosest) {
                                *tClosest = t; *nObj = nn;
                                const Sphere& s = sc->spheres[(size_t)prim];
                                Float3 kd = s.material.Kd;
                                Float3 col = (kd.X == 0.0f && kd.Y == 0.0f && kd.Z == 0.0f) ? s.albedo : kd;
                                if (s.material.HasDiffuseMap != 0 && s.material.DiffuseTexIndex >= 0 && s.material.DiffuseTexIndex < sc->texInfosLength()) {
                                    const float PI = 3.14159265358979323846f;
                                    float u = 0.5f + orc_atan2(nn.Z, nn.X) / (2.0f * PI);
                                    float v = orc_acos(XMin(1.0f, XMax(-1.0f, nn.Y))) / PI;
                                    float aTmp;
                                    col = SampleTextureLinearRGB_A(sc->texInfoAt(s.material.DiffuseTexIndex), u, v, &aTmp);
                                }
                                *albedo = col; *shading = s.shading; *ior = s.ior > 0.0f ? s.ior : 1.0f; *primOut = prim;
                            }
                        }
                    }
                    cur = n.skipIndex;
                } else cur = n.left;
            } else cur = n.skipIndex;
        }
        return *tClosest < 1e29f;
    }

    // :173-237
    bool TraverseBLAS_Tri_Textured(const Ray& rayObj, int blasStart, int blasEnd, float* tClosest, Float3* nObj, Float3* albedo, int* triOut, float* buOut, float* bvOut) const {
        *tClosest = 1e30f; *nObj = Float3(); *albedo = Float3(0.85f, 0.85f, 0.85f); *triOut = -1; *buOut = 0; *bvOut = 0;
        int cur = blasStart;
        while (cur != -1 && cur < blasEnd) {
            const BvhNode& n = sc->blasNodes[(size_t)cur];
            if (box(rayObj, n, 0.001f, *tClosest)) {
                if (n.count > 0) {
                    int end = n.first + n.count;
                    for (int i = n.first; i < end; i++) {
                        int triIndex = sc->triPrimIdx[(size_t)i];
                        MeshTri tri = sc->meshTris[(size_t)triIndex];
                        Float3 v0 = sc->meshPositions[(size_t)tri.i0], v1 = sc->meshPositions[(size_t)tri.i1], v2 = sc->meshPositions[(size_t)tri.i2];
                        float t; Float3 nn; float bu, bv;
                        CNT(tris);
                        if (IntersectTriangleMT_Bary(rayObj, v0, v1, v2, &t, &nn, &bu, &bv)) {
                            int midx = sc->triMatIndex[(size_t)triIndex];
                            const MaterialRecord& mat = sc->materials[(size_t)midx];
                            if (t > 0.001f && t < *tClosest) {
                                MeshTriUV tuv = sc->meshTriUVs[(size_t)triIndex];
                                Float2 t0 = sc->meshTexcoords[(size_t)tuv.t0], t1 = sc->meshTexcoords[(size_t)tuv.t1], t2 = sc->meshTexcoords[(size_t)tuv.t2];
                                float w = 1.0f - bu - bv;
                                float uu = t0.X * w + t1.X * bu + t2.X * bv;
                                float vv = t0.Y * w + t1.Y * bu + t2.Y * bv;
                                float alpha = 1.0f;
                                Float3 kdCol = mat.Kd;
                                if (mat.HasDiffuseMap != 0 && mat.DiffuseTexIndex >= 0 && mat.DiffuseTexIndex < sc->texInfosLength())
                                    kdCol = SampleTextureLinear(sc->texInfoAt(mat.DiffuseTexIndex), uu, vv);
                                if (mat.HasAlphaMap != 0 && mat.AlphaTexIndex >= 0 && mat.AlphaTexIndex < sc->texInfosLength())
                                    alpha = SampleMaskLinear(sc->texInfoAt(mat.AlphaTexIndex), uu, vv);
                                if (alpha < mat.AlphaCutoff) { continue; }
                                *tClosest = t; *nObj = nn;
                                if (mat.TwoSided != 0 && Dot(*nObj, rayObj.dir) > 0.0f) *nObj = *nObj * -1.0f;
                                *albedo = kdCol; *triOut = triIndex; *buOut = bu; *bvOut = bv;
                            }
                        }
                    }
                    cur = n.skipIndex;
                } else cur = n.left;
            } else cur = n.skipIndex;
        }
        return *tClosest < 1e29f;
    }

    // :30-86
    bool TraceClosest(const Ray& wray, HitInfo* h) const {
        h->closestT = 1e30f; h->bestNormal = Float3(); h->bestAlbedo = Float3(1, 1, 1); h->bestObjId = -1; h->bestShade = 0; h->bestIor = 1.0f;
        h->instIdx = -1; h->primIdx = -1;
        int cur = sc->tlasNodes.empty() ? -1 : 0;
        while (cur != -1) {
            const BvhNode& n = sc->tlasNodes[(size_t)cur];
            if (box(wray, n, 0.001f, h->closestT)) {
                if (n.count > 0) {
                    int end = n.first + n.count;
                    for (int i = n.first; i < end; i++) {
                        int instIndex = sc->tlasInstanceIndices[(size_t)i];
                        const InstanceRecord& inst = sc->instances[(size_t)instIndex];
                        Ray iray = TransformRay(inst.worldToObject, wray);
                        float scale = inst.uniformScale > 0.0f ? inst.uniformScale : 1.0f;
                        float tObjClosest; Float3 normalObj, albedo; int triLocal; float bu, bv; int shade; float ior; bool hit; int prim = -1;
                        int blasStart = inst.blasRoot, blasEnd = blasStart + inst.blasNodeCount;
                        if (inst.type == BLAS_SPHERESET) {
                            triLocal = -1; bu = 0; bv = 0; shade = 0; ior = 1.0f;
                            hit = TraverseBLAS_Sphere(iray, blasStart, blasEnd, &tObjClosest, &normalObj, &albedo, &shade, &ior, &prim);
                        } else {
                            shade = 0; ior = 1.0f;
                            hit = TraverseBLAS_Tri_Textured(iray, blasStart, blasEnd, &tObjClosest, &normalObj, &albedo, &triLocal, &bu, &bv);
                            prim = triLocal;
                            if (hit && sc->triMaterials) {   // RT_FLAG_TRI_MATERIALS extension (not in the reference, which forces Lambert at :61)
                                const MaterialRecord& m = sc->materials[(size_t)sc->triMatIndex[(size_t)triLocal]];
                                shade = m.Shading; ior = m.IOR > 0.0f ? m.IOR : 1.0f;
                            }
                        }
                        if (hit) {
                            float tWorld = tObjClosest / scale;
                            if (tWorld < h->closestT) {
                                h->closestT = tWorld;
                                h->bestNormal = Normalize(DV_TransformVector(inst.objectToWorld, normalObj));
                                h->bestAlbedo = albedo; h->bestObjId = triLocal; h->bestShade = shade; h->bestIor = ior;
                                h->instIdx = instIndex; h->primIdx = prim;
                            }
                        }
                    }
                    cur = n.skipIndex;
                } else cur = n.left;
            } else cur = n.skipIndex;
        }
        return h->closestT < 1e29f;
    }

    // :240-267
    bool AnyHit_Sphere(const Ray& rayObj, int blasStart, int blasEnd, float tMaxObj) const {
        int cur = blasStart;
        while (cur != -1 && cur < blasEnd) {
            const BvhNode& n = sc->blasNodes[(size_t)cur];
            if (box(rayObj, n, 0.001f, tMaxObj)) {
                if (n.count > 0) {
                    int end = n.first + n.count;
                    for (int i = n.first; i < end; i++) {
                        int prim = sc->spherePrimIdx[(size_t)i];
                        float t; Float3 _n;
                        CNT(spheres);
                        if (IntersectSphere(rayObj, sc->spheres[(size_t)prim], &t, &_n)) {
                            if (t > 0.001f && t < tMaxObj) return true;
                        }
                    }
                    cur = n.skipIndex;
                } else cur = n.left;
            } else cur = n.skipIndex;
        }
        return false;
    }
    // :270-327
    bool AnyHit_Tri_Textured(const Ray& rayObj, int blasStart, int blasEnd, float tMaxObj) const {
        int cur = blasStart;
        while (cur != -1 && cur < blasEnd) {
            const BvhNode& n = sc->blasNodes[(size_t)cur];
            if (box(rayObj, n, 0.001f, tMaxObj)) {
                if (n.count > 0) {
                    int end = n.first + n.count;
                    for (int i = n.first; i < end; i++) {
                        int triIndex = sc->triPrimIdx[(size_t)i];
                        MeshTri tri = sc->meshTris[(size_t)triIndex];
                        Float3 v0 = sc->meshPositions[(size_t)tri.i0], v1 = sc->meshPositions[(size_t)tri.i1], v2 = sc->meshPositions[(size_t)tri.i2];
                        float t; Float3 nn; float bu, bv;
                        CNT(tris);
                        if (IntersectTriangleMT_Bary(rayObj, v0, v1, v2, &t, &nn, &bu, &bv)) {
                            if (t <= 0.001f || t >= tMaxObj) continue;
                            int midx = sc->triMatIndex[(size_t)triIndex];
                            const MaterialRecord& mat = sc->materials[(size_t)midx];
                            if (mat.HasAlphaMap != 0 && mat.AlphaTexIndex >= 0 && mat.AlphaTexIndex < sc->texInfosLength()) {
                                MeshTriUV tuv = sc->meshTriUVs[(size_t)triIndex];
                                Float2 t0 = sc->meshTexcoords[(size_t)tuv.t0], t1 = sc->meshTexcoords[(size_t)tuv.t1], t2 = sc->meshTexcoords[(size_t)tuv.t2];
                                float w = 1.0f - bu - bv;
                                float uu = t0.X * w + t1.X * bu + t2.X * bv;
                                float vv = t0.Y * w + t1.Y * bu + t2.Y * bv;
                                float aPoint = SampleMaskPoint(sc->texInfoAt(mat.AlphaTexIndex), uu, vv);
                                float cutoff = mat.AlphaCutoff;
                                const float Band = 0.10f;
                                if (aPoint < cutoff - Band) { continue; }
                                if (aPoint >= cutoff + Band) { return true; }
                                float aLin = SampleMaskLinear(sc->texInfoAt(mat.AlphaTexIndex), uu, vv);
                                if (aLin < cutoff) { continue; }
                            }
                            return true;
                        }
                    }
                    cur = n.skipIndex;
                } else cur = n.left;
            } else cur = n.skipIndex;
        }
        return false;
    }
    // :89-121
    bool ShadowOcclusion(const Ray& srayWorld, float tMaxWorld) const {
        int cur = sc->tlasNodes.empty() ? -1 : 0;
        while (cur != -1) {
            const BvhNode& n = sc->tlasNodes[(size_t)cur];
            if (box(srayWorld, n, 0.001f, tMaxWorld)) {
                if (n.count > 0) {
                    int end = n.first + n.count;
                    for (int i = n.first; i < end; i++) {
                        int instIndex = sc->tlasInstanceIndices[(size_t)i];
                        const InstanceRecord& inst = sc->instances[(size_t)instIndex];
                        Ray srayObj = TransformRay(inst.worldToObject, srayWorld);
                        float scale = inst.uniformScale > 0.0f ? inst.uniformScale : 1.0f;
                        float tMaxObj = tMaxWorld * scale;
                        bool blocked = (inst.type == BLAS_SPHERESET)
                            ? AnyHit_Sphere(srayObj, inst.blasRoot, inst.blasRoot + inst.blasNodeCount, tMaxObj)
                            : AnyHit_Tri_Textured(srayObj, inst.blasRoot, inst.blasRoot + inst.blasNodeCount, tMaxObj);
                        if (blocked) return true;
                    }
                    cur = n.skipIndex;
                } else cur = n.left;
            } else cur = n.skipIndex;
        }
        return false;
    }
};

// ------------------------------------------------------------------ Engine/RTRay.cs
struct Reservoir { Float3 L, wi; float pdf, w, wSum; int m, lightId; };   // :171-179
static_assert(sizeof(Reservoir) == 44, "Reservoir");

struct GBuffer {   // GpuGBuffer :80-109 (host arrays)
    std::vector<Float3> worldPos, normalWS, baseColor; std::vector<int> matId, objId, hitMask;
    // parity extras
    std::vector<int> primId, instId; std::vector<float> primaryT;
};

struct IntegratorParams {   // :129-169
    int width, height, frame;
    Camera cam, prevCam;
    Views views;
    const GBuffer* gb;
    Float3 dirLightDir, dirLightRadiance, skyTintTop, skyTintBottom;
    const Reservoir* resPrev; Reservoir* resCur; long resLen;
    int enableTemporalReuse, enableSpatialReuse, rngLockNoise, spp;

    Float3 PrimaryRayDir(int index) const {   // :148-154
        int x = index % width, y = index / width;
        float u = ((float)x + 0.5f) / (float)XMax(1, width);
        float v = ((float)y + 0.5f) / (float)XMax(1, height);
        return GenerateRay(cam, u, v).dir;
    }
    Float3 ViewDirFromCam(Float3 posWS) const { return Normalize(posWS - cam.origin); }   // :156
    float DistanceFromCamera(Float3 posWS) const { Float3 d = posWS - cam.origin; return sqrtf(d.X * d.X + d.Y * d.Y + d.Z * d.Z); } // :158-162
    Float3 SkyWeighted(Float3 dir) const { float tbg = 0.5f * (dir.Y + 1.0f); return skyTintBottom * (1.0f - tbg) + skyTintTop * tbg; } // :164-168
};

static const float PI = 3.14159265358979323846f;       // :183
static const float INV_PI = 0.31830988618379067154f;   // :184
static const float EPS_N = 0.0025f;                    // :185
static const float EPS_MIN = 1e-6f;                    // :186

static inline int PackRGBA8(Float3 c) {   // :66-76
    auto ToByte = [](float x) { float cc = XMin(1.0f, XMax(0.0f, x)); return (int)(255.99f * cc); };
    int R = ToByte(c.X), G = ToByte(c.Y), B = ToByte(c.Z);
    return (int)((255u << 24) | ((uint32_t)R << 16) | ((uint32_t)G << 8) | (uint32_t)B);
}
static inline int FloatToI16(float x) { float cl = XMax(0.0f, XMin(65535.0f, x * 1000.0f)); return (int)cl & 0xFFFF; }   // :609-613
static inline float I16ToFloat(int v) { return (float)v / 1000.0f; }                                                      // :615
static inline Ray MakeRayWithNormalOffset(Float3 origin, Float3 n, Float3 dir, float epsN) {   // :552-558
    Float3 d = Normalize(dir);
    float s = Dot(n, d) >= 0.0f ? 1.0f : -1.0f;
    Float3 o = origin + n * (epsN * s);
    Ray r; r.origin = o; r.dir = d; r.invDir = InvDir(d); return r;
}
static inline Float3 Reflect(Float3 I, Float3 N) { return I - N * (2.0f * Dot(I, N)); }   // :561
static inline bool Refract(Float3 I, Float3 N, float etaI, float etaT, Float3* T) {        // :564-572
    float eta = etaI / etaT;
    float cosI = -Dot(I, N);
    float k = 1.0f - eta * eta * (1.0f - cosI * cosI);
    if (k < 0.0f) { *T = Float3(); return false; }
    *T = Normalize(I * eta + N * (eta * cosI - sqrtf(k)));
    return true;
}
static inline float SchlickFresnel(float cos, float etaI, float etaT) {   // :575-583
    float r0 = (etaI - etaT) / (etaI + etaT);
    r0 = r0 * r0;
    float oneMinusCos = 1.0f - cos;
    float oneMinusCos2 = oneMinusCos * oneMinusCos;
    float oneMinusCos5 = oneMinusCos2 * oneMinusCos2 * oneMinusCos;
    return r0 + (1.0f - r0) * oneMinusCos5;
}
static inline void OrthonormalBasis(Float3 n, Float3* t, Float3* b) {     // :601-606
    Float3 up = fabsf(n.Y) < 0.999f ? Float3(0, 1, 0) : Float3(1, 0, 0);
    *t = Normalize(Cross(up, n));
    *b = Cross(n, *t);
}
static inline Float3 SampleHemisphereCosine(Float3 n, RNG& rng) {         // :586-598
    float r1 = rng.NextFloat(), r2 = rng.NextFloat();
    float phi = 2.0f * PI * r1;
    float cosTheta = sqrtf(1.0f - r2);
    float sinTheta = sqrtf(r2);
    float sphi, cphi; orc_sincos(phi, &sphi, &cphi);
    float x = cphi * sinTheta;
    float y = sphi * sinTheta;
    float z = cosTheta;
    Float3 t, b; OrthonormalBasis(n, &t, &b);
    Float3 v = t * x + b * y + n * z;
    return Normalize(v);
}
static inline float Luminance(Float3 c) { return 0.2126f * c.X + 0.7152f * c.Y + 0.0722f * c.Z; }   // :627
static inline float CosHemispherePdf(Float3 n, Float3 wi) { float nl = XMax(0.0f, Dot(n, wi)); return nl * INV_PI; } // :630-634
static inline uint32_t Hash(uint32_t x) { x ^= x >> 17; x *= 0xed5ad4bbu; x ^= x >> 11; x *= 0xac4c1b51u; x ^= x >> 15; x *= 0x31848babu; x ^= x >> 14; return x; } // :637-641
static inline uint32_t Hash3(uint32_t a, uint32_t b, uint32_t c) { return Hash(a ^ Hash(b ^ Hash(c))); }   // :643
static inline Float3 SafeColor(Float3 c) {   // :646-655
    float x = std::isfinite((double)c.X) ? c.X : 0.0f;
    float y = std::isfinite((double)c.Y) ? c.Y : 0.0f;
    float z = std::isfinite((double)c.Z) ? c.Z : 0.0f;
    x = XMin(1e6f, XMax(-1e6f, x)); y = XMin(1e6f, XMax(-1e6f, y)); z = XMin(1e6f, XMax(-1e6f, z));
    return Float3(x, y, z);
}

// per-path parity trace (not in the reference; observation only)
struct PathAov { int segCount; int term; uint32_t hash; };
static inline void fold(uint32_t& h, uint32_t v) { h = (h ^ v) * 16777619u; }

static bool Visible(const IntegratorParams& k, Float3 origin, Float3 n, Float3 wi, PathAov* aov) {   // :618-624
    float nl = Dot(n, wi);
    if (nl <= 0.0f) return false;
    Ray s = MakeRayWithNormalOffset(origin, n, wi, EPS_N);
    CNT(raysShadow);
    bool vis = !k.views.ShadowOcclusion(s, 1e29f);
    if (aov) fold(aov->hash, 0x100u | (vis ? 1u : 0u));
    return vis;
}
static inline Reservoir NewReservoirDefault() { Reservoir r; memset(&r, 0, sizeof(r)); return r; }   // :330-335
static int ReprojectToPrevPixel(Float3 posWS, const IntegratorParams& k) {   // :339-360
    Float3 p = posWS - k.prevCam.origin;
    float x = Dot(p, k.prevCam.right), y = Dot(p, k.prevCam.up), z = Dot(p, k.prevCam.forward);
    if (z <= 1e-4f) return -1;
    float tanHalfFov = orc_tan(0.5f * k.prevCam.fovYRadians);
    float ndcX = x / (z * tanHalfFov * k.prevCam.aspect);
    float ndcY = y / (z * tanHalfFov);
    float fx = 0.5f * (ndcX + 1.0f) * (float)k.width;
    float fy = 0.5f * (ndcY + 1.0f) * (float)k.height;
    int px = (int)fx, py = (int)fy;
    if ((uint32_t)px >= (uint32_t)k.width || (uint32_t)py >= (uint32_t)k.height) return -1;
    return py * k.width + px;
}
static bool SpatialCompatible(const IntegratorParams& k, int idxA, int idxB, Float3 nA) {   // :363-374
    int objA = k.gb->objId[(size_t)idxA], objB = k.gb->objId[(size_t)idxB];
    if (objA == objB) return true;
    Float3 nB = Normalize(k.gb->normalWS[(size_t)idxB]);
    float ndot = Dot(nA, nB);
    if (ndot < 0.85f) return false;
    float zA = k.DistanceFromCamera(k.gb->worldPos[(size_t)idxA]);
    float zB = k.DistanceFromCamera(k.gb->worldPos[(size_t)idxB]);
    float rel = fabsf(zA - zB) / XMax(1e-3f, zA);
    return rel < 0.05f;
}
static void Neighbor8(int rot, int radius, int* dx, int* dy) {   // :377-391
    int r = radius;
    auto RX = [](int x, int y, int R) { return R == 0 ? x : (R == 1 ? -y : (R == 2 ? -x : y)); };
    auto RY = [](int x, int y, int R) { return R == 0 ? y : (R == 1 ? x : (R == 2 ? -y : -x)); };
    const int sx[8] = {-r, r, 0, 0, -r, r, -r, r}, sy[8] = {0, 0, -r, r, -r, -r, r, r};
    for (int i = 0; i < 8; i++) { dx[i] = RX(sx[i], sy[i], rot); dy[i] = RY(sx[i], sy[i], rot); }
}
static void ReservoirUpdate(Reservoir& r, Float3 wi, float pdfSel, Float3 Li, float scoreS, int multiplicity, int lightId, RNG& rng) {   // :394-405
    float add = scoreS;
    float newSum = r.wSum + add;
    float acceptP = (newSum > 0.0f) ? add / newSum : 0.0f;
    if (rng.NextFloat() < acceptP) { r.wi = wi; r.pdf = pdfSel; r.L = Li; r.w = scoreS; r.lightId = lightId; }
    r.wSum = newSum;
    r.m = r.m + XMax(1, multiplicity);
}
static void ImportFromPrevReservoir(int prevIdx, int curIdx, const IntegratorParams& k, Float3 n, Float3 albedo, float mixLocal, float mixDelta, RNG& rng, Reservoir& r) {   // :408-435
    if (prevIdx < 0 || k.resLen <= prevIdx) return;
    if (!SpatialCompatible(k, curIdx, prevIdx, n)) return;
    Reservoir pr = k.resPrev[prevIdx];
    if (!(pr.m > 0 && pr.w > 0.0f && pr.wSum > 0.0f)) return;
    Float3 wi = pr.wi;
    int lid = pr.lightId == 2 ? 2 : 1;
    Float3 LiImp = (lid == 2) ? k.dirLightRadiance : k.SkyWeighted(wi);
    float nl = XMax(0.0f, Dot(n, wi));
    float pdfHere = (lid == 2) ? XMax(EPS_MIN, mixDelta) : XMax(EPS_MIN, CosHemispherePdf(n, wi) * mixLocal);
    Float3 f_over_p = albedo * LiImp * ((nl / pdfHere) * INV_PI);
    float sHere = Luminance(f_over_p);
    float Wsrc = pr.wSum / ((float)XMax(1, pr.m) * XMax(EPS_MIN, pr.w));
    float eff = sHere * Wsrc;
    ReservoirUpdate(r, wi, pdfHere, LiImp, eff, 1, lid, rng);
}
static Float3 ReSTIR_Direct(int index, const IntegratorParams& k, Float3 pos, Float3 n, Float3 albedo, RNG& rng, Reservoir* outRes, PathAov* aov) {   // :438-543
    const int LocalCandidates = 8, DeltaCandidates = 1, TotalNew = LocalCandidates + DeltaCandidates;
    float mixLocal = (float)LocalCandidates / (float)TotalNew;
    float mixDelta = (float)DeltaCandidates / (float)TotalNew;
    Reservoir r = NewReservoirDefault();
    for (int i = 0; i < LocalCandidates; i++) {   // (1)
        Float3 wi = SampleHemisphereCosine(n, rng);
        float nl = XMax(0.0f, Dot(n, wi));
        float pdfLocal = XMax(EPS_MIN, CosHemispherePdf(n, wi));
        float pdfSel = XMax(EPS_MIN, pdfLocal * mixLocal);
        Float3 LiLoc = k.SkyWeighted(wi);
        Float3 f_over_p = albedo * LiLoc * ((nl / pdfSel) * INV_PI);
        float s = Luminance(f_over_p);
        ReservoirUpdate(r, wi, pdfSel, LiLoc, s, 1, 1, rng);
    }
    {   // (2)
        Float3 wi = Normalize(k.dirLightDir);
        float nl = XMax(0.0f, Dot(n, wi));
        float pdfSel = XMax(EPS_MIN, mixDelta);
        Float3 LiDir = k.dirLightRadiance;
        Float3 f_over_p = albedo * LiDir * ((nl / pdfSel) * INV_PI);
        float s = Luminance(f_over_p);
        ReservoirUpdate(r, wi, pdfSel, LiDir, s, 1, 2, rng);
    }
    if (k.enableTemporalReuse != 0) {   // (3)
        int prevIdx = ReprojectToPrevPixel(pos, k);
        if (prevIdx >= 0) ImportFromPrevReservoir(prevIdx, index, k, n, albedo, mixLocal, mixDelta, rng, r);
    }
    if (k.enableSpatialReuse != 0) {    // (4)
        uint32_t h = Hash3((uint32_t)index, (uint32_t)k.frame, 0xB31F5AB1u);
        int rot = (int)(h & 3u);
        int radius = 1 + (int)((h >> 2) & 1u);
        int x0 = index % k.width, y0 = index / k.width;
        int dx[8], dy[8]; Neighbor8(rot, radius, dx, dy);
        for (int i = 0; i < 8; i++) {
            int ni = ((uint32_t)(x0 + dx[i]) < (uint32_t)k.width && (uint32_t)(y0 + dy[i]) < (uint32_t)k.height) ? (y0 + dy[i]) * k.width + (x0 + dx[i]) : -1;
            ImportFromPrevReservoir(ni, index, k, n, albedo, mixLocal, mixDelta, rng, r);
        }
    }
    Float3 contrib(0, 0, 0);   // (5)
    if (r.m > 0 && r.wSum > 0.0f && r.w > 0.0f) {
        Float3 wiSel = r.wi;
        int lidSel = r.lightId == 2 ? 2 : 1;
        float nlSel = XMax(0.0f, Dot(n, wiSel));
        if (nlSel > 0.0f && Visible(k, pos, n, wiSel, aov)) {
            float mixLocal2 = (float)LocalCandidates / (float)(LocalCandidates + DeltaCandidates);
            float mixDelta2 = (float)DeltaCandidates / (float)(LocalCandidates + DeltaCandidates);
            float pdfSel = (lidSel == 2) ? XMax(EPS_MIN, mixDelta2) : XMax(EPS_MIN, CosHemispherePdf(n, wiSel) * mixLocal2);
            Float3 LiSel = (lidSel == 2) ? k.dirLightRadiance : k.SkyWeighted(wiSel);
            Float3 f_over_p = albedo * LiSel * ((nlSel / pdfSel) * INV_PI);
            float W = r.wSum / (float)XMax(1, r.m) / XMax(EPS_MIN, r.w);
            contrib = f_over_p * W;
        }
    }
    *outRes = r;
    return contrib;
}
static bool TraceNext(const IntegratorParams& k, const Ray& ray, Float3& pos, Float3& nrm, Float3& alb, int& shade, float& ior, PathAov* aov) {   // :659-671
    HitInfo h;
    CNT(raysBounce);
    bool hit2 = k.views.TraceClosest(ray, &h);
    if (aov) { aov->segCount++; if (hit2) { fold(aov->hash, (uint32_t)h.instIdx); fold(aov->hash, (uint32_t)h.primIdx); } else fold(aov->hash, 0xFFFFFFFFu); }
    if (!hit2) return false;
    pos = ray.origin + ray.dir * h.closestT;
    nrm = Normalize(h.bestNormal);
    alb = h.bestAlbedo; shade = h.bestShade; ior = h.bestIor;
    return true;
}

// PrimaryVisibilityKernel :188-201
static void PrimaryVisibilityKernel(int index, int width, int height, const Camera& cam, const Views& views, GBuffer& gb) {
    int x = index % width, y = index / width;   // GBufferParams.PrimaryRay :120-126
    float u = ((float)x + 0.5f) / (float)XMax(1, width);
    float v = ((float)y + 0.5f) / (float)XMax(1, height);
    Ray wray = GenerateRay(cam, u, v);
    HitInfo h;
    CNT(raysPrimary);
    bool hit = views.TraceClosest(wray, &h);
    gb.primId[(size_t)index] = h.primIdx; gb.instId[(size_t)index] = h.instIdx; gb.primaryT[(size_t)index] = h.closestT;
    if (!hit) {   // StoreMiss :100-108
        gb.hitMask[(size_t)index] = 0;
        gb.worldPos[(size_t)index] = wray.origin + wray.dir * 1e6f;
        gb.normalWS[(size_t)index] = Float3(0, 1, 0);
        gb.baseColor[(size_t)index] = Float3(0, 0, 0);
        gb.matId[(size_t)index] = -1; gb.objId[(size_t)index] = -1;
        return;
    }
    Float3 posWS = wray.origin + wray.dir * h.closestT;
    int packedMat = (h.bestShade & 0xFFFF) | (FloatToI16(h.bestIor) << 16);
    gb.hitMask[(size_t)index] = 1;   // StoreHit :90-98
    gb.worldPos[(size_t)index] = posWS; gb.normalWS[(size_t)index] = h.bestNormal; gb.baseColor[(size_t)index] = h.bestAlbedo;
    gb.matId[(size_t)index] = packedMat; gb.objId[(size_t)index] = h.bestObjId;
}

struct FbOut { int* color; float* depth; int* objectId; float* radiance; uint8_t* segCount; uint8_t* termCode; uint32_t* pathHash; long aovStride; long outIndex; };

// PathTraceKernel :203-325
static void PathTraceKernel(int index, const IntegratorParams& k, int MaxDepth, const FbOut& fb) {
    Float3 Lframe;
    for (int s = 0; s < XMax(1, k.spp); s++) {
        RNG rng = RNG::CreateFromIndex1D(index, k.width, k.height, k.frame, (uint32_t)s, 0xC0FFEEu, k.rngLockNoise);
        PathAov aov; aov.segCount = 0; aov.term = 2; aov.hash = 0x811C9DC5u;
        PathAov* pa = (fb.segCount || fb.termCode || fb.pathHash) ? &aov : nullptr;
        if (k.gb->hitMask[(size_t)index] == 0) {
            Float3 vdir = k.PrimaryRayDir(index);
            Lframe = Lframe + SafeColor(k.SkyWeighted(vdir));
            aov.term = 0;
        } else {
            Float3 pos = k.gb->worldPos[(size_t)index];
            Float3 nrm = Normalize(k.gb->normalWS[(size_t)index]);
            Float3 alb = k.gb->baseColor[(size_t)index];
            int packedMat = k.gb->matId[(size_t)index];
            int shade = packedMat & 0xFFFF;
            float ior = I16ToFloat((packedMat >> 16) & 0xFFFF);
            Float3 Li(0, 0, 0), throughput(1, 1, 1);
            Float3 I = k.ViewDirFromCam(pos);
            bool wroteReservoir = false;
            for (int depth = 0; depth < MaxDepth; depth++) {
                if (shade == SHADING_MIRROR) {   // :235-244
                    Float3 dirR = Reflect(I, nrm);
                    Ray ray = MakeRayWithNormalOffset(pos, nrm, dirR, EPS_N);
                    throughput = throughput * alb;
                    if (!TraceNext(k, ray, pos, nrm, alb, shade, ior, pa)) { Li = Li + throughput * k.SkyWeighted(ray.dir); aov.term = 1; break; }
                    I = ray.dir; continue;
                }
                if (shade == SHADING_GLASS) {    // :246-275
                    Float3 Nuse = nrm;
                    bool outside = Dot(I, nrm) < 0.0f;
                    if (!outside) Nuse = Nuse * -1.0f;
                    float etaI = outside ? 1.0f : (ior > 0.0f ? ior : 1.5f);
                    float etaT = outside ? (ior > 0.0f ? ior : 1.5f) : 1.0f;
                    Float3 dirR = Reflect(I, Nuse);
                    Float3 dirT;
                    bool refrOk = Refract(I, Nuse, etaI, etaT, &dirT);
                    float cosI = fabsf(Dot(I, Nuse));
                    float Fr = SchlickFresnel(cosI, etaI, etaT);
                    float xi = rng.NextFloat();
                    Ray ray = (!refrOk || xi < Fr) ? MakeRayWithNormalOffset(pos, Nuse, dirR, EPS_N) : MakeRayWithNormalOffset(pos, -Nuse, dirT, EPS_N);
                    if (refrOk && xi >= Fr) {
                        Float3 transTint = (alb.X == 0.0f && alb.Y == 0.0f && alb.Z == 0.0f) ? Float3(1, 1, 1) : alb;
                        float etaScale = (etaI * etaI) / (etaT * etaT);
                        throughput = throughput * transTint * etaScale;
                    }
                    if (!TraceNext(k, ray, pos, nrm, alb, shade, ior, pa)) { Li = Li + throughput * k.SkyWeighted(ray.dir); aov.term = 1; break; }
                    I = ray.dir; continue;
                }
                {   // :278-298
                    Reservoir outRes;
                    if (wroteReservoir) {
                        IntegratorParams kLocal = k;
                        kLocal.enableTemporalReuse = 0; kLocal.enableSpatialReuse = 0;
                        Float3 direct = ReSTIR_Direct(index, kLocal, pos, nrm, alb, rng, &outRes, pa);
                        Li = Li + throughput * direct;
                    } else {
                        Float3 direct = ReSTIR_Direct(index, k, pos, nrm, alb, rng, &outRes, pa);
                        Li = Li + throughput * direct;
                        if (k.resLen > index) { if (k.resCur) k.resCur[index] = outRes; wroteReservoir = true; }
                    }
                }
                {   // :301-317
                    Float3 wi = SampleHemisphereCosine(nrm, rng);
                    Ray ray = MakeRayWithNormalOffset(pos, nrm, wi, EPS_N);
                    throughput = throughput * alb;
                    if (depth >= 3) {
                        float maxC = XMax(throughput.X, XMax(throughput.Y, throughput.Z));
                        maxC = XClamp(maxC, 0.05f, 0.98f);
                        if (rng.NextFloat() > maxC) { throughput = Float3(0, 0, 0); aov.term = 3; break; }
                        throughput = throughput * (1.0f / maxC);
                    }
                    if (!TraceNext(k, ray, pos, nrm, alb, shade, ior, pa)) { Li = Li + throughput * k.SkyWeighted(ray.dir); aov.term = 1; break; }
                    I = ray.dir; continue;
                }
            }
            Lframe = Lframe + SafeColor(Li);
        }
        if (fb.segCount) fb.segCount[(size_t)s * fb.aovStride + fb.outIndex] = (uint8_t)aov.segCount;
        if (fb.termCode) fb.termCode[(size_t)s * fb.aovStride + fb.outIndex] = (uint8_t)aov.term;
        if (fb.pathHash) fb.pathHash[(size_t)s * fb.aovStride + fb.outIndex] = aov.hash;
    }
    Float3 Lout = Lframe * (1.0f / (float)XMax(1, k.spp));
    if (fb.color) fb.color[fb.outIndex] = PackRGBA8(Lout);                                                     // GpuFramebuffer.Store :59-64
    if (fb.depth) fb.depth[fb.outIndex] = k.DistanceFromCamera(k.gb->worldPos[(size_t)index]);
    if (fb.objectId) fb.objectId[fb.outIndex] = k.gb->objId[(size_t)index];
    if (fb.radiance) { fb.radiance[3 * fb.outIndex] = Lout.X; fb.radiance[3 * fb.outIndex + 1] = Lout.Y; fb.radiance[3 * fb.outIndex + 2] = Lout.Z; }
}

// Camera.CreateCamera / Translate / UpdateDerived  Engine/Camera.cs:19-47,121-126,184-191
static void UpdateDerived(Camera& c, float aspectIn, float fovYRadIn) {
    c.forward = Normalize((c.lowerLeft + c.horizontal * 0.5f + c.vertical * 0.5f) - c.origin);
    c.up = Normalize(c.vertical);
    c.right = Normalize(Cross(c.forward, c.up));
    c.aspect = aspectIn; c.fovYRadians = fovYRadIn;
}
static Camera CreateCameraLookAt(int width, int height, float fovDegrees, Float3 origin, Float3 lookAt) {
    float aspect = (float)width / (float)XMax(1, height);
    float theta = fovDegrees * (3.14159274f / 180.0f);   // XMath.PI is the float constant
    float halfHeight = tanf(0.5f * theta);              // host code: XMath.Tan on the CPU = MathF.Tan
    float halfWidth = aspect * halfHeight;
    Float3 upHint(0, 1, 0);
    Float3 w = Normalize(origin - lookAt);
    Float3 u = Normalize(Cross(upHint, w));
    Float3 v = Cross(w, u);
    Camera cam; memset(&cam, 0, sizeof(cam));
    cam.origin = origin;
    cam.lowerLeft = origin - u * halfWidth - v * halfHeight - w;
    cam.horizontal = u * (2.0f * halfWidth);
    cam.vertical = v * (2.0f * halfHeight);
    UpdateDerived(cam, aspect, theta);
    return cam;
}
static void BakeCameraDerived(Camera& c, int pixelW, int pixelH) {   // Engine/RTRenderer.cs:241-263
    Float3 center = c.lowerLeft + c.horizontal * 0.5f + c.vertical * 0.5f;
    Float3 forward = Normalize(center - c.origin);
    Float3 up = Normalize(c.vertical);
    Float3 right = Normalize(Cross(forward, up));
    float focusDist = Length(center - c.origin);
    float halfHeight = 0.5f * Length(c.vertical);
    float tanHalfFov = (focusDist > 1e-6f) ? (halfHeight / focusDist) : halfHeight;
    float fovY = 2.0f * atanf(tanHalfFov);
    float aspect = (Length(c.horizontal) > 1e-6f && Length(c.vertical) > 1e-6f) ? (Length(c.horizontal) / Length(c.vertical)) : ((float)pixelW / (float)XMax(1, pixelH));
    c.forward = forward; c.up = up; c.right = right; c.fovYRadians = fovY; c.aspect = aspect;
}

}   // namespace orc

// =====================================================================================
// C API (ctypes)
// =====================================================================================
using namespace orc;

struct OrcConfig {
    int32_t width, height, frame, spp, maxDepth, rngLockNoise, enableTemporalReuse, enableSpatialReuse;
    Float3 dirLightDir, dirLightRadiance, skyTintTop, skyTintBottom;
    uint32_t flags;            // bit0 = RT_FLAG_TRI_MATERIALS
    int32_t x0, y0, x1, y1;    // crop window [x0,x1) x [y0,y1); outputs are crop-sized, row-major from (x0,y0)
    int32_t threads;           // 0 = hardware_concurrency
    int32_t noCull;            // debugging: culling-free traversal (slow)
};
struct OrcOutputs {
    int32_t* rgba8; float* depth; int32_t* objId; float* radiance;   // radiance: 3 floats per px
    int32_t* primId; int32_t* instId; float* primaryT; int32_t* hitMask;
    float* gbPos; float* gbNrm; float* gbAlb; int32_t* gbMat;
    uint8_t* segCount; uint8_t* termCode; uint32_t* pathHash;        // [max(1,spp)][npx]
    Reservoir* resPrev; Reservoir* resCur;                           // full-frame only (crop must be the full image)
    uint64_t counters[8];   // raysPrimary, raysBounce, raysShadow, nodes, tris, spheres, -, -
    double seconds[2];      // primary pass, integrator pass (wall clock)
};

ORC_API Scene* orc_scene_new() { return new Scene(); }
ORC_API void orc_scene_free(Scene* s) { delete s; }
ORC_API void orc_scene_clear(Scene* s) { s->Clear(); }
ORC_API void orc_scene_build_default(Scene* s) { s->BuildDefaultScene(); }
ORC_API int orc_scene_add_texture(Scene* s, int w, int h, const RGBA32* px) { return s->AddTexture(w, h, px); }
ORC_API int orc_scene_add_sphere(Scene* s, const Sphere* sp) { return s->AddSphere(*sp); }
ORC_API void orc_scene_add_sphere_instance(Scene* s, const int* ids, int n, const Affine3x4* o2w) { s->instances.push_back(s->BuildSphereInstance(ids, n, *o2w)); }
ORC_API void orc_scene_add_mesh_instance(Scene* s, const Float3* pos, int nPos, const MeshTri* tris, int nTris, const Float2* uv, int nUV,
                                         const MeshTriUV* triUVs, const int* triMat, const MaterialRecord* mats, int nMats, const Affine3x4* o2w) {
    s->AddMeshInstance(pos, nPos, tris, nTris, uv, nUV, triUVs, triMat, mats, nMats, *o2w);
}
ORC_API void orc_scene_rebuild_tlas(Scene* s) { s->RebuildTLAS(); }
ORC_API long orc_scene_sort_ties(Scene* s) { return s->sortTies; }
// which: 0..14 in SceneDeviceViews order; returns element count and the host pointer.
ORC_API int64_t orc_scene_array(Scene* s, int which, const void** ptr) {
    switch (which) {
        case 0: *ptr = s->tlasNodes.data(); return (int64_t)s->tlasNodes.size();
        case 1: *ptr = s->tlasInstanceIndices.data(); return (int64_t)s->tlasInstanceIndices.size();
        case 2: *ptr = s->instances.data(); return (int64_t)s->instances.size();
        case 3: *ptr = s->blasNodes.data(); return (int64_t)s->blasNodes.size();
        case 4: *ptr = s->spherePrimIdx.data(); return (int64_t)s->spherePrimIdx.size();
        case 5: *ptr = s->spheres.data(); return (int64_t)s->spheres.size();
        case 6: *ptr = s->triPrimIdx.data(); return (int64_t)s->triPrimIdx.size();
        case 7: *ptr = s->meshPositions.data(); return (int64_t)s->meshPositions.size();
        case 8: *ptr = s->meshTris.data(); return (int64_t)s->meshTris.size();
        case 9: *ptr = s->meshTexcoords.data(); return (int64_t)s->meshTexcoords.size();
        case 10: *ptr = s->meshTriUVs.data(); return (int64_t)s->meshTriUVs.size();
        case 11: *ptr = s->triMatIndex.data(); return (int64_t)s->triMatIndex.size();
        case 12: *ptr = s->materials.data(); return (int64_t)s->materials.size();
        case 13: *ptr = s->texels.data(); return (int64_t)s->texels.size();
        case 14: *ptr = s->texInfos.data(); return (int64_t)s->texInfos.size();
    }
    *ptr = nullptr; return -1;
}

ORC_API void orc_camera_create(int width, int height, float fovDegrees, const float* origin, const float* lookAt, Camera* out) {
    *out = CreateCameraLookAt(width, height, fovDegrees, Float3(origin[0], origin[1], origin[2]), Float3(lookAt[0], lookAt[1], lookAt[2]));
}
ORC_API void orc_camera_translate(Camera* c, float dx, float dy, float dz) {   // Engine/Camera.cs:121-126
    Float3 d(dx, dy, dz); c->origin = c->origin + d; c->lowerLeft = c->lowerLeft + d; UpdateDerived(*c, c->aspect, c->fovYRadians);
}
ORC_API void orc_camera_bake(Camera* c, int w, int h) { BakeCameraDerived(*c, w, h); }

// ---- present chain: RTRenderer.BlitKernel / BilinearUpsampleKernel (Engine/RTRenderer.cs:281-346), RTTaa.TaaResolveKernel (Engine/RTTaa.cs:117-262) ----
namespace orc {
static inline Float3 UnpackRGB(int rgba8) {   // RTRenderer.cs:322-328
    float r = (float)((rgba8 >> 16) & 255) * (1.0f / 255.0f), g = (float)((rgba8 >> 8) & 255) * (1.0f / 255.0f), b = (float)(rgba8 & 255) * (1.0f / 255.0f);
    return Float3(r, g, b);
}
static inline int ClampI(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static void BilinearUpsampleKernel(int index, const int* srcRGBA8, int srcW, int srcH, int* dstRGBA8, int dstW, int dstH) {   // RTRenderer.cs:287-320
    int x = index % dstW, y = index / dstW;
    float u = (((float)x + 0.5f) * (float)srcW / (float)dstW) - 0.5f;
    float v = (((float)y + 0.5f) * (float)srcH / (float)dstH) - 0.5f;
    int x0 = ClampI((int)floorf(u), 0, srcW - 1), y0 = ClampI((int)floorf(v), 0, srcH - 1);
    int x1 = ClampI(x0 + 1, 0, srcW - 1), y1 = ClampI(y0 + 1, 0, srcH - 1);
    float tx = XClamp(u - (float)x0, 0.0f, 1.0f), ty = XClamp(v - (float)y0, 0.0f, 1.0f);
    Float3 c00 = UnpackRGB(srcRGBA8[y0 * srcW + x0]), c10 = UnpackRGB(srcRGBA8[y0 * srcW + x1]);
    Float3 c01 = UnpackRGB(srcRGBA8[y1 * srcW + x0]), c11 = UnpackRGB(srcRGBA8[y1 * srcW + x1]);
    Float3 cx0 = c00 * (1.0f - tx) + c10 * tx;
    Float3 cx1 = c01 * (1.0f - tx) + c11 * tx;
    Float3 c = cx0 * (1.0f - ty) + cx1 * ty;
    dstRGBA8[index] = PackRGBA8(c);
}
static inline Float3 UnpackSRGB(int rgba) {   // RTTaa.cs:236-246
    float r = (float)((rgba >> 16) & 255) / 255.0f, g = (float)((rgba >> 8) & 255) / 255.0f, b = (float)(rgba & 255) / 255.0f;
    r = (r <= 0.04045f) ? (r / 12.92f) : orc_pow((r + 0.055f) / 1.055f, 2.4f);
    g = (g <= 0.04045f) ? (g / 12.92f) : orc_pow((g + 0.055f) / 1.055f, 2.4f);
    b = (b <= 0.04045f) ? (b / 12.92f) : orc_pow((b + 0.055f) / 1.055f, 2.4f);
    return Float3(r, g, b);
}
static inline int PackSRGB(Float3 c) {   // RTTaa.cs:248-262
    float rL = XMax(0.0f, XMin(1.0f, c.X)), gL = XMax(0.0f, XMin(1.0f, c.Y)), bL = XMax(0.0f, XMin(1.0f, c.Z));
    float r = (rL <= 0.0031308f) ? 12.92f * rL : 1.055f * orc_pow(rL, 1.0f / 2.4f) - 0.055f;
    float g = (gL <= 0.0031308f) ? 12.92f * gL : 1.055f * orc_pow(gL, 1.0f / 2.4f) - 0.055f;
    float b = (bL <= 0.0031308f) ? 12.92f * bL : 1.055f * orc_pow(bL, 1.0f / 2.4f) - 0.055f;
    int R = (int)rintf(XMax(0.0f, XMin(1.0f, r)) * 255.0f);   // XMath.Round: half to even
    int G = (int)rintf(XMax(0.0f, XMin(1.0f, g)) * 255.0f);
    int B = (int)rintf(XMax(0.0f, XMin(1.0f, b)) * 255.0f);
    return (int)((255u << 24) | ((uint32_t)R << 16) | ((uint32_t)G << 8) | (uint32_t)B);
}
static inline Float3 CatRom(Float3 a, Float3 b, float t) { float tt = t * (2.0f - t); return a * (1.0f - tt) + b * tt; }   // :228-233
static Float3 SampleCatRomSRGB(const int* a, int w, int h, float x, float y) {   // :209-226
    int x1 = ClampI((int)floorf(x), 0, w - 1), y1 = ClampI((int)floorf(y), 0, h - 1);
    float fx = x - (float)x1, fy = y - (float)y1;
    Float3 c00 = UnpackSRGB(a[y1 * w + x1]);
    Float3 c10 = UnpackSRGB(a[y1 * w + XMin(x1 + 1, w - 1)]);
    Float3 c01 = UnpackSRGB(a[XMin(y1 + 1, h - 1) * w + x1]);
    Float3 c11 = UnpackSRGB(a[XMin(y1 + 1, h - 1) * w + XMin(x1 + 1, w - 1)]);
    Float3 cx0 = CatRom(c00, c10, fx), cx1 = CatRom(c01, c11, fx);
    return CatRom(cx0, cx1, fy);
}
struct TaaParams { int* outColor; const int* inColorLow; const int* inObjIdLow; int* historyColor; int* historyObjId; int outW, outH, inW, inH; float feedback, sharpness, clampK; int isFirstFrame; };
static void TaaResolveKernel(int idx, const TaaParams& p) {   // :117-179
    int outW = p.outW;
    int px = idx % outW, py = idx / outW;
    float sx = ((float)px + 0.5f) * ((float)p.inW / (float)outW) - 0.5f;
    float sy = ((float)py + 0.5f) * ((float)p.inH / (float)p.outH) - 0.5f;
    Float3 cur = SampleCatRomSRGB(p.inColorLow, p.inW, p.inH, sx, sy);
    Float3 nmin = cur, nmax = cur;
    for (int oy = -1; oy <= 1; oy++) for (int ox = -1; ox <= 1; ox++) {
        if (ox == 0 && oy == 0) continue;
        Float3 c = SampleCatRomSRGB(p.inColorLow, p.inW, p.inH, sx + (float)ox * 0.5f, sy + (float)oy * 0.5f);
        nmin = Float3(XMin(nmin.X, c.X), XMin(nmin.Y, c.Y), XMin(nmin.Z, c.Z));
        nmax = Float3(XMax(nmax.X, c.X), XMax(nmax.Y, c.Y), XMax(nmax.Z, c.Z));
    }
    int ix = ClampI((int)rintf(sx), 0, p.inW - 1), iy = ClampI((int)rintf(sy), 0, p.inH - 1);   // SampleNearestObj :200-205
    int objId = p.inObjIdLow[iy * p.inW + ix];
    Float3 hist = UnpackSRGB(p.historyColor[idx]);
    int histObj = p.historyObjId[idx];
    bool reset = (p.isFirstFrame != 0) || (histObj != objId);
    float k = p.clampK;
    Float3 cmin(nmin.X - k * 0.0f, nmin.Y - k * 0.0f, nmin.Z - k * 0.0f), cmax(nmax.X + k * 0.0f, nmax.Y + k * 0.0f, nmax.Z + k * 0.0f);   // Clamp :191-198
    Float3 histClamped(XMin(cmax.X, XMax(cmin.X, hist.X)), XMin(cmax.Y, XMax(cmin.Y, hist.Y)), XMin(cmax.Z, XMax(cmin.Z, hist.Z)));
    float a = reset ? 1.0f : p.feedback;
    Float3 accum = histClamped * (1.0f - a) + cur * a;                                                         // Lerp
    Float3 sharpen = accum * (1.0f + 2.0f * p.sharpness) - (nmin + nmax) * (0.5f * p.sharpness);
    accum = accum * (1.0f - p.sharpness) + sharpen * p.sharpness;                                             // Mix
    p.outColor[idx] = PackSRGB(accum);
    p.historyColor[idx] = p.outColor[idx];
    p.historyObjId[idx] = objId;
}
}   // namespace orc
ORC_API float orc_math_pow(float x, float y) { return orc_pow(x, y); }
// dst may not alias src.  The blit path (RTRenderer.cs:225-226, equal sizes) is a plain copy and needs no oracle.
ORC_API void orc_bilinear_upsample(const int* src, int srcW, int srcH, int* dst, int dstW, int dstH) {
    for (int i = 0; i < dstW * dstH; i++) BilinearUpsampleKernel(i, src, srcW, srcH, dst, dstW, dstH);
}
// RTTaa.ResolveUpsample (RTTaa.cs:49-96) with its fixed tunables passed in; history arrays are updated in place.
ORC_API void orc_taa_resolve(int* outColor, const int* lowColor, const int* lowObjId, int inW, int inH, int outW, int outH,
                             int* historyColor, int* historyObjId, int isFirstFrame, float feedback, float sharpness, float clampK) {
    TaaParams p; p.outColor = outColor; p.inColorLow = lowColor; p.inObjIdLow = lowObjId; p.historyColor = historyColor; p.historyObjId = historyObjId;
    p.outW = outW; p.outH = outH; p.inW = inW; p.inH = inH; p.feedback = feedback; p.sharpness = sharpness; p.clampK = clampK; p.isFirstFrame = isFirstFrame;
    for (int i = 0; i < outW * outH; i++) TaaResolveKernel(i, p);
}

// KAT taps
ORC_API uint32_t orc_rng_seed(int px, int py, int frame, uint32_t sample, uint32_t salt, int lockNoise) { return RNG::CreateFromPixel(px, py, frame, sample, salt, lockNoise).state; }
ORC_API void orc_rng_stream(uint32_t seed, int n, uint32_t* outU, float* outF) {
    RNG r = RNG::Create(seed);
    for (int i = 0; i < n; i++) { RNG c = r; uint32_t u = r.NextUInt(); outU[i] = u; outF[i] = c.NextFloat(); }
}
ORC_API int orc_pack_rgba8(float r, float g, float b) { return PackRGBA8(Float3(r, g, b)); }
ORC_API int orc_intersect_triangle(const float* o, const float* d, const float* v0, const float* v1, const float* v2, float* out4) {
    Ray r; r.origin = Float3(o[0], o[1], o[2]); r.dir = Float3(d[0], d[1], d[2]); r.invDir = InvDir(r.dir);
    float t, bu, bv; Float3 n;
    bool h = IntersectTriangleMT_Bary(r, Float3(v0[0], v0[1], v0[2]), Float3(v1[0], v1[1], v1[2]), Float3(v2[0], v2[1], v2[2]), &t, &n, &bu, &bv);
    out4[0] = t; out4[1] = bu; out4[2] = bv; out4[3] = n.Y; return h ? 1 : 0;
}
ORC_API int orc_intersect_sphere(const float* o, const float* d, const float* c, float radius, float* out4) {
    Ray r; r.origin = Float3(o[0], o[1], o[2]); r.dir = Float3(d[0], d[1], d[2]); r.invDir = InvDir(r.dir);
    Sphere s; memset(&s, 0, sizeof(s)); s.center = Float3(c[0], c[1], c[2]); s.radius = radius;
    float t; Float3 n; bool h = IntersectSphere(r, s, &t, &n);
    out4[0] = t; out4[1] = n.X; out4[2] = n.Y; out4[3] = n.Z; return h ? 1 : 0;
}
ORC_API int orc_intersect_aabb(const float* o, const float* d, const float* bmin, const float* bmax, float tMin, float tMax) {
    Ray r; r.origin = Float3(o[0], o[1], o[2]); r.dir = Float3(d[0], d[1], d[2]); r.invDir = InvDir(r.dir);
    return IntersectAABB(r, Float3(bmin[0], bmin[1], bmin[2]), Float3(bmax[0], bmax[1], bmax[2]), tMin, tMax) ? 1 : 0;
}
ORC_API void orc_math_sincos(float x, float* s, float* c) { orc_sincos(x, s, c); }
ORC_API float orc_math_atan2(float y, float x) { return orc_atan2(y, x); }
ORC_API float orc_math_acos(float x) { return orc_acos(x); }
ORC_API void orc_sample_hemisphere(const float* n, uint32_t seed, float* out3) {
    RNG r = RNG::Create(seed); Float3 v = SampleHemisphereCosine(Float3(n[0], n[1], n[2]), r); out3[0] = v.X; out3[1] = v.Y; out3[2] = v.Z;
}
// Trace one world ray through TraceClosest (cull=1 reference traversal, cull=0 culling-free).
ORC_API int orc_trace_closest(Scene* s, const float* o, const float* d, int cull, unsigned flags, float* tOut, int* instOut, int* primOut) {
    bool savedNoCull = s->noCull, savedTM = s->triMaterials;
    s->noCull = !cull; s->triMaterials = (flags & 1u) != 0;
    Views v; v.sc = s;
    Ray r; r.origin = Float3(o[0], o[1], o[2]); r.dir = Float3(d[0], d[1], d[2]); r.invDir = InvDir(r.dir);
    HitInfo h; bool hit = v.TraceClosest(r, &h);
    s->noCull = savedNoCull; s->triMaterials = savedTM;
    *tOut = h.closestT; *instOut = h.instIdx; *primOut = h.primIdx;
    return hit ? 1 : 0;
}
ORC_API int orc_hardware_threads() { return (int)std::max(1u, std::thread::hardware_concurrency()); }

// The two launches of RTRenderer.RenderDirectToPbo (Engine/RTRenderer.cs:152-153, 181-205) over a crop window.
ORC_API int orc_render(Scene* scene, const Camera* cam, const Camera* prevCam, const OrcConfig* cfg, OrcOutputs* out) {
    const int W = cfg->width, H = cfg->height;
    int x0 = cfg->x0, y0 = cfg->y0, x1 = cfg->x1, y1 = cfg->y1;
    if (x1 <= x0 || y1 <= y0) { x0 = 0; y0 = 0; x1 = W; y1 = H; }
    const int cw = x1 - x0, ch = y1 - y0;
    const bool reuse = cfg->enableTemporalReuse || cfg->enableSpatialReuse;
    if (reuse && (cw != W || ch != H)) return -1;
    scene->noCull = cfg->noCull != 0;
    scene->triMaterials = (cfg->flags & 1u) != 0;

    // GBuffer is indexed by the GLOBAL pixel index so the integrator restatement stays literal.
    // With reuse off only crop pixels are touched; allocate lazily as a full frame.
    GBuffer gb; size_t N = (size_t)W * H;
    gb.worldPos.resize(N); gb.normalWS.resize(N); gb.baseColor.resize(N); gb.matId.resize(N); gb.objId.resize(N); gb.hitMask.resize(N);
    gb.primId.resize(N); gb.instId.resize(N); gb.primaryT.resize(N);

    int T = cfg->threads > 0 ? cfg->threads : (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<Counters> cnts((size_t)T);
    Views views; views.sc = scene;

    auto runRows = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t]() { g_cnt = &cnts[(size_t)t]; for (int y = y0 + t; y < y1; y += T) fn(y); g_cnt = nullptr; });
        for (auto& x : th) x.join();
    };
    auto tA = std::chrono::steady_clock::now();
    runRows([&](int y) { for (int x = x0; x < x1; x++) PrimaryVisibilityKernel(y * W + x, W, H, *cam, views, gb); });
    auto tB = std::chrono::steady_clock::now();

    IntegratorParams k;
    k.width = W; k.height = H; k.frame = cfg->frame; k.cam = *cam; k.prevCam = prevCam ? *prevCam : *cam; k.views = views; k.gb = &gb;
    k.dirLightDir = cfg->dirLightDir; k.dirLightRadiance = cfg->dirLightRadiance; k.skyTintTop = cfg->skyTintTop; k.skyTintBottom = cfg->skyTintBottom;
    k.resPrev = out->resPrev; k.resCur = out->resCur; k.resLen = (out->resPrev && out->resCur) ? (long)N : 0;
    // The reference always owns reservoir buffers (Engine/Framebuffer.cs:85-97), so "k.resCur.L.Length > index" holds and
    // wroteReservoir flips after the first Lambert vertex; with both reuse flags off that has no observable effect.
    k.enableTemporalReuse = cfg->enableTemporalReuse; k.enableSpatialReuse = cfg->enableSpatialReuse; k.rngLockNoise = cfg->rngLockNoise; k.spp = cfg->spp;

    const long npx = (long)cw * ch;
    runRows([&](int y) {
        for (int x = x0; x < x1; x++) {
            FbOut fb; long oi = (long)(y - y0) * cw + (x - x0);
            fb.color = out->rgba8; fb.depth = out->depth; fb.objectId = out->objId; fb.radiance = out->radiance;
            fb.segCount = out->segCount; fb.termCode = out->termCode; fb.pathHash = out->pathHash; fb.aovStride = npx; fb.outIndex = oi;
            PathTraceKernel(y * W + x, k, cfg->maxDepth, fb);
        }
    });
    auto tC = std::chrono::steady_clock::now();

    for (int y = y0; y < y1; y++) for (int x = x0; x < x1; x++) {
        size_t gi = (size_t)y * W + x; long oi = (long)(y - y0) * cw + (x - x0);
        if (out->primId) out->primId[oi] = gb.primId[gi];
        if (out->instId) out->instId[oi] = gb.instId[gi];
        if (out->primaryT) out->primaryT[oi] = gb.primaryT[gi];
        if (out->hitMask) out->hitMask[oi] = gb.hitMask[gi];
        if (out->gbPos) { out->gbPos[3 * oi] = gb.worldPos[gi].X; out->gbPos[3 * oi + 1] = gb.worldPos[gi].Y; out->gbPos[3 * oi + 2] = gb.worldPos[gi].Z; }
        if (out->gbNrm) { out->gbNrm[3 * oi] = gb.normalWS[gi].X; out->gbNrm[3 * oi + 1] = gb.normalWS[gi].Y; out->gbNrm[3 * oi + 2] = gb.normalWS[gi].Z; }
        if (out->gbAlb) { out->gbAlb[3 * oi] = gb.baseColor[gi].X; out->gbAlb[3 * oi + 1] = gb.baseColor[gi].Y; out->gbAlb[3 * oi + 2] = gb.baseColor[gi].Z; }
        if (out->gbMat) out->gbMat[oi] = gb.matId[gi];
    }
    memset(out->counters, 0, sizeof(out->counters));
    for (auto& c : cnts) {
        out->counters[0] += c.raysPrimary; out->counters[1] += c.raysBounce; out->counters[2] += c.raysShadow;
        out->counters[3] += c.nodes; out->counters[4] += c.tris; out->counters[5] += c.spheres;
    }
    out->seconds[0] = std::chrono::duration<double>(tB - tA).count();
    out->seconds[1] = std::chrono::duration<double>(tC - tB).count();
    scene->noCull = false;
    return 0;
}
