// engine.cpp — C++ mirror of the reference's host-side Engine surface (see engine.h) + a flat C
// wrapper (eng_*) so Python (tests, bench.py) can drive it through ctypes.
// Compiled with -ffp-contract=off -fno-fast-math: host float math follows C# evaluation order.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace ILGPU_Raytracing {
namespace Engine {

namespace {
// ---- Float3.cs helpers (host) -----------------------------------------------------------------------------------------
inline Float3 F3(float x, float y, float z) { Float3 r; r.X = x; r.Y = y; r.Z = z; return r; }
inline Float3 Add(Float3 a, Float3 b) { return F3(a.X + b.X, a.Y + b.Y, a.Z + b.Z); }
inline Float3 Sub(Float3 a, Float3 b) { return F3(a.X - b.X, a.Y - b.Y, a.Z - b.Z); }
inline Float3 Mul(Float3 a, float s) { return F3(a.X * s, a.Y * s, a.Z * s); }
inline Float3 Cross(Float3 a, Float3 b) { return F3(a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X); }
inline float Dot(Float3 a, Float3 b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }
inline float Length(Float3 v) { return sqrtf(v.X * v.X + v.Y * v.Y + v.Z * v.Z); }
inline Float3 Normalize(Float3 v) { float inv = 1.0f / sqrtf(fmaxf(1e-20f, v.X * v.X + v.Y * v.Y + v.Z * v.Z)); return F3(v.X * inv, v.Y * inv, v.Z * inv); }
inline Float3 Min3(Float3 a, Float3 b) { return F3(fminf(a.X, b.X), fminf(a.Y, b.Y), fminf(a.Z, b.Z)); }
inline Float3 Max3(Float3 a, Float3 b) { return F3(fmaxf(a.X, b.X), fmaxf(a.Y, b.Y), fmaxf(a.Z, b.Z)); }
inline Float3 Center(Float3 a, Float3 b) { return F3(0.5f * (a.X + b.X), 0.5f * (a.Y + b.Y), 0.5f * (a.Z + b.Z)); }
const float kFloatMax = 3.402823466e+38f;
const float kXMathPI = 3.14159274f;   // XMath.PI (float)

inline Float3 TransformPoint(const Affine3x4& m, Float3 p) {   // Scene.cs:640-645
    return F3(m.m00 * p.X + m.m01 * p.Y + m.m02 * p.Z + m.m03, m.m10 * p.X + m.m11 * p.Y + m.m12 * p.Z + m.m13, m.m20 * p.X + m.m21 * p.Y + m.m22 * p.Z + m.m23);
}
inline Float3 TransformVector(const Affine3x4& m, Float3 v) {  // Scene.cs:647-652
    return F3(m.m00 * v.X + m.m01 * v.Y + m.m02 * v.Z, m.m10 * v.X + m.m11 * v.Y + m.m12 * v.Z, m.m20 * v.X + m.m21 * v.Y + m.m22 * v.Z);
}
void TransformAABB(const Affine3x4& m, Float3 bmin, Float3 bmax, Float3* outMin, Float3* outMax) {   // Scene.cs:560-580
    const Float3 c[8] = {F3(bmin.X, bmin.Y, bmin.Z), F3(bmax.X, bmin.Y, bmin.Z), F3(bmin.X, bmax.Y, bmin.Z), F3(bmin.X, bmin.Y, bmax.Z),
                         F3(bmax.X, bmax.Y, bmin.Z), F3(bmin.X, bmax.Y, bmax.Z), F3(bmax.X, bmin.Y, bmax.Z), F3(bmax.X, bmax.Y, bmax.Z)};
    Float3 mn = F3(kFloatMax, kFloatMax, kFloatMax), mx = F3(-kFloatMax, -kFloatMax, -kFloatMax);
    for (int i = 0; i < 8; i++) { Float3 w = TransformPoint(m, c[i]); mn = Min3(mn, w); mx = Max3(mx, w); }
    *outMin = mn; *outMax = mx;
}
Affine3x4 InvertRigidOrUniform(const Affine3x4& m, float* uniformScale) {   // Scene.cs:616-638
    float sx = Length(F3(m.m00, m.m10, m.m20)), sy = Length(F3(m.m01, m.m11, m.m21)), sz = Length(F3(m.m02, m.m12, m.m22));
    *uniformScale = (sx + sy + sz) / 3.0f;
    float inv = *uniformScale > 0.0f ? 1.0f / *uniformScale : 1.0f;
    Float3 r0 = Normalize(F3(m.m00, m.m10, m.m20)), r1 = Normalize(F3(m.m01, m.m11, m.m21)), r2 = Normalize(F3(m.m02, m.m12, m.m22));
    Affine3x4 im; memset(&im, 0, sizeof(im));
    im.m00 = r0.X * inv; im.m01 = r1.X * inv; im.m02 = r2.X * inv;
    im.m10 = r0.Y * inv; im.m11 = r1.Y * inv; im.m12 = r2.Y * inv;
    im.m20 = r0.Z * inv; im.m21 = r1.Z * inv; im.m22 = r2.Z * inv;
    Float3 it = Mul(TransformVector(im, F3(m.m03, m.m13, m.m23)), -1.0f);
    im.m03 = it.X; im.m13 = it.Y; im.m23 = it.Z;
    return im;
}

// ---- Array.Sort(int[], start, count, IComparer<int>) ------------------------------------------------------------------
// .NET's ArraySortHelper<T>.IntrospectiveSort (introsort: median-of-three quicksort, insertion sort for partitions <= 16,
// heapsort at depth 2*(log2(n)+1)), restated so that equal keys land where the reference's unstable sort puts them.
// Keys are the comparer's float (centroid coordinate); Compare = (a<b) ? -1 : (a>b) ? 1 : 0.
struct KeySorter {
    const float* key;   // indexed by the VALUE stored in idx (a position into primIdx / an instance index)
    long* ties;   // number of equal neighbours after the sort (an unstable sort may order those either way)
    int cmp(int a, int b) const { float va = key[a], vb = key[b]; if (va < vb) return -1; if (va > vb) return 1; return 0; }
    void swapIfGreater(int* k, int i, int j) const { if (cmp(k[i], k[j]) > 0) std::swap(k[i], k[j]); }
    void insertionSort(int* k, int n) const {
        for (int i = 0; i < n - 1; i++) { int t = k[i + 1]; int j = i; while (j >= 0 && cmp(t, k[j]) < 0) { k[j + 1] = k[j]; j--; } k[j + 1] = t; }
    }
    void downHeap(int* k, int i, int n) const {
        int d = k[i - 1];
        while (i <= n >> 1) { int child = 2 * i; if (child < n && cmp(k[child - 1], k[child]) < 0) child++; if (!(cmp(d, k[child - 1]) < 0)) break; k[i - 1] = k[child - 1]; i = child; }
        k[i - 1] = d;
    }
    void heapSort(int* k, int n) const {
        for (int i = n >> 1; i >= 1; i--) downHeap(k, i, n);
        for (int i = n; i > 1; i--) { std::swap(k[0], k[i - 1]); downHeap(k, 1, i - 1); }
    }
    int pickPivotAndPartition(int* k, int n) const {
        int hi = n - 1, middle = hi >> 1;
        swapIfGreater(k, 0, middle); swapIfGreater(k, 0, hi); swapIfGreater(k, middle, hi);
        int pivot = k[middle];
        std::swap(k[middle], k[hi - 1]);
        int left = 0, right = hi - 1;
        while (left < right) {
            while (cmp(k[++left], pivot) < 0) {}
            while (cmp(pivot, k[--right]) < 0) {}
            if (left >= right) break;
            std::swap(k[left], k[right]);
        }
        if (left != hi - 1) std::swap(k[left], k[hi - 1]);
        return left;
    }
    void introSort(int* k, int n, int depthLimit) const {
        int partitionSize = n;
        while (partitionSize > 1) {
            if (partitionSize <= 16) {
                if (partitionSize == 2) { swapIfGreater(k, 0, 1); return; }
                if (partitionSize == 3) { swapIfGreater(k, 0, 1); swapIfGreater(k, 0, 2); swapIfGreater(k, 1, 2); return; }
                insertionSort(k, partitionSize); return;
            }
            if (depthLimit == 0) { heapSort(k, partitionSize); return; }
            depthLimit--;
            int p = pickPivotAndPartition(k, partitionSize);
            introSort(k + p + 1, partitionSize - (p + 1), depthLimit);
            partitionSize = p;
        }
    }
    void sort(int* k, int n) const {
        if (n > 1) { int lg = 0; for (unsigned v = (unsigned)n; v >>= 1;) lg++; introSort(k, n, 2 * (lg + 1)); }
        if (ties) for (int i = 1; i < n; i++) if (key[k[i]] == key[k[i - 1]]) (*ties)++;
    }
};

void check(int status) {
    if (status != RT_OK) throw RtNativeException(status, rt_last_error());
}
}   // namespace

Affine3x4 AffineIdentity() { Affine3x4 a; memset(&a, 0, sizeof(a)); a.m00 = 1.0f; a.m11 = 1.0f; a.m22 = 1.0f; return a; }

// ======================================================================================================= Camera
void Camera::UpdateDerived(float aspectIn, float fovYRadIn) {   // Camera.cs:184-191
    forward = Normalize(Sub(Add(Add(lowerLeft, Mul(horizontal, 0.5f)), Mul(vertical, 0.5f)), origin));
    up = Normalize(vertical);
    right = Normalize(Cross(forward, up));
    aspect = aspectIn; fovYRadians = fovYRadIn;
}
static void OrthoBasis(Float3 forward, Float3 upHint, Float3* u, Float3* v, Float3* w) {   // Camera.cs:193-205
    Float3 f = Normalize(forward);
    Float3 up = upHint;
    if (fabsf(Dot(f, up)) > 0.999f) { up = F3(0, 1, 0); if (fabsf(Dot(f, up)) > 0.999f) up = F3(1, 0, 0); }
    *u = Normalize(Cross(f, up));
    *v = Normalize(Cross(*u, f));
    *w = F3(-f.X, -f.Y, -f.Z);
}
Camera Camera::CreateCamera(int width, int height, float fovDegrees) {   // Camera.cs:19-47
    return CreateCameraAt(width, height, fovDegrees, F3(0.0f, 1.0f, 3.0f), F3(0.0f, 0.5f, 0.0f));
}
// Camera.CreateCamera with its hard-wired origin / lookAt (Camera.cs:26-27) made parameters; same arithmetic.
Camera Camera::CreateCameraAt(int width, int height, float fovDegrees, Float3 origin, Float3 lookAt) {
    float aspect = (float)width / (float)std::max(1, height);
    float theta = fovDegrees * (kXMathPI / 180.0f);
    float halfHeight = tanf(0.5f * theta);
    float halfWidth = aspect * halfHeight;
    Float3 upHint = F3(0.0f, 1.0f, 0.0f);
    Float3 w = Normalize(Sub(origin, lookAt));
    Float3 u = Normalize(Cross(upHint, w));
    Float3 v = Cross(w, u);
    Camera cam; memset(&cam, 0, sizeof(cam));
    cam.origin = origin;
    cam.lowerLeft = Sub(Sub(Sub(origin, Mul(u, halfWidth)), Mul(v, halfHeight)), w);
    cam.horizontal = Mul(u, 2.0f * halfWidth);
    cam.vertical = Mul(v, 2.0f * halfHeight);
    cam.UpdateDerived(aspect, theta);
    return cam;
}
Camera Camera::LookAt(Float3 origin, Float3 lookAt, Float3 up, float vfovDegrees, float aspect, float focusDist) {   // Camera.cs:99-119
    float theta = vfovDegrees * (kXMathPI / 180.0f);
    float halfHeight = tanf(0.5f * theta);
    float halfWidth = aspect * halfHeight;
    Float3 forward = Normalize(Sub(lookAt, origin));
    Float3 u, v, w; OrthoBasis(forward, up, &u, &v, &w);
    Camera c; memset(&c, 0, sizeof(c));
    c.origin = origin;
    c.horizontal = Mul(u, 2.0f * halfWidth);
    c.vertical = Mul(v, 2.0f * halfHeight);
    c.lowerLeft = Add(Sub(Sub(origin, Mul(u, halfWidth)), Mul(v, halfHeight)), Mul(forward, focusDist));
    c.forward = Normalize(Sub(Add(Add(c.lowerLeft, Mul(c.horizontal, 0.5f)), Mul(c.vertical, 0.5f)), origin));
    c.right = Normalize(Cross(c.forward, v));
    c.up = Normalize(v);
    c.aspect = aspect; c.fovYRadians = theta;
    return c;
}
void Camera::Translate(Float3 delta) { origin = Add(origin, delta); lowerLeft = Add(lowerLeft, delta); UpdateDerived(aspect, fovYRadians); }   // :121-126
void Camera::SetFov(float vfovDegrees, float aspectIn) {   // Camera.cs:128-145
    Float3 centre = Sub(Add(Add(lowerLeft, Mul(horizontal, 0.5f)), Mul(vertical, 0.5f)), origin);
    float focusDist = Length(centre);
    Float3 fwd = Normalize(centre);
    Float3 upv = Normalize(vertical);
    float theta = vfovDegrees * (kXMathPI / 180.0f);
    float halfHeight = tanf(0.5f * theta);
    float halfWidth = aspectIn * halfHeight;
    Float3 u, v, w; OrthoBasis(fwd, upv, &u, &v, &w);
    horizontal = Mul(u, 2.0f * halfWidth);
    vertical = Mul(v, 2.0f * halfHeight);
    lowerLeft = Add(Sub(Sub(origin, Mul(u, halfWidth)), Mul(v, halfHeight)), Mul(fwd, focusDist));
    UpdateDerived(aspectIn, theta);
}
static Float3 RotateAroundAxis(Float3 v, Float3 axis, float angleRad) {   // Camera.cs:207-216
    Float3 a = Normalize(axis);
    float c = cosf(angleRad), s = sinf(angleRad);
    return Add(Add(Mul(v, c), Mul(Cross(a, v), s)), Mul(a, Dot(a, v) * (1.0f - c)));
}
void Camera::RotateYawPitch(float yawDegrees, float pitchDegrees) {   // Camera.cs:147-180
    float halfWidth = 0.5f * Length(horizontal), halfHeight = 0.5f * Length(vertical);
    Float3 centre = Sub(Add(Add(lowerLeft, Mul(horizontal, 0.5f)), Mul(vertical, 0.5f)), origin);
    float focusDist = Length(centre);
    Float3 fwd = Normalize(centre), upVec = Normalize(vertical), rightVec = Normalize(Cross(fwd, upVec)), worldUp = F3(0, 1, 0);
    float yaw = yawDegrees * (kXMathPI / 180.0f), pitch = pitchDegrees * (kXMathPI / 180.0f);
    if (fabsf(Dot(fwd, worldUp)) > 0.999f) worldUp = Normalize(Cross(rightVec, fwd));
    fwd = RotateAroundAxis(fwd, worldUp, yaw);
    upVec = RotateAroundAxis(upVec, worldUp, yaw);
    rightVec = Normalize(Cross(fwd, upVec));
    upVec = Normalize(Cross(rightVec, fwd));
    fwd = RotateAroundAxis(fwd, rightVec, pitch);
    upVec = Normalize(Cross(rightVec, fwd));
    Float3 u, v, w; OrthoBasis(fwd, upVec, &u, &v, &w);
    horizontal = Mul(u, 2.0f * halfWidth);
    vertical = Mul(v, 2.0f * halfHeight);
    lowerLeft = Add(Sub(Sub(origin, Mul(u, halfWidth)), Mul(v, halfHeight)), Mul(fwd, focusDist));
    UpdateDerived(aspect, fovYRadians);
}
void BakeCameraDerived(Camera& c, int pixelW, int pixelH) {   // RTRenderer.cs:241-263
    Float3 center = Add(Add(c.lowerLeft, Mul(c.horizontal, 0.5f)), Mul(c.vertical, 0.5f));
    Float3 forward = Normalize(Sub(center, c.origin));
    Float3 up = Normalize(c.vertical);
    Float3 right = Normalize(Cross(forward, up));
    float focusDist = Length(Sub(center, c.origin));
    float halfHeight = 0.5f * Length(c.vertical);
    float tanHalfFov = (focusDist > 1e-6f) ? (halfHeight / focusDist) : halfHeight;
    float fovY = 2.0f * atanf(tanHalfFov);
    float aspect = (Length(c.horizontal) > 1e-6f && Length(c.vertical) > 1e-6f) ? (Length(c.horizontal) / Length(c.vertical)) : ((float)pixelW / (float)std::max(1, pixelH));
    c.forward = forward; c.up = up; c.right = right; c.fovYRadians = fovY; c.aspect = aspect;
}

// ======================================================================================================= Scene
Scene::Scene(rt_ctx* native) : _native(native) {}

void Scene::Reset() { Clear(); }
void Scene::Clear() {   // Scene.cs:85-96
    _topologyVersion++;
    hTLASNodes.clear(); hTLASInstanceIndices.clear(); hInstances.clear(); hBLASNodes.clear(); hSpherePrimIndices.clear(); hSpheres.clear();
    hTriPrimIndices.clear(); hMeshPositions.clear(); hMeshTris.clear(); hMeshTexcoords.clear(); hMeshTriUVs.clear(); hTriMaterialIndex.clear();
    hMaterials.clear(); hTexInfos.clear(); hTexels.clear();
    _sortTies = 0;
}
int Scene::AddTexture(int width, int height, const RGBA32* texels) {
    _topologyVersion++;
    if (!texels) throw ArgumentNullException("texels");
    if (width <= 0 || height <= 0) throw ArgumentOutOfRangeException("texture size");
    TexInfo ti; ti.Offset = (int)hTexels.size(); ti.Width = width; ti.Height = height;
    hTexels.insert(hTexels.end(), texels, texels + (size_t)width * height);
    hTexInfos.push_back(ti);
    return (int)hTexInfos.size() - 1;
}
int Scene::AddSphere(const Sphere& s) { _topologyVersion++; int id = (int)hSpheres.size(); hSpheres.push_back(s); hSpherePrimIndices.push_back(id); return id; }   // :315-321

int Scene::BuildBLASNodeRecursive(int* idx, int start, int count, const Float3* bminPre, const Float3* bmaxPre, int parentSkip, bool spheres) {   // Scene.cs:405-467
    std::vector<int>& primIdx = spheres ? hSpherePrimIndices : hTriPrimIndices;
    int nodeIndex = (int)hBLASNodes.size();
    BLASNode node; node.first = -1; node.count = 0; node.left = -1; node.right = -1; node.skipIndex = parentSkip;
    Float3 nbMin = F3(kFloatMax, kFloatMax, kFloatMax), nbMax = F3(-kFloatMax, -kFloatMax, -kFloatMax);
    if (bminPre) {   // indexed by POSITION, as the reference does (quirk 1 of SURVEY.md §8a)
        for (int i = start; i < start + count; i++) { nbMin = Min3(nbMin, bminPre[i]); nbMax = Max3(nbMax, bmaxPre[i]); }
    } else {
        for (int i = start; i < start + count; i++) {
            const MeshTri& t = hMeshTris[primIdx[idx[i]]];   // BoundsOfTriangle :597-605
            Float3 v0 = hMeshPositions[t.i0], v1 = hMeshPositions[t.i1], v2 = hMeshPositions[t.i2];
            nbMin = Min3(nbMin, Min3(v0, Min3(v1, v2))); nbMax = Max3(nbMax, Max3(v0, Max3(v1, v2)));
        }
    }
    node.boundsMin = nbMin; node.boundsMax = nbMax;
    hBLASNodes.push_back(node);
    if (count <= 4) {   // LeafThreshold :436
        int leafStart = (int)primIdx.size();
        for (int i = start; i < start + count; i++) primIdx.push_back(primIdx[idx[i]]);
        hBLASNodes[nodeIndex].first = leafStart; hBLASNodes[nodeIndex].count = count; hBLASNodes[nodeIndex].skipIndex = parentSkip;
        return nodeIndex;
    }
    Float3 extent = Sub(nbMax, nbMin);
    int axis = 0;
    if (extent.Y > extent.X && extent.Y >= extent.Z) axis = 1;
    else if (extent.Z > extent.X && extent.Z >= extent.Y) axis = 2;
    if (spheres) {   // BLASPrimComparatorSpheres :512-526: key = centre[axis] of spheres[primIdx[a]]
        std::vector<float> key(primIdx.size());
        for (int i = start; i < start + count; i++) { const Sphere& s = hSpheres[primIdx[idx[i]]]; key[idx[i]] = axis == 0 ? s.center.X : (axis == 1 ? s.center.Y : s.center.Z); }
        KeySorter ks = {key.data(), &_sortTies}; ks.sort(idx + start, count);
    } else {         // BLASPrimComparatorTris :528-543: key = centroid[axis] of triangle primIdx[a]; idx holds identity positions
        KeySorter ks = {_triKey[axis].data(), &_sortTies}; ks.sort(idx + start, count);
    }
    int mid = start + (count >> 1);
    int rightRoot = BuildBLASNodeRecursive(idx, mid, count - (mid - start), bminPre, bmaxPre, parentSkip, spheres);
    int leftRoot = BuildBLASNodeRecursive(idx, start, mid - start, bminPre, bmaxPre, rightRoot, spheres);
    hBLASNodes[nodeIndex].left = leftRoot; hBLASNodes[nodeIndex].right = rightRoot; hBLASNodes[nodeIndex].skipIndex = parentSkip;
    return nodeIndex;
}
void Scene::BuildBLAS_Spheres(int primStart, int primCount) {   // Scene.cs:381-396
    std::vector<int> idx((size_t)primCount);
    for (int i = 0; i < primCount; i++) idx[i] = primStart + i;
    std::vector<Float3> bmin((size_t)primCount), bmax((size_t)primCount);
    for (int i = 0; i < primCount; i++) {
        const Sphere& s = hSpheres[hSpherePrimIndices[primStart + i]];
        bmin[i] = F3(s.center.X - s.radius, s.center.Y - s.radius, s.center.Z - s.radius);
        bmax[i] = F3(s.center.X + s.radius, s.center.Y + s.radius, s.center.Z + s.radius);
    }
    BuildBLASNodeRecursive(idx.data(), 0, primCount, bmin.data(), bmax.data(), -1, true);
}
void Scene::BuildBLAS_Triangles(int primStart, int primCount) {   // Scene.cs:398-403
    std::vector<int> idx((size_t)primCount);
    for (int i = 0; i < primCount; i++) idx[i] = primStart + i;
    // centroid keys, CenterOfTriangle (Scene.cs:607-614): (v0+v1+v2)/3 per component, keyed by POSITION in primIdx
    // (primIdx[primStart+i] == primStart+i for the single mesh the reference supports, quirk 2).
    for (int a = 0; a < 3; a++) _triKey[a].assign(hTriPrimIndices.size(), 0.0f);
    for (int i = 0; i < primCount; i++) {
        const MeshTri& t = hMeshTris[hTriPrimIndices[primStart + i]];
        Float3 v0 = hMeshPositions[t.i0], v1 = hMeshPositions[t.i1], v2 = hMeshPositions[t.i2];
        _triKey[0][primStart + i] = (v0.X + v1.X + v2.X) / 3.0f; _triKey[1][primStart + i] = (v0.Y + v1.Y + v2.Y) / 3.0f; _triKey[2][primStart + i] = (v0.Z + v1.Z + v2.Z) / 3.0f;
    }
    BuildBLASNodeRecursive(idx.data(), 0, primCount, nullptr, nullptr, -1, false);
}
InstanceRecord Scene::BuildSphereInstance(const int* sphereIds, int n, const Affine3x4& objectToWorld) {   // Scene.cs:323-356
    if (!sphereIds) throw ArgumentNullException("sphereIds");
    if (n <= 0) throw ArgumentOutOfRangeException("sphereIds.Length");
    Float3 bmin = F3(kFloatMax, kFloatMax, kFloatMax), bmax = F3(-kFloatMax, -kFloatMax, -kFloatMax);
    for (int i = 0; i < n; i++) {
        if (sphereIds[i] < 0 || sphereIds[i] >= (int)hSpheres.size()) throw ArgumentOutOfRangeException("sphereIds");
        const Sphere& s = hSpheres[sphereIds[i]];
        bmin = Min3(bmin, F3(s.center.X - s.radius, s.center.Y - s.radius, s.center.Z - s.radius));
        bmax = Max3(bmax, F3(s.center.X + s.radius, s.center.Y + s.radius, s.center.Z + s.radius));
    }
    int primStart = sphereIds[0], primCount = n;
    int blasStart = (int)hBLASNodes.size();
    BuildBLAS_Spheres(primStart, primCount);
    int blasCount = (int)hBLASNodes.size() - blasStart;
    Float3 wmin, wmax; TransformAABB(objectToWorld, bmin, bmax, &wmin, &wmax);
    float uniScale; Affine3x4 worldToObject = InvertRigidOrUniform(objectToWorld, &uniScale);
    InstanceRecord inst; memset(&inst, 0, sizeof(inst));
    inst.type = RT_BLAS_SPHERESET; inst.blasRoot = blasStart; inst.blasNodeCount = blasCount; inst.primIndexFirst = primStart; inst.primIndexCount = primCount;
    inst.objectToWorld = objectToWorld; inst.worldToObject = worldToObject; inst.uniformScale = uniScale; inst.worldBoundsMin = wmin; inst.worldBoundsMax = wmax;
    return inst;
}
void Scene::AddSphereInstance(const int* sphereIds, int n, const Affine3x4& objectToWorld) { _topologyVersion++; hInstances.push_back(BuildSphereInstance(sphereIds, n, objectToWorld)); }

int Scene::BuildTLASNodeRecursive(int* idx, int start, int count, int parentSkip) {   // Scene.cs:469-510
    int nodeIndex = (int)hTLASNodes.size();
    TLASNode node; node.first = -1; node.count = 0; node.left = -1; node.right = -1; node.skipIndex = parentSkip;
    Float3 nbMin = F3(kFloatMax, kFloatMax, kFloatMax), nbMax = F3(-kFloatMax, -kFloatMax, -kFloatMax);
    for (int i = start; i < start + count; i++) { const InstanceRecord& r = hInstances[idx[i]]; nbMin = Min3(nbMin, r.worldBoundsMin); nbMax = Max3(nbMax, r.worldBoundsMax); }
    node.boundsMin = nbMin; node.boundsMax = nbMax;
    hTLASNodes.push_back(node);
    if (count <= 2) {   // LeafThreshold :486
        hTLASNodes[nodeIndex].first = start; hTLASNodes[nodeIndex].count = count; hTLASNodes[nodeIndex].skipIndex = parentSkip;
        return nodeIndex;
    }
    Float3 extent = Sub(nbMax, nbMin);
    int axis = 0;
    if (extent.Y > extent.X && extent.Y >= extent.Z) axis = 1;
    else if (extent.Z > extent.X && extent.Z >= extent.Y) axis = 2;
    std::vector<float> key(hInstances.size());   // TLASInstComparator :545-558
    for (int i = start; i < start + count; i++) { Float3 c = Center(hInstances[idx[i]].worldBoundsMin, hInstances[idx[i]].worldBoundsMax); key[idx[i]] = axis == 0 ? c.X : (axis == 1 ? c.Y : c.Z); }
    KeySorter ks = {key.data(), &_sortTies}; ks.sort(idx + start, count);
    int mid = start + (count >> 1);
    int rightRoot = BuildTLASNodeRecursive(idx, mid, count - (mid - start), parentSkip);
    int leftRoot = BuildTLASNodeRecursive(idx, start, mid - start, rightRoot);
    hTLASNodes[nodeIndex].left = leftRoot; hTLASNodes[nodeIndex].right = rightRoot; hTLASNodes[nodeIndex].skipIndex = parentSkip;
    return nodeIndex;
}
void Scene::RebuildTLAS() {   // Scene.cs:358-368
    _topologyVersion++;
    int n = (int)hInstances.size();
    hTLASInstanceIndices.resize((size_t)n);
    for (int i = 0; i < n; i++) hTLASInstanceIndices[i] = i;
    hTLASNodes.clear();
    if (n > 0) BuildTLASNodeRecursive(hTLASInstanceIndices.data(), 0, n, -1);
}
void Scene::LoadMeshInstance(const Float3* positions, int nPositions, const MeshTri* tris, int nTris, const Float2* texcoords, int nTexcoords,
                             const MeshTriUV* triUVs, const int* triMaterialIndex, const MaterialRecord* materials, int nMaterials, const Affine3x4& objectToWorld) {   // Scene.cs:144-256
    if (!positions || !tris || !texcoords || !triUVs || !materials) throw ArgumentNullException("mesh arrays");
    _topologyVersion++;
    if (nPositions <= 0 || nTris <= 0 || nTexcoords <= 0 || nMaterials <= 0) throw ArgumentOutOfRangeException("mesh array length");
    if (!hMeshTris.empty()) throw InvalidOperationException("the reference supports one mesh per scene: a second LoadObjInstance breaks its primIdx identity assumption (Scene.cs:384,401,439-440)");
    int baseVertex = (int)hMeshPositions.size(), baseTri = (int)hMeshTris.size(), baseUV = (int)hMeshTexcoords.size(), baseMat = (int)hMaterials.size();
    hMeshPositions.insert(hMeshPositions.end(), positions, positions + nPositions);
    hMeshTexcoords.insert(hMeshTexcoords.end(), texcoords, texcoords + nTexcoords);
    for (int i = 0; i < nTris; i++) {
        MeshTri t = tris[i];
        if (t.i0 < 0 || t.i1 < 0 || t.i2 < 0 || t.i0 >= nPositions || t.i1 >= nPositions || t.i2 >= nPositions) throw ArgumentOutOfRangeException("triangle vertex index");
        t.i0 += baseVertex; t.i1 += baseVertex; t.i2 += baseVertex; hMeshTris.push_back(t);
        MeshTriUV tuv = triUVs[i];
        if (tuv.t0 < 0 || tuv.t1 < 0 || tuv.t2 < 0 || tuv.t0 >= nTexcoords || tuv.t1 >= nTexcoords || tuv.t2 >= nTexcoords) throw ArgumentOutOfRangeException("triangle texcoord index");
        tuv.t0 += baseUV; tuv.t1 += baseUV; tuv.t2 += baseUV; hMeshTriUVs.push_back(tuv);
        int ml = triMaterialIndex ? triMaterialIndex[i] : 0;
        if (ml < 0 || ml >= nMaterials) throw ArgumentOutOfRangeException("triangle material index");
        hTriMaterialIndex.push_back(baseMat + ml);
        hTriPrimIndices.push_back(baseTri + i);
    }
    hMaterials.insert(hMaterials.end(), materials, materials + nMaterials);
    int blasStart = (int)hBLASNodes.size();
    BuildBLAS_Triangles(baseTri, nTris);
    int blasCount = (int)hBLASNodes.size() - blasStart;
    Float3 bmin = F3(kFloatMax, kFloatMax, kFloatMax), bmax = F3(-kFloatMax, -kFloatMax, -kFloatMax);   // ComputeMeshBounds :582-595
    for (int i = 0; i < nTris; i++) {
        const MeshTri& t = hMeshTris[baseTri + i];
        Float3 v0 = hMeshPositions[t.i0], v1 = hMeshPositions[t.i1], v2 = hMeshPositions[t.i2];
        bmin = Min3(bmin, Min3(v0, Min3(v1, v2))); bmax = Max3(bmax, Max3(v0, Max3(v1, v2)));
    }
    Float3 wmin, wmax; TransformAABB(objectToWorld, bmin, bmax, &wmin, &wmax);
    float uniScale; Affine3x4 worldToObject = InvertRigidOrUniform(objectToWorld, &uniScale);
    InstanceRecord r; memset(&r, 0, sizeof(r));
    r.type = RT_BLAS_TRIMESH; r.blasRoot = blasStart; r.blasNodeCount = blasCount; r.primIndexFirst = baseTri; r.primIndexCount = nTris;
    r.objectToWorld = objectToWorld; r.worldToObject = worldToObject; r.uniformScale = uniScale; r.worldBoundsMin = wmin; r.worldBoundsMax = wmax;
    hInstances.push_back(r);
    RebuildTLAS();
}
void Scene::BuildDefaultScene() {   // Scene.cs:83-142
    _topologyVersion++;
    Clear();
    auto checker = [&](int w, int h, int step, RGBA32 c0, RGBA32 c1) {   // AddCheckerTexture :98-109
        std::vector<RGBA32> px((size_t)w * h);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) px[(size_t)y * w + x] = ((((x / step) + (y / step)) & 1) == 0) ? c0 : c1;
        return AddTexture(w, h, px.data());
    };
    RGBA32 white = {255, 255, 255, 255}, grey = {20, 20, 20, 255}, blue = {40, 40, 200, 255}, yellow = {200, 200, 40, 255};
    int checker0 = checker(256, 256, 16, white, grey);
    int checker1 = checker(256, 256, 8, blue, yellow);
    auto mat = [](Float3 kd, int hasMap, int tex) { MaterialRecord m; m.Kd = kd; m.HasDiffuseMap = hasMap; m.DiffuseTexIndex = tex; m.Shading = RT_SHADING_LAMBERT; m.IOR = 1.0f; m.HasAlphaMap = 0; m.AlphaTexIndex = -1; m.AlphaCutoff = 0.5f; m.TwoSided = 0; return m; };
    MaterialRecord matGround = mat(F3(1, 1, 1), 1, checker0), matRed = mat(F3(0.8f, 0.3f, 0.3f), 0, -1), matGreen = mat(F3(0.3f, 0.8f, 0.3f), 0, -1);
    MaterialRecord matTex = mat(F3(1, 1, 1), 1, checker1), matWhite = mat(F3(1, 1, 1), 0, -1);
    auto sph = [](Float3 c, float r, Float3 alb, MaterialRecord m, int shading, float ior) { Sphere s; s.center = c; s.radius = r; s.albedo = alb; s.material = m; s.shading = shading; s.ior = ior; return s; };
    int ids[6];
    ids[0] = AddSphere(sph(F3(0.0f, -1000.5f, 0.0f), 1000.0f, F3(1, 1, 1), matGround, RT_SHADING_LAMBERT, 1.0f));
    ids[1] = AddSphere(sph(F3(-0.9f, 0.5f, -0.2f), 0.5f, F3(0.8f, 0.3f, 0.3f), matRed, RT_SHADING_LAMBERT, 1.0f));
    ids[2] = AddSphere(sph(F3(0.9f, 0.35f, 0.2f), 0.35f, F3(0.3f, 0.8f, 0.3f), matGreen, RT_SHADING_LAMBERT, 1.0f));
    ids[3] = AddSphere(sph(F3(0.0f, 0.75f, 0.6f), 0.75f, F3(1, 1, 1), matTex, RT_SHADING_LAMBERT, 1.0f));
    ids[4] = AddSphere(sph(F3(-1.8f, 0.5f, 0.8f), 0.5f, F3(1, 1, 1), matWhite, RT_SHADING_MIRROR, 1.0f));
    ids[5] = AddSphere(sph(F3(1.8f, 0.5f, -0.8f), 0.5f, F3(1, 1, 1), matWhite, RT_SHADING_GLASS, 1.5f));
    Affine3x4 I = AffineIdentity();
    for (int i = 0; i < 6; i++) hInstances.push_back(BuildSphereInstance(&ids[i], 1, I));
    // TryAddSponzaFromKnownLocations (:654-674): no .obj ships with the reference; the OBJ path is out of scope here.
    RebuildTLAS();
}
void Scene::FillDesc(RtSceneDesc* d) const {
    memset(d, 0, sizeof(*d));
    d->tlasNodes = hTLASNodes.data(); d->nTlasNodes = (int64_t)hTLASNodes.size();
    d->tlasInstanceIndices = hTLASInstanceIndices.data(); d->nTlasInstanceIndices = (int64_t)hTLASInstanceIndices.size();
    d->instances = hInstances.data(); d->nInstances = (int64_t)hInstances.size();
    d->blasNodes = hBLASNodes.data(); d->nBlasNodes = (int64_t)hBLASNodes.size();
    d->spherePrimIdx = hSpherePrimIndices.data(); d->nSpherePrimIdx = (int64_t)hSpherePrimIndices.size();
    d->spheres = hSpheres.data(); d->nSpheres = (int64_t)hSpheres.size();
    d->triPrimIdx = hTriPrimIndices.data(); d->nTriPrimIdx = (int64_t)hTriPrimIndices.size();
    d->meshPositions = hMeshPositions.data(); d->nMeshPositions = (int64_t)hMeshPositions.size();
    d->meshTris = hMeshTris.data(); d->nMeshTris = (int64_t)hMeshTris.size();
    d->meshTexcoords = hMeshTexcoords.data(); d->nMeshTexcoords = (int64_t)hMeshTexcoords.size();
    d->meshTriUVs = hMeshTriUVs.data(); d->nMeshTriUVs = (int64_t)hMeshTriUVs.size();
    d->triMatIndex = hTriMaterialIndex.data(); d->nTriMatIndex = (int64_t)hTriMaterialIndex.size();
    d->materials = hMaterials.data(); d->nMaterials = (int64_t)hMaterials.size();
    d->texels = hTexels.data(); d->nTexels = (int64_t)hTexels.size();
    d->texInfos = hTexInfos.data(); d->nTexInfos = (int64_t)hTexInfos.size();
}
void Scene::UploadAll() {   // Scene.cs:258-279
    if (!_native) throw InvalidOperationException("Scene has no native context (host-only scene)");
    RtSceneDesc d; FillDesc(&d);
    check(rt_scene_upload_ex(_native, &d, DeviceBuild ? (uint32_t)RT_BUILD_DEVICE_LBVH : 0u));
    _uploadedVersion = _topologyVersion;
}
// New vertex positions for the loaded mesh, same count and triangles (the reference has no such entry point: its meshes never move;
// this is what makes RebuildPolicy.ForceRefit mean something, BvhManager.cs:13-27)
void Scene::SetMeshPositions(const Float3* positions, int nPositions) {
    if (!positions) throw ArgumentNullException("positions");
    if ((size_t)nPositions != hMeshPositions.size()) throw ArgumentOutOfRangeException("positions: the vertex count must stay the same");
    std::copy(positions, positions + nPositions, hMeshPositions.begin());
}
bool Scene::CanRefit() const { return _native && _uploadedVersion == _topologyVersion && !hMeshPositions.empty(); }
void Scene::RefitUpload() {
    if (!_native) throw InvalidOperationException("Scene has no native context (host-only scene)");
    if (!CanRefit()) throw InvalidOperationException("refit needs the topology of the last UploadAll");
    check(rt_scene_refit(_native, hMeshPositions.data(), (int64_t)hMeshPositions.size()));
}

// ======================================================================================================= Framebuffer
void Framebuffer::DownloadToCpu(int slot) {   // Framebuffer.cs:148-156
    if (slot != 0) throw ArgumentOutOfRangeException("slot");   // the reference allocates 3 slots but only ever uses slot 0 (RTRenderer.cs:164)
    size_t nb = 0;
    const int bc = _gathered ? RT_BUF_GATHERED_RGBA8 : RT_BUF_RGBA8, bd = _gathered ? RT_BUF_GATHERED_DEPTH : RT_BUF_DEPTH, bo = _gathered ? RT_BUF_GATHERED_OBJID : RT_BUF_OBJID;
    check(rt_buffer_bytes(_native, bc, &nb));
    size_t n = nb / 4;
    _cpuColor.resize(n); _cpuDepth.resize(n); _cpuObjectId.resize(n);
    check(rt_download(_native, bc, _cpuColor.data(), nb));
    check(rt_download(_native, bd, _cpuDepth.data(), nb));
    check(rt_download(_native, bo, _cpuObjectId.data(), nb));
}
void Framebuffer::BindCpuTargets(int* color, float* depth, int* objectId, size_t n) {
    check(rt_bind_readback(_native, RT_BUF_RGBA8, color, n * 4));
    check(rt_bind_readback(_native, RT_BUF_DEPTH, depth, n * 4));
    check(rt_bind_readback(_native, RT_BUF_OBJID, objectId, n * 4));
    _boundColor = color; _boundDepth = depth; _boundObjectId = objectId; _boundN = n;
}
void Framebuffer::DownloadToCpu(int slot, int* color, float* depth, int* objectId, size_t n) {
    if (slot != 0) throw ArgumentOutOfRangeException("slot");
    if (!color || !depth || !objectId) throw ArgumentNullException("destination");
    if (color == _boundColor && depth == _boundDepth && objectId == _boundObjectId && n == _boundN) { check(rt_sync(_native)); return; }   // bound targets: the frame already copied them
    size_t nb = 0;
    const int bc = _gathered ? RT_BUF_GATHERED_RGBA8 : RT_BUF_RGBA8, bd = _gathered ? RT_BUF_GATHERED_DEPTH : RT_BUF_DEPTH, bo = _gathered ? RT_BUF_GATHERED_OBJID : RT_BUF_OBJID;
    check(rt_buffer_bytes(_native, bc, &nb));
    if (nb / 4 != n) throw ArgumentOutOfRangeException("n");
    check(rt_download_async(_native, bc, color, nb));   // three copies, one wait
    check(rt_download_async(_native, bd, depth, nb));
    check(rt_download_async(_native, bo, objectId, nb));
    check(rt_sync(_native));
}

// ======================================================================================================= RTRenderer
RTRenderer::RTRenderer(int deviceIndex, int windowWidth, int windowHeight) {   // RTRenderer.cs:63-92
    int dev = deviceIndex;
    check(rt_create(&dev, 1, &_native));
    _sceneManager = new SceneManager(_native);
    _sceneManager->BuildDefaultScene();               // :71
    _sceneManager->Commit(RebuildPolicy::Auto);       // :72
    int w = std::max(1, windowWidth), h = std::max(1, windowHeight);
    _camera = Camera::CreateCamera(w, h, 60.0f);      // :78
    _camera.Translate(F3(1, 0, -4));                  // :79
    _prevCamera = _camera;
    _framebuffer = new Framebuffer(_native);
    memset(&_lastCfg, 0, sizeof(_lastCfg));
}
RTRenderer::~RTRenderer() {
    delete _framebuffer; delete _sceneManager;
    if (_native) rt_destroy(_native);
}
void RTRenderer::Synchronize() { check(rt_sync(_native)); }
void RTRenderer::NewCommunicatorId(void* id128) { if (!id128) throw ArgumentNullException("id128"); check(rt_comm_get_unique_id(id128, RT_COMM_ID_BYTES)); }
void RTRenderer::InitMultiGpu(const void* uniqueId, int rank, int worldSize) {
    if (!uniqueId) throw ArgumentNullException("uniqueId");
    if (worldSize < 1 || rank < 0 || rank >= worldSize) throw ArgumentOutOfRangeException("rank");
    check(rt_comm_init(_native, uniqueId, RT_COMM_ID_BYTES, rank, worldSize));
    Rank = rank; WorldSize = worldSize; _multiGpu = true;
    _framebuffer->SetGathered(worldSize > 1);
}

void RTRenderer::RenderDirectToPbo(void* pboDevicePtr, int width, int height, int frame, float dt) {   // RTRenderer.cs:105-237
    int outW = std::max(1, width), outH = std::max(1, height);
    int inW = std::max(1, (int)rintf((float)outW * RenderScale)), inH = std::max(1, (int)rintf((float)outH * RenderScale));   // :113-116 (XMath.Round)
    BakeCameraDerived(_camera, inW, inH);              // :123-124
    BakeCameraDerived(_prevCamera, inW, inH);
    int temporalSeed = (RngLockNoise == 0) ? 0 : FixedSeed;   // :166 (Random.Shared.Next() replaced by a caller-chosen seed)
    float dtClamped = fmaxf(fminf(dt, 0.1f), 0.0f);    // :169
    _sunAzimuth += _sunSpeedRadPerSec * dtClamped;
    const float TwoPi = 6.28318530717958647692f;
    if (_sunAzimuth >= TwoPi) _sunAzimuth -= TwoPi; else if (_sunAzimuth < 0.0f) _sunAzimuth += TwoPi;
    Float3 sunDir = Normalize(F3(cosf(_sunAzimuth) * cosf(_sunElevation), sinf(_sunElevation), sinf(_sunAzimuth) * cosf(_sunElevation)));   // :174-178
    RtRenderConfig cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.width = inW; cfg.height = inH; cfg.frame = frame; cfg.spp = Spp; cfg.maxDepth = MaxDepth; cfg.rngLockNoise = temporalSeed;
    cfg.enableTemporalReuse = EnableTemporalReuse; cfg.enableSpatialReuse = EnableSpatialReuse;
    cfg.dirLightDir = sunDir; cfg.dirLightRadiance = F3(10, 10, 10); cfg.skyTintTop = F3(0.5f, 0.7f, 1.0f); cfg.skyTintBottom = F3(1.0f, 1.0f, 1.0f);   // :191-194
    cfg.flags = Flags; cfg.tileSize = TileSize; cfg.rank = Rank; cfg.worldSize = WorldSize; cfg.samplesPerPass = SamplesPerPass;
    check(rt_render(_native, &_camera, &_prevCamera, &cfg));   // the two kernel launches :152-153, :181-205
    _lastCfg = cfg;
    if (WorldSize > 1 && _multiGpu)   // every rank: its tiles' colour + depth + objectId to rank 0 (NCCL, inside the library)
        check(rt_gather_frame(_native, 0, RT_GATHER_RGBA8 | RT_GATHER_DEPTH_OBJID));
    if (WorldSize <= 1 || (_multiGpu && Rank == 0)) {   // Present (:208-231): TAAU resolve, or blit / bilinear upsample, into the mapped PBO (or the core's own buffer when headless)
        RtPresentConfig pc; memset(&pc, 0, sizeof(pc));
        pc.mode = EnableTAAU ? RT_PRESENT_TAAU : RT_PRESENT_COPY; pc.outWidth = outW; pc.outHeight = outH;
        pc.feedback = 0.075f; pc.sharpness = 0.10f; pc.clampK = 1.25f;   // RTTaa.cs:80-82
        check(rt_present(_native, &pc, pboDevicePtr, pboDevicePtr ? (size_t)outW * outH * 4 : 0));
    }
    if (!AsyncSubmit) check(rt_sync(_native));         // _cuda.Synchronize() :233 (AsyncSubmit: the caller waits later - Synchronize() - so that frames queue back to back)
    _prevCamera = _camera;                             // :236
}

}   // namespace Engine
}   // namespace ILGPU_Raytracing

// =========================================================================================================== C wrapper
using namespace ILGPU_Raytracing::Engine;
#define ENG_API extern "C" __attribute__((visibility("default")))
static thread_local std::string g_engErr;
template <class F> static int guard(F&& f) {
    try { f(); return 0; }
    catch (const RtNativeException& e) { g_engErr = std::string("RtNativeException: ") + e.what(); return e.status; }
    catch (const ArgumentException& e) { g_engErr = std::string("ArgumentException: ") + e.what(); return -108; }
    catch (const ArgumentNullException& e) { g_engErr = std::string("ArgumentNullException: ") + e.what(); return -101; }
    catch (const ArgumentOutOfRangeException& e) { g_engErr = std::string("ArgumentOutOfRangeException: ") + e.what(); return -102; }
    catch (const InvalidOperationException& e) { g_engErr = std::string("InvalidOperationException: ") + e.what(); return -103; }
    catch (const FileNotFoundException& e) { g_engErr = std::string("FileNotFoundException: ") + e.what(); return -104; }
    catch (const InvalidDataException& e) { g_engErr = std::string("InvalidDataException: ") + e.what(); return -105; }
    catch (const EndOfStreamException& e) { g_engErr = std::string("EndOfStreamException: ") + e.what(); return -106; }
    catch (const FormatException& e) { g_engErr = std::string("FormatException: ") + e.what(); return -107; }
    catch (const std::exception& e) { g_engErr = e.what(); return -100; }
}
ENG_API const char* eng_last_error() { return g_engErr.c_str(); }
// host-only Scene (no native context): lets CPU tests exercise the builders
ENG_API Scene* eng_scene_new_hostonly() { return new Scene(nullptr); }
ENG_API void eng_scene_free_hostonly(Scene* s) { delete s; }
ENG_API int eng_scene_build_default(Scene* s) { return guard([&] { s->BuildDefaultScene(); }); }
ENG_API int eng_scene_add_texture(Scene* s, int w, int h, const RGBA32* px, int* outIndex) { return guard([&] { *outIndex = s->AddTexture(w, h, px); }); }
ENG_API int eng_scene_add_sphere(Scene* s, const Sphere* sp, int* outIndex) { return guard([&] { if (!sp) throw ArgumentNullException("sphere"); *outIndex = s->AddSphere(*sp); }); }
ENG_API int eng_scene_add_sphere_instance(Scene* s, const int* ids, int n, const Affine3x4* o2w) { return guard([&] { s->AddSphereInstance(ids, n, o2w ? *o2w : AffineIdentity()); }); }
ENG_API int eng_scene_load_mesh_instance(Scene* s, const Float3* pos, int nPos, const MeshTri* tris, int nTris, const Float2* uv, int nUV, const MeshTriUV* triUVs,
                                         const int* triMat, const MaterialRecord* mats, int nMats, const Affine3x4* o2w) {
    return guard([&] { s->LoadMeshInstance(pos, nPos, tris, nTris, uv, nUV, triUVs, triMat, mats, nMats, o2w ? *o2w : AffineIdentity()); });
}
ENG_API int eng_scene_load_obj_instance(Scene* s, const char* path, const Affine3x4* o2w, float uniformScale) {
    return guard([&] { if (!path) throw ArgumentNullException("objPath"); s->LoadObjInstance(path, o2w ? *o2w : AffineIdentity(), uniformScale); });
}
ENG_API int eng_scene_rebuild_tlas(Scene* s) { return guard([&] { s->RebuildTLAS(); }); }
ENG_API int eng_scene_upload_all(Scene* s) { return guard([&] { s->UploadAll(); }); }
ENG_API long eng_scene_sort_ties(Scene* s) { return s->SortTies(); }
ENG_API void eng_scene_fill_desc(Scene* s, RtSceneDesc* d) { s->FillDesc(d); }
ENG_API void eng_scene_reset(Scene* s) { s->Reset(); }   // SceneManager.ReplaceScene(new Scene) analogue (SceneManager.cs:32-36)

ENG_API void eng_camera_create(int w, int h, float fov, RtCamera* out) { Camera c = Camera::CreateCamera(w, h, fov); memcpy(out, &c, sizeof(RtCamera)); }
ENG_API void eng_camera_create_at(int w, int h, float fov, const float* origin, const float* lookAt, RtCamera* out) {
    Float3 o = {origin[0], origin[1], origin[2]}, l = {lookAt[0], lookAt[1], lookAt[2]};
    Camera c = Camera::CreateCameraAt(w, h, fov, o, l); memcpy(out, &c, sizeof(RtCamera));
}
ENG_API void eng_camera_look_at(const float* origin, const float* lookAt, const float* up, float fov, float aspect, float focus, RtCamera* out) {
    Float3 o = {origin[0], origin[1], origin[2]}, l = {lookAt[0], lookAt[1], lookAt[2]}, u = {up[0], up[1], up[2]};
    Camera c = Camera::LookAt(o, l, u, fov, aspect, focus); memcpy(out, &c, sizeof(RtCamera));
}
ENG_API void eng_camera_translate(RtCamera* cam, float dx, float dy, float dz) { Camera c; memcpy(&c, cam, sizeof(RtCamera)); Float3 d = {dx, dy, dz}; c.Translate(d); memcpy(cam, &c, sizeof(RtCamera)); }
ENG_API void eng_camera_set_fov(RtCamera* cam, float fov, float aspect) { Camera c; memcpy(&c, cam, sizeof(RtCamera)); c.SetFov(fov, aspect); memcpy(cam, &c, sizeof(RtCamera)); }
ENG_API void eng_camera_rotate_yaw_pitch(RtCamera* cam, float yaw, float pitch) { Camera c; memcpy(&c, cam, sizeof(RtCamera)); c.RotateYawPitch(yaw, pitch); memcpy(cam, &c, sizeof(RtCamera)); }
ENG_API void eng_camera_bake(RtCamera* cam, int w, int h) { Camera c; memcpy(&c, cam, sizeof(RtCamera)); BakeCameraDerived(c, w, h); memcpy(cam, &c, sizeof(RtCamera)); }

ENG_API int eng_renderer_new(int deviceIndex, int w, int h, RTRenderer** out) { return guard([&] { *out = new RTRenderer(deviceIndex, w, h); }); }
ENG_API void eng_renderer_free(RTRenderer* r) { delete r; }
ENG_API rt_ctx* eng_renderer_native(RTRenderer* r) { return r->Native(); }
ENG_API Scene* eng_renderer_scene(RTRenderer* r) { return &r->Scenes().GetScene(); }
ENG_API int eng_renderer_commit(RTRenderer* r) { return guard([&] { r->Scenes().Commit(RebuildPolicy::Auto); }); }
ENG_API int eng_renderer_commit_policy(RTRenderer* r, int policy) {
    return guard([&] {
        if (policy < 0 || policy > 2) throw ArgumentOutOfRangeException("policy");
        r->Scenes().Commit((RebuildPolicy)policy);
    });
}
ENG_API int eng_scene_set_mesh_positions(Scene* s, const Float3* pos, int n) { return guard([&] { s->SetMeshPositions(pos, n); }); }
ENG_API int eng_scene_can_refit(Scene* s) { return s->CanRefit() ? 1 : 0; }
ENG_API void eng_scene_set_device_build(Scene* s, int on) { s->DeviceBuild = on != 0; }
ENG_API void eng_renderer_get_camera(RTRenderer* r, RtCamera* out) { memcpy(out, &r->Cam(), sizeof(RtCamera)); }
ENG_API void eng_renderer_set_camera(RTRenderer* r, const RtCamera* in) { memcpy(static_cast<RtCamera*>(&r->Cam()), in, sizeof(RtCamera)); }
ENG_API void eng_renderer_set_sun_params(RTRenderer* r, float speed, float elevation) { r->SetSunParams(speed, elevation); }
// knobs: 0 RenderScale(float bits not used) ... use a struct instead
struct EngKnobs { float renderScale; int enableTemporalReuse, enableSpatialReuse, rngLockNoise, fixedSeed, spp, maxDepth; unsigned flags; int tileSize, rank, worldSize, samplesPerPass, enableTAAU, asyncSubmit; };
ENG_API void eng_renderer_set_knobs(RTRenderer* r, const EngKnobs* k) {
    r->RenderScale = k->renderScale; r->EnableTemporalReuse = k->enableTemporalReuse; r->EnableSpatialReuse = k->enableSpatialReuse; r->RngLockNoise = k->rngLockNoise;
    r->FixedSeed = k->fixedSeed; r->Spp = k->spp; r->MaxDepth = k->maxDepth; r->Flags = k->flags; r->TileSize = k->tileSize; r->Rank = k->rank; r->WorldSize = k->worldSize; r->SamplesPerPass = k->samplesPerPass; r->EnableTAAU = k->enableTAAU != 0; r->AsyncSubmit = k->asyncSubmit != 0;
}
ENG_API int eng_renderer_new_communicator_id(void* id128) { return guard([&] { RTRenderer::NewCommunicatorId(id128); }); }
ENG_API int eng_renderer_init_multi_gpu(RTRenderer* r, const void* id128, int rank, int worldSize) { return guard([&] { r->InitMultiGpu(id128, rank, worldSize); }); }
ENG_API int eng_renderer_render_direct_to_pbo(RTRenderer* r, void* pbo, int w, int h, int frame, float dt) { return guard([&] { r->RenderDirectToPbo(pbo, w, h, frame, dt); }); }
ENG_API void eng_renderer_last_config(RTRenderer* r, RtRenderConfig* out) { *out = r->LastConfig(); }
ENG_API int eng_framebuffer_bind_cpu_targets(RTRenderer* r, int* color, float* depth, int* objId, size_t n) { return guard([&] { r->Frame().BindCpuTargets(color, depth, objId, n); }); }
ENG_API int eng_framebuffer_download_to_cpu(RTRenderer* r, int slot, int* color, float* depth, int* objId, size_t n) {
    return guard([&] {
        Framebuffer& f = r->Frame();
        if (color && depth && objId) { f.DownloadToCpu(slot, color, depth, objId, n); return; }
        f.DownloadToCpu(slot);
        if (f.CpuColor().size() != n) throw ArgumentOutOfRangeException("n");
        if (color) memcpy(color, f.CpuColor().data(), n * 4);
        if (depth) memcpy(depth, f.CpuDepth().data(), n * 4);
        if (objId) memcpy(objId, f.CpuObjectId().data(), n * 4);
    });
}
