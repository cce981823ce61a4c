"""Test infrastructure: a SECOND, independent restatement of the reference's hot path, written in scalar Python/numpy
float32 straight from the C# sources (not from oracle/rt_oracle.cpp), for scenes of spheres and one triangle mesh with
identity instances - the reference's default scene among them.  It pins the C++ oracle where the reference itself cannot (no tests, no fixtures, cannot
run here): two restatements made separately from the same lines must agree.

  Camera.CreateCamera / Translate            Engine/Camera.cs:19-47, 121-126
  Ray.GenerateRay, RNG                       Engine/RTUtils.cs:13-17, 20-138
  IntersectSphere, TraverseBLAS_Sphere       Engine/SceneDeviceViews.cs:517-537, 124-170 (hit acceptance, colour, sphere uv)
  IntersectTriangleMT_Bary, TraverseBLAS_Tri_Textured, AnyHit_Tri_Textured
                                             Engine/SceneDeviceViews.cs:540-558, 173-237, 270-327 (textures, alpha cut-out, two-sided)
  SampleTextureLinear(RGB_A), SampleMaskLinear / Point, TexelRaw
                                             Engine/SceneDeviceViews.cs:330-472
  PrimaryVisibilityKernel, PathTraceKernel   Engine/RTRay.cs:188-325, ReSTIR_Direct :438-543 with temporal / spatial reuse
                                             (:339-435, 475-516; BakeCameraDerived RTRenderer.cs:241-263), helpers :546-671
  GpuFramebuffer.Store / PackRGBA8           Engine/RTRay.cs:59-76
  TaaResolveKernel, BilinearUpsampleKernel   Engine/RTTaa.cs:117-262, Engine/RTRenderer.cs:287-346 (vectorised, end of the file)

Differences by construction: every primitive is tested (no TLAS / BLAS culling: the boxes only prune; scenes with exactly
coincident primitives, whose equal-t ties the visiting order decides, are therefore out of its reach), and the transcendental
functions are numpy's, so the comparison partner is the oracle's libm build and radiance is compared with the north-star
tolerance (1e-4 relative RMS), ids and bounce counts exactly.  Slow (a few ms per path vertex): tiny images only."""
import math

import numpy as np

f32 = np.float32
PI = f32(3.14159265358979323846)
INV_PI = f32(0.31830988618379067154)
EPS_N = f32(0.0025)
EPS_MIN = f32(1e-6)
M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------- Float3
class V:
    __slots__ = ("x", "y", "z")

    def __init__(self, x, y, z):
        self.x, self.y, self.z = f32(x), f32(y), f32(z)

    def __add__(self, o): return V(self.x + o.x, self.y + o.y, self.z + o.z)
    def __sub__(self, o): return V(self.x - o.x, self.y - o.y, self.z - o.z)
    def __neg__(self): return V(-self.x, -self.y, -self.z)

    def __mul__(self, o):
        if isinstance(o, V):
            return V(self.x * o.x, self.y * o.y, self.z * o.z)
        o = f32(o)
        return V(self.x * o, self.y * o, self.z * o)


def dot(a, b): return a.x * b.x + a.y * b.y + a.z * b.z
def cross(a, b): return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x)
def fmax(a, b): return f32(max(float(a), float(b)))
def fmin(a, b): return f32(min(float(a), float(b)))


def normalize(v):   # Float3.cs:91-95; XMath.Rsqrt = 1 / sqrt
    inv = f32(1.0) / np.sqrt(fmax(f32(1e-20), v.x * v.x + v.y * v.y + v.z * v.z))
    return V(v.x * inv, v.y * inv, v.z * inv)


# ---------------------------------------------------------------------------------------------------------------- RNG
def _rotl(v, r): return ((v << (r & 31)) | (v >> ((32 - r) & 31))) & M32


def _splitmix32(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return (x ^ (x >> 32)) & M32


def _pcg(x):
    x ^= x >> 16; x = (x * 0x7FEB352D) & M32; x ^= x >> 15; x = (x * 0x846CA68B) & M32; x ^= x >> 16
    return x


def _hash32(x):
    x ^= x >> 17; x = (x * 0xED5AD4BB) & M32; x ^= x >> 11; x = (x * 0xAC4C1B51) & M32; x ^= x >> 15; x = (x * 0x31848BAB) & M32; x ^= x >> 14
    return x


class RNG:
    def __init__(self, index, width, height, frame, sample, salt, lock_noise):   # CreateFromIndex1D -> CreateFromPixel -> Create
        px, py = index % max(1, width), index // max(1, width)
        f = 0 if lock_noise != 0 else frame & M32
        ln = lock_noise & M32
        ln0 = (_hash32(ln) ^ ((ln * 0x1B873593) & M32)) if lock_noise != 0 else 0
        ln1 = ((_rotl(ln, 7) * 0x85EBCA6B) & M32) if lock_noise != 0 else 0
        a = px ^ 0xB5297A4D
        b = ((py * 0x68E31DA4) & M32) ^ ((f * 0x9E3779B1 + 0x85EBCA6B) & M32) ^ ln0
        c = ((sample ^ 0xC2B2AE35) + _rotl(px, 16)) & M32
        d = (((salt ^ 0x27D4EB2F) + _rotl(py, 8)) & M32) ^ ln1          # C# precedence: + before ^
        s0 = _splitmix32((((a << 32) | b) ^ 0xD1B54A32D192ED03) & M64)
        s1 = _splitmix32((((c << 32) | d) ^ 0x94D049BB133111EB) & M64)
        s = _pcg(s0 ^ ((_rotl(s1, 13) + 0x9E3779B1) & M32)) | 1
        self.state = s if s != 0 else 1

    def next_float(self):
        x = self.state
        x ^= (x << 13) & M32; x ^= x >> 17; x ^= (x << 5) & M32
        self.state = x if x != 0 else 1
        return f32(self.state & 0xFFFFFF) * f32(1.0 / 16777216.0)


# ---------------------------------------------------------------------------------------------------------------- camera, rays
class Camera:
    def __init__(self, width, height, fov_degrees, translate=None, origin=(0.0, 1.0, 3.0), look_at=(0.0, 0.5, 0.0)):   # CreateCamera (+ Translate); origin / lookAt are constants there
        aspect = f32(width) / f32(max(1, height))
        theta = f32(fov_degrees) * (PI / f32(180.0))
        half_h = f32(np.tan(f32(0.5) * theta))
        half_w = aspect * half_h
        origin, look_at, up = V(*origin), V(*look_at), V(0, 1, 0)
        w = normalize(origin - look_at)
        u = normalize(cross(up, w))
        v = cross(w, u)
        self.origin = origin
        self.lower_left = origin - u * half_w - v * half_h - w
        self.horizontal = u * (f32(2.0) * half_w)
        self.vertical = v * (f32(2.0) * half_h)
        if translate is not None:
            d = V(*translate)
            self.origin = self.origin + d
            self.lower_left = self.lower_left + d
        # RTRenderer.BakeCameraDerived (RTRenderer.cs:241-263): what the kernels read of the PREVIOUS camera for reprojection
        center = self.lower_left + self.horizontal * f32(0.5) + self.vertical * f32(0.5)
        self.forward = normalize(center - self.origin)
        self.up = normalize(self.vertical)
        self.right = normalize(cross(self.forward, self.up))

        def length(v):
            return np.sqrt(v.x * v.x + v.y * v.y + v.z * v.z)
        focus, half = length(center - self.origin), f32(0.5) * length(self.vertical)
        self.fov_y = f32(2.0) * f32(np.arctan(half / focus if focus > f32(1e-6) else half))
        lh, lv = length(self.horizontal), length(self.vertical)
        self.aspect = lh / lv if (lh > f32(1e-6) and lv > f32(1e-6)) else f32(width) / f32(max(1, height))


class Ray:
    def __init__(self, o, d):
        self.o, self.d = o, d


def primary_ray(cam, index, width, height):   # GBufferParams.PrimaryRay + Ray.GenerateRay
    x, y = index % width, index // width
    u = (f32(x) + f32(0.5)) / f32(max(1, width))
    v = (f32(y) + f32(0.5)) / f32(max(1, height))
    return Ray(cam.origin, normalize(cam.lower_left + cam.horizontal * u + cam.vertical * v - cam.origin))


def ray_normal_offset(origin, n, direction):   # MakeRayWithNormalOffset
    d = normalize(direction)
    s = f32(1.0) if dot(n, d) >= 0 else f32(-1.0)
    return Ray(origin + n * (EPS_N * s), d)


# ---------------------------------------------------------------------------------------------------------------- scene
class Xf:
    """One instance transform: objectToWorld (AFFINE record or None = identity), worldToObject = Scene.InvertRigidOrUniform of it
    (Scene.cs:616-638: columns normalised and NOT transposed, divided by the mean column length), uniformScale; TransformPoint /
    TransformVector / TransformRay (SceneDeviceViews.cs:475-493: the direction is not renormalised)."""

    def __init__(self, o2w=None):
        self.identity = o2w is None
        if self.identity:
            self.scale = f32(1.0)
            return
        m = [[f32(o2w[f"m{r}{c}"]) for c in range(4)] for r in range(3)]
        self.o2w = m

        def col(c):
            return V(m[0][c], m[1][c], m[2][c])

        def length(v):
            return np.sqrt(v.x * v.x + v.y * v.y + v.z * v.z)
        self.scale = (length(col(0)) + length(col(1)) + length(col(2))) / f32(3.0)
        inv = f32(1.0) / self.scale if self.scale > 0 else f32(1.0)
        r0, r1, r2 = normalize(col(0)), normalize(col(1)), normalize(col(2))
        w = [[r0.x * inv, r1.x * inv, r2.x * inv, f32(0.0)], [r0.y * inv, r1.y * inv, r2.y * inv, f32(0.0)], [r0.z * inv, r1.z * inv, r2.z * inv, f32(0.0)]]
        it = self._vec(w, V(m[0][3], m[1][3], m[2][3])) * f32(-1.0)
        w[0][3], w[1][3], w[2][3] = it.x, it.y, it.z
        self.w2o = w

    @staticmethod
    def _vec(m, v):
        return V(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z, m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z)

    @staticmethod
    def _point(m, p):
        return V(m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3], m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3],
                 m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3])

    def ray_to_object(self, ray):
        return ray if self.identity else Ray(self._point(self.w2o, ray.o), self._vec(self.w2o, ray.d))

    def normal_to_world(self, n):   # bestNormal = Normalize(TransformVector(objectToWorld, normalObj))
        return normalize(n if self.identity else self._vec(self.o2w, n))

    def tmax_scale(self):   # "scale = inst.uniformScale > 0f ? inst.uniformScale : 1f"
        return self.scale if self.scale > 0 else f32(1.0)


class Scene:
    """spheres: SPHERE records, one instance each (sphere_xf: an AFFINE objectToWorld per sphere, None = identity);
    textures: list of (h, w, 4) u8 RGBA; mesh: a MeshSpec (positions, tris, texcoords, tri_uvs, tri_mat, materials,
    object_to_world) or None."""

    def __init__(self, spheres, textures, mesh=None, sphere_xf=None):
        self.spheres, self.textures, self.mesh = spheres, textures, mesh
        self.sphere_xf = [Xf(None if sphere_xf is None else x) for x in (sphere_xf if sphere_xf is not None else [None] * len(spheres))]
        if mesh is not None:
            self.pos = [V(*p) for p in mesh.positions]
            self.uv = [(f32(t[0]), f32(t[1])) for t in mesh.texcoords]
            o2w = mesh.object_to_world
            ident = all(abs(float(o2w[f"m{r}{c}"]) - (1.0 if r == c else 0.0)) == 0.0 for r in range(3) for c in range(4))
            self.mesh_xf = Xf(None if ident else o2w)

    def texel(self, tex, x, y):   # TexelRaw: clamp
        h, w, _ = tex.shape
        return tex[max(0, min(h - 1, y)), max(0, min(w - 1, x))]

    def sample_rgb(self, tex, u, v):   # SampleTextureLinearRGB_A (colour part)
        h, w, _ = tex.shape
        fu = u - np.floor(u)
        fv = f32(1.0) - (v - np.floor(v))
        x, y = fu * f32(w - 1), fv * f32(h - 1)
        x0, y0 = int(np.floor(x)), int(np.floor(y))
        x1, y1 = min(w - 1, x0 + 1), min(h - 1, y0 + 1)
        tx, ty = x - f32(x0), y - f32(y0)
        k = f32(1.0) / f32(255.0)

        def c(px):
            return V(f32(px[0]) * k, f32(px[1]) * k, f32(px[2]) * k)
        c00, c10, c01, c11 = c(self.texel(tex, x0, y0)), c(self.texel(tex, x1, y0)), c(self.texel(tex, x0, y1)), c(self.texel(tex, x1, y1))
        cx0 = c00 * (f32(1.0) - tx) + c10 * tx
        cx1 = c01 * (f32(1.0) - tx) + c11 * tx
        return cx0 * (f32(1.0) - ty) + cx1 * ty

    def _texel_luma(self, tex, x, y):   # Luma01(TexelRaw)
        p = self.texel(tex, x, y)
        k = f32(1.0) / f32(255.0)
        return f32(0.2126) * (f32(p[0]) * k) + f32(0.7152) * (f32(p[1]) * k) + f32(0.0722) * (f32(p[2]) * k)

    def sample_mask_linear(self, tex, u, v):   # SampleMaskLinear
        h, w, _ = tex.shape
        fu = u - np.floor(u)
        fv = f32(1.0) - (v - np.floor(v))
        x, y = fu * f32(w - 1), fv * f32(h - 1)
        x0, y0 = int(np.floor(x)), int(np.floor(y))
        x1, y1 = min(w - 1, x0 + 1), min(h - 1, y0 + 1)
        tx, ty = x - f32(x0), y - f32(y0)
        a00, a10, a01, a11 = self._texel_luma(tex, x0, y0), self._texel_luma(tex, x1, y0), self._texel_luma(tex, x0, y1), self._texel_luma(tex, x1, y1)
        ax0 = a00 * (f32(1.0) - tx) + a10 * tx
        ax1 = a01 * (f32(1.0) - tx) + a11 * tx
        return ax0 * (f32(1.0) - ty) + ax1 * ty

    def sample_mask_point(self, tex, u, v):   # SampleMaskPoint; XMath.Round = half to even
        h, w, _ = tex.shape
        fu = u - np.floor(u)
        fv = f32(1.0) - (v - np.floor(v))
        return self._texel_luma(tex, int(np.rint(fu * f32(w - 1))), int(np.rint(fv * f32(h - 1))))

    def _tex(self, has, idx):
        return self.textures[int(idx)] if (has != 0 and 0 <= idx < len(self.textures)) else None

    def intersect_tri(self, ray, k):   # IntersectTriangleMT_Bary -> (t, n, bu, bv)
        i0, i1, i2 = (int(v) for v in self.mesh.tris[k])
        v0, v1, v2 = self.pos[i0], self.pos[i1], self.pos[i2]
        e1, e2 = v1 - v0, v2 - v0
        p = cross(ray.d, e2)
        det = dot(e1, p)
        if abs(det) < f32(1e-8):
            return None
        inv = f32(1.0) / det
        tv = ray.o - v0
        bu = dot(tv, p) * inv
        if bu < 0 or bu > 1:
            return None
        q = cross(tv, e1)
        bv = dot(ray.d, q) * inv
        if bv < 0 or bu + bv > 1:
            return None
        t = dot(e2, q) * inv
        if t <= 0:
            return None
        return t, normalize(cross(e1, e2)), bu, bv

    def _tri_uv(self, k, bu, bv):
        t0, t1, t2 = (self.uv[int(i)] for i in self.mesh.tri_uvs[k])
        w = f32(1.0) - bu - bv
        return t0[0] * w + t1[0] * bu + t2[0] * bv, t0[1] * w + t1[1] * bu + t2[1] * bv

    def closest_tri(self, ray):
        """TraverseBLAS_Tri_Textured with every box taken: (t, nObj, albedo, tri id) or None."""
        best, closest = None, f32(1e30)
        for k in range(len(self.mesh.tris)):
            hit = self.intersect_tri(ray, k)
            if hit is None:
                continue
            t, nn, bu, bv = hit
            mat = self.mesh.materials[int(self.mesh.tri_mat[k])]
            if not (t > f32(0.001) and t < closest):
                continue
            uu, vv = self._tri_uv(k, bu, bv)
            kd = V(*[mat["Kd"][c] for c in "XYZ"])
            tex = self._tex(mat["HasDiffuseMap"], mat["DiffuseTexIndex"])
            if tex is not None:
                kd = self.sample_rgb(tex, uu, vv)          # SampleTextureLinear: the same bilinear RGB fetch
            alpha = f32(1.0)
            atex = self._tex(mat["HasAlphaMap"], mat["AlphaTexIndex"])
            if atex is not None:
                alpha = self.sample_mask_linear(atex, uu, vv)
            if alpha < mat["AlphaCutoff"]:
                continue
            closest = t
            if mat["TwoSided"] != 0 and dot(nn, ray.d) > 0:
                nn = nn * f32(-1.0)
            best = (t, nn, kd, k)
        return best

    def any_tri(self, ray, t_max):   # AnyHit_Tri_Textured
        for k in range(len(self.mesh.tris)):
            hit = self.intersect_tri(ray, k)
            if hit is None:
                continue
            t, _, bu, bv = hit
            if t <= f32(0.001) or t >= t_max:
                continue
            mat = self.mesh.materials[int(self.mesh.tri_mat[k])]
            atex = self._tex(mat["HasAlphaMap"], mat["AlphaTexIndex"])
            if atex is not None:
                uu, vv = self._tri_uv(k, bu, bv)
                a_point, cutoff, band = self.sample_mask_point(atex, uu, vv), mat["AlphaCutoff"], f32(0.10)
                if a_point < cutoff - band:
                    continue
                if a_point >= cutoff + band:
                    return True
                if self.sample_mask_linear(atex, uu, vv) < cutoff:
                    continue
            return True
        return False

    @staticmethod
    def intersect_sphere(ray, s):   # IntersectSphere
        c = V(*[s["center"][k] for k in "XYZ"])
        oc = ray.o - c
        a = dot(ray.d, ray.d)
        b = f32(2.0) * dot(oc, ray.d)
        cc = dot(oc, oc) - s["radius"] * s["radius"]
        disc = b * b - f32(4.0) * a * cc
        if disc < 0:
            return None
        sq = np.sqrt(disc)
        t = (-b - sq) / (f32(2.0) * a)
        if t < f32(0.001):
            t = (-b + sq) / (f32(2.0) * a)
            if t < f32(0.001):
                return None
        p = ray.o + ray.d * t
        return t, normalize(p - c)

    def trace_closest(self, ray):
        """TraceClosest over identity sphere instances: (t, normal, albedo, shade, ior, sphere id) or None."""
        best = None
        closest = f32(1e30)
        for i, s in enumerate(self.spheres):
            xf = self.sphere_xf[i]
            hit = self.intersect_sphere(xf.ray_to_object(ray), s)
            if hit is None:
                continue
            t, nn = hit
            if not (t > f32(0.001) and t < f32(1e30)):     # TraverseBLAS_Sphere: one sphere per instance, tClosest starts at 1e30
                continue
            t = t / xf.tmax_scale()                        # tWorld = tObj / scale
            if not (t < closest):
                continue
            m = s["material"]
            kd = V(*[m["Kd"][k] for k in "XYZ"])
            col = V(*[s["albedo"][k] for k in "XYZ"]) if (kd.x == 0 and kd.y == 0 and kd.z == 0) else kd
            if m["HasDiffuseMap"] != 0 and 0 <= m["DiffuseTexIndex"] < len(self.textures):
                u = f32(0.5) + f32(np.arctan2(nn.z, nn.x)) / (f32(2.0) * PI)
                v = f32(np.arccos(fmin(f32(1.0), fmax(f32(-1.0), nn.y)))) / PI
                col = self.sample_rgb(self.textures[int(m["DiffuseTexIndex"])], u, v)
            closest = t
            best = (t, xf.normal_to_world(nn), col, int(s["shading"]), s["ior"] if s["ior"] > 0 else f32(1.0), i)
        if self.mesh is not None:   # the mesh instance: triangles are always Lambert, ior 1 (TraceClosest :58-62)
            th = self.closest_tri(self.mesh_xf.ray_to_object(ray))
            if th is not None and th[0] / self.mesh_xf.tmax_scale() < closest:
                best = (th[0] / self.mesh_xf.tmax_scale(), self.mesh_xf.normal_to_world(th[1]), th[2], 0, f32(1.0), -1 - th[3])   # ids < 0: triangle -1 - id
        return best

    def occluded(self, ray, t_max):   # ShadowOcclusion / AnyHit_Sphere: tMaxObj = tMaxWorld * scale
        for i, s in enumerate(self.spheres):
            xf = self.sphere_xf[i]
            hit = self.intersect_sphere(xf.ray_to_object(ray), s)
            if hit is not None and hit[0] > f32(0.001) and hit[0] < t_max * xf.tmax_scale():
                return True
        return self.mesh is not None and self.any_tri(self.mesh_xf.ray_to_object(ray), t_max * self.mesh_xf.tmax_scale())


# ---------------------------------------------------------------------------------------------------------------- integrator
class Env:
    def __init__(self, sun_dir):
        self.dir_light_dir = V(*sun_dir)
        self.dir_light_radiance = V(10, 10, 10)
        self.sky_top, self.sky_bottom = V(0.5, 0.7, 1.0), V(1.0, 1.0, 1.0)

    def sky(self, d):   # SkyWeighted
        t = f32(0.5) * (d.y + f32(1.0))
        return self.sky_bottom * (f32(1.0) - t) + self.sky_top * t


def luminance(c): return f32(0.2126) * c.x + f32(0.7152) * c.y + f32(0.0722) * c.z
def cos_pdf(n, wi): return fmax(f32(0.0), dot(n, wi)) * INV_PI


def sample_hemisphere_cosine(n, rng):
    r1, r2 = rng.next_float(), rng.next_float()
    phi = f32(2.0) * PI * r1
    cos_t, sin_t = np.sqrt(f32(1.0) - r2), np.sqrt(r2)
    x, y, z = f32(np.cos(phi)) * sin_t, f32(np.sin(phi)) * sin_t, cos_t
    up = V(0, 1, 0) if abs(n.y) < f32(0.999) else V(1, 0, 0)
    t = normalize(cross(up, n))
    b = cross(n, t)
    return normalize(t * x + b * y + n * z)


def _hash(x):   # RTRay.Hash
    x &= M32
    x ^= x >> 17; x = (x * 0xED5AD4BB) & M32; x ^= x >> 11; x = (x * 0xAC4C1B51) & M32; x ^= x >> 15; x = (x * 0x31848BAB) & M32; x ^= x >> 14
    return x


class Reuse:
    """What ReSTIR_Direct reads when a reuse flag is set: the previous camera, the previous frame's reservoirs, the CURRENT
    frame's G-buffer (SpatialCompatible), the frame number."""

    def __init__(self, temporal, spatial, prev_cam, res_prev, gb, cam, width, height, frame):
        self.temporal, self.spatial, self.prev_cam, self.res_prev, self.gb, self.cam = temporal, spatial, prev_cam, res_prev, gb, cam
        self.width, self.height, self.frame = width, height, frame

    def reproject(self, pos):   # ReprojectToPrevPixel
        c = self.prev_cam
        p = pos - c.origin
        x, y, z = dot(p, c.right), dot(p, c.up), dot(p, c.forward)
        if z <= f32(1e-4):
            return -1
        tan_half = f32(np.tan(f32(0.5) * c.fov_y))
        ndc_x, ndc_y = x / (z * tan_half * c.aspect), y / (z * tan_half)
        px, py = int(f32(0.5) * (ndc_x + f32(1.0)) * f32(self.width)), int(f32(0.5) * (ndc_y + f32(1.0)) * f32(self.height))
        if not (0 <= px < self.width and 0 <= py < self.height):
            return -1
        return py * self.width + px

    def distance(self, idx):   # DistanceFromCamera(gb.worldPos[idx])
        d = self.gb["pos"][idx] - self.cam.origin
        return np.sqrt(d.x * d.x + d.y * d.y + d.z * d.z)

    def compatible(self, a, b, n_a):   # SpatialCompatible
        if self.gb["obj"][a] == self.gb["obj"][b]:
            return True
        if dot(n_a, normalize(self.gb["nrm"][b])) < f32(0.85):
            return False
        z_a, z_b = self.distance(a), self.distance(b)
        return abs(z_a - z_b) / fmax(f32(1e-3), z_a) < f32(0.05)


def restir_direct(scene, env, pos, n, albedo, rng, counters, reuse=None, index=0, out_res=None):
    mix_local, mix_delta = f32(8.0) / f32(9.0), f32(1.0) / f32(9.0)
    r = dict(L=V(0, 0, 0), wi=V(0, 0, 0), pdf=f32(0), w=f32(0), wsum=f32(0), m=0, light=0)

    def update(wi, pdf_sel, li, s, light):   # ReservoirUpdate
        new_sum = r["wsum"] + s
        accept = s / new_sum if new_sum > 0 else f32(0.0)
        if rng.next_float() < accept:
            r.update(wi=wi, pdf=pdf_sel, L=li, w=s, light=light)
        r["wsum"] = new_sum
        r["m"] += 1

    for _ in range(8):
        wi = sample_hemisphere_cosine(n, rng)
        nl = fmax(f32(0.0), dot(n, wi))
        pdf_sel = fmax(EPS_MIN, fmax(EPS_MIN, cos_pdf(n, wi)) * mix_local)
        li = env.sky(wi)
        update(wi, pdf_sel, li, luminance(albedo * li * ((nl / pdf_sel) * INV_PI)), 1)
    wi = normalize(env.dir_light_dir)
    nl = fmax(f32(0.0), dot(n, wi))
    pdf_sel = fmax(EPS_MIN, mix_delta)
    update(wi, pdf_sel, env.dir_light_radiance, luminance(albedo * env.dir_light_radiance * ((nl / pdf_sel) * INV_PI)), 2)

    def import_prev(prev_idx):   # ImportFromPrevReservoir
        if prev_idx < 0 or len(reuse.res_prev) <= prev_idx or not reuse.compatible(index, prev_idx, n):
            return
        pr = reuse.res_prev[prev_idx]
        if not (pr["m"] > 0 and pr["w"] > 0 and pr["wSum"] > 0):
            return
        wi = V(*[pr["wi"][k] for k in "XYZ"])
        lid = 2 if pr["lightId"] == 2 else 1
        li = env.dir_light_radiance if lid == 2 else env.sky(wi)
        nl = fmax(f32(0.0), dot(n, wi))
        pdf_here = fmax(EPS_MIN, mix_delta) if lid == 2 else fmax(EPS_MIN, cos_pdf(n, wi) * mix_local)
        s_here = luminance(albedo * li * ((nl / pdf_here) * INV_PI))
        w_src = pr["wSum"] / (f32(max(1, int(pr["m"]))) * fmax(EPS_MIN, pr["w"]))
        update(wi, pdf_here, li, s_here * w_src, lid)

    if reuse is not None and reuse.temporal:
        prev_idx = reuse.reproject(pos)
        if prev_idx >= 0:
            import_prev(prev_idx)
    if reuse is not None and reuse.spatial:
        h = _hash(index ^ _hash((reuse.frame & M32) ^ _hash(0xB31F5AB1)))   # Hash3(index, frame, 0xB31F5AB1)
        rot, r_ = h & 3, 1 + ((h >> 2) & 1)
        x0, y0 = index % reuse.width, index // reuse.width

        def rx(x, y): return x if rot == 0 else (-y if rot == 1 else (-x if rot == 2 else y))
        def ry(x, y): return y if rot == 0 else (x if rot == 1 else (-y if rot == 2 else -x))
        for (ox, oy) in ((-r_, 0), (r_, 0), (0, -r_), (0, r_), (-r_, -r_), (r_, -r_), (-r_, r_), (r_, r_)):   # Neighbor8
            nx, ny = x0 + rx(ox, oy), y0 + ry(ox, oy)
            import_prev(ny * reuse.width + nx if (0 <= nx < reuse.width and 0 <= ny < reuse.height) else -1)
    if out_res is not None:   # "outRes = r" (the reservoir before the visibility test)
        out_res.update(r)

    contrib = V(0, 0, 0)
    if r["m"] > 0 and r["wsum"] > 0 and r["w"] > 0:
        wi = r["wi"]
        nl = fmax(f32(0.0), dot(n, wi))
        if nl > 0 and dot(n, wi) > 0:   # Visible(): nl <= 0 -> false, else a shadow ray
            counters["shadow"] += 1
            if not scene.occluded(ray_normal_offset(pos, n, wi), f32(1e29)):
                pdf_sel = fmax(EPS_MIN, mix_delta) if r["light"] == 2 else fmax(EPS_MIN, cos_pdf(n, wi) * mix_local)
                li = env.dir_light_radiance if r["light"] == 2 else env.sky(wi)
                f_over_p = albedo * li * ((nl / pdf_sel) * INV_PI)
                w = r["wsum"] / f32(max(1, r["m"])) / fmax(EPS_MIN, r["w"])
                contrib = f_over_p * w
    return contrib


def safe_color(c):
    def one(v):
        v = v if math.isfinite(float(v)) else f32(0.0)
        return fmin(f32(1e6), fmax(f32(-1e6), v))
    return V(one(c.x), one(c.y), one(c.z))


def to_byte(x): return int(f32(255.99) * fmin(f32(1.0), fmax(f32(0.0), x)))


def render(scene, cam, width, height, spp, max_depth, sun_dir, frame=0, lock_noise=1, temporal=0, spatial=0, prev_cam=None, res_prev=None, res_cur=None):
    """PrimaryVisibilityKernel + PathTraceKernel (reuse off).  Returns dict(rgba8, depth, sphere = primary hit (sphere id, or
    -1 - triangle id, or -1 for the sky), radiance, seg = TraceNext calls per sample, counters)."""
    env = Env(sun_dir)
    n_px = width * height
    out = dict(rgba8=np.zeros(n_px, np.int32), depth=np.zeros(n_px, np.float32), sphere=np.full(n_px, -1, np.int32), hit=np.zeros(n_px, bool),
               radiance=np.zeros((n_px, 3), np.float32), seg=np.zeros((spp, n_px), np.uint8))
    counters = dict(bounce=0, shadow=0)
    with np.errstate(all="ignore"):
        # PrimaryVisibilityKernel over the whole image first: the reuse tests read the G-buffer of OTHER pixels
        primary = [scene.trace_closest(primary_ray(cam, i, width, height)) for i in range(n_px)]
        gb = dict(pos=[], nrm=[], obj=[])
        for i, h0 in enumerate(primary):
            r0 = primary_ray(cam, i, width, height)
            gb["pos"].append(r0.o + r0.d * (f32(1e6) if h0 is None else h0[0]))
            gb["nrm"].append(V(0, 1, 0) if h0 is None else h0[1])                      # StoreMiss: normal (0, 1, 0), objId -1
            gb["obj"].append(-1 if (h0 is None or h0[5] >= 0) else -1 - h0[5])        # spheres report objId -1, triangles their id
        reuse = Reuse(temporal, spatial, prev_cam if prev_cam is not None else cam, res_prev, gb, cam, width, height, frame) if (temporal or spatial) else None
        for index in range(n_px):
            ray = primary_ray(cam, index, width, height)
            hit = primary[index]
            l_frame = V(0, 0, 0)
            if hit is None:
                gpos = ray.o + ray.d * f32(1e6)        # GpuGBuffer.StoreMiss
            else:
                t, gn, galb, gshade, gior, sid = hit
                gpos = ray.o + ray.d * t
                out["sphere"][index] = sid
                out["hit"][index] = True
                packed_ior = int(fmax(f32(0.0), fmin(f32(65535.0), gior * f32(1000.0)))) & 0xFFFF   # FloatToI16
            for s in range(max(1, spp)):
                rng = RNG(index, width, height, frame, s, 0xC0FFEE, lock_noise)
                if hit is None:
                    l_frame = l_frame + safe_color(env.sky(primary_ray(cam, index, width, height).d))
                    continue
                pos, nrm, alb, shade, ior = gpos, normalize(gn), galb, gshade, f32(packed_ior) / f32(1000.0)
                li, thr = V(0, 0, 0), V(1, 1, 1)
                I = normalize(pos - cam.origin)
                seg = 0
                wrote_reservoir = False
                for depth in range(max_depth):
                    if shade == 1:      # mirror
                        nxt = ray_normal_offset(pos, nrm, I - nrm * (f32(2.0) * dot(I, nrm)))
                        thr = thr * alb
                    elif shade == 2:    # glass
                        n_use = nrm
                        outside = dot(I, nrm) < 0
                        if not outside:
                            n_use = n_use * f32(-1.0)
                        g = ior if ior > 0 else f32(1.5)
                        eta_i, eta_t = (f32(1.0), g) if outside else (g, f32(1.0))
                        dir_r = I - n_use * (f32(2.0) * dot(I, n_use))
                        eta = eta_i / eta_t
                        cos_i = -dot(I, n_use)
                        k = f32(1.0) - eta * eta * (f32(1.0) - cos_i * cos_i)
                        refr_ok = not (k < 0)
                        dir_t = normalize(I * eta + n_use * (eta * cos_i - np.sqrt(k))) if refr_ok else V(0, 0, 0)
                        c = abs(dot(I, n_use))
                        r0 = (eta_i - eta_t) / (eta_i + eta_t)
                        r0 = r0 * r0
                        om = f32(1.0) - c
                        om2 = om * om
                        fr = r0 + (f32(1.0) - r0) * (om2 * om2 * om)
                        xi = rng.next_float()
                        nxt = ray_normal_offset(pos, n_use, dir_r) if (not refr_ok or xi < fr) else ray_normal_offset(pos, -n_use, dir_t)
                        if refr_ok and xi >= fr:
                            tint = V(1, 1, 1) if (alb.x == 0 and alb.y == 0 and alb.z == 0) else alb
                            thr = thr * tint * ((eta_i * eta_i) / (eta_t * eta_t))
                    else:               # Lambert: ReSTIR-DI + cosine bounce
                        first = not wrote_reservoir and res_cur is not None      # only the first Lambert vertex of a sample imports and publishes
                        res = {} if first else None
                        li = li + thr * restir_direct(scene, env, pos, nrm, alb, rng, counters, reuse if first else None, index, res)
                        if first:
                            rc = res_cur[index]
                            rc["L"], rc["wi"] = (res["L"].x, res["L"].y, res["L"].z), (res["wi"].x, res["wi"].y, res["wi"].z)
                            rc["pdf"], rc["w"], rc["wSum"], rc["m"], rc["lightId"] = res["pdf"], res["w"], res["wsum"], res["m"], res["light"]
                            wrote_reservoir = True
                        wi = sample_hemisphere_cosine(nrm, rng)
                        nxt = ray_normal_offset(pos, nrm, wi)
                        thr = thr * alb
                        if depth >= 3:
                            mc = fmax(thr.x, fmax(thr.y, thr.z))
                            mc = fmax(fmin(mc, f32(0.98)), f32(0.05))
                            if rng.next_float() > mc:
                                thr = V(0, 0, 0)
                                break
                            thr = thr * (f32(1.0) / mc)
                    counters["bounce"] += 1
                    seg += 1
                    h2 = scene.trace_closest(nxt)     # TraceNext
                    if h2 is None:
                        li = li + thr * env.sky(nxt.d)
                        break
                    pos, nrm, alb, shade, ior = nxt.o + nxt.d * h2[0], normalize(h2[1]), h2[2], h2[3], h2[4]
                    I = nxt.d
                out["seg"][s, index] = seg
                l_frame = l_frame + safe_color(li)
            l_out = l_frame * (f32(1.0) / f32(max(1, spp)))
            out["radiance"][index] = (l_out.x, l_out.y, l_out.z)
            dc = gpos - cam.origin
            out["depth"][index] = np.sqrt(dc.x * dc.x + dc.y * dc.y + dc.z * dc.z)
            rgba = (255 << 24) | (to_byte(l_out.x) << 16) | (to_byte(l_out.y) << 8) | to_byte(l_out.z)
            out["rgba8"][index] = rgba - (1 << 32) if rgba >= (1 << 31) else rgba
    out["counters"] = counters
    return out


# ---------------------------------------------------------------------------------------------------------------- present chain
# RTTaa.TaaResolveKernel + helpers (Engine/RTTaa.cs:117-262) and BilinearUpsampleKernel (Engine/RTRenderer.cs:287-346), vectorised
# over the output image with numpy float32 arrays (every operation still rounds to binary32 like the scalar C#).
def _unpack_srgb(rgba):
    out = []
    for sh in (16, 8, 0):
        c = ((rgba >> sh) & 255).astype(np.float32) / f32(255.0)
        lin = np.power((c + f32(0.055)) / f32(1.055), f32(2.4), dtype=np.float32)
        out.append(np.where(c <= f32(0.04045), c / f32(12.92), lin).astype(np.float32))
    return out


def _pack_srgb(c3):
    chans = []
    for c in c3:
        l = np.maximum(f32(0.0), np.minimum(f32(1.0), c))
        s = np.where(l <= f32(0.0031308), f32(12.92) * l, f32(1.055) * np.power(l, f32(1.0) / f32(2.4), dtype=np.float32) - f32(0.055)).astype(np.float32)
        chans.append(np.rint(np.maximum(f32(0.0), np.minimum(f32(1.0), s)) * f32(255.0)).astype(np.int64))   # XMath.Round: half to even
    v = (255 << 24) | (chans[0] << 16) | (chans[1] << 8) | chans[2]
    return np.where(v >= (1 << 31), v - (1 << 32), v).astype(np.int32)


def _catrom(a, b, t):
    tt = t * (f32(2.0) - t)
    return [x * (f32(1.0) - tt) + y * tt for x, y in zip(a, b)]


def _sample_catrom_srgb(img, w, h, x, y):
    x1 = np.clip(np.floor(x).astype(np.int64), 0, w - 1)
    y1 = np.clip(np.floor(y).astype(np.int64), 0, h - 1)
    fx, fy = x - x1.astype(np.float32), y - y1.astype(np.float32)
    x2, y2 = np.minimum(x1 + 1, w - 1), np.minimum(y1 + 1, h - 1)
    c00, c10, c01, c11 = (_unpack_srgb(img[j * w + i]) for j, i in ((y1, x1), (y1, x2), (y2, x1), (y2, x2)))
    return _catrom(_catrom(c00, c10, fx), _catrom(c01, c11, fx), fy)


class Taa:
    """RTTaa: history colour / object id, _historyValid; resolve() = ResolveUpsample + TaaResolveKernel with the reference's tunables."""

    def __init__(self, out_w, out_h):
        self.w, self.h, self.valid = out_w, out_h, False
        self.hist_color, self.hist_obj = np.zeros(out_w * out_h, np.int32), np.zeros(out_w * out_h, np.int32)

    def resolve(self, low_color, low_obj, in_w, in_h, feedback=0.075, sharpness=0.10):
        low_color, low_obj = np.asarray(low_color, np.int32).astype(np.int64), np.asarray(low_obj, np.int32)
        feedback, sharpness = f32(feedback), f32(sharpness)
        idx = np.arange(self.w * self.h)
        px, py = (idx % self.w).astype(np.float32), (idx // self.w).astype(np.float32)
        sx = (px + f32(0.5)) * (f32(in_w) / f32(self.w)) - f32(0.5)
        sy = (py + f32(0.5)) * (f32(in_h) / f32(self.h)) - f32(0.5)
        with np.errstate(all="ignore"):
            cur = _sample_catrom_srgb(low_color, in_w, in_h, sx, sy)
            nmin, nmax = list(cur), list(cur)
            for oy in (-1, 0, 1):
                for ox in (-1, 0, 1):
                    if ox == 0 and oy == 0:
                        continue
                    c = _sample_catrom_srgb(low_color, in_w, in_h, sx + f32(ox) * f32(0.5), sy + f32(oy) * f32(0.5))
                    nmin = [np.minimum(a, b) for a, b in zip(nmin, c)]
                    nmax = [np.maximum(a, b) for a, b in zip(nmax, c)]
            ix = np.clip(np.rint(sx).astype(np.int64), 0, in_w - 1)
            iy = np.clip(np.rint(sy).astype(np.int64), 0, in_h - 1)
            obj = low_obj[iy * in_w + ix]
            hist = _unpack_srgb(self.hist_color.astype(np.int64))
            reset = (not self.valid) | (self.hist_obj != obj)
            clamped = [np.minimum(hi, np.maximum(lo, v)) for v, lo, hi in zip(hist, nmin, nmax)]   # Clamp(): k * 0.0f
            a = np.where(reset, f32(1.0), feedback).astype(np.float32)
            accum = [h * (f32(1.0) - a) + c * a for h, c in zip(clamped, cur)]
            sharp = [x * (f32(1.0) + f32(2.0) * sharpness) - (lo + hi) * (f32(0.5) * sharpness) for x, lo, hi in zip(accum, nmin, nmax)]
            accum = [x * (f32(1.0) - sharpness) + s * sharpness for x, s in zip(accum, sharp)]
            out = _pack_srgb(accum)
        self.hist_color, self.hist_obj, self.valid = out.copy(), obj.astype(np.int32), True
        return out


def bilinear_upsample(src, src_w, src_h, dst_w, dst_h):   # BilinearUpsampleKernel
    src = np.asarray(src, np.int32).astype(np.int64)
    idx = np.arange(dst_w * dst_h)
    x, y = (idx % dst_w).astype(np.float32), (idx // dst_w).astype(np.float32)
    u = ((x + f32(0.5)) * f32(src_w) / f32(dst_w)) - f32(0.5)
    v = ((y + f32(0.5)) * f32(src_h) / f32(dst_h)) - f32(0.5)
    x0, y0 = np.clip(np.floor(u).astype(np.int64), 0, src_w - 1), np.clip(np.floor(v).astype(np.int64), 0, src_h - 1)
    x1, y1 = np.clip(x0 + 1, 0, src_w - 1), np.clip(y0 + 1, 0, src_h - 1)
    tx = np.clip(u - x0.astype(np.float32), f32(0.0), f32(1.0))
    ty = np.clip(v - y0.astype(np.float32), f32(0.0), f32(1.0))

    def unpack(p):
        return [((p >> sh) & 255).astype(np.float32) * (f32(1.0) / f32(255.0)) for sh in (16, 8, 0)]
    c00, c10, c01, c11 = unpack(src[y0 * src_w + x0]), unpack(src[y0 * src_w + x1]), unpack(src[y1 * src_w + x0]), unpack(src[y1 * src_w + x1])
    chans = []
    for k in range(3):
        cx0 = c00[k] * (f32(1.0) - tx) + c10[k] * tx
        cx1 = c01[k] * (f32(1.0) - tx) + c11[k] * tx
        c = cx0 * (f32(1.0) - ty) + cx1 * ty
        chans.append((f32(255.99) * np.minimum(f32(1.0), np.maximum(f32(0.0), c))).astype(np.int64))
    val = (255 << 24) | (chans[0] << 16) | (chans[1] << 8) | chans[2]
    return np.where(val >= (1 << 31), val - (1 << 32), val).astype(np.int32)
