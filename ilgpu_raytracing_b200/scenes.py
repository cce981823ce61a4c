"""Synthetic scenes for the BASELINE.json configs, as plain numpy descriptions.

A SceneSpec is what the reference's host code would hold BEFORE Scene.LoadObjInstance /
BuildSphereInstance run: textures, spheres, one optional triangle mesh, and the list of instances
to create.  It is consumed by the engine mirror (engine.Scene.from_spec) and, in tests, by the
oracle; both build the reference BVH2 arrays from it with their own builders.

Reference quirks honoured by every generator (SURVEY.md §8a "quirks"): one instance per sphere,
all spheres added before the first instance is built, at most one mesh per scene, distinct
centroid keys (no sort ties).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import layouts as L


@dataclass
class MeshSpec:
    positions: np.ndarray   # (n,3) f32
    tris: np.ndarray        # (m,3) i32
    texcoords: np.ndarray   # (k,2) f32
    tri_uvs: np.ndarray     # (m,3) i32
    tri_mat: np.ndarray     # (m,)  i32 indices into materials
    materials: np.ndarray   # MATERIAL[]; texture indices are GLOBAL indices into SceneSpec.textures
    object_to_world: np.ndarray = field(default_factory=L.affine_identity)


@dataclass
class SceneSpec:
    textures: list = field(default_factory=list)            # list of (h,w,4) u8 RGBA
    spheres: np.ndarray = field(default_factory=lambda: np.zeros(0, L.SPHERE))
    sphere_instances: list = field(default_factory=list)    # list of (ids, AFFINE)
    mesh: MeshSpec | None = None
    mesh_first: bool = False                                # create the mesh instance before the sphere instances


def material(kd=(1.0, 1.0, 1.0), diffuse_tex=-1, shading=L.SHADING_LAMBERT, ior=1.0, alpha_tex=-1, alpha_cutoff=0.5, two_sided=0) -> np.ndarray:
    m = np.zeros((), L.MATERIAL)
    m["Kd"] = tuple(np.float32(k) for k in kd)
    m["HasDiffuseMap"] = 1 if diffuse_tex >= 0 else 0
    m["DiffuseTexIndex"] = diffuse_tex
    m["Shading"] = shading
    m["IOR"] = ior
    m["HasAlphaMap"] = 1 if alpha_tex >= 0 else 0
    m["AlphaTexIndex"] = alpha_tex
    m["TwoSided"] = two_sided
    m["AlphaCutoff"] = alpha_cutoff
    return m


def sphere(center, radius, albedo, mat, shading=L.SHADING_LAMBERT, ior=1.0) -> np.ndarray:
    s = np.zeros((), L.SPHERE)
    s["center"] = tuple(np.float32(c) for c in center)
    s["radius"] = radius
    s["albedo"] = tuple(np.float32(c) for c in albedo)
    s["material"] = mat
    s["shading"] = shading
    s["ior"] = ior
    return s


def checker_texture(w, h, step, c0, c1) -> np.ndarray:
    """AddCheckerTexture, Engine/Scene.cs:98-109."""
    y, x = np.mgrid[0:h, 0:w]
    a = (((x // step) + (y // step)) & 1) == 0
    t = np.where(a[..., None], np.array(c0, np.uint8), np.array(c1, np.uint8)).astype(np.uint8)
    return np.ascontiguousarray(t)


def default_spheres() -> tuple[list, np.ndarray]:
    """Textures and the six spheres of Scene.BuildDefaultScene (Engine/Scene.cs:111-125)."""
    tex = [checker_texture(256, 256, 16, (255, 255, 255, 255), (20, 20, 20, 255)),
           checker_texture(256, 256, 8, (40, 40, 200, 255), (200, 200, 40, 255))]
    m_ground, m_red, m_green = material((1, 1, 1), 0), material((0.8, 0.3, 0.3)), material((0.3, 0.8, 0.3))
    m_tex, m_white = material((1, 1, 1), 1), material((1, 1, 1))
    sp = [sphere((0.0, -1000.5, 0.0), 1000.0, (1, 1, 1), m_ground),
          sphere((-0.9, 0.5, -0.2), 0.5, (0.8, 0.3, 0.3), m_red),
          sphere((0.9, 0.35, 0.2), 0.35, (0.3, 0.8, 0.3), m_green),
          sphere((0.0, 0.75, 0.6), 0.75, (1, 1, 1), m_tex),
          sphere((-1.8, 0.5, 0.8), 0.5, (1, 1, 1), m_white, L.SHADING_MIRROR, 1.0),
          sphere((1.8, 0.5, -0.8), 0.5, (1, 1, 1), m_white, L.SHADING_GLASS, 1.5)]
    return tex, np.array(sp, dtype=L.SPHERE)


def default_scene() -> SceneSpec:
    """C1: Scene.BuildDefaultScene without Sponza (no .obj is shipped)."""
    tex, sp = default_spheres()
    return SceneSpec(textures=tex, spheres=sp, sphere_instances=[([i], L.affine_identity()) for i in range(len(sp))])


class _XorShift:
    """RNG.Create(seed) + NextUInt/NextFloat of the reference (Engine/RTUtils.cs:25-49), for scene generation."""

    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFF or 1

    def next_uint(self) -> int:
        x = self.s
        x ^= (x << 13) & 0xFFFFFFFF
        x ^= x >> 17
        x ^= (x << 5) & 0xFFFFFFFF
        self.s = x or 1
        return self.s

    def next_float(self) -> float:
        return float(np.float32(self.next_uint() & 0xFFFFFF) * np.float32(1.0 / 16777216.0))


def _hash2(i, j):
    h = (np.uint32(i) * np.uint32(0x9E3779B1)) ^ (np.uint32(j) * np.uint32(0x85EBCA6B))
    h ^= h >> np.uint32(15)
    h = h * np.uint32(0x2C1B3C6D)
    h ^= h >> np.uint32(12)
    return h


def sphere_grid_scene(n=32) -> SceneSpec:
    """C2: the six default spheres plus an n x n grid of small spheres on the ground, one instance each."""
    tex, base = default_spheres()
    rng = _XorShift(12345)
    sp = list(base)
    white = material((1, 1, 1))
    with np.errstate(over="ignore"):
        for i in range(n):
            for j in range(n):
                r = 0.2 + 0.2 * rng.next_float()
                kind = int(_hash2(i, j) % np.uint32(10))
                kd = (0.15 + 0.8 * rng.next_float(), 0.15 + 0.8 * rng.next_float(), 0.15 + 0.8 * rng.next_float())
                c = (-(n - 1) / 2.0 + i + 0.013 * rng.next_float(), r - 0.5, -(n - 1) / 2.0 + j + 0.013 * rng.next_float())
                if kind <= 5:
                    sp.append(sphere(c, r, kd, material(kd)))
                elif kind <= 7:
                    sp.append(sphere(c, r, (0.95, 0.95, 0.95), white, L.SHADING_MIRROR, 1.0))
                else:
                    sp.append(sphere(c, r, (1, 1, 1), white, L.SHADING_GLASS, 1.5))
    sp = np.array(sp, dtype=L.SPHERE)
    return SceneSpec(textures=tex, spheres=sp, sphere_instances=[([i], L.affine_identity()) for i in range(len(sp))])


def _value_noise(x, z, seed):
    """Smooth lattice value noise in [0,1): hashed lattice values, smoothstep-weighted bilinear blend."""
    xi, zi = np.floor(x).astype(np.int64), np.floor(z).astype(np.int64)
    fx, fz = x - xi, z - zi
    ux, uz = fx * fx * (3 - 2 * fx), fz * fz * (3 - 2 * fz)

    def lat(a, b):
        with np.errstate(over="ignore"):
            h = _hash2((a.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32) ^ np.uint32(seed), (b.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32))
        return (h >> np.uint32(8)).astype(np.float64) / float(1 << 24)

    v00, v10, v01, v11 = lat(xi, zi), lat(xi + 1, zi), lat(xi, zi + 1), lat(xi + 1, zi + 1)
    return (v00 * (1 - ux) + v10 * ux) * (1 - uz) + (v01 * (1 - ux) + v11 * ux) * uz


def terrain_mesh(n_quads=708, seed=0x5EED, extent=50.0, amplitude=6.0, patch_materials=False) -> MeshSpec:
    """C3/C4 mesh: n x n quad height-field over [-extent, extent]^2 (2*n*n triangles), 4 octaves of value
    noise, per-vertex horizontal jitter so triangle centroids are distinct.  708 -> 1 002 528 triangles."""
    nv = n_quads + 1
    gi, gj = np.meshgrid(np.arange(nv), np.arange(nv), indexing="ij")
    cell = 2.0 * extent / n_quads
    with np.errstate(over="ignore"):
        jx = (_hash2(gi.astype(np.uint32) + np.uint32(seed), gj.astype(np.uint32) * np.uint32(3) + np.uint32(1)) >> np.uint32(8)).astype(np.float64) / float(1 << 24)
        jz = (_hash2(gi.astype(np.uint32) * np.uint32(5) + np.uint32(7), gj.astype(np.uint32) + np.uint32(seed)) >> np.uint32(8)).astype(np.float64) / float(1 << 24)
    x = -extent + gi * cell + (jx - 0.5) * 0.04
    z = -extent + gj * cell + (jz - 0.5) * 0.04
    h = np.zeros_like(x)
    amp, freq = 1.0, 1.0 / 25.0
    for o in range(4):
        h += amp * _value_noise(x * freq + 100.0 * o, z * freq - 37.0 * o, seed + o)
        amp *= 0.5
        freq *= 2.0
    y = amplitude * (h / 1.875 - 0.5) * 2.0
    pos = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([gi / n_quads, gj / n_quads], axis=-1).reshape(-1, 2).astype(np.float32)
    qi, qj = np.meshgrid(np.arange(n_quads), np.arange(n_quads), indexing="ij")
    v00 = (qi * nv + qj).reshape(-1)
    v10, v01, v11 = v00 + nv, v00 + 1, v00 + nv + 1
    # wind the triangles so the geometric normal normalize((v1-v0) x (v2-v0)) points up (+Y)
    t0 = np.stack([v00, v01, v10], axis=-1)
    t1 = np.stack([v10, v01, v11], axis=-1)
    tris = np.stack([t0, t1], axis=1).reshape(-1, 3).astype(np.int32)
    if patch_materials:
        # extension variant of C4: material per 16x16-quad patch, 60 % Lambert / 20 % mirror / 20 % glass
        mats = []
        rng = _XorShift(0xBEEF)
        for k in range(64):
            kd = (0.2 + 0.7 * rng.next_float(), 0.2 + 0.7 * rng.next_float(), 0.2 + 0.7 * rng.next_float())
            if k % 10 < 6:
                mats.append(material(kd))
            elif k % 10 < 8:
                mats.append(material((0.9, 0.9, 0.9), shading=L.SHADING_MIRROR))
            else:
                mats.append(material((1, 1, 1), shading=L.SHADING_GLASS, ior=1.5))
        with np.errstate(over="ignore"):
            pm = (_hash2((qi // 16).astype(np.uint32), (qj // 16).astype(np.uint32)) % np.uint32(64)).astype(np.int32).reshape(-1)
        tri_mat = np.repeat(pm, 2).astype(np.int32)
        materials = np.array(mats, dtype=L.MATERIAL)
    else:
        tri_mat = np.zeros(len(tris), np.int32)
        materials = np.array([material((0.7, 0.7, 0.7))], dtype=L.MATERIAL)
    return MeshSpec(pos, tris, uv, tris.copy(), tri_mat, materials)


def terrain_scene(n_quads=708, n_spheres=0, patch_materials=False, seed=0x5EED) -> SceneSpec:
    """C3 (n_spheres=0) and C4 (reference-faithful variant: n_spheres=256 mirror/glass/diffuse spheres over a Lambert
    terrain; extension variant: patch_materials=True with RT_FLAG_TRI_MATERIALS)."""
    mesh = terrain_mesh(n_quads, seed, patch_materials=patch_materials)
    sp = []
    if n_spheres:
        rng = _XorShift(0xC4C4)
        white = material((1, 1, 1))
        side = int(np.ceil(np.sqrt(n_spheres)))
        for k in range(n_spheres):
            i, j = k // side, k % side
            r = 0.8 + 1.2 * rng.next_float()
            c = (-42.0 + 84.0 * (i + 0.15 + 0.7 * rng.next_float()) / side, 5.5 + 6.0 * rng.next_float(), -42.0 + 84.0 * (j + 0.15 + 0.7 * rng.next_float()) / side)
            kind = k % 5
            if kind < 2:
                sp.append(sphere(c, r, (0.95, 0.95, 0.95), white, L.SHADING_MIRROR, 1.0))
            elif kind < 4:
                sp.append(sphere(c, r, (1, 1, 1), white, L.SHADING_GLASS, 1.5))
            else:
                kd = (0.2 + 0.7 * rng.next_float(), 0.2 + 0.7 * rng.next_float(), 0.2 + 0.7 * rng.next_float())
                sp.append(sphere(c, r, kd, material(kd)))
    spheres = np.array(sp, dtype=L.SPHERE) if sp else np.zeros(0, L.SPHERE)
    return SceneSpec(textures=[], spheres=spheres, sphere_instances=[([i], L.affine_identity()) for i in range(len(spheres))], mesh=mesh, mesh_first=True)


# ---- cameras of the configs --------------------------------------------------------------------------
CAMERAS = {
    "C1A": dict(origin=(0.0, 1.0, 3.0), look_at=(0.0, 0.5, 0.0), translate=(1.0, 0.0, -4.0), fov=60.0),   # reference default (RTRenderer.cs:78-79)
    "C1B": dict(origin=(0.0, 1.0, 3.0), look_at=(0.0, 0.5, 0.0), translate=None, fov=60.0),                # un-translated: sees all six spheres
    "C2": dict(origin=(0.0, 4.0, 12.0), look_at=(0.0, 0.0, 0.0), translate=None, fov=60.0),
    "C3": dict(origin=(0.0, 25.0, 70.0), look_at=(0.0, 0.0, 0.0), translate=None, fov=60.0),
}
