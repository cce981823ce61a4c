"""Shared helpers for the tests: build the oracle's scene from a SceneSpec, cameras."""
from __future__ import annotations

import numpy as np

from ilgpu_raytracing_b200 import scenes
from oracle import orc


def oracle_scene_from_spec(spec: scenes.SceneSpec, variant: str = "") -> orc.Scene:
    sc = orc.Scene(variant)
    for t in spec.textures:
        sc.add_texture(t)
    for s in spec.spheres:
        sc.add_sphere(s)

    def add_mesh():
        m = spec.mesh
        sc.add_mesh_instance(m.positions, m.tris, m.texcoords, m.tri_uvs, m.tri_mat, m.materials, m.object_to_world)

    if spec.mesh is not None and spec.mesh_first:
        add_mesh()
    for ids, xf in spec.sphere_instances:
        sc.add_sphere_instance(ids, xf)
    if spec.mesh is not None and not spec.mesh_first:
        add_mesh()
    sc.rebuild_tlas()
    return sc


def oracle_camera(name: str, width: int, height: int) -> np.ndarray:
    c = scenes.CAMERAS[name]
    cam = orc.camera_create(width, height, c["fov"], c["origin"], c["look_at"])
    if c["translate"] is not None:
        orc.camera_translate(cam, *c["translate"])
    return cam


def special_scene(mode: str = "translated") -> scenes.SceneSpec:
    """Exercises what the benchmark scenes do not: textured / alpha-masked / two-sided triangles, exactly duplicated
    triangles (equal-t ties between different primitive ids), a transformed mesh instance and transformed sphere
    instances (textured, mirror, glass).  mode: "identity" | "translated" | "scaled" (rotation + uniform scale).
    Only translations are self-consistent in the reference: tWorld = tObj / scale (SceneDeviceViews.cs:67) and an
    InvertRigidOrUniform that returns a rotation un-transposed (Scene.cs:624-631) make its own box culling depend on the
    visiting order for anything else."""
    transformed = mode != "identity"
    sc_mesh, sc_a, sc_b = (1.7, 0.75, 1.25) if mode == "scaled" else (1.0, 1.0, 1.0)
    rot = 1.0 if mode == "scaled" else 0.0
    from ilgpu_raytracing_b200 import layouts as L
    rs = np.random.RandomState(5)
    tex_checker = scenes.checker_texture(64, 64, 8, (255, 255, 255, 255), (30, 60, 200, 255))
    yy, xx = np.mgrid[0:64, 0:64]
    mask = (((xx - 32) ** 2 + (yy - 32) ** 2) < 24 ** 2).astype(np.uint8) * 255
    tex_mask = np.stack([mask, mask, mask, np.full_like(mask, 255)], axis=-1).astype(np.uint8)
    noise = rs.randint(0, 256, (32, 32, 4)).astype(np.uint8)
    n = 12
    gi, gj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    pos = np.stack([gi * 0.5 - 3.0, 0.15 * np.sin(gi * 0.9) * np.cos(gj * 0.7), gj * 0.5 - 3.0], axis=-1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([gi / 4.0, gj / 4.0], axis=-1).reshape(-1, 2).astype(np.float32)
    qi, qj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    v00 = (qi * (n + 1) + qj).reshape(-1)
    v10, v01, v11 = v00 + n + 1, v00 + 1, v00 + n + 2
    tris = np.stack([np.stack([v00, v01, v10], -1), np.stack([v10, v01, v11], -1)], axis=1).reshape(-1, 3).astype(np.int32)
    tri_mat = (np.arange(len(tris)) // 2 % 4).astype(np.int32)
    # exact duplicates of 20 triangles with a different material: same t, different primitive id -> tie-break by visiting order
    dup = tris[40:60].copy()
    tris = np.concatenate([tris, dup])
    tri_mat = np.concatenate([tri_mat, np.full(len(dup), 1, np.int32)])
    mats = np.array([scenes.material((0.7, 0.7, 0.7)),
                     scenes.material((1, 1, 1), diffuse_tex=0),
                     scenes.material((0.9, 0.5, 0.2), alpha_tex=1, alpha_cutoff=0.5, two_sided=1),
                     scenes.material((0.2, 0.8, 0.3), diffuse_tex=2, two_sided=1)], dtype=L.MATERIAL)
    xf = L.affine_trs((0.3, -0.2, 0.1), rot_y_deg=25.0 * rot, scale=sc_mesh) if transformed else L.affine_identity()
    mesh = scenes.MeshSpec(pos, tris, uv, tris.copy(), tri_mat, mats, xf)
    white = scenes.material((1, 1, 1))
    sp = np.array([scenes.sphere((0.0, 1.2, 0.0), 0.8, (1, 1, 1), scenes.material((1, 1, 1), diffuse_tex=0)),
                   scenes.sphere((-2.0, 1.0, 1.0), 0.7, (0.95, 0.95, 0.95), white, L.SHADING_MIRROR, 1.0),
                   scenes.sphere((2.0, 1.0, -1.0), 0.7, (1, 1, 1), white, L.SHADING_GLASS, 1.5),
                   scenes.sphere((0.0, -1001.0, 0.0), 1000.0, (0.6, 0.6, 0.6), scenes.material((0.6, 0.6, 0.6)))], dtype=L.SPHERE)
    inst = [([0], L.affine_trs((0.5, 0.3, -0.5), rot_y_deg=40.0 * rot, scale=sc_a) if transformed else L.affine_identity()),
            ([1], L.affine_trs((0.0, 0.2, 0.0), rot_y_deg=-15.0 * rot, scale=sc_b) if transformed else L.affine_identity()),
            ([2], L.affine_identity()), ([3], L.affine_identity())]
    return scenes.SceneSpec(textures=[tex_checker, tex_mask, noise], spheres=sp, sphere_instances=inst, mesh=mesh, mesh_first=False)


SPECIAL_CAMERA = dict(origin=(0.5, 3.5, 8.0), look_at=(0.0, 0.3, 0.0), fov=50.0)


def special_camera(width: int, height: int) -> np.ndarray:
    return orc.camera_create(width, height, SPECIAL_CAMERA["fov"], SPECIAL_CAMERA["origin"], SPECIAL_CAMERA["look_at"])
