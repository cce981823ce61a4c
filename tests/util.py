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
