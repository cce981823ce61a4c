"""Mints the golden fixtures under tests/golden/ from the CPU oracle (python -m tests.golden.make_golden).

The reference has no golden vectors of its own and cannot be run here, so these pin the ORACLE
(a regression guard) and give the GPU tests small, committed input/output pairs that do not need the oracle's
build to be bit-stable across compilers.  Each case is tiny: the fixtures stay a few hundred KB in total.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ilgpu_raytracing_b200 import scenes  # noqa: E402
from oracle import orc  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

CASES = {
    # name: (scene kind, camera, width, height, spp, depth, flags)
    "default_c1a": ("default", "C1A", 160, 90, 1, 1, 0),
    "default_c1b_d3": ("default", "C1B", 160, 90, 2, 3, 0),
    "spheres_c2": ("spheres8", "C2", 160, 90, 4, 4, 0),
    "terrain_c4": ("terrain48", "C3", 160, 90, 2, 8, 0),
    "terrain_trimat": ("terrain48m", "C3", 160, 90, 2, 6, 1),
}


def make_spec(kind):
    if kind == "default":
        return scenes.default_scene()
    if kind == "spheres8":
        return scenes.sphere_grid_scene(8)
    if kind == "terrain48":
        return scenes.terrain_scene(48, 9)
    if kind == "terrain48m":
        return scenes.terrain_scene(48, 4, patch_materials=True)
    raise ValueError(kind)


def render_case(name):
    kind, cam, W, H, spp, depth, flags = CASES[name]
    sc = oracle_scene_from_spec(make_spec(kind))
    return orc.render(sc, oracle_camera(cam, W, H), orc.make_config(W, H, spp=spp, max_depth=depth, flags=flags))


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        r = render_case(name)
        np.savez_compressed(os.path.join(out, name + ".npz"), name=name, primId=r.primId, instId=r.instId, rgba8=r.rgba8, segCount=r.segCount,
                            termCode=r.termCode, pathHash=r.pathHash, radiance=r.radiance, depth=r.depth, objId=r.objId)
        print(name, r.counters)


if __name__ == "__main__":
    main()
