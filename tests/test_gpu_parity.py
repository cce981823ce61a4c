"""-m gpu: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from ilgpu_raytracing_b200 import layouts as L
from ilgpu_raytracing_b200 import scenes
from oracle import orc
from tests.parity import assert_parity, download_all
from tests.util import oracle_camera, oracle_scene_from_spec

pytestmark = pytest.mark.gpu


def _run(ctx, sc, cam, W, H, spp, depth, flags=0, box=None, spp_pass=0, label=""):
    ocfg = orc.make_config(W, H, spp=spp, max_depth=depth, flags=flags & 1, crop=box)
    r = orc.render(sc, cam, ocfg)
    cfg = L.make_render_config(W, H, spp=spp, max_depth=depth, flags=flags | L.RT_FLAG_PATH_AOVS, samples_per_pass=spp_pass)
    ctx.render(cam, cfg)
    ctx.sync()
    prod = download_all(ctx)
    assert_parity(r, prod, W, H, box=box, spp=spp, label=label)
    st = ctx.stats()
    if box is None:
        assert st["raysPrimary"] == r.counters["raysPrimary"]
        assert st["raysBounce"] == r.counters["raysBounce"], label
        assert st["raysShadow"] == r.counters["raysShadow"], label
    return r, prod, st


def test_default_scene_c1(gpu_ctx):
    """C1: Scene.BuildDefaultScene, 1280x720, 1 spp, MaxDepth 1, both cameras."""
    sc = orc.Scene()
    sc.build_default()
    gpu_ctx.scene_upload(sc.arrays())
    for camname in ("C1A", "C1B"):
        cam = oracle_camera(camname, 1280, 720)
        _run(gpu_ctx, sc, cam, 1280, 720, 1, 1, label=camname)


@pytest.mark.parametrize("spp,depth,spp_pass", [(4, 4, 0), (5, 3, 2), (2, 8, 1)])
def test_default_scene_multibounce(gpu_ctx, spp, depth, spp_pass):
    sc = orc.Scene()
    sc.build_default()
    gpu_ctx.scene_upload(sc.arrays())
    cam = oracle_camera("C1B", 480, 270)
    _run(gpu_ctx, sc, cam, 480, 270, spp, depth, spp_pass=spp_pass, label=f"default {spp}spp d{depth}")


def test_sphere_grid_c2(gpu_ctx):
    """C2 at reduced size: sphere-only scene, 16 spp, MaxDepth 4 (Russian roulette fires at depth 3)."""
    spec = scenes.sphere_grid_scene(32)
    sc = oracle_scene_from_spec(spec)
    gpu_ctx.scene_upload(sc.arrays())
    cam = oracle_camera("C2", 480, 270)
    r, prod, st = _run(gpu_ctx, sc, cam, 480, 270, 16, 4, label="C2")
    assert (r.termCode == 3).sum() > 0   # roulette exercised


@pytest.mark.parametrize("flags", [0, L.RT_FLAG_TRI_MATERIALS])
def test_terrain_mesh(gpu_ctx, flags):
    """C3/C4 mesh generator at 2*160^2 = 51 200 triangles + 36 spheres, primary only and 8 bounces."""
    spec = scenes.terrain_scene(n_quads=160, n_spheres=36, patch_materials=bool(flags))
    sc = oracle_scene_from_spec(spec)
    gpu_ctx.scene_upload(sc.arrays())
    cam = oracle_camera("C3", 512, 288)
    _run(gpu_ctx, sc, cam, 512, 288, 1, 0, flags=flags, label="terrain primary")
    _run(gpu_ctx, sc, cam, 512, 288, 3, 8, flags=flags, label="terrain 8 bounces")


def test_empty_scene(gpu_ctx):
    gpu_ctx.scene_upload({})
    cam = oracle_camera("C1B", 64, 36)
    sc = orc.Scene()
    _run(gpu_ctx, sc, cam, 64, 36, 2, 2, label="empty")
