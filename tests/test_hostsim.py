"""not gpu: the renderer core's stage bodies (wide-BVH build, traversal, tie-break rule, wavefront state machine),
compiled for the CPU by tests/hostsim, against the oracle.  The same bodies run inside the CUDA kernels; the -m gpu
tests repeat these comparisons through the C ABI on the device."""
import numpy as np
import pytest

from ilgpu_raytracing_b200 import layouts as L
from ilgpu_raytracing_b200 import scenes
from oracle import orc
from tests.hostsim_binding import HostSimScene
from tests.util import oracle_camera, oracle_scene_from_spec, special_camera, special_scene


def _compare(r, h, label=""):
    for k in ("primId", "instId", "primaryT", "gbPos", "gbNrm", "gbAlb", "gbMat", "objId", "depth", "segCount", "termCode", "pathHash", "rgba8"):
        assert np.array_equal(getattr(r, k), h[k]), f"{label}: {k} differs"
    assert np.array_equal(r.radiance, h["radiance"]), f"{label}: radiance not bit-identical"


@pytest.mark.parametrize("spp,depth", [(1, 0), (1, 1), (3, 4), (2, 8)])
def test_default_scene(spp, depth):
    sc = orc.Scene()
    sc.build_default()
    hs = HostSimScene(sc.arrays())
    for camname in ("C1A", "C1B"):
        cam = oracle_camera(camname, 192, 108)
        r = orc.render(sc, cam, orc.make_config(192, 108, spp=spp, max_depth=depth))
        h = hs.render(cam, L.make_render_config(192, 108, spp=spp, max_depth=depth))
        _compare(r, h, f"default {camname} {spp}spp d{depth}")
        assert h["counters"]["raysBounce"] == r.counters["raysBounce"] and h["counters"]["raysShadow"] == r.counters["raysShadow"]


@pytest.mark.parametrize("temporal,spatial", [(1, 1), (1, 0), (0, 1)])
def test_restir_reuse_sequence(temporal, spatial):
    """ReSTIR temporal / spatial reuse (RTRay.cs:475-516) over three frames with a moving camera; see the -m gpu twin."""
    W, H, spp, depth = 128, 72, 3, 3
    sc = orc.Scene()
    sc.build_default()
    hs = HostSimScene(sc.arrays())
    ores = [np.zeros(W * H, orc.RESERVOIR), np.zeros(W * H, orc.RESERVOIR)]
    hres = [np.zeros(W * H, L.RESERVOIR), np.zeros(W * H, L.RESERVOIR)]
    prev = None
    for frame in range(3):
        cam = orc.camera_create(W, H, 60.0, (0.07 * frame, 1.0 + 0.02 * frame, 3.0 - 0.05 * frame), (0.0, 0.5, 0.0))
        orc.camera_bake(cam, W, H)
        prev = cam.copy() if prev is None else prev
        ci = frame & 1
        r = orc.render(sc, cam, orc.make_config(W, H, spp=spp, max_depth=depth, frame=frame, rng_lock_noise=0, temporal=temporal, spatial=spatial),
                       prev_cam=prev, res_prev=ores[ci ^ 1], res_cur=ores[ci])
        h = hs.render(cam, L.make_render_config(W, H, spp=spp, max_depth=depth, frame=frame, rng_lock_noise=0, temporal=temporal, spatial=spatial,
                                                samples_per_pass=2), prev_cam=prev, res_prev=hres[ci ^ 1], res_cur=hres[ci])
        _compare(r, h, f"reuse frame {frame}")
        assert ores[ci].tobytes() == hres[ci].tobytes(), f"frame {frame}: reservoirs differ"
        if frame > 0:
            assert int((ores[ci]["m"] > 9).sum()) > 0
        prev = cam.copy()


def test_present_chain_bodies():
    """Blit / bilinear upsample / TAAU resolve (RTRenderer.cs:281-320, RTTaa.cs:117-262): the core's stage bodies against the oracle
    on a rendered low-res frame, three frames of history, and the pow stand-in they share."""
    from tests.hostsim_binding import lib as hlib
    import ctypes as C
    h = hlib()
    rs = np.random.RandomState(3)
    for x, y in zip(rs.uniform(1e-4, 4.0, 2000).astype(np.float32), rs.choice(np.array([2.4, 1 / 2.4, 0.5, 3.0], np.float32), 2000)):
        assert h.hs_pow(float(x), float(y)) == orc.lib().orc_math_pow(float(x), float(y))
    inW, inH, outW, outH = 86, 48, 128, 72   # renderScale 0.67
    sc = orc.Scene()
    sc.build_default()
    st = orc.TaaState(outW, outH)
    hc, ho = np.zeros(outW * outH, np.int32), np.zeros(outW * outH, np.int32)
    for frame in range(3):
        cam = orc.camera_create(inW, inH, 60.0, (0.05 * frame, 1.0, 3.0), (0.0, 0.5, 0.0))
        r = orc.render(sc, cam, orc.make_config(inW, inH, spp=1, max_depth=2, frame=frame, rng_lock_noise=0), aovs=False)
        want = st.resolve(r.rgba8, r.objId, inW, inH)
        got = np.zeros(outW * outH, np.int32)
        lc, lo = np.ascontiguousarray(r.rgba8), np.ascontiguousarray(r.objId)
        h.hs_taa_resolve(got.ctypes.data, lc.ctypes.data, lo.ctypes.data, inW, inH, outW, outH, hc.ctypes.data, ho.ctypes.data, 1 if frame == 0 else 0, 0.075, 0.10, 1.25)
        assert np.array_equal(got, want), f"TAAU frame {frame}"
        assert np.array_equal(hc, st.hist_color) and np.array_equal(ho, st.hist_obj)
        up = np.zeros(outW * outH, np.int32)
        h.hs_bilinear_upsample(lc.ctypes.data, inW, inH, up.ctypes.data, outW, outH)
        assert np.array_equal(up, orc.bilinear_upsample(r.rgba8, inW, inH, outW, outH))
    assert len(np.unique(want)) > 50 and (want >> 24 & 255 == 255).all()


def test_sphere_grid_with_roulette():
    sc = oracle_scene_from_spec(scenes.sphere_grid_scene(12))
    hs = HostSimScene(sc.arrays())
    cam = oracle_camera("C2", 160, 90)
    r = orc.render(sc, cam, orc.make_config(160, 90, spp=6, max_depth=6))
    h = hs.render(cam, L.make_render_config(160, 90, spp=6, max_depth=6, samples_per_pass=4))
    _compare(r, h, "sphere grid")
    assert (r.termCode == 3).sum() > 0


@pytest.mark.parametrize("flags", [0, L.RT_FLAG_TRI_MATERIALS])
def test_terrain(flags):
    sc = oracle_scene_from_spec(scenes.terrain_scene(64, 16, patch_materials=bool(flags)))
    hs = HostSimScene(sc.arrays())
    st = hs.stats()
    assert st["nPrims"] == 2 * 64 * 64 + 16 and st["nWideNodes"] < st["nPrims"] / 5
    cam = oracle_camera("C3", 160, 90)
    r = orc.render(sc, cam, orc.make_config(160, 90, spp=2, max_depth=8, flags=flags & 1))
    h = hs.render(cam, L.make_render_config(160, 90, spp=2, max_depth=8, flags=flags))
    _compare(r, h, "terrain")
    # the wide BVH visits far fewer nodes than the reference's BVH2
    assert h["counters"]["nodes"] < r.counters["nodes"] / 3


@pytest.mark.parametrize("node_steps", [1, 2, 3])
def test_kernel_lane_schedule_with_queued_primitive_groups(node_steps):
    """The traversal under k_extend's per-lane schedule (n node steps, then ONE primitive step, primitives of further node steps queued
    in the second group slot): same hits, same image as the oracle - the closest hit does not depend on when a queued test runs."""
    from tests.hostsim_binding import set_lane_schedule
    from tests.util import special_camera, special_scene
    try:
        set_lane_schedule(node_steps)
        sc = oracle_scene_from_spec(scenes.terrain_scene(48, 12))
        cam = oracle_camera("C3", 128, 72)
        r = orc.render(sc, cam, orc.make_config(128, 72, spp=2, max_depth=6))
        _compare(r, HostSimScene(sc.arrays()).render(cam, L.make_render_config(128, 72, spp=2, max_depth=6)), f"schedule {node_steps} terrain")
        sc = oracle_scene_from_spec(special_scene("translated"))
        cam = special_camera(120, 72)
        r = orc.render(sc, cam, orc.make_config(120, 72, spp=2, max_depth=4))
        _compare(r, HostSimScene(sc.arrays()).render(cam, L.make_render_config(120, 72, spp=2, max_depth=4)), f"schedule {node_steps} special")
    finally:
        set_lane_schedule(0)


@pytest.mark.parametrize("transformed", ["identity", "translated"])
def test_textures_alpha_ties_instances(transformed):
    """Textured / alpha-masked / two-sided triangles, duplicated triangles (equal-t ties) and translated instances."""
    sc = oracle_scene_from_spec(special_scene(transformed))
    hs = HostSimScene(sc.arrays())
    cam = special_camera(200, 120)
    r = orc.render(sc, cam, orc.make_config(200, 120, spp=3, max_depth=5))
    h = hs.render(cam, L.make_render_config(200, 120, spp=3, max_depth=5))
    _compare(r, h, f"special transformed={transformed}")
    assert (r.hitMask > 0).mean() > 0.5 and len(np.unique(r.instId)) >= 5


def test_tie_break_prefers_reference_visiting_order():
    """Two coincident triangles: the reference keeps the first it visits (strict '<'); the wide traversal must agree."""
    spec = special_scene("identity")
    sc = oracle_scene_from_spec(spec)
    hs = HostSimScene(sc.arrays())
    m = spec.mesh
    n_orig = len(m.tris) - 20
    on_pair = 0
    for k in range(20):
        tri = m.tris[n_orig + k]
        c = m.positions[tri].mean(axis=0)
        o = (c + np.array([0.01, 5.0, 0.02], np.float32)).astype(np.float32)
        d = (c - o) / np.linalg.norm(c - o)
        hit_o, t_o, inst_o, prim_o = sc.trace_closest(o, d.astype(np.float32))
        hit_h, t_h, inst_h, prim_h, _ = hs.trace(o, d.astype(np.float32))
        assert hit_o and hit_h and (t_o, inst_o, prim_o) == (t_h, inst_h, prim_h)
        on_pair += prim_o in (40 + k, n_orig + k)   # (a sphere may sit in front of some of them)
    assert on_pair >= 10


def test_any_hit_matches_shadow_occlusion():
    sc = oracle_scene_from_spec(special_scene("translated"))
    hs = HostSimScene(sc.arrays())
    cam = special_camera(96, 54)
    # shadow visibility is folded into the path hash (0x100 | visible): one Lambert bounce is enough to cover it
    r = orc.render(sc, cam, orc.make_config(96, 54, spp=2, max_depth=1))
    h = hs.render(cam, L.make_render_config(96, 54, spp=2, max_depth=1))
    assert np.array_equal(r.pathHash, h["pathHash"]) and r.counters["raysShadow"] == h["counters"]["raysShadow"] > 0


def test_scaled_instances_match_cull_free_reference():
    """uniformScale != 1: the reference reports tWorld = tObj / scale, which makes its own box culling depend on the visiting
    order (a sphere found first can cull a scaled mesh whose reported t is smaller).  The core returns the order-independent
    minimum of the reference's own hit rule = the oracle with every box test taken."""
    sc = oracle_scene_from_spec(special_scene("scaled"))
    hs = HostSimScene(sc.arrays())
    rs = np.random.RandomState(1)
    differs = 0
    for _ in range(1500):
        o = np.array([rs.uniform(-8, 8), rs.uniform(1, 6), rs.uniform(-8, 8)], np.float32)
        t = np.array([rs.uniform(-4, 4), rs.uniform(-0.5, 1.5), rs.uniform(-4, 4)], np.float32)
        d = ((t - o) / np.linalg.norm(t - o)).astype(np.float32)
        want = sc.trace_closest(o, d, cull=False)
        hit, tt, inst, prim, _ = hs.trace(o, d)
        assert (hit, tt, inst, prim) == want
        differs += want != sc.trace_closest(o, d, cull=True)
    assert differs > 0   # the reference's culled walk really is order dependent on this scene


def test_tile_partition_is_exact():
    """Interleaved screen tiles: the union of the ranks' pixels is the single-context image, bit for bit."""
    sc = oracle_scene_from_spec(scenes.terrain_scene(32, 9))
    hs = HostSimScene(sc.arrays())
    W, H = 200, 104   # not multiples of the tile size
    cam = oracle_camera("C3", W, H)
    full = hs.render(cam, L.make_render_config(W, H, spp=2, max_depth=3))
    for world in (2, 3):
        acc = np.zeros((W * H, 3), np.float32)
        owned = np.zeros(W * H, np.int32)
        for rank in range(world):
            part = hs.render(cam, L.make_render_config(W, H, spp=2, max_depth=3, tile_size=32, rank=rank, world_size=world))
            mine = part["primId"] != -2   # hostsim initialises untouched pixels to -2
            owned += mine
            acc[mine] = part["radiance"][mine]
            assert np.array_equal(part["rgba8"][mine], full["rgba8"][mine])
        assert np.all(owned == 1)
        assert np.array_equal(acc, full["radiance"])


def test_malformed_scene_is_rejected():
    sc = orc.Scene()
    sc.build_default()
    a = sc.arrays()
    bad = dict(a)
    bad["tlasInstanceIndices"] = a["tlasInstanceIndices"].copy()
    bad["tlasInstanceIndices"][0] = 99
    with pytest.raises(ValueError):
        HostSimScene(bad)
    bad = dict(a)
    bad["spherePrimIdx"] = a["spherePrimIdx"].copy()
    bad["spherePrimIdx"][6:] = 77
    with pytest.raises(ValueError):
        HostSimScene(bad)


def test_host_builder_output_is_pinned():
    """The wide BVH the host builder makes (binned SAH with big subtrees built by their own threads into pre-computed index ranges,
    SAH-optimal collapse, quantisation) is the one the sequential builder made: FNV-1a of nodes + primitive records, recorded before
    the threaded build went in.  Same bytes on every run (no race decides an index)."""
    from tests.util import special_scene
    sc = orc.Scene()
    sc.build_default()
    pinned = {"default": (sc, 0xa435527c2ab62924), "grid32": (oracle_scene_from_spec(scenes.sphere_grid_scene(32)), 0xaeea26dd0c1c73f2),
              "terrain96": (oracle_scene_from_spec(scenes.terrain_scene(96, 24)), 0x295cb80c41c21634),
              "special_t": (oracle_scene_from_spec(special_scene("translated")), 0xc28f7b40f58f6a2d),
              "terrain300": (oracle_scene_from_spec(scenes.terrain_scene(300, 64)), 0xd6dae8d89b51afaf)}   # 180 k triangles: the threaded path
    for name, (scene, want) in pinned.items():
        arrays = scene.arrays()
        for _ in range(2):
            assert HostSimScene(arrays).bvh_hash() == want, name


def test_too_deep_tree_is_rebuilt_depth_bounded():
    """a SAH tree deeper than the traversal stack must not fail the commit (the reference's skip-link walk has no
    stack).  Force the case with a small depth limit: the depth-bounded rebuild (median splits, three binary levels per wide
    node) must respect the limit and render the same image as the oracle."""
    spec = scenes.terrain_scene(n_quads=48, n_spheres=9)
    sc = oracle_scene_from_spec(spec)
    normal = HostSimScene(sc.arrays())
    limit = 4                                     # 4 617 primitives: the SAH tree has 5 levels, the bounded one 4 (585 nodes = 1 + 8 + 64 + 512)
    assert normal.stats()["maxDepth"] > limit and normal.stats()["depthBounded"] == 0
    bounded = HostSimScene(sc.arrays(), max_depth=limit)
    st = bounded.stats()
    assert st["depthBounded"] == 1 and st["maxDepth"] <= limit and st["nPrims"] == normal.stats()["nPrims"]
    cam = oracle_camera("C3", 96, 54)
    r = orc.render(sc, cam, orc.make_config(96, 54, spp=2, max_depth=4))
    _compare(r, bounded.render(cam, L.make_render_config(96, 54, spp=2, max_depth=4)), "depth-bounded tree")
