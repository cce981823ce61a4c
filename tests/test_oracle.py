"""not gpu: pin the oracle against every known-answer vector there is for this path, and against itself.

The reference has no tests / golden vectors / fixtures: what exists is (1) the RNG vectors the
survey derived from an independent Python transcription of RTUtils.cs, (2) closed-form answers for the intersectors,
the pack and the default scene, (3) the libm build of the same restatement, (4) the engine mirror's independent
restatement of the scene builders (test_host.py) and (5) golden fixtures minted by tests/golden/make_golden.py.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import orc
from tests.util import oracle_camera, oracle_scene_from_spec
from ilgpu_raytracing_b200 import scenes

RNG_KATS = [  # (px, py, frame, sample, lockNoise) -> seed, first three NextUInt, first NextFloat   (salt 0xC0FFEE)
    ((0, 0, 0, 0, 0), 0xE582D06F, (0x4E629AA8, 0xBBC51253, 0x2860AC14), 0.38517237),
    ((0, 0, 0, 0, 1), 0x0FB18B0D, (0xE4D6B8C5, 0x4A0E2562, 0x5AE66453), 0.83875686),
    ((640, 360, 0, 0, 1), 0x823905A7, (0xF33B28C1, 0x524B39D0, 0x9B554E68), 0.23109061),
    ((640, 360, 7, 3, 0), 0x236275A3, (0xD7124A48, 0x553BA9C5, 0x20602144), 0.07144594),
    ((1279, 719, 0, 15, 1), 0xC0767119, (0xC4FF9053, 0xE8488E09, 0xC8B38B6D), 0.99829596),
]


@pytest.mark.parametrize("args,seed,us,f0", RNG_KATS)
def test_rng_known_answers(args, seed, us, f0):
    L = orc.lib()
    px, py, fr, s, ln = args
    sd = L.orc_rng_seed(px, py, fr, s, 0xC0FFEE, ln)
    assert sd == seed
    u = np.zeros(3, np.uint32)
    f = np.zeros(3, np.float32)
    L.orc_rng_stream(sd, 3, u.ctypes.data, f.ctypes.data)
    assert tuple(int(x) for x in u) == us
    assert abs(float(f[0]) - f0) < 5e-8
    assert np.array_equal(f, (u & 0xFFFFFF).astype(np.float32) * np.float32(1.0 / 16777216.0))


def test_rng_python_transcription_agrees():
    """Independent pure-Python transcription of RTUtils.cs:33-137 vs the C++ oracle on random inputs."""
    M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF

    def rotl(v, r):
        return ((v << (r & 31)) | (v >> ((32 - r) & 31))) & M32

    def splitmix32(x):
        x = (x + 0x9E3779B97F4A7C15) & M64
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
        x ^= x >> 31
        return (x ^ (x >> 32)) & M32

    def pcg(x):
        x ^= x >> 16; x = (x * 0x7FEB352D) & M32; x ^= x >> 15; x = (x * 0x846CA68B) & M32; x ^= x >> 16
        return x

    def hash32(x):
        x ^= x >> 17; x = (x * 0xED5AD4BB) & M32; x ^= x >> 11; x = (x * 0xAC4C1B51) & M32; x ^= x >> 15; x = (x * 0x31848BAB) & M32; x ^= x >> 14
        return x

    def seed(px, py, frame, sample, salt, ln):
        f = 0 if ln != 0 else frame & M32
        lnu = ln & M32
        m0 = (hash32(lnu) ^ ((lnu * 0x1B873593) & M32)) if ln != 0 else 0
        m1 = ((rotl(lnu, 7) * 0x85EBCA6B) & M32) if ln != 0 else 0
        a = px ^ 0xB5297A4D
        b = ((py * 0x68E31DA4) & M32) ^ ((f * 0x9E3779B1 + 0x85EBCA6B) & M32) ^ m0
        c = ((sample ^ 0xC2B2AE35) + rotl(px, 16)) & M32
        d = (((salt ^ 0x27D4EB2F) + rotl(py, 8)) & M32) ^ m1
        s0 = splitmix32(((a << 32) | b) ^ 0xD1B54A32D192ED03)
        s1 = splitmix32(((c << 32) | d) ^ 0x94D049BB133111EB)
        return pcg(s0 ^ ((rotl(s1, 13) + 0x9E3779B1) & M32)) | 1

    rs = np.random.RandomState(7)
    L = orc.lib()
    for _ in range(300):
        px, py, fr, s = int(rs.randint(0, 8192)), int(rs.randint(0, 8192)), int(rs.randint(0, 1000)), int(rs.randint(0, 256))
        ln = int(rs.choice([0, 1, 7, -5, 2 ** 31 - 1, -2 ** 31]))
        assert L.orc_rng_seed(px, py, fr, s, 0xC0FFEE, ln) == seed(px, py, fr, s, 0xC0FFEE, ln)


def test_pack_rgba8():
    L = orc.lib()
    assert L.orc_pack_rgba8(0.0, 0.0, 0.0) & 0xFFFFFFFF == 0xFF000000
    assert L.orc_pack_rgba8(1.0, 1.0, 1.0) & 0xFFFFFFFF == 0xFFFFFFFF
    assert L.orc_pack_rgba8(2.0, -1.0, 0.5) & 0xFFFFFFFF == (0xFF << 24) | (255 << 16) | (0 << 8) | int(np.float32(255.99) * np.float32(0.5))
    assert L.orc_pack_rgba8(float("nan"), 0.25, 0.75) & 0xFF00 == int(np.float32(255.99) * np.float32(0.25)) << 8


def _f(*v):
    return np.array(v, np.float32)


def test_intersectors_closed_form():
    L = orc.lib()
    out = np.zeros(4, np.float32)
    # triangle in the z=5 plane hit straight on: t = 5, bu = bv = 0.25, normal +z component irrelevant (we read n.Y)
    assert L.orc_intersect_triangle(_f(0.25, 0.25, 0).ctypes.data, _f(0, 0, 1).ctypes.data, _f(0, 0, 5).ctypes.data, _f(1, 0, 5).ctypes.data, _f(0, 1, 5).ctypes.data, out.ctypes.data) == 1
    assert out[0] == 5.0 and out[1] == 0.25 and out[2] == 0.25
    # outside the triangle, behind the origin, parallel
    assert L.orc_intersect_triangle(_f(0.8, 0.8, 0).ctypes.data, _f(0, 0, 1).ctypes.data, _f(0, 0, 5).ctypes.data, _f(1, 0, 5).ctypes.data, _f(0, 1, 5).ctypes.data, out.ctypes.data) == 0
    assert L.orc_intersect_triangle(_f(0.25, 0.25, 9).ctypes.data, _f(0, 0, 1).ctypes.data, _f(0, 0, 5).ctypes.data, _f(1, 0, 5).ctypes.data, _f(0, 1, 5).ctypes.data, out.ctypes.data) == 0
    assert L.orc_intersect_triangle(_f(0.25, 0.25, 0).ctypes.data, _f(1, 0, 0).ctypes.data, _f(0, 0, 5).ctypes.data, _f(1, 0, 5).ctypes.data, _f(0, 1, 5).ctypes.data, out.ctypes.data) == 0
    # unit sphere at z=5: front root 4, from inside the back root, normal = (p - c) normalised
    assert L.orc_intersect_sphere(_f(0, 0, 0).ctypes.data, _f(0, 0, 1).ctypes.data, _f(0, 0, 5).ctypes.data, 1.0, out.ctypes.data) == 1
    assert out[0] == 4.0 and tuple(out[1:]) == (0.0, 0.0, -1.0)
    assert L.orc_intersect_sphere(_f(0, 0, 5).ctypes.data, _f(0, 0, 1).ctypes.data, _f(0, 0, 5).ctypes.data, 1.0, out.ctypes.data) == 1
    assert out[0] == 1.0 and tuple(out[1:]) == (0.0, 0.0, 1.0)
    assert L.orc_intersect_sphere(_f(0, 3, 0).ctypes.data, _f(0, 0, 1).ctypes.data, _f(0, 0, 5).ctypes.data, 1.0, out.ctypes.data) == 0
    # slab test incl. the tMin / tMax window and an axis-parallel ray (invDir = 1e8 substitution)
    bmin, bmax = _f(-1, -1, 4), _f(1, 1, 6)
    assert L.orc_intersect_aabb(_f(0, 0, 0).ctypes.data, _f(0, 0, 1).ctypes.data, bmin.ctypes.data, bmax.ctypes.data, 0.001, 1e30) == 1
    assert L.orc_intersect_aabb(_f(0, 0, 0).ctypes.data, _f(0, 0, 1).ctypes.data, bmin.ctypes.data, bmax.ctypes.data, 0.001, 3.9) == 0
    assert L.orc_intersect_aabb(_f(2, 0, 0).ctypes.data, _f(0, 0, 1).ctypes.data, bmin.ctypes.data, bmax.ctypes.data, 0.001, 1e30) == 0
    assert L.orc_intersect_aabb(_f(0, 0, 7).ctypes.data, _f(0, 0, 1).ctypes.data, bmin.ctypes.data, bmax.ctypes.data, 0.001, 1e30) == 0


def _ulps(a, b):
    a, b = np.float32(a), np.float32(b)
    return abs(int(a.view(np.int32)) - int(b.view(np.int32)))


def test_portable_transcendentals_close_to_libm():
    """orc_sincos / atan2 / acos (the pinned stand-ins for XMath on the device) stay within a few ulp of libm."""
    L = orc.lib()
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    for x in np.linspace(0.0, 2 * np.pi, 4001, dtype=np.float32):
        L.orc_math_sincos(float(x), C.byref(s), C.byref(c))
        worst = max(worst, abs(s.value - np.sin(np.float64(x))), abs(c.value - np.cos(np.float64(x))))
    assert worst < 2.5e-7
    rs = np.random.RandomState(3)
    for _ in range(4000):
        y, x = np.float32(rs.uniform(-2, 2)), np.float32(rs.uniform(-2, 2))
        assert abs(L.orc_math_atan2(float(y), float(x)) - np.arctan2(np.float64(y), np.float64(x))) < 6e-7
        v = np.float32(rs.uniform(-1, 1))
        assert abs(L.orc_math_acos(float(v)) - np.arccos(np.float64(v))) < 6e-7
    assert L.orc_math_acos(1.0) == 0.0 and abs(L.orc_math_acos(-1.0) - np.pi) < 3e-7


def test_default_scene_structure():
    """Scene.BuildDefaultScene: 6 spheres, 6 one-node BLASes, 7 TLAS nodes, 2 x 256^2 texels; skip links consistent."""
    sc = orc.Scene()
    sc.build_default()
    a = sc.arrays()
    assert len(a["spheres"]) == 6 and len(a["instances"]) == 6 and len(a["blasNodes"]) == 6 and len(a["tlasNodes"]) == 7
    assert len(a["texels"]) == 2 * 256 * 256 and len(a["texInfos"]) == 2 and len(a["spherePrimIdx"]) == 12
    # walking the TLAS with every box taken visits each instance exactly once
    seen, cur, n = [], 0, a["tlasNodes"]
    while cur != -1:
        if n[cur]["count"] > 0:
            seen += [int(a["tlasInstanceIndices"][i]) for i in range(n[cur]["first"], n[cur]["first"] + n[cur]["count"])]
            cur = int(n[cur]["skipIndex"])
        else:
            cur = int(n[cur]["left"])
    assert sorted(seen) == list(range(6))
    assert sc.sort_ties() > 0   # three spheres share centre y = 0.5: the reference's Array.Sort order matters here


def test_default_camera_quirk():
    """the reference's default camera (CreateCamera + Translate(1,0,-4)) sees only ground and sky."""
    sc = orc.Scene()
    sc.build_default()
    rA = orc.render(sc, oracle_camera("C1A", 320, 180), orc.make_config(320, 180, 1, 1))
    rB = orc.render(sc, oracle_camera("C1B", 320, 180), orc.make_config(320, 180, 1, 1))
    assert set(np.unique(rA.instId)) <= {-1, 0}
    assert set(np.unique(rB.instId)) == {-1, 0, 1, 2, 3, 4, 5}
    assert np.all(rB.objId == -1)   # spheres report objId -1 (SceneDeviceViews.cs:56,73)


def test_libm_variant_within_tolerance():
    """The same restatement with libm sin/cos/atan2/acos: ids identical on almost every path, radiance within 1e-4 rel RMS
    of the pinned oracle -> the choice of transcendental kernels is below the north-star tolerance."""
    spec = scenes.sphere_grid_scene(8)
    a, b = oracle_scene_from_spec(spec), oracle_scene_from_spec(spec, "libm")
    cam = oracle_camera("C2", 240, 135)
    cfg = orc.make_config(240, 135, spp=8, max_depth=4)
    ra, rb = orc.render(a, cam, cfg), orc.render(b, cam, cfg)
    assert np.array_equal(ra.primId, rb.primId)
    assert (ra.pathHash != rb.pathHash).mean() < 2e-3
    num = np.sqrt(np.mean((ra.radiance.astype(np.float64) - rb.radiance) ** 2))
    den = np.sqrt(np.mean(rb.radiance.astype(np.float64) ** 2))
    assert num / den < 2e-3   # differing paths are rare; the bulk is identical
    same = (ra.pathHash == rb.pathHash).all(axis=0)
    num = np.sqrt(np.mean((ra.radiance[same].astype(np.float64) - rb.radiance[same]) ** 2))
    assert num / den < 1e-4


def test_cull_free_trace_agrees_on_generic_rays():
    """TraceClosest with every box test taken (no culling) = the reference traversal, away from degenerate rays."""
    sc = oracle_scene_from_spec(scenes.terrain_scene(24, 9))
    rs = np.random.RandomState(11)
    for _ in range(300):
        o = np.array([rs.uniform(-60, 60), rs.uniform(8, 40), rs.uniform(-60, 60)], np.float32)
        t = np.array([rs.uniform(-50, 50), rs.uniform(-6, 6), rs.uniform(-50, 50)], np.float32)
        d = (t - o) / np.linalg.norm(t - o)
        assert sc.trace_closest(o, d.astype(np.float32), cull=True) == sc.trace_closest(o, d.astype(np.float32), cull=False)


def test_golden_fixtures():
    """tests/golden/*.npz, minted by tests/golden/make_golden.py from this oracle: regression pin for the oracle itself."""
    import glob
    import os
    from tests.golden.make_golden import CASES, render_case
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
    assert len(files) == len(CASES)
    for f in files:
        g = np.load(f)
        r = render_case(str(g["name"]))
        assert np.array_equal(r.primId, g["primId"]) and np.array_equal(r.rgba8, g["rgba8"])
        assert np.array_equal(r.segCount, g["segCount"]) and np.array_equal(r.pathHash, g["pathHash"])
        assert np.array_equal(r.radiance, g["radiance"])


def test_independent_python_restatement_agrees():
    """tests/pyref.py: the hot path restated a second time, in scalar Python straight from the C# sources (brute-force sphere
    scene, numpy transcendentals), against BOTH builds of the oracle on the reference's default scene: primary hit ids, ray
    counts and per-sample bounce counts equal, RGBA8 equal, radiance far inside the north-star tolerance of 1e-4 relative RMS
    (measured: 8e-7 against the pinned build, 3e-6 against the libm build; 96-99 % of the pixels bit-identical).  Mirror, glass,
    textured and Lambert spheres, ReSTIR-DI, roulette (depth 3 of MaxDepth 4) and the sky are all on the way."""
    from tests import pyref
    from ilgpu_raytracing_b200 import layouts
    W, H, spp, depth = 64, 36, 2, 4
    spec = scenes.default_scene()
    got = pyref.render(pyref.Scene(spec.spheres, spec.textures), pyref.Camera(W, H, 60.0), W, H, spp, depth, layouts.default_sun_dir())
    assert {1, 2} <= set(int(s["shading"]) for s in spec.spheres)   # the scene really has a mirror and a glass sphere in it
    for variant in ("", "libm"):
        ref = orc.render(oracle_scene_from_spec(spec, variant), oracle_camera("C1B", W, H), orc.make_config(W, H, spp=spp, max_depth=depth))
        hit = ref.primId >= 0
        assert hit.sum() > W * H // 3 and (~hit).sum() > 0
        assert np.array_equal(got["sphere"][hit], ref.primId[hit]) and np.array_equal(got["hit"], hit)
        assert np.allclose(got["depth"], ref.depth, rtol=2e-6, atol=0)
        assert np.array_equal(got["seg"], ref.segCount)
        assert got["counters"]["bounce"] == ref.counters["raysBounce"] and got["counters"]["shadow"] == ref.counters["raysShadow"]
        assert (ref.termCode == 3).sum() > 0   # roulette fired somewhere
        den = np.sqrt(np.mean(ref.radiance[:, :3].astype(np.float64) ** 2))
        num = np.sqrt(np.mean((got["radiance"].astype(np.float64) - ref.radiance[:, :3]) ** 2))
        assert num / den < 2e-5, (variant, num / den)
        assert (got["rgba8"] != ref.rgba8).mean() < 0.005


def test_independent_python_restatement_agrees_on_triangles():
    """The same second restatement on the triangle path: Moeller-Trumbore, barycentric uv, bilinear colour fetch, alpha cut-out in
    closest-hit (linear mask) and any-hit (point sample, +-0.10 band, then linear), two-sided normals, triangles always Lambert -
    next to textured, mirror and glass spheres (tests.util.special_scene without its exactly duplicated triangles, whose ties only
    the visiting order decides)."""
    import dataclasses
    from tests import pyref
    from tests.util import SPECIAL_CAMERA, special_scene
    from ilgpu_raytracing_b200 import layouts
    W, H, spp, depth = 56, 32, 2, 4
    spec = special_scene("identity")
    m = spec.mesh
    spec = dataclasses.replace(spec, mesh=dataclasses.replace(m, tris=m.tris[:-20], tri_uvs=m.tri_uvs[:-20], tri_mat=m.tri_mat[:-20]))
    cam_p = pyref.Camera(W, H, SPECIAL_CAMERA["fov"], origin=SPECIAL_CAMERA["origin"], look_at=SPECIAL_CAMERA["look_at"])
    got = pyref.render(pyref.Scene(spec.spheres, spec.textures, spec.mesh), cam_p, W, H, spp, depth, layouts.default_sun_dir())
    for variant in ("", "libm"):
        cam_o = orc.camera_create(W, H, SPECIAL_CAMERA["fov"], SPECIAL_CAMERA["origin"], SPECIAL_CAMERA["look_at"], variant=variant)
        ref = orc.render(oracle_scene_from_spec(spec, variant), cam_o, orc.make_config(W, H, spp=spp, max_depth=depth))
        hit = ref.primId >= 0
        tri = hit & (ref.objId >= 0)
        assert tri.sum() > W * H // 8 and (hit & ~tri).sum() > W * H // 50
        assert np.array_equal(got["hit"], hit)
        assert np.array_equal(-1 - got["sphere"][tri], ref.primId[tri]) and np.array_equal(got["sphere"][hit & ~tri], ref.primId[hit & ~tri])
        assert np.array_equal(got["seg"], ref.segCount)
        assert got["counters"]["bounce"] == ref.counters["raysBounce"] and got["counters"]["shadow"] == ref.counters["raysShadow"]
        den = np.sqrt(np.mean(ref.radiance[:, :3].astype(np.float64) ** 2))
        num = np.sqrt(np.mean((got["radiance"].astype(np.float64) - ref.radiance[:, :3]) ** 2))
        assert num / den < 2e-5, (variant, num / den)
        assert (got["rgba8"] != ref.rgba8).mean() < 0.005


@pytest.mark.parametrize("temporal,spatial", [(1, 1), (1, 0), (0, 1)])
def test_independent_python_restatement_agrees_on_restir_reuse(temporal, spatial):
    """The second restatement with ReSTIR temporal / spatial reuse (ReprojectToPrevPixel, SpatialCompatible, Neighbor8,
    ImportFromPrevReservoir, the reservoir ping-pong; RTRay.cs:339-435, 475-516) over three frames of a moving camera: the
    reservoirs every frame leaves behind and the images agree with the oracle."""
    from tests import pyref
    from ilgpu_raytracing_b200 import layouts
    W, H, spp, depth = 40, 24, 2, 3
    spec = scenes.default_scene()
    sc = oracle_scene_from_spec(spec)
    psc = pyref.Scene(spec.spheres, spec.textures)
    ores = [np.zeros(W * H, orc.RESERVOIR), np.zeros(W * H, orc.RESERVOIR)]
    pres = [np.zeros(W * H, orc.RESERVOIR), np.zeros(W * H, orc.RESERVOIR)]
    prev_o = prev_p = None
    imported = 0
    for frame in range(3):
        origin = (0.07 * frame, 1.0 + 0.02 * frame, 3.0 - 0.05 * frame)
        cam_o = orc.camera_create(W, H, 60.0, origin, (0.0, 0.5, 0.0))
        orc.camera_bake(cam_o, W, H)
        cam_p = pyref.Camera(W, H, 60.0, origin=origin, look_at=(0.0, 0.5, 0.0))
        prev_o = cam_o.copy() if prev_o is None else prev_o
        prev_p = cam_p if prev_p is None else prev_p
        ci = frame & 1
        ref = orc.render(sc, cam_o, orc.make_config(W, H, spp=spp, max_depth=depth, frame=frame, rng_lock_noise=0, temporal=temporal, spatial=spatial),
                         prev_cam=prev_o, res_prev=ores[ci ^ 1], res_cur=ores[ci])
        got = pyref.render(psc, cam_p, W, H, spp, depth, layouts.default_sun_dir(), frame=frame, lock_noise=0, temporal=temporal, spatial=spatial,
                           prev_cam=prev_p, res_prev=pres[ci ^ 1], res_cur=pres[ci])
        assert np.array_equal(got["seg"], ref.segCount), frame
        assert got["counters"]["bounce"] == ref.counters["raysBounce"] and got["counters"]["shadow"] == ref.counters["raysShadow"], frame
        a, b = pres[ci], ores[ci]
        assert np.array_equal(a["m"], b["m"]) and np.array_equal(a["lightId"], b["lightId"]), frame
        for f in ("wSum", "w", "pdf"):   # the texture coordinates of the textured spheres go through atan2 / acos: numpy's vs the oracle's kernels
            assert np.allclose(a[f], b[f], rtol=1e-3, atol=1e-7) and (np.abs(a[f] - b[f]) > 2e-5 * np.abs(b[f]) + 1e-7).mean() < 0.02, (frame, f)
        for f in ("wi", "L"):
            for c in "XYZ":
                assert np.allclose(a[f][c], b[f][c], rtol=2e-5, atol=2e-6), (frame, f, c)
        den = np.sqrt(np.mean(ref.radiance[:, :3].astype(np.float64) ** 2))
        num = np.sqrt(np.mean((got["radiance"].astype(np.float64) - ref.radiance[:, :3]) ** 2))
        assert num / den < 2e-5, (frame, num / den)
        assert (got["rgba8"] != ref.rgba8).mean() < 0.01
        imported += int((b["m"] > 9).sum()) if frame > 0 else 0
        prev_o, prev_p = cam_o.copy(), cam_p
    assert imported > 0   # frames 1 and 2 really imported previous reservoirs


def test_independent_python_restatement_agrees_on_present_chain():
    """The second restatement of the present chain (TaaResolveKernel with its two-tap "Catmull-Rom", 3x3 clamp, objId
    disocclusion, sharpening, sRGB pack; BilinearUpsampleKernel) against the oracle over a four-frame TAAU sequence rendered at
    0.67 scale: bilinear bit-exact, TAAU equal except where numpy's pow and the oracle's pinned pow land on opposite sides of an
    8-bit rounding step (at most one code value, rare), which the history then carries."""
    from tests import pyref
    outW, outH = 96, 54
    inW, inH = int(np.rint(np.float32(outW) * np.float32(0.67))), int(np.rint(np.float32(outH) * np.float32(0.67)))
    sc = orc.Scene()
    sc.build_default()
    st_o, st_p = orc.TaaState(outW, outH), pyref.Taa(outW, outH)
    for frame in range(4):
        cam = orc.camera_create(inW, inH, 60.0, (0.05 * frame, 1.0, 3.0), (0.0, 0.5, 0.0))
        low = orc.render(sc, cam, orc.make_config(inW, inH, spp=1, max_depth=2, frame=frame, rng_lock_noise=0), aovs=False)
        assert np.array_equal(pyref.bilinear_upsample(low.rgba8, inW, inH, outW, outH), orc.bilinear_upsample(low.rgba8, inW, inH, outW, outH))
        a, b = st_p.resolve(low.rgba8, low.objId, inW, inH), st_o.resolve(low.rgba8, low.objId, inW, inH)
        assert np.array_equal(st_p.hist_obj, st_o.hist_obj)
        diff = np.stack([np.abs(((a >> sh) & 255) - ((b >> sh) & 255)) for sh in (16, 8, 0)])
        assert diff.max() <= 1 and (diff > 0).mean() < 0.005, (frame, diff.max(), (diff > 0).mean())   # measured: identical on all four frames
        assert ((a >> 24) & 255 == 255).all()


@pytest.mark.parametrize("mode", ["translated", "scaled"])
def test_independent_python_restatement_agrees_on_instances(mode):
    """The second restatement with instance transforms (TransformRay without renormalising, tWorld = tObj / scale, tMaxObj =
    tMaxWorld * scale, Scene.InvertRigidOrUniform's un-transposed rotation, normals through objectToWorld): translated instances
    against the oracle as it is; rotated + scaled ones - where the reference's own box culling depends on the visiting order
     - against the oracle with every box taken, which is what a brute-force restatement computes."""
    import dataclasses
    from tests import pyref
    from tests.util import SPECIAL_CAMERA, special_scene
    from ilgpu_raytracing_b200 import layouts
    W, H, spp, depth = 48, 28, 2, 3
    spec = special_scene(mode)
    m = spec.mesh
    spec = dataclasses.replace(spec, mesh=dataclasses.replace(m, tris=m.tris[:-20], tri_uvs=m.tri_uvs[:-20], tri_mat=m.tri_mat[:-20]))
    sphere_xf = [None] * len(spec.spheres)
    for ids, xf in spec.sphere_instances:
        assert len(ids) == 1
        sphere_xf[int(ids[0])] = xf
    psc = pyref.Scene(spec.spheres, spec.textures, spec.mesh, sphere_xf=sphere_xf)
    got = pyref.render(psc, pyref.Camera(W, H, SPECIAL_CAMERA["fov"], origin=SPECIAL_CAMERA["origin"], look_at=SPECIAL_CAMERA["look_at"]),
                       W, H, spp, depth, layouts.default_sun_dir())
    cam_o = orc.camera_create(W, H, SPECIAL_CAMERA["fov"], SPECIAL_CAMERA["origin"], SPECIAL_CAMERA["look_at"])
    ref = orc.render(oracle_scene_from_spec(spec), cam_o, orc.make_config(W, H, spp=spp, max_depth=depth, no_cull=1 if mode == "scaled" else 0))
    hit = ref.primId >= 0
    tri = hit & (ref.objId >= 0)
    assert tri.sum() > W * H // 10 and (hit & ~tri).sum() > W * H // 50
    assert np.array_equal(got["hit"], hit)
    assert np.array_equal(-1 - got["sphere"][tri], ref.primId[tri]) and np.array_equal(got["sphere"][hit & ~tri], ref.primId[hit & ~tri])
    same = (got["seg"] == ref.segCount).all(axis=0)
    assert same.mean() > 0.995, same.mean()   # a rotated instance's rays go through sin / cos of the transform: rare last-bit decisions
    den = np.sqrt(np.mean(ref.radiance[:, :3].astype(np.float64) ** 2))
    num = np.sqrt(np.mean((got["radiance"][same].astype(np.float64) - ref.radiance[same, :3]) ** 2))
    assert num / den < 1e-4, num / den
