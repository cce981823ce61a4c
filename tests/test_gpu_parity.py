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


@pytest.mark.parametrize("mode", ["identity", "translated"])
def test_textures_alpha_ties_instances(gpu_ctx, mode):
    """Textured / alpha-masked / two-sided triangles, duplicated triangles (equal-t ties) and translated instances."""
    from tests.util import special_camera, special_scene
    sc = oracle_scene_from_spec(special_scene(mode))
    gpu_ctx.scene_upload(sc.arrays())
    _run(gpu_ctx, sc, special_camera(400, 240), 400, 240, 3, 5, label=f"special {mode}")


def test_scaled_rotated_instances_vs_cull_free_reference(gpu_ctx):
    """uniformScale != 1 / rotations: the reference's own culling is visiting-order dependent (see tests/util.special_scene);
    the core must return the order-independent result = the oracle with every box test taken (primary hits)."""
    from tests.util import special_camera, special_scene
    sc = oracle_scene_from_spec(special_scene("scaled"))
    gpu_ctx.scene_upload(sc.arrays())
    W, H = 200, 120
    cam = special_camera(W, H)
    r = orc.render(sc, cam, orc.make_config(W, H, spp=1, max_depth=0, no_cull=1), aovs=False)
    gpu_ctx.render(cam, L.make_render_config(W, H, spp=1, max_depth=0))
    gpu_ctx.sync()
    assert np.array_equal(gpu_ctx.download(L.RT_BUF_PRIM_ID), r.primId)
    assert np.array_equal(gpu_ctx.download(L.RT_BUF_INST_ID), r.instId)
    assert np.array_equal(gpu_ctx.download(L.RT_BUF_PRIMARY_T), r.primaryT)
    assert np.array_equal(gpu_ctx.download(L.RT_BUF_GB_NORMAL), r.gbNrm)


def test_golden_fixtures(gpu_ctx):
    """Committed input/output pairs (tests/golden, minted from the oracle by make_golden.py) through the C ABI."""
    import glob
    import os
    from tests.golden.make_golden import CASES, make_spec
    for f in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))):
        g = np.load(f)
        kind, camname, W, H, spp, depth, flags = CASES[str(g["name"])]
        sc = oracle_scene_from_spec(make_spec(kind))
        gpu_ctx.scene_upload(sc.arrays())
        gpu_ctx.render(oracle_camera(camname, W, H), L.make_render_config(W, H, spp=spp, max_depth=depth, flags=flags | L.RT_FLAG_PATH_AOVS))
        gpu_ctx.sync()
        assert np.array_equal(gpu_ctx.download(L.RT_BUF_PRIM_ID), g["primId"]), f
        assert np.array_equal(gpu_ctx.download(L.RT_BUF_SEG_COUNT).reshape(g["segCount"].shape), g["segCount"]), f
        assert np.array_equal(gpu_ctx.download(L.RT_BUF_PATH_HASH).reshape(g["pathHash"].shape), g["pathHash"]), f
        assert np.array_equal(gpu_ctx.download(L.RT_BUF_RADIANCE)[:, :3], g["radiance"]), f
        assert np.array_equal(gpu_ctx.download(L.RT_BUF_RGBA8), g["rgba8"]), f


def test_tile_partition_and_deinterleave(gpu_ctx):
    """Multi-GPU protocol on one device: every rank's tiles rendered in turn, payloads concatenated rank-major as the NCCL
    gather delivers them, rt_deinterleave_tiles reassembles the frame == the single-context image, bit for bit."""
    import torch
    from ilgpu_raytracing_b200 import native
    sc = oracle_scene_from_spec(scenes.terrain_scene(64, 16))
    gpu_ctx.scene_upload(sc.arrays())
    W, H, tile = 520, 296, 32
    cam = oracle_camera("C3", W, H)
    gpu_ctx.render(cam, L.make_render_config(W, H, spp=3, max_depth=4))
    gpu_ctx.sync()
    full = gpu_ctx.download(L.RT_BUF_RADIANCE).copy()
    full_rgba = gpu_ctx.download(L.RT_BUF_RGBA8).copy()
    for world in (2, 8):
        counts = [native.tiles_owned_pixels(W, H, tile, r, world) for r in range(world)]
        assert sum(counts) == W * H
        max_n = max(counts)
        flat = torch.zeros((world * max_n, 4), dtype=torch.float32, device="cuda")
        for rank in range(world):
            gpu_ctx.render(cam, L.make_render_config(W, H, spp=3, max_depth=4, tile_size=tile, rank=rank, world_size=world))
            gpu_ctx.sync()
            pay = gpu_ctx.download(L.RT_BUF_TILE_RADIANCE)
            assert pay.shape == (counts[rank], 4)
            flat[rank * max_n: rank * max_n + counts[rank]] = torch.from_numpy(pay).cuda()
        out_rad = torch.zeros((W * H, 4), dtype=torch.float32, device="cuda")
        out_rgba = torch.zeros(W * H, dtype=torch.int32, device="cuda")
        gpu_ctx.deinterleave_tiles(flat.data_ptr(), [r * max_n for r in range(world)], world, W, H, tile, out_rad.data_ptr(), out_rgba.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out_rad.cpu().numpy(), full)
        assert np.array_equal(out_rgba.cpu().numpy(), full_rgba)


def test_progressive_accumulation(gpu_ctx):
    """RT_FLAG_ACCUMULATE: accum.rgb = sum of per-frame Lout, accum.w = frame count, RGBA8 = PackRGBA8(accum.rgb / accum.w);
    with rngLockNoise = 0 every frame draws a different stream (RTUtils.cs:122)."""
    sc = orc.Scene()
    sc.build_default()
    gpu_ctx.scene_upload(sc.arrays())
    W, H = 256, 144
    cam = oracle_camera("C1B", W, H)
    total = np.zeros((W * H, 3), np.float32)
    frames = []
    for frame in range(3):
        flags = L.RT_FLAG_ACCUMULATE | (L.RT_FLAG_RESET_ACCUM if frame == 0 else 0)
        gpu_ctx.render(cam, L.make_render_config(W, H, spp=2, max_depth=2, frame=frame, rng_lock_noise=0, flags=flags))
        gpu_ctx.sync()
        lout = gpu_ctx.download(L.RT_BUF_RADIANCE)[:, :3].copy()
        ref = orc.render(sc, cam, orc.make_config(W, H, spp=2, max_depth=2, frame=frame, rng_lock_noise=0), aovs=False)
        assert np.array_equal(lout, ref.radiance)
        frames.append(lout)
        total = total + lout
    assert not np.array_equal(frames[0], frames[1])
    acc = gpu_ctx.download(L.RT_BUF_ACCUM)
    assert np.array_equal(acc[:, :3], total) and np.all(acc[:, 3] == 3.0)
    shown = total * (np.float32(1.0) / acc[:, 3:4])
    want = (255 << 24) | ((np.float32(255.99) * np.clip(shown[:, 0], 0, 1)).astype(np.int64) << 16) | ((np.float32(255.99) * np.clip(shown[:, 1], 0, 1)).astype(np.int64) << 8) | (np.float32(255.99) * np.clip(shown[:, 2], 0, 1)).astype(np.int64)
    assert np.array_equal(gpu_ctx.download(L.RT_BUF_RGBA8).astype(np.int64) & 0xFFFFFFFF, want)
    single_rgba = gpu_ctx.download(L.RT_BUF_RGBA8).copy()

    # C5 protocol: the same progressive sequence tile-partitioned over 4 ranks; after the last frame the gathered
    # payloads (progressive mean per owned pixel) de-interleave to the single-GPU image, bit for bit
    import torch
    from ilgpu_raytracing_b200 import native
    world, tile = 4, 32
    counts = [native.tiles_owned_pixels(W, H, tile, r, world) for r in range(world)]
    max_n = max(counts)
    flat = torch.zeros((world * max_n, 4), dtype=torch.float32, device="cuda")
    ranks = [native.Context(0) for _ in range(world)]   # one context per rank, as in a real multi-GPU run (each keeps its own accumulator)
    for rc in ranks:
        rc.scene_upload(sc.arrays())
    for frame in range(3):
        for rank, rc in enumerate(ranks):
            flags = L.RT_FLAG_ACCUMULATE | (L.RT_FLAG_RESET_ACCUM if frame == 0 else 0)
            rc.render(cam, L.make_render_config(W, H, spp=2, max_depth=2, frame=frame, rng_lock_noise=0, flags=flags, tile_size=tile, rank=rank, world_size=world))
            rc.sync()
            if frame == 2:
                pay = rc.download(L.RT_BUF_TILE_RADIANCE)
                assert np.all(pay[:, 3] == 3.0)
                flat[rank * max_n: rank * max_n + counts[rank]] = torch.from_numpy(pay).cuda()
    for rc in ranks:
        rc.close()
    out_rgba = torch.zeros(W * H, dtype=torch.int32, device="cuda")
    gpu_ctx.deinterleave_tiles(flat.data_ptr(), [r * max_n for r in range(world)], world, W, H, tile, None, out_rgba.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(out_rgba.cpu().numpy(), single_rgba)


def test_error_paths(gpu_ctx):
    from ilgpu_raytracing_b200 import native
    fresh = native.Context(0)
    cam = oracle_camera("C1B", 64, 36)
    with pytest.raises(native.RtError) as e:
        fresh.render(cam, L.make_render_config(64, 36))
    assert e.value.status == L.RT_ERR_INVALID_STATE
    fresh.scene_upload({})
    for bad in (L.make_render_config(0, 36), L.make_render_config(64, 36, max_depth=-1), L.make_render_config(64, 36, rank=2, world_size=2)):
        with pytest.raises(native.RtError) as e:
            fresh.render(cam, bad)
        assert e.value.status == L.RT_ERR_INVALID_ARGUMENT
    with pytest.raises(native.RtError) as e:
        fresh.render(cam, L.make_render_config(64, 36, temporal=1))                    # temporal reuse without prevCam
    assert e.value.status == L.RT_ERR_INVALID_ARGUMENT
    with pytest.raises(native.RtError) as e:
        fresh.render(cam, L.make_render_config(64, 36, spatial=1, rank=0, world_size=2))   # reuse needs the neighbours' reservoirs: no tile partition
    assert e.value.status == L.RT_ERR_UNSUPPORTED
    with pytest.raises(native.RtError) as e:
        fresh.download(L.RT_BUF_RESERVOIR)
    assert e.value.status == L.RT_ERR_INVALID_STATE
    with pytest.raises(native.RtError):
        fresh.download(L.RT_BUF_RGBA8)
    fresh.render(cam, L.make_render_config(64, 36))
    with pytest.raises(native.RtError) as e:
        fresh.download(L.RT_BUF_SEG_COUNT)   # AOVs not requested
    assert e.value.status == L.RT_ERR_INVALID_STATE
    sc = orc.Scene()
    sc.build_default()
    a = sc.arrays()
    a["spherePrimIdx"] = a["spherePrimIdx"].copy()
    a["spherePrimIdx"][6:] = 1234
    with pytest.raises(native.RtError) as e:
        fresh.scene_upload(a)
    assert e.value.status == L.RT_ERR_INVALID_ARGUMENT
    fresh.close()


def test_engine_api_drop_in(gpu_ctx):
    """The reference's host surface: RTRenderer ctor (default scene + camera), RenderDirectToPbo into a CUDA 'PBO',
    Framebuffer.DownloadToCpu -> CpuColor / CpuDepth / CpuObjectId."""
    import torch
    from ilgpu_raytracing_b200 import engine
    W, H = 320, 180
    rdr = engine.RTRenderer(0, W, H)
    assert rdr.camera.tobytes() == oracle_camera("C1A", W, H).tobytes()   # CreateCamera + Translate(1,0,-4), RTRenderer.cs:78-79
    rdr.camera = engine.config_camera("C1B", W, H)
    rdr.configure(spp=2, maxDepth=3, rngLockNoise=1, fixedSeed=7)
    pbo = torch.zeros(W * H, dtype=torch.int32, device="cuda")
    rdr.RenderDirectToPbo(pbo.data_ptr(), W, H, 0, 0.016)
    color, depth, objid = rdr.DownloadToCpu()
    sc = orc.Scene()
    sc.build_default()
    ref = orc.render(sc, oracle_camera("C1B", W, H), orc.make_config(W, H, spp=2, max_depth=3, rng_lock_noise=7), aovs=False)
    assert np.array_equal(color, ref.rgba8) and np.array_equal(depth, ref.depth) and np.array_equal(objid, ref.objId)
    assert np.array_equal(pbo.cpu().numpy(), ref.rgba8)
    # the reference's default flow (RTRenderer.cs:43-44,113-116,208-223): trace at round(out * 0.67), TAAU-resolve into the PBO
    rdr.configure(renderScale=0.67, enableTAAU=1)
    st = orc.TaaState(W, H)
    for frame in (1, 2):
        rdr.RenderDirectToPbo(pbo.data_ptr(), W, H, frame, 0.016)
        cfg = rdr.last_config()
        assert (cfg.width, cfg.height) == (214, 121)
        low = orc.render(sc, oracle_camera("C1B", W, H), orc.make_config(214, 121, spp=2, max_depth=3, frame=frame, rng_lock_noise=7), aovs=False)
        assert np.array_equal(pbo.cpu().numpy(), st.resolve(low.rgba8, low.objId, 214, 121)), f"TAAU present, frame {frame}"
    rdr.configure(enableTAAU=0)
    rdr.RenderDirectToPbo(pbo.data_ptr(), W, H, 3, 0.016)   # bilinear upsample (RTRenderer.cs:227-228)
    assert np.array_equal(pbo.cpu().numpy(), orc.bilinear_upsample(low.rgba8, 214, 121, W, H))
    rdr.RenderDirectToPbo(None, W, H, 3, 0.016)             # headless: the presented image stays in the core
    assert np.array_equal(rdr.native.download(L.RT_BUF_PRESENT), orc.bilinear_upsample(low.rgba8, 214, 121, W, H))
    rdr.close()


@pytest.mark.parametrize("temporal,spatial", [(1, 1), (1, 0), (0, 1)])
def test_restir_reuse_sequence(gpu_ctx, temporal, spatial):
    """ReSTIR temporal + spatial reuse (RTRay.cs:475-516) over a 4-frame sequence with a moving camera: the previous frame's
    reservoirs (ping-pong by frame parity, Framebuffer.cs:127-146; zero-initialised instead of the reference's uninitialised
    memory) feed the imports, only the first Lambert vertex of a sample imports and publishes, the last sample's reservoir
    stays.  Radiance, paths and the reservoir buffers themselves must be bit-identical to the oracle, frame after frame."""
    W, H, spp, depth = 320, 180, 3, 3
    sc = orc.Scene()
    sc.build_default()
    gpu_ctx.scene_upload(sc.arrays())
    res = [np.zeros(W * H, orc.RESERVOIR), np.zeros(W * H, orc.RESERVOIR)]   # A, B
    prev = None
    for frame in range(4):
        cam = orc.camera_create(W, H, 60.0, (0.0 + 0.07 * frame, 1.0 + 0.02 * frame, 3.0 - 0.05 * frame), (0.0, 0.5, 0.0))
        orc.camera_bake(cam, W, H)
        if prev is None:
            prev = cam.copy()
        cur_i = frame & 1
        ocfg = orc.make_config(W, H, spp=spp, max_depth=depth, frame=frame, rng_lock_noise=0, temporal=temporal, spatial=spatial)
        r = orc.render(sc, cam, ocfg, prev_cam=prev, res_prev=res[cur_i ^ 1], res_cur=res[cur_i])
        cfg = L.make_render_config(W, H, spp=spp, max_depth=depth, frame=frame, rng_lock_noise=0, temporal=temporal, spatial=spatial,
                                   flags=L.RT_FLAG_PATH_AOVS | (L.RT_FLAG_RESET_RESERVOIRS if frame == 0 else 0),
                                   samples_per_pass=2 if frame == 2 else 0)   # one frame in two passes
        gpu_ctx.render(cam, cfg, prev_cam=prev)
        gpu_ctx.sync()
        prod = download_all(gpu_ctx)
        assert_parity(r, prod, W, H, spp=spp, label=f"reuse t{temporal} s{spatial} frame {frame}")
        got = gpu_ctx.download(L.RT_BUF_RESERVOIR)
        want = res[cur_i]
        for f in ("m", "lightId", "pdf", "w", "wSum"):
            assert np.array_equal(got[f], want[f]), f"frame {frame}: reservoir field {f} differs"
        for f in ("L", "wi"):
            for c in ("X", "Y", "Z"):
                assert np.array_equal(got[f][c], want[f][c]), f"frame {frame}: reservoir field {f}.{c} differs"
        if frame > 0:
            assert int((want["m"] > 9).sum()) > 0   # imports happened (m counts the 9 new candidates + every accepted import)
        prev = cam.copy()


def test_present_chain(gpu_ctx):
    """rt_present = the tail of RenderDirectToPbo (RTRenderer.cs:208-231): TAAU resolve over 3 frames (history, objId disocclusion,
    sRGB curves), bilinear upsample and blit, into the context's buffer and into a caller-owned 'PBO'; bit-exact vs the oracle."""
    import torch
    inW, inH, outW, outH = 214, 121, 320, 180   # round(320 * 0.67), round(180 * 0.67) as RTRenderer.cs:113-116
    sc = orc.Scene()
    sc.build_default()
    gpu_ctx.scene_upload(sc.arrays())
    st = orc.TaaState(outW, outH)
    pbo = torch.zeros(outW * outH, dtype=torch.int32, device="cuda")
    for frame in range(3):
        cam = orc.camera_create(inW, inH, 60.0, (0.05 * frame, 1.0, 3.0), (0.0, 0.5, 0.0))
        r = orc.render(sc, cam, orc.make_config(inW, inH, spp=2, max_depth=3, frame=frame, rng_lock_noise=0), aovs=False)
        gpu_ctx.render(cam, L.make_render_config(inW, inH, spp=2, max_depth=3, frame=frame, rng_lock_noise=0))
        want = st.resolve(r.rgba8, r.objId, inW, inH)
        if frame < 2:
            gpu_ctx.present(outW, outH, taau=True, reset_history=(frame == 0))
            got = gpu_ctx.download(L.RT_BUF_PRESENT)
        else:
            gpu_ctx.present(outW, outH, taau=True, dst_ptr=pbo.data_ptr(), dst_bytes=pbo.numel() * 4)
            gpu_ctx.sync()
            got = pbo.cpu().numpy()
        assert np.array_equal(got, want), f"TAAU frame {frame}: {(got != want).sum()} px differ"
        gpu_ctx.present(outW, outH, taau=False)
        assert np.array_equal(gpu_ctx.download(L.RT_BUF_PRESENT), orc.bilinear_upsample(r.rgba8, inW, inH, outW, outH))
        gpu_ctx.present(inW, inH, taau=False)
        assert np.array_equal(gpu_ctx.download(L.RT_BUF_PRESENT), r.rgba8)   # blit
    from ilgpu_raytracing_b200 import native
    with pytest.raises(native.RtError) as e:
        gpu_ctx.present(outW, outH, taau=True, dst_ptr=pbo.data_ptr(), dst_bytes=16)
    assert e.value.status == L.RT_ERR_INVALID_ARGUMENT


def test_full_size_properties_c3(gpu_ctx):
    """BASELINE config C3 at full size (1 002 528 triangles, 3840x2160, primary only) through size-independent properties:
    idempotence, ray count, every hit id valid and its t reproducing the depth buffer, and an oracle crop."""
    from ilgpu_raytracing_b200 import engine
    spec = scenes.terrain_scene(708, 0)
    e = engine.Scene().load_spec(spec)
    gpu_ctx.scene_upload(e.arrays())
    W, H = 3840, 2160
    cam = engine.config_camera("C3", W, H)
    cfg = L.make_render_config(W, H, spp=1, max_depth=0)
    gpu_ctx.render(cam, cfg); gpu_ctx.sync()
    prim, t, depth, rgba = gpu_ctx.download(L.RT_BUF_PRIM_ID), gpu_ctx.download(L.RT_BUF_PRIMARY_T), gpu_ctx.download(L.RT_BUF_DEPTH), gpu_ctx.download(L.RT_BUF_RGBA8)
    st = gpu_ctx.stats()
    assert st["raysPrimary"] == W * H and st["raysBounce"] == 0 and st["raysShadow"] == 0
    gpu_ctx.render(cam, cfg); gpu_ctx.sync()
    assert np.array_equal(prim, gpu_ctx.download(L.RT_BUF_PRIM_ID)) and np.array_equal(rgba, gpu_ctx.download(L.RT_BUF_RGBA8))
    hit = prim >= 0
    assert 0.3 < hit.mean() < 0.9 and prim.max() < len(spec.mesh.tris)
    assert np.allclose(depth[hit], t[hit], rtol=1e-5)          # |pos - origin| == t for a unit direction
    assert np.all(t[~hit] == np.float32(1e30))
    box = (1700, 900, 1956, 1044)
    sc = oracle_scene_from_spec(spec)
    r = orc.render(sc, cam, orc.make_config(W, H, spp=1, max_depth=0, crop=box))
    from tests.parity import crop
    assert np.array_equal(crop(prim, W, H, box), r.primId) and np.array_equal(crop(t, W, H, box), r.primaryT)


def test_obj_asset_scene_through_engine(gpu_ctx, tmp_path):
    """an OBJ + MTL + TGA asset loaded from disk by the engine mirror (Scene.LoadObjInstance,
    Scene.cs:144-256) next to the default spheres, rendered by the CUDA core, against the oracle fed with the arrays the
    tests' own loader restatement produces (textured, alpha cut-out, two-sided, mirror and glass materials on triangles)."""
    from ilgpu_raytracing_b200 import engine
    from tests import objfiles
    W, H = 384, 216
    obj, images = objfiles.write_assets(str(tmp_path))
    xf = L.affine_trs((0.2, 0.0, -0.3))
    rdr = engine.RTRenderer(0, W, H)
    rdr.scene.Reset()
    rdr.scene.LoadObjInstance(obj, xf, 1.0)
    rdr.scene.RebuildTLAS()
    rdr.Commit()
    cam = engine.create_camera_at(W, H, 55.0, (0.6, 1.7, 4.6), (0.0, 0.8, 0.0))
    rdr.camera = cam
    sc = oracle_scene_from_spec(objfiles.expected_spec(obj, images, 1.0, xf))
    for flags in (0, L.RT_FLAG_TRI_MATERIALS):
        rdr.configure(spp=3, maxDepth=4, rngLockNoise=1, fixedSeed=5, flags=flags)
        rdr.RenderDirectToPbo(None, W, H, 0, 0.016)
        color, depth, objid = rdr.DownloadToCpu()
        ref = orc.render(sc, cam, orc.make_config(W, H, spp=3, max_depth=4, rng_lock_noise=5, flags=flags & 1), aovs=False)
        assert np.array_equal(objid, ref.objId) and np.array_equal(depth, ref.depth), f"flags {flags}"
        assert np.array_equal(color, ref.rgba8), f"flags {flags}"
        assert (objid >= 0).sum() > W * H // 8       # the asset is actually in view
    rdr.close()


def _moved(positions: np.ndarray, step: int) -> np.ndarray:
    """A large, smooth deformation plus jitter: every triangle moves, many leave their old boxes."""
    p = positions.astype(np.float32).copy()
    rs = np.random.RandomState(100 + step)
    p[:, 1] += (1.5 * np.sin(0.35 * p[:, 0] + step) * np.cos(0.27 * p[:, 2] - step)).astype(np.float32)
    p[:, 0] += (0.3 * np.sin(0.5 * p[:, 2])).astype(np.float32)
    p += rs.uniform(-0.05, 0.05, p.shape).astype(np.float32)
    return p


@pytest.mark.parametrize("kind", ["terrain", "special"])
def test_refit_moved_vertices(gpu_ctx, kind):
    """SURVEY §8f rank 3 (refit half): rt_scene_refit moves the vertices of the uploaded topology and refits the wide BVH on the
    device.  The refitted scene must render exactly like (a) the oracle on a scene REBUILT from the moved vertices and (b) a full
    rt_scene_upload of that scene - hits, counts, G-buffer, radiance, RGBA8 - twice in a row (refit of a refitted tree)."""
    import dataclasses
    from tests.util import special_camera, special_scene
    if kind == "terrain":
        spec, W, H = scenes.terrain_scene(n_quads=96, n_spheres=24), 448, 252
        cam = oracle_camera("C3", W, H)
    else:   # translated (PRIM_XFORM) mesh instance with textured / alpha-tested triangles next to transformed spheres.  The 20 exactly
        # duplicated triangles are left out: their equal-t ties are broken by the visiting order of the BVH2 that was UPLOADED,
        # which a refit keeps, while a rebuilt reference re-sorts them (rt_scene_refit's documented difference)
        spec, W, H = special_scene("translated"), 400, 240
        m = spec.mesh
        spec = dataclasses.replace(spec, mesh=dataclasses.replace(m, tris=m.tris[:-20], tri_uvs=m.tri_uvs[:-20], tri_mat=m.tri_mat[:-20]))
        cam = special_camera(W, H)
    gpu_ctx.scene_upload(oracle_scene_from_spec(spec).arrays())
    pos = spec.mesh.positions
    for step in (1, 2):
        pos = _moved(pos, step) if kind == "terrain" else (pos + np.float32(0.15 * step) * np.sin(pos[:, ::-1] * 3.0 + step).astype(np.float32)).astype(np.float32)
        moved = dataclasses.replace(spec, mesh=dataclasses.replace(spec.mesh, positions=pos))
        sc = oracle_scene_from_spec(moved)
        gpu_ctx.scene_refit(pos)
        _, refit, st_refit = _run(gpu_ctx, sc, cam, W, H, 2, 4, label=f"refit {kind} step {step}")
        gpu_ctx.scene_upload(sc.arrays())
        _, full, st_full = _run(gpu_ctx, sc, cam, W, H, 2, 4, label=f"rebuilt {kind} step {step}")
        for k in full:
            assert np.array_equal(refit[k], full[k]), (k, step)
        assert st_refit["wideNodes"] >= 0 and st_refit["raysBounce"] == st_full["raysBounce"]
        if step == 1:   # the next refit starts from the refitted tree, not from the rebuilt one
            gpu_ctx.scene_upload(oracle_scene_from_spec(spec).arrays())
            gpu_ctx.scene_refit(pos)


def test_refit_errors_and_engine_policy(gpu_ctx):
    """rt_scene_refit error behaviour, and the engine mirror's RebuildPolicy: ForceRefit refits when only positions changed
    (Scene.SetMeshPositions) and falls back to the full upload after any other edit."""
    from ilgpu_raytracing_b200 import engine, native
    spec = scenes.terrain_scene(n_quads=40, n_spheres=6)
    gpu_ctx.scene_upload(oracle_scene_from_spec(spec).arrays())
    n = len(spec.mesh.positions)
    with pytest.raises(native.RtError) as e:
        gpu_ctx.scene_refit(spec.mesh.positions[: n - 1])
    assert e.value.status == L.RT_ERR_INVALID_ARGUMENT
    bad = spec.mesh.positions.copy(); bad[7, 1] = np.nan
    with pytest.raises(native.RtError) as e:
        gpu_ctx.scene_refit(bad)
    assert e.value.status == L.RT_ERR_INVALID_ARGUMENT
    fresh = native.Context(0)
    with pytest.raises(native.RtError) as e:
        fresh.scene_refit(spec.mesh.positions)
    assert e.value.status == L.RT_ERR_INVALID_STATE
    fresh.close()

    W, H = 320, 180
    rdr = engine.RTRenderer(0, W, H)
    rdr.scene.load_spec(spec)
    assert not rdr.scene.CanRefit()
    rdr.Commit(rdr.FORCE_REFIT)                      # nothing uploaded with this topology yet: full upload
    assert rdr.scene.CanRefit()
    cam = engine.config_camera("C3", W, H)
    rdr.camera = cam
    rdr.configure(spp=2, maxDepth=3, rngLockNoise=1, fixedSeed=3)
    pos = _moved(spec.mesh.positions, 5)
    rdr.scene.SetMeshPositions(pos)
    assert rdr.scene.CanRefit()
    rdr.Commit(rdr.FORCE_REFIT)                      # device-side refit
    rdr.RenderDirectToPbo(None, W, H, 0, 0.0)
    color, depth, objid = rdr.DownloadToCpu()
    import dataclasses
    sc = oracle_scene_from_spec(dataclasses.replace(spec, mesh=dataclasses.replace(spec.mesh, positions=pos)))
    ref = orc.render(sc, oracle_camera("C3", W, H), orc.make_config(W, H, spp=2, max_depth=3, rng_lock_noise=3), aovs=False)
    assert np.array_equal(color, ref.rgba8) and np.array_equal(depth, ref.depth) and np.array_equal(objid, ref.objId)
    rdr.scene.SetDeviceBuild(True)                   # the same moved scene, rebuilt on the device: same image
    rdr.Commit(rdr.FORCE_REBUILD)
    rdr.RenderDirectToPbo(None, W, H, 0, 0.0)
    color2, depth2, objid2 = rdr.DownloadToCpu()
    assert np.array_equal(color2, ref.rgba8) and np.array_equal(depth2, ref.depth) and np.array_equal(objid2, ref.objId)
    with pytest.raises(engine.EngineError, match="ArgumentOutOfRangeException"):
        rdr.scene.SetMeshPositions(pos[:-1])
    rdr.scene.AddSphere(scenes.sphere((0, 30, 0), 1.0, (1, 1, 1), 0))   # any other edit ends refit eligibility
    assert not rdr.scene.CanRefit()
    rdr.close()


@pytest.mark.parametrize("kind", ["terrain", "special", "spheres", "terrain_mats"])
def test_device_built_bvh(gpu_ctx, kind):
    """SURVEY §8f rank 3 (build half): rt_scene_upload_ex(RT_BUILD_DEVICE_LBVH) builds the wide BVH on the GPU (Morton order,
    radix tree and PLOC, the one with the cheaper SAH-optimal 8-wide collapse is kept).  Another tree, the same answers: every output equals the oracle's, including the equal-t
    ties of duplicated triangles (the visiting-order ranks are made by the same host stage), and a refit of the device-built
    tree works like a refit of the host-built one."""
    import dataclasses
    from tests.util import special_camera, special_scene
    flags = 0
    if kind == "terrain":
        spec, W, H, cam = scenes.terrain_scene(n_quads=128, n_spheres=30), 448, 252, oracle_camera("C3", 448, 252)
    elif kind == "terrain_mats":
        spec, W, H, cam, flags = scenes.terrain_scene(n_quads=96, n_spheres=0, patch_materials=True), 384, 216, oracle_camera("C3", 384, 216), L.RT_FLAG_TRI_MATERIALS
    elif kind == "special":
        spec, W, H, cam = special_scene("translated"), 400, 240, special_camera(400, 240)
    else:
        spec, W, H, cam = scenes.sphere_grid_scene(24), 384, 216, oracle_camera("C2", 384, 216)
    sc = oracle_scene_from_spec(spec)
    gpu_ctx.scene_upload(sc.arrays(), device_build=True)
    st = gpu_ctx.stats()
    assert st["bvhPrimCount"] > 64 and 0 < st["bvhWideNodeCount"] < st["bvhPrimCount"]
    _run(gpu_ctx, sc, cam, W, H, 1, 0, flags=flags, label=f"device build {kind} primary")
    _run(gpu_ctx, sc, cam, W, H, 3, 6, flags=flags, label=f"device build {kind} 6 bounces")
    if kind == "terrain":
        pos = _moved(spec.mesh.positions, 3)
        gpu_ctx.scene_refit(pos)
        moved = oracle_scene_from_spec(dataclasses.replace(spec, mesh=dataclasses.replace(spec.mesh, positions=pos)))
        _run(gpu_ctx, moved, cam, W, H, 2, 4, label="refit of a device-built tree")
    with pytest.raises(Exception):
        from ilgpu_raytracing_b200 import native
        desc, keep = L.scene_desc_from_arrays(sc.arrays())
        native.check(gpu_ctx._l.rt_scene_upload_ex(gpu_ctx.h, __import__("ctypes").byref(desc), 0x80))   # unknown build flag


@pytest.mark.parametrize("tree", ["radix", "ploc"])
def test_device_built_bvh_both_trees(gpu_ctx, tree, monkeypatch):
    """The device builder makes two binary trees and keeps the cheaper one; RT_DEVICE_TREE (read in rt_create) forces either, so
    both are held to the oracle - on a height field with spheres and on the scene with duplicated triangles and instances -
    whichever of them the cost comparison happens to pick in test_device_built_bvh."""
    from ilgpu_raytracing_b200 import native
    from tests.util import special_camera, special_scene
    monkeypatch.setenv("RT_DEVICE_TREE", tree)
    ctx = native.Context(0)
    try:
        for spec, W, H, cam in ((scenes.terrain_scene(n_quads=150, n_spheres=40), 448, 252, oracle_camera("C3", 448, 252)),
                                (special_scene("translated"), 400, 240, special_camera(400, 240))):
            sc = oracle_scene_from_spec(spec)
            ctx.scene_upload(sc.arrays(), device_build=True)
            ctx.scene_upload(sc.arrays(), device_build=True)   # the second commit reuses the builder's scratch
            st = ctx.stats()
            assert st["bvhPrimCount"] > 64 and 0 < st["bvhWideNodeCount"] < st["bvhPrimCount"]
            _run(ctx, sc, cam, W, H, 1, 0, label=f"device build ({tree}) primary")
            _run(ctx, sc, cam, W, H, 3, 6, label=f"device build ({tree}) 6 bounces")
    finally:
        ctx.close()


def test_device_build_of_a_tiny_scene_takes_the_host_builder(gpu_ctx):
    """Fewer than 64 primitives are not worth a device build: the commit quietly uses the host builder and the images stay the oracle's."""
    sc = oracle_scene_from_spec(scenes.default_scene())
    gpu_ctx.scene_upload(sc.arrays(), device_build=True)
    assert 0 < gpu_ctx.stats()["bvhPrimCount"] < 64
    _run(gpu_ctx, sc, oracle_camera("C1B", 320, 180), 320, 180, 2, 3, label="tiny scene, device build requested")


def test_scene_commits_do_not_leak_device_memory(gpu_ctx):
    """Repeated host builds, device builds and refits of the same scene leave the free device memory where it was."""
    import torch
    spec = scenes.terrain_scene(n_quads=200, n_spheres=8)
    arrays = oracle_scene_from_spec(spec).arrays()
    for dev in (False, True):
        gpu_ctx.scene_upload(arrays, device_build=dev)
        gpu_ctx.scene_refit(spec.mesh.positions)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for it in range(6):
        gpu_ctx.scene_upload(arrays, device_build=bool(it & 1))
        gpu_ctx.scene_refit(spec.mesh.positions)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 32 << 20, f"device memory shrank by {(free0 - free1) >> 20} MiB over 6 commits"


def test_download_async(gpu_ctx):
    """rt_download_async: several read-backs queued, one rt_sync; same bytes as rt_download; staged buffers are refused."""
    import ctypes as C
    import torch
    from ilgpu_raytracing_b200 import native
    W, H = 160, 90
    sc = orc.Scene()
    sc.build_default()
    gpu_ctx.scene_upload(sc.arrays())
    gpu_ctx.render(oracle_camera("C1B", W, H), L.make_render_config(W, H, spp=2, max_depth=2))
    bufs = {w: torch.empty(W * H, dtype=torch.int32).pin_memory() for w in (L.RT_BUF_RGBA8, L.RT_BUF_DEPTH, L.RT_BUF_OBJID)}
    for which, t in bufs.items():
        native.check(gpu_ctx._l.rt_download_async(gpu_ctx.h, which, C.c_void_p(t.data_ptr()), t.numel() * 4))
    gpu_ctx.sync()
    assert np.array_equal(bufs[L.RT_BUF_RGBA8].numpy(), gpu_ctx.download(L.RT_BUF_RGBA8))
    assert np.array_equal(bufs[L.RT_BUF_DEPTH].numpy().view(np.float32), gpu_ctx.download(L.RT_BUF_DEPTH))
    assert np.array_equal(bufs[L.RT_BUF_OBJID].numpy(), gpu_ctx.download(L.RT_BUF_OBJID))
    with pytest.raises(native.RtError) as e:
        native.check(gpu_ctx._l.rt_download_async(gpu_ctx.h, L.RT_BUF_PRIM_ID, C.c_void_p(bufs[L.RT_BUF_RGBA8].data_ptr()), W * H * 4))
    assert e.value.status == L.RT_ERR_UNSUPPORTED
