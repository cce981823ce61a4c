"""not gpu: the C-ABI library loads and exports every symbol include/*.h declares, the layouts agree, the C++ engine
mirror (Scene / Camera / RTRenderer host logic) restates the reference's builders identically to the oracle, errors
surface like the reference's exceptions, and nothing falls back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ilgpu_raytracing_b200 import build, engine, layouts as L, native, scenes
from oracle import orc
from tests.util import oracle_camera, oracle_scene_from_spec, special_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build_all()


def test_abi_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "rtcore_b200.h")).read()
    declared = set(re.findall(r"RT_API\s+(?:const\s+)?\w+\*?\s+(\w+)\s*\(", hdr))
    assert declared == set(native.EXPORTS) and len(declared) == 29
    lib = native.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rt_abi_version() == 1


def test_layout_sizes():
    assert C.sizeof(L.RtSceneDesc) == 15 * 16
    assert C.sizeof(L.RtRenderConfig) == 112 and C.sizeof(L.RtPresentConfig) == 44 and L.RESERVOIR.itemsize == 44
    assert L.CAMERA.itemsize == 92 and L.SPHERE.itemsize == 80 and L.INSTANCE.itemsize == 144 and L.MATERIAL.itemsize == 44 and L.BVHNODE.itemsize == 44


def test_no_cpu_fallback():
    """Without a CUDA device rt_create must fail loudly (RT_ERR_NO_DEVICE); nothing may render on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(native.RtError) as e:
        native.Context(0)
    assert e.value.status == L.RT_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(engine.EngineError) as e2:
        engine.RTRenderer(0, 64, 64)
    assert e2.value.status == L.RT_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ilgpu_raytracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "from oracle" not in src and "import oracle" not in src and "librt_oracle" not in src and not re.search(r'#include\s+"[^"]*oracle', src), f


def test_cameras_match_oracle():
    for name in scenes.CAMERAS:
        for (w, h) in ((1280, 720), (3840, 2160), (333, 77)):
            assert engine.config_camera(name, w, h).tobytes() == oracle_camera(name, w, h).tobytes()
    cam = engine.create_camera(1280, 720, 60.0)
    assert np.allclose([cam["origin"]["X"], cam["origin"]["Y"], cam["origin"]["Z"]], [0, 1, 3])
    assert abs(float(cam["fovYRadians"]) - np.deg2rad(60)) < 1e-6 and abs(float(cam["aspect"]) - 1280 / 720) < 1e-6
    engine.camera_rotate_yaw_pitch(cam, 30.0, -10.0)
    engine.camera_set_fov(cam, 45.0, 16 / 9)
    f = np.array([cam["forward"]["X"], cam["forward"]["Y"], cam["forward"]["Z"]])
    assert abs(np.linalg.norm(f) - 1) < 1e-5


@pytest.mark.parametrize("kind", ["default", "spheres", "terrain", "terrain_mats", "special"])
def test_engine_builders_match_oracle(kind):
    """Two independent restatements of Scene.cs (BLAS/TLAS builders incl. .NET's introsort, instance records, TLAS) -> identical bytes."""
    if kind == "default":
        e = engine.Scene()
        e.BuildDefaultScene()
        o = orc.Scene()
        o.build_default()
    else:
        spec = {"spheres": lambda: scenes.sphere_grid_scene(16), "terrain": lambda: scenes.terrain_scene(80, 25),
                "terrain_mats": lambda: scenes.terrain_scene(64, 0, True), "special": lambda: special_scene("scaled")}[kind]()
        e = engine.Scene().load_spec(spec)
        o = oracle_scene_from_spec(spec)
    ea, oa = e.arrays(), o.arrays()
    for k in ea:
        assert ea[k].tobytes() == oa[k].tobytes(), k
    assert e.sort_ties() == o.sort_ties()


def test_default_scene_spec_equals_builtin():
    a = engine.Scene()
    a.BuildDefaultScene()
    b = engine.Scene().load_spec(scenes.default_scene())
    aa, bb = a.arrays(), b.arrays()
    for k in aa:
        assert aa[k].tobytes() == bb[k].tobytes(), k


def test_engine_errors_follow_reference_exceptions():
    s = engine.Scene()
    with pytest.raises(engine.EngineError, match="ArgumentOutOfRangeException"):
        s.AddSphereInstance([5])
    with pytest.raises(engine.EngineError, match="ArgumentOutOfRangeException"):
        s.AddSphereInstance([])
    m = scenes.terrain_mesh(4)
    bad = m.tris.copy()
    bad[0, 0] = 10 ** 6
    with pytest.raises(engine.EngineError, match="ArgumentOutOfRangeException"):
        s.LoadMeshInstance(m.positions, bad, m.texcoords, m.tri_uvs, m.tri_mat, m.materials)
    s.LoadMeshInstance(m.positions, m.tris, m.texcoords, m.tri_uvs, m.tri_mat, m.materials)
    with pytest.raises(engine.EngineError, match="InvalidOperationException"):   # reference quirk 2: one mesh per scene
        s.LoadMeshInstance(m.positions, m.tris, m.texcoords, m.tri_uvs, m.tri_mat, m.materials)
    with pytest.raises(engine.EngineError, match="InvalidOperationException"):   # host-only scene has nothing to upload to
        s.UploadAll()


def test_tile_ownership_partitions_the_image():
    for (w, h, t) in ((3840, 2160, 32), (7680, 4320, 32), (200, 104, 32), (33, 17, 8)):
        for world in (1, 2, 4, 8):
            counts = [native.tiles_owned_pixels(w, h, t, r, world) for r in range(world)]
            assert sum(counts) == w * h
            if world > 1 and w * h > 10 ** 6:
                assert max(counts) / min(counts) < 1.02   # interleaving balances the load
    with pytest.raises(native.RtError):
        native.tiles_owned_pixels(64, 64, 32, 3, 2)


def test_terrain_scene_shape():
    """C3/C4 mesh: 708 x 708 quads = 1 002 528 triangles, 709^2 vertices, up-facing."""
    m = scenes.terrain_mesh(708)
    assert len(m.tris) == 1002528 and len(m.positions) == 709 * 709
    p = m.positions[m.tris[:1000]]
    n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    assert (n[:, 1] > 0).all()
    assert abs(m.positions[:, 0]).max() < 50.1 and abs(m.positions[:, 1]).max() < 6.5


def test_obj_loader_matches_independent_restatement(tmp_path):
    """Scene.LoadObjInstance (OBJ + MTL + TGA / BMP from disk) against a Python restatement of MeshLoaderOBJ.cs written for
    the tests: polygons, negative indices, v/vt/vn forms, CRLF, usemtl without MTL entry, MTL-only materials, shared and
    missing textures, every TGA flavour the reference decodes.  Same 15 arrays from the engine mirror and the oracle."""
    from tests import objfiles
    obj, images = objfiles.write_assets(str(tmp_path))
    for scale, xf in ((1.0, None), (0.5, L.affine_trs((0.5, 1.0, -2.0)))):
        spec = objfiles.expected_spec(obj, images, scale, xf)
        assert len(spec.mesh.tris) == 2 + 2 + 3 + 1 + 1 + 1 + 1 and len(spec.mesh.materials) == 9 and len(spec.textures) == 6
        got = engine.Scene()
        got.LoadObjInstance(obj, xf, scale)
        want = engine.Scene().load_spec(spec)
        ga, wa, oa = got.arrays(), want.arrays(), oracle_scene_from_spec(spec).arrays()
        for k in ga:
            assert ga[k].tobytes() == wa[k].tobytes() == oa[k].tobytes(), k
    m = ga["materials"]
    assert list(m["Shading"][:5]) == [L.SHADING_LAMBERT, L.SHADING_LAMBERT, L.SHADING_LAMBERT, L.SHADING_GLASS, L.SHADING_MIRROR]
    assert m["HasAlphaMap"][1] == 1 and m["TwoSided"][1] == 1 and abs(m["IOR"][3] - 1.45) < 1e-6
    assert m["HasDiffuseMap"][8] == 0 and m["DiffuseTexIndex"][8] == -1          # 'lost': texture file absent
    assert m["HasAlphaMap"][6] == 0 and m["HasDiffuseMap"][6] == 1 and m["IOR"][6] == 1.0   # 'shares': FLOOR.TGA == floor.tga, Ni <= 0 -> 1


def test_png_textures_decode_exactly(tmp_path):
    """PNG textures (the reference reads them through System.Drawing, MeshLoaderOBJ.cs:463-502; here csrc/host/png_decode.cpp):
    every colour type and bit depth the decoder takes, both interlace methods, all five filter types, stored / fixed / dynamic
    deflate blocks, split IDAT, tRNS as palette alpha and as a colour key - texels byte-identical to the images the files were
    made from.  Damaged files raise what `new Bitmap(file)` raises; 16-bit samples are refused, not guessed."""
    import zlib
    from tests import objfiles
    obj, images = objfiles.write_png_assets(str(tmp_path))
    spec = objfiles.expected_spec(obj, images)
    assert len(spec.textures) == len(images) == 14
    got = engine.Scene()
    got.LoadObjInstance(obj)
    want = engine.Scene().load_spec(spec)
    ga, wa = got.arrays(), want.arrays()
    for k in ga:
        assert ga[k].tobytes() == wa[k].tobytes(), k
    good = open(tmp_path / "rgba8.png", "rb").read()

    def load(name, data):
        (tmp_path / name).write_bytes(data)
        (tmp_path / (name + ".mtl")).write_text(f"newmtl m0\nmap_Kd {name}\n")
        o = tmp_path / (name + ".obj")
        o.write_text(f"mtllib {name}.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nusemtl m0\nf 1/1 2/1 3/1\n")
        engine.Scene().LoadObjInstance(str(o))

    flipped = bytearray(good); flipped[len(good) // 2] ^= 0x40            # inside IDAT: chunk CRC mismatch
    for name, data in (("crc.png", bytes(flipped)), ("cut.png", good[:len(good) - 20]), ("sig.png", b"\x89PNG...." + good[8:])):
        with pytest.raises(engine.EngineError, match="ArgumentException"):
            load(name, data)
    objfiles.write_png(str(tmp_path / "deep.png"), np.zeros((2, 2, 6), np.int64), color_type=2, depth=8)   # 3 x 16-bit samples = 6 bytes per pixel
    deep = bytearray(open(tmp_path / "deep.png", "rb").read())
    deep[24] = 16                                                           # IHDR bit depth -> 16, CRC patched
    deep[29:33] = (zlib.crc32(bytes(deep[12:29])) & 0xFFFFFFFF).to_bytes(4, "big")
    with pytest.raises(engine.EngineError, match="InvalidDataException"):
        load("deep16.png", bytes(deep))


def test_obj_loader_errors_follow_reference_exceptions(tmp_path):
    from tests import objfiles
    obj, _ = objfiles.write_assets(str(tmp_path))
    s = engine.Scene()
    with pytest.raises(engine.EngineError, match="FileNotFoundException"):
        s.LoadObjInstance(str(tmp_path / "nope.obj"))
    with pytest.raises(engine.EngineError, match="FileNotFoundException"):
        s.LoadObjInstance("   ")
    text = open(obj).read()

    def variant(name, obj_text=None, mtl_text=None):
        p = tmp_path / name
        p.write_text(text if obj_text is None else obj_text)
        if mtl_text is not None:
            (tmp_path / (name + ".mtl")).write_text(mtl_text)
        return str(p)

    with pytest.raises(engine.EngineError, match="FormatException"):                 # float.Parse
        engine.Scene().LoadObjInstance(variant("badnum.obj", text.replace("v  2.0 0.0 -2.0", "v  2.0 zero -2.0")))
    with pytest.raises(engine.EngineError, match="FormatException"):                 # int.Parse("") for 'v/'
        engine.Scene().LoadObjInstance(variant("badface.obj", text.replace("f 14/1 15/2 16/3", "f 14/ 15/2 16/3")))
    with pytest.raises(engine.EngineError, match="InvalidOperationException"):
        engine.Scene().LoadObjInstance(variant("empty.obj", "v 0 0 0\nvt 0 0\nusemtl a\n"))
    with pytest.raises(engine.EngineError, match="ArgumentOutOfRangeException"):     # face index past the vertex list
        engine.Scene().LoadObjInstance(variant("range.obj", text.replace("f 14/1 15/2 16/3", "f 14/1 15/2 99/3")))
    # images: colour-mapped TGA, truncated TGA, a format only System.Drawing decodes
    hdr = bytearray(open(tmp_path / "floor.tga", "rb").read())
    hdr[1] = 1
    (tmp_path / "cmap.tga").write_bytes(bytes(hdr))
    (tmp_path / "short.tga").write_bytes(open(tmp_path / "floor.tga", "rb").read()[:40])
    (tmp_path / "pic.jpg").write_bytes(b"\xff\xd8\xff\xe0....")
    for tex, exc in (("cmap.tga", "InvalidDataException"), ("short.tga", "EndOfStreamException"), ("pic.jpg", "InvalidDataException")):
        o = variant("tex_" + tex + ".obj", text.replace("mtllib assets.mtl", f"mtllib tex_{tex}.obj.mtl"), f"newmtl floor\nmap_Kd {tex}\n")
        with pytest.raises(engine.EngineError, match=exc):
            engine.Scene().LoadObjInstance(o)
    # a second mesh in one scene stays refused (reference quirk 2), after a good first load
    s.LoadObjInstance(obj)
    with pytest.raises(engine.EngineError, match="InvalidOperationException"):
        s.LoadObjInstance(obj)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver times beside ours) runs without a GPU and prints ONE JSON line with the
    contract's keys; the product arm must refuse to run without a CUDA device instead of falling back."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"] == "C4" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    import torch
    if not torch.cuda.is_available():
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"],
                             capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert out.returncode != 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
