"""Test infrastructure: writes a small OBJ + MTL + TGA/BMP asset set and states, independently of the C++ loader
(ilgpu_raytracing_b200/csrc/host/mesh_loader_obj.cpp), what the reference's asset path makes of it:
  MeshLoaderOBJ.Load / LoadMtl / LoadTgaBGRA   Engine/MeshLoaderOBJ.cs:67-254, 319-441, 504-593
  Scene.LoadObjInstance                          Engine/Scene.cs:144-256
The expectation is a SceneSpec (decoded arrays), which both the engine mirror and the oracle already accept."""
import os
import struct

import numpy as np

from ilgpu_raytracing_b200 import layouts as L
from ilgpu_raytracing_b200 import scenes


# ---------------------------------------------------------------------------------------------------------------- writers
def _tga_header(w, h, image_type, depth, top, id_bytes=b""):
    desc = (0x20 if top else 0) | (8 if depth == 32 else 0)
    return struct.pack("<BBBHHBHHHHBB", len(id_bytes), 0, image_type, 0, 0, 0, 0, 0, w, h, depth, desc) + id_bytes


def write_tga(path, bgra_top_down: np.ndarray, depth=32, top=False, rle=False, id_bytes=b""):
    """bgra_top_down: (h, w, 4) u8 as the image should DECODE (rows top-down, B G R A)."""
    h, w, _ = bgra_top_down.shape
    rows = bgra_top_down if top else bgra_top_down[::-1]
    if depth == 32:
        px = rows.reshape(-1, 4)
    elif depth == 24:
        px = rows.reshape(-1, 4)[:, :3]
    else:
        px = rows.reshape(-1, 4)[:, :1]
    with open(path, "wb") as f:
        if not rle:
            f.write(_tga_header(w, h, 3 if depth == 8 else 2, depth, top, id_bytes))
            f.write(np.ascontiguousarray(px).tobytes())
            return
        f.write(_tga_header(w, h, 10, depth, top, id_bytes))
        i, n = 0, len(px)
        while i < n:   # packets may cross scan lines, like the reference's decoder allows
            run = 1
            while i + run < n and run < 128 and np.array_equal(px[i + run], px[i]):
                run += 1
            if run >= 2:
                f.write(bytes([0x80 | (run - 1)]) + px[i].tobytes())
                i += run
            else:
                lit = 1
                while i + lit < n and lit < 128 and not (i + lit + 1 < n and np.array_equal(px[i + lit], px[i + lit + 1])):
                    lit += 1
                f.write(bytes([lit - 1]) + np.ascontiguousarray(px[i:i + lit]).tobytes())
                i += lit


def write_bmp24(path, bgra_top_down: np.ndarray):
    h, w, _ = bgra_top_down.shape
    stride = (w * 3 + 3) & ~3
    data = bytearray()
    for y in range(h - 1, -1, -1):
        row = bgra_top_down[y, :, :3].tobytes()
        data += row + b"\0" * (stride - len(row))
    with open(path, "wb") as f:
        f.write(b"BM" + struct.pack("<IHHI", 54 + len(data), 0, 0, 54))
        f.write(struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(data), 2835, 2835, 0, 0))
        f.write(bytes(data))


# ---- PNG writer for the loader tests (every colour type / bit depth the native decoder takes, both interlace methods, every
# filter type, stored / fixed / dynamic deflate blocks) -----------------------------------------------------------------------
def _png_chunk(tag: bytes, data: bytes) -> bytes:
    import zlib
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def _png_filter_rows(rows: list, bpp: int) -> bytes:
    """rows: list of bytes (packed scanlines); filter type cycles 0..4 over the rows."""
    out, prev = bytearray(), None
    for y, row in enumerate(rows):
        f = y % 5
        prev_row = prev if prev is not None else bytes(len(row))
        cur = bytearray(len(row))
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev_row[i]
            c = prev_row[i - bpp] if i >= bpp else 0
            if f == 0:
                pred = 0
            elif f == 1:
                pred = a
            elif f == 2:
                pred = b
            elif f == 3:
                pred = (a + b) >> 1
            else:
                pq = a + b - c
                pa, pb, pc = abs(pq - a), abs(pq - b), abs(pq - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            cur[i] = (v - pred) & 0xFF
        out.append(f)
        out += cur
        prev = row
    return bytes(out)


def _png_pack(samples: np.ndarray, depth: int) -> list:
    """(h, w, channels) sample values -> packed scanlines (most significant bits first)."""
    h, w, ch = samples.shape
    rows = []
    for y in range(h):
        flat = samples[y].reshape(-1)
        if depth == 8:
            rows.append(bytes(int(v) for v in flat))
        else:
            bits = "".join(format(int(v), "0%db" % depth) for v in flat)
            bits += "0" * (-len(bits) % 8)
            rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
    return rows


def write_png(path, samples: np.ndarray, color_type: int, depth: int = 8, interlace: bool = False, palette=None, trns: bytes = b"",
              zmode: str = "dynamic", idat_split: int = 0, extra_chunks=()):
    import zlib
    h, w, ch = samples.shape
    bpp = max(1, ch * depth // 8)
    raw = b""
    if interlace:
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = samples[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                raw += _png_filter_rows(_png_pack(sub, depth), bpp)
    else:
        raw = _png_filter_rows(_png_pack(samples, depth), bpp)
    if zmode == "stored":
        z = zlib.compress(raw, 0)
    elif zmode == "fixed":
        co = zlib.compressobj(9, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)
        z = co.compress(raw) + co.flush()
    else:
        z = zlib.compress(raw, 9)
    out = b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, 1 if interlace else 0))
    for tag, data in extra_chunks:
        out += _png_chunk(tag, data)
    if palette is not None:
        out += _png_chunk(b"PLTE", bytes(int(v) for v in np.asarray(palette, np.uint8).reshape(-1)))
    if trns:
        out += _png_chunk(b"tRNS", trns)
    if idat_split:
        for i in range(0, len(z), idat_split):
            out += _png_chunk(b"IDAT", z[i:i + idat_split])
    else:
        out += _png_chunk(b"IDAT", z)
    out += _png_chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(out)


def write_png_assets(dirpath: str, seed: int = 23) -> tuple:
    """One material per PNG flavour.  Returns (obj path, {file name: decoded (h,w,4) BGRA top-down})."""
    rs = np.random.RandomState(seed)
    images, names = {}, []

    def add(name, bgra, **kw):
        images[name] = bgra.astype(np.uint8)
        names.append(name)
        write_png(os.path.join(dirpath, name), **kw)

    def bgra_from(r, g, b, a):
        return np.stack([b, g, r, a], -1)

    s = rs.randint(0, 256, (9, 13, 4))
    add("rgba8.png", bgra_from(s[..., 0], s[..., 1], s[..., 2], s[..., 3]), samples=s, color_type=6, extra_chunks=((b"gAMA", struct.pack(">I", 45455)), (b"tEXt", b"Comment\0made by tests")))
    s = rs.randint(0, 256, (7, 5, 3))
    add("rgb8_fixed.png", bgra_from(s[..., 0], s[..., 1], s[..., 2], np.full(s.shape[:2], 255)), samples=s, color_type=2, zmode="fixed")
    s = rs.randint(0, 256, (20, 17, 4))
    add("rgba8_adam7.png", bgra_from(s[..., 0], s[..., 1], s[..., 2], s[..., 3]), samples=s, color_type=6, interlace=True, idat_split=97)
    s = rs.randint(0, 256, (6, 11, 1))
    add("grey8_stored.png", bgra_from(s[..., 0], s[..., 0], s[..., 0], np.full(s.shape[:2], 255)), samples=s, color_type=0, zmode="stored")
    for d in (1, 2, 4):
        s = rs.randint(0, 1 << d, (10, 19, 1))
        g = s[..., 0] * 255 // ((1 << d) - 1)
        add("grey%d.png" % d, bgra_from(g, g, g, np.full(g.shape, 255)), samples=s, color_type=0, depth=d)
    s = rs.randint(0, 256, (5, 8, 2))
    add("greyalpha8.png", bgra_from(s[..., 0], s[..., 0], s[..., 0], s[..., 1]), samples=s, color_type=4)
    for d, inter in ((8, False), (4, True), (2, False), (1, False)):
        n = min(1 << d, 40)
        pal = rs.randint(0, 256, (n, 3))
        tr = rs.randint(0, 256, n // 2)   # tRNS shorter than the palette: the rest is opaque
        s = rs.randint(0, n, (12, 21, 1))
        idx = s[..., 0]
        alpha = np.where(idx < len(tr), np.concatenate([tr, np.full(n - len(tr), 255)])[idx], 255)
        add("pal%d%s.png" % (d, "_adam7" if inter else ""), bgra_from(pal[idx, 0], pal[idx, 1], pal[idx, 2], alpha), samples=s, color_type=3, depth=d, interlace=inter,
            palette=pal, trns=bytes(int(v) for v in tr))
    s = rs.randint(0, 4, (6, 6, 3)) * 85
    key = s[2, 3].copy()
    a = np.where((s == key).all(-1), 0, 255)
    add("rgb8_key.png", bgra_from(s[..., 0], s[..., 1], s[..., 2], a), samples=s, color_type=2, trns=struct.pack(">HHH", int(key[0]), int(key[1]), int(key[2])))
    s = rs.randint(0, 4, (4, 9, 1))
    g = s[..., 0] * 85
    add("grey2_key.png", bgra_from(g, g, g, np.where(s[..., 0] == 2, 0, 255)), samples=s, color_type=0, depth=2, trns=struct.pack(">H", 2))
    obj_lines, mtl_lines = ["mtllib png.mtl", "v 0 0 0", "v 1 0 0", "v 0 1 0", "vt 0 0", "vt 1 0", "vt 0 1"], []
    for k, name in enumerate(names):
        obj_lines += ["usemtl m%d" % k, "f 1/1 2/2 3/3"]
        mtl_lines += ["newmtl m%d" % k, "map_Kd " + name]
    with open(os.path.join(dirpath, "png.mtl"), "w", newline="") as f:
        f.write("\n".join(mtl_lines) + "\n")
    obj = os.path.join(dirpath, "png.obj")
    with open(obj, "w", newline="") as f:
        f.write("\n".join(obj_lines) + "\n")
    return obj, images


def _image(rs, w, h, alpha=None):
    img = rs.randint(0, 256, (h, w, 4)).astype(np.uint8)
    img[:, : w // 2] = img[:1, :1]          # flat areas, so the RLE writer emits run packets as well as literal ones
    if alpha is not None:
        img[..., 3] = alpha
    return img


OBJ_TEXT = """# a small asset: quads, a pentagon, negative indices, v/vt, v/vt/vn and v//vn corners, CRLF and comment lines
mtllib assets.mtl
o plate
v -2.0 0.0 -2.0
v  2.0 0.0 -2.0
v  2.0 0.0  2.0
v -2.0 0.0  2.0
vt 0.0 0.0
vt 3.0 0.0
vt 3.0 3.0
vt 0.0 3.0
vn 0 1 0
usemtl floor
f 1/1/1 4/4/1 3/3/1 2/2/1
v -1.0 0.0 -0.5
v  1.0 0.0 -0.5
v  1.0 2.0 -0.5
v -1.0 2.0 -0.5
vt 1 0
vt 1 1
usemtl leaf
f -4/1 -3/5 -2/6 -1/4\r
usemtl notInMtl
v 0.5 0.2 0.8
v 1.5 0.2 0.8
v 1.8 1.0 0.8
v 1.0 1.6 0.8
v 0.2 1.0 0.8
f 9/1 10/2 11/3 12/4 13/5
usemtl glassy
f 5//1 6//1 7//1
usemtl mirrory
v -1.9 0.1 1.0
v -0.9 0.1 1.0
v -1.4 1.3 0.6
f 14/1 15/2 16/3
usemtl floor
f 1/1 2/2 6/5
usemtl grey
f 14/5 16/6 4/4
f 1 2
"""

MTL_TEXT = """# materials
newmtl floor
Kd 0.9 0.8 0.7
map_Kd floor.tga
illum 2
newmtl leaf
Kd 0.2 0.9 0.3
map_Kd leaf_rgb.tga
map_d Leaf_A.tga
d 0.5
newmtl glassy
Kd 1 1 1
Ni 1.45 extra
illum 7
newmtl mirrory
Kd 0.95 0.95 0.95
illum 3
Tr 0.0
newmtl grey
Kd 0.5 0.5 0.5
map_Kd grey8.tga
newmtl shares
map_Kd FLOOR.TGA
map_d missing_alpha.tga
Ni -2
newmtl bmpmat
map_Kd photo.bmp
newmtl lost
Kd 0.1 0.2 0.3
map_Kd not_there.tga
"""


def write_assets(dirpath: str, seed: int = 11) -> tuple:
    """Returns (obj path, {file name: decoded (h,w,4) BGRA top-down})."""
    rs = np.random.RandomState(seed)
    images = {
        "floor.tga": _image(rs, 8, 4),                                                   # 32-bit, bottom-left origin
        "leaf_rgb.tga": _image(rs, 5, 7, alpha=255),                                      # 24-bit, top-left origin, image id field
        "Leaf_A.tga": _image(rs, 16, 16),                                                 # 32-bit RLE, bottom-left origin
        "grey8.tga": np.repeat(rs.randint(0, 256, (6, 3, 1)).astype(np.uint8), 4, axis=2),   # 8-bit greyscale
        "photo.bmp": _image(rs, 7, 3, alpha=255),
    }
    images["Leaf_A.tga"][..., 3] = np.where(rs.rand(16, 16) < 0.5, 0, 255)   # cut-out alpha
    images["grey8.tga"][..., 3] = 255
    write_tga(os.path.join(dirpath, "floor.tga"), images["floor.tga"], 32, top=False)
    write_tga(os.path.join(dirpath, "leaf_rgb.tga"), images["leaf_rgb.tga"], 24, top=True, id_bytes=b"made by tests")
    write_tga(os.path.join(dirpath, "Leaf_A.tga"), images["Leaf_A.tga"], 32, top=False, rle=True)
    write_tga(os.path.join(dirpath, "grey8.tga"), images["grey8.tga"], 8, top=True)
    write_bmp24(os.path.join(dirpath, "photo.bmp"), images["photo.bmp"])
    with open(os.path.join(dirpath, "assets.mtl"), "w", newline="") as f:
        f.write(MTL_TEXT)
    obj = os.path.join(dirpath, "assets.obj")
    with open(obj, "w", newline="") as f:
        f.write(OBJ_TEXT)
    return obj, images


# ---------------------------------------------------------------------------------------------------------------- expectation
def _default_mat():
    return dict(Kd=(0.8, 0.8, 0.8), HasDiffuseMap=0, DiffuseTexIndex=-1, Shading=L.SHADING_LAMBERT, IOR=1.0, HasAlphaMap=0, AlphaTexIndex=-1,
                TwoSided=0, AlphaCutoff=0.5)


def _index(tok, count):
    v = int(tok)
    return v - 1 if v > 0 else count + v


def expected_spec(obj_path: str, images: dict, scale: float = 1.0, object_to_world=None) -> scenes.SceneSpec:
    """What Scene.LoadObjInstance(objPath, objectToWorld, scale) appends to an EMPTY scene, as a SceneSpec."""
    base = os.path.dirname(obj_path)
    pos, uvs, tris, tuvs, tmat = [], [], [], [], []
    names, mats, cur, mtllib = {}, [], -1, None
    for line in open(obj_path, newline="").read().replace("\r\n", "\n").replace("\r", "\n").split("\n"):
        if not line or line[0] == "#":
            continue
        if line.startswith("v "):
            pos.append([np.float32(t) * np.float32(scale) for t in line[2:].split()[:3]])
        elif line.startswith("vt "):
            uvs.append([np.float32(t) for t in line[3:].split()[:2]])
        elif line.startswith("f "):
            fv, ft = [], []
            for tok in line[2:].split(" "):
                if not tok.strip():
                    continue
                parts = tok.strip().split("/")
                fv.append(_index(parts[0], len(pos)))
                ft.append(_index(parts[1], len(uvs)) if len(parts) > 1 and parts[1] else 0)
            for k in range(1, len(fv) - 1):     # fan, winding kept (LoadObjInstance passes flipWinding: false)
                tris.append((fv[0], fv[k], fv[k + 1]))
                tuvs.append((ft[0], ft[k], ft[k + 1]))
                tmat.append(max(cur, 0))
        elif line.startswith("mtllib "):
            mtllib = os.path.join(base, line[7:].strip())
        elif line.startswith("usemtl "):
            n = line[7:].strip()
            if n not in names:
                names[n] = len(mats)
                mats.append(_default_mat())
            cur = names[n]
    diffuse, alpha = {}, {}
    if mtllib and os.path.exists(mtllib):
        loaded, dmap, amap, name, m = {}, {}, {}, None, None
        for line in open(mtllib).read().split("\n"):
            if not line or line[0] == "#":
                continue
            if line.startswith("newmtl "):
                if name is not None:
                    loaded[name] = m
                name, m = line[7:].strip(), _default_mat()
            elif line.startswith("Kd "):
                m["Kd"] = tuple(np.float32(t) for t in line[3:].split()[:3])
            elif line.startswith("map_Kd "):
                dmap[name] = os.path.join(base, line[7:].strip())
                m["HasDiffuseMap"] = 1
            elif line.startswith("map_d "):
                amap[name] = os.path.join(base, line[6:].strip())
                m["HasAlphaMap"], m["TwoSided"] = 1, 1
            elif line.startswith("d "):
                if np.float32(line[2:]) < np.float32(0.999):
                    m["TwoSided"], m["AlphaCutoff"] = 1, 0.5
            elif line.startswith("Tr "):
                if np.float32(1.0) - np.float32(line[3:]) < np.float32(0.999):
                    m["TwoSided"], m["AlphaCutoff"] = 1, 0.5
            elif line.startswith("Ni "):
                ior = np.float32(line[3:].split()[0])
                m["IOR"] = float(ior) if ior > 0 else 1.0
            elif line.startswith("illum "):
                k = int(line[6:])
                m["Shading"] = L.SHADING_GLASS if k >= 5 else (L.SHADING_MIRROR if k >= 3 else L.SHADING_LAMBERT)
        if name is not None:
            loaded[name] = m
        for n, m in loaded.items():
            if n in names:
                mats[names[n]] = m
            else:
                names[n] = len(mats)
                mats.append(m)
        diffuse = {names[n]: p for n, p in dmap.items() if n in names}
        alpha = {names[n]: p for n, p in amap.items() if n in names}
    # loader-level textures: one per distinct path ignoring case; a missing file clears the material's flag
    loader_tex, by_path = [], {}
    for table, has, idx in ((diffuse, "HasDiffuseMap", "DiffuseTexIndex"), (alpha, "HasAlphaMap", "AlphaTexIndex")):
        for mi, p in table.items():
            if p.lower() not in by_path:
                if not os.path.exists(p):
                    mats[mi][has], mats[mi][idx] = 0, -1
                    continue
                by_path[p.lower()] = len(loader_tex)
                loader_tex.append(images[os.path.basename(p)])
            mats[mi][has], mats[mi][idx] = 1, by_path[p.lower()]
            if has == "HasAlphaMap":
                mats[mi]["TwoSided"] = 1
    # scene-level flattening: every material appends its own copy, diffuse before alpha; RGBA32 memory order is R G B A
    textures, out = [], np.zeros(len(mats), L.MATERIAL)
    for i, m in enumerate(mats):
        for has, idx in (("HasDiffuseMap", "DiffuseTexIndex"), ("HasAlphaMap", "AlphaTexIndex")):
            if m[has] and 0 <= m[idx] < len(loader_tex):
                textures.append(np.ascontiguousarray(loader_tex[m[idx]][..., [2, 1, 0, 3]]))
                m[has], m[idx] = 1, len(textures) - 1
            else:
                m[has], m[idx] = 0, -1
        for k, v in m.items():
            out[i][k] = v
    mesh = scenes.MeshSpec(positions=np.array(pos, np.float32).reshape(-1, 3), tris=np.array(tris, np.int32).reshape(-1, 3),
                           texcoords=np.array(uvs, np.float32).reshape(-1, 2), tri_uvs=np.array(tuvs, np.int32).reshape(-1, 3),
                           tri_mat=np.array(tmat, np.int32), materials=out,
                           object_to_world=L.affine_identity() if object_to_world is None else object_to_world)
    return scenes.SceneSpec(textures=textures, mesh=mesh)
