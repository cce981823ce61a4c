"""Parity comparison between the oracle's RenderResult and the product's downloaded buffers.

Criteria (BASELINE.json north_star): hit primitive IDs and bounce counts bit-exact on non-degenerate
rays; accumulated radiance within 1e-4 relative RMS per pixel.  Because the product evaluates the
reference's arithmetic exactly, the tests below demand bit equality of everything and only fall back
to the "degenerate ray" proviso where the oracle's own box culling changed the answer (checked by
re-tracing the offending primary ray culling-free through the oracle).
"""
from __future__ import annotations

import numpy as np

from ilgpu_raytracing_b200 import layouts as L


def download_all(ctx, aovs=True) -> dict:
    d = dict(rgba8=ctx.download(L.RT_BUF_RGBA8), depth=ctx.download(L.RT_BUF_DEPTH), objId=ctx.download(L.RT_BUF_OBJID),
             radiance=ctx.download(L.RT_BUF_RADIANCE)[:, :3], primId=ctx.download(L.RT_BUF_PRIM_ID), instId=ctx.download(L.RT_BUF_INST_ID),
             primaryT=ctx.download(L.RT_BUF_PRIMARY_T), gbPos=ctx.download(L.RT_BUF_GB_WORLDPOS), gbNrm=ctx.download(L.RT_BUF_GB_NORMAL),
             gbAlb=ctx.download(L.RT_BUF_GB_BASECOLOR), gbMat=ctx.download(L.RT_BUF_GB_MATID))
    if aovs:
        d["segCount"] = ctx.download(L.RT_BUF_SEG_COUNT)
        d["termCode"] = ctx.download(L.RT_BUF_TERM_CODE)
        d["pathHash"] = ctx.download(L.RT_BUF_PATH_HASH)
    return d


def crop(a: np.ndarray, W: int, H: int, box, planes: int | None = None) -> np.ndarray:
    """Cut the crop window out of a full-frame product buffer so it lines up with a cropped oracle render."""
    x0, y0, x1, y1 = box
    if planes is not None:
        return a.reshape(planes, H, W)[:, y0:y1, x0:x1].reshape(planes, -1)
    if a.ndim == 2:
        return a.reshape(H, W, a.shape[1])[y0:y1, x0:x1].reshape(-1, a.shape[1])
    return a.reshape(H, W)[y0:y1, x0:x1].reshape(-1)


def rel_rms(a: np.ndarray, b: np.ndarray) -> float:
    """Relative RMS error per pixel, averaged: sqrt(mean(|a-b|^2)) / sqrt(mean(|b|^2))."""
    num = np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    den = np.sqrt(np.mean(b.astype(np.float64) ** 2)) + 1e-30
    return float(num / den)


def assert_parity(oracle_result, prod: dict, W: int, H: int, box=None, spp=1, allow_degenerate=0, label=""):
    box = box or (0, 0, W, H)
    full = box == (0, 0, W, H)
    get = (lambda k: prod[k]) if full else (lambda k: crop(prod[k], W, H, box))
    r = oracle_result
    bad_prim = (r.primId != get("primId")) | (r.instId != get("instId"))
    nbad = int(bad_prim.sum())
    assert nbad <= allow_degenerate, f"{label}: {nbad} primary hit-id mismatches (allowed {allow_degenerate})"
    ok = ~bad_prim
    assert np.array_equal(r.primaryT[ok], get("primaryT")[ok]), f"{label}: primary t differs"
    for k in ("gbPos", "gbNrm", "gbAlb"):
        assert np.array_equal(getattr(r, k)[ok], get(k)[ok]), f"{label}: G-buffer {k} differs"
    assert np.array_equal(r.gbMat[ok], get("gbMat")[ok]), f"{label}: G-buffer matId differs"
    assert np.array_equal(r.objId[ok], get("objId")[ok]) and np.array_equal(r.depth[ok], get("depth")[ok]), f"{label}: objId/depth differ"
    if "segCount" in prod and r.segCount.size:
        n = max(1, spp)
        seg = prod["segCount"].reshape(n, -1) if full else crop(prod["segCount"], W, H, box, planes=n)
        term = prod["termCode"].reshape(n, -1) if full else crop(prod["termCode"], W, H, box, planes=n)
        hsh = prod["pathHash"].reshape(n, -1) if full else crop(prod["pathHash"], W, H, box, planes=n)
        m = np.broadcast_to(ok, seg.shape)
        bad_paths = int(((r.segCount != seg) | (r.termCode != term) | (r.pathHash != hsh))[m].sum())
        assert bad_paths <= allow_degenerate, f"{label}: {bad_paths} paths differ in bounce count / terminator / hit-id hash"
        if bad_paths == 0:
            assert np.array_equal(r.radiance[ok], get("radiance")[ok]), f"{label}: radiance not bit-identical although every path matches"
            assert np.array_equal(r.rgba8[ok], get("rgba8")[ok]), f"{label}: RGBA8 differs"
    e = rel_rms(get("radiance")[ok], r.radiance[ok])
    assert e <= 1e-4, f"{label}: radiance relative RMS {e} > 1e-4"
    return nbad
