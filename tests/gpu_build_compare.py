"""Manual tool (under gpurun): host-built vs device-built wide BVH - commit time and traversal time.

    python tests/gpu_build_compare.py [terrain|debris]

terrain = the C4 scene (a regular height field: the easy case for Morton-order builders); debris = 1 500 tessellated blobs of
very different sizes scattered over a coarse ground (uneven density, many overlapping boxes).  The device tree is chosen with
RT_DEVICE_TREE (ploc | radix | best), RT_DEVICE_COLLAPSE (dp | greedy); RT_BUILD_TIMING=1 prints the commit phases."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import layouts as L, native, scenes  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402


def debris_scene(n_blobs=1500, seed=0xD3B5) -> scenes.SceneSpec:
    rng = np.random.default_rng(seed)
    pos, tris = [], []
    base = 0

    def add(p, t):
        nonlocal base
        pos.append(p.astype(np.float32)); tris.append((t + base).astype(np.int32)); base += len(p)

    g = 64   # coarse ground
    gi, gj = np.meshgrid(np.arange(g + 1), np.arange(g + 1), indexing="ij")
    gp = np.stack([-50 + 100 * gi / g, 0.3 * np.sin(gi * 0.7) * np.cos(gj * 0.5), -50 + 100 * gj / g], -1).reshape(-1, 3)
    qi, qj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    v00 = (qi * (g + 1) + qj).reshape(-1)
    add(gp, np.concatenate([np.stack([v00, v00 + 1, v00 + g + 1], -1), np.stack([v00 + g + 1, v00 + 1, v00 + g + 2], -1)]))
    for _ in range(n_blobs):
        r = float(np.exp(rng.uniform(np.log(0.15), np.log(4.0))))
        c = np.array([rng.uniform(-45, 45), rng.uniform(0.5, 18.0), rng.uniform(-45, 45)])
        nl = int(rng.integers(6, 28)); nm = 2 * nl
        th = np.linspace(0, np.pi, nl + 1)[:, None]; ph = np.linspace(0, 2 * np.pi, nm, endpoint=False)[None, :]
        bump = 1.0 + 0.25 * np.sin(3 * th + rng.uniform(0, 6)) * np.cos(2 * ph + rng.uniform(0, 6))
        p = c + r * bump[..., None] * np.stack([np.sin(th) * np.cos(ph), np.cos(th) * np.ones_like(ph), np.sin(th) * np.sin(ph)], -1)
        p = p.reshape(-1, 3)
        a, b = np.meshgrid(np.arange(nl), np.arange(nm), indexing="ij")
        i00 = (a * nm + b).reshape(-1); i01 = (a * nm + (b + 1) % nm).reshape(-1); i10 = i00 + nm; i11 = i01 + nm
        add(p, np.concatenate([np.stack([i00, i10, i01], -1), np.stack([i01, i10, i11], -1)]))
    P = np.concatenate(pos); T = np.concatenate(tris)
    mesh = scenes.MeshSpec(P, T, np.zeros((1, 2), np.float32), np.zeros_like(T), np.zeros(len(T), np.int32), np.array([scenes.material((0.7, 0.7, 0.7))], dtype=L.MATERIAL))
    return scenes.SceneSpec(textures=[], mesh=mesh, mesh_first=True)


which = sys.argv[1] if len(sys.argv) > 1 else "terrain"
spec = scenes.terrain_scene(n_quads=708, n_spheres=256) if which == "terrain" else debris_scene()
W, H = 3840, 2160
arrays = oracle_scene_from_spec(spec).arrays()
print(which, "triangles", len(spec.mesh.tris), flush=True)
ctx = native.Context(0)
cam = oracle_camera("C3", W, H)
out = {}
for tag, dev in (("host_sah", False), ("device_lbvh", True), ("host_sah_again", False)):
    t0 = time.perf_counter(); ctx.scene_upload(arrays, device_build=dev); t_up = time.perf_counter() - t0
    t0 = time.perf_counter(); ctx.scene_upload(arrays, device_build=dev); t_up2 = time.perf_counter() - t0
    r = dict(upload_s=round(min(t_up, t_up2), 4))
    s0 = ctx.stats(); r["nodes"] = s0["bvhWideNodeCount"]
    for name, spp, depth in (("C3", 1, 0), ("C4x8", 8, 8)):
        cfg = L.make_render_config(W, H, spp=spp, max_depth=depth, flags=L.RT_FLAG_KERNEL_TIMING | L.RT_FLAG_COUNTERS)
        best = None
        for _ in range(3):
            ctx.render(cam, cfg); ctx.sync(); s = ctx.stats()
            if best is None or s["lastRenderMs"] < best["lastRenderMs"]:
                best = s
        rays = best["raysPrimary"] + best["raysBounce"] + best["raysAnyHitTraced"]
        r[name] = dict(ms=round(best["lastRenderMs"], 3), trace_ms=round(best["lastTraceMs"], 3), nodes_per_ray=round(best["wideNodes"] / rays, 2),
                       prims_per_ray=round((best["trisTested"] + best["spheresTested"]) / rays, 2))
    t0 = time.perf_counter(); ctx.scene_refit(arrays["meshPositions"].view("<f4").reshape(-1, 3)); r["refit_s"] = round(time.perf_counter() - t0, 4)
    out[tag] = r
    print(tag, r, flush=True)
json.dump(out, open("gpurun_out/build_compare_%s.json" % which, "w"), indent=1)
