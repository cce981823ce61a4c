"""Manual tool (under gpurun): what each depth of a C4 frame costs the traversal - frames rendered with max_depth = 0..8, differenced:
rays, wide nodes and primitive tests per ray, and traversal time per ray of every depth.  python tests/gpu_depth_profile.py [spp]"""
import json
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import layouts as L, native, scenes  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W, H = 3840, 2160
sc = oracle_scene_from_spec(scenes.terrain_scene(n_quads=708, n_spheres=256))
ctx = native.Context(0)
ctx.scene_upload(sc.arrays())
cam = oracle_camera("C3", W, H)
rows, prev = [], None
for depth in range(0, 9):
    best = None
    for _ in range(3):
        ctx.render(cam, L.make_render_config(W, H, spp=spp, max_depth=depth, flags=L.RT_FLAG_KERNEL_TIMING)); ctx.sync(); s = ctx.stats()
        if best is None or s["lastTraceMs"] < best["lastTraceMs"]:
            best = s
    ctx.render(cam, L.make_render_config(W, H, spp=spp, max_depth=depth, flags=L.RT_FLAG_COUNTERS)); ctx.sync(); c = ctx.stats()
    cur = dict(rays=c["raysPrimary"] + c["raysBounce"] + c["raysAnyHitTraced"], nodes=c["wideNodes"], prims=c["trisTested"] + c["spheresTested"], trace_ms=best["lastTraceMs"], ms=best["lastRenderMs"])
    if prev is not None:
        d = {k: cur[k] - prev[k] for k in cur}
        rows.append(dict(depth=depth, rays_M=round(d["rays"] / 1e6, 2), nodes_per_ray=round(d["nodes"] / max(1, d["rays"]), 2), prims_per_ray=round(d["prims"] / max(1, d["rays"]), 2),
                         trace_ms=round(d["trace_ms"], 3), grays=round(d["rays"] / max(1e-9, d["trace_ms"]) / 1e6, 2), frame_ms=round(d["ms"], 3)))
        print(rows[-1], flush=True)
    prev = cur
json.dump(rows, open("gpurun_out/depth_profile.json", "w"), indent=1)
