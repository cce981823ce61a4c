"""Manual tool (under gpurun): host-built vs device-built wide BVH on the C4 scene - commit time and traversal time."""
import json
import sys
import time

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import layouts as L, native, scenes  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

W, H = 3840, 2160
arrays = oracle_scene_from_spec(scenes.terrain_scene(n_quads=708, n_spheres=256)).arrays()
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
json.dump(out, open("gpurun_out/build_compare.json", "w"), indent=1)
