"""GPU probe (manual tool, run under gpurun): first timings of the wavefront on the 1M-triangle scene.
Uses the oracle's scene builder, so it lives under tests/ and is not part of the product or the bench."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import build, layouts as L, native, scenes  # noqa: E402
from oracle import orc  # noqa: E402
from tests.parity import assert_parity, download_all  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

build.build_core()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 708
out = {}
t = time.time(); spec = scenes.terrain_scene(n_quads=nq, n_spheres=256); out["gen_s"] = time.time() - t
t = time.time(); sc = oracle_scene_from_spec(spec); out["oracle_build_s"] = time.time() - t; out["sort_ties"] = sc.sort_ties()
arrays = sc.arrays()
ctx = native.Context(0)
t = time.time(); ctx.scene_upload(arrays); out["upload_s"] = time.time() - t
W, H = 3840, 2160
cam = oracle_camera("C3", W, H)


def run(spp, depth, flags=0, reps=3, spp_pass=0):
    cfg = L.make_render_config(W, H, spp=spp, max_depth=depth, flags=flags, samples_per_pass=spp_pass)
    res = []
    for _ in range(reps):
        ctx.render(cam, cfg); ctx.sync(); res.append(ctx.stats())
    return res


st = run(1, 0, L.RT_FLAG_COUNTERS, reps=2)[-1]
out["C3_counters"] = {k: st[k] for k in ("raysPrimary", "wideNodes", "trisTested", "spheresTested", "lastRenderMs", "lastTraceMs", "kernelLaunches", "bvhWideNodeCount", "bvhPrimCount", "bvhBytes")}
sts = run(1, 0, 0, reps=5)
out["C3_ms"] = [s["lastRenderMs"] for s in sts]
ms = min(out["C3_ms"]); out["C3_Grays_s"] = st["raysPrimary"] / ms / 1e6
st = run(4, 8, L.RT_FLAG_COUNTERS, reps=1)[-1]
out["C4_4spp_counters"] = {k: st[k] for k in ("raysPrimary", "raysBounce", "raysShadow", "wideNodes", "trisTested", "spheresTested", "lastRenderMs", "lastTraceMs", "kernelLaunches")}
sts = run(4, 8, 0, reps=3)
out["C4_4spp_ms"] = [s["lastRenderMs"] for s in sts]
ms = min(out["C4_4spp_ms"]); s = sts[-1]
out["C4_Grays_s_prim_bounce"] = (s["raysPrimary"] + s["raysBounce"]) / ms / 1e6
out["C4_Grays_s_all"] = (s["raysPrimary"] + s["raysBounce"] + s["raysShadow"]) / ms / 1e6
print(json.dumps(out, indent=1))
# parity on a crop of the full 4K frame against the oracle (1M triangles)
box = (1792, 1008, 2048, 1152)
t = time.time(); r = orc.render(sc, cam, orc.make_config(W, H, spp=2, max_depth=8, crop=box)); out["oracle_crop_s"] = time.time() - t
ctx.render(cam, L.make_render_config(W, H, spp=2, max_depth=8, flags=L.RT_FLAG_PATH_AOVS)); ctx.sync()
prod = download_all(ctx)
nbad = assert_parity(r, prod, W, H, box=box, spp=2, allow_degenerate=4, label="1M crop")
out["crop_parity_bad"] = nbad
out["oracle_counters"] = r.counters; out["oracle_seconds"] = r.seconds
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
