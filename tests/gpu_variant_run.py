"""One timing run of the currently selected library build (RTCORE_B200_LIB): C3 primary-only and C4 at 8 spp.
Goes through native.Context only: the engine mirror links the in-tree core library, which must not be mixed with a variant build."""
import json
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import layouts as L, native, scenes  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

W, H = 3840, 2160
sc = oracle_scene_from_spec(scenes.terrain_scene(n_quads=708, n_spheres=256))
ctx = native.Context(0)
ctx.scene_upload(sc.arrays())
cam = oracle_camera("C3", W, H)
out = {}
for tag, spp, depth in (("C3", 1, 0), ("C4x8", 8, 8), ("C4x32", 32, 8)):
    cfg = L.make_render_config(W, H, spp=spp, max_depth=depth, flags=L.RT_FLAG_KERNEL_TIMING)
    best = None
    for _ in range(4):
        ctx.render(cam, cfg); ctx.sync(); s = ctx.stats()
        if best is None or s["lastRenderMs"] < best["lastRenderMs"]:
            best = s
    rays = best["raysPrimary"] + best["raysBounce"] + best["raysShadow"]
    out[tag] = dict(ms=round(best["lastRenderMs"], 3), trace_ms=round(best["lastTraceMs"], 3), grays_all=round(rays / best["lastRenderMs"] / 1e6, 3))
print(json.dumps(out))
