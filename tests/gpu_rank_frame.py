"""Manual tool (under gpurun): device time of ONE rank's share of a C4 frame on one GPU - rank r of N renders only its interleaved
tiles (no communicator needed), which is what each GPU of an N-GPU job does.  python tests/gpu_rank_frame.py [N] [spp]"""
import json
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import layouts as L, native, scenes  # noqa: E402
from tests.util import oracle_camera, oracle_scene_from_spec  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W, H = 3840, 2160
sc = oracle_scene_from_spec(scenes.terrain_scene(n_quads=708, n_spheres=256))
ctx = native.Context(0)
ctx.scene_upload(sc.arrays())
cam = oracle_camera("C3", W, H)
out = {}
for world in (1, N):
    for rank in ((0,) if world == 1 else (0, world // 2)):
        cfg = L.make_render_config(W, H, spp=spp, max_depth=8, flags=L.RT_FLAG_KERNEL_TIMING, rank=rank, world_size=world)
        best = None
        for _ in range(5):
            ctx.render(cam, cfg); ctx.sync(); s = ctx.stats()
            if best is None or s["lastRenderMs"] < best["lastRenderMs"]:
                best = s
        out[f"rank{rank}of{world}"] = dict(ms=round(best["lastRenderMs"], 3), trace_ms=round(best["lastTraceMs"], 3))
one = out["rank0of1"]["ms"]
for k, v in out.items():
    v["x_ideal"] = round(v["ms"] / (one / int(k.split("of")[1])), 4)
print(json.dumps(out))
