"""Manual tool (under gpurun): device time of every rank's share of an N-way C4 frame on one GPU, for several tile sizes - the
imbalance of the interleaved partition (the slowest rank is what an N-GPU frame takes).  python tests/gpu_rank_balance.py [N] [spp]"""
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


def frame_ms(world, rank, tile):
    cfg = L.make_render_config(W, H, spp=spp, max_depth=8, rank=rank, world_size=world, tile_size=tile)
    best = 1e9
    for _ in range(3):
        ctx.render(cam, cfg); ctx.sync(); best = min(best, ctx.stats()["lastRenderMs"])
    return best


one = frame_ms(1, 0, 32)
out = {"single_gpu_ms": round(one, 3)}
for tile in (64, 32, 16, 8):
    ms = [frame_ms(N, r, tile) for r in range(N)]
    out[f"tile{tile}"] = dict(max=round(max(ms), 3), mean=round(sum(ms) / N, 3), min=round(min(ms), 3), efficiency_bound=round(one / N / max(ms), 4))
    print(tile, out[f"tile{tile}"], flush=True)
json.dump(out, open("gpurun_out/rank_balance.json", "w"), indent=1)
