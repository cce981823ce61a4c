"""not gpu: the N > 1 protocol of bench.py (tile ownership, padded gather to rank 0, rank-major offsets, de-interleave)
run with torch.distributed / gloo, world_size 2, on the CPU.  Each rank renders its own tiles with the host simulator;
rank 0 reassembles the frame and compares it with the single-rank image bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, W, H, tile, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ilgpu_raytracing_b200 import layouts as L, native, scenes
    from tests.hostsim_binding import HostSimScene
    from tests.util import oracle_camera, oracle_scene_from_spec
    sc = oracle_scene_from_spec(scenes.terrain_scene(24, 4))
    hs = HostSimScene(sc.arrays())
    cam = oracle_camera("C3", W, H)
    part = hs.render(cam, L.make_render_config(W, H, spp=2, max_depth=2, tile_size=tile, rank=rank, world_size=world))
    # tile-compacted payload in the rank's own pixel order (what RT_BUF_TILE_RADIANCE holds on the device)
    mine = np.nonzero(part["primId"] != -2)[0]
    counts = [native.tiles_owned_pixels(W, H, tile, r, world) for r in range(world)]
    assert len(mine) == counts[rank]
    max_n = max(counts)
    payload = torch.zeros((max_n, 5), dtype=torch.float32)
    payload[: len(mine), :4] = torch.from_numpy(np.concatenate([part["radiance"][mine], np.ones((len(mine), 1), np.float32)], axis=1))
    payload[: len(mine), 4] = torch.from_numpy(mine.astype(np.float32))   # pixel ids (exact in f32 for this image size)
    gathered = [torch.zeros_like(payload) for _ in range(world)] if rank == 0 else None
    dist.gather(payload, gathered, dst=0)
    if rank == 0:
        full = hs.render(cam, L.make_render_config(W, H, spp=2, max_depth=2))
        img = np.zeros((W * H, 3), np.float32)
        seen = np.zeros(W * H, np.int32)
        for r in range(world):
            g = gathered[r].numpy()[: counts[r]]
            ids = g[:, 4].astype(np.int64)
            img[ids] = g[:, :3]
            seen[ids] += 1
        ok = bool(np.all(seen == 1) and np.array_equal(img, full["radiance"]))
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_tile_gather(tmp_path):
    out = tmp_path / "result.txt"
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 160, 96, 32, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
