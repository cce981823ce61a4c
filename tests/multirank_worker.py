"""Worker of tests/test_gpu_multigpu.py and of manual `gpurun --gpus N` sessions: one process per GPU (torchrun), the NATIVE
multi-GPU path of the C ABI (rt_comm_init + rt_gather_frame over NCCL) against a single-context render on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multirank_worker.py [out.json]

torch.distributed is only the side channel that hands rank 0's communicator id to the other ranks (a C# host would use a pipe)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from ilgpu_raytracing_b200 import layouts as L, native, scenes
    from tests.util import oracle_camera, oracle_scene_from_spec
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")   # side channel only
    ids = [native.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = native.Context(local)
    ctx.comm_init(ids[0], rank, world)
    sc = oracle_scene_from_spec(scenes.terrain_scene(96, 24))
    ctx.scene_upload(sc.arrays())
    W, H, tile, spp, depth = 776, 440, 32, 3, 5
    cam = oracle_camera("C3", W, H)
    report = {"world": world, "cases": []}
    ok = True
    for what, name in ((L.RT_GATHER_RADIANCE | L.RT_GATHER_DEPTH_OBJID, "radiance+depth+objId"), (L.RT_GATHER_RGBA8 | L.RT_GATHER_DEPTH_OBJID, "rgba8+depth+objId"), (L.RT_GATHER_RGBA8, "rgba8")):
        for frame in range(3):   # back-to-back frames: the gather of frame k overlaps the render of frame k + 1 (double-buffered payloads)
            ctx.render(cam, L.make_render_config(W, H, spp=spp, max_depth=depth, frame=frame, rng_lock_noise=0, tile_size=tile, rank=rank, world_size=world))
            ctx.gather_frame(0, what)
        ctx.sync()
        if rank == 0:
            got = {"rgba8": ctx.download(L.RT_BUF_GATHERED_RGBA8)}
            if what & L.RT_GATHER_DEPTH_OBJID:
                got["depth"], got["objId"] = ctx.download(L.RT_BUF_GATHERED_DEPTH), ctx.download(L.RT_BUF_GATHERED_OBJID)
            if what & L.RT_GATHER_RADIANCE:
                got["radiance"] = ctx.download(L.RT_BUF_GATHERED_RADIANCE)
            if what & L.RT_GATHER_DEPTH_OBJID:   # TAAU present of the gathered frame (needs objId)
                ctx.present(W, H, taau=True, reset_history=True)
                got["present"] = ctx.download(L.RT_BUF_PRESENT)
            single = native.Context(local)
            single.scene_upload(sc.arrays())
            single.render(cam, L.make_render_config(W, H, spp=spp, max_depth=depth, frame=2, rng_lock_noise=0))
            single.sync()
            want = {"rgba8": single.download(L.RT_BUF_RGBA8), "depth": single.download(L.RT_BUF_DEPTH), "objId": single.download(L.RT_BUF_OBJID),
                    "radiance": single.download(L.RT_BUF_RADIANCE)}
            single.present(W, H, taau=True, reset_history=True)
            want["present"] = single.download(L.RT_BUF_PRESENT)
            single.close()
            case = {"what": name, "mismatch": {k: int((np.asarray(got[k]) != np.asarray(want[k])).sum()) for k in got}, "gather_ms": ctx.stats()["lastGatherMs"]}
            case["ok"] = all(v == 0 for v in case["mismatch"].values())
            ok = ok and case["ok"]
            report["cases"].append(case)
    # ---- ReSTIR temporal + spatial reuse ACROSS the tile partition (SURVEY 8e caveat / 8f rank 1): every frame the ranks exchange the
    # current G-buffer (after the primary pass) and the reservoirs they wrote (after accumulate) inside rt_render; the gathered
    # sequence must equal the single-GPU sequence frame by frame, reservoirs included.  Matches Engine/RTRay.cs:339-435,476-516.
    from oracle import orc
    dsc = orc.Scene(); dsc.build_default()
    ctx.scene_upload(dsc.arrays())
    single = None
    if rank == 0:
        single = native.Context(local)
        single.scene_upload(dsc.arrays())
    W2, H2 = 648, 364
    prev = None
    reuse_case = {"what": "reuse across tiles, 4 frames, moving camera", "mismatch": {"rgba8": 0, "radiance": 0, "objId": 0, "reservoir_m": 0, "reservoir_wSum": 0}, "imports": 0}
    for frame in range(4):
        cam2 = orc.camera_create(W2, H2, 60.0, (0.0 + 0.07 * frame, 1.0 + 0.02 * frame, 3.0 - 0.05 * frame), (0.0, 0.5, 0.0))
        orc.camera_bake(cam2, W2, H2)
        if prev is None:
            prev = cam2.copy()
        fl = L.RT_FLAG_RESET_RESERVOIRS if frame == 0 else 0
        ctx.render(cam2, L.make_render_config(W2, H2, spp=3, max_depth=3, frame=frame, rng_lock_noise=0, temporal=1, spatial=1, flags=fl, tile_size=tile, rank=rank, world_size=world), prev_cam=prev)
        ctx.gather_frame(0, L.RT_GATHER_RADIANCE | L.RT_GATHER_DEPTH_OBJID)
        ctx.sync()
        if rank == 0:
            single.render(cam2, L.make_render_config(W2, H2, spp=3, max_depth=3, frame=frame, rng_lock_noise=0, temporal=1, spatial=1, flags=fl), prev_cam=prev)
            single.sync()
            m = reuse_case["mismatch"]
            m["rgba8"] += int((ctx.download(L.RT_BUF_GATHERED_RGBA8) != single.download(L.RT_BUF_RGBA8)).sum())
            m["radiance"] += int((ctx.download(L.RT_BUF_GATHERED_RADIANCE)[:, :3] != single.download(L.RT_BUF_RADIANCE)[:, :3]).any(axis=1).sum())
            m["objId"] += int((ctx.download(L.RT_BUF_GATHERED_OBJID) != single.download(L.RT_BUF_OBJID)).sum())
            ra, rb = ctx.download(L.RT_BUF_RESERVOIR), single.download(L.RT_BUF_RESERVOIR)
            m["reservoir_m"] += int((ra["m"] != rb["m"]).sum()); m["reservoir_wSum"] += int((ra["wSum"] != rb["wSum"]).sum())
            if frame > 0:
                reuse_case["imports"] += int((rb["m"] > 9).sum())
        prev = cam2.copy()
    if rank == 0:
        reuse_case["ok"] = all(v == 0 for v in reuse_case["mismatch"].values()) and reuse_case["imports"] > 0
        ok = ok and reuse_case["ok"]
        report["cases"].append(reuse_case)
        single.close()
    report["ok"] = ok
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    if rank == 0:
        print(json.dumps(report))
        if len(sys.argv) > 1:
            json.dump(report, open(sys.argv[1], "w"), indent=1)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
