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
