#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native renderer core.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C4|C3|C2|C1]

A "step" is one frame of the hot path (generate -> extend -> shade/bounce -> accumulate/tone-map) of the
workload BASELINE.json's metric is quoted on: C4 = the synthetic 1 002 528-triangle terrain + 256 mirror /
glass / diffuse spheres at 3840x2160, 64 spp, 8 bounces (reference-faithful variant: triangles Lambert,
SURVEY.md §8d).  metric = Mrays/s counting primary + bounce rays (TraceClosest calls; shadow rays are
reported beside it), value = whole-job rays / device time with the scene resident in HBM.

  e2e          same metric through the engine API (RTRenderer.RenderDirectToPbo + Framebuffer.DownloadToCpu)
               with host camera/config in and the 12 B/px framebuffer (RGBA8 + depth + objId) read back to
               pinned host memory every step; at N > 1 the step is render + NCCL gather + de-interleave on rank 0 + the
               gathered RGBA8 image (4 B/px) read back to pinned host memory on rank 0
  roofline     extend (wide-BVH traversal) kernels: algorithmic node/primitive/queue bytes per frame divided
               by their summed CUDA-event duration, against the measured HBM copy bandwidth
  cpu_baseline the CPU oracle (line-by-line restatement of the reference kernels, all host threads) on a
               bounded crop of the same frame
  --impl reference   times only that CPU restatement (the reference itself is C#/.NET + ILGPU + OpenGL and
               cannot be built or run in this image: no dotnet, no display)

N > 1 (torchrun, one process per GPU): the image is split into interleaved 32x32 screen tiles, the scene is
replicated, every rank renders its tiles and the float4 tile payloads are gathered to rank 0 over NCCL and
de-interleaved there, all inside the timed region ("strong" scaling: the frame is fixed); the gather of frame k runs on
its own stream while frame k + 1 renders, and the region ends when the last gathered image is complete.

Frames are submitted back to back (rt_render does not wait for the frame), so `value` is device throughput; the e2e leg waits
for every frame, and on a shared host single steps pick up rare 10-40 ms stalls - hence `ms_per_step_median` beside the mean.
Switches for diagnosis (not for reported numbers): RT_BENCH_TILE, RT_BENCH_NO_OVERLAP, RT_BENCH_NO_CLOCKS, RT_BENCH_NO_REFIT,
RT_BENCH_DEBUG (per-step e2e wall / device times on stderr).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: scene, camera, width, height, spp, depth, oracle crop (x0,y0,x1,y1) and spp of the CPU sample (~10-30 s on 16 cores),
    # ref_crop: the smaller per-step sample of --impl reference (a few seconds per step)
    "C1": dict(scene="default", cam="C1A", w=1280, h=720, spp=1, depth=1, crop=None, cpu_spp=1),
    "C2": dict(scene="spheres", cam="C2", w=1920, h=1080, spp=16, depth=4, crop=(480, 270, 1440, 810), cpu_spp=16),
    "C3": dict(scene="terrain", cam="C3", w=3840, h=2160, spp=1, depth=0, crop=(960, 540, 2880, 1620), cpu_spp=1),
    # C4: ONE sample for the cpu_baseline leg, the --impl reference arm and the parity block: the central 512x288 of the frame at the
    # full 64 spp / depth 8 (9.4 M paths, ~7 s on 16 host threads)
    "C4": dict(scene="terrain+spheres", cam="C3", w=3840, h=2160, spp=64, depth=8, crop=(1664, 936, 2176, 1224), cpu_spp=64),
    # C5: 7680x4320 progressive accumulation, a step = one 16-spp frame of the 16-frame / 256-spp sequence (rngLockNoise = 0,
    # frame index advancing, float4 accumulator), extension variant of the scene (per-patch Lambert / mirror / glass triangle materials)
    "C5": dict(scene="terrain+spheres+mats", cam="C3", w=7680, h=4320, spp=16, depth=8, crop=(3712, 2088, 3968, 2232), cpu_spp=16,
               progressive=True, tri_materials=True),
}


def workload_config(name, wl, spec, world, tile, spp=None):
    """The `config` object of a bench line: the same dict from both arms (the driver compares them)."""
    progressive = bool(wl.get("progressive"))
    return {"workload": name, "scene": wl["scene"], "triangles": int(len(spec.mesh.tris)) if spec.mesh is not None else 0, "spheres": int(len(spec.spheres)),
            "width": wl["w"], "height": wl["h"], "spp": wl["spp"] if spp is None else spp, "max_depth": wl["depth"],
            "variant": "extension (RT_FLAG_TRI_MATERIALS: per-patch Lambert / mirror / glass triangles)" if wl.get("tri_materials") else "reference-faithful (triangles Lambert)",
            "progressive": "float4 accumulation, one 16-spp frame of the 256-spp sequence per step" if progressive else None,
            "partition": f"interleaved {tile}x{tile} screen tiles x{world}, scene replicated" if world > 1 else "single GPU",
            "l2": "no explicit flush: per-step path state + queues (GBs) exceed the 126 MB L2; the BVH is meant to stay resident"}


def make_spec(kind: str):
    from ilgpu_raytracing_b200 import scenes
    if kind == "default":
        return scenes.default_scene()
    if kind == "spheres":
        return scenes.sphere_grid_scene(32)
    if kind == "terrain":
        return scenes.terrain_scene(708, 0)
    if kind == "terrain+spheres":
        return scenes.terrain_scene(708, 256)
    if kind == "terrain+spheres+mats":
        return scenes.terrain_scene(708, 256, patch_materials=True)
    raise ValueError(kind)


def peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


RAY_RECORD_BYTES = 32   # what k_extend reads per ray: origin | slot, direction | slot (2 x float4; the box-test reciprocal is derived in the kernel)


def csrc_hash() -> str:
    """sha256 over the DEVICE code of a frame: the headers its kernels are made of and rtcore.cu down to its "host side" marker (the C ABI and
    launch plumbing below it do not change what a kernel executes).  Ties an ncu capture to the kernels it profiled."""
    import hashlib
    h = hashlib.sha256()
    base = os.path.join(ROOT, "ilgpu_raytracing_b200", "csrc")
    for name in ("rt_core.h", "rt_traverse.h", "rt_wavefront.h", "rt_tiles.h"):   # rt_build.h holds the scene-commit kernels (refit / build): no part of a frame
        h.update(name.encode()); h.update(open(os.path.join(base, name), "rb").read())
    src = open(os.path.join(base, "rtcore.cu"), "rb").read()
    marker = b"// ------------------------------------------------------------------------------------------------ host side"
    cut = src.find(marker)
    h.update(b"rtcore.cu[kernels]"); h.update(src if cut < 0 else src[:cut])
    return h.hexdigest()[:16]


CAPTURE_FILE = os.path.join(ROOT, "profiles", "r02_extend_capture_c4.json")


def capture_status(name, args, world) -> str:
    if name != "C4" or args.spp:
        return "no capture for this workload (ncu captures are taken for the headline workload, C4)"
    if not os.path.exists(CAPTURE_FILE):
        return "no capture committed"
    t = json.load(open(CAPTURE_FILE))
    return f"STALE: capture of sources {t.get('csrc_hash')} != current {csrc_hash()} - refused (re-run profiles/capture.py)"


def load_capture(name, args, world):
    """The committed ncu capture of all k_extend launches of one C4 frame (profiles/capture.py), ONLY if it was taken from the
    kernel sources in this tree: a capture of older kernels is refused, not silently divided by this run's launches."""
    if name != "C4" or args.spp or not os.path.exists(CAPTURE_FILE):
        return None
    t = json.load(open(CAPTURE_FILE))
    if t.get("csrc_hash") != csrc_hash() or (world != 1 and not t.get("rays_traced")):
        return None
    return {"thread_instructions": float(t["thread_instructions"]), "warp_instructions": float(t["warp_instructions"]), "rays_traced": t.get("rays_traced"),
            "dram_bytes": float(t["dram_read_bytes"]) + float(t["dram_write_bytes"]), "launches": int(t["launches"]), "file": "profiles/" + os.path.basename(CAPTURE_FILE)}


def measured_peaks():
    p = os.path.join(ROOT, "profiles", "r02_gpu_peaks.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    iss = d.get("issue_lane_inst_per_s_T", {})
    return {"fp32_ffma_tflops": d.get("fp32_ffma_tflops"), "fp32_unfused_tflops": d.get("fp32_mul_add_unfused_tflops"),
            "issue_lane_inst_per_s_T": max(v for k, v in iss.items() if k != "ffma2_instructions") if iss else None,
            "l2_random": d.get("l2_read_gbs_random_80B_records_58MB"), "l2_stream": d.get("l2_read_gbs_stream_58MB")}


class ClockSampler:
    """SM clocks / throttle reasons sampled every 200 ms during the timed region: through NVML in-process (a light call; spawning
    nvidia-smi five times a second perturbs sub-millisecond frames), nvidia-smi as the fallback."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.rows = []      # (sm_mhz, sm_max_mhz, [reason names])
        self._stop = threading.Event()
        self._idx = gpu_index
        self._t = None
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = get(self._h)
        masks = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}   # NVML clocks-event-reason bits
        self.rows.append((float(sm), float(mx), [k for k, m in masks.items() if bits & m]))

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self._idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(",")]
        if len(r) >= 6 and r[0].replace(".", "").isdigit():
            self.rows.append((float(r[0]), float(r[1]), [self.NAMES[i] for i in range(4) if r[2 + i].lower().startswith("active")]))

    def start(self):
        def run():
            while not self._stop.is_set():
                try:
                    self._sample_nvml() if self._nvml else self._sample_smi()
                except Exception:
                    pass
                self._stop.wait(0.2)

        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        reasons = sorted({n for r in self.rows for n in r[2]})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm), source=self.source)


def cpu_oracle_sample(wl: dict, threads: int = 0, aovs: bool = False) -> dict:
    """Time the CPU oracle on the workload's crop: Mrays/s (primary + bounce) over the two passes.  The rendered crop is returned
    too (`result`): the parity block compares the device's frame with it."""
    from oracle import orc
    from tests.util import oracle_camera, oracle_scene_from_spec
    sc = oracle_scene_from_spec(make_spec(wl["scene"]))
    cam = oracle_camera(wl["cam"], wl["w"], wl["h"])
    crop = wl["crop"] or (0, 0, wl["w"], wl["h"])
    cfg = orc.make_config(wl["w"], wl["h"], spp=wl["cpu_spp"], max_depth=wl["depth"], crop=crop, threads=threads,
                          flags=1 if wl.get("tri_materials") else 0, rng_lock_noise=0 if wl.get("progressive") else 1)
    t0 = time.perf_counter()
    r = orc.render(sc, cam, cfg, aovs=aovs)
    wall = time.perf_counter() - t0
    secs = r.seconds[0] + r.seconds[1]
    rays = r.counters["raysPrimary"] + r.counters["raysBounce"]
    return dict(result=r, crop=crop, mrays=rays / secs / 1e6, rays=rays, rays_shadow=r.counters["raysShadow"], seconds=secs, wall=wall,
                cores=threads or orc.lib().orc_hardware_threads(), counters=r.counters,
                sample=f"crop {crop[2] - crop[0]}x{crop[3] - crop[1]} of the {wl['w']}x{wl['h']} frame at {wl['cpu_spp']} spp, depth {wl['depth']}", scene=sc, cam=cam)


def parity_vs_oracle(ctx, wl: dict, cam, cfg_for, oracle: dict) -> dict:
    """The device's frame of the TIMED configuration against the oracle crop the cpu_baseline leg rendered (same scene, camera, seed,
    spp, depth): primary hit ids, depth / objId, float radiance and RGBA8 of the frame exactly as timed, then per-sample bounce
    counts, terminators and hit-id path hashes from one more frame with RT_FLAG_PATH_AOVS.  Zeros everywhere = bit-exact."""
    from ilgpu_raytracing_b200 import layouts as L
    from tests.parity import crop as cut, rel_rms
    r, box = oracle["result"], oracle["crop"]
    W, H, spp = wl["w"], wl["h"], max(1, wl["cpu_spp"])
    ctx.render(cam, cfg_for(0, 0))
    ctx.sync()
    prim, inst = cut(ctx.download(L.RT_BUF_PRIM_ID), W, H, box), cut(ctx.download(L.RT_BUF_INST_ID), W, H, box)
    rgba, rad = cut(ctx.download(L.RT_BUF_RGBA8), W, H, box), cut(ctx.download(L.RT_BUF_RADIANCE)[:, :3], W, H, box)
    dep, oid = cut(ctx.download(L.RT_BUF_DEPTH), W, H, box), cut(ctx.download(L.RT_BUF_OBJID), W, H, box)
    out = {"against": "CPU oracle (restatement of Engine/RTRay.cs:188-325), the crop of the cpu_baseline sample", "crop": list(box), "px": int(prim.size), "spp": spp,
           "id_mismatch": int(((prim != r.primId) | (inst != r.instId)).sum()), "depth_objid_mismatch": int(((dep != r.depth) | (oid != r.objId)).sum()),
           "rgba8_mismatch": int((rgba != r.rgba8).sum()), "radiance_px_not_bit_identical": int((rad != r.radiance).any(axis=1).sum()), "rel_rms": rel_rms(rad, r.radiance)}
    if r.segCount.size:
        ctx.render(cam, cfg_for(L.RT_FLAG_PATH_AOVS, 0))
        ctx.sync()
        seg, term, hsh = (cut(ctx.download(w), W, H, box, planes=spp) for w in (L.RT_BUF_SEG_COUNT, L.RT_BUF_TERM_CODE, L.RT_BUF_PATH_HASH))
        out["paths"] = int(seg.size)
        out["path_mismatch"] = int(((seg != r.segCount) | (term != r.termCode) | (hsh != r.pathHash)).sum())
        out["aov_frame_radiance_px_not_bit_identical"] = int((cut(ctx.download(L.RT_BUF_RADIANCE)[:, :3], W, H, box) != r.radiance).any(axis=1).sum())
    out["ok"] = all(out.get(k, 0) == 0 for k in ("id_mismatch", "depth_objid_mismatch", "rgba8_mismatch", "path_mismatch")) and out["rel_rms"] <= 1e-4
    return out


def fast_shading_leg(ctx, wl: dict, cam, cfg_for, oracle: dict, rays_pb: float) -> dict:
    """The opt-in tolerance mode (RT_FLAG_FAST_SHADING): same frame, the ReSTIR sky candidates scored with FMA + hardware special functions.
    Reported beside the headline (which stays bit-exact): device time, and against the oracle crop ids / bounce counts (must be exact)
    and the radiance relative RMS (north_star: <= 1e-4)."""
    from ilgpu_raytracing_b200 import layouts as L
    from tests.parity import crop as cut, rel_rms
    r, box = oracle["result"], oracle["crop"]
    W, H, spp = wl["w"], wl["h"], max(1, wl["cpu_spp"])
    ctx.render(cam, cfg_for(L.RT_FLAG_FAST_SHADING | L.RT_FLAG_PATH_AOVS, 0))
    ctx.sync()
    prim = cut(ctx.download(L.RT_BUF_PRIM_ID), W, H, box)
    seg, term = (cut(ctx.download(w), W, H, box, planes=spp) for w in (L.RT_BUF_SEG_COUNT, L.RT_BUF_TERM_CODE))
    rad = cut(ctx.download(L.RT_BUF_RADIANCE)[:, :3], W, H, box)
    rgba = cut(ctx.download(L.RT_BUF_RGBA8), W, H, box)
    ms = []
    for _ in range(4):
        ctx.render(cam, cfg_for(L.RT_FLAG_FAST_SHADING, 0))
        ctx.sync()
        ms.append(ctx.stats()["lastRenderMs"])
    best = min(ms[1:])
    return {"flag": "RT_FLAG_FAST_SHADING (opt-in; the headline numbers above are the bit-exact mode)", "ms_per_step": best, "value": rays_pb / (best * 1e-3) / 1e6, "unit": "Mrays/s",
            "id_mismatch": int((prim != r.primId).sum()), "bounce_count_or_terminator_mismatch": int(((seg != r.segCount) | (term != r.termCode)).sum()),
            "rel_rms_vs_oracle": rel_rms(rad, r.radiance), "rgba8_px_differ": int((rgba != r.rgba8).sum()), "px": int(prim.size), "tolerance": 1e-4}


def run_c0(args, rank: int, local_rank: int):
    """Workload C0 = the reference's OWN operating point (VERDICT r1 "next" 7): the engine defaults of Engine/RTRenderer.cs:43-49,113-116,204 -
    a 1280x720 window traced at round(0.67 x) = 858x482, 2 spp, MaxDepth 3, ReSTIR temporal + spatial reuse on, TAAU resolve into the PBO -
    on the default scene (Scene.BuildDefaultScene), through RTRenderer.RenderDirectToPbo with a camera that moves every frame.  The
    un-translated camera is used (the reference's default camera looks away from the five small spheres and sees ground + sky only,
    SURVEY.md section 8a quirk 6).  A step = one displayed frame.  Both submission modes are measured - plain launches and the frame
    graph (RT_FLAG_FRAME_GRAPH: one CUDA-graph launch per frame, node parameters refreshed) - each as the best and the median of five
    batches of back-to-back frames (AsyncSubmit; the shared hosts of the pool add large, irregular submission delays), plus the
    reference's own cadence: per-frame Synchronize() and the presented image read back to pinned host memory (e2e)."""
    import torch
    from ilgpu_raytracing_b200 import engine, layouts as L
    OUT_W, OUT_H = 1280, 720
    if rank != 0:
        return   # replicas only: this workload is one interactive view
    torch.cuda.set_device(local_rank)

    def new_renderer(graph: bool):
        r = engine.RTRenderer(local_rank, OUT_W, OUT_H)
        r.configure(renderScale=0.67, enableTAAU=1, enableTemporalReuse=1, enableSpatialReuse=1, spp=2, maxDepth=3, rngLockNoise=1, fixedSeed=1,
                    flags=(L.RT_FLAG_FRAME_GRAPH if graph else 0))
        return r

    def camera_at(frame: int):
        cam = engine.config_camera("C1B", OUT_W, OUT_H)
        return engine.camera_translate(cam, 0.004 * frame, 0.001 * frame, -0.003 * frame)   # a slow fly-through: the temporal reprojection has work to do

    batch = max(args.steps, 400)
    n_batches = 5
    cams = [camera_at(f) for f in range(args.warmup + (n_batches + 2) * batch + 16)]
    pbo = torch.zeros(OUT_W * OUT_H, dtype=torch.int32, device="cuda")
    host = torch.empty(OUT_W * OUT_H, dtype=torch.int32).pin_memory()
    host_np = host.numpy()
    pbo_ptr = pbo.data_ptr()
    sampler = ClockSampler(local_rank)
    if not os.environ.get("RT_BENCH_NO_CLOCKS"):
        sampler.start()
    modes = {}
    for graph in (False, True):
        rdr = new_renderer(graph)
        ctx = rdr.native
        stream = torch.cuda.ExternalStream(ctx.stream_handle())   # events on the context's own stream
        frame_no = [0]

        def step():
            f = frame_no[0]
            rdr.camera = cams[f]
            rdr.RenderDirectToPbo(pbo_ptr, OUT_W, OUT_H, f, 0.016)
            frame_no[0] += 1

        rdr.configure(asyncSubmit=1)
        for _ in range(args.warmup + 8):
            step()
        rdr.Synchronize()
        per_batch, cpu_us = [], []
        for _ in range(n_batches):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0 = time.thread_time()
            ev0.record(stream)
            for _ in range(batch):
                step()
            ev1.record(stream)
            cpu_us.append((time.thread_time() - c0) / batch * 1e6)
            rdr.Synchronize()
            per_batch.append(ev0.elapsed_time(ev1) / batch * 1e3)
        st = ctx.stats()
        rays_pb = st["raysPrimary"] + st["raysBounce"]
        launches = st["kernelLaunches"] + 1   # + the present kernel
        cfg = rdr.last_config()
        rdr.configure(asyncSubmit=0)          # e2e: per-frame Synchronize() like the reference, plus the displayed image to the host
        per = []
        for i in range(batch + 3):
            t0 = time.perf_counter()
            f = frame_no[0]
            rdr.camera = cams[f]
            rdr.RenderDirectToPbo(None, OUT_W, OUT_H, f, 0.016)     # headless: the presented image stays in the core ...
            frame_no[0] += 1
            ctx.download(L.RT_BUF_PRESENT, out=host_np)             # ... and is read back through the ABI (rt_download, page-locked destination)
            if i >= 3:
                per.append(time.perf_counter() - t0)
        modes[graph] = dict(us_per_frame_best=min(per_batch), us_per_frame_median=float(np.median(per_batch)), host_cpu_us_per_frame=float(np.median(cpu_us)), rays=rays_pb,
                            launches=launches, cfg=(cfg.width, cfg.height), e2e_ms_median=float(np.median(per)) * 1e3, e2e_ms_mean=float(np.mean(per)) * 1e3,
                            device_ms_per_frame_synced=ctx.stats()["lastRenderMs"])
        del stream
        rdr.close()
    clocks = sampler.stop()
    best = min((False, True), key=lambda g: modes[g]["us_per_frame_median"])
    m = modes[best]
    ms_per_step = m["us_per_frame_median"] * 1e-3
    rays_pb, launches, (inW, inH) = m["rays"], m["launches"], m["cfg"]
    e2e_s = m["e2e_ms_median"] * 1e-3

    # parity: the first frames of the same sequence on a fresh renderer (frame graph ON) against the oracle (reuse reservoirs ping-pong, TAAU history)
    parity = cpu = None
    if not args.no_cpu_baseline:
        from oracle import orc
        sc = orc.Scene()
        sc.build_default()
        r2 = new_renderer(True)
        res = [np.zeros(inW * inH, orc.RESERVOIR), np.zeros(inW * inH, orc.RESERVOIR)]
        taa = orc.TaaState(OUT_W, OUT_H)
        mism = {"rgba8_low": 0, "objid": 0, "presented": 0}
        prev = None
        t_cpu = rays_cpu = 0.0
        for frame in range(4):
            r2.camera = cams[frame]
            r2.RenderDirectToPbo(pbo_ptr, OUT_W, OUT_H, frame, 0.016)
            low, _, obj = r2.DownloadToCpu()
            ocam = cams[frame].copy()
            orc.camera_bake(ocam, inW, inH)
            ocfg = orc.make_config(inW, inH, spp=2, max_depth=3, frame=frame, rng_lock_noise=1, temporal=1, spatial=1)
            ref = orc.render(sc, ocam, ocfg, prev_cam=ocam if prev is None else prev, res_prev=res[(frame & 1) ^ 1], res_cur=res[frame & 1], aovs=False)
            t_cpu += ref.seconds[0] + ref.seconds[1]
            rays_cpu += ref.counters["raysPrimary"] + ref.counters["raysBounce"]
            want = taa.resolve(ref.rgba8, ref.objId, inW, inH)
            mism["rgba8_low"] += int((low != ref.rgba8).sum()); mism["objid"] += int((obj != ref.objId).sum()); mism["presented"] += int((pbo.cpu().numpy() != want).sum())
            prev = ocam
        r2.close()
        parity = {"against": "CPU oracle: 4 frames of the same sequence (ReSTIR reuse ping-pong + TAAU history), traced image, objectId and presented image; frame graph on",
                  "frames": 4, "px_low": inW * inH, "px_presented": OUT_W * OUT_H, **{k + "_mismatch": v for k, v in mism.items()}, "ok": all(v == 0 for v in mism.values())}
        cpu = {"value": rays_cpu / t_cpu / 1e6, "unit": "Mrays/s", "cores": orc.lib().orc_hardware_threads(), "kind": "port", "sample": "4 full frames of the sequence (858x482, 2 spp, depth 3, reuse on)",
               "seconds": t_cpu, "frames_per_s": 4 / t_cpu}
    show = lambda d: {k: d[k] for k in ("us_per_frame_best", "us_per_frame_median", "host_cpu_us_per_frame", "e2e_ms_median", "e2e_ms_mean", "device_ms_per_frame_synced")}
    line = {"metric": "Mrays/s (primary+bounce) at the reference's interactive operating point", "value": rays_pb / (ms_per_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": 1,
            "steps": n_batches * batch, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C0", "scene": "Scene.BuildDefaultScene", "window": [OUT_W, OUT_H], "traced": [inW, inH], "spp": 2, "max_depth": 3, "reuse": "temporal + spatial",
                       "present": "TAAU", "camera": "un-translated default camera, moving every frame", "frame_graph": bool(best),
                       "l2": "working set (a few MB) is cache resident by nature: this workload is launch / latency bound, not bandwidth bound"},
            "frames_per_s": 1e3 / ms_per_step, "us_per_frame": ms_per_step * 1e3, "rays_per_step": {"primary_plus_bounce": rays_pb},
            "submission_modes": {"plain_launches": show(modes[False]), "frame_graph": show(modes[True]),
                                 "note": "value / ms_per_step = the median batch of the mode with the lower median; best = the least disturbed batch"},
            "clocks": clocks, "gpu_launches": int(launches * n_batches * batch), "launches_per_frame": int(launches),
            "e2e": {"value": rays_pb / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * L.CAMERA.itemsize + __import__("ctypes").sizeof(L.RtRenderConfig),
                    "d2h_bytes_per_step": OUT_W * OUT_H * 4, "ms_per_step": e2e_s * 1e3, "ms_per_step_median": e2e_s * 1e3},
            "parity": parity, "cpu_baseline": cpu, "roofline": None}
    print(json.dumps(line))


def run_reference(args, wl, name):
    """--impl reference: the CPU restatement of the reference kernels, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    from tests.util import oracle_camera, oracle_scene_from_spec
    spec = make_spec(wl["scene"])
    sc = oracle_scene_from_spec(spec)
    cam = oracle_camera(wl["cam"], wl["w"], wl["h"])
    crop = wl["crop"] or (0, 0, wl["w"], wl["h"])   # the same sample as the b200 arm's cpu_baseline leg
    cfg = orc.make_config(wl["w"], wl["h"], spp=wl["cpu_spp"], max_depth=wl["depth"], crop=crop,
                          flags=1 if wl.get("tri_materials") else 0, rng_lock_noise=0 if wl.get("progressive") else 1)
    rays = secs = 0.0
    for i in range(args.warmup + args.steps):
        r = orc.render(sc, cam, cfg, aovs=False)
        if i >= args.warmup:
            rays += r.counters["raysPrimary"] + r.counters["raysBounce"]
            secs += r.seconds[0] + r.seconds[1]
    v = rays / secs / 1e6
    cores = orc.lib().orc_hardware_threads()
    sample = f"crop {crop[2] - crop[0]}x{crop[3] - crop[1]} of the {wl['w']}x{wl['h']} frame at {wl['cpu_spp']} spp, depth {wl['depth']} per step"
    line = {"impl": "reference", "metric": "Mrays/s (primary+bounce) at 4K" if wl["w"] == 3840 else "Mrays/s (primary+bounce)", "value": v, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name, wl, spec, int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RT_BENCH_TILE", "16")), args.spp or None),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU restatement of the ILGPU kernels (stand-in for ILGPU CPUAccelerator, which cannot run here)"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS) + ["C0"])
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp (debugging; invalidates the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    name = args.workload
    if name == "C0":
        if args.impl == "reference":
            raise SystemExit("--impl reference times workload C4 (the headline); C0's CPU figure is in its own cpu_baseline")
        run_c0(args, int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")))
        return
    wl = dict(WORKLOADS[name])
    if args.spp:
        wl["spp"] = args.spp
    if args.impl == "reference":
        run_reference(args, wl, name)
        return

    import torch
    import torch.distributed as dist
    from ilgpu_raytracing_b200 import engine, layouts as L, native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H, spp, depth = wl["w"], wl["h"], wl["spp"], wl["depth"]
    tile = int(os.environ.get("RT_BENCH_TILE", "16"))   # interleaved screen-tile size (multi-GPU partition)

    # ---- scene + renderer through the engine API (the reference's host surface) -----------------------------------
    rdr = engine.RTRenderer(local_rank, W, H)
    spec = make_spec(wl["scene"])
    t0 = time.perf_counter()
    rdr.scene.load_spec(spec)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    rdr.Commit()
    t_commit = time.perf_counter() - t0
    t_refit = None
    if spec.mesh is not None and not os.environ.get("RT_BENCH_NO_REFIT"):   # Commit(ForceRefit) with the same vertices: the device-side refit path, leaves the scene as it is
        rdr.scene.SetMeshPositions(spec.mesh.positions)
        t0 = time.perf_counter()
        rdr.Commit(rdr.FORCE_REFIT)
        t_refit = time.perf_counter() - t0
    rdr.camera = engine.config_camera(wl["cam"], W, H)
    progressive = bool(wl.get("progressive"))
    base_flags = (L.RT_FLAG_TRI_MATERIALS if wl.get("tri_materials") else 0) | (L.RT_FLAG_ACCUMULATE if progressive else 0)
    lock = 0 if progressive else 1
    rdr.configure(renderScale=1.0, enableTemporalReuse=0, enableSpatialReuse=0, spp=spp, maxDepth=depth, rngLockNoise=lock, fixedSeed=1, flags=base_flags,
                  tileSize=tile, rank=rank, worldSize=world)
    ctx = rdr.native
    stream = torch.cuda.ExternalStream(ctx.stream_handle())   # the context's own stream, seen from torch (events, the timed region); every leg renders on it
    cam = rdr.camera
    # bake the derived camera fields exactly as RenderDirectToPbo does before launching
    engine.lib().eng_camera_bake(cam.ctypes.data_as(__import__("ctypes").c_void_p), W, H)

    def cfg_for(flags=0, frame=0):
        f = base_flags | flags | (L.RT_FLAG_RESET_ACCUM if progressive and frame == 0 else 0)
        return L.make_render_config(W, H, spp=spp, max_depth=depth, frame=frame if progressive else 0, rng_lock_noise=lock, flags=f, tile_size=tile, rank=rank, world_size=world)

    # N > 1: the communicator lives in the LIBRARY (rt_comm_init behind RTRenderer.InitMultiGpu); torch.distributed only hands rank 0's
    # 128-byte id to the other ranks and does the barrier / max-over-ranks bookkeeping of this script
    GATHER = L.RT_GATHER_RGBA8 | L.RT_GATHER_DEPTH_OBJID   # what a displayed frame is in the reference: colour + depth + objectId (Engine/RTRay.cs:59-64), 12 B/px
    if world > 1:
        ids = [engine.RTRenderer.NewCommunicatorId() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, device=torch.device("cuda", local_rank))
        rdr.InitMultiGpu(ids[0], rank, world)

    def step(cfg, what=GATHER):
        """One frame on this rank (async on `stream`); N > 1: plus rt_gather_frame - grouped ncclSend / ncclRecv of the tile payloads to
        rank 0 and the fused de-interleave there, on the library's communicator stream, overlapping the next frame's render."""
        ctx.render(cam, cfg)
        if world > 1:
            ctx.gather_frame(0, what)

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync()                    # the render stream AND the library's communicator stream
        torch.cuda.synchronize()

    def gathered_parity():
        """N > 1: one more frame through the REAL gather path (NCCL inside the library, float4 radiance + depth + objectId), then the
        same frame rendered by rank 0 alone (worldSize = 1); the two must be equal word for word.  Rank 0 returns the verdict."""
        import zlib
        step(cfg_for(0, 0), L.RT_GATHER_RADIANCE | L.RT_GATHER_DEPTH_OBJID)
        barrier()
        if rank != 0:
            return None
        got = {k: ctx.download(w).copy() for k, w in (("rgba8", L.RT_BUF_GATHERED_RGBA8), ("depth", L.RT_BUF_GATHERED_DEPTH), ("objId", L.RT_BUF_GATHERED_OBJID), ("radiance", L.RT_BUF_GATHERED_RADIANCE))}
        f = base_flags | (L.RT_FLAG_RESET_ACCUM if progressive else 0)
        ctx.render(cam, L.make_render_config(W, H, spp=spp, max_depth=depth, frame=0, rng_lock_noise=lock, flags=f, tile_size=tile, rank=0, world_size=1))
        ctx.sync()
        want = {"rgba8": ctx.download(L.RT_BUF_RGBA8), "depth": ctx.download(L.RT_BUF_DEPTH), "objId": ctx.download(L.RT_BUF_OBJID),
                "radiance": ctx.download(L.RT_BUF_ACCUM if progressive else L.RT_BUF_RADIANCE)}
        if progressive:   # the payload is what the pixel shows: the progressive mean (one frame accumulated: Lout * (1 / 1))
            want["radiance"] = want["radiance"].copy(); want["radiance"][:, :3] *= (np.float32(1.0) / want["radiance"][:, 3:4])
        mism = {k: int((got[k][:, :3] != want[k][:, :3]).any(axis=1).sum()) if k == "radiance" else int((got[k] != want[k]).sum()) for k in got}
        return {"against": "the same frame rendered by rank 0 alone (worldSize = 1)", "path": "rt_gather_frame (NCCL send/recv inside the library)", "px": int(got["rgba8"].size),
                "rgba8_mismatch": mism["rgba8"], "depth_mismatch": mism["depth"], "objid_mismatch": mism["objId"], "radiance_px_not_bit_identical": mism["radiance"],
                "crc32_gathered_rgba8": zlib.crc32(got["rgba8"].tobytes()), "crc32_single_gpu_rgba8": zlib.crc32(want["rgba8"].tobytes()), "ok": all(v == 0 for v in mism.values())}

    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step(cfg_for(0, i))
    barrier()
    st0 = ctx.stats()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("RT_BENCH_NO_CLOCKS"):
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for i in range(args.steps):
            step(cfg_for(0, args.warmup + i))   # progressive workloads advance the frame index (new samples every step)
        if world > 1:
            ctx.sync()                          # the timed region ends when the last gathered image is complete on rank 0 (the gather runs on the library's own stream)
        ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    st = ctx.stats()
    per_rank = None
    if world > 1:   # per-rank device time of the last frame + the gather: separates tile imbalance from tails
        t = torch.tensor([st["lastRenderMs"], st["lastGatherMs"]], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rm = [float(x[0]) for x in allt]
        per_rank = {"last_render_ms": rm, "min": min(rm), "max": max(rm), "imbalance": max(rm) / max(1e-9, float(np.mean(rm))) - 1.0,
                    "gather_ms_on_root_stream": float(allt[0][1])}
    rays_pb = st["raysPrimary"] + st["raysBounce"]
    rays_all = rays_pb + st["raysAnyHitTraced"]   # rays actually traced (first-vertex shadow rays answered by a shared sun probe are not)
    rays_ref_calls = rays_pb + st["raysShadow"]   # the reference's TraceClosest + ShadowOcclusion call count for the same frame
    launches = st["kernelLaunches"] * args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        r = torch.tensor([rays_pb, rays_all, launches, rays_ref_calls], dtype=torch.float64, device="cuda")
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rays_pb, rays_all, launches, rays_ref_calls = (float(x) for x in r.tolist())
    ms_per_step = ms / args.steps
    value = rays_pb / (ms_per_step * 1e-3) / 1e6

    # ---- e2e through the engine API with host buffers --------------------------------------------------------------
    n_px = W * H
    pin = [torch.empty(n_px, dtype=torch.int32).pin_memory(), torch.empty(n_px, dtype=torch.float32).pin_memory(), torch.empty(n_px, dtype=torch.int32).pin_memory()]
    pin_np = [p.numpy() for p in pin]
    e2e_frame = [0]
    if rank == 0:   # a standing read-back order: every frame lands in the pinned arrays (depth / objectId overlap the frame, colour follows it)
        rdr.BindCpuTargets(*pin_np)

    def e2e_step():
        # host camera + knobs in; rt_render; N > 1: rt_gather_frame (colour + depth + objectId to rank 0); present; Synchronize()
        rdr.RenderDirectToPbo(None, W, H, e2e_frame[0], 0.0)
        e2e_frame[0] += 1 if progressive else 0
        if rank == 0:
            rdr.DownloadToCpu(*pin_np)               # Framebuffer.DownloadToCpu: RGBA8 + depth + objId (the gathered ones at N > 1) to host: 12 B/px at every N

    for _ in range(min(2, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    per_step = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        e2e_step()
        per_step.append(time.perf_counter() - ts)
        if os.environ.get("RT_BENCH_DEBUG"):
            print(f"e2e step {per_step[-1] * 1e3:.1f} ms, device {ctx.stats()['lastRenderMs']:.1f} ms", file=sys.stderr)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps       # the mean over exactly K steps is the reported number;
    e2e_median_ms = float(np.median(per_step)) * 1e3      # the median is beside it because a shared host adds rare 10-40 ms stalls to single steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = rays_pb / e2e_s / 1e6
    h2d = 2 * L.CAMERA.itemsize + __import__("ctypes").sizeof(L.RtRenderConfig)
    d2h = 12 * n_px   # RGBA8 + depth + objectId, the same at every N (N > 1: the gathered image on rank 0)

    # ---- roofline of the extend kernels: one frame with per-launch events, one with device counters ---------------
    with torch.cuda.stream(stream):
        ctx.render(cam, cfg_for(L.RT_FLAG_KERNEL_TIMING))
    torch.cuda.synchronize()
    st_t = ctx.stats()
    with torch.cuda.stream(stream):
        ctx.render(cam, cfg_for(L.RT_FLAG_COUNTERS))
    torch.cuda.synchronize()
    st_c = ctx.stats()
    # rays the extend kernels actually TRACED (raysShadow counts the reference's ShadowOcclusion calls, of which the first-vertex
    # ones answered by a shared sun probe were never traced: they are charged neither bytes nor instructions)
    n_closest = st_c["raysPrimary"] + st_c["raysBounce"]
    n_anyhit = st_c["raysAnyHitTraced"]
    n_traced = n_closest + n_anyhit
    ray_bytes = RAY_RECORD_BYTES
    # algorithmic bytes of the traversal (shipped layout): 80 B per wide node fetched, 48 B per primitive record
    # tested, the ray record read per traced ray, 16 B hit record written per closest ray, 4 B visibility flag per any-hit ray
    bvh_bytes = 80 * st_c["wideNodes"] + 48 * (st_c["trisTested"] + st_c["spheresTested"])
    alg_bytes = bvh_bytes + ray_bytes * n_traced + 16 * n_closest + 4 * n_anyhit
    trace_ms = st_t["lastTraceMs"]
    trace_s = trace_ms * 1e-3
    n_ext = max(1, st_t["extendLaunchesTimed"])
    pk = peaks()
    mp = measured_peaks()
    cap = load_capture(name, args, world)
    flops = 176.0 * st_c["wideNodes"] + 51.0 * st_c["trisTested"] + 40.0 * st_c["spheresTested"]   # 8 x 22 per wide node (eight slab tests), 51 per triangle, 40 per sphere test
    rate = (lambda x, scale: (x / trace_s / scale) if trace_s > 0 else None)
    frac = (lambda a, p_: (a / p_) if (a is not None and p_) else None)
    # (1) PRIMARY: instruction issue.  lane-instructions of the extend launches of one frame (ncu smsp__thread_inst_executed.sum from the committed
    # capture of THESE sources) / their live CUDA-event time, against the measured lane-instruction issue peak = IPC x active lanes / 32
    issue = None
    if cap is not None:
        # N > 1: this rank traces its tiles' share of the captured frame's rays with the same kernels - the capture's instructions per ray x the rays it traced
        share = (n_traced / float(cap["rays_traced"])) if (world > 1 and cap.get("rays_traced")) else 1.0
        lane_inst, warp_inst = cap["thread_instructions"] * share, cap["warp_instructions"] * share
        cap = dict(cap, dram_bytes=cap["dram_bytes"] * share)
        issue_peak = mp.get("issue_lane_inst_per_s_T") if mp else None
        issue = {"achieved": rate(lane_inst, 1e12), "peak": issue_peak or 148 * 4 * 32 * 1.965e9 / 1e12, "unit": "Tlane-inst/s",
                 "peak_source": "measured: tests/gpu_peaks.py (profiles/r02_gpu_peaks.json), best of FFMA / FMUL+FADD / LOP3+IADD3 issue" if issue_peak else "derived: 148 SM x 4 schedulers x 32 lanes x 1.965 GHz",
                 "lanes_per_instruction": lane_inst / max(1.0, warp_inst), "warp_instructions_per_traced_ray": warp_inst / max(1, n_traced),
                 "ipc_per_scheduler_live": (warp_inst / trace_s / (148 * 4 * 1.965e9)) if trace_s > 0 else None, "capture": cap["file"]}
        issue["frac"] = frac(issue["achieved"], issue["peak"])
    hbm = {"achieved_algorithmic": rate(alg_bytes, 1e9), "peak": pk["hbm_gbs"], "unit": "GB/s", "peak_source": pk["source"],
           "dram_traffic_per_launch": (cap["dram_bytes"] / n_ext) if cap else None, "dram_gbs": rate(cap["dram_bytes"], 1e9) if cap else None}
    hbm["frac_algorithmic"] = frac(hbm["achieved_algorithmic"], pk["hbm_gbs"])
    hbm["frac_dram"] = frac(hbm["dram_gbs"], pk["hbm_gbs"])
    l2 = {"achieved": rate(bvh_bytes, 1e9), "unit": "GB/s", "what": "node + primitive record bytes fetched (served by L1 / L2: the BVH is cache resident)",
          "peak_random_80B_records": mp.get("l2_random") if mp else None, "peak_stream": mp.get("l2_stream") if mp else None,
          "peak_source": "measured: tests/gpu_peaks.py at a 58 MB working set (profiles/r02_gpu_peaks.json)" if mp else None}
    l2["frac_of_random_gather_peak"] = frac(l2["achieved"], l2["peak_random_80B_records"])
    fp32_peak = (mp.get("fp32_ffma_tflops") if mp else None) or 148 * 128 * 2 * 1.965e9 / 1e12
    fp32 = {"achieved": rate(flops, 1e12), "peak": fp32_peak, "unit": "TFLOP/s", "flops_per_traced_ray": flops / max(1, n_traced),
            "peak_source": "measured FFMA rate: tests/gpu_peaks.py (profiles/r02_gpu_peaks.json)" if mp and mp.get("fp32_ffma_tflops") else "derived: 148 SM x 128 FP32 lanes x 2 x 1.965 GHz",
            "peak_unfused_mul_add": mp.get("fp32_unfused_tflops") if mp else None}
    fp32["frac"] = frac(fp32["achieved"], fp32_peak)
    # without a capture of this workload's instruction count the primary fraction is not known: null, not a stand-in (the L2 gather rate used
    # here before reads > 1 on scenes whose BVH sits in L1)
    primary = issue if issue is not None else {"achieved": None, "peak": (mp.get("issue_lane_inst_per_s_T") if mp else None), "unit": "Tlane-inst/s", "frac": None}
    roofline = {"bound": "issue", "kernel": "k_extend (wide-BVH traversal, closest + any-hit)",
                "achieved": primary["achieved"], "peak": primary["peak"], "unit": primary["unit"], "frac": primary["frac"],
                "peak_source": primary.get("peak_source") or l2["peak_source"], "traffic": hbm["dram_traffic_per_launch"],
                "algorithmic_bytes_per_launch": alg_bytes / n_ext, "launches_per_step": n_ext, "avg_launch_ms": trace_ms / n_ext,
                "extend_share_of_step": trace_ms / st_t["lastRenderMs"] if st_t["lastRenderMs"] else None,
                "rays_traced_per_step": {"closest": n_closest, "any_hit": n_anyhit}, "ray_record_bytes": ray_bytes,
                "nodes_per_ray": st_c["wideNodes"] / max(1, n_traced), "prims_per_ray": (st_c["trisTested"] + st_c["spheresTested"]) / max(1, n_traced),
                "issue": issue, "hbm": hbm, "l2": l2, "fp32": fp32,
                "capture_status": "fresh (source hash matches)" if cap else capture_status(name, args, world),
                "note": "k_extend is bound by instruction issue with partly idle lanes (SIMT divergence), fed from L1 / L2: the primary fraction is lane-instructions per second over the "
                        "measured issue peak (= IPC x active lanes / 32).  hbm.frac_algorithmic divides ALGORITHMIC bytes by the HBM copy peak and is cache-served bandwidth, not DRAM: "
                        "hbm.frac_dram is the true DRAM share; l2 compares the node / primitive fetch rate with the measured L2 random-record gather rate."}

    line = None
    parity = fast = None
    if world > 1:   # the image the REAL gather delivered against a single-context render of the same frame on rank 0
        parity = gathered_parity()
        rdr.configure(rank=rank, worldSize=world)
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            c = cpu_oracle_sample(wl, aovs=True)
            parity = parity_vs_oracle(ctx, wl, cam, cfg_for, c)
            fast = fast_shading_leg(ctx, wl, cam, cfg_for, c, rays_pb) if wl["depth"] > 0 else None
            cpu = {"value": c["mrays"], "unit": "Mrays/s", "cores": c["cores"], "kind": "port", "sample": c["sample"], "seconds": c["seconds"],
                   "note": "CPU restatement of the ILGPU kernels (stand-in for ILGPU CPUAccelerator, which cannot run here)"}
        line = {"metric": "Mrays/s (primary+bounce) at 4K" if W == 3840 else "Mrays/s (primary+bounce)", "value": value, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": workload_config(name, wl, spec, world, tile, spp),
                "frames_per_s": 1e3 / ms_per_step, "mrays_per_s_incl_shadow": rays_all / (ms_per_step * 1e-3) / 1e6,
                "rays_per_step": {"primary_plus_bounce": rays_pb, "traced_incl_shadow": rays_all, "reference_calls_incl_shadow": rays_ref_calls},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "ms_per_step_median": e2e_median_ms},
                "gpu_launches": int(launches), "per_rank": per_rank,
                "parity": parity, "fast_shading": fast, "roofline": roofline, "cpu_baseline": cpu,
                "scene_build_s": {"host_bvh2": t_build, "commit_wide_bvh_upload": t_commit, "commit_force_refit": t_refit}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    rdr.close()


if __name__ == "__main__":
    main()
