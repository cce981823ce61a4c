"""Measured FP32 / issue / L2 peaks of the box (manual tool under gpurun; VERDICT r1 "next" 2, SURVEY.md section 8d "measure it
with an FMA micro-benchmark", "also report vs measured L2 bandwidth").

  python tests/gpu_peaks.py [out.json]      ->  profiles/r02_gpu_peaks.json (committed; bench.py reads it for roofline.fp32 / .l2 / issue)

Builds tests/gpu_peaks/peaks.cu with nvcc for sm_100a (in-tree .so) and runs it on cuda:0 while sampling nvidia-smi clocks."""
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gpu_peaks", "peaks.cu")
SO = os.path.join(HERE, "gpu_peaks", "libgpu_peaks.so")
WORKING_SET = 58e6   # bytes: the wide BVH of the 1M-triangle scene


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                        "-shared", "-o", SO, SRC], check=True)
    return SO


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "profiles", "r02_gpu_peaks.json")
    lib = C.CDLL(build())
    lib.pk_measure.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double), C.c_char_p, C.c_int]
    clocks, stop = [], threading.Event()

    def sample():
        while not stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    clocks.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            stop.wait(0.1)

    th = threading.Thread(target=sample, daemon=True)
    th.start()
    vals = (C.c_double * 16)()
    err = C.create_string_buffer(512)
    t0 = time.time()
    rc = lib.pk_measure(0, WORKING_SET, vals, err, 512)
    stop.set(); th.join(timeout=3)
    if rc != 0:
        raise SystemExit("pk_measure failed: " + err.value.decode())
    sm = sorted(float(c[0]) for c in clocks if c and c[0].replace(".", "").isdigit())
    sms, clock_mhz = int(vals[7]), float(vals[8])
    lanes = sms * 128
    res = {
        "what": "micro-benchmarks of tests/gpu_peaks/peaks.cu on cuda:0 (best of 4-5 repetitions each, CUDA events)",
        "sm_count": sms, "sm_clock_mhz_prop": clock_mhz, "l2_bytes": int(vals[9]),
        "fp32_ffma_tflops": 2 * vals[0], "fp32_ffma2_tflops": 2 * vals[1], "fp32_mul_add_unfused_tflops": 2 * vals[2],
        "ffma_lane_ops_per_s_T": vals[0], "alu_pair_lane_ops_per_s_T": vals[3],
        "issue_lane_inst_per_s_T": {"ffma": vals[0], "ffma2_instructions": vals[1] / 2, "fmul_fadd": 2 * vals[2], "lop3_iadd3": 2 * vals[3]},
        "issue_peak_lane_inst_per_s_T_derived": sms * 4 * 32 * clock_mhz * 1e6 / 1e12,
        "fp32_peak_derived_tflops": lanes * 2 * clock_mhz * 1e6 / 1e12,
        "l2_read_gbs_stream_58MB": vals[4], "l2_read_gbs_random_80B_records_58MB": vals[5], "dram_read_gbs_stream_4GB": vals[6],
        "working_set_bytes": WORKING_SET,
        "clocks": {"samples": len(sm), "sm_mhz_median": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None, "sm_mhz_max": sm[-1] if sm else None,
                   "reasons": sorted({c[3] for c in clocks if len(c) > 3})},
        "seconds": time.time() - t0,
    }
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
