"""Manual tuning sweep (under gpurun): build k_extend variants with different macros and time C3 / C4-lite (or VARIANT_SCRIPT)."""
import json
import os
import subprocess
import sys

sys.path.insert(0, ".")
from ilgpu_raytracing_b200 import build  # noqa: E402

variants = [a.split(",") for a in sys.argv[1:]] or [["base"]]
results = {}
for v in variants:
    name = "_".join(v)
    so = f"/tmp/librtcore_{name}.so"
    defs = [f"-D{d}" for d in v if "=" in d]
    cmd = ["nvcc"] + build.NVCC_FLAGS + defs + ["-o", so] + build.CORE_SRCS
    subprocess.run(cmd, check=True)
    env = dict(os.environ, RTCORE_B200_LIB=so)
    if os.environ.get("VARIANT_PYTEST"):   # parity of the variant build: the whole GPU suite against the oracle
        t = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q"], env=env, capture_output=True, text=True)
        print(name, "pytest:", t.stdout.strip().splitlines()[-1] if t.stdout.strip() else t.stderr[-300:], flush=True)
    out = subprocess.run([sys.executable] + os.environ.get("VARIANT_SCRIPT", "tests/gpu_variant_run.py").split(), env=env, capture_output=True, text=True)
    try:
        results[name] = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        results[name] = {"error": out.stderr[-400:]}
    print(name, results[name], flush=True)
json.dump(results, open("gpurun_out/variants.json", "w"), indent=1)
