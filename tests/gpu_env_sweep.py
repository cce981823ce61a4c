"""Manual tuning sweep (under gpurun): run tests/gpu_variant_run.py under different environment settings."""
import json
import os
import subprocess
import sys

res = {}
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split(","):
        if "=" in kv:
            k, v = kv.split("=")
            env[k] = v
    out = subprocess.run([sys.executable, "tests/gpu_variant_run.py"], env=env, capture_output=True, text=True)
    try:
        res[spec] = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        res[spec] = {"error": out.stderr[-300:]}
    print(spec, res[spec], flush=True)
json.dump(res, open("gpurun_out/env_sweep.json", "w"), indent=1)
