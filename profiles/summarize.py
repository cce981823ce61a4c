"""Turns the ncu CSV logs brought back in gpurun_out/ into the small summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  profiles/rNN_launches_c4_frame.txt  [first_launch_id]
  python profiles/summarize.py traffic  gpurun_out/extend_traffic.csv profiles/rNN_extend_traffic_c4.json

launches: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ... python tests/gpu_frame_c4.py 64 2`
traffic : `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,
           smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_extend -c 40 --csv --log-file ...
           python tests/gpu_frame_c4.py 64 1`   (every k_extend launch of ONE C4 frame; bench.py reads the JSON for roofline.traffic / issue)
capture : round 2: `... -k regex:k_extend|k_shade|k_connect -c 80 ... python tests/gpu_frame_c4.py 64 1` + the source hash file written in the same gpurun call
           (`python -c "import bench; print(bench.csrc_hash())" > gpurun_out/rNN_csrc_hash.txt`): python profiles/summarize.py capture <csv> profiles/r02_extend_capture_c4.json <hash file> [rays traced per frame]
Profiler times are cold-cache and serialised: compare shares, not absolutes."""
import collections
import csv
import json
import re
import sys


def rows_of(path):
    out = []
    for r in csv.reader(open(path, errors="replace")):
        if len(r) >= 15 and r[0].isdigit():
            out.append(dict(id=int(r[0]), kernel=re.sub(r"^void |\(.*$", "", r[4]), metric=r[12], unit=r[13], value=float(r[14].replace(",", ""))))
    return out


def to_ms(v, unit):
    return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)


def launches(src, dst, first=None):
    rs = [r for r in rows_of(src) if r["metric"] == "gpu__time_duration.sum"]
    ids = sorted({r["id"] for r in rs})
    if first is None:   # the last frame = from the last k_generate_primary on
        first = max(r["id"] for r in rs if r["kernel"].startswith("k_generate_primary"))
    rs = [r for r in rs if r["id"] >= first]
    tot = collections.OrderedDict()
    for r in rs:
        k = tot.setdefault(r["kernel"], [0, 0.0])
        k[0] += 1
        k[1] += to_ms(r["value"], r["unit"])
    total = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): launches {first}..{ids[-1]} = the last frame of the run.  Times are serialised by the profiler: compare SHARES.\n\n")
        f.write(f"{'kernel':24s}{'launches':>9s}{'total ms':>12s}{'avg us':>11s}{'share':>8s}\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:24s}{n:9d}{ms:12.3f}{ms / n * 1e3:11.1f}{ms / total * 100:7.1f}%\n")
        f.write(f"{'TOTAL':24s}{len(rs):9d}{total:12.3f}\n\n# launch by launch (ms):\n")
        for r in rs:
            f.write(f"{r['id']:5d} {r['kernel']:24s}{to_ms(r['value'], r['unit']):10.3f}\n")
    print(open(dst).read()[:1500])


def traffic(src, dst):
    rs = rows_of(src)
    agg = collections.defaultdict(float)
    n = len({r["id"] for r in rs})
    for r in rs:
        v = r["value"]
        if r["metric"] == "gpu__time_duration.sum":
            v = to_ms(v, r["unit"])
        elif r["unit"].lower().startswith(("kbyte", "mbyte", "gbyte")):
            v *= {"k": 1e3, "m": 1e6, "g": 1e9}[r["unit"][0].lower()]
        agg[r["metric"]] += v
    out = {"command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__issue_active... "
                      "--clock-control none -k regex:k_extend -c 40 python tests/gpu_frame_c4.py 64 1",
           "workload": f"C4 frame (3840x2160, 64 spp, depth 8): all {n} k_extend launches of one frame",
           "launches": n, "sum_duration_ms": agg["gpu__time_duration.sum"], "dram_read_bytes": agg["dram__bytes_read.sum"],
           "dram_write_bytes": agg["dram__bytes_write.sum"], "l2_bytes": agg["lts__t_bytes.sum"], "warp_instructions": agg["smsp__inst_executed.sum"],
           "issue_active_pct_mean": agg["smsp__issue_active.avg.pct_of_peak_sustained_active"] / max(1, n),
           "dram_bytes_per_launch": (agg["dram__bytes_read.sum"] + agg["dram__bytes_write.sum"]) / max(1, n)}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


def capture(src, dst, hash_file, rays_traced=None):
    """Round-2 capture summary (bench.py reads it for roofline.issue / roofline.traffic and REFUSES it when csrc_hash differs
    from the tree's): every launch of one C4 frame of the kernels named on the ncu command line, summed per kernel family; the
    top-level fields are the k_extend totals."""
    rs = rows_of(src)
    fam = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for r in rs:
        k = re.sub(r"<.*$", "", r["kernel"])
        v = r["value"]
        if r["metric"] == "gpu__time_duration.sum":
            v = to_ms(v, r["unit"])
        elif r["unit"].lower().startswith(("kbyte", "mbyte", "gbyte")):
            v *= {"k": 1e3, "m": 1e6, "g": 1e9}[r["unit"][0].lower()]
        fam[k][r["metric"]] += v
        ids[k].add(r["id"])

    def one(k):
        a, n = fam[k], len(ids[k])
        return {"launches": n, "sum_duration_ms": a["gpu__time_duration.sum"], "dram_read_bytes": a["dram__bytes_read.sum"], "dram_write_bytes": a["dram__bytes_write.sum"],
                "l2_bytes": a["lts__t_bytes.sum"], "warp_instructions": a["smsp__inst_executed.sum"], "thread_instructions": a["smsp__thread_inst_executed.sum"],
                "lanes_per_instruction": a["smsp__thread_inst_executed.sum"] / max(1.0, a["smsp__inst_executed.sum"]),
                "issue_active_pct_mean": a["smsp__issue_active.avg.pct_of_peak_sustained_active"] / max(1, n)}
    out = {"csrc_hash": open(hash_file).read().strip(), "workload": "C4 frame (3840x2160, 64 spp, depth 8), one frame",
           "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,"
                      "smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_extend|k_shade|k_connect -c 80 python tests/gpu_frame_c4.py 64 1",
           "note": "durations under ncu are cold-cache and serialised: compare shares, not absolutes"}
    # top level = every traversal launch of the frame: the standalone k_extend launches (primary rays, sun probes) + the fused
    # closest / any-hit k_extend_pair launches (one per depth)
    ext = [k for k in fam if k.startswith("k_extend")]
    tot = collections.defaultdict(float)
    for k in ext:
        for m, v in fam[k].items():
            tot[m] += v
    n_ext = sum(len(ids[k]) for k in ext)
    fam["__extend_all__"], ids["__extend_all__"] = tot, set(range(n_ext))
    out.update(one("__extend_all__"))
    del fam["__extend_all__"], ids["__extend_all__"]
    out["kernels"] = {k: one(k) for k in sorted(fam)}
    if rays_traced:   # closest + any-hit rays of the frame (bench.py roofline.rays_traced_per_step): an N-GPU rank scales the counts by its share
        out["rays_traced"] = int(rays_traced)
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else None)
    elif sys.argv[1] == "capture":
        capture(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else None)
    else:
        traffic(sys.argv[2], sys.argv[3])
