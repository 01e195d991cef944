#!/usr/bin/env python
"""Turn gpurun_out/prof_<tag>.ncu-rep + launches_<tag>.csv into tracked summaries under profiles/.
Runs in the build container (ncu -i reads reports without a GPU)."""
import csv
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


seen = {}
lines = [f"# ncu --set full summary, tag {tag}", "",
         "One launch per kernel (first occurrence after warm-up), `--clock-control none`, B=8 R50 workload "
         "(`bench.py --steps 3 --warmup 3`; tags r02radar / r02chain: `tools/radar_times.py` / "
         "`tools/chain_step.py chain`).  Times under ncu are cold-cache and serialised: compare shares.", ""]
for d in data:
    name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("rcb::", "")
    if name in seen:
        continue
    seen[name] = d
    lines.append(f"## {name}")
    lines.append("")
    lines.append("| metric | value |")
    lines.append("|---|---|")
    for k in KEEP:
        if k in idx and d[idx[k]] not in ("", "n/a"):
            lines.append(f"| {k} | {d[idx[k]]} {units[idx[k]]} |")
    stalls = [(h, float(d[idx[h]])) for h in hdr
              if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")
              and d[idx[h]] not in ("", "n/a")]
    top = sorted(stalls, key=lambda x: -x[1])[:4]
    lines.append("| top stalls (warps per issue) | " + ", ".join(
        f"{h.split('stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.1f}" for h, v in top) + " |")
    lines.append("")
with open(os.path.join(out_dir, f"{tag}_ncu_kernels.md"), "w") as f:
    f.write("\n".join(lines) + "\n")

if "--kernels-only" in sys.argv:   # captures of other programs (radar, the sort-free chain): per-kernel summary only
    print(open(os.path.join(out_dir, f"{tag}_ncu_kernels.md")).read()[:400])
    sys.exit(0)

# DRAM traffic per launch, grouped by bench.py stage
# bench.py's roofline is quoted on ONE kernel: "fwd" = k_fwd_cells, "bwd" = k_pool_bwd_pixels16
stage_of = {"k_cells": "prepare", "k_tile_scatter": "prepare", "k_bucket_sort": "prepare", "k_intervals": "prepare",
            "k_fwd_cells": "fwd", "k_pool_bwd_pixels16": "bwd", "k_pool_bwd_pixels": "bwd"}
traffic = {}
per_kernel = {}
for name, d in seen.items():
    base = name.split("<")[0]
    b = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
        to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    per_kernel[name] = int(b)
    st = stage_of.get(base)
    if base == "k_planes_to_rows":
        continue
    if st:
        traffic[st] = traffic.get(st, 0) + int(b)
traffic["_per_kernel"] = per_kernel
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (the digest bench.py compares against: a capture of other sources is reported as null)
traffic["csrc_sha256_16"] = bench._kernel_sources_digest()
traffic["_source"] = f"profiles/{tag}_ncu_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
with open(os.path.join(out_dir, "traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)

# launch list
src = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(src):
    rows = [l for l in open(src) if not l.startswith("==")]
    out = ["kernel,duration_us,grid,block"]
    tot = {}
    for r in csv.DictReader(rows):
        n = r["Kernel Name"].split("(")[0].replace("void ", "")
        us = float(r["Metric Value"]) / 1000
        out.append(f"\"{n}\",{us:.2f},\"{r['Grid Size']}\",\"{r['Block Size']}\"")
        tot[n] = tot.get(n, 0) + us
    with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
        f.write("\n".join(out) + "\n")
    s = sum(tot.values())
    with open(os.path.join(out_dir, f"{tag}_launch_shares.md"), "w") as f:
        f.write(f"# kernel shares of the step, tag {tag} (ncu gpu__time_duration.sum, --cache-control none: caches as in the un-profiled run)\n\n| kernel | total us | share |\n|---|---|---|\n")
        for n, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| {n} | {v:.1f} | {100 * v / s:.1f}% |\n")
print(open(os.path.join(out_dir, f"{tag}_launch_shares.md")).read())
print(json.dumps({k: v for k, v in traffic.items() if not k.startswith('_')}))
