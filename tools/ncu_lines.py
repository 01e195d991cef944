#!/usr/bin/env python
"""Per-source-line summary of an ncu report's source page.

    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > x.csv
    python tools/ncu_lines.py x.csv [top_n]

Prints, per CUDA source line: warp instructions executed, share, stall samples, the dominant
stall reasons and shared/global wavefront counters -- the view that finds where a kernel's
instructions and stalls come from."""
import csv
import sys
from collections import defaultdict


def main(path, top=45, regions=None):
    rows = list(csv.reader(open(path)))
    cur_file, hdr = None, None
    agg = defaultdict(lambda: defaultdict(float))
    src = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] == "":      # SASS row: already counted in its CUDA line row
            continue
        key = (cur_file, int(r[0]))
        src[key] = r[1].strip()
        # the source text may hold commas / quotes that split the row: align the metrics from the end
        for name, v in zip(hdr[::-1][:len(hdr) - 4], r[::-1]):
            try:
                agg[key][name] += float(v)
            except ValueError:
                pass
    tot_inst = sum(a["Instructions Executed"] for a in agg.values())
    tot_samp = sum(a["# Samples"] for a in agg.values())
    print(f"total warp instructions {tot_inst/1e6:.2f} M, stall samples {tot_samp:.0f}")
    stall_names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    order = sorted(agg, key=lambda k: -(agg[k]["Instructions Executed"] / max(tot_inst, 1) + agg[k]["# Samples"] / max(tot_samp, 1)))
    if regions:   # "name:file:first-last,..." -> instruction / sample share of each line range
        for spec in regions.split(","):
            name, f, rng = spec.split(":")
            a, b = (int(x) for x in rng.split("-"))
            ks = [k for k in agg if k[0] == f and a <= k[1] <= b]
            vi = sum(agg[k]["Instructions Executed"] for k in ks)
            vs = sum(agg[k]["# Samples"] for k in ks)
            print(f"region {name:12s} inst {vi/1e6:6.2f}M {100*vi/max(tot_inst,1):5.1f}%  samples {100*vs/max(tot_samp,1):5.1f}%")
    for k in order[:top]:
        a = agg[k]
        stalls = sorted(((a[s], s[6:]) for s in stall_names if a[s] > 0), reverse=True)[:3]
        st = " ".join(f"{n}:{100*v/max(a['# Samples'],1):.0f}%" for v, n in stalls)
        print(f"{k[0][:18]:18s}:{k[1]:4d} inst {a['Instructions Executed']/1e6:6.2f}M {100*a['Instructions Executed']/max(tot_inst,1):5.1f}%  "
              f"samp {100*a['# Samples']/max(tot_samp,1):5.1f}%  shwf {a['L1 Wavefronts Shared']/1e6:5.2f}M gltag {a['L1 Tag Requests Global']/1e6:5.2f}M  [{st}]  {src[k][:70]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45, sys.argv[3] if len(sys.argv) > 3 else None)
