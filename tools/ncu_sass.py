#!/usr/bin/env python
"""SASS of a line range with executed-instruction counts, from
    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > x.csv
    python tools/ncu_sass.py x.csv file.cu first last"""
import csv
import sys


def num(v):
    try:
        return float(v)
    except ValueError:
        return 0.0


def main(path, fname, a, b):
    rows = list(csv.reader(open(path)))
    hdr, cur, show = None, None, False
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 10:
            continue
        n = len(hdr)
        tail = r[-(n - 4):]
        inst = num(tail[hdr[4:].index("Instructions Executed")])
        samp = num(tail[hdr[4:].index("# Samples")])
        if r[0] != "":
            line = int(r[0])
            show = cur == fname and a <= line <= b
            if show:
                print(f"--- {cur}:{line} {r[1].strip()[:100]}  [{inst/1e6:.2f}M, {samp:.0f} samples]")
        elif show:
            sass = ",".join(r[3:len(r) - (n - 4)])
            print(f"      {sass.strip()[:90]:90s} {inst/1e6:.3f}M {samp:.0f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
