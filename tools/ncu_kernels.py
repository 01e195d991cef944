"""Per-kernel mean durations from an ncu --csv launch list (gpu__time_duration.sum)."""
import collections, csv, sys
rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
t = collections.OrderedDict()
for r in rows:
    n = r['Kernel Name'].split('(')[0].replace('void ', '').replace('rcb::', '')
    t.setdefault(n, []).append(float(r['Metric Value']) / 1000)
for n, v in t.items():
    print(f"{n[:60]:60s} n={len(v):3d} mean={sum(v)/len(v):8.1f} us min={min(v):8.1f}")
