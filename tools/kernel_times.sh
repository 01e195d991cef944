#!/bin/bash
# usage (under gpurun): tools/kernel_times.sh <variant|base> ...  -> ncu kernel durations of the last timed step
for v in "$@"; do
  P=rcbevdet_b200/lib/variants/lib_$v.so
  [ "$v" == "base" ] && P=rcbevdet_b200/lib/librcbevdet_b200.so
  RCB_LIB_PATH=$P ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file gpurun_out/kt_$v.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > /dev/null 2>&1
  python - "$v" <<'PY'
import csv, sys
v = sys.argv[1]
lines = [l for l in open(f'gpurun_out/kt_{v}.csv') if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if 'rcb::' in r['Kernel Name']]
# the last complete step: from the last k_cells on
last = max(i for i, r in enumerate(rows) if 'k_cells' in r['Kernel Name'])
# the step before the last one is complete for sure
prev = max(i for i, r in enumerate(rows[:last]) if 'k_cells' in r['Kernel Name'])
step = rows[prev:last]
print(v, ' '.join('%s=%.1f' % (r['Kernel Name'].split('(')[0].replace('void ', '').replace('rcb::', '')[:18], float(r['Metric Value']) / 1000) for r in step), 'total=%.1f' % (sum(float(r['Metric Value']) for r in step) / 1000))
PY
done
