#!/bin/bash
# usage (under gpurun): tools/kernel_times.sh <variant|base> ...  -> ncu kernel durations of one timed step
for v in "$@"; do
  P=rcbevdet_b200/lib/variants/lib_$v.so
  [ "$v" == "base" ] && P=rcbevdet_b200/lib/librcbevdet_b200.so
  RCB_LIB_PATH=$P ncu --metrics gpu__time_duration.sum --clock-control none -s 38 -c 9 --csv --log-file gpurun_out/kt_$v.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > /dev/null 2>&1
  python - "$v" <<'PY'
import csv, sys
v = sys.argv[1]
lines = [l for l in open(f'gpurun_out/kt_{v}.csv') if not l.startswith('==')]
print(v, ' '.join('%s=%.1f' % (r['Kernel Name'].split('(')[0].replace('void ', '').replace('rcb::', '')[:18], float(r['Metric Value']) / 1000) for r in csv.DictReader(lines)))
PY
done
