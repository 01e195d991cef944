#!/bin/bash
# usage (under gpurun): tools/chain_times.sh <mode> [calib]  -> ncu kernel durations of the last chain step (warm L2)
m=${1:-chain}
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/ct_$m.csv python tools/chain_step.py $m 4 $2 > /dev/null 2>&1
python - "$m" <<'PY'
import csv, sys
m = sys.argv[1]
lines = [l for l in open(f'gpurun_out/ct_{m}.csv') if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if 'rcb::' in r['Kernel Name']]
first = [i for i, r in enumerate(rows) if 'k_cells' in r['Kernel Name'] and (i == 0 or 'k_cells' not in rows[i-1]['Kernel Name'])]
# steps start at a k_cells that follows a backward kernel
starts = [i for i in first if i == 0 or 'bwd' in rows[i-1]['Kernel Name'] or 'planes' in rows[i-1]['Kernel Name']]
step = rows[starts[-1]:]
print(m, ' '.join('%s=%.1f' % (r['Kernel Name'].split('(')[0].split('<')[0].replace('void ', '').replace('rcb::', '')[:18], float(r['Metric Value']) / 1000) for r in step), 'total=%.1f' % (sum(float(r['Metric Value']) for r in step) / 1000))
PY
