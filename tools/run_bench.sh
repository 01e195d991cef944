#!/bin/bash
# usage: tools_run_bench.sh [pytest]  -> runs (optionally tests) + short bench + launch list on the GPU box
set -o pipefail
if [ "$1" == "pytest" ]; then python -m pytest tests -m gpu -q -x 2>&1 | tail -4; fi
python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('RESULT', d['value'], d['ms_per_step'], d['stages_ms'], 'host', d.get('host_launch_ms_per_step'), 'e2e', d['e2e']['value'], d['clocks'])"
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 42 -c 14 --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/ncu.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches_tmp.csv') if not l.startswith('==')]
for row in csv.DictReader(lines):
    print('%-52s %9.1f us  %s %s' % (row['Kernel Name'][:52], float(row['Metric Value'])/1000, row['Grid Size'], row['Block Size']))
PY
