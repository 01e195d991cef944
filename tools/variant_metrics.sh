#!/bin/bash
# usage (under gpurun): tools/variant_metrics.sh <kernel regex> name1 name2 ...
# per variant library: bench RESULT line + a few ncu counters of one launch of the kernel
K=$1; shift
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
for v in "$@"; do
  P=rcbevdet_b200/lib/variants/lib_$v.so
  [ "$v" == "base" ] && P=rcbevdet_b200/lib/librcbevdet_b200.so
  RCB_LIB_PATH=$P python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/err_$v.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('RESULT $v', d['value'], d['ms_per_step'], d['stages_ms'], d.get('variants'))" || tail -5 gpurun_out/err_$v.log
  RCB_LIB_PATH=$P ncu --metrics $M --clock-control none -k regex:$K -s 3 -c 1 --csv --log-file gpurun_out/m_$v.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > /dev/null 2>&1
  python - "$v" <<'PY'
import csv, sys
v = sys.argv[1]
lines = [l for l in open(f'gpurun_out/m_{v}.csv') if not l.startswith('==')]
print('NCU', v, ' '.join('%s=%s' % (r['Metric Name'].split('.')[0].replace('smsp__','').replace('l1tex__','l1_').replace('lts__','l2_'), r['Metric Value']) for r in csv.DictReader(lines)))
PY
done
