#!/bin/bash
# usage (under gpurun): tools/bench_variants.sh name1 name2 ...   -> one RESULT line per variant library
for v in "$@"; do
  P=rcbevdet_b200/lib/variants/lib_$v.so
  [ "$v" == "base" ] && P=rcbevdet_b200/lib/librcbevdet_b200.so
  RCB_LIB_PATH=$P python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/err_$v.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('RESULT $v', d['value'], d['ms_per_step'], d['stages_ms'], d.get('variants'))" || tail -5 gpurun_out/err_$v.log
done
