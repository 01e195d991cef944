#!/bin/bash
for v in variants/lib_*.so; do
  cp $v rcbevdet_b200/lib/librcbevdet_b200.so
  echo -n "$v: "
  python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['stages_ms']['fwd'], d['variants'])"
done
