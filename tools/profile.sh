#!/bin/bash
# Round profile capture (run under gpurun): launch list + one full ncu capture of the hot kernels.
# usage: tools_profile.sh <tag>
TAG=${1:-r01}
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 29 -c 36 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/ncu_launch_$TAG.log 2>&1
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:"k_pool_fwd_tile|k_pool_bwd_pixels16|k_cells_hist|k_digit_offsets|k_radix_scatter|k_bucket_sort|k_intervals|k_planes_to_rows" \
    -s 27 -c 9 -o gpurun_out/prof_$TAG python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
