#!/bin/bash
# Round profile capture (run under gpurun): launch list + one full ncu capture of the hot kernels.
# usage: tools_profile.sh <tag>
TAG=${1:-r01}
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 33 -c 44 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/ncu_launch_$TAG.log 2>&1
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:"k_pool_fwd_tile|k_pool_bwd_pixels16|k_sort_cells_warp|k_point_cells|k_scatter_points|k_scan_cells|k_planes_to_rows|k_sort_cells_cta" \
    -s 33 -c 11 -o gpurun_out/prof_$TAG python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
