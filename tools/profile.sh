#!/bin/bash
# Round profile capture (run under gpurun): launch list + one full ncu capture of every hot-path kernel.
# usage: tools/profile.sh <tag>      then, in the build container: python tools/summarize_profile.py <tag>
TAG=${1:-r02}
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/ncu_launch_$TAG.log 2>&1
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:"k_cells|k_tile_scatter|k_bucket_sort|k_intervals|k_planes_to_rows|k_fwd_cells|k_pool_bwd_pixels16" \
    -s 8 -c 8 -o gpurun_out/prof_$TAG python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
