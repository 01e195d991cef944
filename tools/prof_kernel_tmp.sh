python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/plain_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_sort|k_radix_scatter" -s 4 -c 2 -o gpurun_out/prof_bucket -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
