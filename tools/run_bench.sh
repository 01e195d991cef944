#!/bin/bash
# usage (under gpurun): tools/run_bench.sh [pytest]  -> (optionally the GPU tests) + a short bench + warm-cache kernel times
set -o pipefail
if [ "$1" == "pytest" ]; then python -m pytest tests -m gpu -q 2>&1 | tail -4; fi
tools/bench_variants.sh base
tools/kernel_times.sh base
