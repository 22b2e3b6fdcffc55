#!/bin/bash
# One-call validation on a B200 box (gpurun): GPU parity suite, then a short bench line (no CPU leg).
#   gpurun --timeout 400 -- 'bash tools/gpu_validate.sh'
# Everything is bounded by `timeout`; logs go to gpurun_out/ (merged back by gpurun).
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout ${ACE_VALIDATE_PYTEST_S:-600} python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 400 -rf --durations=8 > gpurun_out/tests_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 60 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?"
