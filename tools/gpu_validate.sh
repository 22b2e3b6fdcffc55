#!/bin/bash
# One-call validation on a B200 box (gpurun): GPU parity suite, smoke, a short bench line, and the PDL A/B.
#   gpurun --timeout 540 -- 'bash tools/gpu_validate.sh'
# Everything is bounded by `timeout`; logs go to gpurun_out/ (merged back by gpurun).
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
date > gpurun_out/validate_start.txt
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
# 1. the whole -m gpu suite, 6 worker processes (the CPU oracle is the slow side of most tests)
timeout 360 python -m pytest tests -m gpu -q -n 6 -p no:cacheprovider --timeout 150 -rf --durations=12 > gpurun_out/tests_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
# 2. smoke (what the driver runs)
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
# 3. bench line without the CPU leg
timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?"
# 4. programmatic dependent launch A/B (option off by default)
ACE_B200_PDL=1 timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_pdl.log 2>&1
echo "smoke pdl exit $?" >> gpurun_out/smoke_pdl.log
ACE_B200_PDL=1 timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err
echo "bench pdl exit $?"
date > gpurun_out/validate_end.txt
