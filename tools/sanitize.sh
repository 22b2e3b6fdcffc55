#!/bin/bash
# compute-sanitizer over the GPU test-suite on a B200 box (gpurun):  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
#   1. memcheck over every -m gpu test except the full-size / long ones (their kernels are the same template instantiations
#      the mid-size cases run; under the sanitizer they would take tens of minutes);
#   2. racecheck + synccheck over the SIMT GEMM, elementwise, ConditionalLayerNorm, corrector, metrics and HEALPix kernels
#      (ACE_B200_FORCE_SIMT=1: these tools do not model the asynchronous tcgen05 / TMA proxies of the tensor-core kernel).
# Logs: gpurun_out/r02_sanitizer_{memcheck,racecheck,synccheck}.log; the summaries are copied to profiles/.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
SKIP='not full_size and not baseline_configs and not quarter_degree and not rollout_40 and not test_block_speed and not faster_and_smaller and not (test_net_tcgen05_path_vs_oracle and 180)'
T1=${ACE_SAN_MEMCHECK_S:-700}
T2=${ACE_SAN_RACE_S:-300}
TOOLS=${ACE_SAN_TOOLS:-memcheck racecheck synccheck}
if [[ " $TOOLS " == *" memcheck "* ]]; then
timeout $T1 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 --error-exitcode 9 \
  python -m pytest tests -m gpu -q -p no:cacheprovider -k "$SKIP" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
fi
RTESTS="tests/test_gpu_corrector.py tests/test_gpu_metrics.py tests/test_gpu_healpix.py tests/test_gpu_sfno.py::test_net_matches_reference_vectors tests/test_gpu_csfno.py::test_reference_goldens tests/test_gpu_stepper.py::test_step_matches_oracle"
for tool in racecheck synccheck; do
  [[ " $TOOLS " == *" $tool "* ]] || continue
  ACE_B200_FORCE_SIMT=1 timeout $T2 compute-sanitizer --tool $tool --target-processes all --print-limit 20 --error-exitcode 9 \
    python -m pytest $RTESTS -q -p no:cacheprovider -k "not 180" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/r02_sanitizer_$tool.log
done
for t in $TOOLS; do echo "== $t"; grep -E "ERROR SUMMARY|passed|failed|exit|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_$t.log | tail -6; done
