#!/bin/bash
# ncu evidence for profiles/ (run on a B200 box: gpurun --timeout 900 -- 'bash tools/ncu_block.sh'):
#   1. per-launch DRAM / L2 / tensor-pipe metrics of the eight GEMMs of one SFNO block at ACE2 1-degree size
#   2. `--set full` of the dominant kernel (mlp.fc2), exported as raw + details CSV (the .ncu-rep stays on the box)
#   3. launch list of the bench command (gpu__time_duration per launch: shares of the step)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg
timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_umma -s 22 -c 8 --csv --log-file gpurun_out/r02_ncu_block_final.csv python tools/gemm_probe.py 2 "{}" > gpurun_out/r02_ncu_block_final.log 2>&1
echo "block metrics exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_umma -s 29 -c 1 -f -o /tmp/fc2_full python tools/gemm_probe.py 2 "{}" > gpurun_out/r02_ncu_fc2_full.log 2>&1
echo "fc2 full exit $?"
ncu -i /tmp/fc2_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_fc2_full_raw.csv 2>/dev/null
ncu -i /tmp/fc2_full.ncu-rep --page details --csv > gpurun_out/r02_ncu_fc2_full_details.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline --sustained-steps 0 --repeats 1 > gpurun_out/r02_launches_final.log 2>&1
echo "launch list exit $?"
