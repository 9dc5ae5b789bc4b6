#!/bin/bash
# round-2 ncu evidence: full-set capture of the hot kernels at the chignolin shapes (tools/profile_kernels.py keeps < 200 MB
# resident) -> gpurun_out/<tag>_ncu_hot_raw.csv, source pages of the two tensor-core kernels, and the launch list of one eager
# step.  Usage (GPU box): bash tools/profile_round2.sh r2
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"message_tc_fwd_kernel|message_fwd_kernel|message_bwd_kernel|gemm_tc_kernel|gemm_nt_stream|gemm_nn_stream|wgrad_grouped|adam_clip|message9" \
    --launch-skip 14 -c 14 -o gpurun_out/${TAG}_hot -f python tools/profile_kernels.py > gpurun_out/${TAG}_hot.log 2>&1
ncu -i gpurun_out/${TAG}_hot.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_hot_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_hot.ncu-rep --page source --csv -k regex:"message_tc_fwd_kernel" > gpurun_out/${TAG}_ncu_source_message_tc_fwd.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_hot.ncu-rep --page source --csv -k regex:"gemm_tc_kernel" > gpurun_out/${TAG}_ncu_source_gemm_tc.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_c2_eager.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-extra > gpurun_out/${TAG}_launches.log 2>&1
ls -la gpurun_out/ | tail -8
