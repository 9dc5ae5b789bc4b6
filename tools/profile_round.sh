#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one eager chignolin step, (2) full-set capture of the hot kernels.
# Usage (on the GPU box): bash tools/profile_round.sh r1b     -> gpurun_out/<tag>_*.csv   (reports stay in /tmp: too big)
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_c2_eager.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"message_fwd_kernel|message_bwd_kernel|gemm_nt_stream|gemm_nn_stream|wgrad_grouped|adam_clip|message9" \
    --launch-skip 600 -c 60 -o /tmp/${TAG}_full -f python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_hot_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
