#!/bin/bash
set -x
mkdir -p gpurun_out
CMD="python bench.py --regions 8 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/launches_r1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_count -s 1 -c 1 -o gpurun_out/prof_tile_count_r1b -f $CMD > gpurun_out/prof_tile_count_r1b.log 2>&1
# decoder LSTM step on the tensor cores: skip pack + 33 encoder steps, take a mid decoder step
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 50 -c 1 -o gpurun_out/prof_tc_gemm_dec_r1b -f $CMD > gpurun_out/prof_tc_gemm_dec_r1b.log 2>&1
# linear_1 (K = 16896): 67th k_tc_gemm launch of the chunk
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 66 -c 1 -o gpurun_out/prof_tc_gemm_lin1_r1b -f $CMD > gpurun_out/prof_tc_gemm_lin1_r1b.log 2>&1
ls -la gpurun_out
