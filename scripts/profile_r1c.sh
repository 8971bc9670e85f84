#!/bin/bash
set -x
mkdir -p gpurun_out
CMD="python bench.py --regions 8 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/launches_r1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_count -s 1 -c 1 -o gpurun_out/prof_tile_count_r1c -f $CMD > gpurun_out/prof_tile_count_r1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 10 -c 1 -o gpurun_out/prof_tc_gemm_enc_r1c -f $CMD > gpurun_out/prof_tc_gemm_enc_r1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 50 -c 1 -o gpurun_out/prof_tc_gemm_dec_r1c -f $CMD > gpurun_out/prof_tc_gemm_dec_r1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 66 -c 1 -o gpurun_out/prof_tc_gemm_lin1_r1c -f $CMD > gpurun_out/prof_tc_gemm_lin1_r1c.log 2>&1
ls -la gpurun_out | tail -12
