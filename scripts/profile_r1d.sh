#!/bin/bash
set -x
mkdir -p gpurun_out
CMD="python bench.py --regions 8 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/launches_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm_p -s 10 -c 1 -o gpurun_out/prof_tcp_enc_r1d -f $CMD > gpurun_out/prof_tcp_enc_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm_p -s 50 -c 1 -o gpurun_out/prof_tcp_dec_r1d -f $CMD > gpurun_out/prof_tcp_dec_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm_p -s 66 -c 1 -o gpurun_out/prof_tcp_lin1_r1d -f $CMD > gpurun_out/prof_tcp_lin1_r1d.log 2>&1
PCMD="python bench.py --config polish --regions 1500 --block 250 --steps 1 --warmup 1 --no-verify"   # (round 1 used scripts/bench_polish.py, since folded into bench.py)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_polish_r1d.csv $PCMD > gpurun_out/launches_polish_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gru_cluster -s 5 -c 1 -o gpurun_out/prof_gru_cluster_dec_r1d -f $PCMD > gpurun_out/prof_gru_cluster_dec_r1d.log 2>&1
ls -la gpurun_out | tail -8
