#!/bin/bash
# Round-1 final variant-path profile: persistent LSTM layer kernel (k_lstm_layer), launch list of bench.py
set -x
mkdir -p gpurun_out
CMD="python bench.py --regions 8 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv $CMD > gpurun_out/launches_r1h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lstm_layer -s 2 -c 1 -o gpurun_out/prof_lstm_enc_r1h -f $CMD > gpurun_out/prof_lstm_enc_r1h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lstm_layer -s 3 -c 1 -o gpurun_out/prof_lstm_dec_r1h -f $CMD > gpurun_out/prof_lstm_dec_r1h.log 2>&1
ls -la gpurun_out | tail -4
