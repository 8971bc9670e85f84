#!/bin/bash
# Round-2 final captures: launch list of the bench step (8 regions, files leg off), k_lstm_layer (fp16 operands) encoder / decoder
# layers, k_tile_count, k_bgzf_inflate (final version), summaries into gpurun_out/ (copied to profiles/).
set -x
mkdir -p gpurun_out
CMD="python bench.py --regions 8 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-verify --no-files"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_variant.csv $CMD > gpurun_out/r2_launches_variant.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lstm_layer -s 2 -c 1 -o gpurun_out/r2_prof_lstm_enc -f $CMD > gpurun_out/r2_prof_lstm_enc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lstm_layer -s 3 -c 1 -o gpurun_out/r2_prof_lstm_dec -f $CMD > gpurun_out/r2_prof_lstm_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_count -s 1 -c 1 -o gpurun_out/r2_prof_tile_count -f $CMD > gpurun_out/r2_prof_tile_count.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bgzf_inflate -s 1 -c 1 -o gpurun_out/r2_prof_inflate_final -f python scripts/bench_inflate.py --steps 1 > gpurun_out/r2_prof_inflate_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_fetch_device.csv python scripts/bench_inflate.py --steps 1 > gpurun_out/r2_launches_fetch_device.log 2>&1
for n in lstm_enc lstm_dec tile_count inflate_final; do
  ncu -i gpurun_out/r2_prof_$n.ncu-rep --page raw --csv > gpurun_out/r2_prof_${n}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r2_prof_$n.ncu-rep > gpurun_out/r2_prof_${n}_summary.txt 2>/dev/null
done
python scripts/ncu_summary.py gpurun_out/r2_launches_variant.csv > gpurun_out/r2_launches_variant_summary.txt 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_launches_fetch_device.csv > gpurun_out/r2_launches_fetch_device_summary.txt 2>/dev/null
rm -f gpurun_out/r2_prof_lstm_enc.ncu-rep gpurun_out/r2_prof_tile_count.ncu-rep
ls -la gpurun_out | tail -12
