#!/bin/bash
# Round-2 encoder follow-up: launch list of one bench step at 32 regions (= one region group of the streaming session) and
# --set full captures of the reworked k_tile_count / k_collect_ops / k_windows / k_cigar_prefix; summaries into gpurun_out/ (copied to profiles/).
set -x
mkdir -p gpurun_out
CMD="python bench.py --regions 32 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-verify --no-files"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2c_launches_variant.csv $CMD > gpurun_out/r2c_launches_variant.log 2>&1
for k in k_tile_count k_collect_ops k_windows k_cigar_prefix; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2c_prof_$k -f $CMD > gpurun_out/r2c_prof_$k.log 2>&1
  ncu -i gpurun_out/r2c_prof_$k.ncu-rep --page raw --csv > gpurun_out/r2c_prof_${k}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r2c_prof_$k.ncu-rep > gpurun_out/r2c_prof_${k}_summary.txt 2>/dev/null
  rm -f gpurun_out/r2c_prof_$k.ncu-rep
done
python scripts/ncu_summary.py gpurun_out/r2c_launches_variant.csv > gpurun_out/r2c_launches_variant_summary.txt 2>/dev/null
ls -la gpurun_out | tail -12
