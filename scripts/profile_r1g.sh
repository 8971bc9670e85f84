#!/bin/bash
# Round-1 front-end profile (final kernels): realignment kernels (k_sw16, k_banded) and get_reads kernels on the polish tiling.
set -x
mkdir -p gpurun_out
CMD="python scripts/bench_frontend.py --contig 200000 --steps 1 --cpu-sample 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_frontend_r1g.csv $CMD > gpurun_out/launches_frontend_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sw16 -c 1 -o gpurun_out/prof_sw16_r1g -f $CMD > gpurun_out/prof_sw16_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_banded -c 1 -o gpurun_out/prof_banded_r1g -f $CMD > gpurun_out/prof_banded_r1g.log 2>&1
ls -la gpurun_out | tail -5
