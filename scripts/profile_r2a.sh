#!/bin/bash
# Round 2: fp16-operand validation, two-product experiment, inflate kernel stand-alone timing + ncu capture
set -x
mkdir -p gpurun_out
python -m pytest tests/test_nets_gpu.py tests/test_tc_gemm_gpu.py tests/test_pipeline_gpu.py -x -q > gpurun_out/r2_fp16_tests.log 2>&1; tail -4 gpurun_out/r2_fp16_tests.log
python scripts/exp_products.py > gpurun_out/r2_exp_products.json 2> gpurun_out/r2_exp_products.err; tail -16 gpurun_out/r2_exp_products.err
python scripts/bench_inflate.py > gpurun_out/r2_bench_inflate.json 2> gpurun_out/r2_bench_inflate.err; cat gpurun_out/r2_bench_inflate.json
ncu --set full --clock-control none --import-source on -k regex:k_bgzf_inflate -s 1 -c 1 -o gpurun_out/r2_prof_inflate -f python scripts/bench_inflate.py --steps 1 > gpurun_out/r2_prof_inflate.log 2>&1
ncu -i gpurun_out/r2_prof_inflate.ncu-rep --page raw --csv > gpurun_out/r2_prof_inflate_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_inflate.ncu-rep --page source --csv > gpurun_out/r2_prof_inflate_source.csv 2>/dev/null
ls -la gpurun_out | tail -6
