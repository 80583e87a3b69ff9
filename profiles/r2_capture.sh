#!/bin/bash
# Round-2 profile captures (run on the B200 box through gpurun; outputs under gpurun_out/):
#   1. ncu --set full of the 9 hot kernels of MMDiT layer 1 at sampler step 0 (batch 64, fp16)
#   2. launch list (gpu__time_duration.sum) of one full bench step
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2|attention_tc5|ln_mod" -s 9 -c 9 -f -o gpurun_out/r2_layer1 python profiles/one_step.py fp16 1 > gpurun_out/r2_layer1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out/r2_*
