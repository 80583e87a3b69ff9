#!/bin/bash
# Encoder SGEMM thread tile: 8 x 16 (<128, 256, 8>, one CTA per SM) vs 8 x 8 (<128, 128, 16>, two CTAs per SM), same library,
# SELFTOK_LINEAR_WIDE=0 selects the old dispatch.  The pre-VQ features must be bit-identical (sha1).
mkdir -p gpurun_out
L=gpurun_out/ab_linear_wide.log
: > $L
for i in 1 2; do
  echo "== wide" >> $L; timeout 200 python profiles/encode_bench.py 2>&1 | tail -1 >> $L
  echo "== 128x128" >> $L; SELFTOK_LINEAR_WIDE=0 timeout 200 python profiles/encode_bench.py 2>&1 | tail -1 >> $L
done
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "linear_f32 or tokens or encode or ragged or pixels_to_tokens or shard" > gpurun_out/ab_linear_wide_tests.log 2>&1
echo "pytest rc=$?" >> $L; tail -2 gpurun_out/ab_linear_wide_tests.log >> $L
cat $L
