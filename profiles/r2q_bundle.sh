#!/bin/bash
# Round-2 closing bundle (one gpurun call): VQ kernel A/B (previous library vs in-tree, ids must be bit-identical), ncu --set full of
# the new VQ kernel, GEMM raster-group knob A/B, the full GPU test suite, then (only if the tests are green) the default bench line.
mkdir -p gpurun_out
L=gpurun_out/r2q_bundle.log
: > $L
for i in 1 2; do
  echo "== vq prev" >> $L; SELFTOK_B200_LIB=$PWD/build/ab/lib_prev.so timeout 200 python profiles/vq_bench.py 2>&1 | tail -1 >> $L
  echo "== vq cur" >> $L; timeout 200 python profiles/vq_bench.py 2>&1 | tail -1 >> $L
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vq_kernel -c 1 -f -o gpurun_out/r2_vq python profiles/vq_bench.py > gpurun_out/r2_vq_ncu.log 2>&1
for gm in 4 2 8 4; do
  echo "== gemm raster GM=$gm" >> $L; SELFTOK_GEMM_GM=$gm timeout 300 python profiles/step_classes.py fp16 50 2>&1 | tail -1 >> $L
done
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; rc=$?
echo "pytest rc=$rc" >> $L; tail -2 gpurun_out/r2q_tests.log >> $L
if [ $rc -eq 0 ]; then
  timeout 600 python bench.py > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; echo "bench rc=$?" >> $L
fi
cat $L
