# encoder kernels: in-tree vs build/ab/lib_simtold.so (the previous kernels_simt.cu); the pre-VQ features must be bit-identical
for i in 1 2; do
echo "== new"; timeout 300 python profiles/encode_bench.py 2>&1 | tail -1
echo "== old"; SELFTOK_B200_LIB=$PWD/build/ab/lib_simtold.so timeout 300 python profiles/encode_bench.py 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention_f32 or fp32 or linear" 2>&1 | tail -2
