# LN+modulate on the packed fp32 pipe (in-tree) vs the scalar version (build/ab/lib_simtold.so): kernel tests, then class totals of
# a 50-step decode on one box, both precisions
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "ln" 2>&1 | tail -2
bash profiles/ab_libs.sh fp16 new=cur old=build/ab/lib_simtold.so new2=cur old2=build/ab/lib_simtold.so
bash profiles/ab_libs.sh bf16x3 new=cur old=build/ab/lib_simtold.so
