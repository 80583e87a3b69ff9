# LN+modulate variants against the scalar one-row-per-warp version (build/ab/lib_simtold.so): kernel + parity tests, then class
# totals of a 50-step decode on one box
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "ln" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tiny_velocity or mid or full_decode" 2>&1 | tail -2
bash profiles/ab_libs.sh fp16 new=cur old=build/ab/lib_simtold.so new2=cur old2=build/ab/lib_simtold.so
