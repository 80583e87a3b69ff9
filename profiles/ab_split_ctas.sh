# split-bf16 attention at 2 CTAs per SM (in-tree) vs 1 CTA per SM (-DSELFTOK_ATTN5_SPLIT_CTAS=1): validate, then A/B on one box
echo "=== kernel tests (in-tree)"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -2
for ns in 3 1; do echo "--- multi-item ns=$ns"; ATTN_CHECK_B=64 timeout 200 python profiles/attn_multiitem_check.py $ns 40 64 100 128 200 600 768 2>&1 | tail -8; done
bash profiles/ab_libs.sh bf16x3 cur=cur one=build/ab/lib_sc1.so cur2=cur one2=build/ab/lib_sc1.so
