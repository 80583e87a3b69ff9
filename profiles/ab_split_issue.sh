for L in s2 s3; do
  export SELFTOK_B200_LIB=$PWD/build/ab/lib_$L.so
  echo "=== $L kernel tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -3
  for ns in 1 3; do echo "--- multi-item ns=$ns"; ATTN_CHECK_B=64 timeout 200 python profiles/attn_multiitem_check.py $ns 40 64 100 128 200 600 768 2>&1 | tail -8; done
done
unset SELFTOK_B200_LIB
bash profiles/ab_libs.sh fp16 cur=cur s2=build/ab/lib_s2.so s3=build/ab/lib_s3.so cur2=cur s2b=build/ab/lib_s2.so s3b=build/ab/lib_s3.so
