#!/bin/bash
# Build a variant library for same-box A/B runs:   profiles/mk_variant.sh <name> <file.cu> [<file.cu> ...]
# Each given .cu replaces the in-tree source of the same role (its basename must start with attn_tc5 / gemm_tc / kernels_simt /
# engine, e.g. build/ab/gemm_tc_hint.cu); everything else links from the in-tree objects (run the normal build first).
# EXTRA="-DFOO" adds compiler flags (macro-selected variants of an in-tree source).
# Result: build/ab/lib_<name>.so  ->  SELFTOK_B200_LIB=$PWD/build/ab/lib_<name>.so python profiles/step_classes.py
set -e
name=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/selftoktokenizer_b200/csrc
mkdir -p $ROOT/build/ab
F="$EXTRA -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -I$C -diag-suppress 177"
declare -A obj=( [kernels_simt]=$C/kernels_simt.o [gemm_tc]=$C/gemm_tc.o [attn_tc5]=$C/attn_tc5.o [engine]=$C/engine.o [vae]=$C/vae.o )
for src in "$@"; do
  b=$(basename $src .cu)
  role=""
  for r in attn_tc5 kernels_simt gemm_tc engine vae; do case $b in $r*) role=$r; break;; esac; done
  [ -n "$role" ] || { echo "cannot map $src to a source role"; exit 1; }
  nvcc $F -c $src -o $ROOT/build/ab/${b}_$name.o
  obj[$role]=$ROOT/build/ab/${b}_$name.o
done
nvcc -shared -o $ROOT/build/ab/lib_$name.so ${obj[kernels_simt]} ${obj[gemm_tc]} ${obj[attn_tc5]} ${obj[engine]} ${obj[vae]} \
  -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
echo built build/ab/lib_$name.so
