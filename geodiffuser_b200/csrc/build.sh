#!/bin/sh
# Builds geodiffuser_b200/libgeodiffuser_b200.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
OUT=${GD_OUT:-../libgeodiffuser_b200.so}      # GD_OUT / GD_OBJ / GD_EXTRA: A/B builds of kernel variants (scripts/), not used by build()
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $ARCH $GD_EXTRA"
OBJ=${GD_OBJ:-_obj}
mkdir -p $OBJ
pids=""
# geometry.cu, postprocess.cu: bit-exact artefacts behind a floating-point pipeline -> no FMA contraction
$NVCC $COMMON -fmad=false -c geometry.cu -o $OBJ/geometry.o 2> $OBJ/geometry.log & pids="$pids $!"
$NVCC $COMMON -fmad=false -c postprocess.cu -o $OBJ/postprocess.o 2> $OBJ/postprocess.log & pids="$pids $!"
for f in attention_mma corr_gemm losses elementwise attention_sm100 corr_sm100 body_norm; do
    if [ -f $f.cu ]; then
        $NVCC $COMMON -c $f.cu -o $OBJ/$f.o 2> $OBJ/$f.log & pids="$pids $!"
    fi
done
fail=0
for p in $pids; do wait $p || fail=1; done
if [ $fail -ne 0 ]; then
    grep -h -E "error|Error" -A3 $OBJ/*.log || cat $OBJ/*.log
    exit 1
fi
$NVCC $ARCH -shared -o $OUT $OBJ/*.o
echo "built $(cd .. && pwd)/libgeodiffuser_b200.so"
