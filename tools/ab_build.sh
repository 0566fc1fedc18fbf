#!/bin/sh
# Developer A/B builds of the kernel library: tools/ab_build.sh NAME "-DFLAG1 -DFLAG2" ...
# -> build/ab/libNAME.so (git-ignored; travels to the GPU box).  Use with B200SP_LIB=build/ab/libNAME.so.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/qat-zstd-plugin_b200/csrc
OUT=$ROOT/build/ab
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
    NAME=$1; FLAGS=$2; shift 2
    $NVCC $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" $FLAGS -c "$SRC/lz77_kernels.cu" -o "$OUT/lz77_$NAME.o"
    $NVCC $ARCH -shared -cudart static -Xlinker -Bsymbolic -o "$OUT/lib$NAME.so" "$OUT/lz77_$NAME.o" "$SRC/seqprod_cuda.o" "$SRC/seqprod_host.o" -lpthread -ldl -lrt
    echo "built $OUT/lib$NAME.so ($FLAGS)"
done
