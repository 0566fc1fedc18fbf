#!/bin/sh
# oracle/refstub/build_ref.sh <path to reference qatseqprod.c>
# Builds oracle/_ref/libqzstd_ref.so = the reference plugin itself (unmodified source, compiled where
# it lies) + the fake QAT driver.  Only possible where /root/reference is mounted; the .so then
# travels to the GPU box with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
set -e
REF="$1"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
mkdir -p "$OUT"
REFDIR="$(dirname "$REF")"
${CC:-gcc} -O2 -fPIC -shared -w -Wl,-Bsymbolic \
    -I"$HERE" -I"$HERE/../../include" -I"$REFDIR" \
    -DDEBUGLEVEL=0 -DREF_SRC="\"$REF\"" \
    "$HERE/ref_wrap.c" "$HERE/fakeqat.c" -o "$OUT/libqzstd_ref.so" -l:libzstd.so.1 -lpthread
echo "built $OUT/libqzstd_ref.so from $REF"
