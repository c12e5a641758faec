#!/bin/bash
# Builds the UNMODIFIED reference (zkDL CUDA sources where they lie under /root/reference) for sm_100 into
# oracle/_ref/ (git-ignored, travels to the GPU box).  Test infrastructure only: the product never links it.
#   oracle/_ref/*.o          reference translation units, Makefile-equivalent flags (-dc -dlto), Makefile:15-37
#   oracle/_ref/demo         the reference's own ./demo binary (Baseline A)
#   oracle/_ref/ref_harness  oracle/ref_harness.cu (ours) linked against the reference objects: feeds injected
#                            challenges/generators through the reference's public API and dumps every result
# The reference has no CPU path, so these binaries only run on the GPU box.
set -e
set -o pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
T=$(python -c 'import torch,os;print(os.path.dirname(torch.__file__))')
mkdir -p "$OUT"
[ -d "$REF" ] || { echo "no $REF here: using prebuilt files in $OUT"; exit 0; }
FLAGS="-arch=sm_100 -std=c++17 -I$T/include -I$T/include/torch/csrc/api/include -I$REF -w"
build_one() {  # $1 = source path, $2 = object
  if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then
    echo "[build_ref] nvcc -dc -dlto $1"; nvcc $FLAGS -x cu -dc -dlto "$1" -o "$2"
  fi
}
JOBS=${JOBS:-4}
pids=()
for f in bls12-381 fr-tensor g1-tensor proof commitment zkfc zkrelu demo; do
  build_one "$REF/$f.cu" "$OUT/$f.o" &
  pids+=($!)
  while [ "$(jobs -rp | wc -l)" -ge "$JOBS" ]; do sleep 1; done
done
build_one "$REF/timer.cpp" "$OUT/timer.o" &
if [ -f "$HERE/ref_harness.cu" ]; then build_one "$HERE/ref_harness.cu" "$OUT/ref_harness.o" & fi
wait
LIBS="-L$T/lib -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -lcudart"
LOPT="--linker-options=-rpath,$T/lib,--copy-dt-needed-entries,--no-as-needed"
COMMON="$OUT/bls12-381.o $OUT/fr-tensor.o $OUT/g1-tensor.o $OUT/proof.o $OUT/commitment.o $OUT/zkfc.o $OUT/zkrelu.o $OUT/timer.o"
link() { # $1 = target, $2 = extra object
  if [ ! -f "$OUT/$1" ] || [ "$2" -nt "$OUT/$1" ]; then
    echo "[build_ref] link $1 (-dlto, several minutes)"; nvcc -arch=sm_100 -std=c++17 $LIBS $COMMON "$2" -o "$OUT/$1" $LOPT -dlto
  fi
}
link demo "$OUT/demo.o" &
if [ -f "$OUT/ref_harness.o" ]; then link ref_harness "$OUT/ref_harness.o" & fi
wait
echo "[build_ref] done"
