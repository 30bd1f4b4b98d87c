#!/bin/bash
# Compiles the reference's own kernel files, in place from /root/reference, for sm_100a
# and links them with oracle/ref_driver.cu into oracle/_ref/dipper_ref.
# Outputs go only to oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).
# No reference source is copied into this repository.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${DIPPER_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF not present, skipping"; exit 0; fi
mkdir -p "$OUT"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -rdc=true --extended-lambda -std=c++17 -O3 -w -ccbin /usr/bin/g++ -I$HERE/stubs -I$REF/src"
SRCS="src/MSA.cu src/mash.cu src/neighborJoining.cu src/placement_close_k.cu src/placement.cu src/matrix_reader.cu src/tree.cpp"
pids=""
# the divide-and-conquer twins (same base names as the files above: objects get a dc_ prefix);
# DC/mash.cpp and DC/placement_close_k.cpp need real TBB and nothing in the kernel objects calls them
DCSRCS="src/divide_and_conquer/placement_close_k.cu src/divide_and_conquer/msa.cu src/divide_and_conquer/mash.cu"
for s in $DCSRCS; do
  o="$OUT/dc_$(basename ${s%.*}).o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    ( $NVCC $FLAGS -x cu -dc "$REF/$s" -o "$o" ) &
    pids="$pids $!"
  fi
done
for s in $SRCS; do
  o="$OUT/$(basename ${s%.*}).o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    ( $NVCC $FLAGS -x cu -dc "$REF/$s" -o "$o" ) &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC $FLAGS -dc "$HERE/ref_driver.cu" -o "$OUT/ref_driver.o"
$NVCC -gencode arch=compute_100a,code=sm_100a -rdc=true -ccbin /usr/bin/g++ -o "$OUT/dipper_ref" "$OUT"/*.o
echo "build_ref: built $OUT/dipper_ref"
# The drop-in test: the SAME main (ref_driver.cu), the reference's unmodified mash_placement.cuh / tree.cpp /
# matrix_reader.cu, but the struct API implemented by integration/mash_placement_b200.cpp on libdipper_b200.so.
LIB="$HERE/../dipper_b200"
if [ -f "$LIB/libdipper_b200.so" ]; then
  mkdir -p "$OUT/shim"
  $NVCC $FLAGS -DDIPPER_REF_SHIM -dc "$HERE/ref_driver.cu" -o "$OUT/shim/ref_driver.o"
  $NVCC $FLAGS -x cu -dc "$HERE/../integration/mash_placement_b200.cpp" -o "$OUT/shim/mash_placement_b200.o"
  $NVCC -gencode arch=compute_100a,code=sm_100a -rdc=true -ccbin /usr/bin/g++ -o "$OUT/dipper_ref_b200" \
      "$OUT/shim/ref_driver.o" "$OUT/shim/mash_placement_b200.o" "$OUT/tree.o" "$OUT/matrix_reader.o" \
      -L"$LIB" -ldipper_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../dipper_b200'
  echo "build_ref: built $OUT/dipper_ref_b200 (reference main + shim + libdipper_b200.so)"
fi
