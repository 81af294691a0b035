#!/bin/bash
# ORACLE — TEST INFRASTRUCTURE ONLY. Compiles the reference's own headers, in place, into
# oracle/_ref/libref.so (git-ignored, travels with the gpurun snapshot). Needs the read-only
# checkout at /root/reference; on the GPU box the prebuilt .so is used as is.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_DIR:-/root/reference}"
OUT="$HERE/../_ref"
[ -d "$REF/include" ] || { echo "reference checkout not found at $REF; skipping"; exit 0; }
mkdir -p "$OUT"
if [ "$1" = "O0" ]; then
  # unoptimised host code: only for tools/texture_fixture.py (the reference's map_Kd result is undefined behaviour that
  # happens to be readable at -O0; see that file)
  nvcc -std=c++17 -O0 -Xcompiler -O0 -gencode arch=compute_100a,code=sm_100a \
       -I "$HERE/stub" -I "$REF/include" -include "$HERE/shim.h" \
       -Xcompiler -fPIC,-w -w -shared -ccbin /usr/bin/g++ \
       -o "$OUT/libref_O0.so" "$HERE/ref_harness.cu"
  echo "built $OUT/libref_O0.so"; exit 0
fi
nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -lineinfo \
     -I "$HERE/stub" -I "$REF/include" -include "$HERE/shim.h" \
     -Xcompiler -fPIC,-w -w -shared -ccbin /usr/bin/g++ \
     -o "$OUT/libref.so" "$HERE/ref_harness.cu"
echo "built $OUT/libref.so"
