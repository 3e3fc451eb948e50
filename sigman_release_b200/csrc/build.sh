#!/bin/bash
# Builds libsgr_b200.so (the C-ABI CUDA library, include/sgr.h) for sm_100a, in-tree next to the package.
# Files listed in EXACT are compiled with --fmad=false: index-determining and per-pixel expressions follow the
# oracle's IEEE fp32 operation order (no compiler-chosen FMA contraction; explicit fmaf where the spec says so).
# The remaining files (gradient kernels compared with a tolerance) keep FMA contraction.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${1:-$HERE/../libsgr_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OBJ="$(mktemp -d)"
trap 'rm -rf "$OBJ"' EXIT
COMMON=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${SGR_NVCC_EXTRA:-})
EXACT="sgr_api sgr_preprocess sgr_binning sgr_blend sgr_blend_simple sgr_knn"
FAST="sgr_gaussian_bwd sgr_attrs sgr_wire"
pids=()
for f in $EXACT; do "$NVCC" "${COMMON[@]}" --fmad=false -c "$HERE/$f.cu" -o "$OBJ/$f.o" & pids+=($!); done
for f in $FAST; do "$NVCC" "${COMMON[@]}" -c "$HERE/$f.cu" -o "$OBJ/$f.o" & pids+=($!); done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ"/*.o
echo "built $OUT"
