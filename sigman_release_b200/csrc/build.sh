#!/bin/bash
# Builds libsgr_b200.so (the C-ABI CUDA library, include/sgr.h) for sm_100a, in-tree next to the package.
# --fmad=false: index-determining and per-pixel expressions follow the oracle's IEEE fp32 operation order
# (no compiler-chosen FMA contraction); explicit fmaf is still used where the spec says so (exp_spec).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${1:-$HERE/../libsgr_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 \
    -Xcompiler -fPIC -shared ${SGR_NVCC_EXTRA:-} \
    -o "$OUT" "$HERE"/sgr_api.cu "$HERE"/sgr_preprocess.cu "$HERE"/sgr_binning.cu "$HERE"/sgr_blend.cu \
    "$HERE"/sgr_blend_simple.cu "$HERE"/sgr_knn.cu
echo "built $OUT"
