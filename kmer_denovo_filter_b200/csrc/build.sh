#!/usr/bin/env bash
# Build libkdf_sm100.so in-tree (cross-compiles for sm_100a without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${KDF_OUT:-${HERE}/../libkdf_sm100.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -std=c++17 -O3 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-O3,-Wall,-fopenmp -shared \
  ${KDF_NVCC_EXTRA:-} \
  -o "${OUT}" "${HERE}/kdf_kernels.cu" "${HERE}/kdf_host.cpp" "${HERE}/kdf_inflate.cpp" -lz -lgomp
echo "built ${OUT}"
