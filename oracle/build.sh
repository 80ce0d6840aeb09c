#!/usr/bin/env bash
# Build the oracle's C/OpenMP twin into oracle/_build/ (git-ignored; travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "${HERE}/_build"
gcc -O3 -march=x86-64-v2 -fopenmp -fPIC -shared -Wall -o "${HERE}/_build/libkdf_oracle.so" "${HERE}/kdf_oracle.c"
echo "built ${HERE}/_build/libkdf_oracle.so"
