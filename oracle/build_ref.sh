#!/usr/bin/env bash
# Builds the REFERENCE itself (not the restatement) into oracle/_ref/ from the sources
# where they lie under $REF (default /root/reference/Deff2DGPU).  Nothing is copied into
# the repo: the reference header is streamed through sed into the compiler's stdin.
# Outputs (all git-ignored, shipped to the GPU box by gpurun):
#   oracle/_ref/libref_cpu.so    reference host code + kernel body on host threads (g++)
#   oracle/_ref/deff2d_ref_cpu   the reference program (its own main) on host threads
#   oracle/_ref/libref_cuda.so   the reference compiled by nvcc for sm_100a (GPU box only)
#   oracle/_ref/deff2d_ref_cuda  the reference program, nvcc, sm_100a
# The reference's own build instructions are a bare `nvcc <file>` (README.md:33); no build
# system is run.  TEST INFRASTRUCTURE ONLY.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference/Deff2DGPU}"
OUT="$HERE/_ref"
if [ ! -f "$REF/Deff2D.cuh" ]; then
    echo "build_ref: $REF/Deff2D.cuh not found; keeping whatever is in $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
# The only textual change: CUDA's triple-chevron launch is not C++.
shim_cuh() { sed -E 's/([A-Za-z_0-9]+)<<<([^,>]+),([^>]+)>>>\(/SHIM_LAUNCH(\1, \2, \3, /' "$REF/Deff2D.cuh"; }
CXXFLAGS="-O2 -ffp-contract=off -fopenmp -w -I$HERE/cuda_shim -I$REF"

# 1. CPU library (harness entry points)
{ shim_cuh; cat "$HERE/ref_harness.inc"; } |
    g++ $CXXFLAGS -shared -fPIC -x c++ - -o "$OUT/libref_cpu.so"
# 2. CPU program: the reference's own main(), minus its #include line (the header text is
#    already in the stream)
{ shim_cuh; grep -v '#include "Deff2D.cuh"' "$REF/Deff2D.cu"; } |
    g++ $CXXFLAGS -x c++ - -o "$OUT/deff2d_ref_cpu"
echo "build_ref: CPU reference built in $OUT"

# 3./4. the real CUDA build (cross-compiled here, runnable only on the GPU box)
if command -v nvcc >/dev/null 2>&1 && [ "${REF_SKIP_CUDA:-0}" != "1" ]; then
    NVFLAGS="-O3 -w -gencode arch=compute_100a,code=sm_100a -I$REF"
    TMPD="$(mktemp -d)"
    trap 'rm -rf "$TMPD"' EXIT
    # nvcc cannot read a translation unit from stdin: stage the two-line TUs in /tmp
    # (outside the repo); they #include the reference header where it lies.
    printf '#include "%s/Deff2D.cuh"\n#include "%s/ref_harness.inc"\n' "$REF" "$HERE" > "$TMPD/ref_lib.cu"
    nvcc $NVFLAGS -shared -Xcompiler -fPIC "$TMPD/ref_lib.cu" -o "$OUT/libref_cuda.so"
    nvcc $NVFLAGS "$REF/Deff2D.cu" -o "$OUT/deff2d_ref_cuda"
    echo "build_ref: CUDA reference built in $OUT"
fi
