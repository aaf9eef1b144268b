#!/usr/bin/env bash
# Builds machupx_b200/liblls.so (CUDA, sm_100a) in-tree.  Used by __graft_entry__.build().
set -euo pipefail
cd "$(dirname "$0")"
SRC=machupx_b200/csrc
OUT=machupx_b200/liblls.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -I include -I $SRC \
  $SRC/*.cu -o $OUT -lcudart "$@"
echo "built $OUT"
