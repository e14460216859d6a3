#!/usr/bin/env bash
# Build libs2f.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${S2F_OUT:-$here/../libs2f.so}"
bdir="${S2F_BUILD_DIR:-build}"
[[ "$bdir" = /* ]] || bdir="$here/$bdir"      # relative names live under csrc/, absolute paths are taken as they are
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC 
       --expt-relaxed-constexpr -Xptxas -v ${S2F_EXTRA_FLAGS:-})
objs=()
mkdir -p "$bdir"
for src in "$here"/*.cu; do
  obj="$bdir/$(basename "${src%.cu}").o"
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "$here/common.cuh" -nt "$obj" || "$here/conv_direct.cuh" -nt "$obj" || "$here/conv_tf32.cuh" -nt "$obj" || "$here/dw_tile.cuh" -nt "$obj" || "$here/../../include/s2f.h" -nt "$obj" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" 2> "$obj.log" || { cat "$obj.log"; exit 1; }
  fi
  objs+=("$obj")
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out" "${objs[@]}" -lcudart_static -ldl -lpthread -lrt
echo "built $out"
