#!/usr/bin/env bash
# Build libdu_b200.so (sm_100a only) next to the Python package.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libdu_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC
       -Xptxas -v --fmad=true)
mkdir -p "$HERE/_obj"
pids=()
for f in du_abi du_moments du_step du_select du_widen du_fused du_fused_pred; do
  ( "$NVCC" "${FLAGS[@]/--shared/-c}" "$@" -o "$HERE/_obj/$f.o" "$HERE/$f.cu" > "$HERE/_obj/$f.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$HERE"/_obj/*.log; exit 1; fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a --shared -o "$OUT" "$HERE"/_obj/du_abi.o "$HERE"/_obj/du_moments.o \
  "$HERE"/_obj/du_step.o "$HERE"/_obj/du_select.o "$HERE"/_obj/du_widen.o "$HERE"/_obj/du_fused.o "$HERE"/_obj/du_fused_pred.o
echo "built $OUT"
