#!/usr/bin/env bash
# Build libdu_b200.so (sm_100a only) next to the Python package.  Usage: csrc/build.sh [extra nvcc flags]
# Objects are rebuilt only when their .cu, any header of csrc/ or include/, or this script is newer, or when they were built
# with other extra flags than this call's (DU_REBUILD=1 forces all).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libdu_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -c -Xcompiler -fPIC
       -Xptxas -v --fmad=true)
SRCS=(du_abi du_moments du_step du_select du_widen du_fused du_fused_pred du_rng)
mkdir -p "$HERE/_obj"
stale() {  # $1 = unit name
  local o="$HERE/_obj/$1.o"
  [ "${DU_REBUILD:-0}" = "1" ] && return 0
  [ -f "$o" ] || return 0
  # an object built with other flags than this call's (e.g. a -DDU_PRED_DEV development build) is stale
  [ "$(cat "$HERE/_obj/$1.flags" 2>/dev/null)" = "$EXTRA" ] || return 0
  local dep
  for dep in "$HERE/$1.cu" "$HERE"/*.cuh "$HERE"/../../include/*.h "${BASH_SOURCE[0]}"; do
    [ "$dep" -nt "$o" ] && return 0
  done
  return 1
}
EXTRA="$*"
pids=()
for f in "${SRCS[@]}"; do
  [ -f "$HERE/$f.cu" ] || continue
  if stale "$f"; then
    ( "$NVCC" "${FLAGS[@]}" "$@" -o "$HERE/_obj/$f.o" "$HERE/$f.cu" > "$HERE/_obj/$f.log" 2>&1 && echo "$EXTRA" > "$HERE/_obj/$f.flags" || { rm -f "$HERE/_obj/$f.o"; exit 1; } ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [ -n "$p" ] && { wait "$p" || rc=1; }; done
if [ $rc -ne 0 ]; then cat "$HERE"/_obj/*.log; exit 1; fi
OBJS=()
for f in "${SRCS[@]}"; do [ -f "$HERE/$f.cu" ] && OBJS+=("$HERE/_obj/$f.o"); done
"$NVCC" -gencode arch=compute_100a,code=sm_100a --shared -o "$OUT" "${OBJS[@]}"
echo "built $OUT"
