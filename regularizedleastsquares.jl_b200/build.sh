#!/bin/bash
# Build librls_b200.so for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/csrc"
OUT="$HERE/lib"
mkdir -p "$OUT" "$HERE/build"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr -Xcompiler -fPIC -Xptxas -v"
OBJS=""
pids=()
for f in rls_context rls_gemv rls_normal rls_normal_tma rls_rowstream rls_tc rls_p2p rls_prox rls_svt rls_linop rls_kaczmarz rls_solvers rls_group; do
  if [ "$SRC/$f.cu" -nt "$HERE/build/$f.o" ] || [ -n "$(find "$SRC" "$HERE/../include" -name '*.cuh' -newer "$HERE/build/$f.o" -o -name '*.h' -newer "$HERE/build/$f.o" 2>/dev/null | head -1)" ] || [ ! -f "$HERE/build/$f.o" ]; then
    $NVCC $FLAGS -c "$SRC/$f.cu" -o "$HERE/build/$f.o" > "$HERE/build/$f.log" 2>&1 &
    pids+=($!)
  fi
  OBJS="$OBJS $HERE/build/$f.o"
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
if [ $fail -ne 0 ]; then
  grep -hE "error|Error" -A3 "$HERE"/build/*.log | head -60
  exit 1
fi
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/librls_b200.so" $OBJS -ldl
# a __device__ function called from host code compiles to exit(1): refuse to ship that
if objdump -d "$OUT/librls_b200.so" --no-show-raw-insn | grep -q "call.*<exit@plt>"; then
  echo "ERROR: host code calls a __device__-only function (exit@plt present)"; exit 1
fi
echo "built $OUT/librls_b200.so"
