#!/usr/bin/env bash
# Builds the reference's OWN main.cpp and Scene.cpp (unmodified, from where they lie) together with integration/Renderer_ptap.cpp
# in place of the reference's Renderer.cpp, and links libptap: the proof that the shim is a drop-in inside the reference tree.
# Assimp is not installed here, so Scene.cpp sees the 80-line OBJ-reading stub in integration/shim/assimp.
# Output: integration/_build/pt_reference_tree (git-ignored; travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$HERE")"
REF="${PTAP_REFERENCE:-/root/reference}/PathTracerAP"
[ -f "$REF/main.cpp" ] || { echo "reference tree not found at $REF" >&2; exit 2; }
OUT="$HERE/_build"; mkdir -p "$OUT"; rm -rf "$OUT/inc"
INC="-I$REF -I$REF/external/include -I$HERE/shim -I$ROOT/include"
nvcc -x cu -std=c++17 -w -gencode arch=compute_100a,code=sm_100a $INC -c "$HERE/Renderer_ptap.cpp" -o "$OUT/Renderer_ptap.o"
g++ -std=c++17 -O2 -ffp-contract=off -w $INC -I/usr/local/cuda/include -c "$REF/Scene.cpp" -o "$OUT/Scene.o"
g++ -std=c++17 -O2 -w $INC -I/usr/local/cuda/include -c "$REF/main.cpp" -o "$OUT/main.o"
g++ "$OUT/main.o" "$OUT/Scene.o" "$OUT/Renderer_ptap.o" -L"$ROOT/pathtracerap_b200" -lptap -Wl,-rpath,'$ORIGIN/../../pathtracerap_b200' \
    -L/usr/local/cuda/lib64 -lcudart -o "$OUT/pt_reference_tree"
echo "built $OUT/pt_reference_tree"
