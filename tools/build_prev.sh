#!/bin/bash
# Build the CUDA sources of another commit into aladin_b200/libalad_b200_prev.so (git-ignored, travels to the GPU box) for a
# same-box A/B against the current build (tools/ab_prev.sh; ALAD_B200_LIB selects the library).  Usage: tools/build_prev.sh [commit]
set -e
REV=${1:-HEAD~1}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$REV" aladin_b200/csrc include | tar -x -C "$TMP"
SRCS=$(cd "$ROOT" && python -c "from aladin_b200 import build; print(' '.join(build.SOURCES))")
FILES=""
for s in $SRCS; do [ -f "$TMP/aladin_b200/csrc/$s" ] && FILES="$FILES $TMP/aladin_b200/csrc/$s"; done
nvcc -shared -Xcompiler -fPIC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --threads 4 \
  -I "$TMP/include" -I "$TMP/aladin_b200/csrc" -o "$ROOT/aladin_b200/libalad_b200_prev.so" $FILES
rm -rf "$TMP"
echo "built aladin_b200/libalad_b200_prev.so from $REV"
