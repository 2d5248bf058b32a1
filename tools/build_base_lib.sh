#!/bin/bash
# Build the library of a git revision into ab_libs/ for same-box A/B timing: tools/build_base_lib.sh [REV] [NAME]
REV=${1:-HEAD}; NAME=${2:-base}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$REV" ebsd_vae_b200/csrc include | tar -x -C "$TMP"
mkdir -p "$ROOT/ab_libs"
make -C "$TMP/ebsd_vae_b200/csrc" OUT="$ROOT/ab_libs/libebsd_$NAME.so" >/dev/null 2>&1 && echo "built ab_libs/libebsd_$NAME.so from $REV"
rm -rf "$TMP"
