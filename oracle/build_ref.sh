#!/bin/sh
# ORACLE -- TEST INFRASTRUCTURE ONLY.
# Stages the UNMODIFIED reference (pure Python: src/*.py + config/*.yml) from where it lies under /root/reference into the
# git-ignored oracle/_ref/, so that it travels to the GPU box with the repository snapshot (the box has no /root/reference).
# Nothing is compiled and nothing under oracle/_ref/ is ever committed or imported by the product path: it is the checker /
# the baseline (bench.py --impl reference, cpu_baseline.kind "reference", the gpu_reference comparator, tests marked `reference`).
#   sh oracle/build_ref.sh [/path/to/reference]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-${LAS_REFERENCE:-/root/reference}}"
if [ ! -d "$REF/src" ]; then
    echo "build_ref: $REF/src not found (GPU box?): keeping whatever is already staged in $HERE/_ref" >&2
    exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$REF/src" "$REF/config" "$HERE/_ref/"
chmod -R u+w "$HERE/_ref"
echo "build_ref: staged $(ls "$HERE/_ref/src" | wc -l) source files from $REF into $HERE/_ref"
