#!/bin/sh
# Installs the UNMODIFIED reference (pure Python) into baseline/_ref for bench.py's reference
# arm and cpu_baseline leg.  baseline/_ref is git-ignored (never in history) but NOT
# gpurun-ignored, so it travels to the GPU box with the snapshot.  /root/reference is
# read-only and setuptools writes build/ and *.egg-info into the source tree: install from a copy.
set -e
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
[ -d "$REF/melvin" ] || { echo "no reference checkout at $REF"; exit 1; }
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/src"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP/src"
rm -rf "$TMP"
echo "installed $(ls "$ROOT/baseline/_ref")"
