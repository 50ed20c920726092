#!/bin/bash
# Installs the UNMODIFIED reference (pure Python, /root/reference) into oracle/_ref/ so that it travels to the GPU
# box (oracle/_ref is git-ignored, not gpurun-ignored).  Test infrastructure only: tests/, smoke() and bench.py's
# CPU arms may import it, the product never does.  The reference tree is read-only, so pip builds from a /tmp copy.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/frankenz" ] || { echo "no reference at $REF: keeping oracle/_ref as it is"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$REF/." "$TMP/"
rm -rf "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP" >/dev/null
rm -rf "$TMP"
echo "reference installed into $HERE/_ref"
