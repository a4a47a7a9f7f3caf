#!/usr/bin/env bash
# Install the UNMODIFIED reference (pure Python, /root/reference) into baseline/_ref so that
# `bench.py --impl reference` and bench.py's cpu_baseline leg time the reference's own
# `dewi.index.DewiIndex(use_ann=False).search` (src/dewi/index.py:50-51 -> backends.py:414-481) and
# `dewi.scorer.DewiScorer` -- not a port.  baseline/_ref is git-ignored (reference sources never enter
# this repository's history) but NOT gpurun-ignored, so the install travels to the GPU box.
#
# Offline: --no-index; --no-deps because the reference's optional dependencies (hnswlib, faiss-cpu, ...)
# have no wheels here and are not needed by the exact path; the source tree is read-only, so the
# build runs from a copy under /tmp.
set -euo pipefail
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
if [ ! -d "$REF/src/dewi" ]; then
  echo "vendor_reference: $REF/src/dewi not found (nothing to install; the port under oracle/ is used instead)" >&2
  exit 0
fi
TMP="$(mktemp -d /tmp/dewi_ref.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/reference"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$ROOT/baseline/_ref" "$TMP/reference"
# the install must be byte-identical to the reference's sources
for f in "$REF"/src/dewi/*.py; do
  cmp -s "$f" "$ROOT/baseline/_ref/dewi/$(basename "$f")" || { echo "vendor_reference: $(basename "$f") differs" >&2; exit 1; }
done
echo "vendor_reference: installed $(ls "$ROOT/baseline/_ref/dewi" | wc -l) files into baseline/_ref/dewi"
