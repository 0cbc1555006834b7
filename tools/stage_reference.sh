#!/usr/bin/env bash
# Stage the UNMODIFIED reference where the probe of tests/helpers.py / bench.py looks for it on the GPU box:
# baseline/_ref/{framework,data}  (git-ignored, travels with the gpurun snapshot; the base contract's install location).
# The reference is five script files with no setup.py / pyproject.toml, so `pip install --target baseline/_ref
# /root/reference` has nothing to build (tried first, for the record); the install degenerates to a plain copy.
# Nothing in the product package imports it: it serves the probe-gated drop-in test (tests/test_dropin_gpu.py)
# and bench.py's baseline arms (cpu_baseline.kind = "reference", gpu_eager_baseline.kind = "reference").
set -u
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DST="$ROOT/baseline/_ref"
if [ "${2:-}" = "--remove" ]; then chmod -R u+w "$DST" 2>/dev/null; rm -rf "$DST"; echo "removed $DST"; exit 0; fi
[ -d "$REF/framework" ] || { echo "no reference at $REF"; exit 1; }
mkdir -p "$DST"
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target "$DST" "$REF" \
  > "$DST/pip_install.log" 2>&1 && echo "pip install succeeded" || echo "pip install: not installable (see $DST/pip_install.log); copying the tree"
chmod -R u+w "$DST" 2>/dev/null
rm -rf "$DST/framework" "$DST/data"
cp -r "$REF/framework" "$DST/framework"
cp -r "$REF/data" "$DST/data"
chmod -R u+w "$DST"
find "$DST" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null
( cd "$REF" && find framework data -type f -exec sha256sum {} + ) > "$DST/SHA256SUMS"
( cd "$DST" && sha256sum -c SHA256SUMS --quiet ) && echo "staged unmodified reference in $DST"
