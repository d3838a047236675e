#!/bin/bash
# Install the UNMODIFIED reference package (rewaifu/resselt, pure Python) into baseline/_ref so that
# `bench.py --impl reference` and the `gpu_library_baseline` leg can import it on the GPU box
# (baseline/_ref is git-ignored but travels with the gpurun snapshot).
#
# The reference's build backend (hatchling) is not in this image's offline wheelhouse, so the documented
#   pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
# fails with "No module named 'hatchling'".  /root/reference is read-only, so we install from a copy under /tmp whose
# pyproject.toml names setuptools as the backend instead; the package sources (resselt/**) are byte-identical
# (checked with diff below).  --no-deps: torch / einops / numpy / safetensors come from the image.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/resselt_ref.XXXXXX)"
cp -r "$SRC"/. "$TMP"/
python - "$TMP/pyproject.toml" <<'EOF'
import sys
p = sys.argv[1]
s = open(p).read()
s = s.replace('requires = ["hatchling"]', 'requires = ["setuptools"]').replace('build-backend = "hatchling.build"', 'build-backend = "setuptools.build_meta"')
s += '\n[tool.setuptools.packages.find]\ninclude = ["resselt*"]\n'
open(p, 'w').write(s)
EOF
rm -rf "$REPO/baseline/_ref"
mkdir -p "$REPO/baseline"
(cd "$TMP" && python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$REPO/baseline/_ref" "$TMP")
diff -rq -x __pycache__ "$SRC/resselt" "$REPO/baseline/_ref/resselt" && echo "baseline/_ref/resselt is identical to $SRC/resselt"
rm -rf "$TMP"
