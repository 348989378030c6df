#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with gpurun):
#     baseline/install_reference.sh [/root/reference]
# The reference ships no packaging metadata (no setup.py / pyproject.toml, no __init__.py files), so `pip install
# /root/reference` cannot work as is.  Following the base contract, the tree is copied to /tmp, a setup.py that ONLY lists
# its directories as namespace packages is added to the copy (no source file of the reference is touched), and pip installs
# that copy with --no-deps (its requirements pin yfinance / matplotlib, which have no wheel here; the hot path needs
# numpy, scipy, numba, joblib, pandas, all present in the image).  Two empty stub modules for the absent yfinance and
# matplotlib (imported at module top by the reference, never used on the hot path) go to baseline/_ref/_stubs.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DEST="$HERE/_ref"
[ -d "$SRC" ] || { echo "reference checkout not found at $SRC" >&2; exit 2; }
TMP="$(mktemp -d /tmp/reference_pkg.XXXXXX)"
cp -r "$SRC"/. "$TMP"/
cat > "$TMP/setup.py" <<'PY'
from setuptools import setup, find_namespace_packages
setup(name="copula-msm-and-copula-garch-var-reference", version="0", py_modules=["main"],
      packages=find_namespace_packages(include=["utils*", "copulas*", "garch*", "kalman_mean_reverting*",
                                                 "markov_switching_multifractal*", "data_loader*"]))
PY
rm -rf "$DEST"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DEST" "$TMP" 2>&1 | tail -3
mkdir -p "$DEST/_stubs/matplotlib"
: > "$DEST/_stubs/yfinance.py"
: > "$DEST/_stubs/matplotlib/__init__.py"
: > "$DEST/_stubs/matplotlib/pyplot.py"
rm -rf "$TMP"
echo "installed: $(find "$DEST" -name '*.py' | wc -l) python files under $DEST"
