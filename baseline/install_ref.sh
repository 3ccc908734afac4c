#!/bin/sh
# Installs the UNMODIFIED reference (rfeinman/pyBPL, /root/reference) into baseline/_ref (git-ignored; it travels to
# the GPU box with gpurun).  Build container only: the GPU box has no /root/reference and uses what is already there.
#   - the package: pip --target from a scratch copy (the checkout is read-only and setuptools writes build/ next to it)
#   - lib_data/  : the hyper-parameter tables pybpl/__init__.py:13-16 expects beside the package (data, not source)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PYBPL_REFERENCE:-/root/reference}"
DST="$HERE/_ref"
[ -d "$REF/pybpl" ] || { echo "install_ref: no reference checkout at $REF" >&2; exit 3; }
if [ -f "$DST/pybpl/__init__.py" ] && [ -d "$DST/lib_data" ] && [ "$1" != "--force" ]; then exit 0; fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"
rm -rf "$DST"
python -m pip install -q --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$DST" "$TMP/src"
cp -r "$REF/lib_data" "$DST/lib_data"
echo "installed reference into $DST"
