#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the reference's own Cython binding (deepgroebner/wrapped.pyx -> CLeadMonomialsEnv,
# the class scripts/train.py constructs by default) UNMODIFIED, from a scratch copy of the reference sources (cythonize
# writes wrapped.cpp beside the .pyx and /root/reference is read-only), into oracle/_ref/deepgroebner_ref/.  Nothing is
# copied into the repository's history: oracle/_ref is git-ignored.  bench.py times it as the `dropin_n1` baseline.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -d "$REF/deepgroebner" ] || { echo "reference sources not present at $REF: keeping prebuilt oracle/_ref (if any)"; exit 0; }
TMP=$(mktemp -d)
cp -r "$REF/deepgroebner" "$REF/setup.py" "$REF/README.md" "$TMP/"
(cd "$TMP" && CFLAGS="-O2" python setup.py -q build_ext --inplace > "$TMP/build.log" 2>&1) || { tail -20 "$TMP/build.log"; exit 1; }
mkdir -p "$HERE/_ref/deepgroebner_ref"
cp "$TMP"/deepgroebner/wrapped*.so "$HERE/_ref/deepgroebner_ref/"
echo "# built from the unmodified reference by oracle/build_cython_ref.sh" > "$HERE/_ref/deepgroebner_ref/__init__.py"
rm -rf "$TMP"
ls "$HERE/_ref/deepgroebner_ref/"
