#!/usr/bin/env bash
# Installs the UNMODIFIED reference (Niccolo-Ajroldi/plainLM) into baseline/_ref with the one offline pip install the
# build contract allows.  baseline/_ref is git-ignored (no reference source enters the history) but travels to the GPU
# box with the gpurun snapshot, where `bench.py --impl reference` and the `gpu_eager_reference` block import it.
#
# The only thing touched is packaging metadata in a scratch copy: the reference's pyproject lists
# packages = ["data", "optim", "engine", "models"] and forgets the sub-package data.datasets, which engine/engine.py:10
# imports (intra_doc_causal_mask) — an installed copy would not even import.  Adding the sub-package to the list is a
# one-word change to pyproject.toml in /tmp; every installed .py file is byte-identical to /root/reference.
set -euo pipefail
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -d "$REF" ] || { echo "install_ref: $REF not found (GPU box?): keeping the existing baseline/_ref"; exit 0; }
TMP=$(mktemp -d /tmp/plainlm_ref.XXXXXX)
cp -r "$REF"/. "$TMP"/
sed -i 's/packages = \["data", /packages = ["data", "data.datasets", /' "$TMP/pyproject.toml"
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$HERE/_ref" "$TMP"
rm -rf "$TMP"
# the driver scripts of the reference (not part of its wheel): train.py and the modules it imports from its own root.
# tests/test_gpu_train_dropin.py runs this unmodified train.py twice — on the drop-in shims and on the reference itself.
for f in train.py utils.py checkpoint_utils.py torch_utils.py; do
  cp "$REF/$f" "$HERE/_ref/$f"
  cmp -s "$HERE/_ref/$f" "$REF/$f" || { echo "MISMATCH $f"; exit 1; }
done
# byte-identity check of every installed source file against the reference tree
( cd "$HERE/_ref" && find data engine models optim -name '*.py' | while read -r f; do cmp -s "$f" "$REF/$f" || { echo "MISMATCH $f"; exit 1; }; done )
find "$HERE/_ref" -name '__pycache__' -type d -prune -exec rm -rf {} +
echo "install_ref: ok -> $HERE/_ref"
