#!/bin/sh
# Stages the reference's own Python tree into oracle/_ref so that it travels to the GPU box.  TEST INFRASTRUCTURE.
#
#   oracle/make_ref.sh [/root/reference]
#
# The reference (shanjiayao/PTT) is pure Python: there is nothing to compile.  oracle/_ref is listed in .gitignore
# (no reference source ever enters this repository's history) but NOT in .gpurunignore, so a `gpurun` snapshot carries
# it like the built .so files.  What is staged: the `ptt` package (models, utils, datasets, config) and `tools/`
# (YAML configs, train / eval loops); docs/ (24 MB of images) is left behind.
# Users: tests marked `reference` (the reference's own nn.Modules on the B200 over the pointnet2_ops._ext drop-in, its
# own build_network with the B200 modules registered), bench.py --impl reference (the reference's own modules on the
# host cores) and bench.py's `reference_modules_gpu` key.  Product code never reads it.
set -eu
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/ptt/models" ]; then
  echo "make_ref: no reference tree at $SRC (nothing staged)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/ptt" "$DST/ptt"
cp -r "$SRC/tools" "$DST/tools"
for f in setup.py requirements.txt README.md; do
  [ -f "$SRC/$f" ] && cp "$SRC/$f" "$DST/$f"
done
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
chmod -R u+w "$DST"
echo "make_ref: staged $(find "$DST" -type f | wc -l) files from $SRC into $DST"
