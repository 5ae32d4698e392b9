#!/bin/bash
# oracle/_ref — the UNMODIFIED reference operator modules, runnable on the GPU box (TEST / BENCH INFRASTRUCTURE).
#
# /root/reference does not exist on the GPU box and the reference package cannot be pip-installed offline (its
# __init__.py imports hydra / xarray / jax / pytorch_lightning), but the operator modules of the hot path are plain
# torch + einops files.  This recipe copies exactly those files, verbatim, into oracle/_ref/ (git-ignored: never part
# of the history; NOT gpurun-ignored: it travels with the snapshot) under stub parent packages, so that
#   * bench.py --impl reference   times the reference's own FNOFactorized2DBlock on the host cores, and
#   * bench.py's eager_gpu leg    times the same module .cuda() (the north star's ">= 10x reference eager" denominator).
# Nothing under fourierflow_b200/ may import it.  Re-run whenever /root/reference changes:  bash oracle/make_ref.sh
set -euo pipefail
REF=${FFNO_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
SRC="$REF/fourierflow/modules"
[ -d "$SRC" ] || { echo "make_ref: $SRC not found (run in the build container)"; exit 0; }
rm -rf "$OUT"
mkdir -p "$OUT/fourierflow/modules/factorized_fno" "$OUT/fourierflow/modules/factorized_cno" "$OUT/fourierflow/modules/zongyi_fno"
# the path itself, then the sibling operators of SURVEY §8 f-4 (tools/sibling_times.py times them next to the mirrors)
for f in feedforward.py linear.py normalizer.py loss.py factorized_fno/grid_2d.py factorized_fno/mesh_2d.py factorized_fno/mesh_3d.py \
         dct.py iphi.py factorized_cno/grid_2d.py factorized_cno/mesh_2d.py factorized_cno/mesh_3d.py factorized_fno/point_cloud_2d.py \
         zongyi_fno/grid_plus_2d.py; do
  cp "$SRC/$f" "$OUT/fourierflow/modules/$f"
done
# stub parents (generated, not copied): the reference's own __init__ files import the whole research stack
: > "$OUT/fourierflow/__init__.py"
: > "$OUT/fourierflow/modules/__init__.py"
: > "$OUT/fourierflow/modules/factorized_fno/__init__.py"
: > "$OUT/fourierflow/modules/factorized_cno/__init__.py"
: > "$OUT/fourierflow/modules/zongyi_fno/__init__.py"
( cd "$OUT" && sha256sum $(find fourierflow -name '*.py' | sort) > MANIFEST.sha256 )
echo "oracle/_ref: $(find "$OUT" -name "*.py" | wc -l) files copied verbatim from $SRC"
