#!/bin/bash
# Stage the UNMODIFIED reference files of the sampling hot path under oracle/_ref/ (git-ignored, NOT gpurun-ignored), so
# that the reference's own SpacedDiffusion.p_sample_loop + CMDM can be timed on the GPU box's host cores (and on its GPU as
# the torch-eager "library" baseline) by `bench.py --impl reference`.  /root/reference does not exist on the GPU box.
#
#   bash oracle/make_ref.sh [reference_root]        (run by __graft_entry__.build() when the reference tree is present)
#
# The file list is the import closure of `model.cmdm`, `model.cfg_sampler`, `diffusion.respace`, `utils.model_util` and
# `utils.rotation_conversions` inside the reference tree (17 files, namespace packages -- the tree has no __init__.py);
# timm / clip / smplx stay stubbed by oracle/ref_shim.py.  Nothing here is read by the product path: only
# oracle/ref_shim.py (test infrastructure) imports from oracle/_ref, and only when /root/reference is absent.
set -e
SRC=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
FILES="
data_loaders/humanml/common/quaternion.py
data_loaders/humanml/common/skeleton.py
data_loaders/humanml/scripts/motion_process.py
data_loaders/humanml/utils/paramUtil.py
diffusion/gaussian_diffusion.py
diffusion/losses.py
diffusion/nn.py
diffusion/respace.py
model/cfg_sampler.py
model/cmdm.py
model/mlp.py
model/rotation2xyz.py
model/smpl.py
model/transformer_utils.py
utils/config.py
utils/model_util.py
utils/rotation_conversions.py
"
if [ ! -d "$SRC/model" ]; then
  echo "make_ref: no reference tree at $SRC (nothing staged)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
for f in $FILES; do
  mkdir -p "$DST/$(dirname "$f")"
  cp "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum $FILES ) > "$DST/SHA256SUMS"
echo "make_ref: staged $(echo $FILES | wc -w) reference files under $DST"
