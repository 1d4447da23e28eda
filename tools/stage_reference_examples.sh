#!/bin/sh
# Copies the reference's example scripts that tests/test_gpu_reference_examples.py runs UNMODIFIED on the GPU into
# an UNTRACKED scratch directory of the repo (git-ignored; it travels to the GPU box with gpurun, where
# /root/reference does not exist).  Nothing here is committed.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference/examples}"
DST="$ROOT/gpurun_scratch/reference_examples"
mkdir -p "$DST"
for f in bortels_threeion_nondim.py bortels_structuredquad_nondim.geo gupta_advection_migration.py BMCSL.py \
         catalyst_layer_Cu_full.py; do
    cp "$SRC/$f" "$DST/"
done
echo "staged into $DST"
