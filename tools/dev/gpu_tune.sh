#!/bin/bash
# ad-hoc tuning sweep (run on the GPU box): launch-bounds variants x maxWait
cd "$(dirname "$0")/.."
for mb in 2 3 4; do
  touch mitsubaer_b200/csrc/mer_render.cu
  make -s -C mitsubaer_b200/csrc EXTRA=-DMER_RENDER_MIN_BLOCKS=$mb 2>&1 | grep -E "error"
  grep -h "Used" mitsubaer_b200/csrc/build/mer_render.ptxas.log | head -2 | tail -1
  for mw in 4 8 12 16; do
    echo -n "minblocks $mb maxwait $mw: "
    MER_MAX_WAIT=$mw timeout 120 python tests/gpu_quick2.py 64 2>&1 | grep "steps/pass  2048 pool       0"
  done
done
touch mitsubaer_b200/csrc/mer_render.cu
make -s -C mitsubaer_b200/csrc
