#!/bin/bash
# ad-hoc: register vs shared-memory stencil cache x min blocks
cd "$(dirname "$0")/.."
for cfg in "0 3" "1 3" "1 4" "1 5" "1 6"; do
  set -- $cfg
  touch mitsubaer_b200/csrc/*.cu
  make -s -j4 -C mitsubaer_b200/csrc EXTRA="-DMER_STENCIL_SMEM=$1 -DMER_RENDER_MIN_BLOCKS=$2" 2>&1 | grep -E "error"
  echo "smem=$1 minblocks=$2: $(grep -h 'Used' mitsubaer_b200/csrc/build/mer_render.ptxas.log | head -2 | tail -1)"
  timeout 120 python tests/gpu_quick2.py 64 2>&1 | grep "steps/pass"
  timeout 200 python tools/sweep_c4.py --modes tricubic --fractions 1e-2,1e-3,1e-4 --rays 4194304 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('   c4', d['h_over_extent'], '%.2f Gsteps/s' % (d['ray_steps_per_sec']/1e9))"
done
touch mitsubaer_b200/csrc/*.cu
make -s -j4 -C mitsubaer_b200/csrc
