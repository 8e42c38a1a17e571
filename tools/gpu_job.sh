#!/bin/bash
# One gpurun call: GPU tests, then bench variants.  Usage: tools/gpu_job.sh <tag> [variants...]
# Every output goes to gpurun_out/<tag>_*.
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
for v in "$@"; do
  name=$(echo "$v" | tr ' =' '__' | tr -cd 'A-Za-z0-9_.-')
  echo "== $v"
  ( eval "env $v timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu $BENCH_ARGS" ) > gpurun_out/${tag}_bench_${name}.json 2> gpurun_out/${tag}_bench_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${name}.json").read().strip().splitlines()[-1])
    print("   value %.2f M samples/s, %.2f G steps/s, e2e %s" % (d["value"]/1e6, d["ray_steps_per_sec"]/1e9, d.get("e2e") and round(d["e2e"]["value"]/1e6,2)))
except Exception as e:
    print("   failed:", e); print(open("gpurun_out/${tag}_bench_${name}.err").read()[-800:])
PY
done
