#!/bin/bash
# bench variants without pytest + one ncu capture.  Usage: tools/gpu_job2.sh <tag> <ncu-env or -> [variants...]
tag=$1; shift
ncuenv=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=$(echo "$v" | tr ' =' '__' | tr -cd 'A-Za-z0-9_.-')
  ( eval "env $v timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu $BENCH_ARGS" ) > gpurun_out/${tag}_bench_${name}.json 2> gpurun_out/${tag}_bench_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${name}.json").read().strip().splitlines()[-1])
    print("$v: value %.2f M samples/s, %.2f G steps/s, e2e %s" % (d["value"]/1e6, d["ray_steps_per_sec"]/1e9, d.get("e2e") and round(d["e2e"]["value"]/1e6,2)))
except Exception as e:
    print("$v failed:", e); print(open("gpurun_out/${tag}_bench_${name}.err").read()[-800:])
PY
done
if [ "$ncuenv" != "-" ]; then
  env $ncuenv timeout 800 ncu --set full --import-source on --clock-control none -k regex:k_render_pass -s 3 -c 1 -f -o gpurun_out/${tag}_pass python bench.py --steps 1 --warmup 1 --no-cpu --spp 64 $BENCH_ARGS > gpurun_out/${tag}_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_pass.ncu-rep --page raw --csv > gpurun_out/${tag}_pass_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_pass.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_pass_src.csv 2>/dev/null
  rm -f gpurun_out/${tag}_pass.ncu-rep
fi
