#!/bin/bash
# ncu --set full capture of one steady-state k_render_pass launch -> gpurun_out/<tag>_pass.ncu-rep (+ raw/source csv)
tag=$1; shift
timeout 800 ncu --set full --import-source on --clock-control none -k regex:k_render_pass -s 3 -c 1 -f -o gpurun_out/${tag}_pass python bench.py --steps 1 --warmup 1 --no-cpu --spp 64 "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_pass.ncu-rep --page raw --csv > gpurun_out/${tag}_pass_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_pass.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_pass_src.csv 2>/dev/null
ls -la gpurun_out/${tag}_pass*
