#!/bin/bash
# launch list + one ncu --set full capture of a kernel.  Usage: tools/gpu_job3.sh <tag> <kernel-regex> <skip> [bench args...]
tag=$1; kre=$2; skip=$3; shift 3
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu --spp ${NCU_SPP:-64} "$@" > gpurun_out/${tag}_launches.log 2>&1
timeout 800 ncu --set full --import-source on --clock-control none -k regex:$kre -s $skip -c 1 -f -o gpurun_out/${tag}_k python bench.py --steps 1 --warmup 0 --no-cpu --spp ${NCU_SPP:-64} "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_k.ncu-rep --page raw --csv > gpurun_out/${tag}_k_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_k.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_k_src.csv 2>/dev/null
rm -f gpurun_out/${tag}_k.ncu-rep
ls -la gpurun_out/${tag}_*
