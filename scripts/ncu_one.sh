#!/bin/bash
# GPU box: one `ncu --set full` capture of one kernel of the bench-size build.  Usage: bash scripts/ncu_one.sh <kernel regex> <name> [launch skip]
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$1 -s ${3:-0} -c 1 \
    -o gpurun_out/ncu_$2 -f python scripts/build_once.py bench > gpurun_out/ncu_$2.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_$2.ncu-rep > gpurun_out/ncu_$2.txt 2>&1
ncu -i gpurun_out/ncu_$2.ncu-rep --page source --csv > gpurun_out/ncu_$2_source.csv 2>/dev/null
