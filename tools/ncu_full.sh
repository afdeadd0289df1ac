#!/bin/bash
# GPU box helper: one `ncu --set full` capture of kernels matching a regex during a 32-image bench step.
#   tools/ncu_full.sh <tag> <kernel regex> [count] [skip]
tag=$1; pat=$2; cnt=${3:-2}; skip=${4:-0}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$pat" -s $skip -c $cnt -f -o gpurun_out/${tag} \
   python bench.py --steps 1 --warmup 0 --batch 32 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
