#!/bin/bash
# GPU box helper: per-kernel times (ncu, duration only) of the per-keypoint kernels for several library variants.
#   tools/bin_times.sh <tag> <variant|default> ...      -> gpurun_out/<tag>_bins.txt
tag=$1; shift
out=gpurun_out/${tag}_bins.txt
: > $out
for v in "$@"; do
  lib=hesaff_b200/variants/$v.so
  [ "$v" = default ] && lib=hesaff_b200/libhesaff_b200.so
  echo "=== $v" >> $out
  HESAFF_LIB=$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_describe|k_affine|k_large|k_sift' -c 24 --csv \
     --log-file gpurun_out/${tag}_$v.csv python bench.py --steps 1 --warmup 0 --batch 32 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py launches gpurun_out/${tag}_$v.csv >> $out 2>&1
done
cat $out
