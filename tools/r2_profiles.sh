#!/bin/bash
# GPU box helper: the round's committed evidence -> gpurun_out/r2_* (copied to profiles/ afterwards).  tools/r2_profiles.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
# launch list of one bench step (chunks of 32 x 1080p), duration only
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 260 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 1 --warmup 0 --batch 32 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt 2>&1
# full captures: first blur (octave 0, 11 taps) + the next three blur levels; the per-keypoint kernels
ncu --set full --clock-control none --import-source on -k regex:'k_blur_tma' -c 5 -f -o gpurun_out/${tag}_blur \
   python bench.py --steps 1 --warmup 0 --batch 32 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py raw gpurun_out/${tag}_blur.ncu-rep > gpurun_out/${tag}_k_blur_tma_ncu.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_describe|k_affine' -c 8 -f -o gpurun_out/${tag}_kp \
   python bench.py --steps 1 --warmup 0 --batch 32 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none -k regex:'k_nms' -c 1 -f -o gpurun_out/${tag}_nms \
   python bench.py --steps 1 --warmup 0 --batch 32 --no-cpu-baseline > /dev/null 2>&1
(python tools/ncu_summary.py raw gpurun_out/${tag}_nms.ncu-rep; python tools/ncu_summary.py raw gpurun_out/${tag}_kp.ncu-rep) > gpurun_out/${tag}_k_describe_affine_ncu.txt 2>&1
# per source line: shared-memory wavefronts and executed instructions of the per-keypoint kernels
ncu -i gpurun_out/${tag}_kp.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${tag}_kp_src.csv 2>/dev/null
(python tools/ncu_smem.py gpurun_out/${tag}_kp_src.csv 14; python tools/ncu_lines.py gpurun_out/${tag}_kp_src.csv 14) > gpurun_out/${tag}_k_describe_affine_lines.txt 2>&1
rm -f gpurun_out/${tag}_kp_src.csv
tail -5 gpurun_out/${tag}_launches.txt
