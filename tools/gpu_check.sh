#!/bin/bash
# GPU box helper: parity subset + stage times + per-kernel times of the default library.  tools/gpu_check.sh <tag> [pytest -k expr]
tag=$1; kexpr=${2:-"patches or detections or large or knobs or golden or f32 or batch_equals or chunked or edge"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/${tag}_pytest.txt 2>&1
tail -15 gpurun_out/${tag}_pytest.txt
python tests/parity_stats.py > gpurun_out/${tag}_parity.txt 2>&1; cat gpurun_out/${tag}_parity.txt
python tools/exp.py "" --batch=256 > gpurun_out/${tag}_exp.txt 2>&1; cat gpurun_out/${tag}_exp.txt
tools/bin_times.sh $tag default > /dev/null 2>&1; cat gpurun_out/${tag}_bins.txt
