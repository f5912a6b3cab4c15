#!/bin/bash
# Run on the GPU box (gpurun): launch list + one full ncu capture of every kernel of a window.
# usage: scripts/gpu_profile.sh <round-tag>     outputs under gpurun_out/
TAG=${1:-r01}
mkdir -p gpurun_out
# launch list: 3 iterations x 7 kernels; skip the first iteration (cold)
ncu --metrics gpu__time_duration.sum --clock-control none -s 7 -c 14 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_step.py 4096 20 3 > gpurun_out/launches_${TAG}.log 2>&1
# full capture of the second iteration's 7 kernels
ncu --set full --clock-control none --import-source on -s 7 -c 7 -f -o gpurun_out/prof_${TAG} \
    python scripts/profile_step.py 4096 20 2 > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out/
