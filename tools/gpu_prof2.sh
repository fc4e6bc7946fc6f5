#!/bin/bash
# usage: gpu_prof2.sh <kernel-regex> [skip] [count] -- ncu full capture of kernels inside un-graphed bslam_iterate calls
mkdir -p gpurun_out
BSLAM_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name "regex:$1" --launch-skip ${2:-2} --launch-count ${3:-2} -f -o gpurun_out/full2 python tools/iter_loop_fused.py 3 > gpurun_out/full2.log 2>&1
tail -3 gpurun_out/full2.log
