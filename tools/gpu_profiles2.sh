#!/bin/bash
# round-2 captures judged under profiles/: launch list of the iteration + full capture of its kernels (un-graphed iterations)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/iter_loop_fused.py 6 > gpurun_out/launches.log 2>&1
BSLAM_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name 'regex:fused_panel|panel_finish|chol_solve|prepare_kernel' --launch-skip 8 --launch-count 4 -f -o gpurun_out/full4 python tools/iter_loop_fused.py 4 > gpurun_out/full4.log 2>&1
tail -2 gpurun_out/full4.log
timeout 300 python tools/chol_trace.py --all > gpurun_out/chol_trace_r2.txt 2>&1; head -6 gpurun_out/chol_trace_r2.txt
