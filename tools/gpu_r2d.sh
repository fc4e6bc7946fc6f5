#!/bin/bash
# round 2: in-process shards test, launch list of the bench command, full ncu capture of the iteration's kernels, Cholesky trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "local" > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/iter_loop_fused.py 6 > gpurun_out/launches.log 2>&1
BSLAM_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name 'regex:fused_panel|panel_finish|chol_solve|prepare_kernel|retract_se3' --launch-skip 10 --launch-count 5 -f -o gpurun_out/full3 python tools/iter_loop_fused.py 4 > gpurun_out/full3.log 2>&1
tail -2 gpurun_out/full3.log
timeout 300 python tools/chol_trace.py --all > gpurun_out/chol_trace_r2.txt 2>&1; head -6 gpurun_out/chol_trace_r2.txt
