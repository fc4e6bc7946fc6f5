#!/bin/bash
# Captures judged under profiles/: launch list of bench.py's command + full capture of the four top kernels.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/iter_loop.py 6 > gpurun_out/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name 'regex:reproj_block|schur_block|lm_finish|chol_solve' --launch-skip 8 --launch-count 4 -f -o gpurun_out/full python tools/iter_loop.py 4 > gpurun_out/full.log 2>&1
tail -n 1 gpurun_out/launches.log; tail -n 1 gpurun_out/full.log
