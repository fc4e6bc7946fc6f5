#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list, ncu full capture of the four top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/iter_loop.py 6 > gpurun_out/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name 'regex:reproj_block|schur_block|lm_finish|chol_solve' --launch-skip 8 --launch-count 4 -f -o gpurun_out/full python tools/iter_loop.py 4 > gpurun_out/full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | head -c 3000
