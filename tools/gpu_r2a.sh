#!/bin/bash
# round 2: fused panel kernels -- tests first, then the whole GPU suite, then the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_fused.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fused.log
tail -25 gpurun_out/pytest_fused.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/time_phases.py > gpurun_out/phases.txt 2>&1; tail -5 gpurun_out/phases.txt
BSLAM_FUSED=0 timeout 300 python tools/time_phases.py > gpurun_out/phases_unfused.txt 2>&1; tail -3 gpurun_out/phases_unfused.txt
