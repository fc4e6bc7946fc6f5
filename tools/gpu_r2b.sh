#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_fused.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fused.log
tail -5 gpurun_out/pytest_fused.log
timeout 300 python tools/time_phases.py > gpurun_out/phases.txt 2>&1; tail -2 gpurun_out/phases.txt
