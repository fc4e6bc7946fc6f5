#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -x -q > gpurun_out/pytest_fin.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fin.log
tail -6 gpurun_out/pytest_fin.log
timeout 300 python tools/time_phases.py > gpurun_out/phases.txt 2>&1; tail -2 gpurun_out/phases.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py -x -q -k "consecutive or small or mixed" > gpurun_out/san_memcheck_r2.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_memcheck_r2.log
