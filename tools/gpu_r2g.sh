#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -x -q -k "fused or config4 or config3 or mixed or panel or damping or small or consecutive or losses or l1" > gpurun_out/pytest_fin.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fin.log
tail -6 gpurun_out/pytest_fin.log
timeout 300 python tools/time_phases.py > gpurun_out/phases.txt 2>&1; tail -2 gpurun_out/phases.txt
