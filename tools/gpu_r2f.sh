#!/bin/bash
# round 2: Cholesky changes -- parity tests that exercise the reduced solve, trace, phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -x -q > gpurun_out/pytest_chol.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_chol.log
tail -6 gpurun_out/pytest_chol.log
timeout 300 python tools/chol_trace.py > gpurun_out/chol_trace_r2b.txt 2>&1; head -6 gpurun_out/chol_trace_r2b.txt
timeout 300 python tools/time_phases.py > gpurun_out/phases.txt 2>&1; tail -2 gpurun_out/phases.txt
