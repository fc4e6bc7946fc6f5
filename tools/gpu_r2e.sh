#!/bin/bash
# round 2: new feature tests (f2-f4, split photometric, reference suite), then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipelines.py tests/test_reference_suite.py -q > gpurun_out/pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_new.log
tail -40 gpurun_out/pytest_new.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pipelines.py --deselect tests/test_reference_suite.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
