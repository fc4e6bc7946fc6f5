#!/bin/bash
# round 2: the multi-GPU tests first (bounded), then the whole GPU suite, the fp64 peak, the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_peak tools/fp64_peak.cu && /tmp/fp64_peak > gpurun_out/fp64_peak.json; cat gpurun_out/fp64_peak.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; head -c 3000 gpurun_out/bench.json
