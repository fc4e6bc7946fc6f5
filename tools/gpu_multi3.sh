#!/bin/bash
# 2 GPUs: fused + multi-GPU tests, then the scaling probe with and without the NVLS multicast transport
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -12 gpurun_out/pytest_multi.log
bash tools/gpu_scale.sh "2 1"
cp gpurun_out/scale_quick.txt gpurun_out/scale_nvls.txt
BSLAM_NVLS=0 bash tools/gpu_scale.sh "2"
