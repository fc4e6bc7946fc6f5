#!/bin/bash
# usage: gpu_prof.sh <kernel-regex> [skip] [count]   -- tests + bench + ncu full capture of the kernels matching the regex
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'reproj_ms', d['roofline']['avg_launch_ms'])
print(d['phase_ms'])
PY
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name "regex:$1" --launch-skip ${2:-8} --launch-count ${3:-4} -f -o gpurun_out/full python tools/iter_loop.py 4 > gpurun_out/full.log 2>&1
tail -2 gpurun_out/full.log
