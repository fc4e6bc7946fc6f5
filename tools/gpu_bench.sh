#!/bin/bash
# the two arms of the bench as the driver runs them (N = 1)
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "ours rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity']['dx_rel_err'], d['parity']['ok'])
print('roofline', d.get('roofline', {}).get('frac'), d.get('roofline_hbm', {}).get('frac'), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('phase', d.get('phase_ms')); print('api', d.get('e2e_problem_solve')); print('series', d['series']); print('cpu', d['cpu_baseline']['value'])
PY
if [ "$1" = "ref" ]; then
  BSLAM_REF_BUDGET_S=${2:-60} timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_ref.err; head -c 1500 gpurun_out/bench_ref.json
fi
