#!/bin/bash
# bench.py as the driver's scaling run launches it: usage gpu_scale_bench.sh "8 4"
mkdir -p gpurun_out
for n in $1; do
  if [ "$n" = "1" ]; then timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; fi
  echo "N=$n rc=$?"; tail -2 gpurun_out/bench_n$n.err | cut -c1-200
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print($n, 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity']['dx_rel_err'], d['parity']['ok'], 'launches', d['gpu_launches'])
    print('   ', d['config']['parallelism'][:150]); print('   ', d['series'])
except Exception as e:
    print($n, 'failed', e)
PY
done
