#!/bin/bash
# usage: gpu_multi.sh N  -- NCCL test + bench at 1..N GPUs (run under `gpurun --gpus N`)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for n in 1 2 4 8; do
  [ $n -gt $1 ] && break
  if [ $n == 1 ]; then python bench.py --steps 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$n.json").read().strip().splitlines()[-1]); print($n, 'it/s', d["value"], 'e2e', d["e2e"]["value"])
except Exception as e:
    print($n, 'FAILED', e)
PY
done
