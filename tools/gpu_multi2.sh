#!/bin/bash
# 2 GPUs: multi-GPU tests (in-process shards, 2 processes / 1 GPU over IPC, NCCL peer + fallback), bench N=1 and N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
timeout 600 python bench.py --steps 50 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 50 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open('gpurun_out/bench_n%d.json' % n).read().strip().splitlines()[-1])
        print(n, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity']['dx_rel_err'], d['parity']['ok'], d['config']['parallelism'][:60], d['series'])
    except Exception as e:
        print(n, 'failed', e)
PY
