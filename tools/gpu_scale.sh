#!/bin/bash
# strong scaling probe on an N-GPU box: usage gpu_scale.sh "1 2 4 8" [track]
mkdir -p gpurun_out
: > gpurun_out/scale_quick.txt
for n in $1; do
  if [ "$n" = "1" ]; then timeout 300 python tools/scale_quick.py ${2:-6} >> gpurun_out/scale_quick.txt 2> gpurun_out/scale_quick_$n.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) tools/scale_quick.py ${2:-6} >> gpurun_out/scale_quick.txt 2> gpurun_out/scale_quick_$n.err; fi
  tail -2 gpurun_out/scale_quick_$n.err | cut -c1-300
done
cat gpurun_out/scale_quick.txt
