#!/bin/bash
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --mode train --steps 3 --warmup 3 > gpurun_out/r2n${N}_train.log 2> gpurun_out/r2n${N}_train.err
grep '^{' gpurun_out/r2n${N}_train.log | cut -c1-400; tail -2 gpurun_out/r2n${N}_train.err | cut -c1-200
