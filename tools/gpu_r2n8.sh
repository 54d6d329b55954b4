#!/bin/bash
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2n${N}_bench.log 2> gpurun_out/r2n${N}_bench.err
timeout 900 $TR bench.py --gpus $N --mode train --steps 3 --warmup 3 > gpurun_out/r2n${N}_train.log 2> gpurun_out/r2n${N}_train.err
timeout 600 $TR bench.py --gpus $N --mode twostage --steps 5 --warmup 3 > gpurun_out/r2n${N}_two.log 2> gpurun_out/r2n${N}_two.err
for f in bench train two; do grep '^{' gpurun_out/r2n${N}_$f.log | cut -c1-330; tail -2 gpurun_out/r2n${N}_$f.err | cut -c1-200; done
