#!/bin/bash
# 2-GPU session: device-guard test, default bench, training step, two-stage frame sharding
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -k "device_not_the_current" 2>&1 | tail -3 > gpurun_out/r2j_pytest.log
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j_bench2.log 2> gpurun_out/r2j_bench2.err
timeout 900 $TR bench.py --gpus 2 --mode train --steps 3 --warmup 3 > gpurun_out/r2j_train2.log 2> gpurun_out/r2j_train2.err
timeout 600 $TR bench.py --gpus 2 --mode twostage --steps 5 --warmup 3 > gpurun_out/r2j_two2.log 2> gpurun_out/r2j_two2.err
cat gpurun_out/r2j_pytest.log; head -c 400 gpurun_out/r2j_bench2.log; echo; tail -c 900 gpurun_out/r2j_train2.log; tail -2 gpurun_out/r2j_train2.err; tail -c 700 gpurun_out/r2j_two2.log; tail -2 gpurun_out/r2j_two2.err
