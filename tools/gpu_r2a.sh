#!/bin/bash
# round-2 GPU session A: full GPU test suite, the reference-on-GPU bar, bench, sphere kernel A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference.py 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
timeout 900 python -m pytest tests/test_gpu_reference.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2a_pytest_ref.log
FP16=1 timeout 300 python tools/sphere_one.py > gpurun_out/r2a_sphere_new.log 2>&1
FP16=1 MODE_B200_SPHERE_SLAB=0 timeout 300 python tools/sphere_one.py > gpurun_out/r2a_sphere_old.log 2>&1
timeout 900 python tools/ref_gpu_bench.py > gpurun_out/r2a_refbench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; tail -15 gpurun_out/r2a_pytest_ref.log; cat gpurun_out/r2a_sphere_new.log gpurun_out/r2a_sphere_old.log; tail -3 gpurun_out/r2a_bench.log | cut -c1-600
