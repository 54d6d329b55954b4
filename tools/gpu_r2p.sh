#!/bin/bash
# re-entry check of HEAD: GPU tests, stand-alone sphere layer timing, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2p_tests.log
FP16=1 timeout 300 python tools/sphere_one.py > gpurun_out/r2p_sphere.log 2>&1
timeout 900 python bench.py > gpurun_out/r2p_bench.log 2>&1
tail -5 gpurun_out/r2p_tests.log; cat gpurun_out/r2p_sphere.log; tail -c 3000 gpurun_out/r2p_bench.log
