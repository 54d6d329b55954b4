#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_slab.py 2>&1 | grep -v Warn | grep "====\|tap\|hdr" > gpurun_out/r2c_debug.log
timeout 900 python -m pytest tests/test_gpu_reference.py -m gpu -q -s -x 2>&1 | grep -v "Warning\|upsample\|meshgrid\|^$" | tail -60 > gpurun_out/r2c_pytest_ref.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_reference.py 2>&1 | tail -60 > gpurun_out/r2c_pytest.log
FP16=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_conv -s 6 -c 2 -o gpurun_out/r2c_sphere python tools/sphere_one.py > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_pytest.log; tail -5 gpurun_out/r2c_pytest_ref.log; cat gpurun_out/r2c_debug.log | tail -30
