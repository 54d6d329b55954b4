#!/bin/bash
timeout 600 python tools/debug_slab.py 2>&1 | grep -v Warn | grep "====\|tap\|bad\|first" > gpurun_out/r2k_debug.log
FP16=1 timeout 300 python tools/sphere_one.py > gpurun_out/r2k_sphere.log 2>&1
FP16=1 MODE_B200_SPHERE_STRIPS=0 timeout 300 python tools/sphere_one.py >> gpurun_out/r2k_sphere.log 2>&1
cat gpurun_out/r2k_debug.log | head -60; cat gpurun_out/r2k_sphere.log
