#!/bin/bash
timeout 600 python tools/debug_slab.py 2>&1 | grep -v Warn | grep "====\|tap\|hdr" > gpurun_out/r2e_debug.log
FP16=1 timeout 300 python tools/sphere_one.py > gpurun_out/r2e_sphere.log 2>&1
timeout 300 python tools/sphere_one.py >> gpurun_out/r2e_sphere.log 2>&1
cat gpurun_out/r2e_debug.log gpurun_out/r2e_sphere.log
