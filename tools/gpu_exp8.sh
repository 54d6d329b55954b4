#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "batch_norm" 2>&1 | tail -3
echo "--- no scatter variant"
MODE_B200_LIB=$PWD/build/variants/noscatter/libmode_b200.so timeout 600 python tools/train_profile.py 2>&1 | grep -v Warn | grep "step:\|sphere_dgrad\|sphere_wgrad"
