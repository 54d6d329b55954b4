#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "batch_norm" 2>&1 | tail -8
timeout 600 python tools/train_profile.py 2>&1 | grep -v Warn | grep "step:\|total kernel\| ms " | head -22
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_reference.py -q -x 2>&1 | tail -3
