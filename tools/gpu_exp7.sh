#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x -k "batch_norm" 2>&1 | tail -15
timeout 600 python tools/train_profile.py 2>&1 | grep -v Warn | grep "step:\|total kernel\|sphere\|bn_\|lerp\| ms " | head -24
timeout 900 python -m pytest tests/test_gpu_model.py -q -x 2>&1 | tail -3
