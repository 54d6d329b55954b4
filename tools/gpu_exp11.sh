#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "costvol" 2>&1 | tail -3
timeout 300 python tools/costvol_one.py 2>&1 | grep -v Warn | tail -2
