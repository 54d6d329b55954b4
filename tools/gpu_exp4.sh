#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "disp_regress" 2>&1 | tail -3
timeout 300 python tools/bench_kernels.py regress 2>&1 | grep -v Warn | tail -1
timeout 900 python -m pytest tests/test_gpu_reference.py tests/test_gpu_model.py -q -x 2>&1 | tail -2
