#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "sphere_conv_tc |ref-gpu|feature stage|EPE|passed|failed|Error|error|FAILED|assert" | grep -v "^  \|print" | head -80 > gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.log 2> gpurun_out/r2h_bench.err
cat gpurun_out/r2h_pytest.log; head -c 700 gpurun_out/r2h_bench.log
