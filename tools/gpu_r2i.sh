#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2i_pytest.log
timeout 900 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/r2i_train1.log 2> gpurun_out/r2i_train1.err
timeout 600 python bench.py --mode twostage --steps 5 --warmup 3 > gpurun_out/r2i_two1.log 2> gpurun_out/r2i_two1.err
timeout 600 python bench.py --mode highres --steps 5 --warmup 3 > gpurun_out/r2i_high1.log 2> gpurun_out/r2i_high1.err
tail -8 gpurun_out/r2i_pytest.log; tail -c 1500 gpurun_out/r2i_train1.log; tail -3 gpurun_out/r2i_train1.err; tail -c 1200 gpurun_out/r2i_two1.log; tail -3 gpurun_out/r2i_two1.err; head -c 600 gpurun_out/r2i_high1.log; tail -3 gpurun_out/r2i_high1.err
