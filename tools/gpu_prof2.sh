#!/bin/bash
O=gpurun_out; mkdir -p $O
ITERS=1 ncu --set full --import-source on --clock-control none -f -k regex:costvol_conv_kernel -s 3 -c 1 -o $O/r02b_costvol python tools/costvol_one.py > $O/r02b_ncu_costvol.log 2>&1
ls -la $O/r02b_costvol.ncu-rep
