#!/bin/bash
O=gpurun_out; mkdir -p $O
export FP16=1
NCU="ncu --set full --import-source on --clock-control none -f"
$NCU -k regex:conv3d_cls_tc -s 3 -c 1 -o $O/r02b_cls python tools/cls_one.py > $O/r02b_ncu_cls.log 2>&1
$NCU -k regex:disp_regress -s 2 -c 1 -o $O/r02b_regress python tools/bench_kernels.py regress > $O/r02b_ncu_regress.log 2>&1
ls -la $O/*.ncu-rep | tail -3
