#!/bin/bash
# Round-end evidence run (one B200): kernel timings, ncu launch list of the bench command, ncu --set full of the dominant kernels.
# Everything lands in gpurun_out/; tools/summarize_profiles.py turns it into profiles/r01_*.md.
set -x
O=gpurun_out
mkdir -p $O
rm -f $O/conv3d_layer_timings_b6.txt
python tools/bench_kernels.py > $O/kernel_timings.txt 2>&1
for c in 64,32,48,256,128,0 32,32,48,256,128,0 32,64,48,256,128,1 64,64,24,128,64,0 64,64,24,128,64,1 64,64,12,64,32,0 64,64,12,64,32,2 64,32,24,128,64,2; do
  NORES=1 CFG=$c BATCH=6 python tools/deconv_one.py >> $O/conv3d_layer_timings_b6.txt 2>&1
done
python tools/sphere_one.py >> $O/conv3d_layer_timings_b6.txt 2>&1
python tools/cls_one.py >> $O/conv3d_layer_timings_b6.txt 2>&1
python tools/costvol_one.py >> $O/conv3d_layer_timings_b6.txt 2>&1
# launch list of the bench command (eager launches so that every kernel is a separate ncu record)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph > $O/bench_under_ncu.log 2>&1
# full captures: dominant conv3d kernel (32->32 stride 1, B=6), transposed conv, sphere conv
ITERS=1 NORES=1 CFG=32,32,48,256,128,0 ncu --set full --import-source on --clock-control none -k regex:conv3d_tc -s 3 -c 1 -o $O/conv3d_s1_b6 -f python tools/deconv_one.py > $O/ncu1.log 2>&1
ITERS=1 ncu --set full --import-source on --clock-control none -k regex:conv3d_tc -s 3 -c 1 -o $O/deconv_b6 -f python tools/deconv_one.py > $O/ncu2.log 2>&1
ITERS=1 ncu --set full --import-source on --clock-control none -k regex:sphere_conv_tc -s 3 -c 1 -o $O/sphere_b12 -f python tools/sphere_one.py > $O/ncu3.log 2>&1
ncu --set full --clock-control none -k regex:stem_conv_tc -c 1 -o $O/stem -f python bench.py --steps 1 --warmup 1 --no-graph > $O/ncu4.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err
tail -1 $O/bench_final.json
