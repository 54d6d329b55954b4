#!/bin/bash
# Round-2 evidence run (one B200): kernel timings, ncu launch list of the bench command, ncu --set full of every kernel class.
# Everything lands in gpurun_out/; tools/summarize_profiles.py r02 turns it into profiles/r02_*.md.
set -x
O=gpurun_out
mkdir -p $O
export FP16=1
rm -f $O/r02_layer_timings_b6.txt
MODE_B200_BENCH_PRECISION=fp16 python tools/bench_kernels.py > $O/r02_kernel_timings.txt 2>&1
for c in 64,32,48,256,128,0 32,32,48,256,128,0 32,64,48,256,128,1 64,64,24,128,64,0 64,64,24,128,64,1 64,64,12,64,32,0 64,64,12,64,32,2 64,32,24,128,64,2; do
  NORES=1 CFG=$c BATCH=6 python tools/deconv_one.py >> $O/r02_layer_timings_b6.txt 2>&1
done
python tools/sphere_one.py >> $O/r02_layer_timings_b6.txt 2>&1
MODE_B200_SPHERE_SLAB=0 python tools/sphere_one.py 2>&1 | sed 's/^/direct-gather kernel only (MODE_B200_SPHERE_SLAB=0): /' >> $O/r02_layer_timings_b6.txt
python tools/cls_one.py >> $O/r02_layer_timings_b6.txt 2>&1
python tools/costvol_one.py >> $O/r02_layer_timings_b6.txt 2>&1
# launch list of the bench command (eager launches so that every kernel is a separate ncu record)
MODE_B200_BENCH_LIGHT=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph > $O/r02_bench_under_ncu.log 2>&1
NCU="ncu --set full --import-source on --clock-control none -f"
ITERS=1 $NCU -k regex:sphere_conv_slab -s 3 -c 1 -o $O/r02_sphere_slab python tools/sphere_one.py > $O/r02_ncu1.log 2>&1
ITERS=1 $NCU -k regex:sphere_conv_tc_kernel -s 3 -c 1 -o $O/r02_sphere_direct python tools/sphere_one.py > $O/r02_ncu2.log 2>&1
ITERS=1 NORES=1 CFG=32,32,48,256,128,0 $NCU -k regex:conv3d_tc -s 3 -c 1 -o $O/r02_conv3d_s1 python tools/deconv_one.py > $O/r02_ncu3.log 2>&1
ITERS=1 $NCU -k regex:conv3d_tc -s 3 -c 1 -o $O/r02_deconv python tools/deconv_one.py > $O/r02_ncu4.log 2>&1
ITERS=1 NORES=1 CFG=64,64,24,128,64,1 $NCU -k regex:conv3d_tc -s 3 -c 1 -o $O/r02_conv3d_s2small python tools/deconv_one.py > $O/r02_ncu5.log 2>&1
$NCU -k regex:conv3d_cls_tc -s 3 -c 1 -o $O/r02_cls python tools/cls_one.py > $O/r02_ncu6.log 2>&1
ITERS=1 $NCU -k regex:costvol_conv_kernel -s 3 -c 1 -o $O/r02_costvol python tools/costvol_one.py > $O/r02_ncu7.log 2>&1
$NCU -k regex:disp_regress_kernel -s 2 -c 1 -o $O/r02_regress python tools/bench_kernels.py regress > $O/r02_ncu8.log 2>&1
$NCU -k regex:cost_volume_bf16 -s 2 -c 1 -o $O/r02_cost_volume python tools/bench_kernels.py cost > $O/r02_ncu9.log 2>&1
MODE_B200_BENCH_LIGHT=1 $NCU -k regex:stem_conv_tc -c 1 -o $O/r02_stem python bench.py --steps 1 --warmup 1 --no-graph > $O/r02_ncu10.log 2>&1
$NCU -k regex:warp_scatter -s 2 -c 1 -o $O/r02_warp python tools/geometry_one.py > $O/r02_ncu11.log 2>&1
ITERS=1 $NCU -k regex:sphere_dgrad -c 1 -o $O/r02_sphere_dgrad python tools/train_one.py > $O/r02_ncu12.log 2>&1
ITERS=1 $NCU -k regex:sphere_wgrad -c 1 -o $O/r02_sphere_wgrad python tools/train_one.py > $O/r02_ncu13.log 2>&1
ITERS=1 $NCU -k regex:disp_regress_bwd -c 1 -o $O/r02_regress_bwd python tools/train_one.py > $O/r02_ncu14.log 2>&1
ITERS=1 $NCU -k regex:cost_volume_bwd -c 1 -o $O/r02_cost_volume_bwd python tools/train_one.py > $O/r02_ncu15.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > $O/r02_clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 > $O/r02_bench_final.json 2> $O/r02_bench_final.err
kill $SMI
tail -c 600 $O/r02_bench_final.json
