#!/bin/bash
# Round-2 FINAL evidence run (one B200): kernel timings, ncu launch list of the bench command, ncu --set full of every kernel class
# (inference and training), the reference on the same GPU, the training-step profile, the final bench line.
# Everything lands in gpurun_out/; `python tools/summarize_profiles.py r02` turns it into profiles/r02_*.md.
set -x
O=gpurun_out
mkdir -p $O
export FP16=1
rm -f $O/r02_layer_timings_b6.txt $O/*.ncu-rep $O/*.raw.csv
MODE_B200_BENCH_PRECISION=fp16 python tools/bench_kernels.py > $O/r02_kernel_timings.txt 2>&1
for c in 64,32,48,256,128,0 32,32,48,256,128,0 32,64,48,256,128,1 64,64,24,128,64,0 64,64,24,128,64,1 64,64,12,64,32,0 64,64,12,64,32,2 64,32,24,128,64,2; do
  NORES=1 CFG=$c BATCH=6 python tools/deconv_one.py >> $O/r02_layer_timings_b6.txt 2>&1
done
python tools/sphere_one.py >> $O/r02_layer_timings_b6.txt 2>&1
MODE_B200_SPHERE_SLAB=0 python tools/sphere_one.py 2>&1 | sed 's/^/direct-gather kernel only (MODE_B200_SPHERE_SLAB=0): /' >> $O/r02_layer_timings_b6.txt
python tools/cls_one.py >> $O/r02_layer_timings_b6.txt 2>&1
python tools/costvol_one.py >> $O/r02_layer_timings_b6.txt 2>&1
python tools/sphere_diag.py 2>&1 | grep "us x\|hdr" > $O/r02_sphere_split.txt
python tools/train_profile.py > $O/r02_train_profile.txt 2>&1
python tools/ref_gpu_bench.py D > $O/r02_ref_train.txt 2>&1
build/mufu_bench > $O/r02_mufu_bench.txt 2>&1
# launch list of the bench command (eager launches so that every kernel is a separate ncu record)
MODE_B200_BENCH_LIGHT=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph > $O/r02_bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none -f"
# gpurun brings back at most 64 MiB: every report is exported to its raw-metric CSV (what tools/summarize_profiles.py reads) and deleted
cap() { name=$1; shift; "$@" > $O/$name.log 2>&1; ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null; rm -f $O/$name.ncu-rep $O/$name.log; }
cap r02_sphere_slab env ITERS=1 $NCU -k regex:sphere_conv_slab -s 3 -c 1 -o $O/r02_sphere_slab python tools/sphere_one.py
cap r02_sphere_direct env ITERS=1 $NCU -k regex:sphere_conv_tc_kernel -s 3 -c 1 -o $O/r02_sphere_direct python tools/sphere_one.py
cap r02_conv3d_s1 env ITERS=1 NORES=1 CFG=32,32,48,256,128,0 $NCU -k regex:conv3d_tc -s 3 -c 1 -o $O/r02_conv3d_s1 python tools/deconv_one.py
cap r02_deconv env ITERS=1 $NCU -k regex:conv3d_tc -s 3 -c 1 -o $O/r02_deconv python tools/deconv_one.py
cap r02_conv3d_s2small env ITERS=1 NORES=1 CFG=64,64,24,128,64,1 $NCU -k regex:conv3d_tc -s 3 -c 1 -o $O/r02_conv3d_s2small python tools/deconv_one.py
cap r02_cls env $NCU -k regex:conv3d_cls_tc -s 3 -c 1 -o $O/r02_cls python tools/cls_one.py
cap r02_costvol env ITERS=1 $NCU -k regex:costvol_conv_kernel -s 3 -c 1 -o $O/r02_costvol python tools/costvol_one.py
cap r02_regress env $NCU -k regex:disp_regress -s 2 -c 1 -o $O/r02_regress python tools/bench_kernels.py regress
cap r02_cost_volume env $NCU -k regex:cost_volume_bf16 -s 2 -c 1 -o $O/r02_cost_volume python tools/bench_kernels.py cost
cap r02_stem env MODE_B200_BENCH_LIGHT=1 $NCU -k regex:stem_conv_tc -c 1 -o $O/r02_stem python bench.py --steps 1 --warmup 1 --no-graph
cap r02_warp env $NCU -k regex:warp_scatter -s 2 -c 1 -o $O/r02_warp python tools/geometry_one.py
# training kernels (512x256, D=96, 2 pairs)
cap r02_sphere_f32 env ITERS=1 $NCU -k regex:sphere_conv_f32_tiled -s 20 -c 1 -o $O/r02_sphere_f32 python tools/train_one.py
cap r02_sphere_dgrad env ITERS=1 $NCU -k regex:sphere_dgrad_f32_tiled -s 4 -c 1 -o $O/r02_sphere_dgrad python tools/train_one.py
cap r02_sphere_wgrad env ITERS=1 $NCU -k regex:sphere_wgrad_f32_tiled -s 4 -c 1 -o $O/r02_sphere_wgrad python tools/train_one.py
cap r02_regress_bwd env ITERS=1 $NCU -k regex:disp_regress_bwd -c 1 -o $O/r02_regress_bwd python tools/train_one.py
cap r02_cost_volume_bwd env ITERS=1 $NCU -k regex:cost_volume_bwd -c 1 -o $O/r02_cost_volume_bwd python tools/train_one.py
cap r02_bn_cl_reduce env ITERS=1 $NCU -k regex:bn_cl_reduce_kernel -s 1 -c 1 -o $O/r02_bn_cl_reduce python tools/train_one.py
cap r02_bn_cl_apply env ITERS=1 $NCU -k regex:bn_cl_apply_kernel -s 1 -c 1 -o $O/r02_bn_cl_apply python tools/train_one.py
cap r02_bn_bwd_apply env ITERS=1 $NCU -k regex:bn_bwd_apply_kernel -s 1 -c 1 -o $O/r02_bn_bwd_apply python tools/train_one.py
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > $O/r02_clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 > $O/r02_bench_final.json 2> $O/r02_bench_final.err
kill $SMI
python bench.py --mode train --steps 3 > $O/r02_bench_train.json 2> $O/r02_bench_train.err
python bench.py --mode twostage > $O/r02_bench_twostage.json 2> $O/r02_bench_twostage.err
python bench.py --mode highres > $O/r02_bench_highres.json 2> $O/r02_bench_highres.err
tail -c 400 $O/r02_bench_final.json; tail -c 300 $O/r02_bench_train.json; ls $O/*.raw.csv | wc -l; du -sh $O
