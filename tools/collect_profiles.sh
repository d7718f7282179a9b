#!/bin/bash
# Regenerates the round-2 evidence under gpurun_out/ (copied to profiles/r02_* afterwards).  One B200, a few minutes.
# usage (from the repo root, on the GPU box): bash tools/collect_profiles.sh
O=gpurun_out
T="timeout 300"
$T python tools/gpu_train_stages.py 4 32            > $O/p_train_stage_timeline.log 2>&1
$T python tools/gpu_train_time.py 4 20              > $O/p_train_time.log 2>&1
$T python tools/gpu_train_time.py 16 10 16         >> $O/p_train_time.log 2>&1
FVT_WGRAD_GROUP=0 $T python tools/gpu_train_time.py 4 20 >> $O/p_train_time.log 2>&1
$T python tools/gpu_wgrad_group.py 20               > $O/p_wgrad_group.log 2>&1
$T python tools/gpu_wgrad_shapes.py 20              > $O/p_wgrad_shapes.log 2>&1
$T python tools/gpu_dgrad_fused.py 20               > $O/p_dgrad_fused_ab.log 2>&1
$T python tools/gpu_small_kernel_floor.py           > $O/p_small_kernel_floor.log 2>&1
# ncu: launch list of two training steps, then full captures of the kernels the review asked for (final build)
$T ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/p_train_step_launches.csv python tools/gpu_train_time.py 4 2 > $O/p_ncu_a.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_group -o $O/p_wgrad_group -f python tools/gpu_wgrad_group.py 1 > $O/p_ncu_b.log 2>&1
FVT_ONLY=conv2 $T ncu --set full --clock-control none --import-source on -k "regex:conv_|bn_bwd" -o $O/p_dgrad_bn_conv2 -f python tools/gpu_dgrad_fused.py 1 > $O/p_ncu_c.log 2>&1
FVT_ONLY="conv2_x" $T ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -o $O/p_wgrad_conv2_single -f python tools/gpu_wgrad_shapes.py 1 > $O/p_ncu_d.log 2>&1
echo collected
