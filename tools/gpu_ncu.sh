#!/bin/bash
# Nsight Compute captures of a round (one GPU; a number printed under ncu is never a bench value):
#   launch lists of one CFG forward and one guidance iteration, and --set full captures of the kernel families.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg_forward.csv python tools/ncu_target.py 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_guidance_iteration.csv python tools/ncu_target.py 1 guidance > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"temporal_attn_mma|cross_attn_mma|gn_stream|layernorm_kernel" -s 420 -c 14 -o gpurun_out/prof_small -f python tools/ncu_target.py 1 > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 470 -c 12 -o gpurun_out/prof_gemm -f python tools/ncu_target.py 1 > gpurun_out/ncu_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 17 -c 3 -o gpurun_out/prof_attn -f python tools/ncu_target.py 1 > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
