#!/bin/bash
# Nsight Compute captures of a round (one GPU; a number printed under ncu is never a bench value):
#   1. launch lists (duration + DRAM bytes of EVERY launch) of one CFG forward and one guidance iteration,
#   2. --set full captures of the kernel families (source-mapped: the library is built with -lineinfo).
# Summaries: python tools/ncu_summarize.py gpurun_out  ->  profiles/r2_*
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_cfg_forward.csv python tools/ncu_target.py 1 > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_guidance_iteration.csv python tools/ncu_target.py 1 guidance > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_self_kernel" -s 5 -c 3 -o gpurun_out/prof_attn_self -f python tools/ncu_target.py 1 > gpurun_out/ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 520 -c 10 -o gpurun_out/prof_gemm -f python tools/ncu_target.py 1 > gpurun_out/ncu_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_stream|layernorm_kernel|temporal_attn_mma|cross_attn_mma" -s 420 -c 12 -o gpurun_out/prof_small -f python tools/ncu_target.py 1 > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
