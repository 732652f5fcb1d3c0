mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d_cfg_forward.csv python tools/ncu_target.py 1 > gpurun_out/ncu36a.log 2>&1; tail -1 gpurun_out/ncu36a.log; wc -l gpurun_out/launches_r1d_cfg_forward.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d_guidance_iteration.csv python tools/ncu_target.py 1 guidance > gpurun_out/ncu36b.log 2>&1; tail -1 gpurun_out/ncu36b.log; wc -l gpurun_out/launches_r1d_guidance_iteration.csv
# full captures: skip the warm-up forward (first ~1100 launches), then a few launches of each family
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"temporal_attn_mma|cross_attn_mma|gn_stream|layernorm_kernel" -s 420 -c 14 -o gpurun_out/prof_small_r1d -f python tools/ncu_target.py 1 > gpurun_out/ncu36c.log 2>&1; tail -1 gpurun_out/ncu36c.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 470 -c 12 -o gpurun_out/prof_gemm_r1d -f python tools/ncu_target.py 1 > gpurun_out/ncu36d.log 2>&1; tail -1 gpurun_out/ncu36d.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 17 -c 3 -o gpurun_out/prof_attn_r1d -f python tools/ncu_target.py 1 > gpurun_out/ncu36e.log 2>&1; tail -1 gpurun_out/ncu36e.log
ls -la gpurun_out/*.ncu-rep
