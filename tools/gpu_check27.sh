mkdir -p gpurun_out
LVD_GN_NO_FUSED=1 PERF_NORM_EAGER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gn_stats|gn_apply" -s 36 -c 4 -o gpurun_out/prof_gn27 -f python tools/perf_norm.py > gpurun_out/ncu27.log 2>&1; tail -3 gpurun_out/ncu27.log
ls -la gpurun_out/*.ncu-rep
