mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_epi_r1 -f python tools/ncu_gemm_one.py > gpurun_out/ncu_epi.log 2>&1; tail -3 gpurun_out/ncu_epi.log
