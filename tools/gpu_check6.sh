mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 6 -c 6 -o gpurun_out/prof_gemm_cases_r1 python tools/ncu_gemm_cases.py > gpurun_out/ncu_cases.log 2>&1; tail -2 gpurun_out/ncu_cases.log
