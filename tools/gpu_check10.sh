mkdir -p gpurun_out
(python tools/perf_gemm_sweep.py; LVD_GEMM_STAGES=3 python tools/perf_gemm_sweep.py; LVD_GEMM_STAGES=2 python tools/perf_gemm_sweep.py) > gpurun_out/sweep1.log 2>&1
cat gpurun_out/sweep1.log
