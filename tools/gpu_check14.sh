mkdir -p gpurun_out
(for d in 0 16 32 48 3 19; do LVD_GEMM_DEBUG=$d timeout 100 python tools/perf_gemm_epi.py; done) > gpurun_out/epi1.log 2>&1; cat gpurun_out/epi1.log
