mkdir -p gpurun_out
for D in 0 1 2 3; do LVD_GEMM_DEBUG=$D timeout 120 python tools/perf_gemm_diag.py; done > gpurun_out/diag20.log 2>&1
cat gpurun_out/diag20.log
