mkdir -p gpurun_out
for D in 3 7 11 15; do LVD_GEMM_DEBUG=$D timeout 120 python tools/perf_gemm_diag.py 2>&1 | grep -v geglu | grep debug; done > gpurun_out/diag21.log 2>&1
cat gpurun_out/diag21.log
