mkdir -p gpurun_out
LVD_GEMM_DEBUG=64 N=320 K=320 PAIR=1 SHOW=400 timeout 100 python tools/trace_gemm.py > gpurun_out/trace22.log 2>&1; head -3 gpurun_out/trace22.log
LVD_GEMM_DEBUG=79 N=320 K=320 PAIR=1 SHOW=400 timeout 100 python tools/trace_gemm.py > gpurun_out/trace22b.log 2>&1; head -3 gpurun_out/trace22b.log
