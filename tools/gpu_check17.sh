mkdir -p gpurun_out
LVD_GEMM_DEBUG=64 SHOW=100000 timeout 100 python tools/trace_gemm.py > gpurun_out/trace1.log 2>&1; head -3 gpurun_out/trace1.log
