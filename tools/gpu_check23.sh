mkdir -p gpurun_out
(RES=0 python tools/phase_gemm.py; RES=1 python tools/phase_gemm.py; M=34560 N=640 K=640 RES=1 python tools/phase_gemm.py) > gpurun_out/phase23.log 2>&1
cat gpurun_out/phase23.log
