mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm.py tests/test_backward_ops.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t12.log; cat gpurun_out/t12.log
(for d in 0 1 2; do LVD_GEMM_DEBUG=$d timeout 120 python tools/perf_gemm_sweep.py; done) > gpurun_out/sweep5.log 2>&1; cat gpurun_out/sweep5.log
