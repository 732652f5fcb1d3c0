mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm.py tests/test_backward_ops.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t11.log; cat gpurun_out/t11.log
timeout 200 python tools/perf_gemm_sweep.py > gpurun_out/sweep2.log 2>&1; cat gpurun_out/sweep2.log
