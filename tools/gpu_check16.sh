mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm.py tests/test_backward_ops.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t16.log; cat gpurun_out/t16.log
timeout 300 python tools/perf_gemm_shapes.py > gpurun_out/shapes3.log 2>&1; cat gpurun_out/shapes3.log
