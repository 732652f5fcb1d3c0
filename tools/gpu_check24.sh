mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_norm_eltwise.py tests/test_backward_ops.py tests/test_determinism.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -5
timeout 200 python tools/perf_norm.py > gpurun_out/norm24.log 2>&1; cat gpurun_out/norm24.log
