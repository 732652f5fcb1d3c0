mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_norm_eltwise.py tests/test_determinism.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
echo FUSED; timeout 200 python tools/perf_norm.py 2>&1 | grep groupnorm | tee gpurun_out/norm26_fused.log
echo UNFUSED; LVD_GN_NO_FUSED=1 timeout 200 python tools/perf_norm.py 2>&1 | grep groupnorm | tee gpurun_out/norm26_unfused.log
