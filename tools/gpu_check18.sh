mkdir -p gpurun_out
ITERS=20 timeout 600 python tools/stress_determinism.py > gpurun_out/stress1.log 2>&1; tail -4 gpurun_out/stress1.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E |unet tiny|passed|failed|FAILED" | head; 
timeout 300 python -m pytest tests/test_unet_forward.py tests/test_guidance.py -m gpu -q -s 2>&1 | grep -E "\[|passed|failed" | head -30
