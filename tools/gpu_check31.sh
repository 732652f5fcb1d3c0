mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention.py tests/test_unet_forward.py tests/test_guidance.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
echo NEW; WHICH=temporal timeout 100 python tools/perf_attn.py 2>&1 | tail -6
echo TILE; LVD_TEMPORAL_TILE=1 WHICH=temporal,self,cross timeout 100 python tools/perf_attn.py 2>&1 | tail -12
