mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention.py tests/test_unet_forward.py tests/test_guidance.py tests/test_pipeline.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
WHICH=cross timeout 100 python tools/perf_attn.py 2>&1 | tail -3
LVD_CROSS_TILE=1 WHICH=cross timeout 100 python tools/perf_attn.py 2>&1 | tail -3
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof33.json > gpurun_out/bench33.json 2> gpurun_out/bench33.err; python -c "
import json; d=json.load(open('gpurun_out/bench33.json')); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['unguided_step'], d['guided_step'], round(d['roofline']['achieved']), d['kernel_families_ms'], d['clocks'])"; tail -3 gpurun_out/bench33.err
