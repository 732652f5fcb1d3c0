mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_norm_eltwise.py tests/test_unet_forward.py tests/test_pipeline.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof38.json > gpurun_out/bench38.json 2> gpurun_out/bench38.err; python -c "
import json; d=json.load(open('gpurun_out/bench38.json')); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['unguided_step'], d['guided_step'], round(d['roofline']['achieved']), d['clocks'])"; tail -3 gpurun_out/bench38.err
