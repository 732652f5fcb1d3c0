mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof28.json > gpurun_out/bench28.json 2> gpurun_out/bench28.err; python -c "
import json; d=json.load(open('gpurun_out/bench28.json')); print(d['value'], d['ms_per_step'], d['unguided_step'], d['guided_step'], d['roofline']['achieved'], d['kernel_families_ms'], d['guidance_families_ms'])"; tail -3 gpurun_out/bench28.err
