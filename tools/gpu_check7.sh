mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider -s 2>&1 | grep -E "FAILED|ERROR|passed|failed|\[unet|\[pipeline|\[guidance|assert|Error|error" | tail -40 > gpurun_out/t9.log; tail -22 gpurun_out/t9.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof9.json > gpurun_out/bench9.json 2> gpurun_out/bench9.err; python -c "
import json; d=json.load(open('gpurun_out/bench9.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['unguided_step'], d['guided_step'], d['roofline']['achieved']); print(d['kernel_families_ms']); print(d['guidance_families_ms'])"; tail -3 gpurun_out/bench9.err
timeout 300 python tools/perf_gemm.py gpurun_out/perf9.json 2>&1 | head -30 | grep -v torch
