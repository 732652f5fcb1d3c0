mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider -s 2>&1 | grep -E "FAILED|ERROR|passed|failed|\[unet|\[golden|\[guidance|assert|Error|error" | tail -70 > gpurun_out/t6.log; tail -45 gpurun_out/t6.log
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof6.json > gpurun_out/bench6.json 2> gpurun_out/bench6.err; python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['kernel_families_ms'])"; tail -3 gpurun_out/bench6.err
