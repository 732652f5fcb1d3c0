timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout 200 python tools/perf_norm.py 2>&1 | grep groupnorm | cut -c1-100
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-video > gpurun_out/bench_gn.json 2> gpurun_out/bench_gn.err; python -c "
import json; d=json.load(open('gpurun_out/bench_gn.json')); print(round(d['value'],1), d['unguided_step']['ms'], d['guided_step']['ms'], d['clocks']['sm_mhz'], d['kernel_families_ms'])"
