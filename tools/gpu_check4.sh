mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider -s 2>&1 | grep -E "FAILED|ERROR|passed|failed|\[pipeline|assert|Error|error" | tail -30 > gpurun_out/t7.log; tail -12 gpurun_out/t7.log
timeout 600 python bench.py --steps 10 --warmup 3 --profile-json gpurun_out/prof7.json > gpurun_out/bench7.json 2> gpurun_out/bench7.err; python -c "
import json; d=json.load(open('gpurun_out/bench7.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['unguided_step'], d['guided_step'], d['roofline']['achieved'], d['cpu_baseline']); print(d['kernel_families_ms']); print(d['guidance_families_ms'])"; tail -3 gpurun_out/bench7.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_guid_r1.csv python tools/ncu_target.py 1 guidance > gpurun_out/ncu_g.log 2>&1; tail -1 gpurun_out/ncu_g.log
