mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider -s 2>&1 | grep -E "FAILED|ERROR|passed|failed|\[unet|\[golden|assert|Error|error" | tail -70 > gpurun_out/t5.log; tail -40 gpurun_out/t5.log
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof5.json > gpurun_out/bench5.json 2> gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['kernel_families_ms'])"; tail -3 gpurun_out/bench5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python tools/ncu_target.py 1 > gpurun_out/ncu1b.log 2>&1; tail -1 gpurun_out/ncu1b.log
