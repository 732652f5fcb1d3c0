mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['steps'], d['warmup'], d['cpu_baseline'], d['video']['seconds_per_video'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
tail -2 gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -3
head -c 400 gpurun_out/bench_ref.json
