mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/t19.log; tail -8 gpurun_out/t19.log
timeout 500 python bench.py --steps 10 --warmup 3 --profile-json gpurun_out/prof19.json > gpurun_out/bench19.json 2> gpurun_out/bench19.err; python -c "
import json; d=json.load(open('gpurun_out/bench19.json')); print(d['value'], d['ms_per_step'], d['unguided_step'], d['guided_step'], d['roofline']['achieved'], d['kernel_families_ms'], d['guidance_families_ms'], d.get('cpu_baseline'))"; tail -3 gpurun_out/bench19.err
timeout 200 python tools/perf_gemm_shapes.py > gpurun_out/shapes19.log 2>&1; tail -40 gpurun_out/shapes19.log
