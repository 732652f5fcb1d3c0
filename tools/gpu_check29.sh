mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
for W in lvd_zeroscope lvd_gligen_zeroscope lvd_gligen_zeroscope_guided lvd_plus_modelscope256; do
timeout 500 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof29_$W.json > gpurun_out/bench29_$W.json 2> gpurun_out/bench29_$W.err; python -c "
import json; d=json.load(open('gpurun_out/bench29_$W.json')); print('$W', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['unguided_step'], d['fuser_step'], d['guided_step'], round(d['roofline']['achieved']), d['gpu_launches'])"; tail -3 gpurun_out/bench29_$W.err
done
