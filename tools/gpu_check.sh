#!/bin/bash
# The GPU-side routine of a round, run on a B200 box as:  gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
#   1. the parity suite (pytest -m gpu; the full-size module's printed error statistics go to gpurun_out/parity.log),
#   2. the bench line + per-kernel-family profile of the headline workload,
#   3. (FULL=1) the other BASELINE workloads and the kernel micro-benchmarks.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_fullsize.py -m gpu --timeout 300 -q --no-header -p no:cacheprovider -s > gpurun_out/parity.log 2>&1
grep -v "^$" gpurun_out/parity.log | tail -60
timeout 900 python -m pytest tests -m gpu --timeout 180 -q --no-header -p no:cacheprovider -x --deselect tests/test_parity_fullsize.py -s > gpurun_out/tests.log 2>&1
tail -4 gpurun_out/tests.log
grep "^\[" gpurun_out/tests.log | head -40
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof.json \
  > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['unguided_step'], d['guided_step'],
      round(d['roofline']['achieved']), d['kernel_families_ms'], d['guidance_families_ms'], d['clocks'])"
tail -3 gpurun_out/bench.err
if [ -n "$FULL" ]; then
  for W in lvd_gligen_zeroscope lvd_gligen_zeroscope_guided lvd_plus_modelscope256; do
    timeout 500 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
    python -c "
import json; d=json.load(open('gpurun_out/bench_$W.json')); print('$W', round(d['value'],1), round(d['e2e']['value'],1), d['unguided_step'], d['fuser_step'], d['guided_step'])"
  done
  timeout 200 python tools/perf_norm.py > gpurun_out/perf_norm.log 2>&1; tail -20 gpurun_out/perf_norm.log
  timeout 200 python tools/perf_attn.py > gpurun_out/perf_attn.log 2>&1; tail -12 gpurun_out/perf_attn.log
  timeout 200 python tools/perf_gemm_shapes.py > gpurun_out/perf_gemm_shapes.log 2>&1; tail -20 gpurun_out/perf_gemm_shapes.log
fi
if [ -n "$FULL" ]; then
  # multi-prompt batching on BASELINE config 4 (lvd-plus ModelScope-256): 4 prompts per GPU, 1 vs 4 prompts per pipeline call
  for P in 1 4; do
    timeout 400 python bench.py --workload lvd_plus_modelscope256 --steps 5 --warmup 3 --no-cpu-baseline --batch-prompts-per-gpu 4 --prompts-per-call $P > gpurun_out/bench_cfg4_ppc$P.json 2> gpurun_out/bench_cfg4_ppc$P.err
    python -c "
import json; d=json.load(open('gpurun_out/bench_cfg4_ppc$P.json')); print('cfg4 prompts-per-call $P', d['prompt_batch'])"
  done
fi
