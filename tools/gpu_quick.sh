#!/bin/bash
# Quick iteration routine: GPU tests (no full-size module), GEMM / attention micro-benchmarks, short bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu --timeout 180 -q --no-header -p no:cacheprovider -x --deselect tests/test_parity_fullsize.py > gpurun_out/tests.log 2>&1
tail -4 gpurun_out/tests.log
timeout 200 python tools/perf_gemm_shapes.py > gpurun_out/perf_gemm_shapes.log 2>&1; cat gpurun_out/perf_gemm_shapes.log
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof.json \
  > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print(round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['unguided_step'], d['guided_step'],
      round(d['roofline']['achieved']), d['kernel_families_ms'], d['guidance_families_ms'], d['clocks'])"
tail -3 gpurun_out/bench.err
