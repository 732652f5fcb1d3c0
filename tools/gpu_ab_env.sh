#!/bin/bash
# Same-box A/B of an environment switch: tools/gpu_ab_env.sh "VAR=1" "VAR=0" (each run twice, alternating)
mkdir -p gpurun_out
for i in 1 2; do
  for v in "$1" "$2"; do
    tag=$(echo "$v" | tr -c 'A-Za-z0-9=_\n' '_')
    env $v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-prompt-batch > gpurun_out/abenv_${tag}_$i.json 2> gpurun_out/abenv_${tag}_$i.err
    python -c "
import json; d=json.load(open('gpurun_out/abenv_${tag}_$i.json'))
print('$v $i', round(d['value'],1), d['unguided_step']['ms'], d['guided_step']['ms'], d['kernel_families_ms']['gemm_tc_kernel'], d['guidance_families_ms']['gemm_tc_kernel'], d['clocks']['sm_mhz'])"
  done
done
