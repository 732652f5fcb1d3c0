#!/bin/bash
# Same-box A/B of two builds of the library: lvd_b200/liblvd_b200.so (new) against lvd_b200/liblvd_b200_prev.so
# (box-to-box SM clocks differ by +-8 % under the power cap, so only runs of one call compare).  Alternates N times.
mkdir -p gpurun_out
cp lvd_b200/liblvd_b200.so /tmp/new.so
for i in 1 2; do
  for v in new prev; do
    if [ $v = new ]; then cp /tmp/new.so lvd_b200/liblvd_b200.so; else cp lvd_b200/liblvd_b200_prev.so lvd_b200/liblvd_b200.so; fi
    timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-prompt-batch > gpurun_out/ab_${v}_$i.json 2> gpurun_out/ab_${v}_$i.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_${v}_$i.json'))
print('$v $i', round(d['value'],1), d['unguided_step']['ms'], d['guided_step']['ms'], d['kernel_families_ms']['gemm_tc_kernel'], d['guidance_families_ms']['gemm_tc_kernel'], d['clocks']['sm_mhz'])"
  done
done
cp /tmp/new.so lvd_b200/liblvd_b200.so
