mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rA -p no:cacheprovider -s 2>&1 | grep -E "FAILED|ERROR|passed|failed|\[unet|\[golden|assert|Error" | tail -60 > gpurun_out/t4.log; tail -30 gpurun_out/t4.log
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/prof4.json > gpurun_out/bench4.json 2> gpurun_out/bench4.err; python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['kernel_families_ms'])"; tail -3 gpurun_out/bench4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/ncu_target.py 1 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log; wc -l gpurun_out/launches_r1.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 500 -c 4 -o gpurun_out/prof_gemm_r1 python tools/ncu_target.py 1 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 70 -c 4 -o gpurun_out/prof_attn_r1 python tools/ncu_target.py 1 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
ls -la gpurun_out | tail -8
