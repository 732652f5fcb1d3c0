mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/prof10.json > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -c 600 gpurun_out/bench10.json; tail -3 gpurun_out/bench10.err
