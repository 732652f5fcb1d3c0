mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t13.log; cat gpurun_out/t13.log
timeout 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/prof15.json > gpurun_out/bench15.json 2> gpurun_out/bench15.err; tail -c 300 gpurun_out/bench15.json; tail -3 gpurun_out/bench15.err
