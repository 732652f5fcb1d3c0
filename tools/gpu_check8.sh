mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_traffic_r1.csv python tools/ncu_target.py 1 > gpurun_out/ncu_t.log 2>&1; tail -1 gpurun_out/ncu_t.log; wc -l gpurun_out/gemm_traffic_r1.csv
