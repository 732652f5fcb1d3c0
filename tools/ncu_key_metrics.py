"""Key metrics of `ncu --set full` captures (.ncu-rep files): python tools/ncu_key_metrics.py a.ncu-rep [b.ncu-rep ...]
One line per captured launch: duration, DRAM traffic / rate, pipe utilisation, occupancy, registers."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"== {rep}")
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        parts = [f"{name[:36]:36s}"]
        for k in KEYS:
            if k in col:
                v = r[col[k]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                u = units[col[k]]
                parts.append(f"{k.split('.')[0].replace('__', ':')}={v}{'' if u in ('', 'block', 'inst') else u}")
        print("  " + "  ".join(parts))
