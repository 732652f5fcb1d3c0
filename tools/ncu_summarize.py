"""Summaries of the Nsight Compute launch lists written by tools/gpu_ncu.sh (duration + DRAM bytes of every launch):
  python tools/ncu_summarize.py <dir with launches_*.csv> <out prefix, e.g. profiles/r2>
writes <prefix>_launch_list_summary.txt (per-kernel totals / shares of the LAST pass of each list) and
<prefix>_gemm_traffic.json (DRAM bytes per gemm_tc_kernel launch over one CFG forward + 1/5 guidance iteration: the
step mix bench.py times)."""
import csv
import json
import sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    launches = OrderedDict()
    for r in rows:
        k = int(r["ID"])
        d = launches.setdefault(k, {"name": r["Kernel Name"], "t": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["t"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        else:
            b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d["rd" if "read" in r["Metric Name"] else "wr"] = b
    return list(launches.values())


def last_pass(launches, marker):
    """The target runs a warm-up pass first: keep the launches after the last occurrence of the pass's first kernel."""
    idx = [i for i, l in enumerate(launches) if marker in l["name"]]
    n_pass = len(idx)
    return launches[idx[-1]:] if idx else launches, n_pass


def short(name):
    n = name.split("(")[0]
    n = n.replace("lvd::", "").replace("void ", "")
    return n.split("<")[0][:44]


def summarize(launches, title):
    tot = sum(l["t"] for l in launches)
    fam = OrderedDict()
    for l in launches:
        d = fam.setdefault(short(l["name"]), {"n": 0, "t": 0.0, "b": 0.0})
        d["n"] += 1
        d["t"] += l["t"]
        d["b"] += l["rd"] + l["wr"]
    out = [f"{title}: {len(launches)} launches, {tot:.2f} ms summed (cold-cache, serialised), "
           f"{sum(d['b'] for d in fam.values()) / 1e9:.1f} GB DRAM traffic"]
    for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["t"]):
        out.append(f"  {k:44s} n={d['n']:4d} ms={d['t']:8.3f} {100 * d['t'] / tot:5.1f}%  "
                   f"{d['b'] / 1e9:7.2f} GB  {d['b'] / max(d['t'], 1e-9) / 1e6:6.0f} GB/s")
    return out, fam


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    fwd, nf = last_pass(load(f"{src}/launches_cfg_forward.csv"), "timestep_embed")
    gui, ng = last_pass(load(f"{src}/launches_guidance_iteration.csv"), "timestep_embed")
    a, fam_f = summarize(fwd, f"cfg_forward (last of {nf} passes)")
    b, fam_g = summarize(gui, f"guidance_iteration (last of {ng} passes)")
    open(prefix + "_launch_list_summary.txt", "w").write("\n".join(a + b) + "\n")
    gf, gg = fam_f.get("gemm_tc_kernel", {"n": 0, "t": 0, "b": 0}), fam_g.get("gemm_tc_kernel", {"n": 0, "t": 0, "b": 0})
    n = gf["n"] + gg["n"] / 5.0
    by = gf["b"] + gg["b"] / 5.0
    json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         "--clock-control none python tools/ncu_target.py 1 [guidance]: every gemm_tc_kernel launch of "
                         "one CFG forward (Zeroscope 24f, B=2) + 1/5 of one guidance iteration (B=1)",
               "launches_cfg_forward": gf["n"], "launches_guidance_iteration": gg["n"],
               "dram_bytes_cfg_forward": gf["b"], "dram_bytes_guidance_iteration": gg["b"],
               "launches_per_step": n, "traffic_bytes_per_launch": by / max(n, 1e-9),
               "ncu_ms_cfg_forward": gf["t"], "ncu_ms_guidance_iteration": gg["t"]},
              open(prefix + "_gemm_traffic.json", "w"), indent=1)
    print("\n".join(a[:8] + b[:8]))


if __name__ == "__main__":
    main()
