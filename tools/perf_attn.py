"""Attention kernels at the Zeroscope shapes (CUDA-graph timed: no host latency between launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def bench(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


which = os.environ.get("WHICH", "temporal,self,cross").split(",")
if "temporal" in which:
    for (b, hw, f, heads) in [(2, 2880, 24, 5), (2, 2880, 24, 8), (2, 720, 24, 10), (2, 180, 24, 20), (2, 1024, 16, 5)]:
        c = heads * 64
        qkv = torch.randn(b * hw * f, 3 * c, device=dev).half()
        out = torch.empty(b * hw * f, c, device=dev, dtype=torch.float16)
        t = bench(lambda: ops.attention(qkv, qkv, heads=heads, nz=b, temporal=True, packed=True, frames=f, hw=hw,
                                        k_col0=c, v_col0=2 * c, out=out))
        gb = qkv.numel() * 2 + out.numel() * 2
        print(f"temporal B={b} hw={hw} F={f} heads={heads}: {t*1e3:7.1f} us  {gb/t/1e6:6.0f} GB/s", flush=True)
if "self" in which:
    for (nz, s_, heads) in [(48, 2880, 5), (48, 720, 10), (48, 180, 20), (24, 2880, 5)]:
        c = heads * 64
        qkv = torch.randn(nz * s_, 3 * c, device=dev).half()
        out = torch.empty(nz * s_, c, device=dev, dtype=torch.float16)
        t = bench(lambda: ops.attention(qkv, qkv, heads=heads, nz=nz, s_q=s_, s_k=s_, k_col0=c, v_col0=2 * c, out=out))
        fl = 4.0 * nz * heads * s_ * s_ * 64
        print(f"self nz={nz} S={s_} heads={heads}: {t*1e3:7.1f} us  {fl/t/1e9:6.0f} TFLOP/s", flush=True)
if "cross" in which:
    for (nz, s_, heads) in [(48, 2880, 5), (48, 720, 10)]:
        c = heads * 64
        q = torch.randn(nz * s_, c, device=dev).half()
        kv = torch.randn(2 * 77, 2 * c, device=dev).half()
        out = torch.empty(nz * s_, c, device=dev, dtype=torch.float16)
        t = bench(lambda: ops.attention(q, kv, heads=heads, nz=nz, s_q=s_, s_k=77, v_col0=c, kv_z_div=24, out=out))
        print(f"cross nz={nz} S={s_} heads={heads}: {t*1e3:7.1f} us", flush=True)
