"""Micro-benchmark of lvd_gemm on the Zeroscope problem sizes; prints TFLOP/s per shape.
CUDA-event timing, 3 warm-ups, L2 flushed between iterations.  Usage: python tools/perf_gemm.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops, repack  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


res = []


def rec(name, flops, ms):
    res.append({"name": name, "ms": ms, "tflops": flops / ms / 1e9})
    print(f"{name:50s} {ms:8.3f} ms {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


for (m, n, k) in [(138240, 320, 320), (34560, 640, 640), (8640, 1280, 1280), (138240, 320, 1280),
                  (34560, 640, 2560), (8640, 1280, 5120), (8192, 8192, 8192)]:
    x = torch.randn(m, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half()
    out = torch.empty(m, n, device=dev, dtype=torch.float16)
    rec(f"linear M={m} N={n} K={k}", 2.0 * m * n * k, timeit(lambda: ops.linear(x, w, out=out)))
    rec(f"  torch.matmul same", 2.0 * m * n * k, timeit(lambda: torch.matmul(x, w.t(), out=out)))
for (m, c) in [(138240, 320), (34560, 640), (8640, 1280)]:
    x = torch.randn(m, c, device=dev).half()
    w = torch.randn(8 * c, c, device=dev).half()
    b = torch.randn(8 * c, device=dev).half()
    wi, bi = repack.geglu_interleave(w, b)
    out = torch.empty(m, 4 * c, device=dev, dtype=torch.float16)
    rec(f"geglu M={m} C={c}", 2.0 * m * 8 * c * c, timeit(lambda: ops.linear(x, wi, bi, act=ops.ACT_GEGLU, out=out)))
for (n_img, h, w_, cin, cout) in [(48, 40, 72, 320, 320), (48, 20, 36, 640, 640), (48, 10, 18, 1280, 1280),
                                  (48, 5, 9, 1280, 1280), (48, 10, 18, 2560, 1280), (48, 40, 72, 960, 320)]:
    x = torch.randn(n_img * h * w_, cin, device=dev).half()
    wk = torch.randn(cout, 9 * cin, device=dev).half()
    out = torch.empty(n_img * h * w_, cout, device=dev, dtype=torch.float16)
    rec(f"conv3x3 N={n_img} {h}x{w_} {cin}->{cout}", 2.0 * n_img * h * w_ * cout * 9 * cin,
        timeit(lambda: ops.conv2d_3x3(x, n_img, h, w_, wk, out=out)))
for (b, f, hw, c) in [(2, 24, 2880, 320), (2, 24, 720, 640), (2, 24, 180, 1280), (2, 24, 45, 1280)]:
    x = torch.randn(b * f * hw, c, device=dev).half()
    wk = torch.randn(c, 3 * c, device=dev).half()
    out = torch.empty(b * f * hw, c, device=dev, dtype=torch.float16)
    rec(f"tconv B={b} F={f} hw={hw} C={c}", 2.0 * b * f * hw * c * 3 * c,
        timeit(lambda: ops.conv_temporal(x, b, f, hw, wk, out=out)))
for (z, s, heads) in [(48, 2880, 5), (48, 720, 10), (48, 180, 20), (48, 45, 20)]:
    c = heads * 64
    qkv = torch.randn(z * s, 3 * c, device=dev).half()
    out = torch.empty(z * s, c, device=dev, dtype=torch.float16)
    rec(f"self-attn z={z} S={s} heads={heads}", 4.0 * z * heads * s * s * 64,
        timeit(lambda: ops.attention(qkv, qkv, heads=heads, nz=z, s_q=s, s_k=s, k_col0=c, v_col0=2 * c, out=out)))
for (b, f, hw, heads) in [(2, 24, 2880, 5), (2, 24, 720, 10), (2, 24, 180, 20)]:
    c = heads * 64
    qkv = torch.randn(b * f * hw, 3 * c, device=dev).half()
    out = torch.empty(b * f * hw, c, device=dev, dtype=torch.float16)
    rec(f"temporal-attn B={b} hw={hw} heads={heads}", 4.0 * b * hw * heads * f * f * 64,
        timeit(lambda: ops.attention(qkv, qkv, heads=heads, nz=b, temporal=True, frames=f, hw=hw, k_col0=c,
                                     v_col0=2 * c, out=out)))
for (samples, rows, c) in [(48, 2880, 320), (2, 24 * 2880, 320), (48, 720, 640), (48, 180, 1280)]:
    x = torch.randn(samples * rows, c, device=dev).half()
    ga = torch.ones(c, device=dev).half()
    out = torch.empty_like(x)
    ms = timeit(lambda: ops.groupnorm(x, samples, ga, ga, 1e-5, silu=True, out=out))
    res.append({"name": f"groupnorm {samples}x{rows}x{c}", "ms": ms, "gbs": 3 * x.numel() * 2 / ms / 1e6})
    print(f"groupnorm {samples}x{rows}x{c}: {ms:.3f} ms, {3 * x.numel() * 2 / ms / 1e6:.0f} GB/s (2 reads + 1 write)")
    ms = timeit(lambda: ops.layernorm(x, ga, ga, out=out))
    res.append({"name": f"layernorm {samples * rows}x{c}", "ms": ms, "gbs": 2 * x.numel() * 2 / ms / 1e6})
    print(f"layernorm {samples * rows}x{c}: {ms:.3f} ms, {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s")
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
