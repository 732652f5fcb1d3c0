"""Tile width vs wave quantisation on the small-M GEMMs (the 10x18 / 5x9 levels; guidance pass = half the rows):
times every admissible block_n for Linear and 3x3-conv shapes, CUDA-graph timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def bench(fn, n=8):
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


cases = []
for n_img, h, w in ((24, 10, 18), (48, 10, 18), (24, 5, 9), (48, 5, 9), (24, 20, 36)):
    for c in ((1280,) if h <= 10 else (640,)):
        cases.append(("linear", n_img, h, w, c, c))
        cases.append(("conv3x3", n_img, h, w, c, c))
        cases.append(("tconv", n_img, h, w, c, c))
for kind, n_img, h, w, cin, cout in cases:
    rows = n_img * h * w
    x = torch.randn(rows, cin, device=dev).half()
    taps = {"linear": 1, "conv3x3": 9, "tconv": 3}[kind]
    wk = (torch.randn(cout, taps * cin, device=dev) * 0.02).half()
    res = []
    for bn in (None, 64, 128, 160, 256):
        if bn is not None and cout % bn:
            continue
        ops.FORCE_BLOCK_N = bn
        if kind == "linear":
            fn = lambda: ops.linear(x, wk)  # noqa: E731
        elif kind == "conv3x3":
            fn = lambda: ops.conv2d_3x3(x, n_img, h, w, wk)  # noqa: E731
        else:
            fn = lambda: ops.conv_temporal(x, n_img // 24, 24, h * w, wk)  # noqa: E731
        t = bench(fn)
        res.append(f"bn={'auto' if bn is None else bn}: {t * 1e3:6.1f} us {2.0 * rows * cout * taps * cin / t / 1e9:5.0f} TF/s")
    ops.FORCE_BLOCK_N = None
    print(f"{kind:8s} M={rows:6d} N={cout} K={taps * cin:5d} | " + " | ".join(res), flush=True)
