"""The Zeroscope GEMM shapes that dominate a step, timed per CTA mode (1 = one CTA per tile, 2 = CTA pairs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops, repack  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def case(name, flops, fn):
    r = []
    for pair in (1, 2):
        ops.FORCE_CTA_PAIR = pair
        r.append(timeit(fn))
    ops.FORCE_CTA_PAIR = 0
    print(f"{name:44s} single {r[0]*1e3:7.1f} us {flops/r[0]/1e9:6.0f} TF/s | pair {r[1]*1e3:7.1f} us {flops/r[1]/1e9:6.0f} TF/s",
          flush=True)


for (m, n, k, res) in [(138240, 960, 320, 0), (138240, 320, 320, 1), (138240, 320, 320, 0), (34560, 640, 640, 1),
                       (34560, 1920, 640, 0), (8640, 1280, 1280, 1), (138240, 320, 1280, 1), (8640, 3840, 1280, 0)]:
    x = torch.randn(m, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half()
    b = torch.randn(n, device=dev).half()
    r_ = torch.randn(m, n, device=dev).half() if res else None
    out = torch.empty(m, n, device=dev, dtype=torch.float16)
    case(f"linear M={m} N={n} K={k} res={res}", 2.0 * m * n * k, lambda: ops.linear(x, w, b, residual=r_, out=out))
for (m, c) in [(138240, 320), (34560, 640), (8640, 1280)]:
    x = torch.randn(m, c, device=dev).half()
    w = torch.randn(8 * c, c, device=dev).half()
    b = torch.randn(8 * c, device=dev).half()
    wi, bi = repack.geglu_interleave(w, b)
    out = torch.empty(m, 4 * c, device=dev, dtype=torch.float16)
    case(f"geglu M={m} C={c}", 2.0 * m * 8 * c * c, lambda: ops.linear(x, wi, bi, act=ops.ACT_GEGLU, out=out))
for (n_img, h, w_, cin, cout) in [(48, 40, 72, 320, 320), (48, 20, 36, 640, 640), (48, 10, 18, 1280, 1280)]:
    x = torch.randn(n_img * h * w_, cin, device=dev).half()
    wk = torch.randn(cout, 9 * cin, device=dev).half()
    out = torch.empty(n_img * h * w_, cout, device=dev, dtype=torch.float16)
    case(f"conv3x3 {n_img}x{h}x{w_} {cin}->{cout}", 2.0 * n_img * h * w_ * cout * 9 * cin,
         lambda: ops.conv2d_3x3(x, n_img, h, w_, wk, out=out))
for (b_, f, hw, c) in [(2, 24, 2880, 320), (2, 24, 720, 640)]:
    x = torch.randn(b_ * f * hw, c, device=dev).half()
    wk = torch.randn(c, 3 * c, device=dev).half()
    out = torch.empty(b_ * f * hw, c, device=dev, dtype=torch.float16)
    case(f"tconv B={b_} F={f} hw={hw} C={c}", 2.0 * b_ * f * hw * c * 3 * c,
         lambda: ops.conv_temporal(x, b_, f, hw, wk, out=out))
