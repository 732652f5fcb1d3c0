"""Attribution for the small-K GEMM shapes: run once per LVD_GEMM_DEBUG value (0 = real, 1 = no TMA loads, 2 = no MMAs,
3 = neither: barrier flow + epilogue only).  Results are garbage for debug != 0; only the times matter."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops, repack  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=9):
    for _ in range(3):
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


dbg = os.environ.get("LVD_GEMM_DEBUG", "0")
for (m, n, k, res) in [(138240, 320, 320, 1), (138240, 320, 320, 0), (34560, 640, 640, 1), (8640, 1280, 1280, 1),
                       (138240, 960, 320, 0), (138240, 320, 1280, 1)]:
    x = torch.randn(m, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half()
    b = torch.randn(n, device=dev).half()
    r_ = torch.randn(m, n, device=dev).half() if res else None
    out = torch.empty(m, n, device=dev, dtype=torch.float16)
    for pair in (1, 2):
        ops.FORCE_CTA_PAIR = pair
        t = timeit(lambda: ops.linear(x, w, b, residual=r_, out=out))
        print(f"debug={dbg} M={m} N={n} K={k} res={res} ctas={pair}: {t*1e3:7.1f} us", flush=True)
for (m, c) in [(138240, 320)]:
    x = torch.randn(m, c, device=dev).half()
    w = torch.randn(8 * c, c, device=dev).half()
    b = torch.randn(8 * c, device=dev).half()
    wi, bi = repack.geglu_interleave(w, b)
    out = torch.empty(m, 4 * c, device=dev, dtype=torch.float16)
    for pair in (1, 2):
        ops.FORCE_CTA_PAIR = pair
        t = timeit(lambda: ops.linear(x, wi, bi, act=ops.ACT_GEGLU, out=out))
        print(f"debug={dbg} geglu M={m} C={c} ctas={pair}: {t*1e3:7.1f} us", flush=True)
