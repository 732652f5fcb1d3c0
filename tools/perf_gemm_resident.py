"""Short-K linears of the C = 320 / 640 levels: streaming walk vs weight-resident walk (include/lvd_b200.h), one CTA per
tile vs CTA pairs."""
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
    return ts[len(ts) // 2] * 1e3


def case(name, flops, fn, ref=None):
    r = []
    outs = []
    for res in (0, 1):
        for pair in (1, 2):
            ops.FORCE_WEIGHT_RESIDENT, ops.FORCE_CTA_PAIR = res, pair
            r.append(timeit(fn))
            outs.append(fn().clone())
    ops.FORCE_WEIGHT_RESIDENT, ops.FORCE_CTA_PAIR = None, 0
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    print(f"{name:40s} stream single {r[0]:7.1f} pair {r[1]:7.1f} | resident single {r[2]:7.1f} pair {r[3]:7.1f} us"
          f"   best {flops / min(r) / 1e6:6.0f} TF/s   outputs identical: {same}", flush=True)


for (m, n, k, res) in [(138240, 960, 320, 0), (138240, 320, 320, 1), (138240, 320, 320, 0), (138240, 512, 512, 1),
                       (138240, 1536, 512, 0), (34560, 640, 640, 1), (34560, 640, 640, 0), (34560, 1920, 640, 0),
                       (69120, 960, 320, 0), (69120, 320, 320, 1), (17280, 640, 640, 1), (8640, 1280, 1280, 1)]:
    x = torch.randn(m, k, device=dev).half()
    w = (torch.randn(n, k, device=dev) * k ** -0.5).half()
    b = torch.randn(n, device=dev).half()
    r_ = torch.randn(m, n, device=dev).half() if res else None
    case(f"linear M={m} N={n} K={k} res={res}", 2.0 * m * n * k, lambda: ops.linear(x, w, b, residual=r_))
for (m, c) in [(138240, 320), (34560, 640), (69120, 320)]:
    x = torch.randn(m, c, device=dev).half()
    w = (torch.randn(8 * c, c, device=dev) * c ** -0.5).half()
    b = torch.randn(8 * c, device=dev).half()
    wi, bi = repack.geglu_interleave(w, b)
    case(f"geglu M={m} C={c}", 2.0 * m * 8 * c * c, lambda: ops.linear(x, wi, bi, act=ops.ACT_GEGLU))
