"""Where does lvd_gemm's time go?  K sweep at one n-tile (constant output traffic) per tile width / CTA mode; the slope
between two K values is the steady-state cost of one 64-deep k-block.  LVD_GEMM_DEBUG=1 skips the TMA loads (tensor
pipe + epilogue only), =2 skips the MMAs (operand delivery + epilogue only).  Results are wrong in those modes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=5):
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


dbg = os.environ.get("LVD_GEMM_DEBUG", "0")
m = 148 * 128 * 8
tiles_per_sm = 8
for n in (256, 320):
    for pair in (1, 2):
        ops.FORCE_CTA_PAIR = pair
        for bn in ((64, 128, 256) if n == 256 else (160,)):
            ops.FORCE_BLOCK_N = bn
            t = {}
            for k in (1024, 4096):
                x = torch.randn(m, k, device=dev).half()
                w = torch.randn(n, k, device=dev).half()
                out = torch.empty(m, n, device=dev, dtype=torch.float16)
                t[k] = timeit(lambda: ops.linear(x, w, out=out))
                del x, w, out
            per_sm_tiles = tiles_per_sm * (n // bn)
            kb_ns = (t[4096] - t[1024]) * 1e6 / (48 * per_sm_tiles)
            print(f"debug={dbg} ctas={pair} N={n} block_n={bn:3d}: K=1024 {t[1024]:.3f} ms, K=4096 {t[4096]:.3f} ms "
                  f"({2.0 * m * n * 4096 / t[4096] / 1e9:6.0f} TFLOP/s) -> {kb_ns:6.1f} ns per k-block per CTA", flush=True)
ops.FORCE_BLOCK_N = None
ops.FORCE_CTA_PAIR = 0
