import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
def timeit(fn, iters=7):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]
ops.FORCE_CTA_PAIR = 1
for (m, n, k, res) in [(138240, 960, 320, 0), (138240, 320, 320, 0), (138240, 320, 320, 1), (138240, 1280, 320, 0)]:
    x = torch.randn(m, k, device=dev).half(); w = torch.randn(n, k, device=dev).half(); b = torch.randn(n, device=dev).half()
    r_ = torch.randn(m, n, device=dev).half() if res else None
    out = torch.empty(m, n, device=dev, dtype=torch.float16)
    for bn in ((160, 64) if n % 256 else (256, 128)):
        ops.FORCE_BLOCK_N = bn
        t = timeit(lambda: ops.linear(x, w, b, residual=r_, out=out))
        print(f"debug={os.environ.get('LVD_GEMM_DEBUG','0'):>2s} M={m} N={n} K={k} res={res} bn={bn}: {t*1e3:7.1f} us", flush=True)
