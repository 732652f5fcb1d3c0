"""Which side bounds the K = 320 linears?  Each shape is timed with parts of lvd_gemm switched off through
LVD_GEMM_DEBUG (bit 0 = no TMA operand loads, 1 = no MMAs, 2 = no output stores, 3 = no residual loads; results are
garbage in those modes).  The flag is read once per process, so every mode runs in a child process."""
import os
import subprocess
import sys

MODES = [0, 1, 2, 4, 8, 3, 5, 6, 7, 15]

if len(sys.argv) > 1 and sys.argv[1] == "child":
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

    out_line = []
    m = 138240
    for (n, k, res) in [(960, 320, 0), (320, 320, 0), (320, 320, 1), (320, 1280, 1)]:
        x = torch.randn(m, k, device=dev).half()
        w = torch.randn(n, k, device=dev).half()
        b = torch.randn(n, device=dev).half()
        r_ = torch.randn(m, n, device=dev).half() if res else None
        out = torch.empty(m, n, device=dev, dtype=torch.float16)
        for pair in (1, 2):
            ops.FORCE_CTA_PAIR = pair
            out_line.append(timeit(lambda: ops.linear(x, w, b, residual=r_, out=out)))
    x = torch.randn(m, 320, device=dev).half()
    w = torch.randn(2560, 320, device=dev).half()
    b = torch.randn(2560, device=dev).half()
    wi, bi = repack.geglu_interleave(w, b)
    out = torch.empty(m, 1280, device=dev, dtype=torch.float16)
    for pair in (1, 2):
        ops.FORCE_CTA_PAIR = pair
        out_line.append(timeit(lambda: ops.linear(x, wi, bi, act=ops.ACT_GEGLU, out=out)))
    print(" ".join(f"{t:7.1f}" for t in out_line), flush=True)
    sys.exit(0)

print("us per launch at M = 138240 (single | pair):   N960 K320 | N320 K320 | N320 K320 +res | N320 K1280 +res | GEGLU C320")
for mode in MODES:
    env = dict(os.environ, LVD_GEMM_DEBUG=str(mode))
    r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    tag = "+".join(nm for bit, nm in ((1, "noload"), (2, "nomma"), (4, "nostore"), (8, "nores")) if mode & bit) or "full"
    print(f"{tag:28s} {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
