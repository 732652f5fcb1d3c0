"""Timeline of one CTA of an lvd_gemm launch (LVD_GEMM_DEBUG=64 must be set): who waits for whom."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
m, n, k = 138240, int(os.environ.get("N", "960")), int(os.environ.get("K", "320"))
x = torch.randn(m, k, device=dev).half()
w = torch.randn(n, k, device=dev).half()
b = torch.randn(n, device=dev).half()
out = torch.empty(m, n, device=dev, dtype=torch.float16)
ops.FORCE_CTA_PAIR = int(os.environ.get("PAIR", "1"))
buf = (ctypes.c_ulonglong * 8192)()
lib = _lib.lib()
for _ in range(3):
    ops.linear(x, w, b, out=out)
    torch.cuda.synchronize()
    nev = lib.lvd_debug_trace_read(buf, 8192)
ev = sorted(((buf[i] & 0xFFFFFFFFFFFF, buf[i] >> 48) for i in range(nev)))
t0 = ev[0][0]
names = {1: "prod tile", 2: "mma wait-acc", 3: "mma go", 4: "mma issued"}
for g_ in range(4):
    for j, nm in enumerate(["wait-tile", "tile-ready", "math-done", "fence+drain", "bar", "released"]):
        names[16 + g_ * 16 + j] = f"epi{g_} {nm}"
print(f"{nev} events")
for t, e in ev[:int(os.environ.get("SHOW", "260"))]:
    print(f"{t - t0:8d} ns  {names.get(e, e)}")
