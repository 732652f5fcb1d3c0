"""Epilogue phase profile of lvd_gemm (library built with LVD_EXTRA_NVCC_FLAGS=-DLVD_GEMM_PHASE_PROF): clock cycles per
phase for each epilogue warp of CTA 0."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
m, n, k = int(os.environ.get("M", "138240")), int(os.environ.get("N", "320")), int(os.environ.get("K", "320"))
res = int(os.environ.get("RES", "0"))
x = torch.randn(m, k, device=dev).half()
w = torch.randn(n, k, device=dev).half()
b = torch.randn(n, device=dev).half()
r_ = torch.randn(m, n, device=dev).half() if res else None
out = torch.empty(m, n, device=dev, dtype=torch.float16)
ops.FORCE_CTA_PAIR = int(os.environ.get("PAIR", "1"))
buf = (ctypes.c_ulonglong * 8192)()
lib = _lib.lib()
for _ in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.linear(x, w, b, residual=r_, out=out)
    e.record()
    torch.cuda.synchronize()
    nev = lib.lvd_debug_trace_read(buf, 8192)
print(f"M={m} N={n} K={k} res={res} pair={ops.FORCE_CTA_PAIR}: {s.elapsed_time(e)*1e3:.1f} us, {nev} words")
names = ["wait-tile", "res-issue", "tmem+math", "release+sts+fence", "drain", "bar", "store+prefetch", "advance", "tail", "-"]
print("warp " + " ".join(f"{x:>18s}" for x in names[:9]) + "   total")
for wp in range(2, 18):
    v = [buf[wp * 16 + i] for i in range(9)]
    print(f"{wp:4d} " + " ".join(f"{x:18d}" for x in v) + f"   {sum(v)}")
