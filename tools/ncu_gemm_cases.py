"""A few representative lvd_gemm launches for `ncu --set full` (one warm-up + one profiled launch each):
0 linear M=138240 N=320 K=320 (+bias)          1 same + residual
2 GEGLU  M=138240 C=320                         3 linear M=138240 N=320 K=1280 (+bias+residual)
4 conv3x3 48x40x72 320->320                     5 linear M=8640 N=1280 K=1280"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from lvd_b200 import ops, repack  # noqa: E402

dev = torch.device("cuda:0")
m = 138240
x = torch.randn(m, 320, device=dev).half()
w = torch.randn(320, 320, device=dev).half()
b = torch.randn(320, device=dev).half()
res = torch.randn(m, 320, device=dev).half()
out = torch.empty(m, 320, device=dev, dtype=torch.float16)
wg = torch.randn(2560, 320, device=dev).half()
bg = torch.randn(2560, device=dev).half()
wi, bi = repack.geglu_interleave(wg, bg)
outg = torch.empty(m, 1280, device=dev, dtype=torch.float16)
x4 = torch.randn(m, 1280, device=dev).half()
w4 = torch.randn(320, 1280, device=dev).half()
wk = torch.randn(320, 9 * 320, device=dev).half()
x5 = torch.randn(8640, 1280, device=dev).half()
w5 = torch.randn(1280, 1280, device=dev).half()
out5 = torch.empty(8640, 1280, device=dev, dtype=torch.float16)
cases = [
    lambda: ops.linear(x, w, b, out=out),
    lambda: ops.linear(x, w, b, residual=res, out=out),
    lambda: ops.linear(x, wi, bi, act=ops.ACT_GEGLU, out=outg),
    lambda: ops.linear(x4, w4, b, residual=res, out=out),
    lambda: ops.conv2d_3x3(x, 48, 40, 72, wk, b, out=out),
    lambda: ops.linear(x5, w5, out=out5),
]
for rep in range(2):
    for c in cases:
        c()
torch.cuda.synchronize()
print("done")
