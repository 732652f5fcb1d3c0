"""Target for `ncu --set full --import-source on -k regex:gemm_tc_kernel`: the short-K linears of the C = 320 / 640
levels, two launches each (the first one warms the caches), in this order: N960 K320 | N320 K320 + residual |
GEGLU C320 | M34560 N640 K640 + residual."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops, repack  # noqa: E402

dev = torch.device("cuda:0")
m = 138240
for (mm, n, k, res, geglu) in [(m, 960, 320, 0, 0), (m, 320, 320, 1, 0), (m, 2560, 320, 0, 1), (34560, 640, 640, 1, 0)]:
    x = torch.randn(mm, k, device=dev).half()
    w = torch.randn(n, k, device=dev).half()
    b = torch.randn(n, device=dev).half()
    if geglu:
        w, b = repack.geglu_interleave(w, b)
    r_ = torch.randn(mm, n, device=dev).half() if res else None
    out = torch.empty(mm, n // 2 if geglu else n, device=dev, dtype=torch.float16)
    for _ in range(2):
        ops.linear(x, w, b, residual=r_, out=out, act=ops.ACT_GEGLU if geglu else ops.ACT_NONE)
    torch.cuda.synchronize()
print("done")
