"""One lvd_gemm launch (the epilogue-bound level-0 QKV projection shape) for `ncu --set full --import-source on`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops
dev = torch.device("cuda:0")
m, n, k = 138240, 960, 320
x = torch.randn(m, k, device=dev).half(); w = torch.randn(n, k, device=dev).half(); b = torch.randn(n, device=dev).half()
out = torch.empty(m, n, device=dev, dtype=torch.float16)
ops.FORCE_CTA_PAIR = int(os.environ.get("PAIR", "1"))
for _ in range(3):
    ops.linear(x, w, b, out=out)
torch.cuda.synchronize()
