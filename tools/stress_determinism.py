"""Run the tiny UNet forward (with and without GLIGEN) many times; any run-to-run difference beyond GroupNorm's
atomic-order noise points at a race."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lvd_b200 import synthetic  # noqa: E402
from lvd_b200.engine import UNetEngine  # noqa: E402

dev = torch.device("cuda:0")
iters = int(os.environ.get("ITERS", "40"))
for gated in (True, False):
    cfg = synthetic.make_cfg("tiny", gated=gated)
    sd16 = {k: v.half() for k, v in synthetic.random_state_dict(cfg, seed=2).items()}
    b, f, h, w = 2, 4, 16, 16
    g = torch.Generator().manual_seed(3)
    x = torch.randn((b, cfg["in_channels"], f, h, w), generator=g).half().to(dev)
    ehs = torch.randn((b, 77, cfg["cross_attention_dim"]), generator=g).half().to(dev)
    gl = None
    if gated:
        boxes = torch.rand((b * f, 30, 4), generator=g).half().to(dev)
        masks = (torch.rand((b * f, 30), generator=g) > 0.6).half().to(dev)
        pe = torch.randn((b * f, 30, cfg["cross_attention_dim"]), generator=g).half().to(dev)
        gl = dict(boxes=boxes, masks=masks, positive_embeddings=pe)
    eng = UNetEngine(cfg, sd16, device=dev)
    ref = eng.forward(x, 300, ehs, gligen=gl).float().clone()
    worst = 0.0
    bad = 0
    junk = []
    for i in range(iters):
        # churn the allocator so that scratch buffers land on memory with different (non-zero) contents
        junk.append(torch.full((1 << 20,), float("nan"), device=dev, dtype=torch.float16))
        if len(junk) > 3:
            junk.pop(0)
        y = eng.forward(x, 300, ehs, gligen=gl).float()
        d = ((y - ref).abs().max() / ref.abs().max()).item()
        if not (d == d) or d > 1e-4:
            bad += 1
            print(f"  gated={gated} iter {i}: rel diff {d:.3e}", flush=True)
        worst = max(worst, d if d == d else 1e9)
    print(f"gated={gated}: {iters} runs, worst run-to-run rel diff {worst:.3e}, {bad} bad", flush=True)
