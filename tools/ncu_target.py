"""Small target for Nsight Compute: one warm-up + N CFG forwards of the Zeroscope-shape UNet (eager launches).
Usage (under gpurun): ncu ... python tools/ncu_target.py [n_forwards]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from lvd_b200 import synthetic  # noqa: E402
from lvd_b200.engine import UNetEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
cfg = synthetic.make_cfg("zeroscope")
eng = UNetEngine(cfg, synthetic.random_state_dict(cfg, seed=0, dtype=torch.float16), device=dev)
g = torch.Generator().manual_seed(1)
x = torch.randn((2, 4, 24, 40, 72), generator=g).half().to(dev)
ehs = torch.randn((2, 77, 1024), generator=g).half().to(dev)
mode = sys.argv[2] if len(sys.argv) > 2 else "forward"
if mode == "forward":
    for _ in range(1 + n):
        eng.forward(x, 500, ehs)
else:  # one warm-up + n guidance iterations (forward prefix + energy + hand-written backward), B=1
    from bench import GUIDANCE_HP, GUIDANCE_KEYS, GUIDANCE_LOSS_SCALE, bear_layout
    from lvd_b200.guidance_engine import GuidanceGraph
    graph = GuidanceGraph(eng)
    bboxes, positions = bear_layout(24)
    for _ in range(1 + n):
        graph.run(x[:1], 951, ehs[1:2], bboxes, positions, GUIDANCE_KEYS, loss_scale=GUIDANCE_LOSS_SCALE,
                  base_attn_dim=(40, 72), **GUIDANCE_HP)
torch.cuda.synchronize()
print("done")
