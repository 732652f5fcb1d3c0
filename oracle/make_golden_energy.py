"""TEST INFRASTRUCTURE: golden vectors of every loss branch of utils/guidance.py, produced by the reference executed
verbatim (oracle/ref_loader.py): `compute_ca_lossv3` on seeded softmax maps with each option set, and dLoss/dP from
torch.autograd through the reference code.  -> tests/golden/energy_variants.npz

The maps are fp32 tensors holding fp16-representable values, so the CUDA kernel (fp16 maps) and the reference (fp32 CPU)
see identical probabilities.  Run:  python -m oracle.make_golden_energy   (needs /root/reference).
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FRAMES, BASE = 6, (8, 12)
KEYS = [("down", 1, 0, 0), ("up", 1, 1, 0)]
SHAPES = {KEYS[0]: (3, 96), KEYS[1]: (2, 24)}  # key -> (heads, hw): 8x12 and 4x6 maps
N_TEXT = 16
POSITIONS = [[2], [5, 6], [9]]
BASE_KW = dict(fg_top_p=0.4, bg_top_p=0.3, fg_weight=1.0, bg_weight=2.0, base_attn_dim=BASE)

VARIANTS = {
    "max": {},
    "max_com": dict(com_loss_scale=0.05),
    "ratio": dict(use_ratio_based_loss=True),
    "ratio_boxdiff_com": dict(use_ratio_based_loss=True, boxdiff_loss_scale=0.7, com_loss_scale=0.03),
    "ce": dict(use_max_based_loss=False, use_ce_based_loss=True),
    "boxdiff_normed": dict(boxdiff_loss_scale=1.5, boxdiff_normed=True),
    "boxdiff_sum_L2": dict(boxdiff_loss_scale=0.5, boxdiff_normed=False, boxdiff_L=2),
    "sync": dict(attn_sync_weight=3.0),
    "sync_all": dict(attn_sync_weight=2.0, boxdiff_loss_scale=1.0, com_loss_scale=0.02),
    "smooth": dict(smooth_attn=True),
    "smooth_sigma": dict(smooth_attn=True, sigma=1.0, com_loss_scale=0.02),
    "renorm": dict(attn_renorm=True, num_tokens=12, renorm_scale=2.0),
    # upsample_scale != 1 only runs with the ratio loss in the reference: the top-k branches multiply the un-flattened
    # (heads, 1, H, W) interpolation result with the flattened mask and raise a shape error (guidance.py:297-347)
    "upsample2_ratio": dict(upsample_scale=2, upsample_mode="bilinear", use_ratio_based_loss=True),
    "upsample2_ratio_com": dict(upsample_scale=2, upsample_mode="bilinear", use_ratio_based_loss=True,
                                com_loss_scale=0.04, boxdiff_loss_scale=0.5),
}
# layouts: "present" = every object has a box in every frame (attn_sync needs that: an empty box makes the reference's
# mean() over an empty crop NaN); "absent" = the second object disappears in odd frames
LAYOUTS = {
    "present": [[(0.05 + 0.1 * f, 0.1, 0.45 + 0.1 * f, 0.6) for f in range(FRAMES)],
                [(0.5, 0.3 + 0.05 * f, 0.95, 0.8 + 0.02 * f) for f in range(FRAMES)],
                [(0.2, 0.55, 0.7, 1.0)] * FRAMES],
    "absent": [[(0.05 + 0.1 * f, 0.1, 0.45 + 0.1 * f, 0.6) for f in range(FRAMES)],
               [(0.0, 0.0, 0.0, 0.0) if f % 2 else (0.5, 0.3, 0.95, 0.8) for f in range(FRAMES)],
               [(0.2, 0.55, 0.7, 1.0)] * FRAMES],
}


def make_maps(seed):
    g = torch.Generator().manual_seed(seed)
    maps = {}
    for k, (heads, hw) in SHAPES.items():
        logits = 2.5 * torch.randn((FRAMES, heads, hw, N_TEXT), generator=g)
        maps[k] = logits.softmax(-1).half().float()
    return maps


def cases():
    for name, kw in VARIANTS.items():
        for lay in LAYOUTS:
            if "sync" in name and lay == "absent":
                continue
            yield f"{name}__{lay}", dict(BASE_KW, **kw), LAYOUTS[lay]


def main():
    warnings.filterwarnings("ignore")
    ns = ref_loader.load()
    out = {}
    for i, (name, kw, bboxes) in enumerate(cases()):
        maps = {k: v.clone().requires_grad_(True) for k, v in make_maps(100 + i).items()}
        loss = ns.guidance.compute_ca_lossv3(maps, bboxes, POSITIONS, KEYS, **kw)
        grads = torch.autograd.grad(loss, [maps[k] for k in KEYS])
        out[name + "__loss"] = np.float64(loss.item())
        for k, g in zip(KEYS, grads):
            out[name + "__grad_" + "_".join(map(str, k))] = g.numpy().astype(np.float32)
        print(f"{name:32s} loss {loss.item():.6f}")
    # the NaN the reference produces for attn_sync with an empty box
    maps = make_maps(999)
    nan_loss = ns.guidance.compute_ca_lossv3(maps, LAYOUTS["absent"], POSITIONS, KEYS,
                                             **dict(BASE_KW, attn_sync_weight=1.0))
    out["sync__absent__loss"] = np.float64(nan_loss.item())
    print("sync with an empty box ->", nan_loss.item())
    np.savez_compressed(os.path.join(GOLD, "energy_variants.npz"), **out)


if __name__ == "__main__":
    main()
