"""TEST INFRASTRUCTURE: generate tests/golden/*.npz by running the REFERENCE's own modules (imported verbatim from
/root/reference through oracle/shim) on the deterministic synthetic weights of oracle/weights.py.

Run in the builder container only:   python -m oracle.make_golden
The fixtures pin (a) the oracle restatement (tests/test_oracle_pin.py, CPU) and (b) the CUDA engine
(tests/test_golden_gpu.py, GPU box -- /root/reference does not exist there).
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, weights  # noqa: E402
from oracle.unet3d_ref import GUIDANCE_KEYS_DEFAULT  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (gated, B, F, h, w, seed)
    "tiny_default_b1_f8_16x24": (False, 1, 8, 16, 24, 11),
    "tiny_default_b2_f6_8x16": (False, 2, 6, 8, 16, 12),
    "tiny_gated_b2_f4_16x16": (True, 2, 4, 16, 16, 13),
}


class FakeScheduler:
    """Only what latent_backward_guidance touches (models/pipelines.py:76,124-132)."""

    def __init__(self):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    def scale_model_input(self, x, t):
        return x


def layout(F, variant):
    if variant == 0:  # one object sweeping left to right (the README "bear" demo shape), token 2
        return [[(0.05 + 0.5 * f / max(F - 1, 1), 0.3, 0.4 + 0.5 * f / max(F - 1, 1), 0.8) for f in range(F)]], [[2]]
    # three objects; the second is absent in odd frames (empty box), the third has two tokens
    b1 = [(0.1, 0.1, 0.45, 0.6)] * F
    b2 = [(0.0, 0.0, 0.0, 0.0) if f % 2 else (0.55, 0.4, 0.95, 0.9) for f in range(F)]
    b3 = [(0.3 + 0.02 * f, 0.5, 0.7 + 0.02 * f, 1.0) for f in range(F)]
    return [b1, b2, b3], [[1], [4], [6, 7]]


def main():
    warnings.filterwarnings("ignore")
    ns = ref_loader.load()
    os.makedirs(GOLD, exist_ok=True)
    spec_dump = {}
    for gated in (False, True):
        for base in ("tiny", "zeroscope"):
            cfg = weights.make_cfg(base, gated=gated)
            spec_dump[f"{base}_{'gated' if gated else 'default'}"] = {k: list(v) for k, v in
                                                                      weights.param_spec(cfg).items()}
    with open(os.path.join(GOLD, "state_dict_spec.json"), "w") as f:
        json.dump(spec_dump, f)
    for name, (gated, B, F, h, w, seed) in CASES.items():
        cfg = weights.make_cfg("tiny", gated=gated)
        sd = weights.random_state_dict(cfg, seed=seed)
        m = ns.UNet3DConditionModel(sample_size=16, block_out_channels=cfg["block_out_channels"],
                                    cross_attention_dim=cfg["cross_attention_dim"],
                                    attention_type="gated" if gated else "default").eval()
        m.load_state_dict(sd)
        g = torch.Generator().manual_seed(seed + 1000)
        x = torch.randn((B, 4, F, h, w), generator=g)
        ehs = torch.randn((B, 77, cfg["cross_attention_dim"]), generator=g)
        out = {"sample": x.numpy(), "ehs": ehs.numpy(), "timestep": np.array(437)}
        kw = {"save_attn_to_dict": {}, "save_keys": list(GUIDANCE_KEYS_DEFAULT)}
        if gated:
            boxes = torch.rand((B * F, 30, 4), generator=g)
            masks = (torch.rand((B * F, 30), generator=g) > 0.6).float()
            pe = torch.randn((B * F, 30, cfg["cross_attention_dim"]), generator=g)
            kw["gligen"] = dict(boxes=boxes, masks=masks, positive_embeddings=pe)
            out.update(gligen_boxes=boxes.numpy(), gligen_masks=masks.numpy(), gligen_pe=pe.numpy())
        with torch.no_grad():
            y = m(x, 437, encoder_hidden_states=ehs, cross_attention_kwargs=kw).sample
        out["unet_out"] = y.numpy()
        for k, v in kw["save_attn_to_dict"].items():
            # keep fixtures small: store the map's object-token columns only (what the energy reads) + a checksum
            out["attn_" + "_".join(map(str, k))] = v[..., [1, 2, 4, 6, 7]].numpy().astype(np.float16)
            out["attnsum_" + "_".join(map(str, k))] = v.sum(dim=(0, 1, 2)).numpy()
        # guidance iterations (B must be 1 on that path: models/pipelines.py:78-83 feeds cond embeddings only)
        if B == 1 and not gated:
            sched = FakeScheduler()
            for variant, hp in enumerate([
                dict(loss_scale=2.5, fg_top_p=0.25, bg_top_p=0.25, fg_weight=1.0, bg_weight=2.0),
                dict(loss_scale=5.0, fg_top_p=0.75, bg_top_p=0.75, fg_weight=1.0, bg_weight=4.0, com_loss_scale=0.03),
            ]):
                bboxes, positions = layout(F, variant)
                hp = dict(hp)
                loss_scale = hp.pop("loss_scale")
                lat, loss = ns.latent_backward_guidance(
                    sched, m, ehs, 0, bboxes, positions, 437, x.clone(), torch.tensor(10000.0), loss_scale=loss_scale,
                    loss_threshold=0.0, max_iter=1, max_index_step=10, guidance_attn_keys=list(GUIDANCE_KEYS_DEFAULT),
                    base_attn_dim=(h, w), **hp)
                out[f"guid{variant}_latents"] = lat.detach().numpy()
                out[f"guid{variant}_loss"] = np.array(loss.item())
                out[f"guid{variant}_alpha"] = np.array(sched.alphas_cumprod[437].item())
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", None) for k, v in out.items() if k.startswith(("unet", "guid"))})


if __name__ == "__main__":
    main()
