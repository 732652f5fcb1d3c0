"""CPU tests (no GPU): the oracle restatement against (a) the golden vectors produced by the reference's own modules
(oracle/make_golden.py) and (b) -- where /root/reference is mounted -- the reference executed live."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import guidance_ref as G
from oracle import ref_loader, weights
from oracle import unet3d_ref as R
from oracle.make_golden import CASES, FakeScheduler, layout

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _key_name(k):
    return "_".join(map(str, k))


def test_param_spec_matches_committed_reference_inventory():
    spec = json.load(open(os.path.join(GOLD, "state_dict_spec.json")))
    for base in ("tiny", "zeroscope"):
        for gated in (False, True):
            cfg = weights.make_cfg(base, gated=gated)
            mine = {k: list(v) for k, v in weights.param_spec(cfg).items()}
            assert mine == spec[f"{base}_{'gated' if gated else 'default'}"]
    z = weights.param_spec(weights.make_cfg("zeroscope"))
    assert len(z) == 1480
    n = sum(int(np.prod(s)) if len(s) else 1 for s in z.values())
    assert abs(n - 1411.23e6) < 0.01e6


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_param_spec_matches_live_reference():
    ns = ref_loader.load()
    for gated in (False, True):
        cfg = weights.make_cfg("zeroscope", gated=gated)
        with torch.device("meta"):
            m = ns.UNet3DConditionModel(sample_size=32, attention_type="gated" if gated else "default")
        ref = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert ref == {k: tuple(v) for k, v in weights.param_spec(cfg).items()}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_forward_matches_golden(name):
    gated, B, F, h, w, seed = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = weights.make_cfg("tiny", gated=gated)
    sd = weights.random_state_dict(cfg, seed=seed)
    gl = None
    if gated:
        gl = dict(boxes=torch.from_numpy(g["gligen_boxes"]), masks=torch.from_numpy(g["gligen_masks"]),
                  positive_embeddings=torch.from_numpy(g["gligen_pe"]))
    saved = {}
    with torch.no_grad():
        y = R.unet_forward(sd, cfg, torch.from_numpy(g["sample"]), int(g["timestep"]), torch.from_numpy(g["ehs"]),
                           save_keys=R.GUIDANCE_KEYS_DEFAULT, saved=saved, gligen=gl)
    ref = torch.from_numpy(g["unet_out"])
    assert (y - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    for k in R.GUIDANCE_KEYS_DEFAULT:
        cols = saved[k][..., [1, 2, 4, 6, 7]]
        assert (cols - torch.from_numpy(g["attn_" + _key_name(k)]).float()).abs().max().item() < 1e-3  # fp16 fixture
        assert (saved[k].sum(dim=(0, 1, 2)) - torch.from_numpy(g["attnsum_" + _key_name(k)])).abs().max().item() < 1e-2


def test_oracle_guidance_matches_golden():
    name = "tiny_default_b1_f8_16x24"
    gated, B, F, h, w, seed = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = weights.make_cfg("tiny")
    sd = weights.random_state_dict(cfg, seed=seed)
    x = torch.from_numpy(g["sample"])
    ehs = torch.from_numpy(g["ehs"])
    for variant, hp in enumerate([
        dict(loss_scale=2.5, fg_top_p=0.25, bg_top_p=0.25, fg_weight=1.0, bg_weight=2.0),
        dict(loss_scale=5.0, fg_top_p=0.75, bg_top_p=0.75, fg_weight=1.0, bg_weight=4.0, com_loss_scale=0.03),
    ]):
        bboxes, positions = layout(F, variant)
        alpha = torch.tensor(float(g[f"guid{variant}_alpha"]))
        lat, loss, grad, _ = G.guidance_step(sd, cfg, x, int(g["timestep"]), ehs, bboxes, positions, alpha,
                                             guidance_attn_keys=R.GUIDANCE_KEYS_DEFAULT, base_attn_dim=(h, w), **hp)
        assert abs(loss.item() - float(g[f"guid{variant}_loss"])) < 1e-3 * abs(float(g[f"guid{variant}_loss"]))
        ref = torch.from_numpy(g[f"guid{variant}_latents"])
        upd = (ref - x).abs().max().item()
        assert upd > 0
        assert (lat - ref).abs().max().item() < 1e-3 * upd
        assert abs(FakeScheduler().alphas_cumprod[437].item() - alpha.item()) < 1e-7


def test_scale_proportion_bankers_rounding_and_hw():
    # utils/utils.py:82-103 uses Python round() (banker's) separately on origin and size
    assert G.scale_proportion((0.25, 0.25, 0.75, 0.75), 10, 18) == (4, 2, 13, 7)
    assert G.scale_proportion((0.0, 0.0, 0.0, 0.0), 20, 36) == (0, 0, 0, 0)
    assert G.scale_proportion((0.9, 0.9, 1.2, 1.3), 20, 36) == (32, 18, 36, 20)
    assert G.get_hw_from_attn_dim(720, (40, 72)) == (20, 36)
    assert G.get_hw_from_attn_dim(180, (40, 72)) == (10, 18)
    assert G.get_hw_from_attn_dim(45, (40, 72)) == (5, 9)
    assert G.get_hw_from_attn_dim(2880, (40, 72)) == (40, 72)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_oracle_matches_live_reference_forward():
    import warnings
    warnings.filterwarnings("ignore")
    ns = ref_loader.load()
    cfg = weights.make_cfg("tiny")
    sd = weights.random_state_dict(cfg, seed=5)
    m = ns.UNet3DConditionModel(sample_size=16, block_out_channels=cfg["block_out_channels"],
                                cross_attention_dim=cfg["cross_attention_dim"]).eval()
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(6)
    x = torch.randn((1, 4, 3, 8, 8), generator=g)
    ehs = torch.randn((1, 77, 128), generator=g)
    with torch.no_grad():
        ref = m(x, 10, encoder_hidden_states=ehs, cross_attention_kwargs={}).sample
        y = R.unet_forward(sd, cfg, x, 10, ehs)
    assert (y - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
