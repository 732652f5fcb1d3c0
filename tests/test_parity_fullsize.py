"""GPU parity at the BASELINE.json shapes (not the tiny configs): the full 1.41 B / 1.62 B-parameter UNets at
  cfg 2  lvd_zeroscope          24 frames 40x72   forward (one CFG branch) + one guidance iteration (loss, dE/dlatent)
  cfg 3  lvd_gligen_zeroscope   24 frames 40x72   forward with the GLIGEN fusers + one guidance iteration (guided variant)
  cfg 4  lvd_plus_modelscope256 16 frames 32x32   forward with fusers (CFG batch 2) + one guidance iteration (benchmark
                                                  hyper-parameters incl. the centre-of-mass terms, 4-object layout)
against the oracle restatement run on the same GPU in fp32 (TF32 off) AND in fp16 (what the reference computes on a
GPU: the comparison SURVEY.md section 8(c) names).  Every comparison prints max-normalised, RMS-relative and
median / p99.9 / max element-wise errors (conftest.err_stats), so the north_star "1e-3 relative fp16" can be judged.

What the numbers say (B200, profiles/r2b_parity_fullsize.txt).  fp16 evaluation of this ~250-layer graph is itself only
reproducible to ~2e-3: the reference's own graph in fp16 (torch) differs from its fp32 evaluation by ||d||/||ref|| =
2.2-2.4e-3, the engine by 1.9-2.1e-3, and the two fp16 runs differ from each other by 2.8-3.1e-3 (independent rounding
noise adds in quadrature).  north_star's "1e-3 relative fp16" is therefore below the noise floor of ANY fp16 run of the
reference, including the reference's; the gate that can be held is "at least as close to the exact (fp32) result as the
reference's own fp16 run":
  * forward:  ||engine - fp32|| / ||fp32|| <= 2.5e-3 (measured 1.9-2.1e-3) and <= 1.05 x the same figure of the
              torch-fp16 run; max|d|/max|ref| <= 2.5e-3 (measured 1.9-2.0e-3) and <= 1.05 x torch-fp16's
  * guidance: |loss - loss_fp32| <= 1e-4 |loss_fp32| (measured <= 2e-6); dE/dlatent with the benchmark
              hyper-parameters: cosine >= 0.9999 and ||g - g_fp32|| <= 1 % ||g_fp32|| (measured 0.4-0.6 %); with the weak
              top-25 % selection -- a discontinuous energy whose selected pixels flip under fp16 noise --
              <= 6 % / cosine >= 0.998 (measured 2.7-3.9 %); and in every case never worse than the torch-fp16 autograd
              run of the oracle (measured 1.3-5.0 %).
"""
import pytest
import torch

from conftest import err_stats, fmt_stats

pytestmark = pytest.mark.gpu

KEYS = [("down", 1, 0, 0), ("down", 2, 0, 0), ("down", 2, 1, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 2, 2, 0)]
HP_WEAK = dict(loss_scale=2.5, fg_top_p=0.25, bg_top_p=0.25, fg_weight=1.0, bg_weight=2.0)            # README.md:68
HP_BENCH = dict(loss_scale=5.0, fg_top_p=0.75, bg_top_p=0.75, fg_weight=1.0, bg_weight=4.0,
                com_loss_scale=0.03)                                                    # README.md:140, lvd.py:92-104


def bear(frames):
    return [[(0.83 * f / (frames - 1) * 0.8, 0.5, 0.83 * f / (frames - 1) * 0.8 + 0.17, 0.695) for f in range(frames)]], [[2]]


def four_objects(frames):
    """Layout (B) of SURVEY.md 8(d): four objects, one absent in half of the frames, one with a two-token phrase."""
    b1 = [(0.05, 0.1, 0.40, 0.55)] * frames
    b2 = [(0.0, 0.0, 0.0, 0.0) if f % 2 else (0.55, 0.35, 0.95, 0.9) for f in range(frames)]
    b3 = [(0.2 + 0.015 * f, 0.55, 0.55 + 0.015 * f, 1.0) for f in range(frames)]
    b4 = [(0.6, 0.02 + 0.01 * f, 0.9, 0.3 + 0.01 * f) for f in range(frames)]
    return [b1, b2, b3, b4], [[1], [4], [6, 7], [9]]


class _Model:
    def __init__(self, gated, dev):
        from lvd_b200 import synthetic
        from lvd_b200.engine import UNetEngine
        self.cfg = synthetic.make_cfg("zeroscope", gated=gated)
        sd16 = synthetic.random_state_dict(self.cfg, seed=0, dtype=torch.float16)
        self.sd16 = {k: v.to(dev) for k, v in sd16.items()}
        self.eng = UNetEngine(self.cfg, sd16, device=dev)
        self._sd32 = None

    @property
    def sd32(self):
        if self._sd32 is None:
            self._sd32 = {k: v.float() for k, v in self.sd16.items()}
        return self._sd32


@pytest.fixture(scope="module")
def zeroscope(cuda):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = _Model(False, cuda)
    yield m
    del m
    torch.cuda.empty_cache()


@pytest.fixture(scope="module")
def gated(cuda):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = _Model(True, cuda)
    yield m
    del m
    torch.cuda.empty_cache()


def _inputs(b, f, h, w, dev, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((b, 4, f, h, w), generator=g).half().to(dev)
    ehs = torch.randn((b, 77, 1024), generator=g).half().to(dev)
    return x, ehs


def _gligen(n_img, dev, seed):
    g = torch.Generator().manual_seed(seed)
    boxes = torch.rand((n_img, 30, 4), generator=g).half().to(dev)
    masks = (torch.rand((n_img, 30), generator=g) > 0.8).half().to(dev)
    pe = torch.randn((n_img, 30, 1024), generator=g).half().to(dev)
    return dict(boxes=boxes, masks=masks, positive_embeddings=pe)


def _forward_case(m, tag, b, f, h, w, t, dev, gligen=None, seed=11):
    from oracle import unet3d_ref as R
    x16, e16 = _inputs(b, f, h, w, dev, seed)
    y = m.eng.forward(x16, t, e16, gligen=gligen)
    torch.cuda.synchronize()
    with torch.no_grad():
        y16 = R.unet_forward(m.sd16, m.cfg, x16, t, e16, gligen=gligen)
        gl32 = None if gligen is None else {k: v.float() for k, v in gligen.items()}
        y32 = R.unet_forward(m.sd32, m.cfg, x16.float(), t, e16.float(), gligen=gl32)
    s_eng, s_t16, s_x = err_stats(y, y32), err_stats(y16, y32), err_stats(y, y16)
    print(f"\n[{tag} forward b={b} f={f} {h}x{w}]"
          f"\n   engine fp16 vs oracle fp32      : {fmt_stats(s_eng)}"
          f"\n   torch fp16 oracle vs oracle fp32: {fmt_stats(s_t16)}"
          f"\n   engine fp16 vs torch fp16 oracle: {fmt_stats(s_x)}")
    assert y.shape == (b, 4, f, h, w) and torch.isfinite(y.float()).all()
    assert s_eng["rms_rel"] <= 2.5e-3 and s_eng["rms_rel"] <= 1.05 * s_t16["rms_rel"]
    assert s_eng["max_norm"] <= 2.5e-3 and s_eng["max_norm"] <= 1.05 * s_t16["max_norm"]
    return s_eng, s_t16, s_x


def _guidance_case(m, tag, f, h, w, t, dev, bboxes, positions, hp, seed=12, max_rel=0.01, min_cos=0.9999):
    from lvd_b200.guidance_engine import GRAD_SCALE, GuidanceGraph
    from oracle import guidance_ref as G
    hp = dict(hp)
    ls = hp.pop("loss_scale")
    x16, e16 = _inputs(1, f, h, w, dev, seed)
    loss, grad = GuidanceGraph(m.eng).run(x16, t, e16, bboxes, positions, KEYS, loss_scale=ls, base_attn_dim=(h, w),
                                          **hp)
    grad = grad / GRAD_SCALE
    torch.cuda.synchronize()
    a = torch.tensor(0.5)
    _, l16, g16, _ = G.guidance_step(m.sd16, m.cfg, x16, t, e16, bboxes, positions, a, guidance_attn_keys=KEYS,
                                     loss_scale=ls, base_attn_dim=(h, w), **hp)
    torch.cuda.empty_cache()
    _, l32, g32, _ = G.guidance_step(m.sd32, m.cfg, x16.float(), t, e16.float(), bboxes, positions, a,
                                     guidance_attn_keys=KEYS, loss_scale=ls, base_attn_dim=(h, w), **hp)
    torch.cuda.empty_cache()

    def cos(u, v):
        u, v = u.float().flatten(), v.float().flatten()
        return (torch.dot(u, v) / (u.norm() * v.norm())).item()

    s_eng, s_t16 = err_stats(grad, g32), err_stats(g16, g32)
    c_eng, c_t16 = cos(grad, g32), cos(g16, g32)
    print(f"\n[{tag} guidance f={f} {h}x{w}, {len(bboxes)} objects] loss engine {loss.item():.5f}  oracle fp32 "
          f"{l32.item():.5f}  oracle fp16 {l16.item():.5f}"
          f"\n   dE/dlatent engine vs oracle fp32         : cos {c_eng:.6f}  {fmt_stats(s_eng)}"
          f"\n   dE/dlatent torch fp16 autograd vs fp32   : cos {c_t16:.6f}  {fmt_stats(s_t16)}")
    assert torch.isfinite(grad).all() and grad.abs().max() > 0
    assert abs(loss.item() - l32.item()) <= 1e-4 * abs(l32.item())
    assert c_eng >= min_cos and s_eng["rms_rel"] <= max_rel
    assert s_eng["rms_rel"] <= s_t16["rms_rel"] and c_eng >= c_t16


def test_cfg2_zeroscope_forward(zeroscope, cuda):
    _forward_case(zeroscope, "cfg2 zeroscope", 1, 24, 40, 72, 951, cuda)


def test_cfg2_zeroscope_guidance(zeroscope, cuda):
    b, p = bear(24)
    # top-25 % selection: a discontinuous energy whose selected pixels flip under ANY fp16 noise (the value moves between
    # 2.7 % and 3.9 % with rounding-level changes of the kernels; torch's fp16 autograd is at 5.0 %)
    _guidance_case(zeroscope, "cfg2 zeroscope weak/bear", 24, 40, 72, 951, cuda, b, p, HP_WEAK, max_rel=0.06,
                   min_cos=0.998)
    b, p = four_objects(24)
    _guidance_case(zeroscope, "cfg2 zeroscope benchmark/4 objects", 24, 40, 72, 780, cuda, b, p, HP_BENCH, seed=13)


def test_cfg3_gligen_zeroscope_forward_and_guidance(gated, cuda):
    gl = _gligen(24, cuda, 5)
    _forward_case(gated, "cfg3 zeroscope+gligen", 1, 24, 40, 72, 932, cuda, gligen=gl, seed=21)
    b, p = four_objects(24)
    _guidance_case(gated, "cfg3 zeroscope+gligen benchmark/4 objects", 24, 40, 72, 932, cuda, b, p, HP_BENCH, seed=22)


def test_cfg4_lvd_plus_modelscope256_forward_and_guidance(gated, cuda):
    gl = _gligen(32, cuda, 6)
    _forward_case(gated, "cfg4 modelscope256+gligen", 2, 16, 32, 32, 913, cuda, gligen=gl, seed=31)
    b, p = four_objects(16)
    _guidance_case(gated, "cfg4 modelscope256 benchmark/4 objects", 16, 32, 32, 913, cuda, b, p, HP_BENCH, seed=32)
