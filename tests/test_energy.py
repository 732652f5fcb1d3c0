"""GPU parity of the fused box-energy kernel (loss and dE/dP) against the oracle restatement of
utils/guidance.py (oracle/guidance_ref.py, autograd for the gradient), fp32 math on fp16 maps."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _layouts(F):
    one = ([[(0.05 + 0.5 * f / max(F - 1, 1), 0.3, 0.4 + 0.5 * f / max(F - 1, 1), 0.8) for f in range(F)]], [[2]])
    b1 = [(0.1, 0.1, 0.45, 0.6)] * F
    b2 = [(0.0, 0.0, 0.0, 0.0) if f % 2 else (0.55, 0.4, 0.95, 0.9) for f in range(F)]
    b3 = [(0.3 + 0.02 * f, 0.5, 0.7 + 0.02 * f, 1.0) for f in range(F)]
    return [one, ([b1, b2, b3], [[1], [4], [6, 7]])]


@pytest.mark.parametrize("F,heads,H,W", [(6, 4, 10, 18), (24, 10, 20, 36), (4, 2, 40, 72), (3, 5, 5, 9)])
@pytest.mark.parametrize("hp", [dict(fg_top_p=0.25, bg_top_p=0.25, fg_weight=1.0, bg_weight=2.0),
                                dict(fg_top_p=0.75, bg_top_p=0.75, fg_weight=1.0, bg_weight=4.0, com_loss_scale=0.03),
                                dict()])
def test_box_energy_loss_and_grad(cuda, F, heads, H, W, hp):
    from lvd_b200.utils import guidance as G
    from lvd_b200 import ops
    from oracle import guidance_ref as R
    g = torch.Generator().manual_seed(F * 100 + H)
    hw = H * W
    logits = torch.randn((F, heads, hw, 77), generator=g) * 2.0
    p16 = torch.softmax(logits, dim=-1).half().to(cuda)
    key = ("down", 1, 0, 0)
    for bboxes, positions in _layouts(F):
        # oracle: fp32 math on the same fp16-rounded map, autograd for dE/dP
        pr = p16.float().clone().requires_grad_(True)
        ref = R.compute_ca_loss({key: pr}, bboxes, positions, [key], base_attn_dim=(H, W), **hp)
        (gref,) = torch.autograd.grad(ref, [pr])
        loss = G.compute_ca_lossv3({key: p16}, bboxes, positions, [key], base_attn_dim=(H, W), **hp)
        assert abs(loss.item() - ref.item()) <= 2e-4 * max(1.0, abs(ref.item())), (loss.item(), ref.item())
        # gradient through the token-column layout
        cols, boxes, wts = G.build_slots(bboxes, positions, F, H, W, cuda)
        ptok = p16[..., cols.long()].permute(3, 0, 1, 2).contiguous()
        dp = torch.zeros_like(ptok)
        l2 = torch.zeros((), dtype=torch.float32, device=cuda)
        kw = {k: v for k, v in hp.items()}
        ops.box_energy(ptok, l2, boxes, wts, frames=F, heads=heads, H=H, W=W, out_scale=1.0 / len(bboxes), dp=dp,
                       grad_scale=1024.0, **kw)
        assert abs(l2.item() - ref.item()) <= 2e-4 * max(1.0, abs(ref.item()))
        gsum = torch.zeros_like(gref)
        for s in range(cols.numel()):
            gsum[..., int(cols[s])] += dp[s].float() / 1024.0
        scale = gref.abs().max().item()
        assert scale > 0
        assert (gsum - gref).abs().max().item() <= 2e-3 * scale
