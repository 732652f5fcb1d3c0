"""Per-op run-to-run determinism on fixed inputs (NaN-poisoned allocator churn between runs): a difference that is not
explained by GroupNorm's shared-memory atomics points at a race or an uninitialised read."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lvd_b200 import ops, repack  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def churn():
    t = [torch.full((1 << 22,), float("nan"), device=dev, dtype=torch.float16) for _ in range(4)]
    del t


def check(name, fn, iters=20):
    ref = fn().float().clone()
    worst, bad = 0.0, 0
    for _ in range(iters):
        churn()
        y = fn().float()
        d = (y - ref).abs().max().item()
        nan = bool(torch.isnan(y).any().item())
        if d != 0.0 or nan:
            bad += 1
        worst = max(worst, d)
    print(f"{name:60s} worst abs diff {worst:.3e}  nondeterministic runs {bad}/{iters}", flush=True)


def r(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).half()


for pair in (0, 1):
    ops.FORCE_CTA_PAIR = pair
    tag = "auto" if pair == 0 else "single"
    for (m, n, k, res) in [(2048, 64, 64, 0), (2048, 128, 128, 1), (2048, 256, 128, 1), (2048, 320, 320, 1),
                           (34560, 640, 640, 1), (8640, 1280, 1280, 1), (300, 192, 64, 0), (154, 640, 1024, 0)]:
        x, w, b = r(m, k), r(n, k, scale=k ** -0.5), r(n)
        rs = r(m, n) if res else None
        check(f"[{tag}] linear M={m} N={n} K={k} res={res}", lambda: ops.linear(x, w, b, residual=rs))
    for (m, c) in [(2048, 64), (8640, 1280)]:
        x, w, b = r(m, c), r(8 * c, c, scale=c ** -0.5), r(8 * c)
        wi, bi = repack.geglu_interleave(w, b)
        check(f"[{tag}] geglu M={m} C={c}", lambda: ops.linear(x, wi, bi, act=ops.ACT_GEGLU))
    for (n_img, h, w_, cin, cout) in [(8, 16, 16, 64, 64), (48, 10, 18, 1280, 1280), (8, 8, 8, 128, 256)]:
        x, wk, b = r(n_img * h * w_, cin), r(cout, 9 * cin, scale=(9 * cin) ** -0.5), r(cout)
        rs, tv = r(n_img * h * w_, cout), r(n_img, cout)
        check(f"[{tag}] conv3x3 {n_img}x{h}x{w_} {cin}->{cout} +temb+res",
              lambda: ops.conv2d_3x3(x, n_img, h, w_, wk, b, rowvec=tv, residual=rs))
    for (b_, f, hw, c) in [(2, 4, 256, 64), (2, 24, 180, 1280)]:
        x, wk = r(b_ * f * hw, c), r(c, 3 * c, scale=(3 * c) ** -0.5)
        rs = r(b_ * f * hw, c)
        check(f"[{tag}] tconv B={b_} F={f} hw={hw} C={c}", lambda: ops.conv_temporal(x, b_, f, hw, wk, residual=rs))
ops.FORCE_CTA_PAIR = 0
for (z, s, heads) in [(8, 256, 1), (8, 286, 2), (48, 720, 10), (8, 64, 4)]:
    c = heads * 64
    qkv = r(z * s, 3 * c)
    check(f"self-attn z={z} S={s} heads={heads}",
          lambda: ops.attention(qkv, qkv, heads=heads, nz=z, s_q=s, s_k=s, k_col0=c, v_col0=2 * c))
for (z, s, heads) in [(8, 256, 2), (48, 720, 10)]:
    c = heads * 64
    q, kv = r(z * s, c), r(2 * 77, 2 * c)
    check(f"cross-attn z={z} S={s} heads={heads}",
          lambda: ops.attention(q, kv, heads=heads, nz=z, s_q=s, s_k=77, v_col0=c, kv_z_div=z // 2))
for (b_, f, hw, heads) in [(2, 8, 256, 1), (2, 24, 180, 20)]:
    c = heads * 64
    qkv = r(b_ * f * hw, 3 * c)
    check(f"temporal-attn packed B={b_} F={f} hw={hw} heads={heads}",
          lambda: ops.attention(qkv, qkv, heads=heads, nz=b_, temporal=True, packed=True, frames=f, hw=hw, k_col0=c,
                                v_col0=2 * c))
for (samples, rows, c) in [(8, 256, 64), (48, 720, 640), (2, 24 * 180, 1280)]:
    x, ga, be = r(samples * rows, c), r(c), r(c)
    check(f"groupnorm samples={samples} rows={rows} C={c}", lambda: ops.groupnorm(x, samples, ga, be, 1e-5, silu=True))
    check(f"layernorm rows={samples * rows} C={c}", lambda: ops.layernorm(x, ga, be))
