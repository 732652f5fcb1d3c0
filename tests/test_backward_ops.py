"""GPU parity of the hand-written input-gradient kernels against torch.autograd on a plain PyTorch fp32 model of the
same op (inputs are the fp16-rounded values; gradients come back as fp16).  Tolerance 3e-3 of max|ref grad|."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 3e-3


def _rand(shape, dev, scale=1.0, seed=0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).to(dev).half()


def _grad(fn, x, dy):
    xr = x.float().clone().requires_grad_(True)
    y = fn(xr)
    (g,) = torch.autograd.grad(y, [xr], grad_outputs=dy.float())
    return g


@pytest.mark.parametrize("rows,n,k", [(300, 320, 640), (1000, 2560, 320), (77, 64, 1024)])
def test_linear_dgrad(cuda, rows, n, k):
    from lvd_b200 import ops
    w = _rand((n, k), cuda, k ** -0.5, seed=1)
    dy = _rand((rows, n), cuda, seed=2)
    res = _rand((rows, k), cuda, seed=3)
    out = ops.linear_dgrad(dy, w, residual=res)
    ref = (dy.float() @ w.float()).half().float() + res.float()
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("n_img,h,w,cin,cout", [(2, 8, 8, 64, 64), (3, 10, 18, 128, 320), (4, 5, 9, 320, 64)])
def test_conv3x3_dgrad(cuda, n_img, h, w, cin, cout):
    from lvd_b200 import ops, repack
    wt = _rand((cout, cin, 3, 3), cuda, (9 * cin) ** -0.5, seed=4)
    x = _rand((n_img, cin, h, w), cuda, seed=5)
    dy = _rand((n_img, cout, h, w), cuda, seed=6)
    ref = _grad(lambda t: F.conv2d(t, wt.float(), padding=1), x, dy)
    dyl = dy.permute(0, 2, 3, 1).reshape(-1, cout).contiguous()
    out = ops.conv2d_3x3_dgrad(dyl, n_img, h, w, repack.conv_weight_krsc(wt))
    assert rel_err(out, ref.permute(0, 2, 3, 1).reshape(-1, cin)) < TOL


@pytest.mark.parametrize("n_img,h,w,c", [(2, 8, 8, 64), (3, 10, 18, 128)])
def test_conv3x3_s2_dgrad(cuda, n_img, h, w, c):
    from lvd_b200 import ops, repack
    wt = _rand((c, c, 3, 3), cuda, (9 * c) ** -0.5, seed=7)
    x = _rand((n_img, c, h, w), cuda, seed=8)
    dy = _rand((n_img, c, h // 2, w // 2), cuda, seed=9)
    ref = _grad(lambda t: F.conv2d(t, wt.float(), stride=2, padding=1), x, dy)
    dyl = dy.permute(0, 2, 3, 1).reshape(-1, c).contiguous()
    out = ops.conv2d_3x3_s2_dgrad(dyl, n_img, h, w, repack.s2_dgrad_weights(repack.conv_weight_krsc(wt), c))
    assert rel_err(out, ref.permute(0, 2, 3, 1).reshape(-1, c)) < TOL


@pytest.mark.parametrize("batch,frames,hw,c", [(1, 8, 16, 64), (2, 24, 45, 128)])
def test_conv_temporal_dgrad(cuda, batch, frames, hw, c):
    from lvd_b200 import ops, repack
    wt = _rand((c, c, 3, 1, 1), cuda, (3 * c) ** -0.5, seed=10)
    x = _rand((batch, c, frames, hw, 1), cuda, seed=11)
    dy = _rand((batch, c, frames, hw, 1), cuda, seed=12)
    ref = _grad(lambda t: F.conv3d(t, wt.float(), padding=(1, 0, 0)), x, dy)
    dyl = dy[..., 0].permute(0, 2, 3, 1).reshape(-1, c).contiguous()
    out = ops.conv_temporal_dgrad(dyl, batch, frames, hw, repack.tconv_weight(wt))
    assert rel_err(out, ref[..., 0].permute(0, 2, 3, 1).reshape(-1, c)) < TOL


@pytest.mark.parametrize("samples,rows,c,silu", [(4, 45, 1280, True), (6, 180, 320, False), (2, 24 * 45, 640, True)])
def test_groupnorm_bwd(cuda, samples, rows, c, silu):
    from lvd_b200 import ops
    x = _rand((samples * rows, c), cuda, seed=13, shift=0.5)
    ga = _rand((c,), cuda, seed=14, shift=1.0)
    be = _rand((c,), cuda, seed=15)
    dz = _rand((samples * rows, c), cuda, seed=16)
    acc0 = _rand((samples * rows, c), cuda, seed=17)

    def fn(t):
        y = F.group_norm(t.reshape(samples, rows, c).permute(0, 2, 1), 32, ga.float(), be.float(), 1e-5)
        y = F.silu(y) if silu else y
        return y.permute(0, 2, 1).reshape(samples * rows, c)

    ref = _grad(fn, x, dz)
    out = ops.groupnorm_bwd(x, dz, samples, ga, be, 1e-5, silu=silu)
    assert rel_err(out, ref) < TOL
    acc = acc0.clone()
    ops.groupnorm_bwd(x, dz, samples, ga, be, 1e-5, silu=silu, out=acc, accumulate=True)
    assert rel_err(acc, ref + acc0.float()) < TOL


@pytest.mark.parametrize("rows,c", [(1000, 320), (77, 1280), (50, 512)])
def test_layernorm_bwd(cuda, rows, c):
    from lvd_b200 import ops
    x = _rand((rows, c), cuda, seed=18, shift=0.3)
    ga = _rand((c,), cuda, seed=19, shift=1.0)
    be = _rand((c,), cuda, seed=20)
    dy = _rand((rows, c), cuda, seed=21)
    ref = _grad(lambda t: F.layer_norm(t, (c,), ga.float(), be.float(), 1e-5), x, dy)
    assert rel_err(ops.layernorm_bwd(x, dy, ga), ref) < TOL


def test_geglu_fwd_bwd(cuda):
    from lvd_b200 import ops, repack
    rows, c = 200, 128
    pre_ref = _rand((rows, 8 * c), cuda, seed=22)                    # reference order: [value | gate]
    n = 4 * c
    hf = ops.GEGLU_HALF
    wv = pre_ref[:, :n].reshape(rows, n // hf, hf)
    wg = pre_ref[:, n:].reshape(rows, n // hf, hf)
    pre = torch.cat([wv, wg], dim=2).reshape(rows, 2 * n).contiguous()  # interleaved like repack.geglu_interleave
    dout = _rand((rows, n), cuda, seed=23)
    out = ops.geglu(pre)
    ref = pre_ref[:, :n].float() * F.gelu(pre_ref[:, n:].float())
    assert rel_err(out, ref) < TOL
    g = _grad(lambda t: t[:, :n] * F.gelu(t[:, n:]), pre_ref, dout)
    dpre = ops.geglu_bwd(pre, dout)
    gv = dpre.reshape(rows, n // hf, 2 * hf)[:, :, :hf].reshape(rows, n)
    gg = dpre.reshape(rows, n // hf, 2 * hf)[:, :, hf:].reshape(rows, n)
    assert rel_err(gv, g[:, :n]) < TOL and rel_err(gg, g[:, n:]) < TOL


def test_small_adjoints(cuda):
    from lvd_b200 import ops
    n, h, w, c = 3, 5, 9, 64
    dy = _rand((n * 4 * h * w, c), cuda, seed=24)
    x = _rand((n * h * w, c), cuda, seed=25)
    ref = _grad(lambda t: F.interpolate(t.reshape(n, h, w, c).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
                .permute(0, 2, 3, 1).reshape(-1, c), x, dy)
    assert rel_err(ops.upsample2x_bwd(dy, n, h, w), ref) < TOL
    a, b = _rand((100, 64), cuda, seed=26), _rand((100, 64), cuda, seed=27)
    assert rel_err(ops.add_rows(a, b), a.float() + b.float()) < TOL
    # conv_in adjoint
    bb, f = 2, 3
    wi = _rand((320, 4, 3, 3), cuda, 1 / 6, seed=28)
    lat = _rand((bb, 4, f, h, w), cuda, seed=29)
    drows = _rand((bb * f * h * w, 320), cuda, seed=30)

    def fn(t):
        y = F.conv2d(t.permute(0, 2, 1, 3, 4).reshape(bb * f, 4, h, w), wi.float(), padding=1)
        return y.permute(0, 2, 3, 1).reshape(-1, 320)

    ref = _grad(fn, lat, drows)
    out = ops.conv_in_bwd(drows, wi, bb, f, h, w)
    assert out.dtype == torch.float32 and rel_err(out, ref) < TOL


def _attn(q, k, v, heads):
    z, sq, c = q.shape
    qh = q.view(z, sq, heads, 64).transpose(1, 2)
    kh = k.view(z, -1, heads, 64).transpose(1, 2)
    vh = v.view(z, -1, heads, 64).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(z, sq, c), p


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("n_seq,frames,heads", [(20, 24, 2), (7, 16, 1), (5, 8, 3), (3, 32, 1), (2880, 24, 1)])
def test_temporal_attention_bwd(cuda, n_seq, frames, heads, tc):
    from lvd_b200 import ops
    c = heads * 64
    qkv = _rand((n_seq * frames, 3 * c), cuda, seed=31)
    d_o = _rand((n_seq * frames, c), cuda, seed=32)

    def fn(t):
        q, k, v = (t[:, i * c:(i + 1) * c].reshape(n_seq, frames, c) for i in range(3))
        return _attn(q, k, v, heads)[0].reshape(n_seq * frames, c)

    ref = _grad(fn, qkv, d_o)
    out = ops.temporal_attention_bwd(qkv, d_o, heads=heads, n_seq=n_seq, frames=frames, tensor_core=tc)
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("b,frames,s,heads", [(1, 3, 180, 2), (2, 2, 45, 4), (1, 2, 720, 1)])
def test_cross_attention_bwd_with_energy_seed(cuda, b, frames, s, heads, tc):
    from lvd_b200 import ops
    c = heads * 64
    z = b * frames
    q = _rand((z * s, c), cuda, seed=33)
    kv = _rand((b * 77, 2 * c), cuda, seed=34)
    d_o = _rand((z * s, c), cuda, seed=35)
    cols = (2, 5, 40)
    dp_tok = _rand((len(cols), z, heads, s), cuda, seed=36)
    k = kv[:, :c].reshape(b, 77, c).repeat_interleave(frames, 0).float()
    v = kv[:, c:].reshape(b, 77, c).repeat_interleave(frames, 0).float()
    qr = q.float().clone().requires_grad_(True)
    o, p = _attn(qr.reshape(z, s, c), k, v, heads)
    seed = (o.reshape(z * s, c) * d_o.float()).sum()
    for i, col in enumerate(cols):
        seed = seed + (p[..., col] * dp_tok[i].float()).sum()
    (ref,) = torch.autograd.grad(seed, [qr])
    out = ops.cross_attention_bwd(q, kv, d_o, heads=heads, nz=z, s_q=s, s_k=77, kv_z_div=frames, dp_tok=dp_tok,
                                  slot_cols=cols, tensor_core=tc)
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("z,s,heads", [(2, 128, 1), (2, 180, 2), (1, 720, 2), (3, 45, 1), (1, 300, 1)])
def test_self_attention_bwd(cuda, z, s, heads):
    from lvd_b200 import ops
    c = heads * 64
    qkv = _rand((z * s, 3 * c), cuda, seed=37)
    d_o = _rand((z * s, c), cuda, seed=38)
    lse = torch.zeros((z, heads, s), dtype=torch.float32, device=cuda)
    o = ops.attention(qkv, qkv, heads=heads, nz=z, s_q=s, s_k=s, k_col0=c, v_col0=2 * c, lse=lse)

    def fn(t):
        q, k, v = (t[:, i * c:(i + 1) * c].reshape(z, s, c) for i in range(3))
        return _attn(q, k, v, heads)[0].reshape(z * s, c)

    ref = _grad(fn, qkv, d_o)
    out = ops.self_attention_bwd(qkv, d_o, o, lse, heads=heads, nz=z, s=s)
    for i, name in enumerate("qkv"):
        e = rel_err(out[:, i * c:(i + 1) * c], ref[:, i * c:(i + 1) * c])
        assert e < 5e-3, (name, e)
