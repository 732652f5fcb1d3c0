"""GPU parity of the normalisation / thin-end kernels against PyTorch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _rand(shape, dev, scale=1.0, seed=0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).to(dev).half()


@pytest.mark.parametrize("samples,rows,c,silu", [(4, 45, 1280, True), (48, 180, 320, False), (2, 24 * 45, 640, True),
                                                 (3, 2880, 960, True), (2, 300, 1920, False), (1, 17, 2560, True),
                                                 (2, 64, 512, False)])
def test_groupnorm(cuda, samples, rows, c, silu):
    from lvd_b200 import ops
    x = _rand((samples * rows, c), cuda, seed=1, shift=0.7)
    ga = _rand((c,), cuda, seed=2, shift=1.0)
    be = _rand((c,), cuda, seed=3)
    out = ops.groupnorm(x, samples, ga, be, 1e-5, silu=silu)
    xr = x.float().reshape(samples, rows, c).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, ga.float(), be.float(), 1e-5)
    if silu:
        ref = F.silu(ref.half().float())
    ref = ref.permute(0, 2, 1).reshape(samples * rows, c)
    assert rel_err(out, ref) < TOL


def test_groupnorm_strided_concat_slice(cuda):
    from lvd_b200 import ops
    samples, rows, c1, c2 = 3, 90, 640, 320
    buf = _rand((samples * rows, c1 + c2), cuda, seed=4)
    ga = _rand((c1 + c2,), cuda, seed=5, shift=1.0)
    be = _rand((c1 + c2,), cuda, seed=6)
    out = ops.groupnorm(buf, samples, ga, be, 1e-6)
    ref = F.group_norm(buf.float().reshape(samples, rows, -1).permute(0, 2, 1), 32, ga.float(), be.float(), 1e-6)
    assert rel_err(out, ref.permute(0, 2, 1).reshape(samples * rows, -1)) < TOL


@pytest.mark.parametrize("rows,c", [(1000, 320), (77, 1280), (4321, 640), (50, 512), (9, 1024)])
def test_layernorm(cuda, rows, c):
    from lvd_b200 import ops
    x = _rand((rows, c), cuda, seed=7, shift=0.3)
    ga = _rand((c,), cuda, seed=8, shift=1.0)
    be = _rand((c,), cuda, seed=9)
    out = ops.layernorm(x, ga, be)
    ref = F.layer_norm(x.float(), (c,), ga.float(), be.float(), 1e-5)
    assert rel_err(out, ref) < TOL


def test_conv_in_out(cuda):
    from lvd_b200 import ops, repack
    b, f, h, w = 2, 3, 10, 18
    x = _rand((b, 4, f, h, w), cuda, seed=10)
    wi = _rand((320, 4, 3, 3), cuda, 1 / 6, seed=11)
    bi = _rand((320,), cuda, seed=12)
    rows = ops.conv_in(x, wi, bi)
    x4 = x.permute(0, 2, 1, 3, 4).reshape(b * f, 4, h, w).float()
    ref = F.conv2d(x4, wi.float(), bi.float(), padding=1).permute(0, 2, 3, 1).reshape(b * f * h * w, 320)
    assert rel_err(rows, ref) < TOL
    rows_tc = ops.conv_in_tc(x, repack.conv_in_weight_padded(wi), bi)  # im2col + K = 64 GEMM path
    assert rel_err(rows_tc, ref) < TOL
    wo = _rand((4, 320, 3, 3), cuda, (9 * 320) ** -0.5, seed=13)
    bo = _rand((4,), cuda, seed=14)
    y = ops.conv_out(rows, repack.conv_weight_krsc(wo), bo, b, f, h, w)
    r4 = rows.float().reshape(b * f, h, w, 320).permute(0, 3, 1, 2)
    ref2 = F.conv2d(r4, wo.float(), bo.float(), padding=1).reshape(b, f, 4, h, w).permute(0, 2, 1, 3, 4)
    assert y.shape == (b, 4, f, h, w)
    assert rel_err(y, ref2) < TOL


def test_upsample_timestep_silu(cuda):
    from lvd_b200 import ops
    n, h, w, c = 3, 5, 9, 64
    x = _rand((n * h * w, c), cuda, seed=15)
    y = ops.upsample2x(x, n, h, w)
    ref = F.interpolate(x.float().reshape(n, h, w, c).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1).reshape(-1, c))
    t = torch.tensor([951.0, 20.0, 1.0], device=cuda)
    e = ops.timestep_embed(t, 320)
    import math
    freq = torch.exp(-math.log(10000.0) * torch.arange(160, device=cuda, dtype=torch.float32) / 160)
    a = t[:, None] * freq[None]
    ref = torch.cat([a.cos(), a.sin()], -1)
    assert (e.float() - ref).abs().max().item() < 2e-3
    v = _rand((1000,), cuda, seed=16, scale=3.0)
    assert rel_err(ops.silu(v), F.silu(v.float())) < TOL


@pytest.mark.gpu
def test_conv_out_tensor_core_path_matches_small_kernel(cuda):
    """conv_out through lvd_gemm (output channels zero-padded to a 64-wide tile) + lvd_rows_to_ncfhw against the
    one-warp-per-pixel kernel and a torch fp32 conv."""
    from lvd_b200 import ops, repack
    g = torch.Generator().manual_seed(9)
    b, f, h, w, cin, cout = 2, 3, 10, 18, 320, 4
    x = torch.randn((b * f * h * w, cin), generator=g).half().to(cuda)
    wo = (torch.randn((cout, cin, 3, 3), generator=g) * 0.02).half().to(cuda)
    bo = (torch.randn((cout,), generator=g) * 0.1).half().to(cuda)
    wk = repack.conv_weight_krsc(wo)
    y_small = ops.conv_out(x, wk, bo, b, f, h, w)
    wp, bp = repack.pad_out_channels(wk, bo)
    y_tc = ops.conv_out_tc(x, wp, bp, cout, b, f, h, w)
    xi = x.float().reshape(b * f, h, w, cin).permute(0, 3, 1, 2)
    ref = torch.nn.functional.conv2d(xi, wo.float(), bo.float(), padding=1)
    ref = ref.reshape(b, f, cout, h, w).permute(0, 2, 1, 3, 4)
    assert y_tc.shape == (b, cout, f, h, w)
    assert rel_err(y_tc, ref) < 3e-3 and rel_err(y_small, ref) < 3e-3


@pytest.mark.parametrize("kind,shape,samples", [
    ("conv2d", (48, 20, 36, 640), 2),     # resnet conv2 -> the temporal conv layer's 5-D norm (sample = 24 frames)
    ("conv2d", (16, 16, 24, 320), 1),     # guidance pass, one prompt
    ("tconv", (2, 24, 720, 640), 2),      # temporal conv -> next norm of the chain
    ("tconv", (1, 16, 1024, 320), 1),
    ("conv2d_res", (48, 10, 18, 1280), 2),
])
def test_groupnorm_statistics_from_the_gemm_epilogue(cuda, kind, shape, samples):
    """GroupNorm whose statistics come from the producing conv's epilogue (lvd_gemm gn_partial ->
    lvd_groupnorm_apply) against the two-pass GroupNorm of the same tensor: same statistics up to fp32 summation order,
    bit-reproducible from run to run, and the partial buffer is back at zero afterwards."""
    from lvd_b200 import ops
    g = torch.Generator().manual_seed(5)
    if kind.startswith("conv2d"):
        n_img, h, w, c = shape
        rows = n_img * h * w
        x = torch.randn((rows, c), generator=g).half().to(cuda)
        wk = (torch.randn((c, 9 * c), generator=g) * 0.02).half().to(cuda)
        res = torch.randn((rows, c), generator=g).half().to(cuda) if kind.endswith("res") else None
        run = lambda gn: ops.conv2d_3x3(x, n_img, h, w, wk, residual=res, gn=gn)  # noqa: E731
    else:
        b, f, hw, c = shape
        rows = b * f * hw
        x = torch.randn((rows, c), generator=g).half().to(cuda)
        wk = (torch.randn((c, 3 * c), generator=g) * 0.03).half().to(cuda)
        run = lambda gn: ops.conv_temporal(x, b, f, hw, wk, gn=gn)  # noqa: E731
    gamma = (1 + 0.1 * torch.randn(c, generator=g)).half().to(cuda)
    beta = (0.1 * torch.randn(c, generator=g)).half().to(cuda)
    assert ops.GN_EPILOGUE
    y_plain = run(None)
    ref_stats = ops.gn_stats_buffer(samples, cuda)
    z_ref = ops.groupnorm(y_plain, samples, gamma, beta, 1e-5, silu=True, stats_out=ref_stats)
    outs = []
    for rep in range(3):
        y = run((samples, 32))
        assert getattr(y, "_lvd_gn", None) is not None, "the epilogue path was not taken"
        assert torch.equal(y, y_plain)
        st = ops.gn_stats_buffer(samples, cuda)
        z = ops.groupnorm(y, samples, gamma, beta, 1e-5, silu=True, stats_out=st)
        outs.append((z, st))
        assert ops._gn_pending[cuda] is None
        assert float(ops._gn_part[cuda].abs().max()) == 0.0  # self-cleaning
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[2][1])  # bit-reproducible
    z, st = outs[0]
    assert ((st - ref_stats).abs() / ref_stats.abs().clamp_min(1e-3)).max().item() < 1e-4
    assert rel_err(z, z_ref) <= 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("rows,c,n,kind", [
    (2880, 320, 960, "plain"),        # norm1 -> fused QKV (no bias)
    (1000, 640, 640, "plain"),        # ragged last row tile
    (2 * 8 * 160, 320, 960, "fm2sm"),  # temporal transformer: the consumer permutes rows, the producer un-permutes
    (2304, 320, 2560, "geglu"),       # norm3 -> ff.net.0 (GEGLU)
    (720, 1280, 10240, "geglu"),
])
def test_layernorm_folded_into_the_consuming_linear(cuda, rows, c, n, kind):
    """LN(x) @ W.T + b computed on the raw x (weight = gamma o W, epilogue rstd * (acc - mean * colsum) + bias') with the
    row statistics taken from the epilogue of the linear that produced x, against LayerNorm kernel + plain linear and
    against fp32 torch."""
    from lvd_b200 import ops, repack
    assert ops.LN_FOLD
    g = torch.Generator().manual_seed(11)
    a = torch.randn((rows, c), generator=g).half().to(cuda)
    res = (torch.randn((rows, c), generator=g) * 2 + 0.7).half().to(cuda)  # rows with a mean well away from 0
    wp = (torch.randn((c, c), generator=g) * 0.05).half().to(cuda)
    bp = (torch.randn(c, generator=g) * 0.1).half().to(cuda)
    gamma = (1 + 0.2 * torch.randn(c, generator=g)).half().to(cuda)
    beta = (0.2 * torch.randn(c, generator=g)).half().to(cuda)
    w = (torch.randn((n, c), generator=g) * 0.05).half().to(cuda)
    b = (torch.randn(n, generator=g) * 0.1).half().to(cuda) if kind == "geglu" else None
    act = ops.ACT_NONE
    if kind == "geglu":
        w, b = repack.geglu_interleave(w, b)
        act = ops.ACT_GEGLU
    batch, frames, hw = 2, 8, 160
    sm2fm = ("sm2fm", batch, frames, hw) if kind == "fm2sm" else None
    fm2sm = ("fm2sm", batch, frames, hw) if kind == "fm2sm" else None
    # producer: x = a @ wp.T + bp + res, emitting the row statistics of x
    x_plain = ops.linear(a, wp, bp, residual=res, row_map=sm2fm)
    x = ops.linear(a, wp, bp, residual=res, row_map=sm2fm, ln_emit=True)
    assert torch.equal(x, x_plain)
    stats = ops.ln_rowstats(x)
    assert stats is not None
    xf = x.float()
    mean_ref = xf.mean(1)
    rstd_ref = 1.0 / torch.sqrt(xf.var(1, unbiased=False) + 1e-5)
    assert (stats[:, 0] - mean_ref).abs().max().item() < 2e-5 * (1 + mean_ref.abs().max().item())
    assert ((stats[:, 1] - rstd_ref).abs() / rstd_ref).max().item() < 1e-4
    folded = ops.ln_fold(w, b, gamma, beta)
    y = ops.linear(x, folded.w, act=act, row_map=fm2sm, ln=(stats, folded))
    y_two = ops.linear(ops.layernorm(x, gamma, beta), w, b, act=act, row_map=fm2sm)
    # fp32 reference of the reference's own op order (LayerNorm, then the linear) on the same fp16 x
    ln = F.layer_norm(xf, (c,), gamma.float(), beta.float(), 1e-5)
    ref_rows = ln
    if kind == "fm2sm":
        ref_rows = ln.reshape(batch, frames, hw, c).permute(0, 2, 1, 3).reshape(rows, c)
    if kind == "geglu":
        # fp32 GEGLU on the interleaved weight: every block of 256 columns = [128 value | 128 gate]
        z = ref_rows @ w.float().t() + b.float()
        zz = z.reshape(rows, n // 256, 2, 128)
        ref = (zz[:, :, 0] * F.gelu(zz[:, :, 1])).reshape(rows, n // 2)
    else:
        ref = ref_rows @ w.float().t()
    e_fold, e_two = rel_err(y, ref), rel_err(y_two, ref)
    print("[ln fold %s rows=%d c=%d n=%d] folded-vs-fp32 %.3e   LN-kernel+linear-vs-fp32 %.3e" % (kind, rows, c, n, e_fold, e_two))
    assert e_fold <= max(1.2 * e_two, 6e-4)
    # a second producer into the same statistics buffer invalidates the first tensor's handle (no stale statistics)
    x2 = ops.linear(a, wp, bp, residual=res, row_map=sm2fm, ln_emit=True)
    assert ops.ln_rowstats(x) is None and ops.ln_rowstats(x2) is not None
