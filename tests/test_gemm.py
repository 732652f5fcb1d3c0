"""GPU parity of the tcgen05 contraction kernel (lvd_gemm) against a plain PyTorch fp32 reference of the same op.
Tolerance: outputs are fp16 (rel. rounding 2^-11 ~ 4.9e-4); we require max|err|/max|ref| <= 2e-3."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-3


@pytest.fixture(autouse=True, params=["auto", "single_cta"])
def cta_mode(request):
    """Every case runs through both instantiations of the kernel: CTA pairs (cta_group::2, picked automatically when a
    problem has >= 2 row tiles) and one CTA per tile."""
    from lvd_b200 import ops
    ops.FORCE_CTA_PAIR = 1 if request.param == "single_cta" else 0
    yield
    ops.FORCE_CTA_PAIR = 0


def _rand(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).half()


@pytest.mark.parametrize("rows,n,k", [(128, 64, 64), (300, 128, 64), (1000, 320, 320), (4096, 256, 512),
                                      (8640, 1280, 1280), (77 * 2, 640, 1024)])
def test_linear_plain(cuda, rows, n, k):
    from lvd_b200 import ops
    x = _rand((rows, k), cuda, seed=1)
    w = _rand((n, k), cuda, k ** -0.5, seed=2)
    out = ops.linear(x, w)
    ref = x.float() @ w.float().t()
    assert rel_err(out, ref) < TOL


def test_linear_bias_residual_strided(cuda):
    from lvd_b200 import ops
    rows, n, k = 777, 320, 640
    buf = _rand((rows, k + 64), cuda, seed=3)
    x = buf[:, 64:]                      # column slice: ld = k + 64
    w = _rand((n, k), cuda, k ** -0.5, seed=4)
    b = _rand((n,), cuda, seed=5)
    res = _rand((rows, n), cuda, seed=6)
    outbuf = torch.zeros((rows, 2 * n), dtype=torch.float16, device=cuda)
    out = ops.linear(x, w, b, residual=res, out=outbuf[:, n:])
    ref = (x.float() @ w.float().t() + b.float()).half().float() + res.float()
    assert rel_err(out, ref) < TOL
    assert outbuf[:, :n].abs().max().item() == 0.0


def test_linear_mn_major(cuda):
    from lvd_b200 import ops
    rows, n, k = 500, 256, 192
    x = _rand((rows, k), cuda, seed=7)
    wt = _rand((k, n), cuda, k ** -0.5, seed=8)   # [K, N], N contiguous
    out = ops.linear(x, wt, b_mn_major=True)
    ref = x.float() @ wt.float()
    assert rel_err(out, ref) < TOL


def test_linear_silu_gamma(cuda):
    from lvd_b200 import ops
    rows, n, k = 130, 512, 1088
    x = _rand((rows, k), cuda, seed=9)
    w = _rand((n, k), cuda, k ** -0.5, seed=10)
    b = _rand((n,), cuda, seed=11)
    out = ops.linear(x, w, b, act=ops.ACT_SILU, gamma=0.37)
    ref = F.silu((x.float() @ w.float().t() + b.float()).half().float()).half().float() * 0.37
    assert rel_err(out, ref) < TOL


def test_linear_geglu(cuda):
    from lvd_b200 import ops, repack
    rows, c = 333, 320
    x = _rand((rows, c), cuda, seed=12)
    w = _rand((8 * c, c), cuda, c ** -0.5, seed=13)
    b = _rand((8 * c,), cuda, seed=14)
    wi, bi = repack.geglu_interleave(w, b)
    out = ops.linear(x, wi, bi, act=ops.ACT_GEGLU)
    y = (x.float() @ w.float().t() + b.float()).half().float()
    val, gate = y.chunk(2, dim=-1)
    ref = val * F.gelu(gate).half().float()
    assert out.shape == (rows, 4 * c)
    assert rel_err(out, ref) < TOL


def test_linear_rowvec(cuda):
    from lvd_b200 import ops
    n_img, hw, n, k = 6, 45, 128, 128
    x = _rand((n_img * hw, k), cuda, seed=15)
    w = _rand((n, k), cuda, k ** -0.5, seed=16)
    rv = _rand((n_img, n), cuda, seed=17)
    out = ops.linear(x, w, rowvec=rv, rows_per_vec=hw)
    ref = (x.float() @ w.float().t()).half().float() + rv.float().repeat_interleave(hw, 0)
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("n_img,h,w,cin,cout", [(2, 8, 8, 64, 64), (4, 10, 18, 128, 128), (3, 5, 9, 64, 320),
                                                (2, 40, 72, 320, 320), (48, 20, 36, 64, 64)])
def test_conv3x3(cuda, n_img, h, w, cin, cout):
    from lvd_b200 import ops, repack
    x = _rand((n_img, cin, h, w), cuda, seed=18)
    wt = _rand((cout, cin, 3, 3), cuda, (9 * cin) ** -0.5, seed=19)
    b = _rand((cout,), cuda, seed=20)
    temb = _rand((n_img, cout), cuda, seed=21)
    res = _rand((n_img * h * w, cout), cuda, seed=22)
    xl = x.permute(0, 2, 3, 1).reshape(n_img * h * w, cin).contiguous()
    out = ops.conv2d_3x3(xl, n_img, h, w, repack.conv_weight_krsc(wt), b, rowvec=temb, residual=res)
    ref = F.conv2d(x.float(), wt.float(), b.float(), padding=1).half().float() + temb.float()[:, :, None, None]
    ref = ref.half().float().permute(0, 2, 3, 1).reshape(n_img * h * w, cout) + res.float()
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("n_img,h,w,c", [(2, 8, 8, 64), (3, 10, 18, 128), (4, 40, 72, 64)])
def test_conv3x3_stride2(cuda, n_img, h, w, c):
    from lvd_b200 import ops, repack
    x = _rand((n_img, c, h, w), cuda, seed=23)
    wt = _rand((c, c, 3, 3), cuda, (9 * c) ** -0.5, seed=24)
    b = _rand((c,), cuda, seed=25)
    xl = x.permute(0, 2, 3, 1).reshape(n_img * h * w, c).contiguous()
    out = ops.conv2d_3x3_s2(xl, n_img, h, w, repack.conv_weight_krsc(wt), b)
    ref = F.conv2d(x.float(), wt.float(), b.float(), stride=2, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, c)
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("batch,frames,hw,c", [(1, 8, 16, 64), (2, 24, 45, 128), (2, 16, 64, 64), (1, 24, 180, 320)])
def test_conv_temporal(cuda, batch, frames, hw, c):
    from lvd_b200 import ops, repack
    x = _rand((batch, c, frames, hw, 1), cuda, seed=26)
    wt = _rand((c, c, 3, 1, 1), cuda, (3 * c) ** -0.5, seed=27)
    b = _rand((c,), cuda, seed=28)
    xl = x[..., 0].permute(0, 2, 3, 1).reshape(batch * frames * hw, c).contiguous()
    res = _rand((batch * frames * hw, c), cuda, seed=29)
    out = ops.conv_temporal(xl, batch, frames, hw, repack.tconv_weight(wt), b, residual=res)
    ref = F.conv3d(x.float(), wt.float(), b.float(), padding=(1, 0, 0)).half().float()
    ref = ref[..., 0].permute(0, 2, 3, 1).reshape(-1, c) + res.float()
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("batch,frames,hw", [(2, 24, 45), (1, 16, 64), (2, 8, 300)])
def test_linear_row_maps(cuda, batch, frames, hw):
    """frame-major <-> sequence-major row permutations folded into the GEMM (temporal transformer permutes)."""
    from lvd_b200 import ops
    rows, k, n = batch * frames * hw, 128, 192
    x = _rand((rows, k), cuda, seed=30)
    w = _rand((n, k), cuda, k ** -0.5, seed=31)
    res = _rand((rows, n), cuda, seed=32)
    y = x.float() @ w.float().t()
    out = ops.linear(x, w, row_map=("fm2sm", batch, frames, hw))
    ref = y.reshape(batch, frames, hw, n).permute(0, 2, 1, 3).reshape(rows, n)
    assert rel_err(out, ref) < TOL
    out2 = ops.linear(x, w, residual=res, row_map=("sm2fm", batch, frames, hw))
    ref2 = y.half().float().reshape(batch, hw, frames, n).permute(0, 2, 1, 3).reshape(rows, n) + res.float()
    assert rel_err(out2, ref2) < TOL


@pytest.mark.parametrize("rows,n,k,res,geglu", [(1000, 960, 320, False, False), (777, 320, 320, True, False),
                                                (600, 640, 640, True, False), (515, 2560, 320, False, True)])
def test_linear_fp32_bias_copy_is_bit_identical(cuda, rows, n, k, res, geglu):
    """The specialised epilogues read an fp32 copy of the bias (lvd_gemm_desc.bias_f32) instead of converting the fp16
    vector per row and slab: the same numbers, so the outputs must be bit-identical."""
    from lvd_b200 import ops, repack
    x = _rand((rows, k), cuda, seed=41)
    w = _rand((n, k), cuda, k ** -0.5, seed=42)
    b = _rand((n,), cuda, seed=43)
    r_ = _rand((rows, n), cuda, seed=44) if res else None
    act = ops.ACT_NONE
    if geglu:
        w, b = repack.geglu_interleave(w, b)
        act = ops.ACT_GEGLU
    try:
        ops.BIAS_F32 = False
        a = ops.linear(x, w, b, residual=r_, act=act)
    finally:
        ops.BIAS_F32 = True
    c = ops.linear(x, w, b, residual=r_, act=act)
    assert torch.equal(a, c)
    if not geglu:
        ref = (x.float() @ w.float().t() + b.float()).half().float() + (r_.float() if res else 0.0)
        assert rel_err(c, ref) < TOL


@pytest.mark.parametrize("rows,n,k,res,pair", [(20000, 960, 320, False, 1), (20000, 320, 320, True, 2),
                                               (9000, 640, 640, True, 2), (9000, 640, 640, False, 1)])
def test_linear_weight_resident_walk_is_bit_identical(cuda, rows, n, k, res, pair):
    """lvd_gemm_desc.weight_resident: every CTA (pair) keeps one B tile in shared memory for the whole launch and only
    A streams; a different tile-to-CTA assignment of the same arithmetic."""
    from lvd_b200 import ops
    x = _rand((rows, k), cuda, seed=51)
    w = _rand((n, k), cuda, k ** -0.5, seed=52)
    b = _rand((n,), cuda, seed=53)
    r_ = _rand((rows, n), cuda, seed=54) if res else None
    try:
        ops.FORCE_CTA_PAIR = pair
        ops.FORCE_WEIGHT_RESIDENT = 0
        a = ops.linear(x, w, b, residual=r_)
        ops.FORCE_WEIGHT_RESIDENT = 1
        c = ops.linear(x, w, b, residual=r_)
    finally:
        ops.FORCE_CTA_PAIR, ops.FORCE_WEIGHT_RESIDENT = 0, None
    assert torch.equal(a, c)
    ref = (x.float() @ w.float().t() + b.float()).half().float() + (r_.float() if res else 0.0)
    assert rel_err(c, ref) < TOL
