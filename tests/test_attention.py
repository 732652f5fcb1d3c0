"""GPU parity of the tcgen05 attention kernel against a plain PyTorch fp32 reference of the same op.
Outputs are fp16 and P is fp16 inside the kernel (as in the reference's fp16 bmm): tolerance 3e-3 of max|ref|."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 3e-3


def _rand(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).half()


def _ref_attn(q, k, v, heads, scale=0.125):
    # q [z, sq, heads*64], k/v [z, sk, heads*64]
    z, sq, _ = q.shape
    sk = k.shape[1]
    qh = q.float().view(z, sq, heads, 64).transpose(1, 2)
    kh = k.float().view(z, sk, heads, 64).transpose(1, 2)
    vh = v.float().view(z, sk, heads, 64).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    o = p @ vh
    return o.transpose(1, 2).reshape(z, sq, heads * 64), p


@pytest.mark.parametrize("z,s,heads", [(2, 128, 1), (3, 45, 2), (2, 180, 2), (2, 720, 2), (1, 2880, 5), (2, 210, 3),
                                       (2, 1024, 1), (4, 16, 2)])
def test_self_attention(cuda, z, s, heads):
    from lvd_b200 import ops
    c = heads * 64
    qkv = _rand((z * s, 3 * c), cuda, seed=1)
    out = ops.attention(qkv, qkv, heads=heads, nz=z, s_q=s, s_k=s, q_col0=0, k_col0=c, v_col0=2 * c)
    q, k, v = (qkv[:, i * c:(i + 1) * c].reshape(z, s, c) for i in range(3))
    ref, _ = _ref_attn(q, k, v, heads)
    assert rel_err(out, ref.reshape(z * s, c)) < TOL


@pytest.mark.parametrize("b,frames,s,heads,sk", [(1, 4, 180, 2, 77), (2, 3, 45, 4, 77), (1, 2, 720, 3, 77),
                                                 (1, 2, 300, 1, 107)])
def test_cross_attention_with_maps(cuda, b, frames, s, heads, sk):
    from lvd_b200 import ops
    c = heads * 64
    z = b * frames
    q = _rand((z * s, c), cuda, seed=2)
    kv = _rand((b * sk, 2 * c), cuda, seed=3)
    toks = (2, 5, 40, sk - 1)
    p_full = torch.zeros((z, heads, s, sk), dtype=torch.float16, device=cuda)
    p_tok = torch.zeros((len(toks), z, heads, s), dtype=torch.float16, device=cuda)
    out = ops.attention(q, kv, heads=heads, nz=z, s_q=s, s_k=sk, k_col0=0, v_col0=c, kv_z_div=frames,
                        p_full=p_full, p_tok=p_tok, tok_cols=toks)
    k = kv[:, :c].reshape(b, sk, c).repeat_interleave(frames, 0)
    v = kv[:, c:].reshape(b, sk, c).repeat_interleave(frames, 0)
    ref, p = _ref_attn(q.reshape(z, s, c), k, v, heads)
    assert rel_err(out, ref.reshape(z * s, c)) < TOL
    assert rel_err(p_full, p) < TOL
    for i, t in enumerate(toks):
        assert rel_err(p_tok[i], p[..., t]) < TOL


@pytest.mark.parametrize("b,frames,s,heads,sk", [(1, 4, 180, 2, 77), (2, 3, 45, 4, 77), (1, 2, 2880, 5, 77),
                                                 (1, 2, 301, 1, 107), (2, 2, 37, 2, 30), (1, 3, 50, 1, 128),
                                                 (1, 1, 16, 1, 1)])
def test_cross_attention_without_maps(cuda, b, frames, s, heads, sk):
    """No probability output requested: the warp-level mma.sync kernel (attention_temporal.cu) -- ragged query counts,
    key counts on both sides of every template boundary (32 / 80 / 128 keys), a single key."""
    from lvd_b200 import ops
    c = heads * 64
    z = b * frames
    q = _rand((z * s, c), cuda, seed=12)
    kv = _rand((b * sk, 2 * c), cuda, seed=13)
    out = ops.attention(q, kv, heads=heads, nz=z, s_q=s, s_k=sk, k_col0=0, v_col0=c, kv_z_div=frames)
    k = kv[:, :c].reshape(b, sk, c).repeat_interleave(frames, 0)
    v = kv[:, c:].reshape(b, sk, c).repeat_interleave(frames, 0)
    ref, _ = _ref_attn(q.reshape(z, s, c), k, v, heads)
    assert rel_err(out, ref.reshape(z * s, c)) < TOL


@pytest.mark.parametrize("b,frames,hw,heads", [(1, 24, 45, 2), (2, 24, 180, 1), (2, 16, 64, 2), (1, 8, 16, 1),
                                               (1, 24, 2880, 1), (1, 24, 7, 1)])
def test_temporal_attention(cuda, b, frames, hw, heads):
    from lvd_b200 import ops
    c = heads * 64
    rows = b * frames * hw
    qkv = _rand((rows, 3 * c), cuda, seed=4)
    out = ops.attention(qkv, qkv, heads=heads, nz=b, temporal=True, frames=frames, hw=hw, q_col0=0, k_col0=c,
                        v_col0=2 * c)

    def seq(t):  # rows ((b*F+f)*hw+p) -> [b*hw, F, c]
        return t.reshape(b, frames, hw, c).permute(0, 2, 1, 3).reshape(b * hw, frames, c)

    q, k, v = (seq(qkv[:, i * c:(i + 1) * c]) for i in range(3))
    ref, _ = _ref_attn(q, k, v, heads)
    ref = ref.reshape(b, hw, frames, c).permute(0, 2, 1, 3).reshape(rows, c)
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("b,frames,hw,heads", [(1, 24, 45, 2), (2, 24, 180, 1), (2, 16, 64, 2), (1, 8, 16, 1),
                                               (1, 24, 2880, 1), (1, 24, 7, 1), (1, 32, 9, 1), (2, 32, 33, 3),
                                               (3, 8, 5, 5)])
def test_temporal_attention_packed(cuda, b, frames, hw, heads):
    """mode 2: rows ((b*hw + p)*F + f); sequences are runs of F consecutive rows."""
    from lvd_b200 import ops
    c = heads * 64
    rows = b * hw * frames
    qkv = _rand((rows, 3 * c), cuda, seed=5)
    out = ops.attention(qkv, qkv, heads=heads, nz=b, temporal=True, packed=True, frames=frames, hw=hw, q_col0=0,
                        k_col0=c, v_col0=2 * c)
    q, k, v = (qkv[:, i * c:(i + 1) * c].reshape(b * hw, frames, c) for i in range(3))
    ref, _ = _ref_attn(q, k, v, heads)
    assert rel_err(out, ref.reshape(rows, c)) < TOL


def test_self_attention_lse(cuda):
    from lvd_b200 import ops
    z, s, heads = 2, 300, 2
    c = heads * 64
    qkv = _rand((z * s, 3 * c), cuda, seed=6)
    lse = torch.zeros((z, heads, s), dtype=torch.float32, device=cuda)
    ops.attention(qkv, qkv, heads=heads, nz=z, s_q=s, s_k=s, k_col0=c, v_col0=2 * c, lse=lse)
    q, k = (qkv[:, i * c:(i + 1) * c].reshape(z, s, heads, 64).transpose(1, 2).float() for i in range(2))
    ref = torch.logsumexp(q @ k.transpose(-1, -2) * 0.125, dim=-1) / 0.6931471805599453
    assert (lse - ref).abs().max().item() < 2e-3


def test_attn_processor_mirror_matches_reference_semantics(cuda):
    """AttnProcessor.__call__ boundary: fused path, materialising path with save_attn_to_dict / return probs."""
    from lvd_b200.models.attention_processor import Attention
    torch.manual_seed(0)
    b, s, c, heads, ctx = 4, 90, 128, 2, 128
    sd = {"to_q.weight": torch.randn(c, c) * c ** -0.5, "to_k.weight": torch.randn(c, ctx) * ctx ** -0.5,
          "to_v.weight": torch.randn(c, ctx) * ctx ** -0.5, "to_out.0.weight": torch.randn(c, c) * c ** -0.5,
          "to_out.0.bias": torch.randn(c) * 0.1}
    attn = Attention(query_dim=c, cross_attention_dim=ctx, heads=heads, dim_head=64).load_state_dict(sd)
    x = _rand((b, s, c), cuda, seed=7)
    e = _rand((b, 77, ctx), cuda, seed=8)
    w = {k: v.half().float().to(cuda) for k, v in sd.items()}

    def ref(xq, ctxt):
        q = xq.float() @ w["to_q.weight"].t()
        k = ctxt.float() @ w["to_k.weight"].t()
        v = ctxt.float() @ w["to_v.weight"].t()
        o, p = _ref_attn(q, k, v, heads)
        return o @ w["to_out.0.weight"].t() + w["to_out.0.bias"], p

    saved = {}
    out = attn(x, encoder_hidden_states=e, attn_key=["down", 1, 0, 0], save_attn_to_dict=saved,
               save_keys=[("down", 1, 0, 0)])
    o_ref, p_ref = ref(x, e)
    assert rel_err(out, o_ref) < 4e-3
    assert saved[("down", 1, 0, 0)].shape == (b, heads, s, 77)
    assert rel_err(saved[("down", 1, 0, 0)], p_ref) < 4e-3
    out2 = attn(x, encoder_hidden_states=e, attn_key=["down", 1, 1, 0], save_attn_to_dict={}, save_keys=[])
    assert rel_err(out2, o_ref) < 4e-3                     # unsaved key -> fused path, same numbers
    out3, p3 = attn(x, encoder_hidden_states=e, return_attntion_probs=True, return_token_ca_only=5,
                    return_cond_ca_only=True)
    assert p3.shape == (b // 2, heads, s, 1) and rel_err(p3[..., 0], p_ref[b // 2:, ..., 5]) < 4e-3
    self_attn = Attention(query_dim=c, heads=heads, dim_head=64).load_state_dict(
        {**sd, "to_k.weight": sd["to_q.weight"] * 0.5, "to_v.weight": sd["to_out.0.weight"]})
    ws = {**w, "to_k.weight": w["to_q.weight"] * 0.5, "to_v.weight": w["to_out.0.weight"]}
    q = x.float() @ ws["to_q.weight"].t()
    o, _ = _ref_attn(q, x.float() @ ws["to_k.weight"].t(), x.float() @ ws["to_v.weight"].t(), heads)
    assert rel_err(self_attn(x), o @ ws["to_out.0.weight"].t() + ws["to_out.0.bias"]) < 4e-3
