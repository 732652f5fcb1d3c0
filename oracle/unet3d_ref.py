"""TEST INFRASTRUCTURE: plain-PyTorch restatement of the reference's per-timestep UNet3DConditionModel.forward.

Functional (a `state_dict` in, tensors out), dtype-generic (fp32 on CPU is the oracle; fp16 on the GPU reproduces what
the reference computes there), NCHW like the reference so that it shares no layout code with the channels-last CUDA
engine it checks.  Each function cites the reference lines it follows; pinned against the reference itself by
tests/test_oracle_pin.py (builder container) and against tests/golden/*.npz (everywhere).
"""
import math

import torch
import torch.nn.functional as F

GUIDANCE_KEYS_DEFAULT = [("down", 1, 0, 0), ("down", 2, 0, 0), ("down", 2, 1, 0), ("up", 1, 0, 0), ("up", 1, 1, 0),
                         ("up", 2, 2, 0)]  # generation/lvd.py:66-73


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _gn(sd, p, x, eps, groups=32):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def timestep_embedding(t, dim):
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0) (unet_3d_condition.py:304,712)."""
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t[:, None].float() * freq[None]
    return torch.cat([a.cos(), a.sin()], dim=-1)


def attention(sd, p, x, ctx, heads, save=None):
    """models/attention_processor.py:344-430 (fast) and :476-589 (slow, materialising).  Returns (out, probs|None);
    probs [batch, heads, S_q, S_k] are produced exactly like get_attention_scores (:222-258): softmax of
    alpha-scaled scores, cast to the working dtype."""
    b, sq, _ = x.shape
    c = x if ctx is None else ctx
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(c, sd[p + ".to_k.weight"])
    v = F.linear(c, sd[p + ".to_v.weight"])
    d = q.shape[-1] // heads
    q = q.view(b, sq, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    probs = scores.float().softmax(dim=-1).to(q.dtype)
    o = torch.matmul(probs, v).transpose(1, 2).reshape(b, sq, heads * d)
    o = _lin(sd, p + ".to_out.0", o)
    return o, (probs if save else None)


def feed_forward(sd, p, x):
    """FeedForward with GEGLU (models/attention.py:355-376): first half = value, second half = gate, erf GELU."""
    h, gate = _lin(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return _lin(sd, p + ".net.2", h * F.gelu(gate))


def gated_self_attention(sd, p, x, objs, heads):
    """GatedSelfAttentionDense.forward (models/attention.py:44-60)."""
    n_visual = x.shape[1]
    objs = _lin(sd, p + ".linear", objs)
    a, _ = attention(sd, p + ".attn", _ln(sd, p + ".norm1", torch.cat([x, objs], dim=1)), None, heads)
    x = x + sd[p + ".alpha_attn"].tanh() * a[:, :n_visual, :]
    x = x + sd[p + ".alpha_dense"].tanh() * feed_forward(sd, p + ".ff", _ln(sd, p + ".norm2", x))
    return x


def basic_block(sd, p, x, ctx, heads, *, double_self, key=None, saved=None, save_keys=None, objs=None,
                fuser_enabled=True):
    """BasicTransformerBlock.forward (models/attention.py:179-276)."""
    a, _ = attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, heads)
    x = a + x
    if objs is not None and (p + ".fuser.linear.weight") in sd and fuser_enabled:
        x = gated_self_attention(sd, p + ".fuser", x, objs, heads)
    want = (not double_self) and saved is not None and save_keys is not None and key in save_keys
    a, probs = attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), None if double_self else ctx, heads, save=want)
    if want:
        saved[key] = probs  # (B*F, heads, hw, 77)  attention_processor.py:553-586
    x = a + x
    x = feed_forward(sd, p + ".ff", _ln(sd, p + ".norm3", x)) + x
    return x


def transformer_2d(sd, p, x, ctx, heads, key, saved, save_keys, objs, fuser_enabled):
    """Transformer2DModel.forward, continuous input + linear projection (models/transformer_2d.py:309-370)."""
    n, c, h, w = x.shape
    res = x
    t = _gn(sd, p + ".norm", x, 1e-6).permute(0, 2, 3, 1).reshape(n, h * w, c)
    t = _lin(sd, p + ".proj_in", t)
    t = basic_block(sd, p + ".transformer_blocks.0", t, ctx, heads, double_self=False, key=key + (0,), saved=saved,
                    save_keys=save_keys, objs=objs, fuser_enabled=fuser_enabled)
    t = _lin(sd, p + ".proj_out", t)
    return t.reshape(n, h, w, c).permute(0, 3, 1, 2) + res


def transformer_temporal(sd, p, x, frames, heads):
    """TransformerTemporalModel.forward (models/transformer_temporal.py:142-189): GroupNorm over the 5-D tensor,
    sequences along the frame axis, both attentions are self-attention; the final reshape is the inverse of the
    input permute."""
    bf, c, h, w = x.shape
    b = bf // frames
    res = x
    t = x.reshape(b, frames, c, h, w).permute(0, 2, 1, 3, 4)
    t = _gn(sd, p + ".norm", t, 1e-6)
    t = t.permute(0, 3, 4, 2, 1).reshape(b * h * w, frames, c)
    t = _lin(sd, p + ".proj_in", t)
    t = basic_block(sd, p + ".transformer_blocks.0", t, None, heads, double_self=True)
    t = _lin(sd, p + ".proj_out", t)
    t = t.reshape(b, h, w, frames, c).permute(0, 3, 4, 1, 2).reshape(bf, c, h, w)
    return t + res


def resnet(sd, p, x, temb, eps):
    """diffusers 0.27.2 ResnetBlock2D.forward."""
    h = F.conv2d(F.silu(_gn(sd, p + ".norm1", x, eps)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = h + _lin(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, p + ".norm2", h, eps)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def temp_conv(sd, p, x, frames):
    """diffusers 0.27.2 TemporalConvLayer.forward: 4 x [GN(5-D) -> SiLU -> Conv3d(3,1,1)]
    + identity."""
    bf, c, h, w = x.shape
    t = x.reshape(bf // frames, frames, c, h, w).permute(0, 2, 1, 3, 4)
    idt = t
    for name, idx in (("conv1", 2), ("conv2", 3), ("conv3", 3), ("conv4", 3)):
        t = F.silu(_gn(sd, f"{p}.{name}.0", t, 1e-5))
        t = F.conv3d(t, sd[f"{p}.{name}.{idx}.weight"], sd[f"{p}.{name}.{idx}.bias"], padding=(1, 0, 0))
    t = idt + t
    return t.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


def position_net(sd, boxes, masks, positive_embeddings):
    """PositionNet.forward, text-only (models/unet_3d_condition.py:119-150) with FourierEmbedder (:47-64)."""
    masks = masks.unsqueeze(-1)
    freq = (100 ** (torch.arange(8, device=boxes.device) / 8)).to(boxes.dtype)[None, None, None]
    xf = freq * boxes.unsqueeze(-1)
    emb = torch.stack((xf.sin(), xf.cos()), dim=-1).permute(0, 1, 3, 4, 2).reshape(*boxes.shape[:2], -1)
    emb = emb * masks + (1 - masks) * sd["position_net.null_position_feature"].view(1, 1, -1)
    pos = positive_embeddings * masks + (1 - masks) * sd["position_net.null_positive_feature"].view(1, 1, -1)
    h = torch.cat([pos, emb], dim=-1)
    h = F.silu(_lin(sd, "position_net.linears.0", h))
    h = F.silu(_lin(sd, "position_net.linears.2", h))
    return _lin(sd, "position_net.linears.4", h)


def unet_forward(sd, cfg, sample, timestep, encoder_hidden_states, *, save_keys=None, saved=None, gligen=None,
                 fuser_enabled=True, stop_after_key=None):
    """UNet3DConditionModel.forward (models/unet_3d_condition.py:642-859).  sample (B,4,F,h,w).
    `saved` (dict) receives the cross-attention probabilities of the keys in `save_keys`
    (key = (kind, stage, layer, 0)).  `stop_after_key`: the guidance pass needs nothing after that key
    (the reference computes the dead tail anyway, models/pipelines.py:99)."""
    boc = list(cfg["block_out_channels"])
    hd = cfg["attention_head_dim"]
    eps = cfg["norm_eps"]
    b, _, frames, h, w = sample.shape
    dt = sample.dtype
    t = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(b)
    temb = timestep_embedding(t, boc[0]).to(dt)
    temb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", temb)))
    temb = temb.repeat_interleave(frames, dim=0)
    ctx = encoder_hidden_states.repeat_interleave(frames, dim=0)
    x = sample.permute(0, 2, 1, 3, 4).reshape(b * frames, -1, h, w)
    x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    x = transformer_temporal(sd, "transformer_in", x, frames, 8)
    objs = position_net(sd, **gligen) if gligen is not None else None

    class _Stop(Exception):
        pass

    def layer(p, j, x, kind, stage, attn):
        x = resnet(sd, f"{p}.resnets.{j}", x, temb, eps)
        x = temp_conv(sd, f"{p}.temp_convs.{j}", x, frames)
        if attn:
            c = x.shape[1]
            x = transformer_2d(sd, f"{p}.attentions.{j}", x, ctx, c // hd, (kind, stage, j), saved, save_keys, objs,
                               fuser_enabled)
            if stop_after_key is not None and (kind, stage, j, 0) == tuple(stop_after_key):
                raise _Stop()
            x = transformer_temporal(sd, f"{p}.temp_attentions.{j}", x, frames, c // hd)
        return x

    try:
        skips = [x]
        nb = len(boc)
        for i in range(nb):
            p = f"down_blocks.{i}"
            for j in range(cfg["layers_per_block"]):
                x = layer(p, j, x, "down", i, i < nb - 1)
                skips.append(x)
            if i < nb - 1:
                x = F.conv2d(x, sd[f"{p}.downsamplers.0.conv.weight"], sd[f"{p}.downsamplers.0.conv.bias"], stride=2,
                             padding=1)
                skips.append(x)
        # UNetMidBlock3DCrossAttn.forward (unet_3d_blocks.py:255-291)
        x = resnet(sd, "mid_block.resnets.0", x, temb, eps)
        x = temp_conv(sd, "mid_block.temp_convs.0", x, frames)
        x = transformer_2d(sd, "mid_block.attentions.0", x, ctx, x.shape[1] // hd, ("mid", 0, 0), saved, save_keys,
                           objs, fuser_enabled)
        x = transformer_temporal(sd, "mid_block.temp_attentions.0", x, frames, x.shape[1] // hd)
        x = resnet(sd, "mid_block.resnets.1", x, temb, eps)
        x = temp_conv(sd, "mid_block.temp_convs.1", x, frames)
        for i in range(nb):
            p = f"up_blocks.{i}"
            for j in range(cfg["layers_per_block"] + 1):
                x = torch.cat([x, skips.pop()], dim=1)
                x = layer(p, j, x, "up", i, i > 0)
            if i < nb - 1:
                x = F.interpolate(x, scale_factor=2.0, mode="nearest")
                x = F.conv2d(x, sd[f"{p}.upsamplers.0.conv.weight"], sd[f"{p}.upsamplers.0.conv.bias"], padding=1)
    except _Stop:
        return None
    x = F.silu(_gn(sd, "conv_norm_out", x, eps))
    x = F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    return x.reshape(b, frames, -1, h, w).permute(0, 2, 1, 3, 4)
