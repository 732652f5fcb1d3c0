"""Parameter inventory of the reference UNet3DConditionModel and deterministic synthetic weights.

`param_spec(cfg)` restates, name by name, the `state_dict` the reference module builds
(/root/reference/models/unet_3d_condition.py:228-445, unet_3d_blocks.py, transformer_2d.py:137-243,
transformer_temporal.py:60-107, attention.py:26-60,100-177 and the diffusers 0.27.2 resnet/embedding modules);
tests/test_oracle_pin.py checks it key-by-key and shape-by-shape against the reference where /root/reference exists,
and against the committed key list (tests/golden/state_dict_spec_*.json) everywhere else.

`random_state_dict(cfg, seed)` fills it with seeded values (no checkpoint is available offline): weights ~ N(0, 1/fan_in)
(x gain), norm scales ~ 1 + 0.1 N, biases ~ 0.1 N; the parameters the reference zero-initialises (TemporalConvLayer.conv4,
fuser alpha_attn / alpha_dense, PositionNet null features) get non-zero values so that those paths are exercised
(SURVEY.md section 8c).
"""
from collections import OrderedDict

import torch

ZEROSCOPE = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 cross_attention_dim=1024, attention_head_dim=64, norm_num_groups=32, norm_eps=1e-5, gated=False)
# small configuration with the same topology for fast CPU oracles / golden vectors
TINY = dict(ZEROSCOPE, block_out_channels=(64, 128, 256, 256), cross_attention_dim=128)


def make_cfg(base="zeroscope", **over):
    cfg = dict(ZEROSCOPE if base == "zeroscope" else TINY)
    cfg.update(over)
    return cfg


def _lin(spec, name, out_f, in_f, bias=True):
    spec[name + ".weight"] = (out_f, in_f)
    if bias:
        spec[name + ".bias"] = (out_f,)


def _norm(spec, name, c):
    spec[name + ".weight"] = (c,)
    spec[name + ".bias"] = (c,)


def _conv2d(spec, name, co, ci, k):
    spec[name + ".weight"] = (co, ci, k, k)
    spec[name + ".bias"] = (co,)


def _attention(spec, p, dim, ctx_dim):
    _lin(spec, p + ".to_q", dim, dim, bias=False)
    _lin(spec, p + ".to_k", dim, ctx_dim, bias=False)
    _lin(spec, p + ".to_v", dim, ctx_dim, bias=False)
    _lin(spec, p + ".to_out.0", dim, dim)


def _ff(spec, p, dim):
    _lin(spec, p + ".net.0.proj", 8 * dim, dim)
    _lin(spec, p + ".net.2", dim, 4 * dim)


def _basic_block(spec, p, dim, ctx_dim, gated, cross_dim_for_fuser):
    # BasicTransformerBlock (models/attention.py:100-177): norm1, attn1 (self), norm2, attn2, norm3, ff, [fuser]
    _norm(spec, p + ".norm1", dim)
    _attention(spec, p + ".attn1", dim, dim)
    _norm(spec, p + ".norm2", dim)
    _attention(spec, p + ".attn2", dim, ctx_dim)
    _norm(spec, p + ".norm3", dim)
    _ff(spec, p + ".ff", dim)
    if gated:
        f = p + ".fuser"
        spec[f + ".alpha_attn"] = ()
        spec[f + ".alpha_dense"] = ()
        _lin(spec, f + ".linear", dim, cross_dim_for_fuser)
        _attention(spec, f + ".attn", dim, dim)
        _ff(spec, f + ".ff", dim)
        _norm(spec, f + ".norm1", dim)
        _norm(spec, f + ".norm2", dim)


def _spatial_transformer(spec, p, c, cfg):
    _norm(spec, p + ".norm", c)
    _lin(spec, p + ".proj_in", c, c)
    _basic_block(spec, p + ".transformer_blocks.0", c, cfg["cross_attention_dim"], cfg["gated"],
                 cfg["cross_attention_dim"])
    _lin(spec, p + ".proj_out", c, c)


def _temporal_transformer(spec, p, c, inner):
    _norm(spec, p + ".norm", c)
    _lin(spec, p + ".proj_in", inner, c)
    _basic_block(spec, p + ".transformer_blocks.0", inner, inner, False, None)
    _lin(spec, p + ".proj_out", c, inner)


def _resnet(spec, p, cin, cout, temb):
    _norm(spec, p + ".norm1", cin)
    _conv2d(spec, p + ".conv1", cout, cin, 3)
    _lin(spec, p + ".time_emb_proj", cout, temb)
    _norm(spec, p + ".norm2", cout)
    _conv2d(spec, p + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv2d(spec, p + ".conv_shortcut", cout, cin, 1)


def _temp_conv(spec, p, c):
    for name, idx in (("conv1", 2), ("conv2", 3), ("conv3", 3), ("conv4", 3)):
        _norm(spec, f"{p}.{name}.0", c)
        spec[f"{p}.{name}.{idx}.weight"] = (c, c, 3, 1, 1)
        spec[f"{p}.{name}.{idx}.bias"] = (c,)


def up_block_resnet_channels(cfg):
    """(in_channels, out_channels) of the three resnets of every up block (unet_3d_blocks.py:555-575)."""
    boc = list(cfg["block_out_channels"])
    rc = list(reversed(boc))
    out = []
    output_channel = rc[0]
    for i in range(len(boc)):
        prev_output_channel = output_channel
        output_channel = rc[i]
        input_channel = rc[min(i + 1, len(boc) - 1)]
        layers = []
        for j in range(cfg["layers_per_block"] + 1):
            res_skip = input_channel if j == cfg["layers_per_block"] else output_channel
            rin = prev_output_channel if j == 0 else output_channel
            layers.append((rin + res_skip, output_channel))
        out.append(layers)
    return out


def param_spec(cfg):
    spec = OrderedDict()
    boc = list(cfg["block_out_channels"])
    c0 = boc[0]
    temb = 4 * c0
    nb = len(boc)
    _conv2d(spec, "conv_in", c0, cfg["in_channels"], 3)
    _lin(spec, "time_embedding.linear_1", temb, c0)
    _lin(spec, "time_embedding.linear_2", temb, temb)
    _temporal_transformer(spec, "transformer_in", c0, 8 * cfg["attention_head_dim"])
    out_c = c0
    for i in range(nb):
        in_c, out_c = out_c, boc[i]
        p = f"down_blocks.{i}"
        has_attn = i < nb - 1
        for j in range(cfg["layers_per_block"]):
            if has_attn:
                _spatial_transformer(spec, f"{p}.attentions.{j}", out_c, cfg)
        for j in range(cfg["layers_per_block"]):
            _resnet(spec, f"{p}.resnets.{j}", in_c if j == 0 else out_c, out_c, temb)
        for j in range(cfg["layers_per_block"]):
            if has_attn:
                _temporal_transformer(spec, f"{p}.temp_attentions.{j}", out_c, out_c)
        for j in range(cfg["layers_per_block"]):
            _temp_conv(spec, f"{p}.temp_convs.{j}", out_c)
        if i < nb - 1:
            _conv2d(spec, f"{p}.downsamplers.0.conv", out_c, out_c, 3)
    cm = boc[-1]
    _resnet(spec, "mid_block.resnets.0", cm, cm, temb)
    _temp_conv(spec, "mid_block.temp_convs.0", cm)
    _spatial_transformer(spec, "mid_block.attentions.0", cm, cfg)
    _temporal_transformer(spec, "mid_block.temp_attentions.0", cm, cm)
    _resnet(spec, "mid_block.resnets.1", cm, cm, temb)
    _temp_conv(spec, "mid_block.temp_convs.1", cm)
    upc = up_block_resnet_channels(cfg)
    for i in range(nb):
        p = f"up_blocks.{i}"
        has_attn = i > 0
        for j, (rin, rout) in enumerate(upc[i]):
            _resnet(spec, f"{p}.resnets.{j}", rin, rout, temb)
            _temp_conv(spec, f"{p}.temp_convs.{j}", rout)
            if has_attn:
                _spatial_transformer(spec, f"{p}.attentions.{j}", rout, cfg)
                _temporal_transformer(spec, f"{p}.temp_attentions.{j}", rout, rout)
        if i < nb - 1:
            _conv2d(spec, f"{p}.upsamplers.0.conv", upc[i][0][1], upc[i][0][1], 3)
    _norm(spec, "conv_norm_out", c0)
    _conv2d(spec, "conv_out", cfg["out_channels"], c0, 3)
    if cfg["gated"]:
        cd = cfg["cross_attention_dim"]
        _lin(spec, "position_net.linears.0", 512, cd + 64)
        _lin(spec, "position_net.linears.2", 512, 512)
        _lin(spec, "position_net.linears.4", cd, 512)
        spec["position_net.null_positive_feature"] = (cd,)
        spec["position_net.null_position_feature"] = (64,)
    return spec


def random_state_dict(cfg, seed=0, dtype=torch.float32, gain=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = OrderedDict()
    for name, shape in param_spec(cfg).items():
        leaf = name.rsplit(".", 1)[-1]
        is_norm = (".norm" in name or name.startswith("conv_norm_out") or
                   (".temp_convs." in name and name.rsplit(".", 2)[-2] == "0"))
        if leaf in ("alpha_attn", "alpha_dense"):
            t = torch.tensor(0.6 if leaf == "alpha_attn" else -0.4)
        elif leaf.startswith("null_"):
            t = 0.2 * torch.randn(shape, generator=g)
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)
        sd[name] = t.to(dtype)
    return sd


# ---------------------------------------------------------------------------------------------- VAE decoder
# diffusers 0.27.2 AutoencoderKL (SD-1.x VAE config used by Zeroscope / ModelScope: SURVEY.md Appendix B.8), decoder
# half only: `post_quant_conv` + `decoder.*` state_dict names and shapes.
VAE_SD = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
              norm_num_groups=32, scaling_factor=0.18215)
VAE_TINY = dict(VAE_SD, block_out_channels=(64, 64, 128, 128))


def vae_param_spec(cfg):
    spec = OrderedDict()
    boc = list(cfg["block_out_channels"])
    lc = cfg["latent_channels"]
    top = boc[-1]
    _conv2d(spec, "post_quant_conv", lc, lc, 1)
    _conv2d(spec, "decoder.conv_in", top, lc, 3)

    def resnet(p, cin, cout):
        _norm(spec, p + ".norm1", cin)
        _conv2d(spec, p + ".conv1", cout, cin, 3)
        _norm(spec, p + ".norm2", cout)
        _conv2d(spec, p + ".conv2", cout, cout, 3)
        if cin != cout:
            _conv2d(spec, p + ".conv_shortcut", cout, cin, 1)

    resnet("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    _norm(spec, a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _lin(spec, a + "." + n, top, top)
    resnet("decoder.mid_block.resnets.1", top, top)
    rev = boc[::-1]
    prev = rev[0]
    for i, c in enumerate(rev):
        for j in range(cfg["layers_per_block"] + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i != len(rev) - 1:
            _conv2d(spec, f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
        prev = c
    _norm(spec, "decoder.conv_norm_out", boc[0])
    _conv2d(spec, "decoder.conv_out", cfg["out_channels"], boc[0], 3)
    return spec


def random_vae_state_dict(cfg, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = OrderedDict()
    for name, shape in vae_param_spec(cfg).items():
        leaf = name.rsplit(".", 1)[-1]
        is_norm = "norm" in name.rsplit(".", 2)[-2]
        if is_norm and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * fan_in ** -0.5
        sd[name] = t.to(dtype)
    return sd
