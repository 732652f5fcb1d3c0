"""TEST INFRASTRUCTURE -- plain-PyTorch restatement of the DECODE half of diffusers 0.27.2 `AutoencoderKL`
(`models/autoencoders/autoencoder_kl.py::AutoencoderKL.decode`, `models/autoencoders/vae.py::Decoder`,
`models/unets/unet_2d_blocks.py::{UNetMidBlock2D, UpDecoderBlock2D}`, `models/resnet.py::ResnetBlock2D`,
`models/attention_processor.py::Attention`), the VAE behind the reference's `decode_latents`
(models/controllable_pipeline_text_to_video_synth.py:374-400).

PINNED against an independent public implementation that IS installed offline: transformers' `JanusVQVAEDecoder`, a port
of the CompVis / taming-transformers `Decoder` -- the code the SD-1.x VAE was trained with, of which diffusers'
AutoencoderKL decoder is a renaming (tests/test_vae.py::test_vae_oracle_pinned_against_transformers_compvis_decoder:
same weights under the other naming, strict load, images equal to 2e-5 at the tiny and the published SD-1.x width:
block_out_channels (128, 256, 512, 512), layers_per_block 2, 32 groups, eps 1e-6, one attention head of width 512 in
the mid block with residual connection, nearest-2x up-sampling + 3x3 conv after the first three up blocks).  What stays
unverifiable offline: diffusers' own file (not vendored by the reference, not installable) and the scaling_factor
0.18215 / key names of the published checkpoints (restated from the published config).  Only tests/ and bench.py's
checker legs may import this module."""
import torch
import torch.nn.functional as F


def _gn(sd, p, x, silu):
    y = F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)
    return F.silu(y) if silu else y


def _resnet(sd, p, x):
    """ResnetBlock2D(temb_channels=None): conv2(silu(norm2(conv1(silu(norm1(x)))))) + shortcut(x)"""
    h = F.conv2d(_gn(sd, p + ".norm1", x, True), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(_gn(sd, p + ".norm2", h, True), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def _attention(sd, p, x):
    """Attention(heads=1, dim_head=C, norm_num_groups=32, residual_connection=True, bias=True, upcast_softmax=True)"""
    n, c, h, w = x.shape
    t = _gn(sd, p + ".group_norm", x, False).reshape(n, c, h * w).transpose(1, 2)
    q = t @ sd[p + ".to_q.weight"].t() + sd[p + ".to_q.bias"]
    k = t @ sd[p + ".to_k.weight"].t() + sd[p + ".to_k.bias"]
    v = t @ sd[p + ".to_v.weight"].t() + sd[p + ".to_v.bias"]
    a = torch.softmax((q @ k.transpose(1, 2)).float() * c ** -0.5, dim=-1).to(v.dtype)
    o = (a @ v) @ sd[p + ".to_out.0.weight"].t() + sd[p + ".to_out.0.bias"]
    return x + o.transpose(1, 2).reshape(n, c, h, w)


def decode(sd, cfg, z):
    """z (N, 4, h, w) -> (N, 3, 8h, 8w); `sd` holds the diffusers-named tensors in the compute dtype."""
    x = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = F.conv2d(x, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    x = _resnet(sd, "decoder.mid_block.resnets.0", x)
    x = _attention(sd, "decoder.mid_block.attentions.0", x)
    x = _resnet(sd, "decoder.mid_block.resnets.1", x)
    nb = len(cfg["block_out_channels"])
    for i in range(nb):
        for j in range(cfg["layers_per_block"] + 1):
            x = _resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i != nb - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)
    x = _gn(sd, "decoder.conv_norm_out", x, True)
    return F.conv2d(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
