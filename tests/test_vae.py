"""VAE decode -- the native decoder (lvd_b200/vae.py) against the plain-PyTorch
restatement of diffusers' AutoencoderKL.decode (oracle/vae_ref.py; parity unpinned: diffusers is not vendored).
Tolerance as for the UNet: err(engine fp16 vs oracle fp32) <= max(3e-3, 1.5 x err(torch fp16 vs oracle fp32))."""
import pytest
import torch

from conftest import rel_err


def test_vae_spec_matches_published_decoder_size():
    from lvd_b200 import synthetic
    spec = synthetic.vae_param_spec(synthetic.VAE_SD)
    n = sum(torch.Size(s).numel() for s in spec.values())
    assert len(spec) == 140 and n == 49_490_199  # SD-1.x AutoencoderKL: post_quant_conv + decoder
    assert spec["decoder.mid_block.attentions.0.to_q.weight"] == (512, 512)
    assert spec["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"] == (256, 512, 1, 1)
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in spec


def test_vae_oracle_shapes_cpu():
    from lvd_b200 import synthetic
    from oracle import vae_ref
    cfg = synthetic.VAE_TINY
    sd = synthetic.random_vae_state_dict(cfg, seed=1)
    y = vae_ref.decode(sd, cfg, torch.randn(2, 4, 8, 8))
    assert y.shape == (2, 3, 64, 64) and torch.isfinite(y).all()


def _janus_decoder(cfg, sd):
    """The CompVis / taming-transformers VAE decoder as shipped in the installed `transformers` (JanusVQVAEDecoder:
    conv_in -> mid (block_1, attn_1, block_2) -> up levels of num_res_blocks + 1 ResnetBlocks with nearest-2x + conv
    up-sampling -> norm_out -> swish -> conv_out) -- the architecture the SD-1.x VAE checkpoints were trained with and
    that diffusers' AutoencoderKL re-implements under other parameter names.  Janus adds attention blocks to the
    lowest-resolution up level; the SD VAE has none there (attn_resolutions = []), so they are removed.  Weights are the
    diffusers-named tensors of `sd`, renamed."""
    from torch import nn
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    from transformers.models.janus.modeling_janus import JanusVQVAEDecoder
    boc = list(cfg["block_out_channels"])
    base = boc[0]
    jc = JanusVQVAEConfig(base_channels=base, channel_multiplier=[c // base for c in boc],
                          num_res_blocks=cfg["layers_per_block"], latent_channels=cfg["latent_channels"],
                          out_channels=cfg["out_channels"], dropout=0.0)
    dec = JanusVQVAEDecoder(jc).eval()
    for lvl in dec.up:
        lvl.attn = nn.ModuleList()
    ren = {}

    def res(src, dst):
        for a in ("norm1", "conv1", "norm2", "conv2"):
            for t in ("weight", "bias"):
                ren[f"{dst}.{a}.{t}"] = sd[f"{src}.{a}.{t}"]
        if f"{src}.conv_shortcut.weight" in sd:
            ren[f"{dst}.nin_shortcut.weight"] = sd[f"{src}.conv_shortcut.weight"]
            ren[f"{dst}.nin_shortcut.bias"] = sd[f"{src}.conv_shortcut.bias"]

    for t in ("weight", "bias"):
        ren[f"conv_in.{t}"] = sd[f"decoder.conv_in.{t}"]
        ren[f"norm_out.{t}"] = sd[f"decoder.conv_norm_out.{t}"]
        ren[f"conv_out.{t}"] = sd[f"decoder.conv_out.{t}"]
        ren[f"mid.attn_1.norm.{t}"] = sd[f"decoder.mid_block.attentions.0.group_norm.{t}"]
    for a, b in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
        ren[f"mid.attn_1.{a}.weight"] = sd[f"decoder.mid_block.attentions.0.{b}.weight"][:, :, None, None]
        ren[f"mid.attn_1.{a}.bias"] = sd[f"decoder.mid_block.attentions.0.{b}.bias"]
    res("decoder.mid_block.resnets.0", "mid.block_1")
    res("decoder.mid_block.resnets.1", "mid.block_2")
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"] + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", f"up.{i}.block.{j}")
        if i != len(boc) - 1:
            for t in ("weight", "bias"):
                ren[f"up.{i}.upsample.conv.{t}"] = sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.{t}"]
    missing, unexpected = dec.load_state_dict(ren, strict=True), None
    return dec


def test_vae_oracle_pinned_against_transformers_compvis_decoder():
    """Pins oracle/vae_ref.py (the restatement of diffusers' AutoencoderKL.decode the CUDA decoder is checked against)
    to an INDEPENDENT public implementation of the same decoder that is installed offline: transformers'
    JanusVQVAEDecoder, a port of the CompVis / taming-transformers `Decoder` (the code the SD-1.x VAE was trained with;
    diffusers' AutoencoderKL is a renaming of it).  Same weights (every tensor of the decoder is consumed: strict load),
    same latents -> same images to fp32 rounding, at the tiny and at the published SD-1.x width."""
    from lvd_b200 import synthetic
    from oracle import vae_ref
    for cfg, hw in ((synthetic.VAE_TINY, (8, 12)), (synthetic.VAE_SD, (4, 6))):
        sd = synthetic.random_vae_state_dict(cfg, seed=3)
        dec = _janus_decoder(cfg, sd)
        z = torch.randn((2, 4) + hw, generator=torch.Generator().manual_seed(4))
        with torch.no_grad():
            want = dec(torch.nn.functional.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))
            got = vae_ref.decode(sd, cfg, z)
        assert got.shape == want.shape == (2, 3, 8 * hw[0], 8 * hw[1])
        err = ((got - want).abs().max() / want.abs().max()).item()
        assert err < 2e-5, err
    # the decoder's tensor inventory equals the published SD-1.x VAE decoder's (diffusers names): 138 decoder tensors
    # + post_quant_conv, 49 490 199 parameters
    spec = synthetic.vae_param_spec(synthetic.VAE_SD)
    n = sum(int(torch.tensor(s).prod()) if len(s) else 1 for s in spec.values())
    assert len(spec) == 140 and n == 49_490_199


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w,slice_images", [(3, 8, 16, 0), (2, 16, 16, 1), (5, 8, 8, 2)])
def test_vae_decode_matches_oracle(cuda, n, h, w, slice_images):
    from lvd_b200 import synthetic
    from lvd_b200.vae import AutoencoderKL
    from oracle import vae_ref
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = synthetic.VAE_TINY
    sd = synthetic.random_vae_state_dict(cfg, seed=3)
    sd16 = {k: v.half() for k, v in sd.items()}
    vae = AutoencoderKL(block_out_channels=cfg["block_out_channels"]).to(cuda)
    vae.load_state_dict(dict(sd16, **{"encoder.conv_in.weight": torch.zeros(1)}))  # encoder tensors are ignored
    if slice_images:
        vae.enable_slicing(slice_images)
    z = torch.randn((n, 4, h, w), generator=torch.Generator().manual_seed(4)).half().to(cuda)
    y = vae.decode(z).sample
    with torch.no_grad():
        y32 = vae_ref.decode({k: v.float().to(cuda) for k, v in sd16.items()}, cfg, z.float())
        y16 = vae_ref.decode({k: v.to(cuda) for k, v in sd16.items()}, cfg, z)
    e, e16 = rel_err(y, y32), rel_err(y16, y32)
    print(f"\n[vae tiny n={n} {h}x{w}] engine-vs-fp32 {e:.3e}   torch-fp16-vs-fp32 {e16:.3e}")
    assert y.shape == (n, 3, 8 * h, 8 * w) and torch.isfinite(y.float()).all()
    assert e <= max(3e-3, 1.5 * e16)


@pytest.mark.gpu
def test_pipeline_decodes_through_native_vae(cuda):
    """decode_latents (controllable_pipeline_text_to_video_synth.py:374-400) + the np output conversion on the mirror."""
    from lvd_b200 import synthetic
    from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline
    from lvd_b200.models.unet_3d_condition import UNet3DConditionModel
    from lvd_b200.vae import AutoencoderKL
    from oracle import vae_ref, weights
    cfg = weights.make_cfg("tiny")
    unet = UNet3DConditionModel(sample_size=16, block_out_channels=cfg["block_out_channels"],
                                cross_attention_dim=cfg["cross_attention_dim"])
    unet.load_state_dict(weights.random_state_dict(cfg, seed=7))
    vcfg = synthetic.VAE_TINY
    vsd = {k: v.half() for k, v in synthetic.random_vae_state_dict(vcfg, seed=8).items()}
    vae = AutoencoderKL(block_out_channels=vcfg["block_out_channels"])
    vae.load_state_dict(vsd)
    pipe = TextToVideoSDPipeline(unet=unet, vae=vae).to("cuda")
    g = torch.Generator().manual_seed(9)
    F, h, w = 8, 8, 16
    emb = torch.randn((2, 77, cfg["cross_attention_dim"]), generator=g)
    lat0 = torch.randn((1, 4, F, h, w), generator=g)
    lat = pipe(prompt_embeds=emb[1:], negative_prompt_embeds=emb[:1], height=h * 8, width=w * 8, num_frames=F,
               num_inference_steps=2, latents=lat0.clone(), output_type="latent").frames
    frames = pipe(prompt_embeds=emb[1:], negative_prompt_embeds=emb[:1], height=h * 8, width=w * 8, num_frames=F,
                  num_inference_steps=2, latents=lat0.clone(), output_type="np").frames
    assert frames.shape == (1, F, h * 8, w * 8, 3) and frames.dtype.name == "float32"
    assert frames.min() >= 0.0 and frames.max() <= 1.0
    z = (lat.float() / vcfg["scaling_factor"]).permute(0, 2, 1, 3, 4).reshape(F, 4, h, w)
    with torch.no_grad():
        ref = vae_ref.decode({k: v.float().to(cuda) for k, v in vsd.items()}, vcfg, z.to(cuda))
    ref = ((ref.clamp(-1, 1) + 1) / 2).permute(0, 2, 3, 1).cpu().numpy()[None]
    assert abs(frames - ref).max() <= 1e-2
