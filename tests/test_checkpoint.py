"""Host-side loading logic (no GPU): local diffusers-format directories, state_dict validation of the VAE / text
encoder mirrors (same error style as torch's load_state_dict), no hub access."""
import json
import os

import pytest
import torch


def test_load_component_errors_and_formats(tmp_path):
    from safetensors.torch import save_file
    from lvd_b200.checkpoint import init_kwargs, load_component
    root = str(tmp_path)
    with pytest.raises(OSError, match="local checkpoints only"):
        load_component(root, "unet")
    os.makedirs(root + "/vae")
    json.dump({"block_out_channels": [64, 64], "scaling_factor": 0.5, "unknown_key": 1}, open(root + "/vae/config.json", "w"))
    with pytest.raises(OSError, match="no weight file"):
        load_component(root, "vae")
    torch.save({"a": torch.ones(2)}, root + "/vae/diffusion_pytorch_model.bin")
    cfg, sd = load_component(root, "vae")
    assert cfg["scaling_factor"] == 0.5 and torch.equal(sd["a"], torch.ones(2))
    save_file({"a": torch.zeros(2)}, root + "/vae/diffusion_pytorch_model.safetensors")  # safetensors wins over .bin
    _, sd = load_component(root, "vae")
    assert torch.equal(sd["a"], torch.zeros(2))

    class M:
        def __init__(self, block_out_channels=(1,), scaling_factor=1.0):
            pass

    assert init_kwargs(M, cfg) == {"block_out_channels": [64, 64], "scaling_factor": 0.5}


def test_vae_mirror_validates_state_dict():
    from lvd_b200 import synthetic
    from lvd_b200.vae import AutoencoderKL
    cfg = synthetic.VAE_TINY
    sd = synthetic.random_vae_state_dict(cfg, seed=0)
    vae = AutoencoderKL(block_out_channels=cfg["block_out_channels"])
    vae._device = None  # host-only: validation without building the engine
    vae.load_state_dict(dict(sd, **{"encoder.conv_in.weight": torch.zeros(1), "quant_conv.bias": torch.zeros(1)}))
    bad = dict(sd)
    del bad["decoder.conv_out.bias"]
    with pytest.raises(RuntimeError, match="missing"):
        vae.load_state_dict(bad)
    with pytest.raises(RuntimeError, match="unexpected"):
        vae.load_state_dict(dict(sd, extra=torch.zeros(1)))
    with pytest.raises(RuntimeError, match="size mismatch"):
        vae.load_state_dict(dict(sd, **{"decoder.conv_out.bias": torch.zeros(7)}))
    # published SD-1.x / Zeroscope VAE checkpoints carry the deprecated AttentionBlock names (query / key / value /
    # proj_attn, the oldest as 1x1 convs): converted on load like diffusers 0.27.2 does
    legacy = {}
    ren = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}
    for k, v in sd.items():
        for new, old in ren.items():
            if f".attentions.0.{new}." in k:
                k = k.replace(f".attentions.0.{new}.", f".attentions.0.{old}.")
                if k.endswith(".weight"):
                    v = v[:, :, None, None]
                break
        legacy[k] = v
    assert any(".query." in k for k in legacy) and not any(".to_q." in k for k in legacy)
    vae.load_state_dict(legacy)
    for k, v in sd.items():
        assert torch.equal(vae._state[k], v), k
    with pytest.raises(RuntimeError, match="no weights"):
        AutoencoderKL().decode(torch.zeros(1, 4, 8, 8))
    with pytest.raises(NotImplementedError):
        AutoencoderKL(norm_num_groups=16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        AutoencoderKL().to("cpu")


def test_text_encoder_mirror_validates_state_dict_and_pools_at_eos():
    from lvd_b200.text_encoder import CLIPTextModel, clip_param_spec
    kw = dict(vocab_size=100, hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=1,
              max_position_embeddings=77)
    enc = CLIPTextModel(**kw, eos_token_id=99)
    enc._device = None
    sd = {k: torch.zeros(s) for k, s in clip_param_spec(kw).items()}
    enc.load_state_dict(dict(sd, **{"text_model.embeddings.position_ids": torch.arange(77)}))  # legacy buffer accepted
    with pytest.raises(RuntimeError, match="unexpected"):
        enc.load_state_dict(dict(sd, bogus=torch.zeros(1)))
    with pytest.raises(RuntimeError, match="size mismatch"):
        enc.load_state_dict(dict(sd, **{"text_model.final_layer_norm.bias": torch.zeros(3)}))
    with pytest.raises(RuntimeError, match="no weights"):
        CLIPTextModel(**kw)(torch.zeros((1, 77), dtype=torch.int64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        CLIPTextModel(**kw).to("cpu")


def test_pipeline_from_pretrained_rejects_hub_names():
    from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline
    with pytest.raises(OSError, match="local"):
        TextToVideoSDPipeline.from_pretrained("cerspense/zeroscope_v2_576w", unet=object())
