"""Denoising loop mirror (TextToVideoSDPipeline.__call__) on the GPU against a loop assembled from the oracle pieces
(oracle UNet fp32 + reference loss / autograd guidance + DPM-Solver++ restatement).  Errors compound over steps, so
the gate is on the final latents after a few steps: max|a-b| <= 2e-2 max|b| (reported per run)."""
import pytest
import torch

from conftest import rel_err


def test_check_inputs_errors_match_reference_messages():
    from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline

    class _U:
        device = None

        class config:
            sample_size = 32
            in_channels = 4

    p = TextToVideoSDPipeline(unet=_U())
    with pytest.raises(ValueError, match="divisible by 8"):
        p.check_inputs("a", 30, 64, 1)
    with pytest.raises(ValueError, match="positive integer"):
        p.check_inputs("a", 64, 64, 0)
    with pytest.raises(ValueError, match="Cannot forward both"):
        p.check_inputs("a", 64, 64, 1, prompt_embeds=torch.zeros(1))
    with pytest.raises(ValueError, match="Provide either"):
        p.check_inputs(None, 64, 64, 1)
    with pytest.raises(ValueError, match="Unexpected latents shape"):
        p.prepare_latents(1, 4, 8, 64, 64, torch.float16, "cpu", None, latents=torch.zeros(1, 4, 8, 9, 8))


@pytest.mark.gpu
def test_pipeline_guided_loop_matches_oracle_loop(cuda):
    from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline
    from lvd_b200.models.unet_3d_condition import UNet3DConditionModel
    from oracle import guidance_ref as G
    from oracle import scheduler_ref
    from oracle import unet3d_ref as R
    from oracle import weights
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = weights.make_cfg("tiny")
    sd = weights.random_state_dict(cfg, seed=31)
    unet = UNet3DConditionModel(sample_size=16, block_out_channels=cfg["block_out_channels"],
                                cross_attention_dim=cfg["cross_attention_dim"])
    unet.load_state_dict(sd)
    pipe = TextToVideoSDPipeline(unet=unet).to("cuda")
    gen = torch.Generator().manual_seed(32)
    F, h, w = 8, 16, 16
    lat0 = torch.randn((1, 4, F, h, w), generator=gen)
    emb = torch.randn((2, 77, 128), generator=gen)
    bboxes = [[(0.1 + 0.04 * f, 0.2, 0.5 + 0.04 * f, 0.8) for f in range(F)]]
    positions = [[2, 3]]
    keys = R.GUIDANCE_KEYS_DEFAULT
    hp = dict(fg_top_p=0.25, bg_top_p=0.25, fg_weight=1.0, bg_weight=2.0, base_attn_dim=(h, w))
    steps, gsteps = 4, 2
    out = pipe(prompt_embeds=emb[1:], negative_prompt_embeds=emb[:1], height=h * 8, width=w * 8, num_frames=F,
               num_inference_steps=steps, guidance_scale=9.0, latents=lat0.clone(), output_type="latent",
               backward_guidance_kwargs=dict(bboxes=bboxes, object_positions=positions, guidance_attn_keys=keys,
                                             loss_scale=2.5, loss_threshold=0.0, max_iter=1, max_index_step=gsteps,
                                             **hp)).frames
    # oracle loop (fp32, same fp16-rounded weights / inputs)
    sd32 = {k: v.half().float().to(cuda) for k, v in sd.items()}
    sched = scheduler_ref.DPMSolverRef()
    sched.set_timesteps(steps)
    lat = lat0.half().float().to(cuda)
    e32 = emb.half().float().to(cuda)
    for i, t in enumerate(sched.timesteps):
        if i < gsteps:
            lat, _, _, _ = G.guidance_step(sd32, cfg, lat, int(t), e32[1:], bboxes, positions,
                                           sched.alphas_cumprod[int(t)].to(cuda), guidance_attn_keys=keys,
                                           loss_scale=2.5, **hp)
        with torch.no_grad():
            eps = R.unet_forward(sd32, cfg, torch.cat([lat, lat]), int(t), e32)
            lat = sched.step(scheduler_ref.cfg_combine(eps[:1], eps[1:], 9.0), lat)
    e = rel_err(out, lat)
    print(f"\n[pipeline {steps} steps, {gsteps} guided] final latents engine-vs-oracle: {e:.3e}")
    assert out.shape == (1, 4, F, h, w) and torch.isfinite(out.float()).all()
    assert e <= 2e-2


@pytest.mark.gpu
def test_from_pretrained_local_checkpoint_runs_like_the_runner(cuda, tmp_path):
    """generation/lvd.py:39-48 against a local diffusers-format directory: UNet3DConditionModel.from_pretrained(
    model_key, subfolder="unet").to(fp16); TextToVideoSDPipeline.from_pretrained(model_key, unet=unet, torch_dtype=fp16);
    scheduler = DPMSolverMultistepScheduler.from_config(pipe.scheduler.config); pipe.to("cuda"); enable_vae_slicing();
    a call with prompt embeddings returns np frames.  Hub names (non-local paths) raise."""
    import json
    import os

    from safetensors.torch import save_file
    from lvd_b200 import synthetic
    from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline
    from lvd_b200.models.unet_3d_condition import UNet3DConditionModel
    from lvd_b200.scheduler import DPMSolverMultistepScheduler
    from lvd_b200.text_encoder import clip_param_spec
    from oracle import weights
    root = str(tmp_path / "ckpt")
    ucfg = weights.make_cfg("tiny")
    os.makedirs(root + "/unet")
    json.dump({"_class_name": "UNet3DConditionModel", "sample_size": 16, "block_out_channels": ucfg["block_out_channels"],
               "cross_attention_dim": ucfg["cross_attention_dim"], "attention_head_dim": 64},
              open(root + "/unet/config.json", "w"))
    save_file({k: v.half().contiguous() for k, v in weights.random_state_dict(ucfg, seed=51).items()},
              root + "/unet/diffusion_pytorch_model.safetensors")
    vcfg = synthetic.VAE_TINY
    os.makedirs(root + "/vae")
    json.dump({"_class_name": "AutoencoderKL", "block_out_channels": vcfg["block_out_channels"],
               "scaling_factor": 0.18215, "sample_size": 256}, open(root + "/vae/config.json", "w"))
    save_file({k: v.half().contiguous() for k, v in synthetic.random_vae_state_dict(vcfg, seed=52).items()},
              root + "/vae/diffusion_pytorch_model.safetensors")
    tcfg = dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                max_position_embeddings=77, hidden_act="gelu", eos_token_id=999)
    os.makedirs(root + "/text_encoder")
    json.dump(dict(tcfg, architectures=["CLIPTextModel"]), open(root + "/text_encoder/config.json", "w"))
    g = torch.Generator().manual_seed(53)
    save_file({k: (0.05 * torch.randn(s, generator=g)).half() for k, s in clip_param_spec(tcfg).items()},
              root + "/text_encoder/model.safetensors")
    os.makedirs(root + "/scheduler")
    json.dump({"_class_name": "DDIMScheduler", "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear",
               "num_train_timesteps": 1000, "steps_offset": 1, "timestep_spacing": "leading",
               "prediction_type": "epsilon", "clip_sample": False}, open(root + "/scheduler/scheduler_config.json", "w"))

    unet = UNet3DConditionModel.from_pretrained(root, subfolder="unet").to(torch.float16)
    pipe = TextToVideoSDPipeline.from_pretrained(root, unet=unet, torch_dtype=torch.float16)
    pipe.scheduler = DPMSolverMultistepScheduler.from_config(pipe.scheduler.config)
    pipe.to("cuda")
    pipe.enable_vae_slicing()
    pipe.guidance_models = None
    assert pipe.scheduler.config["timestep_spacing"] == "leading" and pipe.scheduler.config["steps_offset"] == 1
    assert pipe.vae is not None and pipe.text_encoder is not None and pipe.tokenizer is None
    emb = pipe.text_encoder(torch.randint(1, 998, (2, 77), generator=g))[0]
    frames = pipe(prompt_embeds=emb[1:], negative_prompt_embeds=emb[:1], height=64, width=128, num_frames=8,
                  num_inference_steps=2, output_type="np",
                  generator=torch.Generator().manual_seed(54)).frames
    assert frames.shape == (1, 8, 64, 128, 3) and (frames >= 0).all() and (frames <= 1).all()
    with pytest.raises(OSError, match="local"):
        TextToVideoSDPipeline.from_pretrained("cerspense/zeroscope_v2_576w", unet=unet)
    with pytest.raises(OSError, match="local"):
        UNet3DConditionModel.from_pretrained("cerspense/zeroscope_v2_576w", subfolder="unet")


@pytest.mark.gpu
def test_pipeline_multi_prompt_batch_matches_single_prompt_calls(cuda):
    """B = 2 prompts through ONE pipeline call (CFG batch 2B per UNet call, guidance in lock-step, per-prompt layouts)
    against the same prompts generated one per call."""
    from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline
    from lvd_b200.models.unet_3d_condition import UNet3DConditionModel
    from lvd_b200.scheduler import DPMSolverMultistepScheduler
    from oracle import weights
    cfg = weights.make_cfg("tiny")
    unet = UNet3DConditionModel(sample_size=16, block_out_channels=cfg["block_out_channels"],
                                cross_attention_dim=cfg["cross_attention_dim"])
    unet.to(cuda)
    unet.load_state_dict({k: v.half() for k, v in weights.random_state_dict(cfg, seed=91).items()})
    g = torch.Generator().manual_seed(92)
    B, F, h, w = 2, 8, 16, 16
    lat = torch.randn((B, 4, F, h, w), generator=g).half()
    emb = torch.randn((B, 77, 128), generator=g).half()
    neg = torch.randn((1, 77, 128), generator=g).half()
    layouts = [([[(0.1 + 0.04 * f, 0.2, 0.6, 0.7) for f in range(F)]], [[2]]),
               ([[(0.5, 0.1, 0.9, 0.6)] * F, [(0.0, 0.5, 0.4, 1.0)] * F], [[3], [5, 6]])]
    gk = dict(loss_scale=5.0, loss_threshold=0.0, max_iter=1, max_index_step=3, fg_top_p=0.75, bg_top_p=0.75,
              fg_weight=1.0, bg_weight=4.0, com_loss_scale=0.03, base_attn_dim=(h, w),
              guidance_attn_keys=[("down", 1, 0, 0), ("down", 2, 1, 0), ("up", 1, 1, 0)])
    common = dict(height=h * 8, width=w * 8, num_frames=F, num_inference_steps=5, guidance_scale=7.0,
                  output_type="latent")

    def pipe():
        return TextToVideoSDPipeline(unet=unet, scheduler=DPMSolverMultistepScheduler()).to(cuda)

    both = pipe()(prompt_embeds=emb, negative_prompt_embeds=neg.expand(B, -1, -1), latents=lat.clone(),
                  backward_guidance_kwargs=dict(gk, bboxes=[l[0] for l in layouts],
                                                object_positions=[l[1] for l in layouts]), **common).frames
    assert both.shape == (B, 4, F, h, w)
    for i in range(B):
        one = pipe()(prompt_embeds=emb[i:i + 1], negative_prompt_embeds=neg, latents=lat[i:i + 1].clone(),
                     backward_guidance_kwargs=dict(gk, bboxes=layouts[i][0], object_positions=layouts[i][1]),
                     **common).frames
        e = ((both[i].float() - one[0].float()).abs().max() / one.float().abs().max()).item()
        print(f"\\n[pipeline batch prompt {i}] batched vs single call: {e:.3e}")
        assert e <= 1e-2
