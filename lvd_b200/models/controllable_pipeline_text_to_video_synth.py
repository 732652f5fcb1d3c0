"""Mirror of the reference's `models/controllable_pipeline_text_to_video_synth.py::TextToVideoSDPipeline` for the hot
path: the denoising loop (`__call__`, reference lines 543-979; loop body 836-958) with backward guidance, GLIGEN
scheduling, classifier-free guidance and the scheduler step -- all on the native engine.

What is native here: everything inside the loop (guidance iterations, CFG UNet forward, fused CFG-combine +
DPM-Solver++ update).  What stays a caller-supplied torch module, exactly as in the reference: the CLIP text encoder /
tokenizer (`_encode_prompt`, reference :197-372; GLIGEN phrase embeddings :736-814) and the VAE (`decode_latents`
:374-400).  They are SURVEY.md section 8(f) "next" items; without them pass `prompt_embeds` / `negative_prompt_embeds`
and `output_type="latent"`.

Kept: the `__call__` signature (incl. the `lvd_gligen_*` spellings generation/lvd_gligen.py:126-128 uses),
`check_inputs` / `prepare_latents` ValueErrors (:425-533), `enable_fuser` (:535-539), `enable_vae_slicing`, `.scheduler`,
`.tokenizer`, `.guidance_models`, the `custom_latent_backward_guidance` hook (:572-574,827-831), callbacks.
"""
import warnings
from dataclasses import dataclass
from typing import Any

import torch

from ..scheduler import DPMSolverMultistepScheduler
from .pipelines import latent_backward_guidance, latent_backward_guidance_batch


@dataclass
class TextToVideoSDPipelineOutput:
    frames: Any


class TextToVideoSDPipeline:
    def __init__(self, vae=None, text_encoder=None, tokenizer=None, unet=None, scheduler=None):
        if unet is None:
            raise ValueError("TextToVideoSDPipeline needs a lvd_b200 UNet3DConditionModel")
        self.vae, self.text_encoder, self.tokenizer, self.unet = vae, text_encoder, tokenizer, unet
        self.scheduler = scheduler if scheduler is not None else DPMSolverMultistepScheduler()
        self.vae_scale_factor = 8
        self.guidance_models = None
        self._device = unet.device

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, unet=None, torch_dtype=None, revision=None,
                        trust_remote_code=None, **kw):
        """`TextToVideoSDPipeline.from_pretrained(model_key, unet=unet, torch_dtype=torch.float16)` as the runners call
        it (generation/lvd.py:39-46) for a LOCAL diffusers-format directory: `vae/` and `text_encoder/` load into the
        native mirrors, `tokenizer/` through transformers' CLIPTokenizer (host code), `scheduler/scheduler_config.json`
        is kept as `.scheduler.config` so that `DPMSolverMultistepScheduler.from_config(pipe.scheduler.config)` works.
        Components whose sub-directory is missing are left as None.  No hub access."""
        import json
        import os
        from ..text_encoder import CLIPTextModel
        from ..vae import AutoencoderKL
        from .unet_3d_condition import UNet3DConditionModel
        root = pretrained_model_name_or_path
        if not os.path.isdir(root):
            raise OSError(f"{root} is not a local directory: lvd_b200 loads local checkpoints only (no hub access)")
        if torch_dtype not in (None, torch.float16):
            raise NotImplementedError("lvd_b200 computes in fp16 (the dtype the reference runners select)")
        if unet is None:
            unet = UNet3DConditionModel.from_pretrained(root, subfolder="unet")
        vae = AutoencoderKL.from_pretrained(root, subfolder="vae") if os.path.isdir(os.path.join(root, "vae")) else None
        text_encoder = CLIPTextModel.from_pretrained(root, subfolder="text_encoder") \
            if os.path.isdir(os.path.join(root, "text_encoder")) else None
        tokenizer = None
        if os.path.isdir(os.path.join(root, "tokenizer")):
            from transformers import CLIPTokenizer
            tokenizer = CLIPTokenizer.from_pretrained(os.path.join(root, "tokenizer"))
        pipe = cls(vae=vae, text_encoder=text_encoder, tokenizer=tokenizer, unet=unet)
        sc = os.path.join(root, "scheduler", "scheduler_config.json")
        if os.path.isfile(sc):
            class _CheckpointScheduler:  # what the checkpoint ships (DDIM): only its config is consumed
                config = json.load(open(sc))
            pipe.checkpoint_scheduler_config = _CheckpointScheduler.config
            pipe.scheduler = DPMSolverMultistepScheduler.from_config(_CheckpointScheduler.config)
        return pipe

    # ---- small surface the runners touch (generation/lvd.py:39-53)
    def to(self, device):
        d = torch.device(device)
        if d.type != "cuda":
            raise RuntimeError("lvd_b200 has no CPU path")
        self._device = d if d.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.unet.to(self._device)
        for m in (self.vae, self.text_encoder):
            if m is not None:
                m.to(self._device)
        return self

    @property
    def device(self):
        return self._device

    def enable_vae_slicing(self):
        if self.vae is not None and hasattr(self.vae, "enable_slicing"):
            self.vae.enable_slicing()

    def enable_fuser(self, enabled=True):
        """reference :535-539"""
        for module in self.unet.modules():
            if type(module).__name__ == "GatedSelfAttentionDense":
                module.enabled = enabled

    # ---- input handling
    def check_inputs(self, prompt, height, width, callback_steps, negative_prompt=None, prompt_embeds=None,
                     negative_prompt_embeds=None):
        """reference :425-497"""
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_steps is None or not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type"
                             f" {type(callback_steps)}.")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make "
                             "sure to only forward one of the two.")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds`"
                             " undefined.")
        if prompt is not None and not isinstance(prompt, (str, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError("Cannot forward both `negative_prompt` and `negative_prompt_embeds`.")
        if prompt_embeds is not None and negative_prompt_embeds is not None:
            if prompt_embeds.shape != negative_prompt_embeds.shape:
                raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed "
                                 f"directly, but got: `prompt_embeds` {prompt_embeds.shape} != `negative_prompt_embeds`"
                                 f" {negative_prompt_embeds.shape}.")

    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt=None,
                       prompt_embeds=None, negative_prompt_embeds=None):
        """reference :197-372 -- returns cat([uncond, cond]).  Text encoding itself is the caller's CLIP module."""
        if prompt_embeds is None:
            if self.tokenizer is None or self.text_encoder is None:
                raise ValueError("no text encoder attached: pass `prompt_embeds` / `negative_prompt_embeds`")
            batch = [prompt] if isinstance(prompt, str) else list(prompt)

            def enc(texts):
                ids = self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length,
                                     truncation=True, return_tensors="pt").input_ids.to(device)
                with torch.no_grad():
                    return self.text_encoder(ids)[0]

            prompt_embeds = enc(batch)
            if do_classifier_free_guidance and negative_prompt_embeds is None:
                neg = [""] * len(batch) if negative_prompt is None else (
                    [negative_prompt] if isinstance(negative_prompt, str) else list(negative_prompt))
                if len(neg) != len(batch):
                    raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(neg)}, but `prompt`:"
                                     f" {prompt} has batch size {len(batch)}.")
                negative_prompt_embeds = enc(neg)
        prompt_embeds = prompt_embeds.to(device=device, dtype=torch.float16)
        prompt_embeds = prompt_embeds.repeat_interleave(num_images_per_prompt, dim=0)
        if do_classifier_free_guidance:
            if negative_prompt_embeds is None:
                raise ValueError("classifier-free guidance needs `negative_prompt_embeds` when no text encoder is attached")
            negative_prompt_embeds = negative_prompt_embeds.to(device=device, dtype=torch.float16)
            negative_prompt_embeds = negative_prompt_embeds.repeat_interleave(num_images_per_prompt, dim=0)
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds])
        return prompt_embeds

    def prepare_latents(self, batch_size, num_channels_latents, num_frames, height, width, dtype, device, generator,
                        latents=None):
        """reference :499-533"""
        shape = (batch_size, num_channels_latents, num_frames, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                             f"effective batch size of {batch_size}. Make sure the batch size matches the length of the"
                             " generators.")
        if latents is None:
            gdev = generator.device if generator is not None and not isinstance(generator, list) else device
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        else:
            if tuple(latents.shape) != shape:
                raise ValueError(f"Unexpected latents shape, got {tuple(latents.shape)}, expected {shape}")
            latents = latents.to(device=device, dtype=dtype)
        return latents * self.scheduler.init_noise_sigma

    def decode_latents(self, latents):
        """reference :374-400 (needs the caller's AutoencoderKL)."""
        if self.vae is None:
            raise ValueError("no VAE attached: use output_type='latent'")
        latents = 1 / self.vae.config.scaling_factor * latents
        b, c, f, h, w = latents.shape
        lat = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        image = self.vae.decode(lat).sample
        video = image[None, :].reshape((b, f, -1) + image.shape[2:]).permute(0, 2, 1, 3, 4)
        return video.float()

    def _gligen_inputs(self, gligen_phrases, gligen_boxes, repeat_batch, do_cfg, device):
        """reference :736-814 -- per-frame padded boxes / pooled phrase embeddings / masks, duplicated for CFG with the
        unconditional half masked out.  Identical phrase lists are encoded once (the reference re-encodes per frame)."""
        max_objs = 30
        cd = self.unet.config.cross_attention_dim
        boxes_all, emb_all, masks_all = [], [], []
        cache = {}
        for phrases_f, boxes_f in zip(gligen_phrases, gligen_boxes):
            if len(boxes_f) > max_objs:
                warnings.warn(f"More than {max_objs} objects found. Only first {max_objs} objects will be processed.",
                              FutureWarning)
                phrases_f, boxes_f = phrases_f[:max_objs], boxes_f[:max_objs]
            n = len(boxes_f)
            boxes = torch.zeros(max_objs, 4, device=device, dtype=torch.float16)
            emb = torch.zeros(max_objs, cd, device=device, dtype=torch.float16)
            masks = torch.zeros(max_objs, device=device, dtype=torch.float16)
            if n:
                if self.tokenizer is None or self.text_encoder is None:
                    raise ValueError("GLIGEN phrases need the text encoder (pooled phrase embeddings)")
                key = tuple(phrases_f)
                if key not in cache:
                    tok = self.tokenizer(list(phrases_f), padding=True, return_tensors="pt").to(device)
                    with torch.no_grad():
                        cache[key] = self.text_encoder(**tok).pooler_output.to(torch.float16)
                boxes[:n] = torch.tensor(boxes_f, device=device)
                emb[:n] = cache[key]
                masks[:n] = 1
            boxes = boxes.unsqueeze(0).expand(repeat_batch, -1, -1).clone()
            emb = emb.unsqueeze(0).expand(repeat_batch, -1, -1).clone()
            masks = masks.unsqueeze(0).expand(repeat_batch, -1).clone()
            if do_cfg:
                boxes, emb, masks = torch.cat([boxes] * 2), torch.cat([emb] * 2), torch.cat([masks] * 2)
                masks[:repeat_batch] = 0
            boxes_all.append(boxes)
            emb_all.append(emb)
            masks_all.append(masks)
        return {"boxes": torch.stack(boxes_all, dim=1).flatten(0, 1),
                "positive_embeddings": torch.stack(emb_all, dim=1).flatten(0, 1),
                "masks": torch.stack(masks_all, dim=1).flatten(0, 1)}

    # ---- the loop
    @torch.no_grad()
    def __call__(self, prompt=None, height=None, width=None, num_frames=16, num_inference_steps=50, guidance_scale=9.0,
                 gligen_scheduled_sampling_beta=0.3, gligen_phrases=None, gligen_boxes=None, negative_prompt=None,
                 eta=0.0, generator=None, latents=None, prompt_embeds=None, negative_prompt_embeds=None,
                 output_type="np", return_dict=True, callback=None, guidance_callback=None, callback_steps=1,
                 cross_attention_kwargs=None, backward_guidance_kwargs=None, aux_backward_guidance_kwargs=None,
                 guidance_type="main", return_guidance_saved_attn=False, custom_latent_backward_guidance=None,
                 backward_guidance_kwargs_custom=None, both_attn_and_custom=False, gligen_inputs=None,
                 lvd_gligen_scheduled_sampling_beta=None, lvd_gligen_phrases=None, lvd_gligen_boxes=None):
        # generation/lvd_gligen.py:126-128 spells these with an `lvd_` prefix
        if lvd_gligen_boxes is not None:
            gligen_boxes, gligen_phrases = lvd_gligen_boxes, lvd_gligen_phrases
        if lvd_gligen_scheduled_sampling_beta is not None:
            gligen_scheduled_sampling_beta = lvd_gligen_scheduled_sampling_beta
        if aux_backward_guidance_kwargs is not None or both_attn_and_custom:
            raise NotImplementedError("aux / combined custom guidance is not supported by lvd_b200")
        ss = self.unet.config.sample_size or 32
        height = height or ss * self.vae_scale_factor
        width = width or ss * self.vae_scale_factor
        num_images_per_prompt = 1
        self.check_inputs(prompt, height, width, callback_steps, negative_prompt, prompt_embeds, negative_prompt_embeds)
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None:
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        device = self._device
        do_cfg = guidance_scale > 1.0
        cross_attention_kwargs = dict(cross_attention_kwargs) if cross_attention_kwargs is not None else {}
        prompt_embeds = self._encode_prompt(prompt, device, num_images_per_prompt, do_cfg, negative_prompt,
                                            prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds)
        if do_cfg:
            uncond_prompt_embeds, cond_prompt_embeds = prompt_embeds.chunk(2)
        else:
            cond_prompt_embeds = prompt_embeds
        self.scheduler.set_timesteps(num_inference_steps, device=None)
        timesteps = self.scheduler.timesteps
        latents = self.prepare_latents(batch_size * num_images_per_prompt, self.unet.config.in_channels, num_frames,
                                       height, width, torch.float16, device, generator, latents)
        if gligen_inputs is not None:  # pre-computed boxes / positive_embeddings / masks (B*F or 2*B*F rows)
            cross_attention_kwargs["gligen"] = gligen_inputs
        elif gligen_boxes:
            cross_attention_kwargs["gligen"] = self._gligen_inputs(gligen_phrases, gligen_boxes,
                                                                   batch_size * num_images_per_prompt, do_cfg, device)
        num_grounding_steps = int(gligen_scheduled_sampling_beta * len(timesteps))
        self.enable_fuser(True)
        loss_attn = torch.tensor(10000.0)
        backward_guidance = custom_latent_backward_guidance if custom_latent_backward_guidance \
            else latent_backward_guidance
        # Extension (the reference guides one prompt per call, its loss asserts n_frames == len(boxes)): with several
        # prompts in the batch, `bboxes` / `object_positions` of backward_guidance_kwargs hold ONE LAYOUT PER PROMPT and
        # the prompts are guided in lock-step by one batched pass per iteration (latent_backward_guidance_batch)
        batched_guidance = backward_guidance_kwargs is not None and latents.shape[0] > 1 and \
            custom_latent_backward_guidance is None
        if batched_guidance:
            if len(backward_guidance_kwargs["bboxes"]) != latents.shape[0]:
                raise ValueError("batched guidance: backward_guidance_kwargs['bboxes'] / ['object_positions'] must hold "
                                 f"one layout per prompt ({latents.shape[0]})")
            if return_guidance_saved_attn:
                raise NotImplementedError("return_guidance_saved_attn with several prompts per call")
            loss_attn = torch.full((latents.shape[0],), 10000.0)
        fused_step = isinstance(self.scheduler, DPMSolverMultistepScheduler)
        for i, t in enumerate(timesteps):
            if i == num_grounding_steps:
                self.enable_fuser(False)
            assert latents.shape[1] == 4, f"latent channel mismatch: {latents.shape}"
            if backward_guidance_kwargs is not None:
                if guidance_type != "main":
                    raise ValueError(f"Unsupported guidance type: {guidance_type}")
                if batched_guidance:
                    ret = latent_backward_guidance_batch(self.scheduler, self.unet, cond_prompt_embeds, latents=latents,
                                                         index=i, t=t, loss=loss_attn, **backward_guidance_kwargs)
                else:
                    ret = backward_guidance(self.scheduler, self.unet, cond_prompt_embeds, latents=latents, index=i,
                                            t=t, loss=loss_attn, return_saved_attn=return_guidance_saved_attn,
                                            **backward_guidance_kwargs)
                latents, loss_attn = ret[0], ret[1]
                if guidance_callback is not None and i % callback_steps == 0:
                    guidance_callback(i, t, latents, None, ret[2] if return_guidance_saved_attn else None)
            latent_model_input = torch.cat([latents] * 2) if do_cfg else latents
            latent_model_input = self.scheduler.scale_model_input(latent_model_input, t)
            noise_pred = self.unet(latent_model_input, t, encoder_hidden_states=prompt_embeds,
                                   cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
            if fused_step:
                # CFG combine (:926-930) + scheduler.step (:941-943) in one kernel; the permutes around the step in
                # the reference (:933-950) only reshape an element-wise update
                if do_cfg:
                    half = noise_pred.shape[0] // 2
                    latents = self.scheduler.step_cfg(noise_pred[:half].contiguous(), noise_pred[half:].contiguous(),
                                                      latents.contiguous(), guidance_scale)
                else:
                    latents = self.scheduler.step_cfg(noise_pred.contiguous(), None, latents.contiguous(), 1.0)
            else:
                if do_cfg:
                    noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)
                    noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)
                latents = self.scheduler.step(noise_pred, t, latents).prev_sample
            if callback is not None and i % callback_steps == 0:
                callback(i, t, latents)
        if output_type == "latent":
            out = latents
        elif output_type == "np" and hasattr(self.vae, "decode_video_np"):
            # native VAE (lvd_b200.vae.AutoencoderKL): decode + tensor2vid's denormalise / channels-last in one pass.
            # (The reference decodes twice, :963 and :969, and keeps the second result; once gives the same frames.)
            out = self.vae.decode_video_np(latents)
        else:
            video = self.decode_latents(latents)
            out = video if output_type == "pt" else video.permute(0, 2, 3, 4, 1).clamp(-1, 1).add(1).div(2).cpu().numpy()
        if not return_dict:
            return (out,)
        return TextToVideoSDPipelineOutput(frames=out)
