"""VAE decode on the native kernels -- SURVEY.md section 8(f) rank 1.

`TextToVideoSDPipeline.decode_latents` (controllable_pipeline_text_to_video_synth.py:374-400) calls
`self.vae.decode(latents).sample` on (B*F, 4, h, w) latents; the VAE is diffusers 0.27.2 `AutoencoderKL` (not vendored
by the reference; architecture restated in oracle/vae_ref.py).  The decoder is convolutions, GroupNorm+SiLU, nearest
up-sampling and one single-head 512-wide attention, so it runs on the UNet's kernels (`lvd_gemm` implicit-GEMM convs,
streaming GroupNorm, `lvd_upsample2x`, conv_in / conv_out ends) plus `lvd_softmax_rows` / `lvd_pointwise_small`:

  post_quant_conv (1x1, 4->4) -> conv_in (3x3, 4->512) -> mid [resnet, attention, resnet]
  -> 4 up blocks of 3 resnets (512, 512, 256, 128 channels; nearest-2x + conv after the first three)
  -> GroupNorm + SiLU -> conv_out (3x3, 128->3)

Activations are channels-last row matrices like the UNet's.  `AutoencoderKL` below mirrors the slice of the diffusers
class the pipeline touches: `.config.scaling_factor`, `.decode(z).sample`, `.enable_slicing()`, `.to()`,
`load_state_dict` (diffusers names; encoder / quant_conv tensors are accepted and ignored).  No CPU path.
"""
from dataclasses import dataclass

import torch

from . import _lib, ops, repack, synthetic
from .engine import LayerWeights, _h


class VAEDecoderEngine:
    def __init__(self, cfg, state_dict, device):
        _lib.lib()
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("lvd_b200 VAE decoder needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = dict(cfg)
        self.dev = dev = torch.device(device)
        sd = state_dict
        boc = list(cfg["block_out_channels"])
        if any(c % 64 for c in boc):
            raise NotImplementedError("VAE decoder channels must be multiples of 64")

        def g(name):
            return _h(sd[name], dev)

        def norm(p):
            return LayerWeights(w=g(p + ".weight"), b=g(p + ".bias"))

        def conv(p):
            return LayerWeights(w=repack.conv_weight_krsc(g(p + ".weight")), b=g(p + ".bias"))

        def resnet(p):
            L = LayerWeights(norm1=norm(p + ".norm1"), conv1=conv(p + ".conv1"), norm2=norm(p + ".norm2"),
                             conv2=conv(p + ".conv2"))
            if p + ".conv_shortcut.weight" in sd:
                w = g(p + ".conv_shortcut.weight")
                L["shortcut"] = LayerWeights(w=w.reshape(w.shape[0], w.shape[1]).contiguous(),
                                             b=g(p + ".conv_shortcut.bias"))
            return L

        w = LayerWeights()
        pq = g("post_quant_conv.weight")
        w["post_quant"] = LayerWeights(w=pq.reshape(pq.shape[0], pq.shape[1]).contiguous(), b=g("post_quant_conv.bias"))
        w["conv_in"] = LayerWeights(w_pad=repack.conv_in_weight_padded(g("decoder.conv_in.weight")),
                                    b=g("decoder.conv_in.bias"))
        w["mid0"] = resnet("decoder.mid_block.resnets.0")
        w["mid1"] = resnet("decoder.mid_block.resnets.1")
        a = "decoder.mid_block.attentions.0"
        c = boc[-1]
        scale = float(c) ** -0.5  # one head of width c; folded into the query projection (scores = scale * Q K^T)
        wq = (sd[a + ".to_q.weight"].float() * scale)
        bq = (sd[a + ".to_q.bias"].float() * scale)
        w["attn"] = LayerWeights(
            norm=norm(a + ".group_norm"),
            qkv=LayerWeights(w=_h(torch.cat([wq, sd[a + ".to_k.weight"].float(), sd[a + ".to_v.weight"].float()], 0), dev),
                             b=_h(torch.cat([bq, sd[a + ".to_k.bias"].float(), sd[a + ".to_v.bias"].float()], 0), dev)),
            out=LayerWeights(w=g(a + ".to_out.0.weight"), b=g(a + ".to_out.0.bias")), c=c)
        ups = []
        for i in range(len(boc)):
            blk = LayerWeights(resnets=[resnet(f"decoder.up_blocks.{i}.resnets.{j}")
                                        for j in range(cfg["layers_per_block"] + 1)])
            if f"decoder.up_blocks.{i}.upsamplers.0.conv.weight" in sd:
                blk["up"] = conv(f"decoder.up_blocks.{i}.upsamplers.0.conv")
            ups.append(blk)
        w["ups"] = ups
        w["norm_out"] = norm("decoder.conv_norm_out")
        wko = repack.conv_weight_krsc(g("decoder.conv_out.weight"))
        wpad, bpad = repack.pad_out_channels(wko, g("decoder.conv_out.bias"))
        w["conv_out"] = LayerWeights(w_pad=wpad, b_pad=bpad, cout=wko.shape[0])
        self.w = w

    # ---- blocks (diffusers ResnetBlock2D without time embedding, eps 1e-6; Attention with residual_connection)
    def _resnet(self, L, x, n_img, h, w_):
        t = ops.groupnorm(x, n_img, L.norm1.w, L.norm1.b, 1e-6, silu=True)
        t = ops.conv2d_3x3(t, n_img, h, w_, L.conv1.w, L.conv1.b)
        t = ops.groupnorm(t, n_img, L.norm2.w, L.norm2.b, 1e-6, silu=True)
        sc = ops.linear(x, L.shortcut.w, L.shortcut.b) if "shortcut" in L else x
        return ops.conv2d_3x3(t, n_img, h, w_, L.conv2.w, L.conv2.b, residual=sc)

    def _attention(self, A, x, n_img, hw):
        if hw % 64:
            raise ValueError(f"VAE mid-block attention needs h*w % 64 == 0 (got {hw})")
        c = A.c
        gx = ops.groupnorm(x, n_img, A.norm.w, A.norm.b, 1e-6, silu=False)
        qkv = ops.linear(gx, A.qkv.w, A.qkv.b)
        scores = torch.empty((n_img * hw, hw), dtype=torch.float16, device=self.dev)
        for i in range(n_img):  # one image = one (hw x hw) problem; K / V differ per image
            r = slice(i * hw, (i + 1) * hw)
            ops.linear(qkv[r, 0:c], qkv[r, c:2 * c], out=scores[r])
        _lib.check(_lib.lib().lvd_softmax_rows(scores.data_ptr(), scores.stride(0), scores.shape[0], hw, 1.0,
                                               ops._stream()), "softmax_rows")
        o = torch.empty((n_img * hw, c), dtype=torch.float16, device=self.dev)
        for i in range(n_img):
            r = slice(i * hw, (i + 1) * hw)
            ops.linear(scores[r], qkv[r, 2 * c:3 * c], b_mn_major=True, n=c, out=o[r])
        return ops.linear(o, A.out.w, A.out.b, residual=x)

    def decode(self, z):
        """z (N, 4, h, w) fp16 (already divided by scaling_factor) -> (N, 3, 8h, 8w) fp16."""
        w = self.w
        z = z.to(device=self.dev, dtype=torch.float16).contiguous()
        n, lc, h, w_ = z.shape
        z1 = torch.empty_like(z)
        _lib.check(_lib.lib().lvd_pointwise_small(z.data_ptr(), w.post_quant.w.data_ptr(), w.post_quant.b.data_ptr(),
                                                  z1.data_ptr(), n, lc, lc, h * w_, ops._stream()), "pointwise_small")
        x = ops.conv_in_tc(z1.view(n, lc, 1, h, w_), w.conv_in.w_pad, w.conv_in.b)
        x = self._resnet(w.mid0, x, n, h, w_)
        x = self._attention(w.attn, x, n, h * w_)
        x = self._resnet(w.mid1, x, n, h, w_)
        for blk in w.ups:
            for L in blk.resnets:
                x = self._resnet(L, x, n, h, w_)
            if "up" in blk:
                x = ops.upsample2x(x, n, h, w_)
                h, w_ = 2 * h, 2 * w_
                x = ops.conv2d_3x3(x, n, h, w_, blk.up.w, blk.up.b)
        x = ops.groupnorm(x, n, w.norm_out.w, w.norm_out.b, 1e-6, silu=True)
        co = w.conv_out
        return ops.conv_out_tc(x, co.w_pad, co.b_pad, co.cout, n, 1, h, w_).view(n, co.cout, h, w_)


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class _Config(dict):
    __getattr__ = dict.__getitem__


class AutoencoderKL:
    """The slice of diffusers' AutoencoderKL that TextToVideoSDPipeline touches (decode side)."""

    def __init__(self, latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32, scaling_factor=0.18215):
        if norm_num_groups != 32:
            raise NotImplementedError("lvd_b200 AutoencoderKL: 32 GroupNorm groups only")
        self.config = _Config(latent_channels=latent_channels, out_channels=out_channels,
                              block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                              norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        self._engine = None
        self._state = None
        self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        self.slice_images = 0  # 0 = decode all images at once; enable_slicing() bounds the activation memory

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **kw):
        """Local diffusers-format directory (`<path>/<subfolder>/config.json` + weights); no hub access."""
        from .checkpoint import init_kwargs, load_component
        cfg, sd = load_component(pretrained_model_name_or_path, subfolder)
        model = cls(**init_kwargs(cls, cfg))
        model.load_state_dict(sd)
        return model

    @property
    def dtype(self):
        return torch.float16

    def to(self, *args, **kw):
        for a in list(args) + list(kw.values()):
            if isinstance(a, (str, torch.device)):
                d = torch.device(a)
                if d.type != "cuda":
                    raise RuntimeError("lvd_b200 has no CPU path")
                self._device = d if d.index is not None else torch.device("cuda", torch.cuda.current_device())
        if self._state is not None and self._engine is None and self._device is not None:
            self._build()
        return self

    def enable_slicing(self, images=4):
        """diffusers decodes one image per slice; here a slice is `images` frames (activations of a 576x320 frame at
        128 channels are 47 MB, so a handful per slice keeps the working set to a few GB)."""
        self.slice_images = int(images)

    def disable_slicing(self):
        self.slice_images = 0

    # diffusers < 0.17 `AttentionBlock` parameter names, still present in the published SD-1.x / Zeroscope / ModelScope
    # VAE checkpoints; diffusers 0.27.2 converts them on load (`_convert_deprecated_attention_blocks`)
    _LEGACY_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}

    @classmethod
    def _convert_legacy(cls, state_dict):
        out = {}
        for k, v in state_dict.items():
            parts = k.split(".")
            if len(parts) >= 3 and "attentions" in parts and parts[-2] in cls._LEGACY_ATTN:
                k = ".".join(parts[:-2] + [cls._LEGACY_ATTN[parts[-2]], parts[-1]])
            if ".attentions." in k and k.endswith(".weight") and v.dim() == 4 and v.shape[2:] == (1, 1) and \
                    ".to_" in k:
                v = v.reshape(v.shape[0], v.shape[1])  # the oldest checkpoints store the projections as 1x1 convs
            out[k] = v
        return out

    def load_state_dict(self, state_dict, strict=True):
        spec = synthetic.vae_param_spec(self.config)
        state_dict = self._convert_legacy(state_dict)
        missing = [k for k in spec if k not in state_dict]
        unexpected = [k for k in state_dict if k not in spec and not k.startswith(("encoder.", "quant_conv."))]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        for k, shp in spec.items():
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {tuple(shp)}")
        self._state = {k: state_dict[k] for k in spec if k in state_dict}
        self._engine = None
        if self._device is not None:
            self._build()
        return self

    def _build(self):
        self._engine = VAEDecoderEngine(self.config, self._state, self._device)
        self._state = None

    def decode_video_np(self, latents):
        """(B, 4, F, h, w) scheduler latents -> np.float32 [B, F, 8h, 8w, 3] in [0, 1]: `decode_latents` followed by
        `tensor2vid(..., "np")` (controllable_pipeline_text_to_video_synth.py:374-400, 66-88)."""
        b, c, f, h, w = latents.shape
        lat = (latents / self.config.scaling_factor).permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        img = self.decode(lat).sample
        n, co, hh, ww = img.shape
        out = torch.empty((b, f, hh, ww, co), dtype=torch.float32, device=img.device)
        _lib.check(_lib.lib().lvd_frames_postprocess(img.data_ptr(), out.data_ptr(), n, co, hh * ww, ops._stream()),
                   "frames_postprocess")
        return out.cpu().numpy()

    def decode(self, z, return_dict=True):
        if self._engine is None:
            raise RuntimeError("AutoencoderKL has no weights: call load_state_dict() first")
        n = z.shape[0]
        step = self.slice_images if self.slice_images > 0 else n
        outs = [self._engine.decode(z[i:i + step]) for i in range(0, n, step)]
        out = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return DecoderOutput(sample=out) if return_dict else (out,)
