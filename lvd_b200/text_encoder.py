"""CLIP text encoder on the native kernels -- SURVEY.md section 8(f) rank 2.

`TextToVideoSDPipeline._encode_prompt` (controllable_pipeline_text_to_video_synth.py:197-372) runs
`self.text_encoder(input_ids)[0]` on the prompt / negative prompt, and the GLIGEN preparation (:740-763) takes
`self.text_encoder(**tokens).pooler_output` of the object phrases.  The text encoder is transformers' `CLIPTextModel`
(OpenCLIP ViT-H text tower in the Zeroscope / ModelScope checkpoints: 1024 wide, 16 heads of 64, 23 layers, GELU MLP of
4096, causal attention over 77 tokens).  `CLIPTextModel` below mirrors the slice of that class the pipeline touches
(`__call__(input_ids, attention_mask=None)` -> output with `[0]` / `.last_hidden_state` and `.pooler_output`,
`.config`, `.dtype`, `.to()`, `load_state_dict` with transformers' parameter names) over the C-ABI kernels:
`lvd_embed_tokens`, `lvd_layernorm`, `lvd_gemm` (fused q|k|v projection, out projection + residual, fc1 + GELU epilogue,
fc2 + residual) and `lvd_attention` with `causal=1` (warp-level mma.sync kernel, attention_temporal.cu).  The tokenizer
stays the caller's (pure host code).  No CPU path.
"""
from dataclasses import dataclass

import torch

from . import _lib, ops
from .engine import LayerWeights, _h


def clip_param_spec(cfg):
    d, ff = cfg["hidden_size"], cfg["intermediate_size"]
    spec = {"text_model.embeddings.token_embedding.weight": (cfg["vocab_size"], d),
            "text_model.embeddings.position_embedding.weight": (cfg["max_position_embeddings"], d)}
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            spec[p + f"self_attn.{n}.weight"] = (d, d)
            spec[p + f"self_attn.{n}.bias"] = (d,)
        for n in ("layer_norm1", "layer_norm2"):
            spec[p + n + ".weight"] = (d,)
            spec[p + n + ".bias"] = (d,)
        spec[p + "mlp.fc1.weight"] = (ff, d)
        spec[p + "mlp.fc1.bias"] = (ff,)
        spec[p + "mlp.fc2.weight"] = (d, ff)
        spec[p + "mlp.fc2.bias"] = (d,)
    spec["text_model.final_layer_norm.weight"] = (d,)
    spec["text_model.final_layer_norm.bias"] = (d,)
    return spec


class CLIPTextEngine:
    def __init__(self, cfg, state_dict, device):
        _lib.lib()
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("lvd_b200 text encoder needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = dict(cfg)
        self.dev = dev = torch.device(device)
        d, heads = cfg["hidden_size"], cfg["num_attention_heads"]
        if d != heads * 64 or d % 64 or cfg["intermediate_size"] % 64:
            raise NotImplementedError("lvd_b200 text encoder: 64-wide heads, widths multiples of 64")
        if cfg.get("hidden_act", "gelu") != "gelu":
            raise NotImplementedError(f"hidden_act={cfg.get('hidden_act')!r} (the OpenCLIP ViT-H tower uses gelu)")
        sd = state_dict

        def g(n):
            return _h(sd[n], dev)

        self.tok = g("text_model.embeddings.token_embedding.weight")
        self.pos = g("text_model.embeddings.position_embedding.weight")
        self.layers = []
        for i in range(cfg["num_hidden_layers"]):
            p = f"text_model.encoder.layers.{i}."
            a = p + "self_attn."
            self.layers.append(LayerWeights(
                ln1=LayerWeights(w=g(p + "layer_norm1.weight"), b=g(p + "layer_norm1.bias")),
                qkv=LayerWeights(w=torch.cat([g(a + "q_proj.weight"), g(a + "k_proj.weight"), g(a + "v_proj.weight")], 0),
                                 b=torch.cat([g(a + "q_proj.bias"), g(a + "k_proj.bias"), g(a + "v_proj.bias")], 0)),
                out=LayerWeights(w=g(a + "out_proj.weight"), b=g(a + "out_proj.bias")),
                ln2=LayerWeights(w=g(p + "layer_norm2.weight"), b=g(p + "layer_norm2.bias")),
                fc1=LayerWeights(w=g(p + "mlp.fc1.weight"), b=g(p + "mlp.fc1.bias")),
                fc2=LayerWeights(w=g(p + "mlp.fc2.weight"), b=g(p + "mlp.fc2.bias"))))
        self.ln_f = LayerWeights(w=g("text_model.final_layer_norm.weight"), b=g("text_model.final_layer_norm.bias"))

    def forward(self, input_ids):
        """input_ids (B, S) integer -> last_hidden_state (B, S, D) fp16 (after final_layer_norm)."""
        cfg = self.cfg
        ids = input_ids.to(device=self.dev, dtype=torch.int64).contiguous()
        b, s = ids.shape
        d, heads = cfg["hidden_size"], cfg["num_attention_heads"]
        if s > cfg["max_position_embeddings"] or s > 128:
            raise ValueError(f"sequence length {s} exceeds the position table / the 128-key attention kernel")
        eps = cfg.get("layer_norm_eps", 1e-5)
        x = torch.empty((b * s, d), dtype=torch.float16, device=self.dev)
        _lib.check(_lib.lib().lvd_embed_tokens(ids.data_ptr(), self.tok.data_ptr(), self.pos.data_ptr(), x.data_ptr(),
                                               b * s, s, d, self.tok.shape[0], ops._stream()), "embed_tokens")
        for L in self.layers:  # CLIPEncoderLayer: pre-LN, causal self-attention, GELU MLP
            t = ops.layernorm(x, L.ln1.w, L.ln1.b, eps)
            qkv = ops.linear(t, L.qkv.w, L.qkv.b)
            a = ops.attention(qkv, qkv, heads=heads, nz=b, s_q=s, s_k=s, q_col0=0, k_col0=d, v_col0=2 * d,
                              scale=0.125, causal=True)
            x = ops.linear(a, L.out.w, L.out.b, residual=x)
            t = ops.layernorm(x, L.ln2.w, L.ln2.b, eps)
            t = ops.linear(t, L.fc1.w, L.fc1.b, act=ops.ACT_GELU)
            x = ops.linear(t, L.fc2.w, L.fc2.b, residual=x)
        return ops.layernorm(x, self.ln_f.w, self.ln_f.b, eps).view(b, s, d)


@dataclass
class TextEncoderOutput:
    last_hidden_state: torch.Tensor
    pooler_output: torch.Tensor

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


class _Config(dict):
    __getattr__ = dict.__getitem__


class CLIPTextModel:
    """The slice of transformers.CLIPTextModel that TextToVideoSDPipeline touches."""

    def __init__(self, vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                 num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5,
                 eos_token_id=49407):
        self.config = _Config(vocab_size=vocab_size, hidden_size=hidden_size, intermediate_size=intermediate_size,
                              num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                              max_position_embeddings=max_position_embeddings, hidden_act=hidden_act,
                              layer_norm_eps=layer_norm_eps, eos_token_id=eos_token_id)
        self._engine = None
        self._state = None
        self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **kw):
        """Local transformers-format directory (`<path>/<subfolder>/config.json` + weights); no hub access."""
        from .checkpoint import init_kwargs, load_component
        cfg, sd = load_component(pretrained_model_name_or_path, subfolder)
        model = cls(**init_kwargs(cls, cfg))
        model.load_state_dict({k: v for k, v in sd.items() if k.startswith("text_model.")})
        return model

    @property
    def dtype(self):
        return torch.float16

    @property
    def device(self):
        return self._device

    def to(self, *args, **kw):
        for a in list(args) + list(kw.values()):
            if isinstance(a, (str, torch.device)):
                dv = torch.device(a)
                if dv.type != "cuda":
                    raise RuntimeError("lvd_b200 has no CPU path")
                self._device = dv if dv.index is not None else torch.device("cuda", torch.cuda.current_device())
        if self._state is not None and self._engine is None and self._device is not None:
            self._build()
        return self

    def load_state_dict(self, state_dict, strict=True):
        spec = clip_param_spec(self.config)
        missing = [k for k in spec if k not in state_dict]
        unexpected = [k for k in state_dict if k not in spec and not k.endswith("position_ids")]
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
        self._engine = CLIPTextEngine(self.config, self._state, self._device)
        self._state = None

    def __call__(self, input_ids=None, attention_mask=None, **kw):
        """attention_mask is accepted and ignored exactly where the reference's calls make it a no-op: the pipeline
        passes none for prompts (:236-262 -- the OpenCLIP config has no `use_attention_mask`), and padded GLIGEN phrases
        only read positions up to their EOS token, which causal attention shields from the padding behind it."""
        if self._engine is None:
            raise RuntimeError("CLIPTextModel has no weights: call load_state_dict() first")
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        ids = input_ids.to(self._device)
        h = self._engine.forward(ids)
        eos = self.config.eos_token_id
        if eos == 2:  # legacy configs (transformers modeling_clip.py): EOS is the highest id
            idx = ids.to(torch.int64).argmax(dim=-1)
        else:
            idx = (ids == eos).int().argmax(dim=-1)
        pooled = h[torch.arange(h.shape[0], device=h.device), idx]
        return TextEncoderOutput(last_hidden_state=h, pooler_output=pooled)
