"""Mirror of the reference's models/unet_3d_condition.py public surface (UNet3DConditionModel, UNet3DConditionOutput)
on top of `lvd_b200.engine.UNetEngine`.

Kept from the reference (file:line are into /root/reference/models/unet_3d_condition.py):
  * constructor keywords and defaults of `UNet3DConditionModel.__init__` (:228-257), exposed as `.config.<name>`;
  * `forward(sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None, attention_mask=None,
    cross_attention_kwargs=None, down_block_additional_residuals=None, mid_block_additional_residual=None,
    return_dict=True)` (:642-654) returning `UNet3DConditionOutput(sample)` or `(sample,)` (:856-859);
  * `cross_attention_kwargs` protocol: the dict is required (the reference writes `["attn_key"]` into it, :731);
    `save_attn_to_dict` + `save_keys` store the cross-attention probabilities (B*F, heads, hw, 77) under
    `tuple(attn_key)` (models/attention_processor.py:553-586); `gligen` carries boxes / masks / positive_embeddings;
  * `.modules()` yields objects whose class name is `GatedSelfAttentionDense` with a writable `.enabled`
    (the pipeline toggles them, controllable_pipeline_text_to_video_synth.py:535-539);
  * `load_state_dict` takes the reference's parameter names unchanged.
`return_token_ca_only`, `return_cond_ca_only`, `offload_cross_attn_to_cpu` slice / move the stored maps as the reference
processor does (models/attention_processor.py:556-577); `cross_attn_save_hidden_states` is accepted and ignored (it only
parks an activation on the processor object for debugging, :457-458).
Not implemented (raise, never fall back): attention_mask, class_labels, timestep_cond, additional residuals,
attn_process_fn.
"""
from dataclasses import dataclass

import torch

from .. import synthetic
from ..engine import UNetEngine


@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


class _Config(dict):
    __getattr__ = dict.__getitem__


class GatedSelfAttentionDense:
    """Handle of one GLIGEN fuser (models/attention.py:26-60); only the `enabled` switch lives on the host."""

    def __init__(self, owner):
        self._owner = owner
        self._enabled = True

    @property
    def enabled(self):
        return self._enabled

    @enabled.setter
    def enabled(self, v):
        self._enabled = bool(v)
        if self._owner._engine is not None:
            self._owner._engine.fuser_enabled = all(f._enabled for f in self._owner._fusers)


class UNet3DConditionModel:
    _supports_gradient_checkpointing = False

    def __init__(self, sample_size=None, in_channels=4, out_channels=4,
                 down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D",
                                   "DownBlock3D"),
                 up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
                 block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1,
                 mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1024,
                 attention_head_dim=64, num_attention_heads=None, attention_type="default"):
        if num_attention_heads is not None:
            raise NotImplementedError("At the moment it is not possible to define the number of attention heads via "
                                      "`num_attention_heads`")  # unet_3d_condition.py:262-265
        if len(down_block_types) != len(up_block_types) or len(block_out_channels) != len(down_block_types):
            raise ValueError("Must provide the same number of `down_block_types`, `up_block_types` and "
                             "`block_out_channels`")
        if (tuple(down_block_types) != ("CrossAttnDownBlock3D",) * 3 + ("DownBlock3D",) or
                tuple(up_block_types) != ("UpBlock3D",) + ("CrossAttnUpBlock3D",) * 3 or act_fn != "silu" or
                norm_num_groups != 32 or attention_head_dim != 64 or mid_block_scale_factor != 1 or
                downsample_padding != 1 or in_channels > 8 or out_channels > 8):
            raise NotImplementedError("lvd_b200 implements the Zeroscope / ModelScope UNet3D topology only")
        if attention_type not in ("default", "gated"):
            raise NotImplementedError(f"attention_type={attention_type!r}")
        self.config = _Config(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                              down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
                              block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                              downsample_padding=downsample_padding, mid_block_scale_factor=mid_block_scale_factor,
                              act_fn=act_fn, norm_num_groups=norm_num_groups, norm_eps=norm_eps,
                              cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
                              num_attention_heads=num_attention_heads, attention_type=attention_type)
        self.sample_size = sample_size
        self._cfg = dict(in_channels=in_channels, out_channels=out_channels,
                         block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                         cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
                         norm_num_groups=norm_num_groups, norm_eps=norm_eps, gated=attention_type == "gated")
        self._engine = None
        self._state = None
        self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        n_fusers = 16 if attention_type == "gated" else 0
        self._fusers = [GatedSelfAttentionDense(self) for _ in range(n_fusers)]
        # Launch sequencing only: a forward whose shapes repeat is replayed from a CUDA graph (the ~1100 kernel launches
        # of one forward cost more host time than the GPU needs to run them).  Same kernels, same results.
        self.use_cuda_graph = True
        self._graphs = {}

    # ---- torch.nn.Module-like surface the runners / pipeline touch
    @property
    def dtype(self):
        return torch.float16

    @property
    def device(self):
        return self._device

    def modules(self):
        yield self
        yield from self._fusers

    def eval(self):
        return self

    def to(self, *args, **kw):
        for a in list(args) + list(kw.values()):
            if isinstance(a, torch.dtype) and a != torch.float16:
                raise NotImplementedError("lvd_b200 computes in fp16 (the dtype the reference runners select)")
            if isinstance(a, (str, torch.device)):
                d = torch.device(a)
                if d.type != "cuda":
                    raise RuntimeError("lvd_b200 has no CPU path")
                self._device = d if d.index is not None else torch.device("cuda", torch.cuda.current_device())
        if self._state is not None and self._engine is None:
            self._build()
        return self

    def state_dict_spec(self):
        return synthetic.param_spec(self._cfg)

    def load_state_dict(self, state_dict, strict=True):
        spec = self.state_dict_spec()
        missing = [k for k in spec if k not in state_dict]
        unexpected = [k for k in state_dict if k not in spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        for k, shp in spec.items():
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {tuple(shp)}")
        self._state = state_dict
        self._engine = None
        if self._device is not None:
            self._build()
        return self

    def _build(self):
        if self._device is None:
            raise RuntimeError("lvd_b200.UNet3DConditionModel needs a CUDA device (sm_100a); there is no CPU path")
        self._engine = UNetEngine(self._cfg, self._state, device=self._device)
        self._state = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, revision=None, torch_dtype=None, **kw):
        """Loads a LOCAL diffusers-format directory (config.json + diffusion_pytorch_model.{safetensors,bin}); there is
        no network access, so hub names that are not local paths raise."""
        from ..checkpoint import init_kwargs, load_component
        cfg, sd = load_component(pretrained_model_name_or_path, subfolder)
        model = cls(**init_kwargs(cls, cfg))
        model.load_state_dict(sd)
        return model

    # ---- forward
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict=True):
        if self._engine is None:
            raise RuntimeError("UNet3DConditionModel has no weights: call load_state_dict()/from_pretrained() first")
        for name, v in (("class_labels", class_labels), ("timestep_cond", timestep_cond),
                        ("attention_mask", attention_mask),
                        ("down_block_additional_residuals", down_block_additional_residuals),
                        ("mid_block_additional_residual", mid_block_additional_residual)):
            if v is not None:
                raise NotImplementedError(f"lvd_b200 UNet3DConditionModel.forward: `{name}` is not supported")
        if cross_attention_kwargs is None:
            # the reference crashes here too (it writes cross_attention_kwargs["attn_key"], :731)
            raise TypeError("cross_attention_kwargs must be a dict")
        kw = cross_attention_kwargs
        if kw.get("attn_process_fn") is not None:
            raise NotImplementedError("cross_attention_kwargs['attn_process_fn'] is not supported by lvd_b200")
        kw["attn_key"] = ["up", len(self._cfg["block_out_channels"]) - 1]  # what the reference leaves behind
        save_dict = kw.get("save_attn_to_dict")
        save_keys = kw.get("save_keys")
        tok_only = kw.get("return_token_ca_only")
        cond_only = bool(kw.get("return_cond_ca_only", False))
        offload = bool(kw.get("offload_cross_attn_to_cpu", False))
        capture = None
        pending = []
        # models/attention_processor.py:578-586: a map is stored only when BOTH save_attn_to_dict and save_keys are
        # given and the key is listed (save_keys=None takes the materialising path there but stores nothing)
        if save_dict is not None and save_keys is not None and len(save_keys) > 0:
            wanted = {tuple(k) for k in save_keys}
            dev = self._engine.dev
            if cond_only:
                assert sample.shape[0] % 2 == 0, f"Samples are not in pairs: {sample.shape[0]} samples"

            def capture(key, n_img, heads, hw):
                if key in wanted:
                    t = torch.empty((n_img, heads, hw, encoder_hidden_states.shape[1]), dtype=torch.float16,
                                    device=dev)
                    pending.append((key, t))
                    return dict(p_full=t)
                return {}

        gligen = kw.get("gligen")
        if capture is None and self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
            out = self._forward_graphed(sample, timestep, encoder_hidden_states, gligen)
        else:
            out = self._engine.forward(sample, timestep, encoder_hidden_states, capture=capture, gligen=gligen)
        for key, p in pending:  # the slicing the reference processor applies before storing (:556-577)
            if tok_only is not None:
                p = p[:, :, :, tok_only:tok_only + 1] if isinstance(tok_only, int) else p[:, :, :, tok_only]
            if cond_only:
                p = p[p.shape[0] // 2:]
            if offload:
                p = p.cpu()
            save_dict[key] = p
        if not return_dict:
            return (out,)
        return UNet3DConditionOutput(sample=out)

    def _forward_graphed(self, sample, timestep, ehs, gligen):
        """Eager on the first call of a shape, then captured once and replayed: inputs are copied into the graph's
        static buffers, the output is cloned out of its static buffer."""
        eng = self._engine
        if gligen is not None and not self._cfg["gated"]:
            raise ValueError("`gligen` inputs given but the model has no GLIGEN weights (attention_type='default')")
        use_gligen = gligen is not None and eng.fuser_enabled
        key = (tuple(sample.shape), tuple(ehs.shape), use_gligen,
               tuple(tuple(v.shape) for v in (gligen["boxes"], gligen["positive_embeddings"], gligen["masks"]))
               if use_gligen else None)
        ent = self._graphs.get(key)
        if ent is None:  # first sighting: run eagerly (also warms allocator and kernel attributes)
            self._graphs[key] = "seen"
            return eng.forward(sample, timestep, ehs, gligen=gligen if use_gligen else None)
        dev = eng.dev
        b, _, frames = sample.shape[:3]
        if ent == "seen":
            st = dict(x=torch.empty(tuple(sample.shape), dtype=torch.float16, device=dev),
                      t=torch.empty((b * frames,), dtype=torch.float32, device=dev),
                      ehs=torch.empty(tuple(ehs.shape), dtype=torch.float16, device=dev), gl=None)
            if use_gligen:
                st["gl"] = {k: torch.empty(tuple(gligen[k].shape), dtype=torch.float16, device=dev)
                            for k in ("boxes", "positive_embeddings", "masks")}
            self._fill_static(st, sample, timestep, ehs, gligen if use_gligen else None)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            from .. import ops
            scope = ops.PROF_SCOPE
            if scope is not None:  # bench.py's instrumented replay: timing-event nodes are captured with the kernels
                ops.PROF_SCOPE = "unet:fusers" if use_gligen else "unet"
            try:
                with torch.cuda.graph(g):
                    # disabled fusers never read the grounding tokens (models/attention.py:44-46): PositionNet is skipped
                    st["out"] = eng.forward(st["x"], st["t"], st["ehs"], gligen=st["gl"])
            finally:
                ops.PROF_SCOPE = scope
            st["graph"] = g
            self._graphs[key] = ent = st
        self._fill_static(ent, sample, timestep, ehs, gligen if use_gligen else None)
        ent["graph"].replay()
        return ent["out"].clone()

    @staticmethod
    def _fill_static(st, sample, timestep, ehs, gligen):
        st["x"].copy_(sample, non_blocking=True)
        st["ehs"].copy_(ehs, non_blocking=True)
        b = sample.shape[0]
        frames = sample.shape[2]
        if torch.is_tensor(timestep) and timestep.numel() > 1:
            tt = timestep.to(device=st["t"].device, dtype=torch.float32).reshape(-1)
            st["t"].copy_(tt.expand(b).repeat_interleave(frames) if tt.numel() == b else tt)
        else:
            st["t"].fill_(float(timestep))
        if gligen is not None:
            for k, v in st["gl"].items():
                v.copy_(gligen[k], non_blocking=True)

    __call__ = forward
