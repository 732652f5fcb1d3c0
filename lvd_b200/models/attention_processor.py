"""Mirror of the reference's models/attention_processor.py plugin boundary: `Attention` (weight holder with the
reference's parameter names) and `AttnProcessor.__call__` with the reference's full keyword list (:432-449).

Dispatch follows :454-474: the materialising path (attention map returned / stored under `tuple(attn_key)`) is taken
iff the call is cross-attention and a map is requested; every other call is the fused path.  Both run the same tcgen05
kernel (`lvd_attention`); the map, when requested, is emitted by the kernel itself -- exactly the fp16
probabilities that multiply V, like `get_attention_scores` + `bmm` (:222-258,:536).
"""
import torch

from .. import ops


class Attention:
    """models/attention_processor.py:27-165 (q/k/v without bias, out projection with bias, scale = dim_head**-0.5)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 out_bias=True, processor=None, **unsupported):
        for k, v in unsupported.items():
            if v not in (None, False, 0.0, 1.0, 32, 1e-5, True) or k in ("added_kv_proj_dim", "norm_num_groups"):
                if v:
                    raise NotImplementedError(f"Attention option {k}={v!r} is not used by the 3-D UNet")
        if bias or dim_head != 64:
            raise NotImplementedError("lvd_b200 Attention: q/k/v bias and head sizes other than 64 are not supported")
        self.query_dim = query_dim
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.inner_dim = heads * dim_head
        self.out_bias = out_bias
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.w = None
        self.processor = processor if processor is not None else AttnProcessor()

    def set_processor(self, processor):
        self.processor = processor

    def load_state_dict(self, sd, device="cuda"):
        g = lambda k: sd[k].detach().to(device=device, dtype=torch.float16).contiguous()  # noqa: E731
        self.w = dict(q=g("to_q.weight"), k=g("to_k.weight"), v=g("to_v.weight"), o=g("to_out.0.weight"),
                      ob=g("to_out.0.bias") if self.out_bias else None)
        self.w["kv"] = torch.cat([self.w["k"], self.w["v"]], 0).contiguous()
        self.w["qkv"] = torch.cat([self.w["q"], self.w["kv"]], 0).contiguous() \
            if self.cross_attention_dim == self.query_dim else None
        return self

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, return_attntion_probs=False,
                **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, return_attntion_probs=return_attntion_probs,
                              **cross_attention_kwargs)

    __call__ = forward


class AttnProcessor:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 return_attntion_probs=False, attn_key=None, attn_process_fn=None, return_cond_ca_only=False,
                 return_token_ca_only=None, offload_cross_attn_to_cpu=False, save_attn_to_dict=None, save_keys=None,
                 enable_flash_attn=True, cross_attn_save_hidden_states=False):
        if attn.w is None:
            raise RuntimeError("Attention has no weights: call load_state_dict() first")
        if attention_mask is not None or attn_process_fn is not None:
            raise NotImplementedError("attention_mask / attn_process_fn are not supported by lvd_b200")
        if hidden_states.dim() != 3:
            raise NotImplementedError("lvd_b200 AttnProcessor takes (batch, sequence, channels) inputs")
        cross_attn = encoder_hidden_states is not None
        if cross_attn and cross_attn_save_hidden_states:
            self.hidden_states = hidden_states
        want_map = cross_attn and ((save_attn_to_dict is not None and (save_keys is None or
                                                                      tuple(attn_key) in save_keys))
                                   or return_attntion_probs)
        b, s, c = hidden_states.shape
        dev = hidden_states.device
        x = hidden_states.to(torch.float16).reshape(b * s, c).contiguous()
        w = attn.w
        ci = attn.inner_dim
        probs = None
        if not cross_attn:
            qkv = ops.linear(x, w["qkv"])
            a = ops.attention(qkv, qkv, heads=attn.heads, nz=b, s_q=s, s_k=s, k_col0=ci, v_col0=2 * ci,
                              scale=attn.scale)
        else:
            e = encoder_hidden_states.to(device=dev, dtype=torch.float16)
            sk = e.shape[1]
            q = ops.linear(x, w["q"])
            kv = ops.linear(e.reshape(b * sk, e.shape[2]).contiguous(), w["kv"])
            if want_map:
                if sk > 128:
                    raise NotImplementedError("attention maps are emitted for up to 128 context tokens")
                probs = torch.empty((b, attn.heads, s, sk), dtype=torch.float16, device=dev)
            a = ops.attention(q, kv, heads=attn.heads, nz=b, s_q=s, s_k=sk, k_col0=0, v_col0=ci, scale=attn.scale,
                              p_full=probs)
        out = ops.linear(a, w["o"], w["ob"]).reshape(b, s, c)
        if return_attntion_probs or save_attn_to_dict is not None:
            p = probs
            if p is not None:
                if return_token_ca_only is not None:
                    if isinstance(return_token_ca_only, int):
                        p = p[:, :, :, return_token_ca_only:return_token_ca_only + 1]
                    else:
                        p = p[:, :, :, return_token_ca_only]
                if return_cond_ca_only:
                    assert b % 2 == 0, f"Samples are not in pairs: {b} samples"
                    p = p[b // 2:]
                if offload_cross_attn_to_cpu:
                    p = p.cpu()
                if save_attn_to_dict is not None and save_keys is not None and tuple(attn_key) in save_keys:
                    save_attn_to_dict[tuple(attn_key)] = p
            if return_attntion_probs:
                return out, p
        return out

    def free(self):
        if hasattr(self, "hidden_states"):
            del self.hidden_states


AttentionProcessor = AttnProcessor
