"""Mirror of the reference's `models/pipelines.py::latent_backward_guidance` (lines 21-150): same arguments, same
loop, same return values -- but each iteration is one call into the native guidance pass
(`lvd_b200.guidance_engine.GuidanceGraph.run`: forward prefix + fused box energy + hand-written backward) followed by
the fused latent update kernel; PyTorch autograd is never involved.

Host syncs: the reference calls `loss.item()` in the while-condition (and on every print); here that is the only
device-to-host copy (4 bytes per iteration) because the threshold test genuinely needs the scalar on the host.
"""
import warnings

import torch

from .. import ops
from ..guidance_engine import GRAD_SCALE, GuidanceGraph
from ..utils.guidance import _check_kwargs

_graphs = {}


def _graph_for(unet):
    eng = getattr(unet, "_engine", unet)
    if eng is None:
        raise RuntimeError("UNet3DConditionModel has no weights loaded")
    g = _graphs.get(id(eng))
    if g is None:
        g = GuidanceGraph(eng)
        _graphs[id(eng)] = g
    return g


def latent_backward_guidance(scheduler, unet, cond_embeddings, index, bboxes, object_positions, t, latents, loss,
                             loss_scale=30, loss_threshold=0.2, max_iter=5, max_index_step=10,
                             cross_attention_kwargs=None, guidance_attn_keys=None, verbose=False,
                             return_saved_attn=False, clear_cache=False, **kwargs):
    """return_saved_attn is not supported by the native path (the maps are never materialised); everything else
    follows models/pipelines.py:21-150."""
    if return_saved_attn:
        raise NotImplementedError("lvd_b200 latent_backward_guidance does not materialise attention maps; call "
                                  "UNet3DConditionModel.forward with save_attn_to_dict for visualisation")
    if cross_attention_kwargs:
        raise NotImplementedError("extra cross_attention_kwargs on the guidance pass are not supported")
    _check_kwargs(kwargs)
    iteration = 0
    if not torch.is_tensor(loss):
        loss = torch.tensor(float(loss))
    if index < max_index_step:
        if isinstance(max_iter, list):
            max_iter = max_iter[index]
        graph = _graph_for(unet)
        if verbose:
            print(f"time index {index}, loss: {loss.item()/loss_scale:.3f} (de-scaled with scale {loss_scale:.1f}), "
                  f"loss threshold: {loss_threshold:.3f}")
        t_int = int(t)
        while loss.item() / loss_scale > loss_threshold and iteration < max_iter and index < max_index_step:
            x = scheduler.scale_model_input(latents, t)
            loss, grad = graph.run(x, t, cond_embeddings, bboxes, object_positions, guidance_attn_keys,
                                   loss_scale=loss_scale, base_attn_dim=kwargs.get("base_attn_dim", (40, 72)),
                                   fg_top_p=kwargs.get("fg_top_p", 0.2), bg_top_p=kwargs.get("bg_top_p", 0.2),
                                   fg_weight=kwargs.get("fg_weight", 1.0), bg_weight=kwargs.get("bg_weight", 1.0),
                                   com_loss_scale=kwargs.get("com_loss_scale", 0.0))
            if hasattr(scheduler, "alphas_cumprod"):
                warnings.warn("Using guidance scaled with alphas_cumprod")
                alpha_prod_t = float(scheduler.alphas_cumprod[t_int])
                scale = (1 - alpha_prod_t) ** 0.5
            else:
                warnings.warn("No scaling in guidance is performed")
                scale = 1.0
            latents, _ = ops.latent_update(latents.to(device=grad.device, dtype=torch.float16).contiguous(), grad,
                                           scale, 1.0 / GRAD_SCALE)
            iteration += 1
            if verbose:
                print(f"time index {index}, loss: {loss.item()/loss_scale:.3f}, loss threshold: {loss_threshold:.3f}, "
                      f"iteration: {iteration}")
    return latents, loss
