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
from ..utils.guidance import energy_options

def _graph_for(unet):
    """The guidance pass of an engine lives ON the engine object (captured graphs and their activations are released
    with it)."""
    eng = getattr(unet, "_engine", unet)
    if eng is None:
        raise RuntimeError("UNet3DConditionModel has no weights loaded")
    g = getattr(eng, "_guidance_graph", None)
    if g is None:
        g = GuidanceGraph(eng)
        eng._guidance_graph = g
    return g


def latent_backward_guidance(scheduler, unet, cond_embeddings, index, bboxes, object_positions, t, latents, loss,
                             loss_scale=30, loss_threshold=0.2, max_iter=5, max_index_step=10,
                             cross_attention_kwargs=None, guidance_attn_keys=None, verbose=False,
                             return_saved_attn=False, clear_cache=False, **kwargs):
    """models/pipelines.py:21-150.  `return_saved_attn` ("first" / "last", :86-98) makes the pass materialise the guided
    maps and returns them on the CPU like the reference; `cross_attention_kwargs` may carry the map-saving keys the
    runners pass along (they are overridden by the guidance pass's own, :66-72) -- anything that would change the forward
    (gligen, attn_process_fn) is not part of this path in the reference runners and raises."""
    for k in (cross_attention_kwargs or {}):
        if k not in ("save_attn_to_dict", "save_keys", "offload_cross_attn_to_cpu", "enable_flash_attn"):
            raise NotImplementedError(f"cross_attention_kwargs[{k!r}] on the guidance pass is not supported")
    saved_to_return = None
    energy_options(kwargs)  # validates the loss keywords (same errors as the reference's own assertions)
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
            if return_saved_attn and return_saved_attn not in ("first", "last"):
                raise ValueError(return_saved_attn)  # :97-98
            want_maps = (return_saved_attn == "first" and iteration == 0) or \
                        (return_saved_attn == "last" and iteration == max_iter - 1)
            loss, grad = graph.run(x, t, cond_embeddings, bboxes, object_positions, guidance_attn_keys,
                                   loss_scale=loss_scale, return_saved_attn=want_maps, **kwargs)
            if want_maps:
                saved_to_return = {k: v.detach().cpu() for k, v in graph.saved_attn.items()}
            if torch.isnan(loss):
                print("**Loss is NaN**")
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
    if return_saved_attn:
        if saved_to_return is None:  # the reference's local is unbound when the selected iteration never ran (:148-149)
            raise UnboundLocalError("cannot access local variable 'saved_attn_to_return' where it is not associated "
                                    "with a value")
        return latents, loss, saved_to_return
    return latents, loss


def latent_backward_guidance_batch(scheduler, unet, cond_embeddings, index, bboxes, object_positions, t, latents, loss,
                                   loss_scale=30, loss_threshold=0.2, max_iter=5, max_index_step=10,
                                   guidance_attn_keys=None, verbose=False, **kwargs):
    """B prompts guided in lock-step (an extension of this repo: the reference guides one prompt per call).  `latents`
    (B, 4, F, h, w), `cond_embeddings` (B, 77, D), `bboxes` / `object_positions` = one layout per prompt, `loss` = B
    carried losses (CPU tensor / list).  Per prompt this evaluates exactly the while-loop of latent_backward_guidance
    (models/pipelines.py:59-137) -- same threshold test, same iteration cap, same update -- but every iteration is ONE
    batched guidance pass over all B prompts; a prompt whose loop has ended keeps its latents and its loss while the
    others continue.  Host syncs: one B-float device-to-host copy per iteration."""
    energy_options(kwargs)
    b = latents.shape[0]
    loss = torch.as_tensor(loss, dtype=torch.float32).reshape(-1).clone().cpu()
    if loss.numel() == 1 and b > 1:
        loss = loss.expand(b).clone()
    assert loss.numel() == b and len(bboxes) == b and len(object_positions) == b
    if index >= max_index_step:
        return latents, loss
    if isinstance(max_iter, list):
        max_iter = max_iter[index]
    graph = _graph_for(unet)
    iteration = [0] * b
    t_int = int(t)
    if hasattr(scheduler, "alphas_cumprod"):
        scale = (1 - float(scheduler.alphas_cumprod[t_int])) ** 0.5
    else:
        scale = 1.0

    def active(i):
        return loss[i].item() / loss_scale > loss_threshold and iteration[i] < max_iter

    while any(active(i) for i in range(b)):
        act = [active(i) for i in range(b)]
        x = scheduler.scale_model_input(latents, t)
        new_loss, grad = graph.run(x, t, cond_embeddings, bboxes, object_positions, guidance_attn_keys,
                                   loss_scale=loss_scale, **kwargs)
        cur = latents.to(device=grad.device, dtype=torch.float16).contiguous()
        new_lat, _ = ops.latent_update(cur, grad, scale, 1.0 / GRAD_SCALE)
        keep = torch.tensor(act, device=grad.device).reshape(b, 1, 1, 1, 1)
        latents = torch.where(keep, new_lat, cur)  # selection only: prompts whose loop has ended keep their latents
        host = new_loss.reshape(-1).cpu()
        for i in range(b):
            if act[i]:
                loss[i] = host[i]
                iteration[i] += 1
        if verbose:
            print(f"time index {index}, losses: {[round(v / loss_scale, 3) for v in loss.tolist()]}, iterations: "
                  f"{iteration}")
    return latents, loss
