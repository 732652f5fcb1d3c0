"""TEST INFRASTRUCTURE: restatement of the reference's cross-attention box-energy guidance.

  compute_ca_loss       <- utils/guidance.py:529-574 (compute_ca_lossv3) + :160-526 (add_ca_loss_per_attn_map_to_loss),
                           every branch: max-based / ratio / CE losses, smoothing (utils/attn.py:91-157), renorm,
                           upsampling, attention sync, BoxDiff corner constraint, centre-of-mass terms
  scale_proportion      <- utils/utils.py:82-103     get_hw_from_attn_dim <- utils/utils.py:253-256
  guidance_step         <- models/pipelines.py:59-137 (one iteration of latent_backward_guidance: forward with
                           grad, energy, autograd.grad w.r.t. the latents, x -= sqrt(1-alpha_bar_t) * grad)

Written against plain torch so that autograd provides dE/dlatent for the checker; the CUDA engine computes the same
gradient with hand-written backward kernels and is compared with this.
"""
import math

import torch
import torch.nn.functional as F


def scale_proportion(obj_box, H, W):
    x_min, y_min = round(obj_box[0] * W), round(obj_box[1] * H)
    box_w, box_h = round((obj_box[2] - obj_box[0]) * W), round((obj_box[3] - obj_box[1]) * H)
    x_max, y_max = x_min + box_w, y_min + box_h
    x_min, y_min = max(x_min, 0), max(y_min, 0)
    x_max, y_max = min(x_max, W), min(y_max, H)
    return x_min, y_min, x_max, y_max


def get_hw_from_attn_dim(attn_dim, base_attn_dim):
    scale = int(math.sqrt((base_attn_dim[0] * base_attn_dim[1]) / attn_dim))
    return base_attn_dim[0] // scale, base_attn_dim[1] // scale


def _center_of_mass(x, h_range, w_range):
    com_h = (x.sum(dim=2) * h_range).sum(dim=-1) / x.sum(dim=(1, 2))
    com_w = (x.sum(dim=1) * w_range).sum(dim=-1) / x.sum(dim=(1, 2))
    return com_h, com_w


def gaussian_kernel2d(kernel_size=3, sigma=0.5):
    """utils/attn.py:112-134 (GaussianSmoothing.__init__): note the non-standard exponent -((x - mean) / (2 sigma))^2."""
    ax = torch.arange(kernel_size, dtype=torch.float32)
    mean = (kernel_size - 1) / 2
    g1 = 1 / (sigma * math.sqrt(2 * math.pi)) * torch.exp(-(((ax - mean) / (2 * sigma)) ** 2))
    k = g1[:, None] * g1[None, :]
    return k / k.sum()


def smooth_map(attn_map, kernel_size=3, sigma=0.5):
    """guidance.py:209-218: the (hw x n_text) plane of every (frame, head) is reflect-padded by 1 and blurred with a
    depthwise kernel_size x kernel_size kernel in the map's dtype -- the blur mixes neighbouring FLATTENED pixel indices
    and neighbouring text tokens (that is what the reference computes)."""
    n_f, heads, hw, n_text = attn_map.shape
    w = gaussian_kernel2d(kernel_size, sigma).to(attn_map.dtype).to(attn_map.device)
    x = F.pad(attn_map, (1, 1, 1, 1), mode="reflect")
    return F.conv2d(x, w.view(1, 1, kernel_size, kernel_size).repeat(heads, 1, 1, 1), groups=heads)


def ca_loss_per_map(attn_map, bboxes, object_positions, *, base_attn_dim, fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0,
                    bg_weight=1.0, com_loss_scale=0.0, use_ratio_based_loss=False, use_max_based_loss=True,
                    use_ce_based_loss=False, eps=1.0e-2, smooth_attn=False, kernel_size=3, sigma=0.5, upsample_scale=1,
                    upsample_mode="bilinear", attn_renorm=False, num_tokens=None, renorm_scale=2.0,
                    attn_sync_weight=0.0, boxdiff_loss_scale=0.0, boxdiff_normed=True, boxdiff_L=1):
    """attn_map (F, heads, hw, 77).  Returns sum over objects of obj_loss / n_tokens (guidance.py:160-526)."""
    n_f, heads, hw, _ = attn_map.shape
    if smooth_attn:
        attn_map = smooth_map(attn_map, kernel_size, sigma)
    if attn_renorm:  # :220-222
        attn_map = F.softmax(attn_map[..., 1:num_tokens - 1] * renorm_scale, dim=-1)
    H_attn, W_attn = get_hw_from_attn_dim(hw, base_attn_dim)
    H, W = int(H_attn * upsample_scale), int(W_attn * upsample_scale)
    dev = attn_map.device
    h_range = torch.arange(H, device=dev, dtype=attn_map.dtype)[None]
    w_range = torch.arange(W, device=dev, dtype=attn_map.dtype)[None]
    k_min = int(upsample_scale * upsample_scale)
    loss = torch.zeros((), device=dev, dtype=torch.float32)
    for obj_idx, obj_boxes in enumerate(bboxes):
        assert n_f == len(obj_boxes)
        obj_loss = 0
        for f, box in enumerate(obj_boxes):
            f1 = min(f + 1, len(obj_boxes) - 1)
            mask = torch.zeros((H, W), device=dev)
            x0, y0, x1, y1 = scale_proportion(box, H, W)
            mask[y0:y1, x0:x1] = 1
            if boxdiff_loss_scale > 0:  # :251-266
                cmx = torch.zeros((1, W), device=dev)
                cmy = torch.zeros((1, H), device=dev)
                for e in (x0, x1):
                    cmx[:, max(e - boxdiff_L, 0):min(e + boxdiff_L + 1, W)] = 1.0
                for e in (y0, y1):
                    cmy[:, max(e - boxdiff_L, 0):min(e + boxdiff_L + 1, H)] = 1.0
            mask_t1 = torch.zeros((H, W), device=dev)
            tx0, ty0, tx1, ty1 = scale_proportion(obj_boxes[f1], H, W)
            mask_t1[ty0:ty1, tx0:tx1] = 1
            for tok in object_positions[obj_idx]:
                col = tok - 1 if attn_renorm else tok  # SOT was dropped by the renormalisation (:281-285)
                ca = attn_map[f, :, :, col]
                ca_t1 = attn_map[f1, :, :, col]
                if upsample_scale != 1:  # :297-311
                    ca = F.interpolate(ca.view(heads, 1, H_attn, W_attn), size=(H, W), mode=upsample_mode)
                    ca_t1 = F.interpolate(ca_t1.view(heads, 1, H_attn, W_attn), size=(H, W), mode=upsample_mode)
                if use_ratio_based_loss:  # :312-323
                    ca = ca.reshape(heads, H, W)
                    act = (ca * mask).reshape(heads, -1).sum(dim=-1) / (ca.reshape(heads, -1).sum(dim=-1) + eps)
                    obj_loss = obj_loss + torch.mean((1 - act) ** 2)
                else:
                    k_fg = int((mask.sum() * fg_top_p).long().clamp_(min=k_min))
                    k_bg = int(((1 - mask).sum() * bg_top_p).long().clamp_(min=k_min))
                    m1 = mask.view(1, -1)
                    ca = ca.reshape(heads, -1)
                    if use_max_based_loss:  # :341-353
                        obj_loss = obj_loss + (1 - (ca * m1).topk(k=k_fg).values.mean(dim=1)).sum(dim=0) * fg_weight
                        obj_loss = obj_loss + ((ca * (1 - m1)).topk(k=k_bg).values.mean(dim=1)).sum(dim=0) * bg_weight
                    elif use_ce_based_loss:  # :363-396
                        ca = torch.clamp(ca, min=eps, max=1 - eps)
                        fg = (-torch.log(torch.clamp((m1 * ca).topk(k=k_fg).values, min=eps))).mean(dim=1).sum(dim=0)
                        bg = -torch.log(1 - ((1 - m1) * ca).topk(k=k_bg).values.mean(dim=1)).sum(dim=0)
                        obj_loss = obj_loss + fg * fg_weight + bg * bg_weight
                    else:
                        raise ValueError("Unknown loss: no loss selected")
                if attn_sync_weight != 0.0:  # :401-426
                    assert not attn_renorm, "attn_sync with attn_renorm not implemented together"
                    if f != len(obj_boxes) - 1:
                        # the crop is the NEXT frame's rectangle: the reference's x_min..y_max were last assigned in
                        # its loop over frame_boxes_t1 (:268-271) when the sync term reads them (:414-418)
                        a1 = ca.view(heads, H_attn, W_attn)[:, ty0:ty1, tx0:tx1]
                        a2 = attn_map[f + 1, :, :, tok].view(heads, H_attn, W_attn)[:, ty0:ty1, tx0:tx1]
                        obj_loss = obj_loss + ((a1 - a2) ** 2).mean(dim=(1, 2)).sum(dim=0) * attn_sync_weight
                if boxdiff_loss_scale > 0:  # :433-465
                    c3 = ca.view(heads, H, W)
                    mx_x = c3.max(dim=1).values.to(torch.float32)
                    mx_y = c3.max(dim=2).values.to(torch.float32)
                    mm_x = mask[None].max(dim=1).values
                    mm_y = mask[None].max(dim=2).values
                    red = (lambda t: t.mean()) if boxdiff_normed else (lambda t: t.sum())
                    cc = red((mx_x - mm_x).abs() * cmx) + red((mx_y - mm_y).abs() * cmy)
                    obj_loss = obj_loss + cc * boxdiff_loss_scale
                if com_loss_scale > 0 and mask.sum() > 0:  # :468-522
                    ca2 = ca.view(heads, H, W).to(torch.float32)
                    ch, cw = _center_of_mass(ca2, h_range, w_range)
                    mh, mw = _center_of_mass(mask[None], h_range, w_range)
                    obj_loss = obj_loss + (((ch[..., None, None] - mh) ** 2).mean() +
                                           ((cw[..., None, None] - mw) ** 2).mean()) * com_loss_scale
                    if mask_t1.sum() > 0:
                        c1 = ca_t1.view(heads, H, W).to(torch.float32)
                        ch1, cw1 = _center_of_mass(c1, h_range, w_range)
                        mh1, mw1 = _center_of_mass(mask_t1[None], h_range, w_range)
                        obj_loss = obj_loss + ((((ch1 - ch)[..., None, None] - (mh1 - mh)) ** 2).mean() +
                                               (((cw1 - cw)[..., None, None] - (mw1 - mw)) ** 2).mean()) * com_loss_scale
        loss = loss + obj_loss / len(object_positions[obj_idx])
    return loss


def compute_ca_loss(saved_attn, bboxes, object_positions, guidance_attn_keys, **kw):
    """compute_ca_lossv3 (guidance.py:529-574): mean over objects and keys."""
    first = next(iter(saved_attn.values()))
    loss = torch.zeros((), device=first.device, dtype=torch.float32)
    n_obj = len(bboxes)
    if n_obj == 0:
        return loss
    for key in guidance_attn_keys:
        loss = loss + ca_loss_per_map(saved_attn[key], bboxes, object_positions, **kw)
    if len(guidance_attn_keys) > 0:
        loss = loss / (n_obj * len(guidance_attn_keys))
    return loss


def guidance_step(sd, cfg, latents, t, cond_embeddings, bboxes, object_positions, alpha_prod_t, *, guidance_attn_keys,
                  loss_scale=30.0, **loss_kw):
    """One iteration of the while-loop in latent_backward_guidance (models/pipelines.py:65-137).
    Returns (new_latents, loss (scaled), grad, saved maps)."""
    from . import unet3d_ref as R

    lat = latents.detach().clone().requires_grad_(True)
    saved = {}
    with torch.enable_grad():
        R.unet_forward(sd, cfg, lat, t, cond_embeddings, save_keys=list(guidance_attn_keys), saved=saved,
                       stop_after_key=guidance_attn_keys[-1] if _is_last_in_order(guidance_attn_keys) else None)
        loss = compute_ca_loss(saved, bboxes, object_positions, guidance_attn_keys, **loss_kw) * loss_scale
        grad = torch.autograd.grad(loss, [lat])[0]
    scale = (1 - alpha_prod_t) ** 0.5
    new_lat = latents - scale * grad
    return new_lat.detach(), loss.detach(), grad.detach(), {k: v.detach() for k, v in saved.items()}


_ORDER = ([("down", i, j, 0) for i in range(3) for j in range(2)] + [("mid", 0, 0, 0)] +
          [("up", i, j, 0) for i in range(1, 4) for j in range(3)])


def _is_last_in_order(keys):
    """True when keys[-1] is the last of `keys` in execution order, so the forward may stop right after it."""
    idx = [_ORDER.index(tuple(k)) for k in keys]
    return idx[-1] == max(idx)
