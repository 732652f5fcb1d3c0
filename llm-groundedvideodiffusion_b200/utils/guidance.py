"""Mirror of the reference's utils/guidance.py hot-path entry point `compute_ca_lossv3` (same name, argument meaning
and error behaviour) on top of the fused CUDA energy kernel (`lvd_box_energy`).

The reference loops in Python over keys x objects x frames x tokens with ~20 tiny kernels and a host sync each
(utils/guidance.py:231-524); here one launch per guided key evaluates every (object, token, frame, head) column.
Supported: the default max-based (top-k) loss and the centre-of-mass terms.  The deprecated / CLI-unreachable
variants (ratio-based, CE, smoothing, renorm, upsample, attn-sync, BoxDiff) raise NotImplementedError.
"""
from collections.abc import Iterable

import torch

from .. import ops
from .utils import get_hw_from_attn_dim, scale_proportion

_UNSUPPORTED = dict(use_ratio_based_loss=False, use_ce_based_loss=False, exclude_bg_heads=False, smooth_attn=False,
                    upsample_scale=1, attn_renorm=False, attn_sync_weight=0.0, boxdiff_loss_scale=0.0)


def build_slots(bboxes, object_positions, n_frames, H, W, device):
    """Flatten (object, token) pairs into slots and rasterise every per-frame box at resolution (H, W).
    Returns (cols [n_slots] int32, boxes [n_slots, F, 4] int32, weights [n_slots] fp32) on `device`."""
    cols, rects, wts = [], [], []
    for obj_idx, obj_boxes in enumerate(bboxes):
        assert n_frames == len(obj_boxes), (
            f"Number of frames {n_frames} mismatches with number of frames in box condition {len(obj_boxes)}")
        per_frame = []
        for frame_boxes in obj_boxes:
            assert not isinstance(frame_boxes[0], Iterable), \
                "We currently process different boxes with the same name separately."
            per_frame.append(scale_proportion(frame_boxes, H=H, W=W))
        toks = object_positions[obj_idx]
        for tok in toks:
            cols.append(int(tok))
            rects.append(per_frame)
            wts.append(1.0 / len(toks))
    return (torch.tensor(cols, dtype=torch.int32, device=device),
            torch.tensor(rects, dtype=torch.int32, device=device).reshape(len(cols), n_frames, 4).contiguous(),
            torch.tensor(wts, dtype=torch.float32, device=device))


def _check_kwargs(kw):
    for k, default in _UNSUPPORTED.items():
        if k in kw and kw[k] != default:
            raise NotImplementedError(f"lvd_b200 compute_ca_lossv3: option {k}={kw[k]!r} is not implemented "
                                      "(only the default max-based loss and com_loss_scale are)")
    if kw.get("use_max_based_loss", True) is not True:
        raise ValueError("Unknown loss: no loss selected")


def compute_ca_lossv3(saved_attn, bboxes, object_positions, guidance_attn_keys, index=None, verbose=False, **kwargs):
    """utils/guidance.py:529-574.  `saved_attn[key]` is the map the attention processor stored:
    (frames, heads, hw, n_text) fp16 on the GPU.  Returns the (un-scaled) loss as a 0-d fp32 CUDA tensor."""
    _check_kwargs(kwargs)
    dev = torch.device("cuda", torch.cuda.current_device())
    loss = torch.zeros((), dtype=torch.float32, device=dev)
    object_number = len(bboxes)
    if object_number == 0:
        return loss
    base_attn_dim = kwargs.get("base_attn_dim", (40, 72))
    num_attn = len(guidance_attn_keys)
    for attn_key in guidance_attn_keys:
        attn_map = saved_attn[attn_key]
        if not attn_map.is_cuda:
            attn_map = attn_map.cuda()
        attn_map = attn_map.squeeze(dim=0).contiguous()
        n_f, heads, hw, n_text = attn_map.shape
        H, W = get_hw_from_attn_dim(attn_dim=hw, base_attn_dim=base_attn_dim)
        cols, boxes, wts = build_slots(bboxes, object_positions, n_f, H, W, attn_map.device)
        ops.box_energy(attn_map, loss, boxes, wts, frames=n_f, heads=heads, H=H, W=W, slot_col=cols, n_text=n_text,
                       fg_top_p=kwargs.get("fg_top_p", 0.2), bg_top_p=kwargs.get("bg_top_p", 0.2),
                       fg_weight=kwargs.get("fg_weight", 1.0), bg_weight=kwargs.get("bg_weight", 1.0),
                       com_loss_scale=kwargs.get("com_loss_scale", 0.0),
                       out_scale=1.0 / (object_number * num_attn))
    return loss
