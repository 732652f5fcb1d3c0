"""Mirror of the reference's utils/guidance.py hot-path entry point `compute_ca_lossv3` (same name, argument meaning
and error behaviour) on top of the fused CUDA energy kernel (`lvd_box_energy`).

The reference loops in Python over keys x objects x frames x tokens with ~20 tiny kernels and a host sync each
(utils/guidance.py:231-524); here one launch per guided key evaluates every (object, token, frame, head) column.
Supported: the max-based (top-k), ratio-based and CE losses, attention sync, the BoxDiff corner constraint and the
centre-of-mass terms (everything generate.py's flags and the runners' `run()` keywords can reach, generate.py:78-102,
generation/lvd.py:88-142), and the map pre-processing variants only a direct library call can select: smooth_attn
(GaussianSmoothing, utils/attn.py:91-157), attn_renorm and upsample_scale (lvd_attn_smooth / _renorm / _columns).
"""
import re
from collections.abc import Iterable

import torch

from .. import ops
from .utils import get_hw_from_attn_dim, scale_proportion

# keyword -> default of add_ca_loss_per_attn_map_to_loss (utils/guidance.py:160-193)
_DEFAULTS = dict(use_ratio_based_loss=False, use_max_based_loss=True, use_ce_based_loss=False, fg_top_p=0.2,
                 bg_top_p=0.2, fg_weight=1.0, bg_weight=1.0, eps=1.0e-2, index=None, attn_key=None,
                 exclude_bg_heads=False, smooth_attn=False, kernel_size=3, sigma=0.5, upsample_scale=1,
                 upsample_mode="bilinear", base_attn_dim=(40, 72), attn_renorm=False, num_tokens=None,
                 renorm_scale=2.0, attn_sync_weight=0.0, boxdiff_loss_scale=0.0, boxdiff_normed=True, boxdiff_L=1,
                 com_loss_scale=0.0, verbose=False)


class _Pre(dict):
    __getattr__ = dict.__getitem__

    @property
    def active(self):
        return self.smooth is not None or self.renorm is not None or self.upsample_scale != 1


def energy_options(kw, with_pre=False):
    """Validate the loss keywords of `add_ca_loss_per_attn_map_to_loss` and translate them into `ops.box_energy`
    options (and, with_pre=True, the map pre-processing plan: smoothing weights, renormalisation, upsampling).  Same
    errors as the reference for its own assertions and for the combinations that crash there."""
    unknown = [k for k in kw if k not in _DEFAULTS]
    if unknown:
        raise TypeError(f"add_ca_loss_per_attn_map_to_loss() got an unexpected keyword argument '{unknown[0]}'")
    o = dict(_DEFAULTS)
    o.update(kw)
    assert not o["exclude_bg_heads"], "exclude_bg_heads is deprecated"
    if o["use_ratio_based_loss"]:
        kind = ops.LOSS_RATIO
    elif o["use_max_based_loss"]:
        kind = ops.LOSS_MAX
    elif o["use_ce_based_loss"]:
        kind = ops.LOSS_CE
        if o["attn_sync_weight"] != 0.0 or o["boxdiff_loss_scale"] > 0 or o["com_loss_scale"] > 0:
            raise NotImplementedError("lvd_b200 compute_ca_lossv3: the CE loss combined with attn_sync / BoxDiff / "
                                      "centre-of-mass terms is not implemented")
    else:
        raise ValueError("Unknown loss: no loss selected")
    pre = _Pre(smooth=None, renorm=None, upsample_scale=o["upsample_scale"])
    if o["smooth_attn"]:
        pre["smooth"] = ops.gaussian_weights(o["kernel_size"], o["sigma"])
    if o["attn_renorm"]:
        assert o["attn_sync_weight"] == 0.0, "attn_sync with attn_renorm not implemented together"
        if o["num_tokens"] is None:
            raise TypeError("attn_renorm needs num_tokens")  # the reference fails on `None - 1` (:221)
        pre["renorm"] = (int(o["num_tokens"]), float(o["renorm_scale"]))
    if o["upsample_scale"] != 1:
        if o["upsample_mode"] != "bilinear":
            raise NotImplementedError("lvd_b200 compute_ca_lossv3: upsample_mode other than 'bilinear'")
        if kind != ops.LOSS_RATIO or o["attn_sync_weight"] != 0.0:
            # guidance.py:297-347,414: the top-k branches multiply the un-flattened (heads, 1, H, W) interpolation result
            # with the flattened mask, and the sync term views it at the un-scaled size -- both raise in the reference
            raise RuntimeError("upsample_scale != 1 only runs with use_ratio_based_loss (and without attn_sync) in the "
                               "reference: the other branches fail on mismatching shapes (utils/guidance.py:297-347)")
    opts = dict(loss_kind=kind, fg_top_p=float(o["fg_top_p"]), bg_top_p=float(o["bg_top_p"]),
                fg_weight=float(o["fg_weight"]), bg_weight=float(o["bg_weight"]), eps=float(o["eps"]),
                k_min=int(o["upsample_scale"] * o["upsample_scale"]),
                attn_sync_weight=float(o["attn_sync_weight"]),
                boxdiff_loss_scale=float(max(o["boxdiff_loss_scale"], 0.0)), boxdiff_normed=bool(o["boxdiff_normed"]),
                boxdiff_L=int(o["boxdiff_L"]), com_loss_scale=float(max(o["com_loss_scale"], 0.0)))
    return (opts, pre) if with_pre else opts


def map_energy(attn_map, loss, bboxes, object_positions, base_attn_dim, opts, pre, *, out_scale, want_grad=False,
               grad_scale=1.0, tables=None):
    """Energy of ONE saved map [F, heads, hw, n_text] accumulated into `loss`; returns grad_scale * dE/dP as a dense
    fp16 map when want_grad.  Plain options: one `lvd_box_energy` launch on the map itself.  With smoothing /
    renormalisation / upsampling: pre-process (lvd_attn_smooth -> lvd_attn_renorm -> lvd_attn_columns), evaluate the
    energy on the gathered columns, and run the adjoints in reverse."""
    n_f, heads, hw, n_text = attn_map.shape
    H_attn, W_attn = get_hw_from_attn_dim(attn_dim=hw, base_attn_dim=base_attn_dim)
    s = pre.upsample_scale
    H, W = int(H_attn * s), int(W_attn * s)
    cols, boxes, wts = tables if tables is not None else build_slots(bboxes, object_positions, n_f, H, W,
                                                                     attn_map.device)
    if not pre.active:
        dp = torch.zeros_like(attn_map) if want_grad else None
        ops.box_energy(attn_map, loss, boxes, wts, frames=n_f, heads=heads, H=H, W=W, slot_col=cols, n_text=n_text,
                       out_scale=out_scale, dp=dp, grad_scale=grad_scale, **opts)
        return dp
    m = attn_map
    if pre.smooth is not None:
        m = ops.attn_smooth(m, pre.smooth)
    sm = m
    if pre.renorm is not None:
        m = ops.attn_renorm(m, *pre.renorm)
        cols = cols - 1  # SOT was dropped (utils/guidance.py:281-285)
    c = ops.attn_columns(m, cols, H_attn, W_attn, H, W)
    dc = torch.empty_like(c) if want_grad else None
    ops.box_energy(c, loss, boxes, wts, frames=n_f, heads=heads, H=H, W=W, out_scale=out_scale, dp=dc,
                   grad_scale=grad_scale, **opts)
    if not want_grad:
        return None
    d = ops.attn_columns_bwd(dc, cols, m.shape[3], H_attn, W_attn, H, W,
                             half_out=pre.renorm is None and pre.smooth is None)
    if pre.renorm is not None:
        d = ops.attn_renorm_bwd(m, d, sm.shape[3], pre.renorm[0], pre.renorm[1], half_out=pre.smooth is None)
    if pre.smooth is not None:
        d = ops.attn_smooth_bwd(d, pre.smooth, half_out=True)
    return d


def padded_slot_count(n):
    """Slot tables of the graph-replayed guidance pass are padded to a power of two so that layouts with different
    object / token counts share a captured graph (padding slots carry weight 0 and are skipped by the kernel)."""
    p = 1
    while p < n:
        p *= 2
    return p


def build_slots(bboxes, object_positions, n_frames, H, W, device, pad_to=None, fold_objects=False):
    """Flatten (object, token) pairs into slots and rasterise every per-frame box at resolution (H, W).
    Returns (cols [n_slots] int32, boxes [n_slots, F, 4] int32, weights [n_slots] fp32) on `device`.
    weights = 1 / len(tokens of the object) (utils/guidance.py:524), additionally / n_objects when fold_objects (the
    compute_ca_lossv3 normalisation :569-572 moved into the table, so the kernel's out_scale no longer depends on the
    layout); pad_to appends weight-0 slots."""
    cols, rects, wts = [], [], []
    n_obj = len(bboxes)
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
            wts.append(1.0 / len(toks) / (n_obj if fold_objects else 1))
    if pad_to is not None:
        assert pad_to >= len(cols)
        while len(cols) < pad_to:
            cols.append(cols[0] if cols else 0)
            rects.append([(0, 0, 0, 0)] * n_frames)
            wts.append(0.0)
    return (torch.tensor(cols, dtype=torch.int32, device=device),
            torch.tensor(rects, dtype=torch.int32, device=device).reshape(len(cols), n_frames, 4).contiguous(),
            torch.tensor(wts, dtype=torch.float32, device=device))


def build_slots_batch(layouts, n_frames, H, W, device, pad_to=None):
    """Slot tables of a batch of prompts guided in one pass: `layouts` = [(bboxes, object_positions)] per prompt.
    Returns (cols, boxes, weights, frame0, loss_index): the tables of `build_slots` (1 / n_objects folded in, per
    prompt) concatenated over the prompts, plus per slot the first frame of its prompt inside the batched frame axis
    and the index of its prompt's loss accumulator."""
    cols, boxes, wts, fr0, lidx = [], [], [], [], []
    for b, (bb, pos) in enumerate(layouts):
        if len(bb) == 0:
            continue
        c, r, w = build_slots(bb, pos, n_frames, H, W, torch.device("cpu"), fold_objects=True)
        cols.append(c)
        boxes.append(r)
        wts.append(w)
        fr0 += [b * n_frames] * c.numel()
        lidx += [b] * c.numel()
    cols, boxes, wts = torch.cat(cols), torch.cat(boxes), torch.cat(wts)
    n = cols.numel()
    if pad_to is not None and pad_to > n:
        k = pad_to - n
        cols = torch.cat([cols, cols[:1].expand(k)])
        boxes = torch.cat([boxes, torch.zeros((k, n_frames, 4), dtype=torch.int32)])
        wts = torch.cat([wts, torch.zeros(k)])
        fr0 += [0] * k
        lidx += [0] * k
    dev = torch.device(device)
    return (cols.contiguous().to(dev), boxes.contiguous().to(dev), wts.contiguous().to(dev),
            torch.tensor(fr0, dtype=torch.int32, device=dev), torch.tensor(lidx, dtype=torch.int32, device=dev))


def _lossv3(saved_attn, bboxes, object_positions, guidance_attn_keys, want_grad, grad_scale, kwargs):
    opts, pre = energy_options(kwargs, with_pre=True)
    dev = torch.device("cuda", torch.cuda.current_device())
    loss = torch.zeros((), dtype=torch.float32, device=dev)
    grads = {}
    object_number = len(bboxes)
    if object_number == 0:
        return loss, grads
    base_attn_dim = kwargs.get("base_attn_dim", (40, 72))
    num_attn = len(guidance_attn_keys)
    for attn_key in guidance_attn_keys:
        attn_map = saved_attn[attn_key]
        if not attn_map.is_cuda:
            attn_map = attn_map.cuda()
        attn_map = attn_map.squeeze(dim=0).contiguous()
        grads[attn_key] = map_energy(attn_map, loss, bboxes, object_positions, base_attn_dim, opts, pre,
                                     out_scale=1.0 / (object_number * num_attn), want_grad=want_grad,
                                     grad_scale=grad_scale)
    return loss, grads


def compute_ca_lossv3(saved_attn, bboxes, object_positions, guidance_attn_keys, index=None, verbose=False, **kwargs):
    """utils/guidance.py:529-574.  `saved_attn[key]` is the map the attention processor stored:
    (frames, heads, hw, n_text) fp16 on the GPU.  Returns the (un-scaled) loss as a 0-d fp32 CUDA tensor."""
    return _lossv3(saved_attn, bboxes, object_positions, guidance_attn_keys, False, 1.0, kwargs)[0]


def compute_ca_lossv3_with_grad(saved_attn, bboxes, object_positions, guidance_attn_keys, grad_scale=1.0, **kwargs):
    """The loss of compute_ca_lossv3 and grad_scale * dLoss/dP per key (dense fp16 maps): what torch.autograd.grad gives
    a caller of the reference who differentiates the loss w.r.t. the saved maps (models/pipelines.py:119)."""
    return _lossv3(saved_attn, bboxes, object_positions, guidance_attn_keys, True, grad_scale, kwargs)


# ------------------------------------------------------------------------------------------------ host: prompt tokens
def _plural(word):
    """`inflect.engine().plural` when inflect is installed (the reference's dependency, utils/guidance.py:12-14);
    otherwise the regular English rules -- only reached for phrases whose last word is not in the prompt."""
    try:
        import inflect
        return inflect.engine().plural(word)
    except Exception:
        if re.search(r"(s|x|z|ch|sh)$", word):
            return word + "es"
        if re.search(r"[^aeiou]y$", word):
            return word[:-1] + "ies"
        return word + "s"


def get_token_map(tokenizer, prompt, verbose=False, padding="do_not_pad"):
    """Token strings of `prompt` incl. <bos>/<eos> (utils/guidance.py:18-41)."""
    ids = tokenizer([prompt], padding=padding, max_length=77, return_tensors="np")["input_ids"][0]
    out = []
    for i, tok_id in enumerate(ids.tolist()):
        tok = tokenizer._convert_id_to_token(tok_id)
        if verbose:
            print(f"{i}, {tok} ({tok_id})")
        out.append(tok)
    return out


def refine_phrase(prompt, phrase, verbose=False):
    """Find a form of `phrase` that occurs in `prompt` as whole words: as is, without its digits, its last word, that
    word's plural (utils/guidance.py:44-66).  Returns (found, refined_phrase)."""
    def present(p):
        return re.search(r"\b" + p + r"\b", prompt) is not None

    cand = phrase
    if present(cand):
        return True, cand
    cand = cand.strip("0123456789 ")
    if present(cand):
        return True, cand
    last = cand.split(" ")[-1]
    if verbose:
        print(f'**Phrase "{cand}" is not part of the prompt "{prompt}", using the last word "{last}" only**')
    cand = last
    if present(cand):
        return True, cand
    cand = _plural(cand)
    if verbose:
        print(f'**Phrase is still not part of the prompt "{prompt}", using plural "{cand}" instead**')
    if present(cand):
        return True, cand
    print(f'**Phrase is still not part of the prompt "{prompt}"**')
    return False, cand


def get_phrase_indices(tokenizer, prompt, phrases, verbose=False, words=None, include_eos=False, token_map=None,
                       return_word_token_indices=False):
    """Token positions of every phrase inside the tokenised prompt (utils/guidance.py:69-151): the phrase is tokenised
    on its own and located as a sub-sequence of the prompt's token strings."""
    if token_map is None:
        token_map = get_token_map(tokenizer, prompt=prompt, verbose=verbose, padding="do_not_pad")
    joined = " ".join(token_map)
    positions, word_indices = [], []
    for i, phrase in enumerate(phrases):
        found, refined = refine_phrase(prompt, phrase)
        assert found, "phrase {obj} not found in the prompt {prompt}"
        sub = get_token_map(tokenizer, prompt=refined, verbose=verbose, padding="do_not_pad")[1:-1]  # no <bos>/<eos>
        sub_joined = " ".join(sub)
        if verbose:
            print("Full str:", joined, "Substr:", sub_joined, "Phrase:", phrases)
        first = len(joined[: joined.index(sub_joined) - 1].split(" "))  # tokens before the match
        pos = list(range(first, first + len(sub)))
        if include_eos:
            pos.append(token_map.index(tokenizer.eos_token))
        positions.append(pos)
        if return_word_token_indices:
            if words is None:
                idx = positions[0][-1]
                print(f'Picking the last token "{token_map[idx]}" ({idx}) as attention token for extracting attention '
                      "for SAM, which might not be the right one")
            else:
                wt = get_token_map(tokenizer, prompt=words[i], verbose=verbose, padding="do_not_pad")
                idx = first + sub.index(wt[-2])
            if verbose:
                print("so_token_index:", idx)
            word_indices.append(idx)
    if return_word_token_indices:
        return positions, word_indices
    return positions
