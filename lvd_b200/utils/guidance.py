"""Mirror of the reference's utils/guidance.py hot-path entry point `compute_ca_lossv3` (same name, argument meaning
and error behaviour) on top of the fused CUDA energy kernel (`lvd_box_energy`).

The reference loops in Python over keys x objects x frames x tokens with ~20 tiny kernels and a host sync each
(utils/guidance.py:231-524); here one launch per guided key evaluates every (object, token, frame, head) column.
Supported: the default max-based (top-k) loss and the centre-of-mass terms.  The deprecated / CLI-unreachable
variants (ratio-based, CE, smoothing, renorm, upsample, attn-sync, BoxDiff) raise NotImplementedError.
"""
import re
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
