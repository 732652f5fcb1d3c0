"""Host preparation of one work item: LLM layout -> `Condition` (prompt, per-object per-frame boxes, phrases, token
positions).  Mirror of the hot-path part of the reference's utils/parse.py (:220-363: `convert_box`, `interpolate_box`,
`parsed_layout_to_condition`; same names, arguments and results) -- the plotting helpers of that file (:21-217) are
visualisation and out of scope.

Bit-exactness: the interpolated boxes feed `scale_proportion`'s Python `round()`, so the float64 arithmetic here is kept
operation-for-operation what numpy evaluates in the reference (`np.interp` on `t/(n_in-1) + r` knots queried at
`linspace(0, repeat-1e-5, n_out)`); tests/test_host_prep.py checks all 1000 cached GPT layouts against the live
reference and against committed digests.
"""
from collections import namedtuple

import numpy as np

from . import guidance

Condition = namedtuple("Condition", ["prompt", "boxes", "phrases", "object_positions", "token_map"])


def convert_box(box, height, width):
    """(x, y, w, h) in layout pixels -> (x_min, y_min, x_max, y_max) as fractions (utils/parse.py:220-227)."""
    x_min, y_min = box[0] / width, box[1] / height
    return x_min, y_min, x_min + box[2] / width, y_min + box[3] / height


def interpolate_box(box, num_input_frames=6, num_output_frames=24, repeat=1):
    """`box` maps key-frame index -> (x_min, y_min, x_max, y_max); returns `num_output_frames` boxes, all-zero where the
    object is absent from the nearest key frame (utils/parse.py:230-248)."""
    keys = np.sort(list(box.keys()))
    knots = np.concatenate([keys / (num_input_frames - 1) + r for r in range(repeat)])
    query = np.linspace(0, repeat - 1e-5, num_output_frames)  # the 1e-5 keeps the last query inside the last period
    present = np.isin(np.floor((query % 1.0) * num_input_frames), keys)
    out = np.zeros((num_output_frames, 4))
    for c in range(4):
        vals = np.array([box[k][c] for k in keys] * repeat)
        out[:, c] = np.interp(query, knots, vals) * present
    return out.tolist()


def parsed_layout_to_condition(parsed_layout, height, width, num_parsed_layout_frames=6, num_condition_frames=24,
                               interpolate_boxes=True, tokenizer=None, output_phrase_per_timestep=False,
                               add_background_to_prompt=True, strip_phrases=False, verbose=False):
    """utils/parse.py:251-363.  Objects are keyed by their `id`; one that is missing from a key frame disappears there."""
    prompt = parsed_layout["Prompt"]
    if add_background_to_prompt and parsed_layout["Background keyword"]:
        prompt += f", {parsed_layout['Background keyword']} background"
    order, phrase_of, box_of = [], {}, {}
    for t in range(num_parsed_layout_frames):
        for obj in parsed_layout[f"Frame {t + 1}"]:
            oid = obj["id"]
            label = obj["name"] if "name" in obj else obj["keyword"]
            if oid not in box_of:
                order.append(oid)
                box_of[oid] = {}
                phrase_of[oid] = {} if output_phrase_per_timestep else label  # first occurrence names the object
            box_of[oid][t] = convert_box(obj["box"], height=height, width=width)
            if output_phrase_per_timestep:
                phrase_of[oid][t] = label
    boxes = [box_of[i] for i in order]
    phrases = [phrase_of[i] for i in order]
    before = boxes
    if interpolate_boxes:
        assert not output_phrase_per_timestep, "box interpolation with phrase per timestep is not implemented"
        rep = parsed_layout.get("Repeat", 1)
        boxes = [interpolate_box(b, num_parsed_layout_frames, num_condition_frames, repeat=rep) for b in boxes]
    token_map = object_positions = None
    if tokenizer is not None:
        for phrase in phrases:
            found, _ = guidance.refine_phrase(prompt, phrase, verbose=True)
            if not found:  # make the phrase attendable: append it behind a "|" separator
                prompt += "| " + phrase
                print(f'**Adding {phrase} to the prompt. Using prompt: "{prompt}"')
        token_map = guidance.get_token_map(tokenizer, prompt=prompt, verbose=verbose, padding="do_not_pad")
        object_positions = guidance.get_phrase_indices(tokenizer, prompt, phrases, token_map=token_map, verbose=verbose)
    if verbose:
        print("prompt:", prompt)
        print("boxes (before interpolation):", before)
        if verbose >= 2:
            print("boxes (after interpolation):", np.round(np.array(boxes), 2))
        print("phrases:", phrases)
        if object_positions is not None:
            print("object_positions:", object_positions)
    if strip_phrases:
        phrases = [p.strip("1234567890 ") for p in phrases]
    return Condition(prompt, boxes, phrases, object_positions, token_map)
