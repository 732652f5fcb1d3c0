"""The two hot-path helpers of the reference's utils/utils.py (same names, same arithmetic): box -> integer mask
rectangle at an attention resolution.  Host integer work; must be bit-exact (Python banker's `round`)."""
import math


def scale_proportion(obj_box, H, W, use_legacy=False):
    """utils/utils.py:82-103."""
    if use_legacy:
        return int(obj_box[0] * W), int(obj_box[1] * H), int(obj_box[2] * W), int(obj_box[3] * H)
    x_min, y_min = round(obj_box[0] * W), round(obj_box[1] * H)
    box_w, box_h = round((obj_box[2] - obj_box[0]) * W), round((obj_box[3] - obj_box[1]) * H)
    x_max, y_max = x_min + box_w, y_min + box_h
    x_min, y_min = max(x_min, 0), max(y_min, 0)
    x_max, y_max = min(x_max, W), min(y_max, H)
    return x_min, y_min, x_max, y_max


def get_hw_from_attn_dim(attn_dim, base_attn_dim):
    """utils/utils.py:253-256."""
    scale = int(math.sqrt((base_attn_dim[0] * base_attn_dim[1]) / attn_dim))
    return base_attn_dim[0] // scale, base_attn_dim[1] // scale
