"""TEST INFRASTRUCTURE: golden vectors of the HOST side of the path, produced by the reference executed verbatim
(oracle/ref_loader.py) in the builder container:

  tests/golden/host_boxes.npz   20 000 boxes x attention resolutions -> `utils.utils.scale_proportion` rectangles
                                (incl. products that land exactly on .5 to exercise Python's banker's rounding) and
                                the `get_hw_from_attn_dim` table of every (base_attn_dim, level) the runners use
  tests/golden/host_parse.json  the 2 x 500 cached GPT layouts (cache/cache_lvd_v0.1_*.json) parsed by the reference's
                                `get_parsed_layout` + `parsed_layout_to_condition` at 24 and 16 frames: per response a
                                SHA-256 over the float64 boxes / phrases / prompt, full values for the first 8

Run:  python -m oracle.make_golden_host     (needs /root/reference; never runs on the GPU box)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
RESOLUTIONS = [(40, 72), (20, 36), (10, 18), (5, 9), (32, 32), (16, 16), (8, 8), (4, 4), (64, 64)]
ATTN_DIMS = [((40, 72), [2880, 720, 180, 45]), ((32, 32), [1024, 256, 64, 16]), ((64, 64), [4096, 1024, 256, 64])]
CACHES = ["cache_lvd_v0.1_gpt-4-1106-preview.json", "cache_lvd_v0.1_gpt-3.5-turbo.json"]


def random_boxes(n, seed=0):
    """Half random floats (some outside [0,1], some degenerate / all-zero), half k/(2*W)-style coordinates whose
    products with W or H are exact .5 ties."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4), dtype=np.float64)
    res = np.zeros((n, 2), dtype=np.int64)
    for i in range(n):
        H, W = RESOLUTIONS[rng.integers(len(RESOLUTIONS))]
        res[i] = (H, W)
        mode = rng.integers(4)
        if mode == 0:
            x0, y0 = rng.uniform(-0.1, 1.0, 2)
            out[i] = (x0, y0, x0 + rng.uniform(0, 0.8), y0 + rng.uniform(0, 0.8))
        elif mode == 1:  # ties: (2k+1)/(2W)
            x0 = (2 * rng.integers(0, W) + 1) / (2 * W)
            y0 = (2 * rng.integers(0, H) + 1) / (2 * H)
            out[i] = (x0, y0, x0 + (2 * rng.integers(0, W) + 1) / (2 * W), y0 + (2 * rng.integers(0, H) + 1) / (2 * H))
        elif mode == 2:  # layout-style: integer pixels of a 512 canvas
            x, y, w, h = rng.integers(0, 512, 4)
            out[i] = (x / 512, y / 512, x / 512 + w / 512, y / 512 + h / 512)
        else:
            out[i] = 0.0 if rng.integers(2) else rng.uniform(0, 1, 4)
    return out, res


def digest(cond):
    h = hashlib.sha256()
    h.update(cond.prompt.encode())
    h.update(json.dumps(cond.phrases).encode())
    h.update(np.asarray(cond.boxes, dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    ns = ref_loader.load()
    U = ns.utils
    boxes, res = random_boxes(20000)
    rect = np.array([U.scale_proportion(tuple(b.tolist()), H=int(h), W=int(w)) for b, (h, w) in zip(boxes, res)],
                    dtype=np.int64)
    legacy = np.array([U.scale_proportion(tuple(b.tolist()), H=int(h), W=int(w), use_legacy=True)
                       for b, (h, w) in zip(boxes[:2000], res[:2000])], dtype=np.int64)
    hw = np.array([[b[0], b[1], d, *U.get_hw_from_attn_dim(attn_dim=d, base_attn_dim=b)] for b, ds in ATTN_DIMS
                   for d in ds], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "host_boxes.npz"), boxes=boxes, res=res, rect=rect, legacy=legacy, hw=hw)

    ref_loader.stub_host_modules()
    from utils import parse as ref_parse
    from utils.llm import get_parsed_layout as ref_get_parsed_layout
    out = {}
    for name in CACHES:
        cache = json.load(open(os.path.join(ref_loader.REF_ROOT, "cache", name)))
        items = []
        n = 0
        for prompt, responses in cache.items():
            for resp in responses:
                layout, rebuilt = ref_get_parsed_layout(prompt, max_partial_response_retries=1, override_response=resp,
                                                        json_template=False)
                rec = {"prompt": prompt, "rebuilt_sha": hashlib.sha256(rebuilt.encode()).hexdigest()}
                for frames in (24, 16):
                    c = ref_parse.parsed_layout_to_condition(layout, tokenizer=None, height=512, width=512,
                                                             num_condition_frames=frames, verbose=False)
                    rec[f"sha{frames}"] = digest(c)
                    if n < 8:
                        rec[f"boxes{frames}"] = c.boxes
                        rec["phrases"] = c.phrases
                        rec["cond_prompt"] = c.prompt
                items.append(rec)
                n += 1
        out[name] = items
        print(name, len(items), "responses")
    json.dump(out, open(os.path.join(GOLD, "host_parse.json"), "w"))


if __name__ == "__main__":
    main()
