"""CPU tests of the host side of the path: the PRODUCT functions
(`lvd_b200.utils.utils.scale_proportion / get_hw_from_attn_dim`, `lvd_b200.utils.parse`, `lvd_b200.utils.llm`,
`lvd_b200.utils.guidance.{refine_phrase,get_token_map,get_phrase_indices}`) bit-exact against

  * the committed golden vectors the reference produced (tests/golden/host_boxes.npz, host_parse.json;
    generator: oracle/make_golden_host.py) -- everywhere, and
  * the reference executed live (oracle/ref_loader.py) where /root/reference is mounted.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import ref_loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_scale_proportion_20000_boxes_bit_exact_vs_reference_golden():
    from lvd_b200.utils.utils import get_hw_from_attn_dim, scale_proportion
    g = np.load(os.path.join(GOLD, "host_boxes.npz"))
    boxes, res, rect = g["boxes"], g["res"], g["rect"]
    assert boxes.shape == (20000, 4)
    ties = 0
    for b, (h, w), want in zip(boxes, res, rect):
        got = scale_proportion(tuple(b.tolist()), H=int(h), W=int(w))
        assert tuple(int(v) for v in got) == tuple(int(v) for v in want), (b, h, w, got, want)
        ties += (b[0] * w) % 1.0 == 0.5
    assert ties > 2000  # the banker's-rounding cases are really in the set
    for b, (h, w), want in zip(boxes[:2000], res[:2000], g["legacy"]):
        assert tuple(scale_proportion(tuple(b.tolist()), H=int(h), W=int(w), use_legacy=True)) == tuple(want)
    for bh, bw, d, H, W in g["hw"].tolist():
        assert get_hw_from_attn_dim(attn_dim=d, base_attn_dim=(bh, bw)) == (H, W)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_scale_proportion_vs_live_reference():
    from lvd_b200.utils.utils import get_hw_from_attn_dim, scale_proportion
    from oracle.make_golden_host import ATTN_DIMS, random_boxes
    U = ref_loader.load().utils
    boxes, res = random_boxes(10000, seed=123)  # a different draw than the committed vectors
    for b, (h, w) in zip(boxes, res):
        bb = tuple(b.tolist())
        assert tuple(scale_proportion(bb, H=int(h), W=int(w))) == tuple(U.scale_proportion(bb, H=int(h), W=int(w)))
    for base, dims in ATTN_DIMS:
        for d in dims:
            assert get_hw_from_attn_dim(d, base) == tuple(U.get_hw_from_attn_dim(attn_dim=d, base_attn_dim=base))


def _digest(cond):
    h = hashlib.sha256()
    h.update(cond.prompt.encode())
    h.update(json.dumps(cond.phrases).encode())
    h.update(np.asarray(cond.boxes, dtype=np.float64).tobytes())
    return h.hexdigest()


def test_parse_mirror_bit_exact_on_the_1000_cached_layouts():
    """Responses are re-derived from the digests' prompts only where the cache exists; the committed file carries the
    responses' results, and for the first 8 per cache the full boxes."""
    from lvd_b200.utils import llm, parse
    gold = json.load(open(os.path.join(GOLD, "host_parse.json")))
    assert sum(len(v) for v in gold.values()) == 1000
    if not ref_loader.available():
        pytest.skip("LLM response caches live in /root/reference/cache (not redistributed); digests checked there")
    for name, items in gold.items():
        cache = json.load(open(os.path.join(ref_loader.REF_ROOT, "cache", name)))
        it = iter(items)
        for prompt, responses in cache.items():
            for resp in responses:
                rec = next(it)
                assert rec["prompt"] == prompt
                layout, rebuilt = llm.get_parsed_layout(prompt, max_partial_response_retries=1, override_response=resp,
                                                        json_template=False)
                assert hashlib.sha256(rebuilt.encode()).hexdigest() == rec["rebuilt_sha"]
                for frames in (24, 16):
                    c = parse.parsed_layout_to_condition(layout, tokenizer=None, height=512, width=512,
                                                         num_condition_frames=frames, verbose=False)
                    assert _digest(c) == rec[f"sha{frames}"], (name, prompt, frames)
                    if f"boxes{frames}" in rec:
                        assert c.boxes == rec[f"boxes{frames}"] and c.phrases == rec["phrases"]
                        assert c.prompt == rec["cond_prompt"]


class _WordTokenizer:
    """Whitespace tokenizer with the CLIP tokenizer's call surface (ids, <bos>/<eos>, `_convert_id_to_token`), enough to
    drive get_token_map / get_phrase_indices identically through the reference and the mirror."""
    eos_token = "<|endoftext|>"

    def __init__(self):
        self.vocab = {"<|startoftext|>": 0, self.eos_token: 1}

    def __call__(self, texts, padding="do_not_pad", max_length=77, return_tensors="np"):
        ids = []
        for t in texts:
            row = [0]
            for wd in t.lower().replace(",", " ,").replace("|", " |").split():
                row.append(self.vocab.setdefault(wd + "</w>", len(self.vocab)))
            row.append(1)
            ids.append(row)
        return {"input_ids": np.array(ids)}

    def _convert_id_to_token(self, i):
        return {v: k for k, v in self.vocab.items()}[i]


def test_phrase_indices_and_refine_phrase():
    from lvd_b200.utils import guidance as G
    tok = _WordTokenizer()
    prompt = "a brown bear and two red cars on a road, forest background"
    assert G.refine_phrase(prompt, "brown bear") == (True, "brown bear")
    assert G.refine_phrase(prompt, "bear 2") == (True, "bear")
    assert G.refine_phrase(prompt, "big red car")[0] is True and G.refine_phrase(prompt, "big red car")[1] == "cars"
    assert G.refine_phrase(prompt, "zebra")[0] is False
    pos = G.get_phrase_indices(tok, prompt, ["brown bear", "cars", "road"])
    assert pos == [[2, 3], [7], [10]]
    pos, widx = G.get_phrase_indices(tok, prompt, ["brown bear"], words=["bear"], return_word_token_indices=True)
    assert pos == [[2, 3]] and widx == [3]
    assert G.get_phrase_indices(tok, prompt, ["road"], include_eos=True) == [[10, 14]]


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_host_token_helpers_vs_live_reference():
    from lvd_b200.utils import guidance as G
    from lvd_b200.utils import llm, parse
    ns = ref_loader.load()
    ref_loader.stub_host_modules()
    from utils import parse as ref_parse  # the reference's, through sys.path set by ref_loader
    RG = ns.guidance
    tok = _WordTokenizer()
    cache = json.load(open(os.path.join(ref_loader.REF_ROOT, "cache", "cache_lvd_v0.1_gpt-4-1106-preview.json")))
    n = 0
    for prompt, responses in list(cache.items())[:60]:
        layout, _ = llm.get_parsed_layout(prompt, override_response=responses[0])
        try:
            want = ref_parse.parsed_layout_to_condition(layout, tokenizer=tok, height=512, width=512, verbose=False)
        except Exception as e:  # the reference fails on this layout (e.g. phrase not locatable): so must the mirror
            with pytest.raises(type(e)):
                parse.parsed_layout_to_condition(layout, tokenizer=tok, height=512, width=512, verbose=False)
            continue
        got = parse.parsed_layout_to_condition(layout, tokenizer=tok, height=512, width=512, verbose=False)
        assert got.prompt == want.prompt and got.boxes == want.boxes and got.phrases == want.phrases
        assert got.object_positions == want.object_positions and got.token_map == want.token_map
        for ph in got.phrases:
            assert G.refine_phrase(got.prompt, ph) == RG.refine_phrase(got.prompt, ph)
        n += 1
    assert n >= 50


def test_layout_response_parser_edge_cases():
    from lvd_b200.utils import llm
    resp = ("Reasoning: x\nFrame 1: [{'id': 0, 'name': 'cat', 'box': [1, 2, 3, 4]}] - moves right\n"
            "Frame 2: - [{'id': 0, 'name': 'cat', 'box': [2, 2, 3, 4]}]\nFrame 3: []\nFrame 4: []\nFrame 5: []\n"
            "Frame 6: []\nBackground keyword: `garden`\n")
    layout, _ = llm.get_parsed_layout("a cat", override_response=resp)
    assert layout["Frame 1"] == [{"id": 0, "name": "cat", "box": [1, 2, 3, 4]}]
    assert layout["Frame 2"][0]["box"] == [2, 2, 3, 4] and layout["Frame 3"] == []
    assert layout["Background keyword"] == "garden" and layout["Prompt"] == "a cat"
    with pytest.raises(ValueError):
        llm.get_parsed_layout("a cat", override_response="Frame 1: []\nFrame 2: []")
    with pytest.raises(NotImplementedError):
        llm.get_parsed_layout("a cat")
    j, _ = llm.get_parsed_layout("p", override_response='Response: {"Frame 1": [], "Background keyword": "sky"}',
                                 json_template=True)
    assert j == {"Frame 1": [], "Background keyword": "sky", "Prompt": "p"}
