"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds the CPU checker for the hot path: a plain PyTorch fp32 restatement of the reference algorithm
(`unet3d_ref.py`, `guidance_ref.py`, `scheduler_ref.py`), the deterministic synthetic weights (`weights.py`), and --
usable only inside the builder container where /root/reference is mounted -- a loader that imports the reference's
own files verbatim through a `diffusers` shim (`ref_loader.py`, `shim/`) to pin the restatement and to generate the
golden vectors under tests/golden/ (`make_golden.py`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product package (lvd_b200/) never does: its ops fail loudly without the CUDA library.

Parity status: the reference ships NO tests, golden vectors or fixtures, so the pin is
"reference code executed here": tests/golden/*.npz are outputs of the reference's own modules run in this container
by oracle/make_golden.py; the un-vendored diffusers 0.27.2 pieces inside them are restated from the published
algorithm (shim/diffusers/models/resnet.py, embeddings.py) and cannot be diffed offline -- for those pieces parity
is "unpinned" in the strict sense and DESIGN.md says so.
"""
