"""VAE decode at the Zeroscope shape: 24 latent frames 40x72 -> 24 RGB frames 320x576 (SD-1.x decoder, 49.5 M params,
~1.24 TFLOP per frame), CUDA-event timed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import synthetic  # noqa: E402
from lvd_b200.vae import AutoencoderKL  # noqa: E402

dev = torch.device("cuda:0")
cfg = synthetic.VAE_SD
vae = AutoencoderKL().to(dev)
vae.load_state_dict({k: v.half() for k, v in synthetic.random_vae_state_dict(cfg, seed=0).items()})
frames, h, w = int(os.environ.get("FRAMES", "24")), int(os.environ.get("H", "40")), int(os.environ.get("W", "72"))
z = torch.randn((frames, 4, h, w), generator=torch.Generator().manual_seed(1)).half().to(dev)
for sl in (0, 8, 4):
    if sl:
        vae.enable_slicing(sl)
    for _ in range(2):
        y = vae.decode(z).sample
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        y = vae.decode(z).sample
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 3
    print(f"decode {frames} frames {h}x{w} -> {tuple(y.shape)} slice={sl or 'all'}: {ms:.1f} ms "
          f"({frames / ms * 1e3:.0f} frames/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB), "
          f"finite={bool(torch.isfinite(y.float()).all())}", flush=True)
