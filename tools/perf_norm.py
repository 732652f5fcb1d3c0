"""GroupNorm / LayerNorm at the Zeroscope shapes: achieved GB/s (algorithmic bytes = read x + write y).  Launches are
issued back to back (the GPU never idles on host latency) over rotating buffers larger than L2 ("cold") and over one
buffer ("warm": as in the UNet, where the producer just wrote the input)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lvd_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def bench(fn_of_i, n=24):
    """n launches captured into one CUDA graph: no host latency between kernels."""
    for i in range(4):
        fn_of_i(i)
    torch.cuda.synchronize()
    if os.environ.get("PERF_NORM_EAGER"):  # under ncu: a few eager launches, no graph
        return 1.0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn_of_i(i)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def run(name, rows, c, make_fn, passes):
    nbuf = max(2, int(400e6 // (rows * c * 2 * 2)) + 1)
    xs = [torch.randn(rows, c, device=dev).half() for _ in range(nbuf)]
    ys = [torch.empty(rows, c, device=dev, dtype=torch.float16) for _ in range(nbuf)]
    cold = bench(lambda i: make_fn(xs[i % nbuf], ys[i % nbuf]))
    warm = bench(lambda i: make_fn(xs[0], ys[0]))
    gb = rows * c * 2 * 2 / 1e9
    print(f"{name:34s} cold {cold*1e3:7.1f} us {gb/cold*1e3:6.0f} GB/s | warm {warm*1e3:7.1f} us {gb/warm*1e3:6.0f} GB/s"
          f"   ({passes} HBM passes; 2-pass floor at 6.5 TB/s = {gb/6.5*1e3:5.1f} us)", flush=True)


for (rows, c) in [(138240, 320), (138240, 512), (34560, 640), (8640, 1280), (2160, 1280)]:
    g_ = torch.randn(c, device=dev).half()
    b_ = torch.randn(c, device=dev).half()
    run(f"layernorm rows={rows} C={c}", rows, c, lambda x, y: ops.layernorm(x, g_, b_, out=y), 2)
for (samples, rps, c) in [(48, 2880, 320), (2, 69120, 320), (48, 2880, 640), (48, 2880, 960), (48, 720, 640),
                          (2, 17280, 640), (48, 720, 1280), (48, 180, 1280), (2, 4320, 1280), (48, 180, 2560),
                          (48, 45, 1280), (2, 1080, 1280)]:
    g_ = torch.randn(c, device=dev).half()
    b_ = torch.randn(c, device=dev).half()
    run(f"groupnorm s={samples} rows/s={rps} C={c}", samples * rps, c,
        lambda x, y: ops.groupnorm(x, samples, g_, b_, 1e-5, silu=True, out=y), 3)
