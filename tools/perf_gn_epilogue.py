"""conv -> GroupNorm(+SiLU) pairs with the statistics taken from the conv's epilogue vs the two-pass GroupNorm, at the
shapes of the Zeroscope forward (two videos) and of the guidance pass (one video).  Prints microseconds per pair
(CUDA events around 20 pairs, inputs rotated over 3 buffers so nothing but the conv output is L2-warm).
   python tools/perf_gn_epilogue.py            # timings
   ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/perf_gn_epilogue.py 1   # kernel split
"""
import sys

import torch

sys.path.insert(0, ".")
from lvd_b200 import ops  # noqa: E402

SHAPES = [  # kind, (b, f, hw, c) or (n_img, h, w, c), samples
    ("tconv", (2, 24, 2880, 320), 2),
    ("tconv", (1, 24, 2880, 320), 1),
    ("tconv", (2, 24, 720, 640), 2),
    ("tconv", (2, 24, 180, 1280), 2),
    ("conv2d", (48, 40, 72, 320), 48),
    ("conv2d", (48, 20, 36, 640), 48),
    ("conv2d", (24, 40, 72, 320), 24),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    for kind, shape, samples in SHAPES:
        c = shape[-1]
        if kind == "conv2d":
            n_img, h, w, _ = shape
            rows = n_img * h * w
            wk = (torch.randn((c, 9 * c), generator=g) * 0.02).half().to(dev)
            xs = [torch.randn((rows, c), generator=g).half().to(dev) for _ in range(3)]
            run = lambda x, gn: ops.conv2d_3x3(x, n_img, h, w, wk, gn=gn)  # noqa: E731
        else:
            b, f, hw, _ = shape
            rows = b * f * hw
            wk = (torch.randn((c, 3 * c), generator=g) * 0.03).half().to(dev)
            xs = [torch.randn((rows, c), generator=g).half().to(dev) for _ in range(3)]
            run = lambda x, gn: ops.conv_temporal(x, b, f, hw, wk, gn=gn)  # noqa: E731
        gamma = torch.ones(c).half().to(dev)
        beta = torch.zeros(c).half().to(dev)
        out = torch.empty((rows, c), dtype=torch.float16, device=dev)
        line = "%-7s %-22s s=%-3d" % (kind, "x".join(map(str, shape)), samples)
        for label, gn in (("two-pass", None), ("epilogue", (samples, 32))):
            taken = [True]

            def pair(i):
                y = run(xs[i % 3], gn)
                if gn is not None and getattr(y, "_lvd_gn", None) is None:
                    taken[0] = False
                ops.groupnorm(y, samples, gamma, beta, 1e-5, silu=True, out=out)
                return y
            for i in range(3 if reps > 1 else 0):
                pair(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                pair(i)
            e1.record()
            torch.cuda.synchronize()
            line += "   %s pair %7.1f us%s" % (label, 1e3 * e0.elapsed_time(e1) / reps, "" if taken[0] else " (not taken)")
        print(line, flush=True)


if __name__ == "__main__":
    main()
