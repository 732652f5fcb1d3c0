import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def rel_err(a, b):
    """max |a-b| / max |b| in fp32 -- the normalised error every parity test reports."""
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def err_stats(a, b):
    """Error of `a` against `b`, every figure relative to the SIGNAL scale of b (element-wise |a-b|/|b| is undefined
    where b crosses zero): max_norm = max|a-b| / max|b|; rms_rel = ||a-b|| / ||b||; median / p999 / max of
    |a_i-b_i| / rms(b)."""
    import torch
    a = a.float().flatten()
    b = b.float().flatten()
    d = (a - b).abs()
    rms = b.pow(2).mean().sqrt().clamp_min(1e-30)
    if d.numel() > 8_000_000:  # torch.quantile input limit: a fixed stride sample keeps the estimate deterministic
        dq = d[:: d.numel() // 8_000_000 + 1]
    else:
        dq = d
    q = torch.quantile(dq, torch.tensor([0.5, 0.999], device=d.device))
    return dict(max_norm=(d.max() / b.abs().max().clamp_min(1e-30)).item(), rms_rel=(d.norm() / b.norm()).item(),
                median=(q[0] / rms).item(), p999=(q[1] / rms).item(), max=(d.max() / rms).item())


def fmt_stats(s):
    return ("max|d|/max|ref| %.2e  ||d||/||ref|| %.2e  |d|/rms(ref): median %.2e p99.9 %.2e max %.2e" %
            (s["max_norm"], s["rms_rel"], s["median"], s["p999"], s["max"]))
