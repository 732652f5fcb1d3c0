"""Weight repacking from the reference `state_dict` layouts to the layouts the kernels read.
Pure data movement, done once at load time."""
import torch


def conv_weight_krsc(w):
    """Conv2d weight (Cout, Cin, kh, kw) -> [Cout, kh*kw*Cin] with K ordered (ky, kx, ci)."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous()


def tconv_weight(w):
    """Conv3d weight (Cout, Cin, 3, 1, 1) -> [Cout, 3*Cin] with K ordered (kt, ci)."""
    co, ci, kt = w.shape[:3]
    return w.reshape(co, ci, kt).permute(0, 2, 1).reshape(co, kt * ci).contiguous()


def geglu_interleave(w, b, half=128):
    """GEGLU proj weight (8C, C) = [value rows | gate rows] -> rows regrouped per `half` so that every 2*half-row
    block is [half value rows | the matching half gate rows]; the GEMM epilogue (block_n = 2*half) then multiplies
    value * gelu(gate) in-tile (models/attention.py:355-376)."""
    n2, k = w.shape
    n = n2 // 2
    assert n % half == 0
    wv, wg = w[:n].reshape(n // half, half, k), w[n:].reshape(n // half, half, k)
    wi = torch.cat([wv, wg], dim=1).reshape(n2, k).contiguous()
    bi = None
    if b is not None:
        bv, bg = b[:n].reshape(n // half, half), b[n:].reshape(n // half, half)
        bi = torch.cat([bv, bg], dim=1).reshape(n2).contiguous()
    return wi, bi


def s2_dgrad_weights(wk, cin):
    """Per input-pixel parity class (a, b) of a 3x3 / stride-2 / pad-1 conv: the taps that touch it, gathered into a
    contiguous [Cout, ntaps*Cin] weight, and for each tap the shift (sx, sy) from the input half-coordinate (X, Y) to
    the output pixel that used it:  iy = 2*oy + ky - 1  =>  a=0: ky=1, oy=Y;  a=1: ky=0 -> oy=Y+1, ky=2 -> oy=Y."""
    sel = {0: [(1, 0)], 1: [(0, 1), (2, 0)]}  # parity -> [(k, shift)]
    out = {}
    for a in (0, 1):
        for b in (0, 1):
            cols, shifts = [], []
            for ky, sy in sel[a]:
                for kx, sx in sel[b]:
                    t = ky * 3 + kx
                    cols.append(wk[:, t * cin:(t + 1) * cin])
                    shifts.append((sx, sy))
            out[(a, b)] = (torch.cat(cols, dim=1).contiguous(), shifts)
    return out


def pad_out_channels(wk, bias, n=64):
    """[Cout, K] weight and [Cout] bias -> zero-padded to n output channels (one GEMM tile wide)."""
    co, k = wk.shape
    wp = torch.zeros((n, k), dtype=wk.dtype, device=wk.device)
    wp[:co] = wk
    bp = torch.zeros((n,), dtype=bias.dtype, device=bias.device)
    bp[:co] = bias
    return wp, bp


def conv_in_weight_padded(w):
    """conv_in weight (Cout, Cin <= 7, 3, 3) -> [Cout, 64]: OIHW flattened (column ci*9 + ky*3 + kx), zero-padded to the
    64-deep contraction block of lvd_gemm (ops.conv_in_tc)."""
    co = w.shape[0]
    flat = w.reshape(co, -1)
    wp = torch.zeros((co, 64), dtype=w.dtype, device=w.device)
    wp[:, :flat.shape[1]] = flat
    return wp
