"""Thin Python launch helpers over the C-ABI (one function per kernel family).

Activations are 2-D fp16 CUDA tensors `[rows, C]` (channels-last; rows = ((b*F+f)*H + y)*W + x) that may be
column slices of a wider buffer (stride(0) = ld >= C, stride(1) == 1).  torch is used for device memory and the
current stream only; all arithmetic happens in liblvd_b200.so.
"""
import ctypes
import functools
import os

import torch

from . import _lib
from ._lib import GemmDesc

ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_GELU = 0, 1, 2, 3
_NUM_SMS = 148  # B200

# Optional per-launch instrumentation (bench.py's roofline / kernel-family numbers): when PROF is a list, every launch
# appends [kernel family, algorithmic flops, algorithmic bytes, start event, end event, tag, scope] recorded on the
# launching stream.  Eager launches get ordinary events; launches issued while the stream is being CAPTURED get
# `external=True` events, which become event-record nodes of the CUDA graph -- after every replay of that graph the pair
# reads the duration of that kernel node in replay conditions (bench.py's instrumented-replay region).  PROF_SCOPE names
# the graph being captured so the reader knows which replays refresh which pairs.
PROF = None
PROF_SCOPE = None
FORCE_BLOCK_N = None  # tools/perf_gemm_sweep.py only: override the tile width of every lvd_gemm launch
FORCE_CTA_PAIR = 0    # 0 auto / 1 single-CTA tiles / 2 CTA pairs (tests and tools/perf_gemm_sweep.py)
GEGLU_HALF = 128      # value/gate interleave granularity of GEGLU projections (repack.geglu_interleave)


class _Timed:
    def __init__(self, family, flops=0.0, nbytes=0.0, tag=None):
        self.rec = None
        if PROF is not None:
            ext = torch.cuda.is_current_stream_capturing()
            self.rec = [family, float(flops), float(nbytes), torch.cuda.Event(enable_timing=True, external=ext),
                        torch.cuda.Event(enable_timing=True, external=ext), tag, PROF_SCOPE if ext else None]

    def __enter__(self):
        if self.rec is not None:
            self.rec[3].record()
        return self

    def __exit__(self, *a):
        if self.rec is not None:
            self.rec[4].record()
            PROF.append(self.rec)
        return False


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _check_act(t, name="tensor"):
    if not (t.is_cuda and t.dtype == torch.float16 and t.dim() == 2 and t.stride(1) == 1):
        raise ValueError(f"{name}: expected a 2-D fp16 CUDA tensor with unit column stride, got "
                         f"{t.dtype} {tuple(t.shape)} strides {t.stride()} on {t.device}")
    if t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
        raise ValueError(f"{name}: row stride must be a multiple of 8 and base 16-byte aligned")


FORCE_WEIGHT_RESIDENT = None  # tests / tools: 0 / 1 overrides weight_resident_policy
# measured: no gain on any shape of the step, hence off unless LVD_WEIGHT_RESIDENT=1
WEIGHT_RESIDENT_ON = os.environ.get("LVD_WEIGHT_RESIDENT", "0") != "0"
WEIGHT_RESIDENT_MAX_K = int(os.environ.get("LVD_WEIGHT_RESIDENT_MAX_K", "640"))
BIAS_F32 = True  # hand lvd_gemm an fp32 copy of the bias (tests switch it off to compare with the fp16 read)


def weight_resident_policy(n_taps, k, n, mn_major):
    """Whether an lvd_gemm launch asks for the weight-resident walk (include/lvd_b200.h): the short-K linears, whose
    operand delivery from L2 is the bound.  LVD_WEIGHT_RESIDENT=0 switches it off (A/B)."""
    return WEIGHT_RESIDENT_ON and n_taps == 1 and not mn_major and k <= WEIGHT_RESIDENT_MAX_K


def pick_block_n(n, act=ACT_NONE, mn_major=False, m_tiles=None, k_contract=None):
    """Tile width of an lvd_gemm launch.  Wide tiles move fewer operand bytes per FLOP, but the kernel is
    persistent -- one CTA (or CTA pair) per SM walks the tile list -- so what counts is the number of tile ROUNDS the
    slowest SM runs.  Cost model: rounds x (block_n + 78), the per-tile time fitted to the measured 128- / 160- /
    256-wide rates of the current kernel (tools/perf_gemm_blockn.py: M = 8640, N = 1280 conv3x3 190 / 175 / 184 us,
    model 1.08 : 1 : 1.05; the constant was 180 before the role loops moved to uniform registers, which made the
    narrower tiles relatively cheaper)."""
    if act == ACT_GEGLU:
        return 2 * GEGLU_HALF
    cands = [bn for bn in ((256, 128, 64) if mn_major else (256, 160, 128, 64)) if n % bn == 0]
    if not cands:
        raise ValueError(f"N={n} is not a multiple of 64")
    if m_tiles is None or len(cands) == 1:
        return cands[0]
    pair = m_tiles >= 2 and (k_contract or 0) >= 640  # mirrors the CTA-pair choice of lvd_gemm (gemm.cu)
    units = _NUM_SMS // 2 if pair else _NUM_SMS
    rows = (m_tiles + 1) // 2 if pair else m_tiles
    best = None
    for bn in cands:
        if bn == 64 and len(cands) > 1 and n > 64:
            continue  # 64-wide tiles never win on these shapes (operand traffic)
        rounds = -(-(rows * (n // bn)) // units)
        cost = rounds * (bn + 78)
        if best is None or cost < best[0]:
            best = (cost, bn)
    return best[1]


@functools.lru_cache(maxsize=None)
def choose_box(dims, max_rows=128, div=None):
    """Pick per-axis box extents (product <= 128) for output extents `dims` (fastest axis first) that maximise
    the fraction of useful rows over all tiles.  `div[i]` (optional) constrains the box extent of axis i to divisors of
    div[i] -- used to keep a tile inside one GroupNorm sample (ops.gemm_raw gn=...)."""
    dims = tuple(int(d) for d in dims)
    best = None

    def rec(i, cur, rows):
        nonlocal best
        if i == len(dims):
            tiles = 1
            for d, b in zip(dims, cur):
                tiles *= -(-d // b)
            total = 1
            for d in dims:
                total *= d
            eff = total / (tiles * 128.0)
            key = (round(eff, 6), rows, cur[0])
            if best is None or key > best[0]:
                best = (key, tuple(cur))
            return
        d = dims[i]
        for b in range(1, min(d, max_rows // rows, 256) + 1):
            if div is not None and div[i] is not None and div[i] % b:
                continue
            # only divisors, or the full extent, or power-of-two chunks: keeps the search small
            if d % b == 0 or b == d or (b & (b - 1)) == 0:
                rec(i + 1, cur + [b], rows * b)

    rec(0, [], 1)
    return best[1]


def _box_efficiency(dims, box):
    tiles, total = 1, 1
    for d, b in zip(dims, box):
        tiles *= -(-int(d) // int(b))
        total *= int(d)
    return total / (tiles * 128.0)


def _box_for_gn(dims, gn, sample_axis, per_sample):
    """Box for a conv whose output feeds a GroupNorm with `gn = (samples, groups)`: the unconstrained choice, unless a box
    whose extent along `sample_axis` divides `per_sample` (so that tiles stay inside one sample) is equally efficient."""
    free = choose_box(tuple(dims))
    if gn is None or not GN_EPILOGUE:
        return free, None
    div = tuple(per_sample if i == sample_axis else None for i in range(len(dims)))
    tied = choose_box(tuple(dims), 128, div)
    if _box_efficiency(dims, tied) + 1e-9 >= _box_efficiency(dims, free):
        return tied, gn
    return free, None


# ---- GroupNorm statistics from the producing GEMM's epilogue (lvd_gemm gn_partial ->
# lvd_groupnorm_apply): a conv whose output feeds a GroupNorm is launched with `gn=(samples, groups)`; its epilogue
# leaves per-warp partial sums in a zero-state buffer and the output tensor carries a handle; `groupnorm` on that
# tensor then skips its statistics pass.  LVD_GN_EPILOGUE=0 switches the path off (A/B).
import os as _os

GN_EPILOGUE = _os.environ.get("LVD_GN_EPILOGUE", "1") != "0"
_gn_part = {}
_gn_part_retired = []   # outgrown buffers stay alive: captured CUDA graphs hold their addresses
_gn_pending = {}        # device -> handle whose partials have not been consumed yet


class _GnHandle:
    __slots__ = ("ptr", "shape", "samples", "groups", "slots", "rows_per_sample", "buf")


def _gn_part_buffer(dev, nfloats):
    buf = _gn_part.get(dev)
    if buf is None or buf.numel() < nfloats:
        if buf is not None:
            _gn_part_retired.append(buf)
        buf = torch.zeros(max(nfloats, 1 << 21), dtype=torch.float32, device=dev)  # zero is the buffer's rest state
        _gn_part[dev] = buf
    return buf


def gn_reset(dev):
    """Called at the start of a forward / guidance pass: if an earlier pass died between a producing GEMM and its
    GroupNorm, bring the partial-statistics buffer back to its zero rest state."""
    dev = torch.device(dev)
    if _gn_pending.get(dev) is not None:
        _gn_pending[dev] = None
        if dev in _gn_part:
            _gn_part[dev].zero_()


# ---- LayerNorm folded into the linear that consumes it (lvd_gemm ln_stats / lnstat_partial, lvd_ln_rowstats):
#   LN(x) @ W.T + b  =  rstd * (x @ (gamma o W).T - mean * colsum) + (W @ beta + b)
# the producer of x (a linear with ln_emit=True) leaves per-row {sum, sum of squares} per 32-column slab, ln_rowstats
# folds them into {mean, rstd}, and the consumer runs on the raw x: the LayerNorm pass over the tensor disappears.
# LVD_LN_FOLD=0 switches the path off (A/B).
LN_FOLD = _os.environ.get("LVD_LN_FOLD", "1") != "0"
_ln_part = {}
_ln_stats = {}
_ln_gen = {}  # per (device, buffer): emissions so far -- a handle is valid only until the next producer reuses the buffer


class LnFolded:
    """w [N, K] fp16 = gamma o W;  cs [N] fp32 = row sums of the ROUNDED w;  lb [N] fp32 = W @ beta + bias."""
    __slots__ = ("w", "cs", "lb")


def ln_fold(w, bias, gamma, beta):
    f = LnFolded()
    wf = w.float()
    f.w = (wf * gamma.float()[None, :]).half().contiguous()
    f.cs = f.w.float().sum(1).contiguous()
    f.lb = (wf @ beta.float()).contiguous()
    if bias is not None:
        f.lb = (f.lb + bias.float()).contiguous()
    return f


def _ln_part_buffer(dev, floats2):
    dev = torch.device(dev)
    key = (dev, int(floats2))
    buf = _ln_part.get(key)
    if buf is None:
        buf = torch.empty((int(floats2), 2), dtype=torch.float32, device=dev)
        _ln_part[key] = buf
    return buf


def ln_rowstats(x, eps=1e-5):
    """{mean, rstd} per row of x [rows, C] from the partial statistics its producer left (linear(..., ln_emit=True));
    None when x carries none (the caller then runs the LayerNorm kernel)."""
    h = getattr(x, "_lvd_ln", None)
    if h is None or not LN_FOLD:
        return None
    part, slabs, rows, ptr, gen = h
    if ptr != x.data_ptr() or x.shape[0] != rows or x.shape[1] != slabs * 32 or _ln_gen.get(part.data_ptr()) != gen:
        return None
    key = (x.device, rows)
    stats = _ln_stats.get(key)
    if stats is None:
        stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device)
        _ln_stats[key] = stats
    with _Timed("layernorm", 0.0, 8.0 * rows * (slabs + 1), "ln_rowstats r%d c%d" % (rows, slabs * 32)):
        _lib.check(_lib.lib().lvd_ln_rowstats(part.data_ptr(), slabs, rows, slabs * 32, float(eps), stats.data_ptr(),
                                              _stream()), "ln_rowstats")
    return stats


def _gn_plan(out_dims, a_box, out_row_stride, samples, groups, n):
    """(tiles_per_sample, subslots, rows_per_sample) when every row tile of the launch lies inside one GroupNorm sample
    and the tiles of a sample are consecutive in tile order; None otherwise (the consumer then runs its own statistics
    pass)."""
    dims = [int(v) for v in out_dims]
    box = [int(a_box[i + 1]) for i in range(4)]
    rows = dims[0] * dims[1] * dims[2] * dims[3]
    if rows % samples or n % groups:
        return None
    rps = rows // samples
    prod = 1
    for i in range(4):  # output rows must be numbered in tile-axis order
        if dims[i] > 1 and int(out_row_stride[i]) != prod:
            return None
        prod *= dims[i]
    prod, tps, i = 1, 1, 0
    while i < 4 and prod * dims[i] <= rps:  # axes that lie wholly inside a sample
        prod *= dims[i]
        tps *= -(-dims[i] // box[i])
        i += 1
    if prod != rps:  # the sample ends inside axis i: it must cover whole boxes of that axis
        if i >= 4 or rps % prod:
            return None
        cnt = rps // prod
        if dims[i] % cnt or cnt % box[i]:
            return None
        tps *= cnt // box[i]
        i += 1
    if any(box[k] != 1 for k in range(i, 4) if dims[k] > 1):  # axes that index samples: one sample per tile
        return None
    gw = n // groups
    return tps, (gw + 30) // 32 + 1, rps


def _bias_f32(bias):
    """fp32 copy of a bias vector for the specialised epilogues of lvd_gemm.  The copy hangs on the bias tensor OBJECT
    (weights are persistent objects: converted once), so it lives exactly as long as the bias does -- a captured CUDA
    graph that has the copy's address baked in belongs to the same engine as the weight and dies with it.  (A global
    cache with eviction would free copies that captured graphs still read.)"""
    try:
        version = bias._version
    except RuntimeError:  # inference tensors carry no version counter (and cannot be updated in place outside inference mode)
        version = -1
    hit = getattr(bias, "_lvd_f32", None)
    if hit is None or hit[1] != version or hit[2] != bias.data_ptr():
        hit = (bias.float().contiguous(), version, bias.data_ptr())
        bias._lvd_f32 = hit
    return hit[0].data_ptr()


def gemm_raw(*, a, a_dims, a_strides, a_box, taps, k_per_tap, w, n, out, out_dims, out_row_stride,
             rv_stride=(0, 0, 0, 0), bias=None, act=ACT_NONE, rowvec=None, gamma=1.0, residual=None,
             b_mn_major=False, block_n=None, ldb=None, b_k_tap_stride=0, b_n_tap_stride=0, gn=None, ln=None,
             ln_emit=False):
    d = GemmDesc()
    d.a = a.data_ptr()
    for i in range(5):
        d.a_dims[i] = int(a_dims[i])
        d.a_strides[i] = int(a_strides[i])
        d.a_box[i] = int(a_box[i])
    d.n_taps = len(taps)
    d.k_per_tap = int(k_per_tap)
    for t, off in enumerate(taps):
        for i in range(5):
            d.tap_off[t][i] = int(off[i])
    d.b = w.data_ptr()
    d.b_mn_major = 1 if b_mn_major else 0
    if block_n is None:
        m_tiles = 1
        for i in range(4):
            m_tiles *= -(-int(out_dims[i]) // max(1, int(a_box[i + 1])))
        block_n = FORCE_BLOCK_N or pick_block_n(n, act, b_mn_major, m_tiles, len(taps) * int(k_per_tap))
    d.block_n = int(block_n)
    d.ldb = int(ldb if ldb is not None else w.stride(0))
    d.n = int(n)
    d.act = int(act)
    for i in range(4):
        d.out_dims[i] = int(out_dims[i])
        d.rv_stride[i] = int(rv_stride[i])
        d.out_row_stride[i] = int(out_row_stride[i])
    d.bias = bias.data_ptr() if bias is not None else None
    d.bias_f32 = _bias_f32(bias) if BIAS_F32 and bias is not None and rowvec is None and gamma == 1.0 and \
        act in (ACT_NONE, ACT_GEGLU) else None
    d.rowvec = rowvec.data_ptr() if rowvec is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.out = out.data_ptr()
    d.ld_rowvec = rowvec.stride(0) if rowvec is not None else 0
    d.ld_res = residual.stride(0) if residual is not None else 0
    d.ld_out = out.stride(0)
    d.gamma = float(gamma)
    d.b_k_tap_stride = int(b_k_tap_stride)
    d.b_n_tap_stride = int(b_n_tap_stride)
    d.cta_pair = int(FORCE_CTA_PAIR)
    if FORCE_WEIGHT_RESIDENT is not None:
        d.weight_resident = int(FORCE_WEIGHT_RESIDENT)
    else:
        d.weight_resident = 1 if weight_resident_policy(len(taps), int(k_per_tap), int(n), b_mn_major) else 0
    if gn is not None and GN_EPILOGUE and act != ACT_GEGLU and len(taps) * int(k_per_tap) >= 960:
        # (the 5-k-block linears are bound by their epilogue: the extra column-sum pass would cost them more than the
        # statistics pass it saves)
        samples, groups = gn
        plan = _gn_plan(out_dims, a_box, out_row_stride, samples, groups, n)
        if plan is not None:
            tps, sub, rps = plan
            slots = tps * 4 * sub
            dev = out.device
            if _gn_pending.get(dev) is not None:
                raise RuntimeError("lvd_b200: GroupNorm partial statistics of a previous producer were never consumed")
            buf = _gn_part_buffer(dev, samples * slots * groups * 2)
            d.gn_partial = buf.data_ptr()
            d.gn_group_width, d.gn_groups = n // groups, int(groups)
            d.gn_tiles_per_sample, d.gn_subslots = int(tps), int(sub)
            h = _GnHandle()
            h.ptr, h.shape, h.samples, h.groups = out.data_ptr(), tuple(out.shape), int(samples), int(groups)
            h.slots, h.rows_per_sample, h.buf = slots, rps, buf
            out._lvd_gn = h
            _gn_pending[dev] = h
    if ln is not None:  # LayerNorm folded into this linear: (stats [rows] {mean, rstd}, colsum, bias, A-row strides)
        stats, cs, lb, strides = ln
        assert bias is None and residual is None and rowvec is None and gamma == 1.0 and len(taps) == 1
        assert cs.dtype == torch.float32 and lb.dtype == torch.float32 and cs.numel() == n and lb.numel() == n
        d.ln_stats, d.ln_colsum, d.ln_bias = stats.data_ptr(), cs.data_ptr(), lb.data_ptr()
        for i in range(4):
            d.ln_row_stride[i] = int(strides[i])
    if ln_emit and LN_FOLD and act != ACT_GEGLU:
        rows_total = 1
        for v in out_dims:
            rows_total *= int(v)
        part = _ln_part_buffer(out.device, (n // 32) * rows_total)
        d.lnstat_partial, d.lnstat_rows = part.data_ptr(), rows_total
        for i in range(4):
            d.lnstat_row_stride[i] = int(out_row_stride[i])
        gen = _ln_gen.get(part.data_ptr(), 0) + 1
        _ln_gen[part.data_ptr()] = gen
        out._lvd_ln = (part, n // 32, rows_total, out.data_ptr(), gen)
    if PROF is not None:
        m = 1
        for v in out_dims:
            m *= int(v)
        kt = len(taps) * int(k_per_tap)
        n_out = n // 2 if act == ACT_GEGLU else n
        nbytes = 2.0 * (m * int(k_per_tap) + n * kt + m * n_out * (2 if residual is not None else 1))
        tag = "m%d n%d k%d taps%d act%d res%d mn%d bn%d" % (m, n, int(k_per_tap), len(taps), act,
                                                           residual is not None, bool(b_mn_major), block_n)
        with _Timed("gemm_tc_kernel", 2.0 * m * n * kt, nbytes, tag):
            _lib.check(_lib.lib().lvd_gemm(ctypes.byref(d), _stream()), "gemm")
        return out
    _lib.check(_lib.lib().lvd_gemm(ctypes.byref(d), _stream()), "gemm")
    return out


def _alloc_out(rows, n, like):
    return torch.empty((rows, n), dtype=torch.float16, device=like.device)


def linear(x, w, bias=None, *, act=ACT_NONE, residual=None, gamma=1.0, out=None, rowvec=None, rows_per_vec=0,
           b_mn_major=False, n=None, row_map=None, ln=None, ln_emit=False):
    """out = epilogue(x @ w.T).  x [rows, K]; w [N, K] (or [K, N] when b_mn_major).  GEGLU expects the
    value/gate-interleaved weight produced by `repack.geglu_interleave` and returns N/2 columns.
    ln_emit: the epilogue also leaves the row statistics of `out` for a LayerNorm folded into the next linear
    (`ln_rowstats(out)`); ln = (stats, LnFolded): x is the RAW LayerNorm input, w the folded weight (`ln_fold`)."""
    _check_act(x, "x")
    rows, k = x.shape
    if n is None:
        n = w.shape[1] if b_mn_major else w.shape[0]
    n_out = n // 2 if act == ACT_GEGLU else n
    if out is None:
        out = _alloc_out(rows, n_out, x)
    _check_act(out, "out")
    ld = x.stride(0)
    big = ld * max(rows, 1)
    rv = (0, 0, 0, 0)
    a_dims, a_strides, a_box = (k, rows, 1, 1, 1), (1, ld, big, big, big), (64, 128, 1, 1, 1)
    out_dims, ors = (rows, 1, 1, 1), (1, 0, 0, 0)
    if row_map is not None:
        # Row permutation folded into the launch: the A tile is a (pixels x frames) box of the input order and the
        # epilogue writes / reads residual rows in the output order -- this is how the temporal transformer's
        # permutes (transformer_temporal.py:142-156,176-182) cost nothing.
        kind, batch, frames, hw = row_map
        assert rows == batch * frames * hw and rowvec is None
        if kind == "fm2sm":      # input rows ((b*F+f)*hw+p)  ->  output rows ((b*hw+p)*F+f)
            bp, bf, bb = choose_box((hw, frames, batch))
            a_dims, a_strides = (k, hw, frames, batch, 1), (1, ld, ld * hw, ld * hw * frames, big)
            a_box = (64, bp, bf, bb, 1)
            out_dims, ors = (hw, frames, batch, 1), (frames, 1, hw * frames, 0)
        elif kind == "sm2fm":    # input rows ((b*hw+p)*F+f)  ->  output rows ((b*F+f)*hw+p)
            bf, bp, bb = choose_box((frames, hw, batch))
            a_dims, a_strides = (k, frames, hw, batch, 1), (1, ld, ld * frames, ld * hw * frames, big)
            a_box = (64, bf, bp, bb, 1)
            out_dims, ors = (frames, hw, batch, 1), (hw, 1, hw * frames, 0)
        else:
            raise ValueError(kind)
    if rowvec is not None:
        # rows are grouped in blocks of rows_per_vec consecutive rows sharing one row-vector
        assert rows_per_vec > 0 and rows % rows_per_vec == 0
        bx = choose_box((rows_per_vec, rows // rows_per_vec))
        a_dims = (k, rows_per_vec, rows // rows_per_vec, 1, 1)
        a_strides = (1, ld, ld * rows_per_vec, big, big)
        a_box = (64, bx[0], bx[1], 1, 1)
        out_dims, ors, rv = (rows_per_vec, rows // rows_per_vec, 1, 1), (1, rows_per_vec, 0, 0), (0, 1, 0, 0)
    if ln is not None:
        stats, folded = ln
        assert w is folded.w and stats.numel() == 2 * rows
        ln = (stats, folded.cs, folded.lb, [a_strides[i + 1] // ld if a_dims[i + 1] > 1 else 0 for i in range(4)])
    return gemm_raw(a=x, a_dims=a_dims, a_strides=a_strides, a_box=a_box, taps=[(0, 0, 0, 0, 0)], k_per_tap=k, w=w,
                    n=n, out=out, out_dims=out_dims, out_row_stride=ors, rv_stride=rv, bias=bias, act=act,
                    rowvec=rowvec, gamma=gamma, residual=residual, b_mn_major=b_mn_major, ln=ln, ln_emit=ln_emit)


TAPS_3X3 = [(0, dx, dy, 0, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


def conv2d_3x3(x, n_img, h, w_, wk, bias=None, *, rowvec=None, residual=None, out=None, gn=None):
    """3x3 / stride 1 / pad 1 convolution as 9 shifted TMA boxes.  x [n_img*h*w_, Cin]; wk [Cout, 9*Cin] with K
    ordered (ky, kx, ci).  rowvec [n_img, Cout] is added per image (time-embedding projection)."""
    _check_act(x, "x")
    cin = x.shape[1]
    cout = wk.shape[0]
    assert wk.shape[1] == 9 * cin and x.shape[0] == n_img * h * w_
    if out is None:
        out = _alloc_out(x.shape[0], cout, x)
    ld = x.stride(0)
    # gn = (samples, groups): images per GroupNorm sample = n_img / samples (1: per-frame norm, F: the 5-D norms)
    (bw, bh, bn), gn = _box_for_gn((w_, h, n_img), gn, 2, n_img // gn[0] if gn else 1)
    return gemm_raw(a=x, a_dims=(cin, w_, h, n_img, 1), a_strides=(1, ld, ld * w_, ld * w_ * h, ld * w_ * h * n_img),
                    a_box=(64, bw, bh, bn, 1), taps=TAPS_3X3, k_per_tap=cin, w=wk, n=cout, out=out,
                    out_dims=(w_, h, n_img, 1), out_row_stride=(1, w_, w_ * h, 0), rv_stride=(0, 0, 1, 0), bias=bias,
                    rowvec=rowvec, residual=residual, gn=gn)


def conv2d_3x3_s2(x, n_img, h, w_, wk, bias=None, *, out=None):
    """3x3 / stride 2 / pad 1 (diffusers Downsample2D).  The input is viewed as (2C, W/2, 2, H/2, N) so that the
    column/row parity of each tap is a coordinate; no space-to-depth copy is made."""
    _check_act(x, "x")
    cin = x.shape[1]
    cout = wk.shape[0]
    assert h % 2 == 0 and w_ % 2 == 0 and x.stride(0) == cin, "stride-2 conv needs a dense even-sized input"
    ho, wo = h // 2, w_ // 2
    if out is None:
        out = _alloc_out(n_img * ho * wo, cout, x)
    taps = []
    for ky in range(3):
        pa, sy = ((1, -1), (0, 0), (1, 0))[ky]  # input row 2*oy + ky - 1 = 2*(oy+sy) + pa
        for kx in range(3):
            pb, sx = ((1, -1), (0, 0), (1, 0))[kx]
            taps.append((pb * cin, sx, pa, sy, 0))
    bw, bh, bn = choose_box((wo, ho, n_img))
    return gemm_raw(a=x, a_dims=(2 * cin, wo, 2, ho, n_img),
                    a_strides=(1, 2 * cin, w_ * cin, 2 * w_ * cin, h * w_ * cin), a_box=(64, bw, 1, bh, bn), taps=taps,
                    k_per_tap=cin, w=wk, n=cout, out=out, out_dims=(wo, 1, ho, n_img),
                    out_row_stride=(1, 0, wo, wo * ho), bias=bias)


TAPS_T = [(0, 0, df, 0, 0) for df in (-1, 0, 1)]


def conv_temporal(x, batch, frames, hw, wk, bias=None, *, residual=None, out=None, gn=None):
    """Conv3d (3,1,1) pad (1,0,0) over the frame axis.  x [(batch*frames)*hw, C]; wk [Cout, 3*Cin], K = (kt, ci)."""
    _check_act(x, "x")
    cin = x.shape[1]
    cout = wk.shape[0]
    assert x.shape[0] == batch * frames * hw
    if out is None:
        out = _alloc_out(x.shape[0], cout, x)
    ld = x.stride(0)
    (bp, bf, bb), gn = _box_for_gn((hw, frames, batch), gn, 2, batch // gn[0] if gn else 1)
    return gemm_raw(a=x, a_dims=(cin, hw, frames, batch, 1),
                    a_strides=(1, ld, ld * hw, ld * hw * frames, ld * hw * frames * batch), a_box=(64, bp, bf, bb, 1),
                    taps=TAPS_T, k_per_tap=cin, w=wk, n=cout, out=out, out_dims=(hw, frames, batch, 1),
                    out_row_stride=(1, hw, hw * frames, 0), bias=bias, residual=residual, gn=gn)


# ----------------------------------------------------------------------------------------------- normalisation
_gn_ws = {}
_gn_ws_retired = []  # outgrown workspaces stay alive: captured CUDA graphs hold their raw pointers


def _gn_workspace(dev, samples, groups):
    """Scratch for the GroupNorm partial statistics.  64 MB up front covers 240 samples (a CFG batch of 5 videos x 24
    frames); a larger request allocates a new buffer and RETIRES the old one without freeing it, because UNet /
    guidance graphs captured earlier replay with the old address baked in."""
    need = samples * 1040 * groups * 8
    ws = _gn_ws.get(dev)
    if ws is None or ws.numel() < need:
        if ws is not None:
            _gn_ws_retired.append(ws)
        ws = torch.empty(max(need, 1 << 26), dtype=torch.uint8, device=dev)
        _gn_ws[dev] = ws
    return ws


def groupnorm(x, samples, gamma, beta, eps, *, groups=32, silu=False, out=None, stats_out=None):
    """GroupNorm (+ optional SiLU) with statistics per (sample, group); x [samples*rows_per_sample, C].  When x was
    written by a GEMM launched with `gn=(samples, groups)` its statistics come from that GEMM's epilogue and only the
    normalisation pass runs."""
    _check_act(x, "x")
    rows, c = x.shape
    assert rows % samples == 0
    if out is None:
        out = _alloc_out(rows, c, x)
    h = getattr(x, "_lvd_gn", None)
    pend = _gn_pending.get(x.device)
    if pend is not None:
        if h is not pend or h.ptr != x.data_ptr() or h.shape != tuple(x.shape) or h.samples != samples or \
                h.groups != groups:
            raise RuntimeError("lvd_b200: a GEMM left GroupNorm partial statistics for another tensor / another "
                               "GroupNorm geometry than the one being normalised")
        _gn_pending[x.device] = None
        stats = stats_out if stats_out is not None else gn_stats_buffer(samples, x.device, groups)
        with _Timed("groupnorm", 0.0, 2.0 * rows * c * 2, "gn(epilogue stats) s%d r%d c%d" % (samples, rows // samples, c)):
            ws = _gn_workspace(x.device, samples, groups)
            _lib.check(_lib.lib().lvd_groupnorm_apply(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), samples,
                                                      rows // samples, c, groups, gamma.data_ptr(), beta.data_ptr(),
                                                      float(eps), 1 if silu else 0, h.buf.data_ptr(), h.slots,
                                                      ws.data_ptr(), stats.data_ptr(), _stream()), "groupnorm_apply")
        return out
    ws = _gn_workspace(x.device, samples, groups)
    with _Timed("groupnorm", 0.0, 2.0 * rows * c * 2, "gn s%d r%d c%d" % (samples, rows // samples, c)):
        _lib.check(_lib.lib().lvd_groupnorm(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), samples,
                                            rows // samples, c, groups, gamma.data_ptr(), beta.data_ptr(), float(eps),
                                            1 if silu else 0, stats_out.data_ptr() if stats_out is not None else None,
                                            ws.data_ptr(), ws.numel(), _stream()), "groupnorm")
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _check_act(x, "x")
    rows, c = x.shape
    if out is None:
        out = _alloc_out(rows, c, x)
    with _Timed("layernorm", 0.0, 2.0 * rows * c * 2, "ln r%d c%d" % (rows, c)):
        _lib.check(_lib.lib().lvd_layernorm(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, c,
                                            gamma.data_ptr(), beta.data_ptr(), float(eps), _stream()), "layernorm")
    return out


# ----------------------------------------------------------------------------------------------- attention
def attention(q, kv, *, heads, nz, s_q=0, s_k=0, q_col0=0, k_col0=0, v_col0=0, kv_z_div=1, temporal=False, frames=0,
              hw=0, scale=0.125, out=None, o_col0=0, p_full=None, p_tok=None, tok_cols=(), packed=False, lse=None,
              causal=False, tok_cols_dev=None):
    """softmax(Q K^T * scale) V with 64-wide heads.  q / kv are 2-D row buffers; q/k/v live at column offsets."""
    _check_act(q, "q")
    _check_act(kv, "kv")
    d = _lib.AttnDesc()
    rows = q.shape[0]
    if out is None:
        out = _alloc_out(rows, heads * 64, q)
    d.q, d.kv, d.out = q.data_ptr(), kv.data_ptr(), out.data_ptr()
    d.p_full = p_full.data_ptr() if p_full is not None else None
    d.p_tok = p_tok.data_ptr() if p_tok is not None else None
    d.ld_q, d.ld_kv, d.ld_o = q.stride(0), kv.stride(0), out.stride(0)
    d.q_width, d.kv_width = q.shape[1], kv.shape[1]
    d.s_q, d.s_k, d.hw, d.frames = int(s_q), int(s_k), int(hw), int(frames)
    d.mode = (2 if packed else 1) if temporal else 0
    d.lse = lse.data_ptr() if lse is not None else None
    d.q_col0, d.k_col0, d.v_col0, d.o_col0 = int(q_col0), int(k_col0), int(v_col0), int(o_col0)
    d.heads, d.nz, d.kv_z_div = int(heads), int(nz), int(kv_z_div)
    d.n_tok = len(tok_cols)
    for i, t in enumerate(tok_cols):
        d.tok_cols[i] = int(t)
    if tok_cols_dev is not None:  # int32 device list read at run time (p_tok's leading extent gives the count)
        assert tok_cols_dev.dtype == torch.int32 and tok_cols_dev.is_cuda and p_tok is not None
        d.n_tok = int(tok_cols_dev.numel())
        d.tok_cols_dev = tok_cols_dev.data_ptr()
    d.scale = float(scale)
    d.causal = 1 if causal else 0
    if temporal:
        fl, fam = 4.0 * nz * hw * heads * frames * frames * 64, "attention_temporal"
    else:
        fl, fam = 4.0 * nz * heads * s_q * s_k * 64, ("attention_self" if kv is q else "attention_cross")
    with _Timed(fam, fl, 2.0 * (q.shape[0] * heads * 64 * 2 + kv.shape[0] * heads * 128),
                "%s nz%d h%d sq%d sk%d f%d" % (fam, nz, heads, s_q, s_k, frames)):
        _lib.check(_lib.lib().lvd_attention(ctypes.byref(d), _stream()), "attention")
    return out


# ----------------------------------------------------------------------------------------------- thin ends / eltwise
def conv_in(sample, weight, bias, out=None):
    """sample (B, Cin, F, H, W) fp16 contiguous -> rows [B*F*H*W, Cout]; weight OIHW."""
    b, cin, f, h, w_ = sample.shape
    assert sample.is_contiguous() and sample.dtype == torch.float16
    cout = weight.shape[0]
    if out is None:
        out = torch.empty((b * f * h * w_, cout), dtype=torch.float16, device=sample.device)
    with _Timed("eltwise", 0.0, 2.0 * (sample.numel() + out.shape[0] * cout), "conv_in"):
        _lib.check(_lib.lib().lvd_conv_in(sample.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                          out.stride(0), b, cin, f, h, w_, cout, _stream()), "conv_in")
    return out


def conv_in_tc(sample, w_pad, bias, out=None):
    """conv_in on the tensor-core kernel: im2col of the <= 7-channel sample into 64 columns, then one K = 64 GEMM against
    `w_pad` [Cout, 64] (repack.conv_in_weight_padded).  Same result layout as conv_in."""
    b, cin, f, h, w_ = sample.shape
    assert sample.is_contiguous() and sample.dtype == torch.float16 and cin * 9 <= 64
    col = torch.empty((b * f * h * w_, 64), dtype=torch.float16, device=sample.device)
    with _Timed("eltwise", 0.0, 2.0 * (sample.numel() + col.numel()), "im2col_small"):
        _lib.check(_lib.lib().lvd_im2col_small(sample.data_ptr(), col.data_ptr(), b, cin, f, h, w_, _stream()),
                   "im2col_small")
    return linear(col, w_pad, bias, out=out)


def conv_out(x, wk, bias, batch, frames, h, w_):
    """rows [B*F*H*W, Cin] -> (B, Cout, F, H, W); wk [Cout, 9*Cin] (K = ky,kx,ci)."""
    _check_act(x, "x")
    cin = x.shape[1]
    cout = wk.shape[0]
    out = torch.empty((batch, cout, frames, h, w_), dtype=torch.float16, device=x.device)
    _lib.check(_lib.lib().lvd_conv_out(x.data_ptr(), x.stride(0), wk.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                       batch, cin, frames, h, w_, cout, _stream()), "conv_out")
    return out


def conv_out_tc(x, wk_pad, bias_pad, cout, batch, frames, h, w_):
    """conv_out on the tensor-core kernel: wk_pad [64, 9*Cin] / bias_pad [64] are the conv_out weight / bias with the
    output channels zero-padded to one 64-wide tile (repack.pad_out_channels); returns (B, cout, F, H, W)."""
    rows = conv2d_3x3(x, batch * frames, h, w_, wk_pad, bias_pad)
    out = torch.empty((batch, cout, frames, h, w_), dtype=torch.float16, device=x.device)
    with _Timed("eltwise", 0.0, 4.0 * out.numel(), "rows_to_ncfhw"):
        _lib.check(_lib.lib().lvd_rows_to_ncfhw(rows.data_ptr(), rows.stride(0), out.data_ptr(), batch, cout, frames,
                                                h * w_, _stream()), "rows_to_ncfhw")
    return out


def upsample2x(x, n_img, h, w_, out=None):
    _check_act(x, "x")
    c = x.shape[1]
    if out is None:
        out = _alloc_out(n_img * 4 * h * w_, c, x)
    with _Timed("eltwise", 0.0, 2.0 * 5 * x.shape[0] * c, "upsample2x"):
        _lib.check(_lib.lib().lvd_upsample2x(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), n_img, h, w_, c,
                                             _stream()), "upsample2x")
    return out


def timestep_embed(t_f32, dim):
    out = torch.empty((t_f32.numel(), dim), dtype=torch.float16, device=t_f32.device)
    _lib.check(_lib.lib().lvd_timestep_embed(t_f32.data_ptr(), out.data_ptr(), t_f32.numel(), dim, _stream()),
               "timestep_embed")
    return out


def silu(x):
    assert x.is_contiguous() and x.dtype == torch.float16
    y = torch.empty_like(x)
    _lib.check(_lib.lib().lvd_silu(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "silu")
    return y


def copy_rows(x, out):
    _check_act(x, "x")
    _check_act(out, "out")
    with _Timed("eltwise", 0.0, 4.0 * x.shape[0] * x.shape[1], "copy_rows"):
        _lib.check(_lib.lib().lvd_copy_rows(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), x.shape[0],
                                            x.shape[1], _stream()), "copy_rows")
    return out


# ----------------------------------------------------------------------------------------------- box energy
LOSS_MAX, LOSS_RATIO, LOSS_CE = 0, 1, 2


def box_energy(p, loss, boxes, slot_weight, *, frames, heads, H, W, slot_col=None, n_text=0, fg_top_p=0.2,
               bg_top_p=0.2, fg_weight=1.0, bg_weight=1.0, com_loss_scale=0.0, out_scale=1.0, dp=None,
               grad_scale=1.0, loss_kind=LOSS_MAX, k_min=1, eps=1.0e-2, attn_sync_weight=0.0, boxdiff_loss_scale=0.0,
               boxdiff_normed=True, boxdiff_L=1, p_frames=0, slot_frame0=None, slot_loss=None):
    """Accumulate the box energy of one attention map into `loss` (0-d fp32 CUDA tensor) and optionally write
    grad_scale * dE/dP into `dp`.  slot_col given -> `p` is the full map [F, heads, hw, n_text]; otherwise `p` holds
    token columns [n_slots, F, heads, hw]."""
    d = _lib.EnergyDesc()
    assert p.is_cuda and p.dtype == torch.float16 and p.is_contiguous()
    assert boxes.dtype == torch.int32 and boxes.is_contiguous() and slot_weight.dtype == torch.float32
    d.p = p.data_ptr()
    d.slot_col = slot_col.data_ptr() if slot_col is not None else None
    d.boxes = boxes.data_ptr()
    d.slot_weight = slot_weight.data_ptr()
    d.loss = loss.data_ptr()
    d.dp = dp.data_ptr() if dp is not None else None
    d.layout = 1 if slot_col is not None else 0
    d.n_text = int(n_text)
    d.n_slots = int(boxes.shape[0])
    d.frames, d.heads, d.H, d.W = int(frames), int(heads), int(H), int(W)
    d.fg_top_p, d.bg_top_p = float(fg_top_p), float(bg_top_p)
    d.fg_weight, d.bg_weight = float(fg_weight), float(bg_weight)
    d.com_loss_scale, d.out_scale, d.grad_scale = float(com_loss_scale), float(out_scale), float(grad_scale)
    d.loss_kind, d.k_min = int(loss_kind), int(k_min)
    d.boxdiff_normed, d.boxdiff_L = 1 if boxdiff_normed else 0, int(boxdiff_L)
    d.eps, d.attn_sync_weight, d.boxdiff_loss_scale = float(eps), float(attn_sync_weight), float(boxdiff_loss_scale)
    d.p_frames = int(p_frames)
    d.slot_frame0 = slot_frame0.data_ptr() if slot_frame0 is not None else None
    d.slot_loss = slot_loss.data_ptr() if slot_loss is not None else None
    with _Timed("box_energy", 0.0, 2.0 * p.numel()):
        _lib.check(_lib.lib().lvd_box_energy(ctypes.byref(d), _stream()), "box_energy")
    return loss


# ----------------------------------------------------------------------------------------------- energy pre-processing
def gaussian_weights(kernel_size=3, sigma=0.5):
    """The 3x3 depthwise kernel of utils/attn.py:112-134 (GaussianSmoothing.__init__; the exponent is
    -((x - mean) / (2 sigma))^2 there, not the textbook -(x - mean)^2 / (2 sigma^2)), row-major, as Python floats."""
    import math
    if kernel_size != 3:
        # the reference pads by exactly 1 and asserts the shape is unchanged (utils/guidance.py:213-218)
        raise AssertionError("smooth_attn: only kernel_size=3 keeps the map shape (utils/guidance.py:218)")
    mean = (kernel_size - 1) / 2
    g1 = [1 / (sigma * math.sqrt(2 * math.pi)) * math.exp(-(((i - mean) / (2 * sigma)) ** 2))
          for i in range(kernel_size)]
    tot = sum(a * b for a in g1 for b in g1)
    return [a * b / tot for a in g1 for b in g1]


def _w9(weights):
    return (ctypes.c_float * 9)(*[float(v) for v in weights])


def attn_smooth(p, weights):
    """p [F, heads, hw, n_text] fp16 -> blurred map of the same shape."""
    assert p.is_cuda and p.dtype == torch.float16 and p.is_contiguous() and p.dim() == 4
    out = torch.empty_like(p)
    _lib.check(_lib.lib().lvd_attn_smooth(p.data_ptr(), out.data_ptr(), p.shape[0] * p.shape[1], p.shape[2], p.shape[3],
                                          _w9(weights), _stream()), "attn_smooth")
    return out


def attn_smooth_bwd(d_out, weights, half_out):
    assert d_out.dtype == torch.float32 and d_out.is_contiguous()
    dp = torch.empty(d_out.shape, dtype=torch.float16 if half_out else torch.float32, device=d_out.device)
    _lib.check(_lib.lib().lvd_attn_smooth_bwd(d_out.data_ptr(), dp.data_ptr(), 1 if half_out else 0,
                                              d_out.shape[0] * d_out.shape[1], d_out.shape[2], d_out.shape[3],
                                              _w9(weights), _stream()), "attn_smooth_bwd")
    return dp


def attn_renorm(p, num_tokens, scale):
    assert p.is_cuda and p.dtype == torch.float16 and p.is_contiguous() and p.dim() == 4
    out = torch.empty(p.shape[:3] + (num_tokens - 2,), dtype=torch.float16, device=p.device)
    _lib.check(_lib.lib().lvd_attn_renorm(p.data_ptr(), out.data_ptr(), p.shape[0] * p.shape[1] * p.shape[2], p.shape[3],
                                          int(num_tokens), float(scale), _stream()), "attn_renorm")
    return out


def attn_renorm_bwd(r_out, d_r, n_text, num_tokens, scale, half_out):
    assert d_r.dtype == torch.float32 and d_r.is_contiguous() and r_out.is_contiguous()
    dp = torch.empty(r_out.shape[:3] + (n_text,), dtype=torch.float16 if half_out else torch.float32,
                     device=r_out.device)
    _lib.check(_lib.lib().lvd_attn_renorm_bwd(r_out.data_ptr(), d_r.data_ptr(), dp.data_ptr(), 1 if half_out else 0,
                                              r_out.shape[0] * r_out.shape[1] * r_out.shape[2], int(n_text),
                                              int(num_tokens), float(scale), _stream()), "attn_renorm_bwd")
    return dp


def attn_columns(p, slot_col, h_in, w_in, h_out, w_out):
    """p [F, heads, h_in*w_in, n_cols] -> [n_slots, F, heads, h_out*w_out] (bilinear when the sizes differ)."""
    assert p.is_cuda and p.dtype == torch.float16 and p.is_contiguous() and slot_col.dtype == torch.int32
    assert p.shape[2] == h_in * w_in
    cols = torch.empty((slot_col.numel(), p.shape[0], p.shape[1], h_out * w_out), dtype=torch.float16, device=p.device)
    _lib.check(_lib.lib().lvd_attn_columns(p.data_ptr(), slot_col.data_ptr(), cols.data_ptr(), slot_col.numel(),
                                           p.shape[0] * p.shape[1], p.shape[3], h_in, w_in, h_out, w_out, _stream()),
               "attn_columns")
    return cols


def attn_columns_bwd(d_cols, slot_col, n_cols, h_in, w_in, h_out, w_out, half_out):
    assert d_cols.dtype == torch.float16 and d_cols.is_contiguous()
    n_slots, frames, heads, _ = d_cols.shape
    dm = torch.empty((frames, heads, h_in * w_in, n_cols), dtype=torch.float16 if half_out else torch.float32,
                     device=d_cols.device)
    _lib.check(_lib.lib().lvd_attn_columns_bwd(d_cols.data_ptr(), slot_col.data_ptr(), dm.data_ptr(),
                                               1 if half_out else 0, n_slots, frames * heads, n_cols, h_in, w_in, h_out,
                                               w_out, _stream()), "attn_columns_bwd")
    return dm


# ----------------------------------------------------------------------------------------------- GLIGEN helpers
def copy_rows_seq(x, out, n, *, src_seq, src_off, dst_seq, dst_off, groups):
    _check_act(x, "x")
    _check_act(out, "out")
    _lib.check(_lib.lib().lvd_copy_rows_seq(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), groups * n,
                                            x.shape[1], n, src_seq, src_off, dst_seq, dst_off, _stream()),
               "copy_rows_seq")
    return out


def gligen_prep(boxes, masks, positive_embeddings, null_positive, null_position):
    """boxes (N, n_obj, 4), masks (N, n_obj), positive_embeddings (N, n_obj, D) -> [N*n_obj, D+64] fp16."""
    n, no, d = positive_embeddings.shape
    dev = null_positive.device
    b = boxes.to(device=dev, dtype=torch.float16).contiguous()
    m = masks.to(device=dev, dtype=torch.float16).contiguous()
    pe = positive_embeddings.to(device=dev, dtype=torch.float16).contiguous()
    out = torch.empty((n * no, d + 64), dtype=torch.float16, device=dev)
    _lib.check(_lib.lib().lvd_gligen_prep(b.data_ptr(), m.data_ptr(), pe.data_ptr(), null_positive.data_ptr(),
                                          null_position.data_ptr(), out.data_ptr(), n * no, d, _stream()),
               "gligen_prep")
    return out


# =============================================================================================== backward (dgrad) ops
def _gn_workspace2(dev, samples, groups):
    return _gn_workspace(dev, 2 * samples, groups)


# The adjoint GEMMs contract over the forward's output channels.  Reading the forward weight MN-major needs no copy but
# restricts the tile width to multiples of 64 (N = 320 -> 64-wide tiles at ~450 TFLOP/s); with 180 GB of HBM a K-major
# transposed copy per weight (made once, on first use: +2.8 GB for the whole guidance prefix) lets the adjoints run the
# same 160/256-wide tiles and CTA pairs as the forward.  DGRAD_TRANSPOSED = False restores the zero-copy path.
DGRAD_TRANSPOSED = True


def _dgrad_weight(w, taps):
    """forward weight [Cout, taps*Cin] (K ordered (tap, ci)) -> [Cin, taps*Cout] (K ordered (tap, co)); the copy is
    cached on the weight tensor object itself, so it lives exactly as long as the weight."""
    cache = getattr(w, "_lvd_dgrad", None)
    if cache is None:
        cache = {}
        w._lvd_dgrad = cache
    t = cache.get(taps)
    if t is None:
        cout = w.shape[0]
        cin = w.shape[1] // taps
        t = w.reshape(cout, taps, cin).permute(2, 1, 0).reshape(cin, taps * cout).contiguous()
        cache[taps] = t
    return t


def linear_dgrad(dy, w, *, residual=None, out=None, row_map=None):
    """dx = dy @ w  for a forward y = x @ w.T (w [N_out, K_in], as stored for the forward)."""
    if DGRAD_TRANSPOSED:
        return linear(dy, _dgrad_weight(w, 1), residual=residual, out=out, row_map=row_map)
    return linear(dy, w, b_mn_major=True, n=w.shape[1], residual=residual, out=out, row_map=row_map)


def conv2d_3x3_dgrad(dy, n_img, h, w_, wk, *, residual=None, out=None):
    """Adjoint of conv2d_3x3 w.r.t. its input: 9 shifted boxes of dy (negated offsets) against the same
    [Cout, 9*Cin] weight read MN-major with a per-tap column offset."""
    _check_act(dy, "dy")
    cout = wk.shape[0]
    cin = wk.shape[1] // 9
    assert dy.shape == (n_img * h * w_, cout)
    if out is None:
        out = _alloc_out(dy.shape[0], cin, dy)
    ld = dy.stride(0)
    bw, bh, bn = choose_box((w_, h, n_img))
    taps = [(0, -dx, -dy_, 0, 0) for dy_ in (-1, 0, 1) for dx in (-1, 0, 1)]
    if DGRAD_TRANSPOSED:
        return gemm_raw(a=dy, a_dims=(cout, w_, h, n_img, 1),
                        a_strides=(1, ld, ld * w_, ld * w_ * h, ld * w_ * h * n_img), a_box=(64, bw, bh, bn, 1),
                        taps=taps, k_per_tap=cout, w=_dgrad_weight(wk, 9), n=cin, out=out, out_dims=(w_, h, n_img, 1),
                        out_row_stride=(1, w_, w_ * h, 0), residual=residual)
    return gemm_raw(a=dy, a_dims=(cout, w_, h, n_img, 1), a_strides=(1, ld, ld * w_, ld * w_ * h, ld * w_ * h * n_img),
                    a_box=(64, bw, bh, bn, 1), taps=taps, k_per_tap=cout, w=wk, n=cin, out=out,
                    out_dims=(w_, h, n_img, 1), out_row_stride=(1, w_, w_ * h, 0), residual=residual, b_mn_major=True,
                    b_k_tap_stride=0, b_n_tap_stride=cin)


def conv_temporal_dgrad(dy, batch, frames, hw, wk, *, residual=None, out=None):
    _check_act(dy, "dy")
    cout = wk.shape[0]
    cin = wk.shape[1] // 3
    if out is None:
        out = _alloc_out(dy.shape[0], cin, dy)
    ld = dy.stride(0)
    bp, bf, bb = choose_box((hw, frames, batch))
    taps = [(0, 0, -df, 0, 0) for df in (-1, 0, 1)]
    if DGRAD_TRANSPOSED:
        return gemm_raw(a=dy, a_dims=(cout, hw, frames, batch, 1),
                        a_strides=(1, ld, ld * hw, ld * hw * frames, ld * hw * frames * batch),
                        a_box=(64, bp, bf, bb, 1), taps=taps, k_per_tap=cout, w=_dgrad_weight(wk, 3), n=cin, out=out,
                        out_dims=(hw, frames, batch, 1), out_row_stride=(1, hw, hw * frames, 0), residual=residual)
    return gemm_raw(a=dy, a_dims=(cout, hw, frames, batch, 1),
                    a_strides=(1, ld, ld * hw, ld * hw * frames, ld * hw * frames * batch), a_box=(64, bp, bf, bb, 1),
                    taps=taps, k_per_tap=cout, w=wk, n=cin, out=out, out_dims=(hw, frames, batch, 1),
                    out_row_stride=(1, hw, hw * frames, 0), residual=residual, b_mn_major=True, b_k_tap_stride=0,
                    b_n_tap_stride=cin)


def conv2d_3x3_s2_dgrad(dy, n_img, h, w_, wk_parity, *, out=None):
    """Adjoint of conv2d_3x3_s2 (input h x w_, output h/2 x w_/2).  One launch per input-pixel parity class (a, b);
    `wk_parity[(a, b)]` = ([Cout, ntaps*Cin] weight slice, [(sx, sy)] output-pixel shifts) from
    repack.s2_dgrad_weights."""
    _check_act(dy, "dy")
    ho, wo = h // 2, w_ // 2
    cout = dy.shape[1]
    any_w = wk_parity[(0, 0)][0]
    cin = any_w.shape[1]
    if out is None:
        out = _alloc_out(n_img * h * w_, cin, dy)
    ld = dy.stride(0)
    bw, bh, bn = choose_box((wo, ho, n_img))
    for (a, b), (wsub, shifts) in wk_parity.items():
        taps = [(0, sx, sy, 0, 0) for (sx, sy) in shifts]
        if DGRAD_TRANSPOSED:
            gemm_raw(a=dy, a_dims=(cout, wo, ho, n_img, 1),
                     a_strides=(1, ld, ld * wo, ld * wo * ho, ld * wo * ho * n_img), a_box=(64, bw, bh, bn, 1),
                     taps=taps, k_per_tap=cout, w=_dgrad_weight(wsub, len(shifts)), n=cin, out=out[a * w_ + b:],
                     out_dims=(wo, ho, n_img, 1), out_row_stride=(2, 2 * w_, h * w_, 0))
            continue
        gemm_raw(a=dy, a_dims=(cout, wo, ho, n_img, 1), a_strides=(1, ld, ld * wo, ld * wo * ho, ld * wo * ho * n_img),
                 a_box=(64, bw, bh, bn, 1), taps=taps, k_per_tap=cout, w=wsub, n=cin, out=out[a * w_ + b:],
                 out_dims=(wo, ho, n_img, 1), out_row_stride=(2, 2 * w_, h * w_, 0), b_mn_major=True,
                 b_k_tap_stride=0, b_n_tap_stride=cin)
    return out


def gn_stats_buffer(samples, dev, groups=32):
    return torch.empty((samples, groups, 2), dtype=torch.float32, device=dev)


def groupnorm_bwd(x, dz, samples, gamma, beta, eps, *, groups=32, silu=False, out=None, accumulate=False, stats=None):
    _check_act(x, "x")
    _check_act(dz, "dz")
    rows, c = x.shape
    if out is None:
        assert not accumulate
        out = _alloc_out(rows, c, x)
    ws = _gn_workspace2(x.device, samples, groups)
    with _Timed("groupnorm_bwd", 0.0, 2.0 * rows * c * 4, "gnb s%d r%d c%d" % (samples, rows // samples, c)):
        _lib.check(_lib.lib().lvd_groupnorm_bwd(x.data_ptr(), x.stride(0), dz.data_ptr(), dz.stride(0), out.data_ptr(),
                                                out.stride(0), samples, rows // samples, c, groups, gamma.data_ptr(),
                                                beta.data_ptr(), float(eps), 1 if silu else 0, 1 if accumulate else 0,
                                                stats.data_ptr() if stats is not None else None, ws.data_ptr(),
                                                ws.numel(), _stream()), "groupnorm_bwd")
    return out


def layernorm_bwd(x, dy, gamma, eps=1e-5, *, out=None, accumulate=False):
    _check_act(x, "x")
    _check_act(dy, "dy")
    rows, c = x.shape
    if out is None:
        assert not accumulate
        out = _alloc_out(rows, c, x)
    with _Timed("layernorm_bwd", 0.0, 2.0 * rows * c * 3, "lnb r%d c%d" % (rows, c)):
        _lib.check(_lib.lib().lvd_layernorm_bwd(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), out.data_ptr(),
                                                out.stride(0), rows, c, gamma.data_ptr(), float(eps),
                                                1 if accumulate else 0, _stream()), "layernorm_bwd")
    return out


def geglu(pre, out=None):
    _check_act(pre, "pre")
    rows, n2 = pre.shape
    if out is None:
        out = _alloc_out(rows, n2 // 2, pre)
    with _Timed("eltwise", 0.0, 2.0 * rows * (n2 + n2 // 2), "geglu"):
        _lib.check(_lib.lib().lvd_geglu(pre.data_ptr(), pre.stride(0), out.data_ptr(), out.stride(0), rows, n2 // 2,
                                        _stream()), "geglu")
    return out


def geglu_bwd(pre, dout, out=None):
    _check_act(pre, "pre")
    _check_act(dout, "dout")
    rows, n2 = pre.shape
    if out is None:
        out = _alloc_out(rows, n2, pre)
    with _Timed("eltwise", 0.0, 2.0 * rows * (2 * n2 + n2 // 2), "geglu_bwd"):
        _lib.check(_lib.lib().lvd_geglu_bwd(pre.data_ptr(), pre.stride(0), dout.data_ptr(), dout.stride(0), out.data_ptr(),
                                            out.stride(0), rows, n2 // 2, _stream()), "geglu_bwd")
    return out


def add_rows(a, b, out=None):
    _check_act(a, "a")
    _check_act(b, "b")
    if out is None:
        out = _alloc_out(a.shape[0], a.shape[1], a)
    with _Timed("eltwise", 0.0, 6.0 * a.shape[0] * a.shape[1], "add_rows"):
        _lib.check(_lib.lib().lvd_add_rows(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                                           out.stride(0), a.shape[0], a.shape[1], _stream()), "add_rows")
    return out


def upsample2x_bwd(dy, n_img, h, w_):
    """dy rows of the (2h x 2w_) image -> dx rows of the (h x w_) image."""
    _check_act(dy, "dy")
    out = _alloc_out(n_img * h * w_, dy.shape[1], dy)
    with _Timed("eltwise", 0.0, 2.0 * 5 * out.shape[0] * dy.shape[1], "upsample2x_bwd"):
        _lib.check(_lib.lib().lvd_upsample2x_bwd(dy.data_ptr(), dy.stride(0), out.data_ptr(), out.stride(0), n_img, h, w_,
                                                 dy.shape[1], _stream()), "upsample2x_bwd")
    return out


def conv_in_bwd(drows, weight, batch, frames, h, w_):
    _check_act(drows, "drows")
    cout, cin = weight.shape[0], weight.shape[1]
    out = torch.empty((batch, cin, frames, h, w_), dtype=torch.float32, device=drows.device)
    _lib.check(_lib.lib().lvd_conv_in_bwd(drows.data_ptr(), drows.stride(0), weight.data_ptr(), out.data_ptr(), batch,
                                          cin, frames, h, w_, cout, _stream()), "conv_in_bwd")
    return out


def latent_update(latents, grad_f32, coef, unscale, want_grad=False):
    out = torch.empty_like(latents)
    gh = torch.empty_like(latents) if want_grad else None
    _lib.check(_lib.lib().lvd_latent_update(latents.data_ptr(), grad_f32.data_ptr(), out.data_ptr(),
                                            gh.data_ptr() if gh is not None else None, latents.numel(), float(coef),
                                            float(unscale), _stream()), "latent_update")
    return out, gh


def attn_rowdot(o, d_o, heads):
    _check_act(o, "o")
    _check_act(d_o, "d_o")
    out = torch.empty((o.shape[0], heads), dtype=torch.float32, device=o.device)
    _lib.check(_lib.lib().lvd_attn_rowdot(o.data_ptr(), o.stride(0), d_o.data_ptr(), d_o.stride(0), out.data_ptr(),
                                          o.shape[0], heads, _stream()), "attn_rowdot")
    return out


def self_attention_bwd(qkv, d_o, o, lse, *, heads, nz, s, scale=0.125, out=None):
    """qkv [nz*s, 3C] (q|k|v); returns d_qkv of the same shape."""
    _check_act(qkv, "qkv")
    _check_act(d_o, "d_o")
    c = heads * 64
    if out is None:
        out = torch.empty_like(qkv)
    dvec = attn_rowdot(o, d_o, heads)
    with _Timed("attention_self_bwd", 10.0 * nz * heads * s * s * 64, 0.0):
        _lib.check(_lib.lib().lvd_self_attn_bwd(qkv.data_ptr(), qkv.stride(0), qkv.shape[1], 0, c, 2 * c, d_o.data_ptr(),
                                                d_o.stride(0), lse.data_ptr(), dvec.data_ptr(), out.data_ptr(),
                                                out.stride(0), s, heads, nz, float(scale), _stream()), "self_attn_bwd")
    return out


def temporal_attention_bwd(qkv, d_o, *, heads, n_seq, frames, scale=0.125, out=None, tensor_core=True):
    _check_act(qkv, "qkv")
    _check_act(d_o, "d_o")
    c = heads * 64
    if out is None:
        out = torch.empty_like(qkv)
    with _Timed("attention_temporal_bwd", 10.0 * n_seq * heads * frames * frames * 64, 0.0):
        _lib.check(_lib.lib().lvd_temporal_attn_bwd(qkv.data_ptr(), qkv.stride(0), qkv.shape[1] if tensor_core else 0, 0, c,
                                                    2 * c, d_o.data_ptr(),
                                                    d_o.stride(0), out.data_ptr(), out.stride(0), n_seq, heads, frames,
                                                    float(scale), _stream()), "temporal_attn_bwd")
    return out


def cross_attention_bwd(q, kv, d_o, *, heads, nz, s_q, s_k, kv_z_div, dp_tok=None, slot_cols=(), scale=0.125, out=None,
                        tensor_core=True, dp_dense=None, slot_cols_dev=None):
    """dQ of the cross-attention; dE/dP enters either as token columns (dp_tok [n_slots, nz, heads, s_q] + slot_cols) or
    as a dense map (dp_dense [nz, heads, s_q, s_k])."""
    _check_act(q, "q")
    _check_act(kv, "kv")
    _check_act(d_o, "d_o")
    c = heads * 64
    if out is None:
        out = _alloc_out(q.shape[0], c, q)
    cols = (ctypes.c_int32 * max(1, len(slot_cols)))(*[int(x) for x in slot_cols])
    n_slots = len(slot_cols) if dp_tok is not None else 0
    if slot_cols_dev is not None and dp_tok is not None:
        assert slot_cols_dev.dtype == torch.int32 and slot_cols_dev.is_cuda
        n_slots = int(slot_cols_dev.numel())
        cols = (ctypes.c_int32 * max(1, n_slots))(*([0] * max(1, n_slots)))
    if dp_dense is not None:
        assert dp_tok is None and dp_dense.dtype == torch.float16 and dp_dense.is_contiguous()
        assert tuple(dp_dense.shape) == (nz, heads, s_q, s_k) and tensor_core
        dp_tok, n_slots = dp_dense, -1
    with _Timed("attention_cross_bwd", 6.0 * nz * heads * s_q * s_k * 64, 0.0):
        _lib.check(_lib.lib().lvd_cross_attn_bwd(q.data_ptr(), q.stride(0), q.shape[1] if tensor_core else 0, kv.data_ptr(),
                                                 kv.stride(0), kv.shape[1] if tensor_core else 0, 0, c,
                                                 d_o.data_ptr(), d_o.stride(0),
                                                 dp_tok.data_ptr() if dp_tok is not None else None, cols,
                                                 n_slots, out.data_ptr(),
                                                 out.stride(0), s_q, s_k, heads, nz, kv_z_div, float(scale),
                                                 slot_cols_dev.data_ptr() if (slot_cols_dev is not None and
                                                                              n_slots > 0) else None, _stream()),
                   "cross_attn_bwd")
    return out
