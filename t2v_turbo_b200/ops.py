"""Operator layer: torch tensors in / out, every arithmetic op is a libt2v_b200.so call.

Activations are channels-last bf16.  A frame batch is ``[N, H, W, C]`` (N = B*T); the same memory
is the token matrix ``[N*H*W, C]`` for every per-token op (Linear / LayerNorm / GEGLU), the
spatial sequences ``[N, H*W, C]`` and — through strides only — the temporal sequences
``[B*H*W, T, C]``.  None of the reference's ``rearrange(...).contiguous()`` copies
(attention.py:379,386,475-511; openaimodel3d.py:38-40,251-253) exist here.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import os as _os

import torch

from . import _lib
from ._lib import (AttnDesc, GemmDesc, GroupNormDesc, LayerNormDesc, ShortAttnDesc, SmallLinearDesc,
                   check, lib, ptr, stream_ptr)

BF16 = torch.bfloat16

# ----------------------------------------------------------------------------- launch accounting
KERNELS_PER_CALL = {"groupnorm": 2}     # stats + apply (the cluster path and split-K GEMMs are counted as one: a lower bound)
LAUNCHES = 0                            # kernels of libt2v_b200.so enqueued so far (incl. during graph capture)
_PROF = None                            # list of (family, flops, ev_start, ev_end) while profiling
_FLOPS: dict = {}
_TAG: dict = {}
_BYTES: dict = {}


def start_profile():
    """Per-call CUDA-event timing by kernel family (eager mode only; used by bench.py for the roofline)."""
    global _PROF
    _PROF = []


def stop_profile(by_tag=False):
    """-> {family (or family+shape tag): dict(calls, ms, flops)} after synchronising."""
    global _PROF
    recs, _PROF = _PROF, None
    torch.cuda.synchronize()
    out = {}
    for fam, flops, e0, e1, tag, nbytes in recs:
        d = out.setdefault((fam + " " + tag) if by_tag else fam, dict(calls=0, ms=0.0, flops=0, bytes=0))
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += flops
        d["bytes"] += nbytes       # algorithmic HBM bytes (each operand / result counted once), where the wrapper states them
    return out


def _launch(family, flops, fn, *args):
    global LAUNCHES
    LAUNCHES += KERNELS_PER_CALL.get(family, 1)
    if _PROF is None:
        check(fn(*args), "t2v_" + family)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), "t2v_" + family)
    e1.record()
    _PROF.append((family, flops, e0, e1, _TAG.pop(family, ""), _BYTES.pop(family, 0)))


# ----------------------------------------------------------------------------- tile planning
@functools.lru_cache(maxsize=None)
def plan_box(sizes: tuple, fixed: tuple = (None, None, None, None)) -> tuple:
    """Pick the <=128-point tile box over a (x1..x4) point grid that wastes the fewest MMA rows."""
    cands = []
    for s, f in zip(sizes, fixed):
        if f is not None:
            cands.append([f])
            continue
        c = {d for d in range(1, min(s, 128) + 1) if s % d == 0}
        c |= {1 << k for k in range(8) if (1 << k) <= max(1, min(128, 2 * s))}
        cands.append(sorted(c))
    best, best_key = None, None
    total = math.prod(sizes)
    for b1 in cands[0]:
        for b2 in cands[1]:
            if b1 * b2 > 128:
                break
            for b3 in cands[2]:
                if b1 * b2 * b3 > 128:
                    break
                for b4 in cands[3]:
                    rows = b1 * b2 * b3 * b4
                    if rows > 128:
                        break
                    if rows % 8:
                        continue
                    tiles = 1
                    for s, b in zip(sizes, (b1, b2, b3, b4)):
                        tiles *= -(-s // b)
                    eff = total / (tiles * 128)
                    key = (round(eff, 6), b1, b2, b3)
                    if best_key is None or key > best_key:
                        best_key, best = key, (b1, b2, b3, b4)
    if best is None:
        raise ValueError(f"no tile box for point grid {sizes}")
    return best


def _fill(arr, vals):
    for i, v in enumerate(vals):
        arr[i] = int(v)


def _as_pair(x):
    if isinstance(x, (tuple, list)):
        assert len(x) == 2
        return x[0], x[1]
    return x, None


def _gemm_raw(*, a, a_ch, a_ch_total, a_size, a_stride, box, taps, tap_ch_off, w, n_rows, out,
              o_size, o_stride, n_out, bias=None, bias_row_stride=0, bias_dim=-1, bias_div=1,
              residual=None, r_stride=None, alpha=1.0, flags=0, block_n=0, b_batches=1,
              b_batch_stride=0, b_batch_dim=-1, b_row_stride=0, split_k=0, ln=None, row_accum=None, col_accum=None):
    d = GemmDesc()
    a0, a1 = a
    d.a[0] = a0.data_ptr()
    d.a[1] = a1.data_ptr() if a1 is not None else None
    _fill(d.a_ch, a_ch)
    _fill(d.a_ch_total, a_ch_total)
    _fill(d.a_size, a_size)
    _fill(d.a_stride[0], a_stride[0])
    _fill(d.a_stride[1], a_stride[1] if a_stride[1] is not None else (0, 0, 0, 0))
    _fill(d.box, box)
    d.n_taps = len(taps)
    for t, off in enumerate(taps):
        _fill(d.tap_off[t], off)
        d.tap_ch_off[t] = int(tap_ch_off[t]) if tap_ch_off is not None else 0
    d.b = w.data_ptr()
    d.b_rows = n_rows
    d.b_batches = b_batches
    d.b_batch_stride = b_batch_stride
    d.b_batch_dim = b_batch_dim
    d.b_row_stride = b_row_stride
    d.out = out.data_ptr()
    _fill(d.o_size, o_size)
    _fill(d.o_stride, o_stride)
    d.n_out = n_out
    d.bias = ptr(bias)
    d.bias_row_stride = bias_row_stride
    d.bias_dim = bias_dim
    d.bias_div = bias_div
    d.residual = ptr(residual)
    _fill(d.r_stride, r_stride if r_stride is not None else o_stride)
    d.alpha = alpha
    d.flags = flags | _lib.WS_CLEAN
    d.block_n = block_n
    d.split_k = split_k
    d.tune = GEMM_TUNE
    if ln is not None:   # folded LayerNorm: (row_stats [M,2] fp32, col_sum [N] fp32[, (channels, eps) if the stats are raw sums])
        d.row_stats, d.col_sum = ln[0].data_ptr(), ln[1].data_ptr()
        if len(ln) > 2:
            d.ln_raw, d.ln_channels, d.ln_eps = 1, int(ln[2][0]), float(ln[2][1])
    if row_accum is not None:   # fp32 [M,2], zeroed by the caller: += (sum, sum of squares) of each output row
        d.row_accum = row_accum.data_ptr()
    if col_accum is not None:   # (fp32 [samples, N, 2] zeroed by the caller, cs_mult): GroupNorm statistics of the output
        d.col_accum = col_accum[0].data_ptr()
        _fill(d.cs_mult, col_accum[1])
    ws = _splitk_workspace(out.device)
    d.workspace = ws.data_ptr()
    d.workspace_bytes = ws.numel() * 4
    _FLOPS["gemm"] = 2 * math.prod(int(v) for v in o_size) * int(n_rows) * len(taps) * (int(a_ch[0]) + int(a_ch[1]))
    if _PROF is not None:
        m_pts, k_ch = math.prod(int(v) for v in o_size), int(a_ch[0]) + int(a_ch[1])
        # algorithmic bytes: the input activation once (taps re-read it through L2, not HBM), the weights once, the output,
        # the residual operand if any (bf16 everywhere)
        _BYTES["gemm"] = 2 * (m_pts * k_ch + int(n_rows) * len(taps) * k_ch * max(1, int(b_batches)) + m_pts * int(n_out) * (2 if residual is not None else 1))
        _TAG["gemm"] = (f"M={math.prod(int(v) for v in o_size)} N={int(n_rows)} K={len(taps)}x{int(a_ch[0]) + int(a_ch[1])} "
                        f"box={tuple(int(b) for b in box)} bn={block_n} flags={flags} res={int(residual is not None)}")
    _launch("gemm", _FLOPS.pop("gemm", 0), lib().t2v_gemm, C.byref(d), stream_ptr())
    return out


# GroupNorm statistics in the producing GEMM's epilogue (T2VGemmDesc.col_accum -> T2VGroupNormDesc.chan_sums).  The
# epilogue reduces each staged 128 x 32 bf16 chunk over its rows; tiles inside one sample (large images: VAE, UNet
# levels 0-1) accumulate in per-warp shared-memory tables flushed once per tile, tiles spanning samples (small images)
# reduce 8-row blocks straight to the global sums.  (Round 1 issued a global vector atomic per 8 rows everywhere: 21 M
# atomics onto 16 KB for one VAE conv, slower than the statistics pass it removed.)  Measured on B200 in round 2 (scripts/
# gemm_bench.py STATS=0/1, profiles/r02_gn_fusion.txt): +1 us on the MMA-bound convs (K >= 2304) but +5 us on the K = 960
# temporal convs and +29 % on the epilogue-bound 128-channel VAE convs; whole-step frames/s off 158.6 / conv 153.8 / all
# 147.3 under the 1 kW power cap.  The standalone statistics pass already streams at 5.6-5.9 TB/s on the tensors that
# matter, so the fusion stays OFF by default.
# T2V_GN_FUSE = off | conv (producers with K >= T2V_GN_FUSE_MIN_K, per-frame consumers) | all.
GN_FUSE = _os.environ.get("T2V_GN_FUSE", "off")
GN_FUSE_MIN_K = int(_os.environ.get("T2V_GN_FUSE_MIN_K", "2304"))


def gn_fuse_producer(k_total, grid=None, fixed=(None, None, None, None), sample_dims=()):
    """Should a GEMM with reduction length k_total over the output point grid `grid` accumulate GroupNorm statistics
    for its consumer?  (The epilogue reduces rows in runs of 8 along dim 0: the tile box must allow that.  sample_dims
    names the dims that index samples; tiles spanning them use the direct-reduction mode.)"""
    if GN_FUSE == "off" or (GN_FUSE == "conv" and k_total < GN_FUSE_MIN_K):
        return False
    return grid is None or plan_box(tuple(int(v) for v in grid), fixed)[0] % 8 == 0


def gn_fuse_temporal():
    return GN_FUSE == "all"


GEMM_TUNE = int(_os.environ.get("T2V_GEMM_TUNE", "0"), 0)   # experiments only: see T2VGemmDesc.tune
_SPLITK_WS: dict = {}
SPLITK_WS_BYTES = 32 << 20


def _splitk_workspace(device):
    """Persistent fp32 scratch for split-K partial sums + tile counters, one per device.  Zeroed ONCE here: every
    split-K call finds it zeroed and leaves it zeroed (T2V_WS_CLEAN: the finalize kernel clears what it reads, so no
    zero kernel runs).  Calls on one device are stream-ordered by the host mirror (one stream at a time)."""
    key = device
    ws = _SPLITK_WS.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("t2v_turbo_b200: split-K workspace must be allocated before CUDA-graph capture (run a warm-up call)")
        ws = torch.zeros(SPLITK_WS_BYTES // 4, device=device, dtype=torch.float32)
        _SPLITK_WS[key] = ws
    return ws


def _check_act(x, name="x"):
    assert x.is_cuda and x.dtype == BF16 and x.is_contiguous(), f"{name}: need contiguous CUDA bf16"


# ----------------------------------------------------------------------------- Linear
def linear(x, w, bias=None, *, residual=None, geglu=False, gelu=False, out=None, out_f32=False,
           alpha=1.0, block_n=0, split_k=0, ln=None, row_accum=None):
    """out[m, :] = epi(x[m, :] @ w.T).  x: [M, K] bf16 (or a pair concatenated along K);
    w: [N, K] bf16 (GEGLU: rows packed by pack_geglu); bias: fp32 [N]; residual: bf16 [M, n_out].
    ln = (row_stats, col_sum[, (channels, eps)]): x is the UN-normalised LayerNorm input and w / bias / col_sum come
    from fold_layernorm; with the third element row_stats holds the raw sums a producer GEMM accumulated.
    row_accum: fp32 [M, 2] (zeroed): this GEMM adds (sum, sum of squares) of each output row (it produces a LayerNorm input)."""
    x0, x1 = _as_pair(x)
    _check_act(x0)
    m = x0.shape[0]
    k0 = x0.shape[1]
    k1 = x1.shape[1] if x1 is not None else 0
    n = w.shape[0]
    assert w.dtype == BF16 and w.is_contiguous() and w.shape[1] == k0 + k1, (w.shape, k0, k1)
    n_out = n // 2 if geglu else n
    if out is None:
        out = torch.empty((m, n_out), device=x0.device, dtype=torch.float32 if out_f32 else BF16)
    flags = (_lib.EPI_GEGLU if geglu else 0) | (_lib.EPI_OUT_F32 if out_f32 else 0) | (_lib.EPI_GELU if gelu else 0)
    return _gemm_raw(
        a=(x0, x1), a_ch=(k0, k1), a_ch_total=(k0, k1), a_size=(m, 1, 1, 1),
        a_stride=((k0, 0, 0, 0), (k1, 0, 0, 0) if x1 is not None else None),
        box=(128, 1, 1, 1), taps=[(0, 0, 0, 0)], tap_ch_off=None, w=w, n_rows=n, out=out,
        o_size=(m, 1, 1, 1), o_stride=(out.stride(0), 0, 0, 0), n_out=n_out, bias=bias,
        residual=residual, r_stride=(residual.stride(0), 0, 0, 0) if residual is not None else None,
        alpha=alpha, flags=flags, block_n=block_n, split_k=split_k, ln=ln, row_accum=row_accum)


def linear_frames(x, w, bias=None, *, hw, residual=None, out=None, stats=None, block_n=0, split_k=0):
    """linear() over a token matrix [n_frames*hw, K] tiled per frame, so that the epilogue can accumulate the per-frame
    per-channel GroupNorm statistics of the output (stats: fp32 [n_frames, N, 2], zeroed)."""
    _check_act(x)
    m, k = x.shape
    assert m % hw == 0
    nf = m // hw
    n = w.shape[0]
    assert w.dtype == BF16 and w.is_contiguous() and w.shape[1] == k
    if out is None:
        out = torch.empty((m, n), device=x.device, dtype=BF16)
    assert out.stride(0) == n and x.stride(0) == k and (residual is None or residual.stride(0) == n)
    box = plan_box((hw, nf, 1, 1))
    return _gemm_raw(
        a=(x, None), a_ch=(k, 0), a_ch_total=(k, 0), a_size=(hw, nf, 1, 1), a_stride=((k, hw * k, 0, 0), None),
        box=box, taps=[(0, 0, 0, 0)], tap_ch_off=None, w=w, n_rows=n, out=out,
        o_size=(hw, nf, 1, 1), o_stride=(n, hw * n, 0, 0), n_out=n, bias=bias, residual=residual,
        block_n=block_n, split_k=split_k, col_accum=(stats, (0, 1, 0, 0)) if stats is not None else None)


def bmm_nt(a, b, *, out=None, alpha=1.0, block_n=0):
    """Batched out[i] = a[i] @ b[i].T.  a: [Bt, M, K], b: [Bt, N, K] bf16 views with arbitrary batch / row
    strides (K contiguous) — used by the VAE AttnBlock (ae_modules.py:55-68) on slices of fused projections."""
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(2) == 1 and b.stride(2) == 1
    bt, m, k = a.shape
    n = b.shape[1]
    assert b.shape[0] == bt and b.shape[2] == k and k % 64 == 0, (a.shape, b.shape)
    if out is None:
        out = torch.empty((bt, m, n), device=a.device, dtype=BF16)
    return _gemm_raw(
        a=(a, None), a_ch=(k, 0), a_ch_total=(k, 0), a_size=(m, bt, 1, 1),
        a_stride=((a.stride(1), a.stride(0), 0, 0), None), box=(128, 1, 1, 1), taps=[(0, 0, 0, 0)], tap_ch_off=None,
        w=b, n_rows=n, out=out, o_size=(m, bt, 1, 1), o_stride=(out.stride(1), out.stride(0), 0, 0),
        n_out=n, alpha=alpha, block_n=block_n, b_batches=bt, b_batch_stride=b.stride(0), b_batch_dim=1,
        b_row_stride=b.stride(1))


# ----------------------------------------------------------------------------- convolutions
_TAPS_3X3 = [(kx - 1, ky - 1, 0, 0) for ky in range(3) for kx in range(3)]
_TAPS_T3 = [(0, kt - 1, 0, 0) for kt in range(3)]


def conv3x3(x, w, bias=None, *, bias_div=1, residual=None, out=None, block_n=0, split_k=0, stats=None, out_f32=False):
    """3x3 / pad 1 / stride 1 conv over [N,H,W,C] (or a channel-concatenated pair).
    w: [Cout, 9*C] packed (tap-major); bias: fp32 [rows, Cout], row = frame // bias_div.
    stats: fp32 [N, Cout, 2] (zeroed): per-frame per-channel (sum, sum of squares) of the output for the next GroupNorm."""
    x0, x1 = _as_pair(x)
    _check_act(x0)
    n, h, wd, c0 = x0.shape
    c1 = x1.shape[3] if x1 is not None else 0
    cout = w.shape[0]
    assert w.shape[1] == 9 * (c0 + c1), (w.shape, c0, c1)
    if out is None:
        out = torch.empty((n, h, wd, cout), device=x0.device, dtype=torch.float32 if out_f32 else BF16)
    box = plan_box((wd, h, n, 1))
    return _gemm_raw(
        a=(x0, x1), a_ch=(c0, c1), a_ch_total=(c0, c1), a_size=(wd, h, n, 1),
        a_stride=((c0, wd * c0, h * wd * c0, 0), (c1, wd * c1, h * wd * c1, 0) if x1 is not None else None),
        box=box, taps=_TAPS_3X3, tap_ch_off=None, w=w, n_rows=cout, out=out,
        o_size=(wd, h, n, 1), o_stride=(cout, wd * cout, h * wd * cout, 0), n_out=cout, bias=bias,
        bias_row_stride=cout if bias is not None else 0, bias_dim=2, bias_div=bias_div,
        residual=residual, block_n=block_n, split_k=split_k, flags=_lib.EPI_OUT_F32 if out_f32 else 0,
        col_accum=(stats, (0, 0, 1, 0)) if stats is not None else None)


def conv3x3_s2(x, w, bias=None, *, out=None, block_n=0, stats=None, pad="sym"):
    """3x3 / stride 2 conv.  pad="sym": padding 1 on every side (UNet Downsample, openaimodel3d.py:65-72);
    pad="br": zero-pad right / bottom by one only, no other padding (KL-VAE Downsample, ae_modules.py:87-105).
    The input is read through a parity view [N, H/2, 2, W/2, 2*C] so every tap is a plain box load and the padding
    is TMA out-of-bounds zero fill."""
    _check_act(x)
    n, h, wd, c = x.shape
    assert h % 2 == 0 and wd % 2 == 0
    cout = w.shape[0]
    assert w.shape[1] == 9 * c
    h2, w2 = h // 2, wd // 2
    if out is None:
        out = torch.empty((n, h2, w2, cout), device=x.device, dtype=BF16)
    taps, choff = [], []
    assert pad in ("sym", "br")
    # source row of output row y and tap k: sym 2y+k-1, br 2y+k  ->  (parity, offset in half-resolution rows)
    src = {0: (1, -1), 1: (0, 0), 2: (1, 0)} if pad == "sym" else {0: (0, 0), 1: (1, 0), 2: (0, 1)}
    for ky in range(3):
        hpar, dh = src[ky]
        for kx in range(3):
            wpar, dw = src[kx]
            taps.append((dw, hpar, dh, 0))
            choff.append(wpar * c)
    b = plan_box((w2, 1, h2, n), fixed=(None, 1, None, None))
    return _gemm_raw(
        a=(x, None), a_ch=(c, 0), a_ch_total=(2 * c, 0), a_size=(w2, 2, h2, n),
        a_stride=((2 * c, wd * c, 2 * wd * c, h * wd * c), None), box=b, taps=taps, tap_ch_off=choff,
        w=w, n_rows=cout, out=out, o_size=(w2, 1, h2, n),
        o_stride=(cout, 0, w2 * cout, h2 * w2 * cout), n_out=cout, bias=bias,
        bias_row_stride=0, bias_dim=-1, block_n=block_n,
        col_accum=(stats, (0, 0, 0, 1)) if stats is not None else None)


# nearest-2x upsample + 3x3 conv == four 2x2 convs on the LOW-resolution input, one per output parity (py, px):
# output row 2y+py reads upsampled rows 2y+py-1 .. 2y+py+1, i.e. source rows {y-1, y, y} (py=0) or {y, y, y+1}
# (py=1); taps that hit the same source row are pre-summed.  Zero padding of the upsampled image maps onto TMA
# out-of-bounds zero fill of the source.  4/9 of the FLOPs and no materialised upsampled tensor.
_UP_ROWS = (((-1, (0,)), (0, (1, 2))), ((0, (0, 1)), (1, (2,))))   # parity -> [(source offset, kernel taps summed)]


def pack_upconv_weight(w):
    """conv weight [Cout, Cin, 3, 3] after a nearest-2x upsample -> bf16 [4 phases][Cout, 4*Cin] (tap-major K)."""
    w = w.detach().float()
    phases = []
    for py in range(2):
        for px in range(2):
            taps = []
            for _, kys in _UP_ROWS[py]:
                for _, kxs in _UP_ROWS[px]:
                    taps.append(sum(w[:, :, ky, kx] for ky in kys for kx in kxs))
            phases.append(torch.cat(taps, dim=1))
    return torch.stack(phases).to(BF16).contiguous()


def upconv3x3(x, w_phases, bias=None, *, out=None, block_n=0, stats=None):
    """Upsample(nearest, 2x) + Conv2d(3x3, pad 1) (openaimodel3d.py:75-108, ae_modules.py:50-63) on [N,H,W,C]
    -> [N,2H,2W,Cout].  w_phases from pack_upconv_weight; bias fp32 [Cout]."""
    _check_act(x)
    n, h, wd, c = x.shape
    cout = w_phases.shape[1]
    assert w_phases.shape == (4, cout, 4 * c), (w_phases.shape, c)
    if out is None:
        out = torch.empty((n, 2 * h, 2 * wd, cout), device=x.device, dtype=BF16)
    box = plan_box((wd, h, n, 1))
    for py in range(2):
        for px in range(2):
            taps = [(dx, dy, 0, 0) for dy, _ in _UP_ROWS[py] for dx, _ in _UP_ROWS[px]]
            _gemm_raw(
                a=(x, None), a_ch=(c, 0), a_ch_total=(c, 0), a_size=(wd, h, n, 1),
                a_stride=((c, wd * c, h * wd * c, 0), None), box=box, taps=taps, tap_ch_off=None,
                w=w_phases[2 * py + px], n_rows=cout, out=out[:, py::2, px::2, :],
                o_size=(wd, h, n, 1), o_stride=(2 * cout, 4 * wd * cout, 4 * h * wd * cout, 0), n_out=cout,
                bias=bias, bias_row_stride=0, bias_dim=-1, block_n=block_n,
                col_accum=(stats, (0, 0, 1, 0)) if stats is not None else None)
    return out


def tconv3(x, w, bias=None, *, residual=None, out=None, block_n=0, split_k=0, stats=None):
    """Conv3d (3,1,1) / pad (1,0,0) over x: [B, T, HW, C] (or a channel-concatenated pair) (TemporalConvBlock,
    openaimodel3d.py:274-296).  w: [Cout, 3*C] packed."""
    x, x1 = _as_pair(x)
    _check_act(x)
    b, t, hw, c = x.shape
    c1 = x1.shape[3] if x1 is not None else 0
    cout = w.shape[0]
    assert w.shape[1] == 3 * (c + c1)
    if out is None:
        out = torch.empty((b, t, hw, cout), device=x.device, dtype=BF16)
    box = plan_box((hw, t, b, 1))
    return _gemm_raw(
        a=(x, x1), a_ch=(c, c1), a_ch_total=(c, c1), a_size=(hw, t, b, 1),
        a_stride=((c, hw * c, t * hw * c, 0), (c1, hw * c1, t * hw * c1, 0) if x1 is not None else None), box=box, taps=_TAPS_T3, tap_ch_off=None, w=w,
        n_rows=cout, out=out, o_size=(hw, t, b, 1), o_stride=(cout, hw * cout, t * hw * cout, 0),
        n_out=cout, bias=bias, residual=residual, block_n=block_n, split_k=split_k,
        col_accum=(stats, (0, 1, t, 0)) if stats is not None else None)   # per-frame sums [B*T, Cout, 2]


def conv3x3_small_cin(x, w, bias, cout):
    """Direct conv for the 4-channel latent (input_blocks.0.0 / VAE conv_in). w: [Cout, 9*Cin] packed."""
    _check_act(x)
    n, h, wd, cin = x.shape
    out = torch.empty((n, h, wd, cout), device=x.device, dtype=BF16)
    _launch("conv3x3_small_cin", _FLOPS.pop("conv3x3_small_cin", 0), lib().t2v_conv3x3_small_cin, x.data_ptr(), w.data_ptr(), ptr(bias), out.data_ptr(), n, h, wd, cin,
                                      cout, stream_ptr())
    return out


# ----------------------------------------------------------------------------- normalisation
_GN_WS: dict = {}


def _gn_workspace(device, n):
    """Zeroed fp32 statistics scratch of the two-kernel GroupNorm, one per device (self-cleaning: t2v_groupnorm leaves it
    zeroed).  Outgrown buffers are kept alive: captured CUDA graphs may still point at them."""
    ws = _GN_WS.get(device)
    if ws is None or ws[-1].numel() < n:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("t2v_turbo_b200: GroupNorm workspace must be allocated before CUDA-graph capture (run a warm-up call)")
        _GN_WS.setdefault(device, []).append(torch.zeros(max(n, 1 << 16), device=device, dtype=torch.float32))
        ws = _GN_WS[device]
    return ws[-1]


def groupnorm(x, gamma, beta, *, rows_per_sample, eps, silu, groups=32, out=None, mode=0, chan_sums=None, chan_group=1):
    """GroupNorm(+SiLU) over token matrix x: [rows, C] (or a pair concatenated along C)."""
    x0, x1 = _as_pair(x)
    x0 = x0.reshape(-1, x0.shape[-1])
    rows, c0 = x0.shape
    c1 = 0
    if x1 is not None:
        x1 = x1.reshape(-1, x1.shape[-1])
        c1 = x1.shape[1]
    if out is None:
        out = torch.empty((rows, c0 + c1), device=x0.device, dtype=BF16)
    d = GroupNormDesc()
    d.x[0] = x0.data_ptr()
    d.x[1] = ptr(x1)
    d.ch[0], d.ch[1] = c0, c1
    d.x_row_stride[0] = x0.stride(0)
    d.x_row_stride[1] = x1.stride(0) if x1 is not None else 0
    d.out = out.data_ptr()
    d.out_row_stride = out.stride(0)
    d.gamma = gamma.data_ptr()
    d.beta = beta.data_ptr()
    d.rows = rows
    d.rows_per_sample = rows_per_sample
    d.groups = groups
    d.eps = eps
    d.silu = 1 if silu else 0
    ws = _gn_workspace(x0.device, 2 * groups * (rows // rows_per_sample) + 1)
    d.workspace = ws.data_ptr()
    d.mode = mode   # 0 auto, 1 two-kernel path, 2 single-kernel cluster path (tests)
    if chan_sums is not None:   # per-channel sums from the producing GEMMs (one per source): statistics pass skipped
        cs0, cs1 = _as_pair(chan_sums)
        d.chan_sums[0] = cs0.data_ptr()
        d.chan_sums[1] = ptr(cs1)
        d.chan_group = chan_group
    if _PROF is not None:
        _TAG["groupnorm"] = f"rows={rows} C={c0 + c1} rps={rows_per_sample} sums={int(chan_sums is not None)}"
    _launch("groupnorm", _FLOPS.pop("groupnorm", 0), lib().t2v_groupnorm, C.byref(d), stream_ptr())
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    rows, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    d = LayerNormDesc()
    d.x, d.x_row_stride = x.data_ptr(), x.stride(0)
    d.out, d.out_row_stride = out.data_ptr(), out.stride(0)
    d.gamma, d.beta = gamma.data_ptr(), beta.data_ptr()
    d.rows, d.channels, d.eps = rows, c, eps
    _launch("layernorm", _FLOPS.pop("layernorm", 0), lib().t2v_layernorm, C.byref(d), stream_ptr())
    return out


# ----------------------------------------------------------------------------- attention
def layernorm_stats(x, eps=1e-5, out=None):
    """Per-row (rstd, -rstd*mean) of x [rows, C] -> fp32 [rows, 2]; consumed by linear(..., ln=...)."""
    rows, c = x.shape
    if out is None:
        out = torch.empty((rows, 2), device=x.device, dtype=torch.float32)
    d = LayerNormDesc()
    d.x, d.x_row_stride = x.data_ptr(), x.stride(0)
    d.rows, d.channels, d.eps = rows, c, eps
    _launch("layernorm", _FLOPS.pop("layernorm", 0), lib().t2v_layernorm_stats, C.byref(d), out.data_ptr(), stream_ptr())
    return out


def fold_layernorm(w, bias, gamma, beta):
    """Fold LayerNorm's affine into the Linear that consumes it (load-time, fp32 math):
    returns (w' bf16 [N,K] = w*gamma, bias' fp32 [N] = w @ beta + bias, col_sum fp32 [N] = sum_k float(w'[n,k]))."""
    w32 = w.detach().float()
    wp = (w32 * gamma.float()[None, :]).to(BF16).contiguous()
    bp = w32 @ beta.float()
    if bias is not None:
        bp = bp + bias.float()
    return wp, bp.contiguous(), wp.float().sum(1).contiguous()


def attention(q, k, v, *, heads, scale, kv_batch_div=1, out=None, causal=False, lse2=None):
    """q: [Bq, Lq, H*64] view, k/v: [Bk, Lk, H*64] views (last dim contiguous; token/batch strides free).
    Returns o: [Bq, Lq, H*64].  lse2: optional fp32 [Bq, H, Lq] receiving the log2-domain row log-sum-exp (kept for
    attention_bwd)."""
    bq, lq, inner = q.shape
    bk, lk, _ = k.shape
    assert inner == heads * 64 and q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    assert bq == bk * kv_batch_div
    if out is None:
        out = torch.empty((bq, lq, inner), device=q.device, dtype=BF16)
    d = AttnDesc()
    d.q, d.k, d.v, d.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    d.batch, d.heads, d.len_q, d.len_k = bq, heads, lq, lk
    d.q_stride_b, d.q_stride_t, d.q_stride_h = q.stride(0), q.stride(1), 64
    d.k_stride_b, d.k_stride_t, d.k_stride_h = k.stride(0), k.stride(1), 64
    d.v_stride_b, d.v_stride_t, d.v_stride_h = v.stride(0), v.stride(1), 64
    d.o_stride_b, d.o_stride_t, d.o_stride_h = out.stride(0), out.stride(1), 64
    d.kv_batch_div = kv_batch_div
    d.scale = scale
    d.causal = 1 if causal else 0
    if lse2 is not None:
        assert lse2.dtype == torch.float32 and lse2.is_contiguous() and tuple(lse2.shape) == (bq, heads, lq)
        d.lse2 = lse2.data_ptr()
    _FLOPS["attn_fwd"] = 4 * bq * heads * lq * lk * 64
    if _PROF is not None:
        _TAG["attn_fwd"] = f"B={bq} H={heads} Lq={lq} Lk={lk}"
    _launch("attn_fwd", _FLOPS.pop("attn_fwd", 0), lib().t2v_attn_fwd, C.byref(d), stream_ptr())
    return out


def attention_temporal(q, k, v, *, b, t, hw, heads, scale, out=None, probs=None):
    """Temporal self-attention over token matrices laid out [(b t hw), H*64]: each of the b*hw pixels
    attends over its t frames (token stride hw*row_stride).  Replaces the '(b hw) t c' regrouping.
    probs: optional [(b*hw*heads), t, t] tensor (bf16 / fp16 / fp32) receiving the attention probabilities in the
    reference's "(b h) i j" layout (attention.py:124-126)."""
    rows, inner = q.shape
    assert rows == b * t * hw and inner == heads * 64
    if out is None:
        out = torch.empty((rows, inner), device=q.device, dtype=BF16)
    d = ShortAttnDesc()
    d.q, d.k, d.v, d.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    d.n_seq_outer, d.n_seq_inner, d.heads, d.len = b, hw, heads, t
    for name, ten in (("q", q), ("k", k), ("v", v), ("o", out)):
        rs = ten.stride(0)
        setattr(d, f"{name}_stride_outer", t * hw * rs)
        setattr(d, f"{name}_stride_inner", rs)
        setattr(d, f"{name}_stride_t", hw * rs)
        setattr(d, f"{name}_stride_h", 64)
    d.scale = scale
    if probs is not None:
        assert probs.is_cuda and probs.is_contiguous() and tuple(probs.shape) == (b * hw * heads, t, t)
        d.probs, d.probs_dtype = probs.data_ptr(), _lib.DTYPE_CODE[probs.dtype]
    _FLOPS["attn_short_fwd"] = 4 * b * hw * heads * t * t * 64
    if _PROF is not None:
        _TAG["attn_short_fwd"] = f"b={b} hw={hw} H={heads} t={t}"
    _launch("attn_short_fwd", _FLOPS.pop("attn_short_fwd", 0), lib().t2v_attn_short_fwd, C.byref(d), stream_ptr())
    return out


# ----------------------------------------------------------------------------- student backward (non-GEMM layers)
def _rows2d(x, name):
    assert x.is_cuda and x.dtype == BF16 and x.dim() == 2 and x.stride(1) == 1, f"{name}: CUDA bf16 [rows, C] with contiguous channels"
    return x


def groupnorm_bwd(x, dy, gamma, beta, *, rows_per_sample, eps, silu, groups=32, dx_add=None, out=None, keep_ws=None):
    """Adjoint of groupnorm (+SiLU) w.r.t. x.  x / dy: bf16 [rows, C] (row strides free); dx_add: optional second gradient
    path summed into the result.  keep_ws: a list that receives the statistics workspace (for groupnorm_affine_grad)."""
    _rows2d(x, "x"); _rows2d(dy, "dy")
    rows, c = x.shape
    if out is None:
        out = torch.empty((rows, c), device=x.device, dtype=BF16)
    n_samples = rows // rows_per_sample
    ws = torch.zeros((n_samples * groups * 4,), device=x.device, dtype=torch.float32)
    d = _lib.GroupNormBwdDesc()
    d.x, d.x_row_stride = x.data_ptr(), x.stride(0)
    d.dy, d.dy_row_stride = dy.data_ptr(), dy.stride(0)
    if dx_add is not None:
        _rows2d(dx_add, "dx_add")
        d.dx_add, d.dx_add_row_stride = dx_add.data_ptr(), dx_add.stride(0)
    d.dx, d.dx_row_stride = out.data_ptr(), out.stride(0)
    d.gamma, d.beta = gamma.data_ptr(), beta.data_ptr()
    d.rows, d.rows_per_sample = rows, rows_per_sample
    d.channels, d.groups, d.eps, d.silu = c, groups, eps, 1 if silu else 0
    d.workspace = ws.data_ptr()
    _launch("groupnorm_bwd", 0, lib().t2v_groupnorm_bwd, C.byref(d), stream_ptr())
    if keep_ws is not None:
        keep_ws.append(ws)
    return out


def layernorm_bwd(x, dy, gamma, eps=1e-5, *, dx_add=None, out=None):
    _rows2d(x, "x"); _rows2d(dy, "dy")
    rows, c = x.shape
    if out is None:
        out = torch.empty((rows, c), device=x.device, dtype=BF16)
    if dx_add is not None:
        _rows2d(dx_add, "dx_add")
    _launch("layernorm_bwd", 0, lib().t2v_layernorm_bwd, x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), _lib.ptr(dx_add),
            dx_add.stride(0) if dx_add is not None else 0, out.data_ptr(), out.stride(0), gamma.data_ptr(), rows, c, eps, stream_ptr())
    return out


def geglu(pre, dout=None, out=None):
    """pre: bf16 [rows, 2*inner] = [a | gate].  dout None: a * gelu(gate) -> [rows, inner]; else d(pre) -> [rows, 2*inner]."""
    _rows2d(pre, "pre")
    rows, two = pre.shape
    inner = two // 2
    if out is None:
        out = torch.empty((rows, inner if dout is None else two), device=pre.device, dtype=BF16)
    if dout is not None:
        _rows2d(dout, "dout")
    _launch("geglu", 0, lib().t2v_geglu, pre.data_ptr(), pre.stride(0), _lib.ptr(dout), dout.stride(0) if dout is not None else 0,
            out.data_ptr(), out.stride(0), rows, inner, stream_ptr())
    return out


def _ew2d(op, a, b, out):
    _rows2d(a, "a")
    rows, c = a.shape
    if out is None:
        out = torch.empty((rows, c), device=a.device, dtype=BF16)
    if b is not None:
        _rows2d(b, "b")
        assert b.shape == a.shape
    _launch("ew2d", 0, lib().t2v_ew2d, op, a.data_ptr(), a.stride(0), _lib.ptr(b), b.stride(0) if b is not None else 0, out.data_ptr(),
            out.stride(0), rows, c, stream_ptr())
    return out


def add(a, b, out=None):
    """a + b on bf16 [rows, C] views (gradient accumulation where an activation feeds two consumers)."""
    return _ew2d(0, a, b, out)


def silu(a, out=None):
    return _ew2d(1, a, None, out)


def silu_bwd(pre, dy, out=None):
    return _ew2d(2, pre, dy, out)


def colsum_samples(x, rows_per_sample, out=None):
    """fp32 [rows / rows_per_sample, C]: per-sample column sums of bf16 [rows, C] (accumulated into `out` when given)."""
    _rows2d(x, "x")
    rows, c = x.shape
    if out is None:
        out = torch.zeros((rows // rows_per_sample, c), device=x.device, dtype=torch.float32)
    _launch("colsum_samples", 0, lib().t2v_colsum_samples, x.data_ptr(), x.stride(0), out.data_ptr(), rows, rows_per_sample, c, stream_ptr())
    return out


def resample2x(x, mode):
    """x: bf16 [n, h, w, C] contiguous.  mode 'sub': [n, h/2, w/2, C] = x[:, ::2, ::2]; 'stuff': [n, 2h, 2w, C] zero stuffing
    (adjoint of 'sub'); 'pool': [n, h/2, w/2, C] 2x2 block sums (adjoint of the nearest 2x upsampling)."""
    assert x.is_cuda and x.dtype == BF16 and x.dim() == 4 and x.is_contiguous()
    n, h, w, c = x.shape
    code = {"sub": 0, "stuff": 1, "pool": 2}[mode]
    ho, wo = (2 * h, 2 * w) if mode == "stuff" else (h // 2, w // 2)
    out = torch.empty((n, ho, wo, c), device=x.device, dtype=BF16)
    _launch("resample2x", 0, lib().t2v_resample2x, code, x.data_ptr(), out.data_ptr(), n, ho, wo, c, stream_ptr())
    return out


def attention_bwd(q, k, v, o, d_o, lse2, *, heads, scale, kv_batch_div=1, need_dq=True, need_dkv=True):
    """Adjoint of attention(): returns (dq, dk, dv) shaped like q / k / v (contiguous).  o, lse2 from the forward."""
    bq, lq, inner = q.shape
    bk, lk, _ = k.shape
    assert inner == heads * 64 and bq == bk * kv_batch_div
    for ten in (q, k, v, o, d_o):
        assert ten.is_cuda and ten.dtype == BF16 and ten.stride(2) == 1
    delta = torch.empty((bq, heads, lq), device=q.device, dtype=torch.float32)
    _launch("attn_delta", 0, lib().t2v_attn_delta, o.data_ptr(), o.stride(0), o.stride(1), 64, d_o.data_ptr(), d_o.stride(0), d_o.stride(1), 64,
            delta.data_ptr(), bq, heads, lq, stream_ptr())
    dq = torch.empty((bq, lq, inner), device=q.device, dtype=BF16) if need_dq else None
    dk = torch.empty((bk, lk, inner), device=q.device, dtype=BF16) if need_dkv else None
    dv = torch.empty((bk, lk, inner), device=q.device, dtype=BF16) if need_dkv else None
    d = _lib.AttnBwdDesc()
    d.q, d.k, d.v, d.d_o = q.data_ptr(), k.data_ptr(), v.data_ptr(), d_o.data_ptr()
    d.lse2, d.delta = lse2.data_ptr(), delta.data_ptr()
    d.dq, d.dk, d.dv = _lib.ptr(dq), _lib.ptr(dk), _lib.ptr(dv)
    d.batch, d.heads, d.len_q, d.len_k = bq, heads, lq, lk
    for name, ten in (("q", q), ("k", k), ("v", v), ("do", d_o), ("dq", dq), ("dk", dk), ("dv", dv)):
        if ten is not None:
            setattr(d, f"{name}_stride_b", ten.stride(0)); setattr(d, f"{name}_stride_t", ten.stride(1)); setattr(d, f"{name}_stride_h", 64)
    d.kv_batch_div, d.scale = kv_batch_div, scale
    flops = (6 * need_dq + 8 * need_dkv) * bq * heads * lq * lk * 64
    _launch("attn_bwd", flops, lib().t2v_attn_bwd, C.byref(d), stream_ptr())
    return dq, dk, dv


def attention_temporal_bwd(q, k, v, d_o, *, b, t, hw, heads, scale):
    """Adjoint of attention_temporal(): q / k / v / d_o are [(b t hw), H*64] views (row strides free); returns contiguous
    (dq, dk, dv)."""
    rows, inner = q.shape
    assert rows == b * t * hw and inner == heads * 64
    outs = [torch.empty((rows, inner), device=q.device, dtype=BF16) for _ in range(3)]
    d = _lib.ShortAttnBwdDesc()
    f = d.fwd
    f.q, f.k, f.v, f.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), d_o.data_ptr()
    f.n_seq_outer, f.n_seq_inner, f.heads, f.len = b, hw, heads, t
    for name, ten in (("q", q), ("k", k), ("v", v), ("o", d_o)):
        rs = ten.stride(0)
        setattr(f, f"{name}_stride_outer", t * hw * rs)
        setattr(f, f"{name}_stride_inner", rs)
        setattr(f, f"{name}_stride_t", hw * rs)
        setattr(f, f"{name}_stride_h", 64)
    f.scale = scale
    # dq / dk / dv are written with the strides of q / k / v: the (contiguous) outputs need contiguous q / k / v
    assert q.stride(0) == k.stride(0) == v.stride(0) == inner, "attention_temporal_bwd: q / k / v must be contiguous [rows, H*64]"
    d.d_o = d_o.data_ptr()
    d.dq, d.dk, d.dv = (o.data_ptr() for o in outs)
    _launch("attn_short_bwd", 10 * b * hw * heads * t * t * 64, lib().t2v_attn_short_bwd, C.byref(d), stream_ptr())
    return tuple(outs)


# ----------------------------------------------------------------------------- small ops
def small_linear(x, w, bias=None, *, add=None, silu_in=False, silu_out=False, round_bf16=True, out=None):
    """fp32 rows x [M,K] times bf16 w [N,K] (embedding MLPs, M = batch)."""
    assert x.dtype == torch.float32 and x.stride(1) == 1 and w.dtype == BF16 and w.is_contiguous()
    m, k = x.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((m, n), device=x.device, dtype=torch.float32)
    d = SmallLinearDesc()
    d.x, d.x_row_stride = x.data_ptr(), x.stride(0)
    d.w, d.bias = w.data_ptr(), ptr(bias)
    d.add, d.add_row_stride = ptr(add), (add.stride(0) if add is not None else 0)
    d.out, d.out_row_stride = out.data_ptr(), out.stride(0)
    d.m, d.n, d.k = m, n, k
    d.silu_in, d.silu_out, d.round_bf16 = int(silu_in), int(silu_out), int(round_bf16)
    _launch("small_linear", _FLOPS.pop("small_linear", 0), lib().t2v_small_linear, C.byref(d), stream_ptr())
    return out


def sinusoidal_embedding(t, freqs, *, sin_first=False, round_bf16=True):
    assert t.dtype == torch.float32 and freqs.dtype == torch.float32
    m, half = t.shape[0], freqs.shape[0]
    out = torch.empty((m, 2 * half), device=t.device, dtype=torch.float32)
    _launch("sinusoidal_embedding", _FLOPS.pop("sinusoidal_embedding", 0), lib().t2v_sinusoidal_embedding, t.data_ptr(), freqs.data_ptr(), out.data_ptr(), m, half,
                                         int(sin_first), int(round_bf16), stream_ptr())
    return out


def bcthw_to_frames(x, scale=1.0):
    b, c, t, h, w = x.shape
    x = x.contiguous()
    out = torch.empty((b * t, h, w, c), device=x.device, dtype=BF16)
    _launch("bcthw_to_frames", _FLOPS.pop("bcthw_to_frames", 0), lib().t2v_bcthw_to_frames, x.data_ptr(), _lib.DTYPE_CODE[x.dtype], out.data_ptr(), b, c, t, h, w,
                                    scale, stream_ptr())
    return out


def bcthw_to_frames_pad(x, c_pad, scale=1.0):
    """[B, C, T, H, W] -> bf16 [B*T, H, W, c_pad] with zero channels appended (c_pad <= 64)."""
    b, c, t, h, w = x.shape
    x = x.contiguous()
    out = torch.empty((b * t, h, w, c_pad), device=x.device, dtype=BF16)
    _launch("bcthw_to_frames_pad", 0, lib().t2v_bcthw_to_frames_pad, x.data_ptr(), _lib.DTYPE_CODE[x.dtype], out.data_ptr(), b, c, c_pad, t, h, w,
            scale, stream_ptr())
    return out


def frames_to_bcthw(x, b, c, dtype):
    n, h, w, cp = x.shape
    t = n // b
    out = torch.empty((b, c, t, h, w), device=x.device, dtype=dtype)
    _launch("frames_to_bcthw", _FLOPS.pop("frames_to_bcthw", 0), lib().t2v_frames_to_bcthw, x.data_ptr(), cp, out.data_ptr(), _lib.DTYPE_CODE[dtype], b, c, t, h, w,
                                    stream_ptr())
    return out


def upsample_nearest2x(x):
    _check_act(x)
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=BF16)
    _launch("upsample_nearest2x", _FLOPS.pop("upsample_nearest2x", 0), lib().t2v_upsample_nearest2x, x.data_ptr(), out.data_ptr(), n, h, w, c, stream_ptr())
    return out


def concat_channels(a, b):
    ca, cb = a.shape[-1], b.shape[-1]
    rows = a.numel() // ca
    out = torch.empty((*a.shape[:-1], ca + cb), device=a.device, dtype=BF16)
    _launch("concat_channels", _FLOPS.pop("concat_channels", 0), lib().t2v_concat_channels, a.data_ptr(), ca, b.data_ptr(), cb, out.data_ptr(), rows, stream_ptr())
    return out


def attention_temporal_probs_bwd(q, k, d_probs, *, b, t, hw, heads, scale):
    """Adjoint of attention_temporal's `probs` export w.r.t. q / k: q, k [(b t hw), H*64] views (row strides free), d_probs fp32
    [(b*hw*heads), t, t] -> contiguous (dq, dk).  t <= 16."""
    rows, inner = q.shape
    assert rows == b * t * hw and inner == heads * 64 and k.shape == q.shape
    assert q.is_cuda and q.dtype == BF16 and k.dtype == BF16 and q.stride(1) == 1 and k.stride(1) == 1
    assert d_probs.dtype == torch.float32 and d_probs.is_contiguous() and tuple(d_probs.shape) == (b * hw * heads, t, t)
    dq = torch.empty((rows, inner), device=q.device, dtype=BF16)
    dk = torch.empty((rows, inner), device=q.device, dtype=BF16)
    _launch("attn_short_probs_bwd", 0, lib().t2v_attn_short_probs_bwd, q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), d_probs.data_ptr(),
            dq.data_ptr(), dk.data_ptr(), b, hw, heads, t, float(scale), stream_ptr())
    return dq, dk


def softmax_bwd_rows_(dp, p, scale=1.0):
    """In place on dp: dS = scale * P * (dP - rowsum(dP * P)) for P = softmax(scale * S) (adjoint of softmax_rows_); bf16 [.., cols]
    contiguous."""
    assert dp.is_cuda and dp.dtype == BF16 and p.dtype == BF16 and dp.is_contiguous() and p.is_contiguous() and dp.shape == p.shape
    cols = dp.shape[-1]
    rows = dp.numel() // cols
    _launch("softmax_bwd_rows", 0, lib().t2v_softmax_bwd_rows, dp.data_ptr(), cols, p.data_ptr(), cols, rows, cols, float(scale), stream_ptr())
    return dp


def bcthw_to_frames_mix(z, scale, mix, bias):
    """[B, C, T, H, W] (any float dtype) -> bf16 frames [B*T, H, W, C_out] with out[o] = sum_c mix[o, c] * scale * z[c] + bias[o]:
    `1 / scale_factor * z` + post_quant_conv 1x1 (ddpm3d.py:669, autoencoder.py:111).  mix [C_out, C] / bias [C_out] fp32."""
    b, c, t, hh, ww = z.shape
    z = z.contiguous()
    assert mix.dtype == torch.float32 and bias.dtype == torch.float32 and mix.is_contiguous() and tuple(mix.shape) == (bias.numel(), c)
    out = torch.empty((b * t, hh, ww, mix.shape[0]), device=z.device, dtype=BF16)
    _launch("bcthw_to_frames_mix", 0, lib().t2v_bcthw_to_frames_mix, z.data_ptr(), _lib.DTYPE_CODE[z.dtype], out.data_ptr(), b, c, t, hh, ww,
            float(scale), mix.data_ptr(), bias.data_ptr(), stream_ptr())
    return out


def softmax_rows_(x, scale=1.0):
    cols = x.shape[-1]
    rows = x.numel() // cols
    _launch("softmax_rows", _FLOPS.pop("softmax_rows", 0), lib().t2v_softmax_rows, x.data_ptr(), rows, cols, cols, scale, stream_ptr())
    return x


def lcm_step(x, eps, noise, *, inv_sqrt_alpha_t, sqrt_beta_t, c_skip, c_out, sqrt_alpha_prev, sqrt_beta_prev):
    assert x.is_contiguous() and eps.is_contiguous() and x.dtype == eps.dtype
    prev = torch.empty_like(x)
    den = torch.empty_like(x)
    _launch("lcm_step", _FLOPS.pop("lcm_step", 0), lib().t2v_lcm_step, x.data_ptr(), eps.data_ptr(), ptr(noise), prev.data_ptr(), den.data_ptr(),
                             x.numel(), _lib.DTYPE_CODE[x.dtype], inv_sqrt_alpha_t, sqrt_beta_t, c_skip, c_out,
                             sqrt_alpha_prev, sqrt_beta_prev, stream_ptr())
    return prev, den


def embedding_gather(table, ids, pos=None):
    """bf16 [n, W] = table[ids] (+ pos[i % ctx]): token + positional embedding (condition.py:262-263)."""
    assert table.is_cuda and table.is_contiguous() and ids.dtype == torch.int64 and ids.is_cuda
    ids = ids.contiguous().view(-1)
    n, width = ids.numel(), table.shape[1]
    out = torch.empty((n, width), device=table.device, dtype=BF16)
    ctx = pos.shape[0] if pos is not None else 1
    if pos is not None:
        assert pos.dtype == table.dtype and pos.is_contiguous() and pos.shape[1] == width
    _launch("embedding_gather", 0, lib().t2v_embedding_gather, table.data_ptr(), ptr(pos), _lib.DTYPE_CODE[table.dtype], ids.data_ptr(),
            out.data_ptr(), n, width, ctx, table.shape[0], stream_ptr())
    return out


def video_to_uint8(video):
    """[B, 3, T, H, W] in [-1, 1] -> uint8 [B, T, H, W, 3] (app.py:90-94: clamp, (v + 1) / 2 * 255, truncate, channels last)."""
    assert video.is_cuda and video.is_contiguous() and video.dim() == 5 and video.shape[1] == 3
    b, _, t, h, w = video.shape
    out = torch.empty((b, t, h, w, 3), device=video.device, dtype=torch.uint8)
    _launch("video_to_uint8", 0, lib().t2v_video_to_uint8, video.data_ptr(), _lib.DTYPE_CODE[video.dtype], out.data_ptr(), b, t, h, w,
            stream_ptr())
    return out


def scale_add_rows(x, a, y=None, b=None):
    """out[r] = rnd(rnd(a[r] * x[r]) + rnd(b[r] * y[r])) with per-row (= per-sample) fp32 scalars a, b on the device;
    x / y: contiguous tensors of one dtype whose leading dim indexes the rows."""
    assert x.is_cuda and x.is_contiguous() and a.dtype == torch.float32 and a.numel() == x.shape[0]
    rows = x.shape[0]
    out = torch.empty_like(x)
    if y is not None:
        assert y.shape == x.shape and y.dtype == x.dtype and y.is_contiguous() and b.dtype == torch.float32 and b.numel() == rows
    _launch("scale_add_rows", 0, lib().t2v_scale_add_rows, x.data_ptr(), ptr(y), a.data_ptr(), ptr(b), out.data_ptr(), rows,
            x.numel() // rows, _lib.DTYPE_CODE[x.dtype], stream_ptr())
    return out


def gaussian_sample(moments, noise, *, b, t, zc, scale, dtype):
    """KL posterior sample / mode (distributions.py:24-42) from fp32 channels-last moments [b*t, h, w, 2*zc]."""
    n, h, w, c2 = moments.shape
    assert n == b * t and c2 == 2 * zc and moments.dtype == torch.float32 and moments.is_contiguous()
    out = torch.empty((b, zc, t, h, w), device=moments.device, dtype=dtype)
    if noise is not None:
        noise = noise.to(device=moments.device, dtype=torch.float32).contiguous()
        assert noise.shape == (n, zc, h, w)
    _launch("gaussian_sample", 0, lib().t2v_gaussian_sample, moments.data_ptr(), ptr(noise), out.data_ptr(),
            _lib.DTYPE_CODE[dtype], b, t, h, w, zc, float(scale), stream_ptr())
    return out


# ----------------------------------------------------------------------------- training (LoRA) ops
def wgrad(a, b, out, *, taps=None, out_strides, alpha=1.0, a_grid=None):
    """out[j, c, tap] += alpha * sum_points a[point + off(tap), c] * b[point, j] (t2v_wgrad).
    a: channels-last bf16 [x4.., x1, C] (any leading point dims, up to 4) or [M, C]; b: bf16 [same points.., r];
    out: fp32 tensor (a slice of the gradient arena) addressed by out_strides = (j_stride, c_stride, tap_stride)."""
    _check_act(a, "a")
    _check_act(b, "b")
    assert out.dtype == torch.float32 and out.is_cuda
    pts = tuple(a.shape[:-1])
    assert tuple(b.shape[:-1]) == pts and len(pts) <= 4
    c, r = a.shape[-1], b.shape[-1]
    if math.prod(pts) < 16:   # fewer points than one MMA k-step (embedding layers, M = batch): zero rows add nothing
        assert len(pts) == 1 and taps is None
        a = torch.cat([a, a.new_zeros(16 - pts[0], c)])
        b = torch.cat([b, b.new_zeros(16 - pts[0], r)])
        pts = (16,)
    grid = tuple(reversed(pts)) + (1,) * (4 - len(pts))       # x1 fastest
    d = _lib.WgradDesc()
    d.a, d.a_ch, d.b, d.b_cols = a.data_ptr(), c, b.data_ptr(), r
    astr, bstr, acc = [], [], 1
    for g in grid:
        astr.append(acc * c)
        bstr.append(acc * r)
        acc *= g
    _fill(d.a_size, grid)
    _fill(d.o_size, grid)
    _fill(d.a_stride, astr)
    _fill(d.b_stride, bstr)
    box = plan_box(grid)
    if math.prod(box) % 16:
        box = plan_box(grid, fixed=(16 if grid[0] % 16 == 0 else None, None, None, None))
    assert math.prod(box) % 16 == 0, f"wgrad: no 16-row tile box for point grid {grid}"
    _fill(d.box, box)
    taps = taps if taps is not None else [(0, 0, 0, 0)]
    d.n_taps = len(taps)
    for t, off in enumerate(taps):
        _fill(d.tap_off[t], off)
    d.out = out.data_ptr()
    d.out_j_stride, d.out_c_stride, d.out_tap_stride = (int(v) for v in out_strides)
    d.alpha = alpha
    _launch("wgrad", 2 * math.prod(grid) * c * r * len(taps), lib().t2v_wgrad, C.byref(d), stream_ptr())
    return out


def wgrad_wide(a, b, out, *, taps=None, out_strides, alpha=1.0):
    """ops.wgrad for a B operand of ANY width (the base-weight gradients of the full fine-tune step: b = dy with Cout columns):
    out[j, c, tap] += alpha * sum_points a[point + off(tap), c] * b[point, j], j < b.shape[-1], as ceil(Cout / 64) t2v_wgrad launches
    over 64-column slices of b read IN PLACE (the B tensor map's row stride is the full width; nothing is copied)."""
    _check_act(a, "a")
    _check_act(b, "b")
    assert out.dtype == torch.float32 and out.is_cuda
    pts = tuple(a.shape[:-1])
    assert tuple(b.shape[:-1]) == pts and len(pts) <= 4
    c, n = a.shape[-1], b.shape[-1]
    assert n % 8 == 0, n
    if math.prod(pts) < 16:   # fewer points than one MMA k-step (embedding layers, M = batch): zero rows add nothing
        assert len(pts) == 1 and taps is None
        a = torch.cat([a, a.new_zeros(16 - pts[0], c)])
        b = torch.cat([b, b.new_zeros(16 - pts[0], n)])
        pts = (16,)
    grid = tuple(reversed(pts)) + (1,) * (4 - len(pts))       # x1 fastest
    astr, bstr, acc = [], [], 1
    for g in grid:
        astr.append(acc * c)
        bstr.append(acc * n)
        acc *= g
    box = plan_box(grid)
    if math.prod(box) % 16:
        box = plan_box(grid, fixed=(16 if grid[0] % 16 == 0 else None, None, None, None))
    assert math.prod(box) % 16 == 0, f"wgrad: no 16-row tile box for point grid {grid}"
    taps = taps if taps is not None else [(0, 0, 0, 0)]
    j_stride, c_stride, tap_stride = (int(v) for v in out_strides)
    for j0 in range(0, n, 64):
        cols = min(64, n - j0)
        d = _lib.WgradDesc()
        d.a, d.a_ch, d.b, d.b_cols = a.data_ptr(), c, b.data_ptr() + 2 * j0, cols
        _fill(d.a_size, grid)
        _fill(d.o_size, grid)
        _fill(d.a_stride, astr)
        _fill(d.b_stride, bstr)
        _fill(d.box, box)
        d.n_taps = len(taps)
        for t, off in enumerate(taps):
            _fill(d.tap_off[t], off)
        d.out = out.data_ptr() + 4 * j0 * j_stride
        d.out_j_stride, d.out_c_stride, d.out_tap_stride = j_stride, c_stride, tap_stride
        d.alpha = alpha
        _launch("wgrad", 2 * math.prod(grid) * c * cols * len(taps), lib().t2v_wgrad, C.byref(d), stream_ptr())
    return out


def groupnorm_affine_grad(x, dy, gamma, beta, stats_ws, dgamma, dbeta, *, rows_per_sample, eps, silu, groups=32):
    """dgamma / dbeta (fp32 [C], accumulated into) of groupnorm(+SiLU); stats_ws: the workspace ops.groupnorm_bwd(..., keep_ws=[])
    filled for the same x (full fine-tune step only)."""
    _rows2d(x, "x"); _rows2d(dy, "dy")
    rows, c = x.shape
    assert dgamma.dtype == torch.float32 and dbeta.dtype == torch.float32 and dgamma.numel() == c and dbeta.numel() == c
    assert dgamma.is_contiguous() and dbeta.is_contiguous()
    _launch("groupnorm_affine_grad", 0, lib().t2v_groupnorm_affine_grad, x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0),
            gamma.data_ptr(), beta.data_ptr(), stats_ws.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), rows, rows_per_sample, c, groups,
            eps, 1 if silu else 0, stream_ptr())


def layernorm_affine_grad(x, dy, dgamma, dbeta, eps=1e-5):
    """dgamma / dbeta (fp32 [C], accumulated into) of layernorm (full fine-tune step only)."""
    _rows2d(x, "x"); _rows2d(dy, "dy")
    rows, c = x.shape
    assert dgamma.dtype == torch.float32 and dbeta.dtype == torch.float32 and dgamma.numel() == c and dbeta.numel() == c
    assert dgamma.is_contiguous() and dbeta.is_contiguous()
    _launch("layernorm_affine_grad", 0, lib().t2v_layernorm_affine_grad, x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0),
            dgamma.data_ptr(), dbeta.data_ptr(), rows, c, eps, stream_ptr())


def ema_update(target, src, rate):
    """target = target * rate + src * (1 - rate) over flat fp32 arenas (update_ema, utils/common_utils.py:308-319)."""
    assert target.is_cuda and src.is_cuda and target.dtype == torch.float32 and src.dtype == torch.float32
    assert target.is_contiguous() and src.is_contiguous() and target.numel() == src.numel()
    _launch("ema_update", 0, lib().t2v_ema_update, target.data_ptr(), src.data_ptr(), target.numel(), float(rate), stream_ptr())
    return target


def scale_mask(x, scale, mask=None, out=None):
    """out = x * scale * mask (bf16; mask uint8 keep-mask or None): the LoRA branch's dropout(...) * scale and its adjoint."""
    assert x.is_cuda and x.dtype == BF16 and x.is_contiguous() and x.numel() % 8 == 0
    if out is None:
        out = torch.empty_like(x)
    if mask is not None:
        assert mask.dtype == torch.uint8 and mask.is_contiguous() and mask.numel() == x.numel()
    _launch("scale_mask", 0, lib().t2v_scale_mask, x.data_ptr(), ptr(mask), out.data_ptr(), x.numel(), float(scale), stream_ptr())
    return out


# ---- in-kernel dropout (training): one device-resident 64-bit seed per device, advanced between steps on the device so that
# captured CUDA graphs draw fresh masks on every replay; every call site takes the next call id (baked into a captured graph).
_DROPOUT_SEED: dict = {}
_DROPOUT_CALLS = [0]


def dropout_seed(device, seed=None):
    """The device-resident seed tensor (int64[1]) of `device`; `seed` (re)initialises it (default: torch.initial_seed())."""
    key = torch.device(device)
    t = _DROPOUT_SEED.get(key)
    if t is None or seed is not None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("t2v_turbo_b200: the dropout seed must be created before CUDA-graph capture (run a warm-up call)")
        val = (torch.initial_seed() if seed is None else int(seed)) & 0x7FFFFFFFFFFFFFFF
        if t is None:
            t = torch.empty(1, device=key, dtype=torch.int64)
            _DROPOUT_SEED[key] = t
        t.fill_(val)
    return t


def dropout_advance(device):
    """Advance the device's dropout seed (one tiny kernel; capturable): call once per training step / forward."""
    dropout_seed(device).add_(0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF)


def dropout_scale(x, p, scale=1.0, out=None, addend=None):
    """-> (out, keep): out = x * scale / (1 - p) * keep (+ addend) with keep ~ Bernoulli(1 - p) drawn inside the kernel (uint8,
    kept for the adjoint `scale_mask(dy, scale / (1 - p), keep)`): nn.Dropout(p)(x) * scale — and the residual add that follows
    the layer — in one pass, no separate mask kernel."""
    assert x.is_cuda and x.dtype == BF16 and x.is_contiguous() and x.numel() % 8 == 0 and 0.0 < p < 1.0
    if addend is not None:
        assert addend.dtype == BF16 and addend.is_contiguous() and addend.numel() == x.numel()
    if out is None:
        out = torch.empty_like(x)
    keep = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    _DROPOUT_CALLS[0] = (_DROPOUT_CALLS[0] + 1) & 0xFFFFFFFF
    _launch("dropout_scale", 0, lib().t2v_dropout_scale, x.data_ptr(), ptr(addend), out.data_ptr(), keep.data_ptr(), x.numel(), float(1.0 - p),
            float(scale) / (1.0 - p), dropout_seed(x.device).data_ptr(), _DROPOUT_CALLS[0], stream_ptr())
    return out, keep


def adamw_step(param, grad, exp_avg, exp_avg_sq, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, step, grad_scale=1.0):
    n = param.numel()
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
    _launch("adamw_step", 0, lib().t2v_adamw_step, param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n,
            float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), float(grad_scale), stream_ptr())


def sum_squares(x, out=None):
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.zeros(1, device=x.device, dtype=torch.float32)
    _launch("sum_squares", 0, lib().t2v_sum_squares, x.data_ptr(), x.numel(), out.data_ptr(), stream_ptr())
    return out


def mse_loss_grad(a, b, *, want_grad=True, grad_scale=1.0):
    """(mean((a - b)^2) as a 1-element fp32 tensor, d loss / d a * grad_scale in a's dtype)."""
    assert a.is_cuda and a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    loss = torch.zeros(1, device=a.device, dtype=torch.float32)
    grad = torch.empty_like(a) if want_grad else None
    _launch("mse_loss_grad", 0, lib().t2v_mse_loss_grad, a.data_ptr(), b.data_ptr(), ptr(grad), loss.data_ptr(), a.numel(),
            _lib.DTYPE_CODE[a.dtype], float(grad_scale), stream_ptr())
    return loss, grad


def huber_loss_grad(a, b, huber_c=0.001, *, want_grad=True, grad_scale=1.0):
    """(mean(sqrt((a - b)^2 + c^2) - c) as a 1-element fp32 tensor, d loss / d a * grad_scale in a's dtype):
    utils/common_utils.py:302-304."""
    assert a.is_cuda and a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    loss = torch.zeros(1, device=a.device, dtype=torch.float32)
    grad = torch.empty_like(a) if want_grad else None
    _launch("huber_loss_grad", 0, lib().t2v_huber_loss_grad, a.data_ptr(), b.data_ptr(), ptr(grad), loss.data_ptr(), a.numel(),
            _lib.DTYPE_CODE[a.dtype], float(huber_c), float(grad_scale), stream_ptr())
    return loss, grad


# ----------------------------------------------------------------------------- weight packing
def pack_conv_weight(w):
    """torch conv weight [Cout, Cin, *k] -> bf16 [Cout, taps*Cin] (tap-major K)."""
    cout, cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    w = w.contiguous()
    out = torch.empty((cout, taps * cin), device=w.device, dtype=BF16)
    _launch("pack_conv_weight", _FLOPS.pop("pack_conv_weight", 0), lib().t2v_pack_conv_weight, w.data_ptr(), _lib.DTYPE_CODE[w.dtype], out.data_ptr(), cout, cin, taps,
                                     stream_ptr())
    return out


def pack_geglu(w, bias):
    """GEGLU proj weight [2*inner, K] (+bias) -> value/gate rows interleaved in blocks of 16."""
    two_inner, k = w.shape
    inner = two_inner // 2
    w = w.contiguous()
    out = torch.empty((two_inner, k), device=w.device, dtype=BF16)
    bout = torch.empty((two_inner,), device=w.device, dtype=torch.float32)
    bias = bias.contiguous()
    _launch("pack_geglu_rows", _FLOPS.pop("pack_geglu_rows", 0), lib().t2v_pack_geglu_rows, w.data_ptr(), _lib.DTYPE_CODE[w.dtype], out.data_ptr(), bias.data_ptr(),
                                    _lib.DTYPE_CODE[bias.dtype], bout.data_ptr(), inner, k, stream_ptr())
    return out, bout
