"""fp32-accumulate CHECK MODE of the denoiser (BASELINE north star: rel-L2 <= 1e-3 vs the reference's fp32 forward).

``UNet3DConditionModel(..., check_mode=True)`` runs the SAME host launch sequence (lavie_b200/unet.py) with this module
in place of ``lavie_b200.ops``.  Activations are split-bf16 triples ``[rows, 3C] = [hi | lo | hi]`` (class ``Triple``),
weights are (hi, lo) bf16 pairs packed along K as ``[Wh | Wh | Wl]``, so every Linear / 1x1 / 3x3 conv still runs on the
product's tcgen05 GEMM / im2col-TMA conv mainloop (``lavie_check_gemm`` / ``lavie_check_conv3x3``: fp32 accumulators out
through the workspace, epilogue in fp32).  Norms, attention cores and the small kernels use the fp32 twins in
csrc/check.cu.  Slower (3x the MMA work, SIMT attention) and only meant for parity attribution; frame sharding is not
available in this mode.  Nothing here computes with torch ops either: torch only splits the WEIGHTS once at packing.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _lib, ops
from ._lib import check
from .ops import _Launch, _epilogue, _ptr, _stream, BF16, F32

LAUNCHES_ATTR = "LAUNCHES"


class Triple:
    """A check-mode activation: ``t`` is bf16 [rows, >= 3*C] holding [hi | lo | hi] from column ``off``; ``C`` = width
    of one block; ``lo`` = column distance hi -> lo (C for a whole tensor, the parent's C for a column slice)."""

    def __init__(self, t: torch.Tensor, C: int, off: int = 0, lo: Optional[int] = None):
        self.t, self.C, self.off, self.lo = t, C, off, (C if lo is None else lo)

    @property
    def rows(self):
        return self.t.shape[0]

    @property
    def ld(self):
        return self.t.stride(0)

    @property
    def device(self):
        return self.t.device

    def ptr(self):
        return self.t.data_ptr() + 2 * self.off

    def full(self) -> bool:
        return self.off == 0 and self.lo == self.C

    def float(self) -> torch.Tensor:
        """hi + lo as fp32 [rows, C] (debug taps only)."""
        hi = self.t[:, self.off:self.off + self.C].float()
        lo = self.t[:, self.off + self.lo:self.off + self.lo + self.C].float()
        return hi + lo

    def row_slice(self, r0: int, r1: Optional[int] = None) -> "Triple":
        return Triple(self.t[r0:r1], self.C, self.off, self.lo)


class WPair:
    """A weight matrix as (hi, lo) bf16; the K-tripled layouts are built on first use and cached."""

    def __init__(self, w: torch.Tensor, device):
        w = w.to(device=device, dtype=F32)
        self.hi = w.to(BF16)
        self.lo = (w - self.hi.float()).to(BF16)
        self.shape = tuple(w.shape)
        self._cache: Dict[tuple, torch.Tensor] = {}

    def gemm_layout(self, k0: int) -> torch.Tensor:
        """[N, K] -> [N, 3K]: [Wh | Wh | Wl] per source (source 0 = the first k0 input features)."""
        key = ("g", k0)
        if key not in self._cache:
            parts = []
            K = self.hi.shape[1]
            for a, b in ((0, k0), (k0, K)):
                if b > a:
                    parts += [self.hi[:, a:b], self.hi[:, a:b], self.lo[:, a:b]]
            self._cache[key] = torch.cat(parts, dim=1).contiguous()
        return self._cache[key]

    def conv_layout(self) -> torch.Tensor:
        """[N, 9*C] in (kh, kw, c) order -> [N, 9 * 3C]: per tap [Wh | Wh | Wl]."""
        key = ("c",)
        if key not in self._cache:
            N, K = self.hi.shape
            hi, lo = self.hi.reshape(N, 9, K // 9), self.lo.reshape(N, 9, K // 9)
            self._cache[key] = torch.cat([hi, hi, lo], dim=2).reshape(N, 3 * K).contiguous()
        return self._cache[key]


_workspaces: Dict[tuple, torch.Tensor] = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _new(rows: int, C: int, device) -> Triple:
    return Triple(torch.empty((rows, 3 * C), dtype=BF16, device=device), C)


# ------------------------------------------------------------------------------------------------ layout helpers
def cols(x: Triple, a: int, b: int) -> Triple:
    """Columns [a, b) of a triple WITHOUT copying: hi at off + a, lo stays one parent-block away."""
    return Triple(x.t, b - a, x.off + a, x.lo)


def weight(w: torch.Tensor, device) -> WPair:
    return WPair(w, device)


def weight_small(w: torch.Tensor, device) -> torch.Tensor:
    return w.to(device=device, dtype=F32).contiguous()


def to_float(x: Triple) -> torch.Tensor:
    return x.float()


def text_input(text: torch.Tensor, ctx: int) -> Triple:
    """encoder_hidden_states fp32 [B, L, ctx] -> triple rows (the product path casts to bf16 instead)."""
    lib = _lib.load()
    x = text.to(dtype=F32).reshape(-1, ctx).contiguous()
    out = _new(x.shape[0], ctx, x.device)
    with _Launch("lavie_check_split3"):
        check(lib.lavie_check_split3(x.data_ptr(), x.shape[0], ctx, out.t.data_ptr(), _stream()), "lavie_check_split3")
    return out


# ------------------------------------------------------------------------------------------------ GEMM / conv
def gemm(a: Triple, w: WPair, *, a2: Optional[Triple] = None, bias=None, row_bias=None, rows_per_batch=1,
         residual: Optional[Triple] = None, geglu=False, out: Optional[Triple] = None, block_n: int = 0,
         stats: bool = False) -> Triple:
    lib = _lib.load()          # (stats: the bf16 path's fused GroupNorm statistics; check mode runs the fp32 pass)
    assert a.full() and (a2 is None or a2.full()), "GEMM inputs must be whole triples"
    M, k0 = a.rows, a.C
    k1 = a2.C if a2 is not None else 0
    assert w.shape[1] == k0 + k1, (w.shape, k0, k1)
    N = w.shape[0]
    n_out = N // 2 if geglu else N
    wt = w.gemm_layout(k0)
    if out is None:
        out = _new(M, n_out, a.device)
    assert out.full() and out.C == n_out and out.rows == M
    ep = _epilogue(bias, row_bias, rows_per_batch, None, geglu) if (bias is not None or row_bias is not None or geglu) \
        else None
    if residual is not None:
        assert residual.full() and residual.C == N
        if ep is None:
            ep = _lib.Epilogue()
            ep.rows_per_batch = 1
        ep.residual = residual.t.data_ptr()
        ep.ld_residual = residual.ld
    ws = _workspace(a.device, ((M + 255) // 256) * 256 * N * 4)
    with _Launch("gemm_bf16_tcgen05", 6.0 * M * N * (k0 + k1), 0.0, f"check gemm M={M} N={N} K={3 * (k0 + k1)}"):
        rc = lib.lavie_check_gemm(a.t.data_ptr(), a.ld, 3 * k0, a2.t.data_ptr() if a2 is not None else None,
                                  a2.ld if a2 is not None else 0, 3 * k1, wt.data_ptr(), out.t.data_ptr(), out.ld, M, N,
                                  ctypes.byref(ep) if ep is not None else None, ws.data_ptr(), ws.numel(), _stream())
    check(rc, "lavie_check_gemm")
    return out


def conv3x3(x: Triple, NF: int, H: int, W: int, w: WPair, *, stride: int = 1, bias=None, row_bias=None,
            rows_per_batch=1, residual: Optional[Triple] = None, out: Optional[Triple] = None, block_n: int = 0,
            stats: bool = False) -> Triple:
    lib = _lib.load()
    assert x.full() and x.ld == 3 * x.C and x.rows == NF * H * W, "conv3x3 needs a contiguous whole triple"
    C = x.C
    assert w.shape[1] == 9 * C
    N = w.shape[0]
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    M = NF * Ho * Wo
    wt = w.conv_layout()
    if out is None:
        out = _new(M, N, x.device)
    ep = _epilogue(bias, row_bias, rows_per_batch, None, False) if (bias is not None or row_bias is not None) else None
    if residual is not None:
        assert residual.full() and residual.C == N
        if ep is None:
            ep = _lib.Epilogue()
            ep.rows_per_batch = 1
        ep.residual = residual.t.data_ptr()
        ep.ld_residual = residual.ld
    ws = _workspace(x.device, ((M + 255) // 256) * 256 * N * 4)
    with _Launch("gemm_bf16_tcgen05", 6.0 * M * N * 9 * C, 0.0, f"check conv3x3 M={M} N={N} K={27 * C}"):
        rc = lib.lavie_check_conv3x3(x.t.data_ptr(), NF, H, W, 3 * C, stride, wt.data_ptr(), out.t.data_ptr(), out.ld, N,
                                     ctypes.byref(ep) if ep is not None else None, ws.data_ptr(), ws.numel(), _stream())
    check(rc, "lavie_check_conv3x3")
    return out


# ------------------------------------------------------------------------------------------------ norms
def groupnorm_scale_shift(x: Triple, samples: int, rows_per_sample: int, gamma, beta, eps: float, groups: int = 32,
                          x2: Optional[Triple] = None):
    lib = _lib.load()
    assert x.full() and (x2 is None or x2.full())
    c0, c1 = x.C, (x2.C if x2 is not None else 0)
    C = c0 + c1
    chunks = lib.lavie_groupnorm_chunks(samples, rows_per_sample)
    partial = torch.empty((samples, chunks, groups, 2), dtype=F32, device=x.device)
    ss = torch.empty((samples, C, 2), dtype=F32, device=x.device)
    with _Launch("lavie_groupnorm_stats"):
        check(lib.lavie_check_groupnorm_stats(x.t.data_ptr(), x.ld, c0, x2.t.data_ptr() if x2 is not None else None,
                                              x2.ld if x2 is not None else 0, c1, samples, rows_per_sample, groups,
                                              partial.data_ptr(), _stream()), "lavie_check_groupnorm_stats")
    with _Launch("lavie_groupnorm_finalize"):
        check(lib.lavie_groupnorm_finalize(partial.data_ptr(), samples, chunks, groups, C,
                                           rows_per_sample * (C // groups), gamma.data_ptr(), beta.data_ptr(), eps,
                                           ss.data_ptr(), _stream()), "lavie_groupnorm_finalize")
    return ss


def groupnorm_apply(x: Triple, scale_shift, samples: int, rows_per_sample: int, silu: bool,
                    x2: Optional[Triple] = None) -> Triple:
    lib = _lib.load()
    c0, c1 = x.C, (x2.C if x2 is not None else 0)
    out = _new(x.rows, c0 + c1, x.device)
    with _Launch("lavie_groupnorm_apply"):
        check(lib.lavie_check_groupnorm_apply(x.t.data_ptr(), x.ld, c0, x2.t.data_ptr() if x2 is not None else None,
                                              x2.ld if x2 is not None else 0, c1, samples, rows_per_sample,
                                              scale_shift.data_ptr(), 1 if silu else 0, out.t.data_ptr(), out.ld,
                                              _stream()), "lavie_check_groupnorm_apply")
    return out


def groupnorm(x, samples, rows_per_sample, gamma, beta, eps, silu, groups=32, x2=None) -> Triple:
    ss = groupnorm_scale_shift(x, samples, rows_per_sample, gamma, beta, eps, groups, x2)
    return groupnorm_apply(x, ss, samples, rows_per_sample, silu, x2)


def layernorm(x: Triple, gamma, beta, eps: float = 1e-5) -> Triple:
    lib = _lib.load()
    assert x.full()
    out = _new(x.rows, x.C, x.device)
    with _Launch("lavie_layernorm_bf16"):
        check(lib.lavie_check_layernorm(x.t.data_ptr(), x.ld, gamma.data_ptr(), beta.data_ptr(), eps, out.t.data_ptr(),
                                        out.ld, x.rows, x.C, _stream()), "lavie_check_layernorm")
    return out


# ------------------------------------------------------------------------------------------------ attention cores
def _attention(q: Triple, k: Triple, v: Triple, strides, batch, heads, Sq, Sk, d, pitch, kv_batch_div, sc_frames, scale,
               rope, bias, out: Triple):
    lib = _lib.load()
    (q_seq, q_batch), (kv_seq, kv_batch), (o_seq, o_batch) = strides
    assert k.lo == v.lo and k.ld == v.ld
    rot = rope.shape[1] if rope is not None else 0
    with _Launch("lavie_attention_bf16", 4.0 * batch * heads * Sq * Sk * (2 if sc_frames else 1) * d, 0.0,
                 f"check attn B={batch} Sq={Sq} Sk={Sk} d={d}"):
        check(lib.lavie_check_attention(q.ptr(), q_seq, q_batch, q.lo, k.ptr(), v.ptr(), kv_seq, kv_batch, k.lo,
                                        out.t.data_ptr(), o_seq, o_batch, out.C, batch, heads, Sq, Sk, d, pitch,
                                        kv_batch_div, sc_frames, scale, _ptr(rope), rot, _ptr(bias), _stream()),
              "lavie_check_attention")
    return out


def attention(q: Triple, k: Triple, v: Triple, batch: int, heads: int, Sq: int, Sk: int, d: int, head_pitch: int,
              kv_batch_div: int = 1, scale: Optional[float] = None, sparse_causal_frames: int = 0) -> Triple:
    out = _new(batch * Sq, heads * d, q.device)
    strides = ((q.ld, Sq * q.ld), (k.ld, Sk * k.ld), (out.ld, Sq * out.ld))
    return _attention(q, k, v, strides, batch, heads, Sq, Sk, d, head_pitch, kv_batch_div, sparse_causal_frames,
                      d ** -0.5 if scale is None else scale, None, None, out)


def temporal_attention(qkv: Triple, B: int, F: int, HW: int, heads: int, d: int, head_pitch: int, rope, bias) -> Triple:
    """the base model's TemporalAttention: q pre-scaled, RoPE on q and k, + rel-pos bias; frames read with stride HW."""
    hp = heads * head_pitch
    out = _new(qkv.rows, heads * d, qkv.device)
    q, k, v = cols(qkv, 0, hp), cols(qkv, hp, 2 * hp), cols(qkv, 2 * hp, 3 * hp)
    for b in range(B):
        sl = slice(b * F * HW, (b + 1) * F * HW)
        qb, kb, vb = (Triple(t.t[sl], t.C, t.off, t.lo) for t in (q, k, v))
        ob = Triple(out.t[sl], out.C)
        strides = ((HW * qkv.ld, qkv.ld), (HW * qkv.ld, qkv.ld), (HW * out.ld, out.ld))
        _attention(qb, kb, vb, strides, HW, heads, F, F, d, head_pitch, 1, 0, d ** -0.5, rope, bias, ob)
    return out


def frame_attention(qkv: Triple, B: int, F: int, HW: int, heads: int, d: int, head_pitch: int) -> Triple:
    """the interpolation model's plain attention over frames."""
    return temporal_attention(qkv, B, F, HW, heads, d, head_pitch, None, None)


# ------------------------------------------------------------------------------------------------ small kernels
timestep_embedding = ops.timestep_embedding


def linear_smallm(x, w: torch.Tensor, bias, silu_in=False, silu_out=False):
    lib = _lib.load()
    M, K = x.shape
    assert w.dtype == F32 and w.is_contiguous() and w.shape[1] == K
    N = w.shape[0]
    out = torch.empty((M, N), dtype=F32, device=x.device)
    with _Launch("lavie_linear_smallm"):
        check(lib.lavie_check_linear_smallm(x.data_ptr(), M, K, w.data_ptr(), _ptr(bias), out.data_ptr(), N, int(silu_in),
                                            int(silu_out), _stream()), "lavie_check_linear_smallm")
    return out


def conv_in(x, w, bias, input_scale=None) -> Triple:
    lib = _lib.load()
    B, Cin, Fr, H, W = x.shape
    Cout = w.shape[0]
    out = _new(B * Fr * H * W, Cout, x.device)
    with _Launch("lavie_conv_in"):
        check(lib.lavie_check_conv_in(x.data_ptr(), _ptr(input_scale), B, Cin, Fr, H, W, w.data_ptr(), bias.data_ptr(),
                                      Cout, out.t.data_ptr(), out.ld, _stream()), "lavie_check_conv_in")
    return out


def conv_out(x: Triple, scale_shift, B: int, Fr: int, H: int, W: int, w, bias):
    lib = _lib.load()
    Cout = w.shape[0]
    out = torch.empty((B, Cout, Fr, H, W), dtype=F32, device=x.device)
    with _Launch("lavie_conv_out"):
        check(lib.lavie_check_conv_out(x.t.data_ptr(), x.ld, scale_shift.data_ptr(), B, Fr, H, W, x.C, w.data_ptr(),
                                       bias.data_ptr(), Cout, out.data_ptr(), _stream()), "lavie_check_conv_out")
    return out


def upsample_nearest2x(x: Triple, NF: int, H: int, W: int) -> Triple:
    """nearest-neighbour gather: the triple is just 3C channels to the regular kernel."""
    y = ops.upsample_nearest2x(x.t, NF, H, W)
    return Triple(y, x.C)
