"""CPU oracle for the MXINT8 exponent-sign pruned-attention hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product path (``mx_quantization_b200``) never does.

It restates, stage by stage, what d9bjo0522/mx_quantization computes on this path
(citations are relative to the reference checkout):

  A1  bf16 pre-rounding            microxscaling/mx/elemwise_ops.py:201-216,243-277
  A2  MX block quantizer (int8)    microxscaling/mx/mx_ops.py:49-99,180-306
                                   microxscaling/mx/elemwise_ops.py:45-86,92-180
                                   microxscaling/mx/formats.py:89-91,116-117
  A3  predictor ctor               funcs/exponent_based_prediction.py:12-38
  A4  exponent_based_sign          microxscaling/examples/deit/exponent_based_prediction.py:135-161
  A5  pred_scores = ex_q @ ex_k^T  workloads/deit/scripts/main.py:118
  A6  top-k                        workloads/deit/scripts/main.py:123  (canonical rule below)
  A7  true scores (mx.matmul)      microxscaling/mx/matmul.py:32-100, main.py:101-102,124
  A8  softmax / scatter / P.V      workloads/deit/scripts/main.py:147-152
  f3  other rankings               partial_Q / partial_K (funcs/exponent_based_prediction.py:274-318,
                                   main.py:111-114) and top-k of the true scores (main.py:130)

Parity pinning: ``tests/golden/make_golden.py`` runs the *unmodified reference* (imported
from /root/reference in the authoring container) on seeded inputs and commits its outputs
as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them
and against the reference's own known-answer vectors
(microxscaling/mx/tests/test_corners_mx.py:61-124, funcs/test_scatter.py:156-174).

Integer stages (codes, exponents, sign words, predicted scores, top-k sets) are exact;
floating stages follow the reference's fp32 torch ops.

Canonical top-k: ``torch.topk`` leaves tie order unspecified (and ties are the norm for
these scores), so the contract is "descending score, ascending key index" == the first k
entries of a stable descending sort.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

BLOCK = 32
ZERO_BLOCK_EXP = -126  # floor(log2(FP32_MIN_NORMAL)), mx_ops.py:83-87


# --------------------------------------------------------------------------------------
# A1: bfloat16 pre-rounding ("round nearest" == half away from zero on the magnitude)
# --------------------------------------------------------------------------------------
def bf16_round_half_away(x: torch.Tensor) -> torch.Tensor:
    """elemwise_ops.py:201-216 with bits=9, exp_bits=8, round='nearest' (:64-65).

    Scaling by the private exponent, adding 0.5 to the magnitude and flooring is, on the
    fp32 bit pattern, "add half of the dropped field, truncate 16 bits".  Valid for
    normals and subnormals (bfloat_subnorms=True); values that would exceed the bf16
    max-normal are outside the contract (reference maps them to Inf).
    """
    b = x.contiguous().view(torch.int32)
    mag = ((b & 0x7FFFFFFF) + 0x8000) & ~0xFFFF
    return (mag | (b & -0x80000000)).view(torch.float32)


# --------------------------------------------------------------------------------------
# A2: shared exponent, as the reference's fp32 floor(log2(.)) really evaluates
# --------------------------------------------------------------------------------------
def _log2_bump_table() -> np.ndarray:
    """jmax[n+128]: largest j such that fp32 RN(log2(2^n * (1 - j*2^-24))) == n.

    mx_ops.py:93-97 computes floor(log2(max)) in fp32.  For a block maximum whose
    mantissa is within j<=jmax(n) ulps below the power of two 2^n the fp32 logarithm
    rounds *up* to n, so the reference's shared exponent is n, not n-1.  The half-ulp of
    a float just below n (n>0) or just above |n| (n<0) gives
    jmax = floor(ln2 * 2^(p-1)) with 2^(p-1) < n <= 2^p, resp. floor(ln2 * 2^p) with
    2^p <= |n| < 2^(p+1); tests/test_oracle_golden.py re-derives it against torch.log2.
    """
    tab = np.zeros(257, dtype=np.int32)
    for n in range(-127, 129):
        if n >= 3:
            p = math.ceil(math.log2(n))
            j = math.floor(math.log(2.0) * 2.0 ** (p - 1))
        elif n <= -2:
            p = math.floor(math.log2(-n))
            j = math.floor(math.log(2.0) * 2.0 ** p)
        else:
            j = 0
        tab[n + 128] = j
    return tab


LOG2_BUMP = _log2_bump_table()


def shared_exponent_from_absmax(amax: torch.Tensor) -> torch.Tensor:
    """floor(log2(amax)) exactly as the fp32 reference evaluates it (int32 result).

    zero -> -126 (mx_ops.py:95: log2(0 + 2^-126)); subnormal maxima are clamped to the
    scale_bits=8 floor of -127 (mx_ops.py:289-291) - bit patterns below 2^-126 are
    outside the parity contract (reference README: "undefined").
    """
    bits = amax.contiguous().view(torch.int32)
    E = (bits >> 23) & 0xFF
    m = bits & 0x7FFFFF
    e = E - 127
    bump_tab = torch.from_numpy(LOG2_BUMP)
    jmax = bump_tab[(e + 1 + 128).clamp(0, 256).to(torch.int64)]
    bump = ((0x800000 - m) <= jmax) & (E > 0)
    e = e + bump.to(torch.int32)
    e = torch.where(E == 0, torch.where(m == 0, ZERO_BLOCK_EXP, -127), e)
    if bool((e > 127).any()):
        raise ValueError("Inf/NaN or 2^128-scale input: outside the MXINT8 parity contract")
    return e.to(torch.int32)


def _blocks_last(x: torch.Tensor, block: int):
    """Zero-pad the last axis to a multiple of ``block`` and split it (mx_ops.py:121-161)."""
    d = x.shape[-1]
    nb = (d + block - 1) // block
    pad = nb * block - d
    if pad:
        x = torch.nn.functional.pad(x, (0, pad))
    return x.reshape(*x.shape[:-1], nb, block), d


def quantize_mxint8(x: torch.Tensor, block: int = BLOCK, bfloat: int = 32,
                    flush_subnorms: bool = False):
    """A1+A2 along the last axis.  Returns (codes int8 [..., d], exps int8 [..., nb]).

    dequantised value = code * 2^(exp - 6); code in [-127, 127] (sign-magnitude, -128
    never produced, formats.py:116-117); exp in [-127, 127].
    """
    if bfloat == 16:
        x = bf16_round_half_away(x)
    elif bfloat != 32:
        raise ValueError("only bfloat in {16, 32} is on the path")
    xb, d = _blocks_last(x.to(torch.float32), block)
    amax = xb.abs().amax(dim=-1)
    e = shared_exponent_from_absmax(amax)                       # [..., nb] int32
    if flush_subnorms:                                          # mx_ops.py:282-283
        xb = xb * (e > -127).to(xb.dtype).unsqueeze(-1)
    e = e.clamp(min=-127)
    scale = torch.ldexp(torch.ones_like(amax), e)               # 2^e, exact (subnormal at -127)
    t = (xb.abs() / scale.unsqueeze(-1)) * 64.0                 # exact power-of-two scaling
    r = torch.floor(t + 0.5)                                    # fp32 add: elemwise_ops.py:64-65
    mag = torch.clamp(r, max=127.0).to(torch.int32)             # clamp to max_norm 127/64
    neg = xb < 0
    codes = torch.where(neg, -mag, mag).to(torch.int8)
    codes = codes.reshape(*codes.shape[:-2], -1)[..., :d]
    return codes.contiguous(), e.to(torch.int8).contiguous()


def fake_quant_mxint4(x: torch.Tensor, block: int = BLOCK, bfloat: int = 32,
                      flush_subnorms: bool = False) -> torch.Tensor:
    """quantize_mx_op(quantize_elemwise_op(x), elem_format='int4', axes=[-1]) as fp32 - the operands of the
    reference's "MXINT4" (Sanger) predictor, funcs/exponent_based_prediction.py:179-199.

    int4: ebits 0, mbits 4, emax 0 (formats.py:86-88) -> the shared exponent is the MXINT8 one; the element
    is sign * min(7, floor(|x| / 2^e * 4 + 0.5)) / 4 (lshift by mbits-2 = 2, elemwise_ops.py:155; nearest
    :64-65; clamp to max_norm 7/4 :163-164): value = c4 * 2^(e-2), c4 in [-7, 7]."""
    if bfloat == 16:
        x = bf16_round_half_away(x)
    elif bfloat != 32:
        raise ValueError("only bfloat in {16, 32} is on the path")
    xb, d = _blocks_last(x.to(torch.float32), block)
    amax = xb.abs().amax(dim=-1)
    e = shared_exponent_from_absmax(amax)
    if flush_subnorms:
        xb = xb * (e > -127).to(xb.dtype).unsqueeze(-1)
    e = e.clamp(min=-127)
    scale = torch.ldexp(torch.ones_like(amax), e)
    t = (xb.abs() / scale.unsqueeze(-1)) * 4.0
    mag = torch.clamp(torch.floor(t + 0.5), max=7.0)
    val = torch.ldexp(torch.where(xb < 0, -mag, mag), (e - 2).unsqueeze(-1))
    return val.reshape(*val.shape[:-2], -1)[..., :d].contiguous()


def dequantize_mxint8(codes: torch.Tensor, exps: torch.Tensor, block: int = BLOCK) -> torch.Tensor:
    """fake-quant fp32 value c * 2^(e-6) (what quantize_mx_op returns)."""
    d = codes.shape[-1]
    e = exps.to(torch.int32).repeat_interleave(block, dim=-1)[..., :d]
    return torch.ldexp(codes.to(torch.float32), e - 6)


def fake_quant_mxint8(x: torch.Tensor, axis: int = -1, block: int = BLOCK, bfloat: int = 32,
                      flush_subnorms: bool = False) -> torch.Tensor:
    """quantize_mx_op(quantize_elemwise_op(x), elem_format='int8', axes=[axis]) as fp32."""
    xt = x.movedim(axis, -1)
    c, e = quantize_mxint8(xt, block, bfloat, flush_subnorms)
    return dequantize_mxint8(c, e, block).movedim(-1, axis)


# --------------------------------------------------------------------------------------
# A3/A4: exponent-sign approximation
# --------------------------------------------------------------------------------------
def predictor_exponents(codes: torch.Tensor, exps: torch.Tensor, block: int = BLOCK) -> torch.Tensor:
    """funcs/exponent_based_prediction.py:33-36: floor(log2(max |MX block|)).

    max|MX| = cmax * 2^(e-6); cmax in [64,127] for every non-zero in-contract block, so the
    result is e; all-zero block -> -126.  (cmax = 127 is 2^-7 below a power of two - far
    outside the log2 round-up window.)
    """
    cb, _ = _blocks_last(codes.to(torch.int32), block)
    cmax = cb.abs().amax(dim=-1)
    lg = torch.floor(torch.log2(cmax.clamp(min=1).to(torch.float64))).to(torch.int32)
    e = exps.to(torch.int32) - 6 + lg
    return torch.where(cmax == 0, ZERO_BLOCK_EXP, e).to(torch.int32)


def sign_words(codes: torch.Tensor, block: int = BLOCK) -> torch.Tensor:
    """bit d of word b = (code[b*32+d] < 0); padding bits are 0.  int64 tensor [..., nb]."""
    assert block == 32
    cb, _ = _blocks_last((codes < 0).to(torch.int64), block)
    w = (cb << torch.arange(block, dtype=torch.int64)).sum(dim=-1)
    return w


def exponent_based_sign(codes: torch.Tensor, exps: torch.Tensor, block: int = BLOCK) -> torch.Tensor:
    """approx = (MX < 0 ? -1 : +1) * 2^e_block; zero and -0 count as +1 (:55-56,:85-89)."""
    d = codes.shape[-1]
    ep = predictor_exponents(codes, exps, block).repeat_interleave(block, dim=-1)[..., :d]
    s = torch.where(codes < 0, -1.0, 1.0).to(torch.float32)
    return torch.ldexp(s, ep)


# --------------------------------------------------------------------------------------
# A5: predicted scores
# --------------------------------------------------------------------------------------
def block_widths(d: int, block: int = BLOCK):
    nb = (d + block - 1) // block
    return [min(block, d - b * block) for b in range(nb)]


def pred_scores_integer(qc, qe, kc, ke, block: int = BLOCK) -> torch.Tensor:
    """score[i,j] = sum_b 2^(eq[i,b]+ek[j,b]) * (n_b - 2*popc(sq[i,b]^sk[j,b])), fp32.

    Same quantity funcs/test_scatter.py:156-174 spells out.  Accumulated over blocks in
    ascending b in fp32 - exact whenever the terms of one (i,j) pair fit a 24-bit window
    (SURVEY 8a edge note ii), where it equals the reference's fp32 matmul bit for bit.
    """
    d = qc.shape[-1]
    widths = block_widths(d, block)
    sq, sk = sign_words(qc, block), sign_words(kc, block)
    eq, ek = predictor_exponents(qc, qe, block), predictor_exponents(kc, ke, block)
    out = None
    for b, n_b in enumerate(widths):
        x = sq[..., :, None, b] ^ sk[..., None, :, b]
        # popcount of a 32-bit value held in int64
        p = torch.zeros_like(x)
        for s in range(32):
            p += (x >> s) & 1
        cnt = (n_b - 2 * p).to(torch.float32)
        w = torch.ldexp(torch.ones_like(cnt), eq[..., :, None, b] + ek[..., None, :, b])
        term = cnt * w
        out = term if out is None else out + term
    return out


def pred_window_ok(qe_pred: torch.Tensor, ke_pred: torch.Tensor, widths) -> torch.Tensor:
    """True where all block terms of pair (i,j) fit one 24-bit window, i.e. where the fp32
    matmul of the reference is exact in ANY summation order (SURVEY 8a edge note ii).
    Outside it (an all-zero block meeting exact cancellation elsewhere) the reference's
    value is decided by BLAS's summation order: parity unpinned."""
    s = qe_pred[..., :, None, :] + ke_pred[..., None, :, :]          # [..., Nq, Nk, nb]
    top = s + torch.tensor([math.ceil(math.log2(w)) + 1 for w in widths])
    return (top.amax(-1) - s.amin(-1)) <= 24


def pred_scores_matmul(qc, qe, kc, ke, block: int = BLOCK) -> torch.Tensor:
    """The reference's own formulation: fp32 matmul of the +-2^e tensors (main.py:118)."""
    aq = exponent_based_sign(qc, qe, block)
    ak = exponent_based_sign(kc, ke, block)
    return aq @ ak.transpose(-2, -1)


def two_step_leading_ones(codes: torch.Tensor, exps: torch.Tensor, block: int = BLOCK) -> torch.Tensor:
    """The reference's EXION emulation, funcs/exponent_based_prediction.py:96-127, on the MXINT8 codes:
        raw   = MX / 2^e_shared * 64                    == the int8 code c
        sign  = torch.sign(MX)                          (0 for a zero code)
        f1    = floor(log2 |c|)
        temp  = where(c - 2^f1 < 0, 0, c - 2^f1)        on the SIGNED code: only c > 0 keeps a remainder
        f2    = floor(log2 temp)                        (temp = 0 -> log2(2^-126): 2^f2 vanishes in fp32)
        value = sign * e_shared * (2^f1 + 2^f2) / 64    e_shared is the exponent's VALUE, as the reference writes it
    """
    d = codes.shape[-1]
    c = codes.to(torch.int64)
    a = c.abs()
    f1 = torch.floor(torch.log2(a.clamp(min=1).to(torch.float64))).to(torch.int64)
    rest = torch.where(c > 0, a - (1 << f1), torch.zeros_like(a))
    f2 = torch.floor(torch.log2(rest.clamp(min=1).to(torch.float64))).to(torch.int64)
    m = (1 << f1) + torch.where(rest > 0, 1 << f2, torch.zeros_like(rest))
    ep = predictor_exponents(codes, exps, block).repeat_interleave(block, dim=-1)[..., :d].to(torch.int64)
    val = torch.sign(c) * ep * m
    return (val.to(torch.float32) / 64.0)                         # |val| < 2^24: exact


def sign_leading_ones(codes: torch.Tensor, exps: torch.Tensor, block: int = BLOCK) -> torch.Tensor:
    """"true_ex": microxscaling/examples/deit/exponent_based_prediction.py:163-178 -
    where(MX < 0, -1, +1) * 2^floor(log2 |MX|) per element; a zero element (either sign) gives +1.0
    (get_true_exponents :98-110 leaves the exponent of zeros at 0).  |MX| = |c| * 2^(e-6)."""
    d = codes.shape[-1]
    c = codes.to(torch.int64)
    f1 = torch.floor(torch.log2(c.abs().clamp(min=1).to(torch.float64))).to(torch.int32)
    e = exps.to(torch.int32).repeat_interleave(block, dim=-1)[..., :d]
    t = torch.where(c == 0, torch.zeros_like(f1), e - 6 + f1)
    return torch.ldexp(torch.where(c < 0, -1.0, 1.0).to(torch.float32), t)


def elsa_scores(qc, qe, kc, ke, projection: torch.Tensor, block: int = BLOCK) -> torch.Tensor:
    """funcs/elsa_approximation.py:103-145 on the MXINT8 values: k-bit sign hashes of MX @ P^T (>= 0 -> 1),
    Hamming distance from the +-1 dot product, angle pi/k * h - 0.127 clamped at 0, and the key norms
    `unsqueeze(-1)`-broadcast over the QUERY rows exactly as the reference writes it (:140; needs Nq == Nk)."""
    mq, mk = dequantize_mxint8(qc, qe, block), dequantize_mxint8(kc, ke, block)
    k_bits = mk.shape[-1]
    P = projection.to(torch.float32)
    s_q = (torch.matmul(mq, P.T) >= 0).to(torch.int8).mul(2).sub(1).float()
    s_k = (torch.matmul(mk, P.T) >= 0).to(torch.int8).mul(2).sub(1).float()
    dots = torch.einsum('bhnk,bhmk->bhnm', s_q, s_k)
    hamming = 0.5 * (k_bits - dots)
    est = (torch.pi / k_bits) * hamming.float()
    corrected = torch.clamp(est - 0.127, min=0)
    return torch.norm(mk, dim=-1).unsqueeze(-1) * torch.cos(corrected)


def pred_scores_mode(qc, qe, kc, ke, pred_mode: str = "ex_pred", block: int = BLOCK) -> torch.Tensor:
    """`pred_scores = ex_quant_q @ ex_quant_k^T` (workloads/deit/scripts/main.py:118) for the predictor
    variants that reuse the MXINT8 codes:
      ex_pred    both sides +-2^e                             funcs/exponent_based_prediction.py:44-94
      partial_Q  Q = MXINT8 value c * 2^(e-6), K = +-2^e      funcs/exponent_based_prediction.py:300-318
      partial_K  Q = +-2^e, K = MXINT8 value                  funcs/exponent_based_prediction.py:274-298
      (MXINT4: both sides MXINT4 values - takes the fp32 inputs, see pruned_attention)
    fp32 matmul, as the reference computes it."""
    if pred_mode == "true_ex":
        return sign_leading_ones(qc, qe, block) @ sign_leading_ones(kc, ke, block).transpose(-2, -1)
    if pred_mode == "two_step_leading_ones":                       # funcs/exponent_based_prediction.py:96-177
        return two_step_leading_ones(qc, qe, block) @ two_step_leading_ones(kc, ke, block).transpose(-2, -1)
    if pred_mode not in ("ex_pred", "partial_Q", "partial_K"):
        raise ValueError(f"pred_mode {pred_mode!r}")
    aq = dequantize_mxint8(qc, qe, block) if pred_mode == "partial_Q" else exponent_based_sign(qc, qe, block)
    ak = dequantize_mxint8(kc, ke, block) if pred_mode == "partial_K" else exponent_based_sign(kc, ke, block)
    return aq @ ak.transpose(-2, -1)


# --------------------------------------------------------------------------------------
# A6: canonical top-k
# --------------------------------------------------------------------------------------
def canonical_topk(scores: torch.Tensor, k: int) -> torch.Tensor:
    """First k of a stable descending sort: ties broken by the lower key index."""
    order = torch.sort(scores, dim=-1, descending=True, stable=True).indices
    return order[..., :k].contiguous()


def idx_to_mask_words(idx: torch.Tensor, n_keys: int) -> torch.Tensor:
    """Row bitmask, int64 holding uint32 words: bit (j%32) of word j//32 set iff key j kept."""
    nw = (n_keys + 31) // 32
    dense = torch.zeros(*idx.shape[:-1], nw * 32, dtype=torch.int64)
    dense.scatter_(-1, idx, 1)
    dense = dense.reshape(*idx.shape[:-1], nw, 32)
    return (dense << torch.arange(32, dtype=torch.int64)).sum(dim=-1)


def mask_words_to_dense(words: torch.Tensor, n_keys: int) -> torch.Tensor:
    bits = (words.to(torch.int64).unsqueeze(-1) >> torch.arange(32, dtype=torch.int64)) & 1
    return bits.reshape(*words.shape[:-1], -1)[..., :n_keys].to(torch.bool)


# --------------------------------------------------------------------------------------
# A7/A8: exact MXINT8 attention over the kept keys
# --------------------------------------------------------------------------------------
def _elemwise_out(x: torch.Tensor, bfloat: int) -> torch.Tensor:
    return bf16_round_half_away(x) if bfloat == 16 else x


def true_scores_dense(q, k, scale: float, bfloat: int = 32, flush: bool = False) -> torch.Tensor:
    """mx.matmul(q, k^T, 'aa') * scale  (matmul.py:45-91; main.py:101-102)."""
    qd = fake_quant_mxint8(q, -1, BLOCK, bfloat, flush)
    kd = fake_quant_mxint8(k, -1, BLOCK, bfloat, flush)   # axes=[-2] of k^T == head_dim
    s = _elemwise_out(qd @ kd.transpose(-2, -1), bfloat)
    return s * scale


def sparse_softmax_pv(vals: torch.Tensor, idx: torch.Tensor, v: torch.Tensor, n_keys: int,
                      bfloat: int = 32, flush: bool = False) -> torch.Tensor:
    """main.py:147-152 with the dense N x N matrices kept (oracle sizes are small).

    P is scattered to a dense row, MX-quantised along keys (blocks of 32 *original* key
    positions, row zero-padded to a multiple of 32); V is MX-quantised along tokens
    (matmul.py:76-83, axes=[-2]); fp32 matmul; output elementwise rounding.
    """
    p = torch.softmax(vals, dim=-1)
    attn = torch.zeros(*vals.shape[:-1], n_keys, dtype=torch.float32)
    attn.scatter_(-1, idx, p)
    pq = fake_quant_mxint8(_elemwise_out(attn, bfloat), -1, BLOCK, 32, flush)
    vq = fake_quant_mxint8(_elemwise_out(v, bfloat), -2, BLOCK, 32, flush)
    return _elemwise_out(pq @ vq, bfloat)


def default_scale(hd: int) -> float:
    """fp32(hd ** -0.5): 0.125 for 64, 0.11785113 for 72 (SURVEY 8a edge note iii)."""
    return float(np.float32(hd ** -0.5))


def pruned_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, top_k: int,
                     scale: Optional[float] = None, bfloat: int = 32, flush: bool = False,
                     idx: Optional[torch.Tensor] = None, use_torch_topk: bool = False,
                     integer_scores: bool = False,
                     key_bias: Optional[torch.Tensor] = None,
                     pred_mode: str = "ex_pred",
                     orthogonal_matrix: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """The whole path on CPU (q, k, v: fp32 (B,H,N,hd)); returns every intermediate.

    key_bias: additive attention bias broadcastable to (B,1,1,Nk), as PixArt's cross-attention
    adds to BOTH the true and the predicted scores before top-k
    (workloads/PixArt/models/MX_transformer_block.py:794-803, 821-822).

    idx: use this kept-key set instead of the predictor's (to compare outputs on the same
    set).  use_torch_topk: leave torch.topk in place as the reference does (timing only).
    integer_scores: rank on pred_scores_integer (block-exact sums; what the CUDA kernel
    computes) instead of the fp32 matmul - identical wherever pred_window_ok holds.
    pred_mode: "ex_pred" | "partial_Q" | "partial_K" | "two_step_leading_ones" | "true_ex" (pred_scores_mode), "MXINT4", or "exact" - the reference's
    approx_flag=False branch, `torch.topk(true_scores, k)` (main.py:130).
    """
    q, k, v = (t.to(torch.float32) for t in (q, k, v))
    n_keys = k.shape[-2]
    scale = default_scale(q.shape[-1]) if scale is None else float(np.float32(scale))
    qc, qe = quantize_mxint8(q, BLOCK, bfloat, flush)
    kc, ke = quantize_mxint8(k, BLOCK, bfloat, flush)
    res: Dict[str, torch.Tensor] = {"q_codes": qc, "q_exps": qe, "k_codes": kc, "k_exps": ke}
    qd, kd = dequantize_mxint8(qc, qe), dequantize_mxint8(kc, ke)
    true = _elemwise_out(qd @ kd.transpose(-2, -1), bfloat) * scale
    if key_bias is not None:
        true = true + key_bias.to(torch.float32)                     # true_scores += attn_bias, :803
    if idx is None:
        if pred_mode == "exact":
            pred = true
        elif pred_mode == "ELSA":                                    # funcs/elsa_approximation.py, main.py:119-121
            pred = elsa_scores(qc, qe, kc, ke, orthogonal_matrix)
        elif pred_mode == "MXINT4":                                  # funcs/exponent_based_prediction.py:179-199
            pred = fake_quant_mxint4(q, BLOCK, bfloat, flush) @ fake_quant_mxint4(k, BLOCK, bfloat, flush).transpose(-2, -1)
        elif pred_mode != "ex_pred":
            pred = pred_scores_mode(qc, qe, kc, ke, pred_mode)
        else:
            pred = (pred_scores_integer if integer_scores else pred_scores_matmul)(qc, qe, kc, ke)
        if key_bias is not None and pred_mode != "exact":
            pred = pred + key_bias.to(torch.float32)                 # fp32 add, :822
        res["pred_scores"] = pred
        if use_torch_topk:
            idx = torch.topk(pred, top_k, dim=-1, largest=True, sorted=True).indices
        else:
            idx = canonical_topk(pred, top_k)
    res["idx"] = idx
    vals = true.gather(-1, idx)
    res["true_vals"] = vals
    res["out"] = sparse_softmax_pv(vals, idx, v, n_keys, bfloat, flush)
    return res


def mx_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
              bfloat: int = 32, flush: bool = False) -> torch.Tensor:
    """Forward of the reference's mx.Linear for MXINT8 activations and weights
    (microxscaling/mx/linear.py:20-103): element-wise A1 on input / weight / bias (:29-47), MX
    quantization of both along in_features (:57-72, axes=[-1]), fp32 F.linear (:83), A1 on the
    output (:85-87), + bias and A1 again (:89-93)."""
    x, weight = x.to(torch.float32), weight.to(torch.float32)
    qx = fake_quant_mxint8(x, axis=-1, bfloat=bfloat, flush_subnorms=flush)
    qw = fake_quant_mxint8(weight, axis=-1, bfloat=bfloat, flush_subnorms=flush)
    out = _elemwise_out(torch.nn.functional.linear(qx, qw), bfloat)
    if bias is not None:
        out = _elemwise_out(out + _elemwise_out(bias.to(torch.float32), bfloat), bfloat)
    return out
