"""``exponent_approximation`` - same constructor and method signatures as the reference class
(funcs/exponent_based_prediction.py:11-340 and the example copy
microxscaling/examples/deit/exponent_based_prediction.py:135-178), computed from libmxprune's outputs.

The reference class fake-quantises Q and K in its constructor and every method materialises two dense fp32
tensors so that the caller can do ``@`` / ``torch.topk`` itself (workloads/deit/scripts/main.py:107-123).
That formulation is kept for callers (and comparison scripts) that still want the tensors; the PRODUCT path is
:func:`mx_quantization_b200.pruned_attention` / :meth:`predict_topk`, which never materialises them - every
ranking below is also available there as ``pred_mode=...`` (mxprune_predict_wide.cuh).

All arithmetic of the dense returns is exact: the MXINT8 codes and block exponents come from the CUDA quantizer
(``mxp_quantize_mxint8``, bit-exact with ``quantize_mx_op``), and each method is integer arithmetic on them
followed by one exact scaling by a power of two (tests/test_gpu_parity.py::test_predictor_class_methods compares
them with the reference-generated fixtures bit for bit).
"""
import torch

from . import ops
from .specs import resolve_specs


def _expand_blocks(t: torch.Tensor, d: int) -> torch.Tensor:
    """(…, nb) per-block values -> (…, d) per-element (32-wide blocks, last one partial)."""
    return t.repeat_interleave(32, dim=-1)[..., :d]


class exponent_approximation:  # noqa: N801  (name fixed by the reference API)
    def __init__(self, Q: torch.Tensor, K: torch.Tensor, mx_specs):
        self._sp = resolve_specs(mx_specs)            # reject unsupported configurations up front
        self.mx_specs = mx_specs
        self.Q = Q
        self.K = K
        self._q = None
        self._k = None

    # -- compact integer form ---------------------------------------------------------
    def _quant(self):
        if self._q is None:
            self._q = ops.quantize_mxint8(self.Q, self.mx_specs, with_signs=True)
            self._k = ops.quantize_mxint8(self.K, self.mx_specs, with_signs=True)
        return self._q, self._k

    @property
    def codes(self):
        (qc, _, _), (kc, _, _) = self._quant()
        return qc, kc

    @property
    def exps(self):
        (_, qe, _), (_, ke, _) = self._quant()
        return qe, ke

    @property
    def signbits(self):
        (_, _, qs), (_, _, ks) = self._quant()
        return qs, ks

    # -- dense building blocks (device tensors, exact) -----------------------------------
    def _pred_exps(self, codes, exps):
        """shared_exponent_* of the reference ctor (:35-36): floor(log2(max |MX block|)) = the A2 exponent, -126 for
        an all-zero (or flushed) block."""
        d = codes.shape[-1]
        pad = (-d) % 32
        c = torch.nn.functional.pad(codes, (0, pad)).reshape(*codes.shape[:-1], -1, 32)
        dead = (c == 0).all(-1)
        return torch.where(dead, torch.full_like(exps, -126), exps)

    def _mx(self, codes, exps):
        """MX_Q / MX_K of the ctor (:18-31): the fake-quantised values c * 2^(e-6)."""
        e = _expand_blocks(exps.to(torch.int32), codes.shape[-1])
        return torch.ldexp(codes.to(torch.float32), e - 6)

    def _exp_sign(self, codes, exps):
        e = _expand_blocks(self._pred_exps(codes, exps).to(torch.int32), codes.shape[-1])
        return torch.ldexp(torch.where(codes < 0, -1.0, 1.0).to(torch.float32), e)

    # -- reference API ------------------------------------------------------------------
    def exponent_based_sign(self):
        """(approx_Q, approx_K): fp32 tensors shaped like Q, K with values (code < 0 ? -1 : +1) * 2^e_block
        (funcs/exponent_based_prediction.py:44-94; working body: the example copy :135-161)."""
        return ops.exp_sign_approx(self.Q, self.mx_specs), ops.exp_sign_approx(self.K, self.mx_specs)

    def partial_K(self):  # noqa: N802
        """Q = exponent-sign approximation, K = MXINT8 values (funcs/exponent_based_prediction.py:274-298)."""
        (qc, qe, _), (kc, ke, _) = self._quant()
        return self._exp_sign(qc, qe), self._mx(kc, ke)

    def partial_Q(self):  # noqa: N802
        """Q = MXINT8 values, K = exponent-sign approximation (funcs/exponent_based_prediction.py:300-318)."""
        (qc, qe, _), (kc, ke, _) = self._quant()
        return self._mx(qc, qe), self._exp_sign(kc, ke)

    def MXINT4(self):  # noqa: N802
        """Sanger: both sides fake-quantised to MXINT4 (funcs/exponent_based_prediction.py:179-199).  int4 has emax 0
        like int8 (formats.py:86-88), so the block exponent is the MXINT8 one; the element is
        sign * min(7, floor(|x| * 2^(2-e) + 0.5)) * 2^(e-2), x after the bf16 pre-rounding when bfloat == 16."""
        outs = []
        for x, (_, exps, _) in zip((self.Q, self.K), self._quant()):
            xr = x.to(torch.float32)
            if self._sp.bfloat_bits == 16:      # A1: round-half-away on the magnitude (elemwise_ops.py:64-65, 201-216)
                bits = xr.contiguous().view(torch.int32)
                xr = ((bits + 0x8000) & -65536).view(torch.float32)
            e = _expand_blocks(exps.to(torch.int32), x.shape[-1])
            if self._sp.flush:
                xr = torch.where(e > -127, xr, torch.zeros_like(xr))
            t = torch.ldexp(xr.abs(), 2 - e)                     # |x| / 2^e * 4, exact scaling
            mag = torch.clamp(torch.floor(t + 0.5), max=7.0)
            outs.append(torch.ldexp(torch.where(xr < 0, -mag, mag), e - 2))
        return outs[0], outs[1]

    def two_step_leading_ones(self):
        """EXION emulation, literally as the reference writes it (funcs/exponent_based_prediction.py:96-127):
        value = sign(c) * e_shared * (2^f1 + 2^f2) / 64 with c the int8 code, f1 = floor(log2 |c|), the second term
        only for POSITIVE codes with a remainder (`temp` is formed from the signed code), e_shared the exponent's value."""
        outs = []
        for codes, exps, _ in self._quant():
            c = codes.to(torch.int32)
            a = c.abs()
            f1 = 31 - _clz32(a.clamp(min=1))
            rest = torch.where(c > 0, a - (1 << f1), torch.zeros_like(a))
            f2 = 31 - _clz32(rest.clamp(min=1))
            m = (1 << f1) + torch.where(rest > 0, 1 << f2, torch.zeros_like(rest))
            ep = _expand_blocks(self._pred_exps(codes, exps).to(torch.int32), codes.shape[-1])
            outs.append((torch.sign(c) * ep * m).to(torch.float32) / 64.0)      # |.| < 2^24: exact
        return outs[0], outs[1]

    def exponent_based_sign_leading_ones(self):
        """"true_ex" (PixArt, MX_transformer_block.py:663-664; only the example copy of the predictor file has it,
        microxscaling/examples/deit/exponent_based_prediction.py:163-178): where(MX < 0, -1, +1) * 2^floor(log2 |MX|) per
        element, a zero element giving +1.0."""
        outs = []
        for codes, exps, _ in self._quant():
            c = codes.to(torch.int32)
            f1 = 31 - _clz32(c.abs().clamp(min=1))
            e = _expand_blocks(exps.to(torch.int32), codes.shape[-1])
            t = torch.where(c == 0, torch.zeros_like(f1), e - 6 + f1)
            outs.append(torch.ldexp(torch.where(c < 0, -1.0, 1.0).to(torch.float32), t))
        return outs[0], outs[1]

    # -- the fused replacement ------------------------------------------------------------
    def predict_topk(self, k: int, return_idx: bool = False, pred_mode: str = "ex_pred"):
        """Fused replacement for one of the methods above + ``@`` + ``torch.topk`` (no dense tensors)."""
        return ops.predict_topk(self.Q, self.K, self.mx_specs, k, return_idx=return_idx, pred_mode=pred_mode)


def _clz32(x: torch.Tensor) -> torch.Tensor:
    """Count of leading zero bits of positive int32 values (binary search; exact, no floating point)."""
    x = x.to(torch.int32)
    n = torch.zeros_like(x)
    for s in (16, 8, 4, 2, 1):
        hi = x >> s
        move = hi != 0
        x = torch.where(move, hi, x)
        n = n + torch.where(move, torch.full_like(n, s), torch.zeros_like(n))
    return 31 - n
