"""``exponent_approximation`` - same constructor and ``exponent_based_sign()`` signature as
funcs/exponent_based_prediction.py:11-94 of the reference, computed by libmxprune.

The reference class fake-quantises Q and K in its constructor and materialises the dense
+-2^e tensors in ``exponent_based_sign`` so that the caller can do ``@`` / ``torch.topk``
itself (workloads/deit/scripts/main.py:107-123).  That formulation is kept for callers that
still want the tensors; the product path is :func:`mx_quantization_b200.pruned_attention`, which
never materialises them.  Extra accessors expose the compact integer form.
"""
import torch

from . import ops
from .specs import resolve_specs


class exponent_approximation:  # noqa: N801  (name fixed by the reference API)
    def __init__(self, Q: torch.Tensor, K: torch.Tensor, mx_specs):
        resolve_specs(mx_specs)            # reject unsupported configurations up front
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

    # -- reference API ------------------------------------------------------------------
    def exponent_based_sign(self):
        """Returns (approx_Q, approx_K): fp32 tensors shaped like Q, K with values +-2^e_block."""
        return ops.exp_sign_approx(self.Q, self.mx_specs), ops.exp_sign_approx(self.K, self.mx_specs)

    def predict_topk(self, k: int, return_idx: bool = False):
        """Fused replacement for ``exponent_based_sign`` + ``@`` + ``torch.topk``."""
        return ops.predict_topk(self.Q, self.K, self.mx_specs, k, return_idx=return_idx)

    def _unsupported(self, name):
        raise NotImplementedError(
            f"exponent_approximation.{name}: only pred_mode 'ex_pred' (exponent_based_sign) is on the "
            "B200 hot path; the related-work predictors of the reference are out of scope (SURVEY 8f3)")

    def two_step_leading_ones(self):
        self._unsupported("two_step_leading_ones")

    def MXINT4(self):  # noqa: N802
        self._unsupported("MXINT4")

    def partial_Q(self):  # noqa: N802
        self._unsupported("partial_Q")

    def partial_K(self):  # noqa: N802
        self._unsupported("partial_K")
