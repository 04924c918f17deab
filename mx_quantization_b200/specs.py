"""mx_specs handling at the boundary.

The reference passes a plain dict (workloads/deit/scripts/main.py:719-735) or an ``MxSpecs``
UserDict (microxscaling/mx/specs.py:61-181) down to ``exponent_approximation`` /
``mx.matmul``.  Both are accepted unchanged here.  Only one combination is on the hot path
(SURVEY.md section 5): MXINT8 activations, block 32, 8-bit shared scale from the block max,
round-to-nearest(-away), bfloat in {0/32, 16}.  Anything else raises - no silent fallback.
"""
from collections.abc import Mapping
from typing import NamedTuple


class PathSpecs(NamedTuple):
    bfloat_bits: int      # 16 or 32
    flush: bool           # mx_flush_fp32_subnorms


_DEFAULTS = {   # microxscaling/mx/specs.py:81-120
    "scale_bits": 0, "a_elem_format": None, "w_elem_format": None, "shared_exp_method": "max",
    "block_size": 0, "bfloat": 0, "fp": 0, "bfloat_subnorms": True, "round": "nearest",
    "round_output": "nearest", "round_mx_output": "nearest", "mx_flush_fp32_subnorms": False,
    "custom_cuda": False, "round_weight": "nearest",
}


def resolve_specs(mx_specs) -> PathSpecs:
    if mx_specs is None:
        raise ValueError("mx_specs is required on the MXINT8 pruned-attention path")
    if not isinstance(mx_specs, Mapping):
        raise TypeError(f"mx_specs must be a dict or MxSpecs, got {type(mx_specs).__name__}")
    get = lambda k: mx_specs[k] if k in mx_specs and mx_specs[k] is not None else _DEFAULTS[k]  # noqa: E731

    def need(key, allowed):
        v = get(key)
        if v not in allowed:
            raise ValueError(f"mx_specs[{key!r}]={v!r} is not on the B200 hot path (supported: {allowed})")
        return v

    need("a_elem_format", ("int8",))
    need("block_size", (32,))
    need("scale_bits", (0, 8))                      # 0 is promoted to 8, mx_ops.py:329-332
    need("shared_exp_method", ("max",))
    need("round_output", ("nearest",))
    need("round_mx_output", ("nearest",))
    need("fp", (0,))
    need("bfloat_subnorms", (True,))
    bfloat = need("bfloat", (0, 16, 32))
    flush = bool(get("mx_flush_fp32_subnorms"))
    # custom_cuda selected the reference's own CUDA quantizer; results are identical, so it is
    # accepted and ignored.
    return PathSpecs(16 if bfloat == 16 else 32, flush)


def resolve_linear_specs(mx_specs) -> PathSpecs:
    """mx.Linear additionally quantizes the WEIGHT with ``w_elem_format`` / ``round_weight`` and rounds with ``round``
    (microxscaling/mx/linear.py:36-75): only MXINT8 weights with round-to-nearest are on the path.  ``w_elem_format=None``
    leaves the weight unquantised in the reference and int4 / fp8 give other values - those raise here instead of
    silently running the MXINT8 kernel."""
    sp = resolve_specs(mx_specs)
    get = lambda k: mx_specs[k] if k in mx_specs and mx_specs[k] is not None else _DEFAULTS[k]  # noqa: E731
    for key, allowed in (("w_elem_format", ("int8",)), ("round_weight", ("nearest",)), ("round", ("nearest",))):
        v = get(key)
        if v not in allowed:
            raise ValueError(f"mx_specs[{key!r}]={v!r} is not on the B200 MX Linear path (supported: {allowed})")
    return sp
