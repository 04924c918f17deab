"""Tensor-level entry points over the C ABI (include/mxprune.h).

PyTorch is plumbing here: it owns device memory and the current stream; every computation
happens in libmxprune's sm_100a kernels.  CPU tensors are rejected - there is no CPU path.
"""
from ctypes import c_void_p
from typing import Optional, Tuple

import torch

from . import _lib
from .specs import resolve_linear_specs, resolve_specs


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def _view4(t: torch.Tensor, name: str) -> torch.Tensor:
    """Accept the strided (B,H,N,hd) views the attention modules produce; copy only when the
    layout is outside what the kernels address (innermost stride 1, 16-byte aligned rows)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor")
    if t.device.type != "cuda":
        raise ValueError(f"{name}: expected a CUDA tensor (device={t.device}); there is no CPU fallback")
    if t.dim() != 4:
        raise ValueError(f"{name}: expected a 4-D (B,H,N,head_dim) tensor, got shape {tuple(t.shape)}")
    if t.dtype != torch.float32:
        raise ValueError(f"{name}: expected float32 (the reference's fake-quant dtype), got {t.dtype}")
    ok = t.stride(-1) == 1 and t.data_ptr() % 16 == 0 and all(s % 4 == 0 for s in t.stride()[:3])
    return t if ok else t.contiguous()


def _strides(t: torch.Tensor) -> Tuple[int, int, int]:
    return t.stride(0), t.stride(1), t.stride(2)


def _same_device(*ts):
    dev = ts[0].device
    for t in ts[1:]:
        if t is not None and t.device != dev:
            raise ValueError("all tensors must live on the same CUDA device")
    return dev


def limits() -> Tuple[int, int]:
    import ctypes
    a, b = ctypes.c_int(), ctypes.c_int()
    _lib.load().mxp_limits(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def set_attention_path(path: str) -> None:
    """'tcgen05' (default: both GEMMs of the exact stage on the tensor cores) or 'cuda_core' (dp4a)."""
    codes = {"tcgen05": 0, "cuda_core": 1}
    if path not in codes:
        raise ValueError(f"attention path must be one of {sorted(codes)}")
    _lib.check(_lib.load().mxp_set_attention_path(codes[path]), "mxp_set_attention_path")


def set_predict_path(path: str) -> None:
    """'tcgen05' (default: predictor scored on the tensor cores, keys selected in registers) or
    'cuda_core' (XOR/POPC kernel; also used automatically outside the tensor-core kernel's domain)."""
    codes = {"tcgen05": 0, "cuda_core": 1}
    if path not in codes:
        raise ValueError(f"predict path must be one of {sorted(codes)}")
    _lib.check(_lib.load().mxp_set_predict_path(codes[path]), "mxp_set_predict_path")


def set_fused_path(mode) -> None:
    """True / 1 (default): the round-2 launch plans where they are measured faster - the fused one-launch kernel (DeiT-shaped
    calls: 129-256 keys staged in 128-row steps, top_k / Nk <= 0.35), the cost-follows-k attention kernel (Nk <= 256, same
    bound), the two-lanes-per-row kernel for Nk > 256 and the sampled fine window of the long-sequence selection; 2: the fused launch
    wherever the shape is in its domain; False / 0: always the three kernels of round 1 with the dense epilogue, the long-sequence
    selection on its radix levels alone.  Results agree (A/B aid)."""
    _lib.check(_lib.load().mxp_set_fused_path(int(mode)), "mxp_set_fused_path")


def last_launch_count() -> int:
    return _lib.load().mxp_last_launch_count()


def quantize_mxint8(x: torch.Tensor, mx_specs, with_signs: bool = False):
    """MXINT8 codes/exponents of ``x`` along head_dim (replaces quantize_mx_op on this path).

    Returns (codes int8 (B,H,N,hd), exps int8 (B,H,N,nb)[, signs int32 (B,H,N,nb)])."""
    sp = resolve_specs(mx_specs)
    lib = _lib.load()
    x = _view4(x, "x")
    B, H, N, hd = x.shape
    nb = (hd + 31) // 32
    with torch.cuda.device(x.device):
        codes = torch.empty((B, H, N, hd), dtype=torch.int8, device=x.device)
        exps = torch.empty((B, H, N, nb), dtype=torch.int8, device=x.device)
        signs = torch.empty((B, H, N, nb), dtype=torch.int32, device=x.device) if with_signs else None
        rc = lib.mxp_quantize_mxint8(_ptr(x), *_strides(x), B, H, N, hd, sp.bfloat_bits, int(sp.flush),
                                     _ptr(codes), _ptr(exps), _ptr(signs), _stream())
    _lib.check(rc, "mxp_quantize_mxint8")
    return (codes, exps, signs) if with_signs else (codes, exps)


def exp_sign_approx(x: torch.Tensor, mx_specs) -> torch.Tensor:
    """Dense (code<0 ? -1 : +1) * 2^e tensor, fp32 (B,H,N,hd)."""
    sp = resolve_specs(mx_specs)
    lib = _lib.load()
    x = _view4(x, "x")
    B, H, N, hd = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((B, H, N, hd), dtype=torch.float32, device=x.device)
        rc = lib.mxp_exp_sign_approx(_ptr(x), *_strides(x), B, H, N, hd, sp.bfloat_bits, int(sp.flush),
                                     _ptr(out), _stream())
    _lib.check(rc, "mxp_exp_sign_approx")
    return out


# The reference's sources of the top-k ranking (workloads/deit/scripts/main.py:105-131): pred_mode strings as
# the reference spells them, plus "exact" for its `top_k and not approx_flag` branch (top-k of the true scores).
PRED_MODES = {"ex_pred": 0, "partial_Q": 1, "partial_K": 2, "exact": 3, "MXINT4": 4, "two_step_leading_ones": 5,
              "true_ex": 6}


def _pred_mode_code(pred_mode: str) -> int:
    if pred_mode not in PRED_MODES:
        raise NotImplementedError(f"pred_mode={pred_mode!r}: built modes are {sorted(PRED_MODES)} and 'ELSA'; "
                                  "there is no fallback")
    return PRED_MODES[pred_mode]


ELSA_THETA_BIAS = 0.127     # funcs/elsa_approximation.py:101


def elsa_rank_cap(d: int) -> float:
    """d - 2 h_c, h_c the largest Hamming distance the reference's clamp maps to angle 0
    (funcs/elsa_approximation.py:135-136: clamp((pi / k) * hamming - theta_bias, min=0), evaluated in fp32 as torch does):
    hash dot products at or above this value tie."""
    h = torch.arange(0, d + 1, dtype=torch.float32)
    corrected = torch.clamp((torch.pi / d) * h - ELSA_THETA_BIAS, min=0)
    hc = int(torch.nonzero(corrected == 0).max())
    return float(d - 2 * hc)


def _elsa_proj(orthogonal_matrix, hd, dev):
    if orthogonal_matrix is None:
        raise ValueError("pred_mode='ELSA' needs orthogonal_matrix (the reference's (head_dim, head_dim) projection)")
    P = orthogonal_matrix.to(device=dev, dtype=torch.float32).contiguous()
    if tuple(P.shape) != (hd, hd):
        raise ValueError(f"orthogonal_matrix must be ({hd}, {hd}), got {tuple(P.shape)}")
    return P


def _qk_shapes(q, k):
    B, H, Nq, hd = q.shape
    if k.shape[0] != B or k.shape[1] != H or k.shape[3] != hd:
        raise ValueError(f"q {tuple(q.shape)} and k {tuple(k.shape)} disagree on (B,H,head_dim)")
    return B, H, Nq, k.shape[2], hd


def predict_scores(q: torch.Tensor, k: torch.Tensor, mx_specs) -> torch.Tensor:
    """Dense predicted scores (B,H,Nq,Nk) - parity aid, O(N^2) output."""
    sp = resolve_specs(mx_specs)
    lib = _lib.load()
    q, k = _view4(q, "q"), _view4(k, "k")
    _same_device(q, k)
    B, H, Nq, Nk, hd = _qk_shapes(q, k)
    with torch.cuda.device(q.device):
        out = torch.empty((B, H, Nq, Nk), dtype=torch.float32, device=q.device)
        rc = lib.mxp_predict_scores(_ptr(q), *_strides(q), _ptr(k), *_strides(k), B, H, Nq, Nk, hd,
                                    sp.bfloat_bits, int(sp.flush), _ptr(out), _stream())
    _lib.check(rc, "mxp_predict_scores")
    return out


def predict_topk(q: torch.Tensor, k: torch.Tensor, mx_specs, top_k: int, return_idx: bool = False,
                 return_codes: bool = False, pred_mode: str = "ex_pred", scale: Optional[float] = None,
                 key_bias: Optional[torch.Tensor] = None, orthogonal_matrix: Optional[torch.Tensor] = None):
    """Fused quantize + predictor + per-row top-k.

    Returns a dict: mask int32 (B,H,Nq,ceil(Nk/32)) [bit j%32 of word j//32 = key j kept],
    optionally idx int32 (B,H,Nq,top_k) ascending key order, and q/k codes+exps.
    pred_mode: "ex_pred" (exponent-sign, default), "partial_Q", "partial_K", "MXINT4" or "exact" (top-k of the
    true scores * scale) - see PRED_MODES."""
    if pred_mode == "ELSA":
        return _predict_topk_elsa(q, k, mx_specs, top_k, return_idx, return_codes, key_bias, orthogonal_matrix)
    mode = _pred_mode_code(pred_mode)
    if mode != 0:
        return _predict_topk_mode(q, k, mx_specs, top_k, mode, scale, return_idx, return_codes, key_bias)
    if key_bias is not None:
        raise ValueError("predict_topk with pred_mode='ex_pred' takes no key_bias (use pruned_attention)")
    sp = resolve_specs(mx_specs)
    lib = _lib.load()
    q, k = _view4(q, "q"), _view4(k, "k")
    dev = _same_device(q, k)
    B, H, Nq, Nk, hd = _qk_shapes(q, k)
    nb, nw = (hd + 31) // 32, (Nk + 31) // 32
    res = {}
    with torch.cuda.device(dev):
        res["mask"] = torch.empty((B, H, Nq, nw), dtype=torch.int32, device=dev)
        if return_idx:
            res["idx"] = torch.empty((B, H, Nq, int(top_k)), dtype=torch.int32, device=dev)
        if return_codes:
            res["q_codes"] = torch.empty((B, H, Nq, hd), dtype=torch.int8, device=dev)
            res["q_exps"] = torch.empty((B, H, Nq, nb), dtype=torch.int8, device=dev)
            res["k_codes"] = torch.empty((B, H, Nk, hd), dtype=torch.int8, device=dev)
            res["k_exps"] = torch.empty((B, H, Nk, nb), dtype=torch.int8, device=dev)
        ws_bytes = lib.mxp_predict_topk_workspace_bytes(B, H, Nq, Nk, hd)
        ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev) if ws_bytes else None
        rc = lib.mxp_predict_topk(_ptr(q), *_strides(q), _ptr(k), *_strides(k), B, H, Nq, Nk, hd, int(top_k),
                                  sp.bfloat_bits, int(sp.flush), _ptr(res["mask"]), _ptr(res.get("idx")),
                                  _ptr(res.get("q_codes")), _ptr(res.get("q_exps")),
                                  _ptr(res.get("k_codes")), _ptr(res.get("k_exps")),
                                  _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "mxp_predict_topk")
    return res


def _key_bias_2d(key_bias, B, Nk, dev):
    if key_bias.dtype != torch.float32 or key_bias.numel() != B * Nk or key_bias.device != dev:
        raise ValueError("key_bias must be an fp32 tensor with B*Nk elements, (B, ..., Nk), on q's device")
    return key_bias.reshape(B, Nk).contiguous()


def _predict_topk_mode(q, k, mx_specs, top_k, mode, scale, return_idx, return_codes, key_bias=None):
    if return_codes:
        raise ValueError("return_codes is available with pred_mode='ex_pred' only")
    sp = resolve_specs(mx_specs)
    lib = _lib.load()
    q, k = _view4(q, "q"), _view4(k, "k")
    dev = _same_device(q, k)
    B, H, Nq, Nk, hd = _qk_shapes(q, k)
    scale = float(hd) ** -0.5 if scale is None else float(scale)
    res = {}
    with torch.cuda.device(dev):
        res["mask"] = torch.empty((B, H, Nq, (Nk + 31) // 32), dtype=torch.int32, device=dev)
        if return_idx:
            res["idx"] = torch.empty((B, H, Nq, int(top_k)), dtype=torch.int32, device=dev)
        kb = None if key_bias is None else _key_bias_2d(key_bias, B, Nk, q.device)
        rc = lib.mxp_predict_topk_mode(_ptr(q), *_strides(q), _ptr(k), *_strides(k), B, H, Nq, Nk, hd, int(top_k),
                                       mode, scale, sp.bfloat_bits, int(sp.flush), _ptr(kb), Nk, _ptr(res["mask"]),
                                       _ptr(res.get("idx")), _ptr(None), 0, _stream())
    _lib.check(rc, "mxp_predict_topk_mode")
    return res


def _predict_topk_elsa(q, k, mx_specs, top_k, return_idx, return_codes, key_bias, orthogonal_matrix):
    if return_codes or key_bias is not None:
        raise ValueError("pred_mode='ELSA' takes neither return_codes nor key_bias")
    sp = resolve_specs(mx_specs)
    lib = _lib.load()
    q, k = _view4(q, "q"), _view4(k, "k")
    dev = _same_device(q, k)
    B, H, Nq, Nk, hd = _qk_shapes(q, k)
    if Nq != Nk:
        raise ValueError("pred_mode='ELSA' needs Nq == Nk (the reference broadcasts the key norms over query rows)")
    P = _elsa_proj(orthogonal_matrix, hd, dev)
    res = {}
    with torch.cuda.device(dev):
        res["mask"] = torch.empty((B, H, Nq, (Nk + 31) // 32), dtype=torch.int32, device=dev)
        if return_idx:
            res["idx"] = torch.empty((B, H, Nq, int(top_k)), dtype=torch.int32, device=dev)
        rc = lib.mxp_predict_topk_elsa(_ptr(q), *_strides(q), _ptr(k), *_strides(k), B, H, Nq, hd, int(top_k),
                                       _ptr(P), elsa_rank_cap(hd), sp.bfloat_bits, int(sp.flush),
                                       _ptr(res["mask"]), _ptr(res.get("idx")), _stream())
    _lib.check(rc, "mxp_predict_topk_elsa")
    return res


def sparse_attention(q_codes, q_exps, k_codes, k_exps, v: torch.Tensor, mask: torch.Tensor, mx_specs,
                     scale: Optional[float] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Exact MXINT8 softmax(QK^T*scale)V over the keys selected by ``mask``."""
    sp = resolve_specs(mx_specs)
    lib = _lib.load()
    v = _view4(v, "v")
    B, H, Nk, hd = v.shape
    if q_codes.dim() != 4:
        raise ValueError(f"q_codes: expected a 4-D (B,H,Nq,head_dim) tensor, got shape {tuple(q_codes.shape)}")
    Nq = q_codes.shape[2]
    nb, nw = (hd + 31) // 32, (Nk + 31) // 32
    dev = _same_device(v, q_codes, q_exps, k_codes, k_exps, mask)
    for name, t, dt, shape in (("q_codes", q_codes, torch.int8, (B, H, Nq, hd)), ("q_exps", q_exps, torch.int8, (B, H, Nq, nb)),
                               ("k_codes", k_codes, torch.int8, (B, H, Nk, hd)), ("k_exps", k_exps, torch.int8, (B, H, Nk, nb)),
                               ("mask", mask, torch.int32, (B, H, Nq, nw))):
        if t.dtype != dt or not t.is_contiguous():
            raise ValueError(f"{name}: expected a contiguous {dt} tensor")
        if tuple(t.shape) != shape:
            raise ValueError(f"{name}: expected shape {shape} (from v {tuple(v.shape)}), got {tuple(t.shape)}")
    scale = float(hd) ** -0.5 if scale is None else float(scale)
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((B, H, Nq, hd), dtype=torch.float32, device=dev)
        elif (tuple(out.shape) != (B, H, Nq, hd) or out.dtype != torch.float32 or out.stride(-1) != 1
              or out.device != dev):
            raise ValueError("out must be an fp32 (B,H,Nq,head_dim) view with innermost stride 1 on the inputs' device")
        ws_bytes = lib.mxp_sparse_attention_workspace_bytes(B, H, Nq, Nk, hd)
        ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
        rc = lib.mxp_sparse_attention(_ptr(q_codes), _ptr(q_exps), _ptr(k_codes), _ptr(k_exps),
                                      _ptr(v), *_strides(v), _ptr(mask), B, H, Nq, Nk, hd,
                                      scale, sp.bfloat_bits, int(sp.flush),
                                      _ptr(out), *_strides(out), _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "mxp_sparse_attention")
    return out


def pruned_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mx_specs, top_k: int,
                     scale: Optional[float] = None, return_mask: bool = False,
                     out: Optional[torch.Tensor] = None, _kernel_ms: Optional[list] = None,
                     key_bias: Optional[torch.Tensor] = None, pred_mode: str = "ex_pred",
                     orthogonal_matrix: Optional[torch.Tensor] = None):
    """MXINT8 exponent-sign predicted top-k attention: q,k,v (B,H,N,hd) fp32 -> out (B,H,Nq,hd).

    ``pred_mode``: what ranks the keys - "ex_pred" (default), "partial_Q", "partial_K", "MXINT4" (the reference's
    pred_mode values, workloads/deit/scripts/main.py:109-118) or "exact" (its approx_flag=False branch:
    top-k of the true scores, main.py:130), or "ELSA" with ``orthogonal_matrix`` (head_dim, head_dim) as the
    reference's modules receive it (main.py:119-121).  Everything after the selection is the same.

    Drop-in for lines 101-152 of workloads/deit/scripts/main.py (DiT models.py:168-225, PixArt
    MX_transformer_block.py:647-710) when mx_quant, top_k, approx_flag and pred_mode=="ex_pred".
    ``out`` may be any fp32 (B,H,Nq,hd) *view* with innermost stride 1, e.g. a permuted
    (B,Nq,H,hd) buffer so that the module's transpose(1,2).reshape(B,N,C) is free.

    ``key_bias``: PixArt cross-attention's additive text mask (MX_transformer_block.py:794-803,
    821-822), any fp32 tensor with B * Nk elements laid out (B, ..., Nk) - e.g. the reference's
    (B,1,1,S) attention_mask; it is added to the true AND the predicted scores.  Nq may differ
    from Nk (Nk <= 256 with a bias).
    """
    sp = resolve_specs(mx_specs)
    lib = _lib.load()
    q, k, v = _view4(q, "q"), _view4(k, "k"), _view4(v, "v")
    dev = _same_device(q, k, v, out)
    B, H, Nq, Nk, hd = _qk_shapes(q, k)
    if tuple(v.shape) != (B, H, Nk, hd):
        raise ValueError(f"v {tuple(v.shape)} must be (B,H,Nk,head_dim) = {(B, H, Nk, hd)}")
    scale = float(hd) ** -0.5 if scale is None else float(scale)
    elsa = pred_mode == "ELSA"
    if elsa and (key_bias is not None or _kernel_ms is not None or Nq != Nk):
        raise ValueError("pred_mode='ELSA' needs Nq == Nk and takes neither key_bias nor the per-kernel profile entry")
    P = _elsa_proj(orthogonal_matrix, hd, dev) if elsa else None
    mode = 0 if elsa else _pred_mode_code(pred_mode)
    if mode != 0 and _kernel_ms is not None:
        raise ValueError("the per-kernel profile entry is available with pred_mode='ex_pred' only")
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((B, H, Nq, hd), dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (B, H, Nq, hd) or out.dtype != torch.float32 or out.stride(-1) != 1:
            raise ValueError("out must be an fp32 (B,H,Nq,head_dim) view with innermost stride 1")
        mask = torch.empty((B, H, Nq, (Nk + 31) // 32), dtype=torch.int32, device=dev) if return_mask else None
        ws_bytes = lib.mxp_pruned_attention_workspace_bytes(B, H, Nq, Nk, hd)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        args = (_ptr(q), *_strides(q), _ptr(k), *_strides(k), _ptr(v), *_strides(v),
                B, H, Nq, Nk, hd, int(top_k), scale, sp.bfloat_bits, int(sp.flush),
                _ptr(out), *_strides(out), _ptr(mask), _ptr(ws), ws_bytes, _stream())
        kb = None if key_bias is None else _key_bias_2d(key_bias, B, Nk, q.device)
        if elsa:
            rc = lib.mxp_pruned_attention_elsa(*args[:12], B, H, Nq, hd, int(top_k), _ptr(P), elsa_rank_cap(hd),
                                               *args[18:])
        elif mode != 0:
            rc = lib.mxp_pruned_attention_mode(*args[:18], mode, *args[18:-4], _ptr(kb), Nk, *args[-4:])
        elif key_bias is not None:
            if _kernel_ms is not None:
                raise ValueError("the per-kernel profile entry takes no key_bias")
            rc = lib.mxp_pruned_attention_biased(*args[:-4], _ptr(kb), Nk, *args[-4:])
        elif _kernel_ms is None:
            rc = lib.mxp_pruned_attention(*args)
        else:       # measurement aid: per-kernel CUDA-event times (synchronises)
            import ctypes
            ms = (ctypes.c_float * 3)()
            rc = lib.mxp_pruned_attention_profile(*args, ms)
            _kernel_ms[:] = [float(x) for x in ms]
    _lib.check(rc, "mxp_pruned_attention")
    return (out, mask) if return_mask else out


# ---- MX Linear -------------------------------------------------------------------
def mx_linear_prepare_weight(weight: torch.Tensor, mx_specs) -> torch.Tensor:
    """MX-quantize an (out_features, in_features) fp32 weight once into the GEMM's operand order."""
    sp = resolve_linear_specs(mx_specs)
    lib = _lib.load()
    if weight.dim() != 2 or weight.dtype != torch.float32 or not weight.is_cuda:
        raise ValueError("weight must be a CUDA fp32 (out_features, in_features) tensor")
    w = weight if weight.stride(1) == 1 else weight.contiguous()
    N, K = w.shape
    with torch.cuda.device(w.device):
        w_op = torch.empty((lib.mxp_mx_linear_weight_bytes(N, K),), dtype=torch.uint8, device=w.device)
        rc = lib.mxp_mx_linear_prepare_weight(_ptr(w), w.stride(0), N, K, sp.bfloat_bits, int(sp.flush),
                                              _ptr(w_op), _stream())
    _lib.check(rc, "mxp_mx_linear_prepare_weight")
    return w_op


def mx_linear(x: torch.Tensor, weight, bias: Optional[torch.Tensor], mx_specs,
              out_features: Optional[int] = None) -> torch.Tensor:
    """Forward of the reference's mx.Linear (microxscaling/mx/linear.py:20-103) for MXINT8:
    y = A1(A1(MXq(A1(x)) @ MXq(A1(W))^T) + A1(bias)).  ``weight`` is either the fp32 (N,K) tensor or
    the operand returned by mx_linear_prepare_weight (then pass out_features)."""
    sp = resolve_linear_specs(mx_specs)
    lib = _lib.load()
    if x.dtype != torch.float32 or not x.is_cuda:
        raise ValueError("x must be a CUDA fp32 tensor")
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    if weight.dtype == torch.uint8:
        if out_features is None:
            raise ValueError("out_features is required with a prepared weight operand")
        w_op, N = weight, int(out_features)
        if not w_op.is_contiguous() or w_op.numel() != lib.mxp_mx_linear_weight_bytes(N, K):
            raise ValueError(f"prepared weight operand: expected {lib.mxp_mx_linear_weight_bytes(N, K)} contiguous bytes for "
                             f"out_features={N}, in_features={K}, got {w_op.numel()}")
    else:
        N = weight.shape[0]
        w_op = mx_linear_prepare_weight(weight, mx_specs)
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("bias must be a contiguous fp32 tensor with out_features elements")
    _same_device(x, w_op, bias)
    with torch.cuda.device(x.device):
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
        ws_bytes = lib.mxp_mx_linear_workspace_bytes(M, N, K)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
        rc = lib.mxp_mx_linear(_ptr(x2), x2.stride(0), M, K, _ptr(w_op), N, _ptr(bias), sp.bfloat_bits,
                               int(sp.flush), _ptr(out), out.stride(0), _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "mxp_mx_linear")
    return out.reshape(*x.shape[:-1], N)
