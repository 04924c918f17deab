"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array written here comes out of the reference's own code:
  * mx.mx_ops.quantize_mx_op / mx.elemwise_ops.quantize_elemwise_op / mx.matmul.matmul
  * funcs.exponent_approximation.__init__            (funcs/exponent_based_prediction.py:12-38)
  * exponent_based_sign                              (working body, microxscaling/examples/deit/
                                                      exponent_based_prediction.py:135-161; the
                                                      funcs/ copy has lines 80-81 commented out)
  * the caller sequence of workloads/deit/scripts/main.py:101-152
with one substitution: ``torch.topk`` -> first k of a stable descending sort (canonical
tie-break, SURVEY 7 "Ties"); the raw ``torch.topk`` indices are stored too.
The GPU box has no /root/reference, so these fixtures are what travels.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "microxscaling"))

from mx.elemwise_ops import quantize_elemwise_op  # noqa: E402
from mx.mx_ops import quantize_mx_op, _quantize_mx  # noqa: E402
from mx.matmul import matmul as mx_matmul  # noqa: E402
from funcs.exponent_based_prediction import exponent_approximation  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "ref_example_pred", os.path.join(REF, "microxscaling/examples/deit/exponent_based_prediction.py"))
_example = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_example)
working_exponent_based_sign = _example.exponent_approximation.exponent_based_sign


def mx_specs(bfloat=32, flush=False):
    # workloads/deit/scripts/main.py:719-735 (bfloat 32), DiT sample.py:36-52 (bfloat 16),
    # PixArt text_local_inference_alpha.py:108-124 (flush True)
    return {
        'w_elem_format': 'int8', 'a_elem_format': 'int8', 'scale_bits': 8,
        'shared_exp_method': 'max', 'block_size': 32, 'bfloat': bfloat, 'fp': 0,
        'bfloat_subnorms': True, 'round': 'nearest', 'round_mx_output': 'nearest',
        'round_output': 'nearest', 'round_weight': 'nearest',
        'mx_flush_fp32_subnorms': flush, 'custom_cuda': False, 'quantize_backprop': False,
    }


def reference_layer(q, k, v, top_k, scale, specs):
    """workloads/deit/scripts/main.py:101-152, reference functions only."""
    out = {}
    true_scores = mx_matmul(q, k.transpose(-2, -1), mx_specs=specs, mode_config='aa')
    true_scores = true_scores * scale
    obj = exponent_approximation(Q=q, K=k, mx_specs=specs)
    out["MX_Q"] = obj.MX_Q.clone()
    out["MX_K"] = obj.MX_K.clone()
    out["shared_exp_Q"] = obj.shared_exponent_Q.squeeze(-1).clone()
    out["shared_exp_K"] = obj.shared_exponent_K.squeeze(-1).clone()
    ex_q, ex_k = working_exponent_based_sign(obj)
    out["approx_Q"], out["approx_K"] = ex_q, ex_k
    pred = ex_q @ ex_k.transpose(-2, -1)
    out["pred_scores"] = pred
    out["topk_idx_torch"] = torch.topk(pred, top_k, dim=-1, largest=True, sorted=True).indices
    idx = torch.sort(pred, dim=-1, descending=True, stable=True).indices[..., :top_k].contiguous()
    out["idx"] = idx
    vals = true_scores.gather(dim=-1, index=idx)
    out["true_vals"] = vals
    attn = torch.zeros_like(true_scores)
    attn.scatter_(-1, idx, torch.softmax(vals, dim=-1).to(attn.dtype))
    out["out"] = mx_matmul(attn, v, mx_specs=specs, mode_config='aa')
    return out


def make_inputs(B, H, N, hd, seed, kind):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, H, N, hd, generator=g)
    k = torch.randn(B, H, N, hd, generator=g)
    v = torch.randn(B, H, N, hd, generator=g)
    if kind == "lognormal":      # per-token scale spread
        q = q * torch.exp(1.5 * torch.randn(B, H, N, 1, generator=g))
        k = k * torch.exp(1.5 * torch.randn(B, H, N, 1, generator=g))
        v = v * torch.exp(0.5 * torch.randn(B, H, N, 1, generator=g))
    if kind == "edges":
        q[0, 0, 3] = 0.0                       # all-zero query row
        k[0, 0, 5] = 0.0                       # all-zero key row
        q[0, 0, 7, 32:64] = 0.0                # all-zero block
        k[0, 1, 2, :32] = 0.0
        q[0, 1, 4, :8] = -1e-6                 # tiny negatives -> code -0
        k[0, 1, 9, 40:50] = -1e-7
        k[0, 0, 11] = k[0, 0, 10]              # duplicated key rows -> exact score ties
        k[0, 0, 12] = k[0, 0, 10]
        v[0, 0, 6] = 0.0
    return q, k, v


CASES = [
    # name,            B  H  N    hd  k   bfloat flush  kind        seed
    ("deit_small",     1, 2, 48,  64, 12, 32, False, "randn",     11),
    ("dit_small",      1, 2, 40,  72, 10, 32, False, "lognormal", 12),
    ("dit_bf16",       1, 2, 33,  72,  9, 16, False, "randn",     13),
    ("pixart_flush",   1, 2, 32,  72,  8, 32, True,  "edges",     14),
    ("deit_edges",     1, 2, 37,  64, 37, 32, False, "edges",     15),
    ("deit_tiny_c1",   1, 3, 197, 64, 40, 32, False, "randn",     16),
]


def boundary_case():
    """Block maxima 1..48 ulps below a power of two: the reference's fp32 floor(log2(.))
    rounds up there (mx_ops.py:93-97); this pins the oracle's LOG2_BUMP table."""
    rows = []
    for n in [-100, -40, -20, -9, -5, -3, -2, -1, 0, 1, 2, 3, 4, 5, 8, 9, 16, 17, 33, 64, 65, 100]:
        for j in [1, 2, 3, 5, 6, 11, 12, 22, 23, 44, 45, 48]:
            bits = np.uint32(((n - 1 + 127) << 23) | (2 ** 23 - j))
            amax = bits.view(np.float32)
            g = np.random.RandomState(n * 100 + j + 10000)
            row = (g.uniform(-1, 1, size=32).astype(np.float32) * amax).astype(np.float32)
            row[g.randint(32)] = amax * (1 if j % 2 else -1)
            rows.append(row)
    x = torch.from_numpy(np.stack(rows))
    y = quantize_mx_op(quantize_elemwise_op(x, mx_specs(), round='nearest'), mx_specs(),
                       elem_format='int8', axes=[-1], round='nearest')
    return {"x": x.numpy(), "MX": y.numpy()}


def main():
    torch.set_num_threads(1)   # deterministic BLAS summation order for the stored floats
    for name, B, H, N, hd, top_k, bfloat, flush, kind, seed in CASES:
        specs = mx_specs(bfloat, flush)
        q, k, v = make_inputs(B, H, N, hd, seed, kind)
        scale = hd ** -0.5
        ref = reference_layer(q, k, v, top_k, scale, specs)
        arrays = {"q": q, "k": k, "v": v, **ref}
        np_arrays = {}
        for key, val in arrays.items():
            a = val.numpy()
            if key.startswith("idx") or key.startswith("topk"):
                a = a.astype(np.int16)
            np_arrays[key] = a
        if name == "deit_tiny_c1":   # keep the big fixture small: drop dense N x N / redundant arrays
            for key in ("pred_scores", "approx_Q", "approx_K", "MX_K", "shared_exp_K"):
                np_arrays.pop(key)
        np_arrays["meta"] = np.array([B, H, N, hd, top_k, bfloat, int(flush)], dtype=np.int64)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **np_arrays)
        print(f"{name}: wrote {os.path.getsize(path) / 1024:.1f} KiB")
    path = os.path.join(HERE, "quantizer_log2_boundary.npz")
    np.savez_compressed(path, **boundary_case())
    print(f"quantizer_log2_boundary: wrote {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
