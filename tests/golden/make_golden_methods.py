"""Generate tests/golden/predictor_methods_*.npz and analysis_overlap.npz: the dense returns of every method of the
reference's predictor class, and the --anal overlap score, from the UNMODIFIED reference on seeded inputs.

    python tests/golden/make_golden_methods.py          (authoring container only: needs /root/reference)

  exponent_based_sign               the example copy's working body (make_golden.py)
  partial_Q / partial_K / MXINT4 / two_step_leading_ones
                                    funcs.exponent_approximation methods (funcs/exponent_based_prediction.py:96-318)
  exponent_based_sign_leading_ones  the example copy (microxscaling/examples/deit/exponent_based_prediction.py:163-178)
  diff_idx_analysis                 funcs/analysis.py:136-157 on (true top-k of the exact scores, predicted top-k)
"""
import os
import sys

import numpy as np
import torch

from make_golden import HERE, REF, _example, exponent_approximation, make_inputs, mx_matmul, mx_specs, working_exponent_based_sign

CASES = [
    # name,                        B  H  N   hd  bfloat flush  kind     seed
    ("predictor_methods_deit",     1, 2, 40, 64, 32, False, "randn", 31),
    ("predictor_methods_dit_bf16", 1, 2, 33, 72, 16, False, "edges", 32),
    ("predictor_methods_pixart",   1, 2, 24, 72, 32, True,  "edges", 33),
]


def main():
    for name, B, H, N, hd, bfloat, flush, kind, seed in CASES:
        q, k, v = make_inputs(B, H, N, hd, seed, kind)
        specs = mx_specs(bfloat, flush)
        arrays = {"q": q.numpy(), "k": k.numpy(), "meta": np.array([B, H, N, hd, 0, bfloat, int(flush)], dtype=np.int64)}
        for m in ("partial_Q", "partial_K", "MXINT4", "two_step_leading_ones"):
            obj = exponent_approximation(Q=q.clone(), K=k.clone(), mx_specs=specs)
            aq, ak = getattr(obj, m)()
            arrays[m + ".Q"], arrays[m + ".K"] = aq.numpy().copy(), ak.numpy().copy()
        obj = exponent_approximation(Q=q.clone(), K=k.clone(), mx_specs=specs)
        aq, ak = working_exponent_based_sign(obj)
        arrays["exponent_based_sign.Q"], arrays["exponent_based_sign.K"] = aq.numpy().copy(), ak.numpy().copy()
        aq, ak = _example.exponent_approximation(Q=q.clone(), K=k.clone(), mx_specs=specs).exponent_based_sign_leading_ones()
        arrays["exponent_based_sign_leading_ones.Q"], arrays["exponent_based_sign_leading_ones.K"] = aq.numpy().copy(), ak.numpy().copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, {k_: a.shape for k_, a in arrays.items() if k_ != "meta"})

    # ---- --anal overlap score (funcs/analysis.py:136-157; caller: workloads/DiT/models.py analysis branch)
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_analysis", os.path.join(REF, "funcs/analysis.py"))
    ana = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ana)
    B, H, N, hd, k_true, k_pred = 3, 2, 48, 64, 8, 12
    q, k, v = make_inputs(B, H, N, hd, 41, "randn")
    specs = mx_specs(32, False)
    true_scores = mx_matmul(q, k.transpose(-2, -1), mx_specs=specs, mode_config='aa') * (1.0 / hd ** 0.5)
    obj = exponent_approximation(Q=q, K=k, mx_specs=specs)
    ex_q, ex_k = working_exponent_based_sign(obj)
    pred = ex_q @ ex_k.transpose(-2, -1)
    true_idx = torch.sort(true_scores, dim=-1, descending=True, stable=True).indices[..., :k_true].contiguous()
    pred_idx = torch.sort(pred, dim=-1, descending=True, stable=True).indices[..., :k_pred].contiguous()
    score = ana.diff_idx_analysis(true_idx, pred_idx)
    np.savez_compressed(os.path.join(HERE, "analysis_overlap.npz"), q=q.numpy(), k=k.numpy(), true_idx=true_idx.numpy(),
                        pred_idx=pred_idx.numpy(), diff_idx_analysis=np.array([score], dtype=np.float64),
                        meta=np.array([B, H, N, hd, k_true, k_pred], dtype=np.int64))
    print("analysis_overlap diff_idx_analysis =", score)


if __name__ == "__main__":
    main()
