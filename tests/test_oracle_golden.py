"""The oracle (oracle/mxint8_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py) and against the reference's
own known-answer vectors.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import mxint8_oracle as O

CASES = ["deit_small", "dit_small", "dit_bf16", "pixart_flush", "deit_edges", "deit_tiny_c1"]


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    d = {k: torch.from_numpy(z[k].astype(np.int64) if z[k].dtype == np.int16 else z[k]) for k in z.files}
    B, H, N, hd, top_k, bfloat, flush = (int(x) for x in z["meta"])
    return d, dict(B=B, H=H, N=N, hd=hd, top_k=top_k, bfloat=bfloat, flush=bool(flush))


@pytest.mark.parametrize("name", CASES)
def test_quantizer_bit_exact(golden_dir, name):
    d, m = load(golden_dir, name)
    qc, qe = O.quantize_mxint8(d["q"], 32, m["bfloat"], m["flush"])
    assert torch.equal(O.dequantize_mxint8(qc, qe), d["MX_Q"])
    assert torch.equal(O.predictor_exponents(qc, qe).to(torch.float32), d["shared_exp_Q"])
    assert int(qc.to(torch.int32).abs().max()) <= 127
    if "MX_K" in d:
        kc, ke = O.quantize_mxint8(d["k"], 32, m["bfloat"], m["flush"])
        assert torch.equal(O.dequantize_mxint8(kc, ke), d["MX_K"])
        assert torch.equal(O.predictor_exponents(kc, ke).to(torch.float32), d["shared_exp_K"])


@pytest.mark.parametrize("name", CASES[:-1])
def test_predictor_bit_exact(golden_dir, name):
    d, m = load(golden_dir, name)
    qc, qe = O.quantize_mxint8(d["q"], 32, m["bfloat"], m["flush"])
    kc, ke = O.quantize_mxint8(d["k"], 32, m["bfloat"], m["flush"])
    assert torch.equal(O.exponent_based_sign(qc, qe), d["approx_Q"])
    assert torch.equal(O.exponent_based_sign(kc, ke), d["approx_K"])
    s_int = O.pred_scores_integer(qc, qe, kc, ke)
    s_mm = O.pred_scores_matmul(qc, qe, kc, ke)
    assert torch.equal(s_mm, d["pred_scores"])
    # block-exact integer formula (what the CUDA kernel computes): bit-equal to the reference
    # wherever the reference's own fp32 sum is order-independent; elsewhere parity is unpinned
    ok = O.pred_window_ok(O.predictor_exponents(qc, qe), O.predictor_exponents(kc, ke),
                          O.block_widths(m["hd"]))
    assert torch.equal(s_int[ok], d["pred_scores"][ok])
    if name not in ("pixart_flush", "deit_edges"):      # no all-zero blocks there
        assert bool(ok.all())
    assert float(ok.float().mean()) > 0.95


@pytest.mark.parametrize("name", CASES)
def test_topk_and_output(golden_dir, name):
    d, m = load(golden_dir, name)
    r = O.pruned_attention(d["q"], d["k"], d["v"], m["top_k"], bfloat=m["bfloat"], flush=m["flush"])
    assert torch.equal(r["idx"], d["idx"])                       # canonical sets AND order
    words = O.idx_to_mask_words(r["idx"], m["N"])
    dense = O.mask_words_to_dense(words, m["N"])
    assert int(dense.sum()) == m["B"] * m["H"] * m["N"] * m["top_k"]
    # raw torch.topk (tie order unspecified): same k-th value, same strictly-greater set
    pred = r["pred_scores"]
    kth = pred.gather(-1, d["idx"][..., -1:])
    kth_t = pred.gather(-1, d["topk_idx_torch"][..., -1:])
    assert torch.equal(kth, kth_t)
    gt = pred > kth
    assert bool((dense | ~gt).all())
    t_dense = torch.zeros_like(dense)
    t_dense.scatter_(-1, d["topk_idx_torch"], True)
    assert bool((t_dense | ~gt).all())
    # fp32 stages: stated tolerance 1e-3 of the reference's max-abs (north star)
    ref_vals, ref_out = d["true_vals"], d["out"]
    assert float((r["true_vals"] - ref_vals).abs().max()) <= 1e-5 * float(ref_vals.abs().max())
    assert float((r["out"] - ref_out).abs().max()) <= 1e-3 * float(ref_out.abs().max())


def test_log2_boundary_exponents(golden_dir):
    z = np.load(os.path.join(golden_dir, "quantizer_log2_boundary.npz"))
    x, ref = torch.from_numpy(z["x"]), torch.from_numpy(z["MX"])
    c, e = O.quantize_mxint8(x)
    assert torch.equal(O.dequantize_mxint8(c, e), ref)
    # the table itself, re-derived from this machine's torch.log2 (what mx_ops.py:93-97 calls)
    for n in range(-125, 128):
        js = np.arange(1, 64, dtype=np.int64)
        bits = (((n - 1 + 127) << 23) | (2 ** 23 - js)).astype(np.uint32)
        xs = torch.from_numpy(bits.view(np.float32).copy())
        ref_e = torch.floor(torch.log2(xs)).to(torch.int32)
        assert torch.equal(O.shared_exponent_from_absmax(xs), ref_e), n


def test_bf16_half_away_kat():
    # microxscaling/mx/tests/test_corners_elemwise.py:137-144 pins "nearest" == half away
    x = torch.tensor([1.0 + 2 ** -8, 1.0 + 2 ** -7 + 2 ** -8, -(1.0 + 2 ** -8), 1.0 + 2 ** -9], dtype=torch.float32)
    y = O.bf16_round_half_away(x)
    exp = torch.tensor([1.0 + 2 ** -7, 1.0 + 2 ** -6, -(1.0 + 2 ** -7), 1.0], dtype=torch.float32)
    assert torch.equal(y, exp)


def test_mxint8_hw_kat():
    """Known-answer vectors of microxscaling/mx/tests/test_corners_mx.py:83-124
    (test_mx_hw_test: int8, block 10, round nearest)."""
    x = np.array([
        [1.0] * 10,
        [1.0] * 5 + [2.0] * 5,
        [-1.0] * 5 + [-2.0] * 5,
        [1.0] * 5 + [-2.0] * 5,
        [1.015625, 1.0234375, 1.03125, 1.0390625, 1.25, 1.2578125, 1.9375, 1.9453125, 1.984375, 1.9921875],
        [-1.984375, -1.9765625, -1.96875, -1.9609375, -1.9375, -1.9296875, -1.75, -1.7421875, -1.0, -1.9921875],
        [1.99609375, 1.98828125, 0.0, 0.00390625, 0.0078125, 0.01171875, -0.015625, -0.01171875, -0.0078125,
         -0.00390625]], dtype=np.float32)
    y = np.array([
        [1.0] * 10,
        [1.0] * 5 + [2.0] * 5,
        [-1.0] * 5 + [-2.0] * 5,
        [1.0] * 5 + [-2.0] * 5,
        [1.015625, 1.03125, 1.03125, 1.046875, 1.25, 1.265625, 1.9375, 1.953125, 1.984375, 1.984375],
        [-1.984375, -1.984375, -1.96875, -1.96875, -1.9375, -1.9375, -1.75, -1.75, -1.0, -1.984375],
        [1.984375, 1.984375, 0.0, 0.0, 0.015625, 0.015625, -0.015625, -0.015625, -0.015625, 0.0]], dtype=np.float32)
    c, e = O.quantize_mxint8(torch.from_numpy(x), block=10)
    assert torch.equal(O.dequantize_mxint8(c, e, block=10), torch.from_numpy(y))


def test_scatter_script_kat():
    """funcs/test_scatter.py (np seed 0, 200x128, block 32) prints
    'S_meta range: [-184.000, 180.000]', 'S_meta zeros: 1707/40000',
    'S_meta negatives: 19064/40000' for sum_b 2^(eq+ek) * <sign_q, sign_k>_b  (:156-174)."""
    np.random.seed(0)
    Q = torch.from_numpy(np.random.randn(200, 128)).float()
    K = torch.from_numpy(np.random.randn(200, 128)).float()
    qc, qe = O.quantize_mxint8(Q)
    kc, ke = O.quantize_mxint8(K)
    S = O.pred_scores_integer(qc, qe, kc, ke)
    assert float(S.min()) == -184.0 and float(S.max()) == 180.0
    assert int((S == 0).sum()) == 1707
    assert int((S < 0).sum()) == 19064


MODE_CASES = ["modes_deit", "modes_dit_bf16", "modes_pixart", "modes_deit_197"]


@pytest.mark.parametrize("mode", ["partial_Q", "partial_K", "MXINT4", "two_step_leading_ones", "true_ex", "exact"])
@pytest.mark.parametrize("name", MODE_CASES)
def test_other_rankings_against_reference(golden_dir, name, mode):
    """partial_Q / partial_K / MXINT4 (funcs/exponent_based_prediction.py:179-199,274-318) and the approx_flag=False branch
    (workloads/deit/scripts/main.py:130), outputs of the unmodified reference
    (tests/golden/make_golden_modes.py)."""
    d, m = load(golden_dir, name)
    r = O.pruned_attention(d["q"], d["k"], d["v"], m["top_k"], bfloat=m["bfloat"], flush=m["flush"], pred_mode=mode)
    if f"{mode}.rank_scores" in d:
        assert torch.equal(r["pred_scores"], d[f"{mode}.rank_scores"])
    assert torch.equal(r["idx"], d[f"{mode}.idx"])
    # raw torch.topk (tie order unspecified) keeps the same strictly-greater set
    pred = r["pred_scores"]
    kth = pred.gather(-1, d[f"{mode}.idx"][..., -1:])
    assert torch.equal(kth, pred.gather(-1, d[f"{mode}.topk_idx_torch"][..., -1:]))
    ref_out = d[f"{mode}.out"]
    assert float((r["out"] - ref_out).abs().max()) <= 1e-3 * float(ref_out.abs().max())


@pytest.mark.parametrize("name", ["elsa_deit", "elsa_dit", "elsa_edges"])
def test_elsa_ranking_against_reference(golden_dir, name):
    """ELSA (funcs/elsa_approximation.py:103-145, main.py:119-124): outputs of the unmodified reference class with a
    projection matrix from its own generator (tests/golden/make_golden_elsa.py)."""
    d, m = load(golden_dir, name)
    r = O.pruned_attention(d["q"], d["k"], d["v"], m["top_k"], bfloat=m["bfloat"], flush=m["flush"],
                           pred_mode="ELSA", orthogonal_matrix=d["P"])
    assert torch.equal(r["pred_scores"], d["rank_scores"])
    assert torch.equal(r["idx"], d["idx"])
    assert float((r["out"] - d["out"]).abs().max()) <= 1e-3 * float(d["out"].abs().max())
    # inside a row the reference's ranking is that of min(s_q . s_k, cap): what the CUDA kernel ranks on
    from mx_quantization_b200.ops import elsa_rank_cap
    qc, qe = O.quantize_mxint8(d["q"], 32, m["bfloat"], m["flush"])
    kc, ke = O.quantize_mxint8(d["k"], 32, m["bfloat"], m["flush"])
    mq, mk = O.dequantize_mxint8(qc, qe), O.dequantize_mxint8(kc, ke)
    s_q = torch.where(mq @ d["P"].T >= 0, 1.0, -1.0)
    s_k = torch.where(mk @ d["P"].T >= 0, 1.0, -1.0)
    dots = torch.clamp(s_q @ s_k.transpose(-2, -1), max=elsa_rank_cap(m["hd"]))
    dots = torch.where((mk.abs().sum(-1) == 0).unsqueeze(-1), torch.zeros_like(dots), dots)
    assert torch.equal(O.canonical_topk(dots, m["top_k"]), d["idx"])


CROSS_CASES = ["pixart_cross", "pixart_cross_k77", "pixart_cross_all"]


@pytest.mark.parametrize("mode", ["partial_Q", "partial_K", "MXINT4", "two_step_leading_ones", "true_ex", "exact"])
def test_cross_attention_other_rankings(mode):
    """PixArt cross-attention with the additive text mask in the other ranking modes ("exact" = the branch an
    excluded timestep takes, MX_transformer_block.py:806,833-834) - oracle vs the reference's outputs."""
    d, m = load_cross("modes_pixart_cross")
    bias = d["key_bias"].reshape(m["B"], 1, 1, m["S"])
    r = O.pruned_attention(d["q"], d["k"], d["v"], m["top_k"], scale=1.0 / (m["hd"] ** 0.5),
                           bfloat=m["bfloat"], flush=m["flush"], key_bias=bias, pred_mode=mode)
    assert torch.equal(r["pred_scores"], d[f"{mode}.rank_scores"])
    assert torch.equal(r["idx"], d[f"{mode}.idx"].to(torch.int64))
    assert float((r["out"] - d[f"{mode}.out"]).abs().max()) <= 1e-6 * float(d[f"{mode}.out"].abs().max())


def load_cross(name):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    d = {k: torch.from_numpy(z[k]) for k in z.files}
    B, H, Nq, S, hd, top_k, bfloat, flush = (int(x) for x in z["meta"])
    return d, dict(B=B, H=H, Nq=Nq, S=S, hd=hd, top_k=top_k, bfloat=bfloat, flush=bool(flush))


@pytest.mark.parametrize("name", CROSS_CASES)
def test_cross_attention_with_key_bias(name):
    """PixArt cross-attention (Nq != Nk, additive text mask on true AND predicted scores,
    workloads/PixArt/models/MX_transformer_block.py:791-859) - oracle vs the reference's outputs."""
    d, m = load_cross(name)
    bias = d["key_bias"].reshape(m["B"], 1, 1, m["S"])
    r = O.pruned_attention(d["q"], d["k"], d["v"], m["top_k"], scale=1.0 / (m["hd"] ** 0.5),
                           bfloat=m["bfloat"], flush=m["flush"], key_bias=bias)
    assert torch.equal(r["pred_scores"], d["pred_scores"])
    assert torch.equal(r["idx"], d["idx"])
    assert torch.equal(r["true_vals"], d["true_vals"])
    assert float((r["out"] - d["out"]).abs().max()) == 0.0
    # the integer restatement of the scores ranks identically
    r2 = O.pruned_attention(d["q"], d["k"], d["v"], m["top_k"], scale=1.0 / (m["hd"] ** 0.5),
                            bfloat=m["bfloat"], flush=m["flush"], key_bias=bias, integer_scores=True)
    assert torch.equal(r2["idx"], d["idx"])


def test_coverage_rate_matches_reference_analysis():
    """coverage rate from the kept-key bitmask == funcs/analysis.py:56-110 total_chosen_k
    (values in tests/golden/analysis_coverage.json were produced by the reference function)."""
    import json
    from mx_quantization_b200.analysis import coverage_rate, topk_overlap
    here = os.path.dirname(os.path.abspath(__file__))
    want = json.load(open(os.path.join(here, "golden", "analysis_coverage.json")))
    for name, cov in want.items():
        z = np.load(os.path.join(here, "golden", name + ".npz"))
        idx = torch.from_numpy(z["idx"].astype(np.int64))
        n_keys = int(z["meta"][2])
        words = O.idx_to_mask_words(idx, n_keys)
        mask = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
        assert abs(coverage_rate(mask) - cov) < 1e-12
        assert bool((topk_overlap(mask, idx) == 1.0).all())
        other = torch.from_numpy(z["topk_idx_torch"].astype(np.int64))
        ov = topk_overlap(mask, other)                  # torch.topk picks other members among ties
        assert float(ov.min()) >= 0.0 and float(ov.max()) <= 1.0


LINEAR_CASES = ["mx_linear_qkv", "mx_linear_bf16", "mx_linear_nobias"]


def load_linear(name):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    M, K, N, has_bias, bfloat, flush = (int(v) for v in z["meta"])
    d = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return d, dict(M=M, K=K, N=N, has_bias=bool(has_bias), bfloat=bfloat, flush=bool(flush))


@pytest.mark.parametrize("name", LINEAR_CASES)
def test_mx_linear_oracle_matches_reference(name):
    """oracle restatement of mx.Linear's forward vs outputs of the reference
    (tests/golden/make_golden_linear.py).  Same BLAS, same thread count -> bit-equal."""
    d, m = load_linear(name)
    torch.set_num_threads(1)
    y = O.mx_linear(d["x"], d["w"], d.get("b"), bfloat=m["bfloat"], flush=m["flush"])
    assert float((y - d["y"]).abs().max()) <= 1e-6 * float(d["y"].abs().max())
    if m["bfloat"] == 32:
        assert torch.equal(y, d["y"])
