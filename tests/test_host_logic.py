"""CPU tests of the host-side logic that needs no GPU: the dense arithmetic of the predictor class (fed with the
oracle's codes in place of the CUDA quantizer's), the --anal figures, and the spec resolvers - against fixtures
generated from the unmodified reference (tests/golden/make_golden_methods.py, make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import mxint8_oracle as O
from mx_quantization_b200 import analysis
from mx_quantization_b200.predictor import _clz32, exponent_approximation
from mx_quantization_b200.specs import PathSpecs, resolve_linear_specs, resolve_specs
from tests.helpers import mx_specs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cpu_predictor(q, k, bfloat, flush):
    """The class without its constructor's device plumbing: codes / exponents from the oracle quantizer."""
    obj = object.__new__(exponent_approximation)
    obj._sp = PathSpecs(bfloat, flush)
    obj.mx_specs = mx_specs(bfloat, flush)
    obj.Q, obj.K = q, k
    qc, qe = O.quantize_mxint8(q, 32, bfloat, flush)
    kc, ke = O.quantize_mxint8(k, 32, bfloat, flush)
    obj._q, obj._k = (qc, qe, O.sign_words(qc)), (kc, ke, O.sign_words(kc))
    return obj


@pytest.mark.parametrize("name", ["predictor_methods_deit", "predictor_methods_dit_bf16", "predictor_methods_pixart"])
def test_predictor_class_dense_math(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, H, N, hd, _, bfloat, flush = (int(x) for x in z["meta"])
    q, k = torch.from_numpy(z["q"]), torch.from_numpy(z["k"])
    obj = _cpu_predictor(q, k, bfloat, bool(flush))
    for method in ("partial_Q", "partial_K", "MXINT4", "two_step_leading_ones", "exponent_based_sign_leading_ones"):
        aq, ak = getattr(obj, method)()
        assert torch.equal(aq, torch.from_numpy(z[method + ".Q"])), method
        assert torch.equal(ak, torch.from_numpy(z[method + ".K"])), method
    # the exponent-sign operand (the GPU method goes through mxp_exp_sign_approx; here its dense restatement)
    (qc, qe, _), (kc, ke, _) = obj._q, obj._k
    assert torch.equal(obj._exp_sign(qc, qe), torch.from_numpy(z["exponent_based_sign.Q"]))
    assert torch.equal(obj._exp_sign(kc, ke), torch.from_numpy(z["exponent_based_sign.K"]))


def test_clz32():
    x = torch.tensor([1, 2, 3, 4, 127, 128, 255, 256, 2 ** 20 + 5, 2 ** 30], dtype=torch.int32)
    want = torch.tensor([31, 30, 30, 29, 25, 24, 24, 23, 11, 1], dtype=torch.int32)
    assert torch.equal(_clz32(x), want)


def test_diff_idx_analysis_reference_value():
    z = np.load(os.path.join(GOLDEN, "analysis_overlap.npz"))
    got = analysis.diff_idx_analysis(torch.from_numpy(z["true_idx"]), torch.from_numpy(z["pred_idx"]))
    assert abs(got - float(z["diff_idx_analysis"][0])) < 1e-12


@pytest.mark.parametrize("rows", [1, 2, 5, 8, 13])
def test_coverage_rate_any_row_count(rows):
    g = torch.Generator().manual_seed(rows)
    dense = torch.rand(2, 3, rows, 70, generator=g) < 0.2
    idxs = [torch.nonzero(dense[b, h, r]).flatten() for b in range(2) for h in range(3) for r in range(rows)]
    words = torch.zeros(2, 3, rows, 3, dtype=torch.int64)
    for n, ids in enumerate(idxs):
        b, h, r = n // (3 * rows), (n // rows) % 3, n % rows
        for j in ids.tolist():
            words[b, h, r, j >> 5] |= 1 << (j & 31)
    mask = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    want = float((dense.any(2).sum(-1).double() / rows).mean())
    assert abs(analysis.coverage_rate(mask) - want) < 1e-12


def test_linear_specs_reject_other_weight_formats():
    ok = mx_specs()
    assert resolve_linear_specs(ok) == resolve_specs(ok)
    for key, val in (("w_elem_format", "int4"), ("w_elem_format", None), ("w_elem_format", "fp8_e4m3"),
                     ("round_weight", "floor"), ("round", "even")):
        bad = dict(ok)
        bad[key] = val
        with pytest.raises(ValueError):
            resolve_linear_specs(bad)
        resolve_specs({k: v for k, v in bad.items() if k != "round"} | {"round": "nearest"}) if key == "round" else resolve_specs(bad)


def test_exp_fast_argument_error_bound():
    """csrc/mxprune_device.cuh::exp_fast_nonpos evaluates 2^(fl(x * fl(log2 e))) for x <= 0 in the long-sequence attention kernel.
    The part of its error that is NOT MUFU.EX2's own (<= 2 ulp) comes from rounding the constant and the product to fp32; this
    checks the bound the header states - relative error <= 0.9e-7 |x| (+ one half ulp for tiny |x|) - in float64 arithmetic, and
    that the weight exp(x) of a key shrinks faster than that error grows (the bound's justification)."""
    import numpy as np
    rng = np.random.default_rng(0)
    x = -np.concatenate([rng.uniform(0, 1, 20000), rng.uniform(1, 20, 20000), rng.uniform(20, 87, 20000)]).astype(np.float32)
    l2e = np.float32(1.4426950408889634)
    arg = (x * l2e).astype(np.float32)                          # __fmul_rn
    approx = np.exp2(arg.astype(np.float64))                    # an exact 2^y on the rounded argument
    exact = np.exp(x.astype(np.float64))
    rel = np.abs(approx / exact - 1.0)
    assert np.all(rel <= 0.9e-7 * np.abs(x.astype(np.float64)) + 6e-8), float(rel.max())
    assert np.all(rel * exact <= 4e-8)                          # absolute effect on a weight in (0, 1]: never above ~1/3 ulp of 1


def _select_model(u, kk, nsamp_keys=256, fbins=128, max_range=1400):
    """Host model of k_select_long_tc's control flow for ONE row of integer keys u (csrc/mxprune_predict_long_tc.cuh): returns
    (threshold T, ties wanted at T, passes over the keys).  Sample -> fine window with clamp bins (bins of 2^fs keys + one
    level for the digit inside the bin) -> radix levels when the threshold lands in a clamp bin."""
    import numpy as np
    n = len(u)
    s = u[:nsamp_keys]
    umin, umax = int(s.min()), int(s.max())
    rng = umax - umin
    csh = int(rng >> 7).bit_length()
    hist = np.bincount((s - umin) >> csh, minlength=fbins)
    ks = max(1, (kk * len(s) + n // 2) // n)
    cum, b = 0, fbins - 1
    while b > 0 and cum + hist[b] < ks:
        cum += hist[b]
        b -= 1
    fs = 0 if rng <= max_range else int(rng // (max_range // 2)).bit_length()
    passes = 0
    if fs <= 6:
        c_est = umin + (b << csh) + ((1 << csh) >> 1)
        lo = ((c_est >> fs) - fbins // 2) << fs
        e = np.clip((u >> fs) - (lo >> fs), 0, fbins - 1)
        h = np.bincount(e, minlength=fbins)
        passes += 1
        cum, b = 0, fbins - 1
        while b > 0 and cum + h[b] < kk:
            cum += h[b]
            b -= 1
        if 0 < b < fbins - 1:
            prefix, krem = (lo >> fs) + b, kk - cum
            if fs:                                              # the digit inside the bin: one level of 2^fs bins
                d = u[(u >> fs) == prefix] & ((1 << fs) - 1)
                h2 = np.bincount(d, minlength=1 << fs)
                passes += 1
                cum, b2 = 0, (1 << fs) - 1
                while b2 > 0 and cum + h2[b2] < krem:
                    cum += h2[b2]
                    b2 -= 1
                prefix, krem = (prefix << fs) | b2, krem - cum
            return prefix, krem, passes + 1
    # radix levels, 6 bits each, most significant digit first
    width = max(1, int(u.max()).bit_length())
    lev = (width + 5) // 6
    prefix, krem = 0, kk
    for lv in range(lev):
        lo = 6 * (lev - 1 - lv)
        d = (u[(u >> (lo + 6)) == prefix] >> lo) & 63
        h = np.bincount(d, minlength=64)
        passes += 1
        cum, b = 0, 63
        while b > 0 and cum + h[b] < krem:
            cum += h[b]
            b -= 1
        prefix, krem = (prefix << 6) | b, krem - cum
    return prefix, krem, passes + 1


def test_long_selection_model_is_exact_on_every_path():
    """Whatever the sample predicts, the row ends with the canonical threshold (the top_k-th largest key) and the number of
    ties still wanted at it - the emit pass then keeps keys > T and the first `ties` keys == T in ascending index (SURVEY 8a:
    the canonical tie rule).  Covers: narrow rows (two passes), wide rows (bins of 2^fs keys, three passes), a sample that
    misleads (radix levels), constant rows, heavy ties."""
    import numpy as np
    rng = np.random.default_rng(5)
    seen = set()
    for trial in range(300):
        n = int(rng.choice([512, 1024, 4096]))
        kind = trial % 6
        sigma = [12, 60, 400, 3000, 60, 60][kind]
        u = np.rint(rng.normal(0, sigma, n)).astype(np.int64)
        if kind == 4:
            u[:256] = np.rint(rng.normal(-8 * sigma, 3, 256))      # the sample sits far below the row
        if kind == 5:
            u = (u // 40) * 40                                      # long tie runs
        if trial % 50 == 49:
            u[:] = 7                                                # every key equal
        u = u - u.min() + 1
        kk = max(1, int(n * float(rng.choice([0.1, 0.25, 0.5]))))
        T, ties, passes = _select_model(u, kk, nsamp_keys=256 if n >= 2048 else 128)
        srt = np.sort(u)[::-1]
        want_T = int(srt[kk - 1])
        assert T == want_T, (trial, kind, n, kk)
        assert ties == kk - int((u > want_T).sum()) and 1 <= ties <= int((u == want_T).sum())
        seen.add(passes)
    assert {2, 3} <= seen and max(seen) >= 4                        # every path was exercised
