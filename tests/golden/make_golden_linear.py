"""Generate tests/golden/mx_linear*.npz: outputs of the UNMODIFIED reference mx.Linear forward
(microxscaling/mx/linear.py) on seeded inputs.  Authoring container only:

    python tests/golden/make_golden_linear.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import mx_specs  # noqa: E402  (also puts the reference on sys.path)
from mx.linear import linear as mx_linear_fn  # noqa: E402

CASES = [
    # name,              M    K    N    bias  bfloat flush seed
    ("mx_linear_qkv",    70, 192, 576, True,  32, False, 41),     # DeiT-tiny qkv shape, bias
    ("mx_linear_bf16",   33, 128, 100, True,  16, False, 42),     # DiT bfloat 16
    ("mx_linear_nobias", 48,  64,  64, False, 32, True,  43),
]


def main():
    torch.set_num_threads(1)
    for name, M, K, N, has_bias, bfloat, flush, seed in CASES:
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(M, K, generator=g) * torch.exp(0.7 * torch.randn(M, 1, generator=g))
        w = torch.randn(N, K, generator=g) * (K ** -0.5)
        b = torch.randn(N, generator=g) * 0.1 if has_bias else None
        y = mx_linear_fn(x, w, bias=b, mx_specs=mx_specs(bfloat, flush))
        arrays = {"x": x.numpy(), "w": w.numpy(), "y": y.detach().numpy(),
                  "meta": np.array([M, K, N, int(has_bias), bfloat, int(flush)], dtype=np.int64)}
        if has_bias:
            arrays["b"] = b.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, y.shape, float(y.abs().max()))


if __name__ == "__main__":
    main()
