"""Generate tests/golden/analysis_coverage.json: the reference's funcs/analysis.py total_chosen_k
(coverage rate, :56-110) evaluated on the canonical index sets stored in the golden fixtures.
Authoring container only (needs /root/reference):

    python tests/golden/make_golden_analysis.py
"""
import importlib.util
import json
import os
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")
spec = importlib.util.spec_from_file_location("ref_analysis", "/root/reference/funcs/analysis.py")
ra = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ra)

out = {}
for name in ["deit_small", "dit_small", "dit_bf16", "deit_tiny_c1"]:
    z = np.load(os.path.join(HERE, name + ".npz"))
    out[name] = ra.total_chosen_k(torch.from_numpy(z["idx"].astype(np.int64)))
json.dump(out, open(os.path.join(HERE, "analysis_coverage.json"), "w"), indent=1)
print(out)
