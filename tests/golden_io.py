"""Loads tests/golden/*.npz (written by oracle/make_golden.py from the reference)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CHAIN_CASES = ["mrw_gauss1d", "mrw_gauss2d_iid", "mrw_gauss2d_diag", "mrw_gauss2d_dense",
               "mlda_gauss2d", "mrw_linear", "mlda_linear", "mrw_lv", "mlda_lv", "mlda_lv_nonfinite",
               "pcn_lv", "pcn_linear_dense", "mrw_linear_big", "mlda_linear_big",
               "mrw_linear_big_rows12", "pcn_linear_big", "mrw_linear_big_dense", "pcn_linear_big_dense"]
# MLDA with two surrogates as the reference runs it (mlda.py:12-43,60-71,112-117)
MLDA3_CASES = ["mlda3_gauss2d", "mlda3_gauss2d_hier", "mlda3_linear"]
# TemperedUnnormalisedPosterior surrogate (chain/target.py:25-43)
TEMPERED_CASES = ["mlda_linear_tempered"]
# adaptive Metropolis through the reference's AdaptiveMRWProposal (chain/adaptive.py:37-64)
AM_CASES = ["am_gauss2d", "am_gauss2d_refresh7", "am_lv", "am_mlda_lv", "am_mlda_linear"]


def load(name):
    f = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(f["meta"]))
    arrays = {k: f[k] for k in f.files if k != "meta"}
    return meta, arrays


def rel_err(a, b):
    """Relative error that treats equal infinities / NaNs as matching."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    with np.errstate(all='ignore'):
        e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    return np.where(same, 0.0, e)
