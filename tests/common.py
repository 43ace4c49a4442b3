"""Shared helpers for the parity tests: loss table, problem builders."""
import numpy as np

from globalsfmpy_b200 import _capi as capi

# name in tests/golden/loss_golden.npz -> gsfm_ra_loss
GOLDEN_LOSSES = {
    "trivial": capi.Loss.make(capi.LOSS_TRIVIAL),
    "huber_0.1": capi.Loss.make(capi.LOSS_HUBER, 0.1),
    "softlone_0.1": capi.Loss.make(capi.LOSS_SOFTLONE, 0.1),
    "cauchy_0.05": capi.Loss.make(capi.LOSS_CAUCHY, 0.05),
    "cauchy_0.5": capi.Loss.make(capi.LOSS_CAUCHY, 0.5),
    "arctan_0.3": capi.Loss.make(capi.LOSS_ARCTAN, 0.3),
    "tolerant_0.5_0.1": capi.Loss.make(capi.LOSS_TOLERANT, 0.5, 0.1),
    "tukey_0.4": capi.Loss.make(capi.LOSS_TUKEY, 0.4),
    "lonehalf_0.7": capi.Loss.make(capi.LOSS_LONEHALF, 0.7),
    "ltwo_0.6": capi.Loss.make(capi.LOSS_LTWO, 0.6),
    "gemanmcclure_0.3_2.0": capi.Loss.make(capi.LOSS_GEMANMCCLURE, 0.3, 2.0),
    "magsac3_0.02": capi.Loss.make(capi.LOSS_MAGSAC3, 0.02),
    "magsac3_0.5": capi.Loss.make(capi.LOSS_MAGSAC3, 0.5),
    "magsac3inv_0.02": capi.Loss.make(capi.LOSS_MAGSAC3, 0.02, inverse=True),
    "magsac4_0.02": capi.Loss.make(capi.LOSS_MAGSAC4, 0.02),
    "magsac4inv_0.02": capi.Loss.make(capi.LOSS_MAGSAC4, 0.02, inverse=True),
    "magsac9_0.02": capi.Loss.make(capi.LOSS_MAGSAC9, 0.02),
    "magsac9_0.3": capi.Loss.make(capi.LOSS_MAGSAC9, 0.3),
    "magsac9inv_0.05": capi.Loss.make(capi.LOSS_MAGSAC9, 0.05, inverse=True),
    "scaled2.5_cauchy_0.05": capi.Loss.make(capi.LOSS_CAUCHY, 0.05, scale=2.5),
}


def assert_close(a, b, rtol, what=""):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{what}: max abs err / max|ref| = {err:.3e} > {rtol:.1e}"
    return err


def rel_err_rows(a, b):
    """max over rows of |a-b|_inf / |b|_inf of the row; rows that are pure rounding noise (spanning-tree
    edges have residual ~1e-16) are scaled by 1e-6 of the global magnitude instead."""
    a, b = np.asarray(a).reshape(len(a), -1), np.asarray(b).reshape(len(b), -1)
    floor = 1e-6 * max(np.abs(b).max(), 1e-300)
    return np.max(np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), floor))
