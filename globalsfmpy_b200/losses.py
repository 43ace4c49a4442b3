"""Mapping of `sfm.LossFunction` objects to the device loss descriptor (gsfm_ra_loss).

The reference evaluates the Python loss object once per edge per evaluation under the GIL
(bind_src/GlobalSfMpy.cpp:36-59).  A CUDA kernel cannot call Python, so the shipped classes of
scripts/loss_functions.py are recognised by class name + attributes and run natively on the device.
The mapping is then VERIFIED: the object's own Evaluate() is sampled at a few squared residuals and
compared with the device kernel; a user subclass that merely shares a name is rejected (there is no CPU
fallback, the caller keeps the Ceres path for it)."""
import math

import numpy as np

from . import _capi as capi


class UnsupportedLoss(TypeError):
    pass


def _unwrap_scaled(loss):
    scale = 1.0
    while type(loss).__name__ == "ScaledLoss":        # loss_functions.py:267-281
        scale *= float(loss.a)
        loss = loss.rho
    return loss, scale


def loss_to_struct(loss, verify=True):
    """gsfm_ra_loss for a LossFunction object (None -> TrivialLoss, as a null ceres loss)."""
    if loss is None:
        return capi.Loss.make(capi.LOSS_TRIVIAL)
    if isinstance(loss, capi.Loss):
        return loss
    inner, scale = _unwrap_scaled(loss)
    n = type(inner).__name__
    g = lambda *names: [float(getattr(inner, k)) for k in names]  # noqa: E731
    try:
        if n == "TrivialLoss":
            L = capi.Loss.make(capi.LOSS_TRIVIAL, scale=scale)
        elif n == "HuberLoss":
            L = capi.Loss.make(capi.LOSS_HUBER, *g("a"), scale=scale)
        elif n == "SoftLOneLoss":
            L = capi.Loss.make(capi.LOSS_SOFTLONE, *g("a"), scale=scale)
        elif n == "CauchyLoss":                         # stores b = a^2 only (:90-92)
            L = capi.Loss.make(capi.LOSS_CAUCHY, math.sqrt(float(inner.b)), scale=scale)
        elif n == "ArctanLoss":
            L = capi.Loss.make(capi.LOSS_ARCTAN, *g("a"), scale=scale)
        elif n == "TolerantLoss":
            L = capi.Loss.make(capi.LOSS_TOLERANT, *g("a", "b"), scale=scale)
        elif n == "TukeyLoss":
            L = capi.Loss.make(capi.LOSS_TUKEY, math.sqrt(float(inner.a_squared)), scale=scale)
        elif n == "LOneHalfLoss":
            L = capi.Loss.make(capi.LOSS_LONEHALF, *g("a"), scale=scale)
        elif n == "LTwoLoss":
            L = capi.Loss.make(capi.LOSS_LTWO, math.sqrt(float(inner.a_sq)), scale=scale)
        elif n == "GemanMcClureLoss":
            L = capi.Loss.make(capi.LOSS_GEMANMCCLURE, math.sqrt(float(inner.a_sq)), float(inner.sigma2), scale=scale)
        elif n in ("MAGSACWeightBasedLoss", "MAGSACWeightBasedLoss4", "MAGSACWeightBasedLoss9"):
            kind = {"MAGSACWeightBasedLoss": capi.LOSS_MAGSAC3, "MAGSACWeightBasedLoss4": capi.LOSS_MAGSAC4,
                    "MAGSACWeightBasedLoss9": capi.LOSS_MAGSAC9}[n]
            L = capi.Loss.make(kind, float(inner.sigma_max), inverse=bool(inner.use_weight_inverse), scale=scale)
        else:
            raise UnsupportedLoss(f"LossFunction subclass {n!r} has no device implementation "
                                  "(ComposedLoss and user-defined losses cannot run inside a CUDA kernel)")
    except AttributeError as e:
        raise UnsupportedLoss(f"{n}: missing attribute {e}") from e
    if verify:
        verify_mapping(loss, L)
    return L


_PROBE = np.array([0.0, 1e-7, 3e-4, 1.7e-3, 0.02, 0.3, 1.0, 7.5, 120.0])


def verify_mapping(loss, L, rtol=1e-9):
    """Compare the Python object's own Evaluate with the device kernel at a few points."""
    if not hasattr(loss, "Evaluate") or getattr(loss, "_gsfm_device_backed", False):
        return
    from . import solver
    dev = solver.eval_loss(L, _PROBE)
    for k, s in enumerate(_PROBE):
        out = [0.0, 0.0, 0.0]
        loss.Evaluate(float(s), out)
        ref = np.array(out, dtype=np.float64)
        if not np.allclose(dev[k], ref, rtol=rtol, atol=1e-11 + 1e-9 * np.abs(ref).max()):
            raise UnsupportedLoss(f"{type(loss).__name__}: Evaluate({s}) = {ref} but the device loss gives {dev[k]}; "
                                  "the class does not behave like the reference loss of that name")
