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


# knots of a tabulated loss: s = 0, then 2^(min_exp + o) (1 + m / per_octave); 2^-80 ... 2^64 covers every squared residual the
# whitened / weighted angular errors can produce (a whitening factor of 1e6 on an angle of pi gives s ~ 1e13 ~ 2^43)
TABLE_MIN_EXP, TABLE_OCTAVES, TABLE_PER_OCTAVE = -80, 144, 32


def table_knots(min_exp=TABLE_MIN_EXP, octaves=TABLE_OCTAVES, per_octave=TABLE_PER_OCTAVE):
    o = np.arange(octaves)[:, None]
    m = np.arange(per_octave)[None, :]
    body = np.ldexp(1.0 + m / per_octave, (min_exp + o)).ravel()
    return np.concatenate([[0.0], body, [np.ldexp(1.0, min_exp + octaves)]])


def _evaluate(loss, s):
    out = [0.0, 0.0, 0.0]
    loss.Evaluate(float(s), out)
    return out


def tabulate_loss(loss, verify=True, rtol=1e-7, per_octave=TABLE_PER_OCTAVE):
    """ANY object with the reference's `Evaluate(sq_norm, out)` (bind_src/GlobalSfMpy.cpp:33-65 accepts every Python subclass
    of sfm.LossFunction) as a device table: the host samples rho, rho', rho'' at the knots ONCE (the reference calls the
    object once per edge per evaluation), the device interpolates with quintic Hermite polynomials
    (GSFM_RA_LOSS_TABULATED, include/gsfm_ra.h).  With verify=True the interpolant is compared with the object at points
    between the knots (on the device when one is present) and a loss the table cannot represent to `rtol` -- a kink or a
    jump inside a cell -- is refused with the measured error instead of being solved approximately."""
    knots = table_knots(per_octave=per_octave)
    tab = np.ascontiguousarray([_evaluate(loss, s) for s in knots], dtype=np.float64)
    if not np.all(np.isfinite(tab)):
        raise UnsupportedLoss(f"{type(loss).__name__}.Evaluate returned non-finite values on [0, 2^{TABLE_MIN_EXP + TABLE_OCTAVES}]")
    L = capi.Loss.make(capi.LOSS_TABULATED)
    L.table = tab.ctypes.data_as(capi._dp)
    L.table_min_exp, L.table_octaves, L.table_per_octave = TABLE_MIN_EXP, TABLE_OCTAVES, per_octave
    L._table_owner = tab            # keep the host table alive as long as the struct
    if verify:
        L._table_error = verify_table(loss, L, rtol)
    return L


def verify_table(loss, L, rtol):
    """Max relative error of the device interpolant against the object at the MIDPOINT OF EVERY CELL of the table (a kink or a
    jump anywhere inside a cell shows up there) plus two points beyond its ends."""
    from . import solver
    knots = table_knots(L.table_min_exp, L.table_octaves, L.table_per_octave)
    probes = np.concatenate([[0.0, 1e-300], 0.5 * (knots[1:-1] + knots[2:])])
    try:
        dev = solver.eval_loss(L, probes)
    except capi.GsfmError as err:
        if err.code == capi.ERR_NO_DEVICE:
            return None                                # no device in this process (CPU-only tooling): nothing to verify against
        raise
    ref = np.array([_evaluate(loss, s) for s in probes])
    # per-column scale: |reference| with a floor of 1e-12 of the column's largest magnitude (rho'' may vanish identically)
    floor = 1e-12 * np.maximum(np.abs(ref).max(axis=0), 1e-300)
    err = np.abs(dev - ref) / np.maximum(np.abs(ref), floor)
    # rho'' enters only the Triggs correction; it is the second derivative of the interpolant and is held to a 100x looser
    # bound, measured against |rho''| + 1e-3 |rho'| / s rather than |rho''| alone (it crosses zero, or vanishes identically)
    with np.errstate(divide="ignore", invalid="ignore"):
        natural = np.where(probes > 0, np.abs(ref[:, 1]) / np.maximum(probes, 1e-300), np.inf)    # |rho'| / s: the scale rho'' lives on
    err2 = np.abs(dev[:, 2] - ref[:, 2]) / np.maximum(np.abs(ref[:, 2]) + 1e-3 * natural, 1e-300)
    worst = float(max(err[:, 0].max(), err[:, 1].max(), err2.max() / 100.0))
    if worst > rtol:
        k = int(np.argmax(np.maximum(err[:, 0], err[:, 1])))
        raise UnsupportedLoss(f"{type(loss).__name__}: the tabulated form deviates from Evaluate by {worst:.2e} relative (bound {rtol:.0e}), e.g. at "
                              f"s = {probes[k]:.6g}: table {dev[k]} vs object {ref[k]} -- a loss with kinks or jumps between knots cannot be "
                              "tabulated; compose it from the closed-form classes instead")
    return worst


def loss_to_struct(loss, verify=True):
    """gsfm_ra_loss for a LossFunction object (None -> TrivialLoss, as a null ceres loss).  The shipped classes of
    scripts/loss_functions.py map to their closed forms, ComposedLoss of two of them to the native composition, anything
    else -- user subclasses, deeper nestings -- to a table of the object's own Evaluate (tabulate_loss)."""
    if loss is None:
        return capi.Loss.make(capi.LOSS_TRIVIAL)
    if isinstance(loss, capi.Loss):
        return loss
    try:
        return _closed_form(loss, verify)
    except UnsupportedLoss:
        if not hasattr(loss, "Evaluate"):
            raise
    return tabulate_loss(loss, verify)


def _closed_form(loss, verify=True):
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
        elif n == "ComposedLoss":                       # loss_functions.py:250-265: rho(s) = f(g(s))
            f, g = _closed_form(inner.f, verify=False), _closed_form(inner.g, verify=False)
            try:
                L = capi.Loss.compose(f, g, scale=scale)
            except ValueError as e:
                raise UnsupportedLoss(str(e)) from e
            if g.kind >= capi.LOSS_MAGSAC3:
                raise UnsupportedLoss("a MAGSAC inner function has no closed form on the device")
        else:
            raise UnsupportedLoss(f"LossFunction subclass {n!r} has no closed form on the device")
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
