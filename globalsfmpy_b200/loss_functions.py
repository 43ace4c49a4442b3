"""Host-side mirror of the reference's loss library (scripts/loss_functions.py): same class names, same
constructor arguments, same `Evaluate(sq_norm, out)` contract (out[0..2] = rho, rho', rho'').  The classes only
carry parameters; `Evaluate` runs the device kernel (gsfm_ra_eval_loss) -- there is no CPU implementation here."""
import math

import numpy as np

from . import _capi as capi


class LossFunction:
    """Base class, stands for `sfm.LossFunction` (bind_src/GlobalSfMpy.cpp:163-165)."""
    _gsfm_device_backed = False

    def __init__(self):
        pass

    def Evaluate(self, sq_norm, out):
        raise NotImplementedError


class _DeviceLoss(LossFunction):
    _gsfm_device_backed = True

    def _struct(self):
        raise NotImplementedError

    def Evaluate(self, sq_norm, out):
        from . import solver
        r = solver.eval_loss(self._struct(), np.array([float(sq_norm)]))[0]
        out[0], out[1], out[2] = float(r[0]), float(r[1]), float(r[2])


class TrivialLoss(_DeviceLoss):
    def _struct(self):
        return capi.Loss.make(capi.LOSS_TRIVIAL)


class HuberLoss(_DeviceLoss):
    def __init__(self, a):
        self.a, self.b = a, a * a

    def _struct(self):
        return capi.Loss.make(capi.LOSS_HUBER, self.a)


class SoftLOneLoss(_DeviceLoss):
    def __init__(self, a):
        self.a, self.b, self.c = a, a * a, 1.0 / (a * a)

    def _struct(self):
        return capi.Loss.make(capi.LOSS_SOFTLONE, self.a)


class CauchyLoss(_DeviceLoss):
    def __init__(self, a):
        self.b, self.c = a * a, 1.0 / (a * a)

    def _struct(self):
        return capi.Loss.make(capi.LOSS_CAUCHY, math.sqrt(self.b))


class ArctanLoss(_DeviceLoss):
    def __init__(self, a):
        self.a, self.b = a, 1 / (a * a)

    def _struct(self):
        return capi.Loss.make(capi.LOSS_ARCTAN, self.a)


class TolerantLoss(_DeviceLoss):
    def __init__(self, a, b):
        assert a >= 0 and b > 0
        self.a, self.b = a, b

    def _struct(self):
        return capi.Loss.make(capi.LOSS_TOLERANT, self.a, self.b)


class TukeyLoss(_DeviceLoss):
    def __init__(self, a):
        self.a_squared = a * a

    def _struct(self):
        return capi.Loss.make(capi.LOSS_TUKEY, math.sqrt(self.a_squared))


class LOneHalfLoss(_DeviceLoss):
    def __init__(self, a):
        self.a, self.sqrt_a = a, math.sqrt(a)

    def _struct(self):
        return capi.Loss.make(capi.LOSS_LONEHALF, self.a)


class LTwoLoss(_DeviceLoss):
    def __init__(self, a, sigma2=None):
        self.a_sq = a * a

    def _struct(self):
        return capi.Loss.make(capi.LOSS_LTWO, math.sqrt(self.a_sq))


class GemanMcClureLoss(_DeviceLoss):
    def __init__(self, a, sigma2):
        self.a_sq, self.sigma2 = a * a, sigma2

    def _struct(self):
        return capi.Loss.make(capi.LOSS_GEMANMCCLURE, math.sqrt(self.a_sq), self.sigma2)


class ScaledLoss(_DeviceLoss):
    def __init__(self, rho, a):
        self.rho, self.a = rho, a

    def _struct(self):
        from .losses import loss_to_struct
        return loss_to_struct(self, verify=False)


class MAGSACWeightBasedLoss(_DeviceLoss):
    _kind = capi.LOSS_MAGSAC3

    def __init__(self, sigma, inverse=False):
        self.sigma_max, self.use_weight_inverse = sigma, inverse

    def _struct(self):
        return capi.Loss.make(self._kind, self.sigma_max, inverse=self.use_weight_inverse)


class MAGSACWeightBasedLoss4(MAGSACWeightBasedLoss):
    _kind = capi.LOSS_MAGSAC4

    def __init__(self, sigma, inverse=True):   # the reference's nu=4 class defaults to the inverse weight (:345)
        super().__init__(sigma, inverse)


class MAGSACWeightBasedLoss9(MAGSACWeightBasedLoss):
    _kind = capi.LOSS_MAGSAC9
