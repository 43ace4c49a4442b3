#!/usr/bin/env python
"""Golden vectors for the robust losses and MAGSAC gamma tables.

Run HERE (build container, /root/reference mounted).  Executes the reference's
UNMODIFIED scripts/loss_functions.py against a stub `GlobalSfMpy` module that
provides exactly what the real binding exports for it (bind_src/GlobalSfMpy.cpp:
163-165 LossFunction, :667 tgamma, :669-691 gamma constants/tables), with the
constants and tables parsed from include/gamma_values.cpp.  Output is DATA:
  tests/golden/loss_golden.npz
    s[K]                     squared residuals fed to every loss
    <name>[K,3]              (rho, rho', rho'') the reference class returned
    gamma{3,4,9}_idx / _val  sampled entries of stored_gamma_values{nu}
    const{3,4,9}             [nu, C, sigma_quantile, upper_incomplete_gamma_of_k, N, precision]
"""
import math, re, sys, types
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = sys.argv[2] if len(sys.argv) > 2 else __file__.rsplit("/", 1)[0] + "/loss_golden.npz"

src = open(f"{REF}/include/gamma_values.cpp").read()
stub = types.ModuleType("GlobalSfMpy")
class LossFunction:           # bind_src/GlobalSfMpy.cpp:163-165
    def __init__(self): pass
stub.LossFunction = LossFunction
stub.tgamma = math.gamma       # bind_src/GlobalSfMpy.cpp:31,667 (C tgamma)
consts = {}
for nu in (3, 4, 9):
    def num(name, cast=float):
        return cast(re.search(rf"constexpr\s+\w+\s+{name}{nu}\s*=\s*([-+0-9.eE]+)\s*;", src).group(1))
    n = num("stored_gamma_number", int)
    body = re.search(rf"stored_gamma_values{nu}\s*=\s*\{{(.*?)\}};", src, re.S).group(1)
    table = [float(t) for t in body.replace("\n", " ").split(",") if t.strip()]
    assert len(table) == n, (nu, len(table), n)
    setattr(stub, f"nu{nu}", num("nu"))
    setattr(stub, f"C{nu}", num("C"))
    setattr(stub, f"sigma_quantile{nu}", num("sigma_quantile"))
    setattr(stub, f"upper_incomplete_gamma_of_k{nu}", num("upper_incomplete_gamma_of_k"))
    setattr(stub, f"stored_gamma_number{nu}", n)
    setattr(stub, f"precision_of_stored_gamma{nu}", num("precision_of_stored_gamma"))
    setattr(stub, f"stored_gamma_values{nu}", table)
    consts[nu] = (np.array(table), np.array([num("nu"), num("C"), num("sigma_quantile"),
                                              num("upper_incomplete_gamma_of_k"), n, num("precision_of_stored_gamma")]))
sys.modules["GlobalSfMpy"] = stub
sys.path.insert(0, f"{REF}/scripts")
import loss_functions as lf   # the reference file, unmodified

rng = np.random.default_rng(20231017)
s = np.concatenate([
    [0.0, 1e-12, 1e-9, 1e-7, 1e-6, 4e-7, 1.2e-6, 1e-4, 4e-4, 1e-3, 4.5e-3, 4.6e-3, 1e-2, 0.04, 0.25, 1.0, 1.0000001, 4.0, 9.87, 100.0, 1e4],
    10.0 ** rng.uniform(-8, 2, 400),
    rng.uniform(0, 6e-3, 300),          # dense over the MAGSAC(0.02) support (k^2 sigma^2 = 4.54e-3)
    (np.arange(0, 40) + 0.5) * 2 * 0.02**2 / 1000,   # exact LUT rounding ties for sigma = 0.02
])
cases = {
    "trivial": lf.TrivialLoss(),
    "huber_0.1": lf.HuberLoss(0.1),
    "softlone_0.1": lf.SoftLOneLoss(0.1),
    "cauchy_0.05": lf.CauchyLoss(0.05),
    "cauchy_0.5": lf.CauchyLoss(0.5),
    "arctan_0.3": lf.ArctanLoss(0.3),
    "tolerant_0.5_0.1": lf.TolerantLoss(0.5, 0.1),
    "tukey_0.4": lf.TukeyLoss(0.4),
    "lonehalf_0.7": lf.LOneHalfLoss(0.7),
    "ltwo_0.6": lf.LTwoLoss(0.6, 1.0),
    "gemanmcclure_0.3_2.0": lf.GemanMcClureLoss(0.3, 2.0),
    "magsac3_0.02": lf.MAGSACWeightBasedLoss(0.02),
    "magsac3_0.5": lf.MAGSACWeightBasedLoss(0.5),
    "magsac3inv_0.02": lf.MAGSACWeightBasedLoss(0.02, True),
    "magsac4_0.02": lf.MAGSACWeightBasedLoss4(0.02, False),
    "magsac4inv_0.02": lf.MAGSACWeightBasedLoss4(0.02),
    "magsac9_0.02": lf.MAGSACWeightBasedLoss9(0.02),
    "magsac9_0.3": lf.MAGSACWeightBasedLoss9(0.3),
    "magsac9inv_0.05": lf.MAGSACWeightBasedLoss9(0.05, True),
    "scaled2.5_cauchy_0.05": lf.ScaledLoss(lf.CauchyLoss(0.05), 2.5),
}
out = {"s": s}
for name, loss in cases.items():
    res = np.zeros((len(s), 3))
    for k, sk in enumerate(s):
        o = [0.0, 0.0, 0.0]
        loss.Evaluate(float(sk), o)
        res[k] = o
    out[name] = res
for nu, (table, c) in consts.items():
    idx = np.unique(np.concatenate([np.arange(0, 64), np.arange(0, len(table), 97), [len(table) - 1]]))
    out[f"gamma{nu}_idx"] = idx.astype(np.int64)
    out[f"gamma{nu}_val"] = table[idx]
    out[f"const{nu}"] = c
np.savez_compressed(OUT, **out)
print("wrote", OUT, "K =", len(s), "losses =", len(cases))
