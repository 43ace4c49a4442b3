#!/usr/bin/env python
"""Madrid_Metropolis (tests/golden fixture): whole-solve wall clock of the one-shot C-ABI call on the device,
ANGLE_AXIS_COVARIANCE + MAGSAC(0.02), dense Cholesky (the bench's `madrid` block without the CPU legs)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from globalsfmpy_b200 import _capi as capi, solver as S, viewgraph as vg  # noqa: E402

g = vg.load_madrid_fixture(os.path.join(ROOT, "tests", "golden", "madrid_metropolis.npz"))
prob = S.make_problem(g, capi.ANGLE_AXIS_COVARIANCE)
o = capi.default_options_py()
o.loss = capi.Loss.make(capi.LOSS_MAGSAC3, 0.02)
o.linear_solver = capi.SOLVER_DENSE_CHOLESKY
S.solve(prob, o, g.omega_init)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    om, s, _ = S.solve(prob, o, g.omega_init)
    ts.append(time.perf_counter() - t0)
print(json.dumps({"gpu_solve_ms_min": 1e3 * min(ts), "gpu_solve_ms_median": 1e3 * sorted(ts)[2], "lm_iterations": s.num_iterations,
                  "final_cost": s.final_cost, "ms_linear": s.ms_linear, "ms_assemble": s.ms_assemble, "launches": s.kernel_launches}))
