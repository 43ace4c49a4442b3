#!/usr/bin/env python
"""Live kernel timings (CUDA events on the solver stream) of the bench workload: K1, K1c, K2 alone, one CG step of the
persistent PCG kernel with its phase split, and a short timed run of whole LM steps.

  [GSFM_RA_LIB=variant.so] python profiles/kernel_times.py [workload] [repeats]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GSFM_RA_PROFILE_PHASES", "1")
import bench  # noqa: E402
from globalsfmpy_b200 import _capi as capi, solver as S  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "syn_10k_1M"
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 50
g, loss, etype = bench.build_workload(name)
prob = S.make_problem(g, etype)
opt = bench.bench_options(loss)
sv = S.Solver(prob, opt)
sv.set_rotations(g.omega_init)
sv.iterate(2)
kt = sv.time_kernels(repeats=repeats)
sv.set_rotations(g.omega_init)
for _ in range(5):
    sv.iterate(1)
t0 = time.perf_counter()
n = 0
lin = 0
for _ in range(100):
    s, _ = sv.iterate(1)
    n += 1
    lin += s.total_linear_iterations
    if s.termination != 0:
        sv.set_rotations(g.omega_init)
dt = time.perf_counter() - t0
kt.update(lib=os.environ.get("GSFM_RA_LIB", "default"), workload=name, us_per_lm_step=1e6 * dt / n, pcg_iterations_per_step=lin / n,
          edges_per_s=g.num_edges * n / dt)
print(json.dumps(kt))
