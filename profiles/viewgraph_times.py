#!/usr/bin/env python
"""Wall clock of the device view-graph shaping (host buffers in, host buffers out) next to the host restatements.
  python profiles/viewgraph_times.py [views edges]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from globalsfmpy_b200 import solver as S, viewgraph as vg  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
g = vg.synthetic_pose_graph(N, E, seed=56, noise_deg=1.0, outlier_fraction=0.1)
rng = np.random.default_rng(1)
matches = rng.integers(10, 300, g.num_edges)
out = {"views": N, "edges": int(g.num_edges)}
S.filter_initial_view_graph(N, g.edge_i[:1000], g.edge_j[:1000], matches[:1000], 30)   # CUDA context, pool
for name, fn in (("filter_device_ms", lambda: S.filter_initial_view_graph(N, g.edge_i, g.edge_j, matches, 30)),
                 ("mst_init_device_ms", lambda: S.init_orientations_mst(N, g.edge_i, g.edge_j, g.omega_ij, matches))):
    fn()
    t0 = time.perf_counter()
    r = fn()
    out[name] = 1e3 * (time.perf_counter() - t0)
    if name.startswith("mst"):
        out["boruvka_rounds"] = r[2]
        om_d = r[0]
if E <= 2000000:
    t0 = time.perf_counter()
    vg.filter_initial_view_graph(np.arange(N), np.stack([g.edge_i, g.edge_j], 1), matches, 30)
    out["filter_host_ms"] = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    om_h = vg.max_spanning_tree_orientations(N, g.edge_i, g.edge_j, g.omega_ij, matches)
    out["mst_init_host_ms"] = 1e3 * (time.perf_counter() - t0)
    out["max_abs_rotation_difference"] = float(np.abs(vg.so3_exp(om_d) - vg.so3_exp(om_h)).max())
print(json.dumps(out))
