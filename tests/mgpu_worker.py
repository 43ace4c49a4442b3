"""Worker for the multi-GPU parity test: run under torchrun, one rank per GPU.
Every rank builds the edge-sharded solver on the same problem; rank 0 also solves it on one GPU; the two
solutions and cost sequences must agree (only the summation order of the cross-GPU reduction differs)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from globalsfmpy_b200 import _capi as capi, solver as S, viewgraph as vg  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {}
    cases = [
        ("aa_cauchy", dict(num_views=400, num_edges=12000, seed=3, noise_deg=1.0, outlier_fraction=0.1), capi.ANGLE_AXIS,
         capi.Loss.make(capi.LOSS_CAUCHY, 0.05)),
        ("cov_softl1", dict(num_views=150, num_edges=3000, seed=5, outlier_fraction=0.05, covariance=True), capi.ANGLE_AXIS_COVARIANCE,
         capi.Loss.make(capi.LOSS_SOFTLONE, 1.0)),
        # an edge count no world size divides: the shards differ by one edge, the replicas must still be bit-identical
        ("aa_uneven", dict(num_views=300, num_edges=7001, seed=11, noise_deg=1.0, outlier_fraction=0.05), capi.ANGLE_AXIS,
         capi.Loss.make(capi.LOSS_CAUCHY, 0.05)),
    ]
    # both exchange modes: fused (peer memory inside the persistent PCG kernel) and ncclAllReduce between kernels
    for (name, kw, etype, loss), fused in [(c, f) for c in cases for f in (True, False)]:
        name = f"{name}_{'fused' if fused else 'nccl'}"
        g = vg.synthetic_pose_graph(**kw)
        prob = S.make_problem(g, etype)
        o = capi.default_options_py()
        o.loss = loss
        o.pcg_rtol = 1e-12
        o.pcg_max_iterations = 2000
        o.linear_solver = capi.SOLVER_PCG   # sharded solves are PCG; keep the single-GPU reference on the same path
        o.device = local
        sh = S.Solver(prob, o, rank=rank, world_size=world)
        sh.connect(dist, fused=fused)
        e0, e1 = sh.edge_range()
        assert (e0, e1) == (g.num_edges * rank // world, g.num_edges * (rank + 1) // world)
        sh.set_rotations(g.omega_init)
        s, tr = sh.iterate(o.max_num_iterations + 1, trace_capacity=256)
        om = sh.get_rotations()
        # replicated state must be bit-identical on every rank
        t = torch.from_numpy(om.copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref), "ranks diverged"
        if rank == 0:
            one, s1, tr1 = S.solve(prob, o, g.omega_init, trace_capacity=256)
            mean, mx = vg.mean_angular_error(one, om)
            out[name] = dict(mean=mean, max=mx, iters=(s.num_iterations, s1.num_iterations), term=(s.termination, s1.termination),
                             cost=(s.final_cost, s1.final_cost),
                             max_cost_rel=max(abs(a.cost - b.cost) / abs(b.cost) for a, b in zip(tr, tr1)))
        sh.close()
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
