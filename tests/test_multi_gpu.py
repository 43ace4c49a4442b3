"""Edge-sharded multi-GPU path.  The GPU test needs >= 2 devices (torchrun, NCCL); the CPU test covers the
sharding arithmetic and the exchange pattern with gloo, world_size 2, using the oracle as the per-shard compute."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_solver_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")][-1]
    res = json.loads(line[len("MGPU_RESULT "):])
    for name, v in res.items():
        assert v["term"][0] == v["term"][1], (name, v)
        assert v["iters"][0] == v["iters"][1], (name, v)
        assert v["max_cost_rel"] < 1e-9, (name, v)
        assert v["mean"] < 1e-7, (name, v)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT)
    from globalsfmpy_b200 import _capi as capi, viewgraph as vg
    from oracle import ra_oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = vg.synthetic_pose_graph(60, 500, seed=17, noise_deg=1.0, outlier_fraction=0.1, covariance=True)
    L = capi.Loss.make(capi.LOSS_MAGSAC3, 0.5)
    E = g.num_edges
    e0, e1 = E * rank // world, E * (rank + 1) // world     # the slicing rule of gsfm_ra_solver_create_sharded
    shard = capi.ProblemArrays(60, g.edge_i[e0:e1], g.edge_j[e0:e1], g.omega_ij[e0:e1], cov6=g.cov6[e0:e1],
                               error_type=capi.ANGLE_AXIS_COVARIANCE)
    cost, grad, hd, rp, col, val = orc.assemble(shard, L, g.omega_init)
    # the per-outer-iteration exchange: one all-reduce of [Hd | g | cost]
    buf = torch.from_numpy(np.concatenate([hd.ravel(), grad.ravel(), [cost]]))
    dist.all_reduce(buf)
    # the per-CG-step exchange: one all-reduce of the partial matvec
    x = np.random.default_rng(1).normal(size=(60, 3))
    y = np.zeros((60, 3))
    for a in range(60):
        for s in range(rp[a], rp[a + 1]):
            y[a] += val[s] @ x[col[s]]
    yt = torch.from_numpy(y)
    dist.all_reduce(yt)
    if rank == 0:
        q.put((buf.numpy().copy(), yt.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_edge_sharding_sums_with_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from globalsfmpy_b200 import _capi as capi, viewgraph as vg
    from oracle import ra_oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, 29541, q)) for r in range(world)]
    for p in procs:
        p.start()
    buf, y = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = vg.synthetic_pose_graph(60, 500, seed=17, noise_deg=1.0, outlier_fraction=0.1, covariance=True)
    L = capi.Loss.make(capi.LOSS_MAGSAC3, 0.5)
    full = capi.ProblemArrays(60, g.edge_i, g.edge_j, g.omega_ij, cov6=g.cov6, error_type=capi.ANGLE_AXIS_COVARIANCE)
    cost, grad, hd, rp, col, val = orc.assemble(full, L, g.omega_init)
    ref = np.concatenate([hd.ravel(), grad.ravel(), [cost]])
    assert np.allclose(buf, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    x = np.random.default_rng(1).normal(size=(60, 3))
    yref = np.zeros((60, 3))
    for a in range(60):
        for s in range(rp[a], rp[a + 1]):
            yref[a] += val[s] @ x[col[s]]
    assert np.allclose(y, yref, rtol=1e-12, atol=1e-12 * np.abs(yref).max())
