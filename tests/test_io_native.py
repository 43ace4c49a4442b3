"""Native readers / writer of the on-disk formats either side of the path (csrc/gsfm_io.cpp) against
the Python host-side restatements: covariance_rot.txt (reference src/uncertainty.cpp:164-229) and the 1DSfM dataset
files (T/io/read_1dsfm.cc:93-412).  Host-only code: these tests need no GPU.  A synthetic dataset is written here; when
the reference checkout is mounted (this container only) the shipped Madrid_Metropolis files are compared as well."""
import os

import numpy as np
import pytest

from globalsfmpy_b200 import _capi as capi, solver, viewgraph as vg

MADRID = "/root/reference/datasets/Madrid_Metropolis"


def _write_dataset(d, rng, n_listed=40, n_pairs=120, n_tracks=300):
    keep = np.sort(rng.choice(n_listed, size=n_listed - 6, replace=False))
    with open(os.path.join(d, "cc.txt"), "w") as f:
        f.write("\n".join(str(int(v)) for v in keep) + "\n")
    with open(os.path.join(d, "list.txt"), "w") as f:
        for k in range(n_listed):
            f.write(f"images/im_{k:04d}.jpg" + (f" 0 {1000.0 + 3.25 * k:.5f}" if k % 3 else "") + "\n")
    pairs = set()
    while len(pairs) < n_pairs:
        a, b = rng.integers(0, n_listed, 2)
        if a != b:
            pairs.add((int(min(a, b)), int(max(a, b))))
    pairs = sorted(pairs)
    R = vg.so3_exp(vg.random_rotation_vectors(rng, len(pairs)))
    t = rng.normal(size=(len(pairs), 3))
    with open(os.path.join(d, "EGs.txt"), "w") as f:
        for (a, b), Rk, tk in zip(pairs, R, t):
            f.write(f"{a} {b} " + " ".join(repr(float(x)) for x in Rk.ravel()) + " " + " ".join(repr(float(x)) for x in tk) + "\n")
        a, b = pairs[0]                                       # a repeated pair: the later entry replaces the stored one
        f.write(f"{a} {b} " + " ".join(repr(float(x)) for x in R[1].ravel()) + " " + " ".join(repr(float(x)) for x in t[1]) + "\n")
    with open(os.path.join(d, "tracks.txt"), "w") as f:
        f.write(f"{n_tracks}\n")
        for k in range(n_tracks):
            n = int(rng.integers(2, 9))
            views = rng.choice(n_listed, size=n, replace=(k % 25 == 0))   # every 25th track may see a view twice: rejected
            f.write(f"{n} " + " ".join(f"{int(v)} {int(rng.integers(0, 5000))}" for v in views) + "\n")


def _python_read(d):
    """The compat module's Python reader as the reference point."""
    import importlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "globalsfmpy_b200", "compat"))
    sfm = importlib.import_module("GlobalSfMpy")
    recon, graph = sfm.Reconstruction(), sfm.ViewGraph()
    sfm.Read1DSFM(d, recon, graph)
    return recon, graph


def _compare_dataset(d):
    nat = solver.read_1dsfm(d)
    recon, graph = _python_read(d)
    edges = graph.GetAllEdges()
    assert sorted(nat["view_ids"].tolist()) == sorted(recon.ViewIds())
    for v, f in zip(nat["view_ids"].tolist(), nat["focal_length_priors"].tolist()):
        prior = getattr(recon.View(int(v)), "focal_length_prior", None)
        assert (prior or 0.0) == pytest.approx(f, rel=0, abs=0)
    keys = [tuple(sorted(p)) for p in nat["pairs"].tolist()]
    assert sorted(keys) == sorted(edges) and len(set(keys)) == len(keys)
    rot_py = np.array([edges[k].rotation_2 for k in keys])
    pos_py = np.array([edges[k].position_2 for k in keys])
    m_py = np.array([edges[k].num_verified_matches for k in keys])
    assert np.abs(vg.so3_exp(nat["rotation_2"]) - vg.so3_exp(rot_py)).max() < 1e-13
    assert np.array_equal(nat["position_2"], pos_py)
    assert np.array_equal(nat["num_verified_matches"], m_py)
    assert solver.read_1dsfm(d, with_matches=False)["num_verified_matches"] is None
    return nat


def test_read_1dsfm_synthetic(tmp_path):
    _write_dataset(str(tmp_path), np.random.default_rng(5))
    nat = _compare_dataset(str(tmp_path))
    assert nat["num_listed_views"] == 40 and len(nat["view_ids"]) == 34 and nat["num_verified_matches"].sum() > 0


@pytest.mark.skipif(not os.path.isdir(MADRID), reason="the reference checkout is only mounted in the build container")
def test_read_1dsfm_madrid_matches_python_reader():
    nat = _compare_dataset(MADRID)
    assert len(nat["view_ids"]) == 394 and len(nat["pairs"]) > 20000


def test_covariance_rot_roundtrip_and_python_parity(tmp_path):
    rng = np.random.default_rng(3)
    n = 57
    ids = np.stack([rng.integers(0, 500, n), rng.integers(500, 1000, n)], axis=1)
    cov6 = rng.normal(size=(n, 6)) * 10.0 ** rng.integers(-12, 3, size=(n, 1))
    cov6[0, 0] = 5e-324                                        # a subnormal and a negative zero survive the bit-cast text
    cov6[1, 1] = -0.0
    rot = rng.normal(size=(n, 3))
    p_native, p_python = str(tmp_path / "native.txt"), str(tmp_path / "python.txt")
    solver.write_covariance_rot(p_native, ids, cov6, rot)
    vg.write_covariance_text(p_python, ids, cov6, rot)
    for path in (p_native, p_python):                          # both writers' files through both readers
        a, c, r = solver.read_covariance_rot(path)
        a2, c2, r2 = vg.parse_covariance_text(path)
        assert np.array_equal(a, ids) and np.array_equal(a2, ids)
        assert np.array_equal(c.view(np.uint64), cov6.view(np.uint64)) and np.array_equal(c2.view(np.uint64), cov6.view(np.uint64))
        assert np.array_equal(r, rot) and np.array_equal(r2, rot)


@pytest.mark.skipif(not os.path.isdir(MADRID), reason="the reference checkout is only mounted in the build container")
def test_covariance_rot_madrid_matches_python_reader():
    a, c, r = solver.read_covariance_rot(os.path.join(MADRID, "covariance_rot.txt"))
    a2, c2, r2 = vg.parse_covariance_text(os.path.join(MADRID, "covariance_rot.txt"))
    assert len(a) == 23783 and np.array_equal(a, a2)
    assert np.array_equal(c.view(np.uint64), c2.view(np.uint64)) and np.array_equal(r.view(np.uint64), r2.view(np.uint64))


def test_readers_fail_loudly(tmp_path):
    with pytest.raises(capi.GsfmError) as e:
        solver.read_covariance_rot(str(tmp_path / "missing.txt"))
    assert e.value.code == capi.ERR_INVALID and "cannot open" in str(e.value)
    with pytest.raises(capi.GsfmError):
        solver.read_1dsfm(str(tmp_path))
    bad = tmp_path / "bad.txt"
    bad.write_text("# h\n# h\n1 2 3 4\n")
    with pytest.raises(capi.GsfmError) as e:
        solver.read_covariance_rot(str(bad))
    assert "truncated" in str(e.value)
