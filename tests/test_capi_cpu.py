"""CPU tests of the drop-in boundary: the library loads, exports every symbol of include/gsfm_ra.h,
agrees with the header's defaults, and FAILS LOUDLY (no CPU fallback) when no CUDA device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from globalsfmpy_b200 import _capi as capi, solver, viewgraph as vg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gsfm_ra.h")).read() + open(os.path.join(ROOT, "include", "gsfm_pa.h")).read()
    declared = set(re.findall(r"\b(gsfm_[rp]a_[a-z_0-9]+)\s*\(", header))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    lib = capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gsfm_ra_abi_version() == capi.ABI_VERSION


def test_default_options_match_ceres_defaults():
    o = capi.Options()
    capi.lib().gsfm_ra_default_options(C.byref(o))
    p = capi.default_options_py()
    for f, _ in capi.Options._fields_:
        if f == "loss":
            continue
        assert getattr(o, f) == getattr(p, f), f
    assert (o.max_num_iterations, o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance) == (200, 1e-6, 1e-10, 1e-8)
    assert o.initial_trust_region_radius == 1e4 and o.min_relative_decrease == 1e-3 and o.jacobi_scaling == 1


def test_struct_sizes_match_header(tmp_path):
    """The ctypes mirror against the C header itself: gcc prints sizeof / a few offsets of every struct of include/gsfm_ra.h."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gsfm_pa.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %d %zu %zu %zu\\n",'
                   'sizeof(gsfm_ra_loss),sizeof(gsfm_ra_problem),sizeof(gsfm_ra_options),sizeof(gsfm_ra_iteration),sizeof(gsfm_ra_summary),'
                   'offsetof(gsfm_ra_loss,table),offsetof(gsfm_ra_options,n_gpus),offsetof(gsfm_ra_summary,num_linear_unconverged),GSFM_RA_ABI_VERSION,'
                   'offsetof(gsfm_ra_problem,fixed_view),sizeof(gsfm_pa_problem),offsetof(gsfm_pa_problem,error_type));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(t) for t in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(capi.Loss), C.sizeof(capi.Problem), C.sizeof(capi.Options), C.sizeof(capi.Iteration), C.sizeof(capi.Summary),
            capi.Loss.table.offset, capi.Options.n_gpus.offset, capi.Summary.num_linear_unconverged.offset, capi.ABI_VERSION,
            capi.Problem.fixed_view.offset, C.sizeof(capi.PositionProblem), capi.PositionProblem.error_type.offset]
    assert got == want


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.lib().gsfm_ra_device_count() == 0
    g = vg.synthetic_pose_graph(6, 10, seed=1)
    prob = solver.make_problem(g, capi.ANGLE_AXIS)
    with pytest.raises(capi.GsfmError) as e:
        solver.solve(prob, capi.default_options_py(), g.omega_init)
    assert e.value.code == capi.ERR_NO_DEVICE
    with pytest.raises(capi.GsfmError) as e:
        solver.eval_loss(capi.Loss.make(capi.LOSS_CAUCHY, 0.1), np.array([0.1]))
    assert e.value.code == capi.ERR_NO_DEVICE


def test_invalid_arguments_are_rejected_before_the_device():
    g = vg.synthetic_pose_graph(6, 10, seed=1)
    prob = capi.ProblemArrays(6, g.edge_i, g.edge_j, g.omega_ij, error_type=capi.ANGLE_AXIS_COVARIANCE)  # no cov6
    with pytest.raises(capi.GsfmError) as e:
        solver.solve(prob, capi.default_options_py(), g.omega_init)
    assert e.value.code == capi.ERR_INVALID
    prob = capi.ProblemArrays(6, g.edge_i, g.edge_j, g.omega_ij, error_type=9)  # not a RotationErrorType
    with pytest.raises(capi.GsfmError) as e:
        solver.solve(prob, capi.default_options_py(), g.omega_init)
    assert e.value.code == capi.ERR_INVALID
    assert [capi.lib().gsfm_ra_residual_dim(t) for t in range(9)] == [4, 9, 3, 3, 3, 3, 3, 3, 3]
