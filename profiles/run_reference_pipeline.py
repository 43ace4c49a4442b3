#!/usr/bin/env python3
"""Run the reference's UNMODIFIED scripts/sfm_pipeline.py (step 1-3, onlyRotationAvg=True) on the shipped Madrid_Metropolis dataset
against THIS repo's GlobalSfMpy-compatible module, and score it with the reference's own metric (compare_orientations:
robust AlignRotations + AngularDifference, src/compare_reconstructions.cpp:149-177, 228-262) against the CPU oracle's solve.

  stage (in the build container, /root/reference mounted):   python profiles/run_reference_pipeline.py --stage
      copies scripts/{sfm_pipeline,loss_functions}.py, flags_1dsfm.yaml and the dataset's text files into scratch/ref_stage/
      (git-ignored, NOT gpurun-ignored: it travels to the GPU box with the snapshot; nothing from it is ever committed)
  run (on the GPU box):                                       python profiles/run_reference_pipeline.py
      prints one JSON line (profiles/r02_reference_pipeline.json keeps the committed copy)
"""
import importlib
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "scratch", "ref_stage")
REF = "/root/reference"


def stage():
    os.makedirs(os.path.join(STAGE, "scripts"), exist_ok=True)
    os.makedirs(os.path.join(STAGE, "datasets", "Madrid_Metropolis"), exist_ok=True)
    for f in ("sfm_pipeline.py", "loss_functions.py"):
        shutil.copy(os.path.join(REF, "scripts", f), os.path.join(STAGE, "scripts", f))
    shutil.copy(os.path.join(REF, "flags_1dsfm.yaml"), os.path.join(STAGE, "flags_1dsfm.yaml"))
    for f in ("EGs.txt", "cc.txt", "list.txt", "tracks.txt", "covariance_rot.txt"):
        shutil.copy(os.path.join(REF, "datasets", "Madrid_Metropolis", f), os.path.join(STAGE, "datasets", "Madrid_Metropolis", f))
    print("staged under", STAGE)


def main():
    if "--stage" in sys.argv:
        return stage()
    if not os.path.isdir(STAGE):
        raise SystemExit("nothing staged: run with --stage in the build container first")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "globalsfmpy_b200", "compat"))      # where the reference scripts expect '../build'
    import GlobalSfMpy as sfm                                                   # noqa: F401  (our module under the reference's name)
    cwd = os.getcwd()
    os.chdir(os.path.join(STAGE, "scripts"))                                    # the scripts use paths relative to scripts/
    sys.path.insert(0, os.path.join(STAGE, "scripts"))
    try:
        pipeline = importlib.import_module("sfm_pipeline")                      # UNMODIFIED reference file
        lf = importlib.import_module("loss_functions")                          # UNMODIFIED reference file
        flag = os.path.join(STAGE, "flags_1dsfm.yaml")
        data = os.path.join(STAGE, "datasets", "Madrid_Metropolis")
        t0 = time.perf_counter()
        recon = pipeline.sfm_with_1dsfm_dataset(flag, data, lf.MAGSACWeightBasedLoss(0.02), lf.HuberLoss(0.1),
                                                sfm.RotationErrorType.ANGLE_AXIS_COVARIANCE, sfm.PositionErrorType.BASELINE,
                                                onlyRotationAvg=True)
        t_pipeline = time.perf_counter() - t0
        summary = sfm._solve.last_summary
    finally:
        os.chdir(cwd)
    # the oracle's solve of the same problem (dense Cholesky = the reference's exact factorisation), into a second Reconstruction
    import numpy as np
    from globalsfmpy_b200 import _capi as capi, solver as S, viewgraph as vg
    from oracle import ra_oracle as orc
    g = vg.load_madrid_fixture(os.path.join(ROOT, "tests", "golden", "madrid_metropolis.npz"))
    prob = S.make_problem(g, capi.ANGLE_AXIS_COVARIANCE)
    o = capi.default_options_py()
    o.loss = capi.Loss.make(capi.LOSS_MAGSAC3, 0.02)
    o.num_threads = os.cpu_count()
    t0 = time.perf_counter()
    om_o, s_o, _ = orc.solve(prob, o, g.omega_init)
    t_oracle = time.perf_counter() - t0
    ref = sfm.Reconstruction()
    names = {}
    for v in recon.ViewIds():
        names[v] = recon.View(v).Name()
    # same view ids / names as the pipeline's reconstruction
    ref._views = {v: type(recon.View(v))(names[v]) for v in recon.ViewIds()}
    sfm.SetOrientations({int(v): om_o[k] for k, v in enumerate(g.view_ids.tolist())}, ref)
    common = sfm.FindCommonEstimatedViewsByName(ref, recon)
    info = sfm.compare_orientations(common, ref, recon, 0.0)
    d = np.array(info.rotation_diff_when_align)
    ours = np.array([recon.View(int(v)).GetOrientationAsAngleAxis() for v in g.view_ids.tolist()])
    out = {"script": "scripts/sfm_pipeline.py (unmodified), sfm_with_1dsfm_dataset(flags_1dsfm.yaml, Madrid_Metropolis, MAGSACWeightBasedLoss(0.02), "
                     "HuberLoss(0.1), ANGLE_AXIS_COVARIANCE, onlyRotationAvg=True)",
           "estimated_views": int(sum(recon.View(v).IsEstimated() for v in recon.ViewIds())), "common_camera": info.common_camera,
           "pipeline_wall_s": t_pipeline, "solve_lm_iterations": summary.num_iterations, "solve_final_cost": summary.final_cost,
           "solve_ms_total": summary.ms_total, "n_gpus_used": summary.n_gpus_used,
           "oracle_lm_iterations": s_o.num_iterations, "oracle_final_cost": s_o.final_cost, "oracle_solve_s": t_oracle,
           "reference_metric_compare_orientations": {"mean_rad": float(d.mean()), "median_rad": float(np.median(d)), "max_rad": float(d.max())},
           "chordal_alignment_mean_rad": vg.mean_angular_error(om_o, ours)[0]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
