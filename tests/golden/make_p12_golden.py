#!/usr/bin/env python3
"""Fixture for the reference's 12-parameter camera rows: imports /root/reference/src, builds the
toy-1 BA input with ``input_bundle_adj(domain, rotate_model="R", last_view=2)`` (seq_utils.py:168-169,182-184: rows
``[R|c].reshape(12)``), perturbs R and c (seeded) and records what the reference computes on it -- ``fun`` at two
points, ``project`` (the [R|c] branch, bundle_adjustment.py:43-45), and its own ``run_bundle_adjustment`` (scipy TRF with
finite differences over the 12-column camera pattern).  Writes tests/golden/toy1_lv2_p12.npz.

    python tests/golden/make_p12_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    ref = mg.import_reference()
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import line_ba_oracle as orc
    dom = mg.build_toy_domain(ref, "toy_example1", "toy_example1_3Dvision.json", 3)
    with contextlib.redirect_stdout(io.StringIO()):
        Tr = ref.tracker.Tracker()
        Tr.comp_tracker(dom, last_view=2)
        dom.tracker = Tr
        ref.seq.triangulate_model(dom, last_view=2)
        inp = ref.seq.input_bundle_adj(dom, rotate_model="R", last_view=2)
    rng = np.random.default_rng(31)
    cams = np.array(inp["camera_params"], dtype=np.float64)
    assert cams.shape[1] == 12
    cams = cams + rng.normal(0, 5e-3, cams.shape) * np.tile([1, 1, 1, 2], 3)[None, :]
    inp = dict(inp); inp["camera_params"] = cams
    flat = mg.flatten_input(inp)
    ba = ref.ba
    with contextlib.redirect_stdout(io.StringIO()):
        points_3d, points_2d_p, cam_pt, pt_idx = ba.point_wise_input(inp["Lines_3d"], inp["lines_2d"], inp["camera_indices"],
                                                                     inp["lines_indices"])
    nc, npts = cams.shape[0], points_3d.shape[0]
    x0 = np.hstack((cams.ravel(), points_3d.ravel()))
    args = (nc, npts, cam_pt, pt_idx, points_2d_p, inp["lines_2d"])
    out = dict(flat)
    out["ref_x0"] = x0
    out["ref_fun_x0"] = ba.fun(x0, *args)
    out["ref_project_x0"] = ba.project(x0[12 * nc:].reshape(npts, 3)[pt_idx], x0[:12 * nc].reshape(nc, 3, 4)[cam_pt])
    x1 = x0 + np.random.default_rng(32).normal(0, 2e-2, x0.shape)
    out["ref_x1"] = x1
    out["ref_fun_x1"] = ba.fun(x1, *args)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        opc, opp, res = ba.run_bundle_adjustment(inp, "linear")
    out["ref_ba_op_camera_params"] = opc
    out["ref_ba_op_points_3d"] = opp
    out["ref_ba_residual"] = res
    out["ref_ba_stdout_head"] = np.array("\n".join(buf.getvalue().splitlines()[:6]))
    r = orc.residuals12(cams, points_3d, flat["camera_indices"], flat["lines_indices"], flat["n"], flat["p0"], flat["p1"])
    err = np.max(np.abs(r - out["ref_fun_x0"]) / np.maximum(np.abs(out["ref_fun_x0"]), 1e-300))
    print(f"toy1_lv2_p12: cams {nc} x 12, lines {npts // 2}, obs {len(r)}; fun(x0) cost {0.5 * np.sum(out['ref_fun_x0'] ** 2):.4e}, "
          f"restatement max rel err {err:.2e}; reference BA cost {0.5 * np.sum(res ** 2):.4e}, returned camera rows {opc.shape}")
    np.savez_compressed(os.path.join(HERE, "toy1_lv2_p12.npz"), **out)


if __name__ == "__main__":
    main()
