#!/usr/bin/env python3
"""Measurements for SURVEY.md 8f row N1 (run on the GPU box; prints one JSON document):

(a) the dict API at BASELINE config [3] size: ``run_bundle_adjustment(dict)`` on 2 M ``Line`` / 200 k ``Line3D`` Python
    objects (the reference's own boundary, src/bundle_adjustment.py:254) against the array API on the same scene --
    how much of a call is the walk over the Python objects;
(b) the incremental loop of src/main.py:111-133 on the toy-1 domain (6 views, 774 2D lines): per added view, the
    device-resident ``IncrementalBA.step`` (objects flattened once, tables / triangulation / observation constants
    stay on the device) against rebuilding and re-uploading the BA input from the objects every time.

    python tests/measure_incremental.py [--obs-scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _P:
    __slots__ = ("n", "points", "p", "coord")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--obs-scale", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import io
    import contextlib
    import torch  # noqa: F401  (CUDA context warm-up happens inside LineBA)
    from amuya_sfm_lines_beta_b200 import LineBA
    from amuya_sfm_lines_beta_b200.bundle_adjustment import run_bundle_adjustment
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    out = {}

    # ---- (a) dict API at 2 M observations
    s = make_synthetic_scene(int(1000 * args.obs_scale) or 10, int(200_000 * args.obs_scale), 10, seed=0)
    t0 = time.perf_counter()
    lines_2d = [_P(n=s.n[i], points=[_P(p=s.p0[i]), _P(p=s.p1[i])]) for i in range(s.n_obs)]
    lines_3d = [_P(points=[_P(coord=s.points_3d[2 * L]), _P(coord=s.points_3d[2 * L + 1])]) for L in range(s.n_lines)]
    t_build = time.perf_counter() - t0
    inp = {"camera_params": s.camera_params, "Lines_3d": lines_3d, "camera_indices": s.camera_indices,
           "lines_indices": s.lines_indices, "lines_2d": lines_2d}
    kw = dict(max_iter=args.iters, ftol=0.0, xtol=0.0, gtol=0.0)
    with contextlib.redirect_stdout(io.StringIO()):
        run_bundle_adjustment(inp, "linear", verbose=0, max_iter=2)           # warm (context, allocations)
        t0 = time.perf_counter()
        opc, opp, res, rep = run_bundle_adjustment(inp, "linear", verbose=0, return_report=True, **kw)
        t_dict = time.perf_counter() - t0
    ba = LineBA(0)
    ba.set_problem(s.n_cameras, s.n_lines, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1)
    ba.lm_solve(s.camera_params, s.points_3d, max_iter=2)
    t0 = time.perf_counter()
    ba.set_problem(s.n_cameras, s.n_lines, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1)
    arr = ba.lm_solve(s.camera_params, s.points_3d, **kw)
    t_arr = time.perf_counter() - t0
    ba.close()
    assert arr.cost == rep.cost
    out["dict_api_2M"] = {"observations": int(s.n_obs), "lm_iterations": rep.iterations,
                          "build_python_objects_s (not part of a call)": round(t_build, 2),
                          "run_bundle_adjustment(dict)_s": round(t_dict, 3), "array_api_s (pageable numpy arrays)": round(t_arr, 3),
                          "object_walk_share": round(1.0 - t_arr / t_dict, 3),
                          "iterations_per_s_dict_api": round(rep.iterations / t_dict, 1),
                          "iterations_per_s_array_api": round(arr.iterations / t_arr, 1)}

    # ---- (b) toy incremental loop
    from test_pipeline_gpu import fake_domain, npz
    from amuya_sfm_lines_beta_b200.glue import (IncrementalBA, fundamental_from_cameras, observations_from_tables,
                                                lines2d_arrays_from_domain)
    trk, tri = npz("trk_toy1_lv5"), npz("tri_toy1_lv5")
    dom = fake_domain(trk, tri)
    inc = IncrementalBA(0)
    inc.step(dom, 1, write_back=False, max_iter=10)                            # warm
    rows = []
    host = LineBA(0)
    import cv2
    for last_view in range(1, 6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        inc.step(dom, last_view, write_back=False, max_iter=10)
        t_step = time.perf_counter() - t0
        # the same stages with the BA input rebuilt from the objects and re-uploaded (round-1 glue)
        t0 = time.perf_counter()
        n, p0, p1, E = lines2d_arrays_from_domain(dom)
        F = np.array([fundamental_from_cameras(dom.K, dom.views[v].P, dom.views[v + 1].P) for v in range(last_view)])
        pairs = [np.array(dom.line_match[str(v) + str(v + 1)][1]) for v in range(last_view)]
        ids, bi, n_new = host.track_table(E, pairs, F)
        cam, line, l2d = observations_from_tables(ids, bi)
        tr0 = np.concatenate([[0], np.cumsum(n_new)])
        l0 = np.concatenate([ids[tr0[v]:tr0[v + 1], v] for v in range(last_view)])
        l1 = np.concatenate([ids[tr0[v]:tr0[v + 1], v + 1] for v in range(last_view)])
        X = host.triangulate_lines(tri["P"][:last_view + 1], np.repeat(np.arange(last_view), n_new), E[l0].astype(np.float32),
                                   E[l1].astype(np.float32))
        cams = np.empty((last_view + 1, 3, 2))
        for i in range(last_view + 1):
            cams[i, :, 0] = cv2.Rodrigues(dom.views[i].P.R)[0].ravel(); cams[i, :, 1] = dom.views[i].P.c
        host.set_problem(last_view + 1, X.shape[0], cam, line, n[l2d], p0[l2d], p1[l2d])
        b = host.lm_solve(cams.reshape(-1, 6), X.reshape(-1, 3).astype(np.float64), max_iter=10)
        t_host = time.perf_counter() - t0
        assert inc.last.cost == b.cost
        rows.append({"last_view": last_view, "observations": int(len(cam)), "step_ms": round(1e3 * t_step, 2),
                     "rebuild_and_reupload_ms": round(1e3 * t_host, 2)})
    inc.close(); host.close()
    out["toy_incremental_loop"] = rows
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
