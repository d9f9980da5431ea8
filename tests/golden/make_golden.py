#!/usr/bin/env python3
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The script imports the reference's own modules (``/root/reference/src``) with the
three shims SURVEY.md section 8c lists (matplotlib / mpl_toolkits / cvxpnpl stubs,
``np.int`` and ``np.float_`` aliases), builds BA inputs from the toy JSON fixtures
through the reference's own ``Tracker`` -> ``triangulate_model`` ->
``input_bundle_adj``, and records what the reference computes:

* ``point_wise_input`` outputs (integer indexing, bit-exact pin),
* ``fun`` at x0 and at a seeded random perturbation,
* ``project`` at x0,
* ``bundle_adjustment_sparsity`` non-zero columns,
* a central finite-difference Jacobian of the reference ``fun`` (subset of rows
  for the larger scenes),
* the outputs of ``run_bundle_adjustment`` as shipped (one small scene).

Nothing from the reference's sources is copied; only its outputs are stored.
"""
import io
import json
import os
import sys
import types
import contextlib

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "cvxpnpl"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__dict__.update(figure=lambda *a, **k: None, plot=lambda *a, **k: None,
                                Axes3D=object, pnl=None)
            sys.modules[name] = mod
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    sys.path.insert(0, os.path.join(REF, "src"))
    import bundle_adjustment as ref_ba
    import seq_utils
    import tracker
    import domain as ref_domain
    import camera as ref_camera
    import view as ref_view
    import point as ref_point
    import line as ref_line
    return types.SimpleNamespace(ba=ref_ba, seq=seq_utils, tracker=tracker, domain=ref_domain,
                                 camera=ref_camera, view=ref_view, point=ref_point, line=ref_line)


def build_toy_domain(ref, toy_dir, json_name, n_views):
    """Mirror of the reference drivers' domain construction
    (main_toy_example1.py:27-103) with the ground-truth cameras from the JSON."""
    with open(os.path.join(REF, "input", toy_dir, json_name)) as fp:
        dd = json.load(fp)
    K = np.array(dd["K"])
    views, points2d, lines2d = [], [], []
    for i in list(dd["views"])[:n_views]:
        gv = dd["views"][i]
        V = ref.view.View()
        V.id = i
        V.name = "pic_{}.png".format(i)
        V.image = np.zeros((2048, 2048, 3), dtype=np.uint8)  # only .shape is read (domain.py:340)
        pv, lv = [], []
        for j, l in enumerate(gv["lines2d"]):
            pt0 = ref.point.Point(); pt0.id = gv["pts2d"][2 * j][0]; pt0.coord = gv["pts2d"][2 * j][1]
            pt1 = ref.point.Point(); pt1.id = gv["pts2d"][2 * j + 1][0]; pt1.coord = gv["pts2d"][2 * j + 1][1]
            pv += [pt0, pt1]
            L = ref.line.Line(); L.id = l[0]; L.local_id = j; L.points = [pt0, pt1]; L.comp_angle()
            lv.append(L)
        V.lines, V.points = lv, pv
        P = ref.camera.Camera(np.array(gv["P"]))
        P.factor(); P.center()
        V.P = P
        views.append(V); points2d += pv; lines2d += lv
    dom = ref.domain.Domain(points2d, lines2d, views, K)
    line_match = {}
    for step in (1, 2):
        for i in range(len(dom.views) - step):
            tup = (int(dom.views[i].id), int(dom.views[i + step].id))
            ids = [[dom.views[i].lines[j].id, dom.views[i + step].lines[j].id]
                   for j in range(len(dom.views[i].lines))]
            line_match[dom.views[i].id + dom.views[i + step].id] = [tup, ids]
    dom.line_match = line_match
    dom.spherical_representation()
    return dom


def ba_input_from_domain(ref, dom, last_view):
    with contextlib.redirect_stdout(io.StringIO()):
        Tr = ref.tracker.Tracker()
        Tr.comp_tracker(dom, last_view=last_view)
        dom.tracker = Tr
        ref.seq.triangulate_model(dom, last_view=last_view)
        return ref.seq.input_bundle_adj(dom, rotate_model="Rodrigues", last_view=last_view)


def flatten_input(inp):
    lines_2d = inp["lines_2d"]
    n = np.array([l.n for l in lines_2d], dtype=np.float64)
    p0 = np.array([l.points[0].p for l in lines_2d], dtype=np.float64)
    p1 = np.array([l.points[1].p for l in lines_2d], dtype=np.float64)
    pts_raw = np.array([np.asarray(L.points[e].coord)[:3] for L in inp["Lines_3d"] for e in (0, 1)])
    return dict(camera_params=np.array(inp["camera_params"], dtype=np.float64),
                points_3d_raw=pts_raw,                       # float32 (seq_utils.py:114-122)
                camera_indices=np.asarray(inp["camera_indices"]),
                lines_indices=np.asarray(inp["lines_indices"]),
                n=n, p0=p0, p1=p1)


def perturb_dict(inp, seed, sig_c=1e-2, sig_w=5e-3):
    """Seeded perturbation of the GT cameras."""
    rng = np.random.default_rng(seed)
    cp = np.array(inp["camera_params"], dtype=np.float64).reshape(-1, 3, 2)
    cp[:, :, 0] += rng.normal(0, sig_w, cp[:, :, 0].shape)
    cp[:, :, 1] += rng.normal(0, sig_c, cp[:, :, 1].shape)
    out = dict(inp)
    out["camera_params"] = cp.reshape(-1, 6)
    return out


def reference_quantities(ref, inp, seed, fd_rows=None, run_ba=False):
    ba = ref.ba
    flat = flatten_input(inp)
    cams = flat["camera_params"]
    with contextlib.redirect_stdout(io.StringIO()):
        points_3d, points_2d_p, cam_pt, pt_idx = ba.point_wise_input(
            inp["Lines_3d"], inp["lines_2d"], inp["camera_indices"], inp["lines_indices"])
    nc, npts = cams.shape[0], points_3d.shape[0]
    x0 = np.hstack((cams.ravel(), points_3d.ravel()))
    args = (nc, npts, cam_pt, pt_idx, points_2d_p, inp["lines_2d"])
    out = dict(flat)
    out["ref_points_3d"] = points_3d
    out["ref_points_3d_dtype"] = np.array(str(points_3d.dtype))
    out["ref_camera_indices_pt"] = cam_pt
    out["ref_pt_indices"] = pt_idx
    out["ref_x0"] = x0
    out["ref_fun_x0"] = ba.fun(x0, *args)
    out["ref_project_x0"] = ba.project(x0[6 * nc:].reshape(npts, 3)[pt_idx],
                                       x0[:6 * nc].reshape(nc, 3, 2)[cam_pt])
    rng = np.random.default_rng(seed + 1000)
    x1 = x0 + rng.normal(0, 2e-2, x0.shape)
    out["ref_x1"] = x1
    out["ref_fun_x1"] = ba.fun(x1, *args)
    m = len(inp["lines_2d"])
    A = ba.bundle_adjustment_sparsity(nc, npts, x0.size, m, inp["camera_indices"], inp["lines_indices"])
    A = A.tocsr(); A.sort_indices()
    out["ref_sparsity_indptr"] = A.indptr
    out["ref_sparsity_indices"] = A.indices
    if fd_rows is not None:
        # central FD of the reference fun at x1, restricted to the columns that can be non-zero
        rows = np.arange(m) if isinstance(fd_rows, str) else np.asarray(fd_rows)
        cols = set()
        for i in rows:
            c, L = int(inp["camera_indices"][i]), int(inp["lines_indices"][i])
            cols.update(range(6 * c, 6 * c + 6)); cols.update(range(6 * nc + 6 * L, 6 * nc + 6 * L + 6))
        cols = np.array(sorted(cols))
        Jfd = np.zeros((rows.size, cols.size))
        for jj, col in enumerate(cols):
            h = 1e-6
            xp = x1.copy(); xp[col] += h
            xm = x1.copy(); xm[col] -= h
            Jfd[:, jj] = (ba.fun(xp, *args)[rows] - ba.fun(xm, *args)[rows]) / (2 * h)
        out["ref_fd_rows"], out["ref_fd_cols"], out["ref_fd_jac_x1"] = rows, cols, Jfd
    if run_ba:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            opc, opp, res = ba.run_bundle_adjustment(inp, "linear")
        out["ref_ba_op_camera_params"] = opc
        out["ref_ba_op_points_3d"] = opp
        out["ref_ba_residual"] = res
        out["ref_ba_stdout_head"] = np.array("\n".join(buf.getvalue().splitlines()[:6]))
    return out


def manual_annotation_scene(ref, seed=7):
    """Config-3 stand-in (the natural JPEGs are missing, so LSD + matching cannot run): natural-image
    SIZE and 2D-segment statistics from the 3 x 234 hand-annotated segments
    (input/manual_line_annotation/*_csv.csv), K from main.py:24, bearings exactly as
    Domain.spherical_representation computes them (domain.py:338-351).
    View 0 uses the real annotated segments; their 3D endpoints are back-projected at seeded depths;
    views 1-2 (true poses unknown without the pixels) observe those 3D segments from a seeded rig with
    endpoint sliding and 1e-3 bearing noise.  Initial cameras/endpoints are seeded perturbations."""
    import csv
    K = np.array([[3.47133831e+03, 0, 2.81264921e+03], [0, 3.46595620e+03, 1.84437179e+03], [0, 0, 1]])
    rows = []
    with open(os.path.join(REF, "input", "manual_line_annotation", "DSC_0342_csv.csv")) as fp:
        for row in csv.DictReader(fp):
            sh = json.loads(row["region_shape_attributes"])
            if "all_points_x" in sh and len(sh["all_points_x"]) >= 2:
                rows.append([sh["all_points_x"][0], sh["all_points_y"][0], sh["all_points_x"][1], sh["all_points_y"][1]])
    seg = np.array(rows, dtype=np.float64)
    cen = K[:2, 2]
    q0 = np.concatenate((seg[:, 0:2] - cen, np.full((len(seg), 1), K[0, 0])), axis=1)
    q1 = np.concatenate((seg[:, 2:4] - cen, np.full((len(seg), 1), K[0, 0])), axis=1)
    q0 /= np.linalg.norm(q0, axis=1)[:, None]; q1 /= np.linalg.norm(q1, axis=1)[:, None]
    keep = np.linalg.norm(np.cross(q0, q1), axis=1) > 1e-6
    q0, q1 = q0[keep], q1[keep]
    nl = len(q0)
    rng = np.random.default_rng(seed)
    depth = rng.uniform(6, 12, nl)
    X0 = q0 * (depth / q0[:, 2])[:, None]
    X1 = q1 * ((depth + rng.normal(0, 0.3, nl)) / q1[:, 2])[:, None]
    cams = np.zeros((3, 3, 2))
    cams[1, :, 0] = [0.02, -0.12, 0.01]; cams[1, :, 1] = [1.2, 0.05, 0.1]
    cams[2, :, 0] = [-0.01, -0.25, 0.02]; cams[2, :, 1] = [2.4, -0.05, 0.3]
    import cv2
    n_l, p0_l, p1_l, ci, li = [], [], [], [], []
    for L in range(nl):
        for v in range(3):
            if v == 0:
                a, b = q0[L], q1[L]
            else:
                R = cv2.Rodrigues(cams[v, :, 0])[0]
                c = cams[v, :, 1]
                s0, s1 = rng.uniform(-0.1, 0.1, 2)
                dX = X1[L] - X0[L]
                a = R @ (X0[L] + s0 * dX - c); b = R @ (X1[L] - s1 * dX - c)
                a = a / np.linalg.norm(a) + rng.normal(0, 1e-3, 3); b = b / np.linalg.norm(b) + rng.normal(0, 1e-3, 3)
                a /= np.linalg.norm(a); b /= np.linalg.norm(b)
            nn = np.cross(a, b)
            n_l.append(nn / np.linalg.norm(nn)); p0_l.append(a); p1_l.append(b); ci.append(v); li.append(L)
    cams0 = cams.copy()
    cams0[:, :, 0] += rng.normal(0, 5e-3, (3, 3)); cams0[:, :, 1] += rng.normal(0, 1e-2, (3, 3))
    pts = np.empty((2 * nl, 3)); pts[0::2] = X0; pts[1::2] = X1
    pts = (pts + rng.normal(0, 2e-2, pts.shape)).astype(np.float32)   # float32 like the reference's triangulation
    return dict(camera_params=cams0.reshape(3, 6), points_3d_raw=pts, camera_indices=np.array(ci, dtype=np.int64),
                lines_indices=np.array(li, dtype=np.int64), n=np.array(n_l), p0=np.array(p0_l), p1=np.array(p1_l))


def main():
    ref = import_reference()
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import line_ba_oracle as orc

    specs = [
        # name, toy dir, json, last_view, seed, fd_rows, run_ba
        ("toy1_lv2", "toy_example1", "toy_example1_3Dvision.json", 2, 11, "all", True),
        ("toy1_lv5", "toy_example1", "toy_example1_3Dvision.json", 5, 12, np.arange(0, 774, 37), False),
        ("toy1_lv9", "toy_example1", "toy_example1_3Dvision.json", 9, 13, None, False),
        ("toy2_lv2", "toy_example2", "toy_example2_3Dvision.json", 2, 21, np.arange(0, 651, 31), False),
    ]
    if "--only-manual" in sys.argv:
        specs = []
    for name, tdir, jname, lv, seed, fd_rows, run_ba in specs:
        dom = build_toy_domain(ref, tdir, jname, lv + 1)
        inp = perturb_dict(ba_input_from_domain(ref, dom, lv), seed)
        q = reference_quantities(ref, inp, seed, fd_rows=fd_rows, run_ba=run_ba)
        # sanity: the restatement agrees with the reference before we commit the fixture
        r = orc.residuals(q["camera_params"], q["ref_points_3d"], q["camera_indices"],
                          q["lines_indices"], q["n"], q["p0"], q["p1"])
        err = np.max(np.abs(r - q["ref_fun_x0"]) / np.maximum(np.abs(q["ref_fun_x0"]), 1e-300))
        print(f"{name}: cams {q['camera_params'].shape[0]} lines {q['ref_points_3d'].shape[0] // 2} "
              f"obs {q['n'].shape[0]}  fun(x0) cost {0.5 * np.sum(q['ref_fun_x0'] ** 2):.4e} "
              f"restatement max rel err {err:.2e}")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **q)
    man = manual_annotation_scene(ref)
    np.savez_compressed(os.path.join(HERE, "manual3.npz"), **man)
    print("manual3: cams 3 lines", man["points_3d_raw"].shape[0] // 2, "obs", man["n"].shape[0])


if __name__ == "__main__":
    main()
