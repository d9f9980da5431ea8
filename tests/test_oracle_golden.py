"""CPU: the numpy oracle against the fixtures generated from the reference itself
(tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from oracle import line_ba_oracle as orc
from conftest import GOLDEN, golden_names

EPS = np.finfo(np.float64).eps


def residual_tol(r_ref):
    """1e-9 relative, plus the floor any fp64 evaluation has: r = sum (2a)^2 + b^2 with a, b dot
    products of unit vectors carrying ~eps absolute error, so |dr| <~ 8 eps sqrt(r)."""
    return 1e-9 * np.abs(r_ref) + 16 * EPS * np.sqrt(np.abs(r_ref)) + 64 * EPS * EPS


@pytest.mark.parametrize("name", golden_names())
def test_indexing_bit_exact(name):
    q = np.load(os.path.join(GOLDEN, name + ".npz"))
    pts, cam_pt, pt_idx = orc.point_wise_input_arrays(q["points_3d_raw"], q["camera_indices"], q["lines_indices"])
    assert np.array_equal(cam_pt, q["ref_camera_indices_pt"])
    assert np.array_equal(pt_idx, q["ref_pt_indices"])
    assert cam_pt.dtype == q["ref_camera_indices_pt"].dtype and pt_idx.dtype == q["ref_pt_indices"].dtype
    assert np.array_equal(pts, q["ref_points_3d"].astype(np.float64))
    # x0 packing (bundle_adjustment.py:311)
    assert np.array_equal(orc._x_pack(q["camera_params"], q["points_3d_raw"]), q["ref_x0"])
    # observation order is track-major, view-ascending (seq_utils.py:190)
    li, ci = q["lines_indices"], q["camera_indices"]
    assert np.all(np.diff(li) >= 0)
    same = np.diff(li) == 0
    assert np.all(np.diff(ci)[same] > 0)


@pytest.mark.parametrize("name", golden_names())
def test_residuals_match_reference_fun(name):
    q = np.load(os.path.join(GOLDEN, name + ".npz"))
    nc = q["camera_params"].shape[0]
    for xk, fk in (("ref_x0", "ref_fun_x0"), ("ref_x1", "ref_fun_x1")):
        x = q[xk]
        r = orc.residuals(x[:6 * nc].reshape(nc, 6), x[6 * nc:].reshape(-1, 3), q["camera_indices"],
                          q["lines_indices"], q["n"], q["p0"], q["p1"])
        assert np.all(np.abs(r - q[fk]) <= residual_tol(q[fk]))


@pytest.mark.parametrize("name", golden_names())
def test_project_matches_reference(name):
    q = np.load(os.path.join(GOLDEN, name + ".npz"))
    om, ce = orc.split_camera_params(q["camera_params"])
    pts = q["ref_points_3d"].astype(np.float64)
    cam_pt, pt_idx = q["ref_camera_indices_pt"], q["ref_pt_indices"]
    ph = orc.project(pts[pt_idx], om[cam_pt], ce[cam_pt])
    assert np.max(np.abs(ph - q["ref_project_x0"])) < 1e-14


@pytest.mark.parametrize("name", golden_names())
def test_sparsity_columns_as_shipped(name):
    q = np.load(os.path.join(GOLDEN, name + ".npz"))
    nc = q["camera_params"].shape[0]
    cols = orc.bundle_adjustment_sparsity_columns(nc, q["camera_indices"], q["lines_indices"])
    indptr, indices = q["ref_sparsity_indptr"], q["ref_sparsity_indices"]
    for i in range(cols.shape[0]):
        assert np.array_equal(np.sort(cols[i]), indices[indptr[i]:indptr[i + 1]])


@pytest.mark.parametrize("name", ["toy1_lv2", "toy1_lv5", "toy2_lv2"])
def test_analytic_jacobian_vs_reference_fd(name):
    q = np.load(os.path.join(GOLDEN, name + ".npz"))
    nc = q["camera_params"].shape[0]
    x1 = q["ref_x1"]
    om, ce = orc.split_camera_params(x1[:6 * nc].reshape(nc, 6))
    pts = x1[6 * nc:].reshape(-1, 6)
    ci, li, n = q["camera_indices"], q["lines_indices"], q["n"]
    r, jc, jl = orc.residuals_jacobian(om, ce, pts, ci, li, n, np.cross(n, q["p0"]), np.cross(n, q["p1"]))
    J = orc.dense_jacobian_reference_layout(jc, jl, ci, li, nc, pts.shape[0])
    assert ((J != 0).sum(axis=1) <= 12).all()
    Ja = J[np.ix_(q["ref_fd_rows"], q["ref_fd_cols"])]
    assert np.max(np.abs(Ja - q["ref_fd_jac_x1"])) < 1e-7 * max(1.0, np.abs(Ja).max())
    # no non-zero outside the columns the FD probe covered
    mask = np.ones(J.shape[1], bool); mask[q["ref_fd_cols"]] = False
    assert np.all(J[np.ix_(q["ref_fd_rows"], np.nonzero(mask)[0])] == 0)


def test_rodrigues_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    w = np.vstack([rng.normal(0, 1.0, (200, 3)), np.zeros((1, 3)), rng.normal(0, 1e-9, (5, 3)),
                   rng.normal(0, 1e-3, (5, 3))])
    R = orc.rodrigues(w)
    for k in range(w.shape[0]):
        assert np.max(np.abs(R[k] - cv2.Rodrigues(w[k])[0])) < 4e-15


def test_right_jacobian_against_finite_differences():
    rng = np.random.default_rng(1)
    for scale in (1.0, 0.05, 1e-4, 0.0):
        w = rng.normal(0, 1, 3) * scale
        d = rng.normal(0, 1, 3)
        G = orc.so3_right_jacobian(w[None])[0]
        R = orc.rodrigues(w[None])[0]
        dK = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
        ana = -R @ dK @ G
        h = 1e-6
        fd = np.stack([(orc.rodrigues((w + h * e)[None])[0] @ d - orc.rodrigues((w - h * e)[None])[0] @ d) / (2 * h)
                       for e in np.eye(3)], axis=1)
        assert np.max(np.abs(ana - fd)) < 1e-8


def test_table_identity_for_camera_dot():
    """jc_i . v == sum_e jl_e . (a x X_e - h) with a = G v_w, h = a x c + v_c -- the identity the
    line-ordered PCG kernel relies on (it never reads jc)."""
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    s = make_synthetic_scene(7, 40, 4, seed=5)
    om, ce = orc.split_camera_params(s.camera_params)
    pts = s.points_3d.reshape(-1, 6)
    ci, li = s.camera_indices, s.lines_indices
    r, jc, jl = orc.residuals_jacobian(om, ce, pts, ci, li, s.n, np.cross(s.n, s.p0), np.cross(s.n, s.p1))
    v = np.random.default_rng(2).normal(size=(7, 6))
    G = orc.so3_right_jacobian(om)
    a = np.einsum('cij,cj->ci', G, v[:, :3])
    h = np.cross(a, ce) + v[:, 3:]
    s_tab = sum(np.einsum('ni,ni->n', jl[:, 3 * e:3 * e + 3], np.cross(a[ci], pts[li, 3 * e:3 * e + 3]) - h[ci])
                for e in (0, 1))
    s_dir = np.einsum('ni,ni->n', jc, v[ci])
    assert np.max(np.abs(s_tab - s_dir)) < 1e-12 * max(1.0, np.abs(s_dir).max())


def test_schur_pcg_solves_the_damped_normal_equations():
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    s = make_synthetic_scene(6, 30, 4, seed=9)
    om, ce = orc.split_camera_params(s.camera_params)
    pts = s.points_3d.reshape(-1, 6)
    ci, li = s.camera_indices, s.lines_indices
    r, jc, jl = orc.residuals_jacobian(om, ce, pts, ci, li, s.n, np.cross(s.n, s.p0), np.cross(s.n, s.p1))
    nc, nl = 6, 30
    U, gc, V, gl = orc.normal_equations(r, jc, jl, ci, li, nc, nl)
    idx = np.arange(6)
    Dc = orc.update_scaling(None, U[:, idx, idx]); Dl = orc.update_scaling(None, V[:, idx, idx])
    lam = 1e-3
    sysm = orc.SchurSystem(r, jc, jl, ci, li, U, gc, V, gl, Dc, Dl, lam)
    dc, q = sysm.pcg(1e-12, 500)
    dl, jd = sysm.back_substitute(dc)
    # dense check
    J = np.zeros((len(r), 6 * nc + 6 * nl))
    for i in range(len(r)):
        J[i, 6 * ci[i]:6 * ci[i] + 6] = jc[i]
        J[i, 6 * nc + 6 * li[i]:6 * nc + 6 * li[i] + 6] = jl[i]
    H = J.T @ J + lam * np.diag(np.concatenate([Dc.ravel(), Dl.ravel()]))
    ref = np.linalg.solve(H, -J.T @ r)
    got = np.concatenate([dc.ravel(), dl.ravel()])
    assert np.max(np.abs(got - ref)) < 1e-8 * np.max(np.abs(ref))
    assert np.allclose(jd, J @ got, rtol=1e-10, atol=1e-14)


def test_lm_oracle_beats_the_reference_as_shipped():
    """P5: on the scene where the reference's own run_bundle_adjustment output was recorded, the LM
    restatement ends below the reference's final cost (the as-shipped pattern stalls, finding F4)."""
    q = np.load(os.path.join(GOLDEN, "toy1_lv2.npz"))
    ref_cost = 0.5 * float(np.sum(q["ref_ba_residual"] ** 2))
    out = orc.lm_solve(q["camera_params"], q["points_3d_raw"].astype(np.float64), q["camera_indices"],
                       q["lines_indices"], q["n"], q["p0"], q["p1"], opts=orc.LMOptions(max_iter=15))
    assert out[3].cost < ref_cost
    assert out[3].cost < 0.05 * out[3].cost_init
    # returned layouts: (Nc,6) interleaved, (2Nl,3), (m,)
    assert out[0].shape == q["ref_ba_op_camera_params"].shape
    assert out[1].shape == q["ref_ba_op_points_3d"].shape
    assert out[2].shape == q["ref_ba_residual"].shape


def test_reference_port_reproduces_as_shipped_behaviour():
    """Finding F2: with the shipped sparsity pattern only the first half of the endpoint rows can
    move.  The vectorised port of the reference call shows the same signature as the recorded run."""
    q = np.load(os.path.join(GOLDEN, "toy1_lv2.npz"))
    pts0 = q["points_3d_raw"].astype(np.float64)
    moved_ref = np.any(q["ref_ba_op_points_3d"] != pts0, axis=1)
    op_cam, op_pts, res, sres, dt = orc.scipy_reference_port(
        q["camera_params"], pts0, q["camera_indices"], q["lines_indices"], q["n"], q["p0"], q["p1"])
    moved = np.any(op_pts != pts0, axis=1)
    half = pts0.shape[0] // 2
    assert moved_ref[:half].all() and not moved_ref[half:].any()
    assert moved[:half].all() and not moved[half:].any()
    ref_cost = 0.5 * float(np.sum(q["ref_ba_residual"] ** 2))
    assert abs(sres.cost - ref_cost) < 0.2 * ref_cost


@pytest.mark.parametrize("loss", ["soft_l1", "huber", "cauchy", "arctan"])
def test_robust_rho_matches_scipy(loss):
    """rho, rho', rho'' against scipy's own implementations (the reference forwards loss_f to scipy,
    bundle_adjustment.py:319) -- this pins the robust-loss arithmetic the CUDA kernel restates."""
    import importlib
    lsq = importlib.import_module("scipy.optimize._lsq.least_squares")
    z = np.concatenate([np.linspace(0, 3, 61), [1.0, 1.0 + 1e-12, 25.0, 1e4]])
    rho = np.empty((3, z.size))
    lsq.IMPLEMENTED_LOSSES[loss](z.copy(), rho, cost_only=False)
    r0, r1, r2 = orc.robust_rho(loss, z)
    assert np.allclose(r0, rho[0], rtol=1e-14, atol=0) and np.allclose(r1, rho[1], rtol=1e-14) and np.allclose(r2, rho[2], rtol=1e-14)
    # the scaled system reproduces the gradient of the robust cost: J~^T f~ = J^T (rho' f)
    rng = np.random.default_rng(0)
    f = rng.normal(0, 1.5, 50); J = rng.normal(size=(50, 6))
    ft, Jt, _ = orc.scale_for_robust_loss(f, J, J, loss)
    assert np.allclose(Jt.T @ ft, J.T @ (orc.robust_rho(loss, f * f)[1] * f), rtol=1e-12)
    # and scipy's scale_for_robust_loss_function gives the same rows / residuals
    from scipy.optimize._lsq.common import scale_for_robust_loss_function
    rho3 = np.stack(orc.robust_rho(loss, f * f))
    Js, fs = scale_for_robust_loss_function(J.copy(), f.copy(), rho3)
    assert np.allclose(Js, Jt, rtol=1e-13) and np.allclose(fs, ft, rtol=1e-13)


def test_robust_lm_downweights_outliers():
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    s = make_synthetic_scene(10, 120, 5, seed=3)
    rng = np.random.default_rng(1)
    bad = rng.choice(s.n_obs, 30, replace=False)           # gross outliers: wrong 2D lines
    s.n[bad] = np.roll(s.n[bad], 1, axis=1); s.p0[bad] = np.roll(s.p0[bad], 1, axis=1); s.p1[bad] = np.roll(s.p1[bad], 1, axis=1)
    lin = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(max_iter=15))
    rob = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(max_iter=15, loss="cauchy", f_scale=1e-3))
    good = np.setdiff1d(np.arange(s.n_obs), bad)
    assert rob[3].cost < rob[3].cost_init
    assert np.sum(rob[2][good] ** 2) < np.sum(lin[2][good] ** 2)   # inliers fit better with the robust loss


def test_p5_oracle_lm_ends_below_the_reference_on_every_fixture():
    """Default options (max_iter 50, ftol 1e-4): final cost below what the reference's scipy call reaches as
    shipped (tests/golden/p5_costs.json, generated by tests/golden/make_p5_costs.py)."""
    import json
    from amuya_sfm_lines_beta_b200.scenes import scene_from_npz, make_synthetic_scene
    p5 = json.load(open(os.path.join(GOLDEN, "p5_costs.json")))["fixtures"]
    for name, ref in p5.items():
        s = make_synthetic_scene(12, 200, 5, seed=1) if name.startswith("synth") else scene_from_npz(os.path.join(GOLDEN, name + ".npz"))
        rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions())[3]
        assert abs(rep.cost_init - ref["cost_init"]) <= 1e-9 * ref["cost_init"]
        assert rep.cost < ref["as_shipped_cost"], (name, rep.cost, ref)


# ---------------------------------------------------------------- 12-parameter [R|c] camera rows
def test_p12_residuals_match_reference_fun():
    """The restatement of ``fun`` with (Nc,12) camera rows against the reference's own outputs at two points
    (fixture made by tests/golden/make_p12_golden.py from input_bundle_adj(rotate_model != "Rodrigues"))."""
    q = np.load(os.path.join(GOLDEN, "toy1_lv2_p12.npz"))
    nc = q["camera_params"].shape[0]
    assert q["camera_params"].shape[1] == 12
    for key in ("x0", "x1"):
        x = q["ref_" + key]
        r = orc.residuals12(x[:12 * nc].reshape(nc, 12), x[12 * nc:].reshape(-1, 3), q["camera_indices"], q["lines_indices"],
                            q["n"], q["p0"], q["p1"])
        ref = q["ref_fun_" + key]
        assert np.all(np.abs(r - ref) <= 1e-9 * np.abs(ref) + 16 * np.finfo(float).eps * np.sqrt(np.abs(ref)))


def test_p12_jacobian_and_solver():
    """Analytic 12-parameter Jacobian against central differences of the restated fun; the dense LM ends below the
    reference's own run_bundle_adjustment on the same input (recorded residual vector)."""
    q = np.load(os.path.join(GOLDEN, "toy1_lv2_p12.npz"))
    cams, pts = q["camera_params"], q["points_3d_raw"].astype(np.float64)
    a = (q["camera_indices"], q["lines_indices"], q["n"], q["p0"], q["p1"])
    r, jc, jl = orc.residuals_jacobian12(cams, pts, *a)
    rng = np.random.default_rng(0)
    for i in rng.integers(0, len(r), 6):
        cam, line = a[0][i], a[1][i]
        for k in range(12):
            h = 1e-6
            cp = cams.copy(); cp[cam, k] += h
            cm = cams.copy(); cm[cam, k] -= h
            fd = (orc.residuals12(cp, pts, *a)[i] - orc.residuals12(cm, pts, *a)[i]) / (2 * h)
            assert abs(fd - jc[i, k]) <= 1e-6 * max(1.0, np.abs(jc).max())
        for k in range(6):
            h = 1e-6
            pp = pts.copy().reshape(-1, 6); pp[line, k] += h
            pm = pts.copy().reshape(-1, 6); pm[line, k] -= h
            fd = (orc.residuals12(cams, pp, *a)[i] - orc.residuals12(cams, pm, *a)[i]) / (2 * h)
            assert abs(fd - jl[i, k]) <= 1e-6 * max(1.0, np.abs(jl).max())
    out = orc.lm_solve12(cams, pts, *a, opts=orc.LMOptions(max_iter=15))
    ref_cost = 0.5 * float(np.sum(q["ref_ba_residual"] ** 2))
    assert out[3].cost < ref_cost
    assert out[0].shape == (3, 12)
