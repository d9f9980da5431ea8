"""Drop-in replacement for the reference module ``src/bundle_adjustment.py``.

Same public names, argument order and returned arrays as the reference
(``/root/reference/src/bundle_adjustment.py``): ``run_bundle_adjustment`` (:254),
``point_wise_input`` (:194), ``fun`` (:78), ``project`` (:23),
``bundle_adjustment_sparsity`` (:144).  The arithmetic runs in liblba.so on the GPU;
there is no CPU fallback.  Use as::

    from amuya_sfm_lines_beta_b200.bundle_adjustment import run_bundle_adjustment

(or put ``amuya_sfm_lines_beta_b200`` on ``sys.path`` so that the reference drivers'
``from bundle_adjustment import run_bundle_adjustment`` -- src/main.py:13 -- resolves here).
"""
from __future__ import annotations

import time

import numpy as np

try:  # package import, or flat import when this directory itself is on sys.path
    from .solver import LineBA
except ImportError:  # pragma: no cover
    from solver import LineBA

_KNOWN_LOSSES = ("linear", "soft_l1", "huber", "cauchy", "arctan")  # scipy least_squares `loss`
_ctx_cache: dict = {}


def _context(device=0) -> LineBA:
    ba = _ctx_cache.get(device)
    if ba is None:
        ba = _ctx_cache[device] = LineBA(device)
    return ba


def _flatten_lines_2d(lines_2d):
    """(m,3) arrays of Line.n and Line.points[e].p -- the constants ``fun`` re-reads from the
    objects on every call (bundle_adjustment.py:119-130).  One pass per attribute with numpy doing the
    copies (2x faster than per-row assignment; 2 M observations: ~1.3 s, all of it Python attribute access --
    callers that keep arrays should use ``LineBA`` / ``glue.IncrementalBA`` instead)."""
    m = len(lines_2d)
    try:
        n = np.array([l.n for l in lines_2d], dtype=np.float64).reshape(m, 3)
        p0 = np.array([l.points[0].p for l in lines_2d], dtype=np.float64).reshape(m, 3)
        p1 = np.array([l.points[1].p for l in lines_2d], dtype=np.float64).reshape(m, 3)
        return n, p0, p1
    except ValueError:      # rows of different shapes ((3,) next to (3,1)): row by row
        pass
    n = np.empty((m, 3)); p0 = np.empty((m, 3)); p1 = np.empty((m, 3))
    for i, l in enumerate(lines_2d):
        n[i] = np.asarray(l.n, dtype=np.float64).ravel()
        p0[i] = np.asarray(l.points[0].p, dtype=np.float64).ravel()
        p1[i] = np.asarray(l.points[1].p, dtype=np.float64).ravel()
    return n, p0, p1


def _points_3d_of(lines_3d):
    """``points_3d`` of ``point_wise_input`` (bundle_adjustment.py:231-233): rows 2L, 2L+1 = endpoints of line L,
    in the dtype the objects hold (float32 after the reference's triangulation)."""
    points_3d = []
    for L in lines_3d:
        points_3d.append(L.points[0].coord[:3])
        points_3d.append(L.points[1].coord[:3])
    return np.array(points_3d)


def point_wise_input(lines_3d, lines_2d, camera_indices, line_indices):
    """Endpoint-wise expansion of the BA input (reference :194-250): same four outputs,
    same dtypes (``points_3d`` keeps the float32 the reference's triangulation produces)."""
    points_3d = _points_3d_of(lines_3d)
    points_2d_p = np.array([l.points[e].p for l in lines_2d for e in (0, 1)])
    camera_indices = np.asarray(camera_indices)
    line_indices = np.asarray(line_indices)
    camera_indices_pt = np.repeat(camera_indices, 2)
    pt_indices = np.empty(2 * line_indices.size, dtype=line_indices.dtype)
    pt_indices[0::2] = 2 * line_indices
    pt_indices[1::2] = 2 * line_indices + 1
    return points_3d, points_2d_p, camera_indices_pt, pt_indices


def project(points, camera_params_, device=0):
    """Reference :23-74.  ``points`` (N,3), ``camera_params_`` (N,3,2) gathered per row
    (column 0 = Rodrigues vector, column 1 = centre).  Returns unit bearings (N,3)."""
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    cams = np.asarray(camera_params_, dtype=np.float64)
    N = points.shape[0]
    if cams.ndim == 3 and cams.shape[2] == 4:       # [R|c] rows (reference :43-45)
        ba = LineBA(device)
        try:
            return ba.project12(points, cams.reshape(N, 12))
        finally:
            ba.close()
    cams = cams.reshape(N, 6)
    # every row gets its own camera and its own degenerate "3D line" (row, row)
    ba = LineBA(device)
    try:
        ci = np.arange(N, dtype=np.int64)
        dummy = np.tile(np.array([[0.0, 0.0, 1.0]]), (N, 1))
        ba.set_problem(N, N, ci, ci, dummy, dummy, dummy)
        out = ba.project(cams, np.repeat(points, 2, axis=0))[0::2]
    finally:
        ba.close()
    return out


def fun(params, n_cameras, n_points, camera_indices_pt, pt_indices, points_2d_p, lines_2d, device=0):
    """Residual vector (reference :78-139), one scalar per observed 2D line."""
    params = np.asarray(params, dtype=np.float64)
    n_cam_par = int((len(params) - 3 * n_points) / n_cameras)
    if n_cam_par not in (6, 12):
        raise ValueError("camera rows must have 6 (Rodrigues vector | centre) or 12 ([R|c]) parameters")
    cams = params[:n_cameras * n_cam_par].reshape(n_cameras, n_cam_par)
    pts = params[n_cameras * n_cam_par:].reshape(n_points, 3)
    ci = np.asarray(camera_indices_pt)[0::2].astype(np.int64)
    li = (np.asarray(pt_indices)[0::2] // 2).astype(np.int64)
    n, p0, p1 = _flatten_lines_2d(lines_2d)
    ba = LineBA(device)
    try:
        ba.set_problem(n_cameras, n_points // 2, ci, li, n, p0, p1)
        r, _ = ba.residuals(cams, pts)
    finally:
        ba.close()
    return r


def bundle_adjustment_sparsity(n_cameras, n_points, n, m, camera_indices, line_indices):
    """The pattern the reference ships (:144-191), including its structure-block columns
    ``6 Nc + 3 L + s`` (SURVEY.md finding F2).  Unused by this solver (analytic Jacobian);
    kept for API parity."""
    from scipy.sparse import lil_matrix
    n_cam_par = int((n - 3 * n_points) / n_cameras)
    A = lil_matrix((m, n), dtype=int)
    i = np.arange(np.asarray(camera_indices).size)
    for s in range(n_cam_par):
        A[i, np.asarray(camera_indices) * n_cam_par + s] = 1
    for s in range(3):
        A[i, n_cameras * n_cam_par + np.asarray(line_indices) * 3 + s] = 1
    return A


def run_bundle_adjustment(input_bundle_adjustment, loss_f, *, device=0, verbose=2, return_report=False,
                          **options):
    """Line bundle adjustment (reference :254-349): same input dict, same three returned
    arrays ``(op_camera_params (Nc,6), op_points_3d (2Nl,3), residual (m,))``.

    The reference's scipy ``least_squares`` call (:319) is replaced by the GPU
    Levenberg-Marquardt / Schur / PCG solver; keyword-only ``options`` are forwarded to
    ``LineBA.lm_solve`` (``max_iter``, ``ftol``, ``xtol``, ``gtol``, ``pcg_tol``, ``f_scale`` ...).
    ``loss_f`` takes scipy's names ('linear', 'soft_l1', 'huber', 'cauchy', 'arctan') with scipy's
    semantics (rho applied to the squared residual, rows rescaled as in
    ``scale_for_robust_loss_function``).
    """
    if loss_f not in _KNOWN_LOSSES:
        raise ValueError("`loss` must be one of {0} or a callable.".format(_KNOWN_LOSSES))  # scipy's message
    print("--------------------RUNING BUNDLE ADJUSTMENT ---------------------------")
    camera_params = np.asarray(input_bundle_adjustment['camera_params'], dtype=np.float64)
    lines_3d = input_bundle_adjustment['Lines_3d']
    camera_indices = np.asarray(input_bundle_adjustment['camera_indices'])
    line_indices = np.asarray(input_bundle_adjustment['lines_indices'])
    lines_2d = input_bundle_adjustment['lines_2d']
    if camera_params.ndim != 2 or camera_params.shape[1] not in (6, 12):
        raise ValueError("camera_params must be (Nc,6) Rodrigues rows or (Nc,12) [R|c] rows (reference :43-45)")

    points_3d = _points_3d_of(lines_3d)   # (point_wise_input's other outputs are not needed: indexing happens on the device)
    n_cameras = camera_params.shape[0]
    n_points = points_3d.shape[0]
    n_lines = len(lines_3d)
    n = camera_params.shape[1] * n_cameras + 3 * n_points
    m = len(lines_2d)
    print("n_cameras: {}".format(n_cameras))
    print("n_points3D: {}".format(n_points))
    print("n_lines3D: {}".format(n_lines))
    print("Total number of parameters: {}".format(n))
    print("Total number of residuals: {}".format(m))

    nn, p0, p1 = _flatten_lines_2d(lines_2d)
    ba = _context(device)
    ba.set_problem(n_cameras, n_lines, camera_indices.astype(np.int64), line_indices.astype(np.int64), nn, p0, p1)
    t0 = time.time()
    res = ba.lm_solve(camera_params, np.asarray(points_3d, dtype=np.float64), verbose=verbose, loss=loss_f, **options)
    t1 = time.time()
    if verbose:
        print(res.message)
        print("Function evaluations {0}, initial cost {1:.4e}, final cost {2:.4e}.".format(
            res.n_jac + res.n_res, res.cost_init, res.cost))
    print("Optimization took {0:.0f} seconds".format(t1 - t0))
    out = (res.camera_params.reshape(n_cameras, camera_params.shape[1]), res.points_3d.reshape(n_points, 3), np.copy(res.residuals))
    return out + (res,) if return_report else out
