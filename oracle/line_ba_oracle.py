"""fp64 numpy restatement of the line bundle-adjustment hot path (TEST INFRASTRUCTURE).

Part A restates the reference's own arithmetic and indexing
(``/root/reference/src/bundle_adjustment.py``); each function cites the lines it
follows.  It is pinned against outputs of the reference itself
(``tests/golden/*.npz`` made by ``tests/golden/make_golden.py``).

Part B restates the Levenberg-Marquardt / Schur / PCG algorithm that this
repository runs on the GPU *instead of* the reference's call into scipy
``least_squares`` (``bundle_adjustment.py:319``).  It is the trajectory oracle:
the CUDA solver must follow it iteration by iteration.

Part C is a vectorised port of the reference's scipy call, used as the timed
CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``).

Never imported by the product package.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np

LAMB = 2.0  # bundle_adjustment.py:120  (lamb = 2)
DBL_EPSILON = float(np.finfo(np.float64).eps)


# ---------------------------------------------------------------------------
# Part A -- the reference's arithmetic and indexing
# ---------------------------------------------------------------------------

def point_wise_input_arrays(points_3d, camera_indices, line_indices):
    """Index expansion of ``point_wise_input`` (bundle_adjustment.py:194-250).

    ``camera_indices_pt = repeat(camera_indices, 2)`` (:238,:242) and
    ``pt_indices = 2*line + {0,1}`` (:239,:243).  Integer work: bit-exact.
    """
    camera_indices = np.asarray(camera_indices)
    line_indices = np.asarray(line_indices)
    camera_indices_pt = np.repeat(camera_indices, 2)
    pt_indices = np.empty(2 * line_indices.size, dtype=line_indices.dtype)
    pt_indices[0::2] = 2 * line_indices
    pt_indices[1::2] = 2 * line_indices + 1
    return np.asarray(points_3d, dtype=np.float64), camera_indices_pt, pt_indices


def split_camera_params(camera_params):
    """Interleaved camera rows ``[w0,c0,w1,c1,w2,c2]`` -> (omega, centre).

    Follows ``reshape((n_cameras,3,2))`` with ``[:,:,0]`` = rvec, ``[:,:,1]`` = c
    (bundle_adjustment.py:111,47-49; producer seq_utils.py:179-181).
    """
    cp = np.asarray(camera_params, dtype=np.float64).reshape(-1, 3, 2)
    return np.ascontiguousarray(cp[:, :, 0]), np.ascontiguousarray(cp[:, :, 1])


def join_camera_params(omega, centre):
    out = np.empty((omega.shape[0], 3, 2), dtype=np.float64)
    out[:, :, 0] = omega
    out[:, :, 1] = centre
    return out.reshape(-1, 6)


def rodrigues(omega):
    """``cv2.Rodrigues`` rvec -> R, vectorised (call site bundle_adjustment.py:48).

    OpenCV (pinned opencv-contrib-python==4.5.1.48, calib3d ``cvRodrigues2``):
    theta = |w|; if theta < DBL_EPSILON: R = I; else r = w/theta and
    R = cos(theta) I + (1-cos(theta)) r r^T + sin(theta) [r]x.
    """
    omega = np.asarray(omega, dtype=np.float64).reshape(-1, 3)
    theta = np.sqrt(omega[:, 0] ** 2 + omega[:, 1] ** 2 + omega[:, 2] ** 2)
    small = theta < DBL_EPSILON
    safe = np.where(small, 1.0, theta)
    r = omega / safe[:, None]
    c = np.cos(theta)
    s = np.sin(theta)
    c1 = 1.0 - c
    rx, ry, rz = r[:, 0], r[:, 1], r[:, 2]
    R = np.empty((omega.shape[0], 3, 3), dtype=np.float64)
    R[:, 0, 0] = c + c1 * rx * rx
    R[:, 0, 1] = c1 * rx * ry - s * rz
    R[:, 0, 2] = c1 * rx * rz + s * ry
    R[:, 1, 0] = c1 * rx * ry + s * rz
    R[:, 1, 1] = c + c1 * ry * ry
    R[:, 1, 2] = c1 * ry * rz - s * rx
    R[:, 2, 0] = c1 * rx * rz - s * ry
    R[:, 2, 1] = c1 * ry * rz + s * rx
    R[:, 2, 2] = c + c1 * rz * rz
    R[small] = np.eye(3)
    return R


def project(points, omega, centre):
    """``project`` (bundle_adjustment.py:23-74): x = R (X - c), p_hat = x/|x|.

    The reference forms P = [R | -R c] and multiplies [X;1] (:51-63); R (X - c)
    is the same quantity.  One row per endpoint-observation.
    """
    R = rodrigues(omega)
    d = points - centre
    x = np.einsum('nij,nj->ni', R, d)
    return x / np.linalg.norm(x, axis=1)[:, None]


def cross_nm(n, p):
    """np.cross(n, p) (bundle_adjustment.py:126,133)."""
    return np.cross(n, p)


def residuals(camera_params, points_3d, camera_indices, line_indices, n, p0, p1):
    """``fun`` (bundle_adjustment.py:78-139), vectorised.

    r_i = sum_e (lamb * p_hat_e . n)^2 + (p_hat_e . (n x p_e))^2  (:125-135).
    ``camera_params`` is (Nc,6) interleaved, ``points_3d`` is (2Nl,3).
    """
    omega, centre = split_camera_params(camera_params)
    pts = np.asarray(points_3d, dtype=np.float64).reshape(-1, 3)
    ci = np.asarray(camera_indices)
    li = np.asarray(line_indices)
    R = rodrigues(omega)[ci]
    c = centre[ci]
    r = np.zeros(ci.shape[0], dtype=np.float64)
    for e, p in ((0, p0), (1, p1)):
        X = pts[2 * li + e]
        x = np.einsum('nij,nj->ni', R, X - c)
        ph = x / np.sqrt(np.einsum('ni,ni->n', x, x))[:, None]
        a = np.einsum('ni,ni->n', ph, n)
        b = np.einsum('ni,ni->n', ph, cross_nm(n, p))
        r += (LAMB * a) ** 2 + b ** 2
    return r


def bundle_adjustment_sparsity_columns(n_cameras, camera_indices, line_indices, n_cam_par=6):
    """Column indices marked by ``bundle_adjustment_sparsity`` (:144-191), as shipped.

    Row i -> ``6 c_i + 0..5`` and (finding F2) ``6 Nc + 3 L_i + 0..2``.
    Returns (m, n_cam_par + 3) int array.
    """
    ci = np.asarray(camera_indices)
    li = np.asarray(line_indices)
    cam_cols = ci[:, None] * n_cam_par + np.arange(n_cam_par)[None, :]
    pt_cols = n_cameras * n_cam_par + li[:, None] * 3 + np.arange(3)[None, :]
    return np.concatenate([cam_cols, pt_cols], axis=1)


# ---- analytic Jacobian (replaces scipy's 2-point FD) ----

def so3_right_jacobian(omega):
    """G(w) with d(R(w) d)/dw = -R [d]x G(w);  G = I - B [w]x + C [w]x^2.

    B = (1-cos t)/t^2 = 2 sin^2(t/2)/t^2,  C = (t - sin t)/t^3; Taylor series
    below t = 0.1 (the toy scenes have a camera with w = 0 exactly).
    """
    omega = np.asarray(omega, dtype=np.float64).reshape(-1, 3)
    t2 = omega[:, 0] ** 2 + omega[:, 1] ** 2 + omega[:, 2] ** 2
    t = np.sqrt(t2)
    small = t < 0.1
    ts = np.where(small, 1.0, t)
    sh = np.sin(0.5 * ts)
    B_big = 2.0 * sh * sh / (ts * ts)
    C_big = (ts - np.sin(ts)) / (ts * ts * ts)
    B_small = 0.5 + t2 * (-1.0 / 24 + t2 * (1.0 / 720 + t2 * (-1.0 / 40320 + t2 * (1.0 / 3628800))))
    C_small = 1.0 / 6 + t2 * (-1.0 / 120 + t2 * (1.0 / 5040 + t2 * (-1.0 / 362880 + t2 * (1.0 / 39916800))))
    B = np.where(small, B_small, B_big)
    C = np.where(small, C_small, C_big)
    wx, wy, wz = omega[:, 0], omega[:, 1], omega[:, 2]
    G = np.empty((omega.shape[0], 3, 3), dtype=np.float64)
    # [w]x^2 = w w^T - t2 I
    G[:, 0, 0] = 1.0 + C * (wx * wx - t2)
    G[:, 0, 1] = B * wz + C * wx * wy
    G[:, 0, 2] = -B * wy + C * wx * wz
    G[:, 1, 0] = -B * wz + C * wx * wy
    G[:, 1, 1] = 1.0 + C * (wy * wy - t2)
    G[:, 1, 2] = B * wx + C * wy * wz
    G[:, 2, 0] = B * wy + C * wx * wz
    G[:, 2, 1] = -B * wx + C * wy * wz
    G[:, 2, 2] = 1.0 + C * (wz * wz - t2)
    return G


def residuals_jacobian(omega, centre, pts, ci, li, n, m0, m1):
    """Residuals and the 12 non-zeros of every Jacobian row.

    ``pts`` is (Nl,6) = [X0,X1]; ``m_e = n x p_e`` precomputed.
    Returns r (m,), jc (m,6) in [d/dw(3), d/dc(3)] order, jl (m,6) = [d/dX0, d/dX1].

    Per endpoint:  d = X - c, x = R d, rho = |x|, ph = x/rho,
    a = ph.n, b = ph.m, e = lamb^2 a^2 + b^2, g = 2 lamb^2 a n + 2 b m,
    u = (g - (g.ph) ph)/rho, de/dX = R^T u, de/dc = -R^T u,
    de/dw = G^T (d x R^T u).
    """
    R = rodrigues(omega)[ci]
    G = so3_right_jacobian(omega)[ci]
    c = centre[ci]
    m = ci.shape[0]
    r = np.zeros(m)
    jl = np.empty((m, 6))
    tq = np.zeros((m, 3))
    l2 = LAMB * LAMB
    for e, me in ((0, m0), (1, m1)):
        X = pts[li, 3 * e:3 * e + 3]
        d = X - c
        x = np.einsum('nij,nj->ni', R, d)
        rho2 = np.einsum('ni,ni->n', x, x)
        inv_rho = 1.0 / np.sqrt(rho2)
        ph = x * inv_rho[:, None]
        a = np.einsum('ni,ni->n', ph, n)
        b = np.einsum('ni,ni->n', ph, me)
        r += l2 * a * a + b * b
        g = (2.0 * l2 * a)[:, None] * n + (2.0 * b)[:, None] * me
        gp = np.einsum('ni,ni->n', g, ph)
        u = (g - gp[:, None] * ph) * inv_rho[:, None]
        je = np.einsum('nji,nj->ni', R, u)  # R^T u
        jl[:, 3 * e:3 * e + 3] = je
        tq += np.cross(d, je)
    jc = np.empty((m, 6))
    jc[:, 0:3] = np.einsum('nji,nj->ni', G, tq)  # G^T tq
    jc[:, 3:6] = -(jl[:, 0:3] + jl[:, 3:6])
    return r, jc, jl


def dense_jacobian_reference_layout(jc, jl, ci, li, n_cameras, n_lines):
    """Scatter rows into the reference's column layout: camera block at
    ``6 c + 2k`` (w_k) / ``6 c + 2k + 1`` (c_k), line block at ``6 Nc + 6 L + 0..5``."""
    m = ci.shape[0]
    J = np.zeros((m, 6 * n_cameras + 6 * n_lines))
    rows = np.arange(m)
    for k in range(3):
        J[rows, 6 * ci + 2 * k] = jc[:, k]
        J[rows, 6 * ci + 2 * k + 1] = jc[:, 3 + k]
    for s in range(6):
        J[rows, 6 * n_cameras + 6 * li + s] = jl[:, s]
    return J


# ---------------------------------------------------------------------------
# Part B -- the LM / Schur / PCG algorithm of this repository (trajectory oracle)
# ---------------------------------------------------------------------------

SYM_IDX = [(a, b) for a in range(6) for b in range(a, 6)]  # 21 upper-triangular entries, row-major


def _sym_pack(M):
    return np.stack([M[:, a, b] for a, b in SYM_IDX], axis=1)


def _sym_unpack(P):
    M = np.empty((P.shape[0], 6, 6))
    for k, (a, b) in enumerate(SYM_IDX):
        M[:, a, b] = P[:, k]
        M[:, b, a] = P[:, k]
    return M


def normal_equations(r, jc, jl, ci, li, n_cameras, n_lines):
    """U_c = sum jc jc^T, g_c = sum jc r (per camera); V_l, g_l likewise per line.
    Returns full (N,6,6) blocks and (N,6) gradients."""
    U = np.zeros((n_cameras, 6, 6))
    V = np.zeros((n_lines, 6, 6))
    np.add.at(U, ci, jc[:, :, None] * jc[:, None, :])
    np.add.at(V, li, jl[:, :, None] * jl[:, None, :])
    gc = np.zeros((n_cameras, 6))
    gl = np.zeros((n_lines, 6))
    np.add.at(gc, ci, jc * r[:, None])
    np.add.at(gl, li, jl * r[:, None])
    return U, gc, V, gl


def update_scaling(D_prev, diag):
    """Marquardt scaling, the analogue of scipy ``x_scale='jac'``
    (scipy/optimize/_lsq/common.py ``compute_jac_scale``): D = sum_i J_ij^2 with
    zero columns set to 1, running maximum over iterations."""
    d = np.where(diag == 0.0, 1.0, diag)
    if D_prev is None:
        return d
    return np.maximum(D_prev, d)


def robust_rho(loss, z):
    """rho(z), rho'(z), rho''(z) of scipy's robust losses (scipy/optimize/_lsq/least_squares.py: soft_l1,
    huber, cauchy, arctan) for z = (f/f_scale)^2."""
    z = np.asarray(z, dtype=np.float64)
    if loss == "soft_l1":
        t = 1.0 + z
        return 2.0 * (np.sqrt(t) - 1.0), t ** -0.5, -0.5 * t ** -1.5
    if loss == "huber":
        big = z > 1.0
        zs = np.where(big, z, 1.0)
        return (np.where(big, 2.0 * np.sqrt(zs) - 1.0, z), np.where(big, zs ** -0.5, 1.0),
                np.where(big, -0.5 * zs ** -1.5, 0.0))
    if loss == "cauchy":
        t = 1.0 + z
        return np.log1p(z), 1.0 / t, -1.0 / (t * t)
    if loss == "arctan":
        t = 1.0 + z * z
        return np.arctan(z), 1.0 / t, -2.0 * z / (t * t)
    raise ValueError(loss)


def robust_cost(r, loss, f_scale=1.0):
    """0.5 * f_scale^2 * sum rho((r/f_scale)^2)  (scipy construct_loss_function, cost_only)."""
    if loss == "linear":
        return 0.5 * float(np.dot(r, r))
    return 0.5 * f_scale * f_scale * float(np.sum(robust_rho(loss, (r / f_scale) ** 2)[0]))


def scale_for_robust_loss(r, jc, jl, loss, f_scale=1.0):
    """scipy ``scale_for_robust_loss_function`` (scipy/optimize/_lsq/common.py:720): rows scaled by
    sqrt(max(rho' + 2 rho'' z, eps)), residuals by rho' over that factor."""
    if loss == "linear":
        return r, jc, jl
    z = (r / f_scale) ** 2
    _, r1, r2 = robust_rho(loss, z)
    js = r1 + 2.0 * r2 * z
    js = np.sqrt(np.maximum(js, DBL_EPSILON))
    return r * r1 / js, jc * js[:, None], jl * js[:, None]


@dataclass
class LMOptions:
    max_iter: int = 50       # damped solves (each followed by one trial evaluation, two when the first is rejected)
    ftol: float = 1e-4       # reference call site bundle_adjustment.py:319
    xtol: float = 1e-8       # scipy default
    gtol: float = 1e-8       # scipy default
    lambda0: float = 1e-2
    lambda_min: float = 1e-6
    lambda_max: float = 1e12
    pcg_tol: float = 1e-2    # eta: stop when |r|_Minv <= eta |r0|_Minv
    pcg_max_iter: int = 200
    min_ratio: float = 1e-4
    loss: str = "linear"    # scipy least_squares `loss` (bundle_adjustment.py:319 forwards loss_f)
    f_scale: float = 1.0
    max_evals: int = 0       # > 0: stop after this many residual(+Jacobian) evaluations (scipy's max_nfev)
    max_backtrack: int = 1   # step-length retries along a rejected step before the damping is raised


# Step policy (the same constants are compiled into the device decision kernel, csrc/lba_kernels.cuh k_lm_decide).
# The residual of this problem is itself a sum of squares r_i = sum_k e_ik^2, so the Gauss-Newton model J^T J of
# 1/2 sum r_i^2 misses the term sum_i r_i grad^2 r_i, which is positive semi-definite and at least HALF of J^T J in
# every direction (Cauchy-Schwarz on grad r_i = 2 sum_k e_ik grad e_ik): near a solution the cost curves >= 1.5x
# more than the model says and the best length along a Gauss-Newton direction is <= 2/3 of the step (SURVEY.md F3:
# gain ratios of 0.3-0.5 for perfectly good steps).  The policy therefore measures that length instead of guessing:
#   * every trial gives the curvature of the cost along the step (parabola through F(x), the slope r.J d and
#     F(x + alpha d)) and with it the best length a_opt along that direction;
#   * the NEXT step is tried at sigma = 1.15 a_opt (clamped to [0.1, 1]) -- the overshoot factor changes slowly
#     from one iteration to the next -- so that rejected trials become rare;
#   * the damping follows a_opt as well: a_opt >= 0.6 with a gain ratio >= 0.5 is what a good model looks like here,
#     consecutive good steps divide lambda by 2, 4, 8 (the analogue of a trust radius that doubles); otherwise
#     lambda grows by a_opt^-2 (at most 10x), which bends the next direction towards the scaled gradient;
#   * a rejected trial is first retried at the measured length along the SAME direction (one evaluation, no solve;
#     scipy's trust-region loop likewise re-solves its 2-D subproblem without a new Gauss-Newton step,
#     scipy/optimize/_lsq/trf.py:509-533) before the damping is raised.
RATIO_GOOD, LENGTH_GOOD = 0.5, 0.6
STREAK_MAX = 3
SIGMA_GAIN, SIGMA_MIN = 1.15, 0.1
LENGTH_CAP = 4.0
GROWTH_MAX = 10.0
ALPHA_LO, ALPHA_HI = 0.1, 0.5


def measured_length(alpha, cost, cost_new, rjd):
    """Best step length along the direction, from the parabola through phi(0) = cost, phi'(0) = r.J delta and
    phi(alpha) = cost_new; None when the parabola is not convex or the trial cost is not finite."""
    if not math.isfinite(cost_new):
        return None
    c2 = (cost_new - cost - rjd * alpha) / (alpha * alpha)
    if not (c2 > 0.0):
        return None
    return -rjd / (2.0 * c2)


def damping_update(ratio, a_opt, streak):
    """After an accepted step: (factor on lambda, new streak, first trial length of the next step)."""
    a = LENGTH_CAP if a_opt is None else min(a_opt, LENGTH_CAP)
    if ratio >= RATIO_GOOD and a >= LENGTH_GOOD:
        streak = streak + 1
        fac = 0.5 ** min(streak, STREAK_MAX)
    else:
        streak = 0
        fac = min(1.0 / (a * a), GROWTH_MAX) if a < 1.0 else 1.0
    sigma = min(1.0, max(SIGMA_MIN, SIGMA_GAIN * a))
    return fac, streak, sigma


def backtrack_alpha(alpha, a_opt):
    """Next length after a rejected trial at `alpha`: the measured one, clamped to [0.1, 0.5] alpha."""
    a = ALPHA_HI * alpha if a_opt is None else a_opt
    return min(max(a, ALPHA_LO * alpha), ALPHA_HI * alpha)


@dataclass
class LMReport:
    iterations: int = 0      # damped solves
    status: int = 0          # 0 max_iter / max_evals, 1 gtol, 2 ftol, 3 xtol, 4 both
    cost_init: float = 0.0
    cost: float = 0.0
    n_jac: int = 0
    n_res: int = 0
    pcg_iters_total: int = 0
    n_accepted: int = 0
    trace: list = field(default_factory=list)   # one dict per trial evaluation


class SchurSystem:
    """Damped normal equations reduced to the cameras, matrix-free.

    S v = lam Dc v + sum_{i in c} jc_i w_i,  w_i = s_i - jl_i . (Vinv_l t_l),
    s_i = jc_i . v_{c_i},  t_l = sum_{i in l} jl_i s_i,
    Vinv_l = (V_l + lam Dl_l)^-1.
    """

    def __init__(self, r, jc, jl, ci, li, U, gc, V, gl, Dc, Dl, lam):
        self.r, self.jc, self.jl, self.ci, self.li = r, jc, jl, ci, li
        self.nc, self.nl = U.shape[0], V.shape[0]
        self.gc, self.gl, self.lam, self.Dc, self.Dl = gc, gl, lam, Dc, Dl
        Vd = V.copy()
        idx = np.arange(6)
        Vd[:, idx, idx] += lam * Dl
        self.Vinv = np.linalg.inv(Vd)
        # rhs  b = -gc + sum_{i in c} jc_i (jl_i . Vinv_l gl_l)
        y = np.einsum('lij,lj->li', self.Vinv, gl)
        wr = np.einsum('ni,ni->n', jl, y[li])
        b = -gc.copy()
        np.add.at(b, ci, jc * wr[:, None])
        self.b = b
        # block-Jacobi preconditioner  M_c = lam Dc + sum (1 - jl_i^T Vinv jl_i) jc_i jc_i^T
        kk = np.einsum('ni,nij,nj->n', jl, self.Vinv[li], jl)
        M = np.zeros((self.nc, 6, 6))
        np.add.at(M, ci, (1.0 - kk)[:, None, None] * jc[:, :, None] * jc[:, None, :])
        M[:, idx, idx] += lam * Dc
        self.Minv = np.linalg.inv(M)

    def matvec(self, v):
        s = np.einsum('ni,ni->n', self.jc, v[self.ci])
        t = np.zeros((self.nl, 6))
        np.add.at(t, self.li, self.jl * s[:, None])
        u = np.einsum('lij,lj->li', self.Vinv, t)
        w = s - np.einsum('ni,ni->n', self.jl, u[self.li])
        out = self.lam * self.Dc * v
        np.add.at(out, self.ci, self.jc * w[:, None])
        return out

    def precond(self, rr):
        return np.einsum('cij,cj->ci', self.Minv, rr)

    def pcg(self, tol, max_iter):
        x = np.zeros_like(self.b)
        rr = self.b.copy()
        z = self.precond(rr)
        p = z.copy()
        rz = float(np.sum(rr * z))
        rz0 = rz
        k = 0
        if rz0 <= 0.0:
            return x, 0
        while k < max_iter:
            q = self.matvec(p)
            alpha = rz / float(np.sum(p * q))
            x += alpha * p
            rr -= alpha * q
            z = self.precond(rr)
            rz_new = float(np.sum(rr * z))
            k += 1
            if rz_new <= tol * tol * rz0:
                break
            beta = rz_new / rz
            p = z + beta * p
            rz = rz_new
        return x, k

    def back_substitute(self, dc):
        """dl = -Vinv (gl + sum_{i in l} jl_i (jc_i . dc));  also J delta per observation."""
        s = np.einsum('ni,ni->n', self.jc, dc[self.ci])
        t = np.zeros((self.nl, 6))
        np.add.at(t, self.li, self.jl * s[:, None])
        dl = -np.einsum('lij,lj->li', self.Vinv, self.gl + t)
        jd = s + np.einsum('ni,ni->n', self.jl, dl[self.li])
        return dl, jd


def lm_solve(camera_params, points_3d, ci, li, n, p0, p1, opts: LMOptions | None = None,
             verbose=False):
    """Levenberg-Marquardt with Marquardt scaling, Schur elimination of the line
    blocks, block-Jacobi PCG on the reduced camera system and a step-length retry on
    rejected steps.

    Replaces the scipy call at bundle_adjustment.py:319.  Termination mirrors
    scipy ``check_termination`` (scipy/optimize/_lsq/common.py:705): ftol when
    dF < ftol*F and ratio > 0.25, xtol when |step| < xtol (xtol + |x|), gtol on
    the inf-norm of the gradient.
    Returns (camera_params (Nc,6) interleaved, points_3d (2Nl,3), residuals, report).
    """
    opts = opts or LMOptions()
    omega, centre = split_camera_params(camera_params)
    pts = np.asarray(points_3d, dtype=np.float64).reshape(-1, 6).copy()
    ci = np.asarray(ci, dtype=np.int64)
    li = np.asarray(li, dtype=np.int64)
    nc, nl = omega.shape[0], pts.shape[0]
    m0 = cross_nm(n, p0)
    m1 = cross_nm(n, p1)
    idx = np.arange(6)

    rep = LMReport()
    r_true, jc, jl = residuals_jacobian(omega, centre, pts, ci, li, n, m0, m1)
    rep.n_jac += 1
    cost = robust_cost(r_true, opts.loss, opts.f_scale)
    r, jc, jl = scale_for_robust_loss(r_true, jc, jl, opts.loss, opts.f_scale)
    rep.cost_init = cost
    U, gc, V, gl = normal_equations(r, jc, jl, ci, li, nc, nl)
    Dc = update_scaling(None, U[:, idx, idx])
    Dl = update_scaling(None, V[:, idx, idx])
    lam = opts.lambda0
    nu = 2.0
    streak = 0
    sigma = 1.0
    status = 0
    it = 0
    nev = 1
    done = False
    while not done and it < opts.max_iter:
        if opts.max_evals and nev >= opts.max_evals:
            break
        g_norm = max(np.abs(gc).max() if nc else 0.0, np.abs(gl).max() if nl else 0.0)
        if g_norm < opts.gtol:
            status = 1
            break
        sys_ = SchurSystem(r, jc, jl, ci, li, U, gc, V, gl, Dc, Dl, lam)
        dc, q = sys_.pcg(opts.pcg_tol, opts.pcg_max_iter)
        dl, jd = sys_.back_substitute(dc)
        rep.pcg_iters_total += q
        it += 1
        rjd = float(np.dot(r, jd))
        jd2 = float(np.dot(jd, jd))
        d2 = float(np.sum(dc * dc) + np.sum(dl * dl))
        x2 = float(np.sum(omega * omega) + np.sum(centre * centre) + np.sum(pts * pts))
        x_norm = math.sqrt(x2)
        alpha = sigma
        bt = 0
        while True:
            if opts.max_evals and nev >= opts.max_evals:
                done = True
                break
            pred = -alpha * rjd - 0.5 * alpha * alpha * jd2
            om_new = omega + alpha * dc[:, 0:3]
            ce_new = centre + alpha * dc[:, 3:6]
            pts_new = pts + alpha * dl
            r_new = residuals_jacobian(om_new, ce_new, pts_new, ci, li, n, m0, m1)[0]
            rep.n_res += 1
            nev += 1
            cost_new = robust_cost(r_new, opts.loss, opts.f_scale)
            step_norm = alpha * math.sqrt(d2)
            actual = cost - cost_new
            finite = math.isfinite(cost_new)
            ratio = actual / pred if (finite and pred > 0.0) else -1.0
            accepted = finite and ratio > opts.min_ratio
            a_opt = measured_length(alpha, cost, cost_new, rjd)
            ftol_ok = finite and (actual < opts.ftol * cost) and (ratio > 0.25)
            xtol_ok = step_norm < opts.xtol * (opts.xtol + x_norm)
            rep.trace.append(dict(it=it - 1, cost=cost, cost_new=cost_new, lam=lam, pcg=q, pred=pred, alpha=alpha,
                                  ratio=ratio, accepted=accepted, step_norm=step_norm, g_norm=g_norm, a_opt=a_opt,
                                  rjd=rjd, jd2=jd2, d2=d2, x2=x2))
            if verbose:
                print(f"{it - 1:4d} cost {cost:.6e} -> {cost_new:.6e} lam {lam:.2e} alpha {alpha:.3f} pcg {q:3d} "
                      f"ratio {ratio:+.3f} step {step_norm:.2e} |g| {g_norm:.2e}")
            retry = False
            if accepted:
                rep.n_accepted += 1
                omega, centre, pts = om_new, ce_new, pts_new
                r_true, jc, jl = residuals_jacobian(omega, centre, pts, ci, li, n, m0, m1)
                r, jc, jl = scale_for_robust_loss(r_true, jc, jl, opts.loss, opts.f_scale)
                rep.n_jac += 1
                U, gc, V, gl = normal_equations(r, jc, jl, ci, li, nc, nl)
                Dc = update_scaling(Dc, U[:, idx, idx])
                Dl = update_scaling(Dl, V[:, idx, idx])
                fac, streak, sigma = damping_update(ratio, a_opt, streak)
                lam = min(max(lam * fac, opts.lambda_min), opts.lambda_max)
                nu = 2.0
            elif bt < opts.max_backtrack:
                bt += 1
                alpha_next = backtrack_alpha(alpha, a_opt)
                retry = True
            else:
                lam = min(lam * GROWTH_MAX * nu, opts.lambda_max)
                nu *= 2.0
                streak = 0
            if accepted:
                cost = cost_new
            if ftol_ok and xtol_ok:
                status = 4
            elif ftol_ok:
                status = 2
            elif xtol_ok:
                status = 3
            if status:
                done = True
                break
            if not retry:
                break
            alpha = alpha_next
    rep.iterations = it
    rep.status = status
    rep.cost = cost
    cams_out = join_camera_params(omega, centre)
    return cams_out, pts.reshape(-1, 3), r_true, rep


# ---------------------------------------------------------------------------
# Part B' -- the 12-parameter [R|c] camera rows (bundle_adjustment.py:43-45): dense LM with the same step policy
# ---------------------------------------------------------------------------

def residuals12(camera_params12, points_3d, camera_indices, line_indices, n, p0, p1):
    """``fun`` with (Nc,12) camera rows: ``reshape((3,4))``, R = [:, :3] (nine free entries), c = [:, 3]
    (bundle_adjustment.py:43-45,111); x = R (X - c) as ``[R | -R c] [X;1]`` (:51-63)."""
    return residuals_jacobian12(camera_params12, points_3d, camera_indices, line_indices, n, p0, p1, with_jac=False)[0]


def residuals_jacobian12(camera_params12, points_3d, ci, li, n, p0, p1, with_jac=True):
    """r (m,), jc (m,12) in the row layout [M_k0 M_k1 M_k2 c_k], jl (m,6):
    u = de/dx,  de/dM_kj = u_k d_j,  de/dX = M^T u,  de/dc = -M^T u."""
    P = np.asarray(camera_params12, dtype=np.float64).reshape(-1, 3, 4)
    M, c = P[:, :, :3][ci], P[:, :, 3][ci]
    pts = np.asarray(points_3d, dtype=np.float64).reshape(-1, 6)[li]
    m = ci.shape[0]
    r = np.zeros(m)
    jc = np.zeros((m, 3, 4)); jl = np.zeros((m, 6))
    l2 = LAMB * LAMB
    for e, pe in enumerate((p0, p1)):
        me = cross_nm(n, pe)
        d = pts[:, 3 * e:3 * e + 3] - c
        x = np.einsum('nij,nj->ni', M, d)
        inv_rho = 1.0 / np.sqrt(np.einsum('ni,ni->n', x, x))
        q = x * inv_rho[:, None]
        a = np.einsum('ni,ni->n', q, n); b = np.einsum('ni,ni->n', q, me)
        r += l2 * a * a + b * b
        if with_jac:
            g = (2.0 * l2 * a)[:, None] * n + (2.0 * b)[:, None] * me
            gp = np.einsum('ni,ni->n', g, q)
            u = (g - gp[:, None] * q) * inv_rho[:, None]
            je = np.einsum('nij,ni->nj', M, u)
            jc[:, :, :3] += u[:, :, None] * d[:, None, :]
            jc[:, :, 3] -= je
            jl[:, 3 * e:3 * e + 3] = je
    return r, jc.reshape(m, 12), jl


def lm_solve12(camera_params12, points_3d, ci, li, n, p0, p1, opts: LMOptions | None = None, verbose=False):
    """Levenberg-Marquardt for the 12-parameter camera rows: dense damped normal equations
    (J^T J + lam D) delta = -J^T r solved by Cholesky, Marquardt scaling D = running max of diag(J^T J), and the
    step policy of ``lm_solve`` (first trial at sigma, measured length, retry along the same direction).
    Returns (camera_params12, points_3d, residuals, report)."""
    opts = opts or LMOptions()
    ci = np.asarray(ci, dtype=np.int64); li = np.asarray(li, dtype=np.int64)
    cams = np.asarray(camera_params12, dtype=np.float64).reshape(-1, 12).copy()
    pts = np.asarray(points_3d, dtype=np.float64).reshape(-1, 6).copy()
    nc, nl, m = cams.shape[0], pts.shape[0], ci.shape[0]
    nn = 12 * nc + 6 * nl

    def normal(r, jc, jl):
        J = np.zeros((m, nn))
        rows = np.arange(m)
        for k in range(12):
            J[rows, 12 * ci + k] = jc[:, k]
        for k in range(6):
            J[rows, 12 * nc + 6 * li + k] = jl[:, k]
        return J, J.T @ J, J.T @ r

    rep = LMReport()
    r_true, jc, jl = residuals_jacobian12(cams, pts, ci, li, n, p0, p1)
    rep.n_jac += 1
    cost = robust_cost(r_true, opts.loss, opts.f_scale)
    r, jc, jl = scale_for_robust_loss(r_true, jc, jl, opts.loss, opts.f_scale)
    rep.cost_init = cost
    J, H, g = normal(r, jc, jl)
    D = np.where(np.diag(H) == 0.0, 1.0, np.diag(H))
    lam, nu, streak, sigma, status, it, nev, done = opts.lambda0, 2.0, 0, 1.0, 0, 0, 1, False
    while not done and it < opts.max_iter:
        if opts.max_evals and nev >= opts.max_evals:
            break
        g_norm = float(np.abs(g).max()) if nn else 0.0
        if g_norm < opts.gtol:
            status = 1
            break
        delta = np.linalg.solve(H + lam * np.diag(D), -g)
        it += 1
        jd = J @ delta
        rjd, jd2, d2 = float(r @ jd), float(jd @ jd), float(delta @ delta)
        x2 = float(np.sum(cams * cams) + np.sum(pts * pts)); x_norm = math.sqrt(x2)
        alpha, bt = sigma, 0
        while True:
            if opts.max_evals and nev >= opts.max_evals:
                done = True
                break
            pred = -alpha * rjd - 0.5 * alpha * alpha * jd2
            cams_new = cams + alpha * delta[:12 * nc].reshape(nc, 12)
            pts_new = pts + alpha * delta[12 * nc:].reshape(nl, 6)
            r_new = residuals12(cams_new, pts_new, ci, li, n, p0, p1)
            rep.n_res += 1; nev += 1
            cost_new = robust_cost(r_new, opts.loss, opts.f_scale)
            step_norm = alpha * math.sqrt(d2)
            actual = cost - cost_new
            finite = math.isfinite(cost_new)
            ratio = actual / pred if (finite and pred > 0.0) else -1.0
            accepted = finite and ratio > opts.min_ratio
            a_opt = measured_length(alpha, cost, cost_new, rjd)
            ftol_ok = finite and (actual < opts.ftol * cost) and (ratio > 0.25)
            xtol_ok = step_norm < opts.xtol * (opts.xtol + x_norm)
            rep.trace.append(dict(it=it - 1, cost=cost, cost_new=cost_new, lam=lam, pcg=0, pred=pred, alpha=alpha, ratio=ratio,
                                  accepted=accepted, step_norm=step_norm, g_norm=g_norm, a_opt=a_opt))
            if verbose:
                print(f"{it - 1:4d} cost {cost:.6e} -> {cost_new:.6e} lam {lam:.2e} alpha {alpha:.3f} ratio {ratio:+.3f}")
            retry = False
            if accepted:
                rep.n_accepted += 1
                cams, pts, cost = cams_new, pts_new, cost_new
                r_true, jc, jl = residuals_jacobian12(cams, pts, ci, li, n, p0, p1)
                r, jc, jl = scale_for_robust_loss(r_true, jc, jl, opts.loss, opts.f_scale)
                rep.n_jac += 1
                J, H, g = normal(r, jc, jl)
                D = np.maximum(D, np.where(np.diag(H) == 0.0, 1.0, np.diag(H)))
                fac, streak, sigma = damping_update(ratio, a_opt, streak)
                lam = min(max(lam * fac, opts.lambda_min), opts.lambda_max)
                nu = 2.0
            elif bt < opts.max_backtrack:
                bt += 1
                alpha_next = backtrack_alpha(alpha, a_opt)
                retry = True
            else:
                lam = min(lam * GROWTH_MAX * nu, opts.lambda_max)
                nu *= 2.0
                streak = 0
            status = 4 if (ftol_ok and xtol_ok) else (2 if ftol_ok else (3 if xtol_ok else 0))
            if status:
                done = True
                break
            if not retry:
                break
            alpha = alpha_next
    rep.iterations, rep.status, rep.cost = it, status, cost
    return cams, pts.reshape(-1, 3), r_true, rep


# ---------------------------------------------------------------------------
# Part C -- vectorised port of the reference's scipy call (timed CPU baseline)
# ---------------------------------------------------------------------------

def _x_pack(camera_params, points_3d):
    return np.hstack((np.asarray(camera_params, dtype=np.float64).ravel(),
                      np.asarray(points_3d, dtype=np.float64).ravel()))   # bundle_adjustment.py:311


def scipy_reference_port(camera_params, points_3d, ci, li, n, p0, p1, mode="as_shipped",
                         max_nfev=None, verbose=0, ftol=1e-4, loss="linear"):
    """``run_bundle_adjustment`` (bundle_adjustment.py:254-349) with ``fun`` vectorised.

    mode="as_shipped": 2-point finite differences with the shipped sparsity
        pattern (finding F2), exactly the reference's ``least_squares`` arguments
        (:319): jac_sparsity=A, x_scale='jac', ftol=1e-4, method='trf', loss='linear'.
    mode="analytic":  same call with the analytic sparse Jacobian (correct
        columns) -- the best-case CPU baseline.
    Returns (op_camera_params, op_points_3d, residual, scipy result, seconds).
    """
    from scipy.optimize import least_squares
    from scipy.sparse import csr_matrix, lil_matrix

    camera_params = np.asarray(camera_params, dtype=np.float64)
    nc = camera_params.shape[0]
    pts3 = np.asarray(points_3d, dtype=np.float64).reshape(-1, 3)
    npts = pts3.shape[0]
    nl = npts // 2
    ci = np.asarray(ci, dtype=np.int64)
    li = np.asarray(li, dtype=np.int64)
    m = ci.shape[0]
    x0 = _x_pack(camera_params, pts3)
    m0 = cross_nm(n, p0)
    m1 = cross_nm(n, p1)

    def f(x):
        return residuals(x[:6 * nc].reshape(nc, 6), x[6 * nc:].reshape(npts, 3), ci, li, n, p0, p1)

    kwargs = dict(verbose=verbose, x_scale='jac', ftol=ftol, method='trf', loss=loss,
                  max_nfev=max_nfev)
    if mode == "as_shipped":
        cols = bundle_adjustment_sparsity_columns(nc, ci, li)
        A = lil_matrix((m, x0.size), dtype=int)
        rows = np.repeat(np.arange(m), cols.shape[1])
        A[rows, cols.ravel()] = 1
        kwargs["jac_sparsity"] = A
    elif mode == "analytic":
        rows = np.repeat(np.arange(m), 12)
        colmat = np.empty((m, 12), dtype=np.int64)
        for k in range(3):
            colmat[:, k] = 6 * ci + 2 * k          # w_k
            colmat[:, 3 + k] = 6 * ci + 2 * k + 1  # c_k
        colmat[:, 6:12] = 6 * nc + 6 * li[:, None] + np.arange(6)[None, :]

        def jac(x):
            om, ce = split_camera_params(x[:6 * nc].reshape(nc, 6))
            _, jc, jl = residuals_jacobian(om, ce, x[6 * nc:].reshape(nl, 6), ci, li, n, m0, m1)
            vals = np.concatenate([jc, jl], axis=1).ravel()
            return csr_matrix((vals, (rows, colmat.ravel())), shape=(m, x0.size))

        kwargs["jac"] = jac
        kwargs["tr_solver"] = "lsmr"
    else:
        raise ValueError(mode)
    t0 = time.time()
    res = least_squares(f, x0, **kwargs)
    dt = time.time() - t0
    op_cam = res.x[:6 * nc].reshape(nc, 6)
    op_pts = res.x[6 * nc:].reshape(npts, 3)
    return op_cam, op_pts, np.copy(res.fun), res, dt
