"""CPU oracle for the two-view endpoint triangulation.  TEST INFRASTRUCTURE ONLY: imported by
tests/ (and nothing in the product path); it restates what the reference computes, it is not shipped.

Reference: ``/root/reference/src/seq_utils.py::triangulate_model`` (:75-135) and ``triangulate_model2`` (:11-72):
for every new track of the view pair (i, i+1) the two endpoints of the matched 2D segments are triangulated
with ``cv2.triangulatePoints(P0, P1, pts_v0, pts_v1)`` (:119, float32 points), and the result is dehomogenised
in float32, ``X[:3] / X[3]`` (:121).

Third-party arithmetic: OpenCV ``cv2.triangulatePoints`` (pinned opencv-contrib-python==4.5.1.48 in the
reference's requirements.txt:4; 4.13.0 installed in the build container), not under /root/reference.  Its
published algorithm (calib3d ``triangulate.cpp``): per point the 4x4 DLT matrix

    A[2j+0, :] = x_j * P_j[2, :] - P_j[0, :]
    A[2j+1, :] = y_j * P_j[2, :] - P_j[1, :]        j = 0, 1

in double precision, ``cv::SVD`` of A (one-sided Jacobi), the right singular vector of the smallest singular
value as the homogeneous point, converted to the depth of the input points (float32 here).

Pinned: against ``cv2.triangulatePoints`` itself (tests, when cv2 is importable) and against the outputs of the
reference's ``triangulate_model`` recorded in ``tests/golden/tri_*.npz`` (make_triangulation_golden.py).
"""
import numpy as np


def dlt_matrix(P0, P1, x0, x1):
    """The 4x4 matrix OpenCV builds for one point seen at x0 (view 0) and x1 (view 1); float64."""
    P0 = np.asarray(P0, dtype=np.float64)
    P1 = np.asarray(P1, dtype=np.float64)
    x0 = np.asarray(x0, dtype=np.float64)   # float32 -> float64 is exact
    x1 = np.asarray(x1, dtype=np.float64)
    return np.stack([x0[0] * P0[2] - P0[0], x0[1] * P0[2] - P0[1],
                     x1[0] * P1[2] - P1[0], x1[1] * P1[2] - P1[1]])


def jacobi_null_vector(A, max_sweeps=30):
    """Right singular vector of the smallest singular value of A (..., 4, 4) by Hestenes' one-sided Jacobi, the
    method behind ``cv::SVD`` (OpenCV ``JacobiSVDImpl_`` in core/src/lapack.cpp, restated): the columns of A are
    rotated in pairs (i < j) until mutually orthogonal, the rotations are accumulated in V, and the column that
    ends with the smallest norm gives the null direction.  Batched over the leading axes with masks; the same
    operation order as csrc/lba_triangulate.cuh."""
    A = np.asarray(A, dtype=np.float64)
    At = np.swapaxes(A, -1, -2).copy()                    # row i of At = column i of A
    Vt = np.broadcast_to(np.eye(4), At.shape).copy()
    W = np.sum(At * At, axis=-1)
    eps = np.finfo(np.float64).eps * 10.0
    for _ in range(max_sweeps):
        changed = False
        for i in range(3):
            for j in range(i + 1, 4):
                a, b = W[..., i], W[..., j]
                p = np.sum(At[..., i, :] * At[..., j, :], axis=-1)
                act = np.abs(p) > eps * np.sqrt(a * b)
                if not np.any(act):
                    continue
                changed = True
                with np.errstate(all="ignore"):
                    p2 = p * 2.0
                    beta = a - b
                    gamma = np.hypot(p2, beta)
                    delta = (gamma - beta) * 0.5
                    s_neg = np.sqrt(delta / gamma)
                    c_neg = p2 / (gamma * s_neg * 2.0)
                    c_pos = np.sqrt((gamma + beta) / (gamma * 2.0))
                    s_pos = p2 / (gamma * c_pos * 2.0)
                c = np.where(act, np.where(beta < 0.0, c_neg, c_pos), 1.0)[..., None]
                s = np.where(act, np.where(beta < 0.0, s_neg, s_pos), 0.0)[..., None]
                ai, aj = At[..., i, :].copy(), At[..., j, :].copy()
                ti, tj = c * ai + s * aj, -s * ai + c * aj
                At[..., i, :], At[..., j, :] = ti, tj
                W[..., i] = np.where(act, np.sum(ti * ti, axis=-1), a)
                W[..., j] = np.where(act, np.sum(tj * tj, axis=-1), b)
                vi, vj = Vt[..., i, :].copy(), Vt[..., j, :].copy()
                Vt[..., i, :], Vt[..., j, :] = c * vi + s * vj, -s * vi + c * vj
        if not changed:
            break
    m = np.argmin(W, axis=-1)                             # first minimum, like the strict '<' scan of the kernel
    return np.take_along_axis(Vt, m[..., None, None], axis=-2)[..., 0, :]


def triangulate_point_h(P0, P1, x0, x1):
    """Homogeneous float32 4-vector, as cv2.triangulatePoints returns it (up to the sign of the vector)."""
    return jacobi_null_vector(dlt_matrix(P0, P1, x0, x1)).astype(np.float32)


def triangulate_point_h_lapack(P0, P1, x0, x1):
    """Independent cross-check: the same null vector from LAPACK's SVD (differs from the Jacobi result by the
    conditioning of the point, ~1e-14 relative, i.e. by an occasional float32 ulp after the conversion)."""
    return np.linalg.svd(dlt_matrix(P0, P1, x0, x1))[2][3].astype(np.float32)


def triangulate_tracks(P, view0, pts0, pts1, view1=None):
    """All endpoints of all tracks.  P (V,3,4); view0 (n,) (view1 defaults to view0 + 1, seq_utils.py:106-107);
    pts0, pts1 (n,2,2) float32 [track][endpoint][x,y].  Returns (X (n,2,3) float32, Xh (n,2,4) float32) with
    X = Xh[:3] / Xh[3] evaluated in float32 like the reference (:121)."""
    P = np.asarray(P, dtype=np.float64)
    view0 = np.asarray(view0, dtype=np.int64)
    view1 = view0 + 1 if view1 is None else np.asarray(view1, dtype=np.int64)
    pts0 = np.asarray(pts0, dtype=np.float32)
    pts1 = np.asarray(pts1, dtype=np.float32)
    n = view0.shape[0]
    Xh = np.empty((n, 2, 4), dtype=np.float32)
    for t in range(n):
        for e in range(2):
            Xh[t, e] = triangulate_point_h(P[view0[t]], P[view1[t]], pts0[t, e], pts1[t, e])
    X = Xh[:, :, :3] / Xh[:, :, 3:4]          # float32 / float32
    return X, Xh


def triangulate_tracks_vectorised(P, view0, pts0, pts1, view1=None):
    """Same numbers with the batched Jacobi (for the sizes a Python loop cannot do)."""
    P = np.asarray(P, dtype=np.float64)
    view0 = np.asarray(view0, dtype=np.int64)
    view1 = view0 + 1 if view1 is None else np.asarray(view1, dtype=np.int64)
    x0 = np.asarray(pts0, dtype=np.float32).astype(np.float64)
    x1 = np.asarray(pts1, dtype=np.float32).astype(np.float64)
    Pa, Pb = P[view0][:, None], P[view1][:, None]                       # (n,1,3,4)
    A = np.stack([x0[..., 0:1] * Pa[:, :, 2] - Pa[:, :, 0], x0[..., 1:2] * Pa[:, :, 2] - Pa[:, :, 1],
                  x1[..., 0:1] * Pb[:, :, 2] - Pb[:, :, 0], x1[..., 1:2] * Pb[:, :, 2] - Pb[:, :, 1]], axis=2)
    Xh = jacobi_null_vector(A).astype(np.float32)
    return Xh[..., :3] / Xh[..., 3:4], Xh


def ulp_distance_f32(a, b):
    """Distance in float32 units in the last place between two float32 arrays (sign-magnitude aware)."""
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a)
    b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)
