"""CPU oracle for the track-table construction.  TEST INFRASTRUCTURE ONLY: imported by tests/
(and by nothing in the product path).

Restates, array-native, what the reference does with Python objects:

* ``verify_duplet`` (/root/reference/src/utils.py:1807-1944): for each matched pair of 2D segments of views
  (v, v+1) the two endpoints of the first segment are mapped to epipolar lines ``F x`` (:1886-1887), intersected
  with the supporting line of the second segment (``line_parametrization`` :1250, ``intersection`` :1279), and the
  match is kept when one endpoint is within ``threshold * min(|I1 - I0|, |p1 - p0|)`` of its intersection
  (criterion 1, :1909) and the endpoint order is consistent (criterion 2, :1911).
* ``Tracker.comp_tracker`` (/root/reference/src/tracker.py:11-96): the ``ids[track, view]`` / ``bin[track, view]``
  tables: pair (0,1) opens one track per inlier match; for v > 0 an inlier match whose first line already sits in
  column v of some track extends every such track (later matches overwrite earlier ones), the others open new
  tracks, appended in match order; ``n_new_tracks[v]`` counts them.
* ``relativeCameraMotion`` (:1483) and ``F_from_KRt2`` (:1039) for the fundamental matrix of a pair.

Two reference behaviours are kept on purpose: absent entries of ``ids`` are 0 and take part in the membership test
``l in Tr_ij[:, v]`` (tracker.py:60), so a match whose first line has id 0 extends every track that has no line in
view v; and a line that sits in column v of several tracks extends all of them (:68-70).  For exactly parallel lines
``intersection`` returns ``False`` and the reference raises (utils.py:1890): such a match is an outlier here.  Within
rounding of parallel the reference divides by noise and its decision is set by the rounding of every operation,
including the summation order inside ``np.dot``; ``duplet_inlier`` reproduces those roundings (pinned by the
recorded flags of 2240 reference matches, 88 of them near-parallel).

Pinned against the reference's own ``Tracker`` on the toy domains (tests/golden/trk_*.npz,
make_tracker_golden.py) and against a re-run of the reference's ``verify_duplet`` on perturbed copies of them.
"""
import math
from fractions import Fraction

import numpy as np

PARALLEL_TOL = 0.0   # opt-in: |sin| of the angle between the segment and the epipolar line below which a match is dropped


def fundamental_from_cameras(K, R0, t0, R1, t1):
    """F of the pair (view0, view1): relativeCameraMotion (utils.py:1504-1509) then F_from_KRt2 (:1059-1065)."""
    K = np.asarray(K, dtype=np.float64)
    R0 = np.asarray(R0, dtype=np.float64); R1 = np.asarray(R1, dtype=np.float64)
    t0 = np.asarray(t0, dtype=np.float64).reshape(3, 1); t1 = np.asarray(t1, dtype=np.float64).reshape(3, 1)
    rR = R1 @ R0.T
    rt = (t1 - rR @ t0).reshape((3, 1))
    rt_x = np.array([[0, -rt[2][0], rt[1][0]], [rt[2][0], 0, -rt[0][0]], [-rt[1][0], rt[0][0], 0]])
    F = np.linalg.inv(K).T @ rt_x @ rR @ np.linalg.inv(K)
    return F / F[2, 2]


def _fma(a, b, c):
    """Correctly rounded a*b + c (Python 3.12 has no math.fma): exact rational arithmetic, rounded once."""
    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def _norm2(dx, dy):
    """np.linalg.norm of a 2-vector: sqrt(x.dot(x)), products and sum rounded one at a time."""
    return math.sqrt(dx * dx + dy * dy)


def duplet_inlier(F, a0, a1, b0, b1, threshold, parallel_tol=PARALLEL_TOL):
    """One match: a0, a1 endpoints (x, y) of the segment in view v, b0, b1 of the segment in view v+1.
    utils.py:1877-1911 in scalar arithmetic, every operation rounded as numpy rounds it in the reference: the rows of
    np.dot(F, [x, y, 1]) (:1886) as the BLAS kernel evaluates them, fma(F_k0, x, F_k1 * y) + F_k2 (pinned by the
    recorded reference flags, which are pure rounding on the near-parallel matches), the rest plain scalar ops."""
    F = np.asarray(F, dtype=np.float64)
    a0 = [float(v) for v in a0]; a1 = [float(v) for v in a1]; b0 = [float(v) for v in b0]; b1 = [float(v) for v in b1]
    # line through b0, b1 (:1250)
    A = b0[1] - b1[1]
    B = b1[0] - b0[0]
    Cc = b0[0] * b1[1] - b1[0] * b0[1]
    inter = []
    for a in (a0, a1):
        l0 = _fma(F[0, 0], a[0], float(F[0, 1]) * a[1]) + float(F[0, 2])
        l1 = _fma(F[1, 0], a[0], float(F[1, 1]) * a[1]) + float(F[1, 2])
        l2 = _fma(F[2, 0], a[0], float(F[2, 1]) * a[1]) + float(F[2, 2])
        D = A * l1 - B * l0                                      # intersection (:1279)
        Dx = -Cc * l1 + B * l2
        Dy = -A * l2 + Cc * l0
        if D == 0.0:                 # the reference raises here (intersection returns False, :1890): outlier by rule
            return False
        if parallel_tol > 0.0 and not abs(D) > parallel_tol * _norm2(A, B) * _norm2(l0, l1):
            return False             # opt-in: drop matches within parallel_tol of parallel
        inter.append((Dx / D, Dy / D))
    (x0, y0), (x1, y1) = inter
    len_line = _norm2(x1 - x0, y1 - y0)
    len_corr = _norm2(b1[0] - b0[0], b1[1] - b0[1])
    thr = threshold * min(len_line, len_corr)
    d11 = _norm2(b0[0] - x0, b0[1] - y0)
    d22 = _norm2(b1[0] - x1, b1[1] - y1)
    d12 = _norm2(b0[0] - x1, b0[1] - y1)
    d21 = _norm2(b1[0] - x0, b1[1] - y0)
    if d11 < thr or d22 < thr:
        if (d11 <= d22 and d22 < d21) or (d22 < d11 and d11 < d12):
            return True
    return False


def verify_duplets(F, endpoints, matches, threshold=0.1):
    """Inlier flag per match.  endpoints (n_lines2d, 2, 2) f64 indexed by the global 2D line id
    (domain.lines2d[id], utils.py:1874-1875); matches (k, 2) int."""
    endpoints = np.asarray(endpoints, dtype=np.float64)
    matches = np.asarray(matches, dtype=np.int64).reshape(-1, 2)
    out = np.zeros(len(matches), dtype=bool)
    for j, (l0, l1) in enumerate(matches):
        out[j] = duplet_inlier(F, endpoints[l0, 0], endpoints[l0, 1], endpoints[l1, 0], endpoints[l1, 1], threshold)
    return out


def comp_tracker(endpoints, matches_per_pair, F_per_pair, threshold=0.1):
    """Sequential restatement of Tracker.comp_tracker (tracker.py:39-96).  matches_per_pair[v] is the (k_v, 2) match
    list of views (v, v+1), F_per_pair[v] their fundamental matrix.  Returns (ids (T, V) int64, bin (T, V) int64,
    n_new_tracks (V-1,) int64, inlier flags per pair)."""
    V = len(matches_per_pair) + 1
    ids = np.zeros((0, V), dtype=np.int64)
    bi = np.zeros((0, V), dtype=np.int64)
    n_new, flags = [], []
    for v in range(V - 1):
        m = np.asarray(matches_per_pair[v], dtype=np.int64).reshape(-1, 2)
        fl = verify_duplets(F_per_pair[v], endpoints, m, threshold)
        flags.append(fl)
        inl = m[fl]
        if v > 0:
            seen = np.array([l in ids[:, v] for l in inl[:, 0]], dtype=bool)          # :60 (zeros included)
            for l in inl[seen]:                                                      # :67-70
                ind = np.where(l[0] == ids[:, v])
                ids[ind, v + 1] = l[1]
                bi[ind, v + 1] = 1
            new = inl[~seen]
        else:
            new = inl
        it_ids = np.zeros((len(new), V), dtype=np.int64)
        it_bin = np.zeros((len(new), V), dtype=np.int64)
        it_ids[:, v:v + 2] = new
        it_bin[:, v:v + 2] = 1
        n_new.append(len(new))
        ids = np.concatenate((ids, it_ids))
        bi = np.concatenate((bi, it_bin))
    return ids, bi, np.asarray(n_new, dtype=np.int64), flags


def comp_tracker_tables(endpoints, matches_per_pair, F_per_pair, threshold=0.1):
    """The same tables by the data-parallel formulation the CUDA path uses (per pair: flags -> prefix sums, a
    presence table over column v, a last-writer table over the inlier matches, extension of the rows, ordered
    append): no sequential dependence inside a pair.  Tested equal to ``comp_tracker``."""
    endpoints = np.asarray(endpoints, dtype=np.float64)
    n_lines = endpoints.shape[0]
    V = len(matches_per_pair) + 1
    cap = sum(len(m) for m in matches_per_pair)
    ids = np.zeros((cap, V), dtype=np.int64)
    bi = np.zeros((cap, V), dtype=np.int64)
    rows = 0
    n_new, flags = [], []
    for v in range(V - 1):
        m = np.asarray(matches_per_pair[v], dtype=np.int64).reshape(-1, 2)
        fl = verify_duplets(F_per_pair[v], endpoints, m, threshold)
        flags.append(fl)
        pos = np.cumsum(fl) - fl                                  # exclusive scan: rank among the inliers
        inl_l1 = m[fl, 1]
        if v > 0:
            present = np.zeros(n_lines, dtype=bool)
            present[ids[:rows, v]] = True                         # zeros of absent entries included
            last = np.full(n_lines, -1, dtype=np.int64)
            sel = fl & present[m[:, 0]]
            np.maximum.at(last, m[sel, 0], pos[sel])              # the last inlier match of a line wins
            j = last[ids[:rows, v]]
            hit = j >= 0
            ids[:rows, v + 1][hit] = inl_l1[j[hit]]
            bi[:rows, v + 1][hit] = 1
            newf = fl & ~present[m[:, 0]]
        else:
            newf = fl
        k = int(newf.sum())
        ids[rows:rows + k, v] = m[newf, 0]
        ids[rows:rows + k, v + 1] = m[newf, 1]
        bi[rows:rows + k, v:v + 2] = 1
        rows += k
        n_new.append(k)
    return ids[:rows], bi[:rows], np.asarray(n_new, dtype=np.int64), flags


def observations_from_tables(ids, bi):
    """The observation list input_bundle_adj derives from the tables (seq_utils.py:186-195): tracks seen in more
    than one view, track-major, view-ascending.  Returns (camera_indices, lines_indices, line_ids)."""
    bi = np.asarray(bi)
    n_tracks = int(np.count_nonzero(bi.sum(axis=1) > 1))
    t_idx, v_idx = np.where(bi[:n_tracks] == 1)
    return v_idx.astype(np.int64), t_idx.astype(np.int64), np.asarray(ids)[t_idx, v_idx].astype(np.int64)
