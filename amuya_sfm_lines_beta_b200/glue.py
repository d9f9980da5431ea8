"""Array-native glue between the reference's ``Domain`` objects and the GPU solver.

The reference rebuilds the BA input with Python loops after every registered view and walks 2 m ``Line``
objects again inside ``run_bundle_adjustment`` (``src/seq_utils.py:138-208``, ``src/bundle_adjustment.py:194-250``).
These helpers produce the same arrays in one vectorised pass, keep one ``LineBA`` context alive across the
incremental loop (device buffers are reused, only grown) and write the result back exactly as
``update_structure_motion`` does (``src/seq_utils.py:211-252``).

They operate on the reference's own objects (``domain.views[i].P.R / .c``, ``domain.tracker.ids / .bin``,
``domain.lines2d[j].n / .points[e].p``, ``domain.lines3d[L].points[e].coord``) and need ``cv2`` for the two
``cv2.Rodrigues`` conversions the reference itself performs.
"""
from __future__ import annotations

import numpy as np

from .solver import LineBA


def observations_from_tables(ids, bin_):
    """The observation list ``input_bundle_adj`` derives from the tracker tables (seq_utils.py:186-195): tracks seen in
    more than one view, track-major and view-ascending.  Returns ``(camera_indices, lines_indices, line_ids)`` int64."""
    bin_ = np.asarray(bin_)
    n_tracks = int(np.count_nonzero(bin_.sum(axis=1) > 1))                    # :188-189
    t_idx, v_idx = np.where(bin_[:n_tracks] == 1)                             # :190
    return v_idx.astype(np.int64), t_idx.astype(np.int64), np.asarray(ids)[t_idx, v_idx].astype(np.int64)


def input_arrays_from_domain(domain, last_view=2):
    """Vectorised ``input_bundle_adj(domain, "Rodrigues", last_view)`` (seq_utils.py:138-208).

    Returns ``dict(camera_params (Nc,6) f64 interleaved, points_3d (2Nl,3), camera_indices (m,) int64,
    lines_indices (m,) int64, n, p0, p1 (m,3) f64, line_ids (m,) int64)`` -- the same values, order and dtypes
    the reference's dict leads to after ``point_wise_input``.
    """
    import cv2

    views = domain.views[:last_view + 1]
    cams = np.empty((len(views), 3, 2), dtype=np.float64)
    for i, v in enumerate(views):
        cams[i, :, 0] = cv2.Rodrigues(np.asarray(v.P.R))[0].ravel()          # :179
        cams[i, :, 1] = np.asarray(v.P.c, dtype=np.float64).ravel()           # :172
    v_idx, t_idx, line_ids = observations_from_tables(domain.tracker.ids, domain.tracker.bin)   # :188-192
    lines2d = domain.lines2d
    n = np.array([lines2d[j].n for j in line_ids], dtype=np.float64).reshape(-1, 3)
    p0 = np.array([lines2d[j].points[0].p for j in line_ids], dtype=np.float64).reshape(-1, 3)
    p1 = np.array([lines2d[j].points[1].p for j in line_ids], dtype=np.float64).reshape(-1, 3)
    pts = np.array([np.asarray(L.points[e].coord)[:3] for L in domain.lines3d for e in (0, 1)])  # bundle_adjustment.py:231-233
    return dict(camera_params=cams.reshape(len(views), 6), points_3d=pts,
                camera_indices=v_idx.astype(np.int64), lines_indices=t_idx.astype(np.int64),
                n=n, p0=p0, p1=p1, line_ids=line_ids)


def update_domain_from_arrays(domain, op_camera_params, op_points_3d, last_view=2, camera_cls=None):
    """``update_structure_motion`` (seq_utils.py:211-252) from the solver's output arrays."""
    import cv2

    if camera_cls is None:
        camera_cls = type(domain.views[0].P)
    K = np.asarray(domain.K)
    for i, v in enumerate(domain.views):
        if i == last_view + 1:
            break
        cam = np.asarray(op_camera_params[i]).reshape(3, 2)
        R = np.array(cv2.Rodrigues(cam[:, 0])[0])                                # :240-241
        c = cam[:, 1]
        t = np.array([np.dot(R, -c)]).reshape((3, 1))                            # :243
        Pi = camera_cls(K @ np.concatenate((R, t), axis=1))
        Pi.factor()
        Pi.center()
        v.P = Pi
    for i, L in enumerate(domain.lines3d):
        L.points[0].coord = op_points_3d[2 * i]                                  # :250-252
        L.points[1].coord = op_points_3d[2 * i + 1]


def fundamental_from_cameras(K, P0, P1):
    """``relativeCameraMotion`` (utils.py:1483-1509) + ``F_from_KRt2`` (:1039-1065) for two reference ``Camera``
    objects: the same numpy expressions, so the same bits."""
    R0, t0, R1, t1 = P0.R, P0.t, P1.R, P1.t
    rR = R1 @ R0.T
    rt = (t1 - rR @ t0).reshape((3, 1))
    rt_x = np.array([[0, -rt[2][0], rt[1][0]], [rt[2][0], 0, -rt[0][0]], [-rt[1][0], rt[0][0], 0]])
    F = np.linalg.inv(K).T @ rt_x @ rR @ np.linalg.inv(K)
    return F / F[2, 2]


def tracker_arrays_from_domain(domain, last_view):
    """What ``Tracker.comp_tracker(domain, last_view)`` reads (tracker.py:44-55): per consecutive view pair the
    fundamental matrix and ``domain.line_match["<v><v+1>"][1]``; the endpoints of ``domain.lines2d`` by list index."""
    endpoints = np.array([[np.asarray(p.coord, dtype=np.float64)[:2] for p in L.points] for L in domain.lines2d])
    F, matches = [], []
    for v in range(last_view):
        F.append(fundamental_from_cameras(domain.K, domain.views[v].P, domain.views[v + 1].P))
        matches.append(np.array(domain.line_match[str(v) + str(v + 1)][1], dtype=np.int64).reshape(-1, 2))
    return dict(endpoints=endpoints, F=np.array(F), matches=matches)


class TrackTables:
    """Stand-in for the reference's ``Tracker`` after ``comp_tracker``: ``ids`` / ``bin`` as float64 arrays (the
    reference's dtype, tracker.py:39-41) and ``n_new_tracks`` as a list of ints."""

    def __init__(self, ids, bi, n_new):
        self.ids = np.asarray(ids, dtype=np.float64)
        self.bin = np.asarray(bi, dtype=np.float64)
        self.n_new_tracks = [int(x) for x in n_new]


def comp_tracker(domain, last_view=2, ba=None, threshold=0.1):
    """``Tracker().comp_tracker(domain, last_view)`` on the GPU; returns a ``TrackTables`` to assign to
    ``domain.tracker`` (main.py:96-98)."""
    a = tracker_arrays_from_domain(domain, last_view)
    own = ba is None
    ba = LineBA(0) if own else ba
    try:
        ids, bi, n_new = ba.track_table(a["endpoints"], a["matches"], a["F"], threshold)
    finally:
        if own:
            ba.close()
    return TrackTables(ids, bi, n_new)


def triangulation_arrays_from_domain(domain, last_view):
    """The arrays the reference's ``triangulate_model`` loop feeds to OpenCV (seq_utils.py:105-117): for every new
    track of every consecutive view pair, the pair's first view and the two float32 endpoint arrays.
    Returns ``dict(P (V,3,4) f64, view0 (n,) int64, pts0, pts1 (n,2,2) f32)``."""
    P = np.array([np.asarray(v.P.P, dtype=np.float64) for v in domain.views[:last_view + 1]])
    ids = np.asarray(domain.tracker.ids)
    view0, l0, l1 = [], [], []
    tr = 0
    for i in range(last_view):                                               # :105-112
        k = int(domain.tracker.n_new_tracks[i])
        pair = ids[tr:tr + k, i:i + 2].astype(np.int64)
        view0.append(np.full(k, i, dtype=np.int64)); l0.append(pair[:, 0]); l1.append(pair[:, 1])
        tr += k
    view0 = np.concatenate(view0) if view0 else np.empty(0, dtype=np.int64)
    l0 = np.concatenate(l0) if l0 else np.empty(0, dtype=np.int64)
    l1 = np.concatenate(l1) if l1 else np.empty(0, dtype=np.int64)
    lines2d = domain.lines2d
    pts0 = np.array([[p.coord for p in lines2d[j].points] for j in l0], dtype=np.float32).reshape(-1, 2, 2)   # :114-115
    pts1 = np.array([[p.coord for p in lines2d[j].points] for j in l1], dtype=np.float32).reshape(-1, 2, 2)   # :116-117
    return dict(P=P, view0=view0, pts0=pts0, pts1=pts1)


def triangulate_model(domain, last_view, ba=None, point3d_cls=None, line3d_cls=None):
    """``triangulate_model(domain, last_view)`` (seq_utils.py:75-135) with the per-line ``cv2.triangulatePoints``
    loop replaced by one batched GPU call: rebuilds ``domain.points3d`` / ``domain.lines3d`` with the same ids,
    order and float32 coordinates.  ``point3d_cls`` / ``line3d_cls`` default to the reference's ``Point3D`` /
    ``Line3D`` (``point.py:14``, ``line.py:50``; they must be importable, as they are inside the reference)."""
    if point3d_cls is None:
        from point import Point3D as point3d_cls
    if line3d_cls is None:
        from line import Line3D as line3d_cls
    a = triangulation_arrays_from_domain(domain, last_view)
    own = ba is None
    ba = LineBA(0) if own else ba
    try:
        X = ba.triangulate_lines(a["P"], a["view0"], a["pts0"], a["pts1"])
    finally:
        if own:
            ba.close()
    domain.lines3d, domain.points3d = [], []
    for t in range(X.shape[0]):                                               # :122-131
        p_ini = point3d_cls(2 * t, X[t, 0])
        p_fin = point3d_cls(2 * t + 1, X[t, 1])
        domain.points3d += [p_ini, p_fin]
        domain.lines3d.append(line3d_cls(t, [p_ini, p_fin]))
    return X


def lines2d_arrays_from_domain(domain, start=0):
    """Flatten ``domain.lines2d[start:]`` once: ``n``, ``p0``, ``p1`` (k,3) f64 (``Line.n`` / ``Line.points[e].p``,
    domain.py:338-351) and the pixel endpoints (k,2,2) f64."""
    L = domain.lines2d[start:]
    n = np.array([l.n for l in L], dtype=np.float64).reshape(-1, 3)
    p0 = np.array([l.points[0].p for l in L], dtype=np.float64).reshape(-1, 3)
    p1 = np.array([l.points[1].p for l in L], dtype=np.float64).reshape(-1, 3)
    E = np.array([[np.asarray(p.coord, dtype=np.float64)[:2] for p in l.points] for l in L], dtype=np.float64).reshape(-1, 2, 2)
    return n, p0, p1, E


class IncrementalBA:
    """One GPU context for the whole incremental reconstruction (src/main.py:111-133 re-runs tracker, triangulation
    and BA after every added view on a growing problem).

    ``solve`` is the array-native drop-in for ``input_bundle_adj`` + ``run_bundle_adjustment`` +
    ``update_structure_motion`` on whatever ``domain.tracker`` / ``domain.lines3d`` currently hold.

    ``step`` is the device-resident version of main.py:118-133 (everything after the PnL registration of the new
    view): the ``Line`` objects are flattened and uploaded ONCE (``lines2d_arrays_from_domain``; only a grown tail
    is flattened again), and per added view the track table, the triangulated endpoints, the observation list and
    the observation constants are produced and consumed on the device -- what crosses PCIe per view is the match
    lists of the new pair, the fundamental matrices and camera rows in, and the optimised cameras, 3D endpoints and
    the track table out (the reference's PnL stage reads them from the ``Domain`` objects)."""

    def __init__(self, device=0):
        self.ba = LineBA(device)
        self.last = None
        self._l2d = None          # host copies of the flattened 2D-line arrays (n, p0, p1, endpoints)
        self._matches = {}        # view pair -> (k,2) int64, flattened once

    # -- array-native BA on the domain's current tracker / lines3d ------------------------------------------------
    def solve(self, domain, last_view, loss_f="linear", write_back=True, **options):
        a = input_arrays_from_domain(domain, last_view)
        nc = a["camera_params"].shape[0]
        nl = a["points_3d"].shape[0] // 2
        self.ba.set_problem(nc, nl, a["camera_indices"], a["lines_indices"], a["n"], a["p0"], a["p1"])
        res = self.ba.lm_solve(a["camera_params"], np.asarray(a["points_3d"], dtype=np.float64), loss=loss_f, **options)
        if write_back:
            update_domain_from_arrays(domain, res.camera_params, res.points_3d, last_view)
        self.last = res
        return res.camera_params, res.points_3d, res.residuals

    # -- device-resident incremental step ----------------------------------------------------------------------------
    def _sync_lines2d(self, domain):
        have = 0 if self._l2d is None else self._l2d[0].shape[0]
        if have == len(domain.lines2d):
            return
        tail = lines2d_arrays_from_domain(domain, have)
        self._l2d = tail if self._l2d is None else tuple(np.concatenate((x, y)) for x, y in zip(self._l2d, tail))
        self.ba.set_lines2d(*self._l2d)

    def step(self, domain, last_view, loss_f="linear", threshold=0.1, write_back=True, point3d_cls=None, line3d_cls=None,
             **options):
        """Tracker -> triangulation -> BA -> update for views 0..last_view (main.py:118-133), device-resident."""
        import cv2
        self._sync_lines2d(domain)
        views = domain.views[:last_view + 1]
        F, matches = [], []
        for v in range(last_view):
            F.append(fundamental_from_cameras(domain.K, views[v].P, views[v + 1].P))
            key = str(v) + str(v + 1)
            if key not in self._matches:
                self._matches[key] = np.array(domain.line_match[key][1], dtype=np.int64).reshape(-1, 2)
            matches.append(self._matches[key])
        ids, bi, n_new = self.ba.track_table_device(matches, np.array(F), threshold)              # tracker.py:11-96
        P = np.array([np.asarray(v.P.P, dtype=np.float64) for v in views])
        X, pts = self.ba.triangulate_tracks(P, ids, n_new)                                       # seq_utils.py:75-135
        cam_idx, line_idx, l2d, n_tracks = self.ba.observations_from_tables(ids, bi)             # seq_utils.py:186-195
        cams = np.empty((len(views), 3, 2), dtype=np.float64)
        for i, v in enumerate(views):
            cams[i, :, 0] = cv2.Rodrigues(np.asarray(v.P.R))[0].ravel()                          # seq_utils.py:179
            cams[i, :, 1] = np.asarray(v.P.c, dtype=np.float64).ravel()
        cams = cams.reshape(len(views), 6)
        nl = int(X.shape[0])
        self.ba.set_problem_lines2d(len(views), nl, cam_idx, line_idx, l2d)
        res = self.ba.lm_solve(None, None, out=(cams, pts), loss=loss_f, **options)              # bundle_adjustment.py:254
        res.camera_params = cams
        res.points_3d = pts.cpu().numpy()
        if write_back:
            if point3d_cls is None:
                from point import Point3D as point3d_cls
            if line3d_cls is None:
                from line import Line3D as line3d_cls
            domain.tracker = TrackTables(ids.cpu().numpy(), bi.cpu().numpy(), n_new)             # main.py:121-123
            Xh = X.cpu().numpy()
            domain.lines3d, domain.points3d = [], []
            for t in range(nl):                                                                   # seq_utils.py:122-131
                p_ini = point3d_cls(2 * t, Xh[t, 0]); p_fin = point3d_cls(2 * t + 1, Xh[t, 1])
                domain.points3d += [p_ini, p_fin]
                domain.lines3d.append(line3d_cls(t, [p_ini, p_fin]))
            update_domain_from_arrays(domain, res.camera_params, res.points_3d, last_view)      # seq_utils.py:211-252
        self.last = res
        self.last_tables = (ids, bi, n_new)
        return res.camera_params, res.points_3d, res.residuals

    def close(self):
        self.ba.close()
