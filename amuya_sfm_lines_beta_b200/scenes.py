"""Synthetic line-SfM scenes for tests and benchmarks.

The scenes are flat arrays in the reference's data model: interleaved camera rows
``[w0,c0,w1,c1,w2,c2]`` (``seq_utils.py:179-181``), endpoint rows ``2L+e``
(``bundle_adjustment.py:231-233``), observations sorted by (line, camera) as
``input_bundle_adj`` emits them (``seq_utils.py:188-195``), unit bearings ``p`` and
unit normals ``n`` as ``Domain.spherical_representation`` defines them
(``domain.py:338-351``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Scene:
    camera_params: np.ndarray     # (Nc,6) f64 interleaved, initial guess
    points_3d: np.ndarray         # (2Nl,3) f64, initial guess
    camera_indices: np.ndarray    # (m,) int64
    lines_indices: np.ndarray     # (m,) int64
    n: np.ndarray                 # (m,3) unit normal of the observed 2D line
    p0: np.ndarray                # (m,3) unit bearing of observed endpoint 0
    p1: np.ndarray                # (m,3)
    gt_camera_params: np.ndarray | None = None
    gt_points_3d: np.ndarray | None = None
    name: str = ""

    @property
    def n_cameras(self):
        return self.camera_params.shape[0]

    @property
    def n_lines(self):
        return self.points_3d.shape[0] // 2

    @property
    def n_obs(self):
        return self.camera_indices.shape[0]

    def arrays(self):
        return (self.camera_params, self.points_3d, self.camera_indices, self.lines_indices,
                self.n, self.p0, self.p1)


def _rotvec_to_matrix(w):
    """Rodrigues formula, same closed form as cv2.Rodrigues (bundle_adjustment.py:48)."""
    w = np.asarray(w, dtype=np.float64).reshape(-1, 3)
    th = np.linalg.norm(w, axis=1)
    small = th < np.finfo(np.float64).eps
    r = w / np.where(small, 1.0, th)[:, None]
    c, s = np.cos(th), np.sin(th)
    K = np.zeros((w.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -r[:, 2], r[:, 1]
    K[:, 1, 0], K[:, 1, 2] = r[:, 2], -r[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -r[:, 1], r[:, 0]
    R = c[:, None, None] * np.eye(3) + (1 - c)[:, None, None] * (r[:, :, None] * r[:, None, :]) \
        + s[:, None, None] * K
    R[small] = np.eye(3)
    return R


def _sample_visibility(rng, n_lines, n_cams, k, variant):
    if k > n_cams:
        raise ValueError("obs_per_line exceeds the number of cameras")
    if variant == "B":  # contiguous window of camera ids (locality)
        start = rng.integers(0, n_cams - k + 1, size=n_lines)
        return start[:, None] + np.arange(k)[None, :]
    cams = np.sort(rng.integers(0, n_cams, size=(n_lines, k)), axis=1)
    while True:
        dup = np.nonzero((np.diff(cams, axis=1) == 0).any(axis=1))[0]
        if dup.size == 0:
            return cams
        cams[dup] = np.sort(rng.integers(0, n_cams, size=(dup.size, k)), axis=1)


def make_synthetic_scene(n_cams, n_lines, obs_per_line=10, seed=0, variant="A",
                         noise=1e-3, slide=0.1, sig_c=1e-2, sig_w=5e-3, sig_x=2e-2):
    """Seeded synthetic scene of the shape BASELINE.json names.

    Config 4 is ``make_synthetic_scene(1000, 200_000, 10)`` (2 M observations),
    config 5 ``make_synthetic_scene(10_000, 2_000_000, 10)``.
    """
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(seed)
    # 3D segments
    X0 = rng.uniform(-2, 2, (n_lines, 3))
    X1 = X0 + rng.uniform(-1, 1, (n_lines, 3))
    while True:
        bad = np.nonzero(np.linalg.norm(X1 - X0, axis=1) < 0.1)[0]
        if bad.size == 0:
            break
        X1[bad] = X0[bad] + rng.uniform(-1, 1, (bad.size, 3))
    # cameras on a shell looking at the origin
    dirs = rng.normal(size=(n_cams, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    centre = dirs * rng.uniform(8, 12, n_cams)[:, None]
    z = -centre / np.linalg.norm(centre, axis=1)[:, None] + rng.normal(0, 0.05, (n_cams, 3))
    z /= np.linalg.norm(z, axis=1)[:, None]
    up = np.tile(np.array([0.0, 0.0, 1.0]), (n_cams, 1))
    par = np.abs(np.einsum('ni,ni->n', up, z)) > 0.95
    up[par] = np.array([0.0, 1.0, 0.0])
    x = np.cross(up, z); x /= np.linalg.norm(x, axis=1)[:, None]
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)             # rows are the camera axes: x_cam = R (X - c)
    omega = Rotation.from_matrix(R).as_rotvec()
    R = _rotvec_to_matrix(omega)

    if np.ndim(obs_per_line) == 0:
        cam_of = _sample_visibility(rng, n_lines, n_cams, obs_per_line, variant)
        li = np.repeat(np.arange(n_lines, dtype=np.int64), obs_per_line)
        ci = cam_of.reshape(-1).astype(np.int64)
    else:
        # a different number of (distinct) observing cameras per line, e.g. a few lines seen from hundreds of views
        ks = np.asarray(obs_per_line, dtype=np.int64)
        if ks.shape != (n_lines,) or ks.max() > n_cams:
            raise ValueError("obs_per_line must be an int or one count <= n_cams per line")
        li = np.repeat(np.arange(n_lines, dtype=np.int64), ks)
        ci = np.concatenate([np.sort(rng.choice(n_cams, int(k), replace=False)) for k in ks]).astype(np.int64)
    m = li.shape[0]

    # observed endpoints slide along the 3D segment (relaxed endpoint constraint)
    s0 = rng.uniform(-slide, slide, m)
    s1 = rng.uniform(-slide, slide, m)
    dX = (X1 - X0)[li]
    E0 = X0[li] + s0[:, None] * dX
    E1 = X1[li] - s1[:, None] * dX
    Rm, cm = R[ci], centre[ci]
    p0 = np.einsum('nij,nj->ni', Rm, E0 - cm)
    p1 = np.einsum('nij,nj->ni', Rm, E1 - cm)
    p0 /= np.linalg.norm(p0, axis=1)[:, None]
    p1 /= np.linalg.norm(p1, axis=1)[:, None]
    p0 += rng.normal(0, noise, p0.shape); p0 /= np.linalg.norm(p0, axis=1)[:, None]
    p1 += rng.normal(0, noise, p1.shape); p1 /= np.linalg.norm(p1, axis=1)[:, None]
    nrm = np.cross(p0, p1)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]

    gt_cams = np.empty((n_cams, 3, 2)); gt_cams[:, :, 0] = omega; gt_cams[:, :, 1] = centre
    gt_pts = np.empty((2 * n_lines, 3)); gt_pts[0::2] = X0; gt_pts[1::2] = X1
    cams0 = gt_cams.copy()
    cams0[:, :, 0] += rng.normal(0, sig_w, (n_cams, 3))
    cams0[:, :, 1] += rng.normal(0, sig_c, (n_cams, 3))
    pts0 = gt_pts + rng.normal(0, sig_x, gt_pts.shape)
    return Scene(camera_params=cams0.reshape(n_cams, 6), points_3d=pts0, camera_indices=ci,
                 lines_indices=li, n=nrm, p0=p0, p1=p1, gt_camera_params=gt_cams.reshape(n_cams, 6),
                 gt_points_3d=gt_pts,
                 name=f"synthetic_{n_cams}c_{n_lines}l_{m}o_seed{seed}_{variant}")


def scene_from_npz(path):
    """Load a golden fixture (tests/golden/*.npz) as a Scene.  The float32 endpoints the
    reference triangulates are promoted to float64 as ``np.hstack`` does
    (bundle_adjustment.py:311)."""
    q = np.load(path)
    pts = np.asarray(q["points_3d_raw"], dtype=np.float64)
    return Scene(camera_params=np.asarray(q["camera_params"], dtype=np.float64), points_3d=pts,
                 camera_indices=np.asarray(q["camera_indices"], dtype=np.int64),
                 lines_indices=np.asarray(q["lines_indices"], dtype=np.int64),
                 n=np.asarray(q["n"]), p0=np.asarray(q["p0"]), p1=np.asarray(q["p1"]),
                 name=str(path))
