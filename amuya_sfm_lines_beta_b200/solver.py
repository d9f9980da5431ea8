"""Array-native host API over liblba.so: one ``LineBA`` object per GPU.

This is the fast path below the drop-in ``run_bundle_adjustment`` (which only has to
flatten the reference's Python objects first).  Arrays use the reference's layouts
(see include/lba.h); numpy arrays and torch tensors (CPU or CUDA) are accepted
wherever a pointer is expected, so device-resident data is never copied to the host.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import LbaOptions, LbaReport, check

STATUS_TEXT = {
    0: "The maximum number of iterations is exceeded.",
    1: "`gtol` termination condition is satisfied.",
    2: "`ftol` termination condition is satisfied.",
    3: "`xtol` termination condition is satisfied.",
    4: "Both `ftol` and `xtol` termination conditions are satisfied.",
}


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _as_ptr(x, np_dtype, torch_dtype_name, shape_size=None, ctx=None):
    """Return (address, keepalive) for a C-contiguous numpy array or torch tensor.  A CUDA tensor must live on the
    context's device; the context's stream is ordered behind the torch stream that (may have) produced it, so a
    tensor written by still-running torch kernels -- or the converted temporary made here -- is complete before
    liblba reads it."""
    if x is None:
        return None, None
    if _is_torch(x):
        import torch
        want = getattr(torch, torch_dtype_name)
        if x.is_cuda and ctx is not None and x.device.index != ctx.device:
            raise ValueError(f"tensor lives on cuda:{x.device.index}, the context on cuda:{ctx.device}")
        t = x if (x.dtype == want and x.is_contiguous()) else x.to(want).contiguous()
        if shape_size is not None and t.numel() != shape_size:
            raise ValueError(f"expected {shape_size} elements, got {t.numel()}")
        if t.is_cuda and ctx is not None:
            check(ctx._lib.lba_wait_stream(ctx._h, torch.cuda.current_stream(t.device).cuda_stream))
        return t.data_ptr(), t
    a = np.ascontiguousarray(x, dtype=np_dtype)
    if shape_size is not None and a.size != shape_size:
        raise ValueError(f"expected {shape_size} elements, got {a.size}")
    return a.ctypes.data, a


@dataclass
class LMResult:
    camera_params: np.ndarray
    points_3d: np.ndarray
    residuals: np.ndarray | None
    iterations: int
    status: int
    message: str
    n_jac: int
    n_res: int
    pcg_iters_total: int
    chol_clamped: int
    n_accepted: int
    cost_init: float
    cost: float
    seconds: float
    trace: list = field(default_factory=list)


class LineBA:
    """One bundle-adjustment context bound to one CUDA device."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.lba_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.m = self.n_cameras = self.n_lines = 0
        self.rank, self.world = 0, 1

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- problem ------------------------------------------------------------------
    def set_problem(self, n_cameras, n_lines, camera_indices, lines_indices, n, p0, p1):
        m = int(camera_indices.shape[0])
        pc, k0 = _as_ptr(camera_indices, np.int64, "int64", m, ctx=self)
        pl, k1 = _as_ptr(lines_indices, np.int64, "int64", m, ctx=self)
        pn, k2 = _as_ptr(n, np.float64, "float64", 3 * m, ctx=self)
        pa, k3 = _as_ptr(p0, np.float64, "float64", 3 * m, ctx=self)
        pb, k4 = _as_ptr(p1, np.float64, "float64", 3 * m, ctx=self)
        check(self._lib.lba_set_problem(self._h, m, int(n_cameras), int(n_lines), pc, pl, pn, pa, pb))
        self.m, self.n_cameras, self.n_lines = m, int(n_cameras), int(n_lines)
        return self

    @staticmethod
    def _cam_width(camera_params):
        """6 (Rodrigues vector | centre, interleaved) or 12 ([R|c] rows, bundle_adjustment.py:43-45)."""
        n = camera_params.numel() if _is_torch(camera_params) else np.asarray(camera_params).size
        shape = tuple(camera_params.shape)
        if len(shape) >= 2 and shape[-1] in (6, 12):
            return int(shape[-1])
        return 6

    def _params(self, camera_params, points_3d, width=6):
        pc, kc = _as_ptr(camera_params, np.float64, "float64", width * self.n_cameras, ctx=self)
        pp, kp = _as_ptr(points_3d, np.float64, "float64", 6 * self.n_lines, ctx=self)
        return pc, pp, (kc, kp)

    def point_wise_indices(self):
        cam_pt = np.empty(2 * self.m, dtype=np.int64)
        pt_idx = np.empty(2 * self.m, dtype=np.int64)
        check(self._lib.lba_point_wise_indices(self._h, cam_pt.ctypes.data, pt_idx.ctypes.data))
        return cam_pt, pt_idx

    # -- track table (independent of set_problem) ------------------------------
    def track_table(self, endpoints, matches_per_pair, F_per_pair, threshold=0.1, flags_in=None, want_flags=False):
        """``Tracker.comp_tracker`` (src/tracker.py:11-96) with ``verify_duplet`` (src/utils.py:1807-1944) on the GPU.
        ``endpoints`` (n_lines2d,2,2) f64 pixel endpoints of ``domain.lines2d[id]``; ``matches_per_pair[v]`` the (k,2)
        int64 matches of views (v, v+1); ``F_per_pair[v]`` their 3x3 fundamental matrix.  Returns
        ``(ids (T,V) int64, bin (T,V) int64, n_new_tracks (V-1,) int64[, flags (M,) bool])``."""
        n_pairs = len(matches_per_pair)
        V = n_pairs + 1
        ms = [np.ascontiguousarray(m, dtype=np.int64).reshape(-1, 2) for m in matches_per_pair]
        ptr = np.zeros(V, dtype=np.int64)
        ptr[1:] = np.cumsum([len(m) for m in ms])
        M = int(ptr[-1])
        matches = np.concatenate(ms) if M else np.zeros((0, 2), dtype=np.int64)
        E = np.ascontiguousarray(endpoints, dtype=np.float64)
        F = np.ascontiguousarray(np.asarray(F_per_pair, dtype=np.float64).reshape(n_pairs, 9))
        fin = None if flags_in is None else np.ascontiguousarray(np.asarray(flags_in).astype(np.uint8))
        if fin is not None and fin.size != M:
            raise ValueError(f"flags_in must have one entry per match ({M})")
        cap = max(M, 1)
        ids = np.empty((cap, V), dtype=np.int64)
        bi = np.empty((cap, V), dtype=np.int64)
        n_new = np.zeros(n_pairs, dtype=np.int64)
        flags = np.zeros(M, dtype=np.uint8) if want_flags else None
        n_tracks = C.c_int64()
        check(self._lib.lba_track_table(self._h, V, int(E.shape[0]), E.ctypes.data, ptr.ctypes.data,
                                        matches.ctypes.data if M else None, F.ctypes.data, float(threshold),
                                        fin.ctypes.data if fin is not None else None, cap, ids.ctypes.data, bi.ctypes.data,
                                        n_new.ctypes.data, C.byref(n_tracks), flags.ctypes.data if want_flags else None))
        T = int(n_tracks.value)
        out = (ids[:T].copy(), bi[:T].copy(), n_new)
        return out + (flags.astype(bool),) if want_flags else out

    # -- BA input resident on the device across the incremental loop ------------
    def set_lines2d(self, n, p0, p1, endpoints=None):
        """Upload ONCE the per-2D-line constants ``Line.n`` / ``Line.points[e].p`` (domain.py:338-351) and, optionally,
        the pixel endpoints of every ``domain.lines2d[id]``; later calls name 2D lines by id."""
        k = int(np.asarray(n).shape[0]) if not _is_torch(n) else int(n.shape[0])
        pn, k0 = _as_ptr(n, np.float64, "float64", 3 * k, ctx=self)
        pa, k1 = _as_ptr(p0, np.float64, "float64", 3 * k, ctx=self)
        pb, k2 = _as_ptr(p1, np.float64, "float64", 3 * k, ctx=self)
        pe, k3 = _as_ptr(endpoints, np.float64, "float64", 4 * k, ctx=self)
        check(self._lib.lba_lines2d_set(self._h, k, pn, pa, pb, pe))
        self.n_lines2d = k
        return self

    def track_table_device(self, matches_per_pair, F_per_pair, threshold=0.1):
        """``track_table`` on the resident endpoint table with the tables LEFT on the device: returns
        ``(ids, bin)`` torch CUDA int64 tensors of shape (T, V) (views of the capacity buffers) and ``n_new_tracks``."""
        import torch
        n_pairs = len(matches_per_pair)
        V = n_pairs + 1
        ms = [np.ascontiguousarray(m, dtype=np.int64).reshape(-1, 2) for m in matches_per_pair]
        ptr = np.zeros(V, dtype=np.int64)
        ptr[1:] = np.cumsum([len(m) for m in ms])
        M = int(ptr[-1])
        matches = np.concatenate(ms) if M else np.zeros((0, 2), dtype=np.int64)
        F = np.ascontiguousarray(np.asarray(F_per_pair, dtype=np.float64).reshape(n_pairs, 9))
        cap = max(M, 1)
        dev = torch.device("cuda", self.device)
        ids = torch.empty((cap, V), dtype=torch.int64, device=dev)
        bi = torch.empty((cap, V), dtype=torch.int64, device=dev)
        check(self._lib.lba_wait_stream(self._h, torch.cuda.current_stream(dev).cuda_stream))
        n_new = np.zeros(n_pairs, dtype=np.int64)
        n_tracks = C.c_int64()
        check(self._lib.lba_track_table(self._h, V, 0, None, ptr.ctypes.data, matches.ctypes.data if M else None,
                                        F.ctypes.data, float(threshold), None, cap, ids.data_ptr(), bi.data_ptr(),
                                        n_new.ctypes.data, C.byref(n_tracks), None))
        T = int(n_tracks.value)
        return ids[:T], bi[:T], n_new

    def observations_from_tables(self, ids, bi):
        """The observation list of ``input_bundle_adj`` (seq_utils.py:186-195) from the track tables.  Device tables
        (torch CUDA int64) give device results ``(camera_indices, lines_indices, line2d_ids)``; host tables host results.
        Also returns ``n_tracks`` (rows seen in more than one view)."""
        on_dev = _is_torch(ids) and ids.is_cuda
        T, V = int(ids.shape[0]), int(ids.shape[1])
        m, nt = C.c_int64(), C.c_int64()
        cap = max(T * V, 1)
        if on_dev:
            import torch
            pi, k0 = _as_ptr(ids, np.int64, "int64", T * V, ctx=self)
            pb, k1 = _as_ptr(bi, np.int64, "int64", T * V, ctx=self)
            out = [torch.empty(cap, dtype=torch.int64, device=ids.device) for _ in range(3)]
            check(self._lib.lba_wait_stream(self._h, torch.cuda.current_stream(ids.device).cuda_stream))
            ptrs = [o.data_ptr() for o in out]
        else:
            a = np.ascontiguousarray(ids, dtype=np.int64); b = np.ascontiguousarray(bi, dtype=np.int64)
            pi, pb = a.ctypes.data, b.ctypes.data
            out = [np.empty(cap, dtype=np.int64) for _ in range(3)]
            ptrs = [o.ctypes.data for o in out]
        check(self._lib.lba_observations_from_tables(self._h, T, V, pi, pb, cap, ptrs[0], ptrs[1], ptrs[2],
                                                     C.byref(m), C.byref(nt)))
        k = int(m.value)
        return out[0][:k], out[1][:k], out[2][:k], int(nt.value)

    def set_problem_lines2d(self, n_cameras, n_lines, camera_indices, lines_indices, line2d_ids):
        """``set_problem`` with n, p0, p1 gathered on the device from the table given to ``set_lines2d``."""
        m = int(camera_indices.shape[0])
        pc, k0 = _as_ptr(camera_indices, np.int64, "int64", m, ctx=self)
        pl, k1 = _as_ptr(lines_indices, np.int64, "int64", m, ctx=self)
        pi, k2 = _as_ptr(line2d_ids, np.int64, "int64", m, ctx=self)
        check(self._lib.lba_set_problem_lines2d(self._h, m, int(n_cameras), int(n_lines), pc, pl, pi))
        self.m, self.n_cameras, self.n_lines = m, int(n_cameras), int(n_lines)
        return self

    def triangulate_tracks(self, P, ids, n_new_tracks, want_points64=True):
        """``triangulate_model`` (seq_utils.py:105-131) from the DEVICE track table ``ids`` and the resident endpoint
        table.  Returns ``(X (n,2,3) float32, points_3d (2n,3) float64)`` as torch CUDA tensors (points_3d: the float32
        coordinates widened to double, the BA's starting point)."""
        import torch
        V = int(P.shape[0])
        pP, k0 = _as_ptr(P, np.float64, "float64", 12 * V, ctx=self)
        nn = np.ascontiguousarray(n_new_tracks, dtype=np.int64)
        n = int(nn.sum())
        dev = ids.device
        X = torch.empty((n, 2, 3), dtype=torch.float32, device=dev)
        pts = torch.empty((2 * n, 3), dtype=torch.float64, device=dev) if want_points64 else None
        check(self._lib.lba_wait_stream(self._h, torch.cuda.current_stream(dev).cuda_stream))
        nt = C.c_int64()
        base = ids._base if ids._base is not None else ids      # the full capacity buffer: rows are V apart
        check(self._lib.lba_triangulate_tracks(self._h, V, pP, base.data_ptr(), nn.ctypes.data, X.data_ptr() if n else None,
                                               pts.data_ptr() if (pts is not None and n) else None, C.byref(nt)))
        return X, pts

    # -- two-view endpoint triangulation (independent of set_problem) ------
    def triangulate_lines(self, P, view0, pts0, pts1, view1=None, want_homogeneous=False, out=None):
        """``cv2.triangulatePoints`` + float32 dehomogenisation for every endpoint of every track
        (reference ``triangulate_model``, src/seq_utils.py:105-127).  ``P`` (V,3,4) f64; ``view0`` (n,) int64
        (``view1`` defaults to ``view0 + 1``); ``pts0``, ``pts1`` (n,2,2) f32 ``[track][endpoint][x,y]``.
        Returns ``X`` (n,2,3) float32 (and ``Xh`` (n,2,4) when asked).  numpy or torch (CPU / CUDA) inputs;
        ``out`` may be a preallocated float32 array / CUDA tensor of n*6 elements.  The device time of the
        kernel is left in ``self.last_triangulate_seconds``."""
        n = int(view0.shape[0])
        n_views = int(P.shape[0])
        pP, k0 = _as_ptr(P, np.float64, "float64", 12 * n_views, ctx=self)
        pv0, k1 = _as_ptr(view0, np.int64, "int64", n, ctx=self)
        pv1, k2 = _as_ptr(view1, np.int64, "int64", n, ctx=self)
        pa, k3 = _as_ptr(pts0, np.float32, "float32", 4 * n, ctx=self)
        pb, k4 = _as_ptr(pts1, np.float32, "float32", 4 * n, ctx=self)
        if out is None:
            X = np.empty((n, 2, 3), dtype=np.float32)
            pX = X.ctypes.data
        else:
            X = out
            pX = out.data_ptr() if _is_torch(out) else out.ctypes.data
            if (out.numel() if _is_torch(out) else out.size) != 6 * n:
                raise ValueError(f"out must hold {6 * n} float32 values")
        Xh = np.empty((n, 2, 4), dtype=np.float32) if want_homogeneous else None
        sec = C.c_double()
        check(self._lib.lba_triangulate_lines(self._h, n, n_views, pP, pv0, pv1, pa, pb, pX,
                                              Xh.ctypes.data if Xh is not None else None, C.byref(sec)))
        self.last_triangulate_seconds = sec.value
        return (X, Xh) if want_homogeneous else X

    # -- evaluation -----------------------------------------------------------------
    def project(self, camera_params, points_3d):
        pc, pp, keep = self._params(camera_params, points_3d)
        out = np.empty((2 * self.m, 3), dtype=np.float64)
        check(self._lib.lba_project(self._h, pc, pp, out.ctypes.data))
        return out

    def residuals(self, camera_params, points_3d, want_residuals=True):
        """``fun`` (bundle_adjustment.py:78-139) and the cost; (Nc,6) Rodrigues rows or (Nc,12) [R|c] rows."""
        w = self._cam_width(camera_params)
        pc, pp, keep = self._params(camera_params, points_3d, w)
        r = np.empty(self.m, dtype=np.float64) if want_residuals else None
        cost = C.c_double()
        fn = self._lib.lba_residuals12 if w == 12 else self._lib.lba_residuals
        check(fn(self._h, pc, pp, r.ctypes.data if r is not None else None, C.byref(cost)))
        return r, cost.value

    def project12(self, points, camera_rows):
        """``project`` for [R|c] rows (bundle_adjustment.py:43-63): ``points`` (N,3), ``camera_rows`` (N,12)."""
        N = int(np.asarray(points).shape[0])
        pp, k0 = _as_ptr(points, np.float64, "float64", 3 * N, ctx=self)
        pc, k1 = _as_ptr(camera_rows, np.float64, "float64", 12 * N, ctx=self)
        out = np.empty((N, 3), dtype=np.float64)
        check(self._lib.lba_project12(self._h, N, pp, pc, out.ctypes.data))
        return out

    def residuals_jacobian(self, camera_params, points_3d, want_jacobian=True):
        pc, pp, keep = self._params(camera_params, points_3d)
        r = np.empty(self.m, dtype=np.float64)
        jc = np.empty((self.m, 6), dtype=np.float64) if want_jacobian else None
        jl = np.empty((self.m, 6), dtype=np.float64) if want_jacobian else None
        check(self._lib.lba_residuals_jacobian(self._h, pc, pp, r.ctypes.data,
                                               jc.ctypes.data if want_jacobian else None,
                                               jl.ctypes.data if want_jacobian else None))
        return r, jc, jl

    def normal_equations(self):
        U = np.empty((self.n_cameras, 21)); gc = np.empty((self.n_cameras, 6))
        V = np.empty((self.n_lines, 21)); gl = np.empty((self.n_lines, 6))
        check(self._lib.lba_normal_equations(self._h, U.ctypes.data, gc.ctypes.data, V.ctypes.data, gl.ctypes.data))
        return U, gc, V, gl

    def schur_solve(self, lam, pcg_tol=1e-2, pcg_max_iter=200):
        dc = np.empty((self.n_cameras, 6)); dl = np.empty((self.n_lines, 6))
        it = C.c_int32()
        check(self._lib.lba_schur_solve(self._h, float(lam), float(pcg_tol), int(pcg_max_iter),
                                        dc.ctypes.data, dl.ctypes.data, C.byref(it)))
        return dc, dl, it.value

    # -- LM -----------------------------------------------------------------------------
    def default_options(self, **kw) -> LbaOptions:
        o = LbaOptions()
        self._lib.lba_default_options(C.byref(o))
        for k, v in kw.items():
            if k == "loss" and isinstance(v, str):
                if v not in _lib.LOSSES:
                    raise ValueError("`loss` must be one of {0} or a callable.".format(tuple(_lib.LOSSES)))
                v = _lib.LOSSES[v]
            if not hasattr(o, k):
                raise TypeError(f"unknown option {k!r}")
            setattr(o, k, v)
        return o

    def lm_solve(self, camera_params, points_3d, want_residuals=True, out=None, residuals_out=None, **options) -> LMResult:
        """Optimise cameras and 3D lines.  ``camera_params`` (Nc,6) interleaved and
        ``points_3d`` (2Nl,3) are not modified; optimised copies are returned (host numpy),
        unless ``out=(cams_buf, pts_buf)`` gives buffers (numpy or torch, any device) that
        are read as the initial guess and overwritten in place; ``residuals_out`` likewise receives ``fun`` at the
        solution (a pinned buffer avoids the pageable-memory staging of a fresh array per call)."""
        o = self.default_options(**options)
        w = self._cam_width(out[0] if out is not None else camera_params)
        if out is None:
            cams = np.array(camera_params, dtype=np.float64, order="C", copy=True).reshape(self.n_cameras, w)
            pts = np.array(points_3d, dtype=np.float64, order="C", copy=True).reshape(2 * self.n_lines, 3)
            pc, pp = cams.ctypes.data, pts.ctypes.data
        else:
            cams, pts = out
            pc, _k0 = _as_ptr(cams, np.float64, "float64", w * self.n_cameras, ctx=self)
            pp, _k1 = _as_ptr(pts, np.float64, "float64", 6 * self.n_lines, ctx=self)
            if (_k0 is not cams) or (_k1 is not pts):
                raise ValueError("out buffers must be contiguous float64")
        if residuals_out is not None:
            # caller-provided (m,) float64 buffer for fun at the solution, e.g. a pinned tensor reused across calls
            pr, r = _as_ptr(residuals_out, np.float64, "float64", self.m, ctx=self)
            if r is not residuals_out:
                raise ValueError("residuals_out must be contiguous float64")
        else:
            r = np.empty(self.m, dtype=np.float64) if want_residuals else None
            pr = r.ctypes.data if r is not None else None
        rep = LbaReport()
        solve = self._lib.lba_lm_solve12 if w == 12 else self._lib.lba_lm_solve   # 12: [R|c] rows, dense path
        check(solve(self._h, pc, pp, C.byref(o), C.byref(rep), pr))
        n = rep.trace_len
        trace = [dict(it=rep.trace_it[i], alpha=rep.trace_alpha[i], cost=rep.trace_cost[i], cost_new=rep.trace_cost_new[i], lam=rep.trace_lambda[i],
                      pcg=rep.trace_pcg[i], pred=rep.trace_pred[i], accepted=bool(rep.trace_accepted[i]),
                      step_norm=rep.trace_step_norm[i], g_norm=rep.trace_g_norm[i]) for i in range(n)]
        return LMResult(cams, pts, r, rep.iterations, rep.status, STATUS_TEXT.get(rep.status, "?"), rep.n_jac,
                        rep.n_res, rep.pcg_iters_total, rep.chol_clamped, rep.n_accepted, rep.cost_init, rep.cost, rep.seconds, trace)

    # -- multi-GPU --------------------------------------------------------------------------
    def init_comm(self, rank=None, world=None, group=None, peer_bytes=64 << 20):
        """Attach an NCCL communicator (the unique id travels over torch.distributed) and, unless
        ``peer_bytes`` is 0, the NVLink peer-memory all-reduce buffers (CUDA IPC handles travel the same way)."""
        import torch
        import torch.distributed as dist
        rank = dist.get_rank(group) if rank is None else rank
        world = dist.get_world_size(group) if world is None else world
        if world == 1:
            return self
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            check(self._lib.lba_comm_unique_id(buf))
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.cuda(self.device)
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        buf2 = (C.c_ubyte * 128).from_buffer_copy(raw)
        check(self._lib.lba_comm_init(self._h, buf2, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)
        import os as _os
        if peer_bytes and world <= 8 and not _os.environ.get("LBA_NO_PEER"):
            # NVLink peer-memory all-reduce for the small per-camera vectors: exchange CUDA IPC handles
            h = (C.c_ubyte * 64)()
            check(self._lib.lba_comm_peer_local(self._h, h, int(peer_bytes)))
            mine = torch.tensor(list(bytes(h)), dtype=torch.uint8)
            if dist.get_backend(group) == "nccl":
                mine = mine.cuda(self.device)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine, group=group)
            raw = b"".join(bytes(t.cpu().tolist()) for t in allh)
            hb = (C.c_ubyte * (64 * world)).from_buffer_copy(raw)
            check(self._lib.lba_comm_peer_attach(self._h, hb, int(rank), int(world)))
            self.peer_allreduce = True
        return self

    # -- instrumentation ------------------------------------------------------------------
    def synchronize(self):
        check(self._lib.lba_synchronize(self._h))

    def launch_count(self) -> int:
        return int(self._lib.lba_launch_count(self._h))

    def set_profiling(self, enable: bool):
        check(self._lib.lba_set_profiling(self._h, 1 if enable else 0))

    def get_profile(self):
        ms = (C.c_double * _lib.LBA_K_COUNT)()
        n = (C.c_int64 * _lib.LBA_K_COUNT)()
        check(self._lib.lba_get_profile(self._h, ms, n))
        return {self._lib.lba_kernel_class_name(k).decode(): dict(ms=ms[k], launches=int(n[k]))
                for k in range(_lib.LBA_K_COUNT)}

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        check(self._lib.lba_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def measure_copy_bandwidth(self, nbytes=1 << 30) -> float:
        v = C.c_double()
        check(self._lib.lba_measure_copy_bandwidth(self._h, int(nbytes), C.byref(v)))
        return v.value
