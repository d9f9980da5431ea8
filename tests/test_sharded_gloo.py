"""CPU, world_size 2 over gloo: the N>1 host path.  Each rank evaluates its shard with the oracle,
per-camera partial sums are all-reduced exactly where the GPU path all-reduces them, and the result
must equal the unsharded evaluation."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import line_ba_oracle as orc
        from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
        from amuya_sfm_lines_beta_b200.sharding import plan_shards, local_problem
        s = make_synthetic_scene(9, 120, 4, seed=4)
        sh = plan_shards(s.lines_indices, s.n_lines, world)[rank]
        ci, li, n, p0, p1, pts = local_problem(sh, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1, s.points_3d)
        om, ce = orc.split_camera_params(s.camera_params)
        r, jc, jl = orc.residuals_jacobian(om, ce, pts.reshape(-1, 6), ci, li, n, np.cross(n, p0), np.cross(n, p1))
        U, gc, V, gl = orc.normal_equations(r, jc, jl, ci, li, s.n_cameras, sh.n_lines)
        buf = torch.from_numpy(np.concatenate([U.ravel(), gc.ravel(), [0.5 * float(r @ r)]]))
        dist.all_reduce(buf)                         # the per-camera J^T J / J^T r all-reduce (+ cost)
        ret[rank] = dict(buf=buf.numpy().copy(), V=V, gl=gl, l0=sh.line_start, l1=sh.line_stop,
                         r=r, idx=sh.obs_index)
    finally:
        dist.destroy_process_group()


def test_world2_sharded_normal_equations_equal_unsharded():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    from oracle import line_ba_oracle as orc
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    s = make_synthetic_scene(9, 120, 4, seed=4)
    om, ce = orc.split_camera_params(s.camera_params)
    r, jc, jl = orc.residuals_jacobian(om, ce, s.points_3d.reshape(-1, 6), s.camera_indices, s.lines_indices,
                                       s.n, np.cross(s.n, s.p0), np.cross(s.n, s.p1))
    U, gc, V, gl = orc.normal_equations(r, jc, jl, s.camera_indices, s.lines_indices, s.n_cameras, s.n_lines)
    full = np.concatenate([U.ravel(), gc.ravel(), [0.5 * float(r @ r)]])
    for rank in range(world):
        out = ret[rank]
        assert np.allclose(out["buf"], full, rtol=1e-12, atol=1e-18)
        assert np.array_equal(out["V"], V[out["l0"]:out["l1"]])      # line blocks are entirely local
        assert np.array_equal(out["gl"], gl[out["l0"]:out["l1"]])
        assert np.array_equal(out["r"], r[out["idx"]])
    assert np.array_equal(ret[0]["buf"], ret[1]["buf"])              # every rank sees identical bits
