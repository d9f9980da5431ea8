#!/usr/bin/env python3
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_multigpu_parity.py

Observations are sharded at 3D-line boundaries, cameras replicated, per-camera partial sums
all-reduced with NCCL inside liblba.so.  Rank 0 also solves the whole scene on its own GPU; the sharded
trajectory must follow it (the all-reduce changes the summation order, so agreement is to rounding).
Exit code 0 on success.

With --stall the script checks the failure path instead: rank 0 starts a solve that no other rank joins; its
device-side waits on the peers are bounded (LBA_SPIN_TIMEOUT_MS), so the call must come back with LBA_ERR_COMM
instead of hanging the GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stall():
    import time
    os.environ["LBA_SPIN_TIMEOUT_MS"] = "300"
    import torch
    import torch.distributed as dist
    from amuya_sfm_lines_beta_b200 import LineBA
    from amuya_sfm_lines_beta_b200._lib import LbaError, LBA_ERR_COMM
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    from amuya_sfm_lines_beta_b200.sharding import plan_shards, local_problem
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = make_synthetic_scene(40, 3000, 8, seed=0)
    sh = plan_shards(s.lines_indices, s.n_lines, world)[rank]
    ci, li, n, p0, p1, pts = local_problem(sh, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1, s.points_3d)
    ba = LineBA(local)
    ba.init_comm(rank, world)
    ba.set_problem(s.n_cameras, sh.n_lines, ci, li, n, p0, p1)
    ok = True
    if rank == 0:
        t0 = time.time()
        try:
            ba.lm_solve(s.camera_params, pts, max_iter=4)
            ok = False
            print("stalled peers: lm_solve returned normally (expected LBA_ERR_COMM)")
        except LbaError as e:
            ok = e.code == LBA_ERR_COMM
            print(f"stalled peers: lm_solve failed after {time.time() - t0:.2f} s with: {e}")
    else:
        time.sleep(3.0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    os._exit(0 if int(flag.item()) else 1)   # the contexts are out of step on purpose: no orderly teardown


def main():
    import torch
    import torch.distributed as dist
    from amuya_sfm_lines_beta_b200 import LineBA
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    from amuya_sfm_lines_beta_b200.sharding import plan_shards, local_problem

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (nc, nl, k, seed) in ((40, 3000, 8, 0), (300, 40000, 10, 1)):
        s = make_synthetic_scene(nc, nl, k, seed=seed)
        sh = plan_shards(s.lines_indices, s.n_lines, world)[rank]
        ci, li, n, p0, p1, pts = local_problem(sh, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1, s.points_3d)
        ba = LineBA(local)
        ba.init_comm(rank, world)
        ba.set_problem(s.n_cameras, sh.n_lines, ci, li, n, p0, p1)
        kw = dict(max_iter=8, ftol=0.0, xtol=0.0, gtol=0.0)
        r_loc, cost0 = ba.residuals(s.camera_params, pts)
        res = ba.lm_solve(s.camera_params, pts, **kw)
        # gather the line parameters and residuals of every shard on rank 0
        gathered = [None] * world
        dist.all_gather_object(gathered, dict(l0=sh.line_start, l1=sh.line_stop, pts=res.points_3d, idx=sh.obs_index,
                                              r=res.residuals, cams=res.camera_params, cost=res.cost,
                                              trace=[t["cost_new"] for t in res.trace]))
        if rank == 0:
            full = LineBA(local)
            full.set_problem(s.n_cameras, s.n_lines, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1)
            _, cost0_full = full.residuals(s.camera_params, s.points_3d)
            ref = full.lm_solve(s.camera_params, s.points_3d, **kw)
            pts_all = np.empty_like(ref.points_3d); r_all = np.empty_like(ref.residuals)
            for g in gathered:
                pts_all[2 * g["l0"]:2 * g["l1"]] = g["pts"]
                r_all[g["idx"]] = g["r"]
                assert np.array_equal(g["cams"], gathered[0]["cams"]), "camera parameters differ between ranks"
                assert g["cost"] == gathered[0]["cost"], "cost differs between ranks"
            tr = np.array(gathered[0]["trace"]); tr_ref = np.array([t["cost_new"] for t in ref.trace])
            e_tr = np.max(np.abs(tr - tr_ref) / tr_ref)
            e_c = abs(res.cost - ref.cost) / ref.cost
            e_p = np.max(np.abs(pts_all - ref.points_3d)) / np.max(np.abs(ref.points_3d))
            e_cam = np.max(np.abs(res.camera_params - ref.camera_params)) / np.max(np.abs(ref.camera_params))
            e_r = np.max(np.abs(r_all - ref.residuals)) / np.max(np.abs(ref.residuals))
            print(f"[world {world}] scene {nc}c/{nl}l/{s.n_obs}o: cost0 rel {abs(cost0 - cost0_full) / cost0_full:.1e} "
                  f"trajectory {e_tr:.1e} final cost {e_c:.1e} cams {e_cam:.1e} pts {e_p:.1e} residuals {e_r:.1e}")
            ok = ok and e_tr < 1e-9 and e_c < 1e-6 and e_p < 1e-6 and e_cam < 1e-6 and abs(cost0 - cost0_full) < 1e-12 * cost0_full
            full.close()
        ba.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    stall() if "--stall" in sys.argv else main()
