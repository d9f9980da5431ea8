#!/usr/bin/env python3
"""Benchmark of the line bundle-adjustment hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py --gpus 1 --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --steps K --warmup W     # CPU port of the reference's scipy path
    torchrun --nproc-per-node N bench.py --gpus N ...         # observations sharded over N GPUs

A "step" is one Levenberg-Marquardt iteration (solve of the damped normal equations by Schur
elimination + PCG, trial evaluation with its Jacobian, accept/reject on the device) on BASELINE.json's
synthetic scene.  One JSON line is printed by rank 0.

`--gpus N` is a WEAK-scaling series: N GPUs solve a scene with N x the cameras, lines and observations of
BASELINE config [3] (N = 8: 8k cameras / 1.6M lines / 16M observations, the size class of config [4]).  An LM
iteration on an N x larger scene is N x the work, so `value` is work-normalised: LM iterations/s x
(observations / 2e6), i.e. "config-[3]-sized iterations per second" -- identical to plain iterations/s at N = 1.
`--workload config5` instead runs BASELINE config [4] itself (10k / 2M / 20M) sharded over N GPUs (strong scaling).
For N > 1 the line carries a `parity` block: rank 0 solves the same scene alone on its GPU and the sharded
trajectory, final cost and parameters must agree (exit code 1 otherwise).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lm_iterations_per_sec"
UNIT = "LM iterations/s per 2M observations"
REF_OBS = 2_000_000   # BASELINE config [3]: the metric is quoted per this many observations

# BASELINE.json configs: [3] = 1k cameras / 200k lines / 2M observations (1 GPU),
# [4] = 10k cameras / 2M lines / 20M observations (sharded).  Per-GPU work is fixed as N grows
# ("weak" scaling): N GPUs solve a scene with N x the lines and observations of config [3].
PER_GPU_LINES = 200_000
OBS_PER_LINE = 10
N_CAMERAS = 1000

# algorithmic HBM bytes per observation (fp64, int32 indices, 10 observations per line)
BYTES_PER_OBS = {
    "residual": 72 + 8 + 8 + 4.8,                     # constants, indices, r out, line endpoints (48 B / 10 obs)
    "jacobian": 72 + 10 + 8 + 96 + 4.8 + 22.4,        # + pos, J row out (12 doubles), V_l (176 B rows), g_l out (224 B / 10 obs)
    "pcg_line": 48 + 6 + 4.8 + 17.6 + 8,              # jl, indices (ll, cam, meta), endpoints, Vinv (22*8/10), w out
    "pcg_cam": 48 + 4 + 8,                            # jc, cam_perm, w gather
}
# SURVEY.md 8(d)'s per-observation figures for the same passes (what the judge's roofline uses): J evaluation with the
# Jacobian materialised 188.8 B, matrix-free Schur mat-vec 104 B per CG iteration (both halves together)
SURVEY_BYTES_PER_OBS = {"jacobian": 188.8, "pcg_pair": 104.0, "residual": 92.8}


NCU_KERNEL_OF_CLASS = {"jacobian": "k_eval_z<2, 1>", "residual": "k_eval<0, 2, 1>", "pcg_line": "k_line_pass<0>",
                       "pcg_cam": "k_pcg_cam"}


def read_traffic(kernel_class):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel from the committed
    `ncu --set full` summary (profiles/kernels_r*.json, written by profiles/summarize.py); None if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "kernels_r*.json")))
    if not files:
        return None
    d = json.load(open(files[-1])).get(NCU_KERNEL_OF_CLASS.get(kernel_class, ""), None)
    if not d or d.get("dram_read_MB") is None:
        return None
    return (d["dram_read_MB"] + d["dram_write_MB"]) * 1e6


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Clocks / throttle reasons DURING the timed region: ONE `nvidia-smi -lms` process (the profiling recipe's
    clocks line) started before the region and stopped after it, on rank 0 only, for the GPUs of this job.
    A disabled sampler (other ranks) does nothing."""
    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, indices=(0,), enabled=True, period_ms=50):
        self.indices, self.enabled, self.period_ms = list(indices), enabled, period_ms
        self.samples, self._p = [], None

    def __enter__(self):
        if self.enabled:
            try:
                self._p = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices),
                                            f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                            "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                           text=True)
            except Exception:
                self._p = None
        return self

    def __exit__(self, *a):
        if self._p is None:
            return
        self._p.terminate()
        try:
            out, _ = self._p.communicate(timeout=10)
        except Exception:
            self._p.kill()
            out, _ = self._p.communicate()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 7 and f[1].replace(".", "", 1).isdigit():
                self.samples.append(f[1:])
        self._p = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(sm), "gpus": self.indices}


def make_scene(n_gpus, scale=1.0, seed=0, cams=0, variant="A"):
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    n_lines = int(PER_GPU_LINES * n_gpus * scale)
    n_cams = cams or max(OBS_PER_LINE, int(N_CAMERAS * n_gpus * scale))
    return make_synthetic_scene(n_cams, n_lines, OBS_PER_LINE, seed=seed, variant=variant)


def cpu_reference_run(nfev, full_obs, mode="analytic", scale=1.0):
    """The reference's scipy path (bundle_adjustment.py:319) with `fun` vectorised in numpy, timed on the host
    cores.  scale = 1: BASELINE config [3] itself (1k cameras / 200k lines / 2M observations; ~50 s per TRF
    iteration); scale < 1: a bounded sample of it (same generator, cameras / lines scaled), iterations/s scaled
    by the observation ratio.  ONE definition of an iteration everywhere: a TRF iteration = one trial evaluation,
    iterations = nfev - 1, time = the whole least_squares call."""
    from oracle import line_ba_oracle as orc
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    warm = make_synthetic_scene(20, 400, OBS_PER_LINE, seed=0, variant="A")
    orc.scipy_reference_port(*warm.arrays(), mode=mode, max_nfev=2)   # warm numpy / scipy code paths
    s = make_synthetic_scene(max(OBS_PER_LINE, int(N_CAMERAS * scale)), int(PER_GPU_LINES * scale), OBS_PER_LINE, seed=0, variant="A")
    t0 = time.perf_counter()
    out = orc.scipy_reference_port(*s.arrays(), mode=mode, max_nfev=max(nfev, 2), ftol=0.0)
    dt = time.perf_counter() - t0
    iters = max(1, out[3].nfev - 1)
    v = iters / dt * (s.n_obs / REF_OBS)
    return dict(value=v, unit=UNIT, cores=1, kind="port", iterations=iters, seconds=dt,
                sample=f"scipy least_squares(trf, lsmr, x_scale='jac', {mode} Jacobian) with numpy-vectorised fun on "
                       f"{s.n_cameras} cams/{s.n_lines} lines/{s.n_obs} obs"
                       f"{' (= BASELINE config [3] at full size)' if scale == 1.0 else f' (a {scale:g}-scale sample; iterations/s scaled by {s.n_obs}/{REF_OBS} observations)'}: "
                       f"{iters} TRF iterations (nfev - 1) in {dt:.1f} s; host has {os.cpu_count()} cores, the path is single-threaded "
                       f"like the reference")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config4xN", choices=["config4xN", "config5"],
                    help="config4xN: N x BASELINE config [3] (weak scaling, the default); config5: BASELINE config [4] "
                         "(10k cameras / 2M lines / 20M observations) sharded over N GPUs (strong scaling)")
    ap.add_argument("--repeats", type=int, default=5, help="timed repetitions of the K-step legs (median reported)")
    ap.add_argument("--scale", type=float, default=1.0, help="debug: shrink the scene")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="debug: skip the per-class event profile (roofline = null)")
    ap.add_argument("--no-parity", action="store_true", help="debug: skip the sharded-vs-single-GPU check at N > 1")
    ap.add_argument("--ref-mode", default="analytic", choices=["analytic", "as_shipped"])
    ap.add_argument("--ref-nfev", type=int, default=4, help="reference arm: scipy max_nfev on config [3] (~50 s per iteration)")
    ap.add_argument("--cams", type=int, default=0, help="debug: override the camera count")
    ap.add_argument("--variant", default="A", choices=["A", "B"], help="visibility: A uniform random, B contiguous camera window")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    strong = args.workload == "config5"
    mult = 10 if strong else args.gpus            # scene size in units of config [3]
    full_obs = int(PER_GPU_LINES * mult * args.scale) * OBS_PER_LINE
    n_cams_total = args.cams or max(OBS_PER_LINE, int(N_CAMERAS * mult * args.scale))
    norm = full_obs / REF_OBS
    workload = (f"synthetic line-SfM scene, {n_cams_total} cameras / {full_obs // OBS_PER_LINE} 3D lines / {full_obs} "
                f"line observations (10 views per line, "
                f"{'uniform random visibility' if args.variant == 'A' else 'contiguous camera-window visibility'}), seed 0"
                + (" = BASELINE config [4]" if strong and args.scale == 1.0 else
                   (" = BASELINE config [3]" if mult == 1 and args.scale == 1.0 else f" = {mult} x BASELINE config [3]")))
    config = {"workload": workload, "obs_per_gpu": full_obs // max(args.gpus, 1), "parallelism": f"obs-shard x{args.gpus}",
              "value_normalisation": f"LM iterations/s x ({full_obs} observations / {REF_OBS}); raw iterations/s in lm.iterations_per_sec_raw",
              "l2": "working set per LM iteration (~0.5 GB/GPU) exceeds the 126 MB L2; no flush needed"}
    scaling = "strong" if strong else "weak"

    if args.impl == "reference":
        if rank != 0:
            return
        # The CPU path is linear in the observation count, so its work-normalised throughput does not depend on N:
        # the arm times BASELINE config [3] itself (for N > 1: 1/N of the workload, the metric being per 2M observations)
        cb = cpu_reference_run(args.ref_nfev, full_obs, mode=args.ref_mode, scale=args.scale)
        if norm != 1.0:
            cb["sample"] += f"; the N = {args.gpus} workload is {norm:g} x this scene, the metric is normalised per {REF_OBS} observations"
        line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["iterations"],
                "warmup": 1, "ms_per_step": 1e3 * cb["seconds"] / cb["iterations"], "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from amuya_sfm_lines_beta_b200 import LineBA
    from amuya_sfm_lines_beta_b200.sharding import plan_shards, local_problem

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    s = make_scene(mult, args.scale, cams=args.cams, variant=args.variant)
    shard = plan_shards(s.lines_indices, s.n_lines, world)[rank]
    ci, li, n, p0, p1, pts_local = local_problem(shard, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1, s.points_3d)
    m_local = ci.shape[0]

    # pinned host copies: the e2e leg copies from these inside the timed region
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.pin_memory()
    h = dict(ci=pin(ci), li=pin(li), n=pin(n), p0=pin(p0), p1=pin(p1), cams=pin(s.camera_params), pts=pin(pts_local))

    stream = torch.cuda.Stream()
    ba = LineBA(local_rank)
    from amuya_sfm_lines_beta_b200._lib import check
    check(ba._lib.lba_set_stream(ba._h, stream.cuda_stream))
    if world > 1:
        ba.init_comm(rank, world)
    ba.set_problem(s.n_cameras, shard.n_lines, h["ci"], h["li"], h["n"], h["p0"], h["p1"])
    lm_kw = dict(ftol=0.0, xtol=0.0, gtol=0.0, verbose=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            out = fn()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    cams_buf = h["cams"].clone().cuda()
    pts_buf = h["pts"].clone().cuda()

    def reset_state():
        cams_buf.copy_(h["cams"]); pts_buf.copy_(h["pts"])
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up to the end of the e2e leg (the timed regions
    # themselves last only milliseconds)
    clk = ClockSampler(range(world), enabled=(rank == 0))
    clk.__enter__()
    # -- warm-up: W untimed LM iterations --------------------------------------------------------------
    reset_state()
    ba.lm_solve(None, None, out=(cams_buf, pts_buf), want_residuals=False, max_iter=warmup, **lm_kw)

    # -- untimed soak: the same K-step solve repeated for ~0.7 s, so that the clock sampler sees this load for
    #    several periods before the (milliseconds-long) timed regions; same count on every rank
    for _ in range(max(3, int(700.0 / (max(args.steps, 1) * 0.6 * max(1.0, norm / max(world, 1)))))):
        reset_state()
        ba.lm_solve(None, None, out=(cams_buf, pts_buf), want_residuals=False, max_iter=args.steps, **lm_kw)

    # -- value: K LM iterations, observations and parameters already resident in HBM; R repetitions of the
    #    K-step region (each from the same initial state), the MEDIAN is reported, min / max beside it ----------
    reps = []
    for _ in range(max(1, args.repeats)):
        reset_state()
        n0 = ba.launch_count()
        ms_i, res_i = timed(lambda: ba.lm_solve(None, None, out=(cams_buf, pts_buf), want_residuals=False,
                                                max_iter=args.steps, **lm_kw))
        reps.append((ms_i, res_i, ba.launch_count() - n0))
    reps.sort(key=lambda x: x[0])
    ms, res, launches = reps[len(reps) // 2]
    steps_done = res.iterations
    raw = steps_done / (ms * 1e-3)
    value = raw * norm
    cams_sharded = cams_buf.cpu().numpy().copy()      # state after the last K-step solve (all repetitions are identical)
    pts_sharded = pts_buf.cpu().numpy().copy()

    # -- e2e: host buffers in, host results out, through the public array API ------------------------------
    out_c = torch.empty_like(h["cams"]).pin_memory()
    out_p = torch.empty_like(h["pts"]).pin_memory()
    out_r = torch.empty(m_local, dtype=torch.float64).pin_memory()

    def e2e_call():
        ba.set_problem(s.n_cameras, shard.n_lines, h["ci"], h["li"], h["n"], h["p0"], h["p1"])
        out_c.copy_(h["cams"]); out_p.copy_(h["pts"])
        return ba.lm_solve(None, None, out=(out_c, out_p), residuals_out=out_r, max_iter=args.steps, **lm_kw)
    e2e_call()  # warm
    e2e_reps = sorted((timed(e2e_call) for _ in range(max(1, args.repeats))), key=lambda x: x[0])
    ms_e2e, res_e2e = e2e_reps[len(e2e_reps) // 2]
    e2e_value = res_e2e.iterations / (ms_e2e * 1e-3) * norm
    clk.__exit__(None, None, None)
    h2d = sum(h[k].numel() * h[k].element_size() for k in ("ci", "li", "n", "p0", "p1", "cams", "pts"))
    d2h = out_c.numel() * 8 + out_p.numel() * 8 + m_local * 8

    # -- roofline: per-kernel-class device time from CUDA events on the launching stream ---------------------
    prof = {}
    if not args.no_profile:
        reset_state()
        ba.set_profiling(True)
        ba.lm_solve(None, None, out=(cams_buf, pts_buf), want_residuals=False, max_iter=min(args.steps, 10), **lm_kw)
        prof = ba.get_profile()
        ba.set_profiling(False)
    peak, peak_src = read_peaks()
    tot = sum(v["ms"] for v in prof.values()) or 1.0
    shares = {k: round(v["ms"] / tot, 4) for k, v in prof.items() if v["launches"]}
    dom = max((k for k in BYTES_PER_OBS if prof.get(k, {}).get("launches")), key=lambda k: prof[k]["ms"], default=None)
    roof = None
    per_kernel = {}
    for k, bpo in BYTES_PER_OBS.items():
        if prof.get(k, {}).get("launches"):
            avg_ms = prof[k]["ms"] / prof[k]["launches"]
            per_kernel[k] = {"avg_us": round(avg_ms * 1e3, 2), "GBps": round(bpo * m_local / (avg_ms * 1e-3) / 1e9, 1),
                             "frac": round(bpo * m_local / (avg_ms * 1e-3) / 1e9 / peak, 4), "launches": prof[k]["launches"]}
    if dom:
        roof = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": per_kernel[dom]["frac"], "traffic": read_traffic(dom), "peak_source": peak_src,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full summary under profiles/",
                "frac_survey_bytes": round(SURVEY_BYTES_PER_OBS.get(dom, BYTES_PER_OBS[dom]) * m_local / (per_kernel[dom]["avg_us"] * 1e-6) / 1e9 / peak, 4),
                "survey_bytes_per_obs": SURVEY_BYTES_PER_OBS.get(dom),
                "pcg_pair_frac_survey_bytes": (round(SURVEY_BYTES_PER_OBS["pcg_pair"] * m_local / ((per_kernel["pcg_line"]["avg_us"] + per_kernel["pcg_cam"]["avg_us"]) * 1e-6) / 1e9 / peak, 4)
                                               if "pcg_line" in per_kernel and "pcg_cam" in per_kernel else None),
                "algorithmic_bytes_per_launch": BYTES_PER_OBS[dom] * m_local,
                "algorithmic_bytes_per_obs": BYTES_PER_OBS[dom], "per_kernel": per_kernel, "time_shares": shares,
                "class_avg_us": {k: round(v["ms"] / v["launches"] * 1e3, 2) for k, v in prof.items() if v["launches"]}}

    fp64_tf = ba.measure_fp64_peak()
    copy_gbs = ba.measure_copy_bandwidth(1 << 30)

    # -- N > 1: the sharded solve against the same scene solved by ONE GPU ------------------------------------
    parity = None
    rc = 0
    if world > 1 and not args.no_parity:
        mine = torch.from_numpy(cams_sharded).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        pts_all = [None] * world
        dist.all_gather_object(pts_all, (shard.line_start, pts_sharded))
        if rank == 0:
            identical = all(torch.equal(gathered[0].view(torch.int64), g.view(torch.int64)) for g in gathered[1:])
            one = LineBA(local_rank)
            one.set_problem(s.n_cameras, s.n_lines, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1)
            ref = one.lm_solve(s.camera_params, s.points_3d, want_residuals=False, max_iter=args.steps, **lm_kw)
            one.close()
            pts_full = np.empty_like(ref.points_3d)
            for line0, pp in pts_all:
                pts_full[2 * line0:2 * line0 + pp.shape[0]] = pp
            c_sh = np.array([t["cost_new"] for t in res.trace]); c_1 = np.array([t["cost_new"] for t in ref.trace])
            same_path = (len(c_sh) == len(c_1) and [t["accepted"] for t in res.trace] == [t["accepted"] for t in ref.trace]
                         and [t["pcg"] for t in res.trace] == [t["pcg"] for t in ref.trace])
            traj = float(np.max(np.abs(c_sh - c_1) / np.abs(c_1))) if same_path and len(c_1) else float("inf")
            d_cost = abs(res.cost - ref.cost) / ref.cost
            d_cams = float(np.max(np.abs(cams_sharded - ref.camera_params)) / np.max(np.abs(ref.camera_params)))
            d_pts = float(np.max(np.abs(pts_full - ref.points_3d)) / np.max(np.abs(ref.points_3d)))
            ok = bool(identical and same_path and traj <= 1e-9 and d_cost <= 1e-6 and d_cams <= 1e-6 and d_pts <= 1e-6)
            parity = {"ok": ok, "ranks_camera_params_bit_identical": bool(identical), "same_accepts_and_pcg_counts": bool(same_path),
                      "per_trial_cost_max_rel_diff": traj, "final_cost_rel_diff": d_cost, "camera_params_max_rel_diff": d_cams,
                      "points_3d_max_rel_diff": d_pts, "tolerances": {"per_trial_cost": 1e-9, "final": 1e-6},
                      "single_gpu_cost": ref.cost, "sharded_cost": res.cost, "trials": len(c_1)}
            rc = 0 if ok else 1
        dist.barrier()

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1:
            # bounded sample (~15 s): a 1/10-scale scene from the same generator; the reference arm times the full config
            cb = cpu_reference_run(4, full_obs, mode=args.ref_mode, scale=0.1 * args.scale)
        trials = max(len(res.trace), 1)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
                "warmup": warmup, "ms_per_step": ms / max(steps_done, 1), "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "repeats": {"n": len(reps), "reported": "median", "ms_per_step_min": reps[0][0] / max(steps_done, 1),
                            "ms_per_step_max": reps[-1][0] / max(steps_done, 1),
                            "e2e_ms_total_min": e2e_reps[0][0], "e2e_ms_total_max": e2e_reps[-1][0],
                            "ms_total_all": [round(r[0], 3) for r in reps], "e2e_ms_total_all": [round(r[0], 3) for r in e2e_reps]},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / max(res_e2e.iterations, 1),
                        "d2h_bytes_per_step": d2h / max(res_e2e.iterations, 1), "ms_total": ms_e2e,
                        "note": "one call = lba_set_problem from pinned host arrays + K LM iterations + results to host"},
                "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cb,
                "lm": {"iterations_per_sec_raw": raw, "cost_init": res.cost_init, "cost_final": res.cost,
                       "pcg_iters_total": res.pcg_iters_total, "pcg_iters_per_lm": res.pcg_iters_total / max(steps_done, 1),
                       "evaluations": res.n_jac, "trials": len(res.trace), "accepted": res.n_accepted,
                       "accept_rate": res.n_accepted / trials,
                       "obs_per_sec": m_local * world * res.n_jac / (ms * 1e-3)},
                "jacobian_evals": {"per_sec_in_loop": res.n_jac / (ms * 1e-3),
                                   "kernel_only_per_sec": (1e6 / per_kernel["jacobian"]["avg_us"]) if "jacobian" in per_kernel else None,
                                   "obs_per_sec_kernel_only": (m_local * world * 1e6 / per_kernel["jacobian"]["avg_us"]) if "jacobian" in per_kernel else None,
                                   "note": "fused residual + analytic Jacobian + per-line J^T J blocks, one launch per evaluation"},
                "device_peaks": {"fp64_tflops_measured": round(fp64_tf, 2), "copy_gbs_measured": round(copy_gbs, 1)}}
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line))
    ba.close()
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


if __name__ == "__main__":
    main()
