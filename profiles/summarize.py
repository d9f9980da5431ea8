#!/usr/bin/env python3
"""Turn the raw ncu outputs in gpurun_out/ into the small, committed summaries in profiles/.

    python profiles/summarize.py r1

reads  gpurun_out/launches_<round>.csv   (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/prof_<round>.ncu-rep   (ncu --set full capture of the main kernels)
writes profiles/launches_<round>.md, profiles/kernels_<round>.json, profiles/kernels_<round>.md
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
go = os.path.join(ROOT, "gpurun_out")
out = os.path.join(ROOT, "profiles")

# ---- launch list
rows = list(csv.reader(open(os.path.join(go, f"launches_{rnd}.csv"))))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("lba::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    agg.setdefault(name, []).append(v)
# The PCG kernels are enqueued one iteration ahead of the host's knowledge and return at once when the device-side
# `done` flag is set: those early-exit launches (a few us) are listed apart, not averaged into the kernel time.
EARLY_EXIT = ("k_line_pass<0>", "k_pcg_cam", "k_cg_step")
stats = collections.OrderedDict()
for k, v in agg.items():
    real = [x for x in v if not (k.startswith(EARLY_EXIT) and x < 0.5 * max(v))]
    early = [x for x in v if k.startswith(EARLY_EXIT) and x < 0.5 * max(v)]
    stats[k] = dict(n=len(real), tot=sum(real), n_early=len(early), tot_early=sum(early))
bench_only = {k: a for k, a in stats.items() if not k.startswith(("k_fp64_peak", "k_stream_copy"))}
tot = sum(a["tot"] + a["tot_early"] for a in bench_only.values())
with open(os.path.join(out, f"launches_{rnd}.md"), "w") as f:
    f.write(f"# ncu launch list, round {rnd}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
            "`python bench.py --steps 6 --warmup 3 --no-cpu-baseline` (config 4: 1k cameras / 200k lines / 2M observations).\n"
            "Per-launch times are cold-cache and serialised: compare SHARES.  The two microbenchmark kernels "
            "(`k_fp64_peak`, `k_stream_copy`) are excluded from the shares.  The three PCG kernels are enqueued ahead of "
            "the host's knowledge of convergence and return at once when the device-side `done` flag is set; those "
            "early-exit launches are counted in their own columns and are not part of `avg us`.\n\n"
            "| kernel | launches | total us | avg us | share | early-exit launches | their avg us |\n|---|---:|---:|---:|---:|---:|---:|\n")
    for k, a in sorted(bench_only.items(), key=lambda x: -(x[1]["tot"] + x[1]["tot_early"])):
        ea = f"{a['tot_early'] / a['n_early']:.2f}" if a["n_early"] else ""
        f.write(f"| `{k}` | {a['n']} | {a['tot']:.1f} | {a['tot'] / max(a['n'], 1):.2f} | {(a['tot'] + a['tot_early']) / tot:.3f} | "
                f"{a['n_early'] or ''} | {ea} |\n")
    f.write(f"\nTotal {tot:.0f} us over the listed launches.\n")

# ---- full captures
WANT = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "regs",
        "launch__grid_size": "grid", "launch__block_size": "block", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct"}


def summarize_full(reps, stem, title):
    """One row per kernel (its longest captured launch) from one or more `ncu --set full` reports."""
    best = {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        H, U = rows[0], rows[1]
        ki = H.index("Kernel Name")
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void ", "").replace("lba::", "")
            d = {}
            for m, short in WANT.items():
                if m in H:
                    val, unit = r[H.index(m)].replace(",", ""), U[H.index(m)]
                    try:
                        x = float(val)
                    except ValueError:
                        continue
                    if short == "time_us":
                        x = {"ns": x / 1e3, "nsecond": x / 1e3, "ms": x * 1e3, "msecond": x * 1e3, "s": x * 1e6, "second": x * 1e6}.get(unit, x)
                    if short.endswith("_MB"):
                        x = {"Kbyte": x / 1e3, "Mbyte": x, "Gbyte": x * 1e3, "byte": x / 1e6}.get(unit, x)
                    d[short] = round(x, 3)
            st = [(H[i].replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i].replace(",", "") or 0))
                  for i, h in enumerate(H) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i]]
            tot_s = sum(v for _, v in st) or 1
            d["top_stalls"] = {k: round(v / tot_s, 3) for k, v in sorted(st, key=lambda x: -x[1])[:5]}
            if name not in best or d.get("time_us", 0) > best[name].get("time_us", 0):   # keep the real (not early-exit) launch
                best[name] = d
    json.dump(best, open(os.path.join(out, f"{stem}.json"), "w"), indent=1)
    with open(os.path.join(out, f"{stem}.md"), "w") as f:
        f.write(f"# {title}\n\n"
                "| kernel | time us | DRAM rd MB | DRAM wr MB | DRAM % | L1TEX % | LTS % | FP64 pipe % | issue % | warps act % | regs | top stalls |\n"
                "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for k, d in sorted(best.items(), key=lambda x: -x[1].get("time_us", 0)):
            f.write(f"| `{k}` | {d.get('time_us')} | {d.get('dram_read_MB')} | {d.get('dram_write_MB')} | {d.get('dram_pct')} | "
                    f"{d.get('l1tex_pct')} | {d.get('lts_pct')} | {d.get('fp64_pipe_pct')} | {d.get('issue_pct')} | {d.get('warps_active_pct')} | "
                    f"{d.get('regs')} | {', '.join(f'{a} {b:.0%}' for a, b in d['top_stalls'].items())} |\n")
    return os.path.join(out, f"{stem}.md")


main_md = summarize_full([os.path.join(go, f"prof_{rnd}.ncu-rep")], f"kernels_{rnd}",
                         f"ncu --set full, round {rnd} (one launch per kernel, config 4)")
extra = [os.path.join(go, n) for n in ("prof_n2.ncu-rep", "prof_n3.ncu-rep") if os.path.exists(os.path.join(go, n))]
next_md = summarize_full(extra, f"kernels_next_{rnd}",
                         f"ncu --set full, round {rnd}: kernels of the \"next\" rows (N2 track table: tests/measure_track_table.py, "
                         "8 views x 250 k lines; N3 triangulation: tests/measure_triangulate.py, 2 M tracks)") if extra else None
print(open(os.path.join(out, f"launches_{rnd}.md")).read())
print(open(main_md).read())
if next_md:
    print(open(next_md).read())
