#!/usr/bin/env python3
"""P5 of the parity ladder: for every fixture record what the reference's scipy path reaches,
(a) as shipped (2-point finite differences with the shipped sparsity pattern, ftol=1e-4) and (b) with the
analytic Jacobian in the correct columns at a fixed budget of Jacobian evaluations.  Uses the vectorised
port of the reference call (oracle.scipy_reference_port; its as-shipped behaviour is checked against the
recorded run of the reference itself in tests/test_oracle_golden.py).  Writes tests/golden/p5_costs.json.

    python tests/golden/make_p5_costs.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import line_ba_oracle as orc  # noqa: E402
from amuya_sfm_lines_beta_b200.scenes import scene_from_npz, make_synthetic_scene  # noqa: E402

NJEV = 15
out = {"njev_budget": NJEV, "fixtures": {}}
scenes = {n: scene_from_npz(os.path.join(HERE, n + ".npz")) for n in ("toy1_lv2", "toy1_lv5", "toy1_lv9", "toy2_lv2", "manual3")}
scenes["synth_12c_200l"] = make_synthetic_scene(12, 200, 5, seed=1)
for name, s in scenes.items():
    r0 = orc.residuals(*s.arrays())
    a = orc.scipy_reference_port(*s.arrays(), mode="as_shipped")
    b = orc.scipy_reference_port(*s.arrays(), mode="analytic", max_nfev=NJEV)
    out["fixtures"][name] = {"cost_init": 0.5 * float(r0 @ r0), "as_shipped_cost": float(a[3].cost), "as_shipped_nfev": int(a[3].nfev),
                             "as_shipped_njev": int(a[3].njev), "analytic_cost": float(b[3].cost), "analytic_njev": int(b[3].njev)}
    print(name, out["fixtures"][name])
json.dump(out, open(os.path.join(HERE, "p5_costs.json"), "w"), indent=1)
