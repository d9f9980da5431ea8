#!/usr/bin/env python3
"""Records the numpy oracle's LM trajectory on BASELINE.json config [3] at full size (1k cameras / 200k lines /
2M observations, seed 0, uniform visibility) for tests/test_gpu_parity.py::test_config4_trajectory_matches_recorded_oracle.
The oracle needs about ten seconds per evaluation at this size, so the GPU test compares against this record
instead of running it.

    python tests/golden/make_config4_trajectory.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import line_ba_oracle as orc  # noqa: E402
from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene  # noqa: E402

opts = dict(max_iter=8, ftol=0.0, xtol=0.0, gtol=0.0)
s = make_synthetic_scene(1000, 200_000, 10, seed=0, variant="A")
cams, pts, r, rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(**opts), verbose=True)
out = dict(options=opts, cost_init=rep.cost_init, cost=rep.cost, iterations=rep.iterations, n_accepted=rep.n_accepted,
           it=[t["it"] for t in rep.trace], accepted=[bool(t["accepted"]) for t in rep.trace],
           pcg=[int(t["pcg"]) for t in rep.trace], alpha=[t["alpha"] for t in rep.trace],
           cost_new=[t["cost_new"] for t in rep.trace], lam=[t["lam"] for t in rep.trace],
           cams_norm=float(np.linalg.norm(cams)), pts_norm=float(np.linalg.norm(pts)))
json.dump(out, open(os.path.join(HERE, "config4_trajectory.json"), "w"), indent=1)
print(out)
