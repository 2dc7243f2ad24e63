#!/usr/bin/env python
"""Schedule diagnostic: main-stream time stamps (head / wait for the BIOGEM block / tracer step) of a few graph-replayed
ocean cycles of cg_run, printed by the library when CG_TRACE=<cycles> is set."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgenie_b200 import Ensemble, materialise  # noqa: E402
import numpy as np
d = tempfile.mkdtemp()
materialise(d, "eb_go_gs_ac_bg_36x36x16")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rng = np.random.default_rng(20261017)
e = Ensemble(d, n_members=M, perturb={"adrag": rng.uniform(2.0, 3.0, M), "diff1": rng.uniform(1600.0, 2500.0, M)})
e.set_tracer_variant("col")
e.run(480 * 4)
e.synchronize()
os.environ["CG_TRACE"] = "12"
e.run(5 * 16)
e.synchronize()
