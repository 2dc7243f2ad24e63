#!/usr/bin/env python
"""Small driver for compute-sanitizer over the kernels smoke() does not reach: the mixed-layer scheme (k_mld_*) and the extended
time-series integrals (k_bg_settle_sur, k_bg_sig2_*).  A few ocean steps / two BIOGEM blocks, two members."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgenie_b200 import Ensemble, materialise  # noqa: E402

d = tempfile.mkdtemp()
materialise(d, "eb_go_gs_36x36x8", overrides={"go_imld": 1})
with Ensemble(d, n_members=2, perturb={"scf": np.array([2.0, 1.7])}) as e:
    e.set_tracer_variant("strict")
    e.run(20)
    e.set_tracer_variant("col")
    e.run(10)
    assert int(e.health().sum()) == 0
    print("mld", float(e.get("mld", 1).min()))
d = tempfile.mkdtemp()
materialise(d, "eb_go_gs_ac_bg_36x36x16")
with Ensemble(d, n_members=2) as e:
    genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
    tick = int(round(1000.0 * genie_timestep))
    dts = float(2 * 5) * genie_timestep
    e.biogem_sig_extended()
    e.biogem_sig_reset()
    for k in range(1, 31):
        if k % 5 == 1:
            e.surflux()
        e.step_embm()
        if k % 5 == 0:
            e.step_seaice()
            e.step_goldstein()
        if k % 10 == 0:
            if k == 10:
                e.biogem_climate_sol()
            e.biogem_forcing(k * tick)
            e.biogem_step(dts, k * tick)
            e.biogem_tracercoupling()
            e.biogem_climate()
            e.biogem_sig_update(dts, 1000.0)
            e.atchem_step(dts)
    s2 = e.get("bg_sig2", 1)
    assert int(e.health().sum()) == 0 and s2[8] > 0
    print("bg_sig2", s2[:9])
print("sanitize_new OK")
