#!/usr/bin/env python
"""Small driver for ncu captures: spins the 36x36x16, 16-tracer, M-member ensemble a few ocean steps
and then runs `--steps` more (the ones ncu should look at with -s/-c or -k filters)."""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgenie_b200 import Ensemble, materialise  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=128)
ap.add_argument("--spin", type=int, default=10, help="ocean steps before the measured ones")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--variant", default="fast")
ap.add_argument("--config", default="eb_go_gs_ac_bg_36x36x16")
ap.add_argument("--profile", action="store_true", help="print CUDA-event time per kernel family")
ap.add_argument("--hash", action="store_true", help="print a sha256 of the final ts (knob variants that claim bit-identity must agree)")
ap.add_argument("--perturb", action="store_true", help="the bench's parameter-perturbed ensemble instead of identical members")
a = ap.parse_args()
d = tempfile.mkdtemp()
materialise(d, a.config)
pert = None
if a.perturb:
    from cgenie_b200.sharding import perturbation_table
    pert = perturbation_table(a.members, biogem="ac_bg" in a.config)
e = Ensemble(d, n_members=a.members, perturb=pert)
e.set_tracer_variant(a.variant)
e.set_graphs(False)
L, I, J, K = e.maxl, e.maxi, e.maxj, e.maxk
if "ac_bg" not in a.config and L > 2:   # passive tracers of the physics-only configurations
    ts = e.get_all("ts").reshape(K, J, I, L, e.member_stride)
    kk, jj, ii = np.meshgrid(np.arange(1, K + 1), np.arange(1, J + 1), np.arange(1, I + 1), indexing="ij")
    for l in range(2, L):
        ts[:, :, :, l, :] = (1.0 + 0.1 * np.sin(2 * np.pi * ii / I) * np.cos(np.pi * jj / J) * (kk / K) * (1 + l / L))[..., None]
    e.put_all("ts", ts)
e.run(5 * a.spin)
e.synchronize()
if a.profile:
    e.profile(True)
e.run(5 * a.steps)
e.synchronize()
if a.profile:
    fam = {f: e.profile_get(f) for f in ("tstepo_flux", "co", "momentum", "embm", "surflux", "seaice", "biogem")}
    print("cfg=%s variant=%s M=%d us/step: " % (os.environ.get("CG_TRACER_CFG", "0"), a.variant, a.members) +
          " ".join("%s=%.1f" % (k, 1e3 * v[0] / a.steps) for k, v in fam.items()))
if a.hash:
    import hashlib
    print("ts sha256", hashlib.sha256(np.ascontiguousarray(e.get_all("ts")).tobytes()).hexdigest()[:16])
print("done", e.launch_count())
