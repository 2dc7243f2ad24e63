#!/usr/bin/env python
"""Experiment: G concurrent 128-member ensembles (one handle each, own streams and graphs) on ONE GPU, driven from G host
threads -- what a 256- or 512-member shard per GPU costs when it is run as 128-member groups."""
import os, sys, tempfile, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cgenie_b200 import Ensemble, materialise  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
M = int(sys.argv[2]) if len(sys.argv) > 2 else 128
years = int(sys.argv[3]) if len(sys.argv) > 3 else 4
d = tempfile.mkdtemp()
materialise(d, "eb_go_gs_ac_bg_36x36x16")
rng = np.random.default_rng(20261017)
ens = []
for g in range(G):
    pert = {"adrag": rng.uniform(2.0, 3.0, M), "diff1": rng.uniform(1600.0, 2500.0, M), "par_bio_k0_PO4": rng.uniform(1.6e-6, 2.4e-6, M)}
    e = Ensemble(d, n_members=M, perturb=pert)
    e.set_tracer_variant("col")
    ens.append(e)
def run(e, n):
    e.run(480 * n)
    e.synchronize()
for phase, n in (("spin", 3), ("timed", years)):
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(e, n)) for e in ens]
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print("%s: G=%d x M=%d, %d years: %.1f ms per year of all members, %.0f model-years/hour, blown %d" %
          (phase, G, M, n, 1e3 * dt / n, G * M * n * 3600.0 / dt, sum(int(e.health().sum()) for e in ens)))
