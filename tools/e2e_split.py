#!/usr/bin/env python
"""Where the end-to-end year goes: the per-module C-ABI path of bench.py's e2e leg timed (a) as the bench does (state up,
480 koverall iterations of module calls, state down), (b) without the copies, (c) the copies alone, next to cg_run."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cgenie_b200 import Ensemble, materialise
from cgenie_b200.sharding import perturbation_table, shard
M = 128
d = tempfile.mkdtemp()
materialise(d, "eb_go_gs_ac_bg_36x36x16")
e = Ensemble(d, n_members=M, perturb=shard(perturbation_table(M, biogem=True), 0, 1, M))
e.set_tracer_variant("col")
kyear = e.nyear * e.ndta
spin = int(sys.argv[1]) if len(sys.argv) > 1 else 100
e.run(kyear * spin)
e.synchronize()
pin = {n: torch.empty(e.field_size(n) * e.member_stride, dtype=torch.float64).pin_memory().numpy() for n in ("ts", "tq", "varice")}
for n in pin:
    e.get_all(n, out=pin[n])
gt = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
tick = int(round(1000.0 * gt)); dts = 10.0 * gt
k0 = [spin * kyear]
def up():
    for n in pin: e.put_all(n, pin[n])
    e.put_all("varice1", pin["varice"]); e.put_all("tq1", pin["tq"])
def down():
    for n in pin: e.get_all(n, out=pin[n])
def modules():
    for k in range(1, kyear + 1):
        if k % 5 == 1: e.surflux()
        e.step_embm()
        if k % 5 == 0:
            e.step_seaice(); e.step_goldstein()
        if k % 10 == 0:
            c = (k0[0] + k) * tick
            e.biogem_forcing(c); e.biogem_step(dts, c); e.biogem_tracercoupling(); e.biogem_climate(); e.atchem_step(dts)
    k0[0] += kyear
def timed(name, fn, n=3):
    fn(); e.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    e.synchronize()
    print("%-44s %.2f ms per year" % (name, 1e3 * (time.perf_counter() - t0) / n), flush=True)
timed("cg_run (graphs, resident)", lambda: (e.run(kyear), k0.__setitem__(0, k0[0] + kyear)))
timed("modules, resident (no copies)", modules)
def host_only():
    t0 = time.perf_counter(); modules(); t1 = time.perf_counter(); e.synchronize()
    print("   host time to enqueue one year of module calls: %.2f ms" % (1e3 * (t1 - t0)))
host_only()
timed("up + modules + down (bench e2e)", lambda: (up(), modules(), down()))
timed("up + down only", lambda: (up(), down()))
timed("up only", up)
timed("down only", down)
