#!/usr/bin/env python
"""Two measurements on one B200 (run under gpurun; output to gpurun_out/age_<tag>.log):

(1) cost of a model year as the ocean ages: the bench's 128 perturbed members are advanced from the uniform initial state
    and every 10 model years one year is timed (device-resident, CUDA events) and one instrumented year gives the
    tracer-step family time -- the convective adjustment is data dependent (70 % of all cells sit in mixed regions in a
    4-year-old ocean);
(2) north-star drift criterion at full length: 100 model years of the unperturbed member and one perturbed member on the
    device against the CPU oracle (two oracle threads next to the device run), global means of T, S, DIC, O2 and pCO2.

  python tools/age_sweep.py [--years 120] [--drift-years 100]
"""
import argparse, os, sys, tempfile, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cgenie_b200 import Ensemble, materialise  # noqa: E402
from cgenie_b200.sharding import perturbation_table, shard  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--years", type=int, default=120)
ap.add_argument("--drift-years", type=int, default=100)
ap.add_argument("--members", type=int, default=128)
args = ap.parse_args()
CFG = "eb_go_gs_ac_bg_36x36x16"
d = tempfile.mkdtemp()
materialise(d, CFG)
L, LA = 16, 8

# ---------------------------------------------------------------- (2) started first: the oracle threads need ~100 s
res = {}
if args.drift_years > 0:
    from oracle_lib import Oracle
    tab = perturbation_table(4, biogem=True)
    members = (0, 3)

    def oracle_run(m):
        kw = {k: float(v[m]) for k, v in tab.items()}
        o = Oracle("worjh2", maxk=16, maxl=16, nyear=96, **{k: v for k, v in kw.items() if not k.startswith("par_bio")})
        o.biogem_setup(**{k: v for k, v in kw.items() if k.startswith("par_bio")})
        t0 = time.perf_counter()
        o.run(480 * args.drift_years)
        res[m] = (o.f("ocn").reshape(-1, L).copy(), o.f("bg_M").copy(), o.f("atm").reshape(-1, LA).copy(), time.perf_counter() - t0)
        o.close()
    oth = [threading.Thread(target=oracle_run, args=(m,)) for m in members]
    for t in oth:
        t.start()
    t0 = time.perf_counter()
    with Ensemble(d, n_members=4, perturb=tab) as e:
        e.set_tracer_variant("col")
        e.run(480 * args.drift_years)
        e.synchronize()
        print("drift: device %d years of 4 members in %.1f s, blown %d" % (args.drift_years, time.perf_counter() - t0, int(e.health().sum())), flush=True)
        dev = {m: (e.get("ocn", m).reshape(-1, L), e.get("bg_M", m), e.get("atm", m).reshape(-1, LA)) for m in members}

# ---------------------------------------------------------------- (1)
M = args.members
pert = shard(perturbation_table(M, biogem=True), 0, 1, M)
with Ensemble(d, n_members=M, perturb=pert) as e:
    e.set_tracer_variant("col")
    kyear = e.nyear * e.ndta
    age = 0
    while age < args.years:
        n = 3 if age == 0 else 8   # the bench's state first (3 warm-up years), then every 10 years
        e.run(kyear * n)
        age += n
        e.synchronize()
        e.timer_start()
        e.run(kyear)
        ms = e.timer_stop_ms()
        age += 1
        e.profile(True)
        e.run(kyear)
        e.profile(False)
        age += 1
        fam = {f: e.profile_get(f)[0] for f in ("tstepo_flux", "co", "momentum", "biogem")}
        print("age %4d years: %.2f ms per model year (%.2f M model-years/hour), tracer step %.1f us, momentum %.1f us, biogem %.1f us per block, blown %d" %
              (age, ms, M * 3.6e6 / ms / 1e6, 1e3 * (fam["tstepo_flux"] + fam["co"]) / e.nyear, 1e3 * fam["momentum"] / e.nyear,
               2e3 * fam["biogem"] / e.nyear, int(e.health().sum())), flush=True)

if args.drift_years > 0:
    for t in oth:
        t.join()
    for m in members:
        ocn_o, M_o, atm_o, dt = res[m]
        ocn_d, M_d, atm_d = dev[m]
        print("drift member %d after %d years (oracle %.0f s):" % (m, args.drift_years, dt))
        for l, name in ((0, "T"), (1, "S"), (2, "DIC"), (6, "O2")):
            mo = float((ocn_o[:, l] * M_o).sum() / M_o.sum())
            md = float((ocn_d[:, l] * M_d).sum() / M_d.sum())
            print("  global mean %-4s oracle %.12e device %.12e rel %.2e" % (name, mo, md, abs(md - mo) / abs(mo)))
        print("  pCO2 oracle %.9e device %.9e rel %.2e" % (atm_o[0, 2], atm_d[0, 2], abs(atm_d[0, 2] - atm_o[0, 2]) / atm_o[0, 2]))
        print("  max |ts - ts_oracle| / max|ts| per tracer: " + " ".join("%.1e" % (np.abs(ocn_d[:, l] - ocn_o[:, l]).max() / max(np.abs(ocn_o[:, l]).max(), 1e-300)) for l in range(L)))
