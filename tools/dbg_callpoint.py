"""Diagnostic behind round 1's open item: the BIOGEM time-series integrals sampled inside the
BIOGEM block, device against oracle, block by block.  Usage: dbg_callpoint.py [blocks] [variant] [nosol]

It established (profiles/dbg_callpoint_r2a.log, with "nosol" = without the biogem_climate_sol call genie.f90 makes ahead of
the first BIOGEM step, as round 1's test loop was written) that the device's ocn / cell masses between step_biogem and
biogem_tracercoupling are bit for bit what the previous block left -- the device holds the step's changes in vdocn exactly as
the Fortran does -- and that device and oracle differed after the very first block (DOM by 100 %): a test-loop bug, not a
device / oracle difference.  The call point is genie.f90's (:401-405: behind biogem_climate, ahead of ATCHEM)."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cgenie_b200 import Ensemble, materialise  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

I = J = 36
K = L = 16
LA = 8
NBLK = int(sys.argv[1]) if len(sys.argv) > 1 else 6
VARIANT = sys.argv[2] if len(sys.argv) > 2 else "col"
NOSOL = len(sys.argv) > 3 and sys.argv[3] == "nosol"
d = tempfile.mkdtemp()
materialise(d, "eb_go_gs_ac_bg_36x36x16")
o = Oracle(world="worjh2", maxk=16, maxl=16, nyear=96)
o.biogem_setup(par_bio_k0_PO4=1.9e-6)
o.L.cgo_biogem_sig_auto(o.h, 1, 1000.0)
with Ensemble(d, n_members=2, perturb={"par_bio_k0_PO4": np.array([1.9e-6, 2.3e-6])}) as e:
    e.set_tracer_variant(VARIANT)
    genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
    tick = int(round(1000.0 * genie_timestep))
    dts = float(2 * 5) * genie_timestep
    e.biogem_sig_reset()
    prev = {n: e.get(n, 0) for n in ("ocn", "bg_M")}
    for blk in range(1, NBLK + 1):
        for k in range(10 * (blk - 1) + 1, 10 * blk + 1):
            if k % 5 == 1:
                e.surflux()
            e.step_embm()
            if k % 5 == 0:
                e.step_seaice()
                e.step_goldstein()
            if k % 10 == 0:
                if k == 10 and not NOSOL:
                    e.biogem_climate_sol()                                # genie.f90:369-370
                e.biogem_forcing(k * tick)
                e.biogem_step(dts, k * tick)
                probe = blk % 2 == 0      # every second block: read the state between step and coupling (a host read joins the streams)
                if probe:
                    now = {n: e.get(n, 0) for n in ("ocn", "bg_M")}
                    same = {n: bool(np.array_equal(now[n], prev[n])) for n in now}
                e.biogem_tracercoupling()
                e.biogem_climate()
                e.biogem_sig_update(dts, 1000.0)                          # genie.f90:401-405
                e.atchem_step(dts)
        o.f("bg_sig")[:] = 0.0
        o.run(10)
        a, b = e.get("bg_sig", 0), o.f("bg_sig").copy()
        e.biogem_sig_reset()
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        msg = "blk %d: bg_sig worst rel %.2e (row %d); DIC glob %.2e sur %.2e ben %.2e" % (
            blk, rel.max(), int(rel.argmax()), rel[3 + 2], rel[3 + L + 2], rel[3 + 2 * L + 2])
        if probe:
            msg += "; state between step_biogem and tracercoupling == state behind the previous block: %s" % same
        print(msg)
        prev = {n: e.get(n, 0) for n in ("ocn", "bg_M")}
        oc = o.f("ocn").reshape(-1, L)
        dv = prev["ocn"].reshape(-1, L)
        sc = np.abs(oc).max(axis=0) + 1e-300
        print("       behind the block: ocn worst per-tracer difference / scale: %s" % " ".join("%.1e" % x for x in (np.abs(dv - oc) / sc).max(axis=0)))
