"""One-off diagnostic: humidity row of the time-series integrals, device against oracle, block by block."""
import sys, os, tempfile
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cgenie_b200 import Ensemble, materialise
from oracle_lib import Oracle
I = J = 36; L = 16; LA = 8
d = tempfile.mkdtemp()
materialise(d, "eb_go_gs_ac_bg_36x36x16")
o = Oracle(world="worjh2", maxk=16, maxl=16, nyear=96)
o.biogem_setup()
with Ensemble(d, n_members=1) as e:
    e.set_tracer_variant("strict")
    dts = float(2 * 5) * 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
    for blk in range(1, 4):
        e.run(10); o.run(10)
        tqd = e.get("tq", 0).reshape(J, I, 2)
        tqo = o.f("tq").reshape(J + 2, I + 2, 2)[1:J + 1, 1:I + 1] if o.f("tq").size == (J + 2) * (I + 2) * 2 else o.f("tq").reshape(J, I, 2)
        sfc = o.f("sfcatm1").reshape(J, I, LA)
        qs = o.f("qstar_atm").reshape(J, I) if True else None
        print("blk", blk, "dev tq2 mean %.6e  ora tq2 mean %.6e  ora sfcatm1(2) mean %.6e  ora qstar mean %.6e  max|dev-ora tq2| %.2e"
              % (tqd[..., 1].mean(), tqo[..., 1].mean(), sfc[..., 1].mean(), qs.mean(), np.abs(tqd[..., 1] - tqo[..., 1]).max()))
        e.biogem_sig_update(dts, 1000.0); o.L.cgo_biogem_sig_update(o.h, 1000.0)
        a, b = e.get("bg_sig", 0), o.f("bg_sig")
        print("   sig rows 51/52 dev %.6e %.6e  ora %.6e %.6e" % (a[51], a[52], b[51], b[52]))
