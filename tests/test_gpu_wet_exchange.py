"""Wet-cell packed exchange of 3-D ocean fields (cg_sync_all_wet_to_host / _from_host): the packed form holds exactly the wet
cells of the dense form, an upload changes wet cells only (both ping-pong buffers of ts), and a state that went out and came
back packed continues bit for bit like one that never left the device.  Run with -m gpu on a B200."""
import numpy as np
import pytest

from cgenie_b200 import Ensemble, materialise

pytestmark = pytest.mark.gpu


def test_wet_exchange_round_trip(built, tmp_path):
    materialise(str(tmp_path), "eb_go_gs_ac_bg_36x36x16")
    M = 5
    pert = {"diff1": np.linspace(1800.0, 2400.0, M)}
    with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e, Ensemble(str(tmp_path), n_members=M, perturb=pert) as f:
        for x in (e, f):
            x.set_tracer_variant("col")
            x.run(100)
        I, J, K, L, MS = e.maxi, e.maxj, e.maxk, e.maxl, e.member_stride
        k1 = e.iconst("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
        wet = (np.arange(1, K + 1)[:, None, None] >= k1[None]).ravel()
        for name, inner in (("ts", L), ("rho", 1), ("u", 3), ("ocn", L)):
            dense = e.get_all(name).reshape(K * J * I, inner, MS)
            packed = e.get_all_wet(name).reshape(-1, inner, MS)
            assert e.wet_size(name) == int(wet.sum()) * inner
            assert np.array_equal(packed, dense[wet]), name
        # upload: wet cells replaced, dry cells untouched
        ts = e.get_all("ts").reshape(K * J * I, L, MS).copy()
        packed = e.get_all_wet("ts").copy()
        e.put_all_wet("ts", packed * 1.5)
        now = e.get_all("ts").reshape(K * J * I, L, MS)
        assert np.array_equal(now[wet], ts[wet] * 1.5) and np.array_equal(now[~wet], ts[~wet])
        e.put_all_wet("ts", packed)
        assert np.array_equal(e.get_all("ts").reshape(K * J * I, L, MS), ts)
        # out and back in, then on: identical to the ensemble that stayed resident
        e.run(50)
        f.run(50)
        for name in ("ts", "ocn", "rho", "tq"):
            assert np.array_equal(e.get_all(name), f.get_all(name)), name
        with pytest.raises(Exception):
            e.get_all_wet("tq")                    # not a 3-D ocean field


def test_double_buffered_exchange(built, tmp_path):
    """cg_exchange_*: what the staged copies deliver is what the blocking calls deliver, a commit into two fields, uploads and
    downloads in flight next to running work, and an ensemble whose state left and re-entered through the staging buffers every
    block goes on bit for bit like one that stayed resident."""
    import torch
    materialise(str(tmp_path), "eb_go_gs_ac_bg_36x36x16")
    M = 6
    pert = {"diff1": np.linspace(1800.0, 2400.0, M)}
    pin = lambda n: torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    with Ensemble(str(tmp_path), n_members=M, perturb=pert) as e, Ensemble(str(tmp_path), n_members=M, perturb=pert) as f:
        for x in (e, f):
            x.set_tracer_variant("col")
            x.run(100)
        MS = e.member_stride
        names = (("ts", True), ("tq", False), ("varice", False))
        out = {n: pin((e.wet_size(n) if w else e.field_size(n)) * MS) for n, w in names}
        for n, w in names:
            e.exchange_begin_download(n, out[n], wet=w)
        e.run(10)                                   # the copies cross while this runs; they hold the state BEFORE it
        e.exchange_wait()
        ref = {n: (f.get_all_wet(n) if w else f.get_all(n)) for n, w in names}       # f stayed at iteration 100
        for n, w in names:
            assert np.array_equal(out[n], ref[n]), n
        # staged uploads: e (now at 110) takes f's state through the staging buffers, tq also into tq1, varice into varice1
        cur = {n: pin(out[n].size) for n, _ in names}
        for n, w in names:
            cur[n][:] = ref[n]
            e.exchange_begin_upload(n, cur[n], wet=w)
        for n, _ in names:
            e.exchange_commit_upload(n, also={"tq": "tq1", "varice": "varice1"}.get(n))
        e.exchange_wait()
        for n, w in names:
            assert np.array_equal(e.get_all_wet(n) if w else e.get_all(n), ref[n]), n
        assert np.array_equal(e.get_all("tq1"), f.get_all("tq")) and np.array_equal(e.get_all("varice1"), f.get_all("varice"))
        with pytest.raises(Exception):
            e.exchange_commit_upload("ts")          # nothing staged
        assert int(e.health().sum()) == 0
