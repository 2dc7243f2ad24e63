"""Per-cell proof of the production tracer step ('col': FMA, re-associated coefficients, reciprocal + Newton divisions)
against the 'strict' variant (reference operation order, bit-identical to the oracle: tests/test_gpu_parity.py) from ONE
common state, for every lane of the ensemble.  TEST INFRASTRUCTURE (used by tests/test_gpu_col_proof.py and smoke()).

Why not simply compare trajectories: the convective adjustment (co, goldstein.f90:2657-2777) tests rho(k) < rho(k-1); where a
column is neutrally stable to the last bit -- the first months from the uniform initial state, or the base of a mixed layer --
a 1-ulp difference flips the decision and the column's tracers are redistributed (conservatively) over other levels.  From
then on the two trajectories differ by O(1e-4) in those cells although every arithmetic step agrees to 1e-13.  So the step is
taken twice from the same state and the columns whose decisions differ are reported (flip rate) and checked for what a flip
must preserve (the column inventory of every tracer), while EVERY other cell is held to the per-step bar per cell."""
import numpy as np


def one_step_both(e):
    """One tstepo in 'strict' and one in 'col' from the ensemble's present state (restored afterwards, variant left at 'col').
    Returns (ts_strict, ts_col, cost_strict, cost_col, cost_before) in the device layout [...][member_stride]."""
    names = ["ts", "rho", "cost"] + (["sst"] if e.L.cg_field_size(e.h, b"sst") > 0 else [])   # everything tstepo rewrites
    snap = {n: e.get_all(n).copy() for n in names}
    out = {}
    for variant in ("strict", "col"):
        for n, a in snap.items():
            e.put_all(n, a)
        e.set_tracer_variant(variant)
        if variant == "col":
            assert e.tracer_variant_active() == "col"
        e._ck(e.L.cg_tracer_step(e.h, 1))
        out[variant] = (e.get_all("ts").copy(), e.get_all("cost").copy())
    for n, a in snap.items():
        e.put_all(n, a)
    return out["strict"][0], out["col"][0], out["strict"][1], out["col"][1], snap["cost"]


def check_step(e, tol=1e-10, floor_frac=1e-3, what="", strict_assert=True):
    """Asserts the per-cell bar outside flipped columns and conservation inside them; returns a dict of statistics.

    per-cell error of tracer l = |col - strict| / max(|strict|, floor_frac * max|tracer l|): relative to the CELL's own value,
    floored for the cells of a field that pass through zero (temperature in degC, the salinity anomaly)."""
    I, J, K, L, M, MS = e.maxi, e.maxj, e.maxk, e.maxl, e.n_members, e.member_stride
    ts_s, ts_c, cost_s, cost_c, cost0 = one_step_both(e)
    k1 = e.iconst("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
    wet3 = np.arange(1, K + 1)[:, None, None] >= k1[None]                       # [K][J][I]
    a = ts_s.reshape(K, J, I, L, MS)[..., :M]
    b = ts_c.reshape(K, J, I, L, MS)[..., :M]
    # A column's convection decisions differ ("flipped") if the two results have different mixed regions.  The adjustment
    # leaves bit-identical T and S on all levels of a region (goldstein.f90:2732-2764 copies the merged box's values), so the
    # region map of a result is: level k continues the region of k-1 iff both T and S are bit-equal.  Equal maps = the same
    # final regions (the convection counter cost is compared too); inside equal regions the values agree to rounding.
    def region_map(x):
        same = (x[1:, :, :, 0, :] == x[:-1, :, :, 0, :]) & (x[1:, :, :, 1, :] == x[:-1, :, :, 1, :])      # [K-1][J][I][M]
        return same & wet3[1:, :, :, None] & wet3[:-1, :, :, None]
    flipped = (region_map(a) != region_map(b)).any(axis=0)                                       # [J][I][M]
    flipped |= (cost_s.reshape(J, I, MS)[..., :M] != cost_c.reshape(J, I, MS)[..., :M])
    events = (cost_s.reshape(J, I, MS)[..., :M] - cost0.reshape(J, I, MS)[..., :M])
    wetcol = (k1 <= K)
    scale = np.abs(np.where(wet3[..., None, None], a, 0.0)).reshape(-1, L, M).max(axis=0)       # [L][M]
    denom = np.maximum(np.abs(a), floor_frac * np.maximum(scale, 1e-300)[None, None, None])
    err = np.abs(b - a) / denom
    ok_cells = wet3[..., None, None] & ~flipped[None, :, :, None, :]
    errok = np.where(ok_cells, err, 0.0)
    worst = float(errok.max())
    wk, wj, wi, wl, wm = np.unravel_index(int(errok.argmax()), errok.shape)
    n_cols = int(wetcol.sum()) * M
    n_flip = int((flipped & wetcol[..., None]).sum())
    stats = {"worst_unflipped": worst, "flip_rate": n_flip / max(n_cols, 1), "flipped_columns": n_flip, "columns": n_cols,
             "convecting_columns": int(((events > 0) & wetcol[..., None]).sum()),
             "worst_at": {"k": int(wk) + 1, "j": int(wj) + 1, "i": int(wi) + 1, "tracer": int(wl), "member": int(wm),
                          "strict": float(a[wk, wj, wi, wl, wm]), "col": float(b[wk, wj, wi, wl, wm]),
                          "tracer_scale": float(scale[wl, wm])},
             "worst_by_tracer": [float(x) for x in errok.max(axis=(0, 1, 2, 4))]}
    # a flipped decision moves tracer between the levels of its column, thickness weighted: column inventories are unchanged
    if n_flip:
        dz = np.asarray(e.const("dz"), dtype=np.float64)[1:K + 1]
        inv_a = (np.where(wet3[..., None, None], a, 0.0) * dz[:, None, None, None, None]).sum(axis=0)   # [J][I][L][M]
        inv_b = (np.where(wet3[..., None, None], b, 0.0) * dz[:, None, None, None, None]).sum(axis=0)
        mag = (np.where(wet3[..., None, None], np.abs(a), 0.0) * dz[:, None, None, None, None]).sum(axis=0)
        d = np.abs(inv_b - inv_a) / np.maximum(mag, 1e-300)
        stats["worst_flipped_inventory"] = float(np.where(flipped[:, :, None, :], d, 0.0).max())
        stats["worst_flipped_cell"] = float(np.where(wet3[..., None, None] & flipped[None, :, :, None, :], err, 0.0).max())
    print("%s col vs strict, one step, %d lanes: worst per-cell error outside flipped columns %.2e; convecting columns %d of %d; "
          "flipped %d (rate %.2e)%s" % (what, M, worst, stats["convecting_columns"], n_cols, n_flip, stats["flip_rate"],
                                        "; in flipped columns: worst cell %.1e, worst column inventory %.1e" %
                                        (stats["worst_flipped_cell"], stats["worst_flipped_inventory"]) if n_flip else ""))
    if worst > tol:
        print("   worst cell:", stats["worst_at"], "per tracer:", ["%.1e" % x for x in stats["worst_by_tracer"]])
    if strict_assert:
        assert worst <= tol, (what, stats)
        if n_flip:
            assert stats["worst_flipped_inventory"] <= 1e-12, (what, stats)
    return stats
