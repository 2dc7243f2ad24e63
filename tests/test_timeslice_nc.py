"""fields_biogem_3d.nc (cg_slice_biogem_write_3d / series.write_timeslice_3d) read back with an independent netCDF reader (scipy) and
held to sub_init_netcdf / sub_save_netcdf / sub_save_netcdf_3d (src/biogem/biogem_data_netCDF.f90:148-459, 1959-2315): dimensions and
variables in the reference's order, the unlimited time axis growing by one record per call, surface level first, fill value on dry
cells, temperature in degrees C, isotopes as delta values, D14C, the salinity-normalised and inventory fields, the carbonate rows in
string_carb's order.  Host code only: the window integrals come from a stand-in for the engine."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from cgenie_b200 import series
from cgenie_b200.restart import OCN_TRACERS, SED_TRACERS, biogem_axes

I, J, K = 6, 5, 4
L, LS = len(OCN_TRACERS), len(SED_TRACERS)
FILL = 9.9692099683868690e+36


class _Engine:
    """Just what write_timeslice_3d asks of an Ensemble: grid constants and the "sl_*" window integrals of one member."""
    maxi, maxj, maxk, maxl = I, J, K, L

    def __init__(self, seed, t, grid_seed=1, K=K):
        self.maxk = K
        r = np.random.default_rng(grid_seed)
        self.k1 = np.full((J + 2, I + 2), 99, dtype=np.int32)
        self.k1[1:J + 1, 1:I + 1] = r.integers(1, K + 2, size=(J, I))       # K + 1 = land
        self.k1[1, 1], self.k1[2, 2] = 1, K + 1
        r = np.random.default_rng(seed)
        sv = np.sin(np.linspace(-np.pi / 2, np.pi / 2, J + 1))
        self.c = {"sv": sv, "s": np.concatenate([[0.0], np.sin(0.5 * (np.arcsin(sv[1:]) + np.arcsin(sv[:-1])))]),
                  "dz": np.concatenate([[0.0], np.linspace(0.4, 0.1, K)]), "dza": np.concatenate([[0.0], np.linspace(0.3, 0.12, K)])}
        wet = (np.arange(1, K + 1)[:, None, None] >= self.k1[1:J + 1, 1:I + 1][None]).astype(float)     # (k, j, i)
        ocn = r.uniform(1e-4, 3e-3, size=(K, J, I, L))
        ocn[..., 0] = r.uniform(271.0, 300.0, size=(K, J, I))
        ocn[..., 1] = r.uniform(33.0, 36.0, size=(K, J, I))
        for iso, bulk, frac in ((3, 2, 0.011), (4, 2, 1.1e-12), (9, 8, 0.0109), (10, 8, 1.0e-12)):
            ocn[..., iso] = ocn[..., bulk] * frac * r.uniform(0.98, 1.02, size=(K, J, I))
        part = r.uniform(1e-7, 1e-6, size=(K, J, I, LS))
        for iso, bulk, frac in ((1, 0, 0.0108), (2, 0, 1.0e-12), (5, 4, 0.0112), (6, 4, 1.1e-12)):
            part[..., iso] = part[..., bulk] * frac
        part[0, 0, 0, 4:7] = 0.0                                                  # no CaCO3 in one cell: the null value
        self.f = {"sl_ocn": (ocn * wet[..., None] * t).ravel(), "sl_part": (part * wet[..., None] * t).ravel(),
                  "sl_carb": (r.uniform(1e-9, 1e-3, size=(K, J, I, 10)) * wet[..., None] * t).ravel(),
                  "sl_carbconst": (r.uniform(1e-10, 1e-2, size=(K, J, I, 17)) * wet[..., None] * t).ravel(), "sl_t": np.array([t])}
        self.wet = wet.astype(bool)

    def iconst(self, name):
        return self.k1.ravel()

    def const(self, name):
        return self.c[name]

    def get(self, name, member=0):
        return self.f[name].copy()


def _delta(tot, iso, std):
    with np.errstate(divide="ignore", invalid="ignore"):
        f = iso / tot
        d = 1000.0 * ((f / (1.0 - f)) / std - 1.0)
    return np.where(tot > 0.999999e-19, d, -0.999999e+19)


def _expect(e, name):
    """The reference's arithmetic for one variable, written independently of the product: (zt, lat, lon), surface first."""
    t = e.f["sl_t"][0]
    ocn = e.f["sl_ocn"].reshape(K, J, I, L) / t
    names = [n for n, _ in OCN_TRACERS]
    if name == "ocn_temp":
        v = ocn[..., 0] - 273.15
    elif name in ("ocn_DIC_13C", "ocn_DOM_C_13C"):
        v = _delta(ocn[..., names.index(name[4:-4])], ocn[..., names.index(name[4:])], 0.011202)
    elif name in ("ocn_DIC_14C", "ocn_DOM_C_14C"):
        v = _delta(ocn[..., names.index(name[4:-4])], ocn[..., names.index(name[4:])], 1.176e-12)
    elif name == "ocn_DIC_D14C":
        d13, d14 = _delta(ocn[..., 2], ocn[..., 3], 0.011202), _delta(ocn[..., 2], ocn[..., 4], 1.176e-12)
        v = 1000.0 * ((1.0 + d14 / 1000.0) * 0.975 ** 2 / (1.0 + d13 / 1000.0) ** 2 - 1.0)
    elif name.endswith("_Snorm"):
        m = series.ocean_mass(e).reshape(K, J, I)
        mean_s = np.sum(e.f["sl_ocn"].reshape(K, J, I, L)[..., 1] * m) / np.sum(m)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = e.f["sl_ocn"].reshape(K, J, I, L)[..., names.index(name[4:-6])] * (mean_s / e.f["sl_ocn"].reshape(K, J, I, L)[..., 1]) / t
    elif name.endswith("_tot"):
        v = series.ocean_mass(e).reshape(K, J, I) * ocn[..., names.index(name[4:-4])]
    elif name.startswith("ocn_"):
        v = ocn[..., names.index(name[4:])]
    elif name.startswith("carb_const_"):
        v = e.f["sl_carbconst"].reshape(K, J, I, 17)[..., series.SL_CARBCONST.index(name[11:])] / t
    elif name.startswith("carb_"):
        v = e.f["sl_carb"].reshape(K, J, I, 10)[..., series.SL_CARB.index(name[5:])] / t
    elif name[-4:] in ("_13C", "_14C"):
        p = e.f["sl_part"].reshape(K, J, I, LS) / t
        sn = [n for n, _ in SED_TRACERS]
        v = _delta(p[..., sn.index(name[9:-4])], p[..., sn.index(name[9:])], 0.011202 if name.endswith("13C") else 1.176e-12)
    else:
        v = e.f["sl_part"].reshape(K, J, I, LS)[..., [n for n, _ in SED_TRACERS].index(name[9:])] / t
    return np.where(e.wet, v, FILL)[::-1]


def test_time_slice_file_two_records(built, tmp_path):
    p = str(tmp_path / "biogem" / "fields_biogem_3d.nc")
    engines = [_Engine(1, 1.0), _Engine(2, 0.5)]       # one grid, two save windows
    series.write_timeslice_3d(engines[0], p, 0.5, run_id="exp1", carbconst=True)
    series.write_timeslice_3d(engines[1], p, 9.5, run_id="exp1", carbconst=True)
    with netcdf_file(p, "r", mmap=False) as f:
        assert list(f.dimensions)[:6] == ["time", "xu", "lon", "lat", "zt", "yu"] and f.dimensions["time"] is None
        assert list(f.dimensions)[6:] == ["lon_edges", "lat_edges", "zt_edges", "xu_edges", "yu_edges", "lat_moc", "zt_moc",
                                          "lat_moc_edges", "zt_moc_edges", "para"]
        assert f.dimensions["lat_moc"] == J + 1 and f.dimensions["zt_moc_edges"] == K + 2 and f.dimensions["para"] == 1
        assert f.title == b"Time averaged integrals" and f.time_unit == b"Year mid-point" and f.experiment_name == b"exp1"
        names = list(f.variables)
        assert names[:15] == ["time", "year", "lon", "lat", "zt", "xu", "yu", "lon_edges", "lat_edges", "zt_edges", "xu_edges", "yu_edges",
                              "grid_level", "grid_mask", "grid_topo"]
        ocn_names = ["ocn_" + n for n, _ in OCN_TRACERS]
        assert names[15:15 + L] == ocn_names and names[15 + L] == "ocn_DIC_D14C"
        snorm = [n for n in names if n.endswith("_Snorm")]
        assert snorm == ["ocn_%s_Snorm" % n for (n, _), ty in zip(OCN_TRACERS, series.OCN_TYPE) if ty == 1]
        tot = [n for n in names if n.endswith("_tot")]
        assert tot == ["ocn_%s_tot" % n for (n, _), ty in zip(OCN_TRACERS, series.OCN_TYPE) if ty >= 1]
        assert [n for n in names if n.startswith("carb_") and not n.startswith("carb_const_")] == ["carb_" + n for n in series.REF_CARB]
        assert [n for n in names if n.startswith("carb_const_")] == ["carb_const_" + n for n in series.REF_CARBCONST]
        assert [n for n in names if n.startswith("bio_part_")] == ["bio_part_" + n for n, _ in SED_TRACERS[:7]]      # no *_frac2 variables
        assert np.array_equal(f.variables["time"][:], [0.5, 9.5]) and f.variables["time"].axis == b"T"
        assert np.array_equal(f.variables["year"][:], [1.0, 10.0])      # NINT rounds halves away from zero
        ax = biogem_axes(I, J, K, engines[0].c["s"], engines[0].c["sv"], engines[0].c["dz"], engines[0].c["dza"])
        for n, a in zip(("lon", "lat", "lon_edges", "lat_edges", "zt", "zt_edges"), ax):
            assert np.array_equal(f.variables[n][:], a), n
        assert np.array_equal(f.variables["xu"][:], ax[2][:I]) and np.array_equal(f.variables["yu"][:], ax[3][:J])
        assert np.array_equal(f.variables["xu_edges"][:I], ax[0]) and f.variables["xu_edges"][I] == ax[0][-1] + 360.0 / I
        assert np.array_equal(f.variables["yu_edges"][:J], ax[1]) and f.variables["yu_edges"][J] == ax[1][-1] + (ax[3][J] - ax[3][J - 1])
        assert f.variables["lon"].edges == b"lon_edges" and f.variables["zt"].units == b"m"
        k1 = engines[0].k1[1:J + 1, 1:I + 1]
        assert np.array_equal(f.variables["grid_level"][:], k1) and f.variables["grid_level"].dimensions == ("lat", "lon")
        assert np.array_equal(f.variables["grid_mask"][:], np.where(k1 <= K, 1.0, FILL).astype(np.float32))
        topo = np.where(k1 <= K, ax[5][np.clip(K - k1 + 1, 0, K)], FILL).astype(np.float32)
        assert np.array_equal(f.variables["grid_topo"][:], topo)
        v = f.variables["ocn_DIC"]
        assert v.dimensions == ("time", "zt", "lat", "lon") and v.data.dtype == np.dtype(">f4") and v.units == b"mol kg-1"
        assert np.array_equal(v.valid_range, np.array([-9.99E+2, 9.99E-1], dtype=np.float32)) and v.missing_value == FILL
        assert v.long_name == b"dissolved inorganic carbon (DIC)"
        assert not hasattr(f.variables["carb_H"], "units") and not hasattr(f.variables["carb_H"], "valid_range")
        assert f.variables["bio_part_POC_13C"].units == b"o/oo" and f.variables["bio_part_POC"].units == b"mol kg-1"
        for rec, e in enumerate(engines):
            for n in names[15:]:
                got = f.variables[n][rec]
                want = _expect(e, n).astype(np.float32)
                ok = (got == want) | (np.abs(got - want) <= 2e-7 * np.abs(want))
                assert np.all(ok), (rec, n, got[~ok][:3], want[~ok][:3])
        assert f.variables["bio_part_CaCO3_13C"][0][K - 1, 0, 0] == np.float32(-0.999999e+19)     # fun_calc_isotope_delta's dum_null


def test_time_slice_refuses_another_grid_and_an_empty_window(built, tmp_path):
    p = str(tmp_path / "fields_biogem_3d.nc")
    e = _Engine(3, 1.0)
    series.write_timeslice_3d(e, p, 0.5, derived=False)
    with netcdf_file(p, "r", mmap=False) as f:
        assert not [n for n in f.variables if n.endswith("_tot") or n.startswith("bio_part_") or n.startswith("carb_const_")]
        assert not hasattr(f, "experiment_name")
    e.f["sl_t"][0] = 0.0
    with pytest.raises(series.SeriesError, match="not positive"):
        series.write_timeslice_3d(e, p, 1.5)
    e2 = _Engine(3, 1.0, K=K - 1)
    with pytest.raises(series.SeriesError):
        series.write_timeslice_3d(e2, p, 1.5)


class _Recorder(_Engine):
    """The stand-in with the two engine calls SliceSaver makes: the integrals grow by dtyr per update and are zeroed by a reset."""

    def __init__(self):
        super().__init__(5, 1.0)
        self.unit = {n: v.copy() for n, v in self.f.items()}
        self.calls = []
        self.biogem_slice_reset()

    def biogem_slice_update(self, dts):
        self.calls.append("u")
        for n in self.f:
            self.f[n] = self.f[n] + self.unit[n] * (dts / series.YR_S)

    def biogem_slice_reset(self):
        self.calls.append("r")
        for n in self.f:
            self.f[n] = np.zeros_like(self.unit[n])


def test_slice_saver_windows(built, tmp_path):
    """A 10-year run of 48 BIOGEM steps per year, slices of one year centred on years 0.5, 4.5 and 9.5 (biogem_save_timeslice.dat lists
    mid-points): three records at those years, each the mean over exactly the 48 steps of its window; steps outside a window never
    touch the device integrals."""
    p = str(tmp_path / "fields_biogem_3d.nc")
    e = _Recorder()
    sv = series.SliceSaver(e, p, t_runtime=10.0, save_times=[0.5, 4.5, 9.5], slice_dt=1.0, run_id="w")
    dts = series.YR_S / 48.0
    got = []
    for istep in range(1, 10 * 48 + 1):
        y = sv.step(dts, int(round(istep * dts * 1000.0)))
        if y is not None:
            got.append((istep, y))
    assert [y for _, y in got] == [0.5, 4.5, 9.5] and sv.saved == [0.5, 4.5, 9.5]
    assert [i for i, _ in got] == [48, 5 * 48, 10 * 48]
    assert e.calls.count("u") == 3 * 48 and e.calls.count("r") == 5 and sv.i == 0      # the stand-in's own, the saver's, one per record
    with netcdf_file(p, "r", mmap=False) as f:
        assert np.array_equal(f.variables["time"][:], [0.5, 4.5, 9.5])
        want = np.where(e.wet, e.unit["sl_ocn"].reshape(K, J, I, L)[..., 2], FILL)[::-1].astype(np.float32)
        for rec in range(3):
            assert np.allclose(f.variables["ocn_DIC"][rec], want, rtol=3e-7, atol=0.0)
    # no date inside the run: nothing is saved, unless ctrl_data_save_slice_autoend adds the last half window
    e2 = _Recorder()
    sv2 = series.SliceSaver(e2, str(tmp_path / "none.nc"), t_runtime=2.0, save_times=[50.5])
    assert sv2.i == 0 and sv2.step(dts, 1000) is None and e2.calls == ["r", "r"]
    sv3 = series.SliceSaver(_Recorder(), str(tmp_path / "end.nc"), t_runtime=2.0, save_times=[50.5], autoend=True)
    assert sv3.i == 1 and sv3.ts == [0.5]
