"""tools/nccompare.py (the reference's regression comparer restated: src/tools/nccompare.f90:200-264, tools/tests.py:131-209)
and tools/dump_for_gfortran.py (oracle results in the reference's own output formats + job recipe) -- CPU only."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import nccompare as ncc  # noqa: E402


def test_float_ulps_is_the_bit_pattern_distance_of_the_float32_casts():
    one = np.float32(1.0)
    up = np.nextafter(one, np.float32(2.0))
    assert ncc.float_ulps(1.0, float(up)) == 1
    assert ncc.float_ulps(1.0, 1.0 + 1e-12) == 0                      # below float32 resolution: the cast removes it
    assert ncc.float_ulps(1.0, float(one + 35 * (up - one))) == 35
    assert ncc.float_ulps(-1.0, float(-up)) == 1
    assert ncc.float_ulps(0.0, -0.0) == 2 ** 31                       # the reference's formula shares this quirk (+0 / -0 patterns)
    x = np.array([1.0, 2.0, 3.0])
    assert list(ncc.float_ulps(x, x)) == [0, 0, 0]


def test_verdict_follows_do_comparison():
    a = np.array([1.0, 1e-3, 5.0])
    assert ncc.values_differ(a, a + 5e-15)[0] is False                # under the absolute tolerance
    assert ncc.values_differ(a, a * (1.0 + 2e-6))[0] is False         # 2e-6 relative = 17 - 34 float32 ulps < 35
    assert ncc.values_differ(a, a * (1.0 + 1e-5))[0] is True          # ~ 84 ulps
    assert ncc.values_differ(a, a * (1.0 + 1e-5), max_ulps=0)[0] is True
    assert ncc.values_differ(a, a * (1.0 + 1e-5), abs_tol=0.0)[0] is False     # min_abs = TINY: nccompare.f90 then never fails


def test_ascii_series_lines(tmp_path):
    p, q = tmp_path / "a.res", tmp_path / "b.res"
    p.write_text(" % time (yr) / global DIC (mol)\n       0.500  0.3001407E+19  0.2242385E-02\n")
    q.write_text(" % time (yr) / global DIC (mol)\n       0.500  0.3001408E+19  0.2242385E-02\n")     # one unit of the 7th digit: 2-3 ulps
    assert ncc.compare_ascii(str(p), str(q), out=open(os.devnull, "w")) is False
    q.write_text(" % time (yr) / global DIC (mol)\n       0.500  0.3002407E+19  0.2242385E-02\n")
    assert ncc.compare_ascii(str(p), str(q), out=open(os.devnull, "w")) is True
    q.write_text(" % another header\n       0.500  0.3001407E+19  0.2242385E-02\n")
    assert ncc.compare_ascii(str(p), str(q), out=open(os.devnull, "w")) is True


def test_dump_for_gfortran_config1(built, tmp_path):
    out = tmp_path / "dumps"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "dump_for_gfortran.py"), "--config", "1", "--years", "1",
                           "--out", str(out)])
    d = out / "config1"
    files = sorted(os.listdir(d / "oracle"))
    assert files == ["embm_restart_2001_01_01.nc", "goldsic_restart_2001_01_01.nc", "goldstein_restart_2001_01_01.nc"]
    keys = dict(ln.split("=", 1) for ln in (d / "user_config").read_text().split("\n") if ln)
    assert keys["go_world"] == '"worbe2"' and keys["ma_dim_GOLDSTEINNLEVS"] == "8" and keys["go_nyear"] == "100"
    assert "nccompare.py" in (d / "RECIPE.md").read_text() and "NOT yet compared" in (d / "RECIPE.md").read_text()
    g = str(d / "oracle" / "goldstein_restart_2001_01_01.nc")
    assert ncc.compare_nc(g, g, out=open(os.devnull, "w")) is False
    # a second file with the temperature perturbed by 1e-5 relative fails, by 1e-7 passes (the reference's own tolerance)
    from scipy.io import netcdf_file
    for fac, differs in ((1.0 + 1e-5, True), (1.0 + 1e-7, False)):
        h = str(tmp_path / "mod.nc")
        with netcdf_file(g, "r", mmap=False) as a, netcdf_file(h, "w") as b:
            for n, size in a.dimensions.items():
                b.createDimension(n, size)
            for n, v in a.variables.items():
                w = b.createVariable(n, v.data.dtype, v.dimensions)
                w[:] = v.data * fac if n == "temp" else v.data
        assert ncc.compare_nc(g, h, out=open(os.devnull, "w")) is differs
