#!/usr/bin/env python
"""Pinning aid (mitigation 4): results of this repo's oracle -- and, with --device on a GPU box, of the CUDA path
-- written in the reference's OWN output formats, next to the exact job recipe, so that anyone with gfortran + netCDF can run
the real cGENIE on the same configuration and compare with the reference's own pass criterion (tools/nccompare.py =
src/tools/nccompare.f90's 6e-15 / 35-float32-ulp rule, tools/tests.py:131-209).

This container has no Fortran compiler and the reference ships neither base configs nor known-good outputs (they live in the
un-vendored cgenie-data / cgenie-test repositories, .travis.yml:11-12), so the comparison cannot be made here: parity stays
"unpinned" until someone runs the recipe.  What this tool guarantees is that the files are what the reference would write for
the same state: restart netCDF layouts (goldstein_data.f90:153-300, embm_data.f90, gold_seaice_data.f90,
biogem_data_netCDF.f90:24-142, atchem_data_netCDF.f90:22-109) and biogem_series_*.res lines (biogem_data_ascii.f90:669-935).

  python tools/dump_for_gfortran.py --config 1 --years 10 --out dumps/      # eb_go_gs 36x36x8 (BASELINE config #1)
  python tools/dump_for_gfortran.py --config 2 --years 2 --out dumps/       # eb_go_gs_ac_bg 36x36x16 (config #2, shortened)
  ... --device                                                             # also run the CUDA path (strict variant) and dump it
  ... --set go_imld=1 --set go_iconv=1                                     # GOLDSTEIN options on top: in the oracle,
                                                                           # the device job and the user_config alike
Config 2 also writes the export / air-sea flux / misc series (fexport_*, fseaair_*, focnatm_*, misc_seaice, misc_opsi, misc_atm_D14C,
misc_SLT: biogem_data_ascii.f90:955-1096, 1245-1340) the reference writes at its default save level.

Output: <out>/config<N>/{oracle,device}/..., <out>/config<N>/user_config (the job's namelist keys in new-job's prefix form),
<out>/config<N>/RECIPE.md."""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CFG = {1: ("eb_go_gs_36x36x8", dict(world="worbe2", maxk=8, maxl=2, nyear=100)),
       2: ("eb_go_gs_ac_bg_36x36x16", dict(world="worjh2", maxk=16, maxl=16, nyear=96))}
I = J = 36


class OracleAsEnsemble:
    """The oracle behind the small part of the Ensemble interface that cgenie_b200.restart / .series use (get one member's
    field in the Fortran-shaped flat order, grid constants), so the same writers serve both sides."""

    def __init__(self, o, okw):
        self.o = o
        self.maxi, self.maxj, self.maxk, self.maxl, self.nyear = I, J, int(okw["maxk"]), int(okw["maxl"]), int(okw["nyear"])

    def iconst(self, name):
        return np.asarray(self.o.i(name), dtype=np.int32)

    def const(self, name):
        return np.asarray(self.o.f(name), dtype=np.float64)

    def get(self, name, member=0):
        o, K, L = self.o, self.maxk, self.maxl
        if name == "ts":
            return o.f("ts").reshape(K + 2, J + 2, I + 2, L)[1:K + 1, 1:J + 1, 1:I + 1, :].ravel().copy()
        if name == "u":
            return o.f("u").reshape(K, J + 1, I + 1, 3)[:, 1:, 1:, :].ravel().copy()
        if name == "tice":
            return o.f("temp_sic").copy()
        if name == "albice":
            return o.f("albd_sic").copy()
        return o.f(name).copy()

    def field_size(self, name):
        return self.get(name).size


OPTS = {}      # --set go_imld=1 ...: GOLDSTEIN options on top of the configuration (oracle, device job and user_config alike)


def _oracle_opts():
    return {k[3:]: v for k, v in OPTS.items() if k.startswith("go_")}


def user_config(cfgname):
    """The job's namelists as new-job user-config lines (`<prefix>_<key>=<value>`, tools/config_utils.py:253-281)."""
    from cgenie_b200 import materialise
    pref = {"data_genie": "ma", "data_GOLD": "go", "data_EMBM": "ea", "data_goldSIC": "gs", "data_GEM": "gm",
            "data_BIOGEM": "bg", "data_ATCHEM": "ac"}
    lines = []
    with tempfile.TemporaryDirectory() as d:
        materialise(d, cfgname, overrides=dict(OPTS) or None)
        for f, p in pref.items():
            path = os.path.join(d, f)
            if not os.path.exists(path):
                continue
            for ln in open(path).read().split("\n"):
                if "=" in ln and not ln.startswith("&"):
                    k, v = ln.rstrip(",").split("=", 1)
                    lines.append("%s_%s=%s" % (p, k, v))
    return lines


def dump_physics(e, outdir, year):
    from cgenie_b200.restart import write_restart
    return write_restart(e, outdir, member=0, date=[2000 + int(year), 1, 1, 0])


def dump_biogem(e, outdir, year):
    from cgenie_b200.restart import write_atchem_restart, write_biogem_restart
    write_biogem_restart(e, os.path.join(outdir, "biogem_restart.nc"), member=0, year=float(year), run_id="cgenie_b200_dump")
    write_atchem_restart(e, os.path.join(outdir, "atchem_restart.nc"), member=0, year=float(year), run_id="cgenie_b200_dump")


def run_oracle(n, years, outdir):
    from cgenie_b200.series import write_series, write_series_ext
    from oracle_lib import Oracle
    cfgname, okw = CFG[n]
    o = Oracle(**okw, **_oracle_opts())
    kyear = 5 * okw["nyear"]
    e = OracleAsEnsemble(o, okw)
    if n == 2:
        o.biogem_setup(par_misc_t_runtime=float(years))
        write_series(outdir, None)                           # headers (sub_init_data_save_runtime)
        k1 = o.i("k1").reshape(J + 2, I + 2)[1:J + 1, 1:I + 1]
        sv = o.f("sv")
        A = 2.0 * np.pi * 6.37e6 ** 2 * (1.0 / I) * (sv[1:J + 1] - sv[0:J])              # phys_ocn(ipo_A,:,:,n_k), biogem_data.f90:1156
        ext = dict(ocn_tot_A=float((A[:, None] * (k1 <= okw["maxk"])).sum()), atlantic=True)
        write_series_ext(outdir, **ext)                      # fexport_*, fseaair_*, focnatm_*, misc_* headers
        o.L.cgo_biogem_sig_auto(o.h, 1, 0.0)                 # par_data_save_ben_Dmin = 0.0 (biogem-defaults.nml)
    for y in range(1, years + 1):
        o.run(kyear)
        if n == 2:
            write_series(outdir, o.f("bg_sig"), t_yr=float(y) - 0.5)      # one save window per model year (par_data_save_sig_dt = 1.0)
            write_series_ext(outdir, o.f("bg_sig"), o.f("bg_sig2"), t_yr=float(y) - 0.5, **ext)
            o.f("bg_sig")[:] = 0.0
            o.f("bg_sig2")[:] = 0.0
    dump_physics(e, outdir, years)
    if n == 2:
        dump_biogem(e, outdir, years)
    return o


def run_device(n, years, outdir):
    from cgenie_b200 import Ensemble, materialise
    from cgenie_b200.series import SeriesSaver
    cfgname, okw = CFG[n]
    with tempfile.TemporaryDirectory() as d:
        materialise(d, cfgname, overrides=dict(OPTS, **({"bg_par_misc_t_runtime": float(years)} if n == 2 else {})) or None)
        with Ensemble(d, n_members=1) as e:
            e.set_tracer_variant("strict")
            kyear = 5 * e.nyear
            if n == 1:
                e.run(kyear * years)
            else:
                gts = 3600.0 * 24.0 * 365.25 / 5.0 / e.nyear
                tick, dts = int(round(1000.0 * gts)), 10.0 * gts
                s = SeriesSaver(e, outdir, t_runtime=float(years), sig_dt=1.0, ben_Dmin=0.0, extended=True, world=okw["world"])
                for k in range(1, kyear * years + 1):            # genie.f90's loop, module by module
                    if k % 5 == 1:
                        e.surflux()
                    e.step_embm()
                    if k % 5 == 0:
                        e.step_seaice()
                        e.step_goldstein()
                    if k % 10 == 0:
                        if k == 10:
                            e.biogem_climate_sol()
                        e.biogem_forcing(k * tick)
                        e.biogem_step(dts, k * tick)
                        e.biogem_tracercoupling()
                        e.biogem_climate()
                        s.step(dts, k * tick)                    # diag_biogem_timeseries_wrapper, genie.f90:401-405
                        e.atchem_step(dts)
            dump_physics(e, outdir, years)
            if n == 2:
                dump_biogem(e, outdir, years)


RECIPE = """# Reproducing these dumps with the real cGENIE (gfortran + netCDF-Fortran required)

Configuration: BASELINE.json config #{n} = `{cfgname}` as frozen in `cgenie_b200/jobdir.py` ({years} model years from the
model's built-in initial state, no restart input).

1. Set up cGENIE as its README describes (`./setup-cgenie`; the job tool needs the `cgenie-data` repository for base configs).
2. Take any base config with the module set `{flags}` on the `{world}` 36x36x{maxk} grid{tnote} and create a user config
   from `user_config` in this directory (one `<prefix>_<key>=<value>` per line: every namelist value the frozen job uses, so
   the base config's own values do not matter).  Timestepping: {nyear} ocean steps per year (`new-job{t100} ...`;
   `tools/config_utils.py:103-162` gives `ma_genie_timestep={gts!r}`, 5:1 atmosphere:ocean steps{dbio}).
3. `./new-job -b <base> -u <user_config> dump_config{n} {years}` then `cd ~/cgenie-jobs/dump_config{n} && ./go run`.
4. Compare the model's end-of-run restart files and (config #2) `output/biogem/biogem_series_*.res` with the files in
   `oracle/` (and `device/`) using the reference's own criterion:

       python tools/nccompare.py -v <job>/output/goldstein/goldstein_restart_*.nc  oracle/goldstein_restart_{y4}_01_01.nc
       python tools/nccompare.py -v <job>/output/embm/embm_restart_*.nc            oracle/embm_restart_{y4}_01_01.nc
       python tools/nccompare.py -v <job>/output/goldsteinseaice/goldsic_restart_*.nc oracle/goldsic_restart_{y4}_01_01.nc{bgcmp}

   (`tools/nccompare.py` restates `src/tools/nccompare.f90:200-264` and `tools/tests.py:169-209`: a variable passes if its
   largest absolute difference is <= 6e-15 or its largest difference is < 35 units in the last place after casting to
   float32; the model's own `build/nccompare.exe -a 6e-15 -r 35` gives the same verdict.)

What the files hold: GOLDSTEIN `temp`, `salinity`, `uvel`, `vvel` (NF90_DOUBLE), EMBM `air_temp`, `humidity`, sea-ice
`sic_height`, `sic_cover`, `sic_temp`, `sic_albedo`{bgwhat}.  The restart date stamp in the
file names is arbitrary (ours is {y4}-01-01): compare by variable, the comparer ignores names it does not find in both files.

Status: NOT yet compared against a gfortran build by anyone (no Fortran compiler in the build container) -- parity of this
repo's oracle with the real model is unpinned until this recipe has been run.
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, choices=[1, 2], required=True)
    ap.add_argument("--years", type=int, default=None)
    ap.add_argument("--out", default="dumps")
    ap.add_argument("--device", action="store_true", help="also run the CUDA path (needs a GPU)")
    ap.add_argument("--set", action="append", default=[], metavar="go_KEY=VALUE",
                    help="a GOLDSTEIN option on top of the configuration, e.g. go_imld=1, go_iconv=1, go_ieos=1, go_iediff=1 (repeatable)")
    a = ap.parse_args()
    for kv in a.set:
        k, v = kv.split("=", 1)
        if not k.startswith("go_"):
            ap.error("--set takes GOLDSTEIN keys (go_...)")
        OPTS[k] = float(v) if "." in v or "e" in v.lower() else int(v)
    n = a.config
    years = a.years if a.years is not None else (10 if n == 1 else 2)
    cfgname, okw = CFG[n]
    base = os.path.join(a.out, "config%d" % n)
    os.makedirs(os.path.join(base, "oracle"), exist_ok=True)
    import __graft_entry__ as g
    g.build(only_missing=True)
    run_oracle(n, years, os.path.join(base, "oracle"))
    if a.device:
        os.makedirs(os.path.join(base, "device"), exist_ok=True)
        run_device(n, years, os.path.join(base, "device"))
    with open(os.path.join(base, "user_config"), "w") as f:
        f.write("\n".join(user_config(cfgname)) + "\n")
    gts = 3600.0 * 24.0 * 365.25 / 5.0 / okw["nyear"]
    with open(os.path.join(base, "RECIPE.md"), "w") as f:
        f.write(RECIPE.format(
            n=n, cfgname=cfgname, years=years, world=okw["world"], maxk=okw["maxk"], nyear=okw["nyear"], gts=gts, y4=2000 + years,
            flags="ebatmos, goldsteinocean, goldsteinseaice" + (", atchem, biogem" if n == 2 else ""),
            tnote=" with the 16 ocean / 9 sediment / 8 atmosphere tracers of `gm_*_select` in `user_config`" if n == 2 else "",
            t100=" -t100" if okw["nyear"] == 100 else "", dbio=", BIOGEM and ATCHEM every 2nd ocean step" if n == 2 else "",
            bgcmp=("\n       python tools/nccompare.py -v <job>/output/biogem/biogem_restart.nc  oracle/biogem_restart.nc   (hint: FLOAT variables)"
                   "\n       python tools/nccompare.py -v <job>/output/atchem/atchem_restart.nc  oracle/atchem_restart.nc"
                   "\n       for f in oracle/biogem_series_*.res; do python tools/nccompare.py <job>/output/biogem/$(basename $f) $f; done") if n == 2 else "",
            bgwhat=(", BIOGEM `ocn_*` / `bio_part_*` and ATCHEM `atm_*` (FLOAT, as the reference stores them), and the yearly lines of "
                    "`biogem_series_ocn_*.res` / `biogem_series_atm_*.res` (save interval 1 yr, `bg_par_data_save_sig_dt=1.0`, "
                    "`bg_par_data_save_ben_Dmin=0.0`)") if n == 2 else ""))
    print("wrote", base)


if __name__ == "__main__":
    main()
