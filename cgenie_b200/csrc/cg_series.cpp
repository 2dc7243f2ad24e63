// cg_series.cpp -- BIOGEM's ASCII time series (.res) of the ocean and atmosphere tracers, host code only.
//
// Follows sub_init_data_save_runtime (src/biogem/biogem_data_ascii.f90:23-110) for the files and header lines and
// sub_data_save_runtime (:669-935) for the data lines: what is printed (inventory, mean, ice-free surface mean, benthic mean;
// T in degrees C; isotopes as delta values) and in which edit descriptors.  The numbers come from the window integrals the
// device accumulates (cg_biogem_sig_update, field "bg_sig").  File names: fun_data_timeseries_filename
// (biogem_lib.f90:1618-1652) = <outdir>/<outfile_name>_series_<name><ext>, ext = string_results_ext = '.res'.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/cgenie_b200.h"

namespace {
constexpr double kNullSmall = 0.999999e-19;   // const_real_nullsmall, gem_cmn.f90:719
constexpr double kNullIso = -999.999;         // const_nulliso, gem_cmn.f90:722
constexpr double kZeroC = 273.15;             // const_zeroC, gem_cmn.f90:690
constexpr double kAtmMol = 1.7692e+020;       // conv_atm_mol, gem_cmn.f90:509
constexpr double kStd13C = 0.011202, kStd14C = 1.176e-12;   // const_standards(11:12), gem_cmn.f90:629-631

std::string g_series_err;
int sfail(const std::string &m) { g_series_err = m; return CG_ERR_IO; }

// a field wider than w loses the optional zero before the decimal point first, then becomes asterisks (Fortran 2008 10.7.2)
std::string fit(std::string s, int w) {
  if ((int)s.size() > w) {
    if (s.compare(0, 2, "0.") == 0) s.erase(0, 1);
    else if (s.compare(0, 3, "-0.") == 0) s.erase(1, 1);
  }
  if ((int)s.size() > w) return std::string(w, '*');
  return std::string(w - s.size(), ' ') + s;
}
// Fw.d
std::string fmt_f(double x, int w, int d) {
  char b[512];
  std::snprintf(b, sizeof b, "%.*f", d, x);
  return fit(b, w);
}
// Ew.d: 0.ddddddE+ee (mantissa in [0.1, 1)); a three-digit exponent drops the 'E' as gfortran does
std::string fmt_e(double x, int w, int d) {
  std::string s;
  if (!std::isfinite(x)) s = std::isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity");
  else if (x == 0.0) s = std::string(std::signbit(x) ? "-0." : "0.") + std::string(d, '0') + "E+00";
  else {
    char b[64];
    std::snprintf(b, sizeof b, "%.*e", d - 1, std::fabs(x));      // D.ddddde[+-]XX, correctly rounded to d digits
    const char *e = std::strchr(b, 'e');
    const int ex = std::atoi(e + 1) + 1;
    std::string digits(1, b[0]);
    digits.append(b + 2, (size_t)(e - (b + 2)));
    char eb[16];
    if (std::abs(ex) < 100) std::snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
    else std::snprintf(eb, sizeof eb, "%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
    s = std::string(x < 0 ? "-0." : "0.") + digits + eb;
  }
  return fit(s, w);
}
// fun_calc_isotope_delta(tot, iso, standard, .FALSE., const_nulliso), gem_util.f90:568-598
double iso_delta(double tot, double iso, double standard) {
  if (tot > kNullSmall) {
    const double f = iso / tot;
    if ((1.0 - f) > kNullSmall) {
      const double R = f / (1.0 - f);
      return 1000.0 * (R / standard - 1.0);
    }
  }
  return kNullIso;
}
int put_line(const std::string &path, const char *mode, const std::string &line) {
  FILE *fp = std::fopen(path.c_str(), mode);
  if (!fp) return sfail("cannot open " + path);
  const bool ok = std::fputs(line.c_str(), fp) >= 0 && std::fputc('\n', fp) != EOF;
  return (std::fclose(fp) == 0 && ok) ? CG_OK : sfail("short write to " + path);
}
}  // namespace

extern "C" const char *cg_series_last_error(void) { return g_series_err.c_str(); }

extern "C" int cg_biogem_series_write(const char *outdir, const char *outfile_name, int create, double t_yr, int n_ocn,
                                      const char *const *ocn_names, const int32_t *ocn_type, const int32_t *ocn_dep, int n_atm,
                                      const char *const *atm_names, const int32_t *atm_type, const int32_t *atm_dep,
                                      const double *sig, int with_sur) {
  if (!outdir || !outfile_name || n_ocn < 0 || n_atm < 0 || (n_ocn && (!ocn_names || !ocn_type || !ocn_dep)) ||
      (n_atm && (!atm_names || !atm_type || !atm_dep)) || (!create && !sig))
    return sfail("cg_biogem_series_write: bad argument");
  std::string base = outdir;
  if (!base.empty() && base.back() != '/') base += '/';            // par_outdir_name = trim(par_outdir_name)//'/', biogem_data.f90:53
  base += std::string(outfile_name) + "_series_";
  const int L = n_ocn;
  // the "bg_sig" layout: int_t_sig, tot_M, tot_M_sur, ocn(L), ocn_sur(L), ocn_ben(L), ocnatm(LA)
  const double *S_ocn = sig ? sig + 3 : nullptr, *S_sur = sig ? sig + 3 + L : nullptr, *S_ben = sig ? sig + 3 + 2 * L : nullptr,
               *S_atm = sig ? sig + 3 + 3 * L : nullptr;
  const double t_sig = (sig && !create) ? sig[0] : 1.0;
  if (!create && !(t_sig > kNullSmall)) return CG_OK;               // biogem.f90:3119: nothing integrated, nothing saved
  const double tot_M = (sig && !create) ? sig[1] / t_sig : 0.0;     // loc_ocn_tot_M, biogem_data_ascii.f90:691
  for (int l = 0; l < n_ocn; l++) {
    const std::string n = ocn_names[l], path = base + "ocn_" + n + ".res";
    const int ty = ocn_type[l];
    if (!(ty == 0 || ty == 1 || (ty >= 11 && ty <= 12))) continue;
    if (create) {
      std::string h;
      if (ty == 0) {
        if (l == 0) h = with_sur ? "% time (yr) / temperature (C) / _surT (C) / _benT (degrees C)" : "% time (yr) / temperature (degrees C)";
        else h = with_sur ? "% time (yr) / salinity (o/oo) / _surS (o/oo) / _benS (o/oo)" : "% time (yr) / salinity (o/oo)";
      } else {
        const std::string u = ty == 1 ? " (mol kg-1)" : " (o/oo)";
        h = "% time (yr) / global " + n + " (mol) / global " + n + u;
        if (with_sur) h += " / surface " + n + u + " / benthic " + n + u;
      }
      if (int rc = put_line(path, "w", " " + h)) return rc;         // list-directed output starts with a blank
      continue;
    }
    std::string line = fmt_f(t_yr, 12, 3);
    if (ty == 0) {
      const double off = l == 0 ? kZeroC : 0.0;                     // io_T is the first selected tracer
      line += fmt_f(l == 0 ? S_ocn[l] / t_sig - off : S_ocn[l] / t_sig, 12, 6);
      if (with_sur)
        line += fmt_f(l == 0 ? S_sur[l] / t_sig - off : S_sur[l] / t_sig, 12, 6) + fmt_f(l == 0 ? S_ben[l] / t_sig - off : S_ben[l] / t_sig, 12, 6);
    } else if (ty == 1) {
      const double v = S_ocn[l] / t_sig;
      line += fmt_e(tot_M * v, 15, 7) + fmt_e(v, 15, 7);
      if (with_sur) line += fmt_e(S_sur[l] / t_sig, 15, 7) + fmt_e(S_ben[l] / t_sig, 15, 7);
    } else {
      const int d = ocn_dep[l];
      if (d < 0 || d >= n_ocn) return sfail("cg_biogem_series_write: isotope without its bulk tracer");
      const double st = ty == 11 ? kStd13C : kStd14C;
      const double frac = S_ocn[l] / t_sig;
      line += fmt_e(tot_M * frac, 15, 7) + fmt_f(iso_delta(S_ocn[d] / t_sig, frac, st), 12, 3);
      if (with_sur)
        line += fmt_f(iso_delta(S_sur[d] / t_sig, S_sur[l] / t_sig, st), 12, 3) + fmt_f(iso_delta(S_ben[d] / t_sig, S_ben[l] / t_sig, st), 12, 3);
    }
    if (int rc = put_line(path, "a", line)) return rc;
  }
  for (int l = 0; l < n_atm; l++) {
    const std::string n = atm_names[l], path = base + "atm_" + n + ".res";
    const int ty = atm_type[l];
    if (!(ty == 0 || ty == 1 || (ty >= 11 && ty <= 12))) continue;
    if (create) {
      std::string h;
      if (ty == 0) h = l == 0 ? "% time (yr) / surface air temperature (degrees C)" : "% time (yr) / surface humidity (?" "?" "?)";
      else h = "% time (yr) / global " + n + " (mol) / global " + n + (ty == 1 ? " (atm)" : " (o/oo)");
      if (int rc = put_line(path, "w", " " + h)) return rc;
      continue;
    }
    std::string line = fmt_f(t_yr, 12, 3);
    const double v = S_atm[l] / t_sig;
    if (ty == 0) line += fmt_f(v, 12, 6);
    else if (ty == 1) line += fmt_e(kAtmMol * v, 15, 7) + fmt_e(v, 15, 7);
    else {
      const int d = atm_dep[l];
      if (d < 0 || d >= n_atm) return sfail("cg_biogem_series_write: isotope without its bulk tracer");
      line += fmt_e(kAtmMol * v, 15, 7) + fmt_f(iso_delta(S_atm[d] / t_sig, v, ty == 11 ? kStd13C : kStd14C), 14, 3);
    }
    if (int rc = put_line(path, "a", line)) return rc;
  }
  return CG_OK;
}

namespace {
// fun_calc_isotope_delta with dum_allow_negative = .TRUE. (fluxes: a negative total is a flux out), gem_util.f90:568-598
double iso_delta_signed(double tot, double iso, double standard) {
  if (std::fabs(tot) > kNullSmall) {
    const double f = iso / tot;
    if ((1.0 - f) > kNullSmall) {
      const double R = f / (1.0 - f);
      return 1000.0 * (R / standard - 1.0);
    }
  }
  return kNullIso;
}
// fun_convert_delta14CtoD14C, gem_util.f90:623-637 (Stuiver and Polach 1977)
double delta14C_to_D14C(double d13C, double d14C) {
  return 1000.0 * ((1.0 + d14C / 1000.0) * (0.975 * 0.975) / ((1.0 + d13C / 1000.0) * (1.0 + d13C / 1000.0)) - 1.0);
}
}  // namespace

// The fexport_*, fseaair_*, focnatm_* and misc_seaice / misc_opsi / misc_atm_D14C / misc_SLT series of sub_init_data_save_runtime
// (biogem_data_ascii.f90:107-197, 320-400) and sub_data_save_runtime (:955-1096, 1245-1340), from the integrals "bg_sig" (int_t_sig,
// the atmosphere rows) and "bg_sig2" (cg_biogem_sig_extended).  sed_type: 1 bio ... 7 scavenged (flux + density), 8 age,
// 11 / 12 isotopes, 9 (frac2) writes nothing; ocn_tot_A = SUM(phys_ocn(ipo_A,:,:,n_k)); opsi_scale = goldstein_dsc * goldstein_usc *
// const_rEarth * 1.0E-6; atlantic != 0 for the topographies whose file carries the Atlantic columns (worbe2, worjh2 ...: :1282-1293).
extern "C" int cg_biogem_series_write_ext(const char *outdir, const char *outfile_name, int create, double t_yr, int n_ocn, int n_sed,
                                          const char *const *sed_names, const int32_t *sed_type, const int32_t *sed_dep, int n_atm,
                                          const char *const *atm_names, const int32_t *atm_type, const int32_t *atm_dep,
                                          const double *sig, const double *sig2, double ocn_tot_A, double opsi_scale, int atlantic) {
  if (!outdir || !outfile_name || n_sed < 0 || n_atm < 0 || n_ocn < 0 || (n_sed && (!sed_names || !sed_type || !sed_dep)) ||
      (n_atm && (!atm_names || !atm_type || !atm_dep)) || (!create && (!sig || !sig2)))
    return sfail("cg_biogem_series_write_ext: bad argument");
  std::string base = outdir;
  if (!base.empty() && base.back() != '/') base += '/';
  base += std::string(outfile_name) + "_series_";
  const double t_sig = (sig && !create) ? sig[0] : 1.0;
  if (!create && !(t_sig > kNullSmall)) return CG_OK;
  const double *S_atm = sig ? sig + 3 + 3 * n_ocn : nullptr;
  const double *X = sig2, *F_exp = sig2 ? sig2 + 8 : nullptr, *F_oa = sig2 ? sig2 + 8 + n_sed : nullptr,
               *F_as = sig2 ? sig2 + 8 + n_sed + n_atm : nullptr;
  // ---- fexport
  for (int l = 0; l < n_sed; l++) {
    const std::string n = sed_names[l], path = base + "fexport_" + n + ".res";
    const int ty = sed_type[l];
    const bool bulk = ty >= 1 && ty <= 7, age = ty == 8, iso = ty >= 11 && ty <= 20;
    if (!(bulk || age || iso)) continue;
    if (create) {
      std::string h;
      if (bulk) h = "% time (yr) / global " + n + " flux (mol yr-1) / global " + n + " density (mol m-2 yr-1)";
      else if (age) h = "% time (yr) / CaCO3 age (yr)";
      else h = "% time (yr) / global " + n + " flux (mol yr-1) / global " + n + " delta (o/oo)";
      if (int rc = put_line(path, "w", " " + h)) return rc;
      continue;
    }
    std::string line = fmt_f(t_yr, 12, 3);
    if (bulk) {
      const double v = F_exp[l] / t_sig;
      line += fmt_e(v, 15, 7) + fmt_e(v / ocn_tot_A, 15, 7);
    } else if (age) {
      const int d = sed_dep[l];
      line += fmt_e((d >= 0 && d < n_sed && F_exp[d] > kNullSmall) ? F_exp[l] / t_sig : 0.0, 15, 7);
    } else {
      const int d = sed_dep[l];
      if (d < 0 || d >= n_sed) return sfail("cg_biogem_series_write_ext: isotope without its bulk tracer");
      const double frac = F_exp[l] / t_sig;
      line += fmt_e(frac, 15, 7) + fmt_f(iso_delta(F_exp[d] / t_sig, frac, ty == 11 ? kStd13C : kStd14C), 14, 3);
    }
    if (int rc = put_line(path, "a", line)) return rc;
  }
  // ---- fseaair (int_diag_airsea_sig) and focnatm, the gases only (l = 3 .. n_l_atm)
  for (int pass = 0; pass < 2; pass++) {
    const double *F = pass == 0 ? F_as : F_oa;
    for (int l = 2; l < n_atm; l++) {
      const std::string n = atm_names[l], path = base + (pass == 0 ? "fseaair_" : "focnatm_") + n + ".res";
      const int ty = atm_type[l];
      const bool bulk = ty == 1, iso = ty >= 11 && ty <= 20;
      if (!(bulk || iso)) continue;
      if (create) {
        std::string h;
        if (pass == 0) {
          if (bulk) h = "% time (yr) / global " + n + " sea->air transfer flux (mol yr-1) / global " + n + " density (mol m-2 yr-1)";
          else h = "% time (yr) / global " + n + " sea->air transfer flux (mol yr-1) / global " + n + " (o/oo)";
        } else {
          if (bulk) h = "% time (yr) / global " + n + " flux (mol yr-1) / global " + n + " density (mol m-2 yr-1) ";
          else h = "% time (yr) / global " + n + " flux (mol yr-1) / global " + n + " (o/oo)";
          h += " NOTE: is the atmospheric forcing flux *net* of the sea-air gas exchange flux.";
        }
        if (int rc = put_line(path, "w", " " + h)) return rc;
        continue;
      }
      std::string line = fmt_f(t_yr, 12, 3);
      if (bulk) {
        const double v = F[l] / t_sig;
        line += fmt_e(v, 15, 7) + fmt_f(v / ocn_tot_A, 12, 3);
      } else {
        const int d = atm_dep[l];
        if (d < 0 || d >= n_atm) return sfail("cg_biogem_series_write_ext: isotope without its bulk tracer");
        const double frac = F[l] / t_sig;
        line += fmt_e(frac, 15, 7) + fmt_f(iso_delta_signed(F[d] / t_sig, frac, ty == 11 ? kStd13C : kStd14C), 14, 3);
      }
      if (int rc = put_line(path, "a", line)) return rc;
    }
  }
  // ---- misc
  int l13 = -1, l14 = -1;
  for (int l = 0; l < n_atm; l++) { if (atm_type[l] == 11 && l13 < 0) l13 = l; if (atm_type[l] == 12 && l14 < 0) l14 = l; }
  if (create) {
    if (int rc = put_line(base + "misc_seaice.res", "w", " % time (yr) / global sea-ice area (m2) / mean sea-ice cover (%) / global sea-ice volume (m3) / mean sea-ice thickness (m)")) return rc;
    if (int rc = put_line(base + "misc_opsi.res", "w", atlantic ? " % time (yr) / global min overturning (Sv) / global max overturning (Sv) / Atlantic min overturning (Sv) / Atlantic max overturning (Sv)"
                                                                 : " % time (yr) / global min overturning (Sv) / global max overturning (Sv)")) return rc;
    if (l14 >= 0) if (int rc = put_line(base + "misc_atm_D14C.res", "w", " % time (yr) / mean isotopic composition (o/oo)")) return rc;
    if (int rc = put_line(base + "misc_SLT.res", "w", " % time (yr) / mean (land) surface air temperature (degrees C)")) return rc;
    return CG_OK;
  }
  {
    std::string line = fmt_f(t_yr, 12, 3) + fmt_e(X[0] / t_sig, 12, 4) + fmt_f(100.0 * (1.0 / ocn_tot_A) * X[0] / t_sig, 9, 3) +
                       fmt_e(X[2] / t_sig, 12, 4) + fmt_f(X[1] / t_sig, 9, 3);
    if (int rc = put_line(base + "misc_seaice.res", "a", line)) return rc;
    line = fmt_f(t_yr, 12, 3) + fmt_f(opsi_scale * X[3] / t_sig, 9, 3) + fmt_f(opsi_scale * X[4] / t_sig, 9, 3);
    if (atlantic) line += fmt_f(opsi_scale * X[5] / t_sig, 9, 3) + fmt_f(opsi_scale * X[6] / t_sig, 9, 3);
    if (int rc = put_line(base + "misc_opsi.res", "a", line)) return rc;
    if (l14 >= 0 && l13 >= 0) {
      const int d = atm_dep[l14];
      const double tot = S_atm[d] / t_sig;
      const double d13 = iso_delta(tot, S_atm[l13] / t_sig, kStd13C), d14 = iso_delta(tot, S_atm[l14] / t_sig, kStd14C);
      if (int rc = put_line(base + "misc_atm_D14C.res", "a", fmt_f(t_yr, 12, 3) + fmt_f(delta14C_to_D14C(d13, d14), 12, 3))) return rc;
    }
    if (int rc = put_line(base + "misc_SLT.res", "a", fmt_f(t_yr, 12, 3) + fmt_f(X[7] / t_sig, 12, 6))) return rc;
  }
  return CG_OK;
}
