// cg_biogem.hpp -- host side of the BIOGEM / ATCHEM hot path: configuration (data_BIOGEM, data_GEM, data_ATCHEM,
// forcing directory), tracer-relationship tables and the scalar forcing update, restated from
//   src/biogem/biogem_data.f90  sub_load_goin_biogem :17-460, sub_init_bio :579-624,
//                               sub_data_update_tracerrelationships :731-920, sub_init_force_restore_atm :2706-2791
//   src/biogem/biogem_box.f90   sub_update_sig :3174-3218, sub_update_force_restore_atm :3333-3366
//   src/biogem/biogem_lib.f90   sub_load_data_t2 :1421-1481
//   src/common/gem_util.f90     sub_def_tracerrelationships :27-274, sub_def_tracer_decay :280-308
//   src/common/gem_data.f90     Schmidt / Bunsen tables :69-136
// Only the tracer selection of the frozen eb_go_gs_ac_bg configuration is accepted; anything else is
// refused at cg_create with CG_ERR_CONFIG (no silent approximation).
#pragma once
#include <string>
#include <vector>

#include "cg_host.hpp"

namespace cg {

struct BgDev;

constexpr double kBgZeroC = 273.15;            // gem_cmn.f90:690
constexpr double kBgNull = -0.999999e19;       // gem_cmn.f90:717
constexpr double kBgNullSmall = 0.999999e-19;  // gem_cmn.f90:719
constexpr double kBgYrS = 3600.0 * (24.0 * 365.25);  // gem_cmn.f90:511-513
constexpr double kBgStd13C = 0.011202, kBgStd14C = 1.176e-12;  // gem_cmn.f90:630-631
constexpr double kBgLambda14C = 1. / 8267.0;   // gem_cmn.f90:648
constexpr double kBgM3Kg = 1027.649;           // gem_cmn.f90:510
constexpr double kBgPi = 3.141592653589793, kBgREarth = 6.37e6;  // gem_cmn.f90:688,696

struct BgConfig {
  bool on = false;
  int L = 0, LS = 0, LA = 0;
  std::vector<int> io, otype, odep;            // 1-based (index 0 unused)
  std::vector<int> is, stype, sdep_id, sdep_ls;
  std::vector<int> ia, atype, adep;
  std::vector<double> ocn_init, atm_init;
  // biogem-defaults.nml values used on the path
  double t_runtime = 1001.0, t_end = 1001.0, c0_PO4 = 0.050E-06, red_POP_POC = 106.0, red_POP_PON = 16.0,
         red_POP_PO2 = -138.0, red_PON_ALK = -1.00, red_DOMfrac = 0.66, red_RDOMfrac = 0.0, red_POC_CaCO3_pP = 0.0,
         DOMlifetime = 0.5, POC_frac2 = 0.05, POC_eL2 = 1000000.0, POC_dfrac2 = 0.0, POC_c0frac2 = 0.1E-6,
         CaCO3_frac2 = 0.5, CaCO3_eL1 = 1000.0, CaCO3_eL2 = 1000000.0, sinkingrate_md = 125.0, remin_k_O2 = 1.0,
         remin_c0_O2 = 8.0E-6, gastransfer_a = 0.310, d13C_DIC_Corg_ef = 25.0, Fgeothermal = 0.0;
  std::vector<double> windspeed;               // (I,J), i fastest
  // atmospheric restoring forcing (uniform fields, piecewise-linear signal)
  std::vector<int> rst_sel;
  std::vector<double> rst_tconst;
  std::vector<std::vector<double>> rst_sig_t, rst_sig_v;   // per la: signal points after sub_load_data_t2
  std::vector<int> rst_sig_i1, rst_sig_i2;                 // force_restore_atm_sig_i (1-based, mutable)
  std::vector<double> rst_target;                          // force_restore_atm at wet points, by la
};

// Parse the BIOGEM/ATCHEM part of a job directory; *cfg.on stays false when flag_biogem is off.
bool load_biogem(const std::string &jobdir, const Params &p, const Grid &g, BgConfig *cfg, std::string *err);

// Tables and step-independent parameters of the device view.
void bg_fill_tables(const BgConfig &c, const Params &p, const Grid &g, BgDev *b);

// biogem_forcing at genie_clock (ms): advances the signal indices and sets rst_target / rst_active in *b.
void bg_forcing(BgConfig *c, long long clock_ms, BgDev *b);

// fun_calc_isotope_fraction, gem_util.f90:604-617
inline double bg_iso_fraction(double delta, double standard) {
  const double R = standard * (1.0 + delta / 1000.0);
  return R / (1.0 + R);
}

}  // namespace cg
