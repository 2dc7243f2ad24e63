// cg_host.hpp -- host side of the B200 hot path: namelist/data-file readers and the
// constants the reference's initialise_* routines build (scalar, I/O-driven, bit-exact).
// Product code: nothing here touches oracle/.
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace cg {

// ---- dimensional scales (goldstein_lib.f90:48-96, embm_lib.f90:36-124) ----
constexpr double kUsc = 0.05, kRsc = 6.37e6, kDsc = 5.0e3, kFsc = 2 * 7.2921e-5, kGsc = 9.81, kRh0sc = 1.0e3;
constexpr double kRhosc = kRh0sc * kFsc * kUsc * kRsc / kGsc / kDsc;
constexpr double kTsc = kRsc / kUsc, kCpsc = 3981.1, kRhoair = 1.25, kRho0 = 1.0e3, kRhoao = kRhoair / kRho0;
constexpr double kM2mm = 1000.0, kMm2m = 1.0 / kM2mm;
constexpr double kRfluxsc = kRsc / (kDsc * kUsc * kRh0sc * kCpsc);
constexpr double kCpa = 1004.0, kConst1 = 3.80e-3, kConst2 = 21.87, kConst3 = 265.5, kConst4 = 17.67, kConst5 = 243.5;
constexpr double kSigma = 5.67e-8, kEmo = 0.94 * kSigma, kEma = 0.85 * kSigma, kTfreez = 0.0;
constexpr double kHlv = 2.501e6, kHlf = 3.34e5, kHls = kHlv + kHlf, kConsic = 2.166, kZeroc = 273.15;
constexpr double kCpoIce = 4044.0, kRhoice = 913.0, kHmin = 0.01, kRhmin = 1.0 / kHmin;
constexpr double kRhooi = kRho0 / kRhoice, kRhoio = kRhoice / kRho0, kRrholf = 1.0 / (kRhoice * kHlf);
constexpr double kCo20 = 278.0e-6, kCh40 = 700.0e-9, kN2o0 = 275.0e-9, kAlphaCh4 = 0.036, kAlphaN2o = 0.12;
constexpr double kTsic = -1.8, kCd = 0.0013;

// ---- Fortran namelist group reader ("key=value," lines between &GROUP and &END) ----
class Namelist {
 public:
  bool load(const std::string &path, std::string *err);
  bool has(const std::string &key) const;
  double num(const std::string &key, double dflt) const;
  int integer(const std::string &key, int dflt) const;
  bool flag(const std::string &key, bool dflt) const;
  std::string str(const std::string &key, const std::string &dflt) const;
 private:
  std::map<std::string, std::string> kv_;
};

// whitespace-separated numbers of a free-format ASCII file
bool read_numbers(const std::string &path, std::vector<double> *out, std::string *err);

// ---- scalar parameters of one ensemble member (namelist values, pre-scaling) ----
struct Params {
  // data_genie
  int maxi = 36, maxj = 36, maxk = 8, maxl = 2;
  int kocn_loop = 5, katm_loop = 1, ksic_loop = 5, conv_kocn_kbiogem = 2, conv_kocn_katchem = 2;
  bool flag_biogem = false, flag_atchem = false;
  double solconst = 1368.0, gn_daysperyear = 365.25, genie_timestep = 3600.0;
  // data_GOLD
  int igrid = 0, nyear = 100;
  double yearlen = 365.25, temp0 = 5.0, temp1 = 5.0, rel = 0.9, scf = 2.0, diff1 = 2000.0, diff2 = 1.0e-5,
         adrag = 2.5, hosing = 0.0, hosing_trend = 0.0, albocn = 0.05, ssmaxsurf = 10.0, ssmaxdeep = 10.0,
         saln0 = 34.9, ediff0 = 0.0, ediffpow1 = 1.0, ediffpow2 = 1.0, ediffvar = 0.0,
         mldpebuoycoeff = 0.15, mldketaucoeff = 2.5, mldwindkedec = 25.0;   // imld = 1 (goldstein-defaults.nml:31-34)
  int nyears_hosing = 0, iconv = 0, imld = 0, iediff = 0, ieos = 0;
  bool diso = true;
  std::string world = "worbe2", go_indir = "input/goldstein";
  // data_EMBM
  int ndta = 5, diffa_len = 0, albedop_skewp = 0, par_wind_polar_avg = 0;
  double rmax = 0.85, diffamp1 = 5.0e6, diffamp2 = 1.0e6, diffwid = 1.0, difflin = 0.1, betaz1 = 0.0, betaz2 = 0.4,
         betam1 = 0.0, betam2 = 0.4, tatm = 10.0, relh0_ocean = 0.0, relh0_land = 0.0, extra1a = -0.03,
         extra1b = 0.17, extra1c = 0.18, scl_fwf = 1.0, z1_embm = 10.0, diffa_scl = 1.0, delf2x = 5.77,
         olr_adj0 = 0.0, olr_adj = 0.0, t_eqm = 12.371, albedop_offs = 0.20, albedop_amp = 0.36,
         albedop_skew = 0.0, albedop_mod2 = 0.0, albedop_mod4 = 0.0, albedop_mod6 = 0.0, par_sich_max = 9999.9,
         par_albsic_min = 0.2, par_albsic_max = 0.7, radfor_scl_co2 = 1.0, radfor_pc_co2_rise = 0.0,
         radfor_scl_ch4 = 1.0, radfor_pc_ch4_rise = 0.0, radfor_scl_n2o = 1.0, radfor_pc_n2o_rise = 0.0;
  bool atchem_radfor = false;
  std::string eb_indir = "input/embm", xu_wstress = "taux_u.interp", yu_wstress = "tauy_u.interp",
              xv_wstress = "taux_v.interp", yv_wstress = "tauy_v.interp", u_wspeed = "uncep.silo",
              v_wspeed = "vncep.silo";
  // data_goldSIC
  double diffsic = 2000.0, par_sica_thresh = 1.0, par_sich_thresh = 1000.0;
  // data_BIOGEM: the perturbable subset; everything else lives in BgConfig
  double par_bio_k0_PO4 = 2.0E-06, par_bio_remin_POC_eL1 = 500.0, par_bio_red_POC_CaCO3 = 0.2;
  bool set(const std::string &name, double v);  // per-member override by name
};

// ---- member-independent grid + masks (goldstein.f90:906-1060, 1104-1137, 1361-1391, 1464-1494) ----
struct Grid {
  int I = 0, J = 0, K = 0, L = 0, nyear = 0, igrid = 0;
  double dphi = 0, rdphi = 0, dzz = 0, dt = 0;  // dt(k) is uniform (goldstein.f90:974-977)
  // 1-D metrics, indexed with the Fortran index (over-allocated)
  std::vector<double> ds, dsv, rds2, dz, s, c, sv, cv, dza, zro, zw, rc, rc2, rcv, rdsv, cv2, rds, rdz, rdza, asurf;
  std::vector<int> k1;    // (0:I+1, 0:J+1)
  std::vector<int> ku;    // (2, I, J)
  std::vector<int> mk;    // (I+1, J)
  std::vector<int> getj;  // (I, J)
  std::vector<int> ips, ipf, ias, iaf;
  int jsf = 1, ntot = 0, intot = 0;
  std::vector<double> rh;  // (3, 0:I+1, 0:J+1)
  int k1at(int i, int j) const { return k1[i + (I + 2) * j]; }
  double rhat(int l, int i, int j) const { return rh[(l - 1) + 3 * (i + (I + 2) * j)]; }
  // builds everything above from dims + k1 (file order rows j=J+1..0)
  void build(int I_, int J_, int K_, int L_, int igrid_, int nyear_, double yearlen, const std::vector<int> &k1file);
};

// island geometry (goldstein.f90:1502-1577)
struct Islands {
  int isles = 0, mpi = 0;
  std::vector<double> psiles;  // gbold(i + j*I) landmass ids, 1-based index
  std::vector<int> npi, lpisl, ipisl, jpisl;
};

// ---- constants that depend on per-member parameters ----
struct MemberConsts {
  Params p;
  double diff1 = 0, diff2 = 0, adrag = 0;  // non-dimensional (goldstein.f90:1400-1407)
  double ec[6] = {0, 0, 0, 0, 0, 0}, rpmesco = 0, rsictscsf = 0;
  double hosing = 0, hosing_trend = 0;
  int nsteps_hosing = 0;
  std::vector<double> ssmax;                  // (K-1)
  // stratification-dependent vertical diffusivity, iediff = 1 | 2 (SUBROUTINE ediff, goldstein.f90:2936-3044, ediffvar = 0)
  double ediff0 = 0.0;                        // non-dimensional
  int ediffpow2i = 0;
  std::vector<double> ediff1p, diffmax;       // (K-1), (K)
  std::vector<double> drag, rtv, rtv3;        // (2,I+1,J), (I,J), (I,J)
  std::vector<double> gap, ratm;              // (nm, 2I+3), (nm, I+1)
  std::vector<double> ubisl, psisl, erisl;    // island unit solves
  std::vector<double> rhosing;                // (I,J)
  // EMBM
  double dtatm = 0, rdtdim = 0, ryear = 0, rfluxsca = 0, rpmesca = 0, ppmin = 0, ppmax = 0, hatmbl1 = 8400.0,
         hatmbl2 = 1800.0, rate_co2 = 0, rate_ch4 = 0, rate_n2o = 0, extra1a = 0, extra1b = 0, extra1c = 0;
  std::vector<double> diffa;   // (2,2,J)
  std::vector<double> albcl, ca, pmeadj, uatm, us_dztau, us_dztav, solfor, lowestlu2, lowestlv3;
  std::vector<double> tau, dztau, dztav, usurf;  // from stresses * scf (goldstein.f90:102-107, embm.f90:2762-2816)
  std::vector<double> mldketau;                  // (I,J) wind energy input of the mixed-layer scheme, imld = 1 (goldstein.f90:112-141)
  std::vector<double> mlddec, mlddecd;           // (K) its decay with depth (goldstein.f90:1675-1686)
  std::vector<int> iroff, jroff;
  // sea ice
  double dtsic = 0, sic_rdtdim = 0, diffsic = 0;
  // initial state
  std::vector<double> ts0;   // (L, I, J, K) interior, Fortran order
  std::vector<double> rho0;  // (I, J, K)
  std::vector<double> tq0;   // (2, I, J)
};

struct WindFiles { std::vector<double> taux_u, tauy_u, taux_v, tauy_v, uncep, vncep; };

// equation of state, goldstein.f90:3048-3061: ieos == 0 (ec[5] == 0), and ieos == 1 with the thermobaricity term ec(5) * t * z
inline double eos(const double *ec, double t, double s) { return ec[1] * t + ec[2] * s + ec[3] * (t * t) + ec[4] * (t * t * t); }
inline double eos_z(const double *ec, int ieos, double t, double s, double z) {
  if (ieos == 0) return eos(ec, t, s);
  return ec[1] * t + ec[2] * s + ec[3] * (t * t) + ec[4] * (t * t * t) + ec[5] * t * z;
}

// builds the per-member constants; `shared_baro`, when non-null and the drag parameters match,
// lets members reuse one barotropic factorisation.
void build_member(const Grid &g, const Islands &isl, const WindFiles &w, const Params &p, MemberConsts *mc,
                  const MemberConsts *shared_baro);

bool load_job(const std::string &jobdir, Params *p, Grid *g, Islands *isl, WindFiles *w, std::string *err);

}  // namespace cg
