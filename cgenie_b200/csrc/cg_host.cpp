// cg_host.cpp -- namelist / data-file readers and the bit-exact host restatement of the
// constants built by initialise_goldstein (goldstein.f90:514-2084), initialise_embm
// (embm.f90:198-2018) and initialise_seaice (gold_seaice.f90:17-508).
// Expression order follows the reference line by line (IEEE fp64, no contraction).
#include "cg_host.hpp"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace cg {

// ------------------------------------------------------------------ readers
static std::string lower(std::string s) {
  for (auto &ch : s) ch = (char)std::tolower((unsigned char)ch);
  return s;
}
static std::string trim(const std::string &s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) a++;
  while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

bool Namelist::load(const std::string &path, std::string *err) {
  std::ifstream f(path);
  if (!f) {
    if (err) *err = "could not open namelist file " + path;
    return false;
  }
  std::string line;
  while (std::getline(f, line)) {
    // strip comments (a '!' outside quotes)
    bool inq = false;
    size_t cut = std::string::npos;
    for (size_t i = 0; i < line.size(); i++) {
      if (line[i] == '"' || line[i] == '\'') inq = !inq;
      if (line[i] == '!' && !inq) { cut = i; break; }
    }
    if (cut != std::string::npos) line = line.substr(0, cut);
    line = trim(line);
    if (line.empty() || line[0] == '&' || line[0] == '/') continue;
    size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    std::string key = lower(trim(line.substr(0, eq)));
    std::string val = trim(line.substr(eq + 1));
    while (!val.empty() && val.back() == ',') val = trim(val.substr(0, val.size() - 1));
    if (val.size() >= 2 && (val.front() == '"' || val.front() == '\'') && val.back() == val.front())
      val = val.substr(1, val.size() - 2);
    kv_[key] = val;
  }
  return true;
}
bool Namelist::has(const std::string &key) const { return kv_.count(lower(key)) != 0; }
double Namelist::num(const std::string &key, double dflt) const {
  auto it = kv_.find(lower(key));
  if (it == kv_.end()) return dflt;
  std::string v = it->second;
  for (auto &ch : v)
    if (ch == 'd' || ch == 'D') ch = 'e';  // Fortran double exponent
  return std::strtod(v.c_str(), nullptr);
}
int Namelist::integer(const std::string &key, int dflt) const {
  auto it = kv_.find(lower(key));
  if (it == kv_.end()) return dflt;
  return (int)std::strtol(it->second.c_str(), nullptr, 10);
}
bool Namelist::flag(const std::string &key, bool dflt) const {
  auto it = kv_.find(lower(key));
  if (it == kv_.end()) return dflt;
  std::string v = lower(it->second);
  if (v.find('t') != std::string::npos || v == "y") return true;
  return false;
}
std::string Namelist::str(const std::string &key, const std::string &dflt) const {
  auto it = kv_.find(lower(key));
  return it == kv_.end() ? dflt : trim(it->second);
}

bool read_numbers(const std::string &path, std::vector<double> *out, std::string *err) {
  std::ifstream f(path);
  if (!f) {
    if (err) *err = "could not open data file " + path;
    return false;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  std::string tok;
  out->clear();
  while (ss >> tok) {
    for (auto &ch : tok)
      if (ch == 'd' || ch == 'D') ch = 'e';
    char *end = nullptr;
    double v = std::strtod(tok.c_str(), &end);
    if (end == tok.c_str()) continue;  // non-numeric token (blank "skipped" lines carry none)
    out->push_back(v);
  }
  return true;
}

bool Params::set(const std::string &n, double v) {
#define S(name) if (n == #name) { name = v; return true; }
  S(diff1) S(diff2) S(adrag) S(scf) S(temp0) S(temp1) S(rel) S(albocn) S(hosing) S(hosing_trend) S(ssmaxsurf)
  S(ssmaxdeep) S(rmax) S(diffamp1) S(diffamp2) S(diffwid) S(difflin) S(betaz1) S(betaz2) S(betam1) S(betam2)
  S(tatm) S(relh0_ocean) S(relh0_land) S(extra1a) S(extra1b) S(extra1c) S(scl_fwf) S(diffa_scl) S(delf2x)
  S(olr_adj0) S(olr_adj) S(t_eqm) S(albedop_offs) S(albedop_amp) S(par_sich_max) S(par_albsic_min)
  S(par_albsic_max) S(radfor_scl_co2) S(radfor_pc_co2_rise) S(diffsic) S(par_sica_thresh) S(par_sich_thresh)
  S(solconst) S(ediff0) S(ediffpow1) S(ediffpow2) S(par_bio_k0_PO4) S(par_bio_remin_POC_eL1) S(par_bio_red_POC_CaCO3)
#undef S
  return false;
}

// ------------------------------------------------------------------ helpers
static inline int isign1(int x) { return x >= 0 ? 1 : -1; }
static inline int nint_(double x) { return (int)(x >= 0 ? std::floor(x + 0.5) : -std::floor(-x + 0.5)); }
// x**n with integer n as gfortran lowers it (square-and-multiply)
static inline double powi_(double x, int m) {
  unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? 1.0 / y : y;
}

// ------------------------------------------------------------------ grid
void Grid::build(int I_, int J_, int K_, int L_, int igrid_, int nyear_, double yearlen, const std::vector<int> &k1file) {
  I = I_; J = J_; K = K_; L = L_; igrid = igrid_; nyear = nyear_;
  const double pi = 4 * std::atan(1.0);
  auto z1d = [&](std::vector<double> &v, int n) { v.assign(n + 3, 0.0); };
  z1d(ds, J); z1d(dsv, J); z1d(rds2, J); z1d(s, J); z1d(c, J); z1d(sv, J); z1d(cv, J); z1d(rc, J); z1d(rc2, J);
  z1d(rcv, J); z1d(rdsv, J); z1d(cv2, J); z1d(rds, J); z1d(asurf, J);
  z1d(dz, K); z1d(dza, K); z1d(zro, K); z1d(zw, K); z1d(rdz, K); z1d(rdza, K);
  // horizontal grid (goldstein.f90:908-969)
  const double th0 = -pi / 2, th1 = pi / 2, s0 = std::sin(th0), s1 = std::sin(th1), phix = 2 * pi;
  dphi = phix / I;
  rdphi = 1.0 / dphi;
  sv[0] = s0;
  cv[0] = std::cos(th0);
  if (igrid == 1) {
    const double dth = (th1 - th0) / J;
    for (int j = 1; j <= J; j++) {
      const double thv = th0 + j * dth, theta = thv - 0.5 * dth;
      sv[j] = std::sin(thv);
      s[j] = std::sin(theta);
      cv[j] = std::cos(thv);
    }
  } else {
    const double dscon = (s1 - s0) / J;
    for (int j = 1; j <= J; j++) {
      sv[j] = s0 + j * dscon;
      cv[j] = std::sqrt(1 - sv[j] * sv[j]);
      s[j] = sv[j] - 0.5 * dscon;
    }
  }
  for (int j = 1; j <= J; j++) {
    ds[j] = sv[j] - sv[j - 1];
    rds[j] = 1.0 / ds[j];
    c[j] = std::sqrt(1 - s[j] * s[j]);
    rc[j] = 1.0 / c[j];
    rc2[j] = rc[j] * rc[j] * rdphi;
    if (j < J) {
      dsv[j] = s[j + 1] - s[j];
      rdsv[j] = 1.0 / dsv[j];
      rcv[j] = 1.0 / cv[j];
      cv2[j] = cv[j] * cv[j] * rdsv[j];
      if (j > 1) rds2[j] = 2.0 / (dsv[j] + dsv[j - 1]);
    }
  }
  for (int j = 1; j <= J; j++) asurf[j] = kRsc * kRsc * ds[j] * dphi;
  dt = 86400.0 * yearlen / (nyear * kTsc);
  // vertical grid (goldstein.f90:983-1060)
  const double ez0 = 0.1;
  const double z1 = ez0 * (std::pow(1.0 + 1 / ez0, 1.0 / K) - 1.0);
  double tv4 = ez0 * (std::pow(z1 / ez0 + 1, 0.5) - 1), tv2 = 0, tv3, tv5;
  zro[K] = -tv4;
  zw[K] = tv2;
  for (int k = 1; k <= K; k++) {
    tv3 = ez0 * (powi_(z1 / ez0 + 1, k) - 1);
    dz[K - k + 1] = tv3 - tv2;
    tv2 = tv3;
    tv5 = ez0 * (std::pow(z1 / ez0 + 1, k + 0.5) - 1);
    if (k < K) dza[K - k] = tv5 - tv4;
    tv4 = tv5;
  }
  for (int k = K; k >= 1; k--) {
    if (k > 1) zro[k - 1] = zro[k] - dza[k - 1];
    zw[k - 1] = zw[k] - dz[k];
  }
  dzz = dz[K] * dza[K - 1] / 2;
  for (int k = 1; k <= K - 1; k++) {
    rdz[k] = 1.0 / dz[k];
    rdza[k] = 1.0 / dza[k];
  }
  rdz[K] = 1.0 / dz[K];
  dza[K] = 0.0;
  // bathymetry, file rows j = J+1 .. 0, periodic wrap (goldstein.f90:1114-1123)
  k1.assign((size_t)(I + 2) * (J + 2), 0);
  auto K1 = [&](int i, int j) -> int & { return k1[i + (I + 2) * j]; };
  size_t p = 0;
  for (int j = J + 1; j >= 0; j--) {
    for (int i = 0; i <= I + 1; i++) K1(i, j) = k1file[p++];
    K1(0, j) = K1(I, j);
    K1(I + 1, j) = K1(1, j);
  }
  ntot = 0;
  intot = 0;
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++)
      if (K1(i, j) <= K) {
        ntot += K - K1(i, j) + 1;
        intot += K - K1(i, j);
      }
  // basin boundaries, no .bmask branch (goldstein.f90:1175-1253)
  ips.assign(J + 2, 0); ipf.assign(J + 2, 0); ias.assign(J + 2, 0); iaf.assign(J + 2, 0);
  ias[J] = nint_(I * 24.0 / 36.0);
  ips[J] = nint_(I * 10.0 / 36.0);
  jsf = 1;
  if (igrid != 0) { ias[J] = 61; ips[J] = 36; jsf = 10; }
  for (int j = 1; j <= J; j++) {
    ips[j] = ips[J]; ipf[j] = ips[j]; ias[j] = ias[J]; iaf[j] = ias[j];
    if (j > nint_(J * 34.0 / 36.0) && j <= nint_(J * 35.0 / 36.0)) ias[j] = nint_(I * 20.0 / 36.0);
    for (int i = 1; i <= I; i++) {
      if (K1(ips[j] - 1, j) <= K) ips[j]--;
      if (K1(ipf[j] + 1, j) <= K) ipf[j]++;
      if (K1(ias[j] - 1, j) <= K) ias[j]--;
      if (K1(iaf[j] + 1, j) <= K) iaf[j]++;
      ips[j] = 1 + (ips[j] - 1 + I) % I;
      ipf[j] = 1 + (ipf[j] - 1 + I) % I;
      ias[j] = 1 + (ias[j] - 1 + I) % I;
      iaf[j] = 1 + (iaf[j] - 1 + I) % I;
    }
    if (igrid == 0) {
      if (ias[j] >= iaf[j] && j <= J / 2) jsf = j;
      if (ips[j] >= ipf[j] && j <= J / 2) jsf = j;
    }
  }
  if (igrid == 0) {
    for (int j = 1; j <= J; j++) {
      if (j > nint_(J * 35.0 / 36.0)) { ips[j] = 1; ipf[j] = 0; ias[j] = 1; iaf[j] = I; }
      if (j > nint_(J * 34.0 / 36.0) && j <= nint_(J * 35.0 / 36.0)) { ips[j] = 1; ipf[j] = 0; }
    }
  } else {
    ips[J] = 1; ipf[J] = 0; ips[J - 1] = 1; ipf[J - 1] = 0; ias[J] = 1; iaf[J] = I;
  }
  // seabed depth and its reciprocals at rho/u/v points (goldstein.f90:1361-1391)
  std::vector<double> h((size_t)3 * (I + 2) * (J + 2), 0.0);
  rh.assign((size_t)3 * (I + 2) * (J + 2), 0.0);
  auto H = [&](int l, int i, int j) -> double & { return h[(l - 1) + 3 * (i + (I + 2) * j)]; };
  auto RH = [&](int l, int i, int j) -> double & { return rh[(l - 1) + 3 * (i + (I + 2) * j)]; };
  for (int j = J + 1; j >= 0; j--)
    for (int i = 0; i <= I + 1; i++)
      if (K1(i, j) <= K) {
        for (int k = K1(i, j); k <= K; k++) H(3, i, j) = H(3, i, j) + dz[k];
        RH(3, i, j) = 1.0 / H(3, i, j);
      }
  for (int j = 0; j <= J + 1; j++)
    for (int i = 0; i <= I; i++) {
      H(1, i, j) = std::min(H(3, i, j), H(3, i + 1, j));
      if (std::max(K1(i, j), K1(i + 1, j)) <= K) RH(1, i, j) = 1.0 / H(1, i, j);
    }
  for (int j = 0; j <= J; j++)
    for (int i = 0; i <= I + 1; i++) {
      H(2, i, j) = std::min(H(3, i, j), H(3, i, j + 1));
      if (std::max(K1(i, j), K1(i, j + 1)) <= K) RH(2, i, j) = 1.0 / H(2, i, j);
    }
  ku.assign((size_t)2 * I * J, 0);
  mk.assign((size_t)(I + 1) * J, 0);
  getj.assign((size_t)I * J, 0);
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      ku[0 + 2 * ((i - 1) + I * (j - 1))] = std::max(K1(i, j), K1(i + 1, j));
      ku[1 + 2 * ((i - 1) + I * (j - 1))] = std::max(K1(i, j), K1(i, j + 1));
    }
  // mk / getj (goldstein.f90:1464-1494)
  auto wetk = [&](int i, int j) { return K1(i, j) * (1 + isign1(K - K1(i, j))) / 2; };
  for (int j = 1; j <= J; j++) {
    for (int i = 1; i <= I; i++) {
      int m = std::max(std::max(std::max(wetk(i, j), wetk(i + 1, j)), std::max(wetk(i - 1, j), wetk(i, j + 1))),
                       wetk(i, j - 1));
      mk[(i - 1) + (I + 1) * (j - 1)] = m * (1 + isign1(K - K1(i, j))) / 2;
    }
    mk[I + (I + 1) * (j - 1)] = mk[0 + (I + 1) * (j - 1)];
  }
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++)
      getj[(i - 1) + I * (j - 1)] =
          (std::max(std::max(K1(i, j), K1(i + 1, j)), std::max(K1(i, j + 1), K1(i + 1, j + 1))) <= K) &&
          (K1(i, j) != K1(i, j + 1) || K1(i, j) != K1(i + 1, j) || K1(i, j) != K1(i + 1, j + 1));
}

// ------------------------------------------------------------------ barotropic operator
namespace {
struct Baro {
  const Grid &g;
  MemberConsts &mc;
  int n, m, nm;
  Baro(const Grid &g_, MemberConsts &mc_) : g(g_), mc(mc_), n(g_.I), m(g_.J + 1), nm(g_.I * (g_.J + 1)) {}
  double &GAP(int k, int l) { return mc.gap[(size_t)(k - 1) + (size_t)nm * (l - 1)]; }
  double &RATM(int k, int l) { return mc.ratm[(size_t)(k - 1) + (size_t)nm * (l - 1)]; }
  double DRAG(int l, int i, int j) const { return mc.drag[(l - 1) + 2 * ((i - 1) + (g.I + 1) * (j - 1))]; }
  // goldstein.f90:3138-3204
  void invert() {
    const int I = g.I, J = g.J, K = g.K;
    const double dphi = g.dphi, rdphi = g.rdphi;
    mc.gap.assign((size_t)nm * (2 * n + 3), 0.0);
    mc.ratm.assign((size_t)nm * (n + 1), 0.0);
    for (int i = 1; i <= I; i++)
      for (int j = 0; j <= J; j++) {
        const int k = i + j * n;
        if (std::max(std::max(g.k1at(i, j), g.k1at(i + 1, j)), std::max(g.k1at(i, j + 1), g.k1at(i + 1, j + 1))) <= K) {
          const double tv = (g.s[j + 1] * g.rhat(1, i, j + 1) - g.s[j] * g.rhat(1, i, j)) / (2.0 * g.dsv[j] * dphi);
          const double tv1 = (g.sv[j] * g.rhat(2, i + 1, j) - g.sv[j] * g.rhat(2, i, j)) / (2.0 * dphi * g.dsv[j]);
          GAP(k, 2) = DRAG(1, i, j) * g.c[j] * g.c[j] * g.rhat(1, i, j) / (g.ds[j] * g.dsv[j]) + tv1;
          int l = n + 1;
          if (i == 1) l = 2 * n + 1;
          GAP(k, l) = DRAG(2, i, j) * g.rcv[j] * g.rcv[j] * rdphi * rdphi * g.rhat(2, i, j) - tv;
          GAP(k, n + 2) = -(DRAG(2, i, j) * g.rhat(2, i, j) + DRAG(2, i + 1, j) * g.rhat(2, i + 1, j)) /
                              (g.cv[j] * g.cv[j] * dphi * dphi) -
                          (DRAG(1, i, j) * g.c[j] * g.c[j] * g.rhat(1, i, j) / g.ds[j] +
                           DRAG(1, i, j + 1) * g.c[j + 1] * g.c[j + 1] * g.rhat(1, i, j + 1) / g.ds[j + 1]) /
                              g.dsv[j];
          l = n + 3;
          if (i == I) l = 3;
          GAP(k, l) = DRAG(2, i + 1, j) * g.rhat(2, i + 1, j) / (g.cv[j] * g.cv[j] * dphi * dphi) + tv;
          GAP(k, 2 * n + 2) =
              DRAG(1, i, j + 1) * g.c[j + 1] * g.c[j + 1] * g.rhat(1, i, j + 1) / (g.ds[j + 1] * g.dsv[j]) - tv1;
        } else {
          GAP(k, n + 2) = 1;
        }
      }
    for (int i = 1; i <= n * m - 1; i++) {
      const int im = std::min(i + n + 1, n * m);
      for (int j = i + 1; j <= im; j++) {
        const double rat = GAP(j, n + 2 - j + i) / GAP(i, n + 2);
        RATM(j, j - i) = rat;
        if (rat != 0)
          for (int k = n + 2 - j + i; k <= 2 * n + 3 - j + i; k++) GAP(j, k) = GAP(j, k) - rat * GAP(i, k + j - i);
      }
    }
  }
  // goldstein.f90:3500-3565; gb is 1-based and destroyed
  void ubarsolv(std::vector<double> &gb, double *ub, double *psi) {
    const int I = g.I, J = g.J;
    auto UB = [&](int l, int i, int j) -> double & { return ub[(l - 1) + 2 * (i + (I + 2) * j)]; };
    auto PSI = [&](int i, int j) -> double & { return psi[i + (I + 1) * j]; };
    for (int i = 1; i <= n * m - 1; i++) {
      const int im = std::min(i + n + 1, n * m);
      for (int j = i + 1; j <= im; j++) gb[j] = gb[j] - RATM(j, j - i) * gb[i];
    }
    gb[n * m] = gb[n * m] / GAP(n * m, n + 2);
    for (int i = n * m - 1; i >= 1; i--) {
      const int km = std::min(n + 1, n * m - i);
      for (int k = 1; k <= km; k++) gb[i] = gb[i] - GAP(i, n + 2 + k) * gb[i + k];
      gb[i] = gb[i] / GAP(i, n + 2);
    }
    for (int j = 0; j <= J; j++) {
      for (int i = 1; i <= I; i++) PSI(i, j) = gb[i + j * n];
      PSI(0, j) = PSI(I, j);
    }
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) UB(1, i, j) = -g.rhat(1, i, j) * g.c[j] * (PSI(i, j) - PSI(i, j - 1)) * g.rds[j];
    for (int j = 1; j <= J - 1; j++)
      for (int i = 1; i <= I; i++) UB(2, i, j) = g.rhat(2, i, j) * (PSI(i, j) - PSI(i - 1, j)) * g.rcv[j] * g.rdphi;
    for (int i = 1; i <= I; i++) { UB(2, i, J) = 0.0; UB(2, i, 0) = 0.0; }
    for (int j = 1; j <= J; j++) {
      UB(2, I + 1, j) = UB(2, 1, j);
      UB(1, 0, j) = UB(1, I, j);
      UB(1, I + 1, j) = UB(1, 1, j);
      UB(2, 0, j) = UB(2, I, j);
    }
    UB(2, I + 1, 0) = UB(2, 1, 0);
    UB(2, 0, 0) = UB(2, I, 0);
  }
  // goldstein_lib.f90:186-241 with indj == 0 (unit-source path integrals at init)
  double island0(const Islands &isl, const double *ub, int is) {
    const int I = g.I;
    auto UB = [&](int l, int i, int j) { return ub[(l - 1) + 2 * (i + (I + 2) * j)]; };
    double e = 0.0;
    for (int p = 1; p <= isl.npi[is]; p++) {
      const int lpi = isl.lpisl[(p - 1) + isl.mpi * (is - 1)], ipi = isl.ipisl[(p - 1) + isl.mpi * (is - 1)],
                jpi = isl.jpisl[(p - 1) + isl.mpi * (is - 1)];
      const int al = std::abs(lpi);
      double cor;
      if (al == 1)
        cor = -g.s[jpi] * 0.25 * (UB(2, ipi, jpi) + UB(2, ipi + 1, jpi) + UB(2, ipi, jpi - 1) + UB(2, ipi + 1, jpi - 1));
      else
        cor = g.sv[jpi] * 0.25 * (UB(1, ipi - 1, jpi) + UB(1, ipi, jpi) + UB(1, ipi - 1, jpi + 1) + UB(1, ipi, jpi + 1));
      // tau == 0 at initialisation time (indj = 0 drops the wind term)
      e = e + isign1(lpi) * (DRAG(al, ipi, jpi) * UB(al, ipi, jpi) + cor - 0 * 0.0) *
                  (g.c[jpi] * g.dphi * (2.0 - al) + g.rcv[jpi] * g.dsv[jpi] * (al - 1.0));
    }
    return e;
  }
};
}  // namespace

// ------------------------------------------------------------------ per-member constants
void build_member(const Grid &g, const Islands &isl, const WindFiles &w, const Params &p, MemberConsts *mcp,
                  const MemberConsts *shared) {
  MemberConsts &mc = *mcp;
  mc.p = p;
  const int I = g.I, J = g.J, K = g.K, L = g.L;
  const double pi = 4 * std::atan(1.0);
  const double syr = p.yearlen * 86400;
  auto K1 = [&](int i, int j) { return g.k1at(i, j); };
  // ---- GOLDSTEIN scalars
  mc.rpmesco = kRsc * p.saln0 / (kDsc * kUsc);
  mc.ec[1] = -0.0559 / kRhosc;
  mc.ec[2] = 0.7968 / kRhosc;
  mc.ec[3] = -0.0063 / kRhosc;
  mc.ec[4] = 3.7315e-5 / kRhosc;
  mc.ec[5] = (p.ieos == 1) ? 2.5e-5 * kDsc / kRhosc : 0.0;   // goldstein.f90:1066-1075
  mc.hosing = p.hosing;
  mc.hosing_trend = p.hosing_trend / (1.0e3 * syr);
  mc.nsteps_hosing = p.nyears_hosing * p.nyear;
  // hosing region (goldstein.f90:1275-1309)
  {
    int jh1 = 0, jh2 = 0;
    const double tv1 = std::sin(50.0 * pi / 180.0), tv2 = std::sin(70.0 * pi / 180.0);
    for (int j = 1; j <= J; j++) {
      if (tv1 >= g.sv[j - 1] && tv1 <= g.sv[j]) jh1 = (((g.sv[j] - tv1) / g.ds[j]) >= 0.5) ? j : j + 1;
      if (tv2 >= g.sv[j - 1] && tv2 <= g.sv[j]) jh2 = (((tv2 - g.sv[j - 1]) / g.ds[j]) >= 0.5) ? j : j - 1;
    }
    mc.rhosing.assign((size_t)I * J, 0.0);
    double area = 0.0;
    for (int j = jh1; j <= jh2; j++)
      for (int i = g.ias[j]; i <= g.iaf[j]; i++)
        if (K1(i, j) <= K) area = area + g.asurf[j];
    for (int j = jh1; j <= jh2; j++)
      for (int i = g.ias[j]; i <= g.iaf[j]; i++)
        if (K1(i, j) <= K) mc.rhosing[(i - 1) + I * (j - 1)] = 1e6 / area;
  }
  // drag (goldstein.f90:1400-1415, drgset 2845-2885)
  mc.adrag = 1.0 / (p.adrag * 86400 * kFsc);
  const bool reuse = shared && shared->p.adrag == p.adrag && !shared->gap.empty();
  if (reuse) {
    mc.drag = shared->drag; mc.rtv = shared->rtv; mc.rtv3 = shared->rtv3;
  } else {
    const double drgf = 3.0;
    const int kmxdrg = K / 2, jeb = 1;
    std::vector<double> tmp((size_t)(I + 1) * (J + 1), 0.0);
    auto T = [&](int i, int j) -> double & { return tmp[i + (I + 1) * j]; };
    for (int j = 0; j <= J; j++)
      for (int i = 0; i <= I; i++) {
        const int kloc2 = std::max(std::max(K1(i, j), K1(i + 1, j)), std::max(K1(i, j + 1), K1(i + 1, j + 1)));
        int kloc4 = K1(i, j);
        for (int j1 = std::max(0, j - 1); j1 <= std::min(J + 1, j + 2); j1++)
          for (int i1 = i - 1; i1 <= i + 2; i1++) kloc4 = std::max(kloc4, K1(1 + (I + i1 - 1) % I, j1));
        if (kloc2 > kmxdrg || std::abs(j - J / 2) <= jeb)
          T(i, j) = mc.adrag * drgf * drgf;
        else if (kloc4 > kmxdrg || std::abs(j - J / 2) == jeb + 1)
          T(i, j) = mc.adrag * drgf;
        else
          T(i, j) = mc.adrag;
      }
    mc.drag.assign((size_t)2 * (I + 1) * J, 0.0);
    auto D = [&](int l, int i, int j) -> double & { return mc.drag[(l - 1) + 2 * ((i - 1) + (I + 1) * (j - 1))]; };
    for (int j = 1; j <= J; j++) {
      for (int i = 1; i <= I; i++) {
        D(1, i, j) = 0.5 * (T(i, j) + T(i, j - 1));
        D(2, i, j) = 0.5 * (T(i, j) + T(i - 1, j));
      }
      D(2, I + 1, j) = D(2, 1, j);
    }
    mc.rtv.assign((size_t)I * J, 0.0);
    mc.rtv3.assign((size_t)I * J, 0.0);
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        mc.rtv[(i - 1) + I * (j - 1)] = 1.0 / (g.s[j] * g.s[j] + D(1, i, j) * D(1, i, j));
        mc.rtv3[(i - 1) + I * (j - 1)] = 1.0 / (g.sv[j] * g.sv[j] + D(2, i, j) * D(2, i, j));
      }
  }
  mc.diff1 = p.diff1 / (kRsc * kUsc);
  mc.diff2 = p.diff2 * kRsc / (kUsc * kDsc * kDsc);
  mc.rsictscsf = kDsc * g.dz[K] * kRho0 * kCpoIce / (17.5 * 86400.0);
  // initial conditions (goldstein.f90:1434-1453)
  mc.ts0.assign((size_t)L * I * J * K, 0.0);
  mc.rho0.assign((size_t)I * J * K, 0.0);
  for (int k = 1; k <= K; k++)
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        const double t = ((j <= J / 2) ? p.temp0 : p.temp1) * 0.5 * (1 + isign1(k - K1(i, j)));
        const size_t c = (size_t)(i - 1) + (size_t)I * ((j - 1) + (size_t)J * (k - 1));
        mc.ts0[0 + (size_t)L * c] = t;
        if (L > 1) mc.ts0[1 + (size_t)L * c] = 0.0;
        mc.rho0[c] = eos_z(mc.ec, p.ieos, t, 0.0, g.zro[k]);   // goldstein.f90:1450
      }
  // barotropic factorisation + unit island solves (goldstein.f90:1788-1814)
  if (isl.isles > 0) {
    if (reuse) {
      mc.gap = shared->gap; mc.ratm = shared->ratm; mc.ubisl = shared->ubisl; mc.psisl = shared->psisl;
      mc.erisl = shared->erisl;
    } else {
      Baro b(g, mc);
      b.invert();
      const int nis = isl.isles;
      mc.ubisl.assign((size_t)2 * (I + 2) * (J + 1) * nis, 0.0);
      mc.psisl.assign((size_t)(I + 1) * (J + 1) * nis, 0.0);
      mc.erisl.assign((size_t)nis * (nis + 1), 0.0);
      std::vector<double> gb(b.nm + 2, 0.0);
      for (int isol = 1; isol <= nis; isol++) {
        for (int j = 0; j <= J; j++)
          for (int i = 1; i <= I; i++) {
            const int k = i + j * I;
            gb[k] = ((int)isl.psiles[k] == isol + 1) ? 1.0 : 0.0;
          }
        double *ub = &mc.ubisl[(size_t)2 * (I + 2) * (J + 1) * (isol - 1)];
        double *ps = &mc.psisl[(size_t)(I + 1) * (J + 1) * (isol - 1)];
        b.ubarsolv(gb, ub, ps);
        for (int is = 1; is <= nis; is++) mc.erisl[(is - 1) + nis * (isol - 1)] = b.island0(isl, ub, is);
      }
      // matinv_gold (goldstein.f90:3452-3467)
      auto E = [&](int a, int bb) -> double & { return mc.erisl[(a - 1) + nis * (bb - 1)]; };
      for (int i = 1; i <= nis - 1; i++)
        for (int j = i + 1; j <= nis; j++)
          for (int k = i + 1; k <= nis; k++) E(j, k) = E(i, i) * E(j, k) - E(j, i) * E(i, k);
    }
  }
  // ssmax (goldstein.f90:2058-2070)
  mc.ssmax.assign(K + 2, 0.0);
  if (p.ssmaxsurf - p.ssmaxdeep < 1.0e-7 && p.ssmaxsurf - p.ssmaxdeep > -1.0e-7) {
    for (int k = 1; k <= K - 1; k++) mc.ssmax[k] = p.ssmaxdeep;
  } else {
    const double mid = 0.5 * (std::log(p.ssmaxsurf) + std::log(p.ssmaxdeep));
    const double dif = 0.5 * (std::log(p.ssmaxsurf) - std::log(p.ssmaxdeep));
    const double efold = 200 / kDsc, dep0 = -300 / kDsc;
    for (int k = 1; k <= K - 1; k++) mc.ssmax[k] = std::exp(mid + dif * std::tanh((g.zw[k] - dep0) / efold));
  }
  // IF (iediff > 0) CALL ediff (goldstein.f90:2053-2055, 2936-3044); ediffvar = 0: ediff1(i,j,k) = ediff1p(k)
  mc.ediff1p.assign(K + 2, 0.0);
  mc.diffmax.assign(K + 2, 0.0);
  if (p.iediff > 0 && p.iediff < 3) {
    mc.ediff0 = p.ediff0 * kRsc / (kUsc * kDsc * kDsc);
    const double ediff10 = mc.diff2 - mc.ediff0;
    for (int k = 1; k <= K - 1; k++) {
      const double dzrho_lev = (-5.5e-3 / kRhosc * kDsc) * std::exp(g.zw[k] * (kDsc / 650.0));
      double ediffk0;
      if (p.iediff == 1) {
        ediffk0 = std::exp(-(g.zw[k] + 2500.0 / kDsc) * (kDsc / 700.0));
        const double ediffklim = 1 / 3.0e0;
        ediffk0 = 1 / ((1 - ediffklim) / ediffk0 + ediffklim);
      } else {
        ediffk0 = 1 + (2 / (4 * std::atan(1.0))) * std::atan(-(g.zw[k] + 2500.0 / kDsc) * (4.5e-3 * kDsc));
      }
      mc.ediff1p[k] = ediff10 * std::pow(ediffk0, p.ediffpow1) * std::pow(-dzrho_lev, p.ediffpow2);
    }
    if (p.ediffpow2 > -1.0e-7 && p.ediffpow2 < 1.0e-7) mc.ediffpow2i = 0;
    else if (p.ediffpow2 > (1.0 - 1.0e-7) && p.ediffpow2 < (1.0 + 1.0e-7)) mc.ediffpow2i = 1;
    else if (p.ediffpow2 > (0.5 - 1.0e-7) && p.ediffpow2 < (0.5 + 1.0e-7)) mc.ediffpow2i = 2;
    else mc.ediffpow2i = -999;
    for (int k = 1; k <= K; k++) mc.diffmax[k] = 0.5 * 0.125 * g.dz[k] * g.dz[k] / g.dt;
  }
  if (w.taux_u.empty()) return;  // tracer-only handle: no atmosphere / sea ice

  // ---- EMBM (embm.f90:739-1475)
  const double tv = 86400.0 * p.yearlen / (p.nyear * kTsc);
  mc.ryear = 1.0 / (p.yearlen * 86400);
  mc.dtatm = tv / p.ndta;
  mc.rdtdim = 1.0 / (kTsc * g.dt);
  const size_t ij = (size_t)I * J;
  auto at2 = [&](std::vector<double> &a, int i, int j) -> double & { return a[(i - 1) + I * (j - 1)]; };
  auto at3 = [&](std::vector<double> &a, int l, int i, int j) -> double & { return a[(l - 1) + 2 * ((i - 1) + I * (j - 1))]; };
  mc.us_dztau.assign(2 * ij, 0.0);
  mc.us_dztav.assign(2 * ij, 0.0);
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      const size_t q = (i - 1) + (size_t)I * (j - 1);
      at3(mc.us_dztau, 1, i, j) = w.taux_u[q];
      at3(mc.us_dztau, 2, i, j) = w.tauy_u[q];
      at3(mc.us_dztav, 1, i, j) = w.taux_v[q];
      at3(mc.us_dztav, 2, i, j) = w.tauy_v[q];
    }
  mc.albcl.assign(ij, 0.0);
  for (int j = 1; j <= J; j++) {
    const double scl = powi_((p.albedop_skew - g.s[j]) / 2.0, p.albedop_skewp);
    const double a = std::asin(g.s[j]);
    const double v = p.albedop_offs + p.albedop_amp * 0.5 *
                                          (1.0 - std::cos(2.0 * a) + scl * p.albedop_mod2 * std::cos(2.0 * a) +
                                           scl * p.albedop_mod4 * std::cos(4.0 * a) + scl * p.albedop_mod6 * std::cos(6.0 * a));
    for (int i = 1; i <= I; i++) at2(mc.albcl, i, j) = v;
  }
  mc.ca.assign(ij, 0.0);
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) at2(mc.ca, i, j) = (K1(i, j) <= K) ? 0.3 : 1.0;
  mc.rate_co2 = p.radfor_pc_co2_rise * 0.01 * kTsc * mc.dtatm * p.ndta * mc.ryear;
  mc.rate_ch4 = p.radfor_pc_ch4_rise * 0.01 * kTsc * mc.dtatm * p.ndta * mc.ryear;
  mc.rate_n2o = p.radfor_pc_n2o_rise * 0.01 * kTsc * mc.dtatm * p.ndta * mc.ryear;
  mc.hatmbl1 = 8400.0;
  mc.rfluxsca = kRsc / (mc.hatmbl1 * kUsc * kRhoair * kCpa);
  // advective winds (embm.f90:985-1024)
  mc.uatm.assign(2 * ij, 0.0);
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      at3(mc.uatm, 1, i, j) = w.uncep[(i - 1) + (size_t)I * (j - 1)];
      at3(mc.uatm, 2, i, j) = w.vncep[(i - 1) + (size_t)I * (j - 1)];
    }
  const bool polar = p.par_wind_polar_avg != 1 && p.par_wind_polar_avg != 2;
  if (polar)
    for (int j = 1; j <= J; j++)
      if (j <= 2 || j >= J - 1)
        for (int l = 1; l <= 2; l++) {
          double t = 0.0;
          for (int i = 1; i <= I; i++) t = t + at3(mc.uatm, l, i, j);
          t = t / I;
          for (int i = 1; i <= I; i++) at3(mc.uatm, l, i, j) = t;
        }
  for (auto &x : mc.uatm) x = x / kUsc;
  if (polar)
    for (int i = 1; i <= I; i++) at3(mc.uatm, 2, i, J) = 0.;
  mc.ppmin = 2.0 / (p.yearlen * 86400.0);
  mc.ppmax = 4.0 / (p.yearlen * 86400.0);
  // diffusivities (embm.f90:1031-1078)
  mc.diffa.assign((size_t)4 * J, 0.0);
  auto DA = [&](int l, int m2, int j) -> double & { return mc.diffa[(l - 1) + 2 * ((m2 - 1) + 2 * (j - 1))]; };
  const double diffend = std::exp(-((0.5 * pi / p.diffwid) * (0.5 * pi / p.diffwid)));
  for (int j = 1; j <= J; j++) {
    const double a = std::asin(g.s[j]), a2 = std::asin(g.sv[j]);
    DA(2, 1, j) = p.diffamp2;
    DA(2, 2, j) = p.diffamp2;
    DA(1, 1, j) = p.diffamp1 * (p.difflin * 2.0 * (a + 0.5 * pi) / pi +
                                (1.0 - p.difflin) * (std::exp(-((a / p.diffwid) * (a / p.diffwid))) - diffend) / (1.0 - diffend));
    DA(1, 2, j) = p.diffamp1 * (p.difflin * 2.0 * (a2 + 0.5 * pi) / pi +
                                (1.0 - p.difflin) * (std::exp(-((a2 / p.diffwid) * (a2 / p.diffwid))) - diffend) / (1.0 - diffend));
    if (p.diffa_len < 0) {
      if (std::sin(pi * (double)p.diffa_len / 180.0) > g.sv[j]) DA(1, 2, j) = p.diffa_scl * DA(1, 2, j);
    } else if (j <= p.diffa_len) {
      DA(1, 2, j) = p.diffa_scl * DA(1, 2, j);
    }
    DA(1, 1, j) = DA(1, 1, j) / (kRsc * kUsc);
    DA(1, 2, j) = DA(1, 2, j) / (kRsc * kUsc);
    DA(2, 1, j) = DA(2, 1, j) / (kRsc * kUsc);
    DA(2, 2, j) = DA(2, 2, j) / (kRsc * kUsc);
    if (g.igrid == 1 || g.igrid == 2) DA(2, 1, j) = std::min(DA(2, 1, j), DA(1, 1, j));
  }
  mc.hatmbl2 = 1800.;
  mc.rpmesca = kRsc * kRho0 / (mc.hatmbl2 * kUsc * kRhoair);
  mc.extra1a = p.scl_fwf * p.extra1a;
  mc.extra1b = p.scl_fwf * p.extra1b;
  mc.extra1c = p.scl_fwf * p.extra1c;
  // Atlantic/Pacific P-E adjustment (embm.f90:1168-1318, igrid == 0)
  mc.pmeadj.assign(ij, 0.0);
  {
    int j1as = g.jsf + 1, j1bs = 0, j1cs = 0;
    const double t20 = std::sin(-20.0 * pi / 180.0), t24 = std::sin(24.0 * pi / 180.0);
    for (int j = 1; j <= J; j++) {
      if (t20 >= g.sv[j - 1] && t20 <= g.sv[j]) j1bs = ((g.sv[j] - t20) / g.ds[j] >= 0.5) ? j : j + 1;
      if (t24 >= g.sv[j - 1] && t24 <= g.sv[j]) j1cs = ((g.sv[j] - t24) / g.ds[j] >= 0.5) ? j : j + 1;
    }
    if (g.igrid == 0) {
      int npa = 0, naa = 0, npb = 0, nab = 0, npc = 0, nac = 0;
      for (int j = j1as; j <= j1bs - 1; j++) { npa += g.ipf[j] - g.ips[j] + 1; naa += g.iaf[j] - g.ias[j] + 1; }
      for (int j = j1bs; j <= j1cs - 1; j++) { npb += g.ipf[j] - g.ips[j] + 1; nab += g.iaf[j] - g.ias[j] + 1; }
      for (int j = j1cs; j <= J; j++) {
        for (int i = g.ips[j]; i <= g.ipf[j]; i++) if (K1(i, j) <= K) npc++;
        for (int i = g.ias[j]; i <= g.iaf[j]; i++) if (K1(i, j) <= K) nac++;
      }
      for (int j = j1as; j <= j1bs - 1; j++) {
        for (int i = g.ips[j]; i <= g.ipf[j]; i++) at2(mc.pmeadj, i, j) = 1.0e6 * mc.extra1a / (npa * g.asurf[j]);
        for (int i = g.ias[j]; i <= g.iaf[j]; i++) at2(mc.pmeadj, i, j) = -1.0e6 * mc.extra1a / (naa * g.asurf[j]);
      }
      for (int j = j1bs; j <= j1cs - 1; j++) {
        for (int i = g.ips[j]; i <= g.ipf[j]; i++) at2(mc.pmeadj, i, j) = 1.0e6 * mc.extra1b / (npb * g.asurf[j]);
        for (int i = g.ias[j]; i <= g.iaf[j]; i++) at2(mc.pmeadj, i, j) = -1.0e6 * mc.extra1b / (nab * g.asurf[j]);
      }
      for (int j = j1cs; j <= J; j++) {
        for (int i = g.ips[j]; i <= g.ipf[j]; i++)
          if (K1(i, j) <= K) at2(mc.pmeadj, i, j) = 1.0e6 * mc.extra1c / (npc * g.asurf[j]);
        for (int i = g.ias[j]; i <= g.iaf[j]; i++)
          if (K1(i, j) <= K) at2(mc.pmeadj, i, j) = -1.0e6 * mc.extra1c / (nac * g.asurf[j]);
      }
    }
  }
  // initial atmosphere (embm.f90:1435-1475); tstar_ocn = ts(1,:,:,maxk) from the ocean init
  mc.tq0.assign(2 * ij, 0.0);
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      const double to = mc.ts0[0 + (size_t)L * ((size_t)(i - 1) + (size_t)I * ((j - 1) + (size_t)J * (K - 1)))];
      at3(mc.tq0, 1, i, j) = p.tatm;
      if (K1(i, j) <= K) {
        if (to > kTsic)
          at3(mc.tq0, 2, i, j) = p.relh0_ocean * kConst1 * std::exp(kConst2 * to / (to + kConst3));
        else
          at3(mc.tq0, 2, i, j) = p.relh0_ocean * kConst1 * std::exp(kConst4 * to / (to + kConst5));
      } else {
        const double t1 = p.tatm;
        if (t1 > 0.0)
          at3(mc.tq0, 2, i, j) = p.relh0_land * kConst1 * std::exp(kConst2 * t1 / (t1 + kConst3));
        else
          at3(mc.tq0, 2, i, j) = p.relh0_land * kConst1 * std::exp(kConst4 * t1 / (t1 + kConst5));
      }
    }
  // runoff routing (embm.f90:3787-3836)
  mc.iroff.assign(ij, 0);
  mc.jroff.assign(ij, 0);
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      int ir = i, jr = j, loop = 0;
      while (K1(ir, jr) > K) {
        const int kv = K1(ir, jr);
        if (kv == 91) ir++;
        else if (kv == 92) jr--;
        else if (kv == 93) ir--;
        else if (kv == 94) jr++;
        if (ir == I + 1) ir = 1;
        else if (ir == 0) ir = I;
        if (++loop > 100000) break;
      }
      mc.iroff[(i - 1) + I * (j - 1)] = ir;
      mc.jroff[(i - 1) + I * (j - 1)] = jr;
    }
  // insolation table (radfor, embm.f90:2383-2522, present-day orbit)
  mc.solfor.assign((size_t)J * p.nyear, 0.0);
  {
    const double osce = 0.0167, oscsob = 0.397789, oscgam = 1.352631, osctau0 = -0.5;
    const double rpi = 1.0 / pi, e2 = osce * osce;
    const double osce1 = osce * (2.0 - 0.25 * e2), osce2 = 1.25 * e2, osce3 = osce * e2 * 13. / 12.;
    const double r4 = (1.0 + 0.5 * e2) / (1.0 - e2), osce4 = r4 * r4;
    const double oscryr = 2.0 * pi / (double)p.nyear, osctau1 = osctau0 + 0.5;
    for (int n = 1; n <= p.nyear; n++) {
      const double osct = ((double)((n - 1) % p.nyear + 1) - (p.nyear * osctau1 / p.gn_daysperyear)) * oscryr;
      for (int j = 1; j <= J; j++) {
        const double oscv = osct + osce1 * std::sin(osct) + osce2 * std::sin(2.0 * osct) + osce3 * std::sin(3.0 * osct);
        const double q = 1.0 + osce * std::cos(oscv);
        const double oscsolf = osce4 * (q * q);
        const double oscsind = oscsob * std::sin(oscv - oscgam);
        const double oscss = oscsind * g.s[j];
        const double osccc = std::sqrt(1.0 - oscsind * oscsind) * g.c[j];
        const double osctt = std::min(1.0, std::max(-1.0, oscss / osccc));
        const double oscday = std::acos(-osctt);
        mc.solfor[(j - 1) + (size_t)J * (n - 1)] = p.solconst * oscsolf * rpi * (oscss * oscday + osccc * std::sin(oscday));
      }
    }
  }
  // wind stresses scaled for the ocean and surface wind speed (goldstein.f90:102-107; embm.f90:2762-2816)
  mc.dztau.assign(2 * ij, 0.0);
  mc.dztav.assign(2 * ij, 0.0);
  mc.tau.assign(2 * ij, 0.0);
  mc.usurf.assign(ij, 0.0);
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      at3(mc.dztau, 1, i, j) = p.scf * at3(mc.us_dztau, 1, i, j) / (kRh0sc * kDsc * kUsc * kFsc) / g.dzz;
      at3(mc.dztau, 2, i, j) = p.scf * at3(mc.us_dztau, 2, i, j) / (kRh0sc * kDsc * kUsc * kFsc) / g.dzz;
      at3(mc.dztav, 1, i, j) = p.scf * at3(mc.us_dztav, 1, i, j) / (kRh0sc * kDsc * kUsc * kFsc) / g.dzz;
      at3(mc.dztav, 2, i, j) = p.scf * at3(mc.us_dztav, 2, i, j) / (kRh0sc * kDsc * kUsc * kFsc) / g.dzz;
      at3(mc.tau, 1, i, j) = at3(mc.dztau, 1, i, j) * g.dzz;
      at3(mc.tau, 2, i, j) = at3(mc.dztav, 2, i, j) * g.dzz;
    }
  for (int j = 1; j <= J; j++) {
    double tv3 = 0.0;
    for (int i = 1; i <= I; i++) {
      const double a = (i == 1) ? (at3(mc.tau, 1, i, j) + at3(mc.tau, 1, I, j)) / 2 : (at3(mc.tau, 1, i, j) + at3(mc.tau, 1, i - 1, j)) / 2;
      const double b = (j == 1) ? at3(mc.tau, 2, i, j) / 2 : (at3(mc.tau, 2, i, j) + at3(mc.tau, 2, i, j - 1)) / 2;
      at2(mc.usurf, i, j) = std::sqrt((std::sqrt(a * a + b * b)) * kRh0sc * kDsc * kUsc * kFsc / (kRhoair * kCd * p.scf));
      tv3 = tv3 + at2(mc.usurf, i, j);
    }
    if (p.par_wind_polar_avg != 2 && (j <= 2 || j >= J - 1))
      for (int i = 1; i <= I; i++) at2(mc.usurf, i, j) = tv3 / I;
  }
  // wind energy input of the mixed-layer scheme (imld = 1, goldstein.f90:112-141: "taken from surflux, but min(j, maxj-1)") and its
  // decay with depth (:1675-1686; mldwindkedec in units of dsc)
  mc.mldketau.assign(ij, 0.0);
  mc.mlddec.assign(K + 2, 0.0);
  mc.mlddecd.assign(K + 2, 0.0);
  if (p.imld == 1) {
    for (int j = 1; j <= J; j++) {
      double tv3 = 0.0;
      for (int i = 1; i <= I; i++) {
        const double tv4 = (i == 1) ? (at3(mc.tau, 1, i, j) + at3(mc.tau, 1, I, j)) * 0.5 : (at3(mc.tau, 1, i, j) + at3(mc.tau, 1, i - 1, j)) * 0.5;
        const double tv2 = (j == 1) ? at3(mc.tau, 2, i, j) * 0.5 : (at3(mc.tau, 2, i, std::min(j, J - 1)) + at3(mc.tau, 2, i, j - 1)) * 0.5;
        const double r = std::sqrt(std::sqrt(tv4 * tv4 + tv2 * tv2));
        at2(mc.mldketau, i, j) = p.mldketaucoeff * (r * r * r);
        tv3 = tv3 + at2(mc.mldketau, i, j);
      }
      if (j <= 2 || j >= J - 1)
        for (int i = 1; i <= I; i++) at2(mc.mldketau, i, j) = tv3 / I;
    }
    const double dec = p.mldwindkedec / kDsc;
    for (int k = K; k >= 1; k--) {
      mc.mlddec[k] = std::exp(g.zro[k] / dec);
      mc.mlddecd[k] = (k < K) ? mc.mlddec[k] / mc.mlddec[k + 1] : mc.mlddec[K];
    }
  }
  // wind speeds as exported to the coupler and read back by step_embm (embm.f90:1680-1681, 49-50)
  mc.lowestlu2.assign(ij, 0.0);
  mc.lowestlv3.assign(ij, 0.0);
  for (int j = 1; j <= J; j++)
    for (int i = 1; i <= I; i++) {
      at2(mc.lowestlu2, i, j) = at3(mc.uatm, 1, i, j) * kUsc;
      at2(mc.lowestlv3, i, j) = at3(mc.uatm, 2, i, j) * kUsc;
    }
  // ---- sea ice (gold_seaice.f90:256-304)
  mc.dtsic = tv;
  mc.sic_rdtdim = 1.0 / (kTsc * mc.dtsic);
  mc.diffsic = p.diffsic / (kRsc * kUsc);
}

// ------------------------------------------------------------------ job directory
static std::string joinp(const std::string &a, const std::string &b) {
  if (a.empty()) return b;
  if (!b.empty() && b[0] == '/') return b;
  return a.back() == '/' ? a + b : a + "/" + b;
}

bool load_job(const std::string &jobdir, Params *p, Grid *g, Islands *isl, WindFiles *w, std::string *err) {
  Namelist main, go, eb, gs;
  if (!main.load(joinp(jobdir, "data_genie"), err)) return false;
  if (!go.load(joinp(jobdir, "data_GOLD"), err)) return false;
  if (!eb.load(joinp(jobdir, "data_EMBM"), err)) return false;
  if (!gs.load(joinp(jobdir, "data_goldSIC"), err)) return false;
  Params d;  // defaults
  p->maxi = main.integer("dim_GOLDSTEINNLONS", d.maxi);
  p->maxj = main.integer("dim_GOLDSTEINNLATS", d.maxj);
  p->maxk = main.integer("dim_GOLDSTEINNLEVS", d.maxk);
  p->maxl = main.integer("dim_GOLDSTEINNTRACS", d.maxl);
  p->kocn_loop = main.integer("kocn_loop", d.kocn_loop);
  p->katm_loop = main.integer("katm_loop", d.katm_loop);
  p->ksic_loop = main.integer("ksic_loop", d.ksic_loop);
  p->conv_kocn_kbiogem = main.integer("conv_kocn_kbiogem", d.conv_kocn_kbiogem);
  p->conv_kocn_katchem = main.integer("conv_kocn_katchem", d.conv_kocn_katchem);
  p->flag_biogem = main.flag("flag_biogem", false);
  p->flag_atchem = main.flag("flag_atchem", false);
  p->solconst = main.num("genie_solar_constant", d.solconst);
  p->genie_timestep = main.num("genie_timestep", d.genie_timestep);
  if (!main.flag("flag_ebatmos", true) || !main.flag("flag_goldsteinocean", true) || !main.flag("flag_goldsteinseaice", true)) {
    if (err) *err = "hot path needs flag_ebatmos, flag_goldsteinocean and flag_goldsteinseaice";
    return false;
  }
  for (const char *off : {"flag_ents", "flag_sedgem", "flag_rokgem", "flag_gemlite"})
    if (main.flag(off, false)) {
      if (err) *err = std::string(off) + " is outside the B200 hot path";
      return false;
    }
#define GN(nl, key) p->key = nl.num(#key, d.key)
#define GI(nl, key) p->key = nl.integer(#key, d.key)
  GI(go, igrid); GI(go, nyear); GN(go, yearlen); GN(go, temp0); GN(go, temp1); GN(go, rel); GN(go, scf);
  p->diff1 = go.num("diff(1)", d.diff1);
  p->diff2 = go.num("diff(2)", d.diff2);
  GN(go, adrag); GN(go, hosing); GN(go, hosing_trend); GI(go, nyears_hosing); GN(go, albocn); GI(go, iconv);
  GI(go, imld); GI(go, iediff); GI(go, ieos); GN(go, ssmaxsurf); GN(go, ssmaxdeep); GN(go, saln0);
  GN(go, ediff0); GN(go, ediffpow1); GN(go, ediffpow2); GN(go, ediffvar);
  GN(go, mldpebuoycoeff); GN(go, mldketaucoeff); GN(go, mldwindkedec);
  p->diso = go.flag("diso", true);
  p->world = go.str("world", d.world);
  p->go_indir = go.str("indir_name", d.go_indir);
  if (p->iconv < 0 || p->iconv > 1 || p->imld < 0 || p->imld > 1 || p->ieos < 0 || p->ieos > 1) {
    if (err) *err = "imld / iconv / ieos outside 0..1 are outside the B200 hot path (SURVEY 8f.4)";
    return false;
  }
  if (p->iediff < 0 || p->iediff > 2 || (p->iediff != 0 && (p->ediffvar < -1.0e-7 || p->ediffvar > 1.0e-7))) {
    if (err) *err = "iediff must be 0, 1 or 2 and ediffvar 0 (no ediffvargrid.dat) on the B200 hot path (SURVEY 8f.4)";
    return false;
  }
  if (lower(go.str("fwanomin", "n")) == "y") {
    if (err) *err = "fwanomin='y' is outside the B200 hot path";
    return false;
  }
  GI(eb, ndta); GN(eb, rmax);
  p->diffamp1 = eb.num("diffamp(1)", d.diffamp1);
  p->diffamp2 = eb.num("diffamp(2)", d.diffamp2);
  GN(eb, diffwid); GN(eb, difflin);
  p->betaz1 = eb.num("betaz(1)", d.betaz1);
  p->betaz2 = eb.num("betaz(2)", d.betaz2);
  p->betam1 = eb.num("betam(1)", d.betam1);
  p->betam2 = eb.num("betam(2)", d.betam2);
  GN(eb, tatm); GN(eb, relh0_ocean); GN(eb, relh0_land); GN(eb, extra1a); GN(eb, extra1b); GN(eb, extra1c);
  GN(eb, scl_fwf); GN(eb, z1_embm); GN(eb, diffa_scl); GI(eb, diffa_len); GN(eb, delf2x); GN(eb, olr_adj0);
  GN(eb, olr_adj); GN(eb, t_eqm); GN(eb, albedop_offs); GN(eb, albedop_amp); GN(eb, albedop_skew);
  GI(eb, albedop_skewp); GN(eb, albedop_mod2); GN(eb, albedop_mod4); GN(eb, albedop_mod6); GN(eb, par_sich_max);
  GN(eb, par_albsic_min); GN(eb, par_albsic_max); GI(eb, par_wind_polar_avg); GN(eb, radfor_scl_co2);
  GN(eb, radfor_pc_co2_rise); GN(eb, radfor_scl_ch4); GN(eb, radfor_pc_ch4_rise); GN(eb, radfor_scl_n2o);
  GN(eb, radfor_pc_n2o_rise);
  p->atchem_radfor = lower(eb.str("atchem_radfor", "n")) == "y";
  p->eb_indir = eb.str("indir_name", d.eb_indir);
  p->xu_wstress = eb.str("xu_wstress", d.xu_wstress);
  p->yu_wstress = eb.str("yu_wstress", d.yu_wstress);
  p->xv_wstress = eb.str("xv_wstress", d.xv_wstress);
  p->yv_wstress = eb.str("yv_wstress", d.yv_wstress);
  p->u_wspeed = eb.str("u_wspeed", d.u_wspeed);
  p->v_wspeed = eb.str("v_wspeed", d.v_wspeed);
  if (eb.integer("orogswitch", 0) != 0 || eb.integer("t_co2", 0) != 0 || eb.flag("useforc", false) ||
      lower(eb.str("orbit_radfor", "n")) == "y" || eb.integer("t_orog", 0) != 0 || eb.integer("t_lice", 0) != 0 ||
      eb.integer("t_d18o", 0) != 0) {
    if (err) *err = "EMBM orography / orbit / CO2-series options are outside the B200 hot path";
    return false;
  }
  GN(gs, diffsic); GN(gs, par_sica_thresh); GN(gs, par_sich_thresh);
  if (gs.flag("impsic", false)) {
    if (err) *err = "impsic=.TRUE. is outside the B200 hot path";
    return false;
  }
#undef GN
#undef GI
  // data files
  const std::string gdir = joinp(jobdir, p->go_indir), edir = joinp(jobdir, p->eb_indir);
  std::vector<double> k1d, ps, pa;
  if (!read_numbers(joinp(gdir, p->world + ".k1"), &k1d, err)) return false;
  const int I = p->maxi, J = p->maxj;
  if ((int)k1d.size() < (I + 2) * (J + 2)) {
    if (err) *err = "bathymetry file too short";
    return false;
  }
  std::vector<int> k1f((size_t)(I + 2) * (J + 2));
  for (size_t q = 0; q < k1f.size(); q++) k1f[q] = (int)k1d[q];
  g->build(I, J, p->maxk, p->maxl, p->igrid, p->nyear, p->yearlen, k1f);
  if (!read_numbers(joinp(gdir, p->world + ".psiles"), &ps, err)) return false;
  if ((int)ps.size() < I * (J + 1)) {
    if (err) *err = "psiles file too short";
    return false;
  }
  isl->psiles.assign((size_t)I * (J + 1) + 2, 0.0);
  isl->isles = 0;
  {
    size_t q = 0;
    for (int j = J; j >= 0; j--)
      for (int i = 1; i <= I; i++) {
        const double v = ps[q++];
        isl->psiles[i + j * I] = v;
        if (v > (double)isl->isles) isl->isles = (int)v;
      }
  }
  isl->isles -= 1;
  isl->mpi = 2 * (I + J);
  if (isl->isles > 0) {
    if (!read_numbers(joinp(gdir, p->world + ".paths"), &pa, err)) return false;
    const int n = isl->isles;
    isl->npi.assign(n + 2, 0);
    isl->lpisl.assign((size_t)isl->mpi * n, 0);
    isl->ipisl.assign((size_t)isl->mpi * n, 0);
    isl->jpisl.assign((size_t)isl->mpi * n, 0);
    size_t q = 0;
    for (int i = 1; i <= n; i++) isl->npi[i] = (int)pa[q++];
    for (int i = 1; i <= n; i++) {
      if (isl->npi[i] > isl->mpi) {
        if (err) *err = "path integral around island too long";
        return false;
      }
      for (int j = 1; j <= isl->npi[i]; j++) {
        if (q + 3 > pa.size()) {
          if (err) *err = "paths file too short";
          return false;
        }
        const int lp = (int)pa[q], ip = (int)pa[q + 1], jp = (int)pa[q + 2];
        q += 3;
        if ((std::abs(lp) != 1 && std::abs(lp) != 2) || ip > I || ip < 0 || jp > J || jp < 0 || g->k1at(ip, jp) > p->maxk) {
          if (err) *err = "bad island path";
          return false;
        }
        isl->lpisl[(j - 1) + isl->mpi * (i - 1)] = lp;
        isl->ipisl[(j - 1) + isl->mpi * (i - 1)] = ip;
        isl->jpisl[(j - 1) + isl->mpi * (i - 1)] = jp;
      }
    }
  }
  const size_t ij = (size_t)I * J;
  struct { const std::string *name; std::vector<double> *dst; } wf[] = {
      {&p->xu_wstress, &w->taux_u}, {&p->yu_wstress, &w->tauy_u}, {&p->xv_wstress, &w->taux_v},
      {&p->yv_wstress, &w->tauy_v}, {&p->u_wspeed, &w->uncep}, {&p->v_wspeed, &w->vncep}};
  for (auto &f : wf) {
    if (!read_numbers(joinp(edir, *f.name), f.dst, err)) return false;
    if (f.dst->size() < ij) {
      if (err) *err = "wind file too short: " + *f.name;
      return false;
    }
  }
  return true;
}

}  // namespace cg
