// cg_biogem.cu -- host-only code (compiled by nvcc so that it can share cg_device.cuh): BIOGEM/ATCHEM configuration,
// tracer tables and forcing.  See cg_biogem.hpp for the reference citations.
#include "cg_biogem.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "cg_device.cuh"

namespace cg {

namespace {
// full-table ids (data/main/tracer_define.{ocn,sed,atm})
enum { IO_T = 1, IO_S = 2, IO_DIC = 3, IO_DIC_13C = 4, IO_DIC_14C = 5, IO_PO4 = 8, IO_O2 = 10, IO_ALK = 12, IO_DOM_C = 15,
       IO_DOM_C_13C = 16, IO_DOM_C_14C = 17, IO_DOM_P = 20, IO_CA = 35, IO_CFC11 = 45, IO_CFC12 = 46, IO_MG = 50 };
enum { IS_POC = 3, IS_POC_13C = 4, IS_POC_14C = 5, IS_POP = 8, IS_CACO3 = 14, IS_CACO3_13C = 15, IS_CACO3_14C = 16,
       IS_POC_FRAC2 = 33, IS_CACO3_FRAC2 = 34 };
enum { IA_T = 1, IA_Q = 2, IA_PCO2 = 3, IA_PCO2_13C = 4, IA_PCO2_14C = 5, IA_PO2 = 6, IA_PCFC11 = 18, IA_PCFC12 = 19 };
struct Def { int id, dep, type; const char *name; };
// the frozen selection: id, dependency, type (columns 2-4 of tracer_define.*)
const Def kOcn[] = {{IO_T, IO_T, 0, "temp"}, {IO_S, IO_S, 0, "sal"}, {IO_DIC, IO_DIC, 1, "DIC"}, {IO_DIC_13C, IO_DIC, 11, "DIC_13C"},
    {IO_DIC_14C, IO_DIC, 12, "DIC_14C"}, {IO_PO4, IO_PO4, 1, "PO4"}, {IO_O2, IO_O2, 1, "O2"}, {IO_ALK, IO_ALK, 1, "ALK"},
    {IO_DOM_C, IO_DOM_C, 1, "DOM_C"}, {IO_DOM_C_13C, IO_DOM_C, 11, "DOM_C_13C"}, {IO_DOM_C_14C, IO_DOM_C, 12, "DOM_C_14C"},
    {IO_DOM_P, IO_DOM_P, 1, "DOM_P"}, {IO_CA, IO_CA, 1, "Ca"}, {IO_CFC11, IO_CFC11, 1, "CFC11"}, {IO_CFC12, IO_CFC12, 1, "CFC12"},
    {IO_MG, IO_MG, 1, "Mg"}};
const Def kSed[] = {{IS_POC, IS_POC, 1, "POC"}, {IS_POC_13C, IS_POC, 11, "POC_13C"}, {IS_POC_14C, IS_POC, 12, "POC_14C"},
    {IS_POP, IS_POP, 3, "POP"}, {IS_CACO3, IS_CACO3, 1, "CaCO3"}, {IS_CACO3_13C, IS_CACO3, 11, "CaCO3_13C"},
    {IS_CACO3_14C, IS_CACO3, 12, "CaCO3_14C"}, {IS_POC_FRAC2, IS_POC_FRAC2, 9, "POC_frac2"}, {IS_CACO3_FRAC2, IS_CACO3_FRAC2, 9, "CaCO3_frac2"}};
const Def kAtm[] = {{IA_T, IA_T, 0, "pT"}, {IA_Q, IA_Q, 0, "pq"}, {IA_PCO2, IA_PCO2, 1, "pCO2"}, {IA_PCO2_13C, IA_PCO2, 11, "pCO2_13C"},
    {IA_PCO2_14C, IA_PCO2, 12, "pCO2_14C"}, {IA_PO2, IA_PO2, 1, "pO2"}, {IA_PCFC11, IA_PCFC11, 1, "pCFC11"}, {IA_PCFC12, IA_PCFC12, 1, "pCFC12"}};
constexpr int kNO = sizeof(kOcn) / sizeof(kOcn[0]), kNS = sizeof(kSed) / sizeof(kSed[0]), kNA = sizeof(kAtm) / sizeof(kAtm[0]);

std::string jp(const std::string &a, const std::string &b) {
  if (a.empty()) return b;
  if (!b.empty() && b[0] == '/') return b;
  return a.back() == '/' ? a + b : a + "/" + b;
}
std::string key_n(const char *base, int n) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "%s(%d)", base, n);
  return buf;
}
// the lines between -START-OF-DATA- and -END-OF-DATA- (sub_check_fileformat, gem_util.f90:356-419)
bool data_lines(const std::string &path, std::vector<std::string> *out, std::string *err) {
  std::ifstream f(path);
  if (!f) { if (err) *err = "could not open " + path; return false; }
  std::string line;
  bool in = false;
  while (std::getline(f, line)) {
    if (line.find("-START-OF-DATA-") != std::string::npos) { in = true; continue; }
    if (line.find("-END-OF-DATA-") != std::string::npos) break;
    if (in && line.find_first_not_of(" \t\r") != std::string::npos) out->push_back(line);
  }
  return true;
}
bool is_true(const std::string &t) { return !t.empty() && (t[0] == 't' || t[0] == 'T' || t == ".true." || t == ".TRUE."); }
int find_id(const std::vector<int> &v, int id) {
  for (size_t q = 1; q < v.size(); q++) if (v[q] == id) return (int)q;
  return 0;
}
}  // namespace

bool load_biogem(const std::string &jobdir, const Params &p, const Grid &g, BgConfig *c, std::string *err) {
  c->on = false;
  if (!p.flag_biogem) return true;
  auto bad = [&](const std::string &m) { if (err) *err = m; return false; };
  if (!p.flag_atchem) return bad("flag_biogem without flag_atchem is outside the B200 hot path (air-sea exchange needs ATCHEM)");
  Namelist gm, bg, ac;
  if (!gm.load(jp(jobdir, "data_GEM"), err)) return false;
  if (!bg.load(jp(jobdir, "data_BIOGEM"), err)) return false;
  if (!ac.load(jp(jobdir, "data_ATCHEM"), err)) return false;
  // tracer selection must be exactly the supported set
  for (int id = 1; id <= 95; id++) {
    bool want = false;
    for (int q = 0; q < kNO; q++) want = want || kOcn[q].id == id;
    if (gm.flag(key_n("ocn_select", id), id <= 2) != want) return bad("data_GEM: ocean tracer selection differs from the supported BIOGEM configuration (ocn_select(" + std::to_string(id) + "))");
  }
  for (int id = 1; id <= 79; id++) {
    bool want = false;
    for (int q = 0; q < kNS; q++) want = want || kSed[q].id == id;
    if (gm.flag(key_n("sed_select", id), false) != want) return bad("data_GEM: particulate tracer selection differs from the supported BIOGEM configuration (sed_select(" + std::to_string(id) + "))");
  }
  for (int id = 1; id <= 19; id++) {
    bool want = false;
    for (int q = 0; q < kNA; q++) want = want || kAtm[q].id == id;
    if (gm.flag(key_n("atm_select", id), id <= 2) != want) return bad("data_GEM: atmosphere tracer selection differs from the supported BIOGEM configuration (atm_select(" + std::to_string(id) + "))");
  }
  if (p.maxl != kNO) return bad("dim_GOLDSTEINNTRACS must equal the number of selected ocean tracers (16)");
  if (g.K > kBgMaxK) return bad("BIOGEM path supports at most 16 levels");
  // options outside the restated path
  if (bg.str("par_bio_prodopt", "1N1T_PO4MM") != "1N1T_PO4MM") return bad("par_bio_prodopt: only 1N1T_PO4MM is on the B200 path");
  if (bg.str("opt_bio_CaCO3toPOCrainratio", "Ridgwelletal2007ab") != "Ridgwelletal2007ab") return bad("opt_bio_CaCO3toPOCrainratio: only Ridgwelletal2007ab is on the B200 path");
  if (bg.str("par_bio_remin_fun", "efolding") != "efolding") return bad("par_bio_remin_fun: only efolding is on the B200 path");
  if (gm.str("par_carbconstset_name", "Mehrbach") != "Mehrbach") return bad("par_carbconstset_name: only Mehrbach is on the B200 path");
  struct { const char *k; bool dflt; } flags[] = {{"ctrl_force_sed_closedsystem", true}, {"ctrl_force_GOLDSTEInTS", true},
      {"ctrl_force_windspeed", true}, {"ctrl_bio_remin_POC_fixed", true}, {"ctrl_bio_remin_CaCO3_fixed", true},
      {"ctrl_misc_Snorm", true}, {"ctrl_force_oldformat", true}, {"ctrl_force_GOLDSTEInTSonly", false},
      {"ctrl_force_seaice", false}, {"ctrl_bio_remin_POC_ballast", false}, {"ctrl_bio_remin_POC_kinetic", false},
      {"ctrl_bio_preformed", false}, {"ctrl_bio_CaCO3precip", false}, {"ctrl_misc_t_BP", false}, {"ctrl_continuing", false},
      {"ctrl_force_solconst", false}, {"ctrl_bio_remin_RDOM_photolysis", false}, {"ctrl_force_CaCO3toPOCrainratio", false}};
  for (auto &f : flags)
    if (bg.flag(f.k, f.dflt) != f.dflt) return bad(std::string("data_BIOGEM: ") + f.k + " differs from the value the B200 path implements");
  if (bg.num("par_misc_brinerejection_frac", 0.0) != 0.0) return bad("par_misc_brinerejection_frac != 0 is outside the B200 path");
  if (bg.num("par_misc_t_start", 0.0) != 0.0) return bad("par_misc_t_start != 0 is outside the B200 path");
  if (p.flag_biogem && p.atchem_radfor) return bad("atchem_radfor = y (CO2 feedback on the EMBM) is outside the B200 path");

  c->L = kNO; c->LS = kNS; c->LA = kNA;
  c->io.assign(c->L + 1, 0); c->otype = c->io; c->odep = c->io;
  c->is.assign(c->LS + 1, 0); c->stype = c->is; c->sdep_id = c->is; c->sdep_ls = c->is;
  c->ia.assign(c->LA + 1, 0); c->atype = c->ia; c->adep = c->ia;
  for (int l = 1; l <= c->L; l++) { c->io[l] = kOcn[l - 1].id; c->otype[l] = kOcn[l - 1].type; }
  for (int l = 1; l <= c->L; l++) c->odep[l] = find_id(c->io, kOcn[l - 1].dep);
  for (int l = 1; l <= c->LS; l++) { c->is[l] = kSed[l - 1].id; c->stype[l] = kSed[l - 1].type; c->sdep_id[l] = kSed[l - 1].dep; }
  for (int l = 1; l <= c->LS; l++) c->sdep_ls[l] = find_id(c->is, kSed[l - 1].dep);
  for (int l = 1; l <= c->LA; l++) { c->ia[l] = kAtm[l - 1].id; c->atype[l] = kAtm[l - 1].type; }
  for (int l = 1; l <= c->LA; l++) c->adep[l] = find_id(c->ia, kAtm[l - 1].dep);
  c->ocn_init.assign(c->L + 1, 0.0);
  c->atm_init.assign(c->LA + 1, 0.0);
  for (int l = 1; l <= c->L; l++) c->ocn_init[l] = bg.num(key_n("ocn_init", c->io[l]), 0.0);
  for (int l = 1; l <= c->LA; l++) c->atm_init[l] = ac.num(key_n("atm_init", c->ia[l]), 0.0);
  c->t_runtime = bg.num("par_misc_t_runtime", 1001.0);
  c->t_end = 0.0 + c->t_runtime;  // biogem.f90:232-236
  c->c0_PO4 = bg.num("par_bio_c0_PO4", c->c0_PO4);
  c->red_POP_POC = bg.num("par_bio_red_POP_POC", c->red_POP_POC);
  c->red_POP_PON = bg.num("par_bio_red_POP_PON", c->red_POP_PON);
  c->red_POP_PO2 = bg.num("par_bio_red_POP_PO2", c->red_POP_PO2);
  c->red_PON_ALK = bg.num("par_bio_red_PON_ALK", c->red_PON_ALK);
  c->red_DOMfrac = bg.num("par_bio_red_DOMfrac", c->red_DOMfrac);
  c->red_RDOMfrac = bg.num("par_bio_red_RDOMfrac", c->red_RDOMfrac);
  c->red_POC_CaCO3_pP = bg.num("par_bio_red_POC_CaCO3_pP", c->red_POC_CaCO3_pP);
  c->DOMlifetime = bg.num("par_bio_remin_DOMlifetime", c->DOMlifetime);
  c->POC_frac2 = bg.num("par_bio_remin_POC_frac2", c->POC_frac2);
  c->POC_eL2 = bg.num("par_bio_remin_POC_eL2", c->POC_eL2);
  c->POC_dfrac2 = bg.num("par_bio_remin_POC_dfrac2", c->POC_dfrac2);
  c->POC_c0frac2 = bg.num("par_bio_remin_POC_c0frac2", c->POC_c0frac2);
  c->CaCO3_frac2 = bg.num("par_bio_remin_CaCO3_frac2", c->CaCO3_frac2);
  c->CaCO3_eL1 = bg.num("par_bio_remin_CaCO3_eL1", c->CaCO3_eL1);
  c->CaCO3_eL2 = bg.num("par_bio_remin_CaCO3_eL2", c->CaCO3_eL2);
  c->sinkingrate_md = bg.num("par_bio_remin_sinkingrate", c->sinkingrate_md);
  c->remin_k_O2 = bg.num("par_bio_remin_k_O2", c->remin_k_O2);
  c->remin_c0_O2 = bg.num("par_bio_remin_c0_O2", c->remin_c0_O2);
  c->gastransfer_a = bg.num("par_gastransfer_a", c->gastransfer_a);
  c->d13C_DIC_Corg_ef = bg.num("par_d13C_DIC_Corg_ef", c->d13C_DIC_Corg_ef);
  c->Fgeothermal = bg.num("par_Fgeothermal", c->Fgeothermal);
  if (c->red_RDOMfrac != 0.0) return bad("par_bio_red_RDOMfrac != 0 needs RDOM tracers, which are outside the supported selection");
  if (c->CaCO3_eL1 < kBgNullSmall && c->CaCO3_eL2 < kBgNullSmall) return bad("saturation-dependent CaCO3 dissolution (eL1 = eL2 = 0) is outside the B200 path");

  // prescribed wind speed: rows j = maxj..1, i = 1..maxi per row (sub_load_data_ij, gem_util.f90:511-536)
  {
    const std::string f = jp(jp(jobdir, bg.str("par_indir_name", "input/biogem")), bg.str("par_windspeed_file", "windspeed.dat"));
    std::vector<double> v;
    if (!read_numbers(f, &v, err)) return false;
    if ((int)v.size() != g.I * g.J) return bad("wind speed file " + f + " has the wrong size");
    c->windspeed.assign(v.size(), 0.0);
    size_t q = 0;
    for (int j = g.J; j >= 1; j--)
      for (int i = 1; i <= g.I; i++) c->windspeed[(size_t)(j - 1) * g.I + (i - 1)] = v[q++];
  }
  // atmospheric forcing configuration, current "old format" (sub_init_tracer_forcing_atm, biogem_data.f90:1346-1357):
  //   ia  restore?  tconst  flux?  scale?  airsea_eqm?
  c->rst_sel.assign(c->LA + 1, 0);
  c->rst_tconst.assign(c->LA + 1, 1.0);
  c->rst_sig_t.assign(c->LA + 1, {});
  c->rst_sig_v.assign(c->LA + 1, {});
  c->rst_sig_i1.assign(c->LA + 1, 0);
  c->rst_sig_i2.assign(c->LA + 1, 0);
  c->rst_target.assign(c->LA + 1, 0.0);
  const std::string fordir = jp(jobdir, bg.str("par_fordir_name", "input/biogem/forcing"));
  {
    std::vector<std::string> lines;
    if (!data_lines(jp(fordir, "configure_forcings_atm.dat"), &lines, err)) return false;
    for (auto &ln : lines) {
      std::istringstream ss(ln);
      int ia = 0;
      std::string rs, fs, sc, eq;
      double tc = 1.0;
      if (!(ss >> ia >> rs >> tc >> fs >> sc >> eq)) return bad("configure_forcings_atm.dat: cannot parse '" + ln + "'");
      if (is_true(fs)) return bad("atmospheric flux forcing is outside the B200 path");
      if (is_true(eq)) return bad("ocnatm_airsea_eqm is outside the B200 path");
      const int la = find_id(c->ia, ia);
      if (!is_true(rs)) continue;
      if (!la || la < 3) return bad("restoring forcing of an unselected atmospheric tracer");
      if (tc < kBgNullSmall) return bad("restoring time constant must not be zero");
      c->rst_sel[la] = 1;
      c->rst_tconst[la] = tc;
    }
  }
  for (int la = 3; la <= c->LA; la++) {
    if (!c->rst_sel[la]) continue;
    const std::string base = jp(fordir, std::string("biogem_force_restore_atm_") + kAtm[la - 1].name);
    // _I / _II fields must be uniform 0 / 1 over the wet points (the only pattern the path supports)
    for (int which = 0; which < 2; which++) {
      std::vector<double> v;
      if (!read_numbers(base + (which ? "_II.dat" : "_I.dat"), &v, err)) return false;
      if ((int)v.size() != g.I * g.J) return bad(base + "_I/_II.dat has the wrong size");
      size_t q = 0;
      for (int j = g.J; j >= 1; j--)
        for (int i = 1; i <= g.I; i++, q++)
          if (g.k1at(i, j) <= g.K && v[q] != (which ? 1.0 : 0.0)) return bad("spatially varying atmospheric restoring fields are outside the B200 path");
    }
    std::vector<std::string> lines;
    if (!data_lines(base + "_sig.dat", &lines, err)) return false;
    std::vector<double> t, val;
    for (auto &ln : lines) {
      std::istringstream ss(ln);
      double a, b2;
      if (!(ss >> a >> b2)) return bad(base + "_sig.dat: cannot parse '" + ln + "'");
      t.push_back(1.0 * a);   // par_atm_force_scale_time = 1
      val.push_back(1.0 * b2); // par_atm_force_scale_val = 1
    }
    const int n = (int)t.size();
    if (n == 0) return bad("PLEASE PUT SOME DATA IN TIME SERIES FILE: " + base + "_sig.dat");
    // sub_load_data_t2, .NOT. ctrl_misc_t_BP (biogem_lib.f90:1467-1476)
    c->rst_sig_t[la].assign(n, 0.0);
    c->rst_sig_v[la].assign(n, 0.0);
    if (t[n - 1] <= t[0]) {
      for (int q = 0; q < n; q++) { c->rst_sig_t[la][q] = c->t_end - t[q]; c->rst_sig_v[la][q] = val[q]; }
    } else {
      for (int q = 0; q < n; q++) { c->rst_sig_t[la][q] = c->t_end - t[n - 1 - q]; c->rst_sig_v[la][q] = val[n - 1 - q]; }
    }
    c->rst_sig_i1[la] = n;
    c->rst_sig_i2[la] = n;
  }
  c->on = true;
  return true;
}

void bg_fill_tables(const BgConfig &c, const Params &p, const Grid &g, BgDev *b) {
  std::memset(b, 0, sizeof(*b));
  b->LS = c.LS; b->LA = c.LA;
  auto L = [&](int id) { return find_id(c.io, id); };
  auto S = [&](int id) { return find_id(c.is, id); };
  auto A = [&](int id) { return find_id(c.ia, id); };
  b->l_DIC = L(IO_DIC); b->l_DIC13 = L(IO_DIC_13C); b->l_DIC14 = L(IO_DIC_14C); b->l_PO4 = L(IO_PO4); b->l_O2 = L(IO_O2);
  b->l_ALK = L(IO_ALK); b->l_DOMC = L(IO_DOM_C); b->l_Ca = L(IO_CA); b->l_Mg = L(IO_MG);
  b->s_POC = S(IS_POC); b->s_POC13 = S(IS_POC_13C); b->s_POC14 = S(IS_POC_14C); b->s_POP = S(IS_POP); b->s_CaCO3 = S(IS_CACO3);
  b->s_CaCO313 = S(IS_CACO3_13C); b->s_CaCO314 = S(IS_CACO3_14C); b->s_POCf2 = S(IS_POC_FRAC2); b->s_CaCO3f2 = S(IS_CACO3_FRAC2);
  b->a_CO2 = A(IA_PCO2); b->a_CO213 = A(IA_PCO2_13C); b->a_CO214 = A(IA_PCO2_14C);
  for (int ls = 1; ls <= c.LS; ls++) { b->stype[ls] = c.stype[ls]; b->sdep_ls[ls] = c.sdep_ls[ls]; b->sdep_id[ls] = c.sdep_id[ls]; }
  for (int la = 1; la <= c.LA; la++) { b->atype[la] = c.atype[la]; b->aid[la] = c.ia[la]; b->adep[la] = c.adep[la]; }
  // conv_sed_ocn after sub_data_update_tracerrelationships without NO3 (gem_util.f90:66-98, biogem_data.f90:749-793)
  const double c_ALK_POP = c.red_PON_ALK * c.red_POP_PON;
  const double c_O2_POP = -4.0 / 2.0, c_O2_PON = 0.0;
  const double c_O2_POC = c.red_POP_PO2 / c.red_POP_POC - c_O2_POP / c.red_POP_POC - c_O2_PON * c.red_POP_PON / c.red_POP_POC;
  struct { int io, is; double v; } rel[] = {
      {IO_DIC, IS_POC, 1.0}, {IO_O2, IS_POC, c_O2_POC}, {IO_DIC_13C, IS_POC_13C, 1.0}, {IO_DIC_14C, IS_POC_14C, 1.0},
      {IO_PO4, IS_POP, 1.0}, {IO_O2, IS_POP, c_O2_POP}, {IO_ALK, IS_POP, c_ALK_POP},
      {IO_DIC, IS_CACO3, 1.0}, {IO_ALK, IS_CACO3, 2.0}, {IO_CA, IS_CACO3, 1.0},
      {IO_DIC_13C, IS_CACO3_13C, 1.0}, {IO_DIC_14C, IS_CACO3_14C, 1.0}};
  for (int ls = 1; ls <= c.LS; ls++) {
    b->n_ls_lo[ls] = 0;
    for (int l = 1; l <= c.L; l++)  // io ascending (fun_recalc_tracerrelationships_i, gem_util.f90:1396-1434)
      for (auto &r : rel)
        if (r.is == c.is[ls] && r.io == c.io[l] && std::fabs(r.v) > kBgNullSmall) {
          b->ls_lo[ls][b->n_ls_lo[ls]] = l;
          b->conv_ls_lo[ls][b->n_ls_lo[ls]] = r.v;
          b->n_ls_lo[ls]++;
          if (!b->lrem_slot[l]) b->lrem_slot[l] = ++b->n_lrem;
        }
  }
  const int dp[][2] = {{IO_DOM_C, IS_POC}, {IO_DOM_C_13C, IS_POC_13C}, {IO_DOM_C_14C, IS_POC_14C}, {IO_DOM_P, IS_POP}};
  for (auto &q : dp) { const int l = L(q[0]), ls = S(q[1]); if (l && ls) { b->dom2pom[l] = ls; b->pom2dom[ls] = l; } }
  const int ao[][2] = {{IA_PCO2, IO_DIC}, {IA_PCO2_13C, IO_DIC_13C}, {IA_PCO2_14C, IO_DIC_14C}, {IA_PO2, IO_O2},
                       {IA_PCFC11, IO_CFC11}, {IA_PCFC12, IO_CFC12}};
  for (auto &q : ao) { const int la = A(q[0]); if (la) b->atm2ocn[la] = L(q[1]); }
  for (int l = 1; l <= c.L; l++) b->lam_ocn[l] = (c.io[l] == IO_DIC_14C || c.io[l] == IO_DOM_C_14C) ? kBgLambda14C : 0.0;
  for (int ls = 1; ls <= c.LS; ls++) b->lam_sed[ls] = (c.is[ls] == IS_POC_14C || c.is[ls] == IS_CACO3_14C) ? kBgLambda14C : 0.0;
  for (int la = 1; la <= c.LA; la++) b->lam_atm[la] = (c.ia[la] == IA_PCO2_14C) ? kBgLambda14C : 0.0;
  // gem_data.f90:69-136
  const double Sc[][5] = {{IA_PCO2, 2073.1, 125.62, 3.6276, 0.043219}, {IA_PO2, 1953.4, 128.00, 3.9918, 0.050091},
                          {IA_PCFC11, 4039.8, 264.70, 8.2552, 0.103590}, {IA_PCFC12, 3713.2, 243.40, 7.5879, 0.095215}};
  const double Bu[][7] = {{IA_PCO2, -60.2409, 93.4517, 23.3585, 0.023517, -0.023656, 0.0047036},
                          {IA_PO2, -58.3877, 85.8079, 23.8439, -0.034892, 0.015568, -0.0019387},
                          {IA_PCFC11, -136.2685, 206.1150, 57.2805, -0.148598, 0.095114, -0.0163396},
                          {IA_PCFC12, -124.4395, 185.4299, 51.6383, -0.149779, 0.094668, -0.0160043}};
  for (int q = 0; q < 4; q++) {
    const int la = A((int)Sc[q][0]);
    if (!la) continue;
    for (int z = 0; z < 4; z++) b->Sc[la][z] = Sc[q][1 + z];
    for (int z = 0; z < 6; z++) b->bunsen[la][z] = Bu[q][1 + z];
  }
  b->c0_PO4 = c.c0_PO4; b->red_POP_POC = c.red_POP_POC; b->red_DOMfrac = c.red_DOMfrac; b->red_RDOMfrac = c.red_RDOMfrac;
  b->red_POC_CaCO3_pP = c.red_POC_CaCO3_pP; b->DOMlifetime = c.DOMlifetime; b->POC_frac2 = c.POC_frac2; b->POC_dfrac2 = c.POC_dfrac2;
  b->POC_c0frac2 = c.POC_c0frac2; b->CaCO3_frac2 = c.CaCO3_frac2;
  b->sinkingrate = c.sinkingrate_md / (1.0 / 365.25);   // par/conv_d_yr (biogem_data.f90:421)
  b->remin_k_O2 = c.remin_k_O2; b->remin_c0_O2 = c.remin_c0_O2; b->gastransfer_a = c.gastransfer_a;
  b->d13C_DIC_Corg_ef = c.d13C_DIC_Corg_ef; b->Fgeothermal = c.Fgeothermal; b->solar_constant = p.solconst; b->dsc = kDsc;
  b->dts = (double)(p.conv_kocn_kbiogem * p.kocn_loop) * p.genie_timestep;
  b->dtyr = b->dts / kBgYrS;
  b->dts_atchem = (double)(p.conv_kocn_katchem * p.kocn_loop) * p.genie_timestep;
  b->dtyr_atchem = b->dts_atchem / kBgYrS;
  for (int l = 1; l <= c.L; l++) b->fd_ocn[l] = std::exp(-b->dtyr * b->lam_ocn[l]);
  for (int ls = 1; ls <= c.LS; ls++) b->fd_sed[ls] = std::exp(-b->dtyr * b->lam_sed[ls]);
  for (int la = 1; la <= c.LA; la++) b->fd_atm[la] = std::exp(-b->dtyr_atchem * b->lam_atm[la]);
  for (int la = 3; la <= c.LA; la++) b->tmod[la] = c.rst_sel[la] ? 1.0 - std::exp(-b->dtyr / c.rst_tconst[la]) : 0.0;
  // sub_init_phys_ocn, biogem_data.f90:1098-1137
  {
    const int K = g.K;
    std::vector<double> dzl(K + 2, 0.0), dzal(K + 2, 0.0);
    for (int k = 1; k <= K; k++) { dzl[k] = g.dz[k]; dzal[k] = g.dza[k]; }
    dzal[K] = dzl[K] / 2.0;
    for (int k = 1; k <= K; k++) {
      double s = 0.0;
      b->dD[k] = kDsc * dzl[k];
      for (int kk = k; kk <= K; kk++) s = s + kDsc * dzl[kk];
      b->Dbot[k] = s;
      b->CaCO3_f1[k] = (1.0 - std::exp(-b->dD[k] / c.CaCO3_eL1));
      b->CaCO3_f2[k] = (1.0 - std::exp(-b->dD[k] / c.CaCO3_eL2));
      b->POC_f2[k] = (1.0 - std::exp(-b->dD[k] / c.POC_eL2));
    }
    b->Dmid_surf = kDsc * dzal[K];
    for (int k = 1; k <= K; k++) {   // phys_ocn(ipo_Dmid,i,j,k) = SUM(goldstein_dsc*loc_grid_dza(k:n_k)), biogem_data.f90:1120
      double s = 0.0;
      for (int kk = k; kk <= K; kk++) s = s + kDsc * dzal[kk];
      b->Dmid[k] = s;
    }
  }
}

// sub_update_sig, biogem_box.f90:3174-3218 (1-based indices into sig)
static void update_sig(double t, const std::vector<double> &sig, int *i1, int *i2, double *x) {
  if (*i1 > 1) {
    if (t < sig[*i1 - 1]) {
      for (;;) {
        *i1 = *i1 - 1;
        if (t > sig[*i1 - 1]) break;
        else if (*i1 == 1) break;
      }
    }
  }
  if (*i2 > 1) {
    if (t < sig[*i2 - 1]) {
      for (;;) {
        *i2 = *i2 - 1;
        if (t >= sig[*i2 - 1]) { *i2 = *i2 + 1; break; }
        else if (*i2 == 1) break;
      }
    }
  }
  if (std::fabs(sig[*i2 - 1] - sig[*i1 - 1]) > kBgNullSmall) *x = (sig[*i2 - 1] - t) / (sig[*i2 - 1] - sig[*i1 - 1]);
  else *x = 0.5;
}

void bg_forcing(BgConfig *c, long long clock_ms, BgDev *b) {
  const double t = c->t_runtime - (double)clock_ms / (1000.0 * kBgYrS);
  for (int la = 3; la <= c->LA; la++) {
    b->rst_active[la] = 0;
    b->rst_target[la] = 0.0;
    if (!c->rst_sel[la]) continue;
    double x;
    update_sig(t, c->rst_sig_t[la], &c->rst_sig_i1[la], &c->rst_sig_i2[la], &x);
    const double sx = (1 - x) * c->rst_sig_v[la][c->rst_sig_i2[la] - 1] + x * c->rst_sig_v[la][c->rst_sig_i1[la] - 1];
    const double f = 0.0 + sx * (1.0 - 0.0);   // I + sig_x*(II - I) at wet points
    if (c->atype[la] == 1) c->rst_target[la] = f;
    else c->rst_target[la] = bg_iso_fraction(f, c->atype[la] == 11 ? kBgStd13C : kBgStd14C) * c->rst_target[c->adep[la]];
    b->rst_target[la] = c->rst_target[la];
    b->rst_active[la] = (c->rst_sig_i1[la] != c->rst_sig_i2[la]) ? 1 : 0;
  }
}

}  // namespace cg
