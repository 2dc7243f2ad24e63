/* cgo_biogem.c -- CPU oracle: BIOGEM + ATCHEM on the tracer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates, for the frozen eb_go_gs_ac_bg configuration (DESIGN.md "BIOGEM configuration"), in the
 * reference's own loop and expression order:
 *   src/biogem/biogem.f90       initialise_biogem :24-527, step_biogem :528-1877, biogem_tracercoupling :1885-2077,
 *                               biogem_forcing :2083-2127, biogem_climate :2132-2239, biogem_climate_sol :2243-2263
 *   src/biogem/biogem_box.f90   sub_calc_solconst :27-40, sub_calc_pv :81-117, fun_calc_ocnatm_flux :123-299,
 *                               sub_calc_bio_uptake :346-1514 (1N1T_PO4MM path), sub_box_remin_redfield :2100-2208,
 *                               sub_box_remin_DOM :2287-2406, sub_box_remin_part :2412-2875, sub_update_sig :3174-3218,
 *                               sub_update_force_restore_atm :3333-3366, sub_biogem_copy_ocntots :3691-3739
 *   src/biogem/biogem_data.f90  sub_init_bio :579-624, sub_data_update_tracerrelationships :731-920,
 *                               sub_init_phys_ocn :1098-1137, sub_init_tracer_ocn_comp :1281-1309, sub_init_carb :2336-2430,
 *                               sub_init_force_restore_atm :2706-2791
 *   src/common/gem_carbchem.f90 sub_calc_carbconst :59-268, sub_adj_carbconst :277-297, sub_calc_carb :326-546,
 *                               sub_calc_carb_RF0 :553-669, sub_calc_carb_r13C/r14C :677-780, fits :790-1180,
 *                               fun_calc_solconst :1390-1429
 *   src/common/gem_util.f90     sub_def_tracerrelationships :27-274, sub_def_tracer_decay :280-308,
 *                               fun_calc_isotope_delta/fraction :568-617, fun_calc_rho :670-682
 *   src/common/gem_data.f90     Schmidt / Bunsen tables :69-136
 *   src/atchem/atchem.f90       step_atchem :63-158, cpl_comp_atmocn :252-264, cpl_flux_ocnatm :306-320
 *   src/atchem/atchem_data.f90  sub_init_phys_atm :195-229, sub_init_tracer_atm_comp :234-261
 * Tracer arrays are held in the compact (selected-tracer) index l / ls / la; the padded full-index arrays of the
 * reference (n_ocn=95, n_sed=79, n_atm=19) are not materialised.  Loops over "DO l=1,n_l_*" keep their order, so
 * every accumulation happens in the reference's sequence.
 * Parity unpinned (see cgo.h). */
#include "cgo_impl.h"

#define BG_PI 3.141592653589793   /* gem_cmn.f90:688 */
#define BG_REARTH 6.37e6          /* gem_cmn.f90:696 */
#define BG_M3_KG 1027.649         /* gem_cmn.f90:510 */
#define BG_ZEROC 273.15           /* gem_cmn.f90:690 */
#define BG_NULL (-0.999999e19)    /* gem_cmn.f90:717 */
#define BG_NULLSMALL 0.999999e-19 /* gem_cmn.f90:719 */
#define BG_YR_S (3600.0 * (24.0 * 365.25)) /* gem_cmn.f90:511-513 */
#define BG_ATM_MOL 1.7692e+020    /* gem_cmn.f90:509 */
#define BG_R 83.145               /* gem_cmn.f90:692 */
#define BG_R_SI 8.3145            /* gem_cmn.f90:694 */
#define BG_V 0.022414             /* gem_cmn.f90:698 */
#define BG_CP 4.1855              /* gem_cmn.f90:700 */
#define BG_CONC_MG 0.05282        /* gem_cmn.f90:705 */
#define BG_CONC_MGTOCA 5.155      /* gem_cmn.f90:708 */
#define BG_LAMBDA_14C (1. / 8267.0) /* gem_cmn.f90:648 */
#define BG_STD_13C 0.011202       /* gem_cmn.f90:630 */
#define BG_STD_14C 1.176e-12      /* gem_cmn.f90:631 */
#define BG_PA_ATM (1.0 / 1.01325e+05) /* gem_cmn.f90:526,545 */
#define BG_ATM_TH 7777.0          /* atchem_lib.f90:86 */

/* sed tracer types, gem_cmn.f90:321-330 */
enum { ST_BIO = 1, ST_ABIO = 2, ST_POM = 3, ST_CACO3 = 4, ST_OPAL = 5, ST_DET = 6, ST_SCAV = 7, ST_AGE = 8, ST_FRAC = 9 };
/* full-table tracer ids (data/main/tracer_define.{ocn,sed,atm}) */
enum { IO_T = 1, IO_S = 2, IO_DIC = 3, IO_DIC_13C = 4, IO_DIC_14C = 5, IO_PO4 = 8, IO_O2 = 10, IO_ALK = 12, IO_DOM_C = 15,
       IO_DOM_C_13C = 16, IO_DOM_C_14C = 17, IO_DOM_P = 20, IO_CA = 35, IO_CFC11 = 45, IO_CFC12 = 46, IO_MG = 50 };
enum { IS_POC = 3, IS_POC_13C = 4, IS_POC_14C = 5, IS_POP = 8, IS_CACO3 = 14, IS_CACO3_13C = 15, IS_CACO3_14C = 16,
       IS_POC_FRAC2 = 33, IS_CACO3_FRAC2 = 34 };
enum { IA_T = 1, IA_Q = 2, IA_PCO2 = 3, IA_PCO2_13C = 4, IA_PCO2_14C = 5, IA_PO2 = 6, IA_PCFC11 = 18, IA_PCFC12 = 19 };

#define BG_MAXL 24
#define BG_MAXLS 16
#define BG_MAXLA 12
/* carbonate constants, carb, isotope ratio slots (only what the path reads) */
enum { CC_K1, CC_K2, CC_K, CC_KB, CC_KW, CC_KSI, CC_KHF, CC_KHSO4, CC_KP1, CC_KP2, CC_KP3, CC_KH2S, CC_KNH4, CC_KCAL, CC_KARG,
       CC_QCO2, CC_QO2, N_CC };
enum { IC_H, IC_CO2, IC_CO3, IC_HCO3, IC_FUG, IC_OHM_CAL, IC_OHM_ARG, IC_DCO3_CAL, IC_DCO3_ARG, IC_RF0, N_IC };
enum { ICI_DIC_R13C, ICI_CO2_R13C, ICI_HCO3_R13C, ICI_CO3_R13C, ICI_DIC_R14C, ICI_CO2_R14C, ICI_HCO3_R14C, ICI_CO3_R14C, N_ICI };

struct cgo_bg {
  int L, LS, LA;
  int io[BG_MAXL + 1], otype[BG_MAXL + 1], odep[BG_MAXL + 1];        /* full id, type, dependency (compact index) */
  int is[BG_MAXLS + 1], stype[BG_MAXLS + 1], sdep[BG_MAXLS + 1];     /* sdep: FULL id of the dependency (as the Fortran tests it) */
  int sdep_ls[BG_MAXLS + 1];
  int ia[BG_MAXLA + 1], atype[BG_MAXLA + 1], adep[BG_MAXLA + 1];
  int l_DIC, l_DIC13, l_DIC14, l_PO4, l_O2, l_ALK, l_DOMC, l_DOMP, l_Ca, l_Mg;
  int s_POC, s_POC13, s_POC14, s_POP, s_CaCO3, s_CaCO313, s_CaCO314, s_POCf2, s_CaCO3f2;
  int a_CO2, a_CO213, a_CO214, a_O2;
  /* tracer relationships (compact) */
  double conv_ls_lo[BG_MAXLS + 1][BG_MAXL + 1];
  int n_ls_lo[BG_MAXLS + 1], ls_lo[BG_MAXLS + 1][6];   /* conv_ls_lo_i, io ascending */
  int dom2pom[BG_MAXL + 1];                            /* ocean l -> sed ls (conv_DOM_POM_i), 0 = none */
  int pom2dom[BG_MAXLS + 1];                           /* sed ls -> ocean l (conv_POM_DOM_i), 0 = none */
  int atm2ocn[BG_MAXLA + 1];                           /* conv_atm_ocn_i(1,ia) */
  double lam_ocn[BG_MAXL + 1], lam_sed[BG_MAXLS + 1], lam_atm[BG_MAXLA + 1];
  double Sc[BG_MAXLA + 1][4], bunsen[BG_MAXLA + 1][6];
  double ocn_init[BG_MAXL + 1], atm_init[BG_MAXLA + 1];
  /* parameters (biogem-defaults.nml) */
  double t_runtime, t_end, k0_PO4, c0_PO4, red_POP_POC, red_POP_PON, red_POP_PO2, red_PON_ALK, red_DOMfrac, red_RDOMfrac,
      red_POC_CaCO3, red_POC_CaCO3_pP, DOMlifetime, POC_frac2, POC_eL1, POC_eL2, POC_dfrac2, POC_c0frac2, CaCO3_frac2,
      CaCO3_eL1, CaCO3_eL2, sinkingrate, remin_k_O2, remin_c0_O2, remin_ci_O2, gastransfer_a, d13C_DIC_Corg_ef, Fgeothermal;
  int kbiogem, katchem;
  double genie_timestep;
  long long clock_ms;
  /* restoring forcing of the atmosphere (worjh2_preindustrial): select, time constant, 2-point signal */
  int rst_sel[BG_MAXLA + 1], rst_sig_i[BG_MAXLA + 1][2];
  double rst_tconst[BG_MAXLA + 1], rst_sig_t[BG_MAXLA + 1][2], rst_sig_v[BG_MAXLA + 1][2], rst_sig_x[BG_MAXLA + 1];
  double *rst_I, *rst_II, *rst_atm;                    /* [la][i][j] */
  /* state */
  double *bio_part, *bio_remin, *bio_settle, *focn;    /* [l|ls][i][j][k] */
  double *red;                                         /* bio_part_red [ls][ls][i][j] */
  double *carb, *carbisor;                             /* surface only: [N_IC|N_ICI][i][j] */
  double *seaice, *seaice_th, *wspeed, *solfor, *fxsw, *mld, *rho_surf, *A, *rA, *windspeed_file;
  double solar_constant;
  double *atm, *sfcatm1, *sfxatm1, *sfxsumatm, *atm_A, *atm_V;   /* [la][i][j] */
  double *sfcocn1, *sfxsed1, *focnatm;                 /* interface / diagnostics [l|ls|la][i][j] */
  double *sfxsumsed, *sfcsumocn, *sfxsumrok1;          /* SEDGEM / ROKGEM interface sums (sediment grid = ocean grid) */
  double *sig;                                         /* time-series integrals: t, tot_M, tot_M_sur, ocn(L), sur(L), ben(L), atm(LA) */
  double *sig2;                                        /* ... seaice, seaice_th, seaice_vol, opsi min / max, opsia min / max, SLT, fexport(LS), focnatm(LA), airsea(LA) */
  int sig_auto; double sig_ben_Dmin;                   /* cgo_run takes the diagnostic where genie.f90 does (after step_biogem) */
  /* time-slice diagnostics (diag_biogem_timeslice): 3-D carbonate system of every wet cell below the surface (the surface
   * cell's is carb / carbisor above, shared with step_biogem as in the reference) and the window integrals */
  double *carb3, *ciso3, *cc3;                         /* [N_IC|N_ICI|N_CC][i][j][k] (level K unused: see carb) */
  double *sl_ocn, *sl_part, *sl_carb, *sl_cc, *sl_ciso, *sl_t;   /* int_ocn / bio_part / carb / carbconst / carbisor _timeslice, int_t_timeslice */
  int slice_auto;
  double Dmid[64];
  double Dbot[64], dD[64], Dmid_surf;
  int go;
};

#define BG (o->bg)
#define OCN(l, i, j, k) o->bg_ocn[((l)-1) + NL * (((i)-1) + NI * (((j)-1) + NJ * ((k)-1)))]
#define DOCN(l, i, j, k) o->bg_vdocn[((l)-1) + NL * (((i)-1) + NI * (((j)-1) + NJ * ((k)-1)))]
#define PHM(i, j, k) o->bg_M[((i)-1) + NI * (((j)-1) + NJ * ((k)-1))]
#define PHRM(i, j, k) o->bg_rM[((i)-1) + NI * (((j)-1) + NJ * ((k)-1))]
#define PHV(i, j, k) o->bg_V[((i)-1) + NI * (((j)-1) + NJ * ((k)-1))]
#define REMIN(l, i, j, k) BG->bio_remin[((l)-1) + NL * (((i)-1) + NI * (((j)-1) + NJ * ((k)-1)))]
#define FOCN(l, i, j, k) BG->focn[((l)-1) + NL * (((i)-1) + NI * (((j)-1) + NJ * ((k)-1)))]
#define PART(ls, i, j, k) BG->bio_part[((ls)-1) + BG->LS * (((i)-1) + NI * (((j)-1) + NJ * ((k)-1)))]
#define SETTLE(ls, i, j, k) BG->bio_settle[((ls)-1) + BG->LS * (((i)-1) + NI * (((j)-1) + NJ * ((k)-1)))]
#define RED(a, b, i, j) BG->red[((a)-1) + BG->LS * (((b)-1) + BG->LS * (((i)-1) + NI * ((j)-1)))]
#define CARB(c, i, j) BG->carb[(c) + N_IC * (((i)-1) + NI * ((j)-1))]
#define CISO(c, i, j) BG->carbisor[(c) + N_ICI * (((i)-1) + NI * ((j)-1))]
#define ATM(la, i, j) BG->atm[((la)-1) + BG->LA * (((i)-1) + NI * ((j)-1))]
#define SFCATM1(la, i, j) BG->sfcatm1[((la)-1) + BG->LA * (((i)-1) + NI * ((j)-1))]
#define SFXATM1(la, i, j) BG->sfxatm1[((la)-1) + BG->LA * (((i)-1) + NI * ((j)-1))]
#define SFXSUMATM(la, i, j) BG->sfxsumatm[((la)-1) + BG->LA * (((i)-1) + NI * ((j)-1))]
#define RSTATM(a, la, i, j) (a)[((la)-1) + BG->LA * (((i)-1) + NI * ((j)-1))]

/* ------------------------------------------------------------------------------------------ gem_util helpers */
static double iso_delta(double tot, double iso, double standard, int allow_negative, double nullv) { /* gem_util.f90:568-598 */
  if (((fabs(tot) > BG_NULLSMALL) && allow_negative) || (tot > BG_NULLSMALL)) {
    const double fr = iso / tot;
    if ((1.0 - fr) > BG_NULLSMALL) {
      const double R = fr / (1.0 - fr);
      return 1000.0 * (R / standard - 1.0);
    }
    return nullv;
  }
  return nullv;
}
static double iso_fraction(double delta, double standard) { /* gem_util.f90:604-617 */
  const double R = standard * (1.0 + delta / 1000.0);
  return R / (1.0 + R);
}
static double calc_rho(double T, double S) { /* gem_util.f90:670-682 */
  const double TC = T - BG_ZEROC;
  return 1000.0 + (0.7968 * S - 0.0559 * TC - 0.0063 * (TC * TC) + 3.7315E-05 * powi_(TC, 3));
}

/* ------------------------------------------------------------------------------------------ gem_carbchem */
static const double dpH2CO3[5] = {-2.550E+1, +1.271E-1, +0.000E+0, -3.080E+0, +8.770E-2};
static const double dpHCO3[5] = {-1.582E+1, -2.190E-2, +0.000E+0, +1.130E+0, -1.475E-1};
static const double dpBO3H3[5] = {-2.948E+1, +1.622E-1, +2.608E-3, -2.840E+0, +0.000E+0};
static const double dpH2O[5] = {-2.002E+1, +1.119E-1, -1.409E-3, -5.130E+0, +7.940E-2};
static const double dpHF[5] = {-9.780E+0, -9.000E-3, -9.420E-4, -3.910E+0, +5.400E-2};
static const double dpHSO4[5] = {-1.803E+1, +4.660E-2, +3.160E-4, -4.530E+0, +9.000E-2};
static const double dpH4SiO4[5] = {-2.948E+1, +1.622E-1, +2.608E-3, -2.840E+0, +0.000E+0};
static const double dpH3PO4[5] = {-1.451E+1, +1.211E-1, -3.210E-4, -2.670E+0, +4.270E-2};
static const double dpH2PO4[5] = {-2.312E+1, +1.758E-1, -2.647E-3, -5.150E+0, +9.000E-2};
static const double dpHPO4[5] = {-2.657E+1, +2.020E-1, -3.042E-3, -4.080E+0, +7.140E-2};
static const double dpCaCO3cal[5] = {-4.876E+1, +5.304E-1, +0.000E+0, -1.176E+1, +3.692E-1};
static const double dpCaCO3arg[5] = {-4.596E+1, +5.304E-1, +0.000E+0, -1.176E+1, +3.692E-1};
static const double dpH2S[5] = {-1.107E+1, +9.000E-3, -9.420E-4, +2.890E+0, +5.400E-2};
static const double dpNH4[5] = {-2.643E+0, +8.890E-1, -9.050E-3, -5.030E+0, +8.140E-2};

static double corr_p(double TC, double P, double rRT, const double *dp) { /* gem_carbchem.f90:1175-1186 */
  return (-(dp[0] + dp[1] * TC + dp[2] * TC * TC) + (5.0E-4 * (dp[3] + dp[4] * TC)) * P) * P * rRT;
}
static double f_Btot(double S) { double v = 0.000416 * S / 35.0; if (v < BG_NULLSMALL) v = BG_NULLSMALL; return v; }
static double f_Ftot(double S) { double v = 0.00007 * S / 35.0; if (v < BG_NULLSMALL) v = BG_NULLSMALL; return v; }
static double f_SO4tot(double S) { double v = 0.02824 * S / 35.0; if (v < BG_NULLSMALL) v = BG_NULLSMALL; return v; }

/* sub_calc_carbconst, Mehrbach set (par_carbconstset_name default), gem_carbchem.f90:59-268 */
static void calc_carbconst(double D, double T_in, double S_in, double *cc) {
  double T = T_in, S = S_in;
  if (T < (BG_ZEROC + 2.0)) T = BG_ZEROC + 2.0;
  if (T > (BG_ZEROC + 35.0)) T = BG_ZEROC + 35.0;
  if (S < 26.0) S = 26.0;
  if (S > 43.0) S = 43.0;
  const double P = D / 10.0;
  const double S_p05 = pow(S, 0.5), S_p15 = pow(S, 1.5), S_p20 = S * S;
  const double T_ln = log(T), T_log = log10(T), rT = 1.0 / T, Tr100 = T / 100.0, TC = T - BG_ZEROC;
  const double rRT = 1.0 / (BG_R * T);
  const double I = (S > BG_NULLSMALL) ? 19.924 * S / (1000.0 - 1.005 * S) : BG_NULLSMALL;
  const double I_p05 = pow(I, 0.5), I_p15 = pow(I, 1.5), I_p20 = I * I;
  double Cl = S_in / 1.80655;
  if (Cl < BG_NULLSMALL) Cl = BG_NULLSMALL;
  const double ION = (Cl > BG_NULLSMALL) ? 0.00147 + 0.03592 * Cl + 0.000068 * Cl * Cl : BG_NULLSMALL;
  const double ION_p05 = pow(ION, 0.5);
  const double m2c = log(1 - 0.001005 * S);
  const double SO4tot = f_SO4tot(S), Ftot = f_Ftot(S);
  const double lnkHSO4 = 141.328 - 4276.1 * rT - 23.093 * T_ln + (324.57 - 13856.0 * rT - 47.986 * T_ln) * I_p05 +
                         (-771.54 + 35474.0 * rT + 114.723 * T_ln) * I - 2698.0 * rT * I_p15 + 1776.0 * rT * I_p20;
  const double lnkHF = 1590.2 / T - 12.641 + 1.525 * ION_p05;
  cc[CC_KHSO4] = exp(lnkHSO4 + m2c);
  const double f2t = log(1.0 + SO4tot / cc[CC_KHSO4]);
  cc[CC_KHF] = exp(lnkHF + m2c + f2t);
  const double f2s = log(1.0 + SO4tot / cc[CC_KHSO4] + Ftot / cc[CC_KHF]);
  const double t2s = -f2t + f2s;
  cc[CC_K1] = exp(log(pow(10.0, -(3670.7 * rT - 62.008 + 9.7944 * T_ln - 0.0118 * S + 0.000116 * S_p20))) + corr_p(TC, P, rRT, dpH2CO3));
  cc[CC_K2] = exp(log(pow(10.0, -(1394.7 * rT + 4.777 - 0.0184 * S + 0.000118 * S_p20))) + corr_p(TC, P, rRT, dpHCO3));
  cc[CC_K] = cc[CC_K1] / cc[CC_K2];
  cc[CC_KB] = exp((148.0248 + 137.194 * S_p05 + 1.62247 * S +
                   (-8966.90 - 2890.51 * S_p05 - 77.942 * S + 1.726 * S_p15 - 0.0993 * S_p20) * rT +
                   (-24.4344 - 25.085 * S_p05 - 0.2474 * S) * T_ln + 0.053105 * S_p05 * T) +
                  m2c + t2s + corr_p(TC, P, rRT, dpBO3H3));
  cc[CC_KW] = exp((148.9802 - 13847.26 * rT - 23.6521 * T_ln + (-5.977 + 118.67 * rT + 1.0495 * T_ln) * S_p05 - 0.01615 * S) +
                  corr_p(TC, P, rRT, dpH2O));
  cc[CC_KSI] = exp((117.40 - 8904.2 * rT - 19.334 * T_ln + (3.5913 - 458.79 * rT) * I_p05 + (-1.5998 + 188.74 * rT) * I +
                    (0.07871 - 12.1652 * rT) * I * I) +
                   m2c + corr_p(TC, P, rRT, dpH4SiO4));
  cc[CC_KHF] = exp(lnkHF + m2c + f2s + corr_p(TC, P, rRT, dpHF));
  cc[CC_KHSO4] = exp(lnkHSO4 + m2c + f2s + corr_p(TC, P, rRT, dpHSO4));
  cc[CC_KP1] = exp((115.54 - 4576.752 / T - 18.453 * T_ln + (0.69171 - 106.736 / T) * S_p05 + (-0.01844 - 0.65643 / T) * S) +
                   corr_p(TC, P, rRT, dpH3PO4));
  cc[CC_KP2] = exp((172.1033 - 8814.715 / T - 27.927 * T_ln + (1.3566 - 160.340 / T) * S_p05 + (-0.05778 + 0.37335 / T) * S) +
                   corr_p(TC, P, rRT, dpH2PO4));
  cc[CC_KP3] = exp((-18.126 - 3070.75 / T + (2.81197 + 17.27039 / T) * S_p05 + (-0.09984 - 44.99486 / T) * S) +
                   corr_p(TC, P, rRT, dpHPO4));
  cc[CC_KH2S] = exp((225.838 - 13275.3 * rT - 34.6435 * T_ln + 0.3449 * S_p05 - 0.0274 * S) + t2s + corr_p(TC, P, rRT, dpH2S));
  cc[CC_KNH4] = exp((-6285.33 * rT + 0.0001635 * T - 0.25444 + (0.46532 - 123.7184 * rT) * S_p05 + (-0.01992 + 3.17556 * rT) * S) +
                    corr_p(TC, P, rRT, dpNH4));
  cc[CC_KCAL] = exp(corr_p(TC, P, rRT, dpCaCO3cal)) *
                pow(10.0, (-171.9065 - 0.077993 * T + 2839.319 * rT + 71.595 * T_log +
                           (-0.77712 + 0.0028426 * T + 178.34 * rT) * S_p05 - 0.07711 * S + 0.0041249 * S_p15));
  cc[CC_KARG] = exp(corr_p(TC, P, rRT, dpCaCO3arg)) *
                pow(10.0, (-171.945 - 0.077993 * T + 2903.293 * rT + 71.595 * T_log +
                           (-0.068393 + 0.0017276 * T + 88.135 * rT) * S_p05 - 0.10018 * S + 0.0059415 * S_p15));
  cc[CC_QCO2] = exp(-60.2409 + 93.4517 * (100 * rT) + 23.3585 * log(Tr100) +
                    S * (0.023517 - 0.023656 * (Tr100) + 0.0047036 * (Tr100 * Tr100)));
  cc[CC_QO2] = exp(-173.9894 + 255.5907 * (100.0 * rT) + 146.4813 * log(Tr100) - 22.2040 * (Tr100) +
                   S * (-0.037362 + 0.016504 * (Tr100) - 0.0020564 * (Tr100 * Tr100)) - log(1.0E6) - log(0.20946));
}
static void adj_carbconst(double Ca, double Mg, double *cc) { /* gem_carbchem.f90:277-297 */
  const double alpha = 3.655E-8;
  double ratio = 1.0;
  if (Ca > BG_NULLSMALL) ratio = Mg / Ca;
  cc[CC_KCAL] = cc[CC_KCAL] - alpha * (BG_CONC_MGTOCA - ratio);
  cc[CC_K1] = (1.0 + 0.155 * (Mg - BG_CONC_MG) / BG_CONC_MG) * cc[CC_K1];
  cc[CC_K2] = (1.0 + 0.422 * (Mg - BG_CONC_MG) / BG_CONC_MG) * cc[CC_K2];
}
/* one pass of the implicit [H] loop body shared by sub_calc_carb and sub_calc_carb_RF0 (gem_carbchem.f90:352-424, 578-640) */
static void carb_iter(double DIC, double ALK, double PO4tot, double SiO2tot, double Btot, double SO4tot, double Ftot,
                      const double *cc, double H, double *co2, double *co3, double *hco3, double *H1, double *H2) {
  const double H_p2 = H * H, H_p3 = H * H_p2;
  const double OH = cc[CC_KW] / H;
  const double H4BO4 = Btot / (1.0 + H / cc[CC_KB]);
  const double H3SiO4 = SiO2tot / (1.0 + H / cc[CC_KSI]);
  const double HSO4 = SO4tot / (1.0 + cc[CC_KHSO4] / H);
  const double HF = Ftot / (1.0 + cc[CC_KHF] / H);
  const double HS = 0.0, NH3 = 0.0; /* H2S, NH4 not selected: totals are 0 (< const_real_nullsmall) */
  const double H3PO4 = PO4tot / (1.0 + cc[CC_KP1] / H + (cc[CC_KP1] * cc[CC_KP2]) / H_p2 + (cc[CC_KP1] * cc[CC_KP2] * cc[CC_KP3]) / H_p3);
  const double HPO4 = PO4tot / (1.0 + H / cc[CC_KP2] + H_p2 / (cc[CC_KP1] * cc[CC_KP2]) + cc[CC_KP3] / H);
  const double PO4 = PO4tot / (1.0 + H / cc[CC_KP3] + H_p2 / (cc[CC_KP2] * cc[CC_KP3]) + H_p3 / (cc[CC_KP1] * cc[CC_KP2] * cc[CC_KP3]));
  const double ALK_DIC = ALK - H4BO4 - OH - HPO4 - 2.0 * PO4 - H3SiO4 - NH3 - HS + H + HSO4 + HF + H3PO4;
  const double k = cc[CC_K];
  const double a = 4.0 * ALK_DIC + DIC * k - ALK_DIC * k;
  const double zed = pow(a * a + 4.0 * (k - 4.0) * (ALK_DIC * ALK_DIC), 0.5);
  *hco3 = (DIC * k - zed) / (k - 4.0);
  *co3 = (ALK_DIC * k - DIC * k - 4.0 * ALK_DIC + zed) / (2.0 * (k - 4.0));
  *co2 = DIC - ALK_DIC + (ALK_DIC * k - DIC * k - 4.0 * ALK_DIC + zed) / (2.0 * (k - 4.0));
  *H1 = cc[CC_K1] * *co2 / *hco3;
  *H2 = cc[CC_K2] * *hco3 / *co3;
}
/* sub_calc_carb; returns 0 on success, 1 when the reference would raise error_stop (ctrl_carbchem_fail) */
static int calc_carb(double DIC, double ALK, double Ca, double PO4tot, double SiO2tot, double Btot, double SO4tot, double Ftot,
                     const double *cc, double *carb) {
  int n = 1;
  double H = carb[IC_H], H_old, co2, co3, hco3, H1, H2;
  for (;;) {
    H_old = H;
    carb_iter(DIC, ALK, PO4tot, SiO2tot, Btot, SO4tot, Ftot, cc, H, &co2, &co3, &hco3, &H1, &H2);
    if ((H1 < BG_NULLSMALL) || (H2 < BG_NULLSMALL)) return 1;
    H = sqrt(H1 * H2);
    if (fabs(1.0 - H / H_old) < (1.0E-8 / H) * 0.001) {
      carb[IC_CO2] = co2; carb[IC_CO3] = co3; carb[IC_HCO3] = hco3;
      carb[IC_FUG] = co2 / cc[CC_QCO2];
      carb[IC_OHM_CAL] = Ca * co3 / cc[CC_KCAL];
      carb[IC_OHM_ARG] = Ca * co3 / cc[CC_KARG];
      carb[IC_H] = H;
      carb[IC_DCO3_CAL] = co3 - cc[CC_KCAL] * 1.0 / Ca;
      carb[IC_DCO3_ARG] = co3 - cc[CC_KARG] * 1.0 / Ca;
      return 0;
    }
    n = n + 1;
    if (n > 100) return 1;
  }
}
static void calc_carb_RF0(double DIC, double ALK, double PO4tot, double SiO2tot, double Btot, double SO4tot, double Ftot,
                          const double *cc, double *carb) { /* gem_carbchem.f90:553-669 */
  int n = 1;
  double H = carb[IC_H], H_old, co2, co3, hco3, H1, H2;
  const double DIC_RF0 = DIC + 1.0e-6;
  for (;;) {
    H_old = H;
    carb_iter(DIC_RF0, ALK, PO4tot, SiO2tot, Btot, SO4tot, Ftot, cc, H, &co2, &co3, &hco3, &H1, &H2);
    H = sqrt(H1 * H2);
    if (fabs(1.0 - H / H_old) < 0.001) {
      carb[IC_RF0] = (co2 / carb[IC_CO2] - 1.0) / (DIC_RF0 / DIC - 1.0);
      return;
    }
    n = n + 1;
    if ((H1 < BG_NULLSMALL) || (H2 < BG_NULLSMALL) || (n > 100)) { carb[IC_RF0] = 0.0; return; }
  }
}
/* sub_calc_carb_r13C (m = 1.0, 13C standard) / sub_calc_carb_r14C (m = 2.0, 14C standard), gem_carbchem.f90:677-780 */
static void calc_carb_riso(double T, double DIC, double DICiso, const double *carb, double m, double standard, double *out4) {
  const double TC = T - BG_ZEROC;
  double d = iso_delta(DIC, DICiso, standard, 0, BG_NULL);
  double e_bg, e_dg, e_cg;
  if (m == 1.0) { e_bg = -0.1141 * TC + 10.78; e_dg = +0.0049 * TC - 1.31; e_cg = -0.052 * TC + 7.22; }
  else { e_bg = 2.0 * (-0.1141 * TC + 10.78); e_dg = 2.0 * (+0.0049 * TC - 1.31); e_cg = 2.0 * (-0.052 * TC + 7.22); }
  const double e_cb = e_cg - e_bg / (1.0 + e_bg * 1.0E-3);
  const double e_db = e_dg - e_bg / (1.0 + e_bg * 1.0E-3);
  const double dHCO3 = (d * DIC - (e_db * carb[IC_CO2] + e_cb * carb[IC_CO3])) /
                       ((1.0 + e_db * 1.0E-3) * carb[IC_CO2] + carb[IC_HCO3] + (1.0 + e_cb * 1.0E-3) * carb[IC_CO3]);
  const double dCO2 = e_db + dHCO3 * (1.0 + e_db * 1.0E-3);
  const double dCO3 = e_cb + dHCO3 * (1.0 + e_cb * 1.0E-3);
  const double rCO2 = iso_fraction(dCO2, standard), rHCO3 = iso_fraction(dHCO3, standard), rCO3 = iso_fraction(dCO3, standard);
  const double rDIC = (rCO2 * carb[IC_CO2] + rHCO3 * carb[IC_HCO3] + rCO3 * carb[IC_CO3]) / DIC;
  out4[0] = rDIC; out4[1] = rCO2; out4[2] = rHCO3; out4[3] = rCO3;
}
static double calc_solconst(const struct cgo_bg *b, int la, double T_in, double S_in, double rho) { /* gem_carbchem.f90:1390-1429 */
  double T, S;
  if (T_in < BG_ZEROC + 2.0) T = BG_ZEROC + 2.0; else if (T_in > (BG_ZEROC + 35.0)) T = BG_ZEROC + 35.0; else T = T_in;
  if (S_in < 26.0) S = 26.0; else if (S_in > 43.0) S = 43.0; else S = S_in;
  const double rT = 1.0 / T, Tr100 = T / 100.0;
  const double *c = b->bunsen[la];
  const double e = exp(c[0] + c[1] * (100 * rT) + c[2] * log(Tr100) + S * (c[3] + c[4] * (Tr100) + c[5] * (Tr100 * Tr100)));
  const int ia = b->ia[la];
  if (ia == IA_PCO2 || ia == IA_PCFC11 || ia == IA_PCFC12) return e;
  return e / (rho * BG_V);
}

/* ------------------------------------------------------------------------------------------ check-value hooks
 * (tests/test_oracle_checkvalues.py: the chemistry above against values printed in the papers it is taken from)
 * cc: N_CC = 17 constants in the order of the enum above; carb: N_IC = 10 values, carb[0] = [H+] seed in, solution out. */
void cgo_test_carbconst(double D, double T, double S, double *cc) { calc_carbconst(D, T, S, cc); }
int cgo_test_calc_carb(double DIC, double ALK, double Ca, double PO4tot, double SiO2tot, double S, const double *cc, double *carb) {
  return calc_carb(DIC, ALK, Ca, PO4tot, SiO2tot, f_Btot(S), f_SO4tot(S), f_Ftot(S), cc, carb);
}
double cgo_test_rho(double T, double S) { return calc_rho(T, S); }
double cgo_test_iso_delta(double tot, double iso, double standard) { return iso_delta(tot, iso, standard, 0, BG_NULL); }
double cgo_test_iso_fraction(double delta, double standard) { return iso_fraction(delta, standard); }

void cgo_biogem_climate(cgo_t *o);
/* ------------------------------------------------------------------------------------------ set-up */
static double bg_par(const char *params, const char *key, double dflt) {
  const size_t kl = strlen(key);
  const char *p = params;
  double v = dflt;
  while (p && *p) {
    const char *e = strchr(p, '\n');
    if (!strncmp(p, key, kl) && p[kl] == '=') v = strtod(p + kl + 1, NULL);
    p = e ? e + 1 : NULL;
  }
  return v;
}
static double *bg_alloc(cgo_t *o, const char *name, long n) { return cgo_alloc(o, name, n); }

static void bg_tables(cgo_t *o) {
  struct cgo_bg *b = BG;
  static const int ocn_sel[][3] = {{IO_T, IO_T, 0}, {IO_S, IO_S, 0}, {IO_DIC, IO_DIC, 1}, {IO_DIC_13C, IO_DIC, 11},
      {IO_DIC_14C, IO_DIC, 12}, {IO_PO4, IO_PO4, 1}, {IO_O2, IO_O2, 1}, {IO_ALK, IO_ALK, 1}, {IO_DOM_C, IO_DOM_C, 1},
      {IO_DOM_C_13C, IO_DOM_C, 11}, {IO_DOM_C_14C, IO_DOM_C, 12}, {IO_DOM_P, IO_DOM_P, 1}, {IO_CA, IO_CA, 1},
      {IO_CFC11, IO_CFC11, 1}, {IO_CFC12, IO_CFC12, 1}, {IO_MG, IO_MG, 1}};
  static const int sed_sel[][3] = {{IS_POC, IS_POC, 1}, {IS_POC_13C, IS_POC, 11}, {IS_POC_14C, IS_POC, 12}, {IS_POP, IS_POP, 3},
      {IS_CACO3, IS_CACO3, 1}, {IS_CACO3_13C, IS_CACO3, 11}, {IS_CACO3_14C, IS_CACO3, 12}, {IS_POC_FRAC2, IS_POC_FRAC2, 9},
      {IS_CACO3_FRAC2, IS_CACO3_FRAC2, 9}};
  static const int atm_sel[][3] = {{IA_T, IA_T, 0}, {IA_Q, IA_Q, 0}, {IA_PCO2, IA_PCO2, 1}, {IA_PCO2_13C, IA_PCO2, 11},
      {IA_PCO2_14C, IA_PCO2, 12}, {IA_PO2, IA_PO2, 1}, {IA_PCFC11, IA_PCFC11, 1}, {IA_PCFC12, IA_PCFC12, 1}};
  int l, ls, la, m;
  b->L = 16; b->LS = 9; b->LA = 8;
  if (NL != b->L) { fprintf(stderr, "cgo_biogem: the frozen BIOGEM configuration needs maxl = 16\n"); abort(); }
  for (l = 1; l <= b->L; l++) { b->io[l] = ocn_sel[l - 1][0]; b->otype[l] = ocn_sel[l - 1][2]; }
  for (l = 1; l <= b->L; l++) for (m = 1; m <= b->L; m++) if (b->io[m] == ocn_sel[l - 1][1]) b->odep[l] = m;
  for (ls = 1; ls <= b->LS; ls++) { b->is[ls] = sed_sel[ls - 1][0]; b->sdep[ls] = sed_sel[ls - 1][1]; b->stype[ls] = sed_sel[ls - 1][2]; }
  for (ls = 1; ls <= b->LS; ls++) for (m = 1; m <= b->LS; m++) if (b->is[m] == b->sdep[ls]) b->sdep_ls[ls] = m;
  for (la = 1; la <= b->LA; la++) { b->ia[la] = atm_sel[la - 1][0]; b->atype[la] = atm_sel[la - 1][2]; }
  for (la = 1; la <= b->LA; la++) for (m = 1; m <= b->LA; m++) if (b->ia[m] == atm_sel[la - 1][1]) b->adep[la] = m;
#define LFIND(id) ({ int r_ = 0, q_; for (q_ = 1; q_ <= b->L; q_++) if (b->io[q_] == (id)) r_ = q_; r_; })
#define SFIND(id) ({ int r_ = 0, q_; for (q_ = 1; q_ <= b->LS; q_++) if (b->is[q_] == (id)) r_ = q_; r_; })
#define AFIND(id) ({ int r_ = 0, q_; for (q_ = 1; q_ <= b->LA; q_++) if (b->ia[q_] == (id)) r_ = q_; r_; })
  b->l_DIC = LFIND(IO_DIC); b->l_DIC13 = LFIND(IO_DIC_13C); b->l_DIC14 = LFIND(IO_DIC_14C); b->l_PO4 = LFIND(IO_PO4);
  b->l_O2 = LFIND(IO_O2); b->l_ALK = LFIND(IO_ALK); b->l_DOMC = LFIND(IO_DOM_C); b->l_DOMP = LFIND(IO_DOM_P);
  b->l_Ca = LFIND(IO_CA); b->l_Mg = LFIND(IO_MG);
  b->s_POC = SFIND(IS_POC); b->s_POC13 = SFIND(IS_POC_13C); b->s_POC14 = SFIND(IS_POC_14C); b->s_POP = SFIND(IS_POP);
  b->s_CaCO3 = SFIND(IS_CACO3); b->s_CaCO313 = SFIND(IS_CACO3_13C); b->s_CaCO314 = SFIND(IS_CACO3_14C);
  b->s_POCf2 = SFIND(IS_POC_FRAC2); b->s_CaCO3f2 = SFIND(IS_CACO3_FRAC2);
  b->a_CO2 = AFIND(IA_PCO2); b->a_CO213 = AFIND(IA_PCO2_13C); b->a_CO214 = AFIND(IA_PCO2_14C); b->a_O2 = AFIND(IA_PO2);
  /* conv_sed_ocn after sub_data_update_tracerrelationships (no NO3): gem_util.f90:66-98 + biogem_data.f90:749-793 */
  {
    const double c_ALK_POP = b->red_PON_ALK * b->red_POP_PON;
    const double c_O2_POP = -4.0 / 2.0, c_O2_PON = 0.0;
    const double c_O2_POC = b->red_POP_PO2 / b->red_POP_POC - c_O2_POP / b->red_POP_POC - c_O2_PON * b->red_POP_PON / b->red_POP_POC;
    struct { int io, is; double v; } rel[] = {
        {IO_DIC, IS_POC, 1.0}, {IO_O2, IS_POC, c_O2_POC}, {IO_DIC_13C, IS_POC_13C, 1.0}, {IO_DIC_14C, IS_POC_14C, 1.0},
        {IO_PO4, IS_POP, 1.0}, {IO_O2, IS_POP, c_O2_POP}, {IO_ALK, IS_POP, c_ALK_POP},
        {IO_DIC, IS_CACO3, 1.0}, {IO_ALK, IS_CACO3, 2.0}, {IO_CA, IS_CACO3, 1.0},
        {IO_DIC_13C, IS_CACO3_13C, 1.0}, {IO_DIC_14C, IS_CACO3_14C, 1.0}};
    for (ls = 1; ls <= b->LS; ls++) {
      b->n_ls_lo[ls] = 0;
      for (l = 1; l <= b->L; l++)  /* io ascending: fun_recalc_tracerrelationships_i, gem_util.f90:1396-1434 */
        for (m = 0; m < (int)(sizeof rel / sizeof rel[0]); m++)
          if (rel[m].is == b->is[ls] && rel[m].io == b->io[l] && fabs(rel[m].v) > BG_NULLSMALL) {
            b->conv_ls_lo[ls][l] = rel[m].v;
            b->ls_lo[ls][b->n_ls_lo[ls]++] = l;
          }
    }
  }
  /* conv_DOM_POM / conv_POM_DOM (gem_util.f90:225-247), all 1.0 */
  {
    static const int dp[][2] = {{IO_DOM_C, IS_POC}, {IO_DOM_C_13C, IS_POC_13C}, {IO_DOM_C_14C, IS_POC_14C}, {IO_DOM_P, IS_POP}};
    for (m = 0; m < 4; m++) { l = LFIND(dp[m][0]); ls = SFIND(dp[m][1]); if (l && ls) { b->dom2pom[l] = ls; b->pom2dom[ls] = l; } }
  }
  /* conv_atm_ocn (gem_util.f90:29-46) */
  {
    static const int ao[][2] = {{IA_PCO2, IO_DIC}, {IA_PCO2_13C, IO_DIC_13C}, {IA_PCO2_14C, IO_DIC_14C}, {IA_PO2, IO_O2},
                                {IA_PCFC11, IO_CFC11}, {IA_PCFC12, IO_CFC12}};
    for (m = 0; m < 6; m++) { la = AFIND(ao[m][0]); if (la) b->atm2ocn[la] = LFIND(ao[m][1]); }
  }
  /* decay constants, gem_util.f90:280-308 */
  for (l = 1; l <= b->L; l++) b->lam_ocn[l] = (b->io[l] == IO_DIC_14C || b->io[l] == IO_DOM_C_14C) ? BG_LAMBDA_14C : 0.0;
  for (ls = 1; ls <= b->LS; ls++) b->lam_sed[ls] = (b->is[ls] == IS_POC_14C || b->is[ls] == IS_CACO3_14C) ? BG_LAMBDA_14C : 0.0;
  for (la = 1; la <= b->LA; la++) b->lam_atm[la] = (b->ia[la] == IA_PCO2_14C) ? BG_LAMBDA_14C : 0.0;
  /* Schmidt number and Bunsen coefficients, gem_data.f90:69-136 */
  {
    static const double Sc[][5] = {{IA_PCO2, 2073.1, 125.62, 3.6276, 0.043219}, {IA_PO2, 1953.4, 128.00, 3.9918, 0.050091},
                                   {IA_PCFC11, 4039.8, 264.70, 8.2552, 0.103590}, {IA_PCFC12, 3713.2, 243.40, 7.5879, 0.095215}};
    static const double Bu[][7] = {{IA_PCO2, -60.2409, 93.4517, 23.3585, 0.023517, -0.023656, 0.0047036},
                                   {IA_PO2, -58.3877, 85.8079, 23.8439, -0.034892, 0.015568, -0.0019387},
                                   {IA_PCFC11, -136.2685, 206.1150, 57.2805, -0.148598, 0.095114, -0.0163396},
                                   {IA_PCFC12, -124.4395, 185.4299, 51.6383, -0.149779, 0.094668, -0.0160043}};
    for (m = 0; m < 4; m++) {
      la = AFIND((int)Sc[m][0]);
      if (!la) continue;
      for (l = 0; l < 4; l++) b->Sc[la][l] = Sc[m][1 + l];
      for (l = 0; l < 6; l++) b->bunsen[la][l] = Bu[m][1 + l];
    }
  }
}

/* initialise_biogem + initialise_atchem for the frozen configuration.  Wind speed (par_windspeed_file) is read from the
 * registry field "bg_windspeed" (maxi,maxj), which the caller fills before this call. */
void cgo_biogem_setup(cgo_t *o, const char *params) {
  struct cgo_bg *b;
  const int I = NI, J = NJ, K = NK;
  const long ij = (long)I * J, n3 = ij * K;
  int i, j, k, l, ls, la;
  if (!params) params = o->params ? o->params : "";
  if (BG) return;
  b = BG = (struct cgo_bg *)calloc(1, sizeof(struct cgo_bg));
  /* biogem-defaults.nml */
  b->t_runtime = bg_par(params, "par_misc_t_runtime", 1001.0);
  b->t_end = 0.0 + b->t_runtime;                 /* ctrl_misc_t_BP=.FALSE., par_misc_t_start=0 (biogem.f90:232-236) */
  b->k0_PO4 = bg_par(params, "par_bio_k0_PO4", 2.0E-06); b->c0_PO4 = bg_par(params, "par_bio_c0_PO4", 0.050E-06);
  b->red_POP_POC = 106.0; b->red_POP_PON = 16.0; b->red_POP_PO2 = -138.0; b->red_PON_ALK = -1.00;
  b->red_DOMfrac = bg_par(params, "par_bio_red_DOMfrac", 0.66); b->red_RDOMfrac = 0.0;
  b->red_POC_CaCO3 = bg_par(params, "par_bio_red_POC_CaCO3", 0.2); b->red_POC_CaCO3_pP = bg_par(params, "par_bio_red_POC_CaCO3_pP", 0.0);
  b->DOMlifetime = bg_par(params, "par_bio_remin_DOMlifetime", 0.5);
  b->POC_frac2 = bg_par(params, "par_bio_remin_POC_frac2", 0.05); b->POC_eL1 = bg_par(params, "par_bio_remin_POC_eL1", 500.0);
  b->POC_eL2 = 1000000.0; b->POC_dfrac2 = 0.0; b->POC_c0frac2 = 0.1E-6;
  b->CaCO3_frac2 = 0.5; b->CaCO3_eL1 = 1000.0; b->CaCO3_eL2 = 1000000.0;
  b->sinkingrate = 125.0 / (1.0 / 365.25);      /* m d-1 -> m yr-1: par/conv_d_yr (biogem_data.f90:421, gem_cmn.f90 conv_d_yr = 1/conv_yr_d) */
  b->remin_k_O2 = 1.0; b->remin_c0_O2 = 8.0E-6; b->remin_ci_O2 = 8.0E-6;
  b->gastransfer_a = 0.310; b->d13C_DIC_Corg_ef = 25.0; b->Fgeothermal = 0.0;
  b->kbiogem = (int)bg_par(params, "conv_kocn_kbiogem", 2); b->katchem = (int)bg_par(params, "conv_kocn_katchem", 2);
  b->genie_timestep = bg_par(params, "genie_timestep", 3600.0 * 24.0 * 365.25 / 5.0 / o->nyear);
  b->clock_ms = 0;
  bg_tables(o);
  /* frozen initial composition (data_BIOGEM bg_ocn_init_*, data_ATCHEM ac_atm_init_*) */
  {
    static const double oi[][2] = {{IO_DIC, 2.244E-03}, {IO_DIC_13C, 0.4}, {IO_DIC_14C, -150.0}, {IO_PO4, 2.159E-06},
        {IO_O2, 1.696E-04}, {IO_ALK, 2.363E-03}, {IO_CA, 1.025E-02}, {IO_MG, 5.282E-02}};
    static const double ai[][2] = {{IA_PCO2, 278.0E-06}, {IA_PCO2_13C, -6.5}, {IA_PCO2_14C, 0.0}, {IA_PO2, 0.2095}};
    char key[64];
    int m;
    for (l = 1; l <= b->L; l++) {
      b->ocn_init[l] = 0.0;
      for (m = 0; m < 8; m++) if ((int)oi[m][0] == b->io[l]) b->ocn_init[l] = oi[m][1];
      snprintf(key, sizeof key, "ocn_init_%d", b->io[l]);
      b->ocn_init[l] = bg_par(params, key, b->ocn_init[l]);
    }
    for (la = 1; la <= b->LA; la++) {
      b->atm_init[la] = 0.0;
      for (m = 0; m < 4; m++) if ((int)ai[m][0] == b->ia[la]) b->atm_init[la] = ai[m][1];
      snprintf(key, sizeof key, "atm_init_%d", b->ia[la]);
      b->atm_init[la] = bg_par(params, key, b->atm_init[la]);
    }
  }
  /* arrays */
  if (!o->bg_ocn) {
    o->bg_ocn = bg_alloc(o, "ocn", n3 * NL); o->bg_vdocn = bg_alloc(o, "vdocn", n3 * NL);
    o->bg_M = bg_alloc(o, "bg_M", n3); o->bg_rM = bg_alloc(o, "bg_rM", n3); o->bg_V = bg_alloc(o, "bg_V", n3);
  }
  b->bio_part = bg_alloc(o, "bio_part", n3 * b->LS); b->bio_remin = bg_alloc(o, "bio_remin", n3 * NL);
  b->bio_settle = bg_alloc(o, "bio_settle", n3 * b->LS); b->focn = bg_alloc(o, "bg_focn", n3 * NL);
  b->red = bg_alloc(o, "bio_part_red", ij * b->LS * b->LS);
  b->carb = bg_alloc(o, "carb", ij * N_IC); b->carbisor = bg_alloc(o, "carbisor", ij * N_ICI);
  b->seaice = bg_alloc(o, "bg_seaice", ij); b->seaice_th = bg_alloc(o, "bg_seaice_th", ij); b->wspeed = bg_alloc(o, "bg_wspeed", ij);
  b->solfor = bg_alloc(o, "bg_solfor", ij); b->fxsw = bg_alloc(o, "bg_fxsw", ij); b->mld = bg_alloc(o, "bg_mld", ij);
  b->rho_surf = bg_alloc(o, "bg_rho_surf", ij); b->A = bg_alloc(o, "bg_A", ij); b->rA = bg_alloc(o, "bg_rA", ij);
  b->windspeed_file = cgo_field(o, "bg_windspeed", NULL);
  b->atm = bg_alloc(o, "atm", ij * b->LA); b->sfcatm1 = bg_alloc(o, "sfcatm1", ij * b->LA);
  b->sfxatm1 = bg_alloc(o, "sfxatm1", ij * b->LA); b->sfxsumatm = bg_alloc(o, "sfxsumatm", ij * b->LA);
  b->atm_A = bg_alloc(o, "atm_A", ij); b->atm_V = bg_alloc(o, "atm_V", ij);
  b->sfcocn1 = bg_alloc(o, "sfcocn1", ij * NL); b->sfxsed1 = bg_alloc(o, "sfxsed1", ij * b->LS); b->focnatm = bg_alloc(o, "focnatm", ij * b->LA);
  b->sig = bg_alloc(o, "bg_sig", 3 + 3 * NL + b->LA);
  b->sig2 = bg_alloc(o, "bg_sig2", 8 + b->LS + 2 * b->LA);
  b->carb3 = bg_alloc(o, "carb3", n3 * N_IC); b->ciso3 = bg_alloc(o, "carbisor3", n3 * N_ICI); b->cc3 = bg_alloc(o, "carbconst3", n3 * N_CC);
  b->sl_ocn = bg_alloc(o, "sl_ocn", n3 * NL); b->sl_part = bg_alloc(o, "sl_part", n3 * b->LS); b->sl_carb = bg_alloc(o, "sl_carb", n3 * N_IC);
  b->sl_cc = bg_alloc(o, "sl_carbconst", n3 * N_CC); b->sl_ciso = bg_alloc(o, "sl_carbisor", n3 * N_ICI); b->sl_t = bg_alloc(o, "sl_t", 1);
  b->sfxsumsed = bg_alloc(o, "sfxsumsed", ij * b->LS); b->sfcsumocn = bg_alloc(o, "sfcsumocn", ij * NL); b->sfxsumrok1 = bg_alloc(o, "sfxsumrok1", ij * NL);
  b->rst_I = bg_alloc(o, "rst_atm_I", ij * b->LA); b->rst_II = bg_alloc(o, "rst_atm_II", ij * b->LA);
  b->rst_atm = bg_alloc(o, "force_restore_atm", ij * b->LA);
  b->solar_constant = o->solconst;
  /* sub_init_phys_ocn, biogem_data.f90:1098-1137 */
  {
    double dzl[66], dzal[66];
    for (k = 0; k <= K + 1; k++) { dzl[k] = 0.0; dzal[k] = 0.0; }
    for (k = 1; k <= K; k++) { dzl[k] = o->dz[k]; dzal[k] = o->dza[k]; }
    dzal[K] = dzl[K] / 2.0;
    for (k = 1; k <= K; k++) {
      double s = 0.0;
      int kk;
      b->dD[k] = CG_DSC * dzl[k];
      for (kk = k; kk <= K; kk++) s = s + CG_DSC * dzl[kk];
      b->Dbot[k] = s;
    }
    b->Dmid_surf = CG_DSC * dzal[K];
    for (k = 1; k <= K; k++) {   /* phys_ocn(ipo_Dmid,i,j,k) = SUM(goldstein_dsc*loc_grid_dza(k:n_k)), biogem_data.f90:1120 */
      double s = 0.0;
      int kk;
      for (kk = k; kk <= K; kk++) s = s + CG_DSC * dzal[kk];
      b->Dmid[k] = s;
    }
  }
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++) {
      A2(b->A, i, j) = 2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / I) * (o->sv[j] - o->sv[j - 1]);
      A2(b->rA, i, j) = 1.0 / A2(b->A, i, j);
      A2(b->rho_surf, i, j) = (K >= K1(i, j)) ? BG_M3_KG : 0.0;
      for (k = K1(i, j); k <= K; k++) {
        PHV(i, j, k) = b->dD[k] * A2(b->A, i, j);
        PHM(i, j, k) = BG_M3_KG * PHV(i, j, k);
        PHRM(i, j, k) = 1.0 / PHM(i, j, k);
      }
    }
  /* sub_init_tracer_ocn_comp :1281-1309, then sub_biogem_copy_tstoocn :3745-3763 */
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++)
      for (k = K1(i, j); k <= K; k++) {
        for (l = 1; l <= b->L; l++) {
          if (b->otype[l] == 1) OCN(l, i, j, k) = b->ocn_init[l];
          else if (b->otype[l] >= 11) {
            const double tot = b->ocn_init[b->odep[l]];
            const double fr = iso_fraction(b->ocn_init[l], b->otype[l] == 11 ? BG_STD_13C : BG_STD_14C);
            OCN(l, i, j, k) = fr * tot;
          }
        }
        OCN(1, i, j, k) = TS(1, i, j, k) + BG_ZEROC;
        OCN(2, i, j, k) = TS(2, i, j, k) + o->saln0;
      }
  /* sub_init_bio :579-624 */
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++) {
      RED(b->s_POP, b->s_POP, i, j) = 1.0; RED(b->s_POC, b->s_POC, i, j) = 1.0; RED(b->s_CaCO3, b->s_CaCO3, i, j) = 1.0;
      RED(b->s_POP, b->s_POC, i, j) = b->red_POP_POC;
      RED(b->s_POC, b->s_POP, i, j) = 1.0 / RED(b->s_POP, b->s_POC, i, j);
      RED(b->s_POC, b->s_CaCO3, i, j) = b->red_POC_CaCO3;
      RED(b->s_CaCO3, b->s_POC, i, j) = 1.0 / RED(b->s_POC, b->s_CaCO3, i, j);
    }
  /* sub_init_carb :2336-2430 (surface layer; deeper carbonate chemistry is diagnostic only) */
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++)
      if (K >= K1(i, j)) {
        double cc[N_CC], *cb = &CARB(0, i, j);
        const double S = OCN(2, i, j, K);
        calc_carbconst(b->Dmid_surf, OCN(1, i, j, K), S, cc);
        adj_carbconst(OCN(b->l_Ca, i, j, K), OCN(b->l_Mg, i, j, K), cc);
        cb[IC_H] = pow(10.0, -7.8);
        calc_carb(OCN(b->l_DIC, i, j, K), OCN(b->l_ALK, i, j, K), OCN(b->l_Ca, i, j, K), OCN(b->l_PO4, i, j, K), 0.0, f_Btot(S),
                  f_SO4tot(S), f_Ftot(S), cc, cb);
        calc_carb_RF0(OCN(b->l_DIC, i, j, K), OCN(b->l_ALK, i, j, K), OCN(b->l_PO4, i, j, K), 0.0, f_Btot(S), f_SO4tot(S),
                      f_Ftot(S), cc, cb);
        calc_carb_riso(OCN(1, i, j, K), OCN(b->l_DIC, i, j, K), OCN(b->l_DIC13, i, j, K), cb, 1.0, BG_STD_13C, &CISO(ICI_DIC_R13C, i, j));
        calc_carb_riso(OCN(1, i, j, K), OCN(b->l_DIC, i, j, K), OCN(b->l_DIC14, i, j, K), cb, 2.0, BG_STD_14C, &CISO(ICI_DIC_R14C, i, j));
      }
  /* sub_init_carb, the cells below the surface (k < n_k): same seed, same calls; read by the time-slice diagnostics only */
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++)
      for (k = K1(i, j); k < K; k++) {
        const long c = (i - 1) + (long)I * ((j - 1) + (long)J * (k - 1));
        double *cc = &b->cc3[N_CC * c], *cb = &b->carb3[N_IC * c];
        const double S = OCN(2, i, j, k);
        calc_carbconst(b->Dmid[k], OCN(1, i, j, k), S, cc);
        adj_carbconst(OCN(b->l_Ca, i, j, k), OCN(b->l_Mg, i, j, k), cc);
        cb[IC_H] = pow(10.0, -7.8);
        calc_carb(OCN(b->l_DIC, i, j, k), OCN(b->l_ALK, i, j, k), OCN(b->l_Ca, i, j, k), OCN(b->l_PO4, i, j, k), 0.0, f_Btot(S),
                  f_SO4tot(S), f_Ftot(S), cc, cb);
        calc_carb_RF0(OCN(b->l_DIC, i, j, k), OCN(b->l_ALK, i, j, k), OCN(b->l_PO4, i, j, k), 0.0, f_Btot(S), f_SO4tot(S),
                      f_Ftot(S), cc, cb);
        calc_carb_riso(OCN(1, i, j, k), OCN(b->l_DIC, i, j, k), OCN(b->l_DIC13, i, j, k), cb, 1.0, BG_STD_13C, &b->ciso3[N_ICI * c + ICI_DIC_R13C]);
        calc_carb_riso(OCN(1, i, j, k), OCN(b->l_DIC, i, j, k), OCN(b->l_DIC14, i, j, k), cb, 2.0, BG_STD_14C, &b->ciso3[N_ICI * c + ICI_DIC_R14C]);
      }
  /* sub_init_force_restore_atm :2706-2791 with data/biogem/worjh2_preindustrial (I = 0, II = 1 at wet points, 2-point signal) */
  {
    static const double sig[][2] = {{IA_PCO2, 2.780000E-04}, {IA_PCO2_13C, -6.50}, {IA_PCO2_14C, 38.4}, {IA_PCFC11, 0.0}, {IA_PCFC12, 0.0}};
    int m;
    for (la = 3; la <= b->LA; la++)
      for (m = 0; m < 5; m++)
        if ((int)sig[m][0] == b->ia[la]) {
          b->rst_sel[la] = 1;
          b->rst_tconst[la] = 0.1;
          /* sub_load_data_t2 (biogem_lib.f90:1471-1476): file times (0, 999999) reversed and measured from par_misc_t_end */
          b->rst_sig_t[la][0] = b->t_end - 1.0 * 999999.0; b->rst_sig_t[la][1] = b->t_end - 1.0 * 0.0;
          b->rst_sig_v[la][0] = 1.0 * sig[m][1]; b->rst_sig_v[la][1] = 1.0 * sig[m][1];
          b->rst_sig_i[la][0] = 2; b->rst_sig_i[la][1] = 2;
          for (i = 1; i <= I; i++)
            for (j = 1; j <= J; j++)
              if (K >= K1(i, j)) { RSTATM(b->rst_I, la, i, j) = 0.0; RSTATM(b->rst_II, la, i, j) = 1.0; }
        }
  }
  /* sub_biogem_copy_ocntots :3691-3739 (ctrl_misc_Snorm) */
  {
    double totV = 0.0, sumSV = 0.0, meanS;
    for (k = 1; k <= K; k++) for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) totV = totV + PHV(i, j, k);
    for (k = 1; k <= K; k++) for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) sumSV = sumSV + OCN(2, i, j, k) * PHV(i, j, k);
    meanS = sumSV / totV;
    for (i = 1; i <= I; i++)
      for (j = 1; j <= J; j++)
        for (k = K1(i, j); k <= K; k++)
          for (l = 3; l <= b->L; l++) {
            TS(l, i, j, k) = OCN(l, i, j, k) * (meanS / OCN(2, i, j, k));
            TS1(l, i, j, k) = TS(l, i, j, k);
          }
  }
  /* initialise_atchem: sub_init_phys_atm :195-229, sub_init_tracer_atm_comp :234-261; genie.f90 initial cpl_comp_atmocn */
  {
    const double th0 = -BG_PI / 2, th1 = BG_PI / 2;
    const double s0 = sin(th0), s1 = sin(th1);
    const double ds = (s1 - s0) / (double)J;
    for (i = 1; i <= I; i++)
      for (j = 1; j <= J; j++) {
        const double svj = s0 + (double)j * ds, svjm = s0 + (double)(j - 1) * ds;
        A2(b->atm_A, i, j) = 2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / (double)I) * (svj - svjm);
        A2(b->atm_V, i, j) = BG_ATM_TH * A2(b->atm_A, i, j);
        for (la = 1; la <= b->LA; la++) {
          if (b->atype[la] == 0) { if (b->ia[la] == IA_T) ATM(la, i, j) = BG_ZEROC; }
          else if (b->atype[la] == 1) ATM(la, i, j) = b->atm_init[la];
          else ATM(la, i, j) = iso_fraction(b->atm_init[la], b->atype[la] == 11 ? BG_STD_13C : BG_STD_14C) * b->atm_init[b->adep[la]];
        }
        for (la = 3; la <= b->LA; la++) SFCATM1(la, i, j) = ATM(la, i, j);
        /* the initial cpl_comp_EMBM_wrapper (genie.f90:90; atchem.f90:270-282): tstar_atm, surf_qstar_atm of initialise_embm */
        SFCATM1(1, i, j) = A2(o->tstar_atm, i, j);
        SFCATM1(2, i, j) = A2(o->qstar_atm, i, j);
      }
  }
  (void)ls;
  b->go = 1;
  cgo_biogem_climate(o);   /* genie.f90:109-112: biogem_climate_wrapper before the main loop */
}

/* kept for the stand-alone tracer-coupling tests: builds ocn from ts without the rest of BIOGEM */
void cgo_biogem_init(cgo_t *o) {
  int i, j, k, l;
  const long n3 = (long)NI * NJ * NK;
  if (!o->bg_ocn) {
    o->bg_ocn = cgo_alloc(o, "ocn", n3 * NL); o->bg_vdocn = cgo_alloc(o, "vdocn", n3 * NL);
    o->bg_M = cgo_alloc(o, "bg_M", n3); o->bg_rM = cgo_alloc(o, "bg_rM", n3); o->bg_V = cgo_alloc(o, "bg_V", n3);
  }
  for (i = 1; i <= NI; i++)
    for (j = 1; j <= NJ; j++)
      for (k = K1(i, j); k <= NK; k++) {
        const double dD = CG_DSC * o->dz[k];
        const double A = 2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / NI) * (o->sv[j] - o->sv[j - 1]);
        PHV(i, j, k) = dD * A;
        PHM(i, j, k) = BG_M3_KG * PHV(i, j, k);
        PHRM(i, j, k) = 1.0 / PHM(i, j, k);
      }
  for (k = 1; k <= NK; k++)
    for (j = 1; j <= NJ; j++)
      for (i = 1; i <= NI; i++) {
        OCN(1, i, j, k) = TS(1, i, j, k) + BG_ZEROC;
        OCN(2, i, j, k) = TS(2, i, j, k) + o->saln0;
        for (l = 3; l <= NL; l++) OCN(l, i, j, k) = TS(l, i, j, k);
      }
}

/* ------------------------------------------------------------------------------------------ biogem_forcing */
static void update_sig(double t, const double *sig, int *sig_i, double *x) { /* biogem_box.f90:3174-3218, 1-based indices */
  if (sig_i[0] > 1) {
    if (t < sig[sig_i[0] - 1]) {
      for (;;) {
        sig_i[0] = sig_i[0] - 1;
        if (t > sig[sig_i[0] - 1]) break;
        else if (sig_i[0] == 1) break;
      }
    }
  }
  if (sig_i[1] > 1) {
    if (t < sig[sig_i[1] - 1]) {
      for (;;) {
        sig_i[1] = sig_i[1] - 1;
        if (t >= sig[sig_i[1] - 1]) { sig_i[1] = sig_i[1] + 1; break; }
        else if (sig_i[1] == 1) break;
      }
    }
  }
  if (fabs(sig[sig_i[1] - 1] - sig[sig_i[0] - 1]) > BG_NULLSMALL) *x = (sig[sig_i[1] - 1] - t) / (sig[sig_i[1] - 1] - sig[sig_i[0] - 1]);
  else *x = 0.5;
}
void cgo_biogem_forcing(cgo_t *o) { /* biogem.f90:2083-2127 + biogem_box.f90:3333-3366 */
  struct cgo_bg *b = BG;
  int la, i, j;
  const double t = b->t_runtime - (double)b->clock_ms / (1000.0 * BG_YR_S);
  for (la = 3; la <= b->LA; la++)
    if (b->rst_sel[la]) {
      double x;
      update_sig(t, b->rst_sig_t[la], b->rst_sig_i[la], &x);
      b->rst_sig_x[la] = (1 - x) * b->rst_sig_v[la][b->rst_sig_i[la][1] - 1] + x * b->rst_sig_v[la][b->rst_sig_i[la][0] - 1];
      for (i = 1; i <= NI; i++)
        for (j = 1; j <= NJ; j++) {
          const double f = RSTATM(b->rst_I, la, i, j) + b->rst_sig_x[la] * (RSTATM(b->rst_II, la, i, j) - RSTATM(b->rst_I, la, i, j));
          if (b->atype[la] == 1) RSTATM(b->rst_atm, la, i, j) = f;
          else RSTATM(b->rst_atm, la, i, j) = iso_fraction(f, b->atype[la] == 11 ? BG_STD_13C : BG_STD_14C) * RSTATM(b->rst_atm, b->adep[la], i, j);
        }
    }
}

/* ------------------------------------------------------------------------------------------ step_biogem pieces */
/* sub_box_remin_redfield (O2 only selected): conv = (k_O2*kO2/loc_k)*conv_ls_lo; returns the scalar factor */
static double remin_redfield_factor(const struct cgo_bg *b, double o2) {
  double loc_k = 0.0;
  const double O2 = dmax2(0.0, o2);
  double kO2 = O2 / (O2 + b->remin_c0_O2);
  loc_k = loc_k + b->remin_k_O2 * kO2;
  if (loc_k < BG_NULLSMALL) loc_k = 1.0;   /* ctrl_bio_remin_POC_kinetic = .FALSE. */
  if (O2 < BG_NULLSMALL) kO2 = 1.0;
  return b->remin_k_O2 * kO2 / loc_k;
}

/* sub_box_remin_DOM :2287-2406 for one column; writes vbio_remin (= remin[l][k]) */
static void remin_DOM(cgo_t *o, int i, int j, double dtyr, double (*remin)[64]) {
  struct cgo_bg *b = BG;
  const int k1 = K1(i, j);
  int k, l, ls, m;
  double part[BG_MAXLS + 1];
  for (l = 1; l <= b->L; l++) for (k = 1; k <= NK; k++) remin[l][k] = 0.0;
  for (k = NK; k >= k1; k--) {
    double ratio;
    for (ls = 1; ls <= b->LS; ls++) part[ls] = 0.0;
    if (b->DOMlifetime > dtyr) ratio = dtyr / b->DOMlifetime; else ratio = 1.0;
    if (OCN(b->l_DOMC, i, j, k) > BG_NULLSMALL) {
      for (l = 3; l <= b->L; l++)
        if (b->dom2pom[l]) {
          part[b->dom2pom[l]] = part[b->dom2pom[l]] + 1.0 * ratio * OCN(l, i, j, k);
          remin[l][k] = remin[l][k] - ratio * OCN(l, i, j, k);
        }
    }
    {
      const double f = remin_redfield_factor(b, OCN(b->l_O2, i, j, k));
      for (ls = 1; ls <= b->LS; ls++)
        for (m = 0; m < b->n_ls_lo[ls]; m++) {
          const int lo = b->ls_lo[ls][m];
          remin[lo][k] = remin[lo][k] + (f * b->conv_ls_lo[ls][lo]) * part[ls];
        }
    }
  }
}

/* sub_box_remin_part :2412-2875 for one column (fixed e-folding profiles, no ballast, no scavenging) */
static void remin_part(cgo_t *o, int i, int j, double dtyr, double (*remin)[64]) {
  struct cgo_bg *b = BG;
  const int k1 = K1(i, j), K = NK, LS = b->LS;
  int k, kk, l, ls, m, klim, min_k;
  double OLD[BG_MAXLS + 1][64], TMP[BG_MAXLS + 1][64], part[BG_MAXLS + 1][64], lremin[BG_MAXL + 1][64], settle[BG_MAXLS + 1][64];
  double part_remin[BG_MAXLS + 1];
  double CaCO3_frac1 = 0.0, CaCO3_frac2 = 0.0, CaCO3_ratio = 0.0, POC_frac1 = 0.0, POC_frac2 = 0.0, POC_ratio = 0.0;
  double sinkingrate, dt_layer = 0.0;
  for (ls = 1; ls <= LS; ls++)
    for (k = 1; k <= K; k++) { OLD[ls][k] = (k >= k1) ? PART(ls, i, j, k) : 0.0; part[ls][k] = 0.0; settle[ls][k] = 0.0; }
  for (l = 1; l <= b->L; l++) for (k = 1; k <= K; k++) lremin[l][k] = 0.0;
  if (dtyr * b->sinkingrate <= CG_DSC) { klim = k1; sinkingrate = b->sinkingrate; }
  else { klim = K; sinkingrate = b->sinkingrate; }
  for (k = K; k >= klim; k--) {
    double part_tot = 0.0;
    part_tot = part_tot + OLD[b->s_POC][k];
    part_tot = part_tot + OLD[b->s_CaCO3][k];
    if (part_tot > BG_NULLSMALL) {
      if (k == k1) min_k = k1 - 1;
      else {
        const double max_D = b->Dbot[k] + dtyr * sinkingrate;
        min_k = k1 - 1;
        for (kk = k - 1; kk >= k1; kk--)
          if (b->Dbot[kk] > max_D) { min_k = kk; break; }
      }
      for (ls = 1; ls <= LS; ls++) for (kk = 1; kk <= K; kk++) TMP[ls][kk] = 0.0;
      for (ls = 1; ls <= LS; ls++) TMP[ls][k] = OLD[ls][k];
      for (kk = k - 1; kk >= min_k; kk--) {
        if (kk >= k1) {
          const double layerratio = b->dD[kk + 1] / b->dD[kk];
          const double dD = b->dD[kk];
          if (sinkingrate > BG_NULLSMALL) dt_layer = dD / sinkingrate;
          (void)dt_layer;
          /* carbonate */
          CaCO3_frac1 = (1.0 - exp(-dD / b->CaCO3_eL1));
          CaCO3_frac2 = (1.0 - exp(-dD / b->CaCO3_eL2));
          CaCO3_ratio = 1.0 - ((1.0 - TMP[b->s_CaCO3f2][kk + 1]) * CaCO3_frac1 + TMP[b->s_CaCO3f2][kk + 1] * CaCO3_frac2);
          l = b->s_CaCO3f2;
          if (TMP[l][kk + 1] > BG_NULLSMALL) TMP[l][kk] = (1.0 - CaCO3_frac2) * TMP[l][kk + 1] / CaCO3_ratio; else TMP[l][kk] = 0.0;
          /* particulate organic matter */
          POC_frac1 = (1.0 - exp(-dD / b->POC_eL1));
          POC_frac2 = (1.0 - exp(-dD / b->POC_eL2));
          POC_ratio = 1.0 - ((1.0 - TMP[b->s_POCf2][kk + 1]) * POC_frac1 + TMP[b->s_POCf2][kk + 1] * POC_frac2);
          l = b->s_POCf2;
          if (TMP[l][kk + 1] > BG_NULLSMALL) TMP[l][kk] = (1.0 - POC_frac2) * TMP[l][kk + 1] / POC_ratio; else TMP[l][kk] = 0.0;
          /* particle concentrations in the layer below :2719-2771 */
          for (ls = 1; ls <= LS; ls++) {
            const int dep_type = b->stype[b->sdep_ls[ls]];
            if ((b->sdep[ls] == IS_POC) || (b->stype[ls] == ST_POM) || (dep_type == IS_POC))
              TMP[ls][kk] = TMP[ls][kk + 1] * layerratio * POC_ratio;
            else if ((b->sdep[ls] == IS_CACO3) || (b->stype[ls] == ST_CACO3) || (dep_type == ST_CACO3))
              TMP[ls][kk] = TMP[ls][kk + 1] * layerratio * CaCO3_ratio;
          }
          for (ls = 1; ls <= LS; ls++) part_remin[ls] = (layerratio * TMP[ls][kk + 1] - TMP[ls][kk]);
          {
            const double f = remin_redfield_factor(b, OCN(b->l_O2, i, j, kk));
            for (ls = 1; ls <= LS; ls++)
              for (m = 0; m < b->n_ls_lo[ls]; m++) {
                const int lo = b->ls_lo[ls][m];
                lremin[lo][kk] = lremin[lo][kk] + (f * b->conv_ls_lo[ls][lo]) * part_remin[ls];
              }
          }
        }
      }
      if (min_k >= k1)
        for (ls = 1; ls <= LS; ls++) part[ls][min_k] = part[ls][min_k] + TMP[ls][min_k];
      for (kk = k; kk >= min_k + 1; kk--)
        for (ls = 1; ls <= LS; ls++) {
          if (b->stype[ls] == ST_FRAC) settle[ls][kk] = settle[ls][kk] + TMP[ls][kk];
          else settle[ls][kk] = settle[ls][kk] + PHM(i, j, kk) * TMP[ls][kk];
        }
    }
  }
  for (ls = 1; ls <= LS; ls++)
    for (k = 1; k <= K; k++) {
      if (k >= k1) PART(ls, i, j, k) = part[ls][k];
      SETTLE(ls, i, j, k) = settle[ls][k];
    }
  for (l = 1; l <= b->L; l++) for (k = 1; k <= K; k++) remin[l][k] = remin[l][k] + lremin[l][k];
}

/* surface carbonate system, solubility, piston velocity, restoring, gas exchange, uptake, interface arrays: the body of
 * the (i,j) loop of step_biogem :985-1802 for one wet column.  Returns 1 if the pH solve failed (error_stop). */
static int surface_column(cgo_t *o, int i, int j, double dtyr, double dts, const double *tmod, double *fatm) {
  struct cgo_bg *b = BG;
  const int K = NK, k1 = K1(i, j);
  int l, ls, la, m;
  double cc[N_CC], solconst[BG_MAXLA + 1], pv[BG_MAXLA + 1], focnatm[BG_MAXLA + 1], fatmocn_[BG_MAXLA + 1], focnatm_[BG_MAXLA + 1];
  double *cb = &CARB(0, i, j);
  const double T = OCN(1, i, j, K), S = OCN(2, i, j, K);
  const double rdtyr = 1.0 / dtyr, rdts = 1.0 / dts;
  /* *** UPDATE AIR-SEA INTERFACE AQUEOUS SYSTEM *** :1026-1104 */
  calc_carbconst(b->Dmid_surf, T, S, cc);
  adj_carbconst(OCN(b->l_Ca, i, j, K), OCN(b->l_Mg, i, j, K), cc);
  for (la = 3; la <= b->LA; la++) {
    solconst[la] = 0.0; pv[la] = 0.0;
    if (b->atype[la] == 1) solconst[la] = calc_solconst(b, la, T, S, A2(b->rho_surf, i, j));
  }
  { /* sub_calc_pv :81-117 */
    double TC = T - BG_ZEROC, TC2, TC3, u2;
    if (TC < 0.0) TC = 0.0;
    if (TC > 30.0) TC = 30.0;
    TC2 = TC * TC; TC3 = TC2 * TC;
    u2 = A2(b->wspeed, i, j) * A2(b->wspeed, i, j);
    for (la = 3; la <= b->LA; la++)
      if (b->atype[la] == 1) {
        const double Sc = b->Sc[la][0] - b->Sc[la][1] * TC + b->Sc[la][2] * TC2 - b->Sc[la][3] * TC3;
        pv[la] = (1.0 / 1.0E+02) * (24.0 * 365.25) * b->gastransfer_a * u2 * pow(Sc * 1.515E-3, -0.5);
      }
  }
  if (calc_carb(OCN(b->l_DIC, i, j, K), OCN(b->l_ALK, i, j, K), OCN(b->l_Ca, i, j, K), OCN(b->l_PO4, i, j, K), 0.0, f_Btot(S),
                f_SO4tot(S), f_Ftot(S), cc, cb))
    return 1;
  calc_carb_RF0(OCN(b->l_DIC, i, j, K), OCN(b->l_ALK, i, j, K), OCN(b->l_PO4, i, j, K), 0.0, f_Btot(S), f_SO4tot(S), f_Ftot(S), cc, cb);
  calc_carb_riso(T, OCN(b->l_DIC, i, j, K), OCN(b->l_DIC13, i, j, K), cb, 1.0, BG_STD_13C, &CISO(ICI_DIC_R13C, i, j));
  calc_carb_riso(T, OCN(b->l_DIC, i, j, K), OCN(b->l_DIC14, i, j, K), cb, 2.0, BG_STD_14C, &CISO(ICI_DIC_R14C, i, j));
  /* *** CALCULATE RESTORING BOUNDARY CONDITIONS *** atmosphere :1119-1146 */
  for (la = 1; la <= b->LA; la++) fatm[la] = 0.0;
  for (la = 3; la <= b->LA; la++)
    if (b->rst_sel[la]) {
      if (b->rst_sig_i[la][0] != b->rst_sig_i[la][1]) {
        double d;
        if (b->atype[la] == 1) { if (RSTATM(b->rst_atm, la, i, j) < 0.0) RSTATM(b->rst_atm, la, i, j) = SFCATM1(la, i, j); }
        else { if (RSTATM(b->rst_atm, la, i, j) <= BG_NULL) RSTATM(b->rst_atm, la, i, j) = SFCATM1(la, i, j); }
        d = (RSTATM(b->rst_atm, la, i, j) - SFCATM1(la, i, j)) * tmod[la];
        fatm[la] = (1.0 / (double)(NI * NJ)) * BG_ATM_MOL * d * rdtyr;
      }
    }
  /* geothermal heat :1265-1270 */
  FOCN(1, i, j, k1) = FOCN(1, i, j, k1) + BG_YR_S * b->Fgeothermal * A2(b->A, i, j) / (1.0E+03 * BG_CP);
  /* *** OCEAN-ATMOSPHERE EXCHANGE FLUXES *** fun_calc_ocnatm_flux :123-299 */
  {
    const double rho = A2(b->rho_surf, i, j), TC = T - BG_ZEROC;
    const double area = (1.0 - A2(b->seaice, i, j)) * A2(b->A, i, j);
    double alpha_as = 0.0, alpha_sa = 0.0;
    for (la = 1; la <= b->LA; la++) { focnatm[la] = 0.0; focnatm_[la] = 0.0; fatmocn_[la] = 0.0; }
    for (la = 3; la <= b->LA; la++) {
      const int lo = b->atm2ocn[la];
      if (b->atype[la] == 1) {
        double loc_atm = solconst[la] * SFCATM1(la, i, j), loc_ocn, buff, deqm, dflux;
        if (lo == b->l_DIC) {
          loc_ocn = cb[IC_CO2];
          if (cb[IC_RF0] > BG_NULLSMALL) buff = 1.0 / (cb[IC_RF0] * cb[IC_CO2] / OCN(b->l_DIC, i, j, K));
          else { loc_ocn = 0.0; loc_atm = 0.0; buff = 1.0; }
        } else { loc_ocn = OCN(lo, i, j, K); buff = 1.0; }
        if (loc_ocn < BG_NULLSMALL) loc_ocn = 0.0;
        if (loc_atm < BG_NULLSMALL) loc_atm = 0.0;
        focnatm_[la] = pv[la] * area * rho * loc_ocn;
        fatmocn_[la] = pv[la] * area * rho * loc_atm;
        deqm = b->dD[K] * A2(b->A, i, j) * rho * buff * fabs(loc_atm - loc_ocn);
        dflux = dtyr * fabs(focnatm_[la] - fatmocn_[la]);
        if (deqm > BG_NULLSMALL) {
          const double r = dflux / deqm;
          if (r > 1.00) { focnatm_[la] = (1.00 / r) * focnatm_[la]; fatmocn_[la] = (1.00 / r) * fatmocn_[la]; }
        }
      } else if (b->ia[la] == IA_PCO2_13C) {
        const double r_atm = SFCATM1(la, i, j) / SFCATM1(b->a_CO2, i, j), r_ocn = CISO(ICI_CO2_R13C, i, j);
        const double R_atm = r_atm / (1.0 - r_atm), R_ocn = r_ocn / (1.0 - r_ocn);
        const double alpha_k = 0.99912, alpha_alpha = 0.99869 + 4.9E-6 * TC;
        alpha_as = alpha_alpha * alpha_k; alpha_sa = alpha_k;
        fatmocn_[la] = (alpha_as * R_atm / (1.0 + alpha_as * R_atm)) * fatmocn_[b->a_CO2];
        focnatm_[la] = (alpha_sa * R_ocn / (1.0 + alpha_sa * R_ocn)) * focnatm_[b->a_CO2];
      } else if (b->ia[la] == IA_PCO2_14C) {
        const double r_atm = SFCATM1(la, i, j) / SFCATM1(b->a_CO2, i, j), r_ocn = CISO(ICI_CO2_R14C, i, j);
        const double R_atm = r_atm / (1.0 - r_atm), R_ocn = r_ocn / (1.0 - r_ocn);
        fatmocn_[la] = ((alpha_as * alpha_as) * R_atm / (1.0 + (alpha_as * alpha_as) * R_atm)) * fatmocn_[b->a_CO2];
        focnatm_[la] = ((alpha_sa * alpha_sa) * R_ocn / (1.0 + (alpha_sa * alpha_sa) * R_ocn)) * focnatm_[b->a_CO2];
      }
      focnatm[la] = focnatm_[la] - fatmocn_[la];
    }
  }
  for (la = 3; la <= b->LA; la++) {
    const int lo = b->atm2ocn[la];
    fatm[la] = fatm[la] + focnatm[la];
    if (lo) FOCN(lo, i, j, K) = FOCN(lo, i, j, K) - 1.0 * focnatm[la];
    b->focnatm[(la - 1) + b->LA * ((i - 1) + NI * (j - 1))] = focnatm[la];
  }
  /* *** SURFACE OCEAN BIOLOGICAL PRODUCTIVITY *** sub_calc_bio_uptake, 1N1T_PO4MM :346-1514 */
  {
    int k_mld = k1, k;
    double dPO4, kPO4, kI, ficefree, uptake[BG_MAXL + 1][64], pDOM[BG_MAXLS + 1][64];
    const double PO4 = OCN(b->l_PO4, i, j, K);
    double DOMfrac = b->red_DOMfrac, RDOMfrac = b->red_RDOMfrac, DOMtotal;
    for (k = K; k >= 1; k--)
      if (b->Dbot[k] >= A2(b->mld, i, j)) { k_mld = k; break; }
    kPO4 = PO4 / (PO4 + b->c0_PO4);
    ficefree = (1.0 - A2(b->seaice, i, j));
    kI = A2(b->solfor, i, j) / b->solar_constant;
    if (PO4 > BG_NULLSMALL) dPO4 = dtyr * ficefree * kI * kPO4 * b->k0_PO4; else dPO4 = 0.0;
    DOMtotal = DOMfrac + RDOMfrac;
    if (DOMtotal > 1.0) { DOMfrac = DOMfrac / DOMtotal; RDOMfrac = 1.0 - DOMfrac; DOMtotal = 1.0; }
    { /* CaCO3:POC rain ratio, 'Ridgwelletal2007ab' :886-893 */
      const double ohm = cb[IC_OHM_CAL];
      if (ohm > 1.0) RED(b->s_POC, b->s_CaCO3, i, j) = (1.0 - DOMtotal) * b->red_POC_CaCO3 * pow(ohm - 1.0, b->red_POC_CaCO3_pP);
      else RED(b->s_POC, b->s_CaCO3, i, j) = 0.0;
    }
    { /* isotopic fractionation :1058-1106 */
      const double Kq = 3.170E-05 + (-1.788E-07) * T + 2.829E-10 * (T * T);
      const double delta_Corg = -b->d13C_DIC_Corg_ef + (b->d13C_DIC_Corg_ef - 0.7) * Kq / cb[IC_CO2];
      double alpha = 1.0 + delta_Corg / 1000.0, R = CISO(ICI_CO2_R13C, i, j) / (1.0 - CISO(ICI_CO2_R13C, i, j));
      double delta_CaCO3;
      RED(b->s_POC, b->s_POC13, i, j) = alpha * R / (1.0 + alpha * R);
      alpha = 1.0 + 2.0 * delta_Corg / 1000.0;
      R = CISO(ICI_CO2_R14C, i, j) / (1.0 - CISO(ICI_CO2_R14C, i, j));
      RED(b->s_POC, b->s_POC14, i, j) = alpha * R / (1.0 + alpha * R);
      delta_CaCO3 = 15.10 - 4232.0 / T;
      alpha = 1.0 + delta_CaCO3 / 1000.0;
      R = CISO(ICI_HCO3_R13C, i, j) / (1.0 - CISO(ICI_HCO3_R13C, i, j));
      RED(b->s_CaCO3, b->s_CaCO313, i, j) = alpha * R / (1.0 + alpha * R);
      alpha = 1.0 + 2.0 * delta_CaCO3 / 1000.0;
      R = CISO(ICI_HCO3_R14C, i, j) / (1.0 - CISO(ICI_HCO3_R14C, i, j));
      RED(b->s_CaCO3, b->s_CaCO314, i, j) = alpha * R / (1.0 + alpha * R);
    }
    for (l = 1; l <= b->L; l++) for (k = k_mld; k <= K; k++) uptake[l][k] = 0.0;
    for (ls = 1; ls <= b->LS; ls++) for (k = k_mld; k <= K; k++) pDOM[ls][k] = 0.0;
    for (k = k_mld; k <= K; k++) {
      /* bulk export :1186-1230 */
      PART(b->s_POC, i, j, k) = RED(b->s_POP, b->s_POC, i, j) * dPO4;
      for (ls = 1; ls <= b->LS; ls++)
        if (b->stype[ls] == ST_BIO) PART(ls, i, j, k) = RED(b->s_POC, ls, i, j) * PART(b->s_POC, i, j, k);
      for (ls = 1; ls <= b->LS; ls++) {
        if (b->stype[ls] == ST_POM) PART(ls, i, j, k) = RED(b->s_POC, ls, i, j) * PART(b->s_POC, i, j, k);
        else if (b->stype[ls] == ST_CACO3) PART(ls, i, j, k) = RED(b->s_CaCO3, ls, i, j) * PART(b->s_CaCO3, i, j, k);
      }
      for (ls = 1; ls <= b->LS; ls++)
        if (b->stype[ls] >= 11) PART(ls, i, j, k) = RED(b->sdep_ls[ls], ls, i, j) * PART(b->sdep_ls[ls], i, j, k);
      /* inorganic uptake :1254-1262 */
      for (ls = 1; ls <= b->LS; ls++)
        for (m = 0; m < b->n_ls_lo[ls]; m++) {
          const int lo = b->ls_lo[ls][m];
          uptake[lo][k] = uptake[lo][k] + b->conv_ls_lo[ls][lo] * PART(ls, i, j, k);
        }
      /* DOM production :1316-1352 */
      for (ls = 1; ls <= b->LS; ls++)
        if (b->pom2dom[ls]) {
          const double r_POM_DOM = 1.0; /* par_bio_red_rP_POM_DOM = par_bio_red_rN_POM_DOM = 1.0 */
          pDOM[ls][k] = r_POM_DOM * DOMfrac * PART(ls, i, j, k);
          REMIN(b->pom2dom[ls], i, j, k) = REMIN(b->pom2dom[ls], i, j, k) + pDOM[ls][k];
        }
      for (ls = 1; ls <= b->LS; ls++) PART(ls, i, j, k) = PART(ls, i, j, k) - (pDOM[ls][k] + 0.0);
      /* initial particulate fraction partitioning :1354-1378 */
      {
        const double kP = PO4 / (PO4 + b->POC_c0frac2);
        PART(b->s_POCf2, i, j, k) = (1.0 - kP) * b->POC_dfrac2 + b->POC_frac2;
        PART(b->s_CaCO3f2, i, j, k) = b->CaCO3_frac2;
      }
      for (l = 3; l <= b->L; l++) REMIN(l, i, j, k) = REMIN(l, i, j, k) - uptake[l][k];
    }
  }
  /* *** INTERFACE ARRAY UPDATE *** :1724-1761 */
  for (la = 3; la <= b->LA; la++) SFXATM1(la, i, j) = A2(b->rA, i, j) * (1.0 / BG_YR_S) * fatm[la];
  for (l = 1; l <= b->L; l++)
    b->sfcocn1[(l - 1) + NL * ((i - 1) + NI * (j - 1))] = OCN(l, i, j, k1) + REMIN(l, i, j, k1) + dtyr * PHRM(i, j, k1) * FOCN(l, i, j, k1);
  for (ls = 1; ls <= b->LS; ls++) {
    const double fs = SETTLE(ls, i, j, k1);
    if (b->stype[ls] == ST_FRAC) b->sfxsed1[(ls - 1) + b->LS * ((i - 1) + NI * (j - 1))] = fs * rdts * dtyr;
    else b->sfxsed1[(ls - 1) + b->LS * ((i - 1) + NI * (j - 1))] = A2(b->rA, i, j) * fs * rdts;
  }
  return 0;
}

/* step_biogem, biogem.f90:528-1877.  Returns non-zero when the reference would stop (carbonate chemistry failure). */
int cgo_biogem_step(cgo_t *o) {
  struct cgo_bg *b = BG;
  const int I = NI, J = NJ, K = NK;
  const long n3 = (long)I * J * K;
  int i, j, k, l, ls, la, m, err = 0;
  const double dts = (double)(b->kbiogem * o->kocn_loop) * b->genie_timestep;
  const double t = b->t_runtime - (double)b->clock_ms / (1000.0 * BG_YR_S);
  const double dtyr = dts / BG_YR_S;
  double fd_ocn[BG_MAXL + 1], fd_sed[BG_MAXLS + 1], tmod[BG_MAXLA + 1];
  double fatm[BG_MAXLA + 1];
  if (!b->go) return 0;
  memset(b->bio_remin, 0, sizeof(double) * n3 * NL);
  memset(b->focn, 0, sizeof(double) * n3 * NL);
  for (l = 1; l <= b->L; l++) fd_ocn[l] = exp(-dtyr * b->lam_ocn[l]);
  for (ls = 1; ls <= b->LS; ls++) fd_sed[ls] = exp(-dtyr * b->lam_sed[ls]);
  for (la = 3; la <= b->LA; la++) tmod[la] = b->rst_sel[la] ? 1.0 - exp(-dtyr / b->rst_tconst[la]) : 0.0;
  /* decay + closed-system sediment return :823-943 */
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++) {
      const int k1 = K1(i, j);
      if (K >= k1) {
        double fsedocn[BG_MAXL + 1];
        for (l = 3; l <= b->L; l++)
          if (fabs(b->lam_ocn[l]) > BG_NULLSMALL)
            for (k = k1; k <= K; k++) FOCN(l, i, j, k) = FOCN(l, i, j, k) - PHM(i, j, k) * (1.0 - fd_ocn[l]) * OCN(l, i, j, k) / dtyr;
        for (ls = 1; ls <= b->LS; ls++)
          if (fabs(b->lam_sed[ls]) > BG_NULLSMALL)
            for (k = k1; k <= K; k++) PART(ls, i, j, k) = fd_sed[ls] * PART(ls, i, j, k);
        for (l = 1; l <= b->L; l++) fsedocn[l] = 0.0;
        {
          const double f = remin_redfield_factor(b, OCN(b->l_O2, i, j, k1));
          for (ls = 1; ls <= b->LS; ls++)
            for (m = 0; m < b->n_ls_lo[ls]; m++) {
              const int lo = b->ls_lo[ls][m];
              fsedocn[lo] = fsedocn[lo] + (f * b->conv_ls_lo[ls][lo]) * SETTLE(ls, i, j, k1);
            }
        }
        for (l = 3; l <= b->L; l++) REMIN(l, i, j, k1) = REMIN(l, i, j, k1) + PHRM(i, j, k1) * fsedocn[l];
      }
    }
  /* (v) loop: DOM and particulate remineralisation :953-969 */
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++)
      if (K >= K1(i, j)) {
        double vremin[BG_MAXL + 1][64];
        remin_DOM(o, i, j, dtyr, vremin);
        remin_part(o, i, j, dtyr, vremin);
        for (l = 1; l <= b->L; l++)
          for (k = K1(i, j); k <= K; k++) REMIN(l, i, j, k) = REMIN(l, i, j, k) + vremin[l][k];
      }
  /* (i,j) loop :985-1802 */
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++)
      if (K >= K1(i, j)) err |= surface_column(o, i, j, dtyr, dts, tmod, fatm);
  /* tracer anomaly :1811-1844 */
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++)
      for (k = K1(i, j); k <= K; k++)
        for (l = 1; l <= b->L; l++) DOCN(l, i, j, k) = REMIN(l, i, j, k) + dtyr * PHRM(i, j, k) * FOCN(l, i, j, k);
  if (t < BG_NULLSMALL) b->go = 0;
  return err;
}

/* biogem.f90:1885-2077 (vdbio_part = 0: no particulate flux forcing) */
void cgo_biogem_tracercoupling(cgo_t *o) {
  const int L = NL;
  int i, j, k, l, n, nv = 0;
  double tot_V, rtot_V, mean_S_OLD, rmean_S_OLD, mean_S_NEW, Sratio, rSratio, s;
  double off[3] = {0.0, BG_ZEROC, o->saln0};
  double *tot_OLD = (double *)calloc(L + 1, 8), *tot_NEW = (double *)calloc(L + 1, 8), *rtot_NEW = (double *)calloc(L + 1, 8);
  int *ci = (int *)calloc((size_t)NI * NJ, sizeof(int)), *cj = (int *)calloc((size_t)NI * NJ, sizeof(int));
  double *partial = (double *)calloc((size_t)NI * NJ, 8);
  if (!o->bg_ocn) cgo_biogem_init(o);
  if (BG && !BG->go) goto done;
  for (i = 1; i <= NI; i++)
    for (j = 1; j <= NJ; j++)
      if (NK >= K1(i, j)) { ci[nv] = i; cj[nv] = j; nv++; }
  /* total ocean volume :1934-1941 */
  for (n = 0; n < nv; n++) {
    s = 0.0;
    for (k = K1(ci[n], cj[n]); k <= NK; k++) s = s + PHV(ci[n], cj[n], k);
    partial[n] = s;
  }
  tot_V = 0.0;
  for (n = 0; n < nv; n++) tot_V = tot_V + partial[n];
  rtot_V = 1.0 / tot_V;
  /* (0) original inventories :1952-1969 */
  for (n = 0; n < nv; n++) {
    s = 0.0;
    for (k = K1(ci[n], cj[n]); k <= NK; k++) s = s + OCN(2, ci[n], cj[n], k) * PHV(ci[n], cj[n], k);
    partial[n] = s * rtot_V;
  }
  mean_S_OLD = 0.0;
  for (n = 0; n < nv; n++) mean_S_OLD = mean_S_OLD + partial[n];
  rmean_S_OLD = 1.0 / mean_S_OLD;
  for (l = 3; l <= L; l++) {
    for (n = 0; n < nv; n++) {
      s = 0.0;
      for (k = K1(ci[n], cj[n]); k <= NK; k++) s = s + OCN(l, ci[n], cj[n], k) * PHM(ci[n], cj[n], k);
      partial[n] = s;
    }
    tot_OLD[l] = 0.0;
    for (n = 0; n < nv; n++) tot_OLD[l] = tot_OLD[l] + partial[n];
  }
  /* (1) salinity-adjusted new inventory :1975-1997 */
  for (l = 3; l <= L; l++) {
    for (n = 0; n < nv; n++) {
      s = 0.0;
      for (k = K1(ci[n], cj[n]); k <= NK; k++)
        s = s + (TS(l, ci[n], cj[n], k) * OCN(2, ci[n], cj[n], k) * rmean_S_OLD) * PHM(ci[n], cj[n], k);
      partial[n] = s;
    }
    tot_NEW[l] = 0.0;
    for (n = 0; n < nv; n++) tot_NEW[l] = tot_NEW[l] + partial[n];
    if (fabs(tot_NEW[l]) < BG_NULLSMALL) rtot_NEW[l] = 0.0; else rtot_NEW[l] = 1.0 / tot_NEW[l];
  }
  /* (2) new T,S and new mean salinity :2001-2025 (ctrl_force_GOLDSTEInTS = .TRUE.) */
  for (n = 0; n < nv; n++) {
    s = 0.0;
    for (k = K1(ci[n], cj[n]); k <= NK; k++)
      s = s + (TS(2, ci[n], cj[n], k) + off[2] + DOCN(2, ci[n], cj[n], k)) * PHV(ci[n], cj[n], k);
    partial[n] = s * rtot_V;
  }
  mean_S_NEW = 0.0;
  for (n = 0; n < nv; n++) mean_S_NEW = mean_S_NEW + partial[n];
  Sratio = mean_S_NEW / mean_S_OLD;
  rSratio = 1.0 / Sratio;
  /* (3) adjust fields :2033-2061 */
  for (n = 0; n < nv; n++) {
    i = ci[n]; j = cj[n];
    for (k = NK; k >= K1(i, j); k--) {
      const double Sold = OCN(2, i, j, k);
      double Snew;
      for (l = 1; l <= 2; l++) {
        const double v = TS(l, i, j, k) + off[l] + DOCN(l, i, j, k);
        OCN(l, i, j, k) = v;
        TS(l, i, j, k) = v - off[l];
      }
      Snew = OCN(2, i, j, k);
      for (l = 3; l <= L; l++) {
        const double lv = TS(l, i, j, k) * Sold * rmean_S_OLD;
        double x = (tot_OLD[l] * rtot_NEW[l]) * lv + DOCN(l, i, j, k);
        x = Sratio * x;
        OCN(l, i, j, k) = x;
        TS(l, i, j, k) = (mean_S_NEW / Snew) * x;
      }
      if (BG) /* :2042-2043 */
        for (l = 1; l <= BG->LS; l++) PART(l, i, j, k) = Sratio * (PART(l, i, j, k) + 0.0);
      PHM(i, j, k) = rSratio * PHM(i, j, k);
      PHRM(i, j, k) = Sratio * PHRM(i, j, k);
      for (l = 1; l <= L; l++) TS1(l, i, j, k) = TS(l, i, j, k);
    }
  }
done:
  free(tot_OLD); free(tot_NEW); free(rtot_NEW); free(ci); free(cj); free(partial);
}

/* biogem_climate_sol :2243-2263 and biogem_climate :2132-2239 (what the step reads afterwards) */
void cgo_biogem_climate_sol(cgo_t *o) {
  struct cgo_bg *b = BG;
  int i, j;
  for (i = 1; i <= NI; i++)
    for (j = 1; j <= NJ; j++) { A2(b->solfor, i, j) = o->go_solfor[j]; A2(b->fxsw, i, j) = A2(o->fxsw, i, j); }
  b->solar_constant = o->solconst;
}
void cgo_biogem_climate(cgo_t *o) {
  struct cgo_bg *b = BG;
  int i, j;
  for (i = 1; i <= NI; i++)
    for (j = 1; j <= NJ; j++) {
      if (NK >= K1(i, j)) {
        A2(b->mld, i, j) = -5000.0 * A2(o->mld, i, j);   /* go_mldta = -5000*mld (goldstein.f90:449); mld = 0 for imld = 0 */
        A2(b->rho_surf, i, j) = calc_rho(OCN(1, i, j, NK), OCN(2, i, j, NK));
      }
      A2(b->solfor, i, j) = o->go_solfor[j];
      A2(b->fxsw, i, j) = A2(o->fxsw, i, j);
      A2(b->seaice, i, j) = A2(o->frac_sic, i, j);
      A2(b->seaice_th, i, j) = A2(o->hght_sic, i, j);
      A2(b->wspeed, i, j) = A2(b->windspeed_file, i, j);   /* ctrl_force_windspeed = .TRUE. */
    }
  b->solar_constant = o->solconst;
  for (i = 1; i <= NI; i++) for (j = 1; j <= NJ; j++) A2(o->cost, i, j) = 0.0;
}

/* cpl_flux_ocnatm, atchem.f90:306-320 */
void cgo_cpl_flux_ocnatm(cgo_t *o) {
  struct cgo_bg *b = BG;
  const double dts = (double)(b->kbiogem * o->kocn_loop) * b->genie_timestep;
  long n;
  for (n = 0; n < (long)NI * NJ * b->LA; n++) { b->sfxsumatm[n] = b->sfxsumatm[n] + dts * b->sfxatm1[n]; b->sfxatm1[n] = 0.0; }
}

/* diag_biogem_timeseries, biogem.f90:2703-3159: the sig_ocn / sig_ocn_sur (+ benthic) / sig_ocnatm integrals of one BIOGEM
 * step (:2771-2917) and int_t_sig (:3082); sums in array element order (i fastest) */
void cgo_biogem_sig_update(cgo_t *o, double ben_Dmin) {
  struct cgo_bg *b = BG;
  const int I = NI, J = NJ, K = NK;
  const double dts = (double)(b->kbiogem * o->kocn_loop) * b->genie_timestep, dtyr = dts / BG_YR_S;
  double *S = b->sig;
  double tot_M = 0.0, tot_M_sur = 0.0, tot_A = 0.0, tot_A_ben = 0.0, tot_A_atm = 0.0, rtot_M, rtot_A, rtot_A_ben, rtot_A_atm;
  double *mask = (double *)calloc((size_t)I * J, 8);
  int i, j, k, l, la;
  for (k = 1; k <= K; k++) for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) if (k >= K1(i, j)) tot_M = tot_M + o->bg_M[(i - 1) + I * ((j - 1) + J * (k - 1))];
  for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) {
    const double A = 2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / I) * (o->sv[j] - o->sv[j - 1]);
    tot_A_atm = tot_A_atm + A;
    if (K1(i, j) > K) continue;
    tot_M_sur = tot_M_sur + o->bg_M[(i - 1) + I * ((j - 1) + J * (K - 1))];
    tot_A = tot_A + (1.0 - A2(b->seaice, i, j)) * A;
    {
      double Dbot = 0.0;
      for (k = K1(i, j); k <= K; k++) Dbot = Dbot + CG_DSC * o->dz[k];
      if (Dbot > ben_Dmin) { mask[(i - 1) + I * (j - 1)] = 1.0; tot_A_ben = tot_A_ben + 1.0 * A; }
    }
  }
  rtot_M = tot_M > BG_NULLSMALL ? 1.0 / tot_M : 0.0; rtot_A = tot_A > BG_NULLSMALL ? 1.0 / tot_A : 0.0;
  rtot_A_ben = tot_A_ben > BG_NULLSMALL ? 1.0 / tot_A_ben : 0.0; rtot_A_atm = tot_A_atm > BG_NULLSMALL ? 1.0 / tot_A_atm : 0.0;
  S[1] = S[1] + dtyr * tot_M;
  S[2] = S[2] + dtyr * tot_M_sur;
  for (l = 1; l <= NL; l++) {
    double s3 = 0.0, ss = 0.0, sb = 0.0;
    for (k = 1; k <= K; k++) for (j = 1; j <= J; j++) for (i = 1; i <= I; i++)
      if (k >= K1(i, j)) s3 = s3 + o->bg_M[(i - 1) + I * ((j - 1) + J * (k - 1))] * OCN(l, i, j, k);
    for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) {
      const double A = 2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / I) * (o->sv[j] - o->sv[j - 1]);
      if (K1(i, j) > K) continue;
      ss = ss + (1.0 - A2(b->seaice, i, j)) * A * OCN(l, i, j, K);
      if (mask[(i - 1) + I * (j - 1)] != 0.0) sb = sb + 1.0 * A * OCN(l, i, j, K1(i, j));
    }
    S[3 + (l - 1)] = S[3 + (l - 1)] + dtyr * s3 * rtot_M;
    S[3 + NL + (l - 1)] = S[3 + NL + (l - 1)] + dtyr * ss * rtot_A;
    S[3 + 2 * NL + (l - 1)] = S[3 + 2 * NL + (l - 1)] + dtyr * sb * rtot_A_ben;
  }
  for (la = 1; la <= b->LA; la++) {
    double sa = 0.0;
    for (j = 1; j <= J; j++) for (i = 1; i <= I; i++)
      sa = sa + 2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / I) * (o->sv[j] - o->sv[j - 1]) * b->sfcatm1[(la - 1) + b->LA * ((i - 1) + I * (j - 1))];
    S[3 + 3 * NL + (la - 1)] = S[3 + 3 * NL + (la - 1)] + dtyr * sa * rtot_A_atm;
  }
  /* the flux, export and "misc" integrals of the same routine: int_fexport_sig (:2870-2876), int_focnatm_sig (:2877-2883, with
   * locij_focnatm :2807-2811), int_diag_airsea_sig (:3058-3062), int_misc_seaice_sig / _th / _vol (:2926-2937), the overturning
   * stream function (:2938-2945, sub_calc_psi biogem_box.f90:3796-3869) and the mean land air temperature (:2946-2964) */
  {
    double *S2 = b->sig2;
    const int LS = b->LS, LA = b->LA;
    int ls;
    double s_ice = 0.0, s_icesum = 0.0, s_vol = 0.0;
    for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) {
      const double A = (K1(i, j) <= K) ? 2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / I) * (o->sv[j] - o->sv[j - 1]) : 0.0;   /* phys_ocn(ipo_A,:,:,n_k) */
      s_ice = s_ice + A * A2(b->seaice, i, j);
      s_icesum = s_icesum + A2(b->seaice, i, j);
      s_vol = s_vol + A2(b->seaice_th, i, j) * A * A2(b->seaice, i, j);
    }
    S2[0] = S2[0] + dtyr * s_ice;
    if (s_icesum > BG_NULLSMALL) S2[1] = S2[1] + dtyr * s_vol / s_ice;
    S2[2] = S2[2] + dtyr * s_vol;
    {
      /* sub_calc_psi: the caller's loc_opsi(0:n_j,0:n_k) takes rows / columns 1.. from the routine, 0 stays 0 */
      double *opsi = (double *)calloc((size_t)(J + 1) * (K + 1), 8), *opsia = (double *)calloc((size_t)(J + 1) * (K + 1), 8);
      double omin = 0.0, omax = 0.0, omina = 0.0, omaxa = 0.0, ou;
#define OP(a, j, k) a[(j) + (J + 1) * (k)]
      for (j = 1; j <= J - 1; j++)
        for (k = 1; k <= K - 1; k++) {
          ou = 0.0;
          for (i = 1; i <= I; i++) ou = ou + o->cv[j] * U(2, i, j, k) * o->dphi;
          OP(opsi, j, k) = OP(opsi, j, k - 1) - o->dz[k] * ou;
        }
      for (j = o->jsf + 1; j <= J - 1; j++)
        for (k = 1; k <= K - 1; k++) {
          ou = 0.0;
          for (i = o->ias[j]; i <= o->iaf[j]; i++) ou = ou + o->cv[j] * U(2, i, j, k) * o->dphi;
          OP(opsia, j, k) = OP(opsia, j, k - 1) - o->dz[k] * ou;
          if (OP(opsia, j, k) < omina && k <= K / 2) omina = OP(opsia, j, k);
          if (OP(opsia, j, k) > omaxa && k <= K / 2) omaxa = OP(opsia, j, k);
        }
      for (j = 0; j <= J; j++) for (k = 0; k <= K; k++) {      /* minval / maxval(loc_opsi(:,:)) */
        if (OP(opsi, j, k) < omin) omin = OP(opsi, j, k);
        if (OP(opsi, j, k) > omax) omax = OP(opsi, j, k);
      }
#undef OP
      S2[3] = S2[3] + dtyr * omin; S2[4] = S2[4] + dtyr * omax;
      S2[5] = S2[5] + dtyr * omina; S2[6] = S2[6] + dtyr * omaxa;
      free(opsi); free(opsia);
    }
    {
      double sg = 0.0, ta = 0.0;
      for (i = 1; i <= I; i++) for (j = 1; j <= J; j++)
        if (K < K1(i, j)) {
          const double A = 2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / I) * (o->sv[j] - o->sv[j - 1]);
          sg = sg + A * SFCATM1(1, i, j);
          ta = ta + A;
        }
      if (ta > BG_NULLSMALL) S2[7] = S2[7] + dtyr * sg / ta; else S2[7] = 0.0;
    }
    for (ls = 1; ls <= LS; ls++) {
      double s = 0.0;
      for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) s = s + SETTLE(ls, i, j, K);
      S2[8 + (ls - 1)] = S2[8 + (ls - 1)] + s;
    }
    for (la = 3; la <= LA; la++) {
      double s = 0.0, sd = 0.0;
      for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) {
        if (K >= K1(i, j))
          s = s + BG_YR_S * (2.0 * BG_PI * (BG_REARTH * BG_REARTH) * (1.0 / I) * (o->sv[j] - o->sv[j - 1])) * SFXATM1(la, i, j);
        sd = sd + b->focnatm[(la - 1) + LA * ((i - 1) + I * (j - 1))];
      }
      S2[8 + LS + (la - 1)] = S2[8 + LS + (la - 1)] + dtyr * s;
      S2[8 + LS + LA + (la - 1)] = S2[8 + LS + LA + (la - 1)] + dtyr * sd;
    }
  }
  S[0] = S[0] + dtyr;
  free(mask);
}

/* diag_biogem_timeslice (biogem.f90:2421-2699), the part that is arithmetic: inside a save window with dum_save, the carbonate
 * system of EVERY wet cell is solved again from the cell's last [H+] (:2478-2567; the surface cell's carb is the array
 * step_biogem seeds its own solve from, so the diagnostic feeds back into the next step exactly as in the reference), then the
 * window integrals grow (:2572-2579): int_ocn, int_bio_part, int_carb, int_carbconst, int_carbisor += dtyr * field, int_t += dtyr.
 * The window bookkeeping (par_data_save_timeslice, :2460-2470, 2627-2696) and the netCDF writer stay with the host. */
void cgo_biogem_slice_update(cgo_t *o) {
  struct cgo_bg *b = BG;
  const int I = NI, J = NJ, K = NK;
  int i, j, k, l, ls, q;
  const double dts = (double)(b->kbiogem * o->kocn_loop) * b->genie_timestep;
  const double dtyr = dts / BG_YR_S;
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++)
      for (k = K1(i, j); k <= K; k++) {
        const long c = (i - 1) + (long)I * ((j - 1) + (long)J * (k - 1));
        double ccs[N_CC];
        double *cc = (k < K) ? &b->cc3[N_CC * c] : ccs;
        double *cb = (k < K) ? &b->carb3[N_IC * c] : &CARB(0, i, j);
        double *ci = (k < K) ? &b->ciso3[N_ICI * c] : &CISO(0, i, j);
        const double S = OCN(2, i, j, k);
        calc_carbconst(b->Dmid[k], OCN(1, i, j, k), S, cc);
        adj_carbconst(OCN(b->l_Ca, i, j, k), OCN(b->l_Mg, i, j, k), cc);
        calc_carb(OCN(b->l_DIC, i, j, k), OCN(b->l_ALK, i, j, k), OCN(b->l_Ca, i, j, k), OCN(b->l_PO4, i, j, k), 0.0, f_Btot(S),
                  f_SO4tot(S), f_Ftot(S), cc, cb);
        if (k == K) {
          /* sub_calc_carb_RF0 is called at every k with the SURFACE cell's arguments (:2533-2545); only the call behind the
           * surface solve leaves a value that survives (it writes carb(ic_RF0,i,j,n_k) alone) */
          const double Ss = OCN(2, i, j, K);
          calc_carb_RF0(OCN(b->l_DIC, i, j, K), OCN(b->l_ALK, i, j, K), OCN(b->l_PO4, i, j, K), 0.0, f_Btot(Ss), f_SO4tot(Ss),
                        f_Ftot(Ss), cc, cb);
        }
        calc_carb_riso(OCN(1, i, j, k), OCN(b->l_DIC, i, j, k), OCN(b->l_DIC13, i, j, k), cb, 1.0, BG_STD_13C, &ci[ICI_DIC_R13C]);
        calc_carb_riso(OCN(1, i, j, k), OCN(b->l_DIC, i, j, k), OCN(b->l_DIC14, i, j, k), cb, 2.0, BG_STD_14C, &ci[ICI_DIC_R14C]);
        for (l = 1; l <= NL; l++) b->sl_ocn[(l - 1) + NL * c] = b->sl_ocn[(l - 1) + NL * c] + dtyr * OCN(l, i, j, k);
        for (ls = 1; ls <= b->LS; ls++) b->sl_part[(ls - 1) + b->LS * c] = b->sl_part[(ls - 1) + b->LS * c] + dtyr * PART(ls, i, j, k);
        for (q = 0; q < N_IC; q++) b->sl_carb[q + N_IC * c] = b->sl_carb[q + N_IC * c] + dtyr * cb[q];
        for (q = 0; q < N_CC; q++) b->sl_cc[q + N_CC * c] = b->sl_cc[q + N_CC * c] + dtyr * cc[q];
        for (q = 0; q < N_ICI; q++) b->sl_ciso[q + N_ICI * c] = b->sl_ciso[q + N_ICI * c] + dtyr * ci[q];
      }
  b->sl_t[0] = b->sl_t[0] + dtyr;
}
void cgo_biogem_slice_auto(cgo_t *o, int on) { BG->slice_auto = on; }

void cgo_biogem_sig_auto(cgo_t *o, int on, double ben_Dmin) { BG->sig_auto = on; BG->sig_ben_Dmin = ben_Dmin; }

/* cpl_flux_ocnsed, sedgem.f90:1029-1068 (loc_scalei = loc_scalej = 1: i1 = i, j1 = j) */
void cgo_cpl_flux_ocnsed(cgo_t *o, double dts) {
  struct cgo_bg *b = BG;
  long n;
  for (n = 0; n < (long)NI * NJ * b->LS; n++) b->sfxsumsed[n] = b->sfxsumsed[n] + dts * b->sfxsed1[n];
}
/* cpl_comp_ocnsed, sedgem.f90:894-937 */
void cgo_cpl_comp_ocnsed(cgo_t *o, int ocnstep, int mbiogem, int msedgem) {
  struct cgo_bg *b = BG;
  const int w = ((ocnstep - mbiogem) % msedgem) / mbiogem;   /* int(MOD(ocnstep - mbiogem, msedgem)/mbiogem) */
  long n;
  for (n = 0; n < (long)NI * NJ * NL; n++) b->sfcsumocn[n] = ((double)w * b->sfcsumocn[n] + b->sfcocn1[n]) / (double)(w + 1);
}
/* reinit_flux_rokocn, rokgem.f90:472-480 */
void cgo_reinit_flux_rokocn(cgo_t *o) {
  struct cgo_bg *b = BG;
  long n;
  for (n = 0; n < (long)NI * NJ * NL; n++) b->sfxsumrok1[n] = 0.0;
}

/* step_atchem :63-158 + cpl_comp_atmocn :252-264 + cpl_comp_EMBM :270-282 */
void cgo_atchem_step(cgo_t *o) {
  struct cgo_bg *b = BG;
  const int I = NI, J = NJ;
  int i, j, la;
  const double dts = (double)(b->katchem * o->kocn_loop) * b->genie_timestep;
  const double dtyr = dts / BG_YR_S;
  const double F14C = 0.0; /* par_atm_F14C */
  double *c_am = (double *)calloc((size_t)I * J, 8), *c_ma = (double *)calloc((size_t)I * J, 8), *fl = (double *)calloc((size_t)I * J * b->LA, 8);
  double fd[BG_MAXLA + 1];
  for (j = 1; j <= J; j++)
    for (i = 1; i <= I; i++) {
      A2(c_am, i, j) = A2(b->atm_V, i, j) / (BG_PA_ATM * BG_R_SI * ATM(1, i, j));
      A2(c_ma, i, j) = 1.0 / A2(c_am, i, j);
    }
  for (la = 1; la <= b->LA; la++) fd[la] = exp(-dtyr * b->lam_atm[la]);
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++) {
      for (la = 3; la <= b->LA; la++)
        if (fabs(b->lam_atm[la]) > BG_NULLSMALL) ATM(la, i, j) = fd[la] * ATM(la, i, j);
      /* sub_calc_generate_14C, atchem_box.f90 */
      fl[(b->a_CO214 - 1) + b->LA * ((i - 1) + I * (j - 1))] =
          fl[(b->a_CO214 - 1) + b->LA * ((i - 1) + I * (j - 1))] + dtyr * (1.0 / (double)(I * J)) * F14C;
    }
  for (la = 3; la <= b->LA; la++) {
    double tot = 0.0, totV = 0.0;
    for (j = 1; j <= J; j++)
      for (i = 1; i <= I; i++)
        ATM(la, i, j) = ATM(la, i, j) + A2(c_ma, i, j) * A2(b->atm_A, i, j) * SFXSUMATM(la, i, j) +
                        A2(c_ma, i, j) * fl[(la - 1) + b->LA * ((i - 1) + I * (j - 1))];
    for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) tot = tot + A2(c_am, i, j) * ATM(la, i, j);
    for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) totV = totV + A2(b->atm_V, i, j);
    for (j = 1; j <= J; j++) for (i = 1; i <= I; i++) ATM(la, i, j) = (tot / totV) * BG_PA_ATM * BG_R_SI * ATM(1, i, j);
  }
  for (i = 1; i <= I; i++)
    for (j = 1; j <= J; j++) {
      for (la = 1; la <= b->LA; la++) SFXSUMATM(la, i, j) = 0.0;
      for (la = 3; la <= b->LA; la++) SFCATM1(la, i, j) = ATM(la, i, j);   /* cpl_comp_atmocn */
      SFCATM1(1, i, j) = A2(o->tstar_atm, i, j);                            /* cpl_comp_EMBM */
      SFCATM1(2, i, j) = A2(o->qstar_atm, i, j);
    }
  free(c_am); free(c_ma); free(fl);
}

/* the BIOGEM / ATCHEM block of one koverall iteration, genie.f90:352-447.  Called by cgo_run after the physics modules. */
int cgo_biogem_koverall(cgo_t *o, long k) {
  struct cgo_bg *b = BG;
  int err = 0;
  if (!b) return 0;
  if (k % (b->kbiogem * o->kocn_loop) == 0) {
    if (k == b->kbiogem * o->kocn_loop) cgo_biogem_climate_sol(o);
    cgo_biogem_forcing(o);
    err = cgo_biogem_step(o);
    cgo_biogem_tracercoupling(o);
    cgo_biogem_climate(o);
    /* diag_biogem_timeseries_wrapper, genie.f90:401-405: behind biogem_climate_wrapper (:387), ahead of cpl_flux_ocnatm_wrapper
     * (:411) and of the ATCHEM step (:446-455) -- the ocean is this block's, sfcatm1 still the previous block's */
    if (b->slice_auto) cgo_biogem_slice_update(o);             /* diag_biogem_timeslice_wrapper, genie.f90:391-395 */
    if (b->sig_auto) cgo_biogem_sig_update(o, b->sig_ben_Dmin);
    cgo_cpl_flux_ocnatm(o);
  }
  if (k % (b->katchem * o->kocn_loop) == 0) cgo_atchem_step(o);
  return err;
}
void cgo_biogem_tick(cgo_t *o) { /* increment_genie_clock, genie_global.f90:401-410 */
  if (BG) BG->clock_ms = BG->clock_ms + (long long)nint_(1000.0 * BG->genie_timestep);
}
double cgo_biogem_scalar(cgo_t *o, const char *name) {
  if (!BG) return NAN;
  if (!strcmp(name, "bg_clock_ms")) return (double)BG->clock_ms;
  if (!strcmp(name, "bg_go")) return BG->go;
  if (!strcmp(name, "bg_LS")) return BG->LS;
  if (!strcmp(name, "bg_LA")) return BG->LA;
  return NAN;
}
