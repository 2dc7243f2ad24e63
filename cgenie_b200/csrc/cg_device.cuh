// cg_device.cuh -- device-side data layout of the B200 hot path.
//
// HBM layout (all fp64): structure-of-arrays with the ENSEMBLE MEMBER as the fastest axis,
// tracer next:
//     ts  [k][j][i][l][m]      rho [k][j][i][m]      u [k][j][i][c][m]      2-D [j][i][m]
// over the interior cells only (i=1..I, j=1..J, k=1..K; the periodic i-halo of the reference is
// index arithmetic, the j/k boundary rows are never read because every neighbour access is
// guarded by the bathymetry mask exactly as in the Fortran).  A warp is 32 members of ONE cell
// and tracer: fully coalesced 256-byte rows, and near-uniform control flow because all members
// share the bathymetry.  `MS` (member stride) = members rounded up to a whole number of 32-member warps (256-byte rows);
// padding lanes carry copies of the last member, so kernels need no member guard for safety.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cg {

constexpr int kMaxJ = 132;  // metric arrays in constant memory (config #5: 128 x 128 x 32)
constexpr int kMaxK = 36;
constexpr int kMaxIsles = 8;  // islands of one topography (the reference's worlds have up to 5)

// Member-independent grid metrics, uploaded to __constant__ memory of every kernel TU.
struct GridC {
  int I, J, K, L, M, MS, nyear, ndta, isles, npi1;
  double dphi, rdphi, dzz, dt;
  double ds[kMaxJ], dsv[kMaxJ], rds2[kMaxJ], s[kMaxJ], c[kMaxJ], sv[kMaxJ], cv[kMaxJ], rc[kMaxJ], rc2[kMaxJ],
      rcv[kMaxJ], rdsv[kMaxJ], cv2[kMaxJ], rds[kMaxJ];
  double dz[kMaxK], dza[kMaxK], rdz[kMaxK], rdza[kMaxK], zw[kMaxK], ssmax[kMaxK];
  double zro[kMaxK];       // depth of the density levels (ieos = 1: rho at zro(k))
  double diffmax[kMaxK];   // iediff > 0: 0.5 * 0.125 * dz^2 / dt (goldstein.f90:3042)
  double mlddec[kMaxK], mlddecd[kMaxK];   // imld = 1: wind-energy decay with depth (goldstein.f90:1675-1686)
};

// Per-member scalar parameters ([MS] each) and per-member 2-D constants ([j][i][m]).
struct MemberP {
  const double *diff1, *diff2, *ec1, *ec2, *ec3, *ec4, *ec5, *rel, *scf, *saln0, *rpmesco, *rsictscsf, *albocn;
  const double *hosing_trend;
  const int *nsteps_hosing;
  const double *ediff0, *ediff1p;   // iediff > 0: [m], [k][m] (ediff1(i,j,k) = ediff1p(k): ediffvar = 0)
  // EMBM / surflux
  const double *dtatm, *rdtdim, *rfluxsca, *rpmesca, *rmax, *betaz1, *betaz2, *betam1, *betam2, *ppmin, *ppmax,
      *delf2x, *olr_adj0, *olr_adj, *t_eqm, *par_sich_max, *par_albsic_min, *par_albsic_max, *rate_co2, *rate_ch4,
      *rate_n2o, *hatmbl2;
  const double *diffa;   // [j][4][m]  (l,m2) -> (l-1)+2*(m2-1)
  // sea ice
  const double *dtsic, *sic_rdtdim, *diffsic, *par_sica_thresh, *par_sich_thresh;
};

// Everything a kernel needs, passed by value.
struct Dev {
  int I, J, K, L, M, MS;
  int nm;         // psi points I*(J+1)
  int nbaro;      // number of distinct barotropic factorisations
  // masks (member independent)
  const unsigned char *k1;   // (0:I+1, 0:J+1)
  const unsigned char *ku;   // (2, I, J)
  const unsigned char *mk;   // (I+1, J)
  const unsigned char *getj; // (I, J)
  const int *iroff_src;      // runoff gather lists: CSR over wet cells, sources in reference order
  const int *iroff_ptr;
  const int *lpisl, *ipisl, *jpisl;  // island paths, [island][mpi]
  const int *npi;                    // [island] points on each path
  int isles, mpi;                    // islands (1 .. kMaxIsles), path stride
  const int *wetcols;        // 0-based (i-1)+I*(j-1) of the wet columns, deepest first
  const int *rowcols;        // the same columns in row-major (j, then i) order: neighbours in the list are neighbours in i
  const int *polcols;        // the same columns, poleward rows first (rows J, 1, J-1, 2 ...): where the ocean convects
  int nwet;
  // tracer state, ping-pong
  double *ts_cur, *ts_new;
  double *tsflux;            // [2][j][i][m] surface flux b.c. for T,S (ts(1:2,:,:,maxk+1))
  double *sst;               // [2][j][i][m] tstar_ocn/sstar_ocn as exported by step_goldstein (goldstein.f90:428-431);
                             // NULL = read ts directly (identical unless BIOGEM rewrites ts in between)
  double *rho, *u, *u1, *cost;
  double *velsum;            // [2][j][i][m] depth means of the baroclinic velocities (k_velc1 -> k_velc2)
  double *usnap;             // [2][j][i][m] surface u, v as exported to the sea-ice step (snapshot taken by k_usnap)
  // momentum
  double *bp, *sbp, *gb, *ub, *psi, *erisl_rhs, *psibc;
  const double *rh;          // (3, 0:I+1, 0:J+1) member independent
  const double *drag, *rtv, *rtv3;   // per member: (2,I+1,J)[m], (I,J)[m]
  const double *tau, *dztau, *dztav; // per member (2,I,J)[m]
  const double *gap, *ratm;  // per factor group: [g][nm][2I+3], [g][nm][I+1]  (row-major band rows)
  const double *ubisl, *psisl, *erisl; // per group
  const int *baro_group;     // [MS]
  const double *rhosing;     // (I,J) member independent
  double *hosing, *fw_anom;  // [MS], unused anomaly kept for completeness
  // atmosphere / surface fluxes / sea ice ([j][i][m] unless noted)
  double *tq, *tq1;          // [2][j][i][m]
  double *tqa;               // [2][j][i][m]
  const double *uatm_u, *uatm_v, *usurf, *albcl, *ca, *pmeadj;
  const double *solfor;      // [nyear][J] member independent
  double *co2, *ch4, *n2o;   // [j][i][m]
  double *pptn, *evap, *evapsic, *fx0a, *fx0o, *fxsen, *fxlw, *fxsw, *fxplw, *tice, *albice, *albedo;
  double *latent_ocn, *sensible_ocn, *netsolar_ocn, *netlong_ocn, *evap_ocn, *precip_ocn, *runoff_ocn, *runoff_land;
  double *latent_atm, *sensible_atm, *netsolar_atm, *netlong_atm, *evap_atm, *precip_atm, *dhght_sic, *dfrac_sic;
  double *varice, *varice1;  // [2][j][i][m]
  double *waterflux_ocn, *conductflux_ocn;
  double *q_pa, *rq_pa;
  // BIOGEM tracer coupling (biogem.f90:1885-2077)
  double *bg_ocn, *bg_vdocn;   // [k][j][i][l][m] BIOGEM-unit tracers and the step anomaly
  double *bg_M, *bg_rM;        // [k][j][i][m] cell mass and reciprocal (rescaled by the salinity ratio per member)
  const double *bg_V;          // [k][j][i] cell volume (member independent)
  const int *bgcols;           // wet columns in BIOGEM order (i outer, j inner), 0-based (i-1)+I*(j-1)
  double *bg_part, *bg_tot;    // reduction scratch: [q][n][m] partial sums, [q][m] totals
  double bg_rtot_V;
  double *bg_biopart;          // BIOGEM particulates [k][j][i][ls][m] (rescaled by tracer coupling), NULL without BIOGEM
  int bg_LS;
  unsigned *comask;          // [j][i][m] region map of the convective adjustment (k_ts_pre -> k_tstep_col passive pass):
                             // bit k-1 = level k lies in a mixed region, bit 16+k-1 = it is the region's top
  int co_prefetch;           // k_co_col: prefetch the column's passive tracers during the decisions (tuning knob)
  int co_local;              // k_co_col: column arrays of the decision loop in local instead of shared memory (CG_CO_LOCAL=1)
  int co_pairwise;           // k_co_col: average the passive tracers pair by pair (round-1 form, CG_CO_PAIR=1) instead of region by region
  int co_skip_stable;        // k_co_col: skip (member, column)s the flux kernel flagged stable in comask (CG_CO_SKIP=0: off)
  int col_deep_first;        // k_tstep_col: blocks in wetcols order (deepest columns first) instead of row-major (CG_COL_ORDER=1)
  int iconv;                 // 1: Mueller convection scheme (coshuffle + depth diagnostic, goldstein.f90:2667-2672, 2781-2841), strict kernels only
  int ieos;                  // 1: thermobaricity term in the equation of state (goldstein.f90:3048-3082), strict kernels only
  int iediff, ediffpow2i;    // stratification-dependent vertical diffusivity (goldstein.f90:2501-2515); 0 = constant diff(2)
  double ediffpow2;
  int imld;                  // 1: Kraus-Turner mixed-layer scheme behind the convective adjustment (tstepo, goldstein.f90:2294-2390)
  double mldpebuoycoeff;
  const double *mldketau;    // [j][i][m] wind energy input (per member: tau scales with scf)
  double *mld_pel1;          // [j][i][m] mldpelayer1: energy of mixing the surface forcing over the top layer
  double *mld_rhoold;        // [k][j][i][m] density of the column before the convective adjustment (mldrhoold)
  double *mld;               // [j][i][m] mixed-layer depth (<= 0, units of dsc); go_mldta = -5000 * mld
  int *mldk;                 // [j][i][m] level holding the mixed-layer base
  int *istep_ocn;            // device-resident ocean step counter (read by graph-replayed kernels)
  MemberP p;
};

// ---- index helpers ----
__host__ __device__ inline size_t cell3(const int I, const int J, int i, int j, int k) {
  return (size_t)((k - 1) * J + (j - 1)) * I + (i - 1);
}
__host__ __device__ inline size_t cell2(const int I, int i, int j) { return (size_t)(j - 1) * I + (i - 1); }

#define CG_K1(v, i, j) ((int)(v).k1[(i) + ((v).I + 2) * (j)])

}  // namespace cg

// ---------------------------------------------------------------------------------------------------------------
// BIOGEM / ATCHEM device view (passed by value next to Dev).  Compact tracer indices as in the reference's
// conv_iselected_io/is/ia maps (gem_cmn.f90:366-381): l = 1..L ocean, ls = 1..LS particulate, la = 1..LA atmosphere.
namespace cg {
constexpr int kBgMaxL = 16, kBgMaxLS = 9, kBgMaxLA = 8, kBgMaxK = 16, kBgMaxRel = 3;
// surface-cell results handed from k_bg_step PART 1 to PART 2: gas fluxes into the ocean, export, DOM production, uptake;
// then the deferred side effects ([H+] seed, focnatm and sfxatm1 of each gas) and the error flag (last slot)
constexpr int kBgSurfH = (kBgMaxLA - 2) + kBgMaxLS + 4 + 7;
constexpr int kBgSurfSlots = kBgSurfH + 1 + 2 * (kBgMaxLA - 2) + 1;
struct BgDev {
  int LS, LA;
  // tracer relationships: conv_ls_lo_i / conv_ls_lo (sed -> ocean, io ascending), DOM <-> POM, atm -> ocean
  int n_ls_lo[kBgMaxLS + 1], ls_lo[kBgMaxLS + 1][kBgMaxRel];
  double conv_ls_lo[kBgMaxLS + 1][kBgMaxRel];
  int dom2pom[kBgMaxL + 1], pom2dom[kBgMaxLS + 1], atm2ocn[kBgMaxLA + 1];
  int stype[kBgMaxLS + 1], sdep_ls[kBgMaxLS + 1], sdep_id[kBgMaxLS + 1];
  int atype[kBgMaxLA + 1], aid[kBgMaxLA + 1], adep[kBgMaxLA + 1];
  int lrem_slot[kBgMaxL + 1], n_lrem;     // ocean tracers that receive remineralisation products -> slot in the local accumulator
  int l_DIC, l_DIC13, l_DIC14, l_PO4, l_O2, l_ALK, l_DOMC, l_Ca, l_Mg;
  int s_POC, s_POC13, s_POC14, s_POP, s_CaCO3, s_CaCO313, s_CaCO314, s_POCf2, s_CaCO3f2;
  int a_CO2, a_CO213, a_CO214;
  // decay / relaxation factors of this step (host: exp(-dtyr*lambda), 1-exp(-dtyr/tau))
  double fd_ocn[kBgMaxL + 1], fd_sed[kBgMaxLS + 1], fd_atm[kBgMaxLA + 1], lam_ocn[kBgMaxL + 1], lam_sed[kBgMaxLS + 1],
      lam_atm[kBgMaxLA + 1], tmod[kBgMaxLA + 1];
  int rst_active[kBgMaxLA + 1];           // restoring selected and inside the signal interval
  double rst_target[kBgMaxLA + 1];        // force_restore_atm at wet points (uniform: I = 0, II = 1)
  double Sc[kBgMaxLA + 1][4], bunsen[kBgMaxLA + 1][6];
  // parameters
  double c0_PO4, red_POP_POC, red_DOMfrac, red_RDOMfrac, red_POC_CaCO3_pP, DOMlifetime, POC_frac2, POC_dfrac2, POC_c0frac2,
      CaCO3_frac2, sinkingrate, remin_k_O2, remin_c0_O2, gastransfer_a, d13C_DIC_Corg_ef, Fgeothermal, solar_constant, dsc;
  double dts, dtyr, dts_atchem, dtyr_atchem;
  double Dbot[kBgMaxK + 1], dD[kBgMaxK + 2], Dmid_surf, Dmid[kBgMaxK + 1];
  double CaCO3_f1[kBgMaxK + 1], CaCO3_f2[kBgMaxK + 1], POC_f2[kBgMaxK + 1];   // 1-exp(-dD(k)/eL), host computed
  const double *POC_f1;                   // [k][m] (par_bio_remin_POC_eL1 is a perturbed parameter)
  const double *k0_PO4, *red_POC_CaCO3;   // [m]
  int nsol;                               // solfor row (1-based) of the last biogem_climate call, 0 = none yet
  // state
  double *bio_part;                       // [k][j][i][ls][m]
  double *settle_k1;                      // [j][i][ls][m]   bio_settle at the deepest wet level
  double *carbH;                          // [j][i][m]       surface [H+], seed of the next pH solve
  double *seaice;                         // [j][i][m]       snapshot taken by biogem_climate
  double *seaice_stage;                   // [j][i][m]       sea-ice cover at biogem_climate's call time (k_bg_stage_seaice)
  double *mld, *mld_stage;                // [j][i][m]       imld = 1: mixed-layer depth biogem_climate takes over (m below the surface,
                                          //                 go_mldta = -5000 * mld) and GOLDSTEIN's mld at that call's time; else NULL
  double *atm_tot;                        // [la][m]         mole-weighted totals of the ATCHEM step (k_bg_atchem2 -> k_bg_atchem3)
  double *tq_stage;                       // [2][j][i][m]    EMBM's T, q at the same time: what cpl_comp_EMBM copies (atchem.f90:270-282)
  const double *wspeed, *A, *rA;          // [j][i] member independent
  double *atm, *sfcatm1, *sfxsumatm;      // [la][j][i][m]
  const double *atm_A, *atm_V;            // [j][i]
  double *sfcocn1, *sfxsed1, *focnatm;    // interface / diagnostics: [l|ls|la][j][i][m]
  double *sfxatm1;                        // [la][j][i][m] sfxatm1 as step_biogem leaves it (biogem.f90:1731-1734); NULL unless the extended
                                          // time-series integrals are on (cg_biogem_sig_extended)
  double *settle_sur;                     // [j][i][ls][m] bio_settle(:,i,j,n_k) of this step (k_bg_settle_sur); NULL likewise
  int *err;                               // [m] carbonate chemistry failure flag (error_stop)
  // packets / cells split of the sweep (k_bg_step PART 3 -> k_bg_cell): remineralisation products of the sinking particles per
  // cell, sediment return per column, wet-column index of a column, pending rescaling of bio_part (see k_bg_cell)
  double *lrem;                           // [k][j][i][7][m]
  double *fsedv;                          // [7][wet column][m]
  const int *colidx;                      // [j][i] -> index in Dev::bgcols, -1 on land
  double *pscale;                         // [m] Sratio of the last coupling, not yet applied to bio_part
  double *surf;                           // [kBgSurfSlots][wet column][m] surface-cell results (k_bg_step PART 1 -> PART 2)
};
// time-slice diagnostics (diag_biogem_timeslice, biogem.f90:2421-2699): the carbonate system of every wet cell and the window
// integrals.  carbH3 / rf03 exist from cg_initialise on (sub_init_carb solves every cell); the integrals on first use.
constexpr int kSlCarb = 10, kSlCC = 17, kSlIso = 8;   // carb(ic_*), carbconst(icc_*), carbisor(ici_*) entries the path fills
struct SliceDev {
  const int *wet;                         // wet cells in ascending cell order
  int nwet3;
  double *carbH3, *rf03;                  // [k][j][i][m] [H+] of the cell's last solve (k < K; the surface cell's is BgDev::carbH), RF0
  double *ocn, *part, *carb, *cc, *iso;   // int_*_timeslice, [cell][q][m]
  double *t;                              // [m] int_t_timeslice
};
// window integrals of BIOGEM's time series (k_bg_sig_*): per-quantity rows of MS members
constexpr int kSigHead = 3;
struct SigDev {
  const int *kbot;                        // [j][i] 0-based level of the bottom cell, K on land
  const double *A, *w_ben;                // [j][i] cell area (all cells); benthic mask * area
  double *raw, *acc;                      // [q][m] sums of this step; integrals of the window
  double rtot_A_ben, rtot_A_atm;          // 1 / SUM(mask_ben * A), 1 / SUM(phys_ocnatm(ipoa_A))
  int LA;
  // extended integrals ("bg_sig2": sea ice, overturning, land temperature, export, air-sea fluxes), on when acc2 != NULL
  double *raw2, *acc2;                    // [q][m]
  double *opsi_stage;                     // [4][m] min / max of the global and the Atlantic overturning stream function at call time
  double *th_stage;                       // [j][i][m] sea-ice thickness at call time
  const int *ias, *iaf;                   // [J + 2] Atlantic columns per row
  int jsf, LS;
  double land_A;                          // SUM(phys_ocnatm(ipoa_A)) over the land cells
};
// layout of "bg_sig2": 0 seaice area, 1 mean thickness, 2 volume, 3 / 4 opsi min / max, 5 / 6 opsia min / max, 7 land air temperature,
// 8 + ls fexport, 8 + LS + la focnatm, 8 + LS + LA + la air-sea gas exchange (la >= 2, 0-based: the gases)
constexpr int kSig2Head = 8;
}  // namespace cg
