/* cgo_impl.h -- internal state of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 * Arrays keep the reference's Fortran shapes; 1-D arrays are over-allocated
 * and indexed directly with the Fortran index, multi-D arrays go through the
 * macros below (column-major, Fortran lower bounds). */
#ifndef CGO_IMPL_H
#define CGO_IMPL_H
#include "cgo.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* goldstein_lib.f90:48-96 / embm_lib.f90:36-124 */
#define CG_PI (4.0 * atan(1.0))
#define CG_USC 0.05
#define CG_RSC 6.37e6
#define CG_DSC 5.0e3
#define CG_FSC (2 * 7.2921e-5)
#define CG_GSC 9.81
#define CG_RH0SC 1.0e3
#define CG_RHOSC (CG_RH0SC * CG_FSC * CG_USC * CG_RSC / CG_GSC / CG_DSC)
#define CG_TSC (CG_RSC / CG_USC)
#define CG_CPSC 3981.1
#define CG_RHOAIR 1.25
#define CG_RHO0 1.0e3
#define CG_RHOAO (CG_RHOAIR / CG_RHO0)
#define CG_M2MM 1000.0
#define CG_MM2M (1.0 / CG_M2MM)
#define CG_RFLUXSC (CG_RSC / (CG_DSC * CG_USC * CG_RH0SC * CG_CPSC))
#define CG_CPA 1004.0
#define CG_CONST1 3.80e-3
#define CG_CONST2 21.87
#define CG_CONST3 265.5
#define CG_CONST4 17.67
#define CG_CONST5 243.5
#define CG_SIGMA 5.67e-8
#define CG_EMO (0.94 * CG_SIGMA)
#define CG_EMA (0.85 * CG_SIGMA)
#define CG_TFREEZ 0.0
#define CG_HLV 2.501e6
#define CG_HLF 3.34e5
#define CG_HLS (CG_HLV + CG_HLF)
#define CG_CONSIC 2.166
#define CG_ZEROC 273.15
#define CG_CPO_ICE 4044.0
#define CG_RHOICE 913.0
#define CG_HMIN 0.01
#define CG_RHMIN (1.0 / CG_HMIN)
#define CG_RHOOI (CG_RHO0 / CG_RHOICE)
#define CG_RHOIO (CG_RHOICE / CG_RHO0)
#define CG_RRHOLF (1.0 / (CG_RHOICE * CG_HLF))
#define CG_CO20 278.0e-6
#define CG_CH40 700.0e-9
#define CG_N2O0 275.0e-9
#define CG_ALPHACH4 0.036
#define CG_ALPHAN2O 0.12
#define CG_TSIC (-1.8)
#define CG_CD 0.0013

#define CGO_MAXL 127   /* tracers a column routine keeps on its stack */

struct cgo_field_ent { const char *name; double *p; long n; };
struct cgo_ifield_ent { const char *name; int *p; long n; };
struct cgo_scalar_ent { const char *name; double *p; };
struct cgo_bg;

struct cgo {
  int maxi, maxj, maxk, maxl;
  /* ---------------- parameters (namelists) ---------------- */
  int igrid, nyear, ndta;
  double yearlen, temp0, temp1, rel, scf, diff[3], adrag_in;
  double hosing, hosing_trend; int nyears_hosing, nsteps_hosing;
  double albocn; int iconv, imld, iediff, ieos, diso;
  double ssmaxsurf, ssmaxdeep, saln0;
  double ediff0, ediffpow1, ediffpow2, ediffvar; int ediffpow2i;   /* iediff = 1 | 2 (goldstein.f90:2936-3044) */
  double mldpebuoycoeff, mldketaucoeff, mldwindkedec;                /* imld = 1 (goldstein.f90:1667-1686) */
  double *dzg, *z2dzg, *rdzg, *mlddec, *mlddecd;                      /* (maxk,maxk) x 3, (maxk) x 2 */
  double *mldketau, *mldpelayer1, *mldpeconv, *mldpebuoy, *mldemix, *mld; int *mldk;   /* (maxi,maxj) */
  double rmax, diffamp[3], diffwid, difflin, betaz[3], betam[3];
  double tatm, relh0_ocean, relh0_land, extra1a, extra1b, extra1c, scl_fwf;
  double z1_embm, diffa_scl; int diffa_len;
  double delf2x, olr_adj0, olr_adj, t_eqm;
  double albedop_offs, albedop_amp, albedop_skew; int albedop_skewp;
  double albedop_mod2, albedop_mod4, albedop_mod6;
  double par_sich_max, par_albsic_min, par_albsic_max; int par_wind_polar_avg;
  double radfor_scl_co2, radfor_pc_co2_rise, radfor_scl_ch4,
      radfor_pc_ch4_rise, radfor_scl_n2o, radfor_pc_n2o_rise;
  double diffsic_in, par_sica_thresh, par_sich_thresh;
  double solconst, gn_daysperyear;
  int kocn_loop, katm_loop, ksic_loop;

  /* ---------------- GOLDSTEIN ---------------- */
  int isles, ntot, intot, jsf, mpi, nm;
  int *k1, *ku, *mk, *ips, *ipf, *ias, *iaf, *getj;
  int *npi, *lpisl, *ipisl, *jpisl;
  double dphi, rdphi, dzz, ec[6], rpmesco, rsictscsf, cd, adrag;
  double dmax; int limps;
  double *dt, *ds, *dsv, *rds2, *dz, *s, *c, *sv, *cv, *dza, *zro, *zw;
  double *rc, *rc2, *rcv, *rdsv, *cv2, *rds, *rdz, *rdza, *asurf, *ssmax;
  double *ediff1p, *diffmax;            /* ediff1(i,j,k) = ediff1p(k) (ediffvar = 0), diffmax(k) */
  double *tf_scratch;                   /* work arrays of cgo_tstepo_flux, allocated once */
  double *rtv, *rtv3, *u, *u1, *ts, *ts1, *rho, *tau, *drag, *dztau, *dztav;
  double *ratm, *gap, *gb, *gbold, *ub, *psi, *rh, *cost, *bp, *sbp;
  double *fw_hosing, *rhosing, *fw_anom, *fw_anom_rate;
  double *psisl, *ubisl, *erisl, *psibc, *albcl_go, *dzu;
  double *psiles;
  int istep_ocn, istep_atm, istep_sic; long koverall;
  int go_lfirst; double go_ini_energy, go_ini_water;
  double test_energy_ocean, test_water_ocean;

  /* ---------------- EMBM + surflux ---------------- */
  double dtatm, rdtdim, ryear, rfluxsca, rpmesca, ppmin, ppmax, hatmbl[3];
  double rate_co2, rate_ch4, rate_n2o;
  double *tq, *tq1, *tqa, *uatm, *diffa, *albcl, *ca, *co2, *ch4, *n2o;
  double *usurf, *pmeadj, *pptn, *evap, *fxsw, *fxplw, *fx0a, *fx0o, *fxsen,
      *fxlata, *fxlw, *qb, *qbsic, *fx0sic, *fx0neto_eb, *evapsic, *tsfreez,
      *qsata, *qsato, *q_pa, *rq_pa, *solfor, *us_dztau, *us_dztav;
  double *eb_tau, *eb_dztau, *eb_dztav; /* EMBM's own copies (embm.f90:2762-2773) */
  int *iroff, *jroff;

  /* ---------------- sea ice ---------------- */
  double dtsic, sic_rdtdim, diffsic;
  double *varice, *varice1, *dtha, *sic_u;

  /* ---------------- BIOGEM (tracer coupling) ---------------- */
  double *bg_ocn, *bg_vdocn, *bg_M, *bg_rM, *bg_V;
  struct cgo_bg *bg;      /* full BIOGEM/ATCHEM state (cgo_biogem.c), NULL unless cgo_biogem_setup was called */
  double *go_solfor;      /* (maxj) solfor of the last surflux call (embm.f90:3728) */
  char *params;           /* copy of the construction parameters */

  /* ---------------- coupling arrays (genie_global) ---------------- */
  double *tstar_ocn, *sstar_ocn, *ustar_ocn, *vstar_ocn, *albedo_ocn;
  double *tstar_atm, *qstar_atm, *hght_sic, *frac_sic, *temp_sic, *albd_sic;
  double *stressxu, *stressyu, *stressxv, *stressyv;
  double *latent_ocn, *sensible_ocn, *netsolar_ocn, *netlong_ocn, *evap_ocn,
      *precip_ocn, *runoff_ocn, *runoff_land, *latent_atm, *sensible_atm,
      *netsolar_atm, *netlong_atm, *evap_atm, *precip_atm, *dhght_sic,
      *dfrac_sic, *waterflux_ocn, *conductflux_ocn, *lowestlu2, *lowestlv3;

  /* registry */
  struct cgo_field_ent fields[256]; int nfields;
  struct cgo_ifield_ent ifields[32]; int nifields;
  struct cgo_scalar_ent scalars[64]; int nscalars;
};

#define NI (o->maxi)
#define NJ (o->maxj)
#define NK (o->maxk)
#define NL (o->maxl)

#define K1(i, j) o->k1[(i) + (NI + 2) * (j)]
#define KU(l, i, j) o->ku[((l)-1) + 2 * (((i)-1) + NI * ((j)-1))]
#define MK(i, j) o->mk[((i)-1) + (NI + 1) * ((j)-1)]
#define GETJ(i, j) o->getj[((i)-1) + NI * ((j)-1)]
#define U(l, i, j, k) o->u[((l)-1) + 3 * ((i) + (NI + 1) * ((j) + (NJ + 1) * ((k)-1)))]
#define U1(l, i, j, k) o->u1[((l)-1) + 3 * ((i) + (NI + 1) * ((j) + (NJ + 1) * ((k)-1)))]
#define TS(l, i, j, k) o->ts[((l)-1) + NL * ((i) + (NI + 2) * ((j) + (NJ + 2) * (k)))]
#define TS1(l, i, j, k) o->ts1[((l)-1) + NL * ((i) + (NI + 2) * ((j) + (NJ + 2) * (k)))]
#define RHO(i, j, k) o->rho[(i) + (NI + 2) * ((j) + (NJ + 2) * (k))]
#define TAU(l, i, j) o->tau[((l)-1) + 2 * (((i)-1) + NI * ((j)-1))]
#define DZTAU(l, i, j) o->dztau[((l)-1) + 2 * (((i)-1) + NI * ((j)-1))]
#define DZTAV(l, i, j) o->dztav[((l)-1) + 2 * (((i)-1) + NI * ((j)-1))]
#define DRAG(l, i, j) o->drag[((l)-1) + 2 * (((i)-1) + (NI + 1) * ((j)-1))]
#define UB(l, i, j) o->ub[((l)-1) + 2 * ((i) + (NI + 2) * (j))]
#define PSI(i, j) o->psi[(i) + (NI + 1) * (j)]
#define RH(l, i, j) o->rh[((l)-1) + 3 * ((i) + (NI + 2) * (j))]
#define BP(i, j, k) o->bp[((i)-1) + (NI + 1) * (((j)-1) + NJ * ((k)-1))]
#define SBP(i, j) o->sbp[((i)-1) + (NI + 1) * ((j)-1)]
#define GAP(k, l) o->gap[((k)-1) + (long)o->nm * ((l)-1)]
#define RATM(k, l) o->ratm[((k)-1) + (long)o->nm * ((l)-1)]
#define A2(a, i, j) (a)[((i)-1) + NI * ((j)-1)]         /* (maxi,maxj)   */
#define A3(a, l, i, j) (a)[((l)-1) + 2 * (((i)-1) + NI * ((j)-1))] /* (2,maxi,maxj) */
#define DIFFA(l, m, j) o->diffa[((l)-1) + 2 * (((m)-1) + 2 * ((j)-1))]
#define SOLFOR(j, n) o->solfor[((j)-1) + NJ * ((n)-1)]

static inline int imax2(int a, int b) { return a > b ? a : b; }
static inline int imin2(int a, int b) { return a < b ? a : b; }
static inline double dmax2(double a, double b) { return a > b ? a : b; }
static inline double dmin2(double a, double b) { return a < b ? a : b; }
/* Fortran SIGN(1,x) for integers */
static inline int isign1(int x) { return x >= 0 ? 1 : -1; }
/* Fortran NINT */
static inline int nint_(double x) { return (int)(x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5)); }
/* x**n for integer n as gfortran lowers it (__builtin_powi: square-and-multiply) */
static inline double powi_(double x, int m) {
  unsigned n = m < 0 ? -(unsigned)m : (unsigned)m;
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) { x = x * x; if (n % 2) y = y * x; }
  return m < 0 ? 1.0 / y : y;
}

/* module internals shared between translation units */
void cgo_goldstein_init(cgo_t *o);
void cgo_embm_init(cgo_t *o, const double *taux_u, const double *tauy_u,
                   const double *taux_v, const double *tauy_v,
                   const double *uncep, const double *vncep);
void cgo_seaice_init(cgo_t *o);
void cgo_biogem_init(cgo_t *o);
void cgo_biogem_setup(cgo_t *o, const char *params);
int cgo_biogem_koverall(cgo_t *o, long k);
void cgo_biogem_tick(cgo_t *o);
double cgo_biogem_scalar(cgo_t *o, const char *name);
void cgo_eos(const cgo_t *o, double t, double s, double z, double *rho);
void cgo_reg(cgo_t *o, const char *name, double *p, long n);
void cgo_ireg(cgo_t *o, const char *name, int *p, long n);
void cgo_sreg(cgo_t *o, const char *name, double *p);
double *cgo_alloc(cgo_t *o, const char *name, long n);
int *cgo_ialloc(cgo_t *o, const char *name, long n);

#endif
