/*
 * cgo.h -- CPU ORACLE for the cGENIE hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement, routine for routine and in the reference's own loop
 * and expression order, of the Fortran hot path named in BASELINE.json:
 *   GOLDSTEIN  src/goldstein/goldstein.f90  (init 514-2084, step 17-479,
 *              tstepo 2280-2432, tstepo_flux 2436-2642, co 2657-2777,
 *              drgset 2845-2885, eos/eosd 3048-3082, get_hosing 3086-3129,
 *              invert 3138-3204, jbar 3217-3315, matinv/matmult 3452-3492,
 *              ubarsolv 3500-3565, velc 3568-3679, wind 3685-3714)
 *              src/goldstein/goldstein_lib.f90 (constants 48-96, island 186-241)
 *   EMBM       src/embm/embm.f90 (init 198-2018, step_embm 22-195,
 *              tstipa 2039-2138, radfor 2383-2522, surflux 2548-3738,
 *              readroff 3787-3836), src/embm/embm_lib.f90 (constants 36-124)
 *   SEA ICE    src/goldsteinseaice/gold_seaice.f90 (init 17-508,
 *              step_seaice 511-733, tstepsic 844-929)
 *   driver     src/genie.f90 253-470 (coupling schedule)
 *
 * PARITY STATUS: "parity unpinned".  The reference tree holds no golden
 * vectors for this path and no Fortran compiler exists in this image, so the
 * oracle cannot be checked against a gfortran build here.
 *
 * Arithmetic model: IEEE fp64, no FMA contraction (compile with
 * -ffp-contract=off, no -ffast-math), integer powers as repeated multiplies,
 * real powers through libm pow -- mirroring gfortran -O3 -fdefault-real-8
 * on baseline x86-64 (platforms/LINUX:7-17).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (cgenie_b200/) never links or calls it.
 */
#ifndef CGO_H
#define CGO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgo cgo_t;

/* Create one model instance (one ensemble member).
 *  params : "key=value\n" list overriding the reference namelist defaults
 *           (data_GOLD / data_EMBM / data_goldSIC / data_BIOGEM keys)
 *  k1     : (maxj+2) rows x (maxi+2) ints in FILE order (first row j=maxj+1)
 *  psiles : (maxj+1) rows x maxi values in file order (first row j=maxj)
 *  npaths : number of islands in paths file, npi[isl] points each,
 *  paths  : concatenated (dir,i,j) triples
 *  taux_u,tauy_u,taux_v,tauy_v : maxi*maxj values, i fastest (embm.f90:845-860)
 *  uncep,vncep : maxi*maxj wind speeds, i fastest (embm.f90:989-995)        */
cgo_t *cgo_create(const char *params, const int *k1, const double *psiles,
                  int npaths, const int *npi, const int *paths,
                  const double *taux_u, const double *tauy_u,
                  const double *taux_v, const double *tauy_v,
                  const double *uncep, const double *vncep);
void cgo_destroy(cgo_t *);

/* module entry points, same order/meaning as genie_loop_wrappers.f90 */
void cgo_surflux(cgo_t *);        /* surflux_wrapper       :7-59    */
void cgo_embm_step(cgo_t *);      /* embm_wrapper          :61-86   */
void cgo_seaice_step(cgo_t *);    /* gold_seaice_wrapper   :94-113  */
void cgo_goldstein_step(cgo_t *); /* goldstein_wrapper     :122-151 */
/* sub-steps exposed for kernel-level parity tests */
void cgo_tstepo(cgo_t *);
void cgo_tstepo_flux(cgo_t *);
void cgo_co(cgo_t *);
void cgo_momentum(cgo_t *);       /* wind..velc of step_goldstein :198-233 */
void cgo_tstipa(cgo_t *);
/* BIOGEM: biogem_tracercoupling_wrapper (genie_loop_wrappers.f90:318-322); cgo_biogem_init builds ocn from ts */
void cgo_biogem_init(cgo_t *);
void cgo_biogem_tracercoupling(cgo_t *);
/* full BIOGEM + ATCHEM (frozen eb_go_gs_ac_bg configuration): fill field "bg_windspeed" first, then set up; afterwards
 * cgo_run also executes the BIOGEM/ATCHEM block of genie.f90:352-447.  params: extra "key=value\n" overrides or NULL. */
void cgo_biogem_setup(cgo_t *, const char *params);
void cgo_biogem_forcing(cgo_t *);
int cgo_biogem_step(cgo_t *);
void cgo_biogem_climate(cgo_t *);
void cgo_cpl_flux_ocnatm(cgo_t *);
void cgo_cpl_flux_ocnsed(cgo_t *, double dts);
void cgo_cpl_comp_ocnsed(cgo_t *, int ocnstep, int mbiogem, int msedgem);
void cgo_reinit_flux_rokocn(cgo_t *);
void cgo_biogem_sig_update(cgo_t *, double ben_Dmin);
void cgo_biogem_sig_auto(cgo_t *, int on, double ben_Dmin);
void cgo_biogem_slice_update(cgo_t *);
void cgo_biogem_slice_auto(cgo_t *, int on);
void cgo_atchem_step(cgo_t *);

/* run n iterations of the genie.f90 koverall loop (one EMBM step each) */
void cgo_run(cgo_t *, long nkoverall);

/* named access to state/constant arrays; returns pointer, writes length */
double *cgo_field(cgo_t *, const char *name, long *n);
int *cgo_ifield(cgo_t *, const char *name, long *n);
double cgo_scalar(cgo_t *, const char *name);
void cgo_set_scalar(cgo_t *, const char *name, double v);

#ifdef __cplusplus
}
#endif
#endif
