/*
 * cgenie_b200.h -- C-ABI of the B200-native cGENIE hot path (libcgenie_b200.so).
 *
 * This is the drop-in boundary: one entry point per module procedure that the
 * reference's coupler calls each coupling step through the argument-less
 * wrappers of src/wrappers/genie_loop_wrappers.f90.  A Fortran shim module
 * with the reference's own public names (the files under fortran/) binds these through
 * ISO_C_BINDING; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - all reals are fp64, INTEGER is int32, the GENIE clock is int64 (ms);
 *  - host arrays are caller-owned, column-major, exactly the shapes the
 *    reference passes (citations per struct below);
 *  - a NULL array pointer (or a NULL io struct) means "leave the field
 *    resident on the GPU": between output/coupling intervals the host passes
 *    only scalars;
 *  - host arrays always carry ensemble member `io_member` (default 0, the
 *    control member the Fortran host drives); the other members advance in
 *    lock-step on the device and never cross the boundary per step;
 *  - every function returns 0 on success; non-zero maps to the reference's
 *    die()/write_status('ERRORED') (src/wrappers/genie_util.f90:15-33).
 *    cg_last_error() returns the message.
 */
#ifndef CGENIE_B200_H
#define CGENIE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cg_handle cg_handle;

enum {
  CG_OK = 0,
  CG_ERR_ARG = 1,          /* bad argument / unknown name            */
  CG_ERR_IO = 2,           /* namelist or data file unreadable       */
  CG_ERR_CONFIG = 3,       /* option outside the B200 hot path       */
  CG_ERR_CUDA = 4,         /* CUDA runtime error                     */
  CG_ERR_STATE = 5,        /* call order violated                    */
  CG_ERR_BLOWUP = 6        /* non-finite state (avs guard, goldstein_diag.f90:52-58) */
};

/* ---- life cycle ------------------------------------------------------- */

/* Parse <jobdir>/data_genie, data_GOLD, data_EMBM, data_goldSIC (and
 * data_BIOGEM/data_ATCHEM/data_GEM when flag_biogem) plus the data files they
 * name (<world>.k1/.psiles/.paths, wind stress *.interp, wind speed *.silo),
 * exactly as initialise_goldstein/initialise_embm/initialise_seaice do
 * (goldstein.f90:514-2084, embm.f90:198-2018, gold_seaice.f90:17-508).
 * Nothing is uploaded yet.
 * Physics options of data_GOLD: iediff = 0 | 1 | 2 with ediffvar = 0 (SUBROUTINE ediff, goldstein.f90:2936-3044),
 * ieos = 0 | 1 (thermobaric equation of state, :3048-3082), iconv = 0 | 1 (Mueller convection scheme, coshuffle :2781-2841) are
 * on the device path -- with one of them on, the tracer step runs the generic kernels (reference operation order), not the
 * shape-specialised column kernels; imld = 1 (krausturner), ediffvar /= 0, fwanomin = 'y', topographies with more than one
 * island, and BIOGEM selections other than the frozen 16-tracer one are refused with CG_ERR_CONFIG.
 * n_members: up to 128 members share one 128-wide member stride; 129 .. 256 and 257 .. 512 run on strides 256 / 512 (128-member
 * tiles of the same kernels, bit-identical per member); more than 512 fall back to the generic-shape kernels.          */
int cg_create(const char *jobdir, int n_members, int device, cg_handle **out);

/* Per-member override of a whitelisted scalar parameter BEFORE cg_initialise:
 * "diff1","diff2","adrag","scf","diffamp1","diffamp2","betaz2","betam2",
 * "rmax","temp0","temp1","diffsic", ... ; values[n_members].               */
int cg_set_member_param(cg_handle *, const char *name, const double *values);

/* Build constants (grid, masks, drag, barotropic factorisation, island
 * solves, insolation table ...) bit-exactly as the reference's initialise_*
 * routines, set the initial state and upload everything.                    */
int cg_initialise(cg_handle *);
int cg_destroy(cg_handle *);
const char *cg_last_error(void);

/* ---- module steps (genie_loop_wrappers.f90) --------------------------- */

/* surflux(...)  embm.f90:2548-2588 via surflux_wrapper :7-59.  All members. */
typedef struct {
  /* OUT (ocean side) */ double *albedo_ocn, *latent_ocn, *sensible_ocn, *netsolar_ocn, *netlong_ocn,
      *evap_ocn, *precip_ocn, *runoff_ocn, *runoff_land;
  /* OUT (atm side)   */ double *latent_atm, *sensible_atm, *netsolar_atm, *netlong_atm, *evap_atm, *precip_atm;
  /* OUT (sea ice)    */ double *dhght_sic, *dfrac_sic, *temp_sic, *albd_sic;
  /* INOUT            */ double *qstar_atm;
} cg_surflux_io; /* every array (maxi,maxj) */
int cg_surflux_step(cg_handle *, int istep, const cg_surflux_io *io);

/* step_embm(...)  embm.f90:22-38 via embm_wrapper :61-86 */
typedef struct {
  double *tstar_atm, *qstar_atm; /* OUT (maxi,maxj) */
} cg_embm_io;
int cg_embm_step(cg_handle *, int istep, const cg_embm_io *io);

/* step_seaice(...)  gold_seaice.f90:511-520 via gold_seaice_wrapper :94-113 */
typedef struct {
  double *hght_sic, *frac_sic, *waterflux_ocn, *conductflux_ocn; /* OUT (maxi,maxj) */
} cg_seaice_io;
int cg_seaice_step(cg_handle *, int istep, const cg_seaice_io *io);

/* step_goldstein(...)  goldstein.f90:17-38 via goldstein_wrapper :122-151 */
typedef struct {
  double *tstar_ocn, *sstar_ocn, *ustar_ocn, *vstar_ocn, *albedo_ocn; /* OUT (maxi,maxj)          */
  double *go_ts;   /* INOUT (maxl,maxi,maxj,maxk): uploaded before, downloaded after the step   */
  double *go_u;    /* OUT   (3,maxi,maxj,maxk)                                                  */
  double *go_rho;  /* OUT   (maxi,maxj,maxk)                                                    */
  double *go_cost; /* INOUT (maxi,maxj)                                                         */
  double *go_psi;  /* OUT   (0:maxi,0:maxj)                                                     */
  double *test_energy_ocean, *test_water_ocean; /* OUT scalars (goldstein.f90:458-478)          */
} cg_goldstein_io;
int cg_goldstein_step(cg_handle *, int istep, const cg_goldstein_io *io);
/* step_goldstein's go_mldta (goldstein.f90:449: -5000 * mld, metres; zero unless imld = 1, the Kraus-Turner mixed-layer scheme
 * of tstepo :2294-2390 and SUBROUTINE krausturner :3337-3442) of one member, (maxi,maxj). */
int cg_goldstein_mldta(cg_handle *, int member, double *go_mldta);
/* io == NULL ("stay resident") calls of the four steps above are deferred: they are executed, in order, at the latest when
 * the next call passes arrays or touches state from the host (cg_sync_*, cg_get_*, cg_synchronize, cg_run, BIOGEM / ATCHEM);
 * a complete cycle surflux, kocn_loop x step_embm, step_seaice, step_goldstein is executed at the step_goldstein call as
 * two CUDA-graph replays.  Results are bit-identical either way; an error of a deferred call is returned by the call that
 * triggers its execution. */

/* BIOGEM / ATCHEM (biogem.f90:528-547, 1885-1890, 2083-2087, 2132-2150; atchem.f90:63-67) */
int cg_biogem_forcing(cg_handle *, int64_t genie_clock_ms);
int cg_biogem_step(cg_handle *, double dts, int64_t genie_clock_ms);
int cg_biogem_tracercoupling(cg_handle *, double *go_ts, double *go_ts1);
int cg_biogem_climate(cg_handle *);
int cg_biogem_climate_sol(cg_handle *);   /* biogem_climate_sol, biogem.f90:2243-2263 (first BIOGEM step only) */
/* cpl_flux_ocnatm (atchem.f90:306-320) is fused into cg_biogem_step; kept so that the wrapper has a target */
int cg_cpl_flux_ocnatm(cg_handle *);
/* The SEDGEM / ROKGEM coupler calls genie.f90:413-427 makes after every BIOGEM step even when those modules are off
 * (SURVEY 8f row 2), on device-resident interface arrays (fields "sfxsumsed", "sfcsumocn", "sfxsumrok1"; sediment grid =
 * ocean grid):  cpl_flux_ocnsed (src/sedgem/sedgem.f90:1029-1068)  sfxsumsed += dts * sfxsed1;
 * cpl_comp_ocnsed (sedgem.f90:894-937) running mean of sfcocn1 into sfcsumocn, ocnstep = koverall / kocn_loop, mbiogem /
 * msedgem = conv_kocn_kbiogem / conv_kocn_ksedgem (genie_loop_wrappers.f90:197-226);  reinit_flux_rokocn
 * (src/rokgem/rokgem.f90:472-480) sfxsumrok1 = 0.  cg_run does not issue them (no consumer without SEDGEM). */
int cg_cpl_flux_ocnsed(cg_handle *, double dts);
int cg_cpl_comp_ocnsed(cg_handle *, int ocnstep, int mbiogem, int msedgem);
int cg_reinit_flux_rokocn(cg_handle *);
/* The 3-D / 2-D sums of diag_biogem_timeseries (src/biogem/biogem.f90:2703-3159; SURVEY 8f row 1, time-series part): one BIOGEM
 * step's contribution to the window integrals int_t_sig, int_ocn_tot_M_sig, int_ocn_tot_M_sur_sig, int_ocn_sig(:),
 * int_ocn_sur_sig(:), int_ocn_ben_sig(:), int_ocnatm_sig(:) (:2851-2917, :3082), on the device.  Call it where genie.f90 calls
 * diag_biogem_timeseries_wrapper (src/genie.f90:401-405): behind cg_biogem_climate, ahead of cg_atchem_step, on the steps that
 * fall in a save window (the window logic of :2760-2769 and biogem_save_sig.dat stay with the caller).  There the ocean sums see
 * this block's tracer coupling and sfcatm1 still holds what the PREVIOUS block's ATCHEM step / cpl_comp_atmocn / cpl_comp_EMBM
 * left (air temperature and humidity included: rows 1-2 of sfcatm1 are kept on the device).  Behind cg_atchem_step (or cg_run) the
 * ocean rows are the same and the atmosphere rows are one coupling interval newer.  Device and oracle integrals agree to 1e-13
 * at either point (tests/test_gpu_z_sig.py, tests/test_gpu_zz_series_year.py).  Field "bg_sig" (cg_sync_to_host) = the integrals
 * of one member in that order, 3 + 3*maxl + n_l_atm values, tracers in the compact selection order; cg_biogem_sig_reset =
 * sub_init_int_timeseries (biogem_data.f90:964-1007).  ben_Dmin = par_data_save_ben_Dmin (m). */
int cg_biogem_sig_update(cg_handle *, double dts, double ben_Dmin);
/* The flux, export and "misc" integrals of the same routine: after cg_biogem_sig_extended (call it before the first BIOGEM step
 * whose integrals are wanted) step_biogem keeps sfxatm1 and the export through the base of the surface layer, and
 * cg_biogem_sig_update also accumulates, as field "bg_sig2" of one member,
 *   [0] int_misc_seaice_sig  [1] ..._th  [2] ..._vol (biogem.f90:2926-2937)   [3] [4] int_misc_opsi_min / max_sig
 *   [5] [6] int_misc_opsia_min / max_sig (:2938-2945; sub_calc_psi, biogem_box.f90:3796-3852)   [7] int_misc_SLT_sig (:2946-2964)
 *   [8 + ls] int_fexport_sig (:2870-2876)   [8 + n_l_sed + la] int_focnatm_sig (:2877-2883)
 *   [8 + n_l_sed + n_l_atm + la] int_diag_airsea_sig (:3058-3062)     (0-based compact indices; la >= 2: the gases).
 * Not on the device: int_carb_sur_sig, the insolation / short-wave means, the sediment-interface and diag_bio / diag_geochem sums. */
int cg_biogem_sig_extended(cg_handle *);
/* diag_biogem_timeslice (src/biogem/biogem.f90:2421-2699; SURVEY 8f row 1, time-slice part), its arithmetic: inside a save
 * window the carbonate system of EVERY wet cell is solved again from the cell's last [H+] (:2478-2567; the surface cell's feeds
 * back into step_biogem's next solve, as in the reference) and the window integrals int_ocn, int_bio_part, int_carb,
 * int_carbconst, int_carbisor, int_t _timeslice grow by dtyr * field (:2572-2579).  Call it where genie.f90 calls
 * diag_biogem_timeslice_wrapper (:391-395: behind cg_biogem_climate, ahead of cg_biogem_sig_update and cg_atchem_step) on the
 * steps of a save window; the window bookkeeping and the netCDF writer stay with the host, which reads the fields "sl_ocn"
 * (maxl,i,j,k), "sl_part" (n_l_sed,...), "sl_carb" (10,...: H, CO2, CO3, HCO3, fug_CO2, ohm_cal, ohm_arg, dCO3_cal, dCO3_arg,
 * RF0), "sl_carbconst" (17,...: k1 k2 k kB kW kSi kHF kHSO4 kP1 kP2 kP3 kH2S kNH4 kcal karg QCO2 QO2), "sl_carbisor" (8,...)
 * and "sl_t".  cg_biogem_slice_reset = sub_init_int_timeslice.  Not on the device: the 2-D interface integrals (focnatm,
 * focnsed ...), the overturning stream functions and the diag_* arrays of :2580-2606. */
int cg_biogem_slice_update(cg_handle *, double dts);
int cg_biogem_slice_reset(cg_handle *);
int cg_biogem_sig_reset(cg_handle *);
/* sub_init_data_save_runtime / sub_data_save_runtime (src/biogem/biogem_data_ascii.f90:23-110, 669-935), the ocn_* and atm_*
 * series: <outdir>/<outfile_name>_series_ocn_<name>.res and ..._atm_<name>.res in the reference's formats
 * (header through list-directed output; f12.3 / f12.6 / e15.7 / f14.3 columns; T in degrees C; isotopes as delta values
 * through fun_calc_isotope_delta, gem_util.f90:568-598).  create != 0: (re)create the files with their header lines;
 * create == 0: append one line per file from the integrals `sig` of one member (the "bg_sig" layout).  *_type: 0 (T, S /
 * temp, humidity), 1 (bulk), 11 (13C), 12 (14C); *_dep: 0-based compact index of an isotope's bulk tracer; with_sur =
 * ctrl_data_save_sig_ocn_sur.  Host only. */
const char *cg_series_last_error(void);
int cg_biogem_series_write(const char *outdir, const char *outfile_name, int create, double t_yr, int n_ocn,
                           const char *const *ocn_names, const int32_t *ocn_type, const int32_t *ocn_dep, int n_atm,
                           const char *const *atm_names, const int32_t *atm_type, const int32_t *atm_dep, const double *sig,
                           int with_sur);
/* ... and the fexport_<sed>, fseaair_<atm>, focnatm_<atm>, misc_seaice, misc_opsi, misc_atm_D14C and misc_SLT series of the same two
 * routines (biogem_data_ascii.f90:107-197, 320-400; 955-1096, 1245-1340) from "bg_sig" (int_t_sig, the atmosphere rows) and
 * "bg_sig2".  sed_type: tracer_define.sed column 4 (1 ... 7 bulk, 8 age, 9 frac2: no file, 11 / 12 isotopes); ocn_tot_A =
 * SUM(phys_ocn(ipo_A,:,:,n_k)) (m2); opsi_scale = goldstein_dsc * goldstein_usc * const_rEarth * 1.0E-6; atlantic != 0 for the
 * topographies of :1282-1283 (worbe2, worjh2 ...).  Host only. */
int cg_biogem_series_write_ext(const char *outdir, const char *outfile_name, int create, double t_yr, int n_ocn, int n_sed,
                               const char *const *sed_names, const int32_t *sed_type, const int32_t *sed_dep, int n_atm,
                               const char *const *atm_names, const int32_t *atm_type, const int32_t *atm_dep, const double *sig,
                               const double *sig2, double ocn_tot_A, double opsi_scale, int atlantic);
/* (re)build BIOGEM's ocn array from the current ts (initialise_biogem, biogem.f90:283-285: T in K, S absolute) */
int cg_biogem_init_ocn(cg_handle *);
int cg_atchem_step(cg_handle *, double dts);

/* ---- whole coupling loop on the device -------------------------------- */
/* n iterations of genie.f90's koverall loop (:117-534, "normal" branch) with
 * no host involvement: surflux / EMBM / sea ice / ocean (/ BIOGEM / ATCHEM)
 * on the reference's MOD(koverall, k*_loop) schedule.                       */
int cg_run(cg_handle *, int64_t n_koverall);

/* Restart: continue the coupling loop from iteration `koverall` (multiple of kocn_loop) after the prognostic fields were
 * restored with cg_sync_from_host; sets koverall, istep_ocn/atm/sic, the GENIE clock and BIOGEM's derived counters.   */
int cg_set_koverall(cg_handle *, int64_t koverall);
/* After T, S of `member` were rewritten from the host (restart): rho = eos(T, S) at the wet cells, as initialise_goldstein
 * does behind inm_netcdf (goldstein.f90:1724-1760: ts1 = ts, rho from eos); the momentum step reads rho before the first
 * tracer step recomputes it.  member < 0 = all members.                                                                  */
int cg_refresh_rho(cg_handle *, int member);

/* ---- state movement (restart / output / coupling intervals) ----------- */
/* Named field of ONE member, in the reference's Fortran shape:
 *  "ts" (maxl,maxi,maxj,maxk)  "u" (3,maxi,maxj,maxk)  "rho" (maxi,maxj,maxk)
 *  "tq" (2,maxi,maxj)  "varice" (2,maxi,maxj)  "psi" (0:maxi,0:maxj)
 *  "cost","tice","usurf","pptn","evap","fx0a","fxlw"... (maxi,maxj), BIOGEM: "ocn","bio_part","atm",...
 * cg_field_size returns the number of doubles.                              */
int64_t cg_field_size(cg_handle *, const char *name);
int cg_sync_to_host(cg_handle *, const char *name, int member, double *dst, int64_t n);
int cg_sync_from_host(cg_handle *, const char *name, int member, const double *src, int64_t n);
/* All members at once, device-native layout [..][member]; pinned-host friendly. */
int cg_sync_all_to_host(cg_handle *, const char *name, double *dst, int64_t n);
int cg_sync_all_from_host(cg_handle *, const char *name, const double *src, int64_t n);
/* The same for 3-D ocean fields ("ts", "u", "rho", "ocn", "bio_part" ...) with the WET cells only, packed on the device:
 * layout [wet cell][inner][member], wet cells in ascending cell order (k, j, i), inner = the field's leading dimension
 * (tracers of ts, components of u; 1 for rho).  worjh2: 12 511 of 20 736 cells are wet, so a state exchange moves 60 % of the
 * bytes.  cg_wet_size = doubles per member (n_wet * inner); n = cg_wet_size * member_stride.  Dry cells are left untouched
 * by the upload.                                                                                                        */
int64_t cg_wet_size(cg_handle *, const char *name);
int cg_sync_all_wet_to_host(cg_handle *, const char *name, double *dst, int64_t n);
int cg_sync_all_wet_from_host(cg_handle *, const char *name, const double *src, int64_t n);
/* Double-buffered exchange of a resident ensemble's state at coupling / output intervals (genie.f90's model of "host copies only
 * at coupling and output intervals", without stalling the device on PCIe).  All members of one field in the device layout
 * (wet != 0: 3-D ocean fields as wet cells, cg_wet_size doubles per member; else cg_field_size).
 *   cg_exchange_begin_upload   async copy of the host buffer into a device staging buffer on a copy stream; returns at once.
 *   cg_exchange_commit_upload  the compute stream waits for that copy and unpacks it into the field (`also`: a second field that
 *                              takes the same data, e.g. "tq1" next to "tq"; NULL / "": none).
 *   cg_exchange_begin_download packs the field into a staging buffer behind the work issued so far, then the copy stream moves it
 *                              to the host buffer while the compute stream goes on.
 *   cg_exchange_wait           blocks until the copy stream is idle (host buffers of begun downloads are valid, those of begun
 *                              uploads reusable).
 * Host buffers must be page-locked (cudaHostAlloc / torch pin_memory) for the copies to overlap. */
int cg_exchange_begin_upload(cg_handle *, const char *name, int wet, const double *src, int64_t n);
int cg_exchange_commit_upload(cg_handle *, const char *name, const char *also);
int cg_exchange_begin_download(cg_handle *, const char *name, int wet, double *dst, int64_t n);
int cg_exchange_wait(cg_handle *);


/* Host-side constants as built by cg_initialise (bit-exactness checks):
 * "dz","dza","s","c","sv","cv","ds","dsv","rc","rc2","cv2","rds","rdsv","zro","zw","ssmax",
 * "drag","rh","gap","ratm","ubisl","psisl","erisl","solfor","diffa","albcl","pmeadj","uatm","ca",...
 * integer: "k1","ku","mk","getj","ips","ipf","ias","iaf","iroff","jroff".   */
int64_t cg_const_size(cg_handle *, const char *name);
int cg_get_const(cg_handle *, const char *name, int member, double *dst, int64_t n);
int cg_get_iconst(cg_handle *, const char *name, int32_t *dst, int64_t n);
int cg_get_dims(cg_handle *, int32_t dims[8]); /* maxi,maxj,maxk,maxl,n_members,member_stride,nyear,ndta */

/* ---- diagnostics (warp-shuffle reductions on the device) -------------- */
/* Per-member volume-weighted global means of every ts tracer: out[n_members*maxl] */
int cg_global_means(cg_handle *, double *out);
/* per-member flags: bit 0 = non-finite ts (blow-up), bit 1 = BIOGEM carbonate chemistry failed (error_stop); out[n_members] */
int cg_health(cg_handle *, int32_t *out);

/* ---- measurement helpers ---------------------------------------------- */
int cg_synchronize(cg_handle *);
/* kernels the library launched since the last reset (the bench's gpu_launches) */
int64_t cg_launch_count(cg_handle *, int reset);
/* CUDA-event timing on the library's own stream */
int cg_timer_start(cg_handle *);
int cg_timer_stop_ms(cg_handle *, double *ms);
/* per-kernel-family event timing: name in {"tstepo_flux","co","momentum","embm","surflux","seaice","biogem"} */
int cg_profile_enable(cg_handle *, int on);
int cg_profile_get(cg_handle *, const char *family, double *total_ms, int64_t *launches);
/* tracer kernel variant: 0 = strict reference operation order (bit-exact vs the oracle),
 *                        1 = fast (FMA + factored isoneutral sums; <=1e-10 relative per step),
 *                        2 = fused column kernel (tstepo_flux + co + SST export in one pass over ts; same tolerance;
 *                            compiled for the 36x36x16, 16-tracer shape, any other shape runs variant 1) */
int cg_set_tracer_variant(cg_handle *, int variant);
/* the variant that actually runs (2 is reported as 1 when the grid shape has no compiled column kernel); -1 = bad handle */
int cg_tracer_variant_active(cg_handle *);
/* cg_run only: apply biogem_tracercoupling's per-cell update inside the step_biogem kernel (bit-identical to the two
 * separate calls, which the per-module entry points always use).  Default OFF: measured slower on B200 -- the step
 * kernel is latency bound at 255 registers and the separate update streams at 3.4 TB/s. */
int cg_set_biogem_fusion(cg_handle *, int on);
/* use CUDA-graph replay of one ocean step inside cg_run (default on) */
int cg_set_graphs(cg_handle *, int on);

/* ---- stand-alone tracer step on caller-provided fields (kernel tests, stress config #5) ---- */
/* One tstepo (flux + convection) on a synthetic grid without the rest of the model:
 * k1 (0:maxi+1,0:maxj+1) int32; u (3,0:maxi,0:maxj,maxk); ts (maxl,0:maxi+1,0:maxj+1,0:maxk+1) per member
 * laid out member-slowest on the host.  Used for BASELINE config #5 and parity tests.       */
int cg_tracer_create(int maxi, int maxj, int maxk, int maxl, int n_members, int device, const int32_t *k1,
                     double diff1, double diff2, int nyear, cg_handle **out);
int cg_tracer_set(cg_handle *, const double *ts, const double *u, const double *tsflux);
int cg_tracer_step(cg_handle *, int nsteps);
int cg_tracer_get(cg_handle *, double *ts, double *rho, double *cost);

/* ---- netCDF restart files in the reference's layout (host only: plain arrays, no handle, no device) ----
 * Replace outm_netcdf / inm_netcdf (src/goldstein/goldstein_data.f90:153-300 / :11-150), outm_netcdf_embm / inm_netcdf_embm
 * (src/embm/embm_data.f90:83-200 / :11-80), outm_netcdf_sic / inm_netcdf_sic (src/goldsteinseaice/gold_seaice_data.f90:100-230 /
 * :11-98) for hosts without netCDF-Fortran (the Python engine; the Fortran model keeps writing its own restarts from the
 * arrays cg_sync_to_host fills).  Files are netCDF-3 classic with the reference's dimension / variable names, types and
 * definition order.  Arrays are column-major as in Fortran: ts (maxl,maxi,maxj,maxk), u (3,maxi,maxj,maxk), tq and varice
 * (2,maxi,maxj), k1 (0:maxi+1,0:maxj+1) int32; lon / lat / depth are the nclon1 / nclat1 / depths1 axes.
 * date = {iyear_rest, imonth_rest, iday, ioffset_rest}.  evap / late / sens may be NULL (written as zeros, not read).
 * Non-zero return: CG_ERR_ARG / CG_ERR_IO with the text in cg_restart_last_error(). */
const char *cg_restart_last_error(void);
int cg_restart_goldstein_write(const char *path, int maxi, int maxj, int maxk, int maxl, const int32_t *k1, const double *lon,
                               const double *lat, const double *depth, const double *ts, const double *u, const double *evap,
                               const double *late, const double *sens, const int32_t date[4]);
/* ts(1:2) and u(1:2) are replaced, everything else in ts / u is left as it is (goldstein_data.f90:88-96) */
int cg_restart_goldstein_read(const char *path, int maxi, int maxj, int maxk, int maxl, double *ts, double *u, double *evap,
                              double *late, double *sens, int32_t date[4]);
int cg_restart_embm_write(const char *path, int maxi, int maxj, const double *lon, const double *lat, const double *tq,
                          const int32_t date[4]);
int cg_restart_embm_read(const char *path, int maxi, int maxj, double *tq, int32_t date[4]);
int cg_restart_seaice_write(const char *path, int maxi, int maxj, const int32_t *k1, const double *lon, const double *lat,
                            const double *varice, const double *tice, const double *albice, const int32_t date[4]);
int cg_restart_seaice_read(const char *path, int maxi, int maxj, double *varice, double *tice, double *albice, int32_t date[4]);
/* The date block on its own: dimension nrecs = 1 and the INT variables ioffset, iyear, imonth, iday over it -- what every module's
 * restart starts with (goldstein_data.f90:247-251) and ALL that genie-main's own restart holds (fname_restart_main,
 * src/main-defaults.nml:38; the reference ships one, data/main/main_restart_0.nc, written by the netCDF library: the writer
 * reproduces that file byte for byte, tests/test_restart_nc.py).  date = {iyear, imonth, iday, ioffset}. */
int cg_restart_date_write(const char *path, const int32_t date[4]);
int cg_restart_date_read(const char *path, int32_t date[4]);
/* BIOGEM's restart (ctrl_ncrst = .TRUE.): sub_data_netCDF_ncrstsave (src/biogem/biogem_data_netCDF.f90:24-142) and the netCDF
 * branch of sub_data_load_rst (src/biogem/biogem_data.f90:438-568).  ocn (n_ocn,n_i,n_j,n_k), bio_part (n_sed,n_i,n_j,n_k);
 * names = string_ocn / string_sed of the selected tracers (tracer_define.ocn / .sed column 1), long names column 5.  The
 * file stores FLOAT variables "ocn_<name>", "bio_part_<name>" (zt, lat, lon), surface first, fill value on dry cells --
 * the reference's restart is single precision.  read: tracers without a variable keep their values (found_* = 0). */
int cg_restart_biogem_write(const char *path, int n_i, int n_j, int n_k, const int32_t *k1, const double *lon, const double *lat,
                            const double *lon_e, const double *lat_e, const double *zt, const double *zt_e, int n_ocn,
                            const char *const *ocn_names, const char *const *ocn_longnames, const double *ocn, int n_sed,
                            const char *const *sed_names, const char *const *sed_longnames, const double *bio_part,
                            double year, const char *run_id);
int cg_restart_biogem_read(const char *path, int n_i, int n_j, int n_k, const int32_t *k1, int n_ocn, const char *const *ocn_names,
                           double *ocn, int32_t *found_ocn, int n_sed, const char *const *sed_names, double *bio_part,
                           int32_t *found_sed);

/* ATCHEM's restart (ctrl_ncrst = .TRUE.): sub_data_netCDF_ncrstsave (src/atchem/atchem_data_netCDF.f90:22-109) and the netCDF
 * branch of sub_data_load_rst (src/atchem/atchem_data.f90:89-189).  atm (n_atm,n_i,n_j); names / long names = string_atm /
 * string_longname_atm of the selected tracers (tracer_define.atm columns 1 and 5).  FLOAT variables "atm_<name>" (lat, lon),
 * no mask.  lon / lat / lon_e (0:n_i) / lat_e (0:n_j): phys_atm's axes (atchem_data.f90:195-229) through edge_maker. */
int cg_restart_atchem_write(const char *path, int n_i, int n_j, const double *lon, const double *lat, const double *lon_e,
                            const double *lat_e, int n_atm, const char *const *atm_names, const char *const *atm_longnames,
                            const double *atm, double year, const char *run_id);
int cg_restart_atchem_read(const char *path, int n_i, int n_j, int n_atm, const char *const *atm_names, double *atm, int32_t *found);
/* Binary restarts (ctrl_ncrst = .FALSE.): one gfortran unformatted sequential record, INTEGER*4 and REAL*8
 * (-fdefault-real-8, platforms/LINUX:9), full double precision.  ATCHEM: atchem_save_rst (src/atchem/atchem.f90:186-198) /
 * sub_data_load_rst (atchem_data.f90:175-181); BIOGEM: biogem_save_restart (src/biogem/biogem.f90:2340-2358) /
 * sub_data_load_rst (biogem_data.f90:540-550).  *_ids: the tracers' global indices (conv_iselected_ia / _io / _is), which
 * the record carries and the reader matches on; tracers of the file the caller did not select are skipped. */
int cg_restart_atchem_write_bin(const char *path, int n_i, int n_j, int n_atm, const int32_t *atm_ids, const double *atm);
int cg_restart_atchem_read_bin(const char *path, int n_i, int n_j, int n_atm, const int32_t *atm_ids, double *atm, int32_t *found);
int cg_restart_biogem_write_bin(const char *path, int n_i, int n_j, int n_k, int n_ocn, const int32_t *ocn_ids, const double *ocn,
                                int n_sed, const int32_t *sed_ids, const double *bio_part);
int cg_restart_biogem_read_bin(const char *path, int n_i, int n_j, int n_k, int n_ocn, const int32_t *ocn_ids, double *ocn,
                               int32_t *found_ocn, int n_sed, const int32_t *sed_ids, double *bio_part, int32_t *found_sed);

/* BIOGEM's 3-D time-slice file fields_biogem_3d.nc: sub_init_netcdf (dd = 3) + sub_save_netcdf + sub_save_netcdf_3d
 * (src/biogem/biogem_data_netCDF.f90:148-277, 282-459, 1959-2315; helpers src/common/gem_netcdf.f90:19-76 sub_opennext, 250-359
 * sub_defvar, 702-744 sub_putvar3d_g, 790-840 sub_adddef_netcdf).  The first call creates `path` (dimensions time = unlimited, xu,
 * lon, lat, zt, yu, the *_edges, lat_moc, zt_moc and their edges, para; the axes; grid_level, grid_mask, grid_topo) and writes
 * record 1; a call on an existing file appends the next record (sub_opennext: ntrec = length of time + 1).  One record = time
 * (year mid-point), year = nint(time), and as FLOAT (zt, lat, lon) variables with the surface level first and the fill value on
 * dry cells: ocn_<name> = int_ocn / int_t (ctrl_data_save_slice_ocn; temperature minus 273.15; isotopes as delta values through
 * fun_calc_isotope_delta, gem_util.f90:568-598; valid_range from tracer_define.ocn), ocn_DIC_D14C when DIC_13C and DIC_14C are
 * both there (fun_convert_delta14CtoD14C, gem_util.f90:623-637), with `mass` (phys_ocn(ipo_M), (n_i,n_j,n_k); NULL leaves the
 * derived fields out = ctrl_data_save_derived off) ocn_<name>_Snorm and ocn_<name>_tot, carb_<name> (ctrl_data_save_slice_carb),
 * carb_const_<name> (ctrl_data_save_slice_carbconst) and, with `mass`, bio_part_<name>.  int_ocn (n_ocn,n_i,n_j,n_k), int_part
 * (n_sed,...), int_carb (n_carb,...), int_carbconst (n_carbconst,...) are the window integrals of ONE member in Fortran order
 * (the fields "sl_ocn", "sl_part", "sl_carb", "sl_carbconst" of cg_biogem_slice_update; the carbonate rows in the ORDER OF THE
 * NAMES passed -- the reference's is string_carb / string_carbconst, gem_cmn.f90:425-455), int_t = int_t_timeslice.  *_type: 0,
 * 1, 11 (13C), 12 (14C) ...; *_dep: 0-based compact index of an isotope's bulk tracer; ocn_mima (2,n_ocn).  n_sed / n_carb /
 * n_carbconst = 0 leave a block out.  Axes as for cg_restart_biogem_write.  Not written: the remin / phys_ocn / settling-flux /
 * diag_geochem / velocity blocks (:2108-2135, 2200-2314).  Host only; errors through cg_restart_last_error(). */
int cg_slice_biogem_write_3d(const char *path, int n_i, int n_j, int n_k, const int32_t *k1, const double *lon, const double *lat,
                             const double *lon_e, const double *lat_e, const double *zt, const double *zt_e, int n_ocn,
                             const char *const *ocn_names, const char *const *ocn_longnames, const char *const *ocn_units,
                             const double *ocn_mima, const int32_t *ocn_type, const int32_t *ocn_dep, const double *int_ocn,
                             int n_sed, const char *const *sed_names, const int32_t *sed_type, const int32_t *sed_dep,
                             const double *int_part, int n_carb, const char *const *carb_names, const double *int_carb,
                             int n_carbconst, const char *const *carbconst_names, const double *int_carbconst, const double *mass,
                             double int_t, double year_mid, const char *run_id);

#ifdef __cplusplus
}
#endif
#endif /* CGENIE_B200_H */
