/* genie_driver.c -- a plain-C host that drives libcgenie_b200.so through include/cgenie_b200.h in the exact call order of
 * the reference's coupler (src/genie.f90:117-534, "normal" branch), with Fortran-shaped (column-major) host arrays passed
 * on "output" steps, exactly as the ISO_C_BINDING shims of fortran/ do.  TEST INFRASTRUCTURE (tests/test_gpu_c_driver.py):
 * no Python, no ctypes, no numpy between the host program and the library -- what a linked genie.exe would see.
 *
 *   genie_driver <jobdir> <n_koverall> <out_every_ocean_steps> <outfile> [n_members]
 *
 * Every `out_every` ocean steps the driver passes host arrays to surflux / step_embm / step_seaice / step_goldstein (the
 * shims do this when MOD(istep, npstp|iwstp|itstp|ianav) == 0); all other calls pass NULL = "stay resident".  It writes, as raw
 * doubles in this order: header {maxl, maxi, maxj, maxk, n_out}, then per output step
 * tstar_ocn, sstar_ocn (maxi*maxj each), tstar_atm, qstar_atm, hght_sic, frac_sic, latent_ocn, go_rho (maxi*maxj*maxk),
 * test_energy_ocean, test_water_ocean; at the end go_ts (maxl*maxi*maxj*maxk) through cg_biogem_tracercoupling's INOUT
 * arrays when BIOGEM is on, else through cg_goldstein_step's last output step; then "ocn" and "atm" if BIOGEM is on. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/cgenie_b200.h"

#define CK(call)                                                                         \
  do {                                                                                   \
    int rc_ = (call);                                                                    \
    if (rc_) { fprintf(stderr, "genie_driver: %s -> %d: %s\n", #call, rc_, cg_last_error()); return 10 + rc_; } \
  } while (0)

extern int cg_get_dims(cg_handle *, int32_t dims[8]);
extern int64_t cg_field_size(cg_handle *, const char *);
extern int cg_sync_to_host(cg_handle *, const char *, int, double *, int64_t);
extern int cg_set_tracer_variant(cg_handle *, int);

static void put(FILE *f, const double *p, size_t n) { fwrite(p, sizeof(double), n, f); }

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: genie_driver <jobdir> <n_koverall> <out_every> <outfile> [n_members]\n"); return 2; }
  const char *jobdir = argv[1];
  const long nk = atol(argv[2]);
  const int out_every = atoi(argv[3]);
  FILE *f = fopen(argv[4], "wb");
  const int members = argc > 5 ? atoi(argv[5]) : 1;
  if (!f) return 3;
  cg_handle *h = NULL;
  CK(cg_create(jobdir, members, 0, &h));
  CK(cg_initialise(h));
  int32_t d[8];
  CK(cg_get_dims(h, d));
  const int I = d[0], J = d[1], K = d[2], L = d[3], nyear = d[6];
  const size_t ij = (size_t)I * J, ijk = ij * K;
  const int biogem = cg_field_size(h, "ocn") > 0;
  /* the timestepping the job tool writes (tools/config_utils.py:103-162): 5 atmosphere steps per ocean step */
  const int kocn = 5, katm = 1, ksic = 5, kbiogem = 2, katchem = 2;
  const double genie_timestep = 3600.0 * 24.0 * 365.25 / 5.0 / nyear;
  const long long tick = (long long)floor(1000.0 * genie_timestep + 0.5);     /* NINT, genie_global.f90:401-410 */
  const double dts_bg = (double)(kbiogem * kocn) * genie_timestep, dts_ac = (double)(katchem * kocn) * genie_timestep;

  double *a2[24];
  for (int q = 0; q < 24; q++) a2[q] = (double *)calloc(ij, sizeof(double));
  double *go_ts = (double *)calloc(ijk * L, sizeof(double)), *go_ts1 = (double *)calloc(ijk * L, sizeof(double));
  double *go_u = (double *)calloc(ijk * 3, sizeof(double)), *go_rho = (double *)calloc(ijk, sizeof(double));
  double *go_psi = (double *)calloc((size_t)(I + 1) * (J + 1), sizeof(double));
  double te = 0.0, tw = 0.0;
  int n_out = 0;
  const double hdr[5] = {(double)L, (double)I, (double)J, (double)K, 0.0};
  put(f, hdr, 5);

  int istep_ocn = 0, istep_atm = 0, istep_sic = 0;
  long long genie_clock = 0;
  for (long k = 1; k <= nk; k++) {
    genie_clock += tick;                                                        /* increment_genie_clock */
    const int out = out_every > 0 && (k % kocn == 0 ? ((k / kocn) % out_every == 0) : (((k + kocn - 1) / kocn) % out_every == 0));
    if (k % kocn == 1) {                                                        /* genie.f90:271-277 */
      istep_ocn++;
      if (out) {
        cg_surflux_io io;
        memset(&io, 0, sizeof io);
        io.latent_ocn = a2[6]; io.albedo_ocn = a2[7]; io.evap_ocn = a2[8];
        CK(cg_surflux_step(h, istep_ocn, &io));
      } else CK(cg_surflux_step(h, istep_ocn, NULL));
    }
    if (k % katm == 0) {                                                        /* :287-291 */
      istep_atm++;
      if (out && k % kocn == 0) {
        cg_embm_io io = {a2[2], a2[3]};
        CK(cg_embm_step(h, istep_atm, &io));
      } else CK(cg_embm_step(h, istep_atm, NULL));
    }
    if (k % ksic == 0) {                                                        /* :297-301 */
      istep_sic++;
      if (out) {
        cg_seaice_io io = {a2[4], a2[5], a2[9], a2[10]};
        CK(cg_seaice_step(h, istep_sic, &io));
      } else CK(cg_seaice_step(h, istep_sic, NULL));
    }
    if (k % kocn == 0) {                                                        /* :307-311 */
      if (out) {
        cg_goldstein_io io;
        memset(&io, 0, sizeof io);
        io.tstar_ocn = a2[0]; io.sstar_ocn = a2[1]; io.ustar_ocn = a2[11]; io.vstar_ocn = a2[12]; io.albedo_ocn = a2[13];
        io.go_u = go_u; io.go_rho = go_rho; io.go_psi = go_psi; io.test_energy_ocean = &te; io.test_water_ocean = &tw;
        CK(cg_goldstein_step(h, istep_ocn, &io));
        put(f, a2[0], ij); put(f, a2[1], ij); put(f, a2[2], ij); put(f, a2[3], ij); put(f, a2[4], ij); put(f, a2[5], ij);
        put(f, a2[6], ij); put(f, go_rho, ijk); put(f, &te, 1); put(f, &tw, 1);
        n_out++;
      } else CK(cg_goldstein_step(h, istep_ocn, NULL));
    }
    if (biogem && k % (kbiogem * kocn) == 0) {                                  /* :360-443 */
      if (k == kbiogem * kocn) CK(cg_biogem_climate_sol(h));
      CK(cg_biogem_forcing(h, genie_clock));
      CK(cg_biogem_step(h, dts_bg, genie_clock));
      if (k + kbiogem * kocn > nk) {   /* last block: go_ts / go_ts1 INOUT through the coupling, as the Fortran wrapper passes them */
        CK(cg_sync_to_host(h, "ts", 0, go_ts, (int64_t)(ijk * L)));
        CK(cg_biogem_tracercoupling(h, go_ts, go_ts1));
      } else CK(cg_biogem_tracercoupling(h, NULL, NULL));
      CK(cg_biogem_climate(h));
      CK(cg_cpl_flux_ocnatm(h));
      CK(cg_cpl_flux_ocnsed(h, dts_bg));
      CK(cg_cpl_comp_ocnsed(h, (int)(k / kocn), kbiogem, 2 * kbiogem));
      CK(cg_reinit_flux_rokocn(h));
    }
    if (biogem && k % (katchem * kocn) == 0) CK(cg_atchem_step(h, dts_ac));     /* :446-455 */
  }
  if (!biogem) CK(cg_sync_to_host(h, "ts", 0, go_ts, (int64_t)(ijk * L)));
  put(f, go_ts, ijk * L);
  if (biogem) {
    const int64_t no = cg_field_size(h, "ocn"), na = cg_field_size(h, "atm");
    double *ocn = (double *)malloc((size_t)no * 8), *atm = (double *)malloc((size_t)na * 8);
    CK(cg_sync_to_host(h, "ocn", 0, ocn, no));
    CK(cg_sync_to_host(h, "atm", 0, atm, na));
    put(f, go_ts1, ijk * L);
    put(f, ocn, (size_t)no);
    put(f, atm, (size_t)na);
  }
  const double nout = (double)n_out;
  fseek(f, 4 * sizeof(double), SEEK_SET);
  put(f, &nout, 1);
  fclose(f);
  CK(cg_destroy(h));
  printf("genie_driver: %ld koverall iterations, %d output steps, %s\n", nk, n_out, biogem ? "BIOGEM on" : "physics only");
  return 0;
}
