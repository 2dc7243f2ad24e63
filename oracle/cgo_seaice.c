/* cgo_seaice.c -- CPU oracle, GOLDSTEIN sea ice.  TEST INFRASTRUCTURE ONLY.
 * Restates src/goldsteinseaice/gold_seaice.f90 (explicit scheme, impsic=.FALSE.).
 * Parity unpinned (see cgo.h). */
#include "cgo_impl.h"

#define VARICE(l, i, j) A3(o->varice, l, i, j)
#define VARICE1(l, i, j) A3(o->varice1, l, i, j)
#define DTHA(l, i, j) A3(o->dtha, l, i, j)
#define SU(l, i, j) o->sic_u[((l)-1) + 2 * ((i) + (NI + 1) * (j))] /* u(2,0:maxi,0:maxj) */

/* gold_seaice.f90:17-508 */
void cgo_seaice_init(cgo_t *o) {
  double tv = 86400.0 * o->yearlen / (o->nyear * CG_TSC);
  o->dtsic = tv;
  o->sic_rdtdim = 1.0 / (CG_TSC * o->dtsic);
  o->diffsic = o->diffsic_in / (CG_RSC * CG_USC);
  /* varice, varice1, tice, albice = 0 (calloc); outputs hght/frac/temp/albd = 0 */
}

/* gold_seaice.f90:844-929 */
static void tstepsic(cgo_t *o) {
  int i, j, l;
  double fe[3], fw[3], fn[3], fwsave[3];
  double *fs = (double *)calloc((size_t)3 * (NI + 1), 8);
#define FS(l, i) fs[(l) + 3 * (i)]
  for (j = 1; j <= NJ; j++) {
    for (l = 1; l <= 2; l++) {
      if (NK >= imax2(K1(NI, j), K1(1, j))) {
        fw[l] = SU(1, NI, j) * o->rc[j] * (VARICE1(l, 1, j) + VARICE1(l, NI, j)) * 0.5;
        if (SU(1, NI, j) >= 0.0) {
          if (VARICE1(2, 1, j) > o->par_sica_thresh) fw[l] = 0;
          if (VARICE1(1, 1, j) > o->par_sich_thresh) fw[l] = 0;
        } else {
          if (VARICE1(2, NI, j) > o->par_sica_thresh) fw[l] = 0;
          if (VARICE1(1, NI, j) > o->par_sich_thresh) fw[l] = 0;
        }
        fw[l] = fw[l] - (VARICE1(l, 1, j) - VARICE1(l, NI, j)) * o->rc[j] * o->rc[j] * o->rdphi * o->diffsic;
      } else {
        fw[l] = 0;
      }
      fwsave[l] = fw[l];
    }
    for (i = 1; i <= NI; i++)
      for (l = 1; l <= 2; l++) {
        if (i == NI) {
          fe[l] = fwsave[l];
        } else if (NK < imax2(K1(i, j), K1(i + 1, j))) {
          fe[l] = 0;
        } else {
          fe[l] = SU(1, i, j) * o->rc[j] * (VARICE1(l, i + 1, j) + VARICE1(l, i, j)) * 0.5;
          if (SU(1, i, j) >= 0.0) {
            if (VARICE1(2, i + 1, j) > o->par_sica_thresh) fe[l] = 0;
            if (VARICE1(1, i + 1, j) > o->par_sich_thresh) fe[l] = 0;
          } else {
            if (VARICE1(2, i, j) > o->par_sica_thresh) fe[l] = 0;
            if (VARICE1(1, i, j) > o->par_sich_thresh) fe[l] = 0;
          }
          fe[l] = fe[l] - (VARICE1(l, i + 1, j) - VARICE1(l, i, j)) * o->rc[j] * o->rc[j] * o->rdphi * o->diffsic;
        }
        if (NK < imax2(K1(i, j), K1(i, j + 1))) {
          fn[l] = 0;
        } else {
          fn[l] = o->cv[j] * SU(2, i, j) * (VARICE1(l, i, j + 1) + VARICE1(l, i, j)) * 0.5;
          if (SU(2, i, j) >= 0.0) {
            if (VARICE1(2, i, j + 1) > o->par_sica_thresh) fn[l] = 0;
            if (VARICE1(1, i, j + 1) > o->par_sich_thresh) fn[l] = 0;
          } else {
            if (VARICE1(2, i, j) > o->par_sica_thresh) fn[l] = 0;
            if (VARICE1(1, i, j) > o->par_sich_thresh) fn[l] = 0;
          }
          fn[l] = fn[l] - o->cv[j] * o->cv[j] * (VARICE1(l, i, j + 1) - VARICE1(l, i, j)) * o->rdsv[j] * o->diffsic;
        }
        if (NK >= K1(i, j))
          VARICE(l, i, j) = VARICE1(l, i, j) -
                            o->dtsic * ((fe[l] - fw[l]) * o->rdphi + (fn[l] - FS(l, i)) * o->rds[j]) +
                            CG_TSC * o->dtsic * DTHA(l, i, j);
        fw[l] = fe[l];
        FS(l, i) = fn[l];
      }
  }
  free(fs);
#undef FS
}

/* gold_seaice.f90:511-733 (diagnostic/file output dropped) */
void cgo_seaice_step(cgo_t *o) {
  int i, j;
  double fw_delta, fx_delta;
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      DTHA(1, i, j) = A2(o->dhght_sic, i, j);
      DTHA(2, i, j) = A2(o->dfrac_sic, i, j);
      SU(1, i, j) = A2(o->ustar_ocn, i, j);
      SU(2, i, j) = A2(o->vstar_ocn, i, j);
    }
  tstepsic(o);
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      fw_delta = 0.0;
      fx_delta = 0.0;
      if (NK >= K1(i, j)) {
        fw_delta = -CG_RHOIO * DTHA(1, i, j);
        VARICE(2, i, j) = dmax2(0.0, dmin2(1.0, VARICE(2, i, j)));
        if (VARICE(1, i, j) < CG_HMIN) {
          fx_delta = -VARICE(1, i, j) * CG_RHOICE * CG_HLF * o->sic_rdtdim;
          fw_delta = fw_delta + VARICE(1, i, j) * CG_RHOIO * o->sic_rdtdim;
          VARICE(1, i, j) = 0.0;
          VARICE(2, i, j) = 0.0;
        }
        VARICE1(1, i, j) = VARICE(1, i, j);
        VARICE1(2, i, j) = VARICE(2, i, j);
        fw_delta = fw_delta * CG_M2MM;
      }
      A2(o->hght_sic, i, j) = VARICE(1, i, j);
      A2(o->frac_sic, i, j) = VARICE(2, i, j);
      A2(o->waterflux_ocn, i, j) = fw_delta;
      A2(o->conductflux_ocn, i, j) = fx_delta;
    }
}
