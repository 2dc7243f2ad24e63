/* cgo_biogem.c -- CPU oracle, BIOGEM pieces on the tracer hot path.  TEST INFRASTRUCTURE ONLY.
 * Restates src/biogem/biogem.f90: biogem_tracercoupling (:1885-2077), the cell geometry of
 * sub_init_phys_ocn (biogem_data.f90:1098-1137) and the ts<->ocn unit offsets (biogem.f90:283-285).
 * The source terms of step_biogem (vdocn) are an input here (zero unless a test sets them).
 * Parity unpinned (see cgo.h). */
#include "cgo_impl.h"

#define BG_PI 3.141592653589793   /* gem_cmn.f90:688 */
#define BG_REARTH 6.37e6          /* gem_cmn.f90:696 */
#define BG_M3_KG 1027.649         /* gem_cmn.f90:510 */
#define BG_ZEROC 273.15           /* gem_cmn.f90:690 */
#define BG_NULLSMALL 0.999999e-19 /* gem_cmn.f90:719 */

#define OCN(l, i, j, k) o->bg_ocn[((l)-1) + NL * (((i)-1) + NI * (((j)-1) + NJ * ((k)-1)))]
#define DOCN(l, i, j, k) o->bg_vdocn[((l)-1) + NL * (((i)-1) + NI * (((j)-1) + NJ * ((k)-1)))]
#define PHM(i, j, k) o->bg_M[((i)-1) + NI * (((j)-1) + NJ * ((k)-1))]
#define PHRM(i, j, k) o->bg_rM[((i)-1) + NI * (((j)-1) + NJ * ((k)-1))]
#define PHV(i, j, k) o->bg_V[((i)-1) + NI * (((j)-1) + NJ * ((k)-1))]

/* biogem_data.f90:1098-1137 and initialise_biogem's ts -> ocn copy */
void cgo_biogem_init(cgo_t *o) {
  int i, j, k, l;
  const long n3 = (long)NI * NJ * NK;
  o->bg_ocn = cgo_alloc(o, "ocn", n3 * NL);
  o->bg_vdocn = cgo_alloc(o, "vdocn", n3 * NL);
  o->bg_M = cgo_alloc(o, "bg_M", n3);
  o->bg_rM = cgo_alloc(o, "bg_rM", n3);
  o->bg_V = cgo_alloc(o, "bg_V", n3);
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

/* biogem.f90:1885-2077 (no particulate tracers: vbio_part lines :2042-2043 dropped) */
void cgo_biogem_tracercoupling(cgo_t *o) {
  const int L = NL;
  int i, j, k, l, n, nv = 0;
  double tot_V, rtot_V, mean_S_OLD, rmean_S_OLD, mean_S_NEW, Sratio, rSratio, s;
  double off[3] = {0.0, BG_ZEROC, o->saln0};
  double *tot_OLD = (double *)calloc(L + 1, 8), *tot_NEW = (double *)calloc(L + 1, 8), *rtot_NEW = (double *)calloc(L + 1, 8);
  int *ci = (int *)calloc((size_t)NI * NJ, sizeof(int)), *cj = (int *)calloc((size_t)NI * NJ, sizeof(int));
  double *partial = (double *)calloc((size_t)NI * NJ, 8);
  double *loc = (double *)calloc((size_t)(L + 1) * (NK + 1), 8); /* loc_vocn of one column */
  if (!o->bg_ocn) cgo_biogem_init(o);
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
      PHM(i, j, k) = rSratio * PHM(i, j, k);
      PHRM(i, j, k) = Sratio * PHRM(i, j, k);
      for (l = 1; l <= L; l++) TS1(l, i, j, k) = TS(l, i, j, k);
    }
  }
  free(tot_OLD); free(tot_NEW); free(rtot_NEW); free(ci); free(cj); free(partial); free(loc);
}
