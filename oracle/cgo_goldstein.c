/* cgo_goldstein.c -- CPU oracle, GOLDSTEIN ocean.  TEST INFRASTRUCTURE ONLY.
 * Restates src/goldstein/goldstein.f90 and goldstein_lib.f90 of the reference
 * (line numbers cited per function).  Parity unpinned (see cgo.h). */
#include "cgo_impl.h"

#define ERISL(a, b) o->erisl[((a)-1) + o->isles * ((b)-1)]
#define PSISL(i, j, n) o->psisl[(i) + (NI + 1) * ((j) + (NJ + 1) * ((n)-1))]
#define UBISL(l, i, j, n) o->ubisl[((l)-1) + 2 * ((i) + (NI + 2) * ((j) + (NJ + 1) * ((n)-1)))]
#define LPISL(p, n) o->lpisl[((p)-1) + o->mpi * ((n)-1)]
#define IPISL(p, n) o->ipisl[((p)-1) + o->mpi * ((n)-1)]
#define JPISL(p, n) o->jpisl[((p)-1) + o->mpi * ((n)-1)]

#define DZG(k, kk) o->dzg[(k) + (NK + 1) * (kk)]
#define Z2DZG(k, kk) o->z2dzg[(k) + (NK + 1) * (kk)]
#define RDZG(k, kk) o->rdzg[(k) + (NK + 1) * (kk)]

/* goldstein.f90:3048-3061: ieos = 0, and ieos = 1 with the thermobaricity term ec(5) * t * z */
void cgo_eos(const cgo_t *o, double t, double s, double z, double *rho) {
  if (o->ieos == 0) *rho = o->ec[1] * t + o->ec[2] * s + o->ec[3] * (t * t) + o->ec[4] * (t * t * t);
  else *rho = o->ec[1] * t + o->ec[2] * s + o->ec[3] * (t * t) + o->ec[4] * (t * t * t) + o->ec[5] * t * z;
}

/* goldstein.f90:3064-3082 */
static void eosd(const cgo_t *o, double t1, double t2, double s1, double s2, double z,
                 double rdz, double *dzrho, double *tec) {
  double tatw = 0.5 * (t1 + t2);
  if (o->ieos == 0) *tec = -o->ec[1] - o->ec[3] * tatw * 2 - o->ec[4] * tatw * tatw * 3;
  else *tec = -o->ec[1] - o->ec[3] * tatw * 2 - o->ec[4] * tatw * tatw * 3 - o->ec[5] * z;
  *dzrho = (o->ec[2] * (s2 - s1) - *tec * (t2 - t1)) * rdz;
}

/* goldstein.f90:2845-2885 */
static void drgset(cgo_t *o, double adrag, double drgf, int kmxdrg, int jeb) {
  int i, j, i1, i1p, j1, kloc2, kloc4;
  double *tmpdrg = (double *)calloc((size_t)(NI + 1) * (NJ + 1), sizeof(double));
#define TMPDRG(i, j) tmpdrg[(i) + (NI + 1) * (j)]
  for (j = 0; j <= NJ; j++)
    for (i = 0; i <= NI; i++) {
      kloc2 = imax2(imax2(K1(i, j), K1(i + 1, j)), imax2(K1(i, j + 1), K1(i + 1, j + 1)));
      kloc4 = K1(i, j);
      for (j1 = imax2(0, j - 1); j1 <= imin2(NJ + 1, j + 2); j1++)
        for (i1 = i - 1; i1 <= i + 2; i1++) {
          i1p = 1 + (NI + i1 - 1) % NI;
          kloc4 = imax2(kloc4, K1(i1p, j1));
        }
      if (kloc2 > kmxdrg || abs(j - NJ / 2) <= jeb)
        TMPDRG(i, j) = adrag * drgf * drgf;
      else if (kloc4 > kmxdrg || abs(j - NJ / 2) == jeb + 1)
        TMPDRG(i, j) = adrag * drgf;
      else
        TMPDRG(i, j) = adrag;
    }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      DRAG(1, i, j) = 0.5 * (TMPDRG(i, j) + TMPDRG(i, j - 1));
      DRAG(2, i, j) = 0.5 * (TMPDRG(i, j) + TMPDRG(i - 1, j));
    }
  for (j = 1; j <= NJ; j++) DRAG(2, NI + 1, j) = DRAG(2, 1, j);
#undef TMPDRG
  free(tmpdrg);
}

/* goldstein.f90:3138-3204 */
static void invert(cgo_t *o) {
  int i, j, k, l, n = NI, m = NJ + 1, im;
  double tv, tv1, rat;
  const double dphi = o->dphi, rdphi = o->rdphi;
  memset(o->gap, 0, sizeof(double) * (size_t)o->nm * (2 * n + 3));
  for (i = 1; i <= NI; i++)
    for (j = 0; j <= NJ; j++) {
      k = i + j * n;
      if (imax2(imax2(K1(i, j), K1(i + 1, j)), imax2(K1(i, j + 1), K1(i + 1, j + 1))) <= NK) {
        tv = (o->s[j + 1] * RH(1, i, j + 1) - o->s[j] * RH(1, i, j)) / (2.0 * o->dsv[j] * dphi);
        tv1 = (o->sv[j] * RH(2, i + 1, j) - o->sv[j] * RH(2, i, j)) / (2.0 * dphi * o->dsv[j]);
        GAP(k, 2) = DRAG(1, i, j) * o->c[j] * o->c[j] * RH(1, i, j) / (o->ds[j] * o->dsv[j]) + tv1;
        l = n + 1;
        if (i == 1) l = 2 * n + 1;
        GAP(k, l) = DRAG(2, i, j) * o->rcv[j] * o->rcv[j] * rdphi * rdphi * RH(2, i, j) - tv;
        GAP(k, n + 2) = -(DRAG(2, i, j) * RH(2, i, j) + DRAG(2, i + 1, j) * RH(2, i + 1, j)) /
                            (o->cv[j] * o->cv[j] * dphi * dphi) -
                        (DRAG(1, i, j) * o->c[j] * o->c[j] * RH(1, i, j) / o->ds[j] +
                         DRAG(1, i, j + 1) * o->c[j + 1] * o->c[j + 1] * RH(1, i, j + 1) / o->ds[j + 1]) /
                            o->dsv[j];
        l = n + 3;
        if (i == NI) l = 3;
        GAP(k, l) = DRAG(2, i + 1, j) * RH(2, i + 1, j) / (o->cv[j] * o->cv[j] * dphi * dphi) + tv;
        GAP(k, 2 * n + 2) =
            DRAG(1, i, j + 1) * o->c[j + 1] * o->c[j + 1] * RH(1, i, j + 1) / (o->ds[j + 1] * o->dsv[j]) - tv1;
      } else {
        GAP(k, n + 2) = 1;
      }
    }
  for (i = 1; i <= n * m - 1; i++) {
    im = imin2(i + n + 1, n * m);
    for (j = i + 1; j <= im; j++) {
      rat = GAP(j, n + 2 - j + i) / GAP(i, n + 2);
      RATM(j, j - i) = rat;
      if (rat != 0)
        for (k = n + 2 - j + i; k <= 2 * n + 3 - j + i; k++) GAP(j, k) = GAP(j, k) - rat * GAP(i, k + j - i);
    }
  }
}

/* goldstein.f90:3500-3565.  ubloc(2,0:maxi+1,0:maxj), psiloc(0:maxi,0:maxj) */
static void ubarsolv(cgo_t *o, double *ubloc, double *psiloc) {
  int i, j, k, n = NI, m = NJ + 1, km, im;
  double *gb = o->gb; /* 1-based */
#define UBL(l, i, j) ubloc[((l)-1) + 2 * ((i) + (NI + 2) * (j))]
#define PSL(i, j) psiloc[(i) + (NI + 1) * (j)]
  for (i = 1; i <= n * m - 1; i++) {
    im = imin2(i + n + 1, n * m);
    for (j = i + 1; j <= im; j++) gb[j] = gb[j] - RATM(j, j - i) * gb[i];
  }
  gb[n * m] = gb[n * m] / GAP(n * m, n + 2);
  for (i = n * m - 1; i >= 1; i--) {
    km = imin2(n + 1, n * m - i);
    for (k = 1; k <= km; k++) gb[i] = gb[i] - GAP(i, n + 2 + k) * gb[i + k];
    gb[i] = gb[i] / GAP(i, n + 2);
  }
  for (j = 0; j <= NJ; j++) {
    for (i = 1; i <= NI; i++) PSL(i, j) = gb[i + j * n];
    PSL(0, j) = PSL(NI, j);
  }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++)
      UBL(1, i, j) = -RH(1, i, j) * o->c[j] * (PSL(i, j) - PSL(i, j - 1)) * o->rds[j];
  for (j = 1; j <= NJ - 1; j++)
    for (i = 1; i <= NI; i++)
      UBL(2, i, j) = RH(2, i, j) * (PSL(i, j) - PSL(i - 1, j)) * o->rcv[j] * o->rdphi;
  for (i = 1; i <= NI; i++) {
    UBL(2, i, NJ) = 0.0;
    UBL(2, i, 0) = 0.0;
  }
  for (j = 1; j <= NJ; j++) {
    UBL(2, NI + 1, j) = UBL(2, 1, j);
    UBL(1, 0, j) = UBL(1, NI, j);
    UBL(1, NI + 1, j) = UBL(1, 1, j);
    UBL(2, 0, j) = UBL(2, NI, j);
  }
  UBL(2, NI + 1, 0) = UBL(2, 1, 0);
  UBL(2, 0, 0) = UBL(2, NI, 0);
#undef PSL
}

/* goldstein_lib.f90:186-241 */
static void island(cgo_t *o, const double *ubloc, double *erisl1, int isl, int indj) {
  int i, k, lpi, ipi, jpi, al;
  double cor, tv1, tv2, e = 0.0;
  for (i = 1; i <= o->npi[isl]; i++) {
    lpi = LPISL(i, isl);
    ipi = IPISL(i, isl);
    jpi = JPISL(i, isl);
    al = abs(lpi);
    if (al == 1)
      cor = -o->s[jpi] * 0.25 *
            (UBL(2, ipi, jpi) + UBL(2, ipi + 1, jpi) + UBL(2, ipi, jpi - 1) + UBL(2, ipi + 1, jpi - 1));
    else
      cor = o->sv[jpi] * 0.25 *
            (UBL(1, ipi - 1, jpi) + UBL(1, ipi, jpi) + UBL(1, ipi - 1, jpi + 1) + UBL(1, ipi, jpi + 1));
    e = e + isign1(lpi) *
                (DRAG(al, ipi, jpi) * UBL(al, ipi, jpi) + cor - indj * TAU(al, ipi, jpi) * RH(al, ipi, jpi)) *
                (o->c[jpi] * o->dphi * (2.0 - al) + o->rcv[jpi] * o->dsv[jpi] * (al - 1.0));
    if (indj == 1) {
      if (al == 1) {
        tv1 = 0.0;
        for (k = KU(1, ipi, jpi); k <= MK(ipi + 1, jpi); k++) tv1 = tv1 + BP(ipi + 1, jpi, k) * o->dz[k];
        for (k = KU(1, ipi, jpi); k <= MK(ipi, jpi); k++) tv1 = tv1 - BP(ipi, jpi, k) * o->dz[k];
        e = e + (SBP(ipi + 1, jpi) - SBP(ipi, jpi) + tv1) * isign1(lpi) * RH(1, ipi, jpi);
      } else {
        tv2 = 0.0;
        for (k = KU(2, ipi, jpi); k <= MK(ipi, jpi + 1); k++) tv2 = tv2 + BP(ipi, jpi + 1, k) * o->dz[k];
        for (k = KU(2, ipi, jpi); k <= MK(ipi, jpi); k++) tv2 = tv2 - BP(ipi, jpi, k) * o->dz[k];
        e = e + (SBP(ipi, jpi + 1) - SBP(ipi, jpi) + tv2) * isign1(lpi) * RH(2, ipi, jpi);
      }
    }
  }
  *erisl1 = e;
}
#undef UBL

/* goldstein.f90:3452-3467 */
static void matinv_gold(cgo_t *o) {
  int i, j, k, nvar = o->isles;
  for (i = 1; i <= nvar - 1; i++)
    for (j = i + 1; j <= nvar; j++)
      for (k = i + 1; k <= nvar; k++) ERISL(j, k) = ERISL(i, i) * ERISL(j, k) - ERISL(j, i) * ERISL(i, k);
}

/* goldstein.f90:3470-3492; rhs = erisl(:,isles+1) */
static void matmult(cgo_t *o) {
  int i, j, nvar = o->isles, r = o->isles + 1;
  for (i = 1; i <= nvar - 1; i++)
    for (j = i + 1; j <= nvar; j++) ERISL(j, r) = ERISL(i, i) * ERISL(j, r) - ERISL(j, i) * ERISL(i, r);
  ERISL(nvar, r) = ERISL(nvar, r) / ERISL(nvar, nvar);
  for (i = nvar - 1; i >= 1; i--) {
    for (j = i + 1; j <= nvar; j++) ERISL(i, r) = ERISL(i, r) - ERISL(i, j) * ERISL(j, r);
    ERISL(i, r) = ERISL(i, r) / ERISL(i, i);
  }
}

/* goldstein.f90:3685-3714 */
static void wind(cgo_t *o) {
  int i, j, k, n = NI, ip1;
  for (i = 1; i <= NI; i++)
    for (j = 0; j <= NJ; j++) {
      ip1 = i % NI + 1;
      k = i + j * n;
      if (imax2(imax2(K1(i, j), K1(i + 1, j)), imax2(K1(i, j + 1), K1(i + 1, j + 1))) <= NK) {
        o->gb[k] = (TAU(2, ip1, j) * RH(2, i + 1, j) - TAU(2, i, j) * RH(2, i, j)) * o->rdphi * o->rcv[j] -
                   (TAU(1, i, j + 1) * o->c[j + 1] * RH(1, i, j + 1) - TAU(1, i, j) * o->c[j] * RH(1, i, j)) *
                       o->rdsv[j];
      } else {
        o->gb[k] = 0;
      }
      o->gbold[k] = o->gb[k];
    }
}

/* goldstein.f90:3217-3315 */
static void jbar(cgo_t *o) {
  int i, j, k, l, n = NI, ip1;
  double tv1, tv2, tv3, tv4;
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++)
      if (K1(i, j) <= NK)
        for (k = K1(i, j) + 1; k <= NK; k++)
          BP(i, j, k) = BP(i, j, k - 1) - (RHO(i, j, k) + RHO(i, j, k - 1)) * o->dza[k - 1] * 0.5;
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++)
      if (MK(i, j) > 0) {
        SBP(i, j) = 0;
        for (k = MK(i, j) + 1; k <= NK; k++) SBP(i, j) = SBP(i, j) + BP(i, j, k) * o->dz[k];
      }
  for (j = 1; j <= NJ; j++) {
    if (K1(1, j) < NK)
      for (k = K1(1, j); k <= NK; k++) BP(NI + 1, j, k) = BP(1, j, k);
    if (K1(1, j) <= NK) SBP(NI + 1, j) = SBP(1, j);
  }
  for (j = 1; j <= NJ - 1; j++)
    for (i = 1; i <= NI; i++) {
      ip1 = i % NI + 1;
      l = i + j * n;
      if (GETJ(i, j)) {
        tv1 = 0;
        for (k = KU(2, ip1, j); k <= MK(ip1, j + 1); k++) tv1 = tv1 + BP(ip1, j + 1, k) * o->dz[k];
        tv2 = 0;
        for (k = KU(2, ip1, j); k <= MK(ip1, j); k++) tv2 = tv2 + BP(ip1, j, k) * o->dz[k];
        tv3 = 0;
        for (k = KU(2, i, j); k <= MK(i, j + 1); k++) tv3 = tv3 + BP(i, j + 1, k) * o->dz[k];
        tv4 = 0;
        for (k = KU(2, i, j); k <= MK(i, j); k++) tv4 = tv4 + BP(i, j, k) * o->dz[k];
        o->gb[l] = o->gbold[l] + ((tv3 + SBP(i, j + 1) - tv4 - SBP(i, j)) * RH(2, i, j) -
                                  (tv1 + SBP(ip1, j + 1) - tv2 - SBP(ip1, j)) * RH(2, ip1, j)) *
                                     o->rdphi * o->rdsv[j];
        tv1 = 0;
        for (k = KU(1, i, j + 1); k <= MK(ip1, j + 1); k++) tv1 = tv1 + BP(ip1, j + 1, k) * o->dz[k];
        tv2 = 0;
        for (k = KU(1, i, j); k <= MK(ip1, j); k++) tv2 = tv2 + BP(ip1, j, k) * o->dz[k];
        tv3 = 0;
        for (k = KU(1, i, j + 1); k <= MK(i, j + 1); k++) tv3 = tv3 + BP(i, j + 1, k) * o->dz[k];
        tv4 = 0;
        for (k = KU(1, i, j); k <= MK(i, j); k++) tv4 = tv4 + BP(i, j, k) * o->dz[k];
        o->gb[l] = o->gb[l] + ((tv1 + SBP(ip1, j + 1) - tv3 - SBP(i, j + 1)) * RH(1, i, j + 1) -
                               (tv2 + SBP(ip1, j) - tv4 - SBP(i, j)) * RH(1, i, j)) *
                                  o->rdphi * o->rdsv[j];
      } else {
        o->gb[l] = o->gbold[l];
      }
    }
}

/* goldstein.f90:3568-3679 */
static void velc(cgo_t *o) {
  int i, j, k, l;
  double tv, tv1, tv2, tv4, tv5, sum[3];
#define DZU(l, k) o->dzu[((l)-1) + 2 * ((k)-1)]
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      sum[1] = 0;
      sum[2] = 0;
      for (k = K1(i, j); k <= NK; k++) {
        if (K1(i + 1, j) > k) {
          tv1 = 0;
          tv2 = 0;
        } else {
          tv2 = -(RHO(i + 1, j, k) - RHO(i, j, k)) * o->rdphi * o->rc[j];
          if (imax2(imax2(K1(i, j - 1), K1(i, j + 1)), imax2(K1(i + 1, j - 1), K1(i + 1, j + 1))) <= k)
            tv1 = -o->c[j] * (RHO(i + 1, j + 1, k) - RHO(i + 1, j - 1, k) + RHO(i, j + 1, k) - RHO(i, j - 1, k)) *
                  o->rds2[j] * 0.25;
          else if (imax2(K1(i, j - 1), K1(i + 1, j - 1)) <= k)
            tv1 = -o->c[j] * (RHO(i + 1, j, k) - RHO(i + 1, j - 1, k) + RHO(i, j, k) - RHO(i, j - 1, k)) *
                  o->rdsv[j - 1] * 0.5;
          else if (imax2(K1(i, j + 1), K1(i + 1, j + 1)) <= k)
            tv1 = -o->c[j] * (RHO(i + 1, j + 1, k) - RHO(i + 1, j, k) + RHO(i, j + 1, k) - RHO(i, j, k)) *
                  o->rdsv[j] * 0.5;
          else
            tv1 = 0;
        }
        if (K1(i, j + 1) > k) {
          tv4 = 0;
          tv5 = 0;
        } else {
          tv4 = -o->cv[j] * (RHO(i, j + 1, k) - RHO(i, j, k)) * o->rdsv[j];
          if (imax2(imax2(K1(i - 1, j), K1(i - 1, j + 1)), imax2(K1(i + 1, j), K1(i + 1, j + 1))) <= k)
            tv5 = -(RHO(i + 1, j + 1, k) - RHO(i - 1, j + 1, k) + RHO(i + 1, j, k) - RHO(i - 1, j, k)) *
                  o->rdphi * 0.25 * o->rcv[j];
          else if (imax2(K1(i - 1, j), K1(i - 1, j + 1)) <= k)
            tv5 = -(RHO(i, j + 1, k) - RHO(i - 1, j + 1, k) + RHO(i, j, k) - RHO(i - 1, j, k)) * o->rdphi * 0.5 *
                  o->rcv[j];
          else if (imax2(K1(i + 1, j), K1(i + 1, j + 1)) <= k)
            tv5 = -(RHO(i + 1, j + 1, k) - RHO(i, j + 1, k) + RHO(i + 1, j, k) - RHO(i, j, k)) * o->rdphi * 0.5 *
                  o->rcv[j];
          else
            tv5 = 0;
        }
        if (k == NK) {
          if (K1(i + 1, j) <= k) {
            tv1 = tv1 - DZTAU(2, i, j);
            tv2 = tv2 - DZTAU(1, i, j);
          }
          if (K1(i, j + 1) <= k) {
            tv4 = tv4 - DZTAV(2, i, j);
            tv5 = tv5 - DZTAV(1, i, j);
          }
        }
        DZU(1, k) = -(o->s[j] * tv1 + DRAG(1, i, j) * tv2) * A2(o->rtv, i, j);
        DZU(2, k) = -(DRAG(2, i, j) * tv4 - o->sv[j] * tv5) * A2(o->rtv3, i, j);
        for (l = 1; l <= 2; l++) {
          if (k == K1(i, j)) {
            U(l, i, j, k) = 0;
          } else {
            U(l, i, j, k) = U(l, i, j, k - 1) + o->dza[k - 1] * (DZU(l, k) + DZU(l, k - 1)) * 0.5;
            sum[l] = sum[l] + o->dz[k] * U(l, i, j, k);
          }
        }
      }
      for (k = K1(i, j); k <= NK; k++) {
        if (K1(i + 1, j) <= k) {
          U(1, i, j, k) = U(1, i, j, k) - sum[1] * RH(1, i, j) + UB(1, i, j);
          U(1, i, j, k) = o->rel * U1(1, i, j, k) + (1.0 - o->rel) * U(1, i, j, k);
          U1(1, i, j, k) = U(1, i, j, k);
        }
        if (K1(i, j + 1) <= k) {
          U(2, i, j, k) = U(2, i, j, k) - sum[2] * RH(2, i, j) + UB(2, i, j);
          U(2, i, j, k) = o->rel * U1(2, i, j, k) + (1.0 - o->rel) * U(2, i, j, k);
          U1(2, i, j, k) = U(2, i, j, k);
        }
      }
    }
  for (j = 1; j <= NJ; j++)
    for (k = K1(1, j); k <= NK; k++) U(1, 0, j, k) = U(1, NI, j, k);
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      tv = 0;
      for (k = K1(i, j); k <= NK - 1; k++) {
        tv1 = (U(1, i, j, k) - U(1, i - 1, j, k)) * o->rdphi * o->rc[j];
        tv2 = (U(2, i, j, k) * o->cv[j] - U(2, i, j - 1, k) * o->cv[j - 1]) * o->rds[j];
        U(3, i, j, k) = tv - o->dz[k] * (tv1 + tv2);
        tv = U(3, i, j, k);
      }
    }
#undef DZU
}

/* goldstein.f90:2436-2642 */
void cgo_tstepo_flux(cgo_t *o) {
  const int L = NL;
  double tv, ups[4], pec[4];
  /* work arrays: allocated once per model (fe, fw, fn, fa, fwsave, dzts: L+1 each; fs; fb; dxts, dyts), zeroed per call as the
   * reference's locals are where it matters (fb = 0, fs = 0 per level) */
  const size_t n1 = (size_t)(L + 1), nfs = n1 * (NI + 1), nfb = n1 * (NI + 1) * (NJ + 1), nd = n1 * 5;
  if (!o->tf_scratch) o->tf_scratch = (double *)calloc(6 * n1 + nfs + nfb + 2 * nd, 8);
  double *fe = o->tf_scratch, *fw = fe + n1, *fn = fw + n1, *fa = fn + n1, *fwsave = fa + n1, *dzts = fwsave + n1;
  double *fs = dzts + n1, *fb = fs + nfs, *dxts = fb + nfb, *dyts = dxts + nd;
  memset(o->tf_scratch, 0, (6 * n1 + nfs + nfb + 2 * nd) * 8);
#define FS(l, i) fs[(l) + (L + 1) * (i)]
#define FB(l, i, j) fb[(l) + (L + 1) * ((i) + (NI + 1) * (j))]
#define DXTS(l, a) dxts[(l) + (L + 1) * (a)]
#define DYTS(l, a) dyts[(l) + (L + 1) * (a)]
  int i, j, k, l, ina, nnp, knp;
  double diffv, tec = 0, scc, dzrho = 0, rdzrho, slim, tv1, dxrho[5], dyrho[5];
  const double *diff = o->diff; /* 1-based */
  scc = 0.0;
  rdzrho = 0.0;
  if (o->diso) {
    scc = o->ec[2];
    o->limps = 0;
  }
  diffv = diff[2];
  o->dmax = 0;
  /* fb = 0 (calloc) */
  for (k = 1; k <= NK; k++) {
    memset(fs, 0, sizeof(double) * (size_t)(L + 1) * (NI + 1));
    for (j = 1; j <= NJ; j++) {
      pec[1] = U(1, NI, j, k) * o->dphi / diff[1];
      ups[1] = pec[1] / (2.0 + fabs(pec[1]));
      for (l = 1; l <= L; l++) {
        if (k >= imax2(K1(NI, j), K1(1, j))) {
          fw[l] = U(1, NI, j, k) * o->rc[j] * ((1.0 - ups[1]) * TS1(l, 1, j, k) + (1.0 + ups[1]) * TS1(l, NI, j, k)) * 0.5;
          fw[l] = fw[l] - (TS1(l, 1, j, k) - TS1(l, NI, j, k)) * o->rc2[j] * diff[1];
        } else {
          fw[l] = 0;
        }
        fwsave[l] = fw[l];
      }
      for (i = 1; i <= NI; i++) {
        if (k >= K1(i, j) && k < NK) {
          eosd(o, TS1(1, i, j, k), TS1(1, i, j, k + 1), TS1(2, i, j, k), TS1(2, i, j, k + 1), o->zw[k], o->rdza[k], &dzrho, &tec);
          if (dzrho < -1.0e-12)
            rdzrho = 1.0 / dzrho;
          else
            rdzrho = -1.0e12;
          if (o->iediff > 0 && o->iediff < 3) {   /* :2501-2515 (ediff1(i,j,k) = ediff1p(k): ediffvar = 0) */
            if (o->ediffpow2i == 0) diffv = o->ediff0 + o->ediff1p[k];
            else if (o->ediffpow2i == 1) diffv = o->ediff0 + o->ediff1p[k] * (-rdzrho);
            else if (o->ediffpow2i == 2) diffv = o->ediff0 + o->ediff1p[k] * sqrt(-rdzrho);
            else diffv = o->ediff0 + o->ediff1p[k] * pow(-rdzrho, o->ediffpow2);
            if (diffv > o->diffmax[k + 1]) diffv = o->diffmax[k + 1];
          }
        }
        pec[1] = U(1, i, j, k) * o->dphi / diff[1];
        ups[1] = pec[1] / (2.0 + fabs(pec[1]));
        pec[2] = U(2, i, j, k) * o->dsv[imin2(j, NJ - 1)] / diff[1];
        ups[2] = pec[2] / (2.0 + fabs(pec[2]));
        pec[3] = U(3, i, j, k) * o->dza[k] / diffv;
        ups[3] = pec[3] / (2.0 + fabs(pec[3]));
        for (l = 1; l <= L; l++) {
          if (i == NI) {
            fe[l] = fwsave[l];
          } else if (k < imax2(K1(i, j), K1(i + 1, j))) {
            fe[l] = 0;
          } else {
            fe[l] = U(1, i, j, k) * o->rc[j] * ((1.0 - ups[1]) * TS1(l, i + 1, j, k) + (1.0 + ups[1]) * TS1(l, i, j, k)) * 0.5;
            fe[l] = fe[l] - (TS1(l, i + 1, j, k) - TS1(l, i, j, k)) * o->rc2[j] * diff[1];
          }
          if (k < imax2(K1(i, j), K1(i, j + 1))) {
            fn[l] = 0;
          } else {
            fn[l] = o->cv[j] * U(2, i, j, k) * ((1.0 - ups[2]) * TS1(l, i, j + 1, k) + (1.0 + ups[2]) * TS1(l, i, j, k)) * 0.5;
            fn[l] = fn[l] - o->cv2[j] * (TS1(l, i, j + 1, k) - TS1(l, i, j, k)) * diff[1];
          }
          if (k < K1(i, j)) {
            fa[l] = 0;
          } else if (k == NK) {
            fa[l] = TS(l, i, j, NK + 1);
          } else {
            fa[l] = U(3, i, j, k) * ((1.0 - ups[3]) * TS1(l, i, j, k + 1) + (1.0 + ups[3]) * TS1(l, i, j, k)) * 0.5;
            fa[l] = fa[l] - (TS1(l, i, j, k + 1) - TS1(l, i, j, k)) * o->rdza[k] * diffv;
          }
        }
        if (o->diso) {
          if (k >= K1(i, j) && k < NK) {
            if (dzrho < -1.0e-12) {
              tv1 = 0.0;
              for (knp = 0; knp <= 1; knp++)
                for (nnp = 0; nnp <= 1; nnp++) {
                  ina = 1 + nnp + 2 * knp;
                  for (l = 1; l <= L; l++) {
                    if (k + knp >= K1(i - 1 + 2 * nnp, j))
                      DXTS(l, ina) = (TS1(l, i + nnp, j, k + knp) - TS1(l, i + nnp - 1, j, k + knp)) * o->rc[j] * o->rdphi;
                    else
                      DXTS(l, ina) = 0.0;
                    if (k + knp >= K1(i, j - 1 + 2 * nnp))
                      DYTS(l, ina) = (TS1(l, i, j + nnp, k + knp) - TS1(l, i, j + nnp - 1, k + knp)) * o->cv[j - 1 + nnp] *
                                     o->rdsv[j + nnp - 1];
                    else
                      DYTS(l, ina) = 0.0;
                  }
                  dxrho[ina] = scc * DXTS(2, ina) - tec * DXTS(1, ina);
                  dyrho[ina] = scc * DYTS(2, ina) - tec * DYTS(1, ina);
                  tv1 = tv1 + dxrho[ina] * dxrho[ina] + dyrho[ina] * dyrho[ina];
                }
              tv1 = 0.25 * tv1 * rdzrho * rdzrho;
              if (tv1 > o->ssmax[k]) {
                slim = o->ssmax[k] * o->ssmax[k] / (tv1 * tv1);
                o->limps = o->limps + 1;
              } else {
                slim = 1.0;
              }
              tv1 = tv1 * slim * diff[1] * o->rdza[k];
              tv = tv1 * o->dt[k] * o->rdza[k];
              if (tv > o->dmax) o->dmax = tv;
              for (l = 1; l <= L; l++) {
                dzts[l] = (TS1(l, i, j, k + 1) - TS1(l, i, j, k)) * o->rdza[k];
                tv = 0;
                for (ina = 1; ina <= 4; ina++)
                  tv = tv + (2 * dzrho * DXTS(l, ina) - dxrho[ina] * dzts[l]) * dxrho[ina] +
                       (2 * dzrho * DYTS(l, ina) - dyrho[ina] * dzts[l]) * dyrho[ina];
                tv = 0.25 * slim * diff[1] * tv / (dzrho * dzrho);
                fa[l] = fa[l] + tv;
              }
            }
          }
        }
        for (l = 1; l <= L; l++) {
          tv = 0;
          if (k >= K1(i, j))
            TS(l, i, j, k) = TS1(l, i, j, k) - o->dt[k] * (-tv + (fe[l] - fw[l]) * o->rdphi + (fn[l] - FS(l, i)) * o->rds[j] +
                                                           (fa[l] - FB(l, i, j)) * o->rdz[k]);
          fw[l] = fe[l];
          FS(l, i) = fn[l];
          FB(l, i, j) = fa[l];
        }
        cgo_eos(o, TS(1, i, j, k), TS(2, i, j, k), o->zro[k], &RHO(i, j, k));
      }
    }
  }
#undef FS
#undef FB
#undef DXTS
#undef DYTS
}

/* goldstein.f90:2657-2777, iconv==0, ieos==0 path */
/* coshuffle (Mueller convection scheme, iconv = 1), goldstein.f90:2781-2841: the surface box is moved down to the level of its own
 * density, the boxes in between move up by its thickness; icosd(i,j) = the deepest such displacement of this call (levels) */
static void coshuffle(cgo_t *o, int *icosd) {
  const int L = NL;
  int i, j, k, k0, l, ipass, maxpass = NK;
  double tv_temp;
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      if (K1(i, j) <= NK) {
        icosd[i + (NI + 1) * j] = 0;
        k0 = 0;
        ipass = 0;
        while (k0 < NK && ipass < maxpass) {
          ipass = ipass + 1;
          k = NK - 1;
          if (o->ieos != 0) {
            cgo_eos(o, TS(1, i, j, NK), TS(2, i, j, NK), o->zw[k], &RHO(i, j, NK));
            cgo_eos(o, TS(1, i, j, k), TS(2, i, j, k), o->zw[k], &RHO(i, j, k));
          }
          while (k >= K1(i, j) && (RHO(i, j, NK) > RHO(i, j, k))) {   /* (the reference tests rho first; with k < k1 it reads a dry level) */
            k = k - 1;
            if (o->ieos != 0) {
              cgo_eos(o, TS(1, i, j, NK), TS(2, i, j, NK), o->zw[k], &RHO(i, j, NK));
              cgo_eos(o, TS(1, i, j, k), TS(2, i, j, k), o->zw[k], &RHO(i, j, k));
            }
          }
          k0 = k + 1;
          if (k0 < NK) {
            for (l = 1; l <= L; l++) {
              tv_temp = TS(l, i, j, NK);
              for (k = NK; k >= k0 + 1; k--)
                TS(l, i, j, k) = ((o->dz[k] - o->dz[NK]) * TS(l, i, j, k) + o->dz[NK] * TS(l, i, j, k - 1)) * o->rdz[k];
              TS(l, i, j, k0) = ((o->dz[k0] - o->dz[NK]) * TS(l, i, j, k0) + o->dz[NK] * tv_temp) * o->rdz[k0];
            }
            for (k = k0; k <= NK; k++) cgo_eos(o, TS(1, i, j, k), TS(2, i, j, k), o->zw[NK - 1], &RHO(i, j, k));
            if (NK - k0 > icosd[i + (NI + 1) * j]) icosd[i + (NI + 1) * j] = NK - k0;
          }
        }
      }
    }
}

void cgo_co(cgo_t *o) {
  const int L = NL;
  int i, j, m, n, ni, lastmix, l, icond = 0;
  int *kk = (int *)calloc(NK + 2, sizeof(int));
  int *icosd = (int *)calloc((size_t)(NI + 1) * (NJ + 1), sizeof(int));
  double *dzm = (double *)calloc(NK + 2, 8), *sum = (double *)calloc(L + 1, 8);
  if (o->iconv == 1) coshuffle(o, icosd);   /* :2667-2672 */
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      if (K1(i, j) <= NK) {
        if (o->iconv == 1) icond = 0;
        kk[K1(i, j) - 1] = 0;
        for (m = K1(i, j); m <= NK; m++) {
          kk[m] = m;
          dzm[m] = o->dz[m];
        }
        m = NK;
        lastmix = 0;
        while (kk[m - 1] > 0 || (lastmix != 0 && kk[m] != NK)) {
          /* added code for thermobaricity (:2692-2698): both boxes at the depth of the interface below box k(m-1)'s top.  With
           * k(m-1) = 0 the reference evaluates level 0 of the halo into rho(i,j,0), which nothing reads: skipped */
          if (o->ieos != 0 && kk[m - 1] > 0) {
            cgo_eos(o, TS(1, i, j, kk[m]), TS(2, i, j, kk[m]), o->zw[kk[m - 1]], &RHO(i, j, kk[m]));
            cgo_eos(o, TS(1, i, j, kk[m - 1]), TS(2, i, j, kk[m - 1]), o->zw[kk[m - 1]], &RHO(i, j, kk[m - 1]));
          }
          if (kk[m - 1] == 0 || RHO(i, j, kk[m]) < RHO(i, j, kk[m - 1])) {
            if (lastmix == 0 || kk[m] == NK)
              m = m - 1;
            else
              m = m + 1;
            lastmix = 0;
          } else {
            lastmix = 1;
            n = m - 1;
            if (o->ieos != 0 && kk[n - 1] > 0) {   /* :2714-2720 */
              cgo_eos(o, TS(1, i, j, kk[n]), TS(2, i, j, kk[n]), o->zw[kk[n - 1]], &RHO(i, j, kk[n]));
              cgo_eos(o, TS(1, i, j, kk[n - 1]), TS(2, i, j, kk[n - 1]), o->zw[kk[n - 1]], &RHO(i, j, kk[n - 1]));
            }
            while (kk[n - 1] > 0 && RHO(i, j, kk[n]) >= RHO(i, j, kk[n - 1])) {
              n = n - 1;
              if (o->ieos != 0 && kk[n - 1] > 0) {   /* :2724-2730 */
                cgo_eos(o, TS(1, i, j, kk[n]), TS(2, i, j, kk[n]), o->zw[kk[n - 1]], &RHO(i, j, kk[n]));
                cgo_eos(o, TS(1, i, j, kk[n - 1]), TS(2, i, j, kk[n - 1]), o->zw[kk[n - 1]], &RHO(i, j, kk[n - 1]));
              }
            }
            for (l = 1; l <= L; l++) sum[l] = TS(l, i, j, kk[m]) * dzm[kk[m]];
            for (ni = 1; ni <= m - n; ni++) {
              for (l = 1; l <= L; l++) sum[l] = sum[l] + TS(l, i, j, kk[m - ni]) * dzm[kk[m - ni]];
              dzm[kk[m]] = dzm[kk[m]] + dzm[kk[m - ni]];
            }
            for (l = 1; l <= L; l++) TS(l, i, j, kk[m]) = sum[l] / dzm[kk[m]];
            cgo_eos(o, TS(1, i, j, kk[m]), TS(2, i, j, kk[m]), o->zw[kk[m - 1]], &RHO(i, j, kk[m]));
            ni = m - 1;
            while (kk[ni + 1] > 0) {
              kk[ni] = kk[ni - m + n];
              ni = ni - 1;
            }
          }
        }
        m = NK - 1;
        for (n = NK - 1; n >= K1(i, j); n--) {
          if (n > kk[m]) {
            for (l = 1; l <= L; l++) TS(l, i, j, n) = TS(l, i, j, kk[m + 1]);
            cgo_eos(o, TS(1, i, j, n), TS(2, i, j, n), o->zw[kk[n - 1]], &RHO(i, j, n));
            if (o->iconv == 1) icond = icond + 1;
            else A2(o->cost, i, j) = A2(o->cost, i, j) + 1.0;
          } else {
            m = m - 1;
          }
        }
        if (o->iconv == 1) {   /* :2766-2770: convection diagnostic needed by biogem = depth of the deepest convection of this call */
          if (icond > icosd[i + (NI + 1) * j]) icosd[i + (NI + 1) * j] = icond;
          A2(o->cost, i, j) = CG_DSC * o->zw[NK - 1 - icosd[i + (NI + 1) * j]];
        }
      }
    }
  free(kk); free(dzm); free(sum); free(icosd);
}

/* SUBROUTINE krausturner, goldstein.f90:3337-3442: the PE released in co and the KE from the wind deepen the mixed layer of
 * column (i, j); tv = ts(:, i, j, 1:maxk) in place.  mldt / mldkt keep their old values on the path that assigns neither. */
static void krausturner(cgo_t *o, int i, int j, double pebuoy, double ketau, double *mldt, int *mldkt, int tvkl) {
  const int L = NL;
  double em, empe, emke, eneed, emr, smix, tmix, qmix[CGO_MAXL + 1];
  double rhou = 0.0, rhol = 0.0, rhomix, mlda, mldb, mldtadd;
  int k, l, n, partmix;
  const int mldpk = NK;
  k = mldpk;
  empe = pebuoy;
  emke = ketau * o->mlddec[k];
  em = empe + emke;
  eneed = 0.0;
  tmix = TS(1, i, j, k);
  smix = TS(2, i, j, k);
  for (l = 1; l <= L; l++) qmix[l] = TS(l, i, j, k);
  cgo_eos(o, tmix, smix, o->zw[k], &rhomix);
  partmix = 1;
  while (eneed < em && em > 0) {
    if (k < mldpk) {
      qmix[1] = tmix;
      qmix[2] = smix;
      for (l = 3; l <= L; l++) qmix[l] = RDZG(NK, k) * (qmix[l] * DZG(NK, k + 1) + TS(l, i, j, k) * o->dz[k]);
    }
    if (k == tvkl) {
      *mldkt = k;
      *mldt = o->zw[k - 1];
      partmix = 0;
      em = -1.0e-8;
      if (k < NK)
        for (l = 1; l <= L; l++)
          for (n = k; n <= NK; n++) TS(l, i, j, n) = qmix[l];
    } else {
      k = k - 1;
      emr = (em - eneed) / em;
      empe = empe * emr;
      emke = emke * emr * o->mlddecd[k];
      em = empe + emke;
      if (o->ieos != 0) cgo_eos(o, tmix, smix, o->zw[k], &rhou); else rhou = rhomix;
      tmix = RDZG(NK, k) * (tmix * DZG(NK, k + 1) + TS(1, i, j, k) * o->dz[k]);
      smix = RDZG(NK, k) * (smix * DZG(NK, k + 1) + TS(2, i, j, k) * o->dz[k]);
      cgo_eos(o, TS(1, i, j, k), TS(2, i, j, k), o->zw[k], &rhol);
      cgo_eos(o, tmix, smix, o->zw[k], &rhomix);
      eneed = Z2DZG(NK, k + 1) * rhou + Z2DZG(k, k) * rhol - Z2DZG(NK, k) * rhomix;
    }
  }
  if (partmix == 1 && em > 0) {
    *mldkt = k;
    mlda = DZG(NK, k + 1) * RDZG(NK, k) * em / eneed;
    mldb = (DZG(NK, k) * RDZG(NK, k + 1) - 1) * mlda;
    for (l = 1; l <= L; l++) {
      TS(l, i, j, NK) = (1 - mldb) * qmix[l] + mldb * TS(l, i, j, k);
      for (n = k + 1; n <= NK - 1; n++) TS(l, i, j, n) = TS(l, i, j, NK);
      TS(l, i, j, k) = mlda * qmix[l] + (1 - mlda) * TS(l, i, j, k);
    }
    if (k < NK) {
      mldtadd = em / (o->zw[k] * (rhol - rhou));
      *mldt = o->zw[k] + mldtadd;
    } else {
      *mldt = 0.0;
    }
  } else if (partmix == 1 && k < NK) {
    *mldkt = k + 1;
    *mldt = o->zw[k];
  }
}

/* goldstein.f90:2280-2432 (dead copies :2311-2316 skipped) */
void cgo_tstepo(cgo_t *o) {
  int i, j, k, l;
  if (o->imld == 1) {   /* energy consumed or released in mixing the surface forcing over the top layer, :2294-2309 */
    for (j = 1; j <= NJ; j++)
      for (i = 1; i <= NI; i++) {
        const double t = TS(1, i, j, NK) - TS(1, i, j, NK + 1), sa = TS(2, i, j, NK) - TS(2, i, j, NK + 1);
        double r;
        cgo_eos(o, t, sa, o->zro[NK], &r);
        A2(o->mldpelayer1, i, j) = (r - RHO(i, j, NK)) * Z2DZG(NK, NK);
      }
  }
  cgo_tstepo_flux(o);
  if (o->imld == 1) {   /* :2320-2390 */
    double *tsold = (double *)malloc(sizeof(double) * 2 * (size_t)NI * NJ * (NK + 1));
#define TSOLD(l, i, j, k) tsold[((l)-1) + 2 * (((i)-1) + (size_t)NI * (((j)-1) + (size_t)NJ * (k)))]
    for (i = 1; i <= NI; i++)
      for (j = 1; j <= NJ; j++)
        if (K1(i, j) <= NK)
          for (k = 1; k <= NK; k++) { TSOLD(1, i, j, k) = TS(1, i, j, k); TSOLD(2, i, j, k) = TS(2, i, j, k); }
    cgo_co(o);
    for (i = 1; i <= NI; i++)
      for (j = 1; j <= NJ; j++)
        if (K1(i, j) <= NK) {
          double rold, rnew;
          A2(o->mldpeconv, i, j) = 0;
          for (k = NK; k > 0; k--) {
            cgo_eos(o, TSOLD(1, i, j, k), TSOLD(2, i, j, k), o->zro[k], &rold);
            cgo_eos(o, TS(1, i, j, k), TS(2, i, j, k), o->zro[k], &rnew);
            A2(o->mldpeconv, i, j) = A2(o->mldpeconv, i, j) + (rnew - rold) * Z2DZG(k, k);
          }
          A2(o->mldpebuoy, i, j) = A2(o->mldpeconv, i, j) + A2(o->mldpelayer1, i, j);
          if (A2(o->mldpebuoy, i, j) > 0.0) A2(o->mldpebuoy, i, j) = A2(o->mldpebuoy, i, j) * o->mldpebuoycoeff;
          A2(o->mldemix, i, j) = A2(o->mldpebuoy, i, j) + A2(o->mldketau, i, j) * o->mlddec[NK];
        }
    free(tsold);
#undef TSOLD
    for (i = 1; i <= NI; i++)
      for (j = 1; j <= NJ; j++)
        if (K1(i, j) <= NK) {
          if (A2(o->mldemix, i, j) > 0.0) {
            krausturner(o, i, j, A2(o->mldpebuoy, i, j), A2(o->mldketau, i, j), &A2(o->mld, i, j), &A2(o->mldk, i, j), K1(i, j));
          } else {
            A2(o->mldk, i, j) = NK;
            if (A2(o->mldpelayer1, i, j) < 0) A2(o->mld, i, j) = o->zw[NK - 1] * (1 - A2(o->mldemix, i, j) / A2(o->mldpelayer1, i, j));
            else A2(o->mld, i, j) = o->zw[NK - 1];
          }
        }
  } else {
    cgo_co(o);
  }
  if (o->ieos != 0) {   /* if thermobaricity is on, make sure rho calculation is vertically local (:2396-2408) */
    for (i = 1; i <= NI; i++)
      for (j = 1; j <= NJ; j++)
        if (K1(i, j) <= NK)
          for (k = 1; k <= NK; k++) cgo_eos(o, TS(1, i, j, k), TS(2, i, j, k), o->zro[k], &RHO(i, j, k));
  }
  for (j = 1; j <= NJ; j++) {
    for (k = K1(0, j); k <= NK; k++) {
      RHO(0, j, k) = RHO(NI, j, k);
      for (l = 1; l <= NL; l++) TS1(l, 0, j, k) = TS(l, NI, j, k);
    }
    for (k = K1(NI + 1, j); k <= NK; k++) {
      RHO(NI + 1, j, k) = RHO(1, j, k);
      for (l = 1; l <= NL; l++) TS1(l, NI + 1, j, k) = TS(l, 1, j, k);
    }
  }
  for (k = 1; k <= NK; k++)
    for (j = 1; j <= NJ; j++)
      for (i = 1; i <= NI; i++)
        for (l = 1; l <= NL; l++)
          if (k >= K1(i, j)) TS1(l, i, j, k) = TS(l, i, j, k);
}

/* goldstein.f90:3086-3129 (file output dropped) */
static void get_hosing(cgo_t *o, int istep) {
  int i, j;
  o->hosing = o->hosing + o->hosing_trend * CG_TSC * o->dt[NK];
  for (i = 1; i <= NI; i++)
    for (j = 1; j <= NJ; j++) {
      if (istep <= o->nsteps_hosing)
        A2(o->fw_hosing, i, j) = CG_M2MM * o->hosing * A2(o->rhosing, i, j);
      else
        A2(o->fw_hosing, i, j) = 0.0;
      A2(o->fw_anom, i, j) = A2(o->fw_anom, i, j) + A2(o->fw_anom_rate, i, j) * CG_TSC * o->dt[NK];
    }
}

/* goldstein.f90:198-233: barotropic + baroclinic momentum */
void cgo_momentum(cgo_t *o) {
  int i, j, isl, n;
  double s;
  wind(o);
  jbar(o);
  ubarsolv(o, o->ub, o->psi);
  for (isl = 1; isl <= o->isles; isl++) island(o, o->ub, &ERISL(isl, o->isles + 1), isl, 1);
  if (o->isles > 1) {
    matmult(o);
    for (isl = 1; isl <= o->isles; isl++) o->psibc[isl] = -ERISL(isl, o->isles + 1);
  } else if (o->isles == 1) {
    o->psibc[1] = -ERISL(1, 2) / ERISL(1, 1);
  }
  for (j = 1; j <= NJ; j++)
    for (i = 0; i <= NI + 1; i++) {
      s = 0.0;
      for (n = 1; n <= o->isles; n++) s = s + UBISL(1, i, j, n) * o->psibc[n];
      UB(1, i, j) = UB(1, i, j) + s;
      s = 0.0;
      for (n = 1; n <= o->isles; n++) s = s + UBISL(2, i, j, n) * o->psibc[n];
      UB(2, i, j) = UB(2, i, j) + s;
    }
  for (j = 0; j <= NJ; j++)
    for (i = 0; i <= NI; i++) {
      s = 0.0;
      for (n = 1; n <= o->isles; n++) s = s + PSISL(i, j, n) * o->psibc[n];
      PSI(i, j) = PSI(i, j) + s;
    }
  velc(o);
}

/* goldstein.f90:17-479 (diagnostic/file output dropped) */
void cgo_goldstein_step(cgo_t *o) {
  int i, j, k;
  const int istep = o->istep_ocn;
  double fx0neto, fwfxneto;
  if (o->go_lfirst) {
    double vsc = o->dphi * CG_RSC * CG_RSC;
    o->go_ini_energy = 0.0;
    o->go_ini_water = 0.0;
    for (k = 1; k <= NK; k++)
      for (j = 1; j <= NJ; j++)
        for (i = 1; i <= NI; i++) {
          o->go_ini_energy = o->go_ini_energy + TS(1, i, j, k) * o->dz[k] * o->ds[j];
          o->go_ini_water = o->go_ini_water - TS(2, i, j, k) * o->dz[k] * o->ds[j];
        }
    o->go_ini_energy = o->go_ini_energy * vsc * CG_DSC * CG_RH0SC * CG_CPSC;
    o->go_ini_water = CG_M2MM * o->go_ini_water * vsc * CG_DSC / o->saln0;
    o->go_lfirst = 0;
  }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      DZTAU(1, i, j) = o->scf * A2(o->stressxu, i, j) / (CG_RH0SC * CG_DSC * CG_USC * CG_FSC) / o->dzz;
      DZTAU(2, i, j) = o->scf * A2(o->stressyu, i, j) / (CG_RH0SC * CG_DSC * CG_USC * CG_FSC) / o->dzz;
      DZTAV(1, i, j) = o->scf * A2(o->stressxv, i, j) / (CG_RH0SC * CG_DSC * CG_USC * CG_FSC) / o->dzz;
      DZTAV(2, i, j) = o->scf * A2(o->stressyv, i, j) / (CG_RH0SC * CG_DSC * CG_USC * CG_FSC) / o->dzz;
      TAU(1, i, j) = DZTAU(1, i, j) * o->dzz;
      TAU(2, i, j) = DZTAV(2, i, j) * o->dzz;
    }
  if (o->imld == 1) {   /* wind energy input of the mixed-layer scheme, :112-141 */
    double tv2, tv3, tv4;
    for (j = 1; j <= NJ; j++) {
      tv3 = 0.0;
      for (i = 1; i <= NI; i++) {
        if (i == 1) tv4 = (TAU(1, i, j) + TAU(1, NI, j)) * 0.5; else tv4 = (TAU(1, i, j) + TAU(1, i - 1, j)) * 0.5;
        if (j == 1) tv2 = TAU(2, i, j) * 0.5; else tv2 = (TAU(2, i, (j < NJ - 1 ? j : NJ - 1)) + TAU(2, i, j - 1)) * 0.5;
        { const double r = sqrt(sqrt(tv4 * tv4 + tv2 * tv2)); A2(o->mldketau, i, j) = o->mldketaucoeff * (r * r * r); }
        tv3 = tv3 + A2(o->mldketau, i, j);
      }
      for (i = 1; i <= NI; i++)
        if (j <= 2 || j >= NJ - 1) A2(o->mldketau, i, j) = tv3 / NI;
    }
  }
  get_hosing(o, istep);
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      fx0neto = A2(o->netsolar_ocn, i, j) + A2(o->sensible_ocn, i, j) + A2(o->netlong_ocn, i, j) +
                A2(o->latent_ocn, i, j) + A2(o->conductflux_ocn, i, j);
      fwfxneto = A2(o->precip_ocn, i, j) + A2(o->evap_ocn, i, j) + A2(o->runoff_ocn, i, j) +
                 A2(o->waterflux_ocn, i, j) + A2(o->fw_hosing, i, j) + A2(o->fw_anom, i, j);
      fwfxneto = fwfxneto * CG_MM2M;
      TS(1, i, j, NK + 1) = -fx0neto * CG_RFLUXSC;
      TS(2, i, j, NK + 1) = fwfxneto * o->rpmesco;
      TS1(1, i, j, NK + 1) = TS(1, i, j, NK + 1);
      TS1(2, i, j, NK + 1) = TS(2, i, j, NK + 1);
    }
  for (j = 1; j <= NJ; j++) {
    int l;
    for (k = K1(0, j); k <= NK; k++)
      for (l = 1; l <= NL; l++) TS1(l, 0, j, k) = TS(l, NI, j, k);
    for (k = K1(NI + 1, j); k <= NK; k++)
      for (l = 1; l <= NL; l++) TS1(l, NI + 1, j, k) = TS(l, 1, j, k);
  }
  cgo_momentum(o);
  cgo_tstepo(o);
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A2(o->ustar_ocn, i, j) = U(1, i, j, NK);
      A2(o->vstar_ocn, i, j) = U(2, i, j, NK);
      if (K1(i, j) <= NK) {
        A2(o->tstar_ocn, i, j) = TS(1, i, j, NK);
        A2(o->sstar_ocn, i, j) = TS(2, i, j, NK);
        A2(o->albedo_ocn, i, j) = o->albocn;
      } else {
        A2(o->tstar_ocn, i, j) = 0.0;
        A2(o->sstar_ocn, i, j) = 0.0;
        A2(o->albedo_ocn, i, j) = 0.0;
      }
    }
  /* energy/water conservation diagnostic (goldstein.f90:458-478), every call */
  {
    double vsc = o->dphi * CG_RSC * CG_RSC, tot_energy = 0.0, tot_water = 0.0, se, sw;
    for (k = 1; k <= NK; k++)
      for (j = 1; j <= NJ; j++) {
        se = 0.0;
        sw = 0.0;
        for (i = 1; i <= NI; i++) {
          se = se + TS(1, i, j, k);
          sw = sw + TS(2, i, j, k);
        }
        tot_energy = tot_energy + se * o->dz[k] * o->ds[j];
        tot_water = tot_water - sw * o->dz[k] * o->ds[j];
      }
    tot_energy = tot_energy * vsc * CG_DSC * CG_RH0SC * CG_CPSC;
    tot_water = CG_M2MM * tot_water * vsc * CG_DSC / o->saln0;
    o->test_energy_ocean = tot_energy - o->go_ini_energy;
    o->test_water_ocean = tot_water - o->go_ini_water;
  }
}

/* goldstein.f90:514-2084 (restart, netCDF, file output and grid-export parts dropped) */
void cgo_goldstein_init(cgo_t *o) {
  int i, j, k, l, kk, isol, isl;
  const double pi = CG_PI;
  double th0, th1, s0, s1, phix, dscon, dth, thv, theta, tv, tv1, tv2, tv3, tv4, tv5, z1, ez0;
  double syr = o->yearlen * 86400;
  int j_hosing[3] = {0, 0, 0};
  double area_hosing;
  o->rpmesco = CG_RSC * o->saln0 / (CG_DSC * CG_USC);
  th0 = -pi / 2;
  th1 = pi / 2;
  s0 = sin(th0);
  s1 = sin(th1);
  phix = 2 * pi;
  o->dphi = phix / NI;
  o->rdphi = 1.0 / o->dphi;
  o->sv[0] = s0;
  o->cv[0] = cos(th0);
  if (o->igrid == 1) {
    dth = (th1 - th0) / NJ;
    for (j = 1; j <= NJ; j++) {
      thv = th0 + j * dth;
      theta = thv - 0.5 * dth;
      o->sv[j] = sin(thv);
      o->s[j] = sin(theta);
      o->cv[j] = cos(thv);
    }
  } else if (o->igrid == 0) {
    dscon = (s1 - s0) / NJ;
    for (j = 1; j <= NJ; j++) {
      o->sv[j] = s0 + j * dscon;
      o->cv[j] = sqrt(1 - o->sv[j] * o->sv[j]);
      o->s[j] = o->sv[j] - 0.5 * dscon;
    }
  }
  for (j = 1; j <= NJ; j++) {
    o->ds[j] = o->sv[j] - o->sv[j - 1];
    o->rds[j] = 1.0 / o->ds[j];
    o->c[j] = sqrt(1 - o->s[j] * o->s[j]);
    o->rc[j] = 1.0 / o->c[j];
    o->rc2[j] = o->rc[j] * o->rc[j] * o->rdphi;
    if (j < NJ) {
      o->dsv[j] = o->s[j + 1] - o->s[j];
      o->rdsv[j] = 1.0 / o->dsv[j];
      o->rcv[j] = 1.0 / o->cv[j];
      o->cv2[j] = o->cv[j] * o->cv[j] * o->rdsv[j];
      if (j > 1) o->rds2[j] = 2.0 / (o->dsv[j] + o->dsv[j - 1]);
    }
  }
  for (j = 1; j <= NJ; j++) o->asurf[j] = CG_RSC * CG_RSC * o->ds[j] * o->dphi;
  tv = 86400.0 * o->yearlen / (o->nyear * CG_TSC);
  for (k = 1; k <= NK; k++) o->dt[k] = tv;
  /* vertical grid :983-1060 */
  ez0 = 0.1;
  z1 = ez0 * (pow(1.0 + 1 / ez0, 1.0 / NK) - 1.0);
  tv4 = ez0 * (pow(z1 / ez0 + 1, 0.5) - 1);
  tv2 = 0;
  tv1 = 0;
  o->zro[NK] = -tv4;
  o->zw[NK] = tv2;
  for (k = 1; k <= NK; k++) {
    tv3 = ez0 * (powi_(z1 / ez0 + 1, k) - 1);
    o->dz[NK - k + 1] = tv3 - tv2;
    tv2 = tv3;
    tv5 = ez0 * (pow(z1 / ez0 + 1, k + 0.5) - 1);
    if (k < NK) o->dza[NK - k] = tv5 - tv4;
    tv4 = tv5;
    tv1 = tv1 + o->dz[NK - k + 1];
  }
  for (k = NK; k >= 1; k--) {
    if (k > 1) o->zro[k - 1] = o->zro[k] - o->dza[k - 1];
    o->zw[k - 1] = o->zw[k] - o->dz[k];
  }
  /* 2-D "depth" grids of the difference between each pair of levels, and of the squares (PE), :1013-1027 */
  for (k = NK; k >= 1; k--)
    for (kk = NK; kk >= 1; kk--) {
      DZG(k, kk) = o->zw[k] - o->zw[kk - 1];
      Z2DZG(k, kk) = -o->zw[k] * o->zw[k] + o->zw[kk - 1] * o->zw[kk - 1];
      if (k != kk - 1) RDZG(k, kk) = 1.0 / DZG(k, kk); else RDZG(k, kk) = 1.0e10;
    }
  o->dzz = o->dz[NK] * o->dza[NK - 1] / 2;
  for (k = 1; k <= NK - 1; k++) {
    o->rdz[k] = 1.0 / o->dz[k];
    o->rdza[k] = 1.0 / o->dza[k];
  }
  o->rdz[NK] = 1.0 / o->dz[NK];
  o->dza[NK] = 0.0;
  o->ec[1] = -0.0559 / CG_RHOSC;
  o->ec[2] = 0.7968 / CG_RHOSC;
  o->ec[3] = -0.0063 / CG_RHOSC;
  o->ec[4] = 3.7315e-5 / CG_RHOSC;
  if (o->ieos == 1) o->ec[5] = 2.5e-5 * CG_DSC / CG_RHOSC; else o->ec[5] = 0.0;   /* :1066-1075 */
  o->hosing_trend = o->hosing_trend / (1.0e3 * syr);
  o->nsteps_hosing = o->nyears_hosing * o->nyear;
  /* k1 already loaded (periodic wrap applied) by cgo_create */
  o->ntot = 0;
  o->intot = 0;
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++)
      if (K1(i, j) <= NK) {
        o->ntot = o->ntot + NK - K1(i, j) + 1;
        o->intot = o->intot + NK - K1(i, j);
      }
  /* basins, no .bmask branch :1175-1253 */
  {
    int *ips = o->ips, *ipf = o->ipf, *ias = o->ias, *iaf = o->iaf;
    ias[NJ] = nint_(NI * 24.0 / 36.0);
    ips[NJ] = nint_(NI * 10.0 / 36.0);
    o->jsf = 1;
    if (o->igrid != 0) {
      ias[NJ] = 61;
      ips[NJ] = 36;
      o->jsf = 10;
    }
    for (j = 1; j <= NJ; j++) {
      ips[j] = ips[NJ];
      ipf[j] = ips[j];
      ias[j] = ias[NJ];
      iaf[j] = ias[j];
      if (j > nint_(NJ * 34.0 / 36.0) && j <= nint_(NJ * 35.0 / 36.0)) ias[j] = nint_(NI * 20.0 / 36.0);
      for (i = 1; i <= NI; i++) {
        if (K1(ips[j] - 1, j) <= NK) ips[j] = ips[j] - 1;
        if (K1(ipf[j] + 1, j) <= NK) ipf[j] = ipf[j] + 1;
        if (K1(ias[j] - 1, j) <= NK) ias[j] = ias[j] - 1;
        if (K1(iaf[j] + 1, j) <= NK) iaf[j] = iaf[j] + 1;
        ips[j] = 1 + (ips[j] - 1 + NI) % NI;
        ipf[j] = 1 + (ipf[j] - 1 + NI) % NI;
        ias[j] = 1 + (ias[j] - 1 + NI) % NI;
        iaf[j] = 1 + (iaf[j] - 1 + NI) % NI;
      }
      if (o->igrid == 0) {
        if (ias[j] >= iaf[j] && j <= NJ / 2) o->jsf = j;
        if (ips[j] >= ipf[j] && j <= NJ / 2) o->jsf = j;
      }
    }
    if (o->igrid == 0)
      for (j = 1; j <= NJ; j++) {
        if (j > nint_(NJ * 35.0 / 36.0)) {
          ips[j] = 1;
          ipf[j] = 0;
          ias[j] = 1;
          iaf[j] = NI;
        }
        if (j > nint_(NJ * 34.0 / 36.0) && j <= nint_(NJ * 35.0 / 36.0)) {
          ips[j] = 1;
          ipf[j] = 0;
        }
      }
    if (o->igrid != 0) {
      ips[NJ] = 1; ipf[NJ] = 0; ips[NJ - 1] = 1; ipf[NJ - 1] = 0; ias[NJ] = 1; iaf[NJ] = NI;
    }
  }
  /* hosing region :1275-1309 */
  tv1 = sin(50.0 * pi / 180.0);
  tv2 = sin(70.0 * pi / 180.0);
  for (j = 1; j <= NJ; j++) {
    if (tv1 >= o->sv[j - 1] && tv1 <= o->sv[j]) {
      if (((o->sv[j] - tv1) / o->ds[j]) >= 0.5) j_hosing[1] = j; else j_hosing[1] = j + 1;
    }
    if (tv2 >= o->sv[j - 1] && tv2 <= o->sv[j]) {
      if (((tv2 - o->sv[j - 1]) / o->ds[j]) >= 0.5) j_hosing[2] = j; else j_hosing[2] = j - 1;
    }
  }
  area_hosing = 0.0;
  for (j = j_hosing[1]; j <= j_hosing[2]; j++)
    for (i = o->ias[j]; i <= o->iaf[j]; i++)
      if (K1(i, j) <= NK) area_hosing = area_hosing + o->asurf[j];
  for (j = j_hosing[1]; j <= j_hosing[2]; j++)
    for (i = o->ias[j]; i <= o->iaf[j]; i++)
      if (K1(i, j) <= NK) A2(o->rhosing, i, j) = 1e6 / area_hosing;
  /* fwanomin == 'n' : fw_anom = fw_anom_rate = 0 */
  /* seabed depth :1361-1391 */
  {
    double *h = (double *)calloc((size_t)3 * (NI + 2) * (NJ + 2), 8);
#define H(l, i, j) h[((l)-1) + 3 * ((i) + (NI + 2) * (j))]
    for (j = NJ + 1; j >= 0; j--)
      for (i = 0; i <= NI + 1; i++)
        if (K1(i, j) <= NK) {
          for (k = K1(i, j); k <= NK; k++) H(3, i, j) = H(3, i, j) + o->dz[k];
          RH(3, i, j) = 1.0 / H(3, i, j);
        }
    for (j = 0; j <= NJ + 1; j++)
      for (i = 0; i <= NI; i++) {
        H(1, i, j) = dmin2(H(3, i, j), H(3, i + 1, j));
        if (imax2(K1(i, j), K1(i + 1, j)) <= NK) RH(1, i, j) = 1.0 / H(1, i, j);
      }
    for (j = 0; j <= NJ; j++)
      for (i = 0; i <= NI + 1; i++) {
        H(2, i, j) = dmin2(H(3, i, j), H(3, i, j + 1));
        if (imax2(K1(i, j), K1(i, j + 1)) <= NK) RH(2, i, j) = 1.0 / H(2, i, j);
      }
#undef H
    free(h);
  }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      KU(1, i, j) = imax2(K1(i, j), K1(i + 1, j));
      KU(2, i, j) = imax2(K1(i, j), K1(i, j + 1));
    }
  o->adrag = 1.0 / (o->adrag_in * 86400 * CG_FSC);
  drgset(o, o->adrag, 3.0, NK / 2, 1);
  o->diff[1] = o->diff[1] / (CG_RSC * CG_USC);
  o->diff[2] = o->diff[2] * CG_RSC / (CG_USC * CG_DSC * CG_DSC);
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A2(o->rtv, i, j) = 1.0 / (o->s[j] * o->s[j] + DRAG(1, i, j) * DRAG(1, i, j));
      A2(o->rtv3, i, j) = 1.0 / (o->sv[j] * o->sv[j] + DRAG(2, i, j) * DRAG(2, i, j));
    }
  /* initial conditions :1434-1453 */
  for (i = 0; i <= NI + 1; i++)
    for (j = 0; j <= NJ + 1; j++) {
      for (k = 0; k <= NK + 1; k++) {
        if (j <= NJ / 2)
          TS(1, i, j, k) = o->temp0 * 0.5 * (1 + isign1(k - K1(i, j)));
        else
          TS(1, i, j, k) = o->temp1 * 0.5 * (1 + isign1(k - K1(i, j)));
        TS(2, i, j, k) = 0.0;
        TS1(1, i, j, k) = TS(1, i, j, k);
        TS1(2, i, j, k) = TS(2, i, j, k);
      }
      for (k = 1; k <= NK; k++) cgo_eos(o, TS(1, i, j, k), TS(2, i, j, k), o->zro[k], &RHO(i, j, k));
    }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) RHO(i, j, 0) = 0;
  /* mk, getj :1464-1494 */
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      int a = K1(i, j) * (1 + isign1(NK - K1(i, j))) / 2;
      int b = K1(i + 1, j) * (1 + isign1(NK - K1(i + 1, j))) / 2;
      int cc = K1(i - 1, j) * (1 + isign1(NK - K1(i - 1, j))) / 2;
      int d = K1(i, j + 1) * (1 + isign1(NK - K1(i, j + 1))) / 2;
      int e = K1(i, j - 1) * (1 + isign1(NK - K1(i, j - 1))) / 2;
      MK(i, j) = imax2(imax2(imax2(a, b), imax2(cc, d)), e);
      MK(i, j) = MK(i, j) * (1 + isign1(NK - K1(i, j))) / 2;
    }
  for (j = 1; j <= NJ; j++) MK(NI + 1, j) = MK(1, j);
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++)
      GETJ(i, j) = (imax2(imax2(K1(i, j), K1(i + 1, j)), imax2(K1(i, j + 1), K1(i + 1, j + 1))) <= NK) &&
                   (K1(i, j) != K1(i, j + 1) || K1(i, j) != K1(i + 1, j) || K1(i, j) != K1(i + 1, j + 1));
  /* islands: gbold <- psiles, isles already counted by cgo_create */
  for (j = NJ; j >= 0; j--)
    for (i = 1; i <= NI; i++) o->gbold[i + j * NI] = o->psiles[(i - 1) + NI * (NJ - j)];
  /* climatological albedo :1604-1608 */
  for (j = 1; j <= NJ; j++) {
    tv = asin(o->s[j]);
    tv2 = 0.2 + 0.36 * 0.5 * (1.0 - cos(2.0 * tv));
    for (i = 1; i <= NI; i++) A2(o->albcl_go, i, j) = tv2;
  }
  o->cd = 0.0013;
  o->rsictscsf = CG_DSC * o->dz[NK] * CG_RHO0 * CG_CPO_ICE / (17.5 * 86400.0);
  /* periodic b.c. :1761-1773 */
  for (k = 1; k <= NK; k++)
    for (j = 1; j <= NJ; j++) {
      RHO(0, j, k) = RHO(NI, j, k);
      RHO(NI + 1, j, k) = RHO(1, j, k);
      for (l = 1; l <= NL; l++) {
        TS(l, 0, j, k) = TS(l, NI, j, k);
        TS(l, NI + 1, j, k) = TS(l, 1, j, k);
        TS1(l, 0, j, k) = TS(l, NI, j, k);
        TS1(l, NI + 1, j, k) = TS(l, 1, j, k);
      }
    }
  if (o->isles > 0) invert(o);
  for (isol = 1; isol <= o->isles; isol++) {
    for (j = 0; j <= NJ; j++)
      for (i = 1; i <= NI; i++) {
        kk = i + j * NI;
        if ((int)o->gbold[kk] == isol + 1) o->gb[kk] = 1.0; else o->gb[kk] = 0.0;
      }
    ubarsolv(o, &UBISL(1, 0, 0, isol), &PSISL(0, 0, isol));
    for (isl = 1; isl <= o->isles; isl++) island(o, &UBISL(1, 0, 0, isol), &ERISL(isl, isol), isl, 0);
  }
  matinv_gold(o);
  /* mixed-layer scheme: wind decay efficiency :1675-1686 (mldwindkedec in units of dsc) */
  o->mldwindkedec = o->mldwindkedec / CG_DSC;
  for (k = NK; k >= 1; k--) {
    o->mlddec[k] = exp(o->zro[k] / o->mldwindkedec);
    if (k < NK) o->mlddecd[k] = o->mlddec[k] / o->mlddec[k + 1]; else o->mlddecd[NK] = o->mlddec[NK];
  }
  /* ssmax :2058-2070 */
  if (o->ssmaxsurf - o->ssmaxdeep < 1.0e-7 && o->ssmaxsurf - o->ssmaxdeep > -1.0e-7) {
    for (k = 1; k <= NK - 1; k++) o->ssmax[k] = o->ssmaxdeep;
  } else {
    double ssmaxmid = 0.5 * (log(o->ssmaxsurf) + log(o->ssmaxdeep));
    double ssmaxdiff = 0.5 * (log(o->ssmaxsurf) - log(o->ssmaxdeep));
    double ssmaxtanhefold = 200 / CG_DSC, ssmaxtanh0dep = -300 / CG_DSC, zssmax;
    for (k = 1; k <= NK - 1; k++) {
      zssmax = (o->zw[k] - ssmaxtanh0dep) / ssmaxtanhefold;
      o->ssmax[k] = exp(ssmaxmid + ssmaxdiff * tanh(zssmax));
    }
  }
  /* IF (iediff > 0) CALL ediff, :2053-2055; SUBROUTINE ediff :2936-3044 with ediffvar = 0 (no ediffvargrid.dat) */
  if (o->iediff > 0 && o->iediff < 3) {
    double ediff10, dzrho_lev, ediffk0 = 0.0, ediffklim;
    o->ediff0 = o->ediff0 * CG_RSC / (CG_USC * CG_DSC * CG_DSC);
    ediff10 = o->diff[2] - o->ediff0;
    for (k = 1; k <= NK - 1; k++) {
      dzrho_lev = (-5.5e-3 / CG_RHOSC * CG_DSC) * exp(o->zw[k] * (CG_DSC / 650.0));
      if (o->iediff == 1) {
        ediffk0 = exp(-(o->zw[k] + 2500.0 / CG_DSC) * (CG_DSC / 700.0));
        ediffklim = 1 / 3.0e0;
        ediffk0 = 1 / ((1 - ediffklim) / ediffk0 + ediffklim);
      } else {
        ediffk0 = 1 + (2 / CG_PI) * atan(-(o->zw[k] + 2500.0 / CG_DSC) * (4.5e-3 * CG_DSC));
      }
      o->ediff1p[k] = ediff10 * pow(ediffk0, o->ediffpow1) * pow(-dzrho_lev, o->ediffpow2);
    }
    if (o->ediffpow2 > -1.0e-7 && o->ediffpow2 < 1.0e-7) o->ediffpow2i = 0;
    else if (o->ediffpow2 > (1.0 - 1.0e-7) && o->ediffpow2 < (1.0 + 1.0e-7)) o->ediffpow2i = 1;
    else if (o->ediffpow2 > (0.5 - 1.0e-7) && o->ediffpow2 < (0.5 + 1.0e-7)) o->ediffpow2i = 2;
    else o->ediffpow2i = -999;
    for (k = 1; k <= NK; k++) o->diffmax[k] = 0.5 * 0.125 * o->dz[k] * o->dz[k] / o->dt[k];
  }
  /* output arguments :2015-2023 */
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A2(o->tstar_ocn, i, j) = TS(1, i, j, NK);
      A2(o->sstar_ocn, i, j) = TS(2, i, j, NK);
      A2(o->ustar_ocn, i, j) = U(1, i, j, NK);
      A2(o->vstar_ocn, i, j) = U(2, i, j, NK);
      A2(o->albedo_ocn, i, j) = o->albocn;
    }
  o->go_lfirst = 1;
}
