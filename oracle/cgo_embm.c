/* cgo_embm.c -- CPU oracle, EMBM atmosphere + surflux.  TEST INFRASTRUCTURE ONLY.
 * Restates src/embm/embm.f90 of the reference for the non-ENTS configuration
 * (flag_ents=.FALSE., orogswitch=0, t_co2=0, useforc=.FALSE.; the ENTS/land,
 * orbit and orography branches are out of scope).
 * Parity unpinned (see cgo.h). */
#include "cgo_impl.h"

/* embm.f90:3771-3777 */
static double ch4_func(double ch4, double n2o) {
  return 0.47 * log(1.0 + 2.01e-5 * pow(ch4 * n2o, 0.75) + 5.31e-15 * ch4 * pow(ch4 * n2o, 1.52));
}

/* embm.f90:3787-3836 */
static void readroff(cgo_t *o) {
  int i, j, loop;
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      int ir = i, jr = j;
      loop = 0;
      while (K1(ir, jr) > NK) {
        int kv = K1(ir, jr);
        if (kv == 91) ir = ir + 1;
        else if (kv == 92) jr = jr - 1;
        else if (kv == 93) ir = ir - 1;
        else if (kv == 94) jr = jr + 1;
        if (ir == NI + 1) ir = 1;
        else if (ir == 0) ir = NI;
        loop = loop + 1;
        if (loop > 100000) { fprintf(stderr, "cgo: problem calculating runoff at %d %d\n", i, j); abort(); }
      }
      o->iroff[(i - 1) + NI * (j - 1)] = ir;
      o->jroff[(i - 1) + NI * (j - 1)] = jr;
    }
}

/* embm.f90:2383-2522 (fixed present-day orbit) */
static void radfor(cgo_t *o) {
  const double pi = CG_PI;
  double osce = 0.0167, oscsob = 0.397789, oscgam = 1.352631, osctau0 = -0.5;
  double rpi, tv, osce1, osce2, osce3, osce4, oscryr, osctau1, osct, oscv, oscsolf, oscsind, oscss, osccc, osctt, oscday;
  int istep, j;
  rpi = 1.0 / pi;
  tv = osce * osce;
  osce1 = osce * (2.0 - 0.25 * tv);
  osce2 = 1.25 * tv;
  osce3 = osce * tv * 13. / 12.;
  osce4 = ((1.0 + 0.5 * tv) / (1.0 - tv)) * ((1.0 + 0.5 * tv) / (1.0 - tv));
  oscryr = 2.0 * pi / (double)o->nyear;
  osctau1 = osctau0 + 0.5;
  for (istep = 1; istep <= o->nyear; istep++) {
    osct = ((double)((istep - 1) % o->nyear + 1) - (o->nyear * osctau1 / o->gn_daysperyear)) * oscryr;
    for (j = 1; j <= NJ; j++) {
      oscv = osct + osce1 * sin(osct) + osce2 * sin(2.0 * osct) + osce3 * sin(3.0 * osct);
      oscsolf = osce4 * ((1.0 + osce * cos(oscv)) * (1.0 + osce * cos(oscv)));
      oscsind = oscsob * sin(oscv - oscgam);
      oscss = oscsind * o->s[j];
      osccc = sqrt(1.0 - oscsind * oscsind) * o->c[j];
      osctt = dmin2(1.0, dmax2(-1.0, oscss / osccc));
      oscday = acos(-osctt);
      SOLFOR(j, istep) = o->solconst * oscsolf * rpi * (oscss * oscday + osccc * sin(oscday));
    }
  }
  /* dosc=.TRUE. in all target configs: no annual averaging */
}

/* embm.f90:198-2018 */
void cgo_embm_init(cgo_t *o, const double *taux_u, const double *tauy_u, const double *taux_v,
                   const double *tauy_v, const double *uncep, const double *vncep) {
  const double pi = CG_PI;
  int i, j, l;
  double tv, tv2, tv3, diffend;
  int j1as = 0, j1bs = 0, j1cs = 0, npac1a, natl1a, npac1b, natl1b, npac1c, natl1c;
  tv = 86400.0 * o->yearlen / (o->nyear * CG_TSC);
  o->ryear = 1.0 / (o->yearlen * 86400);
  o->dtatm = tv / o->ndta;
  o->rdtdim = 1.0 / (CG_TSC * o->dt[NK]);
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A3(o->us_dztau, 1, i, j) = taux_u[(i - 1) + NI * (j - 1)];
      A3(o->us_dztau, 2, i, j) = tauy_u[(i - 1) + NI * (j - 1)];
      A3(o->us_dztav, 1, i, j) = taux_v[(i - 1) + NI * (j - 1)];
      A3(o->us_dztav, 2, i, j) = tauy_v[(i - 1) + NI * (j - 1)];
    }
  /* climatological albedo :911-920 */
  for (j = 1; j <= NJ; j++) {
    double albedop_scl = powi_((o->albedop_skew - o->s[j]) / 2.0, o->albedop_skewp);
    tv = asin(o->s[j]);
    tv2 = o->albedop_offs + o->albedop_amp * 0.5 *
                                (1.0 - cos(2.0 * tv) + albedop_scl * o->albedop_mod2 * cos(2.0 * tv) +
                                 albedop_scl * o->albedop_mod4 * cos(4.0 * tv) + albedop_scl * o->albedop_mod6 * cos(6.0 * tv));
    for (i = 1; i <= NI; i++) A2(o->albcl, i, j) = tv2;
  }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) A2(o->ca, i, j) = (K1(i, j) <= NK) ? 0.3 : 1.0;
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A2(o->co2, i, j) = o->radfor_scl_co2 * CG_CO20;
      A2(o->ch4, i, j) = o->radfor_scl_ch4 * CG_CH40;
      A2(o->n2o, i, j) = o->radfor_scl_n2o * CG_N2O0;
    }
  o->rate_co2 = o->radfor_pc_co2_rise * 0.01 * CG_TSC * o->dtatm * o->ndta * o->ryear;
  o->rate_ch4 = o->radfor_pc_ch4_rise * 0.01 * CG_TSC * o->dtatm * o->ndta * o->ryear;
  o->rate_n2o = o->radfor_pc_n2o_rise * 0.01 * CG_TSC * o->dtatm * o->ndta * o->ryear;
  o->hatmbl[1] = 8400.0;
  o->rfluxsca = CG_RSC / (o->hatmbl[1] * CG_USC * CG_RHOAIR * CG_CPA);
  /* winds :985-1024 */
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A3(o->uatm, 1, i, j) = uncep[(i - 1) + NI * (j - 1)];
      A3(o->uatm, 2, i, j) = vncep[(i - 1) + NI * (j - 1)];
    }
  if (o->par_wind_polar_avg != 1 && o->par_wind_polar_avg != 2)
    for (j = 1; j <= NJ; j++)
      if (j <= 2 || j >= NJ - 1)
        for (l = 1; l <= 2; l++) {
          tv = 0.0;
          for (i = 1; i <= NI; i++) tv = tv + A3(o->uatm, l, i, j);
          tv = tv / NI;
          for (i = 1; i <= NI; i++) A3(o->uatm, l, i, j) = tv;
        }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A3(o->uatm, 1, i, j) = A3(o->uatm, 1, i, j) / CG_USC;
      A3(o->uatm, 2, i, j) = A3(o->uatm, 2, i, j) / CG_USC;
    }
  if (o->par_wind_polar_avg != 1 && o->par_wind_polar_avg != 2)
    for (i = 1; i <= NI; i++) A3(o->uatm, 2, i, NJ) = 0.;
  o->ppmin = 2.0 / (o->yearlen * 86400.0);
  o->ppmax = 4.0 / (o->yearlen * 86400.0);
  /* diffusivities :1031-1078 */
  diffend = exp(-((0.5 * pi / o->diffwid) * (0.5 * pi / o->diffwid)));
  for (j = 1; j <= NJ; j++) {
    tv = asin(o->s[j]);
    tv2 = asin(o->sv[j]);
    DIFFA(2, 1, j) = o->diffamp[2];
    DIFFA(2, 2, j) = o->diffamp[2];
    DIFFA(1, 1, j) = o->diffamp[1] * (o->difflin * 2.0 * (tv + 0.5 * pi) / pi +
                                      (1.0 - o->difflin) * (exp(-((tv / o->diffwid) * (tv / o->diffwid))) - diffend) / (1.0 - diffend));
    DIFFA(1, 2, j) = o->diffamp[1] * (o->difflin * 2.0 * (tv2 + 0.5 * pi) / pi +
                                      (1.0 - o->difflin) * (exp(-((tv2 / o->diffwid) * (tv2 / o->diffwid))) - diffend) / (1.0 - diffend));
    if (o->diffa_len < 0) {
      if (sin(pi * (double)o->diffa_len / 180.0) > o->sv[j]) DIFFA(1, 2, j) = o->diffa_scl * DIFFA(1, 2, j);
    } else {
      if (j <= o->diffa_len) DIFFA(1, 2, j) = o->diffa_scl * DIFFA(1, 2, j);
    }
    DIFFA(1, 1, j) = DIFFA(1, 1, j) / (CG_RSC * CG_USC);
    DIFFA(1, 2, j) = DIFFA(1, 2, j) / (CG_RSC * CG_USC);
    DIFFA(2, 1, j) = DIFFA(2, 1, j) / (CG_RSC * CG_USC);
    DIFFA(2, 2, j) = DIFFA(2, 2, j) / (CG_RSC * CG_USC);
    if (o->igrid == 1 || o->igrid == 2) DIFFA(2, 1, j) = dmin2(DIFFA(2, 1, j), DIFFA(1, 1, j));
  }
  o->hatmbl[2] = 1800.;
  o->rpmesca = CG_RSC * CG_RHO0 / (o->hatmbl[2] * CG_USC * CG_RHOAIR);
  /* usurf from tau==0 at init :1089-1111 -> zeros; recomputed by surflux */
  o->extra1a = o->scl_fwf * o->extra1a;
  o->extra1b = o->scl_fwf * o->extra1b;
  o->extra1c = o->scl_fwf * o->extra1c;
  /* P-E adjustment regions :1168-1318 (igrid==0) */
  j1as = o->jsf + 1;
  tv = sin(-20.0 * pi / 180.0);
  tv2 = sin(24.0 * pi / 180.0);
  for (j = 1; j <= NJ; j++) {
    if (tv >= o->sv[j - 1] && tv <= o->sv[j]) { if ((o->sv[j] - tv) / o->ds[j] >= 0.5) j1bs = j; else j1bs = j + 1; }
    if (tv2 >= o->sv[j - 1] && tv2 <= o->sv[j]) { if ((o->sv[j] - tv2) / o->ds[j] >= 0.5) j1cs = j; else j1cs = j + 1; }
  }
  if (o->igrid == 0) {
    npac1a = 0; natl1a = 0;
    for (j = j1as; j <= j1bs - 1; j++) { npac1a = npac1a + o->ipf[j] - o->ips[j] + 1; natl1a = natl1a + o->iaf[j] - o->ias[j] + 1; }
    npac1b = 0; natl1b = 0;
    for (j = j1bs; j <= j1cs - 1; j++) { npac1b = npac1b + o->ipf[j] - o->ips[j] + 1; natl1b = natl1b + o->iaf[j] - o->ias[j] + 1; }
    npac1c = 0; natl1c = 0;
    for (j = j1cs; j <= NJ; j++) {
      for (i = o->ips[j]; i <= o->ipf[j]; i++) if (K1(i, j) <= NK) npac1c = npac1c + 1;
      for (i = o->ias[j]; i <= o->iaf[j]; i++) if (K1(i, j) <= NK) natl1c = natl1c + 1;
    }
    for (j = j1as; j <= j1bs - 1; j++) {
      for (i = o->ips[j]; i <= o->ipf[j]; i++) A2(o->pmeadj, i, j) = 1.0e6 * o->extra1a / (npac1a * o->asurf[j]);
      for (i = o->ias[j]; i <= o->iaf[j]; i++) A2(o->pmeadj, i, j) = -1.0e6 * o->extra1a / (natl1a * o->asurf[j]);
    }
    for (j = j1bs; j <= j1cs - 1; j++) {
      for (i = o->ips[j]; i <= o->ipf[j]; i++) A2(o->pmeadj, i, j) = 1.0e6 * o->extra1b / (npac1b * o->asurf[j]);
      for (i = o->ias[j]; i <= o->iaf[j]; i++) A2(o->pmeadj, i, j) = -1.0e6 * o->extra1b / (natl1b * o->asurf[j]);
    }
    for (j = j1cs; j <= NJ; j++) {
      for (i = o->ips[j]; i <= o->ipf[j]; i++)
        if (K1(i, j) <= NK) A2(o->pmeadj, i, j) = 1.0e6 * o->extra1c / (npac1c * o->asurf[j]);
      for (i = o->ias[j]; i <= o->iaf[j]; i++)
        if (K1(i, j) <= NK) A2(o->pmeadj, i, j) = -1.0e6 * o->extra1c / (natl1c * o->asurf[j]);
    }
  }
  /* initial atmosphere :1435-1475 */
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      double to = A2(o->tstar_ocn, i, j);
      A3(o->tq, 1, i, j) = o->tatm;
      A3(o->tq1, 1, i, j) = A3(o->tq, 1, i, j);
      if (K1(i, j) <= NK) {
        if (to > CG_TSIC)
          A3(o->tq, 2, i, j) = o->relh0_ocean * CG_CONST1 * exp(CG_CONST2 * to / (to + CG_CONST3));
        else
          A3(o->tq, 2, i, j) = o->relh0_ocean * CG_CONST1 * exp(CG_CONST4 * to / (to + CG_CONST5));
      } else {
        double t1 = A3(o->tq1, 1, i, j);
        if (t1 > 0.0)
          A3(o->tq, 2, i, j) = o->relh0_land * CG_CONST1 * exp(CG_CONST2 * t1 / (t1 + CG_CONST3));
        else
          A3(o->tq, 2, i, j) = o->relh0_land * CG_CONST1 * exp(CG_CONST4 * t1 / (t1 + CG_CONST5));
      }
      A3(o->tq1, 2, i, j) = A3(o->tq, 2, i, j);
    }
  readroff(o);
  radfor(o);
  /* output arguments :1664-1683 */
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A2(o->tstar_atm, i, j) = A3(o->tq1, 1, i, j);
      A2(o->qstar_atm, i, j) = A3(o->tq1, 2, i, j);
      A2(o->stressxu, i, j) = A3(o->us_dztau, 1, i, j);
      A2(o->stressxv, i, j) = A3(o->us_dztav, 1, i, j);
      A2(o->stressyu, i, j) = A3(o->us_dztau, 2, i, j);
      A2(o->stressyv, i, j) = A3(o->us_dztav, 2, i, j);
      A2(o->lowestlu2, i, j) = A3(o->uatm, 1, i, j) * CG_USC;
      A2(o->lowestlv3, i, j) = A3(o->uatm, 2, i, j) * CG_USC;
    }
  (void)tv3;
}

/* embm.f90:2039-2138 */
void cgo_tstipa(cgo_t *o) {
  const int nii = 4;
  const double cimp = 0.5, diffmod0 = 0.0;
  double tv, ups, pec, diffpp, centre, dtloc = o->dtatm;
  int i, j, l, iits;
  size_t nc = (size_t)(NI + 1) * (NJ + 1), n2 = (size_t)(NI + 2) * (NJ + 2);
  double *cie = (double *)calloc(nc, 8), *ciw = (double *)calloc(nc, 8), *cin = (double *)calloc(nc, 8),
         *cis = (double *)calloc(nc, 8), *tq2 = (double *)calloc(n2, 8);
#define CIE(i, j) cie[(i) + (NI + 1) * (j)]
#define CIW(i, j) ciw[(i) + (NI + 1) * (j)]
#define CIN(i, j) cin[(i) + (NI + 1) * (j)]
#define CIS(i, j) cis[(i) + (NI + 1) * (j)]
#define TQ2(i, j) tq2[(i) + (NI + 2) * (j)]
  for (l = 1; l <= 2; l++) {
    for (j = 1; j <= NJ; j++)
      for (i = 1; i <= NI; i++) {
        double pp = dmax2(0.0, dmin2(1.0, (A2(o->pptn, i, j) - o->ppmin) / (o->ppmax - o->ppmin)));
        CIE(i, j) = o->betaz[l] * A3(o->uatm, 1, i, j) * o->rc[j] * 0.5 * o->rdphi;
        diffpp = DIFFA(l, 1, j) + (2 - l) * diffmod0 * pp;
        tv = o->rc[j] * o->rc[j] * o->rdphi * diffpp * o->rdphi;
        pec = o->betaz[l] * A3(o->uatm, 1, i, j) * o->dphi / diffpp;
        ups = pec / (2.0 + fabs(pec));
        CIW(i, j) = CIE(i, j) * (1 + ups) + tv;
        CIE(i, j) = CIE(i, j) * (1 - ups) - tv;
        CIN(i, j) = o->cv[j] * o->betam[l] * A3(o->uatm, 2, i, j) * 0.5;
        diffpp = DIFFA(l, 2, j) + (2 - l) * diffmod0 * pp;
        if (j < NJ) {
          tv = o->cv[j] * o->cv[j] * o->rdsv[j] * DIFFA(l, 2, j);
          pec = o->betam[l] * A3(o->uatm, 2, i, j) * o->dsv[j] / diffpp;
          ups = pec / (2.0 + fabs(pec));
        } else {
          tv = 0.0;
          ups = 0.0;
        }
        CIS(i, j) = CIN(i, j) * (1 + ups) + tv;
        CIN(i, j) = CIN(i, j) * (1 - ups) - tv;
      }
    for (j = 1; j <= NJ; j++) {
      CIE(0, j) = CIE(NI, j);
      CIW(0, j) = CIW(NI, j);
    }
    for (i = 0; i <= NI; i++) { CIN(i, 0) = 0.0; CIS(i, 0) = 0.0; TQ2(i, 0) = 0.0; TQ2(i, NJ + 1) = 0.0; }
    for (iits = 1; iits <= nii; iits++) {
      for (j = 1; j <= NJ; j++)
        for (i = 1; i <= NI; i++) TQ2(i, j) = cimp * A3(o->tq, l, i, j) + (1.0 - cimp) * A3(o->tq1, l, i, j);
      for (j = 1; j <= NJ; j++) { TQ2(0, j) = TQ2(NI, j); TQ2(NI + 1, j) = TQ2(1, j); }
      for (j = 1; j <= NJ; j++)
        for (i = 1; i <= NI; i++) {
          centre = dtloc * (CIW(i, j) - CIE(i - 1, j) + (CIS(i, j) - CIN(i, j - 1)) * o->rds[j]);
          A3(o->tq, l, i, j) =
              (A3(o->tq1, l, i, j) * (1.0 - (1.0 - cimp) * centre) -
               dtloc * (-A3(o->tqa, l, i, j) + CIE(i, j) * TQ2(i + 1, j) - CIW(i - 1, j) * TQ2(i - 1, j) +
                        (CIN(i, j) * TQ2(i, j + 1) - CIS(i, j - 1) * TQ2(i, j - 1)) * o->rds[j])) /
              (1 + cimp * centre);
        }
    }
    for (j = 1; j <= NJ; j++)
      for (i = 1; i <= NI; i++)
        TQ2(i, j) = 0.5 * (TQ2(i, j) + cimp * A3(o->tq, l, i, j) + (1.0 - cimp) * A3(o->tq1, l, i, j));
    for (j = 1; j <= NJ; j++) { TQ2(0, j) = TQ2(NI, j); TQ2(NI + 1, j) = TQ2(1, j); }
    for (j = 1; j <= NJ; j++)
      for (i = 1; i <= NI; i++)
        A3(o->tq, l, i, j) =
            A3(o->tq1, l, i, j) -
            dtloc * (-A3(o->tqa, l, i, j) + CIE(i, j) * TQ2(i + 1, j) - CIW(i - 1, j) * TQ2(i - 1, j) +
                     (CIN(i, j) * TQ2(i, j + 1) - CIS(i, j - 1) * TQ2(i, j - 1)) * o->rds[j]) -
            dtloc * TQ2(i, j) * (CIW(i, j) - CIE(i - 1, j) + (CIS(i, j) - CIN(i, j - 1)) * o->rds[j]);
  }
  memcpy(o->tq1, o->tq, sizeof(double) * 2 * NI * NJ);
  free(cie); free(ciw); free(cin); free(cis); free(tq2);
#undef CIE
#undef CIW
#undef CIN
#undef CIS
#undef TQ2
}

/* embm.f90:22-195 (file output dropped) */
void cgo_embm_step(cgo_t *o) {
  int i, j;
  double qsat;
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A3(o->uatm, 1, i, j) = A2(o->lowestlu2, i, j) / CG_USC;
      A3(o->uatm, 2, i, j) = A2(o->lowestlv3, i, j) / CG_USC;
      A3(o->tqa, 1, i, j) =
          (A2(o->netsolar_atm, i, j) + A2(o->latent_atm, i, j) + A2(o->sensible_atm, i, j) + A2(o->netlong_atm, i, j)) *
          o->rfluxsca;
      A3(o->tqa, 2, i, j) = A2(o->evap_atm, i, j) * CG_MM2M * o->rpmesca;
    }
  cgo_tstipa(o);
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      qsat = CG_CONST1 * exp(CG_CONST4 * A3(o->tq, 1, i, j) / (A3(o->tq, 1, i, j) + CG_CONST5));
      A2(o->q_pa, i, j) = dmin2(A3(o->tq, 2, i, j), o->rmax * qsat);
      A2(o->rq_pa, i, j) = A2(o->q_pa, i, j) / qsat;
    }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A2(o->tstar_atm, i, j) = A3(o->tq, 1, i, j);
      A2(o->qstar_atm, i, j) = A3(o->tq, 2, i, j);
      A2(o->stressxu, i, j) = A3(o->us_dztau, 1, i, j);
      A2(o->stressxv, i, j) = A3(o->us_dztav, 1, i, j);
      A2(o->stressyu, i, j) = A3(o->us_dztau, 2, i, j);
      A2(o->stressyv, i, j) = A3(o->us_dztav, 2, i, j);
    }
}

/* embm.f90:2548-3738, flag_ents=.FALSE. */
void cgo_surflux(cgo_t *o) {
  const int itice = 21;
  const double tol = 1.0e-10, zeroc = 273.15;
  const int istot = o->istep_ocn;
  int i, j, iter, nsol;
  double tv, tv2, tv3, tv0, tv1, rq, ch4_term, n2o_term, alw, salt, albsic, fxswsic, ticold, cesic, chsic, cfxsensic,
      qsatsic, tieqn, dtieq, fxlwsic, fxsensic, fx0sica, atm_latenti, atm_sensiblei, atm_netsoli, atm_netlongi, dhsic,
      ce, ch, fx0oa, atm_latent, atm_sensible, atm_netsol, atm_netlong, dho, meantemp;
  double *runoff = (double *)calloc((size_t)NI * NJ, 8);
  double *atemp = o->tstar_atm, *ashum = o->qstar_atm, *otemp = o->tstar_ocn, *osaln = o->sstar_ocn;
  double *sich = o->hght_sic, *sica = o->frac_sic, *tice = o->temp_sic;
#define ETAU(l, i, j) o->eb_tau[((l)-1) + 2 * (((i)-1) + NI * ((j)-1))]
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A3(o->eb_dztau, 1, i, j) = o->scf * A2(o->stressxu, i, j) / (CG_RH0SC * CG_DSC * CG_USC * CG_FSC) / o->dzz;
      A3(o->eb_dztau, 2, i, j) = o->scf * A2(o->stressyu, i, j) / (CG_RH0SC * CG_DSC * CG_USC * CG_FSC) / o->dzz;
      A3(o->eb_dztav, 1, i, j) = o->scf * A2(o->stressxv, i, j) / (CG_RH0SC * CG_DSC * CG_USC * CG_FSC) / o->dzz;
      A3(o->eb_dztav, 2, i, j) = o->scf * A2(o->stressyv, i, j) / (CG_RH0SC * CG_DSC * CG_USC * CG_FSC) / o->dzz;
      ETAU(1, i, j) = A3(o->eb_dztau, 1, i, j) * o->dzz;
      ETAU(2, i, j) = A3(o->eb_dztav, 2, i, j) * o->dzz;
    }
  for (j = 1; j <= NJ; j++) {
    tv3 = 0.0;
    for (i = 1; i <= NI; i++) {
      if (i == 1) tv = (ETAU(1, i, j) + ETAU(1, NI, j)) / 2; else tv = (ETAU(1, i, j) + ETAU(1, i - 1, j)) / 2;
      if (j == 1) tv2 = ETAU(2, i, j) / 2; else tv2 = (ETAU(2, i, j) + ETAU(2, i, j - 1)) / 2;
      A2(o->usurf, i, j) =
          sqrt((sqrt(tv * tv + tv2 * tv2)) * CG_RH0SC * CG_DSC * CG_USC * CG_FSC / (CG_RHOAIR * CG_CD * o->scf));
      tv3 = tv3 + A2(o->usurf, i, j);
    }
    if (o->par_wind_polar_avg != 2)
      if (j <= 2 || j >= NJ - 1)
        for (i = 1; i <= NI; i++) A2(o->usurf, i, j) = tv3 / NI;
  }
  /* greenhouse gases: option 4, compound increase :2913-2919 */
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      A2(o->co2, i, j) = (1.0 + o->rate_co2) * A2(o->co2, i, j);
      A2(o->ch4, i, j) = (1.0 + o->rate_ch4) * A2(o->ch4, i, j);
      A2(o->n2o, i, j) = (1.0 + o->rate_n2o) * A2(o->n2o, i, j);
    }
  for (i = 0; i < NI * NJ; i++) {
    o->evap[i] = 0.0; o->runoff_land[i] = 0.0; o->latent_ocn[i] = 0.0; o->sensible_ocn[i] = 0.0;
    o->netsolar_ocn[i] = 0.0; o->netlong_ocn[i] = 0.0; o->precip_ocn[i] = 0.0; o->runoff_ocn[i] = 0.0;
    o->latent_atm[i] = 0.0; o->sensible_atm[i] = 0.0; o->netsolar_atm[i] = 0.0; o->netlong_atm[i] = 0.0;
    o->dhght_sic[i] = 0.0; o->dfrac_sic[i] = 0.0; o->albedo_ocn[i] = o->albcl[i]; o->albd_sic[i] = 0.0;
  }
  meantemp = 0.0;
  for (i = 0; i < NI * NJ; i++) meantemp = meantemp + atemp[i];
  meantemp = meantemp / (double)(NJ * NI);
  nsol = (istot - 1) % o->nyear + 1;
  for (j = 1; j <= NJ; j++) o->go_solfor[j] = SOLFOR(j, nsol);   /* embm.f90:3727-3729 */
  for (i = 1; i <= NI; i++)
    for (j = 1; j <= NJ; j++) {
      const double at = A2(atemp, i, j);
      A2(o->qsata, i, j) = CG_CONST1 * exp(CG_CONST4 * at / (at + CG_CONST5));
      A2(o->pptn, i, j) = dmax2(0.0, (A2(ashum, i, j) - o->rmax * A2(o->qsata, i, j)) * CG_RHOAO * o->hatmbl[2] * o->rdtdim);
      A2(ashum, i, j) = dmin2(A2(ashum, i, j), o->rmax * A2(o->qsata, i, j));
      A3(o->tq1, 2, i, j) = A2(ashum, i, j);
      A3(o->tq, 2, i, j) = A3(o->tq1, 2, i, j);
      rq = A2(ashum, i, j) / A2(o->qsata, i, j);
      A2(o->fxsw, i, j) = SOLFOR(j, nsol) * (1.0 - A2(o->albcl, i, j));
      tv0 = 2.43414e2 + rq * (-3.47968e1 + 1.02790e1 * rq);
      tv1 = 2.60065 + rq * (-1.62064 + 6.34856e-1 * rq);
      tv2 = 4.40272e-3 + rq * (-2.26092e-2 + 1.12265e-2 * rq);
      tv3 = -2.05237e-5 + rq * (-9.67000e-5 + 5.62925e-5 * rq);
      ch4_term = CG_ALPHACH4 * (sqrt(1.0e9 * A2(o->ch4, i, j)) - sqrt(1.0e9 * CG_CH40)) -
                 ch4_func(1.0e9 * A2(o->ch4, i, j), 1.0e9 * CG_N2O0) + ch4_func(1.0e9 * CG_CH40, 1.0e9 * CG_N2O0);
      n2o_term = CG_ALPHAN2O * (sqrt(1.0e9 * A2(o->n2o, i, j)) - sqrt(1.0e9 * CG_N2O0)) -
                 ch4_func(1.0e9 * CG_CH40, 1.0e9 * A2(o->n2o, i, j)) + ch4_func(1.0e9 * CG_CH40, 1.0e9 * CG_N2O0);
      A2(o->fxplw, i, j) = tv0 + at * (tv1 + at * (tv2 + at * tv3)) - o->delf2x * log(A2(o->co2, i, j) / CG_CO20) -
                           ch4_term - n2o_term + o->olr_adj * (meantemp - o->t_eqm) - o->olr_adj0;
      A2(o->fxlata, i, j) = CG_RHO0 * A2(o->pptn, i, j) * CG_HLV;
      if (K1(i, j) <= NK) {
        const double ot = A2(otemp, i, j), us = A2(o->usurf, i, j), cca = A2(o->ca, i, j), sa = A2(sica, i, j);
        alw = at + zeroc;
        alw = alw * alw;
        alw = alw * alw;
        alw = CG_EMA * alw;
        salt = o->saln0 + A2(osaln, i, j);
        A2(o->tsfreez, i, j) = salt * (-0.0575 + 0.0017 * sqrt(salt) - 0.0002 * salt);
        A2(o->qb, i, j) = o->rsictscsf * (A2(o->tsfreez, i, j) - ot);
        A2(o->qbsic, i, j) = A2(o->qb, i, j);
        if (sa > 0.0) {
          albsic = dmax2(o->par_albsic_min, dmin2(o->par_albsic_max, 0.40 - 0.04 * at));
          fxswsic = SOLFOR(j, nsol) * (1.0 - albsic);
          for (iter = 1; iter <= itice; iter++) {
            double tz, tc3;
            ticold = A2(tice, i, j);
            cesic = 1.0e-3 * (1.0022 - 0.0822 * (at - ticold) + 0.0266 * us);
            cesic = dmax2(6.0e-5, dmin2(2.19e-3, cesic));
            chsic = 0.94 * cesic;
            cfxsensic = CG_RHOAIR * chsic * CG_CPA * us;
            qsatsic = CG_CONST1 * exp(CG_CONST2 * ticold / (ticold + CG_CONST3));
            A2(o->evapsic, i, j) = dmax2(0.0, (qsatsic - A2(ashum, i, j)) * CG_RHOAO * cesic * us);
            tz = ticold + zeroc;
            tieqn = A2(sich, i, j) * ((1 - cca) * fxswsic + alw - CG_EMO * ((tz * tz) * (tz * tz)) -
                                      cfxsensic * (ticold - at) - CG_RHO0 * CG_HLS * A2(o->evapsic, i, j)) +
                    CG_CONSIC * (A2(o->tsfreez, i, j) - ticold);
            tc3 = ticold + CG_CONST3;
            dtieq = A2(sich, i, j) * (-4.0 * CG_EMO * (tz * tz * tz) - cfxsensic -
                                      CG_HLS * CG_RHOAIR * cesic * us * qsatsic * CG_CONST2 * CG_CONST3 / (tc3 * tc3) * 0.5 *
                                          (1.0 + copysign(1.0, qsatsic - A2(ashum, i, j)))) -
                    CG_CONSIC;
            A2(tice, i, j) = ticold - tieqn / dtieq;
            if (fabs(A2(tice, i, j) - ticold) < tol || (ticold > CG_TFREEZ && tieqn > 0.0)) break;
          }
          A2(tice, i, j) = dmin2(CG_TFREEZ, A2(tice, i, j));
          {
            double tz = A2(tice, i, j) + zeroc;
            fxlwsic = CG_EMO * ((tz * tz) * (tz * tz)) - alw;
          }
          cesic = 1.0e-3 * (1.0022 - 0.0822 * (at - A2(tice, i, j)) + 0.0266 * us);
          cesic = dmax2(6.0e-5, dmin2(2.19e-3, cesic));
          chsic = 0.94 * cesic;
          cfxsensic = CG_RHOAIR * chsic * CG_CPA * us;
          fxsensic = cfxsensic * (A2(tice, i, j) - at);
          qsatsic = CG_CONST1 * exp(CG_CONST2 * A2(tice, i, j) / (A2(tice, i, j) + CG_CONST3));
          A2(o->evapsic, i, j) = dmax2(0.0, (qsatsic - A2(ashum, i, j)) * CG_RHOAO * cesic * us);
          A2(o->fx0sic, i, j) = (1 - cca) * fxswsic - fxsensic - fxlwsic - CG_RHO0 * CG_HLS * A2(o->evapsic, i, j);
          fx0sica = cca * fxswsic + A2(o->fxlata, i, j) + fxsensic + fxlwsic - A2(o->fxplw, i, j);
          atm_latenti = +A2(o->fxlata, i, j);
          atm_sensiblei = +fxsensic;
          atm_netsoli = +cca * fxswsic;
          atm_netlongi = +fxlwsic - A2(o->fxplw, i, j);
          dhsic = CG_RRHOLF * (A2(o->qb, i, j) - A2(o->fx0sic, i, j)) - CG_RHOOI * A2(o->evapsic, i, j);
          if (A2(sich, i, j) >= o->par_sich_max) {
            if (dhsic > 0.0) {
              A2(o->qbsic, i, j) = (0.0 + CG_RHOOI * A2(o->evapsic, i, j)) / CG_RRHOLF + A2(o->fx0sic, i, j);
              dhsic = CG_RRHOLF * (A2(o->qbsic, i, j) - A2(o->fx0sic, i, j)) - CG_RHOOI * A2(o->evapsic, i, j);
            }
          }
        } else {
          albsic = 0.0; fx0sica = 0.0; dhsic = 0.0;
          A2(o->evapsic, i, j) = 0.0;
          A2(tice, i, j) = 0.0;
          atm_latenti = 0.0; atm_sensiblei = 0.0; atm_netsoli = 0.0; atm_netlongi = 0.0;
        }
        {
          double tz = ot + zeroc;
          A2(o->fxlw, i, j) = CG_EMO * ((tz * tz) * (tz * tz)) - alw;
        }
        ce = 1.0e-3 * (1.0022 - 0.0822 * (at - ot) + 0.0266 * us);
        ce = dmax2(6.0e-5, dmin2(2.19e-3, ce));
        ch = 0.94 * ce;
        A2(o->fxsen, i, j) = CG_RHOAIR * ch * CG_CPA * us * (ot - at);
        A2(o->qsato, i, j) = CG_CONST1 * exp(CG_CONST4 * ot / (ot + CG_CONST5));
        A2(o->evap, i, j) = dmax2(0.0, (A2(o->qsato, i, j) - A2(ashum, i, j)) * CG_RHOAO * ce * us);
        fx0oa = cca * A2(o->fxsw, i, j) + A2(o->fxlata, i, j) + A2(o->fxsen, i, j) + A2(o->fxlw, i, j) - A2(o->fxplw, i, j);
        atm_latent = +A2(o->fxlata, i, j);
        atm_sensible = +A2(o->fxsen, i, j);
        atm_netsol = +cca * A2(o->fxsw, i, j);
        atm_netlong = +A2(o->fxlw, i, j) - A2(o->fxplw, i, j);
        A2(o->fx0a, i, j) = (1 - sa) * fx0oa + sa * fx0sica;
        A2(o->latent_atm, i, j) = (sa * atm_latenti) + ((1 - sa) * atm_latent);
        A2(o->sensible_atm, i, j) = (sa * atm_sensiblei) + ((1 - sa) * atm_sensible);
        A2(o->netsolar_atm, i, j) = (sa * atm_netsoli) + ((1 - sa) * atm_netsol);
        A2(o->netlong_atm, i, j) = (sa * atm_netlongi) + ((1 - sa) * atm_netlong);
        A2(o->fx0o, i, j) = (1 - cca) * A2(o->fxsw, i, j) - A2(o->fxsen, i, j) - A2(o->fxlw, i, j) -
                            CG_RHO0 * CG_HLV * A2(o->evap, i, j);
        A2(o->fx0neto_eb, i, j) = sa * A2(o->qbsic, i, j) + (1 - sa) * dmax2(A2(o->qb, i, j), A2(o->fx0o, i, j));
        A2(o->latent_ocn, i, j) =
            (1 - sa) * (-CG_RHO0 * CG_HLV * A2(o->evap, i, j) + dmax2(0.0, A2(o->qb, i, j) - A2(o->fx0o, i, j))) +
            sa * A2(o->qbsic, i, j);
        A2(o->sensible_ocn, i, j) = -((1 - sa) * A2(o->fxsen, i, j));
        A2(o->netsolar_ocn, i, j) = (1 - sa) * (1 - cca) * A2(o->fxsw, i, j);
        A2(o->netlong_ocn, i, j) = -((1 - sa) * A2(o->fxlw, i, j));
        dho = dmax2(0.0, CG_RRHOLF * (A2(o->qb, i, j) - A2(o->fx0o, i, j)));
        A2(o->dhght_sic, i, j) = sa * dhsic + (1 - sa) * dho;
        A2(o->dfrac_sic, i, j) = dmax2(0.0, CG_RHMIN * dho * (1 - sa));
        if (A2(sich, i, j) > 1.0e-12)
          A2(o->dfrac_sic, i, j) = A2(o->dfrac_sic, i, j) + dmin2(0.0, 0.5 * sa * sa * dhsic / A2(sich, i, j));
        A2(o->albedo_ocn, i, j) = sa * albsic + (1 - sa) * A2(o->albcl, i, j);
        A2(o->albd_sic, i, j) = albsic;
      } else {
        A2(o->fx0a, i, j) = A2(o->fxsw, i, j) + A2(o->fxlata, i, j) - A2(o->fxplw, i, j);
        A2(o->latent_atm, i, j) = +A2(o->fxlata, i, j);
        A2(o->sensible_atm, i, j) = +0.0;
        A2(o->netsolar_atm, i, j) = +A2(o->fxsw, i, j);
        A2(o->netlong_atm, i, j) = -A2(o->fxplw, i, j);
        {
          int ir = o->iroff[(i - 1) + NI * (j - 1)], jr = o->jroff[(i - 1) + NI * (j - 1)];
          if (o->igrid != 0)
            A2(runoff, ir, jr) = A2(runoff, ir, jr) + A2(o->pptn, i, j) * o->ds[j] * o->rds[jr];
          else
            A2(runoff, ir, jr) = A2(runoff, ir, jr) + A2(o->pptn, i, j);
        }
        A2(o->runoff_land, i, j) = A2(o->pptn, i, j);
      }
    }
  for (j = 1; j <= NJ; j++)
    for (i = 1; i <= NI; i++) {
      if (K1(i, j) <= NK) {
        A2(o->precip_ocn, i, j) = A2(o->pptn, i, j);
        A2(o->runoff_ocn, i, j) = A2(runoff, i, j) + 0.0;
        A2(o->evap_atm, i, j) = A2(o->evap, i, j) * (1 - A2(sica, i, j)) + A2(o->evapsic, i, j) * A2(sica, i, j);
        A2(o->precip_ocn, i, j) = A2(o->precip_ocn, i, j) + A2(o->pmeadj, i, j);
      } else {
        A2(o->precip_ocn, i, j) = 0.0;
        A2(o->runoff_ocn, i, j) = 0.0;
        A2(o->evap_atm, i, j) = 0.0;
      }
      A2(o->precip_atm, i, j) = A2(o->pptn, i, j);
      A2(o->evap_ocn, i, j) = -A2(o->evap_atm, i, j);
      A2(o->precip_ocn, i, j) = A2(o->precip_ocn, i, j) * CG_M2MM;
      A2(o->evap_ocn, i, j) = A2(o->evap_ocn, i, j) * CG_M2MM;
      A2(o->runoff_ocn, i, j) = A2(o->runoff_ocn, i, j) * CG_M2MM;
      A2(o->runoff_land, i, j) = A2(o->runoff_land, i, j) * CG_M2MM;
      A2(o->precip_atm, i, j) = A2(o->precip_atm, i, j) * CG_M2MM;
      A2(o->evap_atm, i, j) = A2(o->evap_atm, i, j) * CG_M2MM;
    }
  free(runoff);
#undef ETAU
}
