/* cgo_driver.c -- CPU oracle: construction, registry and the genie.f90 coupling
 * schedule (src/genie.f90:253-470).  TEST INFRASTRUCTURE ONLY. */
#include "cgo_impl.h"

void cgo_reg(cgo_t *o, const char *name, double *p, long n) {
  if (o->nfields >= 256) { fprintf(stderr, "cgo: field registry full\n"); abort(); }
  o->fields[o->nfields].name = name; o->fields[o->nfields].p = p; o->fields[o->nfields].n = n; o->nfields++;
}
void cgo_ireg(cgo_t *o, const char *name, int *p, long n) {
  if (o->nifields >= 32) { fprintf(stderr, "cgo: ifield registry full\n"); abort(); }
  o->ifields[o->nifields].name = name; o->ifields[o->nifields].p = p; o->ifields[o->nifields].n = n; o->nifields++;
}
void cgo_sreg(cgo_t *o, const char *name, double *p) {
  if (o->nscalars >= 64) { fprintf(stderr, "cgo: scalar registry full\n"); abort(); }
  o->scalars[o->nscalars].name = name; o->scalars[o->nscalars].p = p; o->nscalars++;
}
/* zeroed array with 64 doubles of slack either side (masked out-of-range reads
 * of the Fortran, e.g. rdsv(0), land never touch foreign memory) */
double *cgo_alloc(cgo_t *o, const char *name, long n) {
  double *p = (double *)calloc((size_t)n + 128, sizeof(double));
  if (!p) abort();
  if (name) cgo_reg(o, name, p + 64, n);
  return p + 64;
}
int *cgo_ialloc(cgo_t *o, const char *name, long n) {
  int *p = (int *)calloc((size_t)n + 128, sizeof(int));
  if (!p) abort();
  if (name) cgo_ireg(o, name, p + 64, n);
  return p + 64;
}

double *cgo_field(cgo_t *o, const char *name, long *n) {
  for (int i = 0; i < o->nfields; i++)
    if (!strcmp(o->fields[i].name, name)) { if (n) *n = o->fields[i].n; return o->fields[i].p; }
  return NULL;
}
int *cgo_ifield(cgo_t *o, const char *name, long *n) {
  for (int i = 0; i < o->nifields; i++)
    if (!strcmp(o->ifields[i].name, name)) { if (n) *n = o->ifields[i].n; return o->ifields[i].p; }
  return NULL;
}
double cgo_scalar(cgo_t *o, const char *name) {
  for (int i = 0; i < o->nscalars; i++)
    if (!strcmp(o->scalars[i].name, name)) return *o->scalars[i].p;
  if (!strcmp(name, "isles")) return o->isles;
  if (!strcmp(name, "jsf")) return o->jsf;
  if (!strcmp(name, "ntot")) return o->ntot;
  if (!strcmp(name, "limps")) return o->limps;
  if (!strcmp(name, "istep_ocn")) return o->istep_ocn;
  return cgo_biogem_scalar(o, name);
}
void cgo_set_scalar(cgo_t *o, const char *name, double v) {
  for (int i = 0; i < o->nscalars; i++)
    if (!strcmp(o->scalars[i].name, name)) { *o->scalars[i].p = v; return; }
  if (!strcmp(name, "istep_ocn")) o->istep_ocn = (int)v;
}

static int parse_kv(const char *params, const char *key, double *out) {
  /* last occurrence of "key=" at line start wins */
  size_t kl = strlen(key);
  const char *p = params;
  int found = 0;
  while (p && *p) {
    const char *e = strchr(p, '\n');
    if (!strncmp(p, key, kl) && p[kl] == '=') { *out = strtod(p + kl + 1, NULL); found = 1; }
    p = e ? e + 1 : NULL;
  }
  return found;
}
#define PD(name, dflt) do { double v_ = (dflt); parse_kv(params, #name, &v_); o->name = v_; } while (0)
#define PI_(name, dflt) do { double v_ = (dflt); parse_kv(params, #name, &v_); o->name = (int)v_; } while (0)

cgo_t *cgo_create(const char *params, const int *k1file, const double *psiles, int npaths, const int *npi,
                  const int *paths, const double *taux_u, const double *tauy_u, const double *taux_v,
                  const double *tauy_v, const double *uncep, const double *vncep) {
  cgo_t *o = (cgo_t *)calloc(1, sizeof(cgo_t));
  int i, j, p, isl;
  double v;
  if (!params) params = "";
  o->params = strdup(params);
  /* dims: main-defaults.nml */
  PI_(maxi, 36); PI_(maxj, 36); PI_(maxk, 8); PI_(maxl, 2);
  /* goldstein-defaults.nml */
  PI_(igrid, 0); PI_(nyear, 100); PD(yearlen, 365.25); PD(temp0, 5.0); PD(temp1, 5.0); PD(rel, 0.9);
  PD(scf, 2.0);
  v = 2000.0; parse_kv(params, "diff1", &v); o->diff[1] = v;
  v = 1.0e-5; parse_kv(params, "diff2", &v); o->diff[2] = v;
  v = 2.5; parse_kv(params, "adrag", &v); o->adrag_in = v;
  PD(hosing, 0.0); PD(hosing_trend, 0.0); PI_(nyears_hosing, 0); PD(albocn, 0.05);
  PI_(iconv, 0); PI_(imld, 0); PI_(iediff, 0); PI_(ieos, 0); PI_(diso, 1);
  PD(ssmaxsurf, 10.0); PD(ssmaxdeep, 10.0); PD(saln0, 34.9);
  PD(ediff0, 0.0); PD(ediffpow1, 1.0); PD(ediffpow2, 1.0); PD(ediffvar, 0.0);
  PD(mldpebuoycoeff, 0.15); PD(mldketaucoeff, 2.5); PD(mldwindkedec, 25.0);
  /* embm-defaults.nml */
  PI_(ndta, 5); PD(rmax, 0.85);
  v = 5.0e6; parse_kv(params, "diffamp1", &v); o->diffamp[1] = v;
  v = 1.0e6; parse_kv(params, "diffamp2", &v); o->diffamp[2] = v;
  PD(diffwid, 1.0); PD(difflin, 0.1);
  v = 0.0; parse_kv(params, "betaz1", &v); o->betaz[1] = v;
  v = 0.4; parse_kv(params, "betaz2", &v); o->betaz[2] = v;
  v = 0.0; parse_kv(params, "betam1", &v); o->betam[1] = v;
  v = 0.4; parse_kv(params, "betam2", &v); o->betam[2] = v;
  PD(tatm, 10.0); PD(relh0_ocean, 0.0); PD(relh0_land, 0.0);
  PD(extra1a, -0.03); PD(extra1b, 0.17); PD(extra1c, 0.18); PD(scl_fwf, 1.0); PD(z1_embm, 10.0);
  PD(diffa_scl, 1.0); PI_(diffa_len, 0); PD(delf2x, 5.77); PD(olr_adj0, 0.0); PD(olr_adj, 0.0); PD(t_eqm, 12.371);
  PD(albedop_offs, 0.20); PD(albedop_amp, 0.36); PD(albedop_skew, 0.0); PI_(albedop_skewp, 0);
  PD(albedop_mod2, 0.0); PD(albedop_mod4, 0.0); PD(albedop_mod6, 0.0);
  PD(par_sich_max, 9999.9); PD(par_albsic_min, 0.2); PD(par_albsic_max, 0.7); PI_(par_wind_polar_avg, 0);
  PD(radfor_scl_co2, 1.0); PD(radfor_pc_co2_rise, 0.0); PD(radfor_scl_ch4, 1.0); PD(radfor_pc_ch4_rise, 0.0);
  PD(radfor_scl_n2o, 1.0); PD(radfor_pc_n2o_rise, 0.0);
  /* goldsteinseaice-defaults.nml */
  v = 2000.0; parse_kv(params, "diffsic", &v); o->diffsic_in = v;
  PD(par_sica_thresh, 1.0); PD(par_sich_thresh, 1000.0);
  /* genie main */
  PD(solconst, 1368.0); PD(gn_daysperyear, 365.25);
  PI_(kocn_loop, 5); PI_(katm_loop, 1); PI_(ksic_loop, 5);
  if (o->iconv < 0 || o->iconv > 1 || o->imld < 0 || o->imld > 1 || o->iediff < 0 || o->iediff > 2 || o->ieos < 0 || o->ieos > 1 ||
      (o->iediff != 0 && (o->ediffvar < -1.0e-7 || o->ediffvar > 1.0e-7))) {
    fprintf(stderr, "cgo: imld / iconv / ieos outside 0..1, iediff outside 0..2 and ediffvar != 0 are outside the restated path\n");
    free(o);
    return NULL;
  }
  {
    const int I = NI, J = NJ, K = NK, L = NL;
    const long ij = (long)I * J;
    o->nm = I * (J + 1);
    o->mpi = 2 * (I + J);
    o->k1 = cgo_ialloc(o, "k1", (long)(I + 2) * (J + 2));
    o->ku = cgo_ialloc(o, "ku", 2 * ij);
    o->mk = cgo_ialloc(o, "mk", (long)(I + 1) * J);
    o->getj = cgo_ialloc(o, "getj", ij);
    o->ips = cgo_ialloc(o, "ips", J + 2); o->ipf = cgo_ialloc(o, "ipf", J + 2);
    o->ias = cgo_ialloc(o, "ias", J + 2); o->iaf = cgo_ialloc(o, "iaf", J + 2);
    o->iroff = cgo_ialloc(o, "iroff", ij); o->jroff = cgo_ialloc(o, "jroff", ij);
#define AL1(f_, n) o->f_ = cgo_alloc(o, #f_, (n))
    AL1(dt, K + 2); AL1(ds, J + 2); AL1(dsv, J + 2); AL1(rds2, J + 2); AL1(dz, K + 2); AL1(s, J + 2); AL1(c, J + 2);
    AL1(sv, J + 2); AL1(cv, J + 2); AL1(dza, K + 2); AL1(zro, K + 2); AL1(zw, K + 2); AL1(rc, J + 2); AL1(rc2, J + 2);
    AL1(rcv, J + 2); AL1(rdsv, J + 2); AL1(cv2, J + 2); AL1(rds, J + 2); AL1(rdz, K + 2); AL1(rdza, K + 2);
    AL1(asurf, J + 2); AL1(ssmax, K + 2); AL1(ediff1p, K + 2); AL1(diffmax, K + 2);
    AL1(rtv, ij); AL1(rtv3, ij);
    AL1(u, 3L * (I + 1) * (J + 1) * K); AL1(u1, 3L * (I + 1) * (J + 1) * K);
    AL1(ts, (long)L * (I + 2) * (J + 2) * (K + 2)); AL1(ts1, (long)L * (I + 2) * (J + 2) * (K + 2));
    AL1(rho, (long)(I + 2) * (J + 2) * (K + 1));
    AL1(tau, 2 * ij); AL1(drag, 2L * (I + 1) * J); AL1(dztau, 2 * ij); AL1(dztav, 2 * ij);
    AL1(ratm, (long)o->nm * (I + 1)); AL1(gap, (long)o->nm * (2 * I + 3));
    AL1(gb, o->nm + 2); AL1(gbold, o->nm + 2);
    AL1(ub, 2L * (I + 2) * (J + 1)); AL1(psi, (long)(I + 1) * (J + 1)); AL1(rh, 3L * (I + 2) * (J + 2));
    AL1(cost, ij); AL1(bp, (long)(I + 1) * J * K); AL1(sbp, (long)(I + 1) * J);
    AL1(fw_hosing, ij); AL1(rhosing, ij); AL1(fw_anom, ij); AL1(fw_anom_rate, ij); AL1(albcl_go, ij);
    AL1(dzu, 2L * K);
    AL1(dzg, (long)(K + 1) * (K + 1)); AL1(z2dzg, (long)(K + 1) * (K + 1)); AL1(rdzg, (long)(K + 1) * (K + 1));
    AL1(mlddec, K + 2); AL1(mlddecd, K + 2);
    AL1(mldketau, ij); AL1(mldpelayer1, ij); AL1(mldpeconv, ij); AL1(mldpebuoy, ij); AL1(mldemix, ij); AL1(mld, ij);
    o->mldk = cgo_ialloc(o, "mldk", ij);
    /* EMBM */
    AL1(tq, 2 * ij); AL1(tq1, 2 * ij); AL1(tqa, 2 * ij); AL1(uatm, 2 * ij); AL1(diffa, 4L * J); AL1(albcl, ij);
    AL1(ca, ij); AL1(co2, ij); AL1(ch4, ij); AL1(n2o, ij); AL1(usurf, ij); AL1(pmeadj, ij); AL1(pptn, ij);
    AL1(evap, ij); AL1(fxsw, ij); AL1(fxplw, ij); AL1(fx0a, ij); AL1(fx0o, ij); AL1(fxsen, ij); AL1(fxlata, ij);
    AL1(fxlw, ij); AL1(qb, ij); AL1(qbsic, ij); AL1(fx0sic, ij); AL1(fx0neto_eb, ij); AL1(evapsic, ij);
    AL1(tsfreez, ij); AL1(qsata, ij); AL1(qsato, ij); AL1(q_pa, ij); AL1(rq_pa, ij);
    AL1(solfor, (long)J * o->nyear); AL1(us_dztau, 2 * ij); AL1(us_dztav, 2 * ij);
    AL1(eb_tau, 2 * ij); AL1(eb_dztau, 2 * ij); AL1(eb_dztav, 2 * ij);
    /* sea ice */
    AL1(varice, 2 * ij); AL1(varice1, 2 * ij); AL1(dtha, 2 * ij); AL1(sic_u, 2L * (I + 1) * (J + 1));
    /* coupling */
    AL1(tstar_ocn, ij); AL1(sstar_ocn, ij); AL1(ustar_ocn, ij); AL1(vstar_ocn, ij); AL1(albedo_ocn, ij);
    AL1(tstar_atm, ij); AL1(qstar_atm, ij); AL1(hght_sic, ij); AL1(frac_sic, ij); AL1(temp_sic, ij); AL1(albd_sic, ij);
    AL1(stressxu, ij); AL1(stressyu, ij); AL1(stressxv, ij); AL1(stressyv, ij);
    AL1(latent_ocn, ij); AL1(sensible_ocn, ij); AL1(netsolar_ocn, ij); AL1(netlong_ocn, ij); AL1(evap_ocn, ij);
    AL1(precip_ocn, ij); AL1(runoff_ocn, ij); AL1(runoff_land, ij); AL1(latent_atm, ij); AL1(sensible_atm, ij);
    AL1(netsolar_atm, ij); AL1(netlong_atm, ij); AL1(evap_atm, ij); AL1(precip_atm, ij); AL1(dhght_sic, ij);
    AL1(dfrac_sic, ij); AL1(waterflux_ocn, ij); AL1(conductflux_ocn, ij); AL1(lowestlu2, ij); AL1(lowestlv3, ij);
    AL1(psiles, (long)I * (J + 1));
    AL1(go_solfor, J + 2);
    cgo_alloc(o, "bg_windspeed", ij);
#undef AL1
    cgo_sreg(o, "dphi", &o->dphi); cgo_sreg(o, "rdphi", &o->rdphi); cgo_sreg(o, "dzz", &o->dzz);
    cgo_sreg(o, "diff1", &o->diff[1]); cgo_sreg(o, "diff2", &o->diff[2]); cgo_sreg(o, "adrag", &o->adrag);
    cgo_sreg(o, "ec1", &o->ec[1]); cgo_sreg(o, "ec2", &o->ec[2]); cgo_sreg(o, "ec3", &o->ec[3]);
    cgo_sreg(o, "ec4", &o->ec[4]); cgo_sreg(o, "rpmesco", &o->rpmesco); cgo_sreg(o, "rsictscsf", &o->rsictscsf);
    cgo_sreg(o, "dmax", &o->dmax); cgo_sreg(o, "dtatm", &o->dtatm); cgo_sreg(o, "rdtdim", &o->rdtdim);
    cgo_sreg(o, "rfluxsca", &o->rfluxsca); cgo_sreg(o, "rpmesca", &o->rpmesca); cgo_sreg(o, "dtsic", &o->dtsic);
    cgo_sreg(o, "diffsic", &o->diffsic); cgo_sreg(o, "sic_rdtdim", &o->sic_rdtdim);
    cgo_sreg(o, "test_energy_ocean", &o->test_energy_ocean); cgo_sreg(o, "test_water_ocean", &o->test_water_ocean);
    cgo_sreg(o, "scf", &o->scf); cgo_sreg(o, "rel", &o->rel);

    /* k1: file rows j = maxj+1 .. 0 (goldstein.f90:1114-1123) */
    for (j = J + 1, p = 0; j >= 0; j--) {
      for (i = 0; i <= I + 1; i++) K1(i, j) = k1file[p++];
      K1(0, j) = K1(I, j);
      K1(I + 1, j) = K1(1, j);
    }
    memcpy(o->psiles, psiles, sizeof(double) * I * (J + 1));
    /* number of islands = max landmass id - 1 (goldstein.f90:1508-1518) */
    o->isles = 0;
    for (p = 0; p < I * (J + 1); p++)
      if (psiles[p] > (double)o->isles) o->isles = (int)psiles[p];
    o->isles = o->isles - 1;
    {
      double tro = 0.0;
      parse_kv(params, "tracer_only", &tro);
      if (tro != 0.0) o->isles = 0;  /* stand-alone tracer step on a synthetic grid: no barotropic/island set-up */
      else if (o->isles < 1 || npaths < o->isles) { fprintf(stderr, "cgo: need >=1 island + paths\n"); return NULL; }
    }
    isl = o->isles;
    o->npi = cgo_ialloc(o, "npi", isl + 2);
    o->lpisl = cgo_ialloc(o, "lpisl", (long)o->mpi * isl);
    o->ipisl = cgo_ialloc(o, "ipisl", (long)o->mpi * isl);
    o->jpisl = cgo_ialloc(o, "jpisl", (long)o->mpi * isl);
    o->psisl = cgo_alloc(o, "psisl", (long)(I + 1) * (J + 1) * isl);
    o->ubisl = cgo_alloc(o, "ubisl", 2L * (I + 2) * (J + 1) * isl);
    o->erisl = cgo_alloc(o, "erisl", (long)isl * (isl + 1));
    o->psibc = cgo_alloc(o, "psibc", isl + 2);
    for (i = 1, p = 0; i <= isl; i++) {
      o->npi[i] = npi[i - 1];
      if (o->npi[i] > o->mpi) { fprintf(stderr, "cgo: island path too long\n"); return NULL; }
      for (j = 1; j <= o->npi[i]; j++, p += 3) {
        o->lpisl[(j - 1) + o->mpi * (i - 1)] = paths[p];
        o->ipisl[(j - 1) + o->mpi * (i - 1)] = paths[p + 1];
        o->jpisl[(j - 1) + o->mpi * (i - 1)] = paths[p + 2];
      }
    }
  }
  /* genie.f90:79-82 order: ocean, atmosphere, sea ice */
  cgo_goldstein_init(o);
  if (taux_u) {
    cgo_embm_init(o, taux_u, tauy_u, taux_v, tauy_v, uncep, vncep);
    cgo_seaice_init(o);
  }
  o->istep_ocn = 0; o->istep_atm = 0; o->istep_sic = 0; o->koverall = 0;
  return o;
}

void cgo_destroy(cgo_t *o) {
  if (!o) return;
  for (int i = 0; i < o->nfields; i++) free(o->fields[i].p - 64);
  for (int i = 0; i < o->nifields; i++) free(o->ifields[i].p - 64);
  free(o->params);
  free(o->tf_scratch);
  free(o->bg);
  free(o);
}

/* genie.f90:117-534, "normal" branch, physics modules */
void cgo_run(cgo_t *o, long n) {
  for (long it = 0; it < n; it++) {
    long k = ++o->koverall;
    cgo_biogem_tick(o);
    if (k % o->kocn_loop == 1) { o->istep_ocn++; cgo_surflux(o); }
    if (k % o->katm_loop == 0) { o->istep_atm++; cgo_embm_step(o); }
    if (k % o->ksic_loop == 0) { o->istep_sic++; cgo_seaice_step(o); }
    if (k % o->kocn_loop == 0) { cgo_goldstein_step(o); }
    if (o->bg && cgo_biogem_koverall(o, k)) { fprintf(stderr, "cgo: BIOGEM carbonate chemistry failed at koverall %ld\n", k); }
  }
}
