// k_physics.cu -- momentum (wind, jbar, barotropic solve, island, velc), EMBM (tstipa),
// surflux and sea-ice kernels, sm_100a.  Compiled with -fmad=false: every expression keeps the
// reference's operation order, so all non-transcendental results are bit-identical to the CPU
// oracle; exp/log/pow differ from glibc by <= 2 ulp.
//
// Reference: src/goldstein/goldstein.f90 (step_goldstein :98-233, jbar :3217-3315,
// ubarsolv :3500-3565, velc :3568-3679, wind :3685-3714, get_hosing :3086-3129),
// src/goldstein/goldstein_lib.f90 (island :186-241), src/embm/embm.f90 (step_embm :22-195,
// tstipa :2039-2138, surflux :2548-3738), src/goldsteinseaice/gold_seaice.f90
// (step_seaice :511-733, tstepsic :844-929).
#include "cg_device.cuh"
#include "cg_host.hpp"

namespace cg {

static __constant__ GridC c_g;

void upload_grid_physics(const GridC &g, cudaStream_t s) { cudaMemcpyToSymbolAsync(c_g, &g, sizeof(GridC), 0, cudaMemcpyHostToDevice, s); }

#define A2I(i, j) (cell2(I, (i), (j)) * MS + m)
#define A3I(l, i, j) (((size_t)((l)-1) * I * J + cell2(I, (i), (j))) * MS + m)
#define RHX(l, i, j) v.rh[((l)-1) + 3 * ((i) + (I + 2) * (j))]
#define DRAGX(l, i, j) v.drag[((size_t)((l)-1) + 2 * (((i)-1) + (I + 1) * ((j)-1))) * MS + m]
#define UBX(l, i, j) v.ub[((size_t)((l)-1) + 2 * ((i) + (I + 2) * (j))) * MS + m]
#define PSIX(i, j) v.psi[((size_t)(i) + (I + 1) * (j)) * MS + m]
#define UX(c, i, j, k) v.u[(cell3(I, J, (i), (j), (k)) * 3 + ((c)-1)) * MS + m]
#define U1X(c, i, j, k) v.u1[(cell3(I, J, (i), (j), (k)) * 2 + ((c)-1)) * MS + m]
#define RHOX(i, j, k) v.rho[cell3(I, J, (i), (j), (k)) * MS + m]
#define BPX(i, j, k) v.bp[cell3(I, J, (i), (j), (k)) * MS + m]
#define SBPX(i, j) v.sbp[A2I(i, j)]
#define KUX(l, i, j) ((int)v.ku[((l)-1) + 2 * (((i)-1) + I * ((j)-1))])
#define MKX(i, j) ((int)v.mk[((i)-1) + (I + 1) * ((j)-1)])
#define DIMS const int I = v.I, J = v.J, K = v.K, MS = v.MS; (void)K;

// ---------------------------------------------------------------- step bookkeeping
// istep_ocn = istep_ocn + 1 (genie.f90:271-277) and the hosing increment (goldstein.f90:3101)
__global__ void k_step_begin(const Dev v) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m == 0) *v.istep_ocn = *v.istep_ocn + 1;
}
// surface ocean velocities as the sea-ice step sees them (ustar_ocn/vstar_ocn, goldstein.f90:428-429): a snapshot, so
// that the momentum step of the same cycle may overwrite u while surflux / EMBM / sea ice are still running
__global__ void __launch_bounds__(128) k_usnap(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || c2 >= I * J) return;
  const size_t o = ((size_t)(K - 1) * I * J + c2) * 3 * MS + m;
  v.usnap[(size_t)c2 * MS + m] = v.u[o];
  v.usnap[((size_t)I * J + c2) * MS + m] = v.u[o + MS];
}
__global__ void k_hosing(const Dev v) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < v.M) v.hosing[m] = v.hosing[m] + v.p.hosing_trend[m] * kTsc * c_g.dt;
}

// ---------------------------------------------------------------- surflux, part 1 (per cell)
__device__ inline double ch4_func(double ch4, double n2o) {
  return 0.47 * log(1.0 + 2.01e-5 * pow(ch4 * n2o, 0.75) + 5.31e-15 * ch4 * pow(ch4 * n2o, 1.52));
}

// global-mean air temperature, SUM(atemp)/REAL(maxj*maxi) in array order (embm.f90:2950)
__global__ void k_meantemp(const Dev v, double *meantemp) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= v.M) return;
  double s = 0.0;
  for (int q = 0; q < I * J; q++) s = s + v.tq[(size_t)q * MS + m];
  meantemp[m] = s / (double)(J * I);
}

__global__ void __launch_bounds__(128) k_surflux1(const Dev v, const double *meantemp_arr) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const size_t q = (size_t)c2 * MS + m, q2 = ((size_t)I * J + c2) * MS + m;  // field 1 / field 2 of (2,i,j)
  const int istot = *v.istep_ocn;
  const int nsol = (istot - 1) % c_g.nyear + 1;
  const double solf = v.solfor[(j - 1) + (size_t)J * (nsol - 1)];
  const double zeroc = 273.15, tol = 1.0e-10;
  const double rmax = v.p.rmax[m], rdtdim = v.p.rdtdim[m], hat2 = v.p.hatmbl2[m];
  // greenhouse gases, compound increase (embm.f90:2913-2919)
  const double co2 = (1.0 + v.p.rate_co2[m]) * v.co2[q], ch4 = (1.0 + v.p.rate_ch4[m]) * v.ch4[q],
               n2o = (1.0 + v.p.rate_n2o[m]) * v.n2o[q];
  v.co2[q] = co2; v.ch4[q] = ch4; v.n2o[q] = n2o;
  const double at = v.tq[q];
  double ashum = v.tq[q2];
  const double qsata = kConst1 * exp(kConst4 * at / (at + kConst5));
  const double pptn = fmax(0.0, (ashum - rmax * qsata) * kRhoao * hat2 * rdtdim);
  ashum = fmin(ashum, rmax * qsata);
  v.tq1[q2] = ashum;
  v.tq[q2] = ashum;
  v.pptn[q] = pptn;
  const double rq = ashum / qsata;
  const double albcl = v.albcl[q];
  const double fxsw = solf * (1.0 - albcl);
  const double tv0 = 2.43414e2 + rq * (-3.47968e1 + 1.02790e1 * rq);
  const double tv1 = 2.60065 + rq * (-1.62064 + 6.34856e-1 * rq);
  const double tv2 = 4.40272e-3 + rq * (-2.26092e-2 + 1.12265e-2 * rq);
  const double tv3 = -2.05237e-5 + rq * (-9.67000e-5 + 5.62925e-5 * rq);
  const double ch4_term = kAlphaCh4 * (sqrt(1.0e9 * ch4) - sqrt(1.0e9 * kCh40)) - ch4_func(1.0e9 * ch4, 1.0e9 * kN2o0) +
                          ch4_func(1.0e9 * kCh40, 1.0e9 * kN2o0);
  const double n2o_term = kAlphaN2o * (sqrt(1.0e9 * n2o) - sqrt(1.0e9 * kN2o0)) - ch4_func(1.0e9 * kCh40, 1.0e9 * n2o) +
                          ch4_func(1.0e9 * kCh40, 1.0e9 * kN2o0);
  const double meantemp = meantemp_arr ? meantemp_arr[m] : 0.0;
  const double fxplw = tv0 + at * (tv1 + at * (tv2 + at * tv3)) - v.p.delf2x[m] * log(co2 / kCo20) - ch4_term - n2o_term +
                       v.p.olr_adj[m] * (meantemp - v.p.t_eqm[m]) - v.p.olr_adj0[m];
  const double fxlata = kRho0 * pptn * kHlv;
  v.fxsw[q] = fxsw;
  v.fxplw[q] = fxplw;
  v.albedo[q] = albcl;
  const int k1c = CG_K1(v, i, j);
  if (k1c <= K) {
    const size_t sC = (size_t)v.L * MS;
    const size_t ot = cell3(I, J, i, j, K) * sC + m;
    const double otemp = v.sst ? v.sst[q] : v.ts_cur[ot], osaln = v.sst ? v.sst[(size_t)I * J * MS + q] : v.ts_cur[ot + MS];
    const double us = v.usurf[q], cca = v.ca[q], sa = v.varice1[q2], sich = v.varice1[q];
    double alw = at + zeroc;
    alw = alw * alw;
    alw = alw * alw;
    alw = kEma * alw;
    const double salt = v.p.saln0[m] + osaln;
    const double tsfreez = salt * (-0.0575 + 0.0017 * sqrt(salt) - 0.0002 * salt);
    const double qb = v.p.rsictscsf[m] * (tsfreez - otemp);
    double qbsic = qb;
    double albsic, fx0sica, dhsic, evapsic, tice, atm_latenti, atm_sensiblei, atm_netsoli, atm_netlongi;
    if (sa > 0.0) {
      albsic = fmax(v.p.par_albsic_min[m], fmin(v.p.par_albsic_max[m], 0.40 - 0.04 * at));
      const double fxswsic = solf * (1.0 - albsic);
      tice = v.tice[q];
      double ticold, cesic, chsic, cfxsensic, qsatsic, tieqn, dtieq;
      for (int iter = 1; iter <= 21; iter++) {
        ticold = tice;
        cesic = 1.0e-3 * (1.0022 - 0.0822 * (at - ticold) + 0.0266 * us);
        cesic = fmax(6.0e-5, fmin(2.19e-3, cesic));
        chsic = 0.94 * cesic;
        cfxsensic = kRhoair * chsic * kCpa * us;
        qsatsic = kConst1 * exp(kConst2 * ticold / (ticold + kConst3));
        evapsic = fmax(0.0, (qsatsic - ashum) * kRhoao * cesic * us);
        const double tz = ticold + zeroc, tc3 = ticold + kConst3;
        tieqn = sich * ((1 - cca) * fxswsic + alw - kEmo * ((tz * tz) * (tz * tz)) - cfxsensic * (ticold - at) -
                        kRho0 * kHls * evapsic) +
                kConsic * (tsfreez - ticold);
        dtieq = sich * (-4.0 * kEmo * (tz * tz * tz) - cfxsensic -
                        kHls * kRhoair * cesic * us * qsatsic * kConst2 * kConst3 / (tc3 * tc3) * 0.5 *
                            (1.0 + copysign(1.0, qsatsic - ashum))) -
                kConsic;
        tice = ticold - tieqn / dtieq;
        if (fabs(tice - ticold) < tol || (ticold > kTfreez && tieqn > 0.0)) break;
      }
      tice = fmin(kTfreez, tice);
      const double tz = tice + zeroc;
      const double fxlwsic = kEmo * ((tz * tz) * (tz * tz)) - alw;
      cesic = 1.0e-3 * (1.0022 - 0.0822 * (at - tice) + 0.0266 * us);
      cesic = fmax(6.0e-5, fmin(2.19e-3, cesic));
      chsic = 0.94 * cesic;
      cfxsensic = kRhoair * chsic * kCpa * us;
      const double fxsensic = cfxsensic * (tice - at);
      qsatsic = kConst1 * exp(kConst2 * tice / (tice + kConst3));
      evapsic = fmax(0.0, (qsatsic - ashum) * kRhoao * cesic * us);
      const double fx0sic = (1 - cca) * fxswsic - fxsensic - fxlwsic - kRho0 * kHls * evapsic;
      fx0sica = cca * fxswsic + fxlata + fxsensic + fxlwsic - fxplw;
      atm_latenti = +fxlata;
      atm_sensiblei = +fxsensic;
      atm_netsoli = +cca * fxswsic;
      atm_netlongi = +fxlwsic - fxplw;
      dhsic = kRrholf * (qb - fx0sic) - kRhooi * evapsic;
      if (sich >= v.p.par_sich_max[m]) {
        if (dhsic > 0.0) {
          qbsic = (0.0 + kRhooi * evapsic) / kRrholf + fx0sic;
          dhsic = kRrholf * (qbsic - fx0sic) - kRhooi * evapsic;
        }
      }
    } else {
      albsic = 0.0; fx0sica = 0.0; dhsic = 0.0; evapsic = 0.0; tice = 0.0;
      atm_latenti = 0.0; atm_sensiblei = 0.0; atm_netsoli = 0.0; atm_netlongi = 0.0;
    }
    const double tzo = otemp + zeroc;
    const double fxlw = kEmo * ((tzo * tzo) * (tzo * tzo)) - alw;
    double ce = 1.0e-3 * (1.0022 - 0.0822 * (at - otemp) + 0.0266 * us);
    ce = fmax(6.0e-5, fmin(2.19e-3, ce));
    const double ch = 0.94 * ce;
    const double fxsen = kRhoair * ch * kCpa * us * (otemp - at);
    const double qsato = kConst1 * exp(kConst4 * otemp / (otemp + kConst5));
    const double evap = fmax(0.0, (qsato - ashum) * kRhoao * ce * us);
    const double fx0oa = cca * fxsw + fxlata + fxsen + fxlw - fxplw;
    const double atm_latent = +fxlata, atm_sensible = +fxsen, atm_netsol = +cca * fxsw, atm_netlong = +fxlw - fxplw;
    v.fx0a[q] = (1 - sa) * fx0oa + sa * fx0sica;
    v.latent_atm[q] = (sa * atm_latenti) + ((1 - sa) * atm_latent);
    v.sensible_atm[q] = (sa * atm_sensiblei) + ((1 - sa) * atm_sensible);
    v.netsolar_atm[q] = (sa * atm_netsoli) + ((1 - sa) * atm_netsol);
    v.netlong_atm[q] = (sa * atm_netlongi) + ((1 - sa) * atm_netlong);
    const double fx0o = (1 - cca) * fxsw - fxsen - fxlw - kRho0 * kHlv * evap;
    v.fx0o[q] = fx0o;
    v.latent_ocn[q] = (1 - sa) * (-kRho0 * kHlv * evap + fmax(0.0, qb - fx0o)) + sa * qbsic;
    v.sensible_ocn[q] = -((1 - sa) * fxsen);
    v.netsolar_ocn[q] = (1 - sa) * (1 - cca) * fxsw;
    v.netlong_ocn[q] = -((1 - sa) * fxlw);
    const double dho = fmax(0.0, kRrholf * (qb - fx0o));
    v.dhght_sic[q] = sa * dhsic + (1 - sa) * dho;
    double dta = fmax(0.0, kRhmin * dho * (1 - sa));
    if (sich > 1.0e-12) dta = dta + fmin(0.0, 0.5 * sa * sa * dhsic / sich);
    v.dfrac_sic[q] = dta;
    v.albedo[q] = sa * albsic + (1 - sa) * albcl;
    v.albice[q] = albsic;
    v.tice[q] = tice;
    v.evap[q] = evap;
    v.evapsic[q] = evapsic;
    v.fxsen[q] = fxsen;
    v.fxlw[q] = fxlw;
    v.runoff_land[q] = 0.0;
  } else {
    v.fx0a[q] = fxsw + fxlata - fxplw;
    v.latent_atm[q] = +fxlata;
    v.sensible_atm[q] = +0.0;
    v.netsolar_atm[q] = +fxsw;
    v.netlong_atm[q] = -fxplw;
    v.evap[q] = 0.0;
    v.latent_ocn[q] = 0.0; v.sensible_ocn[q] = 0.0; v.netsolar_ocn[q] = 0.0; v.netlong_ocn[q] = 0.0;
    v.dhght_sic[q] = 0.0; v.dfrac_sic[q] = 0.0; v.albice[q] = 0.0;
    v.runoff_land[q] = pptn * kM2mm;
  }
}

// surflux part 2: runoff routing as an order-preserving gather + final flux scaling (embm.f90:3351-3357, 3657-3695)
__global__ void __launch_bounds__(128) k_surflux2(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const size_t q = (size_t)c2 * MS + m, q2 = ((size_t)I * J + c2) * MS + m;
  const double pptn = v.pptn[q];
  double pptn_ocn, runoff_ocn, evap_atm;
  if (CG_K1(v, i, j) <= K) {
    double runoff = 0.0;
    for (int p = v.iroff_ptr[c2]; p < v.iroff_ptr[c2 + 1]; p++) runoff = runoff + v.pptn[(size_t)v.iroff_src[p] * MS + m];
    const double sa = v.varice1[q2];
    pptn_ocn = pptn;
    runoff_ocn = runoff + 0.0;
    evap_atm = v.evap[q] * (1 - sa) + v.evapsic[q] * sa;
    pptn_ocn = pptn_ocn + v.pmeadj[q];
  } else {
    pptn_ocn = 0.0; runoff_ocn = 0.0; evap_atm = 0.0;
  }
  const double evap_ocn = -evap_atm;
  v.precip_atm[q] = pptn * kM2mm;
  v.precip_ocn[q] = pptn_ocn * kM2mm;
  v.evap_ocn[q] = evap_ocn * kM2mm;
  v.runoff_ocn[q] = runoff_ocn * kM2mm;
  v.evap_atm[q] = evap_atm * kM2mm;
}

// ---------------------------------------------------------------- EMBM: tstipa, nsteps fused
// One block = one member; thread owns CPT cells; tq2 lives in shared memory (halo rows 0 and J+1
// are zero, the i-halo is index arithmetic).  embm.f90:2039-2138 + step_embm :48-70.
// EXACT: blockDim.x * CPT == I * J, every thread owns CPT cells: no guards in the iteration loops (36 x 36: 648 threads x 2)
template <int CPT, bool EXACT, bool SMC = false>
__device__ __forceinline__ void embm_body(const Dev &v, const int nsteps);
template <int CPT>
__global__ void __launch_bounds__(CPT == 3 ? 448 : 704) k_embm(const Dev v, const int nsteps) { embm_body<CPT, false>(v, nsteps); }
// SMC: the four face coefficients of a cell (constant over the launch) live in shared memory, [field][coefficient][cell], instead
// of registers: at 672 threads a thread has 80 registers and the 36 doubles of per-cell state spilled to local memory (111 of the
// loop's 371 instructions were LDL / STL)
__global__ void __launch_bounds__(704) k_embm_s2(const Dev v, const int nsteps) { embm_body<2, false, true>(v, nsteps); }
// exact forms for 36 x 36: 648 threads x 2 cells (21 warps: a sub-partition holds 6 of them, 16384 / (6 x 32) = 85 -> 80 registers,
// which is why __launch_bounds__ stops there and a larger __maxnreg__ cannot launch) and 432 threads x 3 cells (14 warps, 128 registers)
__global__ void __launch_bounds__(648) k_embm_x2(const Dev v, const int nsteps) { embm_body<2, true>(v, nsteps); }
__global__ void __launch_bounds__(432) k_embm_x3(const Dev v, const int nsteps) { embm_body<3, true>(v, nsteps); }
template <int CPT, bool EXACT, bool SMC>
__device__ __forceinline__ void embm_body(const Dev &v, const int nsteps) {
  DIMS
  extern __shared__ double tq2[];  // (I, 0:J+1) x 2 fields [+ SMC: 8 x I*J coefficients]
  const int m = blockIdx.x;
  const int nc = I * J, tc = blockDim.x;
  const double cimp = 0.5;
  const double dtloc = v.p.dtatm[m], rfluxsca = v.p.rfluxsca[m], rpmesca = v.p.rpmesca[m];
  double tq[2][CPT], tq1[2][CPT], tqa[2][CPT];
  double cie[2][CPT], ciwm[2][CPT], cin[2][CPT], cism[2][CPT], cdiv[2][CPT];
  double *const cs = tq2 + 2 * (I * (J + 2));
#define CSM(q, l, n) cs[((l) * 4 + (q)) * nc + threadIdx.x + (n) * tc]
#define T2(l, i, j) tq2[(l) * (I * (J + 2)) + ((i)-1) + I * (j)]
  for (int q = threadIdx.x; q < I; q += tc) { T2(0, q + 1, 0) = 0.0; T2(0, q + 1, J + 1) = 0.0; T2(1, q + 1, 0) = 0.0; T2(1, q + 1, J + 1) = 0.0; }
#pragma unroll
  for (int n = 0; n < CPT; n++) {
    const int c2 = threadIdx.x + n * tc;
    if (c2 < nc) {
      const int i = c2 % I + 1, j = c2 / I + 1, im = (i > 1) ? i - 1 : I;
      const size_t q = (size_t)c2 * MS + m, q2 = ((size_t)nc + c2) * MS + m;
      tq[0][n] = v.tq[q]; tq[1][n] = v.tq[q2];
      tq1[0][n] = v.tq1[q]; tq1[1][n] = v.tq1[q2];
      tqa[0][n] = (v.netsolar_atm[q] + v.latent_atm[q] + v.sensible_atm[q] + v.netlong_atm[q]) * rfluxsca;
      tqa[1][n] = v.evap_atm[q] * kMm2m * rpmesca;
      const double ua = v.uatm_u[q], uaw = v.uatm_u[A2I(im, j)], va = v.uatm_v[q], vas = (j > 1) ? v.uatm_v[A2I(i, j - 1)] : 0.0;
#pragma unroll
      for (int l = 0; l < 2; l++) {
        const double bz = l ? v.p.betaz2[m] : v.p.betaz1[m], bm = l ? v.p.betam2[m] : v.p.betam1[m];
        const double da1 = v.p.diffa[((size_t)(j - 1) * 4 + l) * MS + m], da2 = v.p.diffa[((size_t)(j - 1) * 4 + l + 2) * MS + m];
        // east face of (i,j) and of (i-1,j)
        double e = bz * ua * c_g.rc[j] * 0.5 * c_g.rdphi;
        double tv = c_g.rc[j] * c_g.rc[j] * c_g.rdphi * da1 * c_g.rdphi;
        double pec = bz * ua * c_g.dphi / da1;
        double ups = pec / (2.0 + fabs(pec));
        const double ciw_c = e * (1 + ups) + tv;
        const double cie_c = e * (1 - ups) - tv;
        e = bz * uaw * c_g.rc[j] * 0.5 * c_g.rdphi;
        pec = bz * uaw * c_g.dphi / da1;
        ups = pec / (2.0 + fabs(pec));
        const double ciw_w = e * (1 + ups) + tv;
        const double cie_w = e * (1 - ups) - tv;
        // north face of (i,j) and of (i,j-1)
        double nn = c_g.cv[j] * bm * va * 0.5;
        if (j < J) {
          tv = c_g.cv[j] * c_g.cv[j] * c_g.rdsv[j] * da2;
          pec = bm * va * c_g.dsv[j] / da2;
          ups = pec / (2.0 + fabs(pec));
        } else {
          tv = 0.0;
          ups = 0.0;
        }
        const double cis_c = nn * (1 + ups) + tv;
        const double cin_c = nn * (1 - ups) - tv;
        double cis_s = 0.0, cin_s = 0.0;
        if (j > 1) {
          const double das = v.p.diffa[((size_t)(j - 2) * 4 + l + 2) * MS + m];
          nn = c_g.cv[j - 1] * bm * vas * 0.5;
          tv = c_g.cv[j - 1] * c_g.cv[j - 1] * c_g.rdsv[j - 1] * das;
          pec = bm * vas * c_g.dsv[j - 1] / das;
          ups = pec / (2.0 + fabs(pec));
          cis_s = nn * (1 + ups) + tv;
          cin_s = nn * (1 - ups) - tv;
        }
        if (SMC) { CSM(0, l, n) = cie_c; CSM(1, l, n) = ciw_w; CSM(2, l, n) = cin_c; CSM(3, l, n) = cis_s; }
        else { cie[l][n] = cie_c; ciwm[l][n] = ciw_w; cin[l][n] = cin_c; cism[l][n] = cis_s; }
        cdiv[l][n] = ciw_c - cie_w + (cis_c - cin_s) * c_g.rds[j];
      }
    }
  }
  // the implicit iterations divide by 1 + cimp * dt * cdiv, the same number for a cell in all 4 iterations of all nsteps steps:
  // one reciprocal, then q = a y, r = a - b q (exact in one FMA), q + r y = RN(a / b) -- the correctly rounded quotient the
  // division sequence returns (Markstein 1990), bit for bit the reference's result at 3 instructions instead of ~28
  double rden[2][CPT];
#pragma unroll
  for (int l = 0; l < 2; l++)
#pragma unroll
    for (int n = 0; n < CPT; n++) rden[l][n] = 1.0 / (1 + cimp * (dtloc * cdiv[l][n]));
  // The two fields (l = 0: temperature, l = 1: humidity) do not see each other inside tstipa, so their five iterations run side
  // by side on two planes of shared memory: half the block barriers of the field-after-field loop and two independent chains
  // per cell and thread.  Per field the operations and their order are the reference's.  The cell's five shared-memory offsets
  // are taken once (the index arithmetic was most of the loop's instructions).
  int oc[CPT], oe[CPT], ow[CPT], on[CPT], os[CPT];
  double rdsj[CPT];
  bool ok[CPT];
  const int plane = I * (J + 2);
#pragma unroll
  for (int n = 0; n < CPT; n++) {
    const int c2 = threadIdx.x + n * tc;
    ok[n] = EXACT || c2 < nc;
    const int cc = ok[n] ? c2 : 0;
    const int i = cc % I + 1, j = cc / I + 1, ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
    oc[n] = (i - 1) + I * j; oe[n] = (ip - 1) + I * j; ow[n] = (im - 1) + I * j; on[n] = (i - 1) + I * (j + 1); os[n] = (i - 1) + I * (j - 1);
    rdsj[n] = c_g.rds[j];
  }
  for (int step = 0; step < nsteps; step++) {
    // four implicit iterations ...
    for (int iits = 0; iits < 4; iits++) {
      __syncthreads();
#pragma unroll
      for (int l = 0; l < 2; l++) {
#pragma unroll
        for (int n = 0; n < CPT; n++)
          if (EXACT || ok[n]) tq2[l * plane + oc[n]] = cimp * tq[l][n] + (1.0 - cimp) * tq1[l][n];
      }
      __syncthreads();
#pragma unroll
      for (int l = 0; l < 2; l++) {
#pragma unroll
        for (int n = 0; n < CPT; n++) {
          if (EXACT || ok[n]) {
            const double *t2 = tq2 + l * plane;
            const double kE = SMC ? CSM(0, l, n) : cie[l][n], kW = SMC ? CSM(1, l, n) : ciwm[l][n], kN = SMC ? CSM(2, l, n) : cin[l][n],
                         kS = SMC ? CSM(3, l, n) : cism[l][n];
            const double flx = -tqa[l][n] + kE * t2[oe[n]] - kW * t2[ow[n]] + (kN * t2[on[n]] - kS * t2[os[n]]) * rdsj[n];
            const double centre = dtloc * cdiv[l][n];
            const double num = tq1[l][n] * (1.0 - (1.0 - cimp) * centre) - dtloc * flx, den = 1 + cimp * centre;
            const double qq = num * rden[l][n];
            tq[l][n] = fma(fma(-den, qq, num), rden[l][n], qq);
          }
        }
      }
    }
    // ... and the explicit corrector on the average of the last two iterates (iits = 5 of the reference's loop)
    __syncthreads();
#pragma unroll
    for (int l = 0; l < 2; l++) {
#pragma unroll
      for (int n = 0; n < CPT; n++)
        if (EXACT || ok[n]) {
          double *t2 = tq2 + l * plane + oc[n];
          *t2 = 0.5 * (*t2 + cimp * tq[l][n] + (1.0 - cimp) * tq1[l][n]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < 2; l++) {
#pragma unroll
      for (int n = 0; n < CPT; n++) {
        if (EXACT || ok[n]) {
          const double *t2 = tq2 + l * plane;
          const double kE = SMC ? CSM(0, l, n) : cie[l][n], kW = SMC ? CSM(1, l, n) : ciwm[l][n], kN = SMC ? CSM(2, l, n) : cin[l][n],
                       kS = SMC ? CSM(3, l, n) : cism[l][n];
          const double flx = -tqa[l][n] + kE * t2[oe[n]] - kW * t2[ow[n]] + (kN * t2[on[n]] - kS * t2[os[n]]) * rdsj[n];
          tq[l][n] = tq1[l][n] - dtloc * flx - dtloc * t2[oc[n]] * cdiv[l][n];
        }
      }
    }
#pragma unroll
    for (int n = 0; n < CPT; n++) { tq1[0][n] = tq[0][n]; tq1[1][n] = tq[1][n]; }
  }
#pragma unroll
  for (int n = 0; n < CPT; n++) {
    const int c2 = threadIdx.x + n * tc;
    if (c2 < nc) {
      const size_t q = (size_t)c2 * MS + m, q2 = ((size_t)nc + c2) * MS + m;
      v.tq[q] = tq[0][n]; v.tq[q2] = tq[1][n];
      v.tq1[q] = tq1[0][n]; v.tq1[q2] = tq1[1][n];
      v.tqa[q] = tqa[0][n]; v.tqa[q2] = tqa[1][n];
    }
  }
#undef T2
}

// ---------------------------------------------------------------- sea ice
// tstepsic (gold_seaice.f90:844-929): varice1 -> varice
__global__ void __launch_bounds__(128) k_seaice1(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  if (K < k1c) return;
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CG_K1(v, ip, j), k1w = CG_K1(v, im, j), k1n = CG_K1(v, i, j + 1), k1s = CG_K1(v, i, j - 1);
  const double sath = v.p.par_sica_thresh[m], shth = v.p.par_sich_thresh[m], diffsic = v.p.diffsic[m], dtsic = v.p.dtsic[m];
  const double rc = c_g.rc[j], rdphi = c_g.rdphi;
  // surface ocean velocities (ustar_ocn = u(1,:,:,maxk), goldstein.f90:428-429)
#define USN(c, ii, jj) v.usnap[((size_t)((c)-1) * I * J + cell2(I, (ii), (jj))) * MS + m]
  const double uE = USN(1, i, j), uW = USN(1, im, j), vN = USN(2, i, j), vS = (j > 1) ? USN(2, i, j - 1) : 0.0;
#undef USN
  const size_t nf = (size_t)I * J * MS;
  const size_t qc = A2I(i, j), qe = A2I(ip, j), qw = A2I(im, j), qn = (j < J) ? A2I(i, j + 1) : qc, qs = (j > 1) ? A2I(i, j - 1) : qc;
  const double hc = v.varice1[qc], ac = v.varice1[qc + nf];
#pragma unroll
  for (int l = 0; l < 2; l++) {
    const size_t lo = l * nf;
    const double c0 = v.varice1[qc + lo];
    double fe = 0.0, fw = 0.0, fn = 0.0, fs = 0.0;
    if (K >= max(k1c, k1e)) {
      const double e0 = v.varice1[qe + lo], he = v.varice1[qe], ae = v.varice1[qe + nf];
      fe = uE * rc * (e0 + c0) * 0.5;
      if (uE >= 0.0) { if (ae > sath) fe = 0; if (he > shth) fe = 0; }
      else { if (ac > sath) fe = 0; if (hc > shth) fe = 0; }
      fe = fe - (e0 - c0) * rc * rc * rdphi * diffsic;
    }
    if (K >= max(k1c, k1w)) {
      const double w0 = v.varice1[qw + lo], hw = v.varice1[qw], aw = v.varice1[qw + nf];
      fw = uW * rc * (c0 + w0) * 0.5;
      if (uW >= 0.0) { if (ac > sath) fw = 0; if (hc > shth) fw = 0; }
      else { if (aw > sath) fw = 0; if (hw > shth) fw = 0; }
      fw = fw - (c0 - w0) * rc * rc * rdphi * diffsic;
    }
    if (K >= max(k1c, k1n)) {
      const double n0 = v.varice1[qn + lo], hn = v.varice1[qn], an = v.varice1[qn + nf];
      fn = c_g.cv[j] * vN * (n0 + c0) * 0.5;
      if (vN >= 0.0) { if (an > sath) fn = 0; if (hn > shth) fn = 0; }
      else { if (ac > sath) fn = 0; if (hc > shth) fn = 0; }
      fn = fn - c_g.cv[j] * c_g.cv[j] * (n0 - c0) * c_g.rdsv[j] * diffsic;
    }
    if (j > 1 && K >= max(k1c, k1s)) {
      const double s0 = v.varice1[qs + lo], hs = v.varice1[qs], as = v.varice1[qs + nf];
      fs = c_g.cv[j - 1] * vS * (c0 + s0) * 0.5;
      if (vS >= 0.0) { if (ac > sath) fs = 0; if (hc > shth) fs = 0; }
      else { if (as > sath) fs = 0; if (hs > shth) fs = 0; }
      fs = fs - c_g.cv[j - 1] * c_g.cv[j - 1] * (c0 - s0) * c_g.rdsv[j - 1] * diffsic;
    }
    const double dtha = l ? v.dfrac_sic[qc] : v.dhght_sic[qc];
    v.varice[qc + lo] = c0 - dtsic * ((fe - fw) * rdphi + (fn - fs) * c_g.rds[j]) + kTsc * dtsic * dtha;
  }
}
// step_seaice post-processing (gold_seaice.f90:605-622, 691-702)
__global__ void __launch_bounds__(128) k_seaice2(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const size_t q = (size_t)c2 * MS + m, q2 = ((size_t)I * J + c2) * MS + m;
  double fw_delta = 0.0, fx_delta = 0.0;
  if (K >= CG_K1(v, i, j)) {
    const double rdt = v.p.sic_rdtdim[m];
    double h = v.varice[q], a = v.varice[q2];
    fw_delta = -kRhoio * v.dhght_sic[q];
    a = fmax(0.0, fmin(1.0, a));
    if (h < kHmin) {
      fx_delta = -h * kRhoice * kHlf * rdt;
      fw_delta = fw_delta + h * kRhoio * rdt;
      h = 0.0;
      a = 0.0;
    }
    v.varice[q] = h; v.varice[q2] = a;
    v.varice1[q] = h; v.varice1[q2] = a;
    fw_delta = fw_delta * kM2mm;
  }
  v.waterflux_ocn[q] = fw_delta;
  v.conductflux_ocn[q] = fx_delta;
}

// ---------------------------------------------------------------- ocean: surface b.c.
// net heat / freshwater flux into ts(1:2,:,:,maxk+1) (goldstein.f90:143-170, get_hosing :3114-3127)
__global__ void __launch_bounds__(128) k_gold_pre(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const size_t q = (size_t)c2 * MS + m;
  const int istep = *v.istep_ocn;
  const double fw_hosing = (istep <= v.p.nsteps_hosing[m]) ? kM2mm * v.hosing[m] * v.rhosing[c2] : 0.0;
  const double fw_anom = 0.0;
  const double fx0neto = v.netsolar_ocn[q] + v.sensible_ocn[q] + v.netlong_ocn[q] + v.latent_ocn[q] + v.conductflux_ocn[q];
  double fwfxneto = v.precip_ocn[q] + v.evap_ocn[q] + v.runoff_ocn[q] + v.waterflux_ocn[q] + fw_hosing + fw_anom;
  fwfxneto = fwfxneto * kMm2m;
  v.tsflux[q] = -fx0neto * kRfluxsc;
  v.tsflux[(size_t)I * J * MS + q] = fwfxneto * v.p.rpmesco[m];
}

// ---------------------------------------------------------------- ocean: momentum
// bottom pressure integrals (jbar, goldstein.f90:3226-3247)
__global__ void __launch_bounds__(128) k_bp(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  if (k1c > K) return;
  const int mk = MKX(i, j);
  double bp = BPX(i, j, k1c), rl = RHOX(i, j, k1c), sbp = 0.0;
  for (int k = k1c + 1; k <= K; k++) {
    const double r = RHOX(i, j, k);
    bp = bp - (r + rl) * c_g.dza[k - 1] * 0.5;
    BPX(i, j, k) = bp;
    rl = r;
    if (mk > 0 && k >= mk + 1) sbp = sbp + bp * c_g.dz[k];
  }
  if (mk > 0) SBPX(i, j) = sbp;
}

// wind stress curl + JEBAR source of the barotropic streamfunction (wind :3695-3713, jbar :3265-3314)
__global__ void __launch_bounds__(128) k_gb(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y * blockDim.y + threadIdx.y;  // 0-based psi point
  if (m >= v.M || p >= v.nm) return;
  const int i = p % I + 1, j = p / I;  // j = 0..J
  const int ip1 = i % I + 1;
  double gb = 0.0;
  if (max(max(CG_K1(v, i, j), CG_K1(v, i + 1, j)), max(CG_K1(v, i, j + 1), CG_K1(v, i + 1, j + 1))) <= K) {
    const size_t nf = (size_t)I * J * MS;
    gb = (v.tau[A2I(ip1, j) + nf] * RHX(2, i + 1, j) - v.tau[A2I(i, j) + nf] * RHX(2, i, j)) * c_g.rdphi * c_g.rcv[j] -
         (v.tau[A2I(i, j + 1)] * c_g.c[j + 1] * RHX(1, i, j + 1) - v.tau[A2I(i, j)] * c_g.c[j] * RHX(1, i, j)) * c_g.rdsv[j];
  }
  if (j >= 1 && j <= J - 1 && v.getj[(i - 1) + I * (j - 1)]) {
    const double gbold = gb;
    double tv1 = 0, tv2 = 0, tv3 = 0, tv4 = 0;
    for (int k = KUX(2, ip1, j); k <= MKX(ip1, j + 1); k++) tv1 = tv1 + BPX(ip1, j + 1, k) * c_g.dz[k];
    for (int k = KUX(2, ip1, j); k <= MKX(ip1, j); k++) tv2 = tv2 + BPX(ip1, j, k) * c_g.dz[k];
    for (int k = KUX(2, i, j); k <= MKX(i, j + 1); k++) tv3 = tv3 + BPX(i, j + 1, k) * c_g.dz[k];
    for (int k = KUX(2, i, j); k <= MKX(i, j); k++) tv4 = tv4 + BPX(i, j, k) * c_g.dz[k];
    gb = gbold + ((tv3 + SBPX(i, j + 1) - tv4 - SBPX(i, j)) * RHX(2, i, j) -
                  (tv1 + SBPX(ip1, j + 1) - tv2 - SBPX(ip1, j)) * RHX(2, ip1, j)) *
                     c_g.rdphi * c_g.rdsv[j];
    tv1 = 0; tv2 = 0; tv3 = 0; tv4 = 0;
    for (int k = KUX(1, i, j + 1); k <= MKX(ip1, j + 1); k++) tv1 = tv1 + BPX(ip1, j + 1, k) * c_g.dz[k];
    for (int k = KUX(1, i, j); k <= MKX(ip1, j); k++) tv2 = tv2 + BPX(ip1, j, k) * c_g.dz[k];
    for (int k = KUX(1, i, j + 1); k <= MKX(i, j + 1); k++) tv3 = tv3 + BPX(i, j + 1, k) * c_g.dz[k];
    for (int k = KUX(1, i, j); k <= MKX(i, j); k++) tv4 = tv4 + BPX(i, j, k) * c_g.dz[k];
    gb = gb + ((tv1 + SBPX(ip1, j + 1) - tv3 - SBPX(i, j + 1)) * RHX(1, i, j + 1) -
               (tv2 + SBPX(ip1, j) - tv4 - SBPX(i, j)) * RHX(1, i, j)) *
                  c_g.rdphi * c_g.rdsv[j];
  }
  v.gb[(size_t)p * MS + m] = gb;
}

// banded LU solve of the streamfunction equation, reference operation order (ubarsolv :3511-3524).
// thread = member; gb is [point][m] so every access is coalesced across the warp, the factors of a
// shared factorisation are warp-uniform (broadcast) loads.
__global__ void __launch_bounds__(32) k_baro_strict(const Dev v) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= v.M) return;
  const int n = v.I, nm = v.nm, MS = v.MS;
  const int bw = n + 1, gw = 2 * n + 3;
  const double *__restrict__ ratm = v.ratm + (size_t)v.baro_group[m] * nm * bw;   // [row][t], t = j-i in 1..n+1
  const double *__restrict__ gap = v.gap + (size_t)v.baro_group[m] * nm * gw;     // [row][col]
  double *__restrict__ gb = v.gb + m;
#define GB(r) gb[(size_t)((r)-1) * MS]
  for (int i = 1; i <= nm - 1; i++) {
    const int im = min(i + n + 1, nm);
    const double gi = GB(i);
    for (int j = i + 1; j <= im; j++) GB(j) = GB(j) - ratm[(size_t)(j - 1) * bw + (j - i - 1)] * gi;
  }
  GB(nm) = GB(nm) / gap[(size_t)(nm - 1) * gw + (n + 1)];
  for (int i = nm - 1; i >= 1; i--) {
    const int km = min(n + 1, nm - i);
    double acc = GB(i);
    const double *g = gap + (size_t)(i - 1) * gw;
    for (int k = 1; k <= km; k++) acc = acc - g[n + 1 + k] * GB(i + k);
    GB(i) = acc / g[n + 1];
  }
#undef GB
}


// Fast barotropic solve: one warp per member, the right-hand side lives in shared memory and both
// sweeps are column-oriented (pivot value broadcast, 37 band updates in parallel across lanes).
// The forward sweep performs exactly the reference's operations (ubarsolv :3511-3516); the back
// substitution applies the same eliminations in column order with reciprocal pivots, so it differs
// from the reference's row-order subtraction by rounding only.  Factors are stored pivot-major:
//   bf[(i-1)*bw + t-1] = ratm(i+t, t)          bb[(i-1)*bw + t-1] = gap(i-t, n+2+t)     rd[i-1] = 1/gap(i, n+2)
constexpr int kBaroPF = 8;
__global__ void __launch_bounds__(128) k_baro_fast(const Dev v, const double *__restrict__ bf_all, const double *__restrict__ bb_all,
                                                   const double *__restrict__ rd_all) {
  extern __shared__ double xs[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + wib;
  if (m >= v.M) return;
  const int n = v.I, nm = v.nm, MS = v.MS, bw = n + 1;
  double *x = xs + (size_t)wib * nm;
  const size_t g = v.baro_group[m];
  const double *__restrict__ bf = bf_all + g * nm * bw, *__restrict__ bb = bb_all + g * nm * bw, *__restrict__ rd = rd_all + g * nm;
  for (int r = lane; r < nm; r += 32) x[r] = v.gb[(size_t)r * MS + m];
  __syncwarp();
  const bool two = lane + 33 <= bw;
  double fa[kBaroPF], fb[kBaroPF], na[kBaroPF], nb[kBaroPF];
  // ---- forward elimination
#pragma unroll
  for (int u = 0; u < kBaroPF; u++) {
    const int i = 1 + u;
    fa[u] = (i <= nm - 1) ? bf[(size_t)(i - 1) * bw + lane] : 0.0;
    fb[u] = (two && i <= nm - 1) ? bf[(size_t)(i - 1) * bw + lane + 32] : 0.0;
  }
  for (int ib = 1; ib <= nm - 1; ib += kBaroPF) {
#pragma unroll
    for (int u = 0; u < kBaroPF; u++) {
      const int i = ib + kBaroPF + u;
      na[u] = (i <= nm - 1) ? bf[(size_t)(i - 1) * bw + lane] : 0.0;
      nb[u] = (two && i <= nm - 1) ? bf[(size_t)(i - 1) * bw + lane + 32] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kBaroPF; u++) {
      const int i = ib + u;
      if (i <= nm - 1) {
        const double gi = x[i - 1];
        const int r1 = i + lane + 1, r2 = i + lane + 33;
        if (r1 <= nm) x[r1 - 1] = x[r1 - 1] - fa[u] * gi;
        if (two && r2 <= nm) x[r2 - 1] = x[r2 - 1] - fb[u] * gi;
      }
      __syncwarp();
    }
#pragma unroll
    for (int u = 0; u < kBaroPF; u++) { fa[u] = na[u]; fb[u] = nb[u]; }
  }
  // ---- back substitution, column oriented
#pragma unroll
  for (int u = 0; u < kBaroPF; u++) {
    const int i = nm - u;
    fa[u] = (i >= 1) ? bb[(size_t)(i - 1) * bw + lane] : 0.0;
    fb[u] = (two && i >= 1) ? bb[(size_t)(i - 1) * bw + lane + 32] : 0.0;
  }
  for (int ib = nm; ib >= 1; ib -= kBaroPF) {
#pragma unroll
    for (int u = 0; u < kBaroPF; u++) {
      const int i = ib - kBaroPF - u;
      na[u] = (i >= 1) ? bb[(size_t)(i - 1) * bw + lane] : 0.0;
      nb[u] = (two && i >= 1) ? bb[(size_t)(i - 1) * bw + lane + 32] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kBaroPF; u++) {
      const int i = ib - u;
      if (i >= 1) {
        const double xi = x[i - 1] * rd[i - 1];
        __syncwarp();  // every lane has read the pivot before lane 0 overwrites it
        const int r1 = i - lane - 1, r2 = i - lane - 33;
        if (r1 >= 1) x[r1 - 1] = x[r1 - 1] - fa[u] * xi;
        if (two && r2 >= 1) x[r2 - 1] = x[r2 - 1] - fb[u] * xi;
        if (lane == 0) x[i - 1] = xi;
      }
      __syncwarp();
    }
#pragma unroll
    for (int u = 0; u < kBaroPF; u++) { fa[u] = na[u]; fb[u] = nb[u]; }
  }
  for (int r = lane; r < nm; r += 32) v.gb[(size_t)r * MS + m] = x[r];
}

// ---------------------------------------------------------------------------------------------------------------
// Barotropic solve, register-resident variant (same arithmetic, operation for operation, as k_baro_fast: the results
// are bit-identical; only the data movement differs).  One warp = one member.  The banded substitutions are a chain of
// 2 x nm dependent pivots; k_baro_fast pays a shared-memory round trip (STS -> LDS) per pivot, here the active window
// of the vector lives in registers -- lane l holds the elements e with e mod 32 == l, two live blocks of 32 (the band
// is 32 < bw = I+1 <= 64 wide) -- and the pivot travels by one warp shuffle: the chain per pivot is
// SHFL -> DMUL -> DADD.  At pivot u of a chunk of 32, lane l updates its "current block" register with factor
// f[(l-u-1) mod 64] and its "next block" register with f[l-u+31] (zero outside the band).  To make both factor reads
// `lane base + immediate`, each pivot's factor row is expanded in shared memory to 96 entries,
//     row'[32+j] = f[j] (j < bw),   row'[j] = f[32+j] (j < bw-32),   zero elsewhere,
// so that the two reads are row'[l+31-u] and row'[l+63-u].  The packed pivot-major rows (bw doubles per pivot) stream
// from HBM through a 4-deep ring of 32-pivot chunks filled by cp.async.bulk (one elected lane, completion on an
// mbarrier, three chunks in flight); the expansion of chunk g+1 is interleaved with the pivots of chunk g.
// The backward sweep runs on the reversed index e' = nm-1-e so that both sweeps share one body.
constexpr int kBaroRing = 4, kBaroRow = 96;
__device__ __forceinline__ unsigned baro_sa(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
struct BaroChunk { int sw, c, row0, rows; };
__device__ __forceinline__ BaroChunk baro_chunk(const int g, const int nch, const int nm) {
  // global chunk g: sweep g / nch (0 forward, 1 backward), chunk c = g % nch covers pivots [32c, 32c+32) of that sweep
  BaroChunk k;
  k.sw = (g >= nch) ? 1 : 0;
  k.c = g - k.sw * nch;
  if (k.sw == 0) { k.row0 = 32 * k.c; k.rows = min(32, nm - k.row0); }
  else { const int hi = nm - 32 * k.c; k.row0 = max(hi - 32, 0); k.rows = hi - k.row0; }
  return k;
}
__device__ __forceinline__ void baro_issue(const int g, const int nch, const int nm, const int bw, const double *bf, const double *bb,
                                           double *stage, unsigned long long *bar) {
  const BaroChunk k = baro_chunk(g, nch, nm);
  const unsigned bytes = (unsigned)(k.rows * bw * 8);
  const unsigned b = baro_sa(bar + (g & (kBaroRing - 1)));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   baro_sa(stage + (size_t)(g & (kBaroRing - 1)) * 32 * bw)),
               "l"((k.sw ? bb : bf) + (size_t)k.row0 * bw), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void baro_wait(const int g, unsigned long long *bar) {
  const unsigned b = baro_sa(bar + (g & (kBaroRing - 1))), par = (unsigned)(g / kBaroRing) & 1u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tBRW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra BRD_%=;\n\tbra BRW_%=;\n\tBRD_%=:\n\t}" ::"r"(b),
      "r"(par)
      : "memory");
}
// the 32 pivots of one chunk.  FULL: this chunk and the chunk being expanded both hold 32 pivots (no per-pivot bounds
// checks); otherwise every pivot / row is guarded.  xe = &x[phys(32c)], xs = +-1 (x is padded by 64 zeros on both
// sides, so the register recycling load needs no guard).
template <bool SW, bool FULL>
__device__ __forceinline__ void baro_steps(double &P, double &N, const double rdv, const double *__restrict__ fb, double *xe,
                                           const int rows, const bool more, const double *src0, const int sstep, const int nrows,
                                           double *dst, const int lane, const bool dup) {
  constexpr int xs = SW ? -1 : 1;
#pragma unroll
  for (int u = 0; u < 32; u++) {
    if (FULL || u < rows) {
      const double fP = fb[u * kBaroRow + 31 - u], fN = fb[u * kBaroRow + 63 - u];
      if (SW) { if (lane == u) P = P * rdv; }                         // xi = x(i) / gap(i, n+2)
      const double gi = __shfl_sync(0xffffffffu, P, u);
      if (lane == u) {
        xe[xs * u] = P;                                               // final value of this sweep
        P = xe[xs * (u + 64)];                                        // the freed register takes element e+64
      }
      P = P - fP * gi;
      N = N - fN * gi;
    }
    // expansion of row u of the next chunk (already landed in the staging ring)
    if (FULL || (more && u < nrows)) {
      const double *src = src0 + u * sstep;
      dst[u * kBaroRow + 32] = src[0];
      if (dup) {
        const double f = src[32];
        dst[u * kBaroRow + 64] = f;
        dst[u * kBaroRow] = f;
      }
    }
  }
}
__global__ void __launch_bounds__(32) k_baro_reg(const Dev v, const double *__restrict__ bf_all, const double *__restrict__ bb_all,
                                                 const double *__restrict__ rd_all) {
  extern __shared__ __align__(128) double bsm[];
  const int lane = threadIdx.x, m = blockIdx.x;
  const int nm = v.nm, MS = v.MS, bw = v.I + 1, nd = bw - 32;
  const bool dup = lane < nd;
  double *stage = bsm;                                           // kBaroRing packed chunks of 32 x bw factors
  double *ex = stage + (size_t)kBaroRing * 32 * bw;              // two expanded chunks of 32 x 96
  double *xpad = ex + 2 * 32 * kBaroRow;                         // 64 zeros | the vector (rhs -> forward result -> solution) | 64 zeros
  double *x = xpad + 64;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(xpad + 128 + ((nm + 1) & ~1));
  const size_t grp = v.baro_group[m];
  const double *__restrict__ bf = bf_all + grp * nm * bw, *__restrict__ bb = bb_all + grp * nm * bw, *__restrict__ rd = rd_all + grp * nm;
  const int nch = (nm + 31) / 32, ntot = 2 * nch;
  if (lane == 0) {
    for (int q = 0; q < kBaroRing; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(baro_sa(bar + q)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g = 0; g < kBaroRing && g < ntot; g++) baro_issue(g, nch, nm, bw, bf, bb, stage, bar);
  }
  for (int r = lane; r < 2 * 32 * kBaroRow; r += 32) ex[r] = 0.0;
  for (int r = lane; r < 64; r += 32) { xpad[r] = 0.0; x[nm + r] = 0.0; }
  for (int r = lane; r < nm; r += 32) x[r] = v.gb[(size_t)r * MS + m];
  __syncwarp();
  {
    // chunk 0 is expanded on its own, every later chunk while the pivots of its predecessor run
    baro_wait(0, bar);
    const BaroChunk k0 = baro_chunk(0, nch, nm);
    for (int u = 0; u < k0.rows; u++) {
      const double *src = stage + u * bw + lane;
      double *dst = ex + u * kBaroRow + lane;
      dst[32] = src[0];
      if (dup) { const double f = src[32]; dst[64] = f; dst[0] = f; }
    }
    __syncwarp();
    if (lane == 0 && kBaroRing < ntot) baro_issue(kBaroRing, nch, nm, bw, bf, bb, stage, bar);
  }
  double P = 0.0, N = 0.0, rdv = 0.0;   // P: block of the current chunk, N: the next block (roles swap every chunk)
  for (int g = 0; g < ntot; g++) {
    const BaroChunk k = baro_chunk(g, nch, nm);
    const int sw = k.sw, c = k.c;
    if (c == 0) {
      __syncwarp();
      const int e0 = lane, e1 = 32 + lane;                            // element e' of this sweep lives at x[sw ? nm-1-e' : e']
      P = x[sw ? nm - 1 - e0 : e0];
      N = x[sw ? nm - 1 - e1 : e1];
      rdv = (sw && lane < nm) ? rd[nm - 1 - lane] : 0.0;
    }
    // reciprocal pivots of the next chunk (backward sweep), one chunk ahead of their use
    double rdn = 0.0;
    if (sw && 32 * (c + 1) + lane < nm) rdn = rd[nm - 1 - 32 * (c + 1) - lane];
    const bool more = g + 1 < ntot;
    BaroChunk kn = k;
    if (more) { kn = baro_chunk(g + 1, nch, nm); baro_wait(g + 1, bar); }
    const double *fb = ex + (g & 1) * 32 * kBaroRow + lane;
    double *xe = x + (sw ? nm - 1 - 32 * c : 32 * c);                // &x[phys(32c)]
    const double *src0 = stage + (size_t)((g + 1) & (kBaroRing - 1)) * 32 * bw + (kn.sw ? (kn.rows - 1) * bw : 0) + lane;
    const int sstep = kn.sw ? -bw : bw;
    double *dst = ex + ((g + 1) & 1) * 32 * kBaroRow + lane;
    const bool full = more && k.rows == 32 && kn.rows == 32;
    if (sw) {
      if (full) baro_steps<true, true>(P, N, rdv, fb, xe, k.rows, more, src0, sstep, kn.rows, dst, lane, dup);
      else baro_steps<true, false>(P, N, rdv, fb, xe, k.rows, more, src0, sstep, kn.rows, dst, lane, dup);
    } else {
      if (full) baro_steps<false, true>(P, N, rdv, fb, xe, k.rows, more, src0, sstep, kn.rows, dst, lane, dup);
      else baro_steps<false, false>(P, N, rdv, fb, xe, k.rows, more, src0, sstep, kn.rows, dst, lane, dup);
    }
    { const double t = P; P = N; N = t; }
    rdv = rdn;
    __syncwarp();                                                     // staging slot of chunk g+1 and ex[g&1] are free
    if (lane == 0 && g + 1 + kBaroRing < ntot) baro_issue(g + 1 + kBaroRing, nch, nm, bw, bf, bb, stage, bar);
  }
  __syncwarp();
  for (int r = lane; r < nm; r += 32) v.gb[(size_t)r * MS + m] = x[r];
}
// Barotropic solve, blocked form.  One warp = one member, lane = row within a block of 32 rows.  The pivot-by-pivot
// substitution is a chain of 2 x nm dependent (shuffle, multiply, subtract) steps; here a block's 32 unknowns come out
// together:   y_blk = Tbb^-1 (b_blk - sum_d T(e', e'-d) y(e'-d))   with the inverse of the block's own 32 x 32 triangle
// precomputed on the host (extended precision) -- per block bw + 32 independent shuffle + FMA pairs on four accumulators
// instead of 32 dependent pivots.  The backward sweep runs on the reversed index so that both share one body.  The
// coefficient slabs ((bw + 32) x 32 doubles per block, contiguous over the 2 nb blocks) stream through a ring of
// cp.async.bulk buffers.  Same equations, different summation order: agrees with k_baro_reg to rounding (<= 1e-13
// relative on psi, tests/test_gpu_col.py), not bit for bit.
constexpr int kBlkRing = 4;
template <int BW>
__global__ void __launch_bounds__(32) k_baro_blk(const Dev v, const double *__restrict__ bk_all, const int nb) {
  extern __shared__ __align__(128) double bsm[];
  constexpr int T = BW + 32;
  const int lane = threadIdx.x, m = blockIdx.x;
  const int nm = v.nm, MS = v.MS, npad = nb * 32, ntot = 2 * nb;
  double *ring = bsm;                                   // kBlkRing slabs of T x 32
  double *x = ring + (size_t)kBlkRing * T * 32;         // the vector, padded to npad
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(x + npad);
  const double *__restrict__ bk = bk_all + (size_t)v.baro_group[m] * ntot * T * 32;
  auto issue = [&](const int g) {
    const unsigned b = baro_sa(bar + (g & (kBlkRing - 1)));
    constexpr unsigned bytes = T * 32 * 8;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     baro_sa(ring + (size_t)(g & (kBlkRing - 1)) * T * 32)),
                 "l"(bk + (size_t)g * T * 32), "r"(bytes), "r"(b)
                 : "memory");
  };
  if (lane == 0) {
    for (int q = 0; q < kBlkRing; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(baro_sa(bar + q)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g = 0; g < kBlkRing && g < ntot; g++) issue(g);
  }
  for (int r = lane; r < npad; r += 32) x[r] = (r < nm) ? v.gb[(size_t)r * MS + m] : 0.0;
  __syncwarp();
  // z = Tbb^-1 b_blk of block g (independent of the unknowns of earlier blocks: computed one block ahead, under the
  // dependent chain of the block before)
  auto zblock = [&](const int g) -> double {
    const int sw = (g >= nb) ? 1 : 0, B = g - sw * nb;
    const double *cf = ring + (size_t)(g & (kBlkRing - 1)) * T * 32 + lane;
    const int ep = 32 * B + lane;
    const double r = x[sw ? npad - 1 - ep : ep];
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
    for (int c = 0; c < 32; c++) {
      const double rc = __shfl_sync(0xffffffffu, r, c);
      const double w = cf[(BW + c) * 32];
      if ((c & 3) == 0) b0 = __fma_rn(w, rc, b0);
      else if ((c & 3) == 1) b1 = __fma_rn(w, rc, b1);
      else if ((c & 3) == 2) b2 = __fma_rn(w, rc, b2);
      else b3 = __fma_rn(w, rc, b3);
    }
    return (b0 + b1) + (b2 + b3);
  };
  double Y1 = 0.0, Y2 = 0.0;
  baro_wait(0, bar);
  double z = zblock(0);
  for (int g = 0; g < ntot; g++) {
    const int sw = (g >= nb) ? 1 : 0, B = g - sw * nb;
    if (B == 0) { Y1 = 0.0; Y2 = 0.0; }
    const double *cf = ring + (size_t)(g & (kBlkRing - 1)) * T * 32 + lane;
    const int ep = 32 * B + lane, phys = sw ? npad - 1 - ep : ep;
    // the next block's z; at the turn of the sweeps it needs this block's result and is taken afterwards
    const bool ahead = (g + 1 < ntot) && (g + 1 != nb);
    double zn = 0.0;
    if (g + 1 < ntot) baro_wait(g + 1, bar);
    if (ahead) zn = zblock(g + 1);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int d = 1; d <= BW; d++) {
      // the value d rows before this block's first row: lane 32-d of the previous block, lane 64-d of the one before
      const double val = (d <= 32) ? __shfl_sync(0xffffffffu, Y1, 32 - d) : __shfl_sync(0xffffffffu, Y2, 64 - d);
      const double c = cf[(d - 1) * 32];
      if ((d & 3) == 0) a0 = __fma_rn(c, val, a0);
      else if ((d & 3) == 1) a1 = __fma_rn(c, val, a1);
      else if ((d & 3) == 2) a2 = __fma_rn(c, val, a2);
      else a3 = __fma_rn(c, val, a3);
    }
    const double y = z - ((a0 + a1) + (a2 + a3));
    x[phys] = y;
    Y2 = Y1; Y1 = y;
    __syncwarp();                                         // the slab of block g is free, x holds this block's result
    if (lane == 0 && g + kBlkRing < ntot) issue(g + kBlkRing);
    if (!ahead && g + 1 < ntot) zn = zblock(g + 1);
    z = zn;
  }
  __syncwarp();
  for (int r = lane; r < nm; r += 32) v.gb[(size_t)r * MS + m] = x[r];
}
// k_baro_blk with FOUR warps per member.  The one-warp form issues ~280 dependent-ish instructions per block of 32 unknowns from
// a single warp (ncu: IPC 0.43 per SM at 3.5 warps per SM, 144 us per launch at 512 members: pure issue latency).  Here warp w
// owns accumulator w of both sums -- the terms d with (d & 3) == w of the substitution and the terms c with (c & 3) == w of
// z = Tbb^-1 b -- so each warp issues a quarter of the instructions, and the four partial sums meet in the SAME order
// ((a0 + a1) + (a2 + a3)) as in the one-warp form: bit-identical results.  The values the one-warp form passes by shuffle
// (the two previous blocks' unknowns, the block's right-hand side) are read from the vector in shared memory (one broadcast
// LDS instead of two SHFL).  Two block barriers per block of unknowns.
template <int BW>
__global__ void __launch_bounds__(128) k_baro_blk4(const Dev v, const double *__restrict__ bk_all, const int nb) {
  extern __shared__ __align__(128) double bsm[];
  constexpr int T = BW + 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, m = blockIdx.x;
  const int nm = v.nm, MS = v.MS, npad = nb * 32, ntot = 2 * nb;
  double *ring = bsm;                                   // kBlkRing slabs of T x 32
  double *x = ring + (size_t)kBlkRing * T * 32;         // the vector, padded to npad
  double *pa = x + npad;                                // [4][32] partial sums of the substitution
  double *pz = pa + 128;                                // [4][32] partial sums of z of the next block
  double *zs = pz + 128;                                // [32] z of the current block
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(zs + 32);
  const double *__restrict__ bk = bk_all + (size_t)v.baro_group[m] * ntot * T * 32;
  auto issue = [&](const int g) {
    const unsigned b = baro_sa(bar + (g & (kBlkRing - 1)));
    constexpr unsigned bytes = T * 32 * 8;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     baro_sa(ring + (size_t)(g & (kBlkRing - 1)) * T * 32)),
                 "l"(bk + (size_t)g * T * 32), "r"(bytes), "r"(b)
                 : "memory");
  };
  if (threadIdx.x == 0) {
    for (int q = 0; q < kBlkRing; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(baro_sa(bar + q)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g = 0; g < kBlkRing && g < ntot; g++) issue(g);
  }
  for (int r = threadIdx.x; r < npad; r += 128) x[r] = (r < nm) ? v.gb[(size_t)r * MS + m] : 0.0;
  __syncthreads();
  // this warp's quarter of z = Tbb^-1 b_blk of block g (terms c = w, w + 4, ...: accumulator w of the one-warp form)
  auto zpart = [&](const int g) -> double {
    const int sw = (g >= nb) ? 1 : 0, B = g - sw * nb;
    const double *cf = ring + (size_t)(g & (kBlkRing - 1)) * T * 32 + lane;
    double b = 0.0;
#pragma unroll
    for (int c = w; c < 32; c += 4) {
      const int ep = 32 * B + c;
      const double rc = x[sw ? npad - 1 - ep : ep];
      b = __fma_rn(cf[(BW + c) * 32], rc, b);
    }
    return b;
  };
  baro_wait(0, bar);
  pz[w * 32 + lane] = zpart(0);
  __syncthreads();
  if (w == 0) zs[lane] = (pz[lane] + pz[32 + lane]) + (pz[64 + lane] + pz[96 + lane]);
  __syncthreads();
  for (int g = 0; g < ntot; g++) {
    const int sw = (g >= nb) ? 1 : 0, B = g - sw * nb;
    const double *cf = ring + (size_t)(g & (kBlkRing - 1)) * T * 32 + lane;
    const int ep = 32 * B + lane, phys = sw ? npad - 1 - ep : ep;
    const bool ahead = (g + 1 < ntot) && (g + 1 != nb);
    if (g + 1 < ntot) baro_wait(g + 1, bar);
    if (ahead) pz[w * 32 + lane] = zpart(g + 1);
    // accumulator w: the terms d with (d & 3) == w, d ascending (d = 4, 8, ... for w = 0)
    double a = 0.0;
#pragma unroll
    for (int d = (w == 0 ? 4 : w); d <= BW; d += 4) {
      // the value d rows before this block's first row (zero ahead of the sweep's first block)
      const int e = 32 * B - d;
      const double val = (e >= 0) ? x[sw ? npad - 1 - e : e] : 0.0;
      a = __fma_rn(cf[(d - 1) * 32], val, a);
    }
    pa[w * 32 + lane] = a;
    __syncthreads();                                      // partial sums complete; every warp has read slab g and the vector
    if (w == 0) {
      const double a0 = pa[lane], a1 = pa[32 + lane], a2 = pa[64 + lane], a3 = pa[96 + lane];
      x[phys] = zs[lane] - ((a0 + a1) + (a2 + a3));
      if (ahead) zs[lane] = (pz[lane] + pz[32 + lane]) + (pz[64 + lane] + pz[96 + lane]);
      if (lane == 0 && g + kBlkRing < ntot) issue(g + kBlkRing);
    }
    __syncthreads();                                      // x holds this block's unknowns
    if (!ahead && g + 1 < ntot) {                         // the turn of the sweeps: z of the first backward block needs them
      pz[w * 32 + lane] = zpart(g + 1);
      __syncthreads();
      if (w == 0) zs[lane] = (pz[lane] + pz[32 + lane]) + (pz[64 + lane] + pz[96 + lane]);
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < nm; r += 128) v.gb[(size_t)r * MS + m] = x[r];
}
static_assert(kBlkRing == kBaroRing, "k_baro_blk reuses baro_wait");

bool baro_reg_ok(const Dev &v) { return v.I + 1 > 32 && v.I + 1 <= 64 && (v.nm % 2) == 0 && v.nm >= 64; }

// psi and barotropic velocity from the solved gb (ubarsolv :3527-3564)
__global__ void __launch_bounds__(128) k_psi2ub(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y * blockDim.y + threadIdx.y;  // over (0:I+1, 0:J)
  if (m >= v.M || p >= (I + 2) * (J + 1)) return;
  const int i = p % (I + 2), j = p / (I + 2);
  const int iw = (i == 0) ? I : ((i == I + 1) ? 1 : i);  // periodic image
#define GBP(ii, jj) v.gb[(size_t)(((ii)-1) + (jj)*I) * MS + m]
  if (i <= I) PSIX(i, j) = GBP(iw, j);
  double u1 = 0.0, u2 = 0.0;
  if (j >= 1) u1 = -RHX(1, iw, j) * c_g.c[j] * (GBP(iw, j) - GBP(iw, j - 1)) * c_g.rds[j];
  if (j >= 1 && j <= J - 1) {
    const int iww = (iw > 1) ? iw - 1 : I;
    u2 = RHX(2, iw, j) * (GBP(iw, j) - GBP(iww, j)) * c_g.rcv[j] * c_g.rdphi;
  }
  if (j >= 1) UBX(1, i, j) = u1;
  UBX(2, i, j) = u2;
#undef GBP
}

// island path integral (island, goldstein_lib.f90:186-241, indj = 1) and psibc (goldstein.f90:203-216).
// One warp per member: lanes evaluate the path points in parallel, lane 0 adds the terms in the
// reference's order (t1(1), t2(1), t1(2), ...), so the sum is the sequential one bit for bit.
__global__ void __launch_bounds__(128) k_island(const Dev v) {
  DIMS
  extern __shared__ double isl_terms[];            // [warps per block][2 * mpi]
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + wib;
  if (m >= v.M) return;
  double *terms = isl_terms + (size_t)wib * 2 * v.mpi;
  const size_t nf = (size_t)I * J * MS;
  const int nis = v.isles;
  double rhs[kMaxIsles];
  for (int is = 0; is < nis; is++) {
    const int np = v.npi[is];
    const int *lp = v.lpisl + (size_t)is * v.mpi, *ip_ = v.ipisl + (size_t)is * v.mpi, *jp = v.jpisl + (size_t)is * v.mpi;
    for (int p = lane; p < np; p += 32) {
      const int lpi = lp[p], ipi = ip_[p], jpi = jp[p];
      const int al = abs(lpi), sg = (lpi >= 0) ? 1 : -1;
      double cor;
      if (al == 1)
        cor = -c_g.s[jpi] * 0.25 * (UBX(2, ipi, jpi) + UBX(2, ipi + 1, jpi) + UBX(2, ipi, jpi - 1) + UBX(2, ipi + 1, jpi - 1));
      else
        cor = c_g.sv[jpi] * 0.25 * (UBX(1, ipi - 1, jpi) + UBX(1, ipi, jpi) + UBX(1, ipi - 1, jpi + 1) + UBX(1, ipi, jpi + 1));
      const double tau = v.tau[A2I(ipi, jpi) + (al - 1) * nf];
      const double t1 = sg * (DRAGX(al, ipi, jpi) * UBX(al, ipi, jpi) + cor - 1 * tau * RHX(al, ipi, jpi)) *
                        (c_g.c[jpi] * c_g.dphi * (2.0 - al) + c_g.rcv[jpi] * c_g.dsv[jpi] * (al - 1.0));
      double t2;
      const int ipw = (ipi < I) ? ipi + 1 : 1;
      if (al == 1) {
        double tv1 = 0.0;
        for (int k = KUX(1, ipi, jpi); k <= MKX(ipi + 1, jpi); k++) tv1 = tv1 + BPX(ipw, jpi, k) * c_g.dz[k];
        for (int k = KUX(1, ipi, jpi); k <= MKX(ipi, jpi); k++) tv1 = tv1 - BPX(ipi, jpi, k) * c_g.dz[k];
        t2 = (SBPX(ipw, jpi) - SBPX(ipi, jpi) + tv1) * sg * RHX(1, ipi, jpi);
      } else {
        double tv2 = 0.0;
        for (int k = KUX(2, ipi, jpi); k <= MKX(ipi, jpi + 1); k++) tv2 = tv2 + BPX(ipi, jpi + 1, k) * c_g.dz[k];
        for (int k = KUX(2, ipi, jpi); k <= MKX(ipi, jpi); k++) tv2 = tv2 - BPX(ipi, jpi, k) * c_g.dz[k];
        t2 = (SBPX(ipi, jpi + 1) - SBPX(ipi, jpi) + tv2) * sg * RHX(2, ipi, jpi);
      }
      terms[2 * p] = t1;
      terms[2 * p + 1] = t2;
    }
    __syncwarp();
    double e = 0.0;
    for (int p = 0; p < 2 * np; p++) e = e + terms[p];   // every lane adds the same terms in the reference's order
    rhs[is] = e;
    __syncwarp();
  }
  if (lane == 0) {
    const double *A = v.erisl + (size_t)v.baro_group[m] * nis * (nis + 1);   // erisl(isl, col), isl fastest, after matinv_gold
#define EA(r, c) A[((r)-1) + nis * ((c)-1)]
    for (int is = 0; is < nis; is++) v.erisl_rhs[(size_t)is * MS + m] = rhs[is];
    if (nis > 1) {
      // matmult (goldstein.f90:3470-3492) on the right-hand side, then psibc(isl) = -erisl(isl, isles+1)
      for (int i = 1; i <= nis - 1; i++)
        for (int j = i + 1; j <= nis; j++) rhs[j - 1] = EA(i, i) * rhs[j - 1] - EA(j, i) * rhs[i - 1];
      rhs[nis - 1] = rhs[nis - 1] / EA(nis, nis);
      for (int i = nis - 1; i >= 1; i--) {
        for (int j = i + 1; j <= nis; j++) rhs[i - 1] = rhs[i - 1] - EA(i, j) * rhs[j - 1];
        rhs[i - 1] = rhs[i - 1] / EA(i, i);
      }
      for (int is = 0; is < nis; is++) v.psibc[(size_t)is * MS + m] = -rhs[is];
    } else {
      v.psibc[m] = -rhs[0] / EA(1, 1);  // isles == 1: psibc(1) = -erisl(1,2)/erisl(1,1)
    }
#undef EA
  }
}

// add the island contribution to ub and psi (goldstein.f90:218-230): ub + SUM(ubisl(:, 1:isles) * psibc(1:isles))
__global__ void __launch_bounds__(128) k_ubadd(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || p >= (I + 2) * (J + 1)) return;
  const int i = p % (I + 2), j = p / (I + 2);
  const int nis = v.isles;
  const size_t g = v.baro_group[m];
  const size_t nub = (size_t)2 * (I + 2) * (J + 1), nps = (size_t)(I + 1) * (J + 1);
  if (nis == 1) {
    const double psibc = v.psibc[m];
    if (j >= 1) {
      const double *ubisl = v.ubisl + g * nub;
      UBX(1, i, j) = UBX(1, i, j) + ubisl[0 + 2 * (i + (I + 2) * j)] * psibc;
      UBX(2, i, j) = UBX(2, i, j) + ubisl[1 + 2 * (i + (I + 2) * j)] * psibc;
    }
    if (i <= I) {
      const double *psisl = v.psisl + g * nps;
      PSIX(i, j) = PSIX(i, j) + psisl[i + (I + 1) * j] * psibc;
    }
    return;
  }
  if (j >= 1) {
    double s1 = 0.0, s2 = 0.0;
    for (int is = 0; is < nis; is++) {
      const double *ubisl = v.ubisl + (g * nis + is) * nub;
      const double pb = v.psibc[(size_t)is * MS + m];
      s1 = s1 + ubisl[0 + 2 * (i + (I + 2) * j)] * pb;
      s2 = s2 + ubisl[1 + 2 * (i + (I + 2) * j)] * pb;
    }
    UBX(1, i, j) = UBX(1, i, j) + s1;
    UBX(2, i, j) = UBX(2, i, j) + s2;
  }
  if (i <= I) {
    double s = 0.0;
    for (int is = 0; is < nis; is++) s = s + v.psisl[(g * nis + is) * nps + i + (I + 1) * j] * v.psibc[(size_t)is * MS + m];
    PSIX(i, j) = PSIX(i, j) + s;
  }
}

// baroclinic velocities, barotropic correction and relaxation (velc :3574-3658)
// Two kernels: k_velc1 needs only rho and constants (the baroclinic shear integral and its depth mean), so it can run
// next to the barotropic solve; k_velc2 adds the barotropic velocity and the time relaxation.
__global__ void __launch_bounds__(128) k_velc1(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  if (k1c > K) return;
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k_e = CG_K1(v, ip, j), k_w = CG_K1(v, im, j), k_n = CG_K1(v, i, j + 1), k_s = CG_K1(v, i, j - 1);
  const int k_ne = CG_K1(v, ip, j + 1), k_se = CG_K1(v, ip, j - 1), k_nw = CG_K1(v, im, j + 1);
  const double sj = c_g.s[j], svj = c_g.sv[j], cj = c_g.c[j], cvj = c_g.cv[j], rdphi = c_g.rdphi, rcj = c_g.rc[j];
  const double rcvj = (j < J) ? c_g.rcv[j] : 0.0, rdsvj = (j < J) ? c_g.rdsv[j] : 0.0, rdsvjm = (j > 1) ? c_g.rdsv[j - 1] : 0.0,
               rds2j = (j > 1 && j < J) ? c_g.rds2[j] : 0.0;
  const double drag1 = DRAGX(1, i, j), drag2 = DRAGX(2, i, j), rtv = v.rtv[A2I(i, j)], rtv3 = v.rtv3[A2I(i, j)];
  const size_t nf = (size_t)I * J * MS;
  const double rel = v.p.rel[m];
  double sum1 = 0, sum2 = 0, dzu1p = 0, dzu2p = 0, u1p = 0, u2p = 0;
  for (int k = k1c; k <= K; k++) {
    double tv1, tv2, tv4, tv5;
    const double rc0 = RHOX(i, j, k);
    if (k_e > k) {
      tv1 = 0;
      tv2 = 0;
    } else {
      const double re = RHOX(ip, j, k);
      tv2 = -(re - rc0) * rdphi * rcj;
      if (max(max(k_s, k_n), max(k_se, k_ne)) <= k)
        tv1 = -cj * (RHOX(ip, j + 1, k) - RHOX(ip, j - 1, k) + RHOX(i, j + 1, k) - RHOX(i, j - 1, k)) * rds2j * 0.25;
      else if (max(k_s, k_se) <= k)
        tv1 = -cj * (re - RHOX(ip, j - 1, k) + rc0 - RHOX(i, j - 1, k)) * rdsvjm * 0.5;
      else if (max(k_n, k_ne) <= k)
        tv1 = -cj * (RHOX(ip, j + 1, k) - re + RHOX(i, j + 1, k) - rc0) * rdsvj * 0.5;
      else
        tv1 = 0;
    }
    if (k_n > k) {
      tv4 = 0;
      tv5 = 0;
    } else {
      const double rn = RHOX(i, j + 1, k);
      tv4 = -cvj * (rn - rc0) * rdsvj;
      if (max(max(k_w, k_nw), max(k_e, k_ne)) <= k)
        tv5 = -(RHOX(ip, j + 1, k) - RHOX(im, j + 1, k) + RHOX(ip, j, k) - RHOX(im, j, k)) * rdphi * 0.25 * rcvj;
      else if (max(k_w, k_nw) <= k)
        tv5 = -(rn - RHOX(im, j + 1, k) + rc0 - RHOX(im, j, k)) * rdphi * 0.5 * rcvj;
      else if (max(k_e, k_ne) <= k)
        tv5 = -(RHOX(ip, j + 1, k) - rn + RHOX(ip, j, k) - rc0) * rdphi * 0.5 * rcvj;
      else
        tv5 = 0;
    }
    if (k == K) {
      if (k_e <= k) {
        tv1 = tv1 - v.dztau[A2I(i, j) + nf];
        tv2 = tv2 - v.dztau[A2I(i, j)];
      }
      if (k_n <= k) {
        tv4 = tv4 - v.dztav[A2I(i, j) + nf];
        tv5 = tv5 - v.dztav[A2I(i, j)];
      }
    }
    const double dzu1 = -(sj * tv1 + drag1 * tv2) * rtv;
    const double dzu2 = -(drag2 * tv4 - svj * tv5) * rtv3;
    double ua, ub_;
    if (k == k1c) {
      ua = 0;
      ub_ = 0;
    } else {
      ua = u1p + c_g.dza[k - 1] * (dzu1 + dzu1p) * 0.5;
      ub_ = u2p + c_g.dza[k - 1] * (dzu2 + dzu2p) * 0.5;
      sum1 = sum1 + c_g.dz[k] * ua;
      sum2 = sum2 + c_g.dz[k] * ub_;
    }
    UX(1, i, j, k) = ua;
    UX(2, i, j, k) = ub_;
    u1p = ua; u2p = ub_; dzu1p = dzu1; dzu2p = dzu2;
  }
  v.velsum[A2I(i, j)] = sum1;
  v.velsum[A2I(i, j) + nf] = sum2;
}
__global__ void __launch_bounds__(128) k_velc2(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  if (k1c > K) return;
  const int ip = (i < I) ? i + 1 : 1;
  const int k_e = CG_K1(v, ip, j), k_n = CG_K1(v, i, j + 1);
  const size_t nf = (size_t)I * J * MS;
  const double rel = v.p.rel[m];
  const double sum1 = v.velsum[A2I(i, j)], sum2 = v.velsum[A2I(i, j) + nf];
  const double rh1 = RHX(1, i, j), rh2 = RHX(2, i, j), ub1 = UBX(1, i, j), ub2 = UBX(2, i, j);
  for (int k = k1c; k <= K; k++) {
    if (k_e <= k) {
      double x = UX(1, i, j, k) - sum1 * rh1 + ub1;
      x = rel * U1X(1, i, j, k) + (1.0 - rel) * x;
      UX(1, i, j, k) = x;
      U1X(1, i, j, k) = x;
    }
    if (k_n <= k) {
      double x = UX(2, i, j, k) - sum2 * rh2 + ub2;
      x = rel * U1X(2, i, j, k) + (1.0 - rel) * x;
      UX(2, i, j, k) = x;
      U1X(2, i, j, k) = x;
    }
  }
}

// vertical velocity from continuity (velc :3668-3678)
__global__ void __launch_bounds__(128) k_w(const Dev v) {
  DIMS
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  if (k1c > K) return;
  const int im = (i > 1) ? i - 1 : I;
  double tv = 0;
  for (int k = k1c; k <= K - 1; k++) {
    const double uw = UX(1, im, j, k), vs = (j > 1) ? UX(2, i, j - 1, k) : 0.0;
    const double tv1 = (UX(1, i, j, k) - uw) * c_g.rdphi * c_g.rc[j];
    const double tv2 = (UX(2, i, j, k) * c_g.cv[j] - vs * c_g.cv[j - 1]) * c_g.rds[j];
    const double w = tv - c_g.dz[k] * (tv1 + tv2);
    UX(3, i, j, k) = w;
    tv = w;
  }
}

// ---------------------------------------------------------------- diagnostics
// volume-weighted global tracer means per member: one warp per (member-chunk of 32? no:) block per
// tracer, warp-shuffle reduction over cells; the sum order is fixed (deterministic).
__global__ void __launch_bounds__(256) k_global_means(const Dev v, double *out) {
  DIMS
  const int m = blockIdx.x, l = blockIdx.y, L = v.L;
  double num = 0.0, den = 0.0;
  const int ncell = I * J * K;
  for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
    const int i = c % I + 1, j = (c / I) % J + 1, k = c / (I * J) + 1;
    if (k >= CG_K1(v, i, j)) {
      const double w = c_g.ds[j] * c_g.dz[k];
      num += v.ts_cur[((size_t)c * L + l) * MS + m] * w;
      den += w;
    }
  }
  __shared__ double sn[8], sd[8];
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_down_sync(0xffffffffu, num, o);
    den += __shfl_down_sync(0xffffffffu, den, o);
  }
  if ((threadIdx.x & 31) == 0) { sn[threadIdx.x >> 5] = num; sd[threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += sn[w]; b += sd[w]; }
    out[(size_t)m * L + l] = a / b;
  }
}

// non-finite guard (mirrors the avs blow-up check, goldstein_diag.f90:52-58)
__global__ void __launch_bounds__(256) k_health(const Dev v, int *flags) {
  DIMS
  const int m = blockIdx.x;
  const size_t n = (size_t)I * J * K * v.L;
  int bad = 0;
  for (size_t c = threadIdx.x; c < n; c += blockDim.x) {
    const double x = v.ts_cur[c * MS + m];
    if (!(fabs(x) < 1.0e20)) bad = 1;
  }
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flags[m] = bad;
}

// ---------------------------------------------------------------- launchers
struct LaunchCtx { cudaStream_t s; long long *count; };

static inline dim3 grid2(const Dev &v, int ncells, dim3 b) { return dim3((v.M + b.x - 1) / b.x, (ncells + b.y - 1) / b.y); }

void launch_step_begin(const Dev &v, cudaStream_t s) { k_step_begin<<<1, 32, 0, s>>>(v); }
void launch_usnap(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_usnap<<<dim3(v.MS / 32, (v.I * v.J + 3) / 4), b, 0, s>>>(v);
}
void launch_hosing(const Dev &v, cudaStream_t s) { k_hosing<<<(v.M + 127) / 128, 128, 0, s>>>(v); }
int launch_surflux(const Dev &v, double *meantemp, bool need_mean, cudaStream_t s) {
  int n = 0;
  if (need_mean) { k_meantemp<<<(v.M + 31) / 32, 32, 0, s>>>(v, meantemp); n++; }
  const dim3 b(32, 4);
  k_surflux1<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v, need_mean ? meantemp : nullptr);
  k_surflux2<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  return n + 2;
}
int launch_embm(const Dev &v, int nsteps, cudaStream_t s) {
  const int nc = v.I * v.J;
  const size_t sm = sizeof(double) * 2 * v.I * (v.J + 2);
  auto thr = [&](int cpt) { return (((nc + cpt - 1) / cpt + 31) / 32) * 32; };
  if (nc <= 704) k_embm<1><<<v.M, thr(1), sm, s>>>(v, nsteps);
  else if (nc <= 1344 && getenv("CG_EMBM_CPT3")) k_embm<3><<<v.M, thr(3), sm, s>>>(v, nsteps);   // 36 x 36: 448 threads x 3 cells, no spills (2 cells x 672 threads: 80 registers, spills)
  else if (nc == 1296 && getenv("CG_EMBM_EXACT")) {   // 36 x 36, every thread owns its cells, no guards, corrector peeled: measured SLOWER
    // (52.8 / 48.2 against 46.2 us per member-year: the peeled loop spills more), kept as a knob
    if (getenv("CG_EMBM_X3")) k_embm_x3<<<v.M, 432, sm, s>>>(v, nsteps);
    else k_embm_x2<<<v.M, 648, sm, s>>>(v, nsteps);
  }
  else if (nc <= 1408 && getenv("CG_EMBM_SMC")) {   // measured: 47.0 against 46.1 us per member-year -- the loop loses 40 % of its instructions and no time
    const size_t sm2 = sm + sizeof(double) * 8 * nc;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_embm_s2, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
    k_embm_s2<<<v.M, thr(2), sm2, s>>>(v, nsteps);
  }
  else if (nc <= 1408) k_embm<2><<<v.M, thr(2), sm, s>>>(v, nsteps);
  else if (nc <= 2816) k_embm<4><<<v.M, thr(4), sm, s>>>(v, nsteps);
  else return -1;
  return 1;
}
int launch_seaice(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_seaice1<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  k_seaice2<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  return 2;
}
// tstar_ocn / sstar_ocn export at the end of step_goldstein (goldstein.f90:428-431)
__global__ void k_sst(const Dev v) {
  const int I = v.I, J = v.J, K = v.K, MS = v.MS;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const size_t q = (size_t)c2 * MS + m;
  if (CG_K1(v, i, j) > K) return;
  const size_t ot = cell3(I, J, i, j, K) * (size_t)v.L * MS + m;
  v.sst[q] = v.ts_cur[ot];
  v.sst[(size_t)I * J * MS + q] = v.ts_cur[ot + MS];
}
int launch_sst(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_sst<<<dim3(v.MS / 32, (v.I * v.J + 3) / 4), b, 0, s>>>(v);
  return 1;
}
// ---------------------------------------------------------------- Kraus-Turner mixed-layer scheme (imld = 1)
// tstepo, goldstein.f90:2294-2390, and SUBROUTINE krausturner, :3337-3442.  Thread = (member, wet column); the column's cells lie
// MS doubles apart per tracer, so a warp reads 256-byte rows.  Reference operation order throughout (-fmad=false): bit-exact
// against the oracle's restatement.  The 2-D "depth" grids dzg / z2dzg / rdzg (:1013-1027) are formed from zw on the fly: the
// same IEEE subtraction, products and division the reference stores.
#define TSN(l, k) v.ts_new[((cell3(I, J, i, j, (k)) * L) + ((l)-1)) * MS + m]
__device__ inline double mld_eos(const Dev &v, const int m, const double t, const double s, const double z) {
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  if (!v.ieos) return ec1 * t + ec2 * s + ec3 * (t * t) + ec4 * (t * t * t);
  return ec1 * t + ec2 * s + ec3 * (t * t) + ec4 * (t * t * t) + v.p.ec5[m] * t * z;
}
__device__ inline double mld_dzg(const int k, const int kk) { return c_g.zw[k] - c_g.zw[kk - 1]; }
__device__ inline double mld_z2dzg(const int k, const int kk) { return -c_g.zw[k] * c_g.zw[k] + c_g.zw[kk - 1] * c_g.zw[kk - 1]; }
__device__ inline double mld_rdzg(const int k, const int kk) { return (k != kk - 1) ? 1.0 / mld_dzg(k, kk) : 1.0e10; }

// before tstepo_flux (:2294-2309): energy consumed or released in mixing the surface forcing over the top layer
__global__ void __launch_bounds__(128) k_mld_pre(const Dev v) {
  DIMS
  const int L = v.L;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  if (CG_K1(v, i, j) > K) return;
  const size_t q = (size_t)c2 * MS + m;
  const double *tsc = v.ts_cur + ((cell3(I, J, i, j, K) * L) * MS + m);
  const double t = tsc[0] - v.tsflux[q], sa = tsc[MS] - v.tsflux[(size_t)I * J * MS + q];
  const double r = mld_eos(v, m, t, sa, c_g.zro[K]);
  v.mld_pel1[q] = (r - RHOX(i, j, K)) * mld_z2dzg(K, K);
}
// behind tstepo_flux, ahead of co (:2320-2331): the reference remembers T and S of every level "only so we can calculate PE
// change"; what it takes from them is eos(T, S, zro(k)) (:2346), kept here instead
__global__ void __launch_bounds__(128) k_mld_save(const Dev v) {
  DIMS
  const int L = v.L;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  if (CG_K1(v, i, j) > K) return;
  for (int k = K; k > 0; k--) v.mld_rhoold[cell3(I, J, i, j, k) * MS + m] = mld_eos(v, m, TSN(1, k), TSN(2, k), c_g.zro[k]);
}
// behind co (:2336-2390): PE released by the convective adjustment, wind energy, then krausturner on the column
constexpr int kMldMaxL = 64;
__global__ void __launch_bounds__(128) k_mld_kt(const Dev v) {
  DIMS
  const int L = v.L;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int tvkl = CG_K1(v, i, j);
  if (tvkl > K) return;
  const size_t q = (size_t)c2 * MS + m;
  double peconv = 0;
  for (int k = K; k > 0; k--) {
    const double rnew = mld_eos(v, m, TSN(1, k), TSN(2, k), c_g.zro[k]);
    peconv = peconv + (rnew - v.mld_rhoold[cell3(I, J, i, j, k) * MS + m]) * mld_z2dzg(k, k);
  }
  const double pel1 = v.mld_pel1[q];
  double pebuoy = peconv + pel1;
  if (pebuoy > 0.0) pebuoy = pebuoy * v.mldpebuoycoeff;
  const double ketau = v.mldketau[q];
  const double emix = pebuoy + ketau * c_g.mlddec[K];
  if (emix > 0.0) {
    // ---- krausturner(ts(:, i, j, :), pebuoy, ketau, mldpk, mld, mldk, k1)
    double qmix[kMldMaxL + 1];
    double em, empe, emke, eneed, emr, smix, tmix, rhou = 0.0, rhol = 0.0, rhomix;
    int k = K, partmix = 1;
    const int mldpk = K;
    empe = pebuoy;
    emke = ketau * c_g.mlddec[k];
    em = empe + emke;
    eneed = 0.0;
    tmix = TSN(1, k);
    smix = TSN(2, k);
    for (int l = 1; l <= L; l++) qmix[l] = TSN(l, k);
    rhomix = mld_eos(v, m, tmix, smix, c_g.zw[k]);
    while (eneed < em && em > 0) {
      if (k < mldpk) {
        qmix[1] = tmix;
        qmix[2] = smix;
        for (int l = 3; l <= L; l++) qmix[l] = mld_rdzg(K, k) * (qmix[l] * mld_dzg(K, k + 1) + TSN(l, k) * c_g.dz[k]);
      }
      if (k == tvkl) {
        v.mldk[q] = k;
        v.mld[q] = c_g.zw[k - 1];
        partmix = 0;
        em = -1.0e-8;
        if (k < K)
          for (int l = 1; l <= L; l++)
            for (int n = k; n <= K; n++) TSN(l, n) = qmix[l];
      } else {
        k = k - 1;
        emr = (em - eneed) / em;
        empe = empe * emr;
        emke = emke * emr * c_g.mlddecd[k];
        em = empe + emke;
        if (v.ieos) rhou = mld_eos(v, m, tmix, smix, c_g.zw[k]); else rhou = rhomix;
        tmix = mld_rdzg(K, k) * (tmix * mld_dzg(K, k + 1) + TSN(1, k) * c_g.dz[k]);
        smix = mld_rdzg(K, k) * (smix * mld_dzg(K, k + 1) + TSN(2, k) * c_g.dz[k]);
        rhol = mld_eos(v, m, TSN(1, k), TSN(2, k), c_g.zw[k]);
        rhomix = mld_eos(v, m, tmix, smix, c_g.zw[k]);
        eneed = mld_z2dzg(K, k + 1) * rhou + mld_z2dzg(k, k) * rhol - mld_z2dzg(K, k) * rhomix;
      }
    }
    if (partmix == 1 && em > 0) {
      v.mldk[q] = k;
      const double mlda = mld_dzg(K, k + 1) * mld_rdzg(K, k) * em / eneed;
      const double mldb = (mld_dzg(K, k) * mld_rdzg(K, k + 1) - 1) * mlda;
      for (int l = 1; l <= L; l++) {
        const double top = (1 - mldb) * qmix[l] + mldb * TSN(l, k);
        TSN(l, K) = top;
        for (int n = k + 1; n <= K - 1; n++) TSN(l, n) = top;
        TSN(l, k) = mlda * qmix[l] + (1 - mlda) * TSN(l, k);
      }
      if (k < K) {
        const double mldtadd = em / (c_g.zw[k] * (rhol - rhou));
        v.mld[q] = c_g.zw[k] + mldtadd;
      } else {
        v.mld[q] = 0.0;
      }
    } else if (partmix == 1 && k < K) {
      v.mldk[q] = k + 1;
      v.mld[q] = c_g.zw[k];
    }
  } else {
    // not enough energy even to homogenise the first layer (:2376-2387)
    v.mldk[q] = K;
    if (pel1 < 0) v.mld[q] = c_g.zw[K - 1] * (1 - emix / pel1);
    else v.mld[q] = c_g.zw[K - 1];
  }
  // "if thermobaricity is on, make sure rho calculation is vertically local" (:2396-2408) follows krausturner in the reference:
  // the convection kernel's pass saw the column before the mixed-layer scheme changed it
  if (v.ieos)
    for (int k = tvkl; k <= K; k++) RHOX(i, j, k) = mld_eos(v, m, TSN(1, k), TSN(2, k), c_g.zro[k]);
}
#undef TSN
int launch_mld_pre(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_mld_pre<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  return 1;
}
int launch_mld_save(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_mld_save<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  return 1;
}
int launch_mld_kt(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_mld_kt<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  return 1;
}

int launch_gold_pre(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_gold_pre<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  return 1;
}
int launch_momentum(const Dev &v, int fast, const double *bf, const double *bb, const double *rd, const double *bk, cudaStream_t s) {
  const dim3 b(32, 4);
  k_bp<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  k_gb<<<grid2(v, v.nm, b), b, 0, s>>>(v);
  int blk = 1;           // CG_BARO_BLK=0: pivot-by-pivot k_baro_reg instead of the blocked solve
  { const char *e = getenv("CG_BARO_BLK"); if (e) blk = atoi(e); }
  if (fast == 2 && blk && bk && v.I + 1 == 37) {
    constexpr int BW = 37;
    const int nb = (v.nm + 31) / 32;
    const size_t smem = sizeof(double) * ((size_t)kBlkRing * (BW + 32) * 32 + (size_t)nb * 32) + 8 * kBlkRing;
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(k_baro_blk<BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_baro_blk4<BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr = true;
    }
    const char *e4 = getenv("CG_BARO_W4");   // CG_BARO_W4=0: the one-warp form (read per launch: the captured graphs keep their choice)
    const int w4 = e4 ? atoi(e4) : 1;
    if (w4) k_baro_blk4<BW><<<v.M, 128, smem + (128 + 128 + 32) * sizeof(double), s>>>(v, bk, nb);
    else k_baro_blk<BW><<<v.M, 32, smem, s>>>(v, bk, nb);
  } else if (fast == 2 && baro_reg_ok(v)) {
    const size_t smem = sizeof(double) * ((size_t)kBaroRing * 32 * (v.I + 1) + 2 * 32 * kBaroRow + 128 + ((v.nm + 1) & ~1)) + 8 * kBaroRing;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_baro_reg, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
    k_baro_reg<<<v.M, 32, smem, s>>>(v, bf, bb, rd);
  } else if (fast) {
    const int wpb = 2;  // members (warps) per block
    k_baro_fast<<<(v.M + wpb - 1) / wpb, 32 * wpb, sizeof(double) * v.nm * wpb, s>>>(v, bf, bb, rd);
  } else {
    k_baro_strict<<<(v.M + 31) / 32, 32, 0, s>>>(v);
  }
  k_psi2ub<<<grid2(v, (v.I + 2) * (v.J + 1), b), b, 0, s>>>(v);
  k_island<<<(v.M + 3) / 4, 128, (size_t)4 * 2 * v.mpi * sizeof(double), s>>>(v);
  k_ubadd<<<grid2(v, (v.I + 2) * (v.J + 1), b), b, 0, s>>>(v);
  return 6;
}
// baroclinic shear (independent of the barotropic solve: may run on another stream, next to launch_momentum)
int launch_velc1(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_velc1<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  return 1;
}
// after both: total velocity, relaxation, vertical velocity
int launch_velc2(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  k_velc2<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  k_w<<<grid2(v, v.I * v.J, b), b, 0, s>>>(v);
  return 2;
}
void launch_global_means(const Dev &v, double *out, cudaStream_t s) { k_global_means<<<dim3(v.M, v.L), 256, 0, s>>>(v, out); }
void launch_health(const Dev &v, int *flags, cudaStream_t s) { k_health<<<v.M, 256, 0, s>>>(v, flags); }

}  // namespace cg
