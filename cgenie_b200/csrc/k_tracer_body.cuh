// k_tracer_body.cuh -- GOLDSTEIN tracer timestep kernels (tstepo_flux + co), sm_100a.
//
// Reference: src/goldstein/goldstein.f90:2436-2642 (tstepo_flux), :2657-2777 (co),
//            :3048-3082 (eos, eosd).
//
// Included by two translation units:
//   k_tracer_strict.cu  (CG_TRACER_FAST=0, nvcc -fmad=false): every expression in the reference's
//                       operation order -> bit-identical to the CPU oracle;
//   k_tracer_fast.cu    (CG_TRACER_FAST=1, FMA on): face fluxes as 2-term stencils with per-cell
//                       coefficients hoisted out of the tracer loop and the isoneutral sum factored,
//                       ~4x fewer fp64 instructions per tracer-cell (<=1e-10 relative per step).
//
// Parallelisation: thread = (member m, column i,j, tracer chunk); the thread marches k upward
// carrying the bottom-face flux fb(l) in registers exactly as the Fortran carries fb(l,i,j);
// the west/south face fluxes are recomputed from the neighbour side with the same expression
// the Fortran stored (fw = fe(i-1), fs = fn(j-1)), so the result is the sequential one.
#pragma once
#include "cg_device.cuh"

namespace cg {

// `static __constant__ GridC c_g;` is defined by the including translation unit

#if CG_TRACER_FAST
#define CG_KNAME(x) x##_fast
#else
#define CG_KNAME(x) x##_strict
#endif

// thread-block shape: x = members (MX lanes), y = cells along i, z = cells along j
template <int LC, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) CG_KNAME(k_tstepo_flux)(const Dev v, const int mx, const int ci) {
  const int I = v.I, J = v.J, K = v.K, L = v.L, MS = v.MS;
  const int nmg = (v.M + mx - 1) / mx;
  const int mg = blockIdx.x % nmg, it = blockIdx.x / nmg;
  const int m = mg * mx + threadIdx.x;
  const int i = it * ci + threadIdx.y + 1;
  const int j = blockIdx.y * blockDim.z + threadIdx.z + 1;
  const int l0 = blockIdx.z * LC;
  if (m >= v.M || i > I || j > J) return;
  const int k1c = CG_K1(v, i, j);
  if (k1c > K) return;
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CG_K1(v, ip, j), k1w = CG_K1(v, im, j), k1n = CG_K1(v, i, j + 1), k1s = CG_K1(v, i, j - 1);

  const double diff1 = v.p.diff1[m];
  double diffv = v.p.diff2[m];
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  const double scc = ec2;
  const double dt = c_g.dt, dphi = c_g.dphi, rdphi = c_g.rdphi;
  const double rc = c_g.rc[j], rc2 = c_g.rc2[j], cvj = c_g.cv[j], cvjm = c_g.cv[j - 1], cv2j = c_g.cv2[j],
               cv2jm = (j > 1) ? c_g.cv2[j - 1] : 0.0, rdsj = c_g.rds[j];
  const double dsvN = c_g.dsv[(j < J - 1) ? j : J - 1], dsvS = c_g.dsv[(j - 1 < J - 1) ? ((j > 1) ? j - 1 : 1) : J - 1];
  const double rdsvN = c_g.rdsv[j], rdsvS = (j > 1) ? c_g.rdsv[j - 1] : 0.0;

  const double *__restrict__ ts1 = v.ts_cur;
  double *__restrict__ tsn = v.ts_new;
  const double *__restrict__ uu = v.u;

  const size_t sL = (size_t)MS;             // tracer stride
  const size_t sC = (size_t)L * MS;         // cell stride
  const size_t sK = (size_t)I * J * sC;     // level stride
  // column base offsets (level k=1) of the five stencil columns, at tracer 0, member m
  const size_t oC = cell3(I, J, i, j, 1) * sC + m;
  const size_t oE = cell3(I, J, ip, j, 1) * sC + m;
  const size_t oW = cell3(I, J, im, j, 1) * sC + m;
  const size_t oN = (j < J) ? cell3(I, J, i, j + 1, 1) * sC + m : oC;
  const size_t oS = (j > 1) ? cell3(I, J, i, j - 1, 1) * sC + m : oC;
  const size_t uC = cell3(I, J, i, j, 1) * 3 * MS + m, uWo = cell3(I, J, im, j, 1) * 3 * MS + m,
               uSo = (j > 1) ? cell3(I, J, i, j - 1, 1) * 3 * MS + m : uC;
  const size_t uK = (size_t)I * J * 3 * MS;

  double fb[LC];
#pragma unroll
  for (int q = 0; q < LC; q++) fb[q] = 0.0;

  for (int k = k1c; k <= K; k++) {
    const size_t ko = (size_t)(k - 1) * sK, kuo = (size_t)(k - 1) * uK;
    const bool topl = (k == K);
    // face openness at this level (k >= k1c holds)
    const bool opE = k >= k1e, opW = k >= k1w, opN = k >= k1n, opS = (j > 1) && (k >= k1s);
    // isoneutral masks one level up
    const bool upE = !topl && (k + 1 >= k1e), upW = !topl && (k + 1 >= k1w), upN = !topl && (k + 1 >= k1n),
               upS = !topl && (k + 1 >= k1s);
    // velocities and upstream weights (goldstein.f90:2517-2523)
    const double uE = uu[uC + kuo], vN = uu[uC + kuo + MS], ww = topl ? 0.0 : uu[uC + kuo + 2 * MS];
    const double uW = uu[uWo + kuo], vS = (j > 1) ? uu[uSo + kuo + MS] : 0.0;
    double pec = uE * dphi / diff1;
    const double upsE = pec / (2.0 + fabs(pec));
    pec = uW * dphi / diff1;
    const double upsW = pec / (2.0 + fabs(pec));
    pec = vN * dsvN / diff1;
    const double upsN = pec / (2.0 + fabs(pec));
    pec = vS * dsvS / diff1;
    const double upsS = pec / (2.0 + fabs(pec));
    const double rdza = topl ? 0.0 : c_g.rdza[k], rdzk = c_g.rdz[k];

    // ---- density slopes from T,S (goldstein.f90:2490-2500, 2561-2603)
    bool iso = false;
    double dzrho = 0.0, slim = 1.0;
    double dxr[4], dyr[4];
    if (!topl) {
      const double t0 = ts1[oC + ko], t1 = ts1[oC + ko + sK], s0 = ts1[oC + ko + sL], s1 = ts1[oC + ko + sK + sL];
      const double tatw = 0.5 * (t0 + t1);
      const double tec = v.ieos ? -ec1 - ec3 * tatw * 2 - ec4 * tatw * tatw * 3 - v.p.ec5[m] * c_g.zw[k]   // eosd, ieos = 1
                                : -ec1 - ec3 * tatw * 2 - ec4 * tatw * tatw * 3;
      dzrho = (ec2 * (s1 - s0) - tec * (t1 - t0)) * rdza;
      if (dzrho < -1.0e-12) {
        iso = true;
        const double rdzrho = 1.0 / dzrho;
        double tv1 = 0.0;
#pragma unroll
        for (int knp = 0; knp <= 1; knp++) {
          const size_t kk = ko + (size_t)knp * sK;
          const double tc = knp ? t1 : t0, sc = knp ? s1 : s0;
#pragma unroll
          for (int nnp = 0; nnp <= 1; nnp++) {
            const int a = nnp + 2 * knp;
            const bool mx_ = nnp ? (knp ? upE : opE) : (knp ? upW : opW);
            const bool my_ = nnp ? (knp ? upN : opN) : (knp ? upS : (k >= k1s));
            double dxt = 0.0, dxs = 0.0, dyt = 0.0, dys = 0.0;
            if (mx_) {
              const size_t on = (nnp ? oE : oW) + kk;
              const double tn = ts1[on], sn = ts1[on + sL];
              dxt = (nnp ? (tn - tc) : (tc - tn)) * rc * rdphi;
              dxs = (nnp ? (sn - sc) : (sc - sn)) * rc * rdphi;
            }
            if (my_) {
              const size_t on = (nnp ? oN : oS) + kk;
              const double tn = ts1[on], sn = ts1[on + sL];
              dyt = (nnp ? (tn - tc) : (tc - tn)) * (nnp ? cvj : cvjm) * (nnp ? rdsvN : rdsvS);
              dys = (nnp ? (sn - sc) : (sc - sn)) * (nnp ? cvj : cvjm) * (nnp ? rdsvN : rdsvS);
            }
            dxr[a] = scc * dxs - tec * dxt;
            dyr[a] = scc * dys - tec * dyt;
            tv1 = tv1 + dxr[a] * dxr[a] + dyr[a] * dyr[a];
          }
        }
        tv1 = 0.25 * tv1 * rdzrho * rdzrho;
        const double ssm = c_g.ssmax[k];
        if (tv1 > ssm) slim = ssm * ssm / (tv1 * tv1);
      }
      if (v.iediff) {   // local vertical diffusivity, goldstein.f90:2496-2515 (ediff1(i,j,k) = ediff1p(k))
        const double rdzrho = iso ? 1.0 / dzrho : -1.0e12;
        const double e0 = v.p.ediff0[m], e1 = v.p.ediff1p[(size_t)k * MS + m];
        if (v.ediffpow2i == 0) diffv = e0 + e1;
        else if (v.ediffpow2i == 1) diffv = e0 + e1 * (-rdzrho);
        else if (v.ediffpow2i == 2) diffv = e0 + e1 * sqrt(-rdzrho);
        else diffv = e0 + e1 * pow(-rdzrho, v.ediffpow2);
        if (diffv > c_g.diffmax[k + 1]) diffv = c_g.diffmax[k + 1];
      }
    }
    // (the reference keeps the last diffv where it computes none -- at k = maxk, where w = 0 and dza = 0 make pec = 0 for any
    // diffusivity and the flux through the top face is the boundary condition)
    pec = ww * c_g.dza[k] / diffv;
    const double upsA = pec / (2.0 + fabs(pec));

#if CG_TRACER_FAST
    // ---- per-cell stencil coefficients (hoisted out of the tracer loop)
    // east/west/north/south/above faces: flux = A * ts(neighbour) + B * ts(centre)
    const double hE = uE * rc * 0.5, dE = rc2 * diff1;
    const double aE = opE ? (hE * (1.0 - upsE) - dE) : 0.0, bE = opE ? (hE * (1.0 + upsE) + dE) : 0.0;
    const double hW = uW * rc * 0.5;
    // west face seen from the west cell: fw = hW*((1-ups)*c + (1+ups)*W) - (c - W)*dE
    const double aW = opW ? (hW * (1.0 + upsW) + dE) : 0.0, bW = opW ? (hW * (1.0 - upsW) - dE) : 0.0;
    const double hN = cvj * vN * 0.5, dN = cv2j * diff1;
    const double aN = opN ? (hN * (1.0 - upsN) - dN) : 0.0, bN = opN ? (hN * (1.0 + upsN) + dN) : 0.0;
    const double hS = cvjm * vS * 0.5, dS = cv2jm * diff1;
    const double aS = opS ? (hS * (1.0 + upsS) + dS) : 0.0, bS = opS ? (hS * (1.0 - upsS) - dS) : 0.0;
    const double hA = ww * 0.5, dA = rdza * diffv;
    double aA = hA * (1.0 - upsA) - dA, bA = hA * (1.0 + upsA) + dA;   // fa = aA*c1 + bA*c0 (k<K)
    // isoneutral: tv = 2 dzrho Sum(dx_a wx_a + dy_a wy_a) - dzts * S2 ; fa += cf * tv
    double wx[4], wy[4], cf = 0.0, s2 = 0.0;
    if (iso) {
      cf = 0.25 * slim * diff1 / (dzrho * dzrho);
      const double gx = rc * rdphi * 2.0 * dzrho * cf;
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const bool nn = a & 1;
        wx[a] = dxr[a] * gx;
        wy[a] = dyr[a] * (nn ? cvj * rdsvN : cvjm * rdsvS) * 2.0 * dzrho * cf;
        s2 += dxr[a] * dxr[a] + dyr[a] * dyr[a];
      }
      s2 = s2 * cf * rdza;
      // fold the centre-column part of dzts*S2 into the vertical coefficients
      aA -= s2;
      bA += s2;
    }
    const double cX = dt * rdphi, cY = dt * rdsj, cZ = dt * rdzk;
#endif

    // ---- tracer loop
    double tnew = 0.0, snew = 0.0;
#pragma unroll
    for (int q = 0; q < LC; q++) {
      const int l = l0 + q;
      if (l < L) {
        const size_t lo = ko + (size_t)l * sL;
        const double c0 = ts1[oC + lo];
        const double E0 = opE ? ts1[oE + lo] : 0.0, W0 = opW ? ts1[oW + lo] : 0.0;
        const double N0 = opN ? ts1[oN + lo] : 0.0, S0 = (opS || (iso && k >= k1s)) ? ts1[oS + lo] : 0.0;
        double c1 = 0.0, E1 = 0.0, W1 = 0.0, N1 = 0.0, S1 = 0.0;
        if (!topl) {
          c1 = ts1[oC + lo + sK];
          if (iso) {
            if (upE) E1 = ts1[oE + lo + sK];
            if (upW) W1 = ts1[oW + lo + sK];
            if (upN) N1 = ts1[oN + lo + sK];
            if (upS) S1 = ts1[oS + lo + sK];
          }
        }
#if CG_TRACER_FAST
        const double fe = aE * E0 + bE * c0;
        const double fw = aW * W0 + bW * c0;
        const double fn = aN * N0 + bN * c0;
        const double fs = aS * S0 + bS * c0;
        double fa;
        if (topl) {
          fa = (l < 2) ? v.tsflux[((size_t)l * I * J + cell2(I, i, j)) * MS + m] : 0.0;
        } else {
          fa = aA * c1 + bA * c0;
          if (iso) {
            double acc = 0.0;
            acc += (opW ? (c0 - W0) : 0.0) * wx[0];
            acc += (opE ? (E0 - c0) : 0.0) * wx[1];
            acc += (upW ? (c1 - W1) : 0.0) * wx[2];
            acc += (upE ? (E1 - c1) : 0.0) * wx[3];
            acc += ((k >= k1s) ? (c0 - S0) : 0.0) * wy[0];
            acc += (opN ? (N0 - c0) : 0.0) * wy[1];
            acc += (upS ? (c1 - S1) : 0.0) * wy[2];
            acc += (upN ? (N1 - c1) : 0.0) * wy[3];
            fa += acc;
          }
        }
        const double tn = c0 - ((fe - fw) * cX + (fn - fs) * cY + (fa - fb[q]) * cZ);
#else
        // east / west (goldstein.f90:2526-2537, 2476-2486, 2633)
        double fe = 0.0, fw = 0.0, fn = 0.0, fs = 0.0, fa;
        if (opE) {
          fe = uE * rc * ((1.0 - upsE) * E0 + (1.0 + upsE) * c0) * 0.5;
          fe = fe - (E0 - c0) * rc2 * diff1;
        }
        if (opW) {
          fw = uW * rc * ((1.0 - upsW) * c0 + (1.0 + upsW) * W0) * 0.5;
          fw = fw - (c0 - W0) * rc2 * diff1;
        }
        // north / south (goldstein.f90:2539-2547, 2634)
        if (opN) {
          fn = cvj * vN * ((1.0 - upsN) * N0 + (1.0 + upsN) * c0) * 0.5;
          fn = fn - cv2j * (N0 - c0) * diff1;
        }
        if (opS) {
          fs = cvjm * vS * ((1.0 - upsS) * c0 + (1.0 + upsS) * S0) * 0.5;
          fs = fs - cv2jm * (c0 - S0) * diff1;
        }
        // above (goldstein.f90:2549-2559)
        if (topl) {
          fa = (l < 2) ? v.tsflux[((size_t)l * I * J + cell2(I, i, j)) * MS + m] : 0.0;
        } else {
          fa = ww * ((1.0 - upsA) * c1 + (1.0 + upsA) * c0) * 0.5;
          fa = fa - (c1 - c0) * rdza * diffv;
          if (iso) {  // goldstein.f90:2609-2621
            const double dzts = (c1 - c0) * rdza;
            double dxt[4], dyt[4];
            dxt[0] = opW ? (c0 - W0) * rc * rdphi : 0.0;
            dxt[1] = opE ? (E0 - c0) * rc * rdphi : 0.0;
            dxt[2] = upW ? (c1 - W1) * rc * rdphi : 0.0;
            dxt[3] = upE ? (E1 - c1) * rc * rdphi : 0.0;
            dyt[0] = (k >= k1s) ? (c0 - S0) * cvjm * rdsvS : 0.0;
            dyt[1] = opN ? (N0 - c0) * cvj * rdsvN : 0.0;
            dyt[2] = upS ? (c1 - S1) * cvjm * rdsvS : 0.0;
            dyt[3] = upN ? (N1 - c1) * cvj * rdsvN : 0.0;
            double tv = 0.0;
#pragma unroll
            for (int a = 0; a < 4; a++)
              tv = tv + (2 * dzrho * dxt[a] - dxr[a] * dzts) * dxr[a] + (2 * dzrho * dyt[a] - dyr[a] * dzts) * dyr[a];
            tv = 0.25 * slim * diff1 * tv / (dzrho * dzrho);
            fa = fa + tv;
          }
        }
        // update (goldstein.f90:2627-2632)
        const double tn = c0 - dt * ((fe - fw) * rdphi + (fn - fs) * rdsj + (fa - fb[q]) * rdzk);
#endif
        tsn[oC + lo] = tn;
        fb[q] = fa;
        if (l == 0) tnew = tn;
        if (l == 1) snew = tn;
      }
    }
    // density of the new state (goldstein.f90:2638)
    if (l0 == 0)
      v.rho[cell3(I, J, i, j, k) * MS + m] =
          v.ieos ? ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew) + v.p.ec5[m] * tnew * c_g.zro[k]
                 : ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);
  }
}


#if CG_TRACER_FAST
// ---------------------------------------------------------------------------------------------
// Tracer step, cooperative variant (fast path for L > 2).
// One block = one wet column x one 32-member tile; its NW warps split the tracers (LW each).  Per
// level the per-cell work that does not depend on the tracer (upstream weights, density slopes,
// slope limiter) is computed ONCE per block: warp w evaluates one of the four slope stencils and one
// face, publishes it through shared memory (double-buffered by level parity, one __syncthreads per
// level), and every warp folds the result into 15 linear coefficients so that the update of one
// tracer-cell is 18 fp64 operations:
//     H  = hC*c0 + hE*E0 + hW*W0 + hN*N0 + hS*S0                       (horizontal flux divergence * dt)
//     fa = sum over the ten stencil values (level k and k+1) of pA[.] * value   (flux through the top face)
//     ts = c0 - (H + (fa - fb) * cZ)
// Level-(k+1) values are kept in registers and become the level-k values of the next iteration, so
// every ts1 value is loaded once per block that needs it.  All offsets are 32-bit.
template <int LW, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) k_tstepo_flux_coop(const Dev v) {
  __shared__ double sh[2][22][32];
  const int I = v.I, J = v.J, K = v.K, L = v.L;
  const unsigned MS = v.MS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x % v.nwet, tile = blockIdx.x / v.nwet;
  const unsigned m = tile * 32 + lane;
  const int c2 = v.wetcols[col];
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CG_K1(v, ip, j), k1w = CG_K1(v, im, j), k1n = CG_K1(v, i, j + 1), k1s = CG_K1(v, i, j - 1);
  const int l0 = w * LW;

  const double diff1 = v.p.diff1[m], diffv = v.p.diff2[m], rdiff1 = 1.0 / diff1, rdiffv = 1.0 / diffv;
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  const double dt = c_g.dt, dphi = c_g.dphi, rdphi = c_g.rdphi;
  const double rc = c_g.rc[j], rc2 = c_g.rc2[j], cvj = c_g.cv[j], cvjm = c_g.cv[j - 1], cv2j = c_g.cv2[j],
               cv2jm = (j > 1) ? c_g.cv2[j - 1] : 0.0;
  const double dsvN = c_g.dsv[(j < J - 1) ? j : J - 1], dsvS = (j > 1) ? c_g.dsv[j - 1] : 0.0;
  const double gyN = cvj * c_g.rdsv[j], gyS = (j > 1) ? cvjm * c_g.rdsv[j - 1] : 0.0, gxx = rc * rdphi;
  const double cX = dt * rdphi, cY = dt * c_g.rds[j];

  const double *__restrict__ ts1 = v.ts_cur;
  double *__restrict__ tsn = v.ts_new;
  const double *__restrict__ uu = v.u;
  const unsigned sL = MS, sC = (unsigned)L * MS, sK = (unsigned)(I * J) * sC, uK = (unsigned)(I * J) * 3u * MS;
  // element offsets of the five stencil columns at level k1c, tracer 0
  const unsigned kb = (unsigned)(k1c - 1);
  unsigned oC = (unsigned)cell3(I, J, i, j, 1) * sC + m + kb * sK;
  unsigned oE = (unsigned)cell3(I, J, ip, j, 1) * sC + m + kb * sK;
  unsigned oW = (unsigned)cell3(I, J, im, j, 1) * sC + m + kb * sK;
  unsigned oN = ((j < J) ? (unsigned)cell3(I, J, i, j + 1, 1) * sC + m : oC - kb * sK) + kb * sK;
  unsigned oS = ((j > 1) ? (unsigned)cell3(I, J, i, j - 1, 1) * sC + m : oC - kb * sK) + kb * sK;
  unsigned uC = (unsigned)cell3(I, J, i, j, 1) * 3u * MS + m + kb * uK;
  unsigned uW = (unsigned)cell3(I, J, im, j, 1) * 3u * MS + m + kb * uK;
  unsigned uS = ((j > 1) ? (unsigned)cell3(I, J, i, j - 1, 1) * 3u * MS + m : uC - kb * uK) + kb * uK;
  unsigned oR = (unsigned)cell3(I, J, i, j, 1) * MS + m + kb * (unsigned)(I * J) * MS;

  // level-k stencil values of this warp's tracers (registers), loaded once at the bottom
  double vc[LW], vE[LW], vW[LW], vN[LW], vS[LW], fb[LW];
  {
    const bool opE = k1c >= k1e, opW = k1c >= k1w, opN = k1c >= k1n, opS = k1c >= k1s;
#pragma unroll
    for (int q = 0; q < LW; q++) {
      const unsigned lo = (unsigned)(l0 + q) * sL;
      const bool on = l0 + q < L;
      vc[q] = on ? ts1[oC + lo] : 0.0;
      vE[q] = (on && opE) ? ts1[oE + lo] : 0.0;
      vW[q] = (on && opW) ? ts1[oW + lo] : 0.0;
      vN[q] = (on && opN) ? ts1[oN + lo] : 0.0;
      vS[q] = (on && opS) ? ts1[oS + lo] : 0.0;
      fb[q] = 0.0;
    }
  }
  // T,S of the centre column at level k (every warp needs them for the slopes)
  double t0 = ts1[oC], s0 = ts1[oC + sL];

  for (int k = k1c; k <= K; k++) {
    const int b = k & 1;
    const bool topl = (k == K);
    const bool opE = k >= k1e, opW = k >= k1w, opN = k >= k1n, opS = k >= k1s;
    const bool upE = !topl && (k + 1 >= k1e), upW = !topl && (k + 1 >= k1w), upN = !topl && (k + 1 >= k1n),
               upS = !topl && (k + 1 >= k1s);
    const double rdza = topl ? 0.0 : c_g.rdza[k];
    double t1 = 0.0, s1 = 0.0, tec = 0.0, dzrho = 0.0;
    if (!topl) {
      t1 = ts1[oC + sK];
      s1 = ts1[oC + sK + sL];
      const double tatw = 0.5 * (t0 + t1);
      tec = -ec1 - ec3 * tatw * 2 - ec4 * tatw * tatw * 3;
      dzrho = (ec2 * (s1 - s0) - tec * (t1 - t0)) * rdza;
    }
    // ---- phase A: this warp's share of the per-cell work
    if (w < 4) {
      // slope stencil a = w: nnp = w & 1 (0: west/south side, 1: east/north side), knp = w >> 1 (level k or k+1)
      double dxr = 0.0, dyr = 0.0;
      if (!topl) {
        const int nnp = w & 1, knp = w >> 1;
        const bool mx_ = nnp ? (knp ? upE : opE) : (knp ? upW : opW);
        const bool my_ = nnp ? (knp ? upN : opN) : (knp ? upS : opS);
        const unsigned kk = knp ? sK : 0u;
        const double tc = knp ? t1 : t0, sc = knp ? s1 : s0;
        if (mx_) {
          const unsigned on = (nnp ? oE : oW) + kk;
          const double tn = ts1[on], sn = ts1[on + sL];
          const double dxt = (nnp ? (tn - tc) : (tc - tn)) * gxx, dxs = (nnp ? (sn - sc) : (sc - sn)) * gxx;
          dxr = ec2 * dxs - tec * dxt;
        }
        if (my_) {
          const unsigned on = (nnp ? oN : oS) + kk;
          const double tn = ts1[on], sn = ts1[on + sL];
          const double gy = nnp ? gyN : gyS;
          const double dyt = (nnp ? (tn - tc) : (tc - tn)) * gy, dys = (nnp ? (sn - sc) : (sc - sn)) * gy;
          dyr = ec2 * dys - tec * dyt;
        }
      }
      sh[b][w][lane] = dxr;
      sh[b][4 + w][lane] = dyr;
      // one face per warp: flux = a * ts(neighbour) + bb * ts(centre)
      double a = 0.0, bb = 0.0;
      if (w == 0) {
        if (opE) {
          const double uE = uu[uC], pec = uE * dphi * rdiff1, ups = pec / (2.0 + fabs(pec)), h = uE * rc * 0.5, d = rc2 * diff1;
          a = h * (1.0 - ups) - d;
          bb = h * (1.0 + ups) + d;
        }
      } else if (w == 1) {
        if (opW) {
          const double uWv = uu[uW], pec = uWv * dphi * rdiff1, ups = pec / (2.0 + fabs(pec)), h = uWv * rc * 0.5, d = rc2 * diff1;
          a = h * (1.0 + ups) + d;   // coefficient of W0
          bb = h * (1.0 - ups) - d;  // coefficient of c0
        }
      } else if (w == 2) {
        if (opN) {
          const double vNv = uu[uC + MS], pec = vNv * dsvN * rdiff1, ups = pec / (2.0 + fabs(pec)), h = cvj * vNv * 0.5, d = cv2j * diff1;
          a = h * (1.0 - ups) - d;
          bb = h * (1.0 + ups) + d;
        }
      } else {
        if (opS) {
          const double vSv = uu[uS + MS], pec = vSv * dsvS * rdiff1, ups = pec / (2.0 + fabs(pec)), h = cvjm * vSv * 0.5, d = cv2jm * diff1;
          a = h * (1.0 + ups) + d;   // coefficient of S0
          bb = h * (1.0 - ups) - d;  // coefficient of c0
        }
      }
      sh[b][8 + 2 * w][lane] = a;
      sh[b][9 + 2 * w][lane] = bb;
    }
    if (w == NW - 1) {  // vertical face
      double aA = 0.0, bA = 0.0;
      if (!topl) {
        const double ww = uu[uC + 2 * MS], pec = ww * c_g.dza[k] * rdiffv, ups = pec / (2.0 + fabs(pec)), h = ww * 0.5, d = rdza * diffv;
        aA = h * (1.0 - ups) - d;
        bA = h * (1.0 + ups) + d;
      }
      sh[b][16][lane] = aA;
      sh[b][17][lane] = bA;
    }
    __syncthreads();
    // ---- phase B: fold into the 15 linear coefficients (each warp, redundantly: ~60 flops, 1-2 divisions)
    const double aE = sh[b][8][lane], bE = sh[b][9][lane], aW = sh[b][10][lane], bW = sh[b][11][lane];
    const double aN = sh[b][12][lane], bN = sh[b][13][lane], aS = sh[b][14][lane], bS = sh[b][15][lane];
    const double hE = aE * cX, hW = -aW * cX, hN = aN * cY, hS = -aS * cY, hC = (bE - bW) * cX + (bN - bS) * cY;
    double pc0 = sh[b][17][lane], pc1 = sh[b][16][lane];
    double pW0 = 0.0, pE0 = 0.0, pS0 = 0.0, pN0 = 0.0, pW1 = 0.0, pE1 = 0.0, pS1 = 0.0, pN1 = 0.0;
    if (dzrho < -1.0e-12) {
      const double x0 = sh[b][0][lane], x1 = sh[b][1][lane], x2 = sh[b][2][lane], x3 = sh[b][3][lane];
      const double y0 = sh[b][4][lane], y1 = sh[b][5][lane], y2 = sh[b][6][lane], y3 = sh[b][7][lane];
      const double tv1 = (((x0 * x0 + y0 * y0) + (x1 * x1 + y1 * y1)) + (x2 * x2 + y2 * y2)) + (x3 * x3 + y3 * y3);
      const double rdz = 1.0 / dzrho, rdz2 = rdz * rdz;
      const double sl = 0.25 * tv1 * rdz2, ssm = c_g.ssmax[k];
      const double slim = (sl > ssm) ? ssm * ssm / (sl * sl) : 1.0;
      const double cf = 0.25 * slim * diff1 * rdz2;
      const double g2 = 2.0 * dzrho * cf, gX = g2 * gxx, gS = g2 * gyS, gN = g2 * gyN;
      const double s2 = tv1 * cf * rdza;
      // fa += (c0-W0)*wx0 + (E0-c0)*wx1 + (c1-W1)*wx2 + (E1-c1)*wx3 + (c0-S0)*wy0 + (N0-c0)*wy1 + (c1-S1)*wy2 + (N1-c1)*wy3 - (c1-c0)*s2
      const double wx0 = x0 * gX, wx1 = x1 * gX, wx2 = x2 * gX, wx3 = x3 * gX;
      const double wy0 = y0 * gS, wy1 = y1 * gN, wy2 = y2 * gS, wy3 = y3 * gN;
      pc0 += (wx0 - wx1) + (wy0 - wy1) + s2;
      pc1 += (wx2 - wx3) + (wy2 - wy3) - s2;
      pW0 = -wx0; pE0 = wx1; pS0 = -wy0; pN0 = wy1; pW1 = -wx2; pE1 = wx3; pS1 = -wy2; pN1 = wy3;
    }
    const double cZ = dt * c_g.rdz[k];
    // ---- phase C: tracers of this warp
    double tnew = 0.0, snew = 0.0;
#pragma unroll
    for (int q = 0; q < LW; q++) {
      const int l = l0 + q;
      if (l < L) {
        const unsigned lo = (unsigned)l * sL;
        const double c0 = vc[q], E0 = vE[q], W0 = vW[q], N0 = vN[q], S0 = vS[q];
        double c1 = 0.0, E1 = 0.0, W1 = 0.0, N1 = 0.0, S1 = 0.0, fa;
        if (!topl) {
          c1 = ts1[oC + sK + lo];
          if (upE) E1 = ts1[oE + sK + lo];
          if (upW) W1 = ts1[oW + sK + lo];
          if (upN) N1 = ts1[oN + sK + lo];
          if (upS) S1 = ts1[oS + sK + lo];
          fa = pc0 * c0 + pc1 * c1 + pW0 * W0 + pE0 * E0 + pS0 * S0 + pN0 * N0 + pW1 * W1 + pE1 * E1 + pS1 * S1 + pN1 * N1;
        } else {
          fa = (l < 2) ? v.tsflux[((unsigned)l * (unsigned)(I * J) + (unsigned)c2) * MS + m] : 0.0;
        }
        const double H = hC * c0 + hE * E0 + hW * W0 + hN * N0 + hS * S0;
        const double tn = c0 - (H + (fa - fb[q]) * cZ);
        tsn[oC + lo] = tn;
        fb[q] = fa;
        vc[q] = c1; vE[q] = E1; vW[q] = W1; vN[q] = N1; vS[q] = S1;
        if (l == 0) tnew = tn;
        if (l == 1) snew = tn;
      }
    }
    if (w == 0) v.rho[oR] = ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);
    t0 = t1; s0 = s1;
    oC += sK; oE += sK; oW += sK; oN += sK; oS += sK; uC += uK; uW += uK; uS += uK; oR += (unsigned)(I * J) * MS;
  }
}
#endif  // CG_TRACER_FAST

// Convective adjustment, goldstein.f90:2657-2777 (iconv == 0, ieos == 0).
// thread = (member, column); the index array k(0:maxk) and dzm live in local memory.
__global__ void __launch_bounds__(128) CG_KNAME(k_co)(const Dev v) {
  const int I = v.I, J = v.J, K = v.K, L = v.L, MS = v.MS;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int c2 = blockIdx.y * blockDim.y + threadIdx.y;  // 0-based column index
  if (m >= v.M || c2 >= I * J) return;
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  if (k1c > K) return;
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  double *__restrict__ ts = v.ts_new;
  double *__restrict__ rho = v.rho;
  const size_t sL = (size_t)MS, sC = (size_t)L * MS, sK = (size_t)I * J * sC;
  const size_t oC = cell3(I, J, i, j, 1) * sC + m;
  const size_t rK = (size_t)I * J * MS, rC = cell3(I, J, i, j, 1) * MS + m;
#define RHOK(k) rho[rC + (size_t)((k)-1) * rK]
#define TSK(l, k) ts[oC + (size_t)((k)-1) * sK + (size_t)(l) * sL]
  int kk[kMaxK + 2];
  double dzm[kMaxK + 2], rl[kMaxK + 2];
  kk[k1c - 1] = 0;
  rl[0] = 0.0;
  for (int q = k1c; q <= K; q++) {
    kk[q] = q;
    dzm[q] = c_g.dz[q];
    rl[q] = RHOK(q);  // whole density column up front: independent loads, then register/local work only
  }
  int mm = K, lastmix = 0;
  bool any = false;
  // ieos = 1 (:2692-2698, 2714-2730): the two boxes of a comparison are brought to one depth -- the interface above box k(m-1) --
  // before every test; the stored column is made vertically local again at the end (tstepo :2396-2408)
  const double ec5 = v.ieos ? v.p.ec5[m] : 0.0;
  auto eosz = [&](const int lev, const double z) {
    const double t = TSK(0, lev), s_ = TSK(1, lev);
    return ec1 * t + ec2 * s_ + ec3 * (t * t) + ec4 * (t * t * t) + ec5 * t * z;
  };
  // iconv = 1, Mueller convection scheme: coshuffle first (:2667-2672, SUBROUTINE coshuffle :2781-2841) -- the surface box sinks to
  // the level of its own density, the boxes it passes move up by its thickness -- on this column, then the adjustment below
  int icosd = 0, icond = 0;
  if (v.iconv == 1) {
    int k0 = 0, ipass = 0;
    while (k0 < K && ipass < K) {
      ipass = ipass + 1;
      int k = K - 1;
      if (v.ieos && k >= k1c) { rl[K] = eosz(K, c_g.zw[k]); rl[k] = eosz(k, c_g.zw[k]); }
      while (k >= k1c && rl[K] > rl[k]) {
        k = k - 1;
        if (v.ieos && k >= k1c) { rl[K] = eosz(K, c_g.zw[k]); rl[k] = eosz(k, c_g.zw[k]); }
      }
      k0 = k + 1;
      if (k0 < K) {
        const double dzK = c_g.dz[K];
        for (int l = 0; l < L; l++) {
          const double tv_temp = TSK(l, K);
          for (int q = K; q >= k0 + 1; q--) TSK(l, q) = ((c_g.dz[q] - dzK) * TSK(l, q) + dzK * TSK(l, q - 1)) * c_g.rdz[q];
          TSK(l, k0) = ((c_g.dz[k0] - dzK) * TSK(l, k0) + dzK * tv_temp) * c_g.rdz[k0];
        }
        for (int q = k0; q <= K; q++) { rl[q] = eosz(q, c_g.zw[K - 1]); RHOK(q) = rl[q]; }
        if (K - k0 > icosd) icosd = K - k0;
      }
    }
  }
  while (kk[mm - 1] > 0 || (lastmix != 0 && kk[mm] != K)) {
    if (v.ieos && kk[mm - 1] > 0) {
      const double z = c_g.zw[kk[mm - 1]];
      rl[kk[mm]] = eosz(kk[mm], z);
      rl[kk[mm - 1]] = eosz(kk[mm - 1], z);
    }
    if (kk[mm - 1] == 0 || rl[kk[mm]] < rl[kk[mm - 1]]) {
      if (lastmix == 0 || kk[mm] == K) mm = mm - 1; else mm = mm + 1;
      lastmix = 0;
    } else {
      lastmix = 1;
      any = true;
      int n = mm - 1;
      if (v.ieos) {
        if (kk[n - 1] > 0) { const double z = c_g.zw[kk[n - 1]]; rl[kk[n]] = eosz(kk[n], z); rl[kk[n - 1]] = eosz(kk[n - 1], z); }
        while (kk[n - 1] > 0 && rl[kk[n]] >= rl[kk[n - 1]]) {
          n = n - 1;
          if (kk[n - 1] > 0) { const double z = c_g.zw[kk[n - 1]]; rl[kk[n]] = eosz(kk[n], z); rl[kk[n - 1]] = eosz(kk[n - 1], z); }
        }
      } else
      while (kk[n - 1] > 0 && rl[kk[n]] >= rl[kk[n - 1]]) n = n - 1;
      // thickness-weighted mix of all tracers over index entries n..mm (:2732-2737)
      double dznew = dzm[kk[mm]];
      for (int ni = 1; ni <= mm - n; ni++) dznew = dznew + dzm[kk[mm - ni]];
      double tmix = 0.0, smix = 0.0;
      for (int l = 0; l < L; l++) {
        double sum = TSK(l, kk[mm]) * dzm[kk[mm]];
        for (int ni = 1; ni <= mm - n; ni++) sum = sum + TSK(l, kk[mm - ni]) * dzm[kk[mm - ni]];
        const double val = sum / dznew;
        TSK(l, kk[mm]) = val;
        if (l == 0) tmix = val;
        if (l == 1) smix = val;
      }
      dzm[kk[mm]] = dznew;
      rl[kk[mm]] = ec1 * tmix + ec2 * smix + ec3 * (tmix * tmix) + ec4 * (tmix * tmix * tmix);
      RHOK(kk[mm]) = rl[kk[mm]];
      int ni = mm - 1;
      while (kk[ni + 1] > 0) {
        kk[ni] = kk[ni - mm + n];
        ni = ni - 1;
      }
    }
  }
  if (any) {
    // fill in T,S values in mixed regions (:2749-2764)
    int mq = K - 1;
    double cnt = 0.0;
    for (int n = K - 1; n >= k1c; n--) {
      if (n > kk[mq]) {
        double tmix = 0.0, smix = 0.0;
        for (int l = 0; l < L; l++) {
          const double val = TSK(l, kk[mq + 1]);
          TSK(l, n) = val;
          if (l == 0) tmix = val;
          if (l == 1) smix = val;
        }
        RHOK(n) = ec1 * tmix + ec2 * smix + ec3 * (tmix * tmix) + ec4 * (tmix * tmix * tmix);
        if (v.iconv == 1) icond = icond + 1; else cnt = cnt + 1.0;
      } else {
        mq = mq - 1;
      }
    }
    // cost(i,j) is incremented by 1.0 per filled level; a sum of small integers is exact
    if (v.iconv != 1) v.cost[cell2(I, i, j) * MS + m] += cnt;
  }
  if (v.iconv == 1) {   // the convection diagnostic BIOGEM reads (:2766-2770): depth of the deepest convection of this call
    if (icond > icosd) icosd = icond;
    v.cost[cell2(I, i, j) * MS + m] = 5.0e3 * c_g.zw[K - 1 - icosd];   // dsc * zw(maxk-1-icosd)
  }
  if (v.ieos) {   // "make sure rho calculation is vertically local", tstepo :2396-2408 (wet levels; the dry ones hold eos(0, 0) = 0)
    for (int k = k1c; k <= K; k++) RHOK(k) = eosz(k, c_g.zro[k]);
  }
#undef RHOK
#undef TSK
}

#if CG_TRACER_FAST
// Convective adjustment, fast variant.  The mixing DECISIONS and the T,S,rho values follow the
// reference algorithm operation for operation on a register/local copy of the column (so they are
// bit-identical to k_co_strict); the passive tracers l >= 3, which never feed back into a decision,
// are then averaged once per final mixed region [lo..hi] (thickness weighted) instead of being
// re-mixed at every incremental merge: one read and one write per tracer-level, rounding-level
// differences only.
__global__ void __launch_bounds__(128) k_co_fast2(const Dev v) {
  const int I = v.I, J = v.J, K = v.K, L = v.L;
  const unsigned MS = v.MS;
  const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
  const int col = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || col >= v.nwet) return;
  const int c2 = v.wetcols[col];
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  double *__restrict__ ts = v.ts_new;
  const unsigned sL = MS, sC = (unsigned)L * MS, sK = (unsigned)(I * J) * sC, rK = (unsigned)(I * J) * MS;
  const unsigned oC = (unsigned)cell3(I, J, i, j, 1) * sC + m, rC = (unsigned)cell3(I, J, i, j, 1) * MS + m;
  int kk[kMaxK + 2], head[kMaxK + 2];
  double dzm[kMaxK + 2], rl[kMaxK + 2], tt[kMaxK + 2], ss[kMaxK + 2];
  kk[k1c - 1] = 0;
  rl[0] = 0.0;
  for (int q = k1c; q <= K; q++) {
    kk[q] = q;
    head[q] = q;
    dzm[q] = c_g.dz[q];
    rl[q] = v.rho[rC + (unsigned)(q - 1) * rK];
    tt[q] = ts[oC + (unsigned)(q - 1) * sK];
    ss[q] = ts[oC + (unsigned)(q - 1) * sK + sL];
  }
  int mm = K, lastmix = 0;
  bool any = false;
  while (kk[mm - 1] > 0 || (lastmix != 0 && kk[mm] != K)) {
    if (kk[mm - 1] == 0 || rl[kk[mm]] < rl[kk[mm - 1]]) {
      if (lastmix == 0 || kk[mm] == K) mm = mm - 1; else mm = mm + 1;
      lastmix = 0;
    } else {
      lastmix = 1;
      any = true;
      int n = mm - 1;
      while (kk[n - 1] > 0 && rl[kk[n]] >= rl[kk[n - 1]]) n = n - 1;
      const int h = kk[mm];
      double sumT = tt[h] * dzm[h], sumS = ss[h] * dzm[h], dznew = dzm[h];
      for (int ni = 1; ni <= mm - n; ni++) {
        const int q = kk[mm - ni];
        sumT = sumT + tt[q] * dzm[q];
        sumS = sumS + ss[q] * dzm[q];
        dznew = dznew + dzm[q];
      }
      dzm[h] = dznew;
      tt[h] = sumT / dznew;
      ss[h] = sumS / dznew;
      rl[h] = ec1 * tt[h] + ec2 * ss[h] + ec3 * (tt[h] * tt[h]) + ec4 * (tt[h] * tt[h] * tt[h]);
      int ni = mm - 1;
      while (kk[ni + 1] > 0) {
        kk[ni] = kk[ni - mm + n];
        ni = ni - 1;
      }
    }
  }
  if (!any) return;
  // fill in (goldstein.f90:2749-2764) and record the head of every level's mixed region
  int mq = K - 1;
  double cnt = 0.0;
  for (int n = K - 1; n >= k1c; n--) {
    if (n > kk[mq]) {
      head[n] = kk[mq + 1];
      cnt = cnt + 1.0;
    } else {
      mq = mq - 1;
    }
  }
  v.cost[(unsigned)c2 * MS + m] += cnt;
  // write back T, S, rho of every level that belongs to a mixed region, then average the other tracers
  int hi = K;
  while (hi >= k1c) {
    int lo = hi;
    while (lo - 1 >= k1c && head[lo - 1] == hi) lo--;
    if (lo < hi) {
      const double tv = tt[hi], sv = ss[hi], rv = rl[hi];
      for (int n = lo; n <= hi; n++) {
        ts[oC + (unsigned)(n - 1) * sK] = tv;
        ts[oC + (unsigned)(n - 1) * sK + sL] = sv;
        v.rho[rC + (unsigned)(n - 1) * rK] = rv;
      }
      double dzt = 0.0;
      for (int n = hi; n >= lo; n--) dzt += c_g.dz[n];
      const double rdzt = 1.0 / dzt;
      for (int l = 2; l < L; l += 2) {
        const bool two = l + 1 < L;
        double a0 = 0.0, a1 = 0.0;
        for (int n = hi; n >= lo; n--) {
          const unsigned o = oC + (unsigned)(n - 1) * sK + (unsigned)l * sL;
          a0 += ts[o] * c_g.dz[n];
          if (two) a1 += ts[o + sL] * c_g.dz[n];
        }
        a0 *= rdzt;
        a1 *= rdzt;
        for (int n = lo; n <= hi; n++) {
          const unsigned o = oC + (unsigned)(n - 1) * sK + (unsigned)l * sL;
          ts[o] = a0;
          if (two) ts[o + sL] = a1;
        }
      }
    }
    hi = lo - 1;
  }
}
#endif  // CG_TRACER_FAST

}  // namespace cg
