// k_tracer_col.cuh -- GOLDSTEIN tracer timestep, column formulation with bulk-copy (TMA) staging.
//
// Reference: src/goldstein/goldstein.f90:2436-2642 (tstepo_flux), :2657-2777 (co), :3048-3082 (eos),
//            :428-431 (tstar_ocn/sstar_ocn export).
//
// One thread block = ONE wet column x NT = MS (<= 128) ensemble members, one thread per member; the thread marches up
// the column.  Differences from k_tstepo_flux_strict (reference order, bit-exact) and k_tstepo_flux_coop:
//   * the grid shape and the member stride are template parameters: every address is `base + immediate`;
//   * the stencil rows of a level are brought into shared memory by cp.async.bulk (one elected thread, 18 copies of
//     1-8 KB per level, completion on an mbarrier), two half-level buffers in flight, so no register is tied up by a
//     load in flight and the LSU only sees conflict-free LDS;
//   * closed faces are handled by pointing the neighbour at the centre column (differences vanish exactly);
//   * the update is organised by FACE: when the level-kk values of a tracer arrive they are used at once for (i) the
//     upper half of the flux through face kk-1/2, (ii) the horizontal divergence of level kk and (iii) the lower half
//     of the flux through face kk+1/2, so only two numbers per tracer (P = lower half of the open face,
//     Q = c - H + fb*cZ) are carried from level to level and no stencil value is kept or reloaded;
//   * per-cell work (upstream weights, isoneutral slopes, slope limiter -> 15 linear coefficients) is done once per
//     (member, cell) and amortised over all L tracers;
//   * the thread that produced a column owns its new T, S, rho: the convective-adjustment DECISIONS (co) run in the
//     tail on that column, T, S, rho of mixed levels are rewritten, and the region map (top level of the mixed region
//     of every level) is left for k_co_passive, which averages the passive tracers of mixed regions in one
//     bandwidth-bound pass.
//
// The per-thread body is plain C++ (CG_HD); tests/ compile it for the host (bulk copies emulated element-wise, same
// addressing code) and check it against the oracle without a GPU (tests/test_col_body_host.py).  The product only
// ever runs the __global__ wrappers in k_tracer_col.cu.
#pragma once
#include "cg_device.cuh"

#if defined(__CUDACC__)
#define CG_HD __host__ __device__ __forceinline__
#else
#define CG_HD inline
#endif
#ifdef __CUDA_ARCH__
#define CG_CLZ(x) __clz((int)(x))
#define CG_FFS(x) __ffs((int)(x))
#else
#define CG_CLZ(x) __builtin_clz(x)
#define CG_FFS(x) __builtin_ffs((int)(x))
#endif

namespace cg {

// 1/x for the cell coefficients: MUFU.RCP64H seed (about 20 bits) + two Newton steps, <= 2 ulp.  The IEEE division
// sequence is a chain of ten dependent fp64 operations at ~40 cycles each on B200 and three of them sit on the critical
// path of every level (upstream weight, 1/dzrho, slope limiter); this one is six.  Normal operands only (the callers'
// arguments are bounded away from 0 and infinity).
CG_HD double col_rcp(const double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}

// ---- staging primitives (device: mbarrier + cp.async.bulk; host emulation: element copies, no barriers)
struct ColStage {
  double *sm;              // staging area of the block
  unsigned long long *bar; // four mbarriers: C0, C1 (T,S + velocities, double buffered), A, B (tracer halves)
  int tid;
};

CG_HD void stage_init(const ColStage &s) {
#ifdef __CUDA_ARCH__
  if (s.tid == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(s.bar);
    for (unsigned q = 0; q < 4; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a + 8 * q), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
#else
  (void)s;
#endif
}
// the elected thread announces `bytes` on barrier `which`
CG_HD void stage_expect(const ColStage &s, const int which, const unsigned bytes) {
#ifdef __CUDA_ARCH__
  const unsigned a = (unsigned)__cvta_generic_to_shared(s.bar) + 8u * which;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
#else
  (void)s; (void)which; (void)bytes;
#endif
}
// copy `rows` consecutive rows of NT doubles (global row stride == NT) to staging row `dstrow`
template <int NT>
CG_HD void stage_copy(const ColStage &s, const int which, const int dstrow, const double *src, const int rows) {
#ifdef __CUDA_ARCH__
  const unsigned d = (unsigned)__cvta_generic_to_shared(s.sm + dstrow * NT);
  const unsigned a = (unsigned)__cvta_generic_to_shared(s.bar) + 8u * which;
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
               "r"((unsigned)(rows * NT * 8)), "r"(a)
               : "memory");
#else
  (void)which;
  for (int r = 0; r < rows; r++) s.sm[(dstrow + r) * NT + s.tid] = src[r * NT + s.tid];
#endif
}
CG_HD void stage_wait(const ColStage &s, const int which, const unsigned parity) {
#ifdef __CUDA_ARCH__
  const unsigned a = (unsigned)__cvta_generic_to_shared(s.bar) + 8u * which;
  asm volatile(
      "{\n\t.reg .pred p;\n\tCGW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra CGD_%=;\n\tbra CGW_%=;\n\tCGD_%=:\n\t}" ::"r"(a),
      "r"(parity)
      : "memory");
#else
  (void)s; (void)which; (void)parity;
#endif
}
CG_HD void stage_sync() {
#ifdef __CUDA_ARCH__
  __syncthreads();
#endif
}
CG_HD bool stage_leader(const ColStage &s) {
#ifdef __CUDA_ARCH__
  return s.tid == 0;
#else
  (void)s;
  return true;   // host emulation: every "thread" copies its own elements
#endif
}

// rows of the staging buffers.  Unit C (double buffered, issued 1.5 levels ahead): T,S of the five columns one level
// up + the five velocities; unit A: tracers 2..LH-1 of the five columns; unit B: tracers LH..L-1.
template <int L>
struct ColRows {
  static constexpr int LH = L / 2;
  static constexpr int lB0 = (LH > 2) ? LH : 2;
  static constexpr int nA = lB0 - 2, nB = L - lB0;
  static constexpr int rowsC = 15;                 // rTS: 5 cells x (T,S); rU: uE, vN, ww (centre), uW (west), vS (south)
  static constexpr int rTS = 0, rU = 10;
  static constexpr int rA = 2 * rowsC;             // + cell * nA + (l - 2)
  static constexpr int rowsA = 5 * nA;
  static constexpr int rB = rA + rowsA;            // + cell * nB + (l - lB0)
  static constexpr int rowsB = 5 * nB;
  static constexpr int rows = rB + rowsB;
};

// One (member, column).  All NT threads of a block call this with the same c2.
template <int I, int J, int K, int L, int MS, int NT>
CG_HD void tstep_column(const Dev &v, const GridC &g, const int c2, const unsigned m, const ColStage &st) {
  static_assert(NT == MS, "a block covers all members of one column");
  using R = ColRows<L>;
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC;
  constexpr long uC3 = 3L * MS, uK = (long)I * J * uC3, rK = (long)I * J * MS;
  const int i = c2 % I + 1, j = c2 / I + 1;
#define CGC_K1(ii, jj) ((int)v.k1[(ii) + (I + 2) * (jj)])
  const int k1c = CGC_K1(i, j);
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CGC_K1(ip, j), k1w = CGC_K1(im, j), k1n = CGC_K1(i, j + 1), k1s = CGC_K1(i, j - 1);
#undef CGC_K1

  const double diff1 = v.p.diff1[m], diffv = v.p.diff2[m], rdiff1 = 1.0 / diff1, rdiffv = 1.0 / diffv;
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  const double dt = g.dt, dphi = g.dphi, rdphi = g.rdphi;
  const double rc = g.rc[j], rc2 = g.rc2[j], cvj = g.cv[j], cvjm = g.cv[j - 1], cv2j = g.cv2[j],
               cv2jm = (j > 1) ? g.cv2[j - 1] : 0.0;
  const double dsvN = g.dsv[(j < J - 1) ? j : J - 1], dsvS = (j > 1) ? g.dsv[j - 1] : 0.0;
  const double gyN = (j < J) ? cvj * g.rdsv[j] : 0.0, gyS = (j > 1) ? cvjm * g.rdsv[j - 1] : 0.0, gxx = rc * rdphi;
  const double cX = dt * rdphi, cY = dt * g.rds[j];
  const double dEh = rc2 * diff1, dNh = cv2j * diff1, dSh = cv2jm * diff1;

  // element offsets of the neighbour columns relative to the centre column (periodic in i)
  const long dE = (i < I) ? sC : -(long)(I - 1) * sC, dW = (i > 1) ? -sC : (long)(I - 1) * sC;
  constexpr long dN = (long)I * sC, dS = -(long)I * sC;
  const long dUW = (i > 1) ? -uC3 : (long)(I - 1) * uC3, dUS = (j > 1) ? -(long)I * uC3 : 0;

  // member-0 row of the centre cell at level 1 (block-uniform: the staging copies move whole rows)
  const double *const ts0 = v.ts_cur + (long)c2 * sC;
  const double *const u0 = v.u + (long)c2 * uC3;
  const double *const sm = st.sm + st.tid;

  // issue the staging units of level `lev` (lev <= K).  Unit C: T,S one level up (or the level itself at the top,
  // where every upper coefficient is zero) and the five velocities; units A / B: the passive tracers.
  auto issueC = [&](const int lev) {
    const int b = (lev - k1c) & 1;
    stage_expect(st, b, (unsigned)(R::rowsC * NT * 8));
    const int lu = (lev < K) ? lev + 1 : K;
    const double *c1 = ts0 + (long)(lu - 1) * sK;
    const int r0 = b * R::rowsC;
    stage_copy<NT>(st, b, r0 + R::rTS + 0, c1, 2);
    stage_copy<NT>(st, b, r0 + R::rTS + 2, c1 + ((lu >= k1e) ? dE : 0), 2);
    stage_copy<NT>(st, b, r0 + R::rTS + 4, c1 + ((lu >= k1w) ? dW : 0), 2);
    stage_copy<NT>(st, b, r0 + R::rTS + 6, c1 + ((lu >= k1n) ? dN : 0), 2);
    stage_copy<NT>(st, b, r0 + R::rTS + 8, c1 + ((lu >= k1s) ? dS : 0), 2);
    const double *pu = u0 + (long)(lev - 1) * uK;
    stage_copy<NT>(st, b, r0 + R::rU + 0, pu, 3);
    stage_copy<NT>(st, b, r0 + R::rU + 3, pu + dUW, 1);
    stage_copy<NT>(st, b, r0 + R::rU + 4, pu + dUS + sL, 1);
  };
  auto issueA = [&](const int lev) {
    if (R::nA == 0) return;
    stage_expect(st, 2, (unsigned)(R::rowsA * NT * 8));
    const double *c0 = ts0 + (long)(lev - 1) * sK + 2 * sL;
    stage_copy<NT>(st, 2, R::rA + 0 * R::nA, c0, R::nA);
    stage_copy<NT>(st, 2, R::rA + 1 * R::nA, c0 + ((lev >= k1e) ? dE : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 2 * R::nA, c0 + ((lev >= k1w) ? dW : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 3 * R::nA, c0 + ((lev >= k1n) ? dN : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 4 * R::nA, c0 + ((lev >= k1s) ? dS : 0), R::nA);
  };
  auto issueB = [&](const int lev) {
    stage_expect(st, 3, (unsigned)(R::rowsB * NT * 8));
    const double *c0 = ts0 + (long)(lev - 1) * sK + R::lB0 * sL;
    stage_copy<NT>(st, 3, R::rB + 0 * R::nB, c0, R::nB);
    stage_copy<NT>(st, 3, R::rB + 1 * R::nB, c0 + ((lev >= k1e) ? dE : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 2 * R::nB, c0 + ((lev >= k1w) ? dW : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 3 * R::nB, c0 + ((lev >= k1n) ? dN : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 4 * R::nB, c0 + ((lev >= k1s) ? dS : 0), R::nB);
  };

  stage_init(st);
  if (stage_leader(st)) {
    issueC(k1c);
    if (k1c < K) issueC(k1c + 1);
    issueA(k1c);
    issueB(k1c);
  }

  // T,S of the five columns at the bottom level (direct loads, once per column)
  double tC0, sC0, tE0, sE0, tW0, sW0, tN0, sN0, tS0, sS0;
  {
    const double *qC = ts0 + (long)(k1c - 1) * sK + m;
    const double *qE = qC + ((k1c >= k1e) ? dE : 0), *qW = qC + ((k1c >= k1w) ? dW : 0);
    const double *qN = qC + ((k1c >= k1n) ? dN : 0), *qS = qC + ((k1c >= k1s) ? dS : 0);
    tC0 = qC[0]; sC0 = qC[sL]; tE0 = qE[0]; sE0 = qE[sL]; tW0 = qW[0]; sW0 = qW[sL];
    tN0 = qN[0]; sN0 = qN[sL]; tS0 = qS[0]; sS0 = qS[sL];
  }
  double *wP = v.ts_new + ((long)(k1c - 2) * (I * J) + c2) * sC + m;   // level kk-1 of the new array
  double *rP = v.rho + ((long)(k1c - 2) * (I * J) + c2) * MS + m;
  // upper-half coefficients of the face below the current level, cZ of the level below
  double uc = 0.0, uE = 0.0, uW = 0.0, uN = 0.0, uS = 0.0, cZp = 0.0;
  double P[L], Q[L];
#pragma unroll
  for (int l = 0; l < L; l++) { P[l] = 0.0; Q[l] = 0.0; }
  unsigned par = 0;

  for (int kk = k1c; kk <= K; kk++) {
    const bool top = (kk == K);
    const bool opE = kk >= k1e, opW = kk >= k1w, opN = kk >= k1n, opS = kk >= k1s;
    const int cb = (kk - k1c) & 1;
    stage_wait(st, cb, ((unsigned)(kk - k1c) >> 1) & 1u);
    const double *const smc = sm + cb * R::rowsC * NT;
    const double tC1 = smc[(R::rTS + 0) * NT], sC1 = smc[(R::rTS + 1) * NT], tE1 = smc[(R::rTS + 2) * NT], sE1 = smc[(R::rTS + 3) * NT];
    const double tW1 = smc[(R::rTS + 4) * NT], sW1 = smc[(R::rTS + 5) * NT], tN1 = smc[(R::rTS + 6) * NT], sN1 = smc[(R::rTS + 7) * NT];
    const double tS1 = smc[(R::rTS + 8) * NT], sS1 = smc[(R::rTS + 9) * NT];
    const double vuE = smc[(R::rU + 0) * NT], vvN = smc[(R::rU + 1) * NT], vww = smc[(R::rU + 2) * NT], vuW = smc[(R::rU + 3) * NT],
                 vvS = smc[(R::rU + 4) * NT];

    // ---- horizontal faces of level kk: flux = a * ts(neighbour) + b * ts(centre)   (goldstein.f90:2517-2547).
    // Branch free (closed faces and the top level through 0/1 masks): the whole level is one basic block, so the
    // scheduler can run the tracers' independent FMA chains under the long dependent chain of the slope terms.
    const double mE = opE ? 1.0 : 0.0, mW = opW ? 1.0 : 0.0, mN = opN ? 1.0 : 0.0, mS = opS ? 1.0 : 0.0, mT = top ? 0.0 : 1.0;
    double hE, hW, hN, hS, hC;
    {
      const double pE = vuE * dphi * rdiff1, pW = vuW * dphi * rdiff1, pN = vvN * dsvN * rdiff1, pS = vvS * dsvS * rdiff1;
      const double uE_ = pE * col_rcp(2.0 + fabs(pE)), uW_ = pW * col_rcp(2.0 + fabs(pW));
      const double uN_ = pN * col_rcp(2.0 + fabs(pN)), uS_ = pS * col_rcp(2.0 + fabs(pS));
      const double gE = vuE * rc * 0.5, gW = vuW * rc * 0.5, gN = cvj * vvN * 0.5, gS = cvjm * vvS * 0.5;
      const double cXE = cX * mE, cXW = cX * mW, cYN = cY * mN, cYS = cY * mS;
      hE = (gE * (1.0 - uE_) - dEh) * cXE;
      hW = -(gW * (1.0 + uW_) + dEh) * cXW;
      hN = (gN * (1.0 - uN_) - dNh) * cYN;
      hS = -(gS * (1.0 + uS_) + dSh) * cYS;
      hC = ((gE * (1.0 + uE_) + dEh) * cXE - (gW * (1.0 - uW_) - dEh) * cXW) +
           ((gN * (1.0 + uN_) + dNh) * cYN - (gS * (1.0 - uS_) - dSh) * cYS);
    }
    // ---- face kk+1/2: vertical advection/diffusion + isoneutral terms (goldstein.f90:2549-2621)
    double lc, lE, lW, lN, lS;        // coefficients of the level-kk values
    double nuc, nuE, nuW, nuN, nuS;   // coefficients of the level-kk+1 values
    {
      const double rdza = top ? 0.0 : g.rdza[kk];
      const double pA = vww * g.dza[kk] * rdiffv, uA_ = pA * col_rcp(2.0 + fabs(pA)), gA = vww * 0.5 * mT, dA = rdza * diffv;
      nuc = gA * (1.0 - uA_) - dA;
      lc = gA * (1.0 + uA_) + dA;
      const double tatw = 0.5 * (tC0 + tC1);
      const double tec = -ec1 - ec3 * tatw * 2 - ec4 * tatw * tatw * 3;
      const double dzrho = (ec2 * (sC1 - sC0) - tec * (tC1 - tC0)) * rdza;
      const bool iso = dzrho < -1.0e-12;                     // false at the top level (rdza = 0)
      const double dzs = iso ? dzrho : -1.0, mI = iso ? 1.0 : 0.0;
      // density slopes on the four stencils; a closed face has neighbour == centre, i.e. a zero difference
      const double x0 = ec2 * ((sC0 - sW0) * gxx) - tec * ((tC0 - tW0) * gxx);
      const double x1 = ec2 * ((sE0 - sC0) * gxx) - tec * ((tE0 - tC0) * gxx);
      const double x2 = ec2 * ((sC1 - sW1) * gxx) - tec * ((tC1 - tW1) * gxx);
      const double x3 = ec2 * ((sE1 - sC1) * gxx) - tec * ((tE1 - tC1) * gxx);
      const double y0 = ec2 * ((sC0 - sS0) * gyS) - tec * ((tC0 - tS0) * gyS);
      const double y1 = ec2 * ((sN0 - sC0) * gyN) - tec * ((tN0 - tC0) * gyN);
      const double y2 = ec2 * ((sC1 - sS1) * gyS) - tec * ((tC1 - tS1) * gyS);
      const double y3 = ec2 * ((sN1 - sC1) * gyN) - tec * ((tN1 - tC1) * gyN);
      const double tv1 = (((x0 * x0 + y0 * y0) + (x1 * x1 + y1 * y1)) + (x2 * x2 + y2 * y2)) + (x3 * x3 + y3 * y3);
      const double rdz = col_rcp(dzs), rdz2 = rdz * rdz;
      const double sl = 0.25 * tv1 * rdz2, ssm = g.ssmax[kk];
      const double slim = (sl > ssm) ? ssm * ssm * col_rcp(sl * sl) : 1.0;
      const double cf = 0.25 * slim * diff1 * rdz2 * mI;
      const double g2 = 2.0 * dzs * cf, gX = g2 * gxx, gS = g2 * gyS, gN = g2 * gyN;
      const double s2 = tv1 * cf * rdza;
      const double wx0 = x0 * gX, wx1 = x1 * gX, wx2 = x2 * gX, wx3 = x3 * gX;
      const double wy0 = y0 * gS, wy1 = y1 * gN, wy2 = y2 * gS, wy3 = y3 * gN;
      lc += (wx0 - wx1) + (wy0 - wy1) + s2;
      nuc += (wx2 - wx3) + (wy2 - wy3) - s2;
      lW = -wx0; lE = wx1; lS = -wy0; lN = wy1;
      nuW = -wx2; nuE = wx3; nuS = -wy2; nuN = wy3;
    }
    const double cZ = dt * g.rdz[kk];
    const bool stv = kk > k1c;

    // ---- tracers: one tracer-cell = 15 FMA + 5
    double tnew = 0.0, snew = 0.0;
#define CG_TRACER(l, cc, EE, WW, NN, SS)                                                      \
  {                                                                                           \
    const double c = (cc), E = (EE), W = (WW), N = (NN), S = (SS);                            \
    const double fab = P[l] + (uc * c + uE * E + uW * W + uN * N + uS * S);                   \
    if (stv) {                                                                                \
      const double tn = Q[l] - fab * cZp;                                                     \
      wP[(l) * sL] = tn;                                                                      \
      if ((l) == 0) tnew = tn;                                                                \
      if ((l) == 1) snew = tn;                                                                \
    }                                                                                         \
    const double Hh = hC * c + hE * E + hW * W + hN * N + hS * S;                             \
    Q[l] = (c - Hh) + fab * cZ;                                                               \
    P[l] = lc * c + lE * E + lW * W + lN * N + lS * S;                                        \
  }
    CG_TRACER(0, tC0, tE0, tW0, tN0, tS0)
    CG_TRACER(1, sC0, sE0, sW0, sN0, sS0)
    if (R::nA > 0) stage_wait(st, 2, par);
#pragma unroll
    for (int l = 2; l < R::lB0; l++) {
      const int r = R::rA + (l - 2);
      CG_TRACER(l, sm[(r + 0 * R::nA) * NT], sm[(r + 1 * R::nA) * NT], sm[(r + 2 * R::nA) * NT], sm[(r + 3 * R::nA) * NT],
                sm[(r + 4 * R::nA) * NT])
    }
    stage_sync();                                           // every thread is done with buffers C[cb] and A
    if (stage_leader(st)) {
      if (kk + 2 <= K) issueC(kk + 2);
      if (!top) issueA(kk + 1);
    }
    stage_wait(st, 3, par);
#pragma unroll
    for (int l = R::lB0; l < L; l++) {
      const int r = R::rB + (l - R::lB0);
      CG_TRACER(l, sm[(r + 0 * R::nB) * NT], sm[(r + 1 * R::nB) * NT], sm[(r + 2 * R::nB) * NT], sm[(r + 3 * R::nB) * NT],
                sm[(r + 4 * R::nB) * NT])
    }
#undef CG_TRACER
    stage_sync();                                           // ... and with buffer B
    if (!top && stage_leader(st)) issueB(kk + 1);
    par ^= 1u;
    if (stv) {
      const double r = ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);   // :2638
      rP[0] = r;
    }
    // ---- shift one level up
    tC0 = tC1; sC0 = sC1; tE0 = tE1; sE0 = sE1; tW0 = tW1; sW0 = sW1; tN0 = tN1; sN0 = sN1; tS0 = tS1; sS0 = sS1;
    uc = nuc; uE = nuE; uW = nuW; uN = nuN; uS = nuS; cZp = cZ;
    wP += sK; rP += rK;
  }
  // ---- top level: the flux through the surface is the boundary condition ts(1:2,:,:,maxk+1)   (:2550-2552)
  {
    double tnew = 0.0, snew = 0.0;
#pragma unroll
    for (int l = 0; l < L; l++) {
      double tn = Q[l];
      if (l < 2) tn -= v.tsflux[((long)l * (I * J) + c2) * MS + m] * cZp;
      wP[l * sL] = tn;
      if (l == 0) tnew = tn;
      if (l == 1) snew = tn;
    }
    const double r = ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);
    rP[0] = r;
  }

}

// Convective adjustment (goldstein.f90:2657-2777, iconv == 0) + SST export (:428-431), one thread per (member, wet
// column), run at high occupancy right after the flux kernel.  The mixing DECISIONS and the T, S, rho values follow
// the reference operation for operation on a local copy of the column; the passive tracers (l >= 2), which never feed
// back into a decision, are averaged once per final mixed region (thickness weighted, summed top-down) instead of
// being re-mixed at every incremental merge -- equal up to rounding.  Every load of a tracer pair is independent of
// the others, so the pass is bandwidth bound.
// Part 1: decisions, T/S/rho of the mixed regions, cost, SST.  Returns the region structure: bit k-1 of `in` = level k
// belongs to a mixed region, `topb` / `botb` = it is the region's top / bottom, rdzt[k-1] = 1 / thickness of the region
// (at its bottom level).  in == 0: nothing mixed.
template <int I, int J, int K, int L, int MS>
CG_HD void co_decide(const Dev &v, const GridC &g, const int c2, const unsigned m, unsigned &in, unsigned &topb, unsigned &botb,
                     double *rdzt) {
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC, rK = (long)I * J * MS;
  in = 0; topb = 0; botb = 0;
  const int k1c = (int)v.k1[(c2 % I + 1) + (I + 2) * (c2 / I + 1)];
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  double *__restrict__ ts = v.ts_new + ((long)c2 * sC + m);         // level-1 cell of this column
  double *__restrict__ rho = v.rho + ((long)c2 * MS + m);
  // The reference keeps a compacted index array k(0:maxk) of the levels that still exist as separate boxes and shifts
  // it down after every merge; here the same set is a bit mask (bit l-1 = level l is a separate box), the neighbours of
  // a level come from clz / ffs, and a merge clears bits -- the sequence of comparisons and merges is the reference's.
  int head[K + 2];
  double dzm[K + 2], tt[K + 2], ss[K + 2], rl[K + 2];
#pragma unroll
  for (int q = 1; q <= K; q++) {
    head[q] = q; dzm[q] = g.dz[q];
    tt[q] = 0.0; ss[q] = 0.0; rl[q] = 0.0;
    if (q >= k1c) {
      tt[q] = ts[(long)(q - 1) * sK];
      ss[q] = ts[(long)(q - 1) * sK + sL];
      rl[q] = rho[(long)(q - 1) * rK];
    }
  }
  rl[0] = 0.0;
#ifdef __CUDA_ARCH__
  // the decisions below are a serial, memory-idle stretch: start pulling the column's passive tracers (written by the
  // flux kernel a moment ago, partly evicted since) towards L2 so that the averaging pass finds them there
  if (v.co_prefetch) {
    for (int k = k1c; k <= K; k++) {
#pragma unroll
      for (int l = 2; l < L; l++) asm volatile("prefetch.global.L2 [%0];" ::"l"(ts + (long)(k - 1) * sK + l * sL));
    }
  }
#endif
  unsigned act = (K >= 32 ? 0xffffffffu : ((1u << K) - 1u)) & ~((1u << (k1c - 1)) - 1u);   // levels k1c..K
#define CG_BELOW(x) ({ const unsigned mb_ = act & ((1u << ((x) - 1)) - 1u); mb_ ? 32 - CG_CLZ(mb_) : 0; })
#define CG_ABOVE(x) ({ const unsigned ma_ = act >> (x); (x) + CG_FFS(ma_); })
  int cur = K, lastmix = 0;
  bool any = false;
  for (;;) {
    const int bl = CG_BELOW(cur);
    if (!(bl > 0 || (lastmix != 0 && cur != K))) break;
    if (bl == 0 || rl[cur] < rl[bl]) {
      if (lastmix == 0 || cur == K) cur = bl; else cur = CG_ABOVE(cur);
      lastmix = 0;
    } else {
      lastmix = 1;
      any = true;
      // extend the unstable run downward (goldstein.f90:2722-2731), then mix it into the top box in the order m-1 .. n
      int lo = bl;
      for (;;) {
        const int b2 = CG_BELOW(lo);
        if (!(b2 > 0 && rl[lo] >= rl[b2])) break;
        lo = b2;
      }
      const int h = cur;
      double sumT = tt[h] * dzm[h], sumS = ss[h] * dzm[h], dznew = dzm[h];
      for (int q = bl;; q = CG_BELOW(q)) {
        sumT = sumT + tt[q] * dzm[q];
        sumS = sumS + ss[q] * dzm[q];
        dznew = dznew + dzm[q];
        if (q == lo) break;
      }
      dzm[h] = dznew;
      tt[h] = sumT / dznew;
      ss[h] = sumS / dznew;
      rl[h] = ec1 * tt[h] + ec2 * ss[h] + ec3 * (tt[h] * tt[h]) + ec4 * (tt[h] * tt[h] * tt[h]);
      act &= ~(((1u << bl) - 1u) & ~((1u << (lo - 1)) - 1u));   // levels lo..bl are now part of box h
    }
  }
  // SST / SSS as exported by step_goldstein (:428-431)
  if (v.sst) {
    v.sst[(long)c2 * MS + m] = tt[K];
    v.sst[((long)(I * J) + c2) * MS + m] = ss[K];
  }
  if (!any) return;
  // fill in (:2749-2764): head[n] = the box level n ended up in (the next separate level above it)
  {
    double cnt = 0.0;
    for (int n = K - 1; n >= k1c; n--) {
      if (!((act >> (n - 1)) & 1u)) {
        head[n] = CG_ABOVE(n);
        cnt = cnt + 1.0;
      }
    }
    v.cost[(long)c2 * MS + m] += cnt;
  }
#undef CG_BELOW
#undef CG_ABOVE
  {
    head[0] = 0;
    for (int k = 1; k < k1c; k++) head[k] = 0;
#pragma unroll
    for (int k = K; k >= 1; k--) {
      rdzt[k - 1] = 0.0;
      if (k >= k1c && head[k] != k) {            // inside a region, below its top
        in |= 1u << (k - 1);
        in |= 1u << (head[k] - 1);
      }
    }
    double dzt = 0.0;
#pragma unroll
    for (int k = K; k >= 1; k--) {
      if ((in >> (k - 1)) & 1u) {
        const int hd = head[k];
        const bool istop = (hd == k), isbot = (head[k - 1] != hd);
        dzt = istop ? g.dz[k] : dzt + g.dz[k];
        if (istop) topb |= 1u << (k - 1);
        if (isbot) { botb |= 1u << (k - 1); rdzt[k - 1] = 1.0 / dzt; }
        // T, S, rho of the region (the reference's own values, kept at the region top)
        ts[(long)(k - 1) * sK] = tt[hd];
        ts[(long)(k - 1) * sK + sL] = ss[hd];
        rho[(long)(k - 1) * rK] = rl[hd];
      }
    }
  }
}

// Part 2: one pair (l, l+1) of passive tracers of one (member, column): thickness-weighted mean over every mixed region,
// summed top-down; all loads of the pair are independent.
template <int I, int J, int K, int L, int MS>
CG_HD void co_passive_pair(const Dev &v, const GridC &g, const int c2, const unsigned m, const unsigned in, const unsigned topb,
                           const unsigned botb, const double *rdzt, const int l) {
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC;
  double *__restrict__ ts = v.ts_new + ((long)c2 * sC + m);
  const bool two = (l + 1 < L);
  double a[K], b[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    a[k] = 0.0; b[k] = 0.0;
    if ((in >> k) & 1u) {
      a[k] = ts[(long)k * sK + l * sL];
      if (two) b[k] = ts[(long)k * sK + (l + 1) * sL];
    }
  }
  double accA = 0.0, accB = 0.0;
#pragma unroll
  for (int k = K - 1; k >= 0; k--) {
    if ((in >> k) & 1u) {
      const double dz = g.dz[k + 1];
      if ((topb >> k) & 1u) { accA = 0.0; accB = 0.0; }
      accA += a[k] * dz;
      accB += b[k] * dz;
      if ((botb >> k) & 1u) { a[k] = accA * rdzt[k]; b[k] = accB * rdzt[k]; }
    }
  }
  double curA = 0.0, curB = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    if ((in >> k) & 1u) {
      if ((botb >> k) & 1u) { curA = a[k]; curB = b[k]; }
      ts[(long)k * sK + l * sL] = curA;
      if (two) ts[(long)k * sK + (l + 1) * sL] = curB;
    }
  }
}

// both parts by one thread (host test harness; single-kernel fallback)
template <int I, int J, int K, int L, int MS>
CG_HD void co_column(const Dev &v, const GridC &g, const int c2, const unsigned m) {
  unsigned in, topb, botb;
  double rdzt[K];
  co_decide<I, J, K, L, MS>(v, g, c2, m, in, topb, botb, rdzt);
  if (in == 0) return;
  for (int l = 2; l < L; l += 2) co_passive_pair<I, J, K, L, MS>(v, g, c2, m, in, topb, botb, rdzt, l);
}

}  // namespace cg
