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
#define CG_POPC(x) __popc(x)
#else
#define CG_CLZ(x) __builtin_clz(x)
#define CG_FFS(x) __builtin_ffs((int)(x))
#define CG_POPC(x) __builtin_popcount(x)
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

CG_HD void stage_init(const ColStage &s, const unsigned nbar = 4) {
#ifdef __CUDA_ARCH__
  if (s.tid == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(s.bar);
    for (unsigned q = 0; q < nbar; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a + 8 * q), "r"(1) : "memory");
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
// Consumer release / producer acquire of a staging buffer without a block barrier: every warp arrives once on an "empty"
// mbarrier (initialised with the number of warps) when it has read the buffer for the last time; only the thread that
// issues the refill waits for the phase to complete, the other warps run on to their next data wait.
CG_HD void stage_init_empty(const ColStage &s, const unsigned first, const unsigned n, const unsigned nwarps) {
#ifdef __CUDA_ARCH__
  if (s.tid == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(s.bar);
    for (unsigned q = first; q < first + n; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a + 8 * q), "r"(nwarps) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
#else
  (void)s; (void)first; (void)n; (void)nwarps;
#endif
}
CG_HD void stage_release(const ColStage &s, const int which) {
#ifdef __CUDA_ARCH__
  __syncwarp();
  if ((s.tid & 31) == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(s.bar) + 8u * which;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
  }
#else
  (void)s; (void)which;
#endif
}
// barrier over one half of a split block (named barrier 1 + half, nthreads threads)
CG_HD void stage_sync_half(const int half, const int nthreads) {
#ifdef __CUDA_ARCH__
  if (half == 0) asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory");
#else
  (void)half; (void)nthreads;
#endif
}
// orders this thread's earlier generic-proxy accesses to shared memory before the bulk copies it issues next
CG_HD void stage_fence() {
#ifdef __CUDA_ARCH__
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
// The thread that issues staging unit `unit` (0 = C, 1 = A, 2 = B): lane 0 of warp `unit` where the block has that many warps.
// One leader issuing all 18 copies of a level executes ~270 instructions more per level than the other warps, which then wait
// for it at the next block barrier (ncu: 18 % of the stall cycles "barrier"); spread over three warps the extra work is even.
template <int NT>
CG_HD bool stage_issuer(const ColStage &s, const int unit) {
#ifdef __CUDA_ARCH__
  constexpr int NW = NT / 32;
  return s.tid == 32 * (unit < NW ? unit : NW - 1);
#else
  (void)s; (void)unit;
  return true;
#endif
}

// Tensor maps of the two fields the kernels stage, by box height (rows per copy); built on the host (k_tracer_col.cu).  Box width =
// the member tile of the kernel that uses them (32: warp-tile form; 128: block form on a member stride > 128).
struct ColMaps {
  alignas(64) unsigned long long ts2[16], tsA[16], tsB[16], u3[16], u1[16];   // CUtensorMap is 128 opaque bytes, 64-byte aligned
};
// box (NT members starting at member c0) x (rows starting at row c1) -> staging rows dstrow ...; host emulation: element-wise
template <int MS, int NT = 32>
CG_HD void stage_box(const ColStage &s, const int which, const int dstrow, const void *tmap, const double *field, const int c0,
                     const int c1, const int rows) {
#ifdef __CUDA_ARCH__
  (void)field; (void)rows;
  const unsigned d = (unsigned)__cvta_generic_to_shared(s.sm + dstrow * NT);
  const unsigned a = (unsigned)__cvta_generic_to_shared(s.bar) + 8u * which;
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(d), "l"(tmap),
               "r"(a), "r"(c0), "r"(c1)
               : "memory");
#else
  (void)which; (void)tmap;
  for (int r = 0; r < rows; r++) s.sm[(dstrow + r) * NT + s.tid] = field[(long)(c1 + r) * MS + c0 + s.tid];
#endif
}

// rows of the staging buffers.  Unit C (double buffered, issued 1.5 levels ahead): T,S of the five columns one level
// up + the five velocities; unit A: tracers 2..LH-1 of the five columns; unit B: tracers LH..L-1.
template <int L, bool TSW = true>
struct ColRows {
  static constexpr int lA0 = TSW ? 2 : 0;          // first tracer (window-local) that comes through the staging units
  static constexpr int LH = L / 2;
  static constexpr int lB0 = (LH > lA0) ? LH : lA0;
  static constexpr int nA = lB0 - lA0, nB = L - lB0;
  static constexpr int rowsC = 15;                 // rTS: 5 cells x (T,S); rU: uE, vN, ww (centre), uW (west), vS (south)
  static constexpr int rTS = 0, rU = 10;
  static constexpr int rA = 2 * rowsC;             // + cell * nA + (l - lA0)
  static constexpr int rowsA = 5 * nA;
  static constexpr int rB = rA + rowsA;            // + cell * nB + (l - lB0)
  static constexpr int rowsB = 5 * nB;
  static constexpr int rows = rB + rowsB;
};

// Per-(member, column) constants of the flux formulation.
struct ColK {
  double diff1, diffv, rdiff1, rdiffv, ec1, ec2, ec3, ec4, dt, dphi;
  double rc, cvj, cvjm, dsvN, dsvS, gyN, gyS, gxx, cX, cY, dEh, dNh, dSh;
};
template <int I, int J>
CG_HD ColK col_consts(const Dev &v, const GridC &g, const unsigned m, const int j) {
  ColK q;
  q.diff1 = v.p.diff1[m]; q.diffv = v.p.diff2[m]; q.rdiff1 = 1.0 / q.diff1; q.rdiffv = 1.0 / q.diffv;
  q.ec1 = v.p.ec1[m]; q.ec2 = v.p.ec2[m]; q.ec3 = v.p.ec3[m]; q.ec4 = v.p.ec4[m];
  q.dt = g.dt; q.dphi = g.dphi;
  const double rdphi = g.rdphi, rc2 = g.rc2[j], cv2j = g.cv2[j], cv2jm = (j > 1) ? g.cv2[j - 1] : 0.0;
  q.rc = g.rc[j]; q.cvj = g.cv[j]; q.cvjm = g.cv[j - 1];
  q.dsvN = g.dsv[(j < J - 1) ? j : J - 1]; q.dsvS = (j > 1) ? g.dsv[j - 1] : 0.0;
  q.gyN = (j < J) ? q.cvj * g.rdsv[j] : 0.0; q.gyS = (j > 1) ? q.cvjm * g.rdsv[j - 1] : 0.0; q.gxx = q.rc * rdphi;
  q.cX = q.dt * rdphi; q.cY = q.dt * g.rds[j];
  q.dEh = rc2 * q.diff1; q.dNh = cv2j * q.diff1; q.dSh = cv2jm * q.diff1;
  return q;
}
// T, S of the centre column and its four neighbours at one level (a closed face carries the centre values)
struct TS5 { double tC, sC, tE, sE, tW, sW, tN, sN, tS, sS; };
// The 15 linear coefficients of one cell: horizontal faces of level kk (h*), lower half (l*) and upper half (nu*) of the
// vertical face kk+1/2, and dt/dz of the level.
struct ColCoef { double hE, hW, hN, hS, hC, lc, lE, lW, lN, lS, nuc, nuE, nuW, nuN, nuS, cZ; };

// goldstein.f90:2517-2547 for one (member, cell): the horizontal faces of a level, flux = a * ts(neighbour) + b * ts(centre).
// Branch free (closed faces through 0/1 masks): the whole level is one basic block, so the scheduler can run the tracers'
// independent FMA chains under the long dependent chain of the slope terms.
CG_HD void col_coefs_h(const ColK &q, const bool opE, const bool opW, const bool opN, const bool opS, const double vuE,
                       const double vvN, const double vuW, const double vvS, double &hE, double &hW, double &hN, double &hS,
                       double &hC) {
  const double mE = opE ? 1.0 : 0.0, mW = opW ? 1.0 : 0.0, mN = opN ? 1.0 : 0.0, mS = opS ? 1.0 : 0.0;
  const double pE = vuE * q.dphi * q.rdiff1, pW = vuW * q.dphi * q.rdiff1, pN = vvN * q.dsvN * q.rdiff1, pS = vvS * q.dsvS * q.rdiff1;
  const double uE_ = pE * col_rcp(2.0 + fabs(pE)), uW_ = pW * col_rcp(2.0 + fabs(pW));
  const double uN_ = pN * col_rcp(2.0 + fabs(pN)), uS_ = pS * col_rcp(2.0 + fabs(pS));
  const double gE = vuE * q.rc * 0.5, gW = vuW * q.rc * 0.5, gN = q.cvj * vvN * 0.5, gS = q.cvjm * vvS * 0.5;
  const double cXE = q.cX * mE, cXW = q.cX * mW, cYN = q.cY * mN, cYS = q.cY * mS;
  hE = (gE * (1.0 - uE_) - q.dEh) * cXE;
  hW = -(gW * (1.0 + uW_) + q.dEh) * cXW;
  hN = (gN * (1.0 - uN_) - q.dNh) * cYN;
  hS = -(gS * (1.0 + uS_) + q.dSh) * cYS;
  hC = ((gE * (1.0 + uE_) + q.dEh) * cXE - (gW * (1.0 - uW_) - q.dEh) * cXW) +
       ((gN * (1.0 + uN_) + q.dNh) * cYN - (gS * (1.0 - uS_) - q.dSh) * cYS);
}
// goldstein.f90:2549-2621: face kk+1/2, vertical advection / diffusion + isoneutral terms.  `a` = T,S at level kk, `b` = one
// level up (a closed face carries the centre values, i.e. a zero difference); l* multiply the level-kk values, nu* the
// level-kk+1 values.
template <int K>
CG_HD void col_coefs_v(const ColK &q, const GridC &g, const int kk, const TS5 &a, const TS5 &b, const double vww, double &lc_,
                       double &lE, double &lW, double &lN, double &lS, double &nuc_, double &nuE, double &nuW, double &nuN,
                       double &nuS) {
  const bool top = (kk == K);
  const double mT = top ? 0.0 : 1.0;
  const double tC0 = a.tC, sC0 = a.sC, tE0 = a.tE, sE0 = a.sE, tW0 = a.tW, sW0 = a.sW, tN0 = a.tN, sN0 = a.sN, tS0 = a.tS, sS0 = a.sS;
  const double tC1 = b.tC, sC1 = b.sC, tE1 = b.tE, sE1 = b.sE, tW1 = b.tW, sW1 = b.sW, tN1 = b.tN, sN1 = b.sN, tS1 = b.tS, sS1 = b.sS;
  const double ec1 = q.ec1, ec2 = q.ec2, ec3 = q.ec3, ec4 = q.ec4, gxx = q.gxx, gyS = q.gyS, gyN = q.gyN;
  const double rdza = top ? 0.0 : g.rdza[kk];
  const double pA = vww * g.dza[kk] * q.rdiffv, uA_ = pA * col_rcp(2.0 + fabs(pA)), gA = vww * 0.5 * mT, dA = rdza * q.diffv;
  double nuc = gA * (1.0 - uA_) - dA;
  double lc = gA * (1.0 + uA_) + dA;
  const double tatw = 0.5 * (tC0 + tC1);
  const double tec = -ec1 - ec3 * tatw * 2 - ec4 * tatw * tatw * 3;
  const double dzrho = (ec2 * (sC1 - sC0) - tec * (tC1 - tC0)) * rdza;
  const bool iso = dzrho < -1.0e-12;                     // false at the top level (rdza = 0)
  const double dzs = iso ? dzrho : -1.0, mI = iso ? 1.0 : 0.0;
  // density slopes on the four stencils
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
  const double cf = 0.25 * slim * q.diff1 * rdz2 * mI;
  const double g2 = 2.0 * dzs * cf, gX = g2 * gxx, gS = g2 * gyS, gN = g2 * gyN;
  const double s2 = tv1 * cf * rdza;
  const double wx0 = x0 * gX, wx1 = x1 * gX, wx2 = x2 * gX, wx3 = x3 * gX;
  const double wy0 = y0 * gS, wy1 = y1 * gN, wy2 = y2 * gS, wy3 = y3 * gN;
  lc += (wx0 - wx1) + (wy0 - wy1) + s2;
  nuc += (wx2 - wx3) + (wy2 - wy3) - s2;
  lc_ = lc; nuc_ = nuc;
  lW = -wx0; lE = wx1; lS = -wy0; lN = wy1;
  nuW = -wx2; nuE = wx3; nuS = -wy2; nuN = wy3;
}
// all 15 coefficients + dt/dz of one (member, cell)
template <int K>
CG_HD void col_coefs(const ColK &q, const GridC &g, const int kk, const bool opE, const bool opW, const bool opN, const bool opS,
                     const TS5 &a, const TS5 &b, const double vuE, const double vvN, const double vww, const double vuW,
                     const double vvS, ColCoef &o) {
  col_coefs_h(q, opE, opW, opN, opS, vuE, vvN, vuW, vvS, o.hE, o.hW, o.hN, o.hS, o.hC);
  col_coefs_v<K>(q, g, kk, a, b, vww, o.lc, o.lE, o.lW, o.lN, o.lS, o.nuc, o.nuE, o.nuW, o.nuN, o.nuS);
  o.cZ = q.dt * g.rdz[kk];
}

// One (member, column).  All NT threads of a block call this with the same c2.
// PV = false: all L tracers, new T, S, rho written unmixed (the convective adjustment follows in k_co_col).
// PV = true : the passive tracers (l >= 2) only, convective mixing applied ON WRITE.  T, S, rho and the region map of the
//   column (v.comask: bit k-1 = level k lies in a mixed region, bit 16+k-1 = it is the region's top) come from
//   ts_pre_column.  While the march is inside a region the running thickness-weighted sum rides in Q (Q = sum + dz * the
//   level's partial update), so no extra accumulator is needed; at the region's top the mean is stored to all of its
//   levels.  A level outside any region has weight 1 and carry 0: its value is exactly the unmixed one.
// TM = true: the block is an NT-member TILE of a member stride MS > NT (one handle holding more than 128 members): the rows of
//   a staging unit are then NT * 8 bytes out of every MS * 8, fetched as 2-D tensor-map boxes (NT members x the unit's rows,
//   cp.async.bulk.tensor.2d / UTMALDG) instead of contiguous bulk copies; m = tile * NT + thread.  Everything else is unchanged.
// LT, L0: the block advances the WINDOW of L tracers starting at tracer L0 of a state that carries LT tracers per cell (grids
//   whose tracer count does not fit one thread's registers: 40 tracers = windows of 16 + 12 + 12).  Every window computes the
//   cell coefficients from T, S (unit C); only the window with L0 = 0 advances T and S themselves, writes rho, the stability flag
//   and SST.  LT = L, L0 = 0: the whole tracer set in one block, as for the 16-tracer BIOGEM configuration.
template <int I, int J, int K, int L, int MS, int NT, bool PV, bool ASYNC_REL = false, bool TM = false, int LT = L, int L0 = 0>
CG_HD void tstep_column(const Dev &v, const GridC &g, const int c2, const unsigned m, const ColStage &st, const ColMaps *tm = nullptr) {
  static_assert(TM || NT == MS, "a block covers all members of one column, or a tile of them through tensor maps");
  static_assert(L0 == 0 || !PV, "tracer windows: not with mix-on-write");
  constexpr bool TSW = (L0 == 0);            // this window holds T and S
  static_assert(!PV || K <= 16, "region map is 16 + 16 bits");
  using R = ColRows<L, TSW>;
  constexpr long sL = MS, sC = (long)LT * MS, sK = (long)I * J * sC;
  constexpr long uC3 = 3L * MS, uK = (long)I * J * uC3, rK = (long)I * J * MS;
  const int i = c2 % I + 1, j = c2 / I + 1;
#define CGC_K1(ii, jj) ((int)v.k1[(ii) + (I + 2) * (jj)])
  const int k1c = CGC_K1(i, j);
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CGC_K1(ip, j), k1w = CGC_K1(im, j), k1n = CGC_K1(i, j + 1), k1s = CGC_K1(i, j - 1);
#undef CGC_K1
  const ColK q = col_consts<I, J>(v, g, m, j);
  const double ec1 = q.ec1, ec2 = q.ec2, ec3 = q.ec3, ec4 = q.ec4;

  // element offsets of the neighbour columns relative to the centre column (periodic in i)
  const long dE = (i < I) ? sC : -(long)(I - 1) * sC, dW = (i > 1) ? -sC : (long)(I - 1) * sC;
  constexpr long dN = (long)I * sC, dS = -(long)I * sC;
  const long dUW = (i > 1) ? -uC3 : (long)(I - 1) * uC3, dUS = (j > 1) ? -(long)I * uC3 : 0;

  // member-0 row of the centre cell at level 1 (block-uniform: the staging copies move whole rows)
  const double *const ts0 = v.ts_cur + (long)c2 * sC;
  const double *const u0 = v.u + (long)c2 * uC3;
  const double *const sm = st.sm + st.tid;

  // TM: first member of the tile, neighbour columns as CELL offsets, the maps
  constexpr int IJ = I * J;
  const int m0 = (int)(m - (unsigned)st.tid);
  const int cE = (i < I) ? 1 : -(I - 1), cW = (i > 1) ? -1 : (I - 1), cUS = (j > 1) ? -I : 0;
  const void *const mTS2 = (TM && tm) ? (const void *)tm->ts2 : nullptr, *const mTSA = (TM && tm) ? (const void *)tm->tsA : nullptr,
                   *const mTSB = (TM && tm) ? (const void *)tm->tsB : nullptr, *const mU3 = (TM && tm) ? (const void *)tm->u3 : nullptr,
                   *const mU1 = (TM && tm) ? (const void *)tm->u1 : nullptr;
  auto colcell = [&](const int lev, const int cell) {             // cell offset of stencil column `cell` at level lev
    return (cell == 0) ? 0 : (cell == 1) ? ((lev >= k1e) ? cE : 0) : (cell == 2) ? ((lev >= k1w) ? cW : 0)
           : (cell == 3) ? ((lev >= k1n) ? I : 0) : ((lev >= k1s) ? -I : 0);
  };
  // issue the staging units of level `lev` (lev <= K).  Unit C: T,S one level up (or the level itself at the top,
  // where every upper coefficient is zero) and the five velocities; units A / B: the passive tracers.
  auto issueC = [&](const int lev) {
    const int b = (lev - k1c) & 1;
    stage_expect(st, b, (unsigned)(R::rowsC * NT * 8));
    const int lu = (lev < K) ? lev + 1 : K;
    const int r0 = b * R::rowsC;
    if (TM) {
      const int cellu = (lu - 1) * IJ + c2, cellv = (lev - 1) * IJ + c2;
#pragma unroll
      for (int cell = 0; cell < 5; cell++)
        stage_box<MS, NT>(st, b, r0 + R::rTS + 2 * cell, mTS2, v.ts_cur, m0, (cellu + colcell(lu, cell)) * LT, 2);
      stage_box<MS, NT>(st, b, r0 + R::rU + 0, mU3, v.u, m0, cellv * 3, 3);
      stage_box<MS, NT>(st, b, r0 + R::rU + 3, mU1, v.u, m0, (cellv + cW) * 3, 1);
      stage_box<MS, NT>(st, b, r0 + R::rU + 4, mU1, v.u, m0, (cellv + cUS) * 3 + 1, 1);
      return;
    }
    const double *c1 = ts0 + (long)(lu - 1) * sK;
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
    if (TM) {
      const int cell0 = (lev - 1) * IJ + c2;
#pragma unroll
      for (int cell = 0; cell < 5; cell++)
        stage_box<MS, NT>(st, 2, R::rA + cell * R::nA, mTSA, v.ts_cur, m0, (cell0 + colcell(lev, cell)) * LT + L0 + R::lA0, R::nA);
      return;
    }
    const double *c0 = ts0 + (long)(lev - 1) * sK + (L0 + R::lA0) * sL;
    stage_copy<NT>(st, 2, R::rA + 0 * R::nA, c0, R::nA);
    stage_copy<NT>(st, 2, R::rA + 1 * R::nA, c0 + ((lev >= k1e) ? dE : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 2 * R::nA, c0 + ((lev >= k1w) ? dW : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 3 * R::nA, c0 + ((lev >= k1n) ? dN : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 4 * R::nA, c0 + ((lev >= k1s) ? dS : 0), R::nA);
  };
  auto issueB = [&](const int lev) {
    stage_expect(st, 3, (unsigned)(R::rowsB * NT * 8));
    if (TM) {
      const int cell0 = (lev - 1) * IJ + c2;
#pragma unroll
      for (int cell = 0; cell < 5; cell++)
        stage_box<MS, NT>(st, 3, R::rB + cell * R::nB, mTSB, v.ts_cur, m0, (cell0 + colcell(lev, cell)) * LT + L0 + R::lB0, R::nB);
      return;
    }
    const double *c0 = ts0 + (long)(lev - 1) * sK + (L0 + R::lB0) * sL;
    stage_copy<NT>(st, 3, R::rB + 0 * R::nB, c0, R::nB);
    stage_copy<NT>(st, 3, R::rB + 1 * R::nB, c0 + ((lev >= k1e) ? dE : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 2 * R::nB, c0 + ((lev >= k1w) ? dW : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 3 * R::nB, c0 + ((lev >= k1n) ? dN : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 4 * R::nB, c0 + ((lev >= k1s) ? dS : 0), R::nB);
  };

  stage_init(st);
  if (ASYNC_REL) stage_init_empty(st, 4, 2, NT / 32);
  if (ASYNC_REL) {
    if (stage_issuer<NT>(st, 1)) {
      issueC(k1c);
      if (k1c < K) issueC(k1c + 1);
      issueA(k1c);
    }
    if (stage_issuer<NT>(st, 2)) issueB(k1c);
  } else {
    if (stage_issuer<NT>(st, 0)) {
      issueC(k1c);
      if (k1c < K) issueC(k1c + 1);
    }
    if (stage_issuer<NT>(st, 1)) issueA(k1c);
    if (stage_issuer<NT>(st, 2)) issueB(k1c);
  }

  // T,S of the five columns at the bottom level (direct loads, once per column)
  TS5 a;
  {
    const double *qC = ts0 + (long)(k1c - 1) * sK + m;
    const double *qE = qC + ((k1c >= k1e) ? dE : 0), *qW = qC + ((k1c >= k1w) ? dW : 0);
    const double *qN = qC + ((k1c >= k1n) ? dN : 0), *qS = qC + ((k1c >= k1s) ? dS : 0);
    a.tC = qC[0]; a.sC = qC[sL]; a.tE = qE[0]; a.sE = qE[sL]; a.tW = qW[0]; a.sW = qW[sL];
    a.tN = qN[0]; a.sN = qN[sL]; a.tS = qS[0]; a.sS = qS[sL];
  }
  double *wP = v.ts_new + ((long)(k1c - 2) * (I * J) + c2) * sC + L0 * sL + m;   // level kk-1 of the new array, first tracer of the window
  double *rP = v.rho + ((long)(k1c - 2) * (I * J) + c2) * MS + m;
  // upper-half coefficients of the face below the current level, cZ of the level below
  double uc = 0.0, uE = 0.0, uW = 0.0, uN = 0.0, uS = 0.0, cZp = 0.0;
  double P[L], Q[L];
#pragma unroll
  for (int l = 0; l < L; l++) { P[l] = 0.0; Q[l] = 0.0; }
  unsigned par = 0;
  // !PV: is any level of the new column statically unstable against the one below (rho(k) >= rho(k-1), the test of
  // goldstein.f90:2700)?  If none is, the convective adjustment is the identity on this (member, column): k_co_col skips it.
  bool unstable = false;
  double rbelow = 0.0;
  // PV: region map of this (member, column); thickness and depth (levels) of the region the march is inside
  const unsigned cmask = PV ? v.comask[(long)c2 * MS + m] : 0u;
  double dzt = 0.0;
  int nreg = 0;

  for (int kk = k1c; kk <= K; kk++) {
    const bool top = (kk == K);
    const bool opE = kk >= k1e, opW = kk >= k1w, opN = kk >= k1n, opS = kk >= k1s;
    const int cb = (kk - k1c) & 1;
    stage_wait(st, cb, ((unsigned)(kk - k1c) >> 1) & 1u);
    const double *const smc = sm + cb * R::rowsC * NT;
    TS5 b;
    b.tC = smc[(R::rTS + 0) * NT]; b.sC = smc[(R::rTS + 1) * NT]; b.tE = smc[(R::rTS + 2) * NT]; b.sE = smc[(R::rTS + 3) * NT];
    b.tW = smc[(R::rTS + 4) * NT]; b.sW = smc[(R::rTS + 5) * NT]; b.tN = smc[(R::rTS + 6) * NT]; b.sN = smc[(R::rTS + 7) * NT];
    b.tS = smc[(R::rTS + 8) * NT]; b.sS = smc[(R::rTS + 9) * NT];
    const double vuE = smc[(R::rU + 0) * NT], vvN = smc[(R::rU + 1) * NT], vww = smc[(R::rU + 2) * NT], vuW = smc[(R::rU + 3) * NT],
                 vvS = smc[(R::rU + 4) * NT];
    ColCoef cf;
    col_coefs<K>(q, g, kk, opE, opW, opN, opS, a, b, vuE, vvN, vww, vuW, vvS, cf);
    const double hE = cf.hE, hW = cf.hW, hN = cf.hN, hS = cf.hS, hC = cf.hC;
    const double lc = cf.lc, lE = cf.lE, lW = cf.lW, lN = cf.lN, lS = cf.lS, cZ = cf.cZ;
    const bool stv = kk > k1c;

    // PV: the level being finalised (f = kk-1) closes a region (store the mean to nst levels), continues one (nothing is
    // stored, the sum is carried) or lies outside (nst = 1, scale = 1); the level being formed enters with weight wcur
    double cZw = cZp, scale = 1.0, wcur = 1.0;
    int nst = 1;
    if (PV) {
      if (stv && ((cmask >> (kk - 2)) & 1u)) {
        const double dzf = g.dz[kk - 1];
        dzt += dzf; nreg++;
        cZw = cZp * dzf;
        if ((cmask >> (kk + 14)) & 1u) { scale = 1.0 / dzt; nst = nreg; nreg = 0; dzt = 0.0; }
        else nst = 0;
      }
      if ((cmask >> (kk - 1)) & 1u) wcur = g.dz[kk];
    }

    // ---- tracers: one tracer-cell = 15 FMA + 5
    double tnew = 0.0, snew = 0.0;
#define CG_TRACER(l, cc, EE, WW, NN, SS)                                                      \
  {                                                                                           \
    const double c = (cc), E = (EE), W = (WW), N = (NN), S = (SS);                            \
    const double fab = P[l] + (uc * c + uE * E + uW * W + uN * N + uS * S);                   \
    double cr = 0.0;                                                                          \
    if (stv) {                                                                                \
      if (!PV) {                                                                              \
        const double tn = Q[l] - fab * cZp;                                                   \
        wP[(l) * sL] = tn;                                                                    \
        if ((l) == 0) tnew = tn;                                                              \
        if ((l) == 1) snew = tn;                                                              \
      } else {                                                                                \
        const double rs = Q[l] - fab * cZw;                                                   \
        if (nst > 0) {                                                                        \
          const double val = rs * scale;                                                      \
          double *w = wP + (l) * sL;                                                          \
          w[0] = val;                                                                         \
          for (int q_ = 1; q_ < nst; q_++) w[-(long)q_ * sK] = val;                           \
        } else cr = rs;                                                                       \
      }                                                                                       \
    }                                                                                         \
    const double Hh = hC * c + hE * E + hW * W + hN * N + hS * S;                             \
    if (!PV) Q[l] = (c - Hh) + fab * cZ;                                                      \
    else Q[l] = wcur * ((c - Hh) + fab * cZ) + cr;                                            \
    P[l] = lc * c + lE * E + lW * W + lN * N + lS * S;                                        \
  }
    if (!PV && TSW) {
      CG_TRACER(0, a.tC, a.tE, a.tW, a.tN, a.tS)
      CG_TRACER(1, a.sC, a.sE, a.sW, a.sN, a.sS)
    }
    if (R::nA > 0) stage_wait(st, 2, par);
#pragma unroll
    for (int l = R::lA0; l < R::lB0; l++) {
      const int r = R::rA + (l - R::lA0);
      CG_TRACER(l, sm[(r + 0 * R::nA) * NT], sm[(r + 1 * R::nA) * NT], sm[(r + 2 * R::nA) * NT], sm[(r + 3 * R::nA) * NT],
                sm[(r + 4 * R::nA) * NT])
    }
    // every warp announces that it is done with buffers C[cb] and A; the leader alone waits for all of them
    if (ASYNC_REL) {
      stage_release(st, 4);
      if (stage_issuer<NT>(st, 1)) {     // only the issuing warp waits for the others; they run on to unit B
        stage_wait(st, 4, par);
        if (kk + 2 <= K) issueC(kk + 2);
        if (!top) issueA(kk + 1);
      }
    } else {
      stage_sync();
      if (kk + 2 <= K && stage_issuer<NT>(st, 0)) issueC(kk + 2);
      if (!top && stage_issuer<NT>(st, 1)) issueA(kk + 1);
    }
    stage_wait(st, 3, par);
#pragma unroll
    for (int l = R::lB0; l < L; l++) {
      const int r = R::rB + (l - R::lB0);
      CG_TRACER(l, sm[(r + 0 * R::nB) * NT], sm[(r + 1 * R::nB) * NT], sm[(r + 2 * R::nB) * NT], sm[(r + 3 * R::nB) * NT],
                sm[(r + 4 * R::nB) * NT])
    }
#undef CG_TRACER
    if (ASYNC_REL) {                                        // ... and with buffer B
      stage_release(st, 5);
      if (stage_issuer<NT>(st, 2)) {
        stage_wait(st, 5, par);
        if (!top) issueB(kk + 1);
      }
    } else {
      stage_sync();
      if (!top && stage_issuer<NT>(st, 2)) issueB(kk + 1);
    }
    par ^= 1u;
    if (!PV && TSW && stv) {
      const double r = ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);   // :2638
      rP[0] = r;
      if (kk - 1 > k1c) unstable = unstable || !(r < rbelow);
      rbelow = r;
    }
    // ---- shift one level up
    a = b;
    uc = cf.nuc; uE = cf.nuE; uW = cf.nuW; uN = cf.nuN; uS = cf.nuS; cZp = cZ;
    wP += sK; rP += rK;
  }
  // ---- top level: the flux through the surface is the boundary condition ts(1:2,:,:,maxk+1)   (:2550-2552)
  if (!PV) {
    double tnew = 0.0, snew = 0.0;
#pragma unroll
    for (int l = 0; l < L; l++) {
      double tn = Q[l];
      if (TSW && l < 2) tn -= v.tsflux[((long)l * (I * J) + c2) * MS + m] * cZp;
      wP[l * sL] = tn;
      if (l == 0) tnew = tn;
      if (l == 1) snew = tn;
    }
    if (TSW) {
      const double r = ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);
      rP[0] = r;
      if (K > k1c) unstable = unstable || !(r < rbelow);
      if (v.comask) v.comask[(long)c2 * MS + m] = unstable ? 1u : 0u;
      // SST / SSS as step_goldstein exports them (:428-431); k_co_col rewrites them where it mixes
      if (v.sst) {
        v.sst[(long)c2 * MS + m] = tnew;
        v.sst[((long)(I * J) + c2) * MS + m] = snew;
      }
    }
  } else {
    // passive tracers have no surface flux; level K is either outside any region or the top of one
    double scale = 1.0;
    int nst = 1;
    if ((cmask >> (K - 1)) & 1u) { dzt += g.dz[K]; nreg++; scale = 1.0 / dzt; nst = nreg; }
#pragma unroll
    for (int l = 2; l < L; l++) {
      const double val = Q[l] * scale;
      double *w = wP + l * sL;
      w[0] = val;
      for (int q_ = 1; q_ < nst; q_++) w[-(long)q_ * sK] = val;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-tile form of the column kernel: one block = ONE WARP = 32 members (a 256-byte slice of every row) of one wet column.
// The arithmetic is tstep_column's.  What changes is the staging: the block's four warps no longer meet at two block
// barriers per level (ncu: 19 % of the stall samples sit behind them and 9 % in the waits for rows that the slowest warp
// requested late); every warp owns its staging rows and mbarriers, re-requests a unit as soon as IT has consumed it, and the
// up to eight one-warp blocks of an SM drift freely against each other.  A warp's slice of a staging unit is a 2-D box of
// the field seen as [row = (cell, tracer)][member]: 32 members x the unit's consecutive tracer rows of one stencil column,
// fetched by one tensor-map copy (cp.async.bulk.tensor.2d, SASS UTMALDG) -- 18 copies per level, as the 128-member block
// issues, but per warp.  (A first version copied the 256-byte row slices with "one warp-wide cp.async.bulk, lane r = row r":
// the instruction takes its operands from the uniform datapath, so the compiler turns per-lane addresses into a serial loop
// over the lanes, 95 copies per warp and level, twice the kernel's instructions: profiles/README_r2.md.)
CG_HD void stage_syncw() {
#ifdef __CUDA_ARCH__
  __syncwarp();
#endif
}

template <int I, int J, int K, int L, int MS>
CG_HD void tstep_column_w(const Dev &v, const GridC &g, const int c2, const unsigned m, const ColStage &st, const ColMaps *tm) {
  constexpr int NT = 32;
  using R = ColRows<L>;
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC;
  constexpr long rK = (long)I * J * MS;
  constexpr int IJ = I * J;
  const int i = c2 % I + 1, j = c2 / I + 1;
#define CGC_K1(ii, jj) ((int)v.k1[(ii) + (I + 2) * (jj)])
  const int k1c = CGC_K1(i, j);
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CGC_K1(ip, j), k1w = CGC_K1(im, j), k1n = CGC_K1(i, j + 1), k1s = CGC_K1(i, j - 1);
#undef CGC_K1
  const ColK q = col_consts<I, J>(v, g, m, j);
  const double ec1 = q.ec1, ec2 = q.ec2, ec3 = q.ec3, ec4 = q.ec4;
  // neighbour columns as CELL offsets (periodic in i); a closed face points at the centre column (differences vanish exactly)
  const int cE = (i < I) ? 1 : -(I - 1), cW = (i > 1) ? -1 : (I - 1);
  constexpr int cN = I, cS = -I;
  const int cUS = (j > 1) ? -I : 0;
  const int m0 = (int)(m - (unsigned)st.tid);                     // first member of this warp's tile
  const double *const sm = st.sm + st.tid;
#ifdef __CUDA_ARCH__
  const bool leader = (st.tid == 0);
#else
  const bool leader = true;                                       // host emulation: every "thread" copies its own elements
#endif
  const void *const mTS2 = tm ? (const void *)tm->ts2 : nullptr, *const mTSA = tm ? (const void *)tm->tsA : nullptr,
                   *const mTSB = tm ? (const void *)tm->tsB : nullptr, *const mU3 = tm ? (const void *)tm->u3 : nullptr,
                   *const mU1 = tm ? (const void *)tm->u1 : nullptr;
  auto colcell = [&](const int lev, const int cell) {             // cell offset of stencil column `cell` at level lev
    return (cell == 0) ? 0 : (cell == 1) ? ((lev >= k1e) ? cE : 0) : (cell == 2) ? ((lev >= k1w) ? cW : 0)
           : (cell == 3) ? ((lev >= k1n) ? cN : 0) : ((lev >= k1s) ? cS : 0);
  };
  // Unit C of level lev: rows 0..9 = T,S one level up of the five columns, rows 10..14 = uE, vN, ww, uW, vS
  auto issueC = [&](const int lev) {
    if (!leader) return;
    const int b = (lev - k1c) & 1;
    stage_expect(st, b, (unsigned)(R::rowsC * NT * 8));
    const int lu = (lev < K) ? lev + 1 : K;
    const int cellu = (lu - 1) * IJ + c2;
    const int r0 = b * R::rowsC;
#pragma unroll
    for (int cell = 0; cell < 5; cell++)
      stage_box<MS>(st, b, r0 + R::rTS + 2 * cell, mTS2, v.ts_cur, m0, (cellu + colcell(lu, cell)) * L, 2);
    const int cellv = (lev - 1) * IJ + c2;
    stage_box<MS>(st, b, r0 + R::rU + 0, mU3, v.u, m0, cellv * 3, 3);
    stage_box<MS>(st, b, r0 + R::rU + 3, mU1, v.u, m0, (cellv + cW) * 3, 1);
    stage_box<MS>(st, b, r0 + R::rU + 4, mU1, v.u, m0, (cellv + cUS) * 3 + 1, 1);
  };
  auto issueA = [&](const int lev) {
    if (R::nA == 0 || !leader) return;
    stage_expect(st, 2, (unsigned)(R::rowsA * NT * 8));
    const int cell0 = (lev - 1) * IJ + c2;
#pragma unroll
    for (int cell = 0; cell < 5; cell++)
      stage_box<MS>(st, 2, R::rA + cell * R::nA, mTSA, v.ts_cur, m0, (cell0 + colcell(lev, cell)) * L + 2, R::nA);
  };
  auto issueB = [&](const int lev) {
    if (!leader) return;
    stage_expect(st, 3, (unsigned)(R::rowsB * NT * 8));
    const int cell0 = (lev - 1) * IJ + c2;
#pragma unroll
    for (int cell = 0; cell < 5; cell++)
      stage_box<MS>(st, 3, R::rB + cell * R::nB, mTSB, v.ts_cur, m0, (cell0 + colcell(lev, cell)) * L + R::lB0, R::nB);
  };

#ifdef __CUDA_ARCH__
  if (st.tid == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(st.bar);
    for (unsigned qb = 0; qb < 4; qb++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a + 8 * qb), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
#endif
  issueC(k1c);
  if (k1c < K) issueC(k1c + 1);
  issueA(k1c);
  issueB(k1c);

  // T,S of the five columns at the bottom level (direct loads, once per column)
  TS5 a;
  {
    const double *qC = v.ts_cur + ((long)(k1c - 1) * IJ + c2) * sC + m;
    const double *qE = qC + (long)colcell(k1c, 1) * sC, *qW = qC + (long)colcell(k1c, 2) * sC;
    const double *qN = qC + (long)colcell(k1c, 3) * sC, *qS = qC + (long)colcell(k1c, 4) * sC;
    a.tC = qC[0]; a.sC = qC[sL]; a.tE = qE[0]; a.sE = qE[sL]; a.tW = qW[0]; a.sW = qW[sL];
    a.tN = qN[0]; a.sN = qN[sL]; a.tS = qS[0]; a.sS = qS[sL];
  }
  double *wP = v.ts_new + ((long)(k1c - 2) * IJ + c2) * sC + m;   // level kk-1 of the new array
  double *rP = v.rho + ((long)(k1c - 2) * IJ + c2) * MS + m;
  double uc = 0.0, uE = 0.0, uW = 0.0, uN = 0.0, uS = 0.0, cZp = 0.0;
  double P[L], Q[L];
#pragma unroll
  for (int l = 0; l < L; l++) { P[l] = 0.0; Q[l] = 0.0; }
  unsigned par = 0;
  bool unstable = false;
  double rbelow = 0.0;

  for (int kk = k1c; kk <= K; kk++) {
    const bool top = (kk == K);
    const bool opE = kk >= k1e, opW = kk >= k1w, opN = kk >= k1n, opS = kk >= k1s;
    const int cb = (kk - k1c) & 1;
    stage_wait(st, cb, ((unsigned)(kk - k1c) >> 1) & 1u);
    const double *const smc = sm + cb * R::rowsC * NT;
    TS5 b;
    b.tC = smc[(R::rTS + 0) * NT]; b.sC = smc[(R::rTS + 1) * NT]; b.tE = smc[(R::rTS + 2) * NT]; b.sE = smc[(R::rTS + 3) * NT];
    b.tW = smc[(R::rTS + 4) * NT]; b.sW = smc[(R::rTS + 5) * NT]; b.tN = smc[(R::rTS + 6) * NT]; b.sN = smc[(R::rTS + 7) * NT];
    b.tS = smc[(R::rTS + 8) * NT]; b.sS = smc[(R::rTS + 9) * NT];
    const double vuE = smc[(R::rU + 0) * NT], vvN = smc[(R::rU + 1) * NT], vww = smc[(R::rU + 2) * NT], vuW = smc[(R::rU + 3) * NT],
                 vvS = smc[(R::rU + 4) * NT];
    ColCoef cf;
    col_coefs<K>(q, g, kk, opE, opW, opN, opS, a, b, vuE, vvN, vww, vuW, vvS, cf);
    const double hE = cf.hE, hW = cf.hW, hN = cf.hN, hS = cf.hS, hC = cf.hC;
    const double lc = cf.lc, lE = cf.lE, lW = cf.lW, lN = cf.lN, lS = cf.lS, cZ = cf.cZ;
    // unit C[cb] has been consumed (its values went through the coefficient arithmetic above): this warp, its only reader,
    // requests the unit two levels up
    stage_syncw();
    if (kk + 2 <= K) issueC(kk + 2);
    const bool stv = kk > k1c;
    double tnew = 0.0, snew = 0.0;
#define CG_TRACERW(l, cc, EE, WW, NN, SS)                                                     \
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
    CG_TRACERW(0, a.tC, a.tE, a.tW, a.tN, a.tS)
    CG_TRACERW(1, a.sC, a.sE, a.sW, a.sN, a.sS)
    if (R::nA > 0) stage_wait(st, 2, par);
#pragma unroll
    for (int l = 2; l < R::lB0; l++) {
      const int r = R::rA + (l - 2);
      CG_TRACERW(l, sm[(r + 0 * R::nA) * NT], sm[(r + 1 * R::nA) * NT], sm[(r + 2 * R::nA) * NT], sm[(r + 3 * R::nA) * NT],
                 sm[(r + 4 * R::nA) * NT])
    }
    stage_syncw();                                           // every lane has read unit A: re-request it for the next level
    if (!top) issueA(kk + 1);
    stage_wait(st, 3, par);
#pragma unroll
    for (int l = R::lB0; l < L; l++) {
      const int r = R::rB + (l - R::lB0);
      CG_TRACERW(l, sm[(r + 0 * R::nB) * NT], sm[(r + 1 * R::nB) * NT], sm[(r + 2 * R::nB) * NT], sm[(r + 3 * R::nB) * NT],
                 sm[(r + 4 * R::nB) * NT])
    }
#undef CG_TRACERW
    stage_syncw();
    if (!top) issueB(kk + 1);
    par ^= 1u;
    if (stv) {
      const double r = ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);   // :2638
      rP[0] = r;
      if (kk - 1 > k1c) unstable = unstable || !(r < rbelow);
      rbelow = r;
    }
    a = b;
    uc = cf.nuc; uE = cf.nuE; uW = cf.nuW; uN = cf.nuN; uS = cf.nuS; cZp = cZ;
    wP += sK; rP += rK;
  }
  {
    double tnew = 0.0, snew = 0.0;
#pragma unroll
    for (int l = 0; l < L; l++) {
      double tn = Q[l];
      if (l < 2) tn -= v.tsflux[((long)l * IJ + c2) * MS + m] * cZp;
      wP[l * sL] = tn;
      if (l == 0) tnew = tn;
      if (l == 1) snew = tn;
    }
    const double r = ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);
    rP[0] = r;
    if (K > k1c) unstable = unstable || !(r < rbelow);
    if (v.comask) v.comask[(long)c2 * MS + m] = unstable ? 1u : 0u;
    if (v.sst) {
      v.sst[(long)c2 * MS + m] = tnew;
      v.sst[((long)IJ + c2) * MS + m] = snew;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Pipelined form of the column kernel (production): the per-cell coefficients are computed ONE LEVEL AHEAD.
//
// In tstep_column the 16 coefficients of a cell -- upstream weights (four divisions), density slopes on the four isoneutral
// stencils, slope limiter (two more divisions in sequence) -- are a dependent chain of ~60 fp64 operations at the head of
// every level, and the tracers' 320 independent FMAs can only start when it has finished: at 8 warps per SM the chain's
// latency is exposed (ncu: 36 % of the stall cycles "wait", IPC 1.2).  Here the thread computes the coefficients of level
// kk+1 WHILE it applies those of level kk: the two instruction streams share basic blocks (horizontal coefficients next to
// tracers 0..L/2-1, vertical / isoneutral coefficients next to tracers L/2..L-1), so the scheduler fills the chain's latency
// with tracer FMAs.  Consequences for the staging: the T,S / velocity unit C is consumed one level earlier (issued two levels
// ahead), and T and S of level kk come from shared memory like every other tracer (unit A = tracers 0..L/2-1) because the
// registers now hold T,S of levels kk+1 and kk+2 for the coefficients.  The slope limiter is taken from 1/tv1 instead of
// 1/sl^2 (slim = (4 ssmax dzrho^2 / tv1)^2), which makes its division independent of 1/dzrho.
// Arithmetic per tracer-cell, face formulation, stability flag and SST export are those of tstep_column.
template <int L>
struct ColRows2 {
  static constexpr int nA = L / 2, nB = L - nA;
  static constexpr int rowsC = 15, rTS = 0, rU = 10;
  static constexpr int rA = 2 * rowsC, rowsA = 5 * nA;      // + cell * nA + l
  static constexpr int rB = rA + rowsA, rowsB = 5 * nB;     // + cell * nB + (l - nA)
  static constexpr int rows = rB + rowsB;
};

// vertical face kk+1/2 as col_coefs_v, with the two divisions of the slope limiter made independent of each other
template <int K>
CG_HD void col_coefs_v2(const ColK &q, const GridC &g, const int kk, const TS5 &a, const TS5 &b, const double vww, double &lc_,
                        double &lE, double &lW, double &lN, double &lS, double &nuc_, double &nuE, double &nuW, double &nuN,
                        double &nuS) {
  const bool top = (kk == K);
  const double mT = top ? 0.0 : 1.0;
  const double ec1 = q.ec1, ec2 = q.ec2, ec3 = q.ec3, ec4 = q.ec4, gxx = q.gxx, gyS = q.gyS, gyN = q.gyN;
  const double rdza = top ? 0.0 : g.rdza[kk];
  const double pA = vww * g.dza[kk] * q.rdiffv, uA_ = pA * col_rcp(2.0 + fabs(pA)), gA = vww * 0.5 * mT, dA = rdza * q.diffv;
  double nuc = gA * (1.0 - uA_) - dA;
  double lc = gA * (1.0 + uA_) + dA;
  const double tatw = 0.5 * (a.tC + b.tC);
  const double tec = -ec1 - ec3 * tatw * 2 - ec4 * tatw * tatw * 3;
  const double dzrho = (ec2 * (b.sC - a.sC) - tec * (b.tC - a.tC)) * rdza;
  const bool iso = dzrho < -1.0e-12;                     // false at the top level (rdza = 0)
  const double dzs = iso ? dzrho : -1.0, mI = iso ? 1.0 : 0.0;
  const double x0 = ec2 * ((a.sC - a.sW) * gxx) - tec * ((a.tC - a.tW) * gxx);
  const double x1 = ec2 * ((a.sE - a.sC) * gxx) - tec * ((a.tE - a.tC) * gxx);
  const double x2 = ec2 * ((b.sC - b.sW) * gxx) - tec * ((b.tC - b.tW) * gxx);
  const double x3 = ec2 * ((b.sE - b.sC) * gxx) - tec * ((b.tE - b.tC) * gxx);
  const double y0 = ec2 * ((a.sC - a.sS) * gyS) - tec * ((a.tC - a.tS) * gyS);
  const double y1 = ec2 * ((a.sN - a.sC) * gyN) - tec * ((a.tN - a.tC) * gyN);
  const double y2 = ec2 * ((b.sC - b.sS) * gyS) - tec * ((b.tC - b.tS) * gyS);
  const double y3 = ec2 * ((b.sN - b.sC) * gyN) - tec * ((b.tN - b.tC) * gyN);
  const double tv1 = (((x0 * x0 + y0 * y0) + (x1 * x1 + y1 * y1)) + (x2 * x2 + y2 * y2)) + (x3 * x3 + y3 * y3);
  const double rdz = col_rcp(dzs), rdz2 = rdz * rdz;
  const double sl = 0.25 * tv1 * rdz2, ssm = g.ssmax[kk];
  // slim = ssmax^2 / sl^2 = (4 ssmax dzrho^2 / tv1)^2: 1 / tv1 does not wait for 1 / dzrho.  sl > ssmax implies tv1 > 0; where
  // the limiter is not active the value is discarded by the select (tv1 == 0 gives a NaN there, never used)
  const double rt = col_rcp(tv1 > 0.0 ? tv1 : 1.0);
  const double sq = 4.0 * ssm * (dzs * dzs) * rt;
  const double slim = (sl > ssm) ? sq * sq : 1.0;
  const double cf = 0.25 * slim * q.diff1 * rdz2 * mI;
  const double g2 = 2.0 * dzs * cf, gX = g2 * gxx, gS = g2 * gyS, gN = g2 * gyN;
  const double s2 = tv1 * cf * rdza;
  const double wx0 = x0 * gX, wx1 = x1 * gX, wx2 = x2 * gX, wx3 = x3 * gX;
  const double wy0 = y0 * gS, wy1 = y1 * gN, wy2 = y2 * gS, wy3 = y3 * gN;
  lc += (wx0 - wx1) + (wy0 - wy1) + s2;
  nuc += (wx2 - wx3) + (wy2 - wy3) - s2;
  lc_ = lc; nuc_ = nuc;
  lW = -wx0; lE = wx1; lS = -wy0; lN = wy1;
  nuW = -wx2; nuE = wx3; nuS = -wy2; nuN = wy3;
}

template <int I, int J, int K, int L, int MS, int NT>
CG_HD void tstep_column2(const Dev &v, const GridC &g, const int c2, const unsigned m, const ColStage &st) {
  static_assert(NT == MS, "a block covers all members of one column");
  using R = ColRows2<L>;
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC;
  constexpr long uC3 = 3L * MS, uK = (long)I * J * uC3, rK = (long)I * J * MS;
  const int i = c2 % I + 1, j = c2 / I + 1;
#define CGC_K1(ii, jj) ((int)v.k1[(ii) + (I + 2) * (jj)])
  const int k1c = CGC_K1(i, j);
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CGC_K1(ip, j), k1w = CGC_K1(im, j), k1n = CGC_K1(i, j + 1), k1s = CGC_K1(i, j - 1);
#undef CGC_K1
  const ColK q = col_consts<I, J>(v, g, m, j);
  const double ec1 = q.ec1, ec2 = q.ec2, ec3 = q.ec3, ec4 = q.ec4;
  const long dE = (i < I) ? sC : -(long)(I - 1) * sC, dW = (i > 1) ? -sC : (long)(I - 1) * sC;
  constexpr long dN = (long)I * sC, dS = -(long)I * sC;
  const long dUW = (i > 1) ? -uC3 : (long)(I - 1) * uC3, dUS = (j > 1) ? -(long)I * uC3 : 0;
  const double *const ts0 = v.ts_cur + (long)c2 * sC;
  const double *const u0 = v.u + (long)c2 * uC3;
  const double *const sm = st.sm + st.tid;

  // unit C of level `lev`: T,S one level up (the level itself at the top) of the five columns + the five velocities of `lev`
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
    stage_expect(st, 2, (unsigned)(R::rowsA * NT * 8));
    const double *c0 = ts0 + (long)(lev - 1) * sK;
    stage_copy<NT>(st, 2, R::rA + 0 * R::nA, c0, R::nA);
    stage_copy<NT>(st, 2, R::rA + 1 * R::nA, c0 + ((lev >= k1e) ? dE : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 2 * R::nA, c0 + ((lev >= k1w) ? dW : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 3 * R::nA, c0 + ((lev >= k1n) ? dN : 0), R::nA);
    stage_copy<NT>(st, 2, R::rA + 4 * R::nA, c0 + ((lev >= k1s) ? dS : 0), R::nA);
  };
  auto issueB = [&](const int lev) {
    stage_expect(st, 3, (unsigned)(R::rowsB * NT * 8));
    const double *c0 = ts0 + (long)(lev - 1) * sK + R::nA * sL;
    stage_copy<NT>(st, 3, R::rB + 0 * R::nB, c0, R::nB);
    stage_copy<NT>(st, 3, R::rB + 1 * R::nB, c0 + ((lev >= k1e) ? dE : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 2 * R::nB, c0 + ((lev >= k1w) ? dW : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 3 * R::nB, c0 + ((lev >= k1n) ? dN : 0), R::nB);
    stage_copy<NT>(st, 3, R::rB + 4 * R::nB, c0 + ((lev >= k1s) ? dS : 0), R::nB);
  };
  auto loadTS = [&](const int cb, TS5 &o) {
    const double *const smc = sm + cb * R::rowsC * NT;
    o.tC = smc[(R::rTS + 0) * NT]; o.sC = smc[(R::rTS + 1) * NT]; o.tE = smc[(R::rTS + 2) * NT]; o.sE = smc[(R::rTS + 3) * NT];
    o.tW = smc[(R::rTS + 4) * NT]; o.sW = smc[(R::rTS + 5) * NT]; o.tN = smc[(R::rTS + 6) * NT]; o.sN = smc[(R::rTS + 7) * NT];
    o.tS = smc[(R::rTS + 8) * NT]; o.sS = smc[(R::rTS + 9) * NT];
  };

  stage_init(st);
  if (stage_issuer<NT>(st, 0)) {
    issueC(k1c);
    if (k1c < K) issueC(k1c + 1);
  }
  if (stage_issuer<NT>(st, 1)) issueA(k1c);
  if (stage_issuer<NT>(st, 2)) issueB(k1c);
  // T,S of the five columns at the bottom level (direct loads, once per column)
  TS5 a;
  {
    const double *qC = ts0 + (long)(k1c - 1) * sK + m;
    const double *qE = qC + ((k1c >= k1e) ? dE : 0), *qW = qC + ((k1c >= k1w) ? dW : 0);
    const double *qN = qC + ((k1c >= k1n) ? dN : 0), *qS = qC + ((k1c >= k1s) ? dS : 0);
    a.tC = qC[0]; a.sC = qC[sL]; a.tE = qE[0]; a.sE = qE[sL]; a.tW = qW[0]; a.sW = qW[sL];
    a.tN = qN[0]; a.sN = qN[sL]; a.tS = qS[0]; a.sS = qS[sL];
  }
  // prologue: the coefficients of the bottom level from unit C(k1c) (buffer 0, first fill)
  double hE, hW, hN, hS, hC, lc, lE, lW, lN, lS, cZ;          // level kk: horizontal faces, lower half of face kk+1/2, dt/dz
  double nuc, nuE, nuW, nuN, nuS;                            // level kk: upper half of face kk+1/2 (used at level kk+1)
  {
    stage_wait(st, 0, 0u);
    TS5 b;
    loadTS(0, b);
    const double *const smc = sm;
    const double vuE = smc[(R::rU + 0) * NT], vvN = smc[(R::rU + 1) * NT], vww = smc[(R::rU + 2) * NT], vuW = smc[(R::rU + 3) * NT],
                 vvS = smc[(R::rU + 4) * NT];
    col_coefs_h(q, k1c >= k1e, k1c >= k1w, k1c >= k1n, k1c >= k1s, vuE, vvN, vuW, vvS, hE, hW, hN, hS, hC);
    col_coefs_v2<K>(q, g, k1c, a, b, vww, lc, lE, lW, lN, lS, nuc, nuE, nuW, nuN, nuS);
    cZ = q.dt * g.rdz[k1c];
    a = b;                                                   // T,S of level k1c+1
    stage_sync();
    if (k1c + 2 <= K && stage_issuer<NT>(st, 0)) issueC(k1c + 2);
  }
  double *wP = v.ts_new + ((long)(k1c - 2) * (I * J) + c2) * sC + m;   // level kk-1 of the new array
  double *rP = v.rho + ((long)(k1c - 2) * (I * J) + c2) * MS + m;
  double uc = 0.0, uE = 0.0, uW = 0.0, uN = 0.0, uS = 0.0, cZp = 0.0;   // upper half of the face below, cZ of the level below
  double P[L], Q[L];
#pragma unroll
  for (int l = 0; l < L; l++) { P[l] = 0.0; Q[l] = 0.0; }
  unsigned par = 0;
  bool unstable = false;
  double rbelow = 0.0;

  for (int kk = k1c; kk <= K; kk++) {
    const bool top = (kk == K);
    const bool stv = kk > k1c;
    // inputs of the NEXT level's coefficients: T,S of level kk+2 and the velocities of level kk+1 (unit C(kk+1))
    const int kn = top ? K : kk + 1;
    const int cb = (kn - k1c) & 1;
    TS5 b = a;
    double vuE = 0.0, vvN = 0.0, vww = 0.0, vuW = 0.0, vvS = 0.0;
    if (!top) {
      stage_wait(st, cb, ((unsigned)(kn - k1c) >> 1) & 1u);
      loadTS(cb, b);
      const double *const smc = sm + cb * R::rowsC * NT;
      vuE = smc[(R::rU + 0) * NT]; vvN = smc[(R::rU + 1) * NT]; vww = smc[(R::rU + 2) * NT]; vuW = smc[(R::rU + 3) * NT];
      vvS = smc[(R::rU + 4) * NT];
    }
    stage_wait(st, 2, par);
    // ---- segment 1: horizontal coefficients of level kk+1 next to tracers 0 .. nA-1 of level kk
    double hE1 = 0.0, hW1 = 0.0, hN1 = 0.0, hS1 = 0.0, hC1 = 0.0;
    if (!top) col_coefs_h(q, kn >= k1e, kn >= k1w, kn >= k1n, kn >= k1s, vuE, vvN, vuW, vvS, hE1, hW1, hN1, hS1, hC1);
    double tnew = 0.0, snew = 0.0;
#define CG_TRACER2(l, cc, EE, WW, NN, SS)                                                     \
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
#pragma unroll
    for (int l = 0; l < R::nA; l++) {
      const int r = R::rA + l;
      CG_TRACER2(l, sm[(r + 0 * R::nA) * NT], sm[(r + 1 * R::nA) * NT], sm[(r + 2 * R::nA) * NT], sm[(r + 3 * R::nA) * NT],
                 sm[(r + 4 * R::nA) * NT])
    }
    stage_sync();                                            // unit A and unit C(kk+1) are consumed
    if (kk + 3 <= K && stage_issuer<NT>(st, 0)) issueC(kk + 3);
    if (!top && stage_issuer<NT>(st, 1)) issueA(kk + 1);
    stage_wait(st, 3, par);
    // ---- segment 2: vertical / isoneutral coefficients of level kk+1 next to tracers nA .. L-1 of level kk
    double lc1 = 0.0, lE1 = 0.0, lW1 = 0.0, lN1 = 0.0, lS1 = 0.0, nuc1 = 0.0, nuE1 = 0.0, nuW1 = 0.0, nuN1 = 0.0, nuS1 = 0.0;
    if (!top) col_coefs_v2<K>(q, g, kn, a, b, vww, lc1, lE1, lW1, lN1, lS1, nuc1, nuE1, nuW1, nuN1, nuS1);
#pragma unroll
    for (int l = R::nA; l < L; l++) {
      const int r = R::rB + (l - R::nA);
      CG_TRACER2(l, sm[(r + 0 * R::nB) * NT], sm[(r + 1 * R::nB) * NT], sm[(r + 2 * R::nB) * NT], sm[(r + 3 * R::nB) * NT],
                 sm[(r + 4 * R::nB) * NT])
    }
#undef CG_TRACER2
    stage_sync();
    if (!top && stage_issuer<NT>(st, 2)) issueB(kk + 1);
    par ^= 1u;
    if (stv) {
      const double r = ec1 * tnew + ec2 * snew + ec3 * (tnew * tnew) + ec4 * (tnew * tnew * tnew);   // :2638
      rP[0] = r;
      if (kk - 1 > k1c) unstable = unstable || !(r < rbelow);
      rbelow = r;
    }
    // ---- shift one level up
    uc = nuc; uE = nuE; uW = nuW; uN = nuN; uS = nuS; cZp = cZ;
    hE = hE1; hW = hW1; hN = hN1; hS = hS1; hC = hC1;
    lc = lc1; lE = lE1; lW = lW1; lN = lN1; lS = lS1;
    nuc = nuc1; nuE = nuE1; nuW = nuW1; nuN = nuN1; nuS = nuS1;
    cZ = q.dt * g.rdz[kn];
    a = b;
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
    if (K > k1c) unstable = unstable || !(r < rbelow);
    if (v.comask) v.comask[(long)c2 * MS + m] = unstable ? 1u : 0u;
    if (v.sst) {
      v.sst[(long)c2 * MS + m] = tnew;
      v.sst[((long)(I * J) + c2) * MS + m] = snew;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Split form of the column kernel: TWO threads per (member, column), 2 * NT threads per block.  Half 0 computes the
// horizontal coefficients of the cell and carries tracers 0 .. L/2-1 (T and S among them), half 1 computes the vertical /
// isoneutral coefficients and carries tracers L/2 .. L-1.  The halves exchange their coefficients through the rows of
// the T,S / velocity staging buffer they have just consumed (each half overwrites only the rows it alone reads, in its own
// member column), so per thread the state is half the tracers' (P, Q) plus half of the coefficient work: <= 128
// registers, 16 warps per SM instead of 8, and the dependent chain of a level is split in two.  Every tracer (T, S too)
// comes from shared memory; each half refills its own two sub-units (named barrier over the half), the block barrier in
// the middle of a level publishes the coefficients, the one at its end frees the buffers.
// BOTH = true (host test harness): one caller plays both halves of a member in turn.
template <int L>
struct SplitRows {
  static constexpr int LH = L / 2, LQ = L / 4;       // tracers per half / per staging sub-unit
  static constexpr int rowsC = 16;                   // T,S one level up of 5 cells (10) | uE vN ww uW vS (5) | spare (1)
  static constexpr int rTS = 0, rU = 10, rX = 15;
  static constexpr int rowsX = 5 * LQ;               // sub-unit (half hh, u): rows rUnit + (2 hh + u) rowsX + cell LQ + s
  static constexpr int rUnit = 2 * rowsC;
  static constexpr int rows = rUnit + 4 * rowsX;
  static constexpr int nbar = 6;                     // mbarriers: C0, C1, then 2 + 2 hh + u
};

template <int I, int J, int K, int L, int MS, int NT, bool BOTH>
CG_HD void tstep_column_split(const Dev &v, const GridC &g, const int c2, const int tid, const ColStage &st) {
  static_assert(NT == MS, "a block covers all members of one column");
  static_assert(L % 4 == 0 && L >= 8, "tracers are split in four staging sub-units");
  using R = SplitRows<L>;
  constexpr int LH = R::LH, LQ = R::LQ, NH = BOTH ? 2 : 1;
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC;
  constexpr long uC3 = 3L * MS, uK = (long)I * J * uC3, rK = (long)I * J * MS;
  const int h = BOTH ? 0 : tid / NT;
  const unsigned m = BOTH ? (unsigned)tid : (unsigned)(tid - h * NT);
  const int hlo = BOTH ? 0 : h, hhi = BOTH ? 2 : h + 1;
  const int i = c2 % I + 1, j = c2 / I + 1;
#define CGC_K1(ii, jj) ((int)v.k1[(ii) + (I + 2) * (jj)])
  const int k1c = CGC_K1(i, j);
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CGC_K1(ip, j), k1w = CGC_K1(im, j), k1n = CGC_K1(i, j + 1), k1s = CGC_K1(i, j - 1);
#undef CGC_K1
  const ColK q = col_consts<I, J>(v, g, m, j);
  const long dE = (i < I) ? sC : -(long)(I - 1) * sC, dW = (i > 1) ? -sC : (long)(I - 1) * sC;
  constexpr long dN = (long)I * sC, dS = -(long)I * sC;
  const long dUW = (i > 1) ? -uC3 : (long)(I - 1) * uC3, dUS = (j > 1) ? -(long)I * uC3 : 0;
  const double *const ts0 = v.ts_cur + (long)c2 * sC;
  const double *const u0 = v.u + (long)c2 * uC3;
  double *const sm = st.sm + m;
#ifdef __CUDA_ARCH__
  const bool lead0 = (tid == 0), lead1 = (tid == NT);
#else
  const bool lead0 = true, lead1 = true;   // host emulation: every caller copies its own elements
#endif

  auto issueC = [&](const int lev) {
    const int b = (lev - k1c) & 1;
    stage_expect(st, b, (unsigned)(15 * NT * 8));
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
  auto issueX = [&](const int hh, const int u, const int lev) {
    const int bar = 2 + 2 * hh + u;
    stage_expect(st, bar, (unsigned)(R::rowsX * NT * 8));
    const double *c0 = ts0 + (long)(lev - 1) * sK + (long)(hh * LH + u * LQ) * sL;
    const int r0 = R::rUnit + (2 * hh + u) * R::rowsX;
    stage_copy<NT>(st, bar, r0 + 0 * LQ, c0, LQ);
    stage_copy<NT>(st, bar, r0 + 1 * LQ, c0 + ((lev >= k1e) ? dE : 0), LQ);
    stage_copy<NT>(st, bar, r0 + 2 * LQ, c0 + ((lev >= k1w) ? dW : 0), LQ);
    stage_copy<NT>(st, bar, r0 + 3 * LQ, c0 + ((lev >= k1n) ? dN : 0), LQ);
    stage_copy<NT>(st, bar, r0 + 4 * LQ, c0 + ((lev >= k1s) ? dS : 0), LQ);
  };

  stage_init(st, R::nbar);
  if (lead0) {
    issueC(k1c);
    if (k1c < K) issueC(k1c + 1);
    issueX(0, 0, k1c);
    issueX(0, 1, k1c);
  }
  if (lead1) {
    issueX(1, 0, k1c);
    issueX(1, 1, k1c);
  }

  double *wP = v.ts_new + ((long)(k1c - 2) * (I * J) + c2) * sC + m;   // level kk-1 of the new array
  double *rP = v.rho + ((long)(k1c - 2) * (I * J) + c2) * MS + m;
  double uc = 0.0, uE = 0.0, uW = 0.0, uN = 0.0, uS = 0.0, cZp = 0.0;
  double P[NH][LH], Q[NH][LH];
#pragma unroll
  for (int a_ = 0; a_ < NH; a_++)
#pragma unroll
    for (int s_ = 0; s_ < LH; s_++) { P[a_][s_] = 0.0; Q[a_][s_] = 0.0; }
  unsigned par = 0;

  for (int kk = k1c; kk <= K; kk++) {
    const bool top = (kk == K), stv = kk > k1c;
    const int cb = (kk - k1c) & 1;
    stage_wait(st, cb, ((unsigned)(kk - k1c) >> 1) & 1u);
    double *const smc = sm + cb * R::rowsC * NT;
    // ---- coefficients: each half its share, written over the rows it has just read
    for (int hh = hlo; hh < hhi; hh++) {
      if (hh == 0) {
        const double vuE = smc[(R::rU + 0) * NT], vvN = smc[(R::rU + 1) * NT], vuW = smc[(R::rU + 3) * NT], vvS = smc[(R::rU + 4) * NT];
        double hE, hW, hN, hS, hC;
        col_coefs_h(q, kk >= k1e, kk >= k1w, kk >= k1n, kk >= k1s, vuE, vvN, vuW, vvS, hE, hW, hN, hS, hC);
        smc[(R::rU + 0) * NT] = hE; smc[(R::rU + 1) * NT] = hN; smc[(R::rU + 3) * NT] = hW; smc[(R::rU + 4) * NT] = hS;
        smc[R::rX * NT] = hC;
      } else {
        stage_wait(st, 2, par);   // T,S of level kk: tracers 0, 1 of sub-unit (0, 0), shared read-only with half 0
        const double *const sa = sm + R::rUnit * NT;
        TS5 a, b;
        a.tC = sa[(0 * LQ + 0) * NT]; a.sC = sa[(0 * LQ + 1) * NT]; a.tE = sa[(1 * LQ + 0) * NT]; a.sE = sa[(1 * LQ + 1) * NT];
        a.tW = sa[(2 * LQ + 0) * NT]; a.sW = sa[(2 * LQ + 1) * NT]; a.tN = sa[(3 * LQ + 0) * NT]; a.sN = sa[(3 * LQ + 1) * NT];
        a.tS = sa[(4 * LQ + 0) * NT]; a.sS = sa[(4 * LQ + 1) * NT];
        b.tC = smc[(R::rTS + 0) * NT]; b.sC = smc[(R::rTS + 1) * NT]; b.tE = smc[(R::rTS + 2) * NT]; b.sE = smc[(R::rTS + 3) * NT];
        b.tW = smc[(R::rTS + 4) * NT]; b.sW = smc[(R::rTS + 5) * NT]; b.tN = smc[(R::rTS + 6) * NT]; b.sN = smc[(R::rTS + 7) * NT];
        b.tS = smc[(R::rTS + 8) * NT]; b.sS = smc[(R::rTS + 9) * NT];
        const double vww = smc[(R::rU + 2) * NT];
        double lc, lE, lW, lN, lS, nuc, nuE, nuW, nuN, nuS;
        col_coefs_v<K>(q, g, kk, a, b, vww, lc, lE, lW, lN, lS, nuc, nuE, nuW, nuN, nuS);
        smc[(R::rTS + 0) * NT] = lc; smc[(R::rTS + 1) * NT] = lE; smc[(R::rTS + 2) * NT] = lW; smc[(R::rTS + 3) * NT] = lN;
        smc[(R::rTS + 4) * NT] = lS; smc[(R::rTS + 5) * NT] = nuc; smc[(R::rTS + 6) * NT] = nuE; smc[(R::rTS + 7) * NT] = nuW;
        smc[(R::rTS + 8) * NT] = nuN; smc[(R::rTS + 9) * NT] = nuS;
      }
    }
    stage_sync();                                           // the coefficients of this level are published
    const double lc = smc[(R::rTS + 0) * NT], lE = smc[(R::rTS + 1) * NT], lW = smc[(R::rTS + 2) * NT], lN = smc[(R::rTS + 3) * NT],
                 lS = smc[(R::rTS + 4) * NT];
    const double hE = smc[(R::rU + 0) * NT], hN = smc[(R::rU + 1) * NT], hW = smc[(R::rU + 3) * NT], hS = smc[(R::rU + 4) * NT],
                 hC = smc[R::rX * NT];
    const double cZ = q.dt * g.rdz[kk];
    // ---- tracers: one tracer-cell = 15 FMA + 5
    for (int hh = hlo; hh < hhi; hh++) {
      const int hs = BOTH ? hh : 0;
      double tnew = 0.0, snew = 0.0;
#pragma unroll
      for (int u = 0; u < 2; u++) {
        stage_wait(st, 2 + 2 * hh + u, par);
        const double *const su = sm + (R::rUnit + (2 * hh + u) * R::rowsX) * NT;
        double *const w = wP + (long)(hh * LH + u * LQ) * sL;
#pragma unroll
        for (int s_ = 0; s_ < LQ; s_++) {
          const int sl = u * LQ + s_;
          const double c = su[(0 * LQ + s_) * NT], E = su[(1 * LQ + s_) * NT], W = su[(2 * LQ + s_) * NT], N = su[(3 * LQ + s_) * NT],
                       S = su[(4 * LQ + s_) * NT];
          const double fab = P[hs][sl] + (uc * c + uE * E + uW * W + uN * N + uS * S);
          if (stv) {
            const double tn = Q[hs][sl] - fab * cZp;
            w[s_ * sL] = tn;
            if (sl == 0) tnew = tn;
            if (sl == 1) snew = tn;
          }
          const double Hh = hC * c + hE * E + hW * W + hN * N + hS * S;
          Q[hs][sl] = (c - Hh) + fab * cZ;
          P[hs][sl] = lc * c + lE * E + lW * W + lN * N + lS * S;
        }
        if (u == 0) {
          stage_sync_half(hh, NT);                          // this half is done with its first sub-unit
          if ((hh == 0 ? lead0 : lead1) && !top) { stage_fence(); issueX(hh, 0, kk + 1); }
        }
      }
      if (hh == 0 && stv) {
        const double r = q.ec1 * tnew + q.ec2 * snew + q.ec3 * (tnew * tnew) + q.ec4 * (tnew * tnew * tnew);   // :2638
        rP[0] = r;
      }
    }
    const double nuc = smc[(R::rTS + 5) * NT], nuE = smc[(R::rTS + 6) * NT], nuW = smc[(R::rTS + 7) * NT], nuN = smc[(R::rTS + 8) * NT],
                 nuS = smc[(R::rTS + 9) * NT];
    stage_sync();                                           // every thread is done with buffer C[cb] and the second sub-units
    if (lead0) {
      stage_fence();
      if (kk + 2 <= K) issueC(kk + 2);
      if (!top) issueX(0, 1, kk + 1);
    }
    if (lead1 && !top) { stage_fence(); issueX(1, 1, kk + 1); }
    par ^= 1u;
    uc = nuc; uE = nuE; uW = nuW; uN = nuN; uS = nuS; cZp = cZ;
    wP += sK; rP += rK;
  }
  // ---- top level: the flux through the surface is the boundary condition ts(1:2,:,:,maxk+1)   (:2550-2552)
  for (int hh = hlo; hh < hhi; hh++) {
    const int hs = BOTH ? hh : 0;
    double tnew = 0.0, snew = 0.0;
#pragma unroll
    for (int sl = 0; sl < LH; sl++) {
      double tn = Q[hs][sl];
      if (hh == 0 && sl < 2) tn -= v.tsflux[((long)sl * (I * J) + c2) * MS + m] * cZp;
      wP[(long)(hh * LH + sl) * sL] = tn;
      if (sl == 0) tnew = tn;
      if (sl == 1) snew = tn;
    }
    if (hh == 0) {
      const double r = q.ec1 * tnew + q.ec2 * snew + q.ec3 * (tnew * tnew) + q.ec4 * (tnew * tnew * tnew);
      rP[0] = r;
    }
  }
}

// T, S pre-pass of the mix-on-write form of the tracer step: one thread = (member, wet column).  The thread marches up
// the column with the same per-cell coefficients as tstep_column (col_coefs), applies them to T and S only, evaluates
// rho (goldstein.f90:2638), runs the convective-adjustment decisions on its column while it is still thread-private
// (co_decide_core), and writes the final T, S, rho of every wet level, SST/SSS, cost and the region map v.comask that the
// passive pass (tstep_column<PV = true>) applies on write.
template <int I, int J, int K, int L, int MS, bool ALL>
CG_HD void co_decide_core(const Dev &v, const GridC &g, const int c2, const unsigned m, const int k1c, double *tt, double *ss,
                          double *rl, unsigned &in, unsigned &topb, unsigned &botb, double *rdzt, const int st = 1,
                          double *dzm_ext = nullptr);

template <int I, int J, int K, int L, int MS>
CG_HD void ts_pre_column(const Dev &v, const GridC &g, const int c2, const unsigned m) {
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC;
  constexpr long uC3 = 3L * MS, uK = (long)I * J * uC3;
  const int i = c2 % I + 1, j = c2 / I + 1;
#define CGC_K1(ii, jj) ((int)v.k1[(ii) + (I + 2) * (jj)])
  const int k1c = CGC_K1(i, j);
  const int ip = (i < I) ? i + 1 : 1, im = (i > 1) ? i - 1 : I;
  const int k1e = CGC_K1(ip, j), k1w = CGC_K1(im, j), k1n = CGC_K1(i, j + 1), k1s = CGC_K1(i, j - 1);
#undef CGC_K1
  const ColK q = col_consts<I, J>(v, g, m, j);
  const long dE = (i < I) ? sC : -(long)(I - 1) * sC, dW = (i > 1) ? -sC : (long)(I - 1) * sC;
  constexpr long dN = (long)I * sC, dS = -(long)I * sC;
  const long dUW = (i > 1) ? -uC3 : (long)(I - 1) * uC3, dUS = (j > 1) ? -(long)I * uC3 : 0;
  const double *const ts0 = v.ts_cur + (long)c2 * sC + m;
  const double *const u0 = v.u + (long)c2 * uC3 + m;
  auto loadTS = [&](const int lev, TS5 &o) {
    const double *qC = ts0 + (long)(lev - 1) * sK;
    const double *qE = qC + ((lev >= k1e) ? dE : 0), *qW = qC + ((lev >= k1w) ? dW : 0);
    const double *qN = qC + ((lev >= k1n) ? dN : 0), *qS = qC + ((lev >= k1s) ? dS : 0);
    o.tC = qC[0]; o.sC = qC[sL]; o.tE = qE[0]; o.sE = qE[sL]; o.tW = qW[0]; o.sW = qW[sL];
    o.tN = qN[0]; o.sN = qN[sL]; o.tS = qS[0]; o.sS = qS[sL];
  };
  struct Vel { double uE, vN, ww, uW, vS; };
  auto loadU = [&](const int lev, Vel &o) {
    const double *pu = u0 + (long)(lev - 1) * uK;
    o.uE = pu[0]; o.vN = pu[sL]; o.ww = pu[2 * sL]; o.uW = pu[dUW]; o.vS = pu[dUS + sL];
  };
  double tt[K + 2], ss[K + 2], rl[K + 2];
#pragma unroll
  for (int k = 0; k < K + 2; k++) { tt[k] = 0.0; ss[k] = 0.0; rl[k] = 0.0; }
  TS5 a, b;
  Vel w;
  loadTS(k1c, a);
  loadTS((k1c < K) ? k1c + 1 : K, b);
  loadU(k1c, w);
  double uc = 0.0, uE = 0.0, uW = 0.0, uN = 0.0, uS = 0.0, cZp = 0.0, PT = 0.0, QT = 0.0, PS = 0.0, QS = 0.0;
  for (int kk = k1c; kk <= K; kk++) {
    // the rows of the next level are requested before this level's dependent chain starts
    TS5 nb = b;
    Vel nw = w;
    if (kk < K) { loadTS((kk + 2 <= K) ? kk + 2 : K, nb); loadU(kk + 1, nw); }
    ColCoef cf;
    col_coefs<K>(q, g, kk, kk >= k1e, kk >= k1w, kk >= k1n, kk >= k1s, a, b, w.uE, w.vN, w.ww, w.uW, w.vS, cf);
    const double fabT = PT + (uc * a.tC + uE * a.tE + uW * a.tW + uN * a.tN + uS * a.tS);
    const double fabS = PS + (uc * a.sC + uE * a.sE + uW * a.sW + uN * a.sN + uS * a.sS);
    if (kk > k1c) {
      const double tn = QT - fabT * cZp, sn = QS - fabS * cZp;
      tt[kk - 1] = tn; ss[kk - 1] = sn;
      rl[kk - 1] = q.ec1 * tn + q.ec2 * sn + q.ec3 * (tn * tn) + q.ec4 * (tn * tn * tn);   // :2638
    }
    const double HT = cf.hC * a.tC + cf.hE * a.tE + cf.hW * a.tW + cf.hN * a.tN + cf.hS * a.tS;
    const double HS = cf.hC * a.sC + cf.hE * a.sE + cf.hW * a.sW + cf.hN * a.sN + cf.hS * a.sS;
    QT = (a.tC - HT) + fabT * cf.cZ;
    QS = (a.sC - HS) + fabS * cf.cZ;
    PT = cf.lc * a.tC + cf.lE * a.tE + cf.lW * a.tW + cf.lN * a.tN + cf.lS * a.tS;
    PS = cf.lc * a.sC + cf.lE * a.sE + cf.lW * a.sW + cf.lN * a.sN + cf.lS * a.sS;
    a = b; b = nb; w = nw;
    uc = cf.nuc; uE = cf.nuE; uW = cf.nuW; uN = cf.nuN; uS = cf.nuS; cZp = cf.cZ;
  }
  {   // top level: surface boundary condition ts(1:2,:,:,maxk+1)   (:2550-2552)
    const double tn = QT - v.tsflux[(long)c2 * MS + m] * cZp, sn = QS - v.tsflux[((long)(I * J) + c2) * MS + m] * cZp;
    tt[K] = tn; ss[K] = sn;
    rl[K] = q.ec1 * tn + q.ec2 * sn + q.ec3 * (tn * tn) + q.ec4 * (tn * tn * tn);
  }
  unsigned in, topb, botb;
  double rdzt[K];
  co_decide_core<I, J, K, L, MS, true>(v, g, c2, m, k1c, tt, ss, rl, in, topb, botb, rdzt);
  v.comask[(long)c2 * MS + m] = in | (topb << 16);
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
template <int I, int J, int K, int L, int MS, bool ALL>
CG_HD void co_decide_core(const Dev &v, const GridC &g, const int c2, const unsigned m, const int k1c, double *tt, double *ss,
                          double *rl, unsigned &in, unsigned &topb, unsigned &botb, double *rdzt, const int st,
                          double *dzm_ext);

// scratch (device): four arrays of (K + 2) x st doubles, [level][thread], this thread's column at offset tid; nullptr: local arrays
template <int I, int J, int K, int L, int MS>
CG_HD void co_decide(const Dev &v, const GridC &g, const int c2, const unsigned m, unsigned &in, unsigned &topb, unsigned &botb,
                     double *rdzt, double *scratch = nullptr, const int st_ = 1) {
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC, rK = (long)I * J * MS;
  const int k1c = (int)v.k1[(c2 % I + 1) + (I + 2) * (c2 / I + 1)];
  const double *__restrict__ ts = v.ts_new + ((long)c2 * sC + m);         // level-1 cell of this column
  const double *__restrict__ rho = v.rho + ((long)c2 * MS + m);
  double tt_loc[K + 2], ss_loc[K + 2], rl_loc[K + 2];
  const int st = scratch ? st_ : 1;
  double *const tt = scratch ? scratch : tt_loc, *const ss = scratch ? scratch + (K + 2) * st : ss_loc,
               *const rl = scratch ? scratch + 2 * (K + 2) * st : rl_loc;
  double *const dzm = scratch ? scratch + 3 * (K + 2) * st : nullptr;
#pragma unroll
  for (int q = 1; q <= K; q++) {
    double t_ = 0.0, s_ = 0.0, r_ = 0.0;
    if (q >= k1c) {
      t_ = ts[(long)(q - 1) * sK];
      s_ = ts[(long)(q - 1) * sK + sL];
      r_ = rho[(long)(q - 1) * rK];
    }
    tt[q * st] = t_; ss[q * st] = s_; rl[q * st] = r_;
  }
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
  co_decide_core<I, J, K, L, MS, false>(v, g, c2, m, k1c, tt, ss, rl, in, topb, botb, rdzt, st, dzm);
}

// The decisions proper, on the column's new T, S, rho held in tt, ss, rl[1..K] (thread-private).  ALL = false: T, S, rho
// of the mixed levels are rewritten in place (they are in memory already); ALL = true: every wet level is written.
template <int I, int J, int K, int L, int MS, bool ALL>
CG_HD void co_decide_core(const Dev &v, const GridC &g, const int c2, const unsigned m, const int k1c, double *tt_, double *ss_,
                          double *rl_, unsigned &in, unsigned &topb, unsigned &botb, double *rdzt, const int st, double *dzm_ext) {
  // Element k of the column arrays lives at base[k * st]: st = 1 for thread-private arrays (host harness, ts_pre_column);
  // k_co_col passes shared-memory arrays laid out [level][thread] (st = block size).  The decision loop indexes them with
  // per-lane values (cur, bl, lo ...): in local memory such an access touches up to 32 cache lines per warp once the members
  // of a warp take different paths, in shared memory it is one conflict-free LDS (bank = thread whatever the level).
#define tt(k) tt_[(k) * st]
#define ss(k) ss_[(k) * st]
#define rl(k) rl_[(k) * st]
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC, rK = (long)I * J * MS;
  in = 0; topb = 0; botb = 0;
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  double *__restrict__ ts = v.ts_new + ((long)c2 * sC + m);         // level-1 cell of this column
  double *__restrict__ rho = v.rho + ((long)c2 * MS + m);
  // The reference keeps a compacted index array k(0:maxk) of the levels that still exist as separate boxes and shifts
  // it down after every merge; here the same set is a bit mask (bit l-1 = level l is a separate box), the neighbours of
  // a level come from clz / ffs, and a merge clears bits -- the sequence of comparisons and merges is the reference's.
  int head[K + 2];
  double dzm_loc[K + 2];
  double *const dzm_ = dzm_ext ? dzm_ext : dzm_loc;
  const int sd = dzm_ext ? st : 1;
#define dzm(k) dzm_[(k) * sd]
#pragma unroll
  for (int q = 1; q <= K; q++) { head[q] = q; dzm(q) = g.dz[q]; }
  rl(0) = 0.0;
  unsigned act = (K >= 32 ? 0xffffffffu : ((1u << K) - 1u)) & ~((1u << (k1c - 1)) - 1u);   // levels k1c..K
#define CG_BELOW(x) ({ const unsigned mb_ = act & ((1u << ((x) - 1)) - 1u); mb_ ? 32 - CG_CLZ(mb_) : 0; })
#define CG_ABOVE(x) ({ const unsigned ma_ = act >> (x); (x) + CG_FFS(ma_); })
  int cur = K, lastmix = 0;
  bool any = false;
  for (;;) {
    const int bl = CG_BELOW(cur);
    if (!(bl > 0 || (lastmix != 0 && cur != K))) break;
    if (bl == 0 || rl(cur) < rl(bl)) {
      if (lastmix == 0 || cur == K) cur = bl; else cur = CG_ABOVE(cur);
      lastmix = 0;
    } else {
      lastmix = 1;
      any = true;
      // extend the unstable run downward (goldstein.f90:2722-2731), then mix it into the top box in the order m-1 .. n
      int lo = bl;
      for (;;) {
        const int b2 = CG_BELOW(lo);
        if (!(b2 > 0 && rl(lo) >= rl(b2))) break;
        lo = b2;
      }
      const int h = cur;
      double sumT = tt(h) * dzm(h), sumS = ss(h) * dzm(h), dznew = dzm(h);
      for (int q = bl;; q = CG_BELOW(q)) {
        sumT = sumT + tt(q) * dzm(q);
        sumS = sumS + ss(q) * dzm(q);
        dznew = dznew + dzm(q);
        if (q == lo) break;
      }
      dzm(h) = dznew;
      tt(h) = sumT / dznew;
      ss(h) = sumS / dznew;
      rl(h) = ec1 * tt(h) + ec2 * ss(h) + ec3 * (tt(h) * tt(h)) + ec4 * (tt(h) * tt(h) * tt(h));
      act &= ~(((1u << bl) - 1u) & ~((1u << (lo - 1)) - 1u));   // levels lo..bl are now part of box h
    }
  }
  // SST / SSS as exported by step_goldstein (:428-431)
  if (v.sst) {
    v.sst[(long)c2 * MS + m] = tt(K);
    v.sst[((long)(I * J) + c2) * MS + m] = ss(K);
  }
  if (!any) {
    if (ALL) {
#pragma unroll
      for (int k = 1; k <= K; k++)
        if (k >= k1c) {
          ts[(long)(k - 1) * sK] = tt(k);
          ts[(long)(k - 1) * sK + sL] = ss(k);
          rho[(long)(k - 1) * rK] = rl(k);
        }
    }
    return;
  }
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
        ts[(long)(k - 1) * sK] = tt(hd);
        ts[(long)(k - 1) * sK + sL] = ss(hd);
        rho[(long)(k - 1) * rK] = rl(hd);
      } else if (ALL && k >= k1c) {
        ts[(long)(k - 1) * sK] = tt(k);
        ts[(long)(k - 1) * sK + sL] = ss(k);
        rho[(long)(k - 1) * rK] = rl(k);
      }
    }
  }
}
#undef tt
#undef ss
#undef rl
#undef dzm

// The decisions in LOCKSTEP form: the column's T, S, rho live in registers (static indices only) and the adjustment is a sequence of
// passes every lane executes alike -- (1) mark every boundary between two boxes whose upper box is not lighter than the lower one
// (the reference's test, goldstein.f90:2712 / 2724: mix unless rho(upper) < rho(lower)), (2) remove the marked boundaries, re-form
// the thickness-weighted means of the boxes that grew (summed top-down over the levels, as the passive tracers' means are) and their
// density, (3) repeat until a pass marks nothing.  A chain of unstable pairs is merged in one pass, as the reference merges a run;
// boxes that did not grow keep their values bit for bit (an unmixed level keeps the rho the flux kernel stored, so ties are decided
// on the reference's operands).  Box means equal those of co_decide_core's walk (one box at a time, down and back up, thread-private
// arrays indexed per lane) to rounding -- sums over levels against the means of means of the sequential merges -- and cost counts the
// same levels whenever the partition is the same, which is a matter of the ORDER of the merges (next paragraph).  ieos = 0, iconv = 0
// only (as every `col` kernel).
// WALK = true: the passes follow the reference's own ORDER of merges -- its walk (one box at a time from the top, after a merge first
// against the box below, then back up one box at a time) is replayed on the mask of unstable boundaries (integer work only: the
// comparisons of a pass are all taken beforehand, in static loops), and a pass merges exactly the run the walk would merge next.  With a
// nonlinear equation of state the final partition can depend on that order (two unstable runs close to each other in one column:
// 0.5 % of random columns with 2 K of noise per level, none in model states); WALK = false merges every unstable run of a pass at once.
template <int I, int J, int K, int L, int MS, bool WALK = false>
CG_HD void co_decide_static(const Dev &v, const GridC &g, const int c2, const unsigned m, unsigned &in, unsigned &topb, unsigned &botb,
                            double *rdzt) {
  static_assert(K <= 31, "level masks are 32 bits");
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC, rK = (long)I * J * MS;
  const int k1c = (int)v.k1[(c2 % I + 1) + (I + 2) * (c2 / I + 1)];
  double *__restrict__ ts = v.ts_new + ((long)c2 * sC + m);
  double *__restrict__ rho = v.rho + ((long)c2 * MS + m);
  const double ec1 = v.p.ec1[m], ec2 = v.p.ec2[m], ec3 = v.p.ec3[m], ec4 = v.p.ec4[m];
  const unsigned wet = ((1u << K) - 1u) & ~((1u << (k1c - 1)) - 1u);          // bit q = level q + 1 is wet
  double T[K], S[K], R[K];
#pragma unroll
  for (int q = 0; q < K; q++) {
    T[q] = 0.0; S[q] = 0.0; R[q] = 0.0;
    if ((wet >> q) & 1u) { T[q] = ts[(long)q * sK]; S[q] = ts[(long)q * sK + sL]; R[q] = rho[(long)q * rK]; }
  }
  unsigned sep = wet & ~(1u << (k1c - 1));      // bit q: a box boundary lies between level q + 1 and level q (q >= k1c)
  int cur = K, lastmix = 0;                     // WALK: the reference's position (top level of the box under test) and flag
  for (;;) {
    unsigned merge = 0u;
#pragma unroll
    for (int q = K - 1; q >= 1; q--)
      if (((sep >> q) & 1u) && !(R[q] < R[q - 1])) merge |= 1u << q;
    if (WALK) {
      // every unstable boundary of the column is in `merge`; pick the run the reference's walk merges next (goldstein.f90:2700-2746).
      // Box tops: level K and every level above a boundary; the boundary between a box and the box whose top is level b is bit b.
      const unsigned unst = merge;
      const unsigned tops = ((sep >> 1) | (1u << (K - 1))) & wet;
      merge = 0u;
      for (;;) {
        const unsigned mb = tops & ((1u << (cur - 1)) - 1u);
        const int bl = mb ? 32 - CG_CLZ(mb) : 0;                              // top level of the box below, 0: none
        if (!(bl > 0 || (lastmix != 0 && cur != K))) break;
        if (bl == 0 || !((unst >> bl) & 1u)) {                                // stable (or nothing below)
          if (lastmix == 0 || cur == K) cur = bl;
          else cur = cur + CG_FFS(tops >> cur);                               // back up one box
          lastmix = 0;
        } else {                                                              // unstable: cur's box takes the run below it
          lastmix = 1;
          int lo = bl;
          for (;;) {
            const unsigned m2 = tops & ((1u << (lo - 1)) - 1u);
            const int b2 = m2 ? 32 - CG_CLZ(m2) : 0;
            if (!(b2 > 0 && ((unst >> b2) & 1u))) break;
            lo = b2;
          }
          merge = sep & ((2u << bl) - 1u) & ~((1u << lo) - 1u);               // boundaries bit lo .. bit bl
          break;
        }
      }
    }
    if (merge == 0u) break;
    sep &= ~merge;
    double sT = 0.0, sS = 0.0, sD = 0.0;
    bool grew = false;
#pragma unroll
    for (int q = K - 1; q >= 0; q--) {
      if ((wet >> q) & 1u) {
        const bool top = (q == K - 1) || ((sep >> (q + 1)) & 1u);
        const bool bot = ((sep >> q) & 1u) || (q + 1 == k1c);
        const double dz = g.dz[q + 1];
        if (top) { sT = T[q] * dz; sS = S[q] * dz; sD = dz; grew = false; }
        else { sT = sT + T[q] * dz; sS = sS + S[q] * dz; sD = sD + dz; grew = grew || ((merge >> (q + 1)) & 1u); }
        if (bot && grew) {
          const double tm = sT / sD, sm = sS / sD;
          T[q] = tm; S[q] = sm;
          R[q] = ec1 * tm + ec2 * sm + ec3 * (tm * tm) + ec4 * (tm * tm * tm);
        }
      }
    }
#pragma unroll
    for (int q = 1; q < K; q++)     // a box's values are those of its bottom level: hand them up
      if (((wet >> q) & 1u) && q + 1 > k1c && !((sep >> q) & 1u)) { T[q] = T[q - 1]; S[q] = S[q - 1]; R[q] = R[q - 1]; }
  }
  if (v.sst) {   // SST / SSS as exported by step_goldstein (:428-431)
    v.sst[(long)c2 * MS + m] = T[K - 1];
    v.sst[((long)(I * J) + c2) * MS + m] = S[K - 1];
  }
  const unsigned bot_all = sep | (1u << (k1c - 1));
  const unsigned top_all = ((bot_all >> 1) | (1u << (K - 1))) & wet;
  in = wet & ~(top_all & bot_all);              // levels of boxes of more than one level
  topb = top_all & in;
  botb = bot_all & in;
  if (in == 0u) return;
  v.cost[(long)c2 * MS + m] += (double)CG_POPC(wet & ~top_all);   // levels that are no longer a box of their own (:2749-2764)
  double dzt = 0.0;
#pragma unroll
  for (int q = K - 1; q >= 0; q--) {
    rdzt[q] = 0.0;
    if ((in >> q) & 1u) {
      dzt = ((topb >> q) & 1u) ? g.dz[q + 1] : dzt + g.dz[q + 1];
      if ((botb >> q) & 1u) rdzt[q] = 1.0 / dzt;
      ts[(long)q * sK] = T[q];
      ts[(long)q * sK + sL] = S[q];
      rho[(long)q * rK] = R[q];
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

// Part 2, region by region: ALL passive tracers of one mixed region at a time.  co_passive_pair walks the column once per
// tracer pair, and each walk's loads wait behind the previous walk's stores (same array): seven dependent trips to memory
// per thread (ncu: 4.5 of 10 stall cycles per instruction "long scoreboard").  Here the loads of a region -- (L - 2) tracers x
// its levels, all independent -- are issued together, then the means are stored: one trip per region, and a column rarely
// has more than one or two.  Same sums in the same order (top-down, thickness weighted) as co_passive_pair.
template <int I, int J, int K, int L, int MS>
CG_HD void co_passive_regions(const Dev &v, const GridC &g, const int c2, const unsigned m, const unsigned topb, const unsigned botb,
                              const double *rdzt) {
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC;
  constexpr int NP = L - 2;
  double *__restrict__ ts = v.ts_new + ((long)c2 * sC + m) + 2 * sL;
  unsigned rem = topb;
  while (rem) {
    const int kt = 31 - CG_CLZ(rem);                                  // topmost region not yet done (0-based level)
    rem &= ~(1u << kt);
    const int kb = 31 - CG_CLZ(botb & ((2u << kt) - 1u));              // its bottom: the nearest bottom mark below the top
    double acc[NP];
#pragma unroll
    for (int l = 0; l < NP; l++) acc[l] = 0.0;
#pragma unroll 4
    for (int k = kt; k >= kb; k--) {
      const double dz = g.dz[k + 1];
      const double *__restrict__ p = ts + (long)k * sK;
#pragma unroll
      for (int l = 0; l < NP; l++) acc[l] += p[l * sL] * dz;
    }
    const double r = rdzt[kb];
#pragma unroll
    for (int l = 0; l < NP; l++) acc[l] = acc[l] * r;
    for (int k = kb; k <= kt; k++) {
      double *__restrict__ p = ts + (long)k * sK;
#pragma unroll
      for (int l = 0; l < NP; l++) p[l * sL] = acc[l];
    }
  }
}

// Part 2 as its own kernel body: ONE passive tracer of one (member, column); the region map comes from comask (bit k-1 = level
// k lies in a mixed region, bit 16+k-1 = it is the region's top; a region's bottom is the level whose lower neighbour is outside
// any region or the top of the next one).  All loads of the thread are independent (one trip to memory), the warps of a block
// are the tracers of 32 members of one column, and a warp none of whose lanes has a region returns at once.  Same sums in the
// same order as co_passive_pair; the region thickness is accumulated top-down exactly as co_decide_core does.
template <int I, int J, int K, int L, int MS>
CG_HD void co_passive_one(const Dev &v, const GridC &g, const int c2, const unsigned m, const int l) {
  static_assert(K <= 16, "region map is 16 + 16 bits");
  constexpr long sL = MS, sC = (long)L * MS, sK = (long)I * J * sC;
  const unsigned cm = v.comask[(long)c2 * MS + m];
  const unsigned in = cm & 0xffffu, topb = cm >> 16;
  if (in == 0u) return;
  double *__restrict__ ts = v.ts_new + ((long)c2 * sC + m) + l * sL;
  double a[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    a[k] = 0.0;
    if ((in >> k) & 1u) a[k] = ts[(long)k * sK];
  }
  double acc = 0.0, dzt = 0.0;
#pragma unroll
  for (int k = K - 1; k >= 0; k--) {
    if ((in >> k) & 1u) {
      const double dz = g.dz[k + 1];
      const bool istop = (topb >> k) & 1u;
      const bool isbot = (k == 0) || !((in >> (k - 1)) & 1u) || ((topb >> (k - 1)) & 1u);
      if (istop) { acc = 0.0; dzt = dz; } else dzt = dzt + dz;
      acc += a[k] * dz;
      if (isbot) a[k] = acc * (1.0 / dzt);
    }
  }
  double cur = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    if ((in >> k) & 1u) {
      const bool isbot = (k == 0) || !((in >> (k - 1)) & 1u) || ((topb >> (k - 1)) & 1u);
      if (isbot) cur = a[k];
      ts[(long)k * sK] = cur;
    }
  }
}

// both parts by one thread (production convection kernel k_co_col; host test harness)
template <int I, int J, int K, int L, int MS, bool DEC_ONLY = false, int PMODE = -1, int LOCKSTEP = 0>   // PMODE: -1 = v.co_pairwise decides, 0 = regions, 1 = pairs; LOCKSTEP: 1 = every unstable run of a pass, 2 = the reference's order
CG_HD void co_column(const Dev &v, const GridC &g, const int c2, const unsigned m, double *scratch = nullptr, const int st = 1) {
  // the flux kernel found every level of this (member, column) stable: nothing to adjust, SST / SSS are exported already
  if (v.co_skip_stable && v.comask && v.comask[(long)c2 * MS + m] == 0u) return;
  unsigned in, topb, botb;
  double rdzt[K];
  if constexpr (LOCKSTEP != 0) co_decide_static<I, J, K, L, MS, LOCKSTEP == 2>(v, g, c2, m, in, topb, botb, rdzt);
  else co_decide<I, J, K, L, MS>(v, g, c2, m, in, topb, botb, rdzt, scratch, st);
  if (DEC_ONLY || v.co_pairwise == 2) {   // decisions only: the region map goes to comask for k_co_passive (one thread per tracer)
    v.comask[(long)c2 * MS + m] = in | (topb << 16);
    return;
  }
  if (in == 0) return;
  if (PMODE == 1 || (PMODE < 0 && v.co_pairwise)) {
    for (int l = 2; l < L; l += 2) co_passive_pair<I, J, K, L, MS>(v, g, c2, m, in, topb, botb, rdzt, l);
  } else if (L > 2) {
    co_passive_regions<I, J, K, L, MS>(v, g, c2, m, topb, botb, rdzt);
  }
}

}  // namespace cg
