// k_biogem.cu -- BIOGEM kernels on the tracer hot path, sm_100a, compiled -fmad=false.
//
// biogem_tracercoupling (src/biogem/biogem.f90:1885-2077): salinity-normalised transfer between the
// GOLDSTEIN tracer array ts and BIOGEM's ocn, with per-member global reductions (mean salinity old /
// new, old and new inventory of every biogeochemical tracer).  The reference sums each wet column
// over k (ascending) and then the column partials in its vocn order (i outer, j inner); the kernels
// keep exactly that order -- thread (member, column) for the partials, thread (member, quantity) for
// the ordered sum over columns -- so every total is bit-identical to the sequential code.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "cg_device.cuh"
#include "cg_host.hpp"

namespace cg {
// Compiled member strides of the fixed-shape (36 x 36 x 16) kernel instances: 128 (one tile), 256, 512 (larger shards of one handle)
static inline bool fix_ms(int MS) { return MS == 128 || MS == 256 || MS == 512; }
static inline bool tc_fix_shape(const Dev &v) { return v.I == 36 && v.J == 36 && v.K == 16 && v.L == 16 && fix_ms(v.MS) && !getenv("CG_BG_NOFIX"); }
// f(integral_constant<int, FMS>): FMS = the handle's member stride where an instance with that stride exists, else 0 (generic kernel)
template <class F>
static inline void fix_dispatch(const bool ok, const int MS, F f) {
  if (ok && MS == 128) f(std::integral_constant<int, 128>());
  else if (ok && MS == 256) f(std::integral_constant<int, 256>());
  else if (ok && MS == 512) f(std::integral_constant<int, 512>());
  else f(std::integral_constant<int, 0>());
}

constexpr double kBgZeroC = 273.15;          // gem_cmn.f90:690
constexpr double kBgNullSmall = 0.999999e-19; // gem_cmn.f90:719

// quantity slots of the reduction scratch
//   0            : sum_k ocn(S)*V * rtot_V              (old mean salinity)        [phase A]
//   1            : sum_k (ts(S)+saln0+docn(S))*V*rtot_V (new mean salinity)        [phase A]
//   2 .. L-1     : sum_k ocn(l)*M, l = 3..L             (old inventories)          [phase A]
//   L .. 2L-3    : sum_k loc_vocn(l)*M, l = 3..L        (salinity-adjusted new)    [phase B]
// FIX: grid shape, tracer count and member stride of the bench configuration as compile-time constants (see k_bg_step)
template <int FMS>
__global__ void __launch_bounds__(128) k_tc_partial(const Dev v, const int phase) {
  constexpr bool FIX = FMS > 0;
  const int I = FIX ? 36 : v.I, J = FIX ? 36 : v.J, K = FIX ? 16 : v.K, L = FIX ? 16 : v.L, MS = FIX ? FMS : v.MS;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || n >= v.nwet) return;
  const int c2 = v.bgcols[n];
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  const size_t sC = (size_t)L * MS, sK = (size_t)I * J * sC;
  const size_t o0 = cell3(I, J, i, j, 1) * sC + m;
  const size_t p0 = cell3(I, J, i, j, 1) * MS + m, pK = (size_t)I * J * MS;
  const size_t v0 = cell3(I, J, i, j, 1), vK = (size_t)I * J;
  const size_t nq = (size_t)v.nwet * MS;
  double *part = v.bg_part + (size_t)n * MS + m;
  // phase 3 / 4 = phase 2 + phase 1 regrouped by what they read (cg_run's pipelined block): 3 = everything that depends on
  // BIOGEM's own state only (old mean salinity, old inventories: ocn, V, M as the PREVIOUS coupling left them), taken
  // one block ahead; 4 = what needs this cycle's ts (new mean salinity, salinity-adjusted new inventories -- the latter
  // need the old mean salinity only, which phase 3 + its ordered sum have produced by then).  Same expressions per slot.
  const bool do_old = phase == 0 || phase == 2 || phase == 3, do_newS = phase == 0 || phase == 2 || phase == 4;
  if (do_old || do_newS) {
    // phase 2 = phase 0 taken BEFORE step_biogem (fused coupling): the salinity anomaly of the frozen configuration is
    // exactly +0.0 (no BIOGEM source or sink of salt), so the sum does not need vdocn
    const double saln0 = v.p.saln0[m];
    double a = 0.0, b = 0.0;
    for (int k = k1c; k <= K; k++) {
      const double V = v.bg_V[v0 + (size_t)(k - 1) * vK];
      if (do_old) a = a + __ldcs(v.bg_ocn + o0 + (size_t)(k - 1) * sK + MS) * V;
      const double dS = (phase == 0) ? v.bg_vdocn[o0 + (size_t)(k - 1) * sK + MS] : 0.0;
      if (do_newS) b = b + (__ldcs(v.ts_cur + o0 + (size_t)(k - 1) * sK + MS) + saln0 + dS) * V;
    }
    if (do_old) part[0] = a * v.bg_rtot_V;
    if (do_newS) part[nq] = b * v.bg_rtot_V;
  }
  if (do_old) {
    // old inventories: level outer / tracer inner, so that the loads of a level are all in flight at once and the
    // per-tracer sums (each still accumulated over k ascending, as the reference does) are independent chains
    if (L <= kBgMaxL) {
      double s[kBgMaxL];
#pragma unroll
      for (int l = 2; l < kBgMaxL; l++) s[l] = 0.0;
      for (int k = k1c; k <= K; k++) {
        const double Mk = v.bg_M[p0 + (size_t)(k - 1) * pK];
        const double *__restrict__ oc = v.bg_ocn + o0 + (size_t)(k - 1) * sK;
#pragma unroll
        for (int l = 2; l < kBgMaxL; l++)
          if (l < L) s[l] = s[l] + __ldcs(oc + (size_t)l * MS) * Mk;   // read once: streaming
      }
#pragma unroll
      for (int l = 2; l < kBgMaxL; l++)
        if (l < L) part[(size_t)l * nq] = s[l];
    } else {
      for (int l = 2; l < L; l++) {
        double s = 0.0;
        for (int k = k1c; k <= K; k++) s = s + v.bg_ocn[o0 + (size_t)(k - 1) * sK + (size_t)l * MS] * v.bg_M[p0 + (size_t)(k - 1) * pK];
        part[(size_t)l * nq] = s;
      }
    }
  }
  if (phase == 1 || phase == 4) {
    const double rmean = 1.0 / v.bg_tot[m];  // loc_ocn_rmean_S_OLD
    if (L <= kBgMaxL) {
      double s[kBgMaxL];
#pragma unroll
      for (int l = 2; l < kBgMaxL; l++) s[l] = 0.0;
      for (int k = k1c; k <= K; k++) {
        const size_t o = o0 + (size_t)(k - 1) * sK;
        const double Mk = v.bg_M[p0 + (size_t)(k - 1) * pK], Sk = v.bg_ocn[o + MS];
        const double *__restrict__ tc = v.ts_cur + o;
#pragma unroll
        for (int l = 2; l < kBgMaxL; l++)
          if (l < L) s[l] = s[l] + (__ldcs(tc + (size_t)l * MS) * Sk * rmean) * Mk;
      }
#pragma unroll
      for (int l = 2; l < kBgMaxL; l++)
        if (l < L) part[(size_t)(L - 2 + l) * nq] = s[l];
    } else {
      for (int l = 2; l < L; l++) {
        double s = 0.0;
        for (int k = k1c; k <= K; k++) {
          const size_t o = o0 + (size_t)(k - 1) * sK;
          s = s + (v.ts_cur[o + (size_t)l * MS] * v.bg_ocn[o + MS] * rmean) * v.bg_M[p0 + (size_t)(k - 1) * pK];
        }
        part[(size_t)(L - 2 + l) * nq] = s;
      }
    }
  }
}

// Sum of n terms per member IN ORDER (term t of lane's member at p[t*stride]), done by one warp: the terms are staged
// through shared memory in tiles of kSumTile x 32 by all warps of the block (coalesced, many loads in flight), then
// warp 0 adds its lane's column of the tile sequentially.  Bit-identical to the reference's scalar loop, but the memory
// latency is paid once per tile instead of once per term.
constexpr int kSumTile = 128, kSumWarps = 8;   // blocks using ordered_sum_block have 32*kSumWarps threads
// The same buffer as two half tiles: while warp 0 adds the terms of one half tile (a dependent chain: ~40 cycles per term on
// B200's fp64 pipe, nothing else to do), warps 1 .. kSumWarps-1 fetch the next half tile, so that the global-memory latency of
// a tile is paid under the additions of the previous one instead of in front of them (k_bg_atchem2: 1296 terms, 94 -> ~35 us).
__device__ double ordered_sum_block(const double *__restrict__ p, const size_t stride, const int n, double *tile /* [kSumTile][32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kHalf = kSumTile / 2, kLoaders = kSumWarps - 1, kPer = (kHalf + kLoaders - 1) / kLoaders;
  const int ntile = (n + kHalf - 1) / kHalf;
  double s = 0.0;
  auto fetch = [&](const int t, const int w, const int nw) {   // terms w, w + nw, ... of half tile t -> buffer t & 1
    const int t0 = t * kHalf, nt = min(kHalf, n - t0);
    double *dst = tile + (t & 1) * kHalf * 32;
    double r[kPer + 1];
#pragma unroll
    for (int u = 0; u < kPer + 1; u++) {   // all loads in flight before the first store
      const int q = w + u * nw;
      r[u] = (q < nt) ? p[(size_t)(t0 + q) * stride + lane] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kPer + 1; u++) {
      const int q = w + u * nw;
      if (q < nt) dst[q * 32 + lane] = r[u];
    }
  };
  if (ntile > 0) fetch(0, warp, kSumWarps);                    // first half tile: all warps (kHalf / kSumWarps <= kPer + 1 terms each)
  __syncthreads();
  for (int t = 0; t < ntile; t++) {
    if (warp == 0) {
      const int nt = min(kHalf, n - t * kHalf);
      const double *src = tile + (t & 1) * kHalf * 32 + lane;
#pragma unroll 16
      for (int q = 0; q < nt; q++) s = s + src[q * 32];
    } else if (t + 1 < ntile) {
      fetch(t + 1, warp - 1, kLoaders);
    }
    __syncthreads();
  }
  return s;  // valid in warp 0
}

// ordered sum over the wet columns: block = (32-member tile, quantity)
// skip >= 0: blockIdx.y = 0 sums quantity `skip`, blockIdx.y >= 1 quantities q0 .. q1-1 (the "new" set: slot 1 and L .. 2L-3);
// skip = -2: quantities q0 .. q1-1 without slot 1 (the "old" set)
__global__ void __launch_bounds__(32 * kSumWarps) k_tc_sum(const Dev v, const int q0, const int q1, const int skip = -1) {
  __shared__ double tile[kSumTile * 32];
  const int MS = v.MS;
  const int m0 = blockIdx.x * 32;
  const int q = skip >= 0 ? (blockIdx.y == 0 ? skip : q0 + (int)blockIdx.y - 1) : q0 + (int)blockIdx.y;
  if (q >= q1 || (skip == -2 && q == 1)) return;
  const double s = ordered_sum_block(v.bg_part + (size_t)q * v.nwet * MS + m0, (size_t)MS, v.nwet, tile);
  if (threadIdx.x < 32) v.bg_tot[(size_t)q * MS + m0 + threadIdx.x] = s;
}

// (2)+(3) of biogem_tracercoupling: new T,S, rescaled tracers, cell masses, ts <- normalised ocn.  Cells are independent
// here, so the grid runs over (member tile, cell) with one warp per cell; the per-member totals are staged once per
// block in shared memory.
// per-member factors of steps (2)+(3), computed once: slots 2L .. 2L+3 = 1/mean_S_OLD, Sratio, 1/Sratio, mean_S_NEW,
// slots 2L+4+l = tot_OLD(l) / tot_NEW(l)
__global__ void k_tc_factors(const Dev v) {
  const int L = v.L, MS = v.MS;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= MS) return;
  double *fac = v.bg_tot + (size_t)2 * L * MS + m;
  const double mean_S_OLD = v.bg_tot[m], mean_S_NEW = v.bg_tot[MS + m];
  const double Sratio = mean_S_NEW / mean_S_OLD;
  fac[0] = 1.0 / mean_S_OLD;
  fac[(size_t)MS] = Sratio;
  fac[(size_t)2 * MS] = 1.0 / Sratio;
  fac[(size_t)3 * MS] = mean_S_NEW;
  for (int l = 2; l < L; l++) {
    const double told = v.bg_tot[(size_t)l * MS + m], tnew = v.bg_tot[(size_t)(L - 2 + l) * MS + m];
    const double rtnew = (fabs(tnew) < kBgNullSmall) ? 0.0 : 1.0 / tnew;
    fac[(size_t)(4 + l) * MS] = told * rtnew;
  }
}
constexpr int kApplyCellsPerWarp = 1, kApplyWarps = 4;
template <int FMS>
__global__ void __launch_bounds__(32 * kApplyWarps, 4) k_tc_apply(const Dev v) {
  __shared__ double s_f[kBgMaxL][32], s_rmean[32], s_sr[32], s_rsr[32], s_mnew[32];
  constexpr bool FIX = FMS > 0;
  const int I = FIX ? 36 : v.I, J = FIX ? 36 : v.J, K = FIX ? 16 : v.K, L = FIX ? 16 : v.L, MS = FIX ? FMS : v.MS;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int m = blockIdx.x * 32 + lane;
  {
    const double *fac = v.bg_tot + (size_t)2 * L * MS + m;
    if (warp == 0) {
      s_rmean[lane] = fac[0];
      s_sr[lane] = fac[(size_t)MS];
      s_rsr[lane] = fac[(size_t)2 * MS];
      s_mnew[lane] = fac[(size_t)3 * MS];
    }
    for (int l = 2 + warp; l < L; l += kApplyWarps) s_f[l][lane] = fac[(size_t)(4 + l) * MS];
  }
  __syncthreads();
  const double rmean_S_OLD = s_rmean[lane], Sratio = s_sr[lane], rSratio = s_rsr[lane], mean_S_NEW = s_mnew[lane];
  const double saln0 = v.p.saln0[m];
  const int ncell = I * J * K;
  const int c0 = (blockIdx.y * kApplyWarps + warp) * kApplyCellsPerWarp;
  for (int c = c0; c < min(c0 + kApplyCellsPerWarp, ncell); c++) {
    const int k = c / (I * J) + 1, r = c % (I * J), j = r / I + 1, i = r % I + 1;
    if (k < CG_K1(v, i, j)) continue;
    const size_t o = (size_t)c * L * MS + m;
    double *__restrict__ ocn = v.bg_ocn + o;
    double *__restrict__ ts = v.ts_cur + o;
    const double *__restrict__ dv = v.bg_vdocn + o;
    if (L <= kBgMaxL) {
      // every load of the cell first (ts and ocn are rewritten in place: the compiler cannot hoist a load of ts above a
      // store to ts by itself), then the arithmetic in the reference's order, then the stores
      double tv[kBgMaxL], dd[kBgMaxL], bpv[kBgMaxLS];
      const double Sold = ocn[MS];
#pragma unroll
      for (int l = 0; l < kBgMaxL; l++)
        if (l < L) { tv[l] = ts[(size_t)l * MS]; dd[l] = dv[(size_t)l * MS]; }
      const double Mc = v.bg_M[(size_t)c * MS + m], rMc = v.bg_rM[(size_t)c * MS + m];
      double *__restrict__ bp = v.bg_biopart ? v.bg_biopart + (size_t)c * v.bg_LS * MS + m : nullptr;
      if (bp) {
#pragma unroll
        for (int ls = 0; ls < kBgMaxLS; ls++)
          if (ls < v.bg_LS) bpv[ls] = bp[(size_t)ls * MS];
      }
      const double Tn = tv[0] + kBgZeroC + dd[0];
      const double Sn = tv[1] + saln0 + dd[1];
      const double rn = mean_S_NEW / Sn;
      ocn[0] = Tn;
      ocn[MS] = Sn;
      ts[0] = Tn - kBgZeroC;
      ts[MS] = Sn - saln0;
#pragma unroll
      for (int l = 2; l < kBgMaxL; l++)
        if (l < L) {
          const double lv = tv[l] * Sold * rmean_S_OLD;
          double x = s_f[l][lane] * lv + dd[l];
          x = Sratio * x;
          ocn[(size_t)l * MS] = x;
          ts[(size_t)l * MS] = rn * x;
        }
      if (bp) {  // biogem.f90:2042-2043 (vdbio_part = 0: no particulate flux forcing)
#pragma unroll
        for (int ls = 0; ls < kBgMaxLS; ls++)
          if (ls < v.bg_LS) bp[(size_t)ls * MS] = Sratio * (bpv[ls] + 0.0);
      }
      v.bg_M[(size_t)c * MS + m] = rSratio * Mc;
      v.bg_rM[(size_t)c * MS + m] = Sratio * rMc;
      continue;
    }
    const double Sold = ocn[MS];
    const double Tn = ts[0] + kBgZeroC + dv[0];
    const double Sn = ts[MS] + saln0 + dv[MS];
    const double rn = mean_S_NEW / Sn;
    ocn[0] = Tn;
    ocn[MS] = Sn;
    ts[0] = Tn - kBgZeroC;
    ts[MS] = Sn - saln0;
    for (int l = 2; l < L; l++) {
      const double lv = ts[(size_t)l * MS] * Sold * rmean_S_OLD;
      double x = s_f[l][lane] * lv + dv[(size_t)l * MS];
      x = Sratio * x;
      ocn[(size_t)l * MS] = x;
      ts[(size_t)l * MS] = rn * x;
    }
    if (v.bg_biopart) {  // biogem.f90:2042-2043 (vdbio_part = 0: no particulate flux forcing)
      double *__restrict__ bp = v.bg_biopart + (size_t)c * v.bg_LS * MS + m;
      for (int ls = 0; ls < v.bg_LS; ls++) bp[(size_t)ls * MS] = Sratio * (bp[(size_t)ls * MS] + 0.0);
    }
    v.bg_M[(size_t)c * MS + m] = rSratio * v.bg_M[(size_t)c * MS + m];
    v.bg_rM[(size_t)c * MS + m] = Sratio * v.bg_rM[(size_t)c * MS + m];
  }
}

// biogem_climate's only state change on this path: reset the convection counter (biogem.f90:2238)
__global__ void k_bg_reset_cost(const Dev v) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < (size_t)v.I * v.J * v.MS) v.cost[q] = 0.0;
}


// =============================================================================================================
// step_biogem (biogem.f90:528-1877) for the frozen configuration: one thread = (member, wet column).  The thread
// runs the column's decay, closed-system sediment return, DOM and particulate remineralisation
// (sub_box_remin_DOM/part, biogem_box.f90:2287-2875), then the surface cell's carbonate chemistry
// (gem_carbchem.f90:59-780), gas exchange (biogem_box.f90:27-299), restoring forcing, biological uptake
// (sub_calc_bio_uptake :346-1514) and writes the tracer anomaly vdocn (:1811-1844).  No cross-column terms exist
// inside step_biogem, so the reference's three sweeps over columns collapse into one pass per column; every
// accumulation keeps the reference's order (see oracle/cgo_biogem.c for the plain restatement).
// =============================================================================================================
namespace bgk {
enum { CC_K1, CC_K2, CC_K, CC_KB, CC_KW, CC_KSI, CC_KHF, CC_KHSO4, CC_KP1, CC_KP2, CC_KP3, CC_KCAL, CC_KARG, CC_QCO2, N_CC };
struct Carb { double H, co2, co3, hco3, ohm_cal, RF0; };
constexpr double kZeroC = 273.15, kNull = -0.999999e19, kNS = 0.999999e-19, kYrS = 3600.0 * (24.0 * 365.25);
constexpr double kAtmMol = 1.7692e+020, kR = 83.145, kRSI = 8.3145, kVmol = 0.022414, kCp = 4.1855, kConcMg = 0.05282,
                 kConcMgtoCa = 5.155, kStd13C = 0.011202, kStd14C = 1.176e-12, kPaAtm = 1.0 / 1.01325e+05;

__device__ __forceinline__ double iso_delta(double tot, double iso, double standard) {  // gem_util.f90:568-598, allow_negative = F
  if (tot > kNS) {
    const double fr = iso / tot;
    if ((1.0 - fr) > kNS) {
      const double R = fr / (1.0 - fr);
      return 1000.0 * (R / standard - 1.0);
    }
    return kNull;
  }
  return kNull;
}
__device__ __forceinline__ double iso_fraction(double delta, double standard) {  // gem_util.f90:604-617
  const double R = standard * (1.0 + delta / 1000.0);
  return R / (1.0 + R);
}
__device__ __forceinline__ double calc_rho(double T, double S) {  // gem_util.f90:670-682
  const double TC = T - kZeroC;
  return 1000.0 + (0.7968 * S - 0.0559 * TC - 0.0063 * (TC * TC) + 3.7315E-05 * ((TC * TC) * TC));
}
__device__ __forceinline__ double corr_p(double TC, double P, double rRT, double d1, double d2, double d3, double d4, double d5) {
  return (-(d1 + d2 * TC + d3 * TC * TC) + (5.0E-4 * (d4 + d5 * TC)) * P) * P * rRT;  // gem_carbchem.f90:1175-1186
}
// a / b with y = 1.0 / b already known (correctly rounded): q = a y is within an ulp or two of the quotient, r = a - b q is exact in
// one FMA, q + r y rounds to RN(a / b) (Markstein 1990; holds for normal operands unless b's significand is all ones) -- three
// instructions for the ~28 of the IEEE division sequence, same result.  Used where one denominator serves several divisions.
__device__ __forceinline__ double div_by(const double a, const double b, const double y) {
  const double q = a * y;
  const double r = fma(-b, q, a);
  return fma(r, y, q);
}
__device__ __forceinline__ double fS(double c, double S) { double x = c * S / 35.0; if (x < kNS) x = kNS; return x; }

// sub_calc_carbconst (Mehrbach) + sub_adj_carbconst, gem_carbchem.f90:59-297; only the constants the surface solve reads
__device__ void carbconst(double D, double T_in, double S_in, double Ca, double Mg, double *cc) {
  double T = T_in, S = S_in;
  if (T < (kZeroC + 2.0)) T = kZeroC + 2.0;
  if (T > (kZeroC + 35.0)) T = kZeroC + 35.0;
  if (S < 26.0) S = 26.0;
  if (S > 43.0) S = 43.0;
  const double P = D / 10.0;
  // x**0.5, x**1.5, 10**y, LOG(10**y) of the reference (libm pow, <= 1 ulp) through sqrt / exp10 / a multiplication: each is
  // within 1 ulp of the exact value too -- the same distance CUDA's pow() keeps from glibc's -- at a fraction of pow()'s
  // instruction count, and without pow()'s special-case branches (members diverge in them)
  const double S_p05 = sqrt(S), S_p15 = S * S_p05, S_p20 = S * S;
  const double T_ln = log(T), T_log = log10(T), rT = 1.0 / T, Tr100 = T / 100.0, TC = T - kZeroC;
  const double rRT = 1.0 / (kR * T);
  const double Ii = (S > kNS) ? 19.924 * S / (1000.0 - 1.005 * S) : kNS;
  const double I_p05 = sqrt(Ii), I_p15 = Ii * I_p05, I_p20 = Ii * Ii;
  double Cl = S_in / 1.80655;
  if (Cl < kNS) Cl = kNS;
  const double ION = (Cl > kNS) ? 0.00147 + 0.03592 * Cl + 0.000068 * Cl * Cl : kNS;
  const double ION_p05 = sqrt(ION);
  const double m2c = log(1 - 0.001005 * S);
  const double SO4tot = fS(0.02824, S), Ftot = fS(0.00007, S);
  const double lnkHSO4 = 141.328 - 4276.1 * rT - 23.093 * T_ln + (324.57 - 13856.0 * rT - 47.986 * T_ln) * I_p05 +
                         (-771.54 + 35474.0 * rT + 114.723 * T_ln) * Ii - 2698.0 * rT * I_p15 + 1776.0 * rT * I_p20;
  const double lnkHF = div_by(1590.2, T, rT) - 12.641 + 1.525 * ION_p05;
  cc[CC_KHSO4] = exp(lnkHSO4 + m2c);
  const double f2t = log(1.0 + SO4tot / cc[CC_KHSO4]);
  cc[CC_KHF] = exp(lnkHF + m2c + f2t);
  const double f2s = log(1.0 + SO4tot / cc[CC_KHSO4] + Ftot / cc[CC_KHF]);
  const double t2s = -f2t + f2s;
  constexpr double kLn10 = 2.302585092994045684;   // LOG(10**y) = y ln 10
  cc[CC_K1] = exp(kLn10 * -(3670.7 * rT - 62.008 + 9.7944 * T_ln - 0.0118 * S + 0.000116 * S_p20) +
                  corr_p(TC, P, rRT, -2.550E+1, +1.271E-1, +0.000E+0, -3.080E+0, +8.770E-2));
  cc[CC_K2] = exp(kLn10 * -(1394.7 * rT + 4.777 - 0.0184 * S + 0.000118 * S_p20) +
                  corr_p(TC, P, rRT, -1.582E+1, -2.190E-2, +0.000E+0, +1.130E+0, -1.475E-1));
  cc[CC_K] = cc[CC_K1] / cc[CC_K2];
  cc[CC_KB] = exp((148.0248 + 137.194 * S_p05 + 1.62247 * S +
                   (-8966.90 - 2890.51 * S_p05 - 77.942 * S + 1.726 * S_p15 - 0.0993 * S_p20) * rT +
                   (-24.4344 - 25.085 * S_p05 - 0.2474 * S) * T_ln + 0.053105 * S_p05 * T) +
                  m2c + t2s + corr_p(TC, P, rRT, -2.948E+1, +1.622E-1, +2.608E-3, -2.840E+0, +0.000E+0));
  cc[CC_KW] = exp((148.9802 - 13847.26 * rT - 23.6521 * T_ln + (-5.977 + 118.67 * rT + 1.0495 * T_ln) * S_p05 - 0.01615 * S) +
                  corr_p(TC, P, rRT, -2.002E+1, +1.119E-1, -1.409E-3, -5.130E+0, +7.940E-2));
  cc[CC_KSI] = exp((117.40 - 8904.2 * rT - 19.334 * T_ln + (3.5913 - 458.79 * rT) * I_p05 + (-1.5998 + 188.74 * rT) * Ii +
                    (0.07871 - 12.1652 * rT) * Ii * Ii) +
                   m2c + corr_p(TC, P, rRT, -2.948E+1, +1.622E-1, +2.608E-3, -2.840E+0, +0.000E+0));
  cc[CC_KHF] = exp(lnkHF + m2c + f2s + corr_p(TC, P, rRT, -9.780E+0, -9.000E-3, -9.420E-4, -3.910E+0, +5.400E-2));
  cc[CC_KHSO4] = exp(lnkHSO4 + m2c + f2s + corr_p(TC, P, rRT, -1.803E+1, +4.660E-2, +3.160E-4, -4.530E+0, +9.000E-2));
  cc[CC_KP1] = exp((115.54 - div_by(4576.752, T, rT) - 18.453 * T_ln + (0.69171 - div_by(106.736, T, rT)) * S_p05 + (-0.01844 - div_by(0.65643, T, rT)) * S) +
                   corr_p(TC, P, rRT, -1.451E+1, +1.211E-1, -3.210E-4, -2.670E+0, +4.270E-2));
  cc[CC_KP2] = exp((172.1033 - div_by(8814.715, T, rT) - 27.927 * T_ln + (1.3566 - div_by(160.340, T, rT)) * S_p05 + (-0.05778 + div_by(0.37335, T, rT)) * S) +
                   corr_p(TC, P, rRT, -2.312E+1, +1.758E-1, -2.647E-3, -5.150E+0, +9.000E-2));
  cc[CC_KP3] = exp((-18.126 - div_by(3070.75, T, rT) + (2.81197 + div_by(17.27039, T, rT)) * S_p05 + (-0.09984 - div_by(44.99486, T, rT)) * S) +
                   corr_p(TC, P, rRT, -2.657E+1, +2.020E-1, -3.042E-3, -4.080E+0, +7.140E-2));
  cc[CC_KCAL] = exp(corr_p(TC, P, rRT, -4.876E+1, +5.304E-1, +0.000E+0, -1.176E+1, +3.692E-1)) *
                exp10(-171.9065 - 0.077993 * T + 2839.319 * rT + 71.595 * T_log +
                      (-0.77712 + 0.0028426 * T + 178.34 * rT) * S_p05 - 0.07711 * S + 0.0041249 * S_p15);
  cc[CC_KARG] = 0.0;  // aragonite saturation is diagnostic only
  cc[CC_QCO2] = exp(-60.2409 + 93.4517 * (100 * rT) + 23.3585 * log(Tr100) +
                    S * (0.023517 - 0.023656 * (Tr100) + 0.0047036 * (Tr100 * Tr100)));
  // sub_adj_carbconst
  double ratio = 1.0;
  if (Ca > kNS) ratio = Mg / Ca;
  cc[CC_KCAL] = cc[CC_KCAL] - 3.655E-8 * (kConcMgtoCa - ratio);
  cc[CC_K1] = (1.0 + 0.155 * (Mg - kConcMg) / kConcMg) * cc[CC_K1];
  cc[CC_K2] = (1.0 + 0.422 * (Mg - kConcMg) / kConcMg) * cc[CC_K2];
}
// reciprocals of the constants of one solve that carb_iter divides by (taken once per solve)
struct CarbR { double rKB, rKP2, rKP3, rK12, rK23, rK123, K12, K23, K123, km4, rkm4, km4x2, rkm4x2; };
__device__ __forceinline__ void carb_recips(const double *cc, CarbR &r) {
  r.K12 = cc[CC_KP1] * cc[CC_KP2]; r.K23 = cc[CC_KP2] * cc[CC_KP3]; r.K123 = cc[CC_KP1] * cc[CC_KP2] * cc[CC_KP3];
  r.rKB = 1.0 / cc[CC_KB]; r.rKP2 = 1.0 / cc[CC_KP2]; r.rKP3 = 1.0 / cc[CC_KP3];
  r.rK12 = 1.0 / r.K12; r.rK23 = 1.0 / r.K23; r.rK123 = 1.0 / r.K123;
  r.km4 = cc[CC_K] - 4.0; r.rkm4 = 1.0 / r.km4; r.km4x2 = 2.0 * r.km4; r.rkm4x2 = 1.0 / r.km4x2;
}
// one pass of the implicit [H] loop (gem_carbchem.f90:352-424 / 578-640); H2S, NH4, SiO2 totals are zero here.  Every quotient
// is the reference's (same operands, correctly rounded); the 16 whose denominator is a constant of the solve or a power of H go
// through div_by with one reciprocal each instead of the full division sequence.  H3SiO4 = 0 / (1 + H / kSi) is +0.0 as written.
__device__ __forceinline__ void carb_iter(double DIC, double ALK, double PO4tot, double Btot, double SO4tot, double Ftot,
                                          const double *cc, const CarbR &rk, double H, double &co2, double &co3, double &hco3, double &H1,
                                          double &H2) {
  const double H_p2 = H * H, H_p3 = H * H_p2;
  const double rH = 1.0 / H, rH2 = 1.0 / H_p2, rH3 = 1.0 / H_p3;
  const double OH = div_by(cc[CC_KW], H, rH);
  const double H4BO4 = Btot / (1.0 + div_by(H, cc[CC_KB], rk.rKB));
  const double H3SiO4 = 0.0;
  const double HSO4 = SO4tot / (1.0 + div_by(cc[CC_KHSO4], H, rH));
  const double HF = Ftot / (1.0 + div_by(cc[CC_KHF], H, rH));
  const double H3PO4 = PO4tot / (1.0 + div_by(cc[CC_KP1], H, rH) + div_by(rk.K12, H_p2, rH2) + div_by(rk.K123, H_p3, rH3));
  const double HPO4 = PO4tot / (1.0 + div_by(H, cc[CC_KP2], rk.rKP2) + div_by(H_p2, rk.K12, rk.rK12) + div_by(cc[CC_KP3], H, rH));
  const double PO4 = PO4tot / (1.0 + div_by(H, cc[CC_KP3], rk.rKP3) + div_by(H_p2, rk.K23, rk.rK23) + div_by(H_p3, rk.K123, rk.rK123));
  const double ALK_DIC = ALK - H4BO4 - OH - HPO4 - 2.0 * PO4 - H3SiO4 - 0.0 - 0.0 + H + HSO4 + HF + H3PO4;
  const double k = cc[CC_K];
  const double a = 4.0 * ALK_DIC + DIC * k - ALK_DIC * k;
  const double zed = sqrt(a * a + 4.0 * (k - 4.0) * (ALK_DIC * ALK_DIC));
  hco3 = div_by(DIC * k - zed, rk.km4, rk.rkm4);
  const double t = ALK_DIC * k - DIC * k - 4.0 * ALK_DIC + zed;
  co3 = div_by(t, rk.km4x2, rk.rkm4x2);
  co2 = DIC - ALK_DIC + co3;
  H1 = cc[CC_K1] * co2 / hco3;
  H2 = cc[CC_K2] * hco3 / co3;
}
// sub_calc_carb + sub_calc_carb_RF0; returns false when the reference sets error_stop
__device__ bool solve_carb(double DIC, double ALK, double Ca, double PO4tot, double S, const double *cc, Carb &c, bool with_RF0) {
  const double Btot = fS(0.000416, S), SO4tot = fS(0.02824, S), Ftot = fS(0.00007, S);
  double H = c.H, H_old, co2, co3, hco3, H1, H2;
  CarbR rk;
  carb_recips(cc, rk);
  int n = 1;
  for (;;) {
    H_old = H;
    carb_iter(DIC, ALK, PO4tot, Btot, SO4tot, Ftot, cc, rk, H, co2, co3, hco3, H1, H2);
    if ((H1 < kNS) || (H2 < kNS)) return false;
    H = sqrt(H1 * H2);
    if (fabs(1.0 - H / H_old) < (1.0E-8 / H) * 0.001) {
      c.co2 = co2; c.co3 = co3; c.hco3 = hco3; c.ohm_cal = Ca * co3 / cc[CC_KCAL]; c.H = H;
      break;
    }
    n = n + 1;
    if (n > 100) return false;
  }
  if (!with_RF0) return true;
  {
    const double DIC_RF0 = DIC + 1.0e-6;
    n = 1;
    for (;;) {
      H_old = H;
      carb_iter(DIC_RF0, ALK, PO4tot, Btot, SO4tot, Ftot, cc, rk, H, co2, co3, hco3, H1, H2);
      H = sqrt(H1 * H2);
      if (fabs(1.0 - H / H_old) < 0.001) { c.RF0 = (co2 / c.co2 - 1.0) / (DIC_RF0 / DIC - 1.0); break; }
      n = n + 1;
      if ((H1 < kNS) || (H2 < kNS) || (n > 100)) { c.RF0 = 0.0; break; }
    }
  }
  return true;
}
// sub_calc_carb_r13C (mult = 1) / r14C (mult = 2): r of CO2(aq) and HCO3-
__device__ void carb_riso(double T, double DIC, double DICiso, const Carb &c, double mult, double standard, double &rCO2, double &rHCO3) {
  const double TC = T - kZeroC;
  const double d = iso_delta(DIC, DICiso, standard);
  double e_bg, e_dg, e_cg;
  if (mult == 1.0) { e_bg = -0.1141 * TC + 10.78; e_dg = +0.0049 * TC - 1.31; e_cg = -0.052 * TC + 7.22; }
  else { e_bg = 2.0 * (-0.1141 * TC + 10.78); e_dg = 2.0 * (+0.0049 * TC - 1.31); e_cg = 2.0 * (-0.052 * TC + 7.22); }
  const double e_cb = e_cg - e_bg / (1.0 + e_bg * 1.0E-3);
  const double e_db = e_dg - e_bg / (1.0 + e_bg * 1.0E-3);
  const double dHCO3 = (d * DIC - (e_db * c.co2 + e_cb * c.co3)) /
                       ((1.0 + e_db * 1.0E-3) * c.co2 + c.hco3 + (1.0 + e_cb * 1.0E-3) * c.co3);
  const double dCO2 = e_db + dHCO3 * (1.0 + e_db * 1.0E-3);
  rCO2 = iso_fraction(dCO2, standard);
  rHCO3 = iso_fraction(dHCO3, standard);
}
__device__ __forceinline__ double redfield_factor(const BgDev &b, double o2) {  // sub_box_remin_redfield, O2 only
  double loc_k = 0.0;
  const double O2 = fmax(0.0, o2);
  double kO2 = O2 / (O2 + b.remin_c0_O2);
  loc_k = loc_k + b.remin_k_O2 * kO2;
  if (loc_k < kNS) loc_k = 1.0;
  if (O2 < kNS) kO2 = 1.0;
  return b.remin_k_O2 * kO2 / loc_k;
}
}  // namespace bgk

// Compact tracer layout of the frozen configuration (cg_biogem.cu builds the tables; bg_layout_ok() on the host refuses
// anything else).  k_bg_step is written against this layout so that the tracer relationships of
// sub_data_update_tracerrelationships (biogem_data.f90:731-920) become straight-line code in the reference's loop order.
namespace lay {
enum { T = 1, S, DIC, DIC13, DIC14, PO4, O2, ALK, DOMC, DOMC13, DOMC14, DOMP, CA, CFC11, CFC12, MG, NL = 16 };
enum { POC = 1, POC13, POC14, POP, CACO3, CACO313, CACO314, POCF2, CACO3F2, NLS = 9 };
enum { A_T = 1, A_Q, A_CO2, A_CO213, A_CO214, A_O2, A_CFC11, A_CFC12, NLA = 8 };
}  // namespace lay
// loc_k_mld of sub_calc_bio_uptake (biogem_box.f90:423-430): the shallowest level whose floor lies at or below the mixed-layer depth;
// export production, its DOM fraction and the nutrient uptake are applied to every level k_mld .. n_k (:1186-1378).  Without the
// mixed-layer scheme (imld = 0: mld = 0) that is the top level alone.
__device__ __forceinline__ int bg_k_mld(const BgDev &b, const int K, const int k1, const size_t q) {
  if (!b.mld) return K;
  const double mld = b.mld[q];
  int km = k1;
  for (int k = K; k >= 1; k--)
    if (b.Dbot[k] >= mld) { km = k; break; }
  return km;
}
bool bg_layout_ok(const BgDev &b, int L) {
  using namespace lay;
  if (L != NL || b.LS != NLS || b.LA != NLA) return false;
  if (b.l_DIC != DIC || b.l_DIC13 != DIC13 || b.l_DIC14 != DIC14 || b.l_PO4 != PO4 || b.l_O2 != O2 || b.l_ALK != ALK ||
      b.l_DOMC != DOMC || b.l_Ca != CA || b.l_Mg != MG) return false;
  if (b.s_POC != POC || b.s_POC13 != POC13 || b.s_POC14 != POC14 || b.s_POP != POP || b.s_CaCO3 != CACO3 ||
      b.s_CaCO313 != CACO313 || b.s_CaCO314 != CACO314 || b.s_POCf2 != POCF2 || b.s_CaCO3f2 != CACO3F2) return false;
  if (b.a_CO2 != A_CO2 || b.a_CO213 != A_CO213 || b.a_CO214 != A_CO214) return false;
  const int want_n[NLS + 1] = {0, 2, 1, 1, 3, 3, 1, 1, 0, 0};
  const int want_lo[NLS + 1][3] = {{0, 0, 0}, {DIC, O2, 0}, {DIC13, 0, 0}, {DIC14, 0, 0}, {PO4, O2, ALK}, {DIC, ALK, CA},
                                   {DIC13, 0, 0}, {DIC14, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int ls = 1; ls <= NLS; ls++) {
    if (b.n_ls_lo[ls] != want_n[ls]) return false;
    for (int r = 0; r < want_n[ls]; r++) if (b.ls_lo[ls][r] != want_lo[ls][r]) return false;
  }
  // unit coefficients that the kernel folds away
  if (b.conv_ls_lo[POC][0] != 1.0 || b.conv_ls_lo[POC13][0] != 1.0 || b.conv_ls_lo[POC14][0] != 1.0 || b.conv_ls_lo[POP][0] != 1.0 ||
      b.conv_ls_lo[CACO3][0] != 1.0 || b.conv_ls_lo[CACO3][2] != 1.0 || b.conv_ls_lo[CACO313][0] != 1.0 || b.conv_ls_lo[CACO314][0] != 1.0)
    return false;
  if (b.dom2pom[DOMC] != POC || b.dom2pom[DOMC13] != POC13 || b.dom2pom[DOMC14] != POC14 || b.dom2pom[DOMP] != POP) return false;
  if (b.atm2ocn[A_CO2] != DIC || b.atm2ocn[A_CO213] != DIC13 || b.atm2ocn[A_CO214] != DIC14 || b.atm2ocn[A_O2] != O2 ||
      b.atm2ocn[A_CFC11] != CFC11 || b.atm2ocn[A_CFC12] != CFC12) return false;
  const int want_st[NLS + 1] = {0, 1, 11, 12, 3, 1, 11, 12, 9, 9};
  for (int ls = 1; ls <= NLS; ls++) if (b.stype[ls] != want_st[ls]) return false;
  return true;
}

// dissolved products of remineralising particulates, accumulated in the order of the reference's (ls, io) loops
struct Rem7 { double dic, d13, d14, po4, o2, alk, ca; };
__device__ __forceinline__ void rem_zero(Rem7 &r) { r.dic = r.d13 = r.d14 = r.po4 = r.o2 = r.alk = r.ca = 0.0; }
// r += (f*conv_ls_lo(lo,ls)) * p(ls) for ls = POC .. CaCO3_14C (conv = 1 folds to f)
__device__ __forceinline__ void rem_add(Rem7 &r, const double f, const double *p, const double fO2POC, const double fO2POP,
                                        const double fALKPOP, const double fALKCa) {
  using namespace lay;
  r.dic = r.dic + f * p[POC];
  r.o2 = r.o2 + fO2POC * p[POC];
  r.d13 = r.d13 + f * p[POC13];
  r.d14 = r.d14 + f * p[POC14];
  r.po4 = r.po4 + f * p[POP];
  r.o2 = r.o2 + fO2POP * p[POP];
  r.alk = r.alk + fALKPOP * p[POP];
  r.dic = r.dic + f * p[CACO3];
  r.alk = r.alk + fALKCa * p[CACO3];
  r.ca = r.ca + f * p[CACO3];
  r.d13 = r.d13 + f * p[CACO313];
  r.d14 = r.d14 + f * p[CACO314];
}

// fuse != 0: biogem_tracercoupling's steps (2)+(3) (biogem.f90:2033-2061, k_tc_apply) are applied to each cell as soon
// as its anomaly is known -- same expressions, same order, bit-identical -- instead of writing vdocn and re-reading ocn,
// ts, vdocn and bio_part in a separate pass.  The global sums of step (1) do not depend on step_biogem's output (the
// salinity anomaly is +0.0), so the caller takes them first (k_tc_partial phases 2 and 1).
// PART = 0: the whole step.  PART = 1: the surface cell only (carbonate chemistry, gas exchange, restoring, export
// production) -- it reads nothing but BIOGEM's own state of the previous block, so cg_run issues it one block ahead, next to
// the latency-bound momentum kernels of the two ocean cycles in between; its results cross to PART = 2 (sediment return +
// water-column sweep, at the nominal time) through b.surf (kBgSurfSlots doubles per member-column).  Same operations in
// the same order: bit-identical to PART = 0.
// FIX: the grid shape and the member stride of the bench configuration (36 x 36 x 16, 128 members) as compile-time
// constants -- every address becomes base + immediate (40 % of the generic kernel's instructions are 64-bit address
// arithmetic).
template <int MINB, int PART, int FMS>
__global__ void __launch_bounds__(128, MINB) k_bg_step(const Dev v, const BgDev b, const int init_only, const int mode) {
  const bool fuse = (mode & 1) != 0, pf = (mode & 2) == 0, pend = (mode & 4) != 0;
  // PART = 3: PART 2 without its per-cell half -- sediment return, the packets of sinking particles, the new particulate field,
  // the settling flux; the remineralisation products of every cell (lrem, 7 numbers) and the column's sediment return go to
  // b.lrem / b.fsedv, and the per-cell half (DOM, decay, anomaly, tracer coupling) runs cell-parallel in k_bg_cell.
  // pend: bio_part still lacks the last coupling's rescaling (Sratio, b.pscale): applied to the values as they are read
  constexpr bool PK = (PART == 3);
  using namespace bgk;
  using namespace lay;
  constexpr bool FIX = FMS > 0;
  const int I = FIX ? 36 : v.I, J = FIX ? 36 : v.J, K = FIX ? 16 : v.K, MS = FIX ? FMS : v.MS;
  constexpr int L = NL, LS = NLS, LA = NLA;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y * blockDim.y + threadIdx.y;
  __shared__ double s_f[L][32], s_rmean[32], s_sr[32], s_rsr[32], s_mnew[32];
  if (fuse) {   // per-member totals of the coupling, staged once per block (blockDim = (32, 4), m < MS always)
    const int lane = threadIdx.x;
    if (threadIdx.y == 0) {
      const double mean_S_OLD = v.bg_tot[m], mean_S_NEW = v.bg_tot[MS + m];
      s_rmean[lane] = 1.0 / mean_S_OLD;
      const double Sratio = mean_S_NEW / mean_S_OLD;
      s_sr[lane] = Sratio;
      s_rsr[lane] = 1.0 / Sratio;
      s_mnew[lane] = mean_S_NEW;
    }
    for (int l = 2 + threadIdx.y; l < L; l += blockDim.y) {
      const double told = v.bg_tot[(size_t)l * MS + m], tnew = v.bg_tot[(size_t)(L - 2 + l) * MS + m];
      const double rtnew = (fabs(tnew) < kBgNullSmall) ? 0.0 : 1.0 / tnew;
      s_f[l][lane] = told * rtnew;
    }
    __syncthreads();
  }
  if (m >= MS || n >= v.nwet) return;
  const int c2 = v.bgcols[n];
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1 = (int)v.k1[i + (I + 2) * j];
  const size_t c2d = cell2(I, i, j);
  const int k_mld = bg_k_mld(b, K, k1, c2d * MS + m);
  const size_t sK = (size_t)I * J * L * MS, o0 = cell3(I, J, i, j, 1) * L * MS + m;
  const size_t qK = (size_t)I * J * LS * MS, q0 = cell3(I, J, i, j, 1) * LS * MS + m;
  const size_t pK = (size_t)I * J * MS, p0 = cell3(I, J, i, j, 1) * MS + m;
#define OCN_(l, k) v.bg_ocn[o0 + (size_t)((k)-1) * sK + (size_t)((l)-1) * MS]
#define DOCN_(l, k) v.bg_vdocn[o0 + (size_t)((k)-1) * sK + (size_t)((l)-1) * MS]
#define PART_(ls, k) b.bio_part[q0 + (size_t)((k)-1) * qK + (size_t)((ls)-1) * MS]
#define M_(k) v.bg_M[p0 + (size_t)((k)-1) * pK]
#define RM_(k) v.bg_rM[p0 + (size_t)((k)-1) * pK]
#define SET1_(ls) b.settle_k1[(c2d * LS + ((ls)-1)) * MS + m]
  const double dtyr = b.dtyr;
  const double Tsf = PK ? 0.0 : OCN_(T, K), Ssf = PK ? 0.0 : OCN_(S, K);
  double cc[N_CC];
  Carb cb;
  if (init_only) {  // sub_init_carb, biogem_data.f90:2336-2430 (surface cell)
    carbconst(b.Dmid_surf, Tsf, Ssf, OCN_(CA, K), OCN_(MG, K), cc);
    cb.H = pow(10.0, -7.8);
    if (!solve_carb(OCN_(DIC, K), OCN_(ALK, K), OCN_(CA, K), OCN_(PO4, K), Ssf, cc, cb, false)) b.err[m] = 1;
    b.carbH[c2d * MS + m] = cb.H;
    return;
  }
  // conv_ls_lo coefficients that are not 1
  const double cO2POC = b.conv_ls_lo[POC][1], cO2POP = b.conv_ls_lo[POP][1], cALKPOP = b.conv_ls_lo[POP][2],
               cALKCa = b.conv_ls_lo[CACO3][1];
#define SC_(slot) b.surf[((size_t)(slot) * (I * J) + n) * MS + m]
  // ---- closed-system sediment return (:887-940) from the settling flux of the previous step
  Rem7 fsed;
  rem_zero(fsed);
  if (PART != 1) {
    const double f = redfield_factor(b, OCN_(O2, k1));
    double st[LS + 1];
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) st[ls] = SET1_(ls);
    rem_add(fsed, f, st, f * cO2POC, f * cO2POP, f * cALKPOP, f * cALKCa);
  }
  const double A = b.A[c2d], rA = b.rA[c2d];
  double focn_surf[LA + 1];   // -conv_atm_ocn*focnatm of each gas, applied to its ocean tracer at the surface (:1612-1616)
  double psurf[LS + 1], dom_add[LS + 1];
  Rem7 uptake;
  if (PART != 2 && PART != 3) {
  // ---- surface cell: carbonate chemistry, solubility, piston velocity (:1026-1104)
  carbconst(b.Dmid_surf, Tsf, Ssf, OCN_(CA, K), OCN_(MG, K), cc);
  cb.H = b.carbH[c2d * MS + m];
  cb.RF0 = 0.0;
  const double DICs = OCN_(DIC, K), PO4s = OCN_(PO4, K);
  if (!solve_carb(DICs, OCN_(ALK, K), OCN_(CA, K), PO4s, Ssf, cc, cb, true)) {
    if (PART == 1) SC_(kBgSurfSlots - 1) = 1.0; else b.err[m] = 1;
    return;
  }
  if (PART == 1) SC_(kBgSurfH) = cb.H; else b.carbH[c2d * MS + m] = cb.H;
  double r13_CO2, r13_HCO3, r14_CO2, r14_HCO3;
  carb_riso(Tsf, DICs, OCN_(DIC13, K), cb, 1.0, kStd13C, r13_CO2, r13_HCO3);
  carb_riso(Tsf, DICs, OCN_(DIC14, K), cb, 2.0, kStd14C, r14_CO2, r14_HCO3);
  const double rho = calc_rho(Tsf, Ssf);   // phys_ocn(ipo_rho) as left by biogem_climate (:2171)
  const double seaice = b.seaice[c2d * MS + m];
  {
    double fatm[LA + 1], focnatm[LA + 1], f_oa[LA + 1], f_ao[LA + 1];
    double TC = Tsf - kZeroC, TC2, TC3;
    const double TCf = Tsf - kZeroC;
    if (TC < 0.0) TC = 0.0;
    if (TC > 30.0) TC = 30.0;
    TC2 = TC * TC; TC3 = TC2 * TC;
    const double ws = b.wspeed[c2d];
    const double u2 = ws * ws;
    const double area = (1.0 - seaice) * A;
    double alpha_as = 0.0, alpha_sa = 0.0;
#pragma unroll
    for (int la = 1; la <= LA; la++) { fatm[la] = 0.0; focnatm[la] = 0.0; f_oa[la] = 0.0; f_ao[la] = 0.0; }
    double sfc[LA + 1];
#pragma unroll
    for (int la = 3; la <= LA; la++) sfc[la] = b.sfcatm1[((size_t)(la - 1) * I * J + c2d) * MS + m];
    // restoring of the atmosphere (:1119-1146)
#pragma unroll
    for (int la = 3; la <= LA; la++)
      if (b.rst_active[la]) {
        double tgt = b.rst_target[la];
        if (b.atype[la] == 1) { if (tgt < 0.0) tgt = sfc[la]; } else { if (tgt <= kNull) tgt = sfc[la]; }
        const double d = (tgt - sfc[la]) * b.tmod[la];
        fatm[la] = (1.0 / (double)(I * J)) * kAtmMol * d * (1.0 / dtyr);
      }
    // air-sea gas exchange (fun_calc_ocnatm_flux :123-299)
    double Ts, Ss;   // fun_calc_solconst clamps (gem_carbchem.f90:1396-1409)
    if (Tsf < kZeroC + 2.0) Ts = kZeroC + 2.0; else if (Tsf > (kZeroC + 35.0)) Ts = kZeroC + 35.0; else Ts = Tsf;
    if (Ssf < 26.0) Ss = 26.0; else if (Ssf > 43.0) Ss = 43.0; else Ss = Ssf;
    const double rT = 1.0 / Ts, Tr100 = Ts / 100.0, lnTr100 = log(Tr100);
#pragma unroll
    for (int la = 3; la <= LA; la++) {
      if (la == A_CO2 || la == A_O2 || la == A_CFC11 || la == A_CFC12) {
        const double *bc = b.bunsen[la];
        double sol = exp(bc[0] + bc[1] * (100 * rT) + bc[2] * lnTr100 + Ss * (bc[3] + bc[4] * (Tr100) + bc[5] * (Tr100 * Tr100)));
        if (la == A_O2) sol = sol / (rho * kVmol);
        const double Sc = b.Sc[la][0] - b.Sc[la][1] * TC + b.Sc[la][2] * TC2 - b.Sc[la][3] * TC3;
        const double pv = (1.0 / 1.0E+02) * (24.0 * 365.25) * b.gastransfer_a * u2 * (1.0 / sqrt(Sc * 1.515E-3));   // (Sc/660)**(-0.5)
        double loc_atm = sol * sfc[la], loc_ocn, buff;
        if (la == A_CO2) {
          loc_ocn = cb.co2;
          if (cb.RF0 > kNS) buff = 1.0 / (cb.RF0 * cb.co2 / DICs);
          else { loc_ocn = 0.0; loc_atm = 0.0; buff = 1.0; }
        } else {
          loc_ocn = OCN_((la == A_O2 ? O2 : (la == A_CFC11 ? CFC11 : CFC12)), K);
          buff = 1.0;
        }
        if (loc_ocn < kNS) loc_ocn = 0.0;
        if (loc_atm < kNS) loc_atm = 0.0;
        f_oa[la] = pv * area * rho * loc_ocn;
        f_ao[la] = pv * area * rho * loc_atm;
        const double deqm = b.dD[K] * A * rho * buff * fabs(loc_atm - loc_ocn);
        const double dflux = dtyr * fabs(f_oa[la] - f_ao[la]);
        if (deqm > kNS) {
          const double r = dflux / deqm;
          if (r > 1.00) { f_oa[la] = (1.00 / r) * f_oa[la]; f_ao[la] = (1.00 / r) * f_ao[la]; }
        }
      } else if (la == A_CO213) {
        const double r_atm = sfc[la] / sfc[A_CO2];
        const double R_atm = r_atm / (1.0 - r_atm), R_ocn = r13_CO2 / (1.0 - r13_CO2);
        const double alpha_k = 0.99912, alpha_alpha = 0.99869 + 4.9E-6 * TCf;
        alpha_as = alpha_alpha * alpha_k; alpha_sa = alpha_k;
        f_ao[la] = (alpha_as * R_atm / (1.0 + alpha_as * R_atm)) * f_ao[A_CO2];
        f_oa[la] = (alpha_sa * R_ocn / (1.0 + alpha_sa * R_ocn)) * f_oa[A_CO2];
      } else if (la == A_CO214) {
        const double r_atm = sfc[la] / sfc[A_CO2];
        const double R_atm = r_atm / (1.0 - r_atm), R_ocn = r14_CO2 / (1.0 - r14_CO2);
        f_ao[la] = ((alpha_as * alpha_as) * R_atm / (1.0 + (alpha_as * alpha_as) * R_atm)) * f_ao[A_CO2];
        f_oa[la] = ((alpha_sa * alpha_sa) * R_ocn / (1.0 + (alpha_sa * alpha_sa) * R_ocn)) * f_oa[A_CO2];
      }
      focnatm[la] = f_oa[la] - f_ao[la];
    }
#pragma unroll
    for (int la = 3; la <= LA; la++) {
      fatm[la] = fatm[la] + focnatm[la];
      focn_surf[la] = 0.0 - 1.0 * focnatm[la];
      const size_t qa = ((size_t)(la - 1) * I * J + c2d) * MS + m;
      // interface (:1731-1734) and cpl_flux_ocnatm (atchem.f90:306-320): sfxsumatm += dts*sfxatm1
      const double sfx = rA * (1.0 / kYrS) * fatm[la];
      if (PART == 1) {   // no side effect outside b.surf: a part issued ahead can be dropped (PART 2 commits)
        SC_(kBgSurfH + 1 + (la - 3)) = focnatm[la];
        SC_(kBgSurfH + 1 + (LA - 2) + (la - 3)) = sfx;
      } else {
        b.focnatm[qa] = focnatm[la];
        b.sfxsumatm[qa] = b.sfxsumatm[qa] + b.dts * sfx;
        if (b.sfxatm1) b.sfxatm1[qa] = sfx;
      }
    }
  }
  // ---- biological uptake at the surface (k_mld = K because mld = 0), sub_calc_bio_uptake 1N1T_PO4MM
  {
    const double kPO4 = PO4s / (PO4s + b.c0_PO4);
    const double ficefree = (1.0 - seaice);
    const double solfor = b.nsol > 0 ? v.solfor[(size_t)(b.nsol - 1) * J + (j - 1)] : 0.0;
    const double kI = solfor / b.solar_constant;
    double dPO4;
    if (PO4s > kNS) dPO4 = dtyr * ficefree * kI * kPO4 * b.k0_PO4[m]; else dPO4 = 0.0;
    double DOMfrac = b.red_DOMfrac, RDOMfrac = b.red_RDOMfrac, DOMtotal = DOMfrac + RDOMfrac;
    if (DOMtotal > 1.0) { DOMfrac = DOMfrac / DOMtotal; RDOMfrac = 1.0 - DOMfrac; DOMtotal = 1.0; }
    double red_POC_CaCO3;
    if (cb.ohm_cal > 1.0) red_POC_CaCO3 = (1.0 - DOMtotal) * b.red_POC_CaCO3[m] * pow(cb.ohm_cal - 1.0, b.red_POC_CaCO3_pP);
    else red_POC_CaCO3 = 0.0;
    const double Kq = 3.170E-05 + (-1.788E-07) * Tsf + 2.829E-10 * (Tsf * Tsf);
    const double delta_Corg = -b.d13C_DIC_Corg_ef + (b.d13C_DIC_Corg_ef - 0.7) * Kq / cb.co2;
    double alpha = 1.0 + delta_Corg / 1000.0, R = r13_CO2 / (1.0 - r13_CO2);
    const double red_POC13 = alpha * R / (1.0 + alpha * R);
    alpha = 1.0 + 2.0 * delta_Corg / 1000.0; R = r14_CO2 / (1.0 - r14_CO2);
    const double red_POC14 = alpha * R / (1.0 + alpha * R);
    const double delta_CaCO3 = 15.10 - 4232.0 / Tsf;
    alpha = 1.0 + delta_CaCO3 / 1000.0; R = r13_HCO3 / (1.0 - r13_HCO3);
    const double red_Ca13 = alpha * R / (1.0 + alpha * R);
    alpha = 1.0 + 2.0 * delta_CaCO3 / 1000.0; R = r14_HCO3 / (1.0 - r14_HCO3);
    const double red_Ca14 = alpha * R / (1.0 + alpha * R);
    // bulk export (:1186-1230): POC currency, CaCO3, POP, isotopes
    psurf[POC] = b.red_POP_POC * dPO4;
    psurf[POC] = 1.0 * psurf[POC];
    psurf[CACO3] = red_POC_CaCO3 * psurf[POC];
    psurf[POP] = (1.0 / b.red_POP_POC) * psurf[POC];
    psurf[POC13] = red_POC13 * psurf[POC];
    psurf[POC14] = red_POC14 * psurf[POC];
    psurf[CACO313] = red_Ca13 * psurf[CACO3];
    psurf[CACO314] = red_Ca14 * psurf[CACO3];
    psurf[POCF2] = 0.0; psurf[CACO3F2] = 0.0;
    rem_zero(uptake);   // inorganic uptake (:1254-1262): conv_sed_ocn*bio_part
    rem_add(uptake, 1.0, psurf, cO2POC, cO2POP, cALKPOP, cALKCa);
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) dom_add[ls] = 0.0;
    dom_add[POC] = 1.0 * DOMfrac * psurf[POC];       // DOM production (:1316-1352), r_POM_DOM = 1
    dom_add[POC13] = 1.0 * DOMfrac * psurf[POC13];
    dom_add[POC14] = 1.0 * DOMfrac * psurf[POC14];
    dom_add[POP] = 1.0 * DOMfrac * psurf[POP];
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) psurf[ls] = psurf[ls] - (dom_add[ls] + 0.0);
    const double kP = PO4s / (PO4s + b.POC_c0frac2);   // initial particulate fraction partitioning (:1354-1378)
    psurf[POCF2] = (1.0 - kP) * b.POC_dfrac2 + b.POC_frac2;
    psurf[CACO3F2] = b.CaCO3_frac2;
  }
  }   // PART != 2
  if (PART == 1) {   // hand the surface cell's results to the sweep kernel
#pragma unroll
    for (int la = 3; la <= LA; la++) SC_(la - 3) = focn_surf[la];
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) SC_(LA - 2 + ls - 1) = psurf[ls];
    SC_(LA - 2 + LS + 0) = dom_add[POC]; SC_(LA - 2 + LS + 1) = dom_add[POC13]; SC_(LA - 2 + LS + 2) = dom_add[POC14];
    SC_(LA - 2 + LS + 3) = dom_add[POP];
    SC_(LA - 2 + LS + 4) = uptake.dic; SC_(LA - 2 + LS + 5) = uptake.d13; SC_(LA - 2 + LS + 6) = uptake.d14;
    SC_(LA - 2 + LS + 7) = uptake.po4; SC_(LA - 2 + LS + 8) = uptake.o2; SC_(LA - 2 + LS + 9) = uptake.alk;
    SC_(LA - 2 + LS + 10) = uptake.ca;
    SC_(kBgSurfSlots - 1) = 0.0;
    return;
  }
  if (PART == 2 || PART == 3) {
    if (SC_(kBgSurfSlots - 1) != 0.0) { b.err[m] = 1; return; }   // the carbonate solve of this column failed (error_stop)
    b.carbH[c2d * MS + m] = SC_(kBgSurfH);
#pragma unroll
    for (int la = 3; la <= LA; la++) {
      const size_t qa = ((size_t)(la - 1) * I * J + c2d) * MS + m;
      b.focnatm[qa] = SC_(kBgSurfH + 1 + (la - 3));
      const double sfx = SC_(kBgSurfH + 1 + (LA - 2) + (la - 3));
      b.sfxsumatm[qa] = b.sfxsumatm[qa] + b.dts * sfx;
      if (b.sfxatm1) b.sfxatm1[qa] = sfx;
    }
#pragma unroll
    for (int la = 1; la <= LA; la++) focn_surf[la] = (la >= 3) ? SC_(la - 3) : 0.0;
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) { psurf[ls] = SC_(LA - 2 + ls - 1); dom_add[ls] = 0.0; }
    dom_add[POC] = SC_(LA - 2 + LS + 0); dom_add[POC13] = SC_(LA - 2 + LS + 1); dom_add[POC14] = SC_(LA - 2 + LS + 2);
    dom_add[POP] = SC_(LA - 2 + LS + 3);
    uptake.dic = SC_(LA - 2 + LS + 4); uptake.d13 = SC_(LA - 2 + LS + 5); uptake.d14 = SC_(LA - 2 + LS + 6);
    uptake.po4 = SC_(LA - 2 + LS + 7); uptake.o2 = SC_(LA - 2 + LS + 8); uptake.alk = SC_(LA - 2 + LS + 9);
    uptake.ca = SC_(LA - 2 + LS + 10);
  }
  if (PK) {
    double *fo = b.fsedv + (size_t)n * MS + m;
    const size_t fq = (size_t)v.nwet * MS;
    fo[0] = fsed.dic; fo[fq] = fsed.d13; fo[2 * fq] = fsed.d14; fo[3 * fq] = fsed.po4; fo[4 * fq] = fsed.o2; fo[5 * fq] = fsed.alk;
    fo[6 * fq] = fsed.ca;
  }
  const double pscale = (PK && pend) ? b.pscale[m] : 1.0;
  // ---- water column, one downward sweep (K -> k1).  sub_box_remin_part (:2412-2875) follows each source layer's
  // particles down to the deepest layer they reach in dt; here all packets in flight are advanced level by level, kept in
  // source order (shallowest source first = the reference's k loop), so every sum over sources at a given layer
  // (remineralisation products, parked particles, settling flux) is accumulated in the reference's order.  Per level the
  // sweep also does sub_box_remin_DOM (:2287-2406), the decay terms (:838-871) and the tracer anomaly (:1811-1844).
  double pk[kBgMaxK][LS + 1];   // TMP(:, current level) of each packet
  int pk_min[kBgMaxK];          // loc_bio_remin_min_k of each packet
  int npk = 0;
  double set1[LS + 1];
  const int klim = (dtyr * b.sinkingrate <= b.dsc) ? k1 : K;
  const double fd13 = b.fd_sed[POC14];   // decay factor of the 14C particulates (POC_14C and CaCO3_14C share lambda)
  const bool sed_decays = fabs(b.lam_sed[POC14]) > kNS;
  const bool ocn_decays = fabs(b.lam_ocn[DIC14]) > kNS;
  const double decay14 = 1.0 - b.fd_ocn[DIC14];
  double ratio;
  if (b.DOMlifetime > dtyr) ratio = dtyr / b.DOMlifetime; else ratio = 1.0;
#pragma unroll
  for (int ls = 1; ls <= LS; ls++) set1[ls] = 0.0;
  for (int kk = K; kk >= k1; kk--) {
    // pull the next level's rows towards the SM while this level is being worked on (no registers tied up)
    if (pf && kk > k1) {
      if (PK) asm volatile("prefetch.global.L1 [%0];" ::"l"(&OCN_(O2, kk - 1)));
      else {
#pragma unroll
      for (int l = 1; l <= L; l++) asm volatile("prefetch.global.L1 [%0];" ::"l"(&OCN_(l, kk - 1)));
      }
      if (fuse) {
#pragma unroll
        for (int l = 1; l <= L; l++) asm volatile("prefetch.global.L1 [%0];" ::"l"(v.ts_cur + (o0 + (size_t)(kk - 2) * sK + (size_t)(l - 1) * MS)));
      }
      if (!PK) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(&M_(kk - 1)));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(&RM_(kk - 1)));
      }
      if (kk - 1 >= klim) {
#pragma unroll
        for (int ls = 1; ls <= LS; ls++) asm volatile("prefetch.global.L1 [%0];" ::"l"(&PART_(ls, kk - 1)));
      }
    }
    Rem7 lrem;
    double pnew[LS + 1];
    rem_zero(lrem);
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) pnew[ls] = 0.0;
    // PART 3 does the per-cell half for the BOTTOM cell only (the bottom-water interface sfcocn1 is an output of step_biogem,
    // biogem.f90:1736-1744: it must stand when the step returns, not when the coupling has run); elsewhere it reads O2 alone
    const bool pk_skip = PK && kk != k1;
    const double Mk = pk_skip ? 0.0 : M_(kk), rM = pk_skip ? 0.0 : RM_(kk);
    double x[L + 1];
    if (pk_skip) x[O2] = OCN_(O2, kk);
    else {
#pragma unroll
      for (int l = 1; l <= L; l++) x[l] = OCN_(l, kk);
    }
    const double f = redfield_factor(b, x[O2]);
    const double fO2POC = f * cO2POC, fO2POP = f * cO2POP, fALKPOP = f * cALKPOP, fALKCa = f * cALKCa;
    // (1) packets from the layers above pass through / stop in layer kk
    if (npk > 0) {
      const double layerratio = b.dD[kk + 1] / b.dD[kk];
      const double Ca_f1 = b.CaCO3_f1[kk], Ca_f2 = b.CaCO3_f2[kk];
      const double PO_f1 = b.POC_f1[(size_t)kk * MS + m], PO_f2 = b.POC_f2[kk];
      int keep = 0;
      for (int p = 0; p < npk; p++) {
        double tp[LS + 1], tcur[LS + 1], pr[LS + 1];
#pragma unroll
        for (int ls = 1; ls <= LS; ls++) tp[ls] = pk[p][ls];
        const double Ca_ratio = 1.0 - ((1.0 - tp[CACO3F2]) * Ca_f1 + tp[CACO3F2] * Ca_f2);
        const double PO_ratio = 1.0 - ((1.0 - tp[POCF2]) * PO_f1 + tp[POCF2] * PO_f2);
        tcur[CACO3F2] = (tp[CACO3F2] > kNS) ? (1.0 - Ca_f2) * tp[CACO3F2] / Ca_ratio : 0.0;
        tcur[POCF2] = (tp[POCF2] > kNS) ? (1.0 - PO_f2) * tp[POCF2] / PO_ratio : 0.0;
        tcur[POC] = tp[POC] * layerratio * PO_ratio;
        tcur[POC13] = tp[POC13] * layerratio * PO_ratio;
        tcur[POC14] = tp[POC14] * layerratio * PO_ratio;
        tcur[POP] = tp[POP] * layerratio * PO_ratio;
        tcur[CACO3] = tp[CACO3] * layerratio * Ca_ratio;
        tcur[CACO313] = tp[CACO313] * layerratio * Ca_ratio;
        tcur[CACO314] = tp[CACO314] * layerratio * Ca_ratio;
#pragma unroll
        for (int ls = 1; ls <= LS; ls++) pr[ls] = (layerratio * tp[ls] - tcur[ls]);
        rem_add(lrem, f, pr, fO2POC, fO2POP, fALKPOP, fALKCa);
        const int mk = pk_min[p];
        if (kk == mk) {                   // deepest layer reached within dt: park the remainder here
#pragma unroll
          for (int ls = 1; ls <= LS; ls++) pnew[ls] = pnew[ls] + tcur[ls];
        } else if (kk == k1) {            // through the base of the deepest layer: settling flux
#pragma unroll
          for (int ls = 1; ls <= LS; ls++) set1[ls] = set1[ls] + ((ls >= POCF2) ? tcur[ls] : Mk * tcur[ls]);
        } else {                          // keeps sinking
#pragma unroll
          for (int ls = 1; ls <= LS; ls++) pk[keep][ls] = tcur[ls];
          pk_min[keep] = mk;
          keep++;
        }
      }
      npk = keep;
    }
    // (2) layer kk as a source (decayed particulates of the previous step, :862-871)
    if (kk >= klim) {
      double old[LS + 1];
#pragma unroll
      for (int ls = 1; ls <= LS; ls++) old[ls] = PART_(ls, kk);
      if (PK && pend) {   // biogem.f90:2042-2043 of the last coupling, not applied yet (k_bg_cell leaves it to the next reader)
#pragma unroll
        for (int ls = 1; ls <= LS; ls++) old[ls] = pscale * (old[ls] + 0.0);
      }
      if (sed_decays) { old[POC14] = fd13 * old[POC14]; old[CACO314] = fd13 * old[CACO314]; }
      double part_tot = 0.0;
      part_tot = part_tot + old[POC];
      part_tot = part_tot + old[CACO3];
      if (part_tot > kNS) {
        if (kk == k1) {
#pragma unroll
          for (int ls = 1; ls <= LS; ls++) set1[ls] = set1[ls] + ((ls >= POCF2) ? old[ls] : Mk * old[ls]);
        } else {
          const double max_D = b.Dbot[kk] + dtyr * b.sinkingrate;
          int min_k = k1 - 1;
          for (int k2 = kk - 1; k2 >= k1; k2--)
            if (b.Dbot[k2] > max_D) { min_k = k2; break; }
#pragma unroll
          for (int ls = 1; ls <= LS; ls++) pk[npk][ls] = old[ls];
          pk_min[npk] = min_k;
          npk++;
        }
      }
    }
    // (3) new particulate field of this layer
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) {
      const double pv = (kk >= k_mld) ? psurf[ls] : pnew[ls];
      PART_(ls, kk) = fuse ? s_sr[threadIdx.x] * (pv + 0.0) : pv;   // biogem.f90:2042-2043 (vdbio_part = 0)
    }
    if (PK) {   // the cell's half of the work is k_bg_cell's: hand over the remineralisation products of the particles
      double *lo = b.lrem + (cell3(I, J, i, j, kk) * 7) * MS + m;
      lo[0] = lrem.dic; lo[MS] = lrem.d13; lo[2 * (size_t)MS] = lrem.d14; lo[3 * (size_t)MS] = lrem.po4; lo[4 * (size_t)MS] = lrem.o2;
      lo[5 * (size_t)MS] = lrem.alk; lo[6 * (size_t)MS] = lrem.ca;
      if (pk_skip) continue;
    }
    // (4) sub_box_remin_DOM for this layer: DOM -> POM -> inorganic products
    const bool has_dom = x[DOMC] > kNS;
    double dom[LS + 1];
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) dom[ls] = 0.0;
    if (has_dom) {
      dom[POC] = dom[POC] + 1.0 * ratio * x[DOMC];
      dom[POC13] = dom[POC13] + 1.0 * ratio * x[DOMC13];
      dom[POC14] = dom[POC14] + 1.0 * ratio * x[DOMC14];
      dom[POP] = dom[POP] + 1.0 * ratio * x[DOMP];
    }
    Rem7 domrem;
    rem_zero(domrem);
    rem_add(domrem, f, dom, fO2POC, fO2POP, fALKPOP, fALKCa);
    // (5) tracer anomaly vdocn(l, kk) = bio_remin + dtyr*rM*focn and the bottom-water interface (:1736-1744)
    const bool bot = (kk == k1), top = (kk == K), prod = (kk >= k_mld);
    double rn_cpl = 0.0;   // mean_S_NEW / Snew of this cell (fused coupling)
#pragma unroll
    for (int l = 1; l <= L; l++) {
      double vrem = 0.0, lr = 0.0, fs = 0.0, up = 0.0, da = 0.0, gas = 0.0;
      bool slot = false;
      switch (l) {
        case DIC: vrem = vrem + domrem.dic; lr = lrem.dic; fs = fsed.dic; up = uptake.dic; slot = true; gas = focn_surf[A_CO2]; break;
        case DIC13: vrem = vrem + domrem.d13; lr = lrem.d13; fs = fsed.d13; up = uptake.d13; slot = true; gas = focn_surf[A_CO213]; break;
        case DIC14: vrem = vrem + domrem.d14; lr = lrem.d14; fs = fsed.d14; up = uptake.d14; slot = true; gas = focn_surf[A_CO214]; break;
        case PO4: vrem = vrem + domrem.po4; lr = lrem.po4; fs = fsed.po4; up = uptake.po4; slot = true; break;
        case O2: vrem = vrem + domrem.o2; lr = lrem.o2; fs = fsed.o2; up = uptake.o2; slot = true; gas = focn_surf[A_O2]; break;
        case ALK: vrem = vrem + domrem.alk; lr = lrem.alk; fs = fsed.alk; up = uptake.alk; slot = true; break;
        case CA: vrem = vrem + domrem.ca; lr = lrem.ca; fs = fsed.ca; up = uptake.ca; slot = true; break;
        case DOMC: if (has_dom) vrem = vrem - ratio * x[l]; da = dom_add[POC]; break;
        case DOMC13: if (has_dom) vrem = vrem - ratio * x[l]; da = dom_add[POC13]; break;
        case DOMC14: if (has_dom) vrem = vrem - ratio * x[l]; da = dom_add[POC14]; break;
        case DOMP: if (has_dom) vrem = vrem - ratio * x[l]; da = dom_add[POP]; break;
        case CFC11: gas = focn_surf[A_CFC11]; break;
        case CFC12: gas = focn_surf[A_CFC12]; break;
        default: break;
      }
      double rem = 0.0;
      if (bot && l >= 3) rem = rem + rM * fs;
      rem = rem + (vrem + lr);
      double focn = 0.0;
      if ((l == DIC14 || l == DOMC14) && ocn_decays) focn = focn - Mk * decay14 * x[l] / dtyr;
      if (l == 1 && bot) focn = focn + kYrS * b.Fgeothermal * A / (1.0E+03 * kCp);
      if (top) focn = focn + gas;
      if (prod && l >= 3) {
        if (l == DOMC || l == DOMC13 || l == DOMC14 || l == DOMP) rem = rem + da;
        rem = rem - (slot ? up : 0.0);
      }
      const double dval = rem + dtyr * rM * focn;
      if (bot) b.sfcocn1[((size_t)(l - 1) * I * J + c2d) * MS + m] = x[l] + rem + dtyr * rM * focn;
      if (PK) {
        // the anomaly itself is k_bg_cell's business
      } else if (!fuse) {
        DOCN_(l, kk) = dval;
      } else {
        double *tsp = v.ts_cur + (o0 + (size_t)(kk - 1) * sK + (size_t)(l - 1) * MS);
        if (l == 1) {
          const double Tn = *tsp + kBgZeroC + dval;
          OCN_(l, kk) = Tn;
          *tsp = Tn - kBgZeroC;
        } else if (l == 2) {
          const double Sn = *tsp + v.p.saln0[m] + dval;
          rn_cpl = s_mnew[threadIdx.x] / Sn;
          OCN_(l, kk) = Sn;
          *tsp = Sn - v.p.saln0[m];
        } else {
          const double lv = *tsp * x[S] * s_rmean[threadIdx.x];
          double xx = s_f[l - 1][threadIdx.x] * lv + dval;
          xx = s_sr[threadIdx.x] * xx;
          OCN_(l, kk) = xx;
          *tsp = rn_cpl * xx;
        }
      }
    }
    if (fuse) {
      M_(kk) = s_rsr[threadIdx.x] * Mk;
      RM_(kk) = s_sr[threadIdx.x] * rM;
    }
  }
#pragma unroll
  for (int ls = 1; ls <= LS; ls++) {
    SET1_(ls) = set1[ls];
    const double rdts = 1.0 / b.dts;
    b.sfxsed1[((size_t)(ls - 1) * I * J + c2d) * MS + m] = (ls >= POCF2) ? set1[ls] * rdts * dtyr : rA * set1[ls] * rdts;
  }
#undef OCN_
#undef DOCN_
#undef PART_
#undef M_
#undef RM_
#undef SET1_
#undef SC_
}

// =============================================================================================================
// The per-cell half of step_biogem's water-column work + biogem_tracercoupling's steps (2)+(3), cell-parallel: one warp = 32
// members of ONE wet cell (as k_tc_apply).  Behind k_bg_step PART 3 (which left the particles' remineralisation products of
// every cell in b.lrem and the column's sediment return in b.fsedv) and the coupling's sums / factors, the thread does what the
// sweep did per level -- sub_box_remin_DOM (biogem_box.f90:2287-2406), the decay terms (biogem.f90:838-871), the tracer anomaly
// (:1811-1844), the bottom-water interface (:1736-1744) -- and applies the coupling to the cell at once (:2033-2061): the
// anomaly never goes to memory (no vdocn round trip), ocn and ts are read once and written once.  Same expressions in the same
// order as k_bg_step (`fuse`) / k_tc_apply: bit-identical results.  bio_part is NOT rescaled here (a pass over 9 tracers for
// one multiplication): the factor stays pending in b.pscale and the next reader applies it (k_bg_step PART 3, k_bg_part_scale).
template <int FMS, int MINB, bool CS = false>
__global__ void __launch_bounds__(32 * kApplyWarps, MINB) k_bg_cell(const Dev v, const BgDev b) {
  using namespace bgk;
  using namespace lay;
  constexpr bool FIX = FMS > 0;
  constexpr int L = NL, LS = NLS, LA = NLA;
  __shared__ double s_f[L][32], s_rmean[32], s_sr[32], s_rsr[32], s_mnew[32];
  const int I = FIX ? 36 : v.I, J = FIX ? 36 : v.J, K = FIX ? 16 : v.K, MS = FIX ? FMS : v.MS;
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int m = blockIdx.x * 32 + lane;
  {
    const double *fac = v.bg_tot + (size_t)2 * L * MS + m;
    if (warp == 0) {
      s_rmean[lane] = fac[0];
      s_sr[lane] = fac[(size_t)MS];
      s_rsr[lane] = fac[(size_t)2 * MS];
      s_mnew[lane] = fac[(size_t)3 * MS];
    }
    for (int l = 2 + warp; l < L; l += kApplyWarps) s_f[l][lane] = fac[(size_t)(4 + l) * MS];
  }
  __syncthreads();
  const int c = blockIdx.y * kApplyWarps + warp;
  if (c >= I * J * K) return;
  const int kk = c / (I * J) + 1, r2 = c % (I * J), j = r2 / I + 1, i = r2 % I + 1;
  const int k1 = (int)v.k1[i + (I + 2) * j];
  if (kk < k1) return;
  const size_t c2d = cell2(I, i, j);
  const int n = b.colidx[c2d];
  if (kk == K && c2d == (size_t)v.bgcols[0]) b.pscale[m] = s_sr[lane];   // the rescaling of bio_part this coupling owes
#define SC_(slot) b.surf[((size_t)(slot) * (I * J) + n) * MS + m]
  if (SC_(kBgSurfSlots - 1) != 0.0) return;   // the carbonate solve of this column failed (error_stop): flagged by PART 3
  const bool bot = (kk == k1), top = (kk == K), prod = (kk >= bg_k_mld(b, K, k1, c2d * MS + m));
  const size_t o = (size_t)c * L * MS + m;
  const double dtyr = b.dtyr;
  double x[L + 1], tv[L + 1];
#pragma unroll
  for (int l = 1; l <= L; l++) {   // read once, written once: streaming (evict-first) accesses keep the pass out of L2's way
    x[l] = CS ? __ldcs(v.bg_ocn + o + (size_t)(l - 1) * MS) : v.bg_ocn[o + (size_t)(l - 1) * MS];
    tv[l] = CS ? __ldcs(v.ts_cur + o + (size_t)(l - 1) * MS) : v.ts_cur[o + (size_t)(l - 1) * MS];
  }
  const double Mk = v.bg_M[(size_t)c * MS + m], rM = v.bg_rM[(size_t)c * MS + m];
  Rem7 lrem, fsed, uptake;
  {
    const double *lo = b.lrem + ((size_t)c * 7) * MS + m;
    if (CS) {
      lrem.dic = __ldcs(lo); lrem.d13 = __ldcs(lo + MS); lrem.d14 = __ldcs(lo + 2 * (size_t)MS); lrem.po4 = __ldcs(lo + 3 * (size_t)MS);
      lrem.o2 = __ldcs(lo + 4 * (size_t)MS); lrem.alk = __ldcs(lo + 5 * (size_t)MS); lrem.ca = __ldcs(lo + 6 * (size_t)MS);
    } else {
      lrem.dic = lo[0]; lrem.d13 = lo[MS]; lrem.d14 = lo[2 * (size_t)MS]; lrem.po4 = lo[3 * (size_t)MS]; lrem.o2 = lo[4 * (size_t)MS];
      lrem.alk = lo[5 * (size_t)MS]; lrem.ca = lo[6 * (size_t)MS];
    }
  }
  rem_zero(fsed);
  rem_zero(uptake);
  if (bot) {
    const double *fo = b.fsedv + (size_t)n * MS + m;
    const size_t fq = (size_t)v.nwet * MS;
    fsed.dic = fo[0]; fsed.d13 = fo[fq]; fsed.d14 = fo[2 * fq]; fsed.po4 = fo[3 * fq]; fsed.o2 = fo[4 * fq]; fsed.alk = fo[5 * fq];
    fsed.ca = fo[6 * fq];
  }
  double focn_surf[LA + 1], dom_add[LS + 1];
#pragma unroll
  for (int la = 1; la <= LA; la++) focn_surf[la] = 0.0;
#pragma unroll
  for (int ls = 1; ls <= LS; ls++) dom_add[ls] = 0.0;
  if (top) {
#pragma unroll
    for (int la = 3; la <= LA; la++) focn_surf[la] = SC_(la - 3);
  }
  if (prod) {
    dom_add[POC] = SC_(LA - 2 + LS + 0); dom_add[POC13] = SC_(LA - 2 + LS + 1); dom_add[POC14] = SC_(LA - 2 + LS + 2);
    dom_add[POP] = SC_(LA - 2 + LS + 3);
    uptake.dic = SC_(LA - 2 + LS + 4); uptake.d13 = SC_(LA - 2 + LS + 5); uptake.d14 = SC_(LA - 2 + LS + 6);
    uptake.po4 = SC_(LA - 2 + LS + 7); uptake.o2 = SC_(LA - 2 + LS + 8); uptake.alk = SC_(LA - 2 + LS + 9);
    uptake.ca = SC_(LA - 2 + LS + 10);
  }
#undef SC_
  const double cO2POC = b.conv_ls_lo[POC][1], cO2POP = b.conv_ls_lo[POP][1], cALKPOP = b.conv_ls_lo[POP][2],
               cALKCa = b.conv_ls_lo[CACO3][1];
  const bool ocn_decays = fabs(b.lam_ocn[DIC14]) > kNS;
  const double decay14 = 1.0 - b.fd_ocn[DIC14];
  double ratio;
  if (b.DOMlifetime > dtyr) ratio = dtyr / b.DOMlifetime; else ratio = 1.0;
  const double A = b.A[c2d];
  const double f = redfield_factor(b, x[O2]);
  const double fO2POC = f * cO2POC, fO2POP = f * cO2POP, fALKPOP = f * cALKPOP, fALKCa = f * cALKCa;
  // (4) sub_box_remin_DOM for this layer: DOM -> POM -> inorganic products
  const bool has_dom = x[DOMC] > kNS;
  double dom[LS + 1];
#pragma unroll
  for (int ls = 1; ls <= LS; ls++) dom[ls] = 0.0;
  if (has_dom) {
    dom[POC] = dom[POC] + 1.0 * ratio * x[DOMC];
    dom[POC13] = dom[POC13] + 1.0 * ratio * x[DOMC13];
    dom[POC14] = dom[POC14] + 1.0 * ratio * x[DOMC14];
    dom[POP] = dom[POP] + 1.0 * ratio * x[DOMP];
  }
  Rem7 domrem;
  rem_zero(domrem);
  rem_add(domrem, f, dom, fO2POC, fO2POP, fALKPOP, fALKCa);
  // (5) tracer anomaly vdocn(l, kk) = bio_remin + dtyr*rM*focn, the bottom-water interface (:1736-1744), and the coupling
  const double saln0 = v.p.saln0[m];
  double rn_cpl = 0.0;   // mean_S_NEW / Snew of this cell
#pragma unroll
  for (int l = 1; l <= L; l++) {
    double vrem = 0.0, lr = 0.0, fs = 0.0, up = 0.0, da = 0.0, gas = 0.0;
    bool slot = false;
    switch (l) {
      case DIC: vrem = vrem + domrem.dic; lr = lrem.dic; fs = fsed.dic; up = uptake.dic; slot = true; gas = focn_surf[A_CO2]; break;
      case DIC13: vrem = vrem + domrem.d13; lr = lrem.d13; fs = fsed.d13; up = uptake.d13; slot = true; gas = focn_surf[A_CO213]; break;
      case DIC14: vrem = vrem + domrem.d14; lr = lrem.d14; fs = fsed.d14; up = uptake.d14; slot = true; gas = focn_surf[A_CO214]; break;
      case PO4: vrem = vrem + domrem.po4; lr = lrem.po4; fs = fsed.po4; up = uptake.po4; slot = true; break;
      case O2: vrem = vrem + domrem.o2; lr = lrem.o2; fs = fsed.o2; up = uptake.o2; slot = true; gas = focn_surf[A_O2]; break;
      case ALK: vrem = vrem + domrem.alk; lr = lrem.alk; fs = fsed.alk; up = uptake.alk; slot = true; break;
      case CA: vrem = vrem + domrem.ca; lr = lrem.ca; fs = fsed.ca; up = uptake.ca; slot = true; break;
      case DOMC: if (has_dom) vrem = vrem - ratio * x[l]; da = dom_add[POC]; break;
      case DOMC13: if (has_dom) vrem = vrem - ratio * x[l]; da = dom_add[POC13]; break;
      case DOMC14: if (has_dom) vrem = vrem - ratio * x[l]; da = dom_add[POC14]; break;
      case DOMP: if (has_dom) vrem = vrem - ratio * x[l]; da = dom_add[POP]; break;
      case CFC11: gas = focn_surf[A_CFC11]; break;
      case CFC12: gas = focn_surf[A_CFC12]; break;
      default: break;
    }
    double rem = 0.0;
    if (bot && l >= 3) rem = rem + rM * fs;
    rem = rem + (vrem + lr);
    double focn = 0.0;
    if ((l == DIC14 || l == DOMC14) && ocn_decays) focn = focn - Mk * decay14 * x[l] / dtyr;
    if (l == 1 && bot) focn = focn + kYrS * b.Fgeothermal * A / (1.0E+03 * kCp);
    if (top) focn = focn + gas;
    if (prod && l >= 3) {
      if (l == DOMC || l == DOMC13 || l == DOMC14 || l == DOMP) rem = rem + da;
      rem = rem - (slot ? up : 0.0);
    }
    const double dval = rem + dtyr * rM * focn;   // (the bottom-water interface sfcocn1 = x + dval was written by PART 3)
    double *tsp = v.ts_cur + (o + (size_t)(l - 1) * MS);
    double *ocp = v.bg_ocn + (o + (size_t)(l - 1) * MS);
    if (l == 1) {
      const double Tn = tv[l] + kBgZeroC + dval;
      if (CS) { __stcs(ocp, Tn); __stcs(tsp, Tn - kBgZeroC); } else { *ocp = Tn; *tsp = Tn - kBgZeroC; }
    } else if (l == 2) {
      const double Sn = tv[l] + saln0 + dval;
      rn_cpl = s_mnew[lane] / Sn;
      if (CS) { __stcs(ocp, Sn); __stcs(tsp, Sn - saln0); } else { *ocp = Sn; *tsp = Sn - saln0; }
    } else {
      const double lv = tv[l] * x[S] * s_rmean[lane];
      double xx = s_f[l - 1][lane] * lv + dval;
      xx = s_sr[lane] * xx;
      if (CS) { __stcs(ocp, xx); __stcs(tsp, rn_cpl * xx); } else { *ocp = xx; *tsp = rn_cpl * xx; }
    }
  }
  v.bg_M[(size_t)c * MS + m] = s_rsr[lane] * Mk;
  v.bg_rM[(size_t)c * MS + m] = s_sr[lane] * rM;
}
// the rescaling of bio_part a coupling through k_bg_cell left pending (biogem.f90:2042-2043), for readers other than k_bg_step
__global__ void __launch_bounds__(256) k_bg_part_scale(const Dev v, const BgDev b) {
  const int MS = v.MS;
  const int m = blockIdx.x * 32 + threadIdx.x;
  const size_t row = (size_t)blockIdx.y * blockDim.y + threadIdx.y;
  const size_t nrow = (size_t)v.I * v.J * v.K * b.LS;
  if (row >= nrow) return;
  double *p = b.bio_part + row * MS + m;
  *p = b.pscale[m] * (*p + 0.0);
}

// biogem_climate (:2132-2239): snapshot the sea-ice fraction, reset the convection counter
// (the cover is read through a staging copy taken by k_bg_stage_seaice at the reference's call time, so that the block
// may run while the next cycle's sea-ice step is already rewriting varice)
__global__ void k_bg_stage_seaice(const Dev v, const BgDev b) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)v.I * v.J * v.MS;
  if (q < n) {
    b.seaice_stage[q] = v.varice[n + q];   // varice(2,:,:) = fractional cover
    if (b.mld_stage) b.mld_stage[q] = v.mld[q];   // GOLDSTEIN's mixed-layer depth behind this cycle's tstepo (imld = 1)
    b.tq_stage[q] = v.tq[q];               // tstar_atm, surf_qstar_atm of this koverall iteration (embm.f90:177-193)
    b.tq_stage[n + q] = v.tq[n + q];
  }
}
__global__ void k_bg_climate(const Dev v, const BgDev b) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)v.I * v.J * v.MS;
  if (q >= n) return;
  b.seaice[q] = b.seaice_stage[q];
  if (b.mld) b.mld[q] = -5000.0 * b.mld_stage[q];   // go_mldta (goldstein.f90:449) -> phys_ocnatm(ipoa_mld) (biogem.f90:2183)
  v.cost[q] = 0.0;
}
// cpl_comp_EMBM (atchem.f90:270-282; genie.f90:454, behind the ATCHEM step): rows 1-2 of sfcatm1 = air temperature and
// humidity of this koverall iteration (staged by k_bg_stage_seaice on the caller's stream)
__global__ void k_bg_cpl_comp_embm(const Dev v, const BgDev b) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)2 * v.I * v.J * v.MS;
  if (q < n) b.sfcatm1[q] = b.tq_stage[q];
}

// step_atchem (atchem.f90:63-158) + cpl_comp_atmocn (:252-264).  k_bg_atchem1: thread = (member, cell, tracer), the
// per-cell decay + flux update and the mole product loc_conv_atm_mol*atm.  k_bg_atchem2: block = (32-member tile, tracer),
// the mole-weighted global sum (:146) taken in array-element order (ordered_sum_block: bit-identical to the sequential
// code), then the homogenised partial pressure is written back to atm and the interface array.
__global__ void __launch_bounds__(256) k_bg_atchem1(const Dev v, const BgDev b, double *__restrict__ scratch) {
  using namespace bgk;
  const int I = v.I, J = v.J, MS = v.MS;
  const int m = blockIdx.x * 32 + threadIdx.x;
  const int c = blockIdx.y * blockDim.y + threadIdx.y;
  const int la = 3 + blockIdx.z;
  const int ij = I * J;
  if (c >= ij) return;
  const size_t q = ((size_t)(la - 1) * ij + c) * MS + m;
  const double c_am = b.atm_V[c] / (kPaAtm * kRSI * b.atm[(size_t)c * MS + m]);
  const double c_ma = 1.0 / c_am;
  double a = b.atm[q];
  if (fabs(b.lam_atm[la]) > kNS) a = b.fd_atm[la] * a;
  const double F14C = 0.0;   // par_atm_F14C (atchem-defaults.nml)
  double fl = 0.0;
  if (la == b.a_CO214) fl = fl + b.dtyr_atchem * (1.0 / (double)(I * J)) * F14C;
  a = a + c_ma * b.atm_A[c] * b.sfxsumatm[q] + c_ma * fl;
  scratch[((size_t)(la - 3) * ij + c) * MS + m] = c_am * a;
}
__global__ void __launch_bounds__(32 * kSumWarps) k_bg_atchem2(const Dev v, const BgDev b, const double atm_totV, const double *__restrict__ scratch) {
  using namespace bgk;
  __shared__ double tile[kSumTile * 32];
  __shared__ double tot_s[32];
  const int I = v.I, J = v.J, MS = v.MS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * 32;
  const int la = 3 + blockIdx.y;
  const int ij = I * J;
  const double tot = ordered_sum_block(scratch + (size_t)(la - 3) * ij * MS + m0, (size_t)MS, ij, tile);
  if (warp == 0) tot_s[lane] = tot;
  __syncthreads();
  if (b.atm_tot) {   // the homogenised field is written by k_bg_atchem3 (one thread per element instead of 162 cells per warp)
    if (warp == 0) b.atm_tot[(size_t)(la - 1) * MS + m0 + lane] = tot;
    return;
  }
  const double t = tot_s[lane];
  double *__restrict__ atm = b.atm + (size_t)(la - 1) * ij * MS + m0 + lane;
  const double *__restrict__ atmT = b.atm + m0 + lane;
  double *__restrict__ sfx = b.sfxsumatm + (size_t)(la - 1) * ij * MS + m0 + lane;
  double *__restrict__ sfc = b.sfcatm1 + (size_t)(la - 1) * ij * MS + m0 + lane;
#pragma unroll 6
  for (int c = warp; c < ij; c += kSumWarps) {
    const double a = (t / atm_totV) * kPaAtm * kRSI * atmT[(size_t)c * MS];
    atm[(size_t)c * MS] = a;
    sfc[(size_t)c * MS] = a;
    sfx[(size_t)c * MS] = 0.0;
  }
}
// atchem.f90:146-150 + cpl_comp_atmocn (:252-264): the homogenised partial pressure of every cell from the member's total
__global__ void __launch_bounds__(256) k_bg_atchem3(const Dev v, const BgDev b, const double atm_totV) {
  using namespace bgk;
  const int MS = v.MS, ij = v.I * v.J;
  const int m = blockIdx.x * 32 + threadIdx.x;
  const int c = blockIdx.y * blockDim.y + threadIdx.y;
  const int la = 3 + blockIdx.z;
  if (c >= ij) return;
  const double t = b.atm_tot[(size_t)(la - 1) * MS + m];
  const size_t q = ((size_t)(la - 1) * ij + c) * MS + m;
  const double a = (t / atm_totV) * kPaAtm * kRSI * b.atm[(size_t)c * MS + m];
  b.atm[q] = a;
  b.sfcatm1[q] = a;
  b.sfxsumatm[q] = 0.0;
}

static bool bg_fix_shape(const Dev &v) { return v.I == 36 && v.J == 36 && v.K == 16 && fix_ms(v.MS) && !getenv("CG_BG_NOFIX"); }
int launch_bg_settle_sur(const Dev &, const BgDev &, int pend, cudaStream_t);   // extended time-series integrals: export through the surface layer's base
int launch_bg_step(const Dev &v, const BgDev &b, int init_only, int fuse, cudaStream_t s) {
  const int nx = init_only ? 0 : launch_bg_settle_sur(v, b, 0, s);
  // registers per thread 255 / 168 / 128 for MINB = 2 / 3 / 4 (CG_BG_MINB overrides; tuning knob)
  static int minb = -1;
  if (minb < 0) { const char *e = getenv("CG_BG_MINB"); minb = e ? atoi(e) : 2; }
  static int nopf = -1;
  if (nopf < 0) nopf = getenv("CG_BG_NOPF") ? 2 : 0;
  fuse |= nopf;
  const dim3 g(v.MS / 32, (v.nwet + 3) / 4), bl(32, 4);
  if (minb == 4) k_bg_step<4, 0, 0><<<g, bl, 0, s>>>(v, b, init_only, fuse);
  else if (minb == 3) k_bg_step<3, 0, 0><<<g, bl, 0, s>>>(v, b, init_only, fuse);
  else fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_step<2, 0, decltype(ms)::value><<<g, bl, 0, s>>>(v, b, init_only, fuse); });
  return 1 + nx;
}
// the two parts of the step (see k_bg_step): surface cell, then sediment return + water-column sweep
int launch_bg_surf(const Dev &v, const BgDev &b, cudaStream_t s) {
  static int minb = -1;   // registers per thread 134 / 128 for 3 / 4 (CG_BG_SURF_MINB; tuning knob)
  if (minb < 0) { const char *e = getenv("CG_BG_SURF_MINB"); minb = e ? atoi(e) : 4; }
  const dim3 g(v.MS / 32, (v.nwet + 3) / 4), bl(32, 4);
  if (minb == 3) k_bg_step<3, 1, 0><<<g, bl, 0, s>>>(v, b, 0, 0);
  else fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_step<4, 1, decltype(ms)::value><<<g, bl, 0, s>>>(v, b, 0, 0); });
  return 1;
}
// packets half of the sweep (PART 3) + its cell half fused with the coupling update (k_bg_cell); pend: see k_bg_step
int launch_bg_packets(const Dev &v, const BgDev &b, int pend, cudaStream_t s) {
  static int nopf = -1;
  if (nopf < 0) nopf = getenv("CG_BG_NOPF") ? 2 : 0;
  static int minb = -1;
  if (minb < 0) { const char *e = getenv("CG_BG_PK_MINB"); minb = e ? atoi(e) : 3; }
  const dim3 g(v.MS / 32, (v.nwet + 3) / 4), bl(32, 4);
  const int mode = nopf | (pend ? 4 : 0);
  const int nx = launch_bg_settle_sur(v, b, pend, s);
  if (minb == 2) fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_step<2, 3, decltype(ms)::value><<<g, bl, 0, s>>>(v, b, 0, mode); });
  else if (minb == 4) fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_step<4, 3, decltype(ms)::value><<<g, bl, 0, s>>>(v, b, 0, mode); });
  else fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_step<3, 3, decltype(ms)::value><<<g, bl, 0, s>>>(v, b, 0, mode); });
  return 1 + nx;
}
int launch_bg_cell(const Dev &v, const BgDev &b, cudaStream_t s) {
  const int ncell = v.I * v.J * v.K;
  static int minb = -1;   // 4: 128 registers (200 bytes of spills), 3: 168 registers
  if (minb < 0) { const char *e = getenv("CG_BG_CELL_MINB"); minb = e ? atoi(e) : 4; }   // measured: 8.49 vs 8.36 M model-years/hour
  const dim3 g(v.MS / 32, (ncell + kApplyWarps - 1) / kApplyWarps), bl(32, kApplyWarps);
  static int cs = -1;     // CG_BG_CELL_CS=1: streaming loads / stores (ld.global.cs / st.global.cs)
  if (cs < 0) { const char *e = getenv("CG_BG_CELL_CS"); cs = e ? atoi(e) : 1; }   // measured: 8.44 against 8.36 M model-years/hour
  if (minb == 4 && cs) fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_cell<decltype(ms)::value, 4, true><<<g, bl, 0, s>>>(v, b); });
  else if (minb == 4) fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_cell<decltype(ms)::value, 4><<<g, bl, 0, s>>>(v, b); });
  else fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_cell<decltype(ms)::value, 3><<<g, bl, 0, s>>>(v, b); });
  return 1;
}
int launch_bg_part_scale(const Dev &v, const BgDev &b, cudaStream_t s) {
  const size_t nrow = (size_t)v.I * v.J * v.K * b.LS;
  k_bg_part_scale<<<dim3(v.MS / 32, (unsigned)((nrow + 7) / 8)), dim3(32, 8), 0, s>>>(v, b);
  return 1;
}
int launch_bg_sweep(const Dev &v, const BgDev &b, cudaStream_t s, int fuse) {
  static int minb = -1;
  if (minb < 0) { const char *e = getenv("CG_BG_SWEEP_MINB"); minb = e ? atoi(e) : 2; }
  static int nopf = -1;
  if (nopf < 0) nopf = getenv("CG_BG_NOPF") ? 2 : 0;
  const dim3 g(v.MS / 32, (v.nwet + 3) / 4), bl(32, 4);
  const int nx = launch_bg_settle_sur(v, b, 0, s);
  if (minb == 3) k_bg_step<3, 2, 0><<<g, bl, 0, s>>>(v, b, 0, nopf | (fuse ? 1 : 0));
  else fix_dispatch(bg_fix_shape(v), v.MS, [&](auto ms) { k_bg_step<2, 2, decltype(ms)::value><<<g, bl, 0, s>>>(v, b, 0, nopf | (fuse ? 1 : 0)); });
  return 1 + nx;
}
// step (1) of biogem_tracercoupling taken BEFORE step_biogem (fused coupling, see k_bg_step)
int launch_tc_sums_first(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  const dim3 gc(v.MS / 32, (v.nwet + 3) / 4);
  const int L = v.L;
  fix_dispatch(tc_fix_shape(v), v.MS, [&](auto ms) { k_tc_partial<decltype(ms)::value><<<gc, b, 0, s>>>(v, 2); });
  k_tc_sum<<<dim3(v.MS / 32, L), 32 * kSumWarps, 0, s>>>(v, 0, L);
  fix_dispatch(tc_fix_shape(v), v.MS, [&](auto ms) { k_tc_partial<decltype(ms)::value><<<gc, b, 0, s>>>(v, 1); });
  k_tc_sum<<<dim3(v.MS / 32, L - 2), 32 * kSumWarps, 0, s>>>(v, L, 2 * L - 2);
  k_tc_factors<<<(v.MS + 127) / 128, 128, 0, s>>>(v);
  return 5;
}
// launch_tc_sums_first in two halves (see k_tc_partial phases 3 / 4): the half that reads BIOGEM's own state only ...
int launch_tc_sums_old(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  const dim3 gc(v.MS / 32, (v.nwet + 3) / 4);
  fix_dispatch(tc_fix_shape(v), v.MS, [&](auto ms) { k_tc_partial<decltype(ms)::value><<<gc, b, 0, s>>>(v, 3); });
  k_tc_sum<<<dim3(v.MS / 32, v.L), 32 * kSumWarps, 0, s>>>(v, 0, v.L, -2);
  return 2;
}
// ... and the half that needs this cycle's ts, + the per-member factors
int launch_tc_sums_new(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  const dim3 gc(v.MS / 32, (v.nwet + 3) / 4);
  const int L = v.L;
  fix_dispatch(tc_fix_shape(v), v.MS, [&](auto ms) { k_tc_partial<decltype(ms)::value><<<gc, b, 0, s>>>(v, 4); });
  k_tc_sum<<<dim3(v.MS / 32, L - 1), 32 * kSumWarps, 0, s>>>(v, L, 2 * L - 2, 1);
  k_tc_factors<<<(v.MS + 127) / 128, 128, 0, s>>>(v);
  return 3;
}
// steps (2)+(3) alone, after launch_tc_sums_first
int launch_tc_apply_only(const Dev &v, cudaStream_t s) {
  const int ncell = v.I * v.J * v.K, per_block = kApplyWarps * kApplyCellsPerWarp;
  fix_dispatch(tc_fix_shape(v), v.MS, [&](auto ms) { k_tc_apply<decltype(ms)::value><<<dim3(v.MS / 32, (ncell + per_block - 1) / per_block), dim3(32, kApplyWarps), 0, s>>>(v); });
  return 1;
}
int launch_bg_stage_seaice(const Dev &v, const BgDev &b, cudaStream_t s) {
  const size_t n = (size_t)v.I * v.J * v.MS;
  k_bg_stage_seaice<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v, b);
  return 1;
}
int launch_bg_climate(const Dev &v, const BgDev &b, cudaStream_t s) {
  const size_t n = (size_t)v.I * v.J * v.MS;
  k_bg_climate<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v, b);
  return 1;
}
int launch_bg_atchem(const Dev &v, const BgDev &b, double atm_totV, cudaStream_t s) {
  // scratch: the reduction buffer of the tracer coupling (idle here), (LA-2)*I*J*MS doubles
  const int ij = v.I * v.J;
  k_bg_atchem1<<<dim3(v.MS / 32, (ij + 7) / 8, b.LA - 2), dim3(32, 8), 0, s>>>(v, b, v.bg_part);
  k_bg_atchem2<<<dim3(v.MS / 32, b.LA - 2), 32 * kSumWarps, 0, s>>>(v, b, atm_totV, v.bg_part);
  if (b.atm_tot) k_bg_atchem3<<<dim3(v.MS / 32, (ij + 7) / 8, b.LA - 2), dim3(32, 8), 0, s>>>(v, b, atm_totV);
  const size_t n = (size_t)2 * ij * v.MS;
  k_bg_cpl_comp_embm<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v, b);
  return b.atm_tot ? 4 : 3;
}

int launch_tracercoupling(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  const dim3 gc(v.MS / 32, (v.nwet + 3) / 4);
  const int L = v.L;
  fix_dispatch(tc_fix_shape(v), v.MS, [&](auto ms) { k_tc_partial<decltype(ms)::value><<<gc, b, 0, s>>>(v, 0); });
  k_tc_sum<<<dim3(v.MS / 32, L), 32 * kSumWarps, 0, s>>>(v, 0, L);
  if (L > 2) {
    fix_dispatch(tc_fix_shape(v), v.MS, [&](auto ms) { k_tc_partial<decltype(ms)::value><<<gc, b, 0, s>>>(v, 1); });
    k_tc_sum<<<dim3(v.MS / 32, L - 2), 32 * kSumWarps, 0, s>>>(v, L, 2 * L - 2);
  }
  {
    const int ncell = v.I * v.J * v.K, per_block = kApplyWarps * kApplyCellsPerWarp;
    k_tc_factors<<<(v.MS + 127) / 128, 128, 0, s>>>(v);
    fix_dispatch(tc_fix_shape(v), v.MS, [&](auto ms) { k_tc_apply<decltype(ms)::value><<<dim3(v.MS / 32, (ncell + per_block - 1) / per_block), dim3(32, kApplyWarps), 0, s>>>(v); });
  }
  return L > 2 ? 6 : 4;
}
int launch_bg_reset_cost(const Dev &v, cudaStream_t s) {
  const size_t n = (size_t)v.I * v.J * v.MS;
  k_bg_reset_cost<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v);
  return 1;
}

// ---- SEDGEM coupler accumulations on the interface arrays (sediment grid = ocean grid) -------------------------------
// mode 0: cpl_flux_ocnsed, sedgem.f90:1029-1068        sum = sum + dts * src               (a = dts)
// mode 1: cpl_comp_ocnsed, sedgem.f90:894-937          sum = (w * sum + src) / (w + 1)     (a = w, b = w + 1)
// Elementwise over [tracer][j][i][member]: 3 x 8 bytes of HBM traffic per element, nothing else.
__global__ void __launch_bounds__(256) k_cpl_ocnsed(double *__restrict__ sum, const double *__restrict__ src, size_t n, double a,
                                                    double b, int mode) {
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x)
    sum[q] = mode ? (a * sum[q] + src[q]) / b : sum[q] + a * src[q];
}
int launch_cpl_ocnsed(double *sum, const double *src, size_t n, double a, double b, int mode, cudaStream_t s) {
  const size_t blocks = (n + 255) / 256;
  k_cpl_ocnsed<<<(unsigned)std::min<size_t>(blocks, 148 * 8), 256, 0, s>>>(sum, src, n, a, b, mode);
  return 1;
}

// ---- BIOGEM time-series integrals (diag_biogem_timeseries, biogem.f90:2836-2917: sig_ocn, sig_ocn_sur / _ben, sig_ocnatm) ------
// Quantities q (kSigHead = 3 ahead of the tracer blocks):
//   0 SUM(M)   1 SUM(M(:,:,n_k))   2 SUM((1-seaice)*A(n_k))
//   3+l        SUM(M*ocn(l))                                      over the wet cells
//   3+L+l      SUM((1-seaice)*A*ocn(l,:,:,n_k))                   ice-free surface
//   3+2L+l     SUM(mask_ben*A*ocn(l,i,j,k1))                      bottom cells deeper than par_data_save_ben_Dmin
//   3+3L+la    SUM(A*sfcatm1(la))                                 whole atmosphere grid
// Pass 1: block = (32-member tile, quantity), lanes = members, 8 warps stride over the cells; the 8 partial sums are added
// in warp order (deterministic; not the reference's element order -- relative difference ~1e-16, bar 1e-10).  Each block
// streams its operands once: M and one tracer of ocn for the 3-D sums (16 B per wet cell and member).
// Pass 2: one thread per member folds the sums into the window integrals in the reference's expression order.
__global__ void __launch_bounds__(256) k_bg_sig_sums(const Dev v, const BgDev b, const SigDev g) {
  __shared__ double part[8][32];
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int I = v.I, J = v.J, K = v.K, L = v.L, MS = v.MS, ij = I * J;
  const int m = blockIdx.x * 32 + lane, q = blockIdx.y;
  double s = 0.0;
  if (q == 0 || (q >= kSigHead && q < kSigHead + L)) {
    const int l = q - kSigHead;
    for (int c2 = warp; c2 < ij; c2 += 8) {
      for (int k = g.kbot[c2]; k < K; k++) {
        const size_t c = (size_t)k * ij + c2;
        const double M = v.bg_M[c * MS + m];
        s = s + (q == 0 ? M : M * v.bg_ocn[(c * L + l) * MS + m]);
      }
    }
  } else if (q == 1) {
    for (int c2 = warp; c2 < ij; c2 += 8)
      if (g.kbot[c2] < K) s = s + v.bg_M[((size_t)(K - 1) * ij + c2) * MS + m];
  } else if (q == 2 || (q >= kSigHead + L && q < kSigHead + 2 * L)) {
    const int l = q - kSigHead - L;
    for (int c2 = warp; c2 < ij; c2 += 8) {
      if (g.kbot[c2] >= K) continue;
      const double w = (1.0 - b.seaice[(size_t)c2 * MS + m]) * g.A[c2];
      s = s + (q == 2 ? w : w * v.bg_ocn[(((size_t)(K - 1) * ij + c2) * L + l) * MS + m]);
    }
  } else if (q < kSigHead + 3 * L) {
    const int l = q - kSigHead - 2 * L;
    for (int c2 = warp; c2 < ij; c2 += 8) {
      const double w = g.w_ben[c2];
      if (w == 0.0) continue;
      s = s + w * v.bg_ocn[(((size_t)g.kbot[c2] * ij + c2) * L + l) * MS + m];
    }
  } else {
    const int la = q - kSigHead - 3 * L;
    // rows 1-2 of sfcatm1 are the air temperature and humidity cpl_comp_EMBM copied behind the last ATCHEM step
    // (k_bg_cpl_comp_embm): at genie.f90's call point (before this block's ATCHEM step) those of the previous block
    for (int c2 = warp; c2 < ij; c2 += 8) s = s + g.A[c2] * b.sfcatm1[((size_t)la * ij + c2) * MS + m];
  }
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
    double t = part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; w++) t = t + part[w][lane];
    g.raw[(size_t)q * MS + m] = t;
  }
}
__global__ void k_bg_sig_acc(const Dev v, const SigDev g, const double dtyr) {
  const int MS = v.MS, L = v.L;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= MS) return;
  const double *r = g.raw + m;
  double *a = g.acc + m;
  const double tot_M = r[0], tot_A = r[(size_t)2 * MS];
  const double rtot_M = tot_M > kBgNullSmall ? 1.0 / tot_M : 0.0, rtot_A = tot_A > kBgNullSmall ? 1.0 / tot_A : 0.0;
  a[0] = a[0] + dtyr;                                            // int_t_sig, :3082
  a[(size_t)MS] = a[(size_t)MS] + dtyr * tot_M;                  // int_ocn_tot_M_sig
  a[(size_t)2 * MS] = a[(size_t)2 * MS] + dtyr * r[(size_t)MS];  // int_ocn_tot_M_sur_sig
  for (int l = 0; l < L; l++) {
    const size_t q0 = (size_t)(kSigHead + l) * MS, q1 = (size_t)(kSigHead + L + l) * MS, q2 = (size_t)(kSigHead + 2 * L + l) * MS;
    a[q0] = a[q0] + dtyr * r[q0] * rtot_M;
    a[q1] = a[q1] + dtyr * r[q1] * rtot_A;
    a[q2] = a[q2] + dtyr * r[q2] * g.rtot_A_ben;
  }
  for (int la = 0; la < g.LA; la++) {
    const size_t q = (size_t)(kSigHead + 3 * L + la) * MS;
    a[q] = a[q] + dtyr * r[q] * g.rtot_A_atm;
  }
}
// diag_biogem_timeslice (biogem.f90:2478-2579): thread = (member, wet cell).  The cell's carbonate system is solved from the
// cell's last [H+] (for the surface cell that is BgDev::carbH, the seed of step_biogem's own solve: the diagnostic feeds back
// into the next step as it does in the reference), then the window integrals grow by dtyr * field.  init != 0: sub_init_carb
// (biogem_data.f90:2336-2430) for the cells below the surface -- seed 10**(-7.8), RF0 of the cell -- no integrals.
__global__ void __launch_bounds__(128) k_bg_slice(const Dev v, const BgDev b, const SliceDev sl, const double dtyr, const int init) {
  using namespace bgk;
  using namespace lay;
  const int I = v.I, J = v.J, K = v.K, MS = v.MS;
  constexpr int L = NL, LS = NLS;
  const int m = blockIdx.x * 32 + threadIdx.x;
  const int w = blockIdx.y * blockDim.y + threadIdx.y;
  if (w >= sl.nwet3) return;
  const int c = sl.wet[w];
  const int k = c / (I * J) + 1;
  const bool surface = (k == K);
  if (init && surface) return;                    // the surface cell's initial solve is k_bg_step's (init_only)
  const size_t oc = (size_t)c * L * MS + m;
  double x[L + 1];
#pragma unroll
  for (int l = 1; l <= L; l++) x[l] = v.bg_ocn[oc + (size_t)(l - 1) * MS];
  double cc[N_CC];
  carbconst(b.Dmid[k], x[T], x[S], x[CA], x[MG], cc);
  // the constants the surface solve never reads (sub_calc_carbconst :205-262), for int_carbconst_timeslice
  double Tc = x[T], Sc = x[S];
  if (Tc < (kZeroC + 2.0)) Tc = kZeroC + 2.0;
  if (Tc > (kZeroC + 35.0)) Tc = kZeroC + 35.0;
  if (Sc < 26.0) Sc = 26.0;
  if (Sc > 43.0) Sc = 43.0;
  const double rT = 1.0 / Tc, T_ln = log(Tc), T_log = log10(Tc), Tr100 = Tc / 100.0, TC = Tc - kZeroC, rRT = 1.0 / (kR * Tc), P = b.Dmid[k] / 10.0;
  const double S_p05 = sqrt(Sc), S_p15 = Sc * S_p05;
  double t2s;
  {
    const double Ii = (Sc > kNS) ? 19.924 * Sc / (1000.0 - 1.005 * Sc) : kNS;
    const double I_p05 = sqrt(Ii), I_p15 = Ii * I_p05, I_p20 = Ii * Ii;
    double Cl = x[S] / 1.80655;
    if (Cl < kNS) Cl = kNS;
    const double ION = (Cl > kNS) ? 0.00147 + 0.03592 * Cl + 0.000068 * Cl * Cl : kNS;
    const double m2c = log(1 - 0.001005 * Sc);
    const double SO4tot = fS(0.02824, Sc), Ftot = fS(0.00007, Sc);
    const double lnkHSO4 = 141.328 - 4276.1 * rT - 23.093 * T_ln + (324.57 - 13856.0 * rT - 47.986 * T_ln) * I_p05 +
                           (-771.54 + 35474.0 * rT + 114.723 * T_ln) * Ii - 2698.0 * rT * I_p15 + 1776.0 * rT * I_p20;
    const double kHSO4f = exp(lnkHSO4 + m2c);
    const double f2t = log(1.0 + SO4tot / kHSO4f);
    const double kHFt = exp(div_by(1590.2, Tc, rT) - 12.641 + 1.525 * sqrt(ION) + m2c + f2t);
    t2s = -f2t + log(1.0 + SO4tot / kHSO4f + Ftot / kHFt);
  }
  const double kH2S = exp((225.838 - 13275.3 * rT - 34.6435 * T_ln + 0.3449 * S_p05 - 0.0274 * Sc) + t2s +
                          corr_p(TC, P, rRT, -1.107E+1, +9.000E-3, -9.420E-4, +2.890E+0, +5.400E-2));
  const double kNH4 = exp((-6285.33 * rT + 0.0001635 * Tc - 0.25444 + (0.46532 - 123.7184 * rT) * S_p05 + (-0.01992 + 3.17556 * rT) * Sc) +
                          corr_p(TC, P, rRT, -2.643E+0, +8.890E-1, -9.050E-3, -5.030E+0, +8.140E-2));
  const double kArg = exp(corr_p(TC, P, rRT, -4.596E+1, +5.304E-1, +0.000E+0, -1.176E+1, +3.692E-1)) *
                      exp10(-171.945 - 0.077993 * Tc + 2903.293 * rT + 71.595 * T_log + (-0.068393 + 0.0017276 * Tc + 88.135 * rT) * S_p05 -
                            0.10018 * Sc + 0.0059415 * S_p15);
  const double qO2 = exp(-173.9894 + 255.5907 * (100.0 * rT) + 146.4813 * log(Tr100) - 22.2040 * (Tr100) +
                         Sc * (-0.037362 + 0.016504 * (Tr100) - 0.0020564 * (Tr100 * Tr100)) - log(1.0E6) - log(0.20946));
  double *Hp = surface ? &b.carbH[((size_t)(c - (size_t)(K - 1) * I * J)) * MS + m] : &sl.carbH3[(size_t)c * MS + m];
  double *RFp = &sl.rf03[(size_t)c * MS + m];
  Carb cb;
  cb.H = init ? pow(10.0, -7.8) : *Hp;
  cb.RF0 = init ? 0.0 : *RFp;
  const bool rf = init || surface;
  double keepRF = cb.RF0;
  if (!solve_carb(x[DIC], x[ALK], x[CA], x[PO4], x[S], cc, cb, rf)) b.err[m] = 1;
  if (!rf) cb.RF0 = keepRF;
  *Hp = cb.H;
  if (rf) *RFp = cb.RF0;
  if (init) return;
  // sub_calc_carb_r13C / r14C in full (gem_carbchem.f90:677-780): ratios of DIC, CO2(aq), HCO3-, CO3--
  double iso[kSlIso];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const double mult = q ? 2.0 : 1.0, standard = q ? kStd14C : kStd13C, DICi = q ? x[DIC14] : x[DIC13];
    const double d = iso_delta(x[DIC], DICi, standard);
    const double e_bg = mult * (-0.1141 * (x[T] - kZeroC) + 10.78), e_dg = mult * (+0.0049 * (x[T] - kZeroC) - 1.31),
                 e_cg = mult * (-0.052 * (x[T] - kZeroC) + 7.22);
    const double e_cb = e_cg - e_bg / (1.0 + e_bg * 1.0E-3), e_db = e_dg - e_bg / (1.0 + e_bg * 1.0E-3);
    const double dHCO3 = (d * x[DIC] - (e_db * cb.co2 + e_cb * cb.co3)) /
                         ((1.0 + e_db * 1.0E-3) * cb.co2 + cb.hco3 + (1.0 + e_cb * 1.0E-3) * cb.co3);
    const double dCO2 = e_db + dHCO3 * (1.0 + e_db * 1.0E-3), dCO3 = e_cb + dHCO3 * (1.0 + e_cb * 1.0E-3);
    const double rCO2 = iso_fraction(dCO2, standard), rHCO3 = iso_fraction(dHCO3, standard), rCO3 = iso_fraction(dCO3, standard);
    iso[4 * q + 0] = (rCO2 * cb.co2 + rHCO3 * cb.hco3 + rCO3 * cb.co3) / x[DIC];
    iso[4 * q + 1] = rCO2; iso[4 * q + 2] = rHCO3; iso[4 * q + 3] = rCO3;
  }
  // ---- window integrals
  {
    double *a = sl.ocn + oc;
#pragma unroll
    for (int l = 1; l <= L; l++) a[(size_t)(l - 1) * MS] = a[(size_t)(l - 1) * MS] + dtyr * x[l];
    const size_t pc = (size_t)c * LS * MS + m;
#pragma unroll
    for (int ls = 0; ls < LS; ls++) sl.part[pc + (size_t)ls * MS] = sl.part[pc + (size_t)ls * MS] + dtyr * b.bio_part[pc + (size_t)ls * MS];
    // carb(ic_*): H, CO2, CO3, HCO3, fug_CO2, ohm_cal, ohm_arg, dCO3_cal, dCO3_arg, RF0   (gem_carbchem.f90:497-512)
    const double carb[kSlCarb] = {cb.H, cb.co2, cb.co3, cb.hco3, cb.co2 / cc[CC_QCO2], cb.ohm_cal, x[CA] * cb.co3 / kArg,
                                  cb.co3 - cc[CC_KCAL] * 1.0 / x[CA], cb.co3 - kArg * 1.0 / x[CA], cb.RF0};
    double *ac = sl.carb + (size_t)c * kSlCarb * MS + m;
#pragma unroll
    for (int q = 0; q < kSlCarb; q++) ac[(size_t)q * MS] = ac[(size_t)q * MS] + dtyr * carb[q];
    // carbconst(icc_*) in the oracle's order: k1 k2 k kB kW kSi kHF kHSO4 kP1 kP2 kP3 kH2S kNH4 kcal karg QCO2 QO2
    const double ccv[kSlCC] = {cc[CC_K1], cc[CC_K2], cc[CC_K], cc[CC_KB], cc[CC_KW], cc[CC_KSI], cc[CC_KHF], cc[CC_KHSO4], cc[CC_KP1],
                               cc[CC_KP2], cc[CC_KP3], kH2S, kNH4, cc[CC_KCAL], kArg, cc[CC_QCO2], qO2};
    double *acc = sl.cc + (size_t)c * kSlCC * MS + m;
#pragma unroll
    for (int q = 0; q < kSlCC; q++) acc[(size_t)q * MS] = acc[(size_t)q * MS] + dtyr * ccv[q];
    double *ai = sl.iso + (size_t)c * kSlIso * MS + m;
#pragma unroll
    for (int q = 0; q < kSlIso; q++) ai[(size_t)q * MS] = ai[(size_t)q * MS] + dtyr * iso[q];
    if (w == 0) sl.t[m] = sl.t[m] + dtyr;
  }
}
int launch_bg_slice(const Dev &v, const BgDev &b, const SliceDev &sl, double dtyr, int init, cudaStream_t s) {
  if (sl.nwet3 <= 0) return 0;
  k_bg_slice<<<dim3(v.MS / 32, (sl.nwet3 + 3) / 4), dim3(32, 4), 0, s>>>(v, b, sl, dtyr, init);
  return 1;
}

// ---- extended time-series integrals ("bg_sig2"; diag_biogem_timeseries, biogem.f90:2870-2883, 2926-2964, 3058-3062) -------------
// bio_settle(:,i,j,n_k) of the sweep that follows on the same stream: the surface layer is a source only, so what settles through
// its base is the layer's particulate field as the sweep reads it (rescaling of the last coupling applied if still pending, 14C
// decayed: sub_box_remin_part, biogem_box.f90:2412-2875, k = n_k; the frac2 arrays pass unweighted, the rest times the cell mass)
__global__ void __launch_bounds__(128) k_bg_settle_sur(const Dev v, const BgDev b, const int pend) {
  using namespace lay;
  const int I = v.I, J = v.J, K = v.K, MS = v.MS;
  constexpr int LS = NLS;
  const int m = blockIdx.x * 32 + threadIdx.x;
  const int n = blockIdx.y * blockDim.y + threadIdx.y;
  if (n >= v.nwet) return;
  const int c2d = v.bgcols[n];
  const size_t c = (size_t)(K - 1) * I * J + c2d;
  double old[LS + 1];
#pragma unroll
  for (int ls = 1; ls <= LS; ls++) old[ls] = b.bio_part[(c * LS + (ls - 1)) * MS + m];
  if (pend) {
    const double pscale = b.pscale[m];
#pragma unroll
    for (int ls = 1; ls <= LS; ls++) old[ls] = pscale * (old[ls] + 0.0);
  }
  if (fabs(b.lam_sed[POC14]) > bgk::kNS) { old[POC14] = b.fd_sed[POC14] * old[POC14]; old[CACO314] = b.fd_sed[POC14] * old[CACO314]; }
  double part_tot = 0.0;
  part_tot = part_tot + old[POC];
  part_tot = part_tot + old[CACO3];
  const double Mk = v.bg_M[c * MS + m];
#pragma unroll
  for (int ls = 1; ls <= LS; ls++)
    b.settle_sur[((size_t)c2d * LS + (ls - 1)) * MS + m] = (part_tot > bgk::kNS) ? ((ls >= POCF2) ? old[ls] : Mk * old[ls]) : 0.0;
}
int launch_bg_settle_sur(const Dev &v, const BgDev &b, int pend, cudaStream_t s) {
  if (!b.settle_sur) return 0;
  k_bg_settle_sur<<<dim3(v.MS / 32, (v.nwet + 3) / 4), dim3(32, 4), 0, s>>>(v, b, pend);
  return 1;
}
// On the CALLER's stream, at the time diag_biogem_timeseries is called: sea-ice thickness (phys_ocnatm(ipoa_seaice_th) =
// hght_sic of this koverall iteration) and the extrema of the overturning stream functions (sub_calc_psi, biogem_box.f90:3796-3852,
// from the velocities biogem_climate copied) -- the next cycle's sea-ice and momentum steps may overwrite both before the BIOGEM
// stream gets to the sums.  Block = 32 members x 32 rows j; each thread integrates its rows upward in the reference's order.
__global__ void __launch_bounds__(1024) k_bg_sig2_stage(const Dev v, const SigDev g, const double *__restrict__ dz, const double *__restrict__ cv,
                                                        const double dphi) {
  __shared__ double red[4][32][33];
  const int I = v.I, J = v.J, K = v.K, MS = v.MS;
  const int lane = threadIdx.x, row = threadIdx.y;
  const int m = blockIdx.x * 32 + lane;
  double omin = 0.0, omax = 0.0, omina = 0.0, omaxa = 0.0;
  for (int j = 1 + row; j <= J - 1; j += 32) {
    double opsi = 0.0, opsia = 0.0;
    const bool atl = j >= g.jsf + 1;
    const int ia0 = g.ias[j], ia1 = g.iaf[j];
    for (int k = 1; k <= K - 1; k++) {
      double ou = 0.0, oua = 0.0;
      const double *u2 = v.u + ((cell3(I, J, 1, j, k) * 3 + 1) * (size_t)MS + m);
      for (int i = 1; i <= I; i++) {
        const double t = cv[j] * u2[(size_t)(i - 1) * 3 * MS] * dphi;
        ou = ou + t;
        if (atl && i >= ia0 && i <= ia1) oua = oua + t;
      }
      opsi = opsi - dz[k] * ou;
      omin = fmin(omin, opsi); omax = fmax(omax, opsi);
      if (atl) {
        opsia = opsia - dz[k] * oua;
        if (k <= K / 2) { omina = fmin(omina, opsia); omaxa = fmax(omaxa, opsia); }
      }
    }
  }
  red[0][row][lane] = omin; red[1][row][lane] = omax; red[2][row][lane] = omina; red[3][row][lane] = omaxa;
  __syncthreads();
  if (row < 4) {
    double r = red[row][0][lane];
    for (int q = 1; q < 32; q++) r = (row & 1) ? fmax(r, red[row][q][lane]) : fmin(r, red[row][q][lane]);
    g.opsi_stage[(size_t)row * MS + m] = r;
  }
  // sea-ice thickness: varice(1,:,:)
  const size_t n = (size_t)I * J * MS;
  for (size_t q = (size_t)blockIdx.x * 1024 + row * 32 + lane; q < n; q += (size_t)gridDim.x * 1024) g.th_stage[q] = v.varice[q];
}
// sums of one step: block = (32-member tile, quantity), 8 warps stride over the cells (partial sums added in warp order)
//   0 SUM(A*seaice)  1 SUM(seaice)  2 SUM(th*A*seaice)  3 SUM over land of A*sfcatm1(T)
//   4 + ls  SUM(bio_settle(ls,:,:,n_k))   4 + LS + la  SUM(conv_yr_s*A*sfxatm1(la)) over the ocean   4 + LS + LA + la  SUM(focnatm(la))
__global__ void __launch_bounds__(256) k_bg_sig2_sums(const Dev v, const BgDev b, const SigDev g) {
  __shared__ double part[8][32];
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int I = v.I, J = v.J, K = v.K, MS = v.MS, ij = I * J, LS = g.LS, LA = g.LA;
  const int m = blockIdx.x * 32 + lane, q = blockIdx.y;
  double s = 0.0;
  for (int c2 = warp; c2 < ij; c2 += 8) {
    const bool wet = g.kbot[c2] < K;
    const size_t c = (size_t)c2 * MS + m;
    if (q == 0) { if (wet) s = s + g.A[c2] * b.seaice[c]; }
    else if (q == 1) s = s + b.seaice[c];
    else if (q == 2) { if (wet) s = s + g.th_stage[c] * g.A[c2] * b.seaice[c]; }
    else if (q == 3) { if (!wet) s = s + g.A[c2] * b.sfcatm1[c]; }
    else if (q < 4 + LS) { if (wet) s = s + b.settle_sur[((size_t)c2 * LS + (q - 4)) * MS + m]; }
    else if (q < 4 + LS + LA) { if (wet) s = s + bgk::kYrS * g.A[c2] * b.sfxatm1[((size_t)(q - 4 - LS) * ij + c2) * MS + m]; }
    else { if (wet) s = s + b.focnatm[((size_t)(q - 4 - LS - LA) * ij + c2) * MS + m]; }
  }
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
    double t = part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; w++) t = t + part[w][lane];
    g.raw2[(size_t)q * MS + m] = t;
  }
}
__global__ void k_bg_sig2_acc(const Dev v, const SigDev g, const double dtyr) {
  const int MS = v.MS, LS = g.LS, LA = g.LA;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= MS) return;
  const double *r = g.raw2 + m;
  double *a = g.acc2 + m;
  a[0] = a[0] + dtyr * r[0];                                                          // int_misc_seaice_sig
  if (r[(size_t)MS] > kBgNullSmall) a[(size_t)MS] = a[(size_t)MS] + dtyr * r[(size_t)2 * MS] / r[0];   // ..._th
  a[(size_t)2 * MS] = a[(size_t)2 * MS] + dtyr * r[(size_t)2 * MS];                  // ..._vol
  for (int q = 0; q < 4; q++) a[(size_t)(3 + q) * MS] = a[(size_t)(3 + q) * MS] + dtyr * g.opsi_stage[(size_t)q * MS + m];
  if (g.land_A > kBgNullSmall) a[(size_t)7 * MS] = a[(size_t)7 * MS] + dtyr * r[(size_t)3 * MS] / g.land_A; else a[(size_t)7 * MS] = 0.0;
  for (int ls = 0; ls < LS; ls++) a[(size_t)(kSig2Head + ls) * MS] = a[(size_t)(kSig2Head + ls) * MS] + r[(size_t)(4 + ls) * MS];
  for (int la = 2; la < LA; la++) {
    const size_t qa = (size_t)(kSig2Head + LS + la) * MS, qd = (size_t)(kSig2Head + LS + LA + la) * MS;
    a[qa] = a[qa] + dtyr * r[(size_t)(4 + LS + la) * MS];
    a[qd] = a[qd] + dtyr * r[(size_t)(4 + LS + LA + la) * MS];
  }
}
int launch_bg_sig2_stage(const Dev &v, const SigDev &g, const double *dz, const double *cv, double dphi, cudaStream_t s) {
  k_bg_sig2_stage<<<v.MS / 32, dim3(32, 32), 0, s>>>(v, g, dz, cv, dphi);
  return 1;
}
int launch_bg_sig2(const Dev &v, const BgDev &b, const SigDev &g, double dtyr, cudaStream_t s) {
  k_bg_sig2_sums<<<dim3(v.MS / 32, 4 + g.LS + 2 * g.LA), dim3(32, 8), 0, s>>>(v, b, g);
  k_bg_sig2_acc<<<(v.MS + 127) / 128, 128, 0, s>>>(v, g, dtyr);
  return 2;
}

int launch_bg_sig(const Dev &v, const BgDev &b, const SigDev &g, double dtyr, cudaStream_t s) {
  const int nq = kSigHead + 3 * v.L + g.LA;
  k_bg_sig_sums<<<dim3(v.MS / 32, nq), dim3(32, 8), 0, s>>>(v, b, g);
  k_bg_sig_acc<<<(v.MS + 127) / 128, 128, 0, s>>>(v, g, dtyr);
  return 2;
}

}  // namespace cg
