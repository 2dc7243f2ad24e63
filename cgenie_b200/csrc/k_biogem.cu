// k_biogem.cu -- BIOGEM kernels on the tracer hot path, sm_100a, compiled -fmad=false.
//
// biogem_tracercoupling (src/biogem/biogem.f90:1885-2077): salinity-normalised transfer between the
// GOLDSTEIN tracer array ts and BIOGEM's ocn, with per-member global reductions (mean salinity old /
// new, old and new inventory of every biogeochemical tracer).  The reference sums each wet column
// over k (ascending) and then the column partials in its vocn order (i outer, j inner); the kernels
// keep exactly that order -- thread (member, column) for the partials, thread (member, quantity) for
// the ordered sum over columns -- so every total is bit-identical to the sequential code.
#include "cg_device.cuh"
#include "cg_host.hpp"

namespace cg {

constexpr double kBgZeroC = 273.15;          // gem_cmn.f90:690
constexpr double kBgNullSmall = 0.999999e-19; // gem_cmn.f90:719

// quantity slots of the reduction scratch
//   0            : sum_k ocn(S)*V * rtot_V              (old mean salinity)        [phase A]
//   1            : sum_k (ts(S)+saln0+docn(S))*V*rtot_V (new mean salinity)        [phase A]
//   2 .. L-1     : sum_k ocn(l)*M, l = 3..L             (old inventories)          [phase A]
//   L .. 2L-3    : sum_k loc_vocn(l)*M, l = 3..L        (salinity-adjusted new)    [phase B]
__global__ void __launch_bounds__(128) k_tc_partial(const Dev v, const int phase) {
  const int I = v.I, J = v.J, K = v.K, L = v.L, MS = v.MS;
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
  if (phase == 0) {
    const double saln0 = v.p.saln0[m];
    double a = 0.0, b = 0.0;
    for (int k = k1c; k <= K; k++) {
      const double V = v.bg_V[v0 + (size_t)(k - 1) * vK];
      a = a + v.bg_ocn[o0 + (size_t)(k - 1) * sK + MS] * V;
      b = b + (v.ts_cur[o0 + (size_t)(k - 1) * sK + MS] + saln0 + v.bg_vdocn[o0 + (size_t)(k - 1) * sK + MS]) * V;
    }
    part[0] = a * v.bg_rtot_V;
    part[nq] = b * v.bg_rtot_V;
    for (int l = 2; l < L; l++) {
      double s = 0.0;
      for (int k = k1c; k <= K; k++) s = s + v.bg_ocn[o0 + (size_t)(k - 1) * sK + (size_t)l * MS] * v.bg_M[p0 + (size_t)(k - 1) * pK];
      part[(size_t)l * nq] = s;
    }
  } else {
    const double rmean = 1.0 / v.bg_tot[m];  // loc_ocn_rmean_S_OLD
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

// ordered sum over the wet columns: thread = (member, quantity)
__global__ void __launch_bounds__(128) k_tc_sum(const Dev v, const int q0, const int q1) {
  const int MS = v.MS;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = q0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || q >= q1) return;
  const double *part = v.bg_part + (size_t)q * v.nwet * MS + m;
  double s = 0.0;
  for (int n = 0; n < v.nwet; n++) s = s + part[(size_t)n * MS];
  v.bg_tot[(size_t)q * MS + m] = s;
}

// (2)+(3) of biogem_tracercoupling: new T,S, rescaled tracers, cell masses, ts <- normalised ocn
__global__ void __launch_bounds__(128) k_tc_apply(const Dev v) {
  const int I = v.I, J = v.J, K = v.K, L = v.L, MS = v.MS;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= MS || n >= v.nwet) return;
  const int c2 = v.bgcols[n];
  const int i = c2 % I + 1, j = c2 / I + 1;
  const int k1c = CG_K1(v, i, j);
  const size_t sC = (size_t)L * MS, sK = (size_t)I * J * sC;
  const size_t o0 = cell3(I, J, i, j, 1) * sC + m;
  const size_t p0 = cell3(I, J, i, j, 1) * MS + m, pK = (size_t)I * J * MS;
  const double saln0 = v.p.saln0[m];
  const double mean_S_OLD = v.bg_tot[m], mean_S_NEW = v.bg_tot[MS + m];
  const double rmean_S_OLD = 1.0 / mean_S_OLD;
  const double Sratio = mean_S_NEW / mean_S_OLD, rSratio = 1.0 / Sratio;
  for (int k = K; k >= k1c; k--) {
    const size_t o = o0 + (size_t)(k - 1) * sK;
    const double Sold = v.bg_ocn[o + MS];
    const double Tn = v.ts_cur[o] + kBgZeroC + v.bg_vdocn[o];
    const double Sn = v.ts_cur[o + MS] + saln0 + v.bg_vdocn[o + MS];
    v.bg_ocn[o] = Tn;
    v.bg_ocn[o + MS] = Sn;
    v.ts_cur[o] = Tn - kBgZeroC;
    v.ts_cur[o + MS] = Sn - saln0;
    for (int l = 2; l < L; l++) {
      const double told = v.bg_tot[(size_t)l * MS + m], tnew = v.bg_tot[(size_t)(L - 2 + l) * MS + m];
      const double rtnew = (fabs(tnew) < kBgNullSmall) ? 0.0 : 1.0 / tnew;
      const double lv = v.ts_cur[o + (size_t)l * MS] * Sold * rmean_S_OLD;
      double x = (told * rtnew) * lv + v.bg_vdocn[o + (size_t)l * MS];
      x = Sratio * x;
      v.bg_ocn[o + (size_t)l * MS] = x;
      v.ts_cur[o + (size_t)l * MS] = (mean_S_NEW / Sn) * x;
    }
    v.bg_M[p0 + (size_t)(k - 1) * pK] = rSratio * v.bg_M[p0 + (size_t)(k - 1) * pK];
    v.bg_rM[p0 + (size_t)(k - 1) * pK] = Sratio * v.bg_rM[p0 + (size_t)(k - 1) * pK];
  }
}

// biogem_climate's only state change on this path: reset the convection counter (biogem.f90:2238)
__global__ void k_bg_reset_cost(const Dev v) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < (size_t)v.I * v.J * v.MS) v.cost[q] = 0.0;
}

int launch_tracercoupling(const Dev &v, cudaStream_t s) {
  const dim3 b(32, 4);
  const dim3 gc(v.MS / 32, (v.nwet + 3) / 4);
  const int L = v.L;
  k_tc_partial<<<gc, b, 0, s>>>(v, 0);
  k_tc_sum<<<dim3(v.MS / 32, (L + 3) / 4), b, 0, s>>>(v, 0, L);
  if (L > 2) {
    k_tc_partial<<<gc, b, 0, s>>>(v, 1);
    k_tc_sum<<<dim3(v.MS / 32, (L - 2 + 3) / 4), b, 0, s>>>(v, L, 2 * L - 2);
  }
  k_tc_apply<<<gc, b, 0, s>>>(v);
  return L > 2 ? 5 : 3;
}
int launch_bg_reset_cost(const Dev &v, cudaStream_t s) {
  const size_t n = (size_t)v.I * v.J * v.MS;
  k_bg_reset_cost<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(v);
  return 1;
}

}  // namespace cg
