// col_host.cpp -- TEST INFRASTRUCTURE: compiles the per-thread body of the fused tracer column kernel
// (cgenie_b200/csrc/k_tracer_col.cuh) for the HOST so that its logic can be checked against the oracle without a GPU
// (tests/test_col_body_host.py).  Never linked into the product library.
#include <cstring>
#include <vector>
#include <cmath>
#include "k_tracer_col.cuh"

using namespace cg;

extern "C" int col_host_step(int MS, const unsigned char *k1, const int *cols, int ncol, const double *ts_cur,
                             double *ts_new, const double *tsflux, double *sst, double *rho, const double *u, double *cost,
                             const double *diff1, const double *diff2, const double *ec /* [4][MS] */,
                             const double *jm /* rc, rc2, cv, cv2, dsv, rdsv, rds: 7 x (J+2) */,
                             const double *km /* dz, dza, rdz, rdza, ssmax: 5 x (K+2) */, double dphi, double rdphi,
                             double dt, int mix) {
  constexpr int I = 36, J = 36, K = 16, L = 16;
  if (MS != 32) return 1;
  static GridC g;
  std::memset(&g, 0, sizeof(g));
  g.I = I; g.J = J; g.K = K; g.L = L; g.M = MS; g.MS = MS;
  g.dphi = dphi; g.rdphi = rdphi; g.dt = dt;
  for (int j = 0; j < J + 2; j++) {
    g.rc[j] = jm[0 * (J + 2) + j]; g.rc2[j] = jm[1 * (J + 2) + j]; g.cv[j] = jm[2 * (J + 2) + j];
    g.cv2[j] = jm[3 * (J + 2) + j]; g.dsv[j] = jm[4 * (J + 2) + j]; g.rdsv[j] = jm[5 * (J + 2) + j];
    g.rds[j] = jm[6 * (J + 2) + j];
  }
  for (int k = 0; k < K + 2; k++) {
    g.dz[k] = km[0 * (K + 2) + k]; g.dza[k] = km[1 * (K + 2) + k]; g.rdz[k] = km[2 * (K + 2) + k];
    g.rdza[k] = km[3 * (K + 2) + k]; g.ssmax[k] = km[4 * (K + 2) + k];
  }
  Dev v;
  std::memset(&v, 0, sizeof(v));
  v.I = I; v.J = J; v.K = K; v.L = L; v.M = MS; v.MS = MS;
  v.k1 = k1;
  v.ts_cur = const_cast<double *>(ts_cur); v.ts_new = ts_new; v.tsflux = const_cast<double *>(tsflux);
  v.sst = sst; v.rho = rho; v.u = const_cast<double *>(u); v.cost = cost;
  v.p.diff1 = diff1; v.p.diff2 = diff2;
  v.p.ec1 = ec; v.p.ec2 = ec + MS; v.p.ec3 = ec + 2 * MS; v.p.ec4 = ec + 3 * MS;
  // the whole "block" (32 members of one column) shares one staging area; bulk copies are emulated element-wise
  std::vector<double> sm((size_t)(ColRows<L>::rows > SplitRows<L>::rows ? ColRows<L>::rows : SplitRows<L>::rows) * 32);
  unsigned long long bar[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (mix == 5) {   // warp-tile form (one warp = 32 members of a column, lane-parallel row copies) + flag + co
    std::vector<unsigned> comask((size_t)I * J * MS, 7u);
    v.comask = comask.data();
    v.co_skip_stable = 1;
    v.co_pairwise = 1;
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) {
        ColStage st{sm.data(), bar, m};
        tstep_column_w<I, J, K, L, 32>(v, g, cols[n], (unsigned)m, st, nullptr);
      }
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) co_column<I, J, K, L, 32>(v, g, cols[n], (unsigned)m);
    return 0;
  }
  if (mix >= 7 && mix <= 9) {   // the convective adjustment alone on the caller's ts_new / rho (7: the reference's walk on thread-private
    v.co_skip_stable = 0;       // arrays, 8: lockstep passes merging every unstable run at once, 9: lockstep passes in the walk's order)
    v.co_pairwise = 0;
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) {
        if (mix == 9) co_column<I, J, K, L, 32, false, 0, 2>(v, g, cols[n], (unsigned)m);
        else if (mix == 8) co_column<I, J, K, L, 32, false, 0, 1>(v, g, cols[n], (unsigned)m);
        else co_column<I, J, K, L, 32, false, 0, 0>(v, g, cols[n], (unsigned)m);
      }
    return 0;
  }
  if (mix == 6) {   // round-1 flux kernel + stability flag, then co with the decisions in lockstep form + region-wise averaging
    std::vector<unsigned> comask((size_t)I * J * MS, 7u);
    v.comask = comask.data();
    v.co_skip_stable = 1;
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) {
        ColStage st{sm.data(), bar, m};
        tstep_column<I, J, K, L, 32, 32, false>(v, g, cols[n], (unsigned)m, st);
      }
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) co_column<I, J, K, L, 32, false, 0, 2>(v, g, cols[n], (unsigned)m);
    return 0;
  }
  if (mix == 4) {   // round-1 flux kernel + stability flag, decisions only, then one "thread" per passive tracer (k_co_passive)
    std::vector<unsigned> comask((size_t)I * J * MS, 7u);
    v.comask = comask.data();
    v.co_skip_stable = 1;
    v.co_pairwise = 2;
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) {
        ColStage st{sm.data(), bar, m};
        tstep_column<I, J, K, L, 32, 32, false>(v, g, cols[n], (unsigned)m, st);
      }
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) co_column<I, J, K, L, 32>(v, g, cols[n], (unsigned)m);
    for (int n = 0; n < ncol; n++)
      for (int l = 2; l < L; l++)
        for (int m = 0; m < MS; m++) co_passive_one<I, J, K, L, 32>(v, g, cols[n], (unsigned)m, l);
    return 0;
  }
  if (mix == 3) {   // pipelined form (coefficients one level ahead) + stability flag, then co on the flagged member-columns
    std::vector<double> sm2((size_t)ColRows2<L>::rows * 32);
    std::vector<unsigned> comask((size_t)I * J * MS, 7u);
    v.comask = comask.data();
    v.co_skip_stable = 1;
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) {
        ColStage st{sm2.data(), bar, m};
        tstep_column2<I, J, K, L, 32, 32>(v, g, cols[n], (unsigned)m, st);
      }
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) co_column<I, J, K, L, 32>(v, g, cols[n], (unsigned)m);
    return 0;
  }
  if (mix == 2) {   // split form (two threads per member-column; here one caller plays both halves), then co
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) {
        ColStage st{sm.data(), bar, m};
        tstep_column_split<I, J, K, L, 32, 32, true>(v, g, cols[n], m, st);
      }
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) co_column<I, J, K, L, 32>(v, g, cols[n], (unsigned)m);
    return 0;
  }
  if (mix) {   // T,S pre-pass + decisions, then the passive tracers mixed on write (k_ts_pre + k_tstep_col<PV>)
    std::vector<unsigned> comask((size_t)I * J * MS, 0u);
    v.comask = comask.data();
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) ts_pre_column<I, J, K, L, 32>(v, g, cols[n], (unsigned)m);
    for (int n = 0; n < ncol; n++)
      for (int m = 0; m < MS; m++) {
        ColStage st{sm.data(), bar, m};
        tstep_column<I, J, K, L, 32, 32, true>(v, g, cols[n], (unsigned)m, st);
      }
    return 0;
  }
  for (int n = 0; n < ncol; n++)
    for (int m = 0; m < MS; m++) {
      ColStage st{sm.data(), bar, m};
      tstep_column<I, J, K, L, 32, 32, false>(v, g, cols[n], (unsigned)m, st);
    }
  for (int n = 0; n < ncol; n++)
    for (int m = 0; m < MS; m++) co_column<I, J, K, L, 32>(v, g, cols[n], (unsigned)m);
  return 0;
}
