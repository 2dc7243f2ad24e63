// cg_capi.cu -- the C-ABI of include/cgenie_b200.h: handle, device state, step scheduling,
// state movement.  No torch types, no oracle code.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges per entry point (the reference brackets the same calls with ITT tasks,
                               // src/genie.f90:13-15, 118-121 under INTEL_PROFILE; src/utils/itt_profile.f90)

#include "../../include/cgenie_b200.h"
#include "cg_device.cuh"
#include "cg_host.hpp"
#include "cg_biogem.hpp"

namespace cg {
// kernels' launchers (k_physics.cu, k_tracer_*.cu, k_biogem.cu)
void upload_grid_physics(const GridC &, cudaStream_t);
void upload_grid_tracer_strict(const GridC &, cudaStream_t);
void upload_grid_tracer_fast(const GridC &, cudaStream_t);
int launch_tstepo_flux_strict(const Dev &, cudaStream_t);
int launch_tstepo_flux_fast(const Dev &, cudaStream_t);
int launch_co_strict(const Dev &, cudaStream_t);
int launch_co_fast(const Dev &, cudaStream_t);
void upload_grid_tracer_col(const GridC &, cudaStream_t);
int launch_tstep_col(const Dev &, cudaStream_t);
bool tstep_col_supported(const Dev &);
void launch_step_begin(const Dev &, cudaStream_t);
void launch_usnap(const Dev &, cudaStream_t);
void launch_hosing(const Dev &, cudaStream_t);
int launch_surflux(const Dev &, double *meantemp, bool need_mean, cudaStream_t);
int launch_embm(const Dev &, int nsteps, cudaStream_t);
int launch_seaice(const Dev &, cudaStream_t);
int launch_gold_pre(const Dev &, cudaStream_t);
int launch_mld_pre(const Dev &, cudaStream_t);
int launch_mld_save(const Dev &, cudaStream_t);
int launch_mld_kt(const Dev &, cudaStream_t);
int launch_sst(const Dev &, cudaStream_t);
int launch_momentum(const Dev &, int fast, const double *bf, const double *bb, const double *rd, const double *bk, cudaStream_t);
int launch_velc1(const Dev &, cudaStream_t);
int launch_velc2(const Dev &, cudaStream_t);
void launch_global_means(const Dev &, double *out, cudaStream_t);
int launch_tracercoupling(const Dev &, cudaStream_t);
int launch_bg_reset_cost(const Dev &, cudaStream_t);
int launch_bg_sig(const Dev &, const BgDev &, const SigDev &, double dtyr, cudaStream_t);
int launch_bg_slice(const Dev &, const BgDev &, const SliceDev &, double dtyr, int init, cudaStream_t);
int launch_cpl_ocnsed(double *sum, const double *src, size_t n, double a, double b, int mode, cudaStream_t);
int launch_bg_step(const Dev &, const BgDev &, int init_only, int fuse, cudaStream_t);
int launch_tc_sums_first(const Dev &, cudaStream_t);
int launch_tc_sums_old(const Dev &, cudaStream_t);
int launch_tc_sums_new(const Dev &, cudaStream_t);
int launch_bg_surf(const Dev &, const BgDev &, cudaStream_t);
int launch_bg_sweep(const Dev &, const BgDev &, cudaStream_t, int fuse = 0);
int launch_bg_packets(const Dev &, const BgDev &, int pend, cudaStream_t);
int launch_bg_cell(const Dev &, const BgDev &, cudaStream_t);
int launch_bg_part_scale(const Dev &, const BgDev &, cudaStream_t);
int launch_tc_apply_only(const Dev &, cudaStream_t);
bool bg_layout_ok(const BgDev &, int L);
int launch_bg_climate(const Dev &, const BgDev &, cudaStream_t);
int launch_bg_sig2_stage(const Dev &, const SigDev &, const double *dz, const double *cv, double dphi, cudaStream_t);
int launch_bg_sig2(const Dev &, const BgDev &, const SigDev &, double dtyr, cudaStream_t);
int launch_bg_stage_seaice(const Dev &, const BgDev &, cudaStream_t);
int launch_bg_atchem(const Dev &, const BgDev &, double atm_totV, cudaStream_t);
void launch_health(const Dev &, int *flags, cudaStream_t);

// member <-> Fortran-shaped staging (gather/scatter one member of a [..][m] field)
struct FieldDesc {
  double *d = nullptr;
  int nd = 0;
  int dims[4] = {1, 1, 1, 1};
  long long strides[4] = {0, 0, 0, 0};  // device element stride (before *MS) of each Fortran dim
  long long count() const { long long n = 1; for (int q = 0; q < nd; q++) n *= dims[q]; return n; }
};
__global__ void k_gather_member(const double *__restrict__ src, double *__restrict__ dst, int m, int MS, int nd, int d0, int d1,
                                int d2, int d3, long long s0, long long s1, long long s2, long long s3, long long n) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  long long r = q;
  const int i0 = r % d0; r /= d0;
  const int i1 = r % d1; r /= d1;
  const int i2 = r % d2; r /= d2;
  const int i3 = (int)r;
  (void)nd; (void)d3;
  dst[q] = src[(i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3) * MS + m];
}
__global__ void k_scatter_member(double *__restrict__ dstf, const double *__restrict__ src, int m, int MS, int nd, int d0, int d1,
                                 int d2, int d3, long long s0, long long s1, long long s2, long long s3, long long n) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  long long r = q;
  const int i0 = r % d0; r /= d0;
  const int i1 = r % d1; r /= d1;
  const int i2 = r % d2; r /= d2;
  const int i3 = (int)r;
  (void)nd; (void)d3;
  dstf[(i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3) * MS + m] = src[q];
}
}  // namespace cg

using namespace cg;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CUDA_OK(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(CG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct ProfFam { double ms = 0; long long n = 0; };

struct cg_handle {
  int device = 0, M = 0, MS = 0;
  bool tracer_only = false, initialised = false;
  Params base;
  Grid g;
  Islands isl;
  WindFiles w;
  std::vector<Params> mp;
  std::vector<MemberConsts> mc;  // one per member (big factor arrays kept on group leaders only)
  std::vector<int> baro_group;
  int nbaro = 0;
  GridC gc;
  Dev dv;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;                 // momentum branch of the graph-captured ocean cycle
  cudaStream_t stream4 = nullptr;                 // BIOGEM / ATCHEM block next to the head of the following cycle (low priority)
  cudaEvent_t evT = nullptr, evBG = nullptr, evBGtail = nullptr;   // evBG: ts is ready for tstepo; evBGtail: the whole block (ATCHEM) is done
  cudaEvent_t evAtchem = nullptr;                 // the ATCHEM step has consumed the staged air temperature / humidity (tq_stage)
  bool atchem_ev_pending = false;
  bool bg_tail_pending = false;
  bool bg_ahead = false;                          // the step kernel of the next BIOGEM block is already in stream4
  bool tc_old_ready = false;                      // ... and the "old" half of its tracer-coupling sums in stream5 (same validity as bg_ahead)
  // per-module path: the same half issued by cg_atchem_step for the next cg_biogem_tracercoupling call (valid until the host
  // writes state); tc_old_pending = its kernels may still be running on stream5 (they own the reduction scratch)
  bool tc_spec_valid = false, tc_old_pending = false;
  // per-module path, lazy cycle: module calls without host arrays are only noted (1 = surflux, 1 + n = n EMBM steps,
  // kocn_loop + 2 = sea ice); if step_goldstein completes the canonical cycle, the cycle is issued as cg_run issues it
  // (two graph replays); anything else replays the noted calls one by one first (flush_lazy)
  int lazy_stage = 0;
  cudaEvent_t evTcOld = nullptr;
  cudaStream_t stream5 = nullptr;                 // tracer-coupling sums next to the BIOGEM step kernel
  cudaEvent_t evFork5 = nullptr, evJoin5 = nullptr;
  bool bg_overlap = true, bg_pending = false, bg_staged = false;
  bool eager = true;                              // per-module entry points: momentum at surflux time, BIOGEM calls on stream4
  bool mom_pending = false, mom_ready = false;    // eager momentum: still running on stream2 / result valid and unconsumed
  bool usnap_valid = false;                       // the surface-velocity snapshot for this cycle's sea-ice step was taken before an early momentum step
  double *u1_snap = nullptr;                      // u1 as it was before an eager momentum step (the one state it reads AND rewrites)
  cudaStream_t stream3 = nullptr;                 // baroclinic shear next to the barotropic solve
  cudaEvent_t evFork = nullptr, evJoin = nullptr, evFork3 = nullptr, evJoin3 = nullptr;
  bool bg_fuse = false;                           // cg_run: tracer coupling fused into the BIOGEM step kernel (slower on B200)
  bool fork_momentum = true;                      // CG_FORK=0 keeps the captured cycle on one stream
  bool forked = false;                            // inside enqueue_cycle with the momentum branch on stream2
  std::vector<void *> allocs;
  std::map<std::string, FieldDesc> fields;
  std::map<std::string, std::vector<double>> hconst;   // host constants for cg_get_const (member 0 / shared)
  std::map<std::string, std::vector<int>> hiconst;
  double *stage = nullptr;
  size_t stage_n = 0;
  int *d_wet3 = nullptr;                          // wet cells in ascending cell order (cg_sync_all_wet_*)
  int n_wet3 = 0;
  double *wet_stage = nullptr;
  size_t wet_stage_n = 0;
  // double-buffered state exchange (cg_exchange_*): per field one device staging buffer per direction, a copy stream and the
  // events that order it against the compute stream
  struct Xchg { double *up = nullptr, *down = nullptr; size_t n = 0; bool wet = false; int inner = 1; cudaEvent_t up_done = nullptr, down_ready = nullptr; bool up_pending = false; };
  std::map<std::string, Xchg> xchg;
  cudaStream_t xstream = nullptr;
  double *d_meantemp = nullptr, *d_means = nullptr;
  double *d_bf = nullptr, *d_bb = nullptr, *d_rd = nullptr;  // pivot-major barotropic factors (fast solve)
  double *d_bk = nullptr;                                    // block slabs of the blocked solve (k_baro_blk)
  int *d_flags = nullptr;
  bool need_mean = false;
  long long launches = 0;
  long long koverall = 0;
  int istep_ocn = 0, istep_atm = 0, istep_sic = 0;
  bool bg_split = false, bg_surf_issued = false;   // pipelined BIOGEM block, split form (see bg_issue_surf)
  // packets / cells form of the sweep (k_bg_step PART 3 + k_bg_cell, CG_BG_PD=0: off): pd_pending = the packets half of a step has
  // run and its cell half (with the coupling update) is owed; part_pending = bio_part lacks the last coupling's rescaling
  bool bg_pd = true, pd_pending = false, part_pending = false;
  std::vector<int> bg_colidx;                      // (i,j) -> index in the BIOGEM column order, -1 on land
  // per-module entry points: the surface part of the NEXT step_biogem is issued speculatively at the end of cg_atchem_step
  // (it has no side effect outside bgd.surf); cg_biogem_step uses it if it is called with the predicted clock and nothing
  // touched the state in between, and drops it otherwise
  bool spec_valid = false;
  long long spec_clock = 0, bg_last_clock = -1;
  int bg_stage = 0;                               // 1 step, 2 tracercoupling, 3 climate seen in the canonical order
  int variant = 0;  // 0 strict, 1 fast (cooperative flux kernel + separate convection), 2 fused column kernel (falls back to 1)
  bool use_graphs = true;
  cudaGraphExec_t graph[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  cudaGraphExec_t graph2[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};  // [variant][parity]
  long long graph_launches[3] = {0, 0, 0};   // launches inside one replayed cycle, per variant
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool profile = false;
  std::map<std::string, ProfFam> prof;
  int io_member = 0;
  // BIOGEM / ATCHEM
  BgConfig bg;
  BgDev bgd;
  std::vector<double> bg_ocn0;   // initial ocn (device layout), dropped after upload
  double atm_totV = 0.0;
  bool bg_go = true;
  SigDev sig{};                       // BIOGEM time-series integrals
  SliceDev slice{};                   // BIOGEM time-slice diagnostics (integrals allocated on first use)
  double sig_ben_Dmin = -1.0;
  double *sig2_dz = nullptr, *sig2_cv = nullptr;   // device copies of dz(0:maxk), cv(0:maxj) for the overturning stream function
  double *sig_w_ben = nullptr;
  double *sfxsumsed = nullptr, *sfcsumocn = nullptr, *sfxsumrok1 = nullptr;   // SEDGEM / ROKGEM interface sums, [ls|l][j][i][m]
  ~cg_handle() {
    cudaSetDevice(device);
    for (auto &gv : graph) for (auto &ge : gv) if (ge) cudaGraphExecDestroy(ge);
    for (auto &gv : graph2) for (auto &ge : gv) if (ge) cudaGraphExecDestroy(ge);
    for (void *p : allocs) cudaFree(p);
    if (wet_stage) cudaFree(wet_stage);
    for (auto &kv : xchg) {
      if (kv.second.up) cudaFree(kv.second.up);
      if (kv.second.down) cudaFree(kv.second.down);
      if (kv.second.up_done) cudaEventDestroy(kv.second.up_done);
      if (kv.second.down_ready) cudaEventDestroy(kv.second.down_ready);
    }
    if (xstream) cudaStreamDestroy(xstream);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (evFork) cudaEventDestroy(evFork);
    if (evJoin) cudaEventDestroy(evJoin);
    if (evFork5) cudaEventDestroy(evFork5);
    if (evJoin5) cudaEventDestroy(evJoin5);
    if (evTcOld) cudaEventDestroy(evTcOld);
    if (stream5) cudaStreamDestroy(stream5);
    if (evBGtail) cudaEventDestroy(evBGtail);
    if (evAtchem) cudaEventDestroy(evAtchem);
    if (evT) cudaEventDestroy(evT);
    if (evBG) cudaEventDestroy(evBG);
    if (stream4) cudaStreamDestroy(stream4);
    if (evFork3) cudaEventDestroy(evFork3);
    if (evJoin3) cudaEventDestroy(evJoin3);
    if (stream3) cudaStreamDestroy(stream3);
    if (stream2) cudaStreamDestroy(stream2);
    if (stream) cudaStreamDestroy(stream);
  }
};

// ------------------------------------------------------------------ device memory helpers
template <typename T>
static int dalloc(cg_handle *h, T **p, size_t n, bool zero = true) {
  void *q = nullptr;
  CUDA_OK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
  if (zero) CUDA_OK(cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
  h->allocs.push_back(q);
  *p = (T *)q;
  return CG_OK;
}
template <typename T>
static int dupload(cg_handle *h, T **p, const std::vector<T> &v) {
  int rc = dalloc(h, p, v.size(), false);
  if (rc) return rc;
  CUDA_OK(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CG_OK;
}
// per-member scalar -> [MS]
static int dparam(cg_handle *h, const double **p, const std::vector<double> &v) {
  std::vector<double> t(h->MS, v.empty() ? 0.0 : v[0]);
  for (int m = 0; m < h->M; m++) t[m] = v[m];
  double *q;
  int rc = dupload(h, &q, t);
  *p = q;
  return rc;
}
// per-member array of n elements (host [m][n]) -> device [n][MS]
static int dmember_array(cg_handle *h, double **p, size_t n, const std::vector<const std::vector<double> *> &src) {
  std::vector<double> t(n * h->MS, 0.0);
  for (int m = 0; m < h->MS; m++) {
    const std::vector<double> &s = *src[std::min(m, h->M - 1)];
    for (size_t q = 0; q < n; q++) t[q * h->MS + m] = s[q];
  }
  return dupload(h, p, t);
}

static void reg_field(cg_handle *h, const std::string &name, double *d, std::initializer_list<int> dims,
                      std::initializer_list<long long> strides) {
  FieldDesc f;
  f.d = d;
  f.nd = (int)dims.size();
  int q = 0;
  for (int x : dims) f.dims[q++] = x;
  q = 0;
  for (long long x : strides) f.strides[q++] = x;
  h->fields[name] = f;
}

static void fill_gridc(cg_handle *h) {
  GridC &c = h->gc;
  const Grid &g = h->g;
  memset(&c, 0, sizeof(c));
  c.I = g.I; c.J = g.J; c.K = g.K; c.L = g.L; c.M = h->M; c.MS = h->MS; c.nyear = g.nyear;
  c.ndta = h->base.ndta; c.isles = h->isl.isles; c.npi1 = h->isl.isles > 0 ? h->isl.npi[1] : 0;
  c.dphi = g.dphi; c.rdphi = g.rdphi; c.dzz = g.dzz; c.dt = g.dt;
  for (int j = 0; j <= g.J + 1 && j < kMaxJ; j++) {
    c.ds[j] = g.ds[j]; c.dsv[j] = g.dsv[j]; c.rds2[j] = g.rds2[j]; c.s[j] = g.s[j]; c.c[j] = g.c[j]; c.sv[j] = g.sv[j];
    c.cv[j] = g.cv[j]; c.rc[j] = g.rc[j]; c.rc2[j] = g.rc2[j]; c.rcv[j] = g.rcv[j]; c.rdsv[j] = g.rdsv[j];
    c.cv2[j] = g.cv2[j]; c.rds[j] = g.rds[j];
  }
  for (int k = 0; k <= g.K + 1 && k < kMaxK; k++) {
    c.dz[k] = g.dz[k]; c.dza[k] = g.dza[k]; c.rdz[k] = g.rdz[k]; c.rdza[k] = g.rdza[k]; c.zw[k] = g.zw[k];
    c.zro[k] = k < (int)g.zro.size() ? g.zro[k] : 0.0;
    c.ssmax[k] = h->mc.empty() ? 0.0 : h->mc[0].ssmax[k];
    c.diffmax[k] = (h->mc.empty() || k >= (int)h->mc[0].diffmax.size()) ? 0.0 : h->mc[0].diffmax[k];
    c.mlddec[k] = (h->mc.empty() || k >= (int)h->mc[0].mlddec.size()) ? 0.0 : h->mc[0].mlddec[k];
    c.mlddecd[k] = (h->mc.empty() || k >= (int)h->mc[0].mlddecd.size()) ? 0.0 : h->mc[0].mlddecd[k];
  }
}
static void upload_grid(cg_handle *h) {
  upload_grid_physics(h->gc, h->stream);
  upload_grid_tracer_strict(h->gc, h->stream);
  upload_grid_tracer_fast(h->gc, h->stream);
  upload_grid_tracer_col(h->gc, h->stream);
}
// Constant memory (the 1-D grid metrics, GridC) is per process and device, shared by every handle.  Handles of the same job
// -- the groups of cgenie_b200.EnsembleGroups, driven concurrently from host threads -- hold identical tables (M / MS are not
// read from constant memory), so nothing is re-uploaded when one of them follows another.  A handle with different tables
// (another job on the same device) waits for the device to drain before it replaces them: handles of different jobs may
// share a device but must not be driven concurrently.
static std::mutex g_active_mu;
static std::map<int, GridC> g_uploaded;   // device -> tables last uploaded
static cg_handle *g_active = nullptr;
static void activate(cg_handle *h) {
  cudaSetDevice(h->device);
  std::lock_guard<std::mutex> lock(g_active_mu);
  if (g_active == h) return;
  GridC want = h->gc;
  want.M = want.MS = 0;
  auto it = g_uploaded.find(h->device);
  if (it == g_uploaded.end() || memcmp(&it->second, &want, sizeof(GridC)) != 0) {
    if (it != g_uploaded.end()) cudaDeviceSynchronize();
    upload_grid(h);
    cudaStreamSynchronize(h->stream);
    g_uploaded[h->device] = want;
  }
  g_active = h;
}

// Order the main stream after everything the per-module entry points or cg_run left running on the side streams.
// Called by every entry point that reads or writes state from the host side.
#define IO0(x) do { int rc0_ = (x); if (rc0_) return rc0_; } while (0)
static int flush_lazy(cg_handle *h);
static int join_side(cg_handle *h) {
  if (h->lazy_stage) IO0(flush_lazy(h));
  if (h->mom_pending) {
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->evJoin, 0));
    h->mom_pending = false;
  }
  if (h->bg_pending) {
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->evBG, 0));
    h->bg_pending = false;
  }
  if (h->bg_tail_pending) {
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->evBGtail, 0));
    h->bg_tail_pending = false;
  }
  if (h->tc_old_pending) {
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->evTcOld, 0));
    h->tc_old_pending = false;
  }
  return CG_OK;
}

// The momentum step of a cycle may have been computed ahead of step_goldstein (eager_momentum).  It reads rho, u1 and constants
// and rewrites u, u1, ub, psi, gb, bp, sbp; of these only u1 carries over from cycle to cycle (velc: u = rel*u1 + (1-rel)*u_new,
// u1 = u, goldstein.f90:3648-3654).  A host write to one of its inputs makes the early result stale: u1 is rolled back to the
// snapshot taken before the early step, which step_goldstein then repeats with the new state -- the relaxation is applied
// once, as in the reference.  Writes to anything else (ts, cost, the atmosphere, BIOGEM's arrays ...) leave it valid.
static bool momentum_input(const char *name) {
  for (const char *n : {"rho", "u", "u1", "ub", "psi", "gb", "bp", "sbp"})
    if (strcmp(name, n) == 0) return true;
  return false;
}
static int drop_momentum(cg_handle *h) {
  if (!h->mom_ready) return CG_OK;
  if (h->mom_pending) { CUDA_OK(cudaStreamWaitEvent(h->stream, h->evJoin, 0)); h->mom_pending = false; }
  if (h->u1_snap)
    CUDA_OK(cudaMemcpyAsync(h->dv.u1, h->u1_snap, (size_t)2 * h->g.I * h->g.J * h->g.K * h->MS * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->mom_ready = false;
  return CG_OK;
}

// host-side constants of member 0, exposed through cg_get_const / cg_get_iconst
static void register_hconst(cg_handle *h, const MemberConsts &c) {
  const Grid &g = h->g;
  auto hc = [&](const char *n, const std::vector<double> &x) { h->hconst[n] = x; };
  h->hiconst["k1"] = g.k1; h->hiconst["ku"] = g.ku; h->hiconst["mk"] = g.mk; h->hiconst["getj"] = g.getj;
  h->hiconst["ips"] = g.ips; h->hiconst["ipf"] = g.ipf; h->hiconst["ias"] = g.ias; h->hiconst["iaf"] = g.iaf;
  h->hiconst["jsf"] = std::vector<int>(1, g.jsf);
  h->hiconst["ntot"] = std::vector<int>(1, g.ntot);
  hc("ds", g.ds); hc("dsv", g.dsv); hc("rds2", g.rds2); hc("dz", g.dz); hc("s", g.s); hc("c", g.c); hc("sv", g.sv);
  hc("cv", g.cv); hc("dza", g.dza); hc("zro", g.zro); hc("zw", g.zw); hc("rc", g.rc); hc("rc2", g.rc2); hc("rcv", g.rcv);
  hc("rdsv", g.rdsv); hc("cv2", g.cv2); hc("rds", g.rds); hc("rdz", g.rdz); hc("rdza", g.rdza); hc("asurf", g.asurf);
  hc("rh", g.rh);
  hc("mldketau", c.mldketau); hc("mlddec", c.mlddec); hc("mlddecd", c.mlddecd);
  hc("scalars", {g.dphi, g.rdphi, g.dzz, g.dt, c.diff1, c.diff2, c.adrag, c.ec[1], c.ec[2], c.ec[3], c.ec[4], c.rpmesco,
                 c.rsictscsf, c.dtatm, c.rdtdim, c.rfluxsca, c.rpmesca, c.dtsic, c.sic_rdtdim, c.diffsic});
  hc("ssmax", c.ssmax); hc("drag", c.drag); hc("rtv", c.rtv); hc("rtv3", c.rtv3); hc("rhosing", c.rhosing);
  hc("ts0", c.ts0); hc("rho0", c.rho0);
  if (!c.gap.empty()) { hc("gap", c.gap); hc("ratm", c.ratm); hc("ubisl", c.ubisl); hc("psisl", c.psisl); hc("erisl", c.erisl); }
  if (!c.tq0.empty()) {
    hc("tau", c.tau); hc("dztau", c.dztau); hc("dztav", c.dztav); hc("usurf", c.usurf); hc("diffa", c.diffa);
    hc("uatm", c.uatm); hc("albcl", c.albcl); hc("ca", c.ca); hc("pmeadj", c.pmeadj); hc("solfor", c.solfor);
    hc("tq0", c.tq0); hc("us_dztau", c.us_dztau); hc("us_dztav", c.us_dztav);
    h->hiconst["iroff"] = c.iroff; h->hiconst["jroff"] = c.jroff;
  }
}

// ------------------------------------------------------------------ life cycle
extern "C" const char *cg_last_error(void) { return g_err.c_str(); }

// Member stride: members rounded up to whole warps; above 128 to the strides the fixed-shape kernels are compiled for (256, 512:
// 128-member tiles of one handle), padding lanes carry copies of the last member.  More than 512 members: generic kernels.
static int member_stride_for(int n_members) {
  const int ms = ((n_members + 31) / 32) * 32;
  if (ms > 128 && ms <= 256) return 256;
  if (ms > 256 && ms <= 512) return 512;
  return ms;
}
extern "C" int cg_create(const char *jobdir, int n_members, int device, cg_handle **out) {
  if (!jobdir || !out || n_members < 1) return fail(CG_ERR_ARG, "cg_create: bad argument");
  std::unique_ptr<cg_handle> h(new cg_handle);
  h->device = device;
  h->M = n_members;
  h->MS = member_stride_for(n_members);
  std::string err;
  if (!load_job(jobdir, &h->base, &h->g, &h->isl, &h->w, &err)) {
    const bool io = err.find("could not open") != std::string::npos || err.find("too short") != std::string::npos;
    return fail(io ? CG_ERR_IO : CG_ERR_CONFIG, err);
  }
  if (h->g.J + 2 > kMaxJ || h->g.K + 2 > kMaxK) return fail(CG_ERR_CONFIG, "grid larger than the compiled metric tables");
  if (h->isl.isles < 1 || h->isl.isles > kMaxIsles) return fail(CG_ERR_CONFIG, "topographies with 1 .. 8 islands are on the B200 path");
  if (!load_biogem(jobdir, h->base, h->g, &h->bg, &err)) {
    const bool io = err.find("could not open") != std::string::npos;
    return fail(io ? CG_ERR_IO : CG_ERR_CONFIG, err);
  }
  if (h->bg.on) {  // the perturbable BIOGEM parameters travel with the member parameter sets
    Namelist nb;
    std::string e2;
    if (nb.load(std::string(jobdir) + "/data_BIOGEM", &e2)) {
      h->base.par_bio_k0_PO4 = nb.num("par_bio_k0_PO4", h->base.par_bio_k0_PO4);
      h->base.par_bio_remin_POC_eL1 = nb.num("par_bio_remin_POC_eL1", h->base.par_bio_remin_POC_eL1);
      h->base.par_bio_red_POC_CaCO3 = nb.num("par_bio_red_POC_CaCO3", h->base.par_bio_red_POC_CaCO3);
    }
  }
  h->mp.assign(n_members, h->base);
  {
    MemberConsts c0;  // member-0 constants are available without a device (bit-exactness checks)
    build_member(h->g, h->isl, h->w, h->base, &c0, nullptr);
    register_hconst(h.get(), c0);
  }
  *out = h.release();
  return CG_OK;
}

extern "C" int cg_set_member_param(cg_handle *h, const char *name, const double *values) {
  if (!h || !name || !values) return fail(CG_ERR_ARG, "cg_set_member_param: bad argument");
  if (h->initialised) return fail(CG_ERR_STATE, "cg_set_member_param after cg_initialise");
  // the slope limits ssmax(k) live in constant memory next to the grid metrics: one table for the whole ensemble
  if (strcmp(name, "ssmaxsurf") == 0 || strcmp(name, "ssmaxdeep") == 0) {
    for (int m = 1; m < h->M; m++)
      if (values[m] != values[0]) return fail(CG_ERR_ARG, std::string(name) + " is shared by all members of a handle (per-member values are not supported)");
    h->base.set(name, values[0]);
  }
  for (int m = 0; m < h->M; m++)
    if (!h->mp[m].set(name, values[m])) return fail(CG_ERR_ARG, std::string("unknown member parameter ") + name);
  return CG_OK;
}

extern "C" int cg_destroy(cg_handle *h) {
  if (!h) return CG_OK;
  { std::lock_guard<std::mutex> lock(g_active_mu); if (g_active == h) g_active = nullptr; }
  delete h;
  return CG_OK;
}

static int build_device(cg_handle *h);

extern "C" int cg_initialise(cg_handle *h) {
  if (!h) return fail(CG_ERR_ARG, "null handle");
  if (h->initialised) return fail(CG_ERR_STATE, "already initialised");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= h->device)
    return fail(CG_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = least urgent
    CUDA_OK(cudaStreamCreateWithPriority(&h->stream4, cudaStreamNonBlocking, lo));
  }
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream5, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&h->evFork5, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evJoin5, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evTcOld, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evT, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evBG, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evBGtail, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evAtchem, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evFork3, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evJoin3, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreate(&h->ev0));
  CUDA_OK(cudaEventCreate(&h->ev1));
  // per-member constants; members sharing adrag share one barotropic factorisation
  h->mc.resize(h->M);
  h->baro_group.assign(h->MS, 0);
  std::vector<int> leaders;
  for (int m = 0; m < h->M; m++) {
    int grp = -1;
    for (size_t q = 0; q < leaders.size(); q++)
      if (h->mp[leaders[q]].adrag == h->mp[m].adrag) { grp = (int)q; break; }
    const MemberConsts *shared = grp >= 0 ? &h->mc[leaders[grp]] : nullptr;
    build_member(h->g, h->isl, h->w, h->mp[m], &h->mc[m], shared);
    if (grp < 0) { grp = (int)leaders.size(); leaders.push_back(m); }
    else { h->mc[m].gap.clear(); h->mc[m].gap.shrink_to_fit(); h->mc[m].ratm.clear(); h->mc[m].ratm.shrink_to_fit(); }
    h->baro_group[m] = grp;
  }
  for (int m = h->M; m < h->MS; m++) h->baro_group[m] = h->baro_group[h->M - 1];
  h->nbaro = (int)leaders.size();
  for (int m = 0; m < h->M; m++)
    if (h->mp[m].olr_adj != 0.0) h->need_mean = true;
  fill_gridc(h);
  register_hconst(h, h->mc[0]);
  int rc = build_device(h);
  if (rc) return rc;
  // keep leaders' factors for cg_get_const, drop the rest of the heavy host arrays
  h->initialised = true;
  activate(h);
  if (h->bg.on) {
    // sub_init_carb (biogem_data.f90:2336-2430) and the biogem_climate call before the main loop (genie.f90:109-112)
    h->launches += launch_bg_step(h->dv, h->bgd, 1, 0, h->stream);
    h->launches += launch_bg_slice(h->dv, h->bgd, h->slice, 0.0, 1, h->stream);   // ... and the cells below the surface
    h->bgd.nsol = 0;
    h->launches += launch_bg_climate(h->dv, h->bgd, h->stream);
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}

static int build_device(cg_handle *h) {
  const Grid &g = h->g;
  const int I = g.I, J = g.J, K = g.K, L = g.L, M = h->M, MS = h->MS;
  const size_t ij = (size_t)I * J, ijk = ij * K;
  Dev &v = h->dv;
  memset(&v, 0, sizeof(v));
  v.I = I; v.J = J; v.K = K; v.L = L; v.M = M; v.MS = MS; v.nm = I * (J + 1); v.nbaro = h->nbaro;
  int rc;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  // ---- masks
  {
    std::vector<unsigned char> k1(g.k1.size()), ku(g.ku.size()), mk(g.mk.size()), gj(g.getj.size());
    for (size_t q = 0; q < k1.size(); q++) k1[q] = (unsigned char)std::min(g.k1[q], 255);
    for (size_t q = 0; q < ku.size(); q++) ku[q] = (unsigned char)std::min(g.ku[q], 255);
    for (size_t q = 0; q < mk.size(); q++) mk[q] = (unsigned char)std::min(g.mk[q], 255);
    for (size_t q = 0; q < gj.size(); q++) gj[q] = (unsigned char)g.getj[q];
    unsigned char *p;
    TRY(dupload(h, &p, k1)); v.k1 = p;
    TRY(dupload(h, &p, ku)); v.ku = p;
    TRY(dupload(h, &p, mk)); v.mk = p;
    TRY(dupload(h, &p, gj)); v.getj = p;
  }
  // ---- per-member scalars
  {
    // wet columns, deepest first (load balance: long columns start early)
    std::vector<int> wc;
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++)
        if (g.k1at(i, j) <= K) wc.push_back((i - 1) + I * (j - 1));
    std::stable_sort(wc.begin(), wc.end(), [&](int a, int b) { return g.k1at(a % I + 1, a / I + 1) < g.k1at(b % I + 1, b / I + 1); });
    v.nwet = (int)wc.size();
    if (wc.empty()) wc.push_back(0);
    int *q;
    TRY(dupload(h, &q, wc));
    v.wetcols = q;
    // the same columns row by row (fused column kernel: the CPB columns of a block are neighbours in i)
    std::sort(wc.begin(), wc.end());
    TRY(dupload(h, &q, wc));
    v.rowcols = q;
    // ... and poleward rows first: the convective adjustment has work at high latitudes only, so its long-running blocks should be
    // the first to be scheduled (the stable columns of the tropics fill the tail of the grid at no cost)
    std::stable_sort(wc.begin(), wc.end(), [&](int a, int b) {
      const int ja = a / I + 1, jb = b / I + 1;
      return std::min(ja - 1, J - ja) < std::min(jb - 1, J - jb);
    });
    TRY(dupload(h, &q, wc));
    v.polcols = q;
  }
  auto col = [&](auto getter) { std::vector<double> t(M); for (int m = 0; m < M; m++) t[m] = getter(h->mc[m]); return t; };
  MemberP &p = v.p;
  TRY(dparam(h, &p.diff1, col([](const MemberConsts &c) { return c.diff1; })));
  TRY(dparam(h, &p.diff2, col([](const MemberConsts &c) { return c.diff2; })));
  TRY(dparam(h, &p.ec1, col([](const MemberConsts &c) { return c.ec[1]; })));
  TRY(dparam(h, &p.ec2, col([](const MemberConsts &c) { return c.ec[2]; })));
  TRY(dparam(h, &p.ec3, col([](const MemberConsts &c) { return c.ec[3]; })));
  TRY(dparam(h, &p.ec4, col([](const MemberConsts &c) { return c.ec[4]; })));
  TRY(dparam(h, &p.ec5, col([](const MemberConsts &c) { return c.ec[5]; })));
  v.ieos = h->base.ieos; v.iconv = h->base.iconv;
  TRY(dparam(h, &p.rel, col([](const MemberConsts &c) { return c.p.rel; })));
  TRY(dparam(h, &p.scf, col([](const MemberConsts &c) { return c.p.scf; })));
  TRY(dparam(h, &p.saln0, col([](const MemberConsts &c) { return c.p.saln0; })));
  TRY(dparam(h, &p.rpmesco, col([](const MemberConsts &c) { return c.rpmesco; })));
  TRY(dparam(h, &p.rsictscsf, col([](const MemberConsts &c) { return c.rsictscsf; })));
  TRY(dparam(h, &p.albocn, col([](const MemberConsts &c) { return c.p.albocn; })));
  TRY(dparam(h, &p.hosing_trend, col([](const MemberConsts &c) { return c.hosing_trend; })));
  v.iediff = h->base.iediff; v.ediffpow2i = h->mc[0].ediffpow2i; v.ediffpow2 = h->base.ediffpow2;
  if (v.iediff) {   // per member: ediff1p scales with diff(2) - ediff0, both perturbable
    TRY(dparam(h, &p.ediff0, col([](const MemberConsts &c) { return c.ediff0; })));
    std::vector<const std::vector<double> *> src;
    for (int m = 0; m < M; m++) src.push_back(&h->mc[m].ediff1p);
    double *q; TRY(dmember_array(h, &q, (size_t)K + 2, src)); p.ediff1p = q;
  }
  {
    std::vector<int> t(MS, 0);
    for (int m = 0; m < M; m++) t[m] = h->mc[m].nsteps_hosing;
    int *q; TRY(dupload(h, &q, t)); p.nsteps_hosing = q;
    TRY(dupload(h, &q, h->baro_group)); v.baro_group = q;
  }
  {
    std::vector<double> t(MS, 0.0);
    for (int m = 0; m < M; m++) t[m] = h->mc[m].hosing;
    TRY(dupload(h, &v.hosing, t));
  }
  // ---- tracer state
  TRY(dalloc(h, &v.ts_cur, ijk * L * MS));
  TRY(dalloc(h, &v.ts_new, ijk * L * MS));
  TRY(dalloc(h, &v.tsflux, 2 * ij * MS));
  TRY(dalloc(h, &v.usnap, 2 * ij * MS));
  TRY(dalloc(h, &v.velsum, 2 * ij * MS));
  TRY(dalloc(h, &v.comask, ij * MS));
  TRY(dalloc(h, &v.rho, ijk * MS));
  TRY(dalloc(h, &v.u, ijk * 3 * MS));
  TRY(dalloc(h, &v.u1, ijk * 2 * MS));
  TRY(dalloc(h, &v.cost, ij * MS));
  TRY(dalloc(h, &v.istep_ocn, 1));
  if (!h->tracer_only) {
    // initial ts/rho: host (L,I,J,K)/(I,J,K) Fortran order == device cell order with l fastest
    std::vector<double> t(ijk * L * MS, 0.0), r(ijk * MS, 0.0);
    for (int m = 0; m < MS; m++) {
      const MemberConsts &c = h->mc[std::min(m, M - 1)];
      for (size_t q = 0; q < ijk * L; q++) t[q * MS + m] = c.ts0[q];
      for (size_t q = 0; q < ijk; q++) r[q * MS + m] = c.rho0[q];
    }
    if (h->bg.on) {
      // initialise_biogem: sub_init_tracer_ocn_comp (biogem_data.f90:1281-1309), sub_biogem_copy_tstoocn (:3745-3763) and
      // the salinity-normalised copy back to ts, sub_biogem_copy_ocntots (biogem_box.f90:3691-3739)
      const BgConfig &bc = h->bg;
      h->bg_ocn0.assign(ijk * L * MS, 0.0);
      std::vector<double> V(ijk, 0.0);
      for (int i = 1; i <= I; i++)
        for (int j = 1; j <= J; j++)
          for (int k = g.k1at(i, j); k <= K; k++)
            V[cell3(I, J, i, j, k)] = (kDsc * g.dz[k]) * (2.0 * kBgPi * (kBgREarth * kBgREarth) * (1.0 / I) * (g.sv[j] - g.sv[j - 1]));
      double totV = 0.0;
      for (size_t q = 0; q < ijk; q++) totV = totV + V[q];   // SUM(phys_ocn(ipo_V,:,:,:)), array element order
      for (int m = 0; m < MS; m++) {
        const double saln0 = h->mp[std::min(m, M - 1)].saln0;
        double sumSV = 0.0;
        for (size_t q = 0; q < ijk; q++) {
          const int k = (int)(q / ij) + 1, jj = (int)((q % ij) / I) + 1, ii = (int)(q % I) + 1;
          if (k < g.k1at(ii, jj)) continue;
          double *o = &h->bg_ocn0[(q * L) * MS + m];
          for (int l = 1; l <= L; l++) {
            if (bc.otype[l] == 1) o[(size_t)(l - 1) * MS] = bc.ocn_init[l];
            else if (bc.otype[l] >= 11)
              o[(size_t)(l - 1) * MS] = bg_iso_fraction(bc.ocn_init[l], bc.otype[l] == 11 ? kBgStd13C : kBgStd14C) * bc.ocn_init[bc.odep[l]];
          }
          o[0] = t[(q * L) * MS + m] + kBgZeroC;
          o[(size_t)MS] = t[(q * L + 1) * MS + m] + saln0;
        }
        for (size_t q = 0; q < ijk; q++) sumSV = sumSV + h->bg_ocn0[(q * L + 1) * MS + m] * V[q];
        const double meanS = sumSV / totV;
        for (size_t q = 0; q < ijk; q++) {
          const int k = (int)(q / ij) + 1, jj = (int)((q % ij) / I) + 1, ii = (int)(q % I) + 1;
          if (k < g.k1at(ii, jj)) continue;
          const double Sc = h->bg_ocn0[(q * L + 1) * MS + m];
          for (int l = 3; l <= L; l++) t[(q * L + (l - 1)) * MS + m] = h->bg_ocn0[(q * L + (l - 1)) * MS + m] * (meanS / Sc);
        }
      }
    }
    CUDA_OK(cudaMemcpy(v.ts_cur, t.data(), t.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(v.ts_new, t.data(), t.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(v.rho, r.data(), r.size() * 8, cudaMemcpyHostToDevice));
  }
  reg_field(h, "ts", v.ts_cur, {L, I, J, K}, {1, L, (long long)L * I, (long long)L * I * J});
  reg_field(h, "rho", v.rho, {I, J, K}, {1, I, (long long)I * J});
  reg_field(h, "u", v.u, {3, I, J, K}, {1, 3, 3LL * I, 3LL * I * J});
  reg_field(h, "u1", v.u1, {2, I, J, K}, {1, 2, 2LL * I, 2LL * I * J});
  reg_field(h, "cost", v.cost, {I, J}, {1, I});
  v.imld = h->base.imld;
  if (v.imld) {   // Kraus-Turner mixed-layer scheme (goldstein.f90:2294-2390, 3337-3442)
    v.mldpebuoycoeff = h->base.mldpebuoycoeff;
    std::vector<const std::vector<double> *> src;
    for (int m = 0; m < M; m++) src.push_back(&h->mc[m].mldketau);
    double *q; TRY(dmember_array(h, &q, ij, src)); v.mldketau = q;
    TRY(dalloc(h, &v.mld_pel1, ij * MS));
    TRY(dalloc(h, &v.mld_rhoold, ijk * MS));
    TRY(dalloc(h, &v.mld, ij * MS));
    TRY(dalloc(h, &v.mldk, ij * MS));
    reg_field(h, "mld", v.mld, {I, J}, {1, I});
    reg_field(h, "mldketau", const_cast<double *>(v.mldketau), {I, J}, {1, I});
    reg_field(h, "mldpelayer1", v.mld_pel1, {I, J}, {1, I});
  }
  reg_field(h, "tsflux", v.tsflux, {2, I, J}, {(long long)ij, 1, I});
  if (L > 2) {
    // BIOGEM tracer coupling state (sub_init_phys_ocn, biogem_data.f90:1098-1137; ts->ocn offsets biogem.f90:283-285)
    const double pi_bg = 3.141592653589793, rEarth = 6.37e6, m3_kg = 1027.649;
    std::vector<double> V(ijk, 0.0), Mm(ijk * MS, 0.0), rMm(ijk * MS, 0.0);
    std::vector<int> cols;
    for (int i = 1; i <= I; i++)
      for (int j = 1; j <= J; j++)
        if (g.k1at(i, j) <= K) {
          cols.push_back((i - 1) + I * (j - 1));
          for (int k = g.k1at(i, j); k <= K; k++) {
            const double dD = kDsc * g.dz[k];
            const double A = 2.0 * pi_bg * (rEarth * rEarth) * (1.0 / I) * (g.sv[j] - g.sv[j - 1]);
            const size_t c = cell3(I, J, i, j, k);
            V[c] = dD * A;
            const double M0 = m3_kg * V[c], rM0 = 1.0 / M0;
            for (int m = 0; m < MS; m++) { Mm[c * MS + m] = M0; rMm[c * MS + m] = rM0; }
          }
        }
    // total volume in the reference's order: per-column sums over k, then over columns (biogem.f90:1934-1941)
    double totV = 0.0;
    for (int c2 : cols) {
      const int i = c2 % I + 1, j = c2 / I + 1;
      double sV = 0.0;
      for (int k = g.k1at(i, j); k <= K; k++) sV = sV + V[cell3(I, J, i, j, k)];
      totV = totV + sV;
    }
    v.bg_rtot_V = 1.0 / totV;
    if (cols.empty()) cols.push_back(0);
    { double *q; TRY(dupload(h, &q, V)); v.bg_V = q; }
    TRY(dupload(h, &v.bg_M, Mm));
    TRY(dupload(h, &v.bg_rM, rMm));
    { int *q; TRY(dupload(h, &q, cols)); v.bgcols = q; }
    h->bg_colidx.assign((size_t)I * J, -1);
    for (int n = 0; n < v.nwet; n++) h->bg_colidx[cols[n]] = n;
    TRY(dalloc(h, &v.bg_ocn, ijk * L * MS));
    TRY(dalloc(h, &v.bg_vdocn, ijk * L * MS));
    TRY(dalloc(h, &v.bg_part, (size_t)2 * L * std::max(v.nwet, 1) * MS));
    TRY(dalloc(h, &v.bg_tot, (size_t)(3 * L + 4) * MS));   // 2L-2 totals, then the per-member factors of k_tc_factors
    reg_field(h, "ocn", v.bg_ocn, {L, I, J, K}, {1, L, (long long)L * I, (long long)L * I * J});
    reg_field(h, "vdocn", v.bg_vdocn, {L, I, J, K}, {1, L, (long long)L * I, (long long)L * I * J});
    reg_field(h, "bg_M", v.bg_M, {I, J, K}, {1, I, (long long)I * J});
    reg_field(h, "bg_rM", v.bg_rM, {I, J, K}, {1, I, (long long)I * J});
    h->hconst["bg_V"] = V;
  }
  if (h->bg.on) {
    const BgConfig &bc = h->bg;
    BgDev &b = h->bgd;
    const int LS = bc.LS, LA = bc.LA;
    bg_fill_tables(bc, h->base, g, &b);
    if (!bg_layout_ok(b, L)) return fail(CG_ERR_CONFIG, "BIOGEM: tracer tables differ from the layout k_bg_step is compiled for");
    // imld = 1 hands BIOGEM a mixed-layer depth (go_mldta) and sub_calc_bio_uptake then spreads the export production over the levels
    // k_mld .. n_k (biogem_box.f90:423-430): bg_k_mld in the sweep and cells kernels
    if (h->base.imld) {
      TRY(dalloc(h, &b.mld, ij * MS));
      TRY(dalloc(h, &b.mld_stage, ij * MS));
      reg_field(h, "bg_mld", b.mld, {I, J}, {1, I});
    }
    if (L > 64) return fail(CG_ERR_CONFIG, "more than 64 tracers");
    CUDA_OK(cudaMemcpy(v.bg_ocn, h->bg_ocn0.data(), h->bg_ocn0.size() * 8, cudaMemcpyHostToDevice));
    h->bg_ocn0.clear(); h->bg_ocn0.shrink_to_fit();
    TRY(dalloc(h, &b.bio_part, ijk * LS * MS));
    TRY(dalloc(h, &b.settle_k1, ij * LS * MS));
    TRY(dalloc(h, &b.carbH, ij * MS));
    TRY(dalloc(h, &b.surf, (size_t)kBgSurfSlots * ij * MS));
    {
      const char *e = getenv("CG_BG_PD");
      h->bg_pd = !(e && atoi(e) == 0);
      TRY(dalloc(h, &b.lrem, ijk * 7 * MS));
      TRY(dalloc(h, &b.fsedv, (size_t)7 * std::max(v.nwet, 1) * MS));
      TRY(dalloc(h, &b.pscale, (size_t)MS));
      int *q; TRY(dupload(h, &q, h->bg_colidx)); b.colidx = q;
    }
    TRY(dalloc(h, &b.seaice, ij * MS));
    TRY(dalloc(h, &b.seaice_stage, ij * MS));
    TRY(dalloc(h, &b.tq_stage, 2 * ij * MS));
    if (!getenv("CG_ATCHEM_FUSED")) TRY(dalloc(h, &b.atm_tot, (size_t)LA * MS));
    TRY(dalloc(h, &b.sfxsumatm, ij * LA * MS));
    TRY(dalloc(h, &b.sfcocn1, ij * L * MS));
    TRY(dalloc(h, &b.sfxsed1, ij * LS * MS));
    TRY(dalloc(h, &b.focnatm, ij * LA * MS));
    TRY(dalloc(h, &b.err, MS));
    v.bg_biopart = b.bio_part; v.bg_LS = LS;
    {
      // SST/SSS as exported by the last step_goldstein (initially the initial state, goldstein.f90:2061-2068)
      std::vector<double> sst(2 * ij * MS, 0.0);
      for (int m = 0; m < MS; m++) {
        const MemberConsts &c = h->mc[std::min(m, M - 1)];
        for (int j = 1; j <= J; j++)
          for (int i = 1; i <= I; i++)
            if (g.k1at(i, j) <= K) {
              const size_t q = cell3(I, J, i, j, K) * L;
              sst[cell2(I, i, j) * MS + m] = c.ts0[q];
              sst[(ij + cell2(I, i, j)) * MS + m] = c.ts0[q + 1];
            }
      }
      TRY(dupload(h, &v.sst, sst));
      reg_field(h, "sst", v.sst, {2, I, J}, {(long long)ij, 1, I});
    }
    { double *q; TRY(dupload(h, &q, bc.windspeed)); b.wspeed = q; }
    {
      std::vector<double> A(ij), rA(ij), aA(ij), aV(ij);
      // sub_init_phys_ocnatm (biogem_data.f90:1143-1161) and sub_init_phys_atm (atchem_data.f90:195-229)
      const double th0 = -kBgPi / 2, th1 = kBgPi / 2;
      const double s0 = std::sin(th0), s1 = std::sin(th1);
      const double ds = (s1 - s0) / (double)J;
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++) {
          const size_t c = cell2(I, i, j);
          A[c] = 2.0 * kBgPi * (kBgREarth * kBgREarth) * (1.0 / I) * (g.sv[j] - g.sv[j - 1]);
          rA[c] = 1.0 / A[c];
          const double svj = s0 + (double)j * ds, svjm = s0 + (double)(j - 1) * ds;
          aA[c] = 2.0 * kBgPi * (kBgREarth * kBgREarth) * (1.0 / (double)I) * (svj - svjm);
          aV[c] = 7777.0 * aA[c];   // par_atm_th, atchem_lib.f90:86
        }
      h->atm_totV = 0.0;
      for (size_t c = 0; c < ij; c++) h->atm_totV = h->atm_totV + aV[c];
      double *q;
      TRY(dupload(h, &q, A)); b.A = q;
      TRY(dupload(h, &q, rA)); b.rA = q;
      TRY(dupload(h, &q, aA)); b.atm_A = q;
      TRY(dupload(h, &q, aV)); b.atm_V = q;
    }
    {
      // sub_init_tracer_atm_comp (atchem_data.f90:234-261) and the initial cpl_comp_atmocn (genie.f90:87-90)
      std::vector<double> atm(ij * LA * MS, 0.0), sfc(ij * LA * MS, 0.0);
      for (int la = 1; la <= LA; la++) {
        double val = 0.0;
        if (bc.atype[la] == 0) { if (bc.ia[la] == 1) val = kBgZeroC; }
        else if (bc.atype[la] == 1) val = bc.atm_init[la];
        else val = bg_iso_fraction(bc.atm_init[la], bc.atype[la] == 11 ? kBgStd13C : kBgStd14C) * bc.atm_init[bc.adep[la]];
        for (size_t c = 0; c < ij * MS; c++) {
          atm[(size_t)(la - 1) * ij * MS + c] = val;
          if (la >= 3) sfc[(size_t)(la - 1) * ij * MS + c] = val;
        }
      }
      // the initial cpl_comp_EMBM (genie.f90:90): tstar_atm, surf_qstar_atm as initialise_embm left them
      for (int m = 0; m < MS; m++) {
        const MemberConsts &c = h->mc[std::min(m, M - 1)];
        for (size_t q = 0; q < ij; q++) {
          sfc[q * MS + m] = c.tq0[0 + 2 * q];
          sfc[(ij + q) * MS + m] = c.tq0[1 + 2 * q];
        }
      }
      TRY(dupload(h, &b.atm, atm));
      TRY(dupload(h, &b.sfcatm1, sfc));
    }
    {
      // perturbable parameters and the e-folding table of POC fraction 1 (1 - exp(-dD(k)/eL1)), [k][m]
      std::vector<double> k0(M), rr(M), f1((size_t)(K + 1) * MS, 0.0);
      for (int m = 0; m < M; m++) { k0[m] = h->mp[m].par_bio_k0_PO4; rr[m] = h->mp[m].par_bio_red_POC_CaCO3; }
      for (int m = 0; m < MS; m++)
        for (int k = 1; k <= K; k++) f1[(size_t)k * MS + m] = (1.0 - std::exp(-b.dD[k] / h->mp[std::min(m, M - 1)].par_bio_remin_POC_eL1));
      TRY(dparam(h, &b.k0_PO4, k0));
      TRY(dparam(h, &b.red_POC_CaCO3, rr));
      { double *q; TRY(dupload(h, &q, f1)); b.POC_f1 = q; }
    }
    reg_field(h, "bio_part", b.bio_part, {LS, I, J, K}, {1, LS, (long long)LS * I, (long long)LS * I * J});
    reg_field(h, "settle_k1", b.settle_k1, {LS, I, J}, {1, LS, (long long)LS * I});
    reg_field(h, "carbH", b.carbH, {I, J}, {1, I});
    reg_field(h, "bg_seaice", b.seaice, {I, J}, {1, I});
    reg_field(h, "atm", b.atm, {LA, I, J}, {(long long)ij, 1, I});
    reg_field(h, "sfcatm1", b.sfcatm1, {LA, I, J}, {(long long)ij, 1, I});
    reg_field(h, "sfxsumatm", b.sfxsumatm, {LA, I, J}, {(long long)ij, 1, I});
    reg_field(h, "focnatm", b.focnatm, {LA, I, J}, {(long long)ij, 1, I});
    reg_field(h, "sfcocn1", b.sfcocn1, {L, I, J}, {(long long)ij, 1, I});
    reg_field(h, "sfxsed1", b.sfxsed1, {LS, I, J}, {(long long)ij, 1, I});
    {
      // time-series integrals (cg_biogem_sig_update): bottom level and area per column, sums and integrals [q][m]
      std::vector<int> kb(ij);
      std::vector<double> Aall(ij);
      double totA = 0.0;
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++) {
          const size_t c = cell2(I, i, j);
          kb[c] = g.k1at(i, j) <= K ? g.k1at(i, j) - 1 : K;
          Aall[c] = 2.0 * kBgPi * (kBgREarth * kBgREarth) * (1.0 / I) * (g.sv[j] - g.sv[j - 1]);   // phys_ocnatm(ipoa_A), biogem_data.f90:1156
        }
      for (int j = 1; j <= J; j++)          // SUM(phys_ocnatm(ipoa_A,:,:)) in array element order
        for (int i = 1; i <= I; i++) totA = totA + Aall[cell2(I, i, j)];
      {
        double totAo = 0.0;                 // SUM(phys_ocn(ipo_A,:,:,n_k)): the ocean cells (loc_ocn_tot_A, biogem_data_ascii.f90:695)
        for (int j = 1; j <= J; j++)
          for (int i = 1; i <= I; i++) if (g.k1at(i, j) <= K) totAo = totAo + Aall[cell2(I, i, j)]; else totAo = totAo + 0.0;
        h->hconst["bg_ocn_tot_A"] = std::vector<double>(1, totAo);
      }
      const int nq = kSigHead + 3 * L + LA;
      int *qi; double *qd;
      TRY(dupload(h, &qi, kb)); h->sig.kbot = qi;
      TRY(dupload(h, &qd, Aall)); h->sig.A = qd;
      TRY(dalloc(h, &h->sig_w_ben, ij)); h->sig.w_ben = h->sig_w_ben;
      TRY(dalloc(h, &h->sig.raw, (size_t)nq * MS));
      TRY(dalloc(h, &h->sig.acc, (size_t)nq * MS));
      h->sig.rtot_A_atm = totA > kBgNullSmall ? 1.0 / totA : 0.0;
      h->sig.LA = LA;
      reg_field(h, "bg_sig", h->sig.acc, {nq}, {1});
    }
    {
      // time-slice diagnostics: [H+] and RF0 of every cell from sub_init_carb on, the wet-cell list
      std::vector<int> wet;
      for (int k = 1; k <= K; k++)
        for (int j = 1; j <= J; j++)
          for (int i = 1; i <= I; i++)
            if (k >= g.k1at(i, j)) wet.push_back((int)cell3(I, J, i, j, k));
      h->slice.nwet3 = (int)wet.size();
      if (wet.empty()) wet.push_back(0);
      int *qi; TRY(dupload(h, &qi, wet)); h->slice.wet = qi;
      TRY(dalloc(h, &h->slice.carbH3, ijk * MS));
      TRY(dalloc(h, &h->slice.rf03, ijk * MS));
      reg_field(h, "carbH3", h->slice.carbH3, {I, J, K}, {1, I, (long long)I * J});
    }
    // genie_sfxsumsed, genie_sfcsumocn, genie_sfxsumrok1 (genie_global.f90) of a job whose sediment grid is the ocean grid
    TRY(dalloc(h, &h->sfxsumsed, ij * LS * MS));
    TRY(dalloc(h, &h->sfcsumocn, ij * L * MS));
    TRY(dalloc(h, &h->sfxsumrok1, ij * L * MS));
    reg_field(h, "sfxsumsed", h->sfxsumsed, {LS, I, J}, {(long long)ij, 1, I});
    reg_field(h, "sfcsumocn", h->sfcsumocn, {L, I, J}, {(long long)ij, 1, I});
    reg_field(h, "sfxsumrok1", h->sfxsumrok1, {L, I, J}, {(long long)ij, 1, I});
  }
  TRY(dalloc(h, &h->d_meantemp, MS));
  TRY(dalloc(h, &h->d_means, (size_t)MS * L));
  TRY(dalloc(h, &h->d_flags, MS));
  if (h->tracer_only) return CG_OK;

  // ---- momentum
  TRY(dalloc(h, &v.bp, ijk * MS));
  TRY(dalloc(h, &v.sbp, ij * MS));
  TRY(dalloc(h, &v.gb, (size_t)v.nm * MS));
  TRY(dalloc(h, &v.ub, (size_t)2 * (I + 2) * (J + 1) * MS));
  TRY(dalloc(h, &v.psi, (size_t)(I + 1) * (J + 1) * MS));
  TRY(dalloc(h, &v.erisl_rhs, (size_t)std::max(h->isl.isles, 1) * MS));   // [island][m]
  TRY(dalloc(h, &v.psibc, (size_t)std::max(h->isl.isles, 1) * MS));
  { double *q; TRY(dupload(h, &q, g.rh)); v.rh = q; }
  auto members = [&](auto getter) {
    std::vector<const std::vector<double> *> s(M);
    for (int m = 0; m < M; m++) s[m] = &getter(h->mc[m]);
    return s;
  };
  { double *q;
    TRY(dmember_array(h, &q, (size_t)2 * (I + 1) * J, members([](const MemberConsts &c) -> const std::vector<double> & { return c.drag; }))); v.drag = q;
    TRY(dmember_array(h, &q, ij, members([](const MemberConsts &c) -> const std::vector<double> & { return c.rtv; }))); v.rtv = q;
    TRY(dmember_array(h, &q, ij, members([](const MemberConsts &c) -> const std::vector<double> & { return c.rtv3; }))); v.rtv3 = q;
  }
  // (2,I,J) host arrays (l fastest) -> device component-major [l][c2][m]
  auto comp_major = [&](const std::vector<double> &a) {
    std::vector<double> t(2 * ij);
    for (size_t c2 = 0; c2 < ij; c2++) { t[c2] = a[0 + 2 * c2]; t[ij + c2] = a[1 + 2 * c2]; }
    return t;
  };
  std::vector<std::vector<double>> tmpA(M), tmpB(M), tmpC(M);
  for (int m = 0; m < M; m++) { tmpA[m] = comp_major(h->mc[m].tau); tmpB[m] = comp_major(h->mc[m].dztau); tmpC[m] = comp_major(h->mc[m].dztav); }
  auto vecs = [&](std::vector<std::vector<double>> &x) { std::vector<const std::vector<double> *> s(M); for (int m = 0; m < M; m++) s[m] = &x[m]; return s; };
  { double *q;
    TRY(dmember_array(h, &q, 2 * ij, vecs(tmpA))); v.tau = q;
    TRY(dmember_array(h, &q, 2 * ij, vecs(tmpB))); v.dztau = q;
    TRY(dmember_array(h, &q, 2 * ij, vecs(tmpC))); v.dztav = q;
  }
  // barotropic factors per group, band rows made contiguous: ratm[row][t], gap[row][col]
  {
    const int nm = v.nm, bw = I + 1, gw = 2 * I + 3;
    std::vector<double> R((size_t)h->nbaro * nm * bw), G((size_t)h->nbaro * nm * gw), UBI, PSI, ER;
    std::vector<int> seen(h->nbaro, 0);
    const int nis = h->isl.isles;
    const size_t nub = (size_t)2 * (I + 2) * (J + 1), nps = (size_t)(I + 1) * (J + 1), ner = (size_t)nis * (nis + 1);
    UBI.resize((size_t)h->nbaro * nis * nub);      // [group][island][...]
    PSI.resize((size_t)h->nbaro * nis * nps);
    ER.resize((size_t)h->nbaro * ner);             // erisl(isl, 1:isles) after matinv_gold, Fortran order (isl fastest)
    for (int m = 0; m < M; m++) {
      const int grp = h->baro_group[m];
      if (seen[grp]) continue;
      seen[grp] = 1;
      const MemberConsts &c = h->mc[m];
      for (int r = 0; r < nm; r++) {
        for (int t = 0; t < bw; t++) R[((size_t)grp * nm + r) * bw + t] = c.ratm[(size_t)r + (size_t)nm * t];
        for (int q = 0; q < gw; q++) G[((size_t)grp * nm + r) * gw + q] = c.gap[(size_t)r + (size_t)nm * q];
      }
      std::copy(c.ubisl.begin(), c.ubisl.begin() + nis * nub, UBI.begin() + (size_t)grp * nis * nub);
      std::copy(c.psisl.begin(), c.psisl.begin() + nis * nps, PSI.begin() + (size_t)grp * nis * nps);
      for (size_t q = 0; q < ner; q++) ER[(size_t)grp * ner + q] = q < c.erisl.size() ? c.erisl[q] : 0.0;
    }
    {
      // pivot-major copies for the warp-cooperative solve
      std::vector<double> BF((size_t)h->nbaro * nm * bw, 0.0), BB((size_t)h->nbaro * nm * bw, 0.0), RD((size_t)h->nbaro * nm, 0.0);
      for (int grp = 0; grp < h->nbaro; grp++) {
        const double *Rg = &R[(size_t)grp * nm * bw], *Gg = &G[(size_t)grp * nm * gw];
        for (int i = 1; i <= nm; i++) {
          RD[(size_t)grp * nm + i - 1] = 1.0 / Gg[(size_t)(i - 1) * gw + (I + 1)];
          for (int t = 1; t <= bw; t++) {
            if (i + t <= nm) BF[((size_t)grp * nm + i - 1) * bw + t - 1] = Rg[(size_t)(i + t - 1) * bw + (t - 1)];
            if (i - t >= 1) BB[((size_t)grp * nm + i - 1) * bw + t - 1] = Gg[(size_t)(i - t - 1) * gw + (I + 1 + t)];
          }
        }
      }
      TRY(dupload(h, &h->d_bf, BF));
      TRY(dupload(h, &h->d_bb, BB));
      TRY(dupload(h, &h->d_rd, RD));
    }
    if (bw <= 64) {
      // Blocked form of the two banded substitutions (k_baro_blk): the rows are taken in blocks of 32, in sweep order
      // (forward: e' = e; backward: e' = npad-1-e, so that both are lower triangular).  Per block the slab holds, lane =
      // row within the block: bw rows W(:, d) = Tbb^-1 x (coefficients of the value d rows before the block), d = 1..bw,
      // then the 32 x 32 inverse Tbb^-1 of the block's own triangle, column by column.  Both are formed in extended
      // precision; padding rows are identity rows.
      const int nb = (nm + 31) / 32, npad = nb * 32, T = bw + 32;
      std::vector<double> BK((size_t)h->nbaro * 2 * nb * T * 32, 0.0);
      for (int grp = 0; grp < h->nbaro; grp++) {
        const double *Rg = &R[(size_t)grp * nm * bw], *Gg = &G[(size_t)grp * nm * gw];
        for (int sw = 0; sw < 2; sw++) {
          auto offd = [&](const int ep, const int d) -> double {   // T(e', e'-d), d >= 1
            if (sw == 0) return (ep < nm && ep - d >= 0 && d <= bw) ? Rg[(size_t)ep * bw + (d - 1)] : 0.0;
            const int e = npad - 1 - ep;
            return (e < nm && e + d < nm && d <= bw) ? Gg[(size_t)e * gw + (I + 1 + d)] : 0.0;
          };
          auto diag = [&](const int ep) -> double {
            if (sw == 0) return 1.0;
            const int e = npad - 1 - ep;
            return e < nm ? Gg[(size_t)e * gw + (I + 1)] : 1.0;
          };
          for (int B = 0; B < nb; B++) {
            double *slab = &BK[(((size_t)grp * 2 + sw) * nb + B) * T * 32];
            long double Ai[32][32];
            for (int c = 0; c < 32; c++)
              for (int l = 0; l < 32; l++) {
                long double acc = (l == c) ? 1.0L : 0.0L;
                for (int k2 = (l - bw > c ? l - bw : c); k2 < l; k2++) acc -= (long double)offd(32 * B + l, l - k2) * Ai[k2][c];
                Ai[l][c] = (l < c) ? 0.0L : acc / (long double)diag(32 * B + l);
              }
            for (int c = 0; c < 32; c++)
              for (int l = 0; l < 32; l++) slab[(size_t)(bw + c) * 32 + l] = (double)Ai[l][c];
            // W = Tbb^-1 x (off-block coefficients): row e' of the block against the value d rows before the block's
            // first row, d = 1..bw, so that the block's unknowns are z - W y_prev with z = Tbb^-1 b_blk
            for (int d = 1; d <= bw; d++) {
              if (32 * B - d < 0) continue;
              for (int l = 0; l < 32; l++) {
                long double acc = 0.0L;
                for (int c = 0; c <= l; c++) acc += Ai[l][c] * (long double)offd(32 * B + c, c + d);   // row c reaches back c + d
                slab[(size_t)(d - 1) * 32 + l] = (double)acc;
              }
            }
          }
        }
      }
      TRY(dupload(h, &h->d_bk, BK));
    }
    double *q;
    TRY(dupload(h, &q, R)); v.ratm = q;
    TRY(dupload(h, &q, G)); v.gap = q;
    TRY(dupload(h, &q, UBI)); v.ubisl = q;
    TRY(dupload(h, &q, PSI)); v.psisl = q;
    TRY(dupload(h, &q, ER)); v.erisl = q;
  }
  { double *q; TRY(dupload(h, &q, h->mc[0].rhosing)); v.rhosing = q; }
  {
    // island paths, [island][mpi] as the host holds them
    const int nis = h->isl.isles, mpi = h->isl.mpi;
    std::vector<int> a((size_t)nis * mpi, 0), b((size_t)nis * mpi, 1), c((size_t)nis * mpi, 1), np(nis, 0);
    for (int is = 1; is <= nis; is++) {
      np[is - 1] = h->isl.npi[is];
      for (int q = 0; q < h->isl.npi[is]; q++) {
        const size_t src = (size_t)q + (size_t)mpi * (is - 1);
        a[src] = h->isl.lpisl[src]; b[src] = h->isl.ipisl[src]; c[src] = h->isl.jpisl[src];
      }
    }
    int *q;
    TRY(dupload(h, &q, a)); v.lpisl = q;
    TRY(dupload(h, &q, b)); v.ipisl = q;
    TRY(dupload(h, &q, c)); v.jpisl = q;
    TRY(dupload(h, &q, np)); v.npi = q;
    v.isles = nis; v.mpi = mpi;
  }
  reg_field(h, "ub", v.ub, {2, I + 2, J + 1}, {1, 2, 2LL * (I + 2)});
  reg_field(h, "psi", v.psi, {I + 1, J + 1}, {1, I + 1});
  reg_field(h, "gb", v.gb, {v.nm}, {1});
  reg_field(h, "bp", v.bp, {I, J, K}, {1, I, (long long)I * J});
  reg_field(h, "sbp", v.sbp, {I, J}, {1, I});

  // ---- EMBM / surflux / sea ice
  TRY(dparam(h, &p.dtatm, col([](const MemberConsts &c) { return c.dtatm; })));
  TRY(dparam(h, &p.rdtdim, col([](const MemberConsts &c) { return c.rdtdim; })));
  TRY(dparam(h, &p.rfluxsca, col([](const MemberConsts &c) { return c.rfluxsca; })));
  TRY(dparam(h, &p.rpmesca, col([](const MemberConsts &c) { return c.rpmesca; })));
  TRY(dparam(h, &p.rmax, col([](const MemberConsts &c) { return c.p.rmax; })));
  TRY(dparam(h, &p.betaz1, col([](const MemberConsts &c) { return c.p.betaz1; })));
  TRY(dparam(h, &p.betaz2, col([](const MemberConsts &c) { return c.p.betaz2; })));
  TRY(dparam(h, &p.betam1, col([](const MemberConsts &c) { return c.p.betam1; })));
  TRY(dparam(h, &p.betam2, col([](const MemberConsts &c) { return c.p.betam2; })));
  TRY(dparam(h, &p.ppmin, col([](const MemberConsts &c) { return c.ppmin; })));
  TRY(dparam(h, &p.ppmax, col([](const MemberConsts &c) { return c.ppmax; })));
  TRY(dparam(h, &p.delf2x, col([](const MemberConsts &c) { return c.p.delf2x; })));
  TRY(dparam(h, &p.olr_adj0, col([](const MemberConsts &c) { return c.p.olr_adj0; })));
  TRY(dparam(h, &p.olr_adj, col([](const MemberConsts &c) { return c.p.olr_adj; })));
  TRY(dparam(h, &p.t_eqm, col([](const MemberConsts &c) { return c.p.t_eqm; })));
  TRY(dparam(h, &p.par_sich_max, col([](const MemberConsts &c) { return c.p.par_sich_max; })));
  TRY(dparam(h, &p.par_albsic_min, col([](const MemberConsts &c) { return c.p.par_albsic_min; })));
  TRY(dparam(h, &p.par_albsic_max, col([](const MemberConsts &c) { return c.p.par_albsic_max; })));
  TRY(dparam(h, &p.rate_co2, col([](const MemberConsts &c) { return c.rate_co2; })));
  TRY(dparam(h, &p.rate_ch4, col([](const MemberConsts &c) { return c.rate_ch4; })));
  TRY(dparam(h, &p.rate_n2o, col([](const MemberConsts &c) { return c.rate_n2o; })));
  TRY(dparam(h, &p.hatmbl2, col([](const MemberConsts &c) { return c.hatmbl2; })));
  TRY(dparam(h, &p.dtsic, col([](const MemberConsts &c) { return c.dtsic; })));
  TRY(dparam(h, &p.sic_rdtdim, col([](const MemberConsts &c) { return c.sic_rdtdim; })));
  TRY(dparam(h, &p.diffsic, col([](const MemberConsts &c) { return c.diffsic; })));
  TRY(dparam(h, &p.par_sica_thresh, col([](const MemberConsts &c) { return c.p.par_sica_thresh; })));
  TRY(dparam(h, &p.par_sich_thresh, col([](const MemberConsts &c) { return c.p.par_sich_thresh; })));
  {
    // diffa host (l, m2, j) l fastest -> device [j][(l-1)+2*(m2-1)][m]: same linear order
    double *q;
    TRY(dmember_array(h, &q, (size_t)4 * J, members([](const MemberConsts &c) -> const std::vector<double> & { return c.diffa; })));
    p.diffa = q;
  }
  // effective advective winds: step_embm resets uatm = lowestlu2/usc every call (embm.f90:49-50)
  std::vector<std::vector<double>> uu(M), vv(M);
  for (int m = 0; m < M; m++) {
    uu[m].resize(ij); vv[m].resize(ij);
    for (size_t c2 = 0; c2 < ij; c2++) { uu[m][c2] = h->mc[m].lowestlu2[c2] / kUsc; vv[m][c2] = h->mc[m].lowestlv3[c2] / kUsc; }
  }
  { double *q;
    TRY(dmember_array(h, &q, ij, vecs(uu))); v.uatm_u = q;
    TRY(dmember_array(h, &q, ij, vecs(vv))); v.uatm_v = q;
    TRY(dmember_array(h, &q, ij, members([](const MemberConsts &c) -> const std::vector<double> & { return c.usurf; }))); v.usurf = q;
    TRY(dmember_array(h, &q, ij, members([](const MemberConsts &c) -> const std::vector<double> & { return c.albcl; }))); v.albcl = q;
    TRY(dmember_array(h, &q, ij, members([](const MemberConsts &c) -> const std::vector<double> & { return c.ca; }))); v.ca = q;
    TRY(dmember_array(h, &q, ij, members([](const MemberConsts &c) -> const std::vector<double> & { return c.pmeadj; }))); v.pmeadj = q;
    TRY(dupload(h, &q, h->mc[0].solfor)); v.solfor = q;
  }
  {
    // runoff gather lists in the reference's accumulation order: i outer, j inner (embm.f90:2955-2956)
    std::vector<std::vector<int>> src(ij);
    for (int i = 1; i <= I; i++)
      for (int j = 1; j <= J; j++)
        if (g.k1at(i, j) > K) {
          const int ir = h->mc[0].iroff[(i - 1) + I * (j - 1)], jr = h->mc[0].jroff[(i - 1) + I * (j - 1)];
          src[(ir - 1) + (size_t)I * (jr - 1)].push_back((i - 1) + I * (j - 1));
        }
    std::vector<int> ptr(ij + 1, 0), flat;
    for (size_t c2 = 0; c2 < ij; c2++) { ptr[c2] = (int)flat.size(); flat.insert(flat.end(), src[c2].begin(), src[c2].end()); }
    ptr[ij] = (int)flat.size();
    if (flat.empty()) flat.push_back(0);
    int *q;
    TRY(dupload(h, &q, ptr)); v.iroff_ptr = q;
    TRY(dupload(h, &q, flat)); v.iroff_src = q;
  }
  for (double **a : {&v.tq, &v.tq1, &v.tqa, &v.varice, &v.varice1}) TRY(dalloc(h, a, 2 * ij * MS));
  for (double **a : {&v.co2, &v.ch4, &v.n2o, &v.pptn, &v.evap, &v.evapsic, &v.fx0a, &v.fx0o, &v.fxsen, &v.fxlw, &v.fxsw, &v.fxplw,
                     &v.tice, &v.albice, &v.albedo, &v.latent_ocn, &v.sensible_ocn, &v.netsolar_ocn, &v.netlong_ocn, &v.evap_ocn,
                     &v.precip_ocn, &v.runoff_ocn, &v.runoff_land, &v.latent_atm, &v.sensible_atm, &v.netsolar_atm, &v.netlong_atm,
                     &v.evap_atm, &v.precip_atm, &v.dhght_sic, &v.dfrac_sic, &v.waterflux_ocn, &v.conductflux_ocn, &v.q_pa, &v.rq_pa})
    TRY(dalloc(h, a, ij * MS));
  {
    std::vector<double> t(2 * ij * MS), c1(ij * MS), c2v(ij * MS), c3(ij * MS);
    for (int m = 0; m < MS; m++) {
      const MemberConsts &c = h->mc[std::min(m, M - 1)];
      for (size_t q = 0; q < ij; q++) {
        t[q * MS + m] = c.tq0[0 + 2 * q];
        t[(ij + q) * MS + m] = c.tq0[1 + 2 * q];
        c1[q * MS + m] = c.p.radfor_scl_co2 * kCo20;
        c2v[q * MS + m] = c.p.radfor_scl_ch4 * kCh40;
        c3[q * MS + m] = c.p.radfor_scl_n2o * kN2o0;
      }
    }
    CUDA_OK(cudaMemcpy(v.tq, t.data(), t.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(v.tq1, t.data(), t.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(v.co2, c1.data(), c1.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(v.ch4, c2v.data(), c2v.size() * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(v.n2o, c3.data(), c3.size() * 8, cudaMemcpyHostToDevice));
  }
  const long long IJ = (long long)ij;
  reg_field(h, "tq", v.tq, {2, I, J}, {IJ, 1, I});
  reg_field(h, "tq1", v.tq1, {2, I, J}, {IJ, 1, I});
  reg_field(h, "tqa", v.tqa, {2, I, J}, {IJ, 1, I});
  reg_field(h, "varice", v.varice, {2, I, J}, {IJ, 1, I});
  reg_field(h, "varice1", v.varice1, {2, I, J}, {IJ, 1, I});
  struct { const char *n; double *d; } two[] = {
      {"co2", v.co2}, {"pptn", v.pptn}, {"evap", v.evap}, {"evapsic", v.evapsic}, {"fx0a", v.fx0a}, {"fx0o", v.fx0o},
      {"fxsen", v.fxsen}, {"fxlw", v.fxlw}, {"fxsw", v.fxsw}, {"fxplw", v.fxplw}, {"tice", v.tice}, {"temp_sic", v.tice},
      {"albice", v.albice}, {"albd_sic", v.albice}, {"albedo_ocn", v.albedo}, {"latent_ocn", v.latent_ocn},
      {"sensible_ocn", v.sensible_ocn}, {"netsolar_ocn", v.netsolar_ocn}, {"netlong_ocn", v.netlong_ocn},
      {"evap_ocn", v.evap_ocn}, {"precip_ocn", v.precip_ocn}, {"runoff_ocn", v.runoff_ocn}, {"runoff_land", v.runoff_land},
      {"latent_atm", v.latent_atm}, {"sensible_atm", v.sensible_atm}, {"netsolar_atm", v.netsolar_atm},
      {"netlong_atm", v.netlong_atm}, {"evap_atm", v.evap_atm}, {"precip_atm", v.precip_atm}, {"dhght_sic", v.dhght_sic},
      {"dfrac_sic", v.dfrac_sic}, {"waterflux_ocn", v.waterflux_ocn}, {"conductflux_ocn", v.conductflux_ocn}};
  for (auto &e : two) reg_field(h, e.n, e.d, {I, J}, {1, I});
#undef TRY
  return CG_OK;
}

// ------------------------------------------------------------------ one-member staging
static int ensure_stage(cg_handle *h, size_t n) {
  if (h->stage_n >= n) return CG_OK;
  void *q;
  CUDA_OK(cudaMalloc(&q, n * sizeof(double)));
  h->allocs.push_back(q);
  h->stage = (double *)q;
  h->stage_n = n;
  return CG_OK;
}

// ---- the water-column sweep and the coupling update in their two forms (see k_bg_cell) ----
static int bg_part_materialise(cg_handle *h, cudaStream_t s) {   // bio_part as the reference holds it (biogem.f90:2042-2043 applied)
  if (!h->part_pending) return 0;
  h->part_pending = false;
  return launch_bg_part_scale(h->dv, h->bgd, s);
}
static int bg_launch_sweep(cg_handle *h, cudaStream_t s, int fuse = 0) {
  if (h->bg_pd && !fuse && !h->pd_pending) {
    const int pend = h->part_pending ? 1 : 0;
    h->part_pending = false;
    h->pd_pending = true;
    return launch_bg_packets(h->dv, h->bgd, pend, s);
  }
  int n = bg_part_materialise(h, s);
  return n + launch_bg_sweep(h->dv, h->bgd, s, fuse);
}
static int bg_launch_apply(cg_handle *h, cudaStream_t s) {
  if (h->pd_pending) {
    h->pd_pending = false;
    h->part_pending = true;
    return launch_bg_cell(h->dv, h->bgd, s);
  }
  return launch_tc_apply_only(h->dv, s);
}
static FieldDesc *find_field(cg_handle *h, const char *name) {
  auto it = h->fields.find(name);
  if (it == h->fields.end()) return nullptr;
  // bio_part is read or written as the reference holds it: apply a rescaling the last coupling left pending (callers have joined
  // the side streams)
  if (h->part_pending && h->initialised && it->first == "bio_part") { activate(h); bg_part_materialise(h, h->stream); }
  // ts / ts1 alias the current ping-pong buffer
  if (it->first == "ts") it->second.d = h->dv.ts_cur;
  return &it->second;
}
extern "C" int64_t cg_field_size(cg_handle *h, const char *name) {
  if (!h || !name) return -1;
  FieldDesc *f = find_field(h, strcmp(name, "ts1") == 0 ? "ts" : name);
  return f ? f->count() : -1;
}
extern "C" int cg_sync_to_host(cg_handle *h, const char *name, int member, double *dst, int64_t n) {
  if (!h || !name || !dst || !h->initialised) return fail(CG_ERR_ARG, "cg_sync_to_host: bad argument");
  IO0(join_side(h));
  FieldDesc *f = find_field(h, strcmp(name, "ts1") == 0 ? "ts" : name);
  if (!f) return fail(CG_ERR_ARG, std::string("unknown field ") + name);
  if (member < 0 || member >= h->M || n != f->count()) return fail(CG_ERR_ARG, "cg_sync_to_host: member/size mismatch");
  activate(h);
  int rc = ensure_stage(h, (size_t)n);
  if (rc) return rc;
  k_gather_member<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(f->d, h->stage, member, h->MS, f->nd, f->dims[0], f->dims[1],
                                                                     f->dims[2], f->dims[3], f->strides[0], f->strides[1],
                                                                     f->strides[2], f->strides[3], n);
  CUDA_OK(cudaMemcpyAsync(dst, h->stage, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}
static int sync_from_host_lane(cg_handle *h, const char *name, int member, const double *src, int64_t n);
extern "C" int cg_sync_from_host(cg_handle *h, const char *name, int member, const double *src, int64_t n) {
  if (!h || member < 0 || member >= h->M) return fail(CG_ERR_ARG, "cg_sync_from_host: member out of range");
  int rc = sync_from_host_lane(h, name, member, src, n);
  // padding lanes mirror the last member (kernels run them unguarded)
  if (!rc && member == h->M - 1)
    for (int q = h->M; q < h->MS && !rc; q++) rc = sync_from_host_lane(h, name, q, src, n);
  return rc;
}
static int sync_from_host_lane(cg_handle *h, const char *name, int member, const double *src, int64_t n) {
  if (!h || !name || !src || !h->initialised) return fail(CG_ERR_ARG, "cg_sync_from_host: bad argument");
  IO0(join_side(h));
  h->spec_valid = false;
  h->tc_spec_valid = false;
  if (momentum_input(name)) IO0(drop_momentum(h));   // a momentum step computed ahead of time is stale: roll u1 back, redo it later
  const bool both = strcmp(name, "ts") == 0 || strcmp(name, "ts1") == 0;
  FieldDesc *f = find_field(h, both ? "ts" : name);
  if (!f) return fail(CG_ERR_ARG, std::string("unknown field ") + name);
  if (member < 0 || member >= h->MS || n != f->count()) return fail(CG_ERR_ARG, "cg_sync_from_host: member/size mismatch");
  activate(h);
  int rc = ensure_stage(h, (size_t)n);
  if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(h->stage, src, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  k_scatter_member<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(f->d, h->stage, member, h->MS, f->nd, f->dims[0], f->dims[1],
                                                                      f->dims[2], f->dims[3], f->strides[0], f->strides[1],
                                                                      f->strides[2], f->strides[3], n);
  if (both)  // ts and ts1 are one field on the device; keep the ping-pong partner's dry cells in line
    k_scatter_member<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->dv.ts_new, h->stage, member, h->MS, f->nd, f->dims[0],
                                                                        f->dims[1], f->dims[2], f->dims[3], f->strides[0],
                                                                        f->strides[1], f->strides[2], f->strides[3], n);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}
extern "C" int cg_sync_all_to_host(cg_handle *h, const char *name, double *dst, int64_t n) {
  if (!h || !name || !dst || !h->initialised) return fail(CG_ERR_ARG, "cg_sync_all_to_host: bad argument");
  IO0(join_side(h));
  FieldDesc *f = find_field(h, name);
  if (!f) return fail(CG_ERR_ARG, std::string("unknown field ") + name);
  if (n != f->count() * h->MS) return fail(CG_ERR_ARG, "cg_sync_all_to_host: size must be field_size*member_stride");
  CUDA_OK(cudaMemcpyAsync(dst, f->d, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}
extern "C" int cg_sync_all_from_host(cg_handle *h, const char *name, const double *src, int64_t n) {
  if (!h || !name || !src || !h->initialised) return fail(CG_ERR_ARG, "cg_sync_all_from_host: bad argument");
  IO0(join_side(h));
  if (momentum_input(name)) IO0(drop_momentum(h));
  h->spec_valid = false;
  h->tc_spec_valid = false;
  FieldDesc *f = find_field(h, name);
  if (!f) return fail(CG_ERR_ARG, std::string("unknown field ") + name);
  if (n != f->count() * h->MS) return fail(CG_ERR_ARG, "cg_sync_all_from_host: size must be field_size*member_stride");
  CUDA_OK(cudaMemcpyAsync(f->d, src, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  if (strcmp(name, "ts") == 0)   // the ping-pong partner's dry cells stay in line: device copy, the host data crosses PCIe once
    CUDA_OK(cudaMemcpyAsync(h->dv.ts_new, f->d, (size_t)n * 8, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}

// ---- wet-cell packed exchange of 3-D ocean fields
namespace cg {
// one block per (wet cell, inner index): MS contiguous doubles each way
__global__ void k_wet_pack(const double *__restrict__ field, double *__restrict__ packed, const int *__restrict__ wet, const int inner,
                           const int MS, const int to_packed, double *__restrict__ field2) {
  const size_t w = blockIdx.x;
  const int q = blockIdx.y;
  const size_t a = ((size_t)wet[w] * inner + q) * MS, b = (w * inner + q) * MS;
  for (int m = threadIdx.x; m < MS; m += blockDim.x) {
    if (to_packed) packed[b + m] = field[a + m];
    else {
      const double x = packed[b + m];
      const_cast<double *>(field)[a + m] = x;
      if (field2) field2[a + m] = x;
    }
  }
}
}  // namespace cg
static int wet_setup(cg_handle *h, const char *name, FieldDesc **fo, int *inner) {
  if (!h || !name || !h->initialised) return fail(CG_ERR_ARG, "wet exchange: bad argument");
  FieldDesc *f = find_field(h, name);
  if (!f) return fail(CG_ERR_ARG, std::string("unknown field ") + name);
  const int I = h->g.I, J = h->g.J, K = h->g.K;
  const int o = f->nd - 3;
  if (o < 0 || o > 1 || f->dims[o] != I || f->dims[o + 1] != J || f->dims[o + 2] != K)
    return fail(CG_ERR_ARG, std::string(name) + " is not a 3-D ocean field");
  const int in = o ? f->dims[0] : 1;
  if ((o && f->strides[0] != 1) || f->strides[o] != in || f->strides[o + 1] != (long long)in * I || f->strides[o + 2] != (long long)in * I * J)
    return fail(CG_ERR_ARG, std::string(name) + " is not stored cell-major");
  if (!h->d_wet3) {
    std::vector<int> wet;
    for (int k = 1; k <= K; k++)
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++)
          if (k >= h->g.k1at(i, j)) wet.push_back((int)cell3(I, J, i, j, k));
    h->n_wet3 = (int)wet.size();
    if (wet.empty()) wet.push_back(0);
    int rc = dupload(h, &h->d_wet3, wet);
    if (rc) return rc;
  }
  *fo = f;
  *inner = in;
  return CG_OK;
}
extern "C" int64_t cg_wet_size(cg_handle *h, const char *name) {
  FieldDesc *f; int in;
  if (wet_setup(h, name, &f, &in)) return -1;
  return (int64_t)h->n_wet3 * in;
}
static int wet_stage(cg_handle *h, size_t n) {
  if (h->wet_stage_n >= n) return CG_OK;
  if (h->wet_stage) cudaFree(h->wet_stage);
  h->wet_stage = nullptr; h->wet_stage_n = 0;
  CUDA_OK(cudaMalloc(&h->wet_stage, n * sizeof(double)));
  h->wet_stage_n = n;
  return CG_OK;
}
extern "C" int cg_sync_all_wet_to_host(cg_handle *h, const char *name, double *dst, int64_t n) {
  FieldDesc *f; int in;
  if (!dst) return fail(CG_ERR_ARG, "cg_sync_all_wet_to_host: bad argument");
  IO0(wet_setup(h, name, &f, &in));
  if (n != (int64_t)h->n_wet3 * in * h->MS) return fail(CG_ERR_ARG, "cg_sync_all_wet_to_host: size must be cg_wet_size * member_stride");
  IO0(join_side(h));
  activate(h);
  IO0(wet_stage(h, (size_t)n));
  if (h->n_wet3 > 0) k_wet_pack<<<dim3(h->n_wet3, in), 128, 0, h->stream>>>(f->d, h->wet_stage, h->d_wet3, in, h->MS, 1, nullptr);
  CUDA_OK(cudaMemcpyAsync(dst, h->wet_stage, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}
extern "C" int cg_sync_all_wet_from_host(cg_handle *h, const char *name, const double *src, int64_t n) {
  FieldDesc *f; int in;
  if (!src) return fail(CG_ERR_ARG, "cg_sync_all_wet_from_host: bad argument");
  IO0(wet_setup(h, name, &f, &in));
  if (n != (int64_t)h->n_wet3 * in * h->MS) return fail(CG_ERR_ARG, "cg_sync_all_wet_from_host: size must be cg_wet_size * member_stride");
  IO0(join_side(h));
  if (momentum_input(name)) IO0(drop_momentum(h));
  h->spec_valid = false;
  h->tc_spec_valid = false;
  activate(h);
  IO0(wet_stage(h, (size_t)n));
  CUDA_OK(cudaMemcpyAsync(h->wet_stage, src, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  // ts and ts1 are one field on the device: both ping-pong buffers take the wet cells
  double *second = strcmp(name, "ts") == 0 ? h->dv.ts_new : nullptr;
  if (h->n_wet3 > 0) k_wet_pack<<<dim3(h->n_wet3, in), 128, 0, h->stream>>>(f->d, h->wet_stage, h->d_wet3, in, h->MS, 0, second);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}

// ---- double-buffered state exchange -------------------------------------------------------------------------------------
// The coupling-interval exchange of a resident ensemble (all members of one field, device layout; 3-D ocean fields as wet cells)
// without stalling the compute stream on PCIe: an upload crosses into a device staging buffer on a copy stream while the model
// runs and is committed (unpacked into the field: a device-to-device pass) where the host program wants the new state; a
// download is packed into a staging buffer on the compute stream and crosses to the host on the copy stream while the next
// interval computes.  Host buffers must be page-locked for the copies to be asynchronous and must stay untouched until
// cg_exchange_wait.  wet != 0: the cg_sync_all_wet_* layout (cg_wet_size doubles per member), else cg_field_size.
static int check_async(cg_handle *h);
static int xchg_setup(cg_handle *h, const char *name, int wet, int64_t n, cg_handle::Xchg **out) {
  if (!h || !name || !h->initialised) return fail(CG_ERR_ARG, "cg_exchange: bad argument");
  FieldDesc *f = nullptr; int in = 1;
  if (wet) { IO0(wet_setup(h, name, &f, &in)); if (n != (int64_t)h->n_wet3 * in * h->MS) return fail(CG_ERR_ARG, "cg_exchange: size must be cg_wet_size * member_stride"); }
  else { f = find_field(h, name); if (!f) return fail(CG_ERR_ARG, std::string("unknown field ") + name); if (n != f->count() * h->MS) return fail(CG_ERR_ARG, "cg_exchange: size must be field_size * member_stride"); }
  cg_handle::Xchg &x = h->xchg[name];
  if (!x.up) {
    activate(h);
    if (!h->xstream) CUDA_OK(cudaStreamCreateWithFlags(&h->xstream, cudaStreamNonBlocking));
    CUDA_OK(cudaMalloc(&x.up, (size_t)n * 8));
    CUDA_OK(cudaMalloc(&x.down, (size_t)n * 8));
    CUDA_OK(cudaEventCreateWithFlags(&x.up_done, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&x.down_ready, cudaEventDisableTiming));
    x.n = (size_t)n; x.wet = wet != 0; x.inner = in;
  }
  if (x.n != (size_t)n || x.wet != (wet != 0)) return fail(CG_ERR_ARG, "cg_exchange: a field keeps the layout of its first exchange");
  *out = &x;
  return CG_OK;
}
extern "C" int cg_exchange_begin_upload(cg_handle *h, const char *name, int wet, const double *src, int64_t n) {
  cg_handle::Xchg *x;
  if (!src) return fail(CG_ERR_ARG, "cg_exchange_begin_upload: bad argument");
  IO0(xchg_setup(h, name, wet, n, &x));
  // the staging buffer may still feed the commit of the previous upload on the compute stream
  if (x->up_pending) return fail(CG_ERR_STATE, "cg_exchange_begin_upload: the previous upload of this field was not committed");
  CUDA_OK(cudaMemcpyAsync(x->up, src, (size_t)n * 8, cudaMemcpyHostToDevice, h->xstream));
  CUDA_OK(cudaEventRecord(x->up_done, h->xstream));
  x->up_pending = true;
  return CG_OK;
}
// commit the staged upload of `name` into field `target` (NULL: `name` itself; e.g. the state staged as "tq" also into "tq1")
static int xchg_commit(cg_handle *h, const char *name, const char *target, bool last) {
  auto it = h->xchg.find(name);
  if (it == h->xchg.end() || !it->second.up_pending) return fail(CG_ERR_STATE, "cg_exchange_commit_upload: nothing staged for this field");
  cg_handle::Xchg &x = it->second;
  const char *tn = target ? target : name;
  IO0(join_side(h));
  if (momentum_input(tn)) IO0(drop_momentum(h));
  h->spec_valid = false;
  h->tc_spec_valid = false;
  activate(h);
  CUDA_OK(cudaStreamWaitEvent(h->stream, x.up_done, 0));
  if (x.wet) {
    FieldDesc *f; int in;
    IO0(wet_setup(h, tn, &f, &in));
    if ((size_t)h->n_wet3 * in * h->MS != x.n) return fail(CG_ERR_ARG, "cg_exchange_commit_upload: target has another layout");
    double *second = strcmp(tn, "ts") == 0 ? h->dv.ts_new : nullptr;
    if (h->n_wet3 > 0) k_wet_pack<<<dim3(h->n_wet3, in), 128, 0, h->stream>>>(f->d, x.up, h->d_wet3, in, h->MS, 0, second);
  } else {
    FieldDesc *f = find_field(h, tn);
    if (!f || (size_t)f->count() * h->MS != x.n) return fail(CG_ERR_ARG, "cg_exchange_commit_upload: target has another layout");
    CUDA_OK(cudaMemcpyAsync(f->d, x.up, x.n * 8, cudaMemcpyDeviceToDevice, h->stream));
    if (strcmp(tn, "ts") == 0) CUDA_OK(cudaMemcpyAsync(h->dv.ts_new, x.up, x.n * 8, cudaMemcpyDeviceToDevice, h->stream));
  }
  if (last) {
    // the next upload into this staging buffer must not start before the compute stream has read it
    CUDA_OK(cudaEventRecord(x.down_ready, h->stream));
    CUDA_OK(cudaStreamWaitEvent(h->xstream, x.down_ready, 0));
    x.up_pending = false;
  }
  return check_async(h);
}
extern "C" int cg_exchange_commit_upload(cg_handle *h, const char *name, const char *also) {
  if (!h || !name || !h->initialised) return fail(CG_ERR_ARG, "cg_exchange_commit_upload: bad argument");
  if (also && *also) IO0(xchg_commit(h, name, also, false));
  return xchg_commit(h, name, nullptr, true);
}
extern "C" int cg_exchange_begin_download(cg_handle *h, const char *name, int wet, double *dst, int64_t n) {
  cg_handle::Xchg *x;
  if (!dst) return fail(CG_ERR_ARG, "cg_exchange_begin_download: bad argument");
  IO0(xchg_setup(h, name, wet, n, &x));
  IO0(join_side(h));
  activate(h);
  // the previous download out of this staging buffer is on the copy stream: in order behind it
  CUDA_OK(cudaEventRecord(x->down_ready, h->xstream));
  CUDA_OK(cudaStreamWaitEvent(h->stream, x->down_ready, 0));
  if (x->wet) {
    FieldDesc *f; int in;
    IO0(wet_setup(h, name, &f, &in));
    if (h->n_wet3 > 0) k_wet_pack<<<dim3(h->n_wet3, in), 128, 0, h->stream>>>(f->d, x->down, h->d_wet3, in, h->MS, 1, nullptr);
  } else {
    FieldDesc *f = find_field(h, name);
    CUDA_OK(cudaMemcpyAsync(x->down, f->d, x->n * 8, cudaMemcpyDeviceToDevice, h->stream));
  }
  CUDA_OK(cudaEventRecord(x->down_ready, h->stream));
  CUDA_OK(cudaStreamWaitEvent(h->xstream, x->down_ready, 0));
  CUDA_OK(cudaMemcpyAsync(dst, x->down, (size_t)n * 8, cudaMemcpyDeviceToHost, h->xstream));
  return check_async(h);
}
extern "C" int cg_exchange_wait(cg_handle *h) {   // every staged copy has crossed PCIe
  if (!h || !h->initialised) return fail(CG_ERR_ARG, "cg_exchange_wait: bad argument");
  if (h->xstream) CUDA_OK(cudaStreamSynchronize(h->xstream));
  return CG_OK;
}

extern "C" int64_t cg_const_size(cg_handle *h, const char *name) {
  if (!h || !name) return -1;
  auto it = h->hconst.find(name);
  if (it != h->hconst.end()) return (int64_t)it->second.size();
  auto jt = h->hiconst.find(name);
  if (jt != h->hiconst.end()) return (int64_t)jt->second.size();
  return -1;
}
extern "C" int cg_get_const(cg_handle *h, const char *name, int member, double *dst, int64_t n) {
  if (!h || !name || !dst) return fail(CG_ERR_ARG, "cg_get_const: bad argument");
  (void)member;
  auto it = h->hconst.find(name);
  if (it == h->hconst.end()) return fail(CG_ERR_ARG, std::string("unknown constant ") + name);
  if (n != (int64_t)it->second.size()) return fail(CG_ERR_ARG, "cg_get_const: size mismatch");
  memcpy(dst, it->second.data(), (size_t)n * 8);
  return CG_OK;
}
extern "C" int cg_get_iconst(cg_handle *h, const char *name, int32_t *dst, int64_t n) {
  if (!h || !name || !dst) return fail(CG_ERR_ARG, "cg_get_iconst: bad argument");
  auto it = h->hiconst.find(name);
  if (it == h->hiconst.end()) return fail(CG_ERR_ARG, std::string("unknown constant ") + name);
  if (n != (int64_t)it->second.size()) return fail(CG_ERR_ARG, "cg_get_iconst: size mismatch");
  for (int64_t q = 0; q < n; q++) dst[q] = it->second[q];
  return CG_OK;
}
extern "C" int cg_get_dims(cg_handle *h, int32_t dims[8]) {
  if (!h || !dims) return fail(CG_ERR_ARG, "cg_get_dims: bad argument");
  dims[0] = h->g.I; dims[1] = h->g.J; dims[2] = h->g.K; dims[3] = h->g.L; dims[4] = h->M; dims[5] = h->MS;
  dims[6] = h->g.nyear; dims[7] = h->base.ndta;
  return CG_OK;
}

// ------------------------------------------------------------------ steps
struct ProfScope {
  cg_handle *h; const char *fam; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(cg_handle *h_, const char *f) : h(h_), fam(f) {
    if (h->profile) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, h->stream); }
  }
  void done(int n) {
    h->launches += n;
    if (h->profile) {
      cudaEventRecord(b, h->stream);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      h->prof[fam].ms += ms;
      h->prof[fam].n += n;
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  }
};

static int do_surflux(cg_handle *h) {
  { ProfScope ps(h, "surflux"); launch_step_begin(h->dv, h->stream); int n = launch_surflux(h->dv, h->d_meantemp, h->need_mean, h->stream); ps.done(n + 1); }
  h->istep_ocn++;
  return CG_OK;
}
static int do_embm(cg_handle *h, int nsteps) {
  ProfScope ps(h, "embm");
  int n = launch_embm(h->dv, nsteps, h->stream);
  if (n < 0) return fail(CG_ERR_CONFIG, "EMBM grid too large for the block-resident kernel");
  ps.done(n);
  h->istep_atm += nsteps;
  return CG_OK;
}
static int do_seaice(cg_handle *h) {
  ProfScope ps(h, "seaice");
  // the sea-ice step advects with the surface velocities exported by the last step_goldstein: a snapshot, taken here
  // unless the forked cycle already took it before the momentum branch started
  int n = 0;
  if (!h->forked && !h->usnap_valid) { launch_usnap(h->dv, h->stream); n++; }
  ps.done(n + launch_seaice(h->dv, h->stream));
  h->istep_sic++;
  return CG_OK;
}
static void do_tstepo(cg_handle *h) {
  if (h->dv.imld) {   // Kraus-Turner mixed-layer scheme (tstepo, goldstein.f90:2294-2390) around the generic flux / convection kernels
    { ProfScope ps(h, "co"); ps.done(launch_mld_pre(h->dv, h->stream)); }
    { ProfScope ps(h, "tstepo_flux"); ps.done(h->variant ? launch_tstepo_flux_fast(h->dv, h->stream) : launch_tstepo_flux_strict(h->dv, h->stream)); }
    ProfScope ps(h, "co");
    int n = launch_mld_save(h->dv, h->stream);
    n += h->variant ? launch_co_fast(h->dv, h->stream) : launch_co_strict(h->dv, h->stream);
    n += launch_mld_kt(h->dv, h->stream);
    std::swap(h->dv.ts_cur, h->dv.ts_new);
    if (h->dv.sst) n += launch_sst(h->dv, h->stream);
    ps.done(n);
    return;
  }
  if (h->variant == 2) {  // fused column kernel: flux + convection + SST export in one pass (compiled shapes only)
    ProfScope ps(h, "tstepo_flux");
    const int n = launch_tstep_col(h->dv, h->stream);
    if (n > 0) {
      ps.done(n);
      std::swap(h->dv.ts_cur, h->dv.ts_new);
      return;
    }
    ps.done(0);
  }
  { ProfScope ps(h, "tstepo_flux"); ps.done(h->variant ? launch_tstepo_flux_fast(h->dv, h->stream) : launch_tstepo_flux_strict(h->dv, h->stream)); }
  { ProfScope ps(h, "co"); ps.done(h->variant ? launch_co_fast(h->dv, h->stream) : launch_co_strict(h->dv, h->stream)); }
  std::swap(h->dv.ts_cur, h->dv.ts_new);
  if (h->dv.sst) { ProfScope ps(h, "co"); ps.done(launch_sst(h->dv, h->stream)); }
}
static void do_gold_pre(cg_handle *h) {
  ProfScope ps(h, "momentum");
  launch_hosing(h->dv, h->stream);
  ps.done(1 + launch_gold_pre(h->dv, h->stream));
}
// s3 != nullptr (graph capture only): the baroclinic shear integral runs on s3 next to the barotropic solve on s
static int do_momentum(cg_handle *h, cudaStream_t s, cudaStream_t s3 = nullptr) {
  ProfScope ps(h, "momentum");
  int n = 0;
  if (s3) {
    CUDA_OK(cudaEventRecord(h->evFork3, s));
    CUDA_OK(cudaStreamWaitEvent(s3, h->evFork3, 0));
    n += launch_velc1(h->dv, s3);
    CUDA_OK(cudaEventRecord(h->evJoin3, s3));
  }
  n += launch_momentum(h->dv, h->variant, h->d_bf, h->d_bb, h->d_rd, h->d_bk, s);
  if (s3) CUDA_OK(cudaStreamWaitEvent(s, h->evJoin3, 0));
  else n += launch_velc1(h->dv, s);
  n += launch_velc2(h->dv, s);
  ps.done(n);
  return CG_OK;
}
static int bg_join(cg_handle *h);
// Snapshot, on the caller's stream, of what the BIOGEM / ATCHEM block reads from the physics at the reference's call time: the
// sea-ice cover (biogem_climate) and the air temperature / humidity (cpl_comp_EMBM).  The block itself may run on its own
// stream while the caller's stream is already in the next cycle's surflux / EMBM / sea-ice steps.  The previous ATCHEM step
// must have consumed the previous snapshot first (it finished a cycle ago; the wait is for ordering, not time).
static int bg_stage_inputs(cg_handle *h, cudaStream_t s) {
  if (h->atchem_ev_pending) { CUDA_OK(cudaStreamWaitEvent(s, h->evAtchem, 0)); h->atchem_ev_pending = false; }
  h->launches += launch_bg_stage_seaice(h->dv, h->bgd, s);
  return CG_OK;
}
static int do_goldstein(cg_handle *h) {
  h->usnap_valid = false;
  do_gold_pre(h);
  if (h->mom_ready) {   // computed on stream2 since the surflux call of this cycle
    if (h->mom_pending) { CUDA_OK(cudaStreamWaitEvent(h->stream, h->evJoin, 0)); h->mom_pending = false; }
    h->mom_ready = false;
  } else {
    int rc = do_momentum(h, h->stream);
    if (rc) return rc;
  }
  { int rc = bg_join(h); if (rc) return rc; }   // tstepo reads the ts an asynchronous tracer coupling rewrote
  do_tstepo(h);
  return CG_OK;
}
// per-module path: start the momentum step of this cycle now (it needs rho of the previous tracer step only)
static int eager_momentum(cg_handle *h) {
  if (!h->eager || h->profile || h->mom_ready || getenv("CG_NOEAGER")) return CG_OK;
  {
    const size_t n = (size_t)2 * h->g.I * h->g.J * h->g.K * h->MS;
    if (!h->u1_snap) IO0(dalloc(h, &h->u1_snap, n, false));
    CUDA_OK(cudaMemcpyAsync(h->u1_snap, h->dv.u1, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  launch_usnap(h->dv, h->stream);
  h->launches++;
  CUDA_OK(cudaEventRecord(h->evFork, h->stream));
  CUDA_OK(cudaStreamWaitEvent(h->stream2, h->evFork, 0));
  int rc = do_momentum(h, h->stream2, h->stream3);
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(h->evJoin, h->stream2));
  h->mom_pending = true;
  h->mom_ready = true;
  h->usnap_valid = true;
  return CG_OK;
}
// per-module path: run one BIOGEM / ATCHEM entry point on stream4, ordered after everything issued so far on the main
// stream; the next tracer step (or any host access) joins.  RAII: restores h->stream and records the completion event.
struct BgAsyncScope {
  cg_handle *h; cudaStream_t save = nullptr; bool on = false, tail = false;
  explicit BgAsyncScope(cg_handle *h_, bool allow, bool tail_ = false) : h(h_), tail(tail_) {
    if (!allow || !h->eager || h->profile || !h->bg.on || h->stream == h->stream4 || getenv("CG_NOEAGER")) return;
    if (cudaEventRecord(h->evT, h->stream) != cudaSuccess || cudaStreamWaitEvent(h->stream4, h->evT, 0) != cudaSuccess) return;
    save = h->stream;
    h->stream = h->stream4;
    on = true;
  }
  ~BgAsyncScope() {
    if (!on) return;
    // tail = ATCHEM: nothing on the physics side reads the atmosphere tracers, only a host access has to wait for it
    cudaEventRecord(tail ? h->evBGtail : h->evBG, h->stream4);
    h->stream = save;
    if (tail) h->bg_tail_pending = true; else h->bg_pending = true;
  }
};
static int check_async(cg_handle *h) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CG_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  (void)h;
  return CG_OK;
}
// one NVTX range per C-ABI entry point (visible in Nsight Systems / Compute timelines; a no-op without a tool attached)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define CG_RANGE() NvtxRange nvtx_range_(__func__)
#define READY(h) do { if (!(h) || !(h)->initialised) return fail(CG_ERR_STATE, "handle not initialised"); if ((h)->tracer_only) return fail(CG_ERR_STATE, "tracer-only handle"); activate(h); } while (0)

static int put(cg_handle *h, const char *name, const double *src) {
  return src ? cg_sync_from_host(h, name, h->io_member, src, cg_field_size(h, name)) : CG_OK;
}
static int get(cg_handle *h, const char *name, double *dst) {
  return dst ? cg_sync_to_host(h, name, h->io_member, dst, cg_field_size(h, name)) : CG_OK;
}
#define IO(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

static bool lazy_ok(const cg_handle *h);
static int lazy_cycle(cg_handle *h);
extern "C" int cg_surflux_step(cg_handle *h, int istep, const cg_surflux_io *io) {
  CG_RANGE();
  READY(h);
  if (h->lazy_stage) IO(flush_lazy(h));   // out of pattern: replay what was noted
  if (istep != h->istep_ocn + 1) {  // the host owns the step counter; keep the device copy in line
    const int v0 = istep - 1;
    CUDA_OK(cudaMemcpyAsync(h->dv.istep_ocn, &v0, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->istep_ocn = v0;
  }
  if (io && io->qstar_atm) {
    // surf_qstar_atm is INOUT: field 2 of tq
    std::vector<double> t(cg_field_size(h, "tq"));
    IO(cg_sync_to_host(h, "tq", h->io_member, t.data(), (int64_t)t.size()));
    for (size_t q = 0; q < t.size() / 2; q++) t[1 + 2 * q] = io->qstar_atm[q];
    IO(cg_sync_from_host(h, "tq", h->io_member, t.data(), (int64_t)t.size()));
  }
  if (!io && lazy_ok(h)) { h->lazy_stage = 1; return CG_OK; }
  IO(eager_momentum(h));
  IO(do_surflux(h));
  IO(check_async(h));
  if (io) {
    IO(get(h, "albedo_ocn", io->albedo_ocn)); IO(get(h, "latent_ocn", io->latent_ocn)); IO(get(h, "sensible_ocn", io->sensible_ocn));
    IO(get(h, "netsolar_ocn", io->netsolar_ocn)); IO(get(h, "netlong_ocn", io->netlong_ocn)); IO(get(h, "evap_ocn", io->evap_ocn));
    IO(get(h, "precip_ocn", io->precip_ocn)); IO(get(h, "runoff_ocn", io->runoff_ocn)); IO(get(h, "runoff_land", io->runoff_land));
    IO(get(h, "latent_atm", io->latent_atm)); IO(get(h, "sensible_atm", io->sensible_atm)); IO(get(h, "netsolar_atm", io->netsolar_atm));
    IO(get(h, "netlong_atm", io->netlong_atm)); IO(get(h, "evap_atm", io->evap_atm)); IO(get(h, "precip_atm", io->precip_atm));
    IO(get(h, "dhght_sic", io->dhght_sic)); IO(get(h, "dfrac_sic", io->dfrac_sic)); IO(get(h, "temp_sic", io->temp_sic));
    IO(get(h, "albd_sic", io->albd_sic));
    if (io->qstar_atm) {
      std::vector<double> t(cg_field_size(h, "tq"));
      IO(cg_sync_to_host(h, "tq", h->io_member, t.data(), (int64_t)t.size()));
      for (size_t q = 0; q < t.size() / 2; q++) io->qstar_atm[q] = t[1 + 2 * q];
    }
  }
  return CG_OK;
}

extern "C" int cg_embm_step(cg_handle *h, int istep, const cg_embm_io *io) {
  CG_RANGE();
  READY(h);
  (void)istep;
  if (!io && h->lazy_stage >= 1 && h->lazy_stage <= h->base.kocn_loop) { h->lazy_stage++; return CG_OK; }
  if (h->lazy_stage) IO(flush_lazy(h));
  IO(do_embm(h, 1));
  IO(check_async(h));
  if (io && (io->tstar_atm || io->qstar_atm)) {
    std::vector<double> t(cg_field_size(h, "tq"));
    IO(cg_sync_to_host(h, "tq", h->io_member, t.data(), (int64_t)t.size()));
    for (size_t q = 0; q < t.size() / 2; q++) {
      if (io->tstar_atm) io->tstar_atm[q] = t[2 * q];
      if (io->qstar_atm) io->qstar_atm[q] = t[1 + 2 * q];
    }
  }
  return CG_OK;
}

extern "C" int cg_seaice_step(cg_handle *h, int istep, const cg_seaice_io *io) {
  CG_RANGE();
  READY(h);
  (void)istep;
  if (!io && h->lazy_stage == h->base.kocn_loop + 1) { h->lazy_stage++; return CG_OK; }
  if (h->lazy_stage) IO(flush_lazy(h));
  IO(do_seaice(h));
  IO(check_async(h));
  if (io) {
    if (io->hght_sic || io->frac_sic) {
      std::vector<double> t(cg_field_size(h, "varice"));
      IO(cg_sync_to_host(h, "varice", h->io_member, t.data(), (int64_t)t.size()));
      for (size_t q = 0; q < t.size() / 2; q++) {
        if (io->hght_sic) io->hght_sic[q] = t[2 * q];
        if (io->frac_sic) io->frac_sic[q] = t[1 + 2 * q];
      }
    }
    IO(get(h, "waterflux_ocn", io->waterflux_ocn));
    IO(get(h, "conductflux_ocn", io->conductflux_ocn));
  }
  return CG_OK;
}

// go_mldta = REAL(-5000.0 * mld) (goldstein.f90:449): mixed-layer depth in metres below the surface as handed to BIOGEM
extern "C" int cg_goldstein_mldta(cg_handle *h, int member, double *go_mldta) {
  CG_RANGE();
  READY(h);
  if (!go_mldta || member < 0 || member >= h->M) return fail(CG_ERR_ARG, "cg_goldstein_mldta: bad argument");
  const size_t ij = (size_t)h->g.I * h->g.J;
  if (!h->dv.imld) {   // mld stays 0 without the scheme
    for (size_t q = 0; q < ij; q++) go_mldta[q] = -5000.0 * 0.0;
    return CG_OK;
  }
  IO(cg_sync_to_host(h, "mld", member, go_mldta, (int64_t)ij));
  for (size_t q = 0; q < ij; q++) go_mldta[q] = -5000.0 * go_mldta[q];
  return CG_OK;
}

extern "C" int cg_goldstein_step(cg_handle *h, int istep, const cg_goldstein_io *io) {
  CG_RANGE();
  READY(h);
  (void)istep;
  if (!io && h->lazy_stage == h->base.kocn_loop + 2 && lazy_ok(h)) return lazy_cycle(h);
  if (h->lazy_stage) IO(flush_lazy(h));
  if (io) { IO(put(h, "ts", io->go_ts)); IO(put(h, "cost", io->go_cost)); }
  IO(do_goldstein(h));
  IO(check_async(h));
  if (io) {
    const int I = h->g.I, J = h->g.J, K = h->g.K, L = h->g.L;
    IO(get(h, "ts", io->go_ts)); IO(get(h, "u", io->go_u)); IO(get(h, "rho", io->go_rho)); IO(get(h, "cost", io->go_cost));
    if (io->go_psi) {
      IO(get(h, "psi", io->go_psi));
      for (int q = 0; q < (I + 1) * (J + 1); q++) io->go_psi[q] = 1592.5 * io->go_psi[q];  // goldstein.f90:450
    }
    if (io->tstar_ocn || io->sstar_ocn || io->ustar_ocn || io->vstar_ocn || io->albedo_ocn || io->test_energy_ocean ||
        io->test_water_ocean) {
      std::vector<double> ts((size_t)L * I * J * K), u((size_t)3 * I * J * K);
      IO(cg_sync_to_host(h, "ts", h->io_member, ts.data(), (int64_t)ts.size()));
      IO(cg_sync_to_host(h, "u", h->io_member, u.data(), (int64_t)u.size()));
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++) {
          const size_t c2 = (size_t)(i - 1) + (size_t)I * (j - 1), c3 = c2 + (size_t)I * J * (K - 1);
          const bool wet = h->g.k1at(i, j) <= K;
          if (io->ustar_ocn) io->ustar_ocn[c2] = u[0 + 3 * c3];
          if (io->vstar_ocn) io->vstar_ocn[c2] = u[1 + 3 * c3];
          if (io->tstar_ocn) io->tstar_ocn[c2] = wet ? ts[0 + (size_t)L * c3] : 0.0;
          if (io->sstar_ocn) io->sstar_ocn[c2] = wet ? ts[1 + (size_t)L * c3] : 0.0;
          if (io->albedo_ocn) io->albedo_ocn[c2] = wet ? h->mp[h->io_member].albocn : 0.0;
        }
      if (io->test_energy_ocean || io->test_water_ocean) {
        // goldstein.f90:458-478 (relative to the initial state, :79-96), host-side at output intervals only
        const Grid &g = h->g;
        auto tot = [&](const std::vector<double> &a, int l, double sgn) {
          double t = 0.0;
          for (int k = 1; k <= K; k++)
            for (int j = 1; j <= J; j++) {
              double s = 0.0;
              for (int i = 1; i <= I; i++) s = s + a[l + (size_t)L * ((size_t)(i - 1) + (size_t)I * ((j - 1) + (size_t)J * (k - 1)))];
              t = t + sgn * s * g.dz[k] * g.ds[j];
            }
          return t;
        };
        auto tot0 = [&](int l, double sgn) {
          double t = 0.0;
          const std::vector<double> &a = h->mc[h->io_member].ts0;
          for (int k = 1; k <= K; k++)
            for (int j = 1; j <= J; j++)
              for (int i = 1; i <= I; i++)
                t = t + sgn * a[l + (size_t)L * ((size_t)(i - 1) + (size_t)I * ((j - 1) + (size_t)J * (k - 1)))] * g.dz[k] * g.ds[j];
          return t;
        };
        const double vsc = g.dphi * kRsc * kRsc, saln0 = h->mp[h->io_member].saln0;
        if (io->test_energy_ocean)
          *io->test_energy_ocean = tot(ts, 0, 1.0) * vsc * kDsc * kRh0sc * kCpsc - tot0(0, 1.0) * vsc * kDsc * kRh0sc * kCpsc;
        if (io->test_water_ocean)
          *io->test_water_ocean = kM2mm * tot(ts, 1, -1.0) * vsc * kDsc / saln0 - kM2mm * tot0(1, -1.0) * vsc * kDsc / saln0;
      }
    }
  }
  return CG_OK;
}

// ---- BIOGEM / ATCHEM -------------------------------------------------------------------------------------------
#define BGREADY(h) do { READY(h); if (!(h)->bg.on) return fail(CG_ERR_CONFIG, "flag_biogem is off in this job"); if ((h)->lazy_stage) IO(flush_lazy(h)); } while (0)
static long long nint_ll(double x) { return (long long)(x >= 0 ? std::floor(x + 0.5) : -std::floor(-x + 0.5)); }

// biogem_forcing(genie_clock), biogem.f90:2083-2127: time-interpolated restoring targets (host scalars, bit-exact)
extern "C" int cg_biogem_forcing(cg_handle *h, int64_t genie_clock_ms) {
  CG_RANGE();
  BGREADY(h);
  bg_forcing(&h->bg, (long long)genie_clock_ms, &h->bgd);
  return CG_OK;
}
// step_biogem(dts, genie_clock, ...), biogem.f90:528-547.  The interface arrays stay on the device
// ("sfcocn1", "sfxsed1", "sfxsumatm", "focnatm" fields); cpl_flux_ocnatm (atchem.f90:306-320) is fused into the kernel.
extern "C" int cg_biogem_step(cg_handle *h, double dts, int64_t genie_clock_ms) {
  CG_RANGE();
  BGREADY(h);
  if (dts != h->bgd.dts) return fail(CG_ERR_ARG, "cg_biogem_step: dts differs from conv_kocn_kbiogem*kocn_loop*genie_timestep");
  const double t = h->bg.t_runtime - (double)genie_clock_ms / (1000.0 * kBgYrS);
  if (!h->bg_go) return CG_OK;   // par_misc_t_go (biogem.f90:1851-1853)
  BgAsyncScope as(h, true);
  {
    ProfScope ps(h, "biogem");
    if (as.on && h->spec_valid && h->spec_clock == (long long)genie_clock_ms) ps.done(bg_launch_sweep(h, h->stream));
    else { const int n = bg_part_materialise(h, h->stream); ps.done(n + launch_bg_step(h->dv, h->bgd, 0, 0, h->stream)); }
  }
  h->spec_valid = false;
  h->bg_last_clock = (long long)genie_clock_ms;
  h->bg_stage = 1;
  if (t < kBgNullSmall) h->bg_go = false;
  return check_async(h);
}
// cpl_flux_ocnatm_wrapper (genie_loop_wrappers.f90:178-183): already applied by cg_biogem_step
extern "C" int cg_cpl_flux_ocnatm(cg_handle *h) { BGREADY(h); return CG_OK; }
// The SEDGEM / ROKGEM coupler calls genie.f90 makes after every BIOGEM step whether or not those modules run
// (genie.f90:413-427).  The interface arrays stay on the device ("sfxsumsed", "sfcsumocn", "sfxsumrok1"
// fields); they follow cg_biogem_step on its stream.  cg_run does not issue them: without SEDGEM nothing reads the sums.
// The stream these calls run on waits for every piece of work still in flight on a side stream (the flags stay set: the
// main stream joins later as it would have), so that sfxsed1 / sfcocn1 are complete whichever schedule produced them.
static int side_wait(cg_handle *h) {
  if (h->mom_pending) CUDA_OK(cudaStreamWaitEvent(h->stream, h->evJoin, 0));
  if (h->bg_pending) CUDA_OK(cudaStreamWaitEvent(h->stream, h->evBG, 0));
  if (h->bg_tail_pending) CUDA_OK(cudaStreamWaitEvent(h->stream, h->evBGtail, 0));
  if (h->tc_old_pending) CUDA_OK(cudaStreamWaitEvent(h->stream, h->evTcOld, 0));
  return CG_OK;
}
// The ocean / atmosphere part of diag_biogem_timeseries (biogem.f90:2703-3159): one BIOGEM step's contribution to the window
// integrals int_t_sig, int_ocn_tot_M(_sur)_sig, int_ocn_sig, int_ocn_sur_sig, int_ocn_ben_sig, int_ocnatm_sig (:2851-2917,
// :3082), taken on the device behind the BIOGEM step.  The caller keeps the reference's save-window logic (:2760-2769, the
// list of biogem_save_sig.dat) and reads the integrals as field "bg_sig" = [int_t_sig, tot_M, tot_M_sur, ocn(L), ocn_sur(L),
// ocn_ben(L), ocnatm(LA)] when a window closes; cg_biogem_sig_reset = sub_init_int_timeseries (biogem_data.f90:964-1007).
// ben_Dmin = par_data_save_ben_Dmin (benthic mask: bottom cells whose floor lies deeper).
extern "C" int cg_biogem_sig_update(cg_handle *h, double dts, double ben_Dmin) {
  CG_RANGE();
  BGREADY(h);
  const Grid &g = h->g;
  const int I = g.I, J = g.J, K = g.K;
  if (ben_Dmin != h->sig_ben_Dmin) {
    std::vector<double> w((size_t)I * J, 0.0);
    double tot = 0.0;
    for (int j = 1; j <= J; j++)
      for (int i = 1; i <= I; i++) {
        const int k1 = g.k1at(i, j);
        if (k1 > K) continue;
        double Dbot = 0.0;                                  // phys_ocn(ipo_Dbot,i,j,k1) = SUM(dsc*dz(k1:n_k)), biogem_data.f90:1123
        for (int k = k1; k <= K; k++) Dbot = Dbot + kDsc * g.dz[k];
        if (Dbot > ben_Dmin) w[cell2(I, i, j)] = 2.0 * kBgPi * (kBgREarth * kBgREarth) * (1.0 / I) * (g.sv[j] - g.sv[j - 1]);
      }
    for (size_t c = 0; c < w.size(); c++) tot = tot + w[c];
    IO(join_side(h));
    CUDA_OK(cudaMemcpy(h->sig_w_ben, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->sig.rtot_A_ben = tot > kBgNullSmall ? 1.0 / tot : 0.0;
    h->sig_ben_Dmin = ben_Dmin;
  }
  // extended integrals: what they take from the physics (velocities, sea-ice thickness) is read NOW, on the caller's stream -- the
  // next cycle's momentum and sea-ice steps may run before the BIOGEM stream gets to the sums.  (The previous call's sums have
  // finished: the caller's stream has joined the BIOGEM stream at least once per ocean step since.)
  if (h->sig.acc2) h->launches += launch_bg_sig2_stage(h->dv, h->sig, h->sig2_dz, h->sig2_cv, h->g.dphi, h->stream);
  // everything the sums read is BIOGEM's own state (ocn, cell masses, the sea-ice snapshot, sfcatm1 incl. the air temperature
  // and humidity rows cpl_comp_EMBM filled at the last block): they run on the BIOGEM stream
  BgAsyncScope as(h, true);
  IO(side_wait(h));
  ProfScope ps(h, "biogem");
  int n = launch_bg_sig(h->dv, h->bgd, h->sig, dts / kBgYrS, h->stream);
  if (h->sig.acc2) n += launch_bg_sig2(h->dv, h->bgd, h->sig, dts / kBgYrS, h->stream);
  ps.done(n);
  return check_async(h);
}
// The flux, export and "misc" integrals of diag_biogem_timeseries on the device as well (field "bg_sig2", layout in
// include/cgenie_b200.h): from this call on step_biogem keeps sfxatm1 and the export through the base of the surface layer, and
// cg_biogem_sig_update also accumulates int_misc_seaice_sig / _th / _vol (biogem.f90:2926-2937), the extrema of the overturning
// stream functions (:2938-2945), int_misc_SLT_sig (:2946-2964), int_fexport_sig (:2870-2876), int_focnatm_sig (:2877-2883) and
// int_diag_airsea_sig (:3058-3062).  Call it before the first BIOGEM step whose window integrals are wanted.
extern "C" int cg_biogem_sig_extended(cg_handle *h) {
  CG_RANGE();
  BGREADY(h);
  if (h->sig.acc2) return CG_OK;
  const Grid &g = h->g;
  const int I = g.I, J = g.J, K = g.K, LS = h->bg.LS, LA = h->bg.LA, MS = h->MS;
  const size_t ij = (size_t)I * J;
  IO(join_side(h));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->spec_valid = false;          // a surface part issued ahead would not have kept its sfxatm1
  const int nq = kSig2Head + LS + 2 * LA;
  IO0(dalloc(h, &h->bgd.sfxatm1, (size_t)LA * ij * MS));
  IO0(dalloc(h, &h->bgd.settle_sur, ij * LS * MS));
  IO0(dalloc(h, &h->sig.raw2, (size_t)(4 + LS + 2 * LA) * MS));
  IO0(dalloc(h, &h->sig.opsi_stage, (size_t)4 * MS));
  IO0(dalloc(h, &h->sig.th_stage, ij * MS));
  std::vector<int> ias(g.ias.begin(), g.ias.end()), iaf(g.iaf.begin(), g.iaf.end());
  ias.resize(J + 2, 1); iaf.resize(J + 2, 0);
  int *qi;
  IO0(dupload(h, &qi, ias)); h->sig.ias = qi;
  IO0(dupload(h, &qi, iaf)); h->sig.iaf = qi;
  std::vector<double> dz(g.dz.begin(), g.dz.begin() + K + 1), cv(g.cv.begin(), g.cv.begin() + J + 1);
  IO0(dupload(h, &h->sig2_dz, dz));
  IO0(dupload(h, &h->sig2_cv, cv));
  h->sig.jsf = g.jsf; h->sig.LS = LS;
  double land = 0.0;   // loc_tot_A of :2950-2958: i outer, j inner
  for (int i = 1; i <= I; i++)
    for (int j = 1; j <= J; j++)
      if (K < g.k1at(i, j)) land = land + 2.0 * kBgPi * (kBgREarth * kBgREarth) * (1.0 / I) * (g.sv[j] - g.sv[j - 1]);
  h->sig.land_A = land;
  double *acc2;
  IO0(dalloc(h, &acc2, (size_t)nq * MS));
  reg_field(h, "bg_sig2", acc2, {nq}, {1});
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->sig.acc2 = acc2;
  return CG_OK;
}
// diag_biogem_timeslice (biogem.f90:2421-2699), its arithmetic: the 3-D carbonate re-solve (:2478-2567) and the growth of the
// window integrals int_ocn / int_bio_part / int_carb / int_carbconst / int_carbisor / int_t _timeslice (:2572-2579) on the
// device.  Call it where genie.f90 calls diag_biogem_timeslice_wrapper (:391-395: behind cg_biogem_climate, ahead of the
// time-series diagnostic and ATCHEM) on the BIOGEM steps inside a save window; the window bookkeeping (:2460-2470, 2627-2696)
// and the netCDF writer stay with the host, which reads the integrals as fields "sl_ocn" (maxl,i,j,k), "sl_part", "sl_carb"
// (10,...), "sl_carbconst" (17,...), "sl_carbisor" (8,...), "sl_t" (1) and divides by sl_t as sub_save_netcdf_3d does.
static int slice_alloc(cg_handle *h) {
  if (h->slice.ocn) return CG_OK;
  const size_t ijk = (size_t)h->g.I * h->g.J * h->g.K, MS = h->MS;
  const int I = h->g.I, J = h->g.J, K = h->g.K, L = h->g.L, LS = h->bg.LS;
  IO0(dalloc(h, &h->slice.ocn, ijk * L * MS));
  IO0(dalloc(h, &h->slice.part, ijk * LS * MS));
  IO0(dalloc(h, &h->slice.carb, ijk * kSlCarb * MS));
  IO0(dalloc(h, &h->slice.cc, ijk * kSlCC * MS));
  IO0(dalloc(h, &h->slice.iso, ijk * kSlIso * MS));
  IO0(dalloc(h, &h->slice.t, MS));
  CUDA_OK(cudaStreamSynchronize(h->stream));   // the zero fills
  const long long ij = (long long)I * J;
  reg_field(h, "sl_ocn", h->slice.ocn, {L, I, J, K}, {1, L, (long long)L * I, (long long)L * ij});
  reg_field(h, "sl_part", h->slice.part, {LS, I, J, K}, {1, LS, (long long)LS * I, (long long)LS * ij});
  reg_field(h, "sl_carb", h->slice.carb, {kSlCarb, I, J, K}, {1, kSlCarb, (long long)kSlCarb * I, (long long)kSlCarb * ij});
  reg_field(h, "sl_carbconst", h->slice.cc, {kSlCC, I, J, K}, {1, kSlCC, (long long)kSlCC * I, (long long)kSlCC * ij});
  reg_field(h, "sl_carbisor", h->slice.iso, {kSlIso, I, J, K}, {1, kSlIso, (long long)kSlIso * I, (long long)kSlIso * ij});
  reg_field(h, "sl_t", h->slice.t, {1}, {1});
  return CG_OK;
}
extern "C" int cg_biogem_slice_update(cg_handle *h, double dts) {
  CG_RANGE();
  BGREADY(h);
  IO(join_side(h));            // allocation + field registration happen on the main stream
  IO(slice_alloc(h));
  h->spec_valid = false;       // the surface [H+] seed changes under a surface part issued ahead
  BgAsyncScope as(h, true);
  IO(side_wait(h));
  ProfScope ps(h, "biogem");
  const int nmat = bg_part_materialise(h, h->stream);   // int_bio_part_timeslice integrates bio_part as the reference holds it
  ps.done(nmat + launch_bg_slice(h->dv, h->bgd, h->slice, dts / kBgYrS, 0, h->stream));
  return check_async(h);
}
extern "C" int cg_biogem_slice_reset(cg_handle *h) {   // sub_init_int_timeslice, biogem_data.f90:1012-1060
  BGREADY(h);
  IO(join_side(h));
  IO(slice_alloc(h));
  const size_t ijk = (size_t)h->g.I * h->g.J * h->g.K, MS = h->MS;
  CUDA_OK(cudaMemsetAsync(h->slice.ocn, 0, ijk * h->g.L * MS * 8, h->stream));
  CUDA_OK(cudaMemsetAsync(h->slice.part, 0, ijk * h->bg.LS * MS * 8, h->stream));
  CUDA_OK(cudaMemsetAsync(h->slice.carb, 0, ijk * kSlCarb * MS * 8, h->stream));
  CUDA_OK(cudaMemsetAsync(h->slice.cc, 0, ijk * kSlCC * MS * 8, h->stream));
  CUDA_OK(cudaMemsetAsync(h->slice.iso, 0, ijk * kSlIso * MS * 8, h->stream));
  CUDA_OK(cudaMemsetAsync(h->slice.t, 0, MS * 8, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}
extern "C" int cg_biogem_sig_reset(cg_handle *h) {
  BGREADY(h);
  BgAsyncScope as(h, true);
  IO(side_wait(h));
  CUDA_OK(cudaMemsetAsync(h->sig.acc, 0, (size_t)(kSigHead + 3 * h->g.L + h->bg.LA) * h->dv.MS * sizeof(double), h->stream));
  if (h->sig.acc2) CUDA_OK(cudaMemsetAsync(h->sig.acc2, 0, (size_t)(kSig2Head + h->bg.LS + 2 * h->bg.LA) * h->dv.MS * sizeof(double), h->stream));
  return CG_OK;
}
// cpl_flux_ocnsed(dts, ...), sedgem.f90:1029-1068: sfxsumsed = sfxsumsed + dts * sfxsed1 (sediment grid = ocean grid)
extern "C" int cg_cpl_flux_ocnsed(cg_handle *h, double dts) {
  CG_RANGE();
  BGREADY(h);
  BgAsyncScope as(h, true);
  IO(side_wait(h));
  ProfScope ps(h, "biogem");
  ps.done(launch_cpl_ocnsed(h->sfxsumsed, h->bgd.sfxsed1, (size_t)h->g.I * h->g.J * h->bg.LS * h->dv.MS, dts, 0.0, 0, h->stream));
  return check_async(h);
}
// cpl_comp_ocnsed(ocnstep, mbiogem, msedgem, ...), sedgem.f90:894-937: running mean of the bottom-water composition over
// the BIOGEM steps of one SEDGEM step; ocnstep = koverall / kocn_loop (genie_loop_wrappers.f90:219-226)
extern "C" int cg_cpl_comp_ocnsed(cg_handle *h, int ocnstep, int mbiogem, int msedgem) {
  CG_RANGE();
  BGREADY(h);
  if (mbiogem <= 0 || msedgem <= 0) return fail(CG_ERR_ARG, "cg_cpl_comp_ocnsed: conv_kocn_kbiogem and conv_kocn_ksedgem must be positive");
  const int w = ((ocnstep - mbiogem) % msedgem) / mbiogem;     // Fortran MOD and C % both truncate towards zero
  BgAsyncScope as(h, true);
  IO(side_wait(h));
  ProfScope ps(h, "biogem");
  ps.done(launch_cpl_ocnsed(h->sfcsumocn, h->bgd.sfcocn1, (size_t)h->g.I * h->g.J * h->g.L * h->dv.MS, (double)w, (double)(w + 1), 1, h->stream));
  return check_async(h);
}
// reinit_flux_rokocn(sfxsumrok1), rokgem.f90:472-480
extern "C" int cg_reinit_flux_rokocn(cg_handle *h) {
  BGREADY(h);
  BgAsyncScope as(h, true);
  IO(side_wait(h));
  CUDA_OK(cudaMemsetAsync(h->sfxsumrok1, 0, (size_t)h->g.I * h->g.J * h->g.L * h->dv.MS * sizeof(double), h->stream));
  return CG_OK;
}
// biogem_tracercoupling(go_ts, go_ts1), biogem.f90:1885-1890.  Host arrays are optional (NULL = resident).
extern "C" int cg_biogem_tracercoupling(cg_handle *h, double *go_ts, double *go_ts1) {
  CG_RANGE();
  if (!h || !h->initialised) return fail(CG_ERR_STATE, "handle not initialised");
  if (h->g.L <= 2 || !h->dv.bg_ocn) return fail(CG_ERR_CONFIG, "tracer coupling needs biogeochemical tracers (maxl > 2)");
  activate(h);
  if (h->lazy_stage) IO(flush_lazy(h));
  if (go_ts) IO(cg_sync_from_host(h, "ts", h->io_member, go_ts, cg_field_size(h, "ts")));
  if (h->bg.on && !h->bg_go) return CG_OK;
  BgAsyncScope as(h, !go_ts && !go_ts1);
  h->bg_stage = (h->bg_stage == 1 && !go_ts && !go_ts1) ? 2 : 0;
  // the sums cg_atchem_step took ahead of time own the reduction scratch until their event has passed
  const bool ahead = as.on && h->tc_spec_valid && h->tc_old_pending;
  if (h->tc_old_pending) { CUDA_OK(cudaStreamWaitEvent(h->stream, h->evTcOld, 0)); h->tc_old_pending = false; }
  h->tc_spec_valid = false;
  {
    ProfScope ps(h, "biogem");
    if (ahead) { const int n = launch_tc_sums_new(h->dv, h->stream); ps.done(n + bg_launch_apply(h, h->stream)); }
    else if (h->pd_pending) { const int n = launch_tc_sums_first(h->dv, h->stream); ps.done(n + bg_launch_apply(h, h->stream)); }
    else ps.done(launch_tracercoupling(h->dv, h->stream));
  }
  IO(check_async(h));
  if (go_ts) IO(cg_sync_to_host(h, "ts", h->io_member, go_ts, cg_field_size(h, "ts")));
  if (go_ts1) IO(cg_sync_to_host(h, "ts", h->io_member, go_ts1, cg_field_size(h, "ts")));
  return CG_OK;
}
// biogem_climate, biogem.f90:2132-2239: the physics it copies (u, rho, sea ice, winds, cost, MLD) is aliased
// in place on the device; the only state change on this path is the reset of the convection counter (:2238).
extern "C" int cg_biogem_climate(cg_handle *h) {
  CG_RANGE();
  if (!h || !h->initialised) return fail(CG_ERR_STATE, "handle not initialised");
  activate(h);
  if (h->lazy_stage) IO(flush_lazy(h));
  if (h->bg.on) {
    // go_solfor of the last surflux call (embm.f90:3727-3729): row MOD(istot-1,nyear)+1 of solfor
    h->bgd.nsol = h->istep_ocn > 0 ? (h->istep_ocn - 1) % h->g.nyear + 1 : 0;
    int n = 0;
    if (!h->bg_staged) IO(bg_stage_inputs(h, h->stream));   // on the caller's stream: at call time
    BgAsyncScope as(h, true);
    h->bg_stage = (h->bg_stage == 2) ? 3 : 0;
    ProfScope ps(h, "biogem");
    ps.done(n + launch_bg_climate(h->dv, h->bgd, h->stream));
  } else {
    ProfScope ps(h, "biogem");
    ps.done(launch_bg_reset_cost(h->dv, h->stream));
  }
  return check_async(h);
}
// biogem_climate_sol_wrapper (genie_loop_wrappers.f90:338-342): insolation only
extern "C" int cg_biogem_climate_sol(cg_handle *h) {
  BGREADY(h);
  h->bgd.nsol = h->istep_ocn > 0 ? (h->istep_ocn - 1) % h->g.nyear + 1 : 0;
  return CG_OK;
}
// (re)build BIOGEM's ocn from the current ts: T in K, S absolute, tracers as they are (initialise_biogem)
extern "C" int cg_biogem_init_ocn(cg_handle *h) {
  if (!h || !h->initialised) return fail(CG_ERR_STATE, "handle not initialised");
  if (h->g.L <= 2 || !h->dv.bg_ocn) return fail(CG_ERR_CONFIG, "needs biogeochemical tracers (maxl > 2)");
  const int I = h->g.I, J = h->g.J, K = h->g.K, L = h->g.L;
  std::vector<double> t((size_t)L * I * J * K);
  for (int m = 0; m < h->M; m++) {
    IO(cg_sync_to_host(h, "ts", m, t.data(), (int64_t)t.size()));
    for (size_t c = 0; c < (size_t)I * J * K; c++) {
      t[c * L] = t[c * L] + 273.15;
      t[c * L + 1] = t[c * L + 1] + h->mp[m].saln0;
    }
    IO(cg_sync_from_host(h, "ocn", m, t.data(), (int64_t)t.size()));
  }
  return CG_OK;
}
// step_atchem(dts, sfxsumatm, sfcatm), atchem.f90:63-67, with cpl_comp_atmocn (:252-264) fused
extern "C" int cg_atchem_step(cg_handle *h, double dts) {
  CG_RANGE();
  BGREADY(h);
  if (dts != h->bgd.dts_atchem) return fail(CG_ERR_ARG, "cg_atchem_step: dts differs from conv_kocn_katchem*kocn_loop*genie_timestep");
  // an ATCHEM step that does not follow this iteration's biogem_climate call (conv_kocn_katchem /= conv_kocn_kbiogem, or
  // a host that calls it out of the canonical order): the air temperature / humidity cpl_comp_EMBM copies are staged here
  if (h->stream != h->stream4 && !h->bg_staged && h->bg_stage != 3) IO(bg_stage_inputs(h, h->stream));
  BgAsyncScope as(h, true, true);
  { ProfScope ps(h, "biogem"); ps.done(launch_bg_atchem(h->dv, h->bgd, h->atm_totV, h->stream)); }
  CUDA_OK(cudaEventRecord(h->evAtchem, h->stream));
  h->atchem_ev_pending = true;
  // the block came in the canonical order (step, coupling, climate, ATCHEM) on the asynchronous stream: issue the surface
  // part of the next step behind it, for the clock one BIOGEM period later
  const Params &p = h->base;
  static const bool nospec = getenv("CG_BG_SPLIT") && atoi(getenv("CG_BG_SPLIT")) == 0;
  if (as.on && h->bg_stage == 3 && h->bg_go && !h->bg_fuse && !nospec && p.conv_kocn_kbiogem == p.conv_kocn_katchem) {
    // ... and, next to it on stream5, the half of the next coupling's sums that reads BIOGEM's own state only (see
    // do_biogem_block_pipelined); ATCHEM borrowed the same scratch, hence the event
    if (!(getenv("CG_TC_AHEAD") && atoi(getenv("CG_TC_AHEAD")) == 0) && h->dv.L > 2 &&
        cudaEventRecord(h->evFork5, h->stream) == cudaSuccess && cudaStreamWaitEvent(h->stream5, h->evFork5, 0) == cudaSuccess) {
      h->launches += launch_tc_sums_old(h->dv, h->stream5);
      CUDA_OK(cudaEventRecord(h->evTcOld, h->stream5));
      h->tc_spec_valid = true;
      h->tc_old_pending = true;
    }
    const long long next = h->bg_last_clock + (long long)p.conv_kocn_kbiogem * p.kocn_loop * nint_ll(1000.0 * p.genie_timestep);
    bg_forcing(&h->bg, next, &h->bgd);
    h->launches += launch_bg_surf(h->dv, h->bgd, h->stream);
    h->spec_valid = true;
    h->spec_clock = next;
  }
  h->bg_stage = 0;
  return check_async(h);
}
// the BIOGEM / ATCHEM block of one koverall iteration (genie.f90:352-447)
static int do_biogem_block(cg_handle *h, long long k) {
  const Params &p = h->base;
  if (!h->bg.on) return CG_OK;
  const long long clock = k * nint_ll(1000.0 * p.genie_timestep);   // increment_genie_clock, genie_global.f90:401-410
  if (k % ((long long)p.conv_kocn_kbiogem * p.kocn_loop) == 0) {
    if (k == (long long)p.conv_kocn_kbiogem * p.kocn_loop) IO(cg_biogem_climate_sol(h));
    IO(cg_biogem_forcing(h, clock));
    static const bool nofuse = getenv("CG_BG_NOFUSE") != nullptr;
    if (h->bg_staged && h->bg_go && !h->bg_fuse && !getenv("CG_BG_NOSPLIT")) {
      // asynchronous block (cg_run): the coupling's global sums do not depend on this step's anomaly (the salinity
      // anomaly is +0.0), so they run on a side stream next to the latency-bound step kernel; bit-identical
      const double t = h->bg.t_runtime - (double)clock / (1000.0 * kBgYrS);
      CUDA_OK(cudaEventRecord(h->evFork5, h->stream));
      CUDA_OK(cudaStreamWaitEvent(h->stream5, h->evFork5, 0));
      h->launches += launch_tc_sums_first(h->dv, h->stream5);
      CUDA_OK(cudaEventRecord(h->evJoin5, h->stream5));
      h->launches += bg_part_materialise(h, h->stream);
      h->launches += launch_bg_step(h->dv, h->bgd, 0, 0, h->stream);
      CUDA_OK(cudaStreamWaitEvent(h->stream, h->evJoin5, 0));
      h->launches += bg_launch_apply(h, h->stream);
      if (t < kBgNullSmall) h->bg_go = false;
      IO(check_async(h));
    } else if (nofuse || !h->bg_fuse) {
      IO(cg_biogem_step(h, h->bgd.dts, clock));
      IO(cg_biogem_tracercoupling(h, nullptr, nullptr));
    } else if (h->bg_go) {
      // step_biogem with biogem_tracercoupling's per-cell update fused in: the coupling's global sums do not depend on
      // this step's anomaly, so they are taken first; bit-identical to the two separate calls (tests/test_gpu_col.py)
      const double t = h->bg.t_runtime - (double)clock / (1000.0 * kBgYrS);
      ProfScope ps(h, "biogem");
      int n = launch_tc_sums_first(h->dv, h->stream);
      n += bg_part_materialise(h, h->stream);
      n += launch_bg_step(h->dv, h->bgd, 0, 1, h->stream);
      ps.done(n);
      if (t < kBgNullSmall) h->bg_go = false;
      IO(check_async(h));
    }
    IO(cg_biogem_climate(h));
  }
  // asynchronous block: the next tracer step needs ts (tracer coupling) and cost (climate), not the atmosphere
  if (h->stream == h->stream4) { CUDA_OK(cudaEventRecord(h->evBG, h->stream4)); h->bg_pending = true; }
  if (k % ((long long)p.conv_kocn_katchem * p.kocn_loop) == 0) IO(cg_atchem_step(h, h->bgd.dts_atchem));
  return CG_OK;
}

// One ocean cycle = kocn_loop iterations of the koverall loop when katm_loop == 1 and
// ksic_loop == kocn_loop (the only schedule tools/config_utils.py:103-162 generates):
// surflux, kocn_loop x EMBM, sea ice, ocean.
// One ocean cycle (kocn_loop iterations of the koverall loop, regular schedule).  fork = true (graph capture only): the
// momentum step -- which needs nothing but rho of the previous tracer step and constants, and is dominated by the
// latency-bound barotropic solve -- runs on a second stream next to surflux / EMBM / sea ice and joins before tstepo.
static int enqueue_cycle_head(cg_handle *h, bool fork) {   // everything of the cycle that precedes tstepo
  const Params &p = h->base;
  if (fork) {
    launch_usnap(h->dv, h->stream);
    h->launches++;
    CUDA_OK(cudaEventRecord(h->evFork, h->stream));
    CUDA_OK(cudaStreamWaitEvent(h->stream2, h->evFork, 0));
    IO(do_momentum(h, h->stream2, h->stream3));
    CUDA_OK(cudaEventRecord(h->evJoin, h->stream2));
    h->forked = true;
  }
  int rc = do_surflux(h);
  if (!rc) rc = do_embm(h, p.kocn_loop);
  if (!rc) rc = do_seaice(h);
  h->forked = false;
  do_gold_pre(h);
  if (fork) CUDA_OK(cudaStreamWaitEvent(h->stream, h->evJoin, 0));
  else if (!rc) rc = do_momentum(h, h->stream);
  return rc;
}
static int enqueue_cycle(cg_handle *h, bool fork) {
  IO(enqueue_cycle_head(h, fork));
  do_tstepo(h);
  return CG_OK;
}
// capture `body` (launches on h->stream) into an executable graph
template <class F>
static int capture_graph(cg_handle *h, cudaGraphExec_t *ge, F body) {
  cudaGraph_t gr;
  CUDA_OK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  int rc = body();
  cudaError_t e = cudaStreamEndCapture(h->stream, &gr);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(CG_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
  CUDA_OK(cudaGraphInstantiate(ge, gr, 0));
  cudaGraphDestroy(gr);
  return CG_OK;
}
// The BIOGEM / ATCHEM block of step n touches nothing the head of cycle n+1 (momentum, surflux, EMBM, sea ice) reads or
// writes -- ts_cur, ocn, bio_part, atm ... on one side; rho, u, tq, varice, sst ... on the other -- except the sea-ice
// cover biogem_climate snapshots, which is staged before the streams part.  So inside cg_run the block goes to its own
// low-priority stream and only tstepo(n+1) waits for it.
static int do_biogem_block_async(cg_handle *h, long long k) {
  const Params &p = h->base;
  if (!h->bg.on) return CG_OK;
  const bool due = k % ((long long)p.conv_kocn_kbiogem * p.kocn_loop) == 0 || k % ((long long)p.conv_kocn_katchem * p.kocn_loop) == 0;
  if (!due) return CG_OK;
  const bool climate_due = k % ((long long)p.conv_kocn_kbiogem * p.kocn_loop) == 0;
  IO(bg_stage_inputs(h, h->stream));
  h->bg_staged = true;
  (void)climate_due;
  CUDA_OK(cudaEventRecord(h->evT, h->stream));
  CUDA_OK(cudaStreamWaitEvent(h->stream4, h->evT, 0));
  cudaStream_t save = h->stream;
  h->stream = h->stream4;
  int rc = do_biogem_block(h, k);
  h->stream = save;
  h->bg_staged = false;
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(h->evBGtail, h->stream4));
  h->bg_tail_pending = true;
  return CG_OK;
}
// Pipelined form of the asynchronous block (kbiogem == katchem, no fusion): step_biogem reads BIOGEM's own state only
// (ocn, bio_part, the atmosphere, the sea-ice snapshot -- all last written by the previous block) and the forcing of its
// clock, never ts, so the step kernel of block n+1 is enqueued on stream4 as soon as block n has been, and runs next to
// the two ocean cycles in between.  At its nominal time only the coupling (sums, update), the climate snapshot and
// ATCHEM remain.  The order of every read and write of every field is the serial one: bit-identical.
// `remaining` = koverall iterations this cg_run call will still execute (the step kernel is only issued ahead if its own
// block is certain to follow inside the same call, so the host never observes state from the future).
static int bg_issue_step(cg_handle *h, long long clock) {
  IO(cg_biogem_forcing(h, clock));
  const double t = h->bg.t_runtime - (double)clock / (1000.0 * kBgYrS);
  if (!h->bg_go) return CG_OK;   // par_misc_t_go (biogem.f90:1851-1853)
  h->launches += bg_part_materialise(h, h->stream);
  h->launches += launch_bg_step(h->dv, h->bgd, 0, 0, h->stream);
  if (t < kBgNullSmall) h->bg_go = false;
  return CG_OK;
}
// Split form of the pipelined block: only the SURFACE part of step_biogem (carbonate chemistry, gas exchange, export
// production: transcendental-bound, no HBM traffic to speak of) is issued one block ahead, where it runs next to the
// latency-bound momentum kernels; the water-column sweep stays at its nominal place.
static int bg_issue_surf(cg_handle *h, long long clock) {
  IO(cg_biogem_forcing(h, clock));
  const double t = h->bg.t_runtime - (double)clock / (1000.0 * kBgYrS);
  h->bg_surf_issued = false;
  if (!h->bg_go) return CG_OK;   // par_misc_t_go (biogem.f90:1851-1853)
  h->launches += launch_bg_surf(h->dv, h->bgd, h->stream);
  h->bg_surf_issued = true;
  if (t < kBgNullSmall) h->bg_go = false;
  return CG_OK;
}
static int do_biogem_block_pipelined(cg_handle *h, long long k, long long remaining) {
  const Params &p = h->base;
  const long long period = (long long)p.conv_kocn_kbiogem * p.kocn_loop;
  if (!h->bg.on || k % period != 0) return CG_OK;
  const long long tick = nint_ll(1000.0 * p.genie_timestep);
  if (k == period) IO(cg_biogem_climate_sol(h));
  IO(bg_stage_inputs(h, h->stream));
  CUDA_OK(cudaEventRecord(h->evT, h->stream));
  CUDA_OK(cudaStreamWaitEvent(h->stream5, h->evT, 0));
  cudaStream_t save = h->stream;
  h->stream = h->stream4;
  h->bg_staged = true;
  int rc = CG_OK;
  do {
    if (!h->bg_ahead) {   // first block of this call: the step kernel at its nominal place
      if (cudaStreamWaitEvent(h->stream4, h->evT, 0) != cudaSuccess) { rc = fail(CG_ERR_CUDA, "stream wait"); break; }
      if ((rc = h->bg_split ? bg_issue_surf(h, k * tick) : bg_issue_step(h, k * tick))) break;
    }
    // CG_BG_FUSE2=1 (experiment): the sweep is not issued ahead but here, behind the coupling's sums, with the per-cell update of
    // biogem_tracercoupling applied as each cell's anomaly is known (k_bg_step's `fuse`): no vdocn round trip, no k_tc_apply
    static const bool fuse2 = getenv("CG_BG_FUSE2") && atoi(getenv("CG_BG_FUSE2")) != 0;
    const bool fuse_now = fuse2 && h->bg_split && h->bg_surf_issued && h->bg_go;
    if (!fuse_now && h->bg_split && h->bg_surf_issued) {   // sediment return + water-column sweep, after this cycle's tracer step
      if (cudaStreamWaitEvent(h->stream4, h->evT, 0) != cudaSuccess) { rc = fail(CG_ERR_CUDA, "stream wait"); break; }
      h->launches += bg_launch_sweep(h, h->stream4);
      h->bg_surf_issued = false;
    }
    const bool old_ready = h->bg_ahead && h->tc_old_ready;
    h->bg_ahead = false;
    h->tc_old_ready = false;
    if (fuse_now) {
      h->launches += old_ready ? launch_tc_sums_new(h->dv, h->stream5) : launch_tc_sums_first(h->dv, h->stream5);
      if (cudaEventRecord(h->evJoin5, h->stream5) != cudaSuccess || cudaStreamWaitEvent(h->stream4, h->evJoin5, 0) != cudaSuccess ||
          cudaStreamWaitEvent(h->stream4, h->evT, 0) != cudaSuccess) { rc = fail(CG_ERR_CUDA, "stream wait"); break; }
      h->launches += bg_launch_sweep(h, h->stream4, 1);
      h->bg_surf_issued = false;
    } else
    if (h->bg_go) {       // biogem_tracercoupling: sums on stream5 (they need ts of this cycle, not the step's anomaly)
      // the sums over BIOGEM's own state (old mean salinity, old inventories) were taken one block ahead if old_ready
      h->launches += old_ready ? launch_tc_sums_new(h->dv, h->stream5) : launch_tc_sums_first(h->dv, h->stream5);
      if (cudaEventRecord(h->evJoin5, h->stream5) != cudaSuccess || cudaStreamWaitEvent(h->stream4, h->evJoin5, 0) != cudaSuccess ||
          cudaStreamWaitEvent(h->stream4, h->evT, 0) != cudaSuccess) { rc = fail(CG_ERR_CUDA, "stream wait"); break; }
      h->launches += bg_launch_apply(h, h->stream4);
    } else if (cudaStreamWaitEvent(h->stream4, h->evT, 0) != cudaSuccess) { rc = fail(CG_ERR_CUDA, "stream wait"); break; }
    if ((rc = cg_biogem_climate(h))) break;
    if (cudaEventRecord(h->evBG, h->stream4) != cudaSuccess) { rc = fail(CG_ERR_CUDA, "event record"); break; }
    h->bg_pending = true;
    if ((rc = cg_atchem_step(h, h->bgd.dts_atchem))) break;
    if (remaining >= period) {   // the step kernel (split form: its surface part) of the next block, one block ahead
      // the half of the next block's coupling sums that reads ocn, V, M only (all final once this block's update has run):
      // on stream5 behind ATCHEM (which borrows the same scratch), next to the transcendental-bound surface part.
      // 144 -> 77 us of kernels between the tracer step and k_tc_apply.  CG_TC_AHEAD=0: off.
      const bool tc_ahead = !(getenv("CG_TC_AHEAD") && atoi(getenv("CG_TC_AHEAD")) == 0);
      if (tc_ahead && h->bg_go && h->dv.L > 2) {
        if (cudaEventRecord(h->evFork5, h->stream4) != cudaSuccess || cudaStreamWaitEvent(h->stream5, h->evFork5, 0) != cudaSuccess) {
          rc = fail(CG_ERR_CUDA, "stream wait"); break;
        }
        h->launches += launch_tc_sums_old(h->dv, h->stream5);
        h->tc_old_ready = true;
      }
      if ((rc = h->bg_split ? bg_issue_surf(h, (k + period) * tick) : bg_issue_step(h, (k + period) * tick))) break;
      // ... and the sweep right behind it (default; CG_BG_SWEEP_EARLY=1: behind the tracer step of the cycle before the
      // block's, 0: at the block's nominal place.  Measured 79.5 / 81.2 / 82.4 ms per model year.)
      if (!fuse2 && h->bg_split && h->bg_surf_issued && !(getenv("CG_BG_SWEEP_EARLY") && atoi(getenv("CG_BG_SWEEP_EARLY")) != 2)) {
        h->launches += bg_launch_sweep(h, h->stream);
        h->bg_surf_issued = false;
      }
      h->bg_ahead = true;
    }
    if (cudaEventRecord(h->evBGtail, h->stream4) != cudaSuccess) { rc = fail(CG_ERR_CUDA, "event record"); break; }
    h->bg_tail_pending = true;
  } while (0);
  h->stream = save;
  h->bg_staged = false;
  if (rc) return rc;
  return check_async(h);
}
static int bg_join(cg_handle *h) {   // order the main stream after an outstanding BIOGEM block
  if (h->bg_pending) {
    CUDA_OK(cudaStreamWaitEvent(h->stream, h->evBG, 0));
    h->bg_pending = false;
  }
  return CG_OK;
}

// the two executable graphs of an ocean cycle for the current variant and ping-pong parity, captured on first use
static int cycle_graphs(cg_handle *h) {
  const int par = (h->dv.ts_cur < h->dv.ts_new) ? 0 : 1;
  cudaGraphExec_t &g1 = h->graph[h->variant][par], &g2 = h->graph2[h->variant][par];
  if (g1 && g2) return CG_OK;
  const long long l0 = h->launches;
  const int i0 = h->istep_ocn, a0 = h->istep_atm, s0 = h->istep_sic;
  const bool fork = h->fork_momentum && !getenv("CG_NOFORK");
  if (g1) { cudaGraphExecDestroy(g1); g1 = nullptr; }
  if (g2) { cudaGraphExecDestroy(g2); g2 = nullptr; }
  IO(capture_graph(h, &g1, [&]() { return enqueue_cycle_head(h, fork); }));
  IO(capture_graph(h, &g2, [&]() { do_tstepo(h); return (int)CG_OK; }));
  h->graph_launches[h->variant] = h->launches - l0;
  h->launches = l0; h->istep_ocn = i0; h->istep_atm = a0; h->istep_sic = s0;
  std::swap(h->dv.ts_cur, h->dv.ts_new);  // undo the swap done while capturing
  return CG_OK;
}
// Per-module path, lazy cycle.  The Fortran host calls surflux, kocn_loop x step_embm, step_seaice, step_goldstein once per
// ocean cycle, mostly with no array to exchange (NULL = stay resident).  Such calls are only noted; when step_goldstein
// completes the canonical sequence the cycle is issued exactly as cg_run issues it -- the captured head graph (momentum
// next to surflux / EMBM / sea ice, 5 EMBM steps in one launch), the join with the BIOGEM block, the tracer-step graph --
// instead of ~30 separate launches.  State is observable from the host only through calls that pass arrays or through the
// sync / get entry points, and all of those replay the noted calls first (flush_lazy).  Bit-identical to both other paths
// (tests/test_gpu_col.py::test_module_by_module_matches_run).  CG_NOLAZY=1: off.
static bool lazy_ok(const cg_handle *h) {
  const Params &p = h->base;
  return h->eager && !h->profile && h->use_graphs && p.katm_loop == 1 && p.ksic_loop == p.kocn_loop && p.kocn_loop > 1 &&
         !getenv("CG_NOEAGER") && !(getenv("CG_NOLAZY") && atoi(getenv("CG_NOLAZY")) != 0);
}
static int flush_lazy(cg_handle *h) {
  const int st = h->lazy_stage, nl = h->base.kocn_loop;
  if (!st) return CG_OK;
  h->lazy_stage = 0;
  IO(eager_momentum(h));
  IO(do_surflux(h));
  for (int q = 0; q < std::min(st - 1, nl); q++) IO(do_embm(h, 1));
  if (st == nl + 2) IO(do_seaice(h));
  return check_async(h);
}
static int lazy_cycle(cg_handle *h) {   // step_goldstein closing a fully noted cycle
  const Params &p = h->base;
  h->lazy_stage = 0;
  const int par = (h->dv.ts_cur < h->dv.ts_new) ? 0 : 1;
  IO(cycle_graphs(h));
  CUDA_OK(cudaGraphLaunch(h->graph[h->variant][par], h->stream));
  IO(bg_join(h));                            // tstepo reads the ts the tracer coupling rewrote
  CUDA_OK(cudaGraphLaunch(h->graph2[h->variant][par], h->stream));
  h->launches += h->graph_launches[h->variant];
  h->istep_ocn++; h->istep_atm += p.kocn_loop; h->istep_sic++;
  std::swap(h->dv.ts_cur, h->dv.ts_new);
  return check_async(h);
}

extern "C" int cg_run(cg_handle *h, int64_t n) {
  CG_RANGE();
  READY(h);
  IO0(join_side(h));
  h->spec_valid = false;
  h->tc_spec_valid = false;
  IO0(drop_momentum(h));  // cg_run computes the momentum step inside its own schedule
  const Params &p = h->base;
  const bool regular = p.katm_loop == 1 && p.ksic_loop == p.kocn_loop && p.kocn_loop > 1;
  const int trace_n = getenv("CG_TRACE") ? atoi(getenv("CG_TRACE")) : 0;
  std::vector<cudaEvent_t> trace_ev;
  std::vector<int> trace_blk;
  struct TraceDump {
    std::vector<cudaEvent_t> &ev; std::vector<int> &blk; cg_handle *h;
    ~TraceDump() {
      if (ev.empty()) return;
      cudaDeviceSynchronize();
      float t0 = 0.f;
      for (size_t c = 0; c + 3 < ev.size(); c += 4) {
        float head, wait, ts, gap = 0.f;
        cudaEventElapsedTime(&head, ev[c], ev[c + 1]);
        cudaEventElapsedTime(&wait, ev[c + 1], ev[c + 2]);
        cudaEventElapsedTime(&ts, ev[c + 2], ev[c + 3]);
        if (c) cudaEventElapsedTime(&gap, ev[c - 1], ev[c]);
        cudaEventElapsedTime(&t0, ev[0], ev[c + 3]);
        fprintf(stderr, "trace cycle %2zu%s: gap %6.1f head %6.1f wait %6.1f tstep %6.1f us  (t = %8.1f)\n", c / 4, blk[c / 4] ? " B" : "  ",
                1e3 * gap, 1e3 * head, 1e3 * wait, 1e3 * ts, 1e3 * t0);
      }
      for (auto e : ev) cudaEventDestroy(e);
    }
  } trace_dump{trace_ev, trace_blk, h};
  while (n > 0) {
    const long long k = h->koverall + 1;
    if (regular && (k % p.kocn_loop) == 1 && n >= p.kocn_loop) {
      if (h->use_graphs && !h->profile) {
        // two graphs per variant and ping-pong parity: the head of the cycle and the tracer step
        const int par = (h->dv.ts_cur < h->dv.ts_new) ? 0 : 1;
        IO(cycle_graphs(h));
        cudaGraphExec_t &g1 = h->graph[h->variant][par], &g2 = h->graph2[h->variant][par];
        // CG_TRACE=<cycles>: main-stream time stamps of the first cycles of this call (head / wait for the BIOGEM block /
        // tracer step), printed to stderr at the end of the call -- a diagnostic of the schedule, not used by the bench
        cudaEvent_t *tev = nullptr;
        if (trace_n > 0 && (int)trace_ev.size() < 4 * trace_n) {
          for (int q = 0; q < 4; q++) { cudaEvent_t e; cudaEventCreate(&e); trace_ev.push_back(e); }
          tev = &trace_ev[trace_ev.size() - 4];
          trace_blk.push_back((h->koverall + p.kocn_loop) % ((long long)p.conv_kocn_kbiogem * p.kocn_loop) == 0);
        }
        if (tev) cudaEventRecord(tev[0], h->stream);
        CUDA_OK(cudaGraphLaunch(g1, h->stream));
        if (tev) cudaEventRecord(tev[1], h->stream);
        IO(bg_join(h));                            // tstepo reads the ts the tracer coupling rewrote
        if (tev) cudaEventRecord(tev[2], h->stream);
        CUDA_OK(cudaGraphLaunch(g2, h->stream));
        if (tev) cudaEventRecord(tev[3], h->stream);
        h->launches += h->graph_launches[h->variant];
        h->istep_ocn++; h->istep_atm += p.kocn_loop; h->istep_sic++;
        std::swap(h->dv.ts_cur, h->dv.ts_new);
        h->koverall += p.kocn_loop;
        n -= p.kocn_loop;
        // split form, cycle before the block's cycle: the sweep kernel of the coming block is issued behind this tracer
        // step, i.e. it runs next to the head of the next cycle (momentum: latency bound, SMs mostly idle) and not at the
        // block's nominal place on the critical path.  It reads and writes BIOGEM's own arrays only.
        if (h->bg_ahead && h->bg_split && h->bg_surf_issued && h->bg.on && !(getenv("CG_BG_FUSE2") && atoi(getenv("CG_BG_FUSE2")) != 0)) {
          const long long period = (long long)p.conv_kocn_kbiogem * p.kocn_loop;
          const bool early = !(getenv("CG_BG_SWEEP_EARLY") && atoi(getenv("CG_BG_SWEEP_EARLY")) == 0);
          if (early && h->koverall % period != 0 && (h->koverall + p.kocn_loop) % period == 0) {
            CUDA_OK(cudaEventRecord(h->evT, h->stream));
            CUDA_OK(cudaStreamWaitEvent(h->stream4, h->evT, 0));
            h->launches += bg_launch_sweep(h, h->stream4);
            h->bg_surf_issued = false;
          }
        }
        if (h->bg_overlap && !getenv("CG_BG_SERIAL")) {
          // CG_BG_PIPE=1 (whole step kernel ahead): no gain on B200 -- the 255-register kernel then shares the SMs with the tracer step
          // default: pipelined block with only the surface part of the step issued ahead (CG_BG_SPLIT=0: off;
          // 99.5 -> 81.4 ms per model year at 128 members, bit-identical)
          if (!h->bg_ahead) { const char *e = getenv("CG_BG_SPLIT"); h->bg_split = e ? atoi(e) != 0 : true; }
          const bool pipe = !h->bg_fuse && p.conv_kocn_kbiogem == p.conv_kocn_katchem && (getenv("CG_BG_PIPE") || h->bg_split);
          if (pipe || h->bg_ahead) IO(do_biogem_block_pipelined(h, h->koverall, n));
          else IO(do_biogem_block_async(h, h->koverall));
        } else {
          IO(do_biogem_block(h, h->koverall));
        }
        continue;
      }
      IO(enqueue_cycle(h, false));
      h->koverall += p.kocn_loop;
      n -= p.kocn_loop;
      IO(do_biogem_block(h, h->koverall));
      continue;
    }
    // general schedule (genie.f90:271-311)
    if (k % p.kocn_loop == 1) IO(do_surflux(h));
    if (k % p.katm_loop == 0) IO(do_embm(h, 1));
    if (k % p.ksic_loop == 0) IO(do_seaice(h));
    if (k % p.kocn_loop == 0) IO(do_goldstein(h));
    h->koverall++;
    n--;
    IO(do_biogem_block(h, h->koverall));
  }
  IO(join_side(h));
  return check_async(h);
}

// Restart support: place the coupling loop at iteration `koverall` (a multiple of kocn_loop).  The prognostic fields are
// restored with cg_sync_from_host; this restores the counters the reference keeps in genie_global (koverall, istep_*,
// genie_clock) and everything BIOGEM derives from them.
extern "C" int cg_set_koverall(cg_handle *h, int64_t koverall) {
  READY(h);
  IO0(join_side(h));
  IO0(drop_momentum(h));
  h->spec_valid = false;
  h->tc_spec_valid = false;
  const Params &p = h->base;
  if (koverall < 0 || koverall % p.kocn_loop != 0) return fail(CG_ERR_ARG, "cg_set_koverall: koverall must be a non-negative multiple of kocn_loop");
  h->koverall = koverall;
  h->istep_ocn = (int)(koverall / p.kocn_loop);
  h->istep_atm = (int)(koverall / p.katm_loop);
  h->istep_sic = (int)(koverall / p.ksic_loop);
  CUDA_OK(cudaMemcpyAsync(h->dv.istep_ocn, &h->istep_ocn, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (h->bg.on) {
    h->bgd.nsol = h->istep_ocn > 0 ? (h->istep_ocn - 1) % h->g.nyear + 1 : 0;
    const long long clock = koverall * nint_ll(1000.0 * p.genie_timestep);
    h->bg_go = !((h->bg.t_runtime - (double)clock / (1000.0 * kBgYrS)) < kBgNullSmall) || koverall == 0;
    for (int la = 3; la <= h->bg.LA; la++)
      if (h->bg.rst_sel[la]) { h->bg.rst_sig_i1[la] = (int)h->bg.rst_sig_t[la].size(); h->bg.rst_sig_i2[la] = h->bg.rst_sig_i1[la]; }
  }
  return CG_OK;
}

extern "C" int cg_refresh_rho(cg_handle *h, int member) {
  READY(h);
  if (member >= h->M) return fail(CG_ERR_ARG, "cg_refresh_rho: no such member");
  const int I = h->g.I, J = h->g.J, K = h->g.K, L = h->g.L;
  std::vector<double> ts((size_t)L * I * J * K), rho((size_t)I * J * K);
  for (int m = (member < 0 ? 0 : member); m < (member < 0 ? h->M : member + 1); m++) {
    IO(cg_sync_to_host(h, "ts", m, ts.data(), (int64_t)ts.size()));
    IO(cg_sync_to_host(h, "rho", m, rho.data(), (int64_t)rho.size()));
    for (int k = 1; k <= K; k++)
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++) {
          if (k < h->g.k1at(i, j)) continue;
          const size_t c = cell3(I, J, i, j, k);
          rho[c] = eos_z(h->mc[m].ec, h->base.ieos, ts[c * L], ts[c * L + 1], h->g.zro[k]);
        }
    IO(cg_sync_from_host(h, "rho", m, rho.data(), (int64_t)rho.size()));
    if (h->dv.sst) {   // SST / SSS as step_goldstein last exported them (tsval, ssval of initialise_goldstein after a restart)
      std::vector<double> sst((size_t)2 * I * J);
      IO(cg_sync_to_host(h, "sst", m, sst.data(), (int64_t)sst.size()));
      for (int j = 1; j <= J; j++)
        for (int i = 1; i <= I; i++) {
          if (h->g.k1at(i, j) > K) continue;
          const size_t c = cell3(I, J, i, j, K);
          sst[0 + 2 * cell2(I, i, j)] = ts[c * L];
          sst[1 + 2 * cell2(I, i, j)] = ts[c * L + 1];
        }
      IO(cg_sync_from_host(h, "sst", m, sst.data(), (int64_t)sst.size()));
    }
  }
  return CG_OK;
}

// ------------------------------------------------------------------ diagnostics, measurement
extern "C" int cg_global_means(cg_handle *h, double *out) {
  if (!h || !h->initialised || !out) return fail(CG_ERR_ARG, "cg_global_means: bad argument");
  IO0(join_side(h));
  activate(h);
  launch_global_means(h->dv, h->d_means, h->stream);
  CUDA_OK(cudaMemcpyAsync(out, h->d_means, (size_t)h->M * h->g.L * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}
extern "C" int cg_health(cg_handle *h, int32_t *out) {
  if (!h || !h->initialised || !out) return fail(CG_ERR_ARG, "cg_health: bad argument");
  IO0(join_side(h));
  activate(h);
  launch_health(h->dv, h->d_flags, h->stream);
  CUDA_OK(cudaMemcpyAsync(out, h->d_flags, (size_t)h->M * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (h->bg.on) {  // carbonate-chemistry failure = the reference's error_stop (gem_carbchem.f90:452-456)
    std::vector<int> e(h->M);
    CUDA_OK(cudaMemcpy(e.data(), h->bgd.err, (size_t)h->M * 4, cudaMemcpyDeviceToHost));
    for (int m = 0; m < h->M; m++) if (e[m]) out[m] |= 2;
  }
  return CG_OK;
}
extern "C" int cg_synchronize(cg_handle *h) {
  if (!h || !h->initialised) return fail(CG_ERR_STATE, "handle not initialised");
  IO0(join_side(h));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  return CG_OK;
}
extern "C" int64_t cg_launch_count(cg_handle *h, int reset) {
  if (!h) return -1;
  const long long n = h->launches;
  if (reset) h->launches = 0;
  return n;
}
extern "C" int cg_timer_start(cg_handle *h) {
  if (!h || !h->initialised) return fail(CG_ERR_STATE, "handle not initialised");
  IO0(join_side(h));
  CUDA_OK(cudaEventRecord(h->ev0, h->stream));
  return CG_OK;
}
extern "C" int cg_timer_stop_ms(cg_handle *h, double *ms) {
  if (!h || !h->initialised || !ms) return fail(CG_ERR_STATE, "handle not initialised");
  IO0(join_side(h));
  CUDA_OK(cudaEventRecord(h->ev1, h->stream));
  CUDA_OK(cudaEventSynchronize(h->ev1));
  float f = 0;
  CUDA_OK(cudaEventElapsedTime(&f, h->ev0, h->ev1));
  *ms = f;
  return CG_OK;
}
extern "C" int cg_profile_enable(cg_handle *h, int on) {
  if (!h) return fail(CG_ERR_ARG, "null handle");
  h->profile = on != 0;
  if (on) h->prof.clear();
  return CG_OK;
}
extern "C" int cg_profile_get(cg_handle *h, const char *family, double *total_ms, int64_t *launches) {
  if (!h || !family) return fail(CG_ERR_ARG, "cg_profile_get: bad argument");
  auto it = h->prof.find(family);
  if (total_ms) *total_ms = it == h->prof.end() ? 0.0 : it->second.ms;
  if (launches) *launches = it == h->prof.end() ? 0 : it->second.n;
  return CG_OK;
}
extern "C" int cg_set_biogem_fusion(cg_handle *h, int on) {
  if (!h) return fail(CG_ERR_ARG, "cg_set_biogem_fusion: bad handle");
  h->bg_fuse = on != 0;
  return CG_OK;
}
extern "C" int cg_tracer_variant_active(cg_handle *h) {
  if (!h) return -1;
  return (h->variant == 2 && !tstep_col_supported(h->dv)) ? 1 : h->variant;
}
extern "C" int cg_set_tracer_variant(cg_handle *h, int variant) {
  if (!h || variant < 0 || variant > 2) return fail(CG_ERR_ARG, "variant must be 0 (strict), 1 (fast) or 2 (fused column kernel)");
  h->variant = variant;
  return CG_OK;
}
extern "C" int cg_set_graphs(cg_handle *h, int on) {
  if (!h) return fail(CG_ERR_ARG, "null handle");
  h->use_graphs = on != 0;
  return CG_OK;
}

// ------------------------------------------------------------------ stand-alone tracer step
extern "C" int cg_tracer_create(int maxi, int maxj, int maxk, int maxl, int n_members, int device, const int32_t *k1, double diff1,
                                double diff2, int nyear, cg_handle **out) {
  if (!k1 || !out || maxi < 2 || maxj < 2 || maxk < 2 || maxl < 2 || n_members < 1)
    return fail(CG_ERR_ARG, "cg_tracer_create: bad argument");
  if (maxj + 2 > kMaxJ || maxk + 2 > kMaxK) return fail(CG_ERR_CONFIG, "grid larger than the compiled metric tables");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device)
    return fail(CG_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
  std::unique_ptr<cg_handle> h(new cg_handle);
  h->device = device;
  h->M = n_members;
  h->MS = ((n_members + 31) / 32) * 32;   // stand-alone tracer step: generic-shape kernels, any stride
  h->tracer_only = true;
  Params p;
  p.maxi = maxi; p.maxj = maxj; p.maxk = maxk; p.maxl = maxl; p.nyear = nyear; p.diff1 = diff1; p.diff2 = diff2;
  h->base = p;
  // k1 is given as (0:maxi+1, 0:maxj+1) column-major; Grid::build wants file order (rows j = J+1..0)
  std::vector<int> k1file((size_t)(maxi + 2) * (maxj + 2));
  size_t q = 0;
  for (int j = maxj + 1; j >= 0; j--)
    for (int i = 0; i <= maxi + 1; i++) k1file[q++] = k1[i + (maxi + 2) * j];
  h->g.build(maxi, maxj, maxk, maxl, 0, nyear, p.yearlen, k1file);
  h->isl.isles = 0;
  h->mp.assign(n_members, p);
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming));
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = least urgent
    CUDA_OK(cudaStreamCreateWithPriority(&h->stream4, cudaStreamNonBlocking, lo));
  }
  CUDA_OK(cudaStreamCreateWithFlags(&h->stream5, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&h->evFork5, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evJoin5, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evTcOld, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evT, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evBG, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evBGtail, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evAtchem, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evFork3, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&h->evJoin3, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreate(&h->ev0));
  CUDA_OK(cudaEventCreate(&h->ev1));
  h->mc.resize(1);
  build_member(h->g, h->isl, h->w, p, &h->mc[0], nullptr);
  h->mc.resize(n_members, h->mc[0]);
  h->baro_group.assign(h->MS, 0);
  fill_gridc(h.get());
  register_hconst(h.get(), h->mc[0]);
  int rc = build_device(h.get());
  if (rc) return rc;
  h->initialised = true;
  activate(h.get());
  CUDA_OK(cudaStreamSynchronize(h->stream));
  *out = h.release();
  return CG_OK;
}

// ts: [m][(maxl,0:maxi+1,0:maxj+1,0:maxk+1)], u: [m][(3,0:maxi,0:maxj,maxk)], tsflux: [m][(2,maxi,maxj)] or NULL
extern "C" int cg_tracer_set(cg_handle *h, const double *ts, const double *u, const double *tsflux) {
  if (!h || !h->initialised) return fail(CG_ERR_STATE, "handle not initialised");
  activate(h);
  const int I = h->g.I, J = h->g.J, K = h->g.K, L = h->g.L, M = h->M;
  const size_t ij = (size_t)I * J, ijk = ij * K;
  if (ts) {
    std::vector<double> t(ijk * L), r(ijk);
    const size_t nts = (size_t)L * (I + 2) * (J + 2) * (K + 2);
    for (int m = 0; m < M; m++) {
      for (int k = 1; k <= K; k++)
        for (int j = 1; j <= J; j++)
          for (int i = 1; i <= I; i++) {
            const size_t c = cell3(I, J, i, j, k);
            const double *src = ts + m * nts + (size_t)L * (i + (size_t)(I + 2) * (j + (size_t)(J + 2) * k));
            for (int l = 0; l < L; l++) t[c * L + l] = src[l];
            r[c] = eos(h->mc[0].ec, src[0], src[1]);
          }
      IO(cg_sync_from_host(h, "ts", m, t.data(), (int64_t)t.size()));
      IO(cg_sync_from_host(h, "rho", m, r.data(), (int64_t)r.size()));
    }
  }
  if (u) {
    std::vector<double> t(ijk * 3);
    const size_t nu = (size_t)3 * (I + 1) * (J + 1) * K;
    for (int m = 0; m < M; m++) {
      for (int k = 1; k <= K; k++)
        for (int j = 1; j <= J; j++)
          for (int i = 1; i <= I; i++) {
            const size_t c = cell3(I, J, i, j, k);
            const double *src = u + m * nu + (size_t)3 * (i + (size_t)(I + 1) * (j + (size_t)(J + 1) * (k - 1)));
            for (int cc = 0; cc < 3; cc++) t[c * 3 + cc] = src[cc];
          }
      IO(cg_sync_from_host(h, "u", m, t.data(), (int64_t)t.size()));
    }
  }
  if (tsflux)
    for (int m = 0; m < M; m++) IO(cg_sync_from_host(h, "tsflux", m, tsflux + (size_t)m * 2 * ij, (int64_t)(2 * ij)));
  return CG_OK;
}
extern "C" int cg_tracer_step(cg_handle *h, int nsteps) {
  if (!h || !h->initialised) return fail(CG_ERR_STATE, "handle not initialised");
  activate(h);
  for (int s = 0; s < nsteps; s++) do_tstepo(h);
  return check_async(h);
}
// ts: [m][(maxl,maxi,maxj,maxk)], rho: [m][(maxi,maxj,maxk)], cost: [m][(maxi,maxj)]; any may be NULL
extern "C" int cg_tracer_get(cg_handle *h, double *ts, double *rho, double *cost) {
  if (!h || !h->initialised) return fail(CG_ERR_STATE, "handle not initialised");
  for (int m = 0; m < h->M; m++) {
    if (ts) IO(cg_sync_to_host(h, "ts", m, ts + (size_t)m * cg_field_size(h, "ts"), cg_field_size(h, "ts")));
    if (rho) IO(cg_sync_to_host(h, "rho", m, rho + (size_t)m * cg_field_size(h, "rho"), cg_field_size(h, "rho")));
    if (cost) IO(cg_sync_to_host(h, "cost", m, cost + (size_t)m * cg_field_size(h, "cost"), cg_field_size(h, "cost")));
  }
  return CG_OK;
}
