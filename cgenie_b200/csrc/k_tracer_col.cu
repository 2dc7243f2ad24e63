// k_tracer_col.cu -- tracer step as two kernels: tstepo_flux with cp.async.bulk staging (one block per wet column), then
// the convective adjustment + SST export (one thread per member-column).  See k_tracer_col.cuh for the formulation.  Compiled with FMA
// contraction (fast variant family, <= 1e-10 per step against the strict kernels / the oracle).
#include <cuda.h>   // CUtensorMap, cuTensorMapEncodeTiled (resolved through cudaGetDriverEntryPoint: no link against libcuda)

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "k_tracer_col.cuh"

namespace cg {

static __constant__ GridC c_g;

void upload_grid_tracer_col(const GridC &g, cudaStream_t s) {
  cudaMemcpyToSymbolAsync(c_g, &g, sizeof(GridC), 0, cudaMemcpyHostToDevice, s);
}

// one block = the MS members of ONE wet column (row-major column order: neighbouring blocks share stencil rows in L2)
template <int I, int J, int K, int L, int MS, int MINB, bool PV, bool AR = false>
__global__ void __launch_bounds__(MS, MINB) k_tstep_col(const Dev v) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColStage st;
  st.sm = reinterpret_cast<double *>(smem_raw);
  st.bar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)ColRows<L>::rows * MS * 8);
  st.tid = threadIdx.x;
  const int c2 = v.col_deep_first ? v.wetcols[blockIdx.x] : v.rowcols[blockIdx.x];
  tstep_column<I, J, K, L, MS, MS, PV, AR>(v, c_g, c2, threadIdx.x, st);
}

// pipelined form (production): coefficients one level ahead, T and S staged like every other tracer (see tstep_column2)
template <int I, int J, int K, int L, int MS, int MINB>
__global__ void __launch_bounds__(MS, MINB) k_tstep_col2(const Dev v) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColStage st;
  st.sm = reinterpret_cast<double *>(smem_raw);
  st.bar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)ColRows2<L>::rows * MS * 8);
  st.tid = threadIdx.x;
  const int c2 = v.col_deep_first ? v.wetcols[blockIdx.x] : v.rowcols[blockIdx.x];
  tstep_column2<I, J, K, L, MS, MS>(v, c_g, c2, threadIdx.x, st);
}

// warp-tile form: one block = one warp = 32 members of one column; consecutive blocks are the MS / 32 tiles of a column
template <int I, int J, int K, int L, int MS>
__global__ void __launch_bounds__(32, 8) k_tstep_colw(const Dev v, const __grid_constant__ ColMaps maps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NTILE = MS / 32;
  ColStage st;
  st.sm = reinterpret_cast<double *>(smem_raw);
  st.bar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)ColRows<L>::rows * 32 * 8);
  st.tid = threadIdx.x;
  const int c2 = v.rowcols[blockIdx.x / NTILE];
  const unsigned m = (blockIdx.x % NTILE) * 32 + threadIdx.x;
  tstep_column_w<I, J, K, L, MS>(v, c_g, c2, m, st, &maps);
}

// tile form of the block kernel: one block = NT = 128 members (tile blockIdx.x / nwet) of one wet column of a handle whose member
// stride is MS = 256 or 512; tile-major block order, so that the stencil rows a tile's neighbouring columns share stay in L2 as
// they do at MS = 128.  Staging through 2-D tensor maps (see tstep_column, TM).
template <int I, int J, int K, int L, int MS, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_tstep_colt(const Dev v, const __grid_constant__ ColMaps maps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColStage st;
  st.sm = reinterpret_cast<double *>(smem_raw);
  st.bar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)ColRows<L>::rows * NT * 8);
  st.tid = threadIdx.x;
  const int tile = blockIdx.x / v.nwet, ci = blockIdx.x - tile * v.nwet;
  const int c2 = v.col_deep_first ? v.wetcols[ci] : v.rowcols[ci];
  tstep_column<I, J, K, L, MS, NT, false, false, true>(v, c_g, c2, (unsigned)(tile * NT) + threadIdx.x, st, &maps);
}

// 2-D tensor maps of a field seen as [row][member] (fp64, row pitch MS doubles), box = `boxw` members x `rows` rows, no swizzle
static bool make_map(void *out128, const void *base, unsigned long long nrows, int MS, int rows, unsigned boxw = 32u) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p) return false;
    fn = (EncodeFn)p;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)MS, (cuuint64_t)nrows};
  const cuuint64_t gstr[1] = {(cuuint64_t)MS * sizeof(double)};
  const cuuint32_t box[2] = {boxw, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUtensorMap tmap;
  const CUresult r = fn(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  memcpy(out128, &tmap, 128);
  return true;
}
template <int I, int J, int K, int L, int MS, unsigned BOXW = 32u, int LW = L, int L0 = 0>
static const ColMaps *col_maps(const Dev &v) {
  static std::mutex mu;
  static std::map<std::pair<const void *, const void *>, ColMaps> cache;   // (ts buffer read this step, u) -> maps
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_pair((const void *)v.ts_cur, (const void *)v.u);
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  ColMaps mp;
  using R = ColRows<LW, L0 == 0>;
  const unsigned long long nts = (unsigned long long)I * J * K * L, nu = (unsigned long long)I * J * K * 3;
  if (!make_map(mp.ts2, v.ts_cur, nts, MS, 2, BOXW) || !make_map(mp.tsA, v.ts_cur, nts, MS, R::nA > 0 ? R::nA : 1, BOXW) ||
      !make_map(mp.tsB, v.ts_cur, nts, MS, R::nB, BOXW) || !make_map(mp.u3, v.u, nu, MS, 3, BOXW) || !make_map(mp.u1, v.u, nu, MS, 1, BOXW))
    return nullptr;
  return &(cache[key] = mp);
}

// split form: two threads per (member, column), 2 * MS threads per block (see tstep_column_split)
template <int I, int J, int K, int L, int MS, int MINB>
__global__ void __launch_bounds__(2 * MS, MINB) k_tstep_split(const Dev v) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColStage st;
  st.sm = reinterpret_cast<double *>(smem_raw);
  st.bar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)SplitRows<L>::rows * MS * 8);
  st.tid = threadIdx.x;
  const int c2 = v.rowcols[blockIdx.x];
  tstep_column_split<I, J, K, L, MS, MS, false>(v, c_g, c2, threadIdx.x, st);
}

// T, S pre-pass + convective-adjustment decisions of the mix-on-write form: one thread per (member, wet column)
template <int I, int J, int K, int L, int MS, int MINB>
__global__ void __launch_bounds__(128, MINB) k_ts_pre(const Dev v) {
  const unsigned m = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ci = blockIdx.y * 4 + (threadIdx.x >> 5);
  if (ci >= v.nwet) return;
  ts_pre_column<I, J, K, L, MS>(v, c_g, v.rowcols[ci], m);
}

// convective adjustment + SST export: one thread per (member, wet column).  (A split into a decisions kernel and a
// (member, column, tracer pair)-parallel averaging kernel was measured slower: +3.8 ms per model year at 128 members.)
template <int I, int J, int K, int L, int MS>
__global__ void __launch_bounds__(128) k_co_col(const Dev v) {
  // the column arrays of the decision loop (T, S, rho, box thickness) in shared memory, [level][thread]: see co_decide_core
  extern __shared__ __align__(16) unsigned char co_smem[];
  const unsigned m = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ci = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);   // blockDim.x / 32 columns per block (1 .. 4)
  if (ci >= v.nwet) return;
  double *scratch = v.co_local ? nullptr : reinterpret_cast<double *>(co_smem) + threadIdx.x;
  co_column<I, J, K, L, MS>(v, c_g, v.rowcols[ci], m, scratch, 128);
}

// the same, one warp per block, compiled for MINB resident blocks per SM (column arrays thread-private)
template <int I, int J, int K, int L, int MS, int MINB, int PMODE = -1, int LOCKSTEP = 0>
__global__ void __launch_bounds__(32, MINB) k_co_col1(const Dev v) {
  const int c2 = v.col_deep_first ? v.polcols[blockIdx.y] : v.rowcols[blockIdx.y];   // (col_deep_first doubles as "poleward rows first" here)
  co_column<I, J, K, L, MS, false, PMODE, LOCKSTEP>(v, c_g, c2, blockIdx.x * 32 + threadIdx.x, nullptr, 1);
}

// two member tiles of one column per block (64 threads): at <= 48 registers 21 blocks = 42 warps are resident per SM, more than the
// hardware's 32 one-warp blocks
template <int I, int J, int K, int L, int MS, int MINB>
__global__ void __launch_bounds__(64, MINB) k_co_col2w(const Dev v) {
  co_column<I, J, K, L, MS, false, 0>(v, c_g, v.rowcols[blockIdx.y], blockIdx.x * 64 + threadIdx.x, nullptr, 1);
}

// ... and its decisions alone (the region maps go to comask for k_co_passive): without the averaging code the kernel needs fewer
// registers and its warps retire sooner
template <int I, int J, int K, int L, int MS, int MINB>
__global__ void __launch_bounds__(32, MINB) k_co_dec1(const Dev v) {
  co_column<I, J, K, L, MS, true>(v, c_g, v.rowcols[blockIdx.y], blockIdx.x * 32 + threadIdx.x, nullptr, 1);
}

// Block-cooperative form of the convective adjustment: one block = 32 members (lanes) of ONE wet column.  Warp 0 takes the
// decisions for its flagged lanes (T, S, rho of the column, serial per lane: latency bound) and leaves the region maps in shared
// memory; then every warp averages one pair of passive tracers over the mixed regions, so that the loads of all seven pairs are
// in flight together instead of one pair after the other in the same thread.  A block none of whose 32 member-columns was flagged
// unstable by the flux kernel returns at once (the flags are read by all warps alike, so the exit is block-uniform).
template <int I, int J, int K, int L, int MS, int NW>
__global__ void __launch_bounds__(32 * NW, 2) k_co_blk(const Dev v) {
  __shared__ unsigned s_in[32], s_top[32], s_bot[32];
  __shared__ double s_rdzt[K][32];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = blockIdx.x * 32 + lane;
  const int c2 = v.rowcols[blockIdx.y];
  const unsigned flag = v.comask[(long)c2 * MS + m];
  if (!__any_sync(0xffffffffu, flag != 0u)) return;
  if (warp == 0) {
    unsigned in = 0, topb = 0, botb = 0;
    double rdzt[K];
#pragma unroll
    for (int k = 0; k < K; k++) rdzt[k] = 0.0;
    if (flag) co_decide<I, J, K, L, MS>(v, c_g, c2, m, in, topb, botb, rdzt);
    s_in[lane] = in; s_top[lane] = topb; s_bot[lane] = botb;
#pragma unroll
    for (int k = 0; k < K; k++) s_rdzt[k][lane] = rdzt[k];
  }
  __syncthreads();
  const unsigned in = s_in[lane];
  if (in == 0u) return;
  const unsigned topb = s_top[lane], botb = s_bot[lane];
  double rdzt[K];
#pragma unroll
  for (int k = 0; k < K; k++) rdzt[k] = s_rdzt[k][lane];
  for (int l = 2 + 2 * (int)warp; l < L; l += 2 * NW) co_passive_pair<I, J, K, L, MS>(v, c_g, c2, m, in, topb, botb, rdzt, l);
}

// passive tracers of the mixed regions, one thread per (member, column, tracer): block = 32 members x (L - 2) tracers of one column
template <int I, int J, int K, int L, int MS>
__global__ void __launch_bounds__(32 * (L - 2)) k_co_passive(const Dev v) {
  const unsigned m = blockIdx.x * 32 + threadIdx.x;
  co_passive_one<I, J, K, L, MS>(v, c_g, v.rowcols[blockIdx.y], m, 2 + (int)threadIdx.y);
}

constexpr int kCoMinbDefault = 516;   // form of the convection kernel (table in go_tiled); CG_CO_MINB overrides

template <int I, int J, int K, int L, int MS>
static int go(const Dev &v, cudaStream_t s, int cfg) {
  constexpr size_t smem = (size_t)ColRows<L>::rows * MS * 8 + 64;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_tstep_col<I, J, K, L, MS, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_tstep_col<I, J, K, L, MS, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_tstep_col<I, J, K, L, MS, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_tstep_col<I, J, K, L, MS, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  // mix-on-write form (CG_COL_MIX=1, opt-in): T, S pre-pass with the convection decisions, then the passive tracers with
  // the mixing applied as they are written -- the second pass over ts of k_co_col disappears (528 MB of DRAM traffic per
  // step instead of 844 MB) but the per-cell coefficients are computed twice; measured slower on B200 (241 us against
  // 217 us at 128 members: both forms are bound by fp64 latency, not by bandwidth; profiles/README_r1.md).
  static int mix = -1;
  if (mix < 0) { const char *e = getenv("CG_COL_MIX"); mix = e ? atoi(e) : 0; }
  if (mix && v.comask && K <= 16) {
    static int minb = -1;   // registers / occupancy of the pre-pass: 2 -> 202 regs, 3 -> 168, 4 -> 128 (spills)
    if (minb < 0) { const char *e = getenv("CG_PRE_MINB"); minb = e ? atoi(e) : 3; }
    const dim3 gp(MS / 32, (v.nwet + 3) / 4);
    if (minb == 2) k_ts_pre<I, J, K, L, MS, 2><<<gp, 128, 0, s>>>(v);
    else if (minb == 4) k_ts_pre<I, J, K, L, MS, 4><<<gp, 128, 0, s>>>(v);
    else k_ts_pre<I, J, K, L, MS, 3><<<gp, 128, 0, s>>>(v);
    k_tstep_col<I, J, K, L, MS, 2, true><<<v.nwet, MS, smem, s>>>(v);
    return 2;
  }
  // split form (CG_COL_SPLIT=1): two threads per member-column, <= 128 registers, 16 warps per SM
  static int split = -1;
  if (split < 0) { const char *e = getenv("CG_COL_SPLIT"); split = e ? atoi(e) : 0; }
  static int colv = -1;   // 1 (default): flux kernel of round 1 + stability flag; 2: pipelined form, coefficients one level ahead
  if (colv < 0) { const char *e = getenv("CG_COL_V"); colv = e ? atoi(e) : 1; }
  static int order = -1, coskip = -1;
  if (order < 0) { const char *e = getenv("CG_COL_ORDER"); order = e ? atoi(e) : 0; }
  if (coskip < 0) { const char *e = getenv("CG_CO_SKIP"); coskip = e ? atoi(e) : 1; }
  Dev v1 = v;
  v1.col_deep_first = order;
  const ColMaps *maps = (colv == 3) ? col_maps<I, J, K, L, MS>(v) : nullptr;
  if (split && MS == 128) {
    constexpr size_t smem2 = (size_t)SplitRows<L>::rows * MS * 8 + 64;
    static bool attr2 = false;
    if (!attr2) {
      cudaFuncSetAttribute(k_tstep_split<I, J, K, L, MS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      cudaFuncSetAttribute(k_tstep_split<I, J, K, L, MS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      attr2 = true;
    }
    if (split == 2) k_tstep_split<I, J, K, L, MS, 1><<<v.nwet, 2 * MS, smem2, s>>>(v);
    else k_tstep_split<I, J, K, L, MS, 2><<<v.nwet, 2 * MS, smem2, s>>>(v);
  } else
  if (colv == 3 && maps) {
    constexpr size_t smemw = (size_t)ColRows<L>::rows * 32 * 8 + 64;
    k_tstep_colw<I, J, K, L, MS><<<v.nwet * (MS / 32), 32, smemw, s>>>(v1, *maps);
  } else
  if (colv == 2) {
    constexpr size_t smem3 = (size_t)ColRows2<L>::rows * MS * 8 + 64;
    static bool attr3 = false;
    if (!attr3) {
      cudaFuncSetAttribute(k_tstep_col2<I, J, K, L, MS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
      cudaFuncSetAttribute(k_tstep_col2<I, J, K, L, MS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
      attr3 = true;
    }
    if (cfg == 1) k_tstep_col2<I, J, K, L, MS, 1><<<v.nwet, MS, smem3, s>>>(v1);
    else k_tstep_col2<I, J, K, L, MS, 2><<<v.nwet, MS, smem3, s>>>(v1);
  } else
  if (cfg == 1) k_tstep_col<I, J, K, L, MS, 1, false><<<v.nwet, MS, smem, s>>>(v1);
  else if (cfg == 2) k_tstep_col<I, J, K, L, MS, 2, false, true><<<v.nwet, MS, smem, s>>>(v1);   // mbarrier buffer release
  else k_tstep_col<I, J, K, L, MS, 2, false><<<v.nwet, MS, smem, s>>>(v1);
  static int copf = -1;
  if (copf < 0) { const char *e = getenv("CG_CO_PF"); copf = e ? atoi(e) : 0; }   // measured: the L2 prefetch costs more than it hides (profiles/README_r1.md)
  Dev v2 = v;
  v2.co_prefetch = copf;
  v2.co_skip_stable = (coskip && !(split && MS == 128)) ? 1 : 0;   // the split form does not write the stability flag
  static int copair = -1;
  if (copair < 0) { const char *e = getenv("CG_CO_PAIR"); copair = e ? atoi(e) : 1; }   // measured: region by region is 6 us slower
  v2.co_pairwise = copair;
  constexpr size_t co_smem_full = (size_t)4 * (K + 2) * 128 * sizeof(double);
  static int colocal = -1;
  if (colocal < 0) {
    const char *e = getenv("CG_CO_LOCAL");
    colocal = e ? atoi(e) : 1;   // measured: shared-memory column arrays change nothing (169.8 vs 170.8 us per tracer step)
    cudaFuncSetAttribute(k_co_col<I, J, K, L, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)co_smem_full);
  }
  v2.co_local = colocal;
  const size_t co_smem_bytes = colocal ? 0 : co_smem_full;
  static int cov = -1;
  if (cov < 0) { const char *e = getenv("CG_CO_V"); cov = e ? atoi(e) : 1; }
  constexpr int NW = (L - 2 + 1) / 2 > 0 ? ((L - 2 + 1) / 2 < 8 ? (L - 2 + 1) / 2 : 8) : 1;   // one warp per passive tracer pair
  if (cov == 2 && v2.co_skip_stable && v.comask) k_co_blk<I, J, K, L, MS, NW><<<dim3(MS / 32, v.nwet), 32 * NW, 0, s>>>(v2);
  else if (cov == 3 && v2.co_skip_stable && v.comask && L > 2) {
    // decisions (thread = member x column), then the passive tracers with one thread per (member, column, tracer)
    v2.co_pairwise = 2;
    k_co_col<I, J, K, L, MS><<<dim3(MS / 32, (v.nwet + 3) / 4), 128, co_smem_bytes, s>>>(v2);
    k_co_passive<I, J, K, L, MS><<<dim3(MS / 32, v.nwet), dim3(32, L - 2), 0, s>>>(v2);
    return 3;
  } else {
    static int wpb = -1;   // see go_tiled
    if (wpb < 0) { const char *e = getenv("CG_CO_WPB"); wpb = e ? atoi(e) : 1; if (wpb < 1 || wpb > 4 || !colocal) wpb = 4; }
    static int lock = -1;  // CG_CO_MINB >= 400: the lockstep form of the decisions, as in go_tiled (one code path for every member stride)
    if (lock < 0) { const char *e = getenv("CG_CO_MINB"); lock = e ? atoi(e) : kCoMinbDefault; }
    if (lock >= 400 && v2.co_skip_stable && v.comask && L > 2) {
      v2.col_deep_first = 0;
      if (lock == 412) k_co_col1<I, J, K, L, MS, 12, 0, 1><<<dim3(MS / 32, v.nwet), 32, 0, s>>>(v2);
      else if (lock >= 500) k_co_col1<I, J, K, L, MS, 16, 0, 2><<<dim3(MS / 32, v.nwet), 32, 0, s>>>(v2);
      else k_co_col1<I, J, K, L, MS, 16, 0, 1><<<dim3(MS / 32, v.nwet), 32, 0, s>>>(v2);
      return 2;
    }
    k_co_col<I, J, K, L, MS><<<dim3(MS / 32, (v.nwet + wpb - 1) / wpb), 32 * wpb, co_smem_bytes, s>>>(v2);
  }
  return 2;
}

// Tracer windows: one block = one wet column x MS members x a WINDOW of LW tracers starting at tracer L0 of a state that carries
// LT tracers per cell (BASELINE config #5: 128 x 128 x 32 with 40 tracers = windows of 16 + 12 + 12; one thread cannot carry the P / Q
// pairs of 40 tracers).  Every window computes the cell coefficients from T, S; the window with L0 = 0 advances T, S, rho and the
// stability flag.
template <int I, int J, int K, int LT, int MS, int LW, int L0, int MINB>
__global__ void __launch_bounds__(MS, MINB) k_tstep_colx(const Dev v) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColStage st;
  st.sm = reinterpret_cast<double *>(smem_raw);
  st.bar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)ColRows<LW, L0 == 0>::rows * MS * 8);
  st.tid = threadIdx.x;
  const int c2 = v.rowcols[blockIdx.x];
  tstep_column<I, J, K, LW, MS, MS, false, false, false, LT, L0>(v, c_g, c2, threadIdx.x, st);
}
template <int I, int J, int K, int LT, int MS, int LW, int L0>
static void launch_window(const Dev &v, cudaStream_t s) {
  constexpr size_t smem = (size_t)ColRows<LW, L0 == 0>::rows * MS * 8 + 64;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_tstep_colx<I, J, K, LT, MS, LW, L0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  k_tstep_colx<I, J, K, LT, MS, LW, L0, 2><<<v.nwet, MS, smem, s>>>(v);
}
// 128 x 128 x 32, 40 tracers: three windows + the convection kernel
template <int MS>
static int go_config5(const Dev &v, cudaStream_t s) {
  constexpr int I = 128, J = 128, K = 32, LT = 40;
  Dev v1 = v;
  v1.col_deep_first = 0;
  launch_window<I, J, K, LT, MS, 16, 0>(v1, s);
  launch_window<I, J, K, LT, MS, 12, 16>(v1, s);
  launch_window<I, J, K, LT, MS, 12, 28>(v1, s);
  Dev v2 = v;
  v2.co_prefetch = 0; v2.co_skip_stable = 1; v2.co_pairwise = 1; v2.co_local = 1;
  k_co_col<I, J, K, LT, MS><<<dim3(MS / 32, v.nwet), 32, 0, s>>>(v2);
  return 4;
}

// ... and the tile form advancing a window of LW tracers starting at L0 (see k_tstep_colx): two windows of 8 tracers need 60 KB of
// staging and ~170 registers per block instead of 100 KB and 238, at the price of computing the cell coefficients twice
template <int I, int J, int K, int LT, int MS, int NT, int LW, int L0, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_tstep_coltw(const Dev v, const __grid_constant__ ColMaps maps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColStage st;
  st.sm = reinterpret_cast<double *>(smem_raw);
  st.bar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)ColRows<LW, L0 == 0>::rows * NT * 8);
  st.tid = threadIdx.x;
  const int tile = blockIdx.x / v.nwet, ci = blockIdx.x - tile * v.nwet;
  const int c2 = v.rowcols[ci];
  tstep_column<I, J, K, LW, MS, NT, false, false, true, LT, L0>(v, c_g, c2, (unsigned)(tile * NT) + threadIdx.x, st, &maps);
}
template <int I, int J, int K, int LT, int MS, int LW, int L0, int MINB>
static bool launch_window_tiled(const Dev &v, cudaStream_t s) {
  constexpr int NT = 128;
  constexpr size_t smem = (size_t)ColRows<LW, L0 == 0>::rows * NT * 8 + 64;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_tstep_coltw<I, J, K, LT, MS, NT, LW, L0, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const ColMaps *maps = col_maps<I, J, K, LT, MS, 128u, LW, L0>(v);
  if (!maps) return false;
  k_tstep_coltw<I, J, K, LT, MS, NT, LW, L0, MINB><<<v.nwet * (MS / NT), NT, smem, s>>>(v, *maps);
  return true;
}

// member strides above 128: the flux kernel in its tile form (128-member tiles through tensor maps), the convection kernel as it is
// (its grid already runs over 32-member tiles)
template <int I, int J, int K, int L, int MS>
static int go_tiled(const Dev &v, cudaStream_t s) {
  constexpr int NT = 128;
  constexpr size_t smem = (size_t)ColRows<L>::rows * NT * 8 + 64;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_tstep_colt<I, J, K, L, MS, NT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_co_col<I, J, K, L, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)4 * (K + 2) * 128 * sizeof(double)));
    attr = true;
  }
  const ColMaps *maps = col_maps<I, J, K, L, MS, 128u>(v);
  if (!maps) return 0;
  static int order = -1, coskip = -1;
  if (order < 0) { const char *e = getenv("CG_COL_ORDER"); order = e ? atoi(e) : 0; }
  if (coskip < 0) { const char *e = getenv("CG_CO_SKIP"); coskip = e ? atoi(e) : 1; }
  Dev v1 = v;
  v1.col_deep_first = order;
  static int win = -1;   // CG_COL_WIN=1: two tracer windows of 8 (3 blocks per SM), 2: the same at 2 blocks per SM
  if (win < 0) { const char *e = getenv("CG_COL_WIN"); win = e ? atoi(e) : 0; }
  if (win == 1 && L == 16) {
    if (!launch_window_tiled<I, J, K, L, MS, 8, 0, 3>(v1, s) || !launch_window_tiled<I, J, K, L, MS, 8, 8, 3>(v1, s)) return 0;
  } else if (win == 2 && L == 16) {
    if (!launch_window_tiled<I, J, K, L, MS, 8, 0, 2>(v1, s) || !launch_window_tiled<I, J, K, L, MS, 8, 8, 2>(v1, s)) return 0;
  } else
  k_tstep_colt<I, J, K, L, MS, NT, 2><<<v.nwet * (MS / NT), NT, smem, s>>>(v1, *maps);
  Dev v2 = v;
  v2.co_prefetch = 0;
  v2.co_skip_stable = coskip ? 1 : 0;
  v2.co_pairwise = 1;
  v2.co_local = 1;
  // One warp = 32 members of one column.  In a stratified ocean most warps find all their member-columns flagged stable and leave
  // at once; with several warps per block the few that work keep the registers of their idle neighbours allocated until the block
  // retires, so the default is one warp per block (CG_CO_WPB = 1 .. 4)
  static int wpb = -1, cov = -1, copair = -1;
  if (wpb < 0) { const char *e = getenv("CG_CO_WPB"); wpb = e ? atoi(e) : 1; if (wpb < 1 || wpb > 4) wpb = 1; }
  if (cov < 0) { const char *e = getenv("CG_CO_V"); cov = e ? atoi(e) : 1; }
  if (copair < 0) { const char *e = getenv("CG_CO_PAIR"); copair = e ? atoi(e) : 1; }
  v2.co_pairwise = copair;
  if ((cov == 4 || cov == 5) && v2.co_skip_stable && v.comask && L > 2) {
    // the same split with the decisions kernel compiled on its own (CG_CO_V=4: 20 blocks per SM, 5: 24)
    if (cov == 4) k_co_dec1<I, J, K, L, MS, 20><<<dim3(MS / 32, v.nwet), 32, 0, s>>>(v2);
    else k_co_dec1<I, J, K, L, MS, 24><<<dim3(MS / 32, v.nwet), 32, 0, s>>>(v2);
    k_co_passive<I, J, K, L, MS><<<dim3(MS / 32, v.nwet), dim3(32, L - 2), 0, s>>>(v2);
    return 3;
  }
  if (cov == 3 && v2.co_skip_stable && v.comask && L > 2) {
    // decisions (thread = member x column), then the passive tracers with one thread per (member, column, tracer)
    v2.co_pairwise = 2;
    k_co_col<I, J, K, L, MS><<<dim3(MS / 32, (v.nwet + wpb - 1) / wpb), 32 * wpb, 0, s>>>(v2);
    k_co_passive<I, J, K, L, MS><<<dim3(MS / 32, v.nwet), dim3(32, L - 2), 0, s>>>(v2);
    return 3;
  }
  // The one-warp-per-block form compiled for more resident blocks per SM (the kernel is bound by latency at 12 warps per SM).
  // Measured on the bench state (profiles/ab_r4a_*.log, ab_r4d.log; us per tracer step, results bit-identical throughout):
  //   averaging form chosen at run time (both compiled in): 164 registers 556.2 | 128 registers 522.8 | 96 (spills) 532.7 | 80 532.1
  //   pairs form alone, 128 registers 519.8;  regions form alone (no spills at any cap): 126 registers 519.3 | 96 515.7 | 80 511.7 |
  //   72 510.8 | 64 registers = 32 warps per SM 510.3 (CG_CO_MINB=232).  CG_CO_MINB=0: the old form.
  // Decisions in lockstep form (co_decide_static: the column in registers, passes every lane executes alike) + regions averaging
  // (profiles/ab_r4j_co_lockstep.log): 166 registers / 12 blocks per SM 490.3 | 128 registers / 16 blocks 486.1 (CG_CO_MINB=416) |
  // 96 registers (spills) / 20 blocks 497.7 -- against 512.5 for the 64-register walk on that box.  That form merges every unstable run
  // of a pass at once: same decisions as the walk in every model state tried (cost equal, no flipped column from the first model year
  // on against the strict kernels), but not by construction (0.5 % of random columns with 2 K of noise per level end in another
  // partition: two unstable runs interacting through the cubic equation of state).  The DEFAULT (CG_CO_MINB=516) is the lockstep form
  // that replays the reference's own order of merges on the mask of unstable boundaries: identical partition on 166k random columns
  // incl. those (tests/test_col_body_host.py), 504.9 us (profiles/ab_r4l_co_lockstep_walk_order.log).  Box means equal to rounding in
  // both, so neither is bit-identical to the 2xx / 1xx forms; every member stride uses the same form (go and go_tiled).
  static int minb = -1;
  if (minb < 0) { const char *e = getenv("CG_CO_MINB"); minb = e ? atoi(e) : kCoMinbDefault; }
  if (minb && wpb == 1) {
    const dim3 g(MS / 32, v.nwet);
    static int copol = -1;   // CG_CO_POLAR=1: blocks of the poleward rows first
    if (copol < 0) { const char *e = getenv("CG_CO_POLAR"); copol = e ? atoi(e) : 0; }
    v2.col_deep_first = copol;
    // CG_CO_MINB = 116 | 120 | 124 | 216 | 220: the averaging form fixed at compile time (1xx pairs, 2xx regions) at 16 / 20 / 24 blocks
    if (minb == 116) { k_co_col1<I, J, K, L, MS, 16, 1><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 120) { k_co_col1<I, J, K, L, MS, 20, 1><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 124) { k_co_col1<I, J, K, L, MS, 24, 1><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 216) { k_co_col1<I, J, K, L, MS, 16, 0><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 220) { k_co_col1<I, J, K, L, MS, 20, 0><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 224) { k_co_col1<I, J, K, L, MS, 24, 0><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 228) { k_co_col1<I, J, K, L, MS, 28, 0><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 232) { k_co_col1<I, J, K, L, MS, 32, 0><<<g, 32, 0, s>>>(v2); return 2; }
    // 4xx: the decisions in lockstep form (column in registers, passes every lane executes alike: co_decide_static) + regions averaging
    // 5xx: the lockstep passes in the reference's own order of merges (co_decide_static<WALK>): identical partition by construction
    if (minb == 512) { k_co_col1<I, J, K, L, MS, 12, 0, 2><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 516) { k_co_col1<I, J, K, L, MS, 16, 0, 2><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 520) { k_co_col1<I, J, K, L, MS, 20, 0, 2><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 412) { k_co_col1<I, J, K, L, MS, 12, 0, 1><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 416) { k_co_col1<I, J, K, L, MS, 16, 0, 1><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 420) { k_co_col1<I, J, K, L, MS, 20, 0, 1><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 424) { k_co_col1<I, J, K, L, MS, 24, 0, 1><<<g, 32, 0, s>>>(v2); return 2; }
    if (minb == 321) { k_co_col2w<I, J, K, L, MS, 21><<<dim3(MS / 64, v.nwet), 64, 0, s>>>(v2); return 2; }   // 48 registers, 42 warps per SM
    if (minb == 320) { k_co_col2w<I, J, K, L, MS, 20><<<dim3(MS / 64, v.nwet), 64, 0, s>>>(v2); return 2; }   // 48 registers, 40 warps per SM
    if (minb == 318) { k_co_col2w<I, J, K, L, MS, 18><<<dim3(MS / 64, v.nwet), 64, 0, s>>>(v2); return 2; }   // 56 registers, 36 warps per SM
    if (minb == 316) { k_co_col2w<I, J, K, L, MS, 16><<<dim3(MS / 64, v.nwet), 64, 0, s>>>(v2); return 2; }   // 64 registers, 32 warps per SM
    if (minb == 16) k_co_col1<I, J, K, L, MS, 16><<<g, 32, 0, s>>>(v2);
    else if (minb == 20) k_co_col1<I, J, K, L, MS, 20><<<g, 32, 0, s>>>(v2);
    else k_co_col1<I, J, K, L, MS, 24><<<g, 32, 0, s>>>(v2);
    return 2;
  }
  k_co_col<I, J, K, L, MS><<<dim3(MS / 32, (v.nwet + wpb - 1) / wpb), 32 * wpb, 0, s>>>(v2);
  return 2;
}

bool tstep_col_supported(const Dev &v) {
  if (v.iediff || v.ieos || v.iconv || v.imld) return false;   // the column kernel takes diff(2) as a per-member constant and has no thermobaric term
  if (v.I == 128 && v.J == 128 && v.K == 32 && v.L == 40 && (v.MS == 32 || v.MS == 64 || v.MS == 128)) return true;   // BASELINE config #5
  return v.I == 36 && v.J == 36 && v.K == 16 && v.L == 16 && (v.MS == 32 || v.MS == 64 || v.MS == 128 || v.MS == 256 || v.MS == 512);
}

// 0 = this grid shape / member stride has no compiled instance (the caller falls back to the generic kernels)
int launch_tstep_col(const Dev &v, cudaStream_t s) {
  static int cfg = -1;
  if (cfg < 0) {
    const char *e = getenv("CG_COL_CFG");
    cfg = e ? atoi(e) : 0;
  }
  if (!tstep_col_supported(v)) return 0;
  if (v.I == 128) return v.MS == 32 ? go_config5<32>(v, s) : v.MS == 64 ? go_config5<64>(v, s) : go_config5<128>(v, s);
  if (v.MS == 256) return go_tiled<36, 36, 16, 16, 256>(v, s);
  if (v.MS == 512) return go_tiled<36, 36, 16, 16, 512>(v, s);
  if (v.MS == 32) return go<36, 36, 16, 16, 32>(v, s, cfg);
  if (v.MS == 64) return go<36, 36, 16, 16, 64>(v, s, cfg);
  return go<36, 36, 16, 16, 128>(v, s, cfg);
}

}  // namespace cg
