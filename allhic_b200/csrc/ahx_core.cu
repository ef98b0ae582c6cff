// ahx_core.cu -- context, upload, and the batched Tour.Evaluate / EvaluateQ kernels.
//
// Kernels (all sm_100a, no tensor cores: the path is gather + fp64 reciprocal/log +
// reduction):
//   eval_m_res_kernel     (ahx_resident.cuh) Tour.Evaluate (evaluate.go:116-143), the product
//                         path: persistent, warp-specialised, pair table resident in shared
//                         memory or streamed through a TMA ring
//   eval_m_x32_kernel     fallback, 32-bit coordinates, one CTA per T tours, table streamed
//   eval_m_sparse_kernel  fallback, fp64 coordinates (groups longer than 2.1 Gb, N > 65535)
//   eval_m_dense_kernel   the same score in position space, gathering dense M
//                         (north_star's literal formulation; one CTA per tour)
//   eval_q_kernel         CLM.EvaluateQ (orientation.go:159-193), one CTA per
//                         (tour, signs) individual
//   build_dense_kernel    scatters the pair table into the dense N x N int32 M
//   probe_* kernels       DFMA / fp64-divide / L2 read peaks for the roofline
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "ahx_ctx.h"
#include "ahx_device.cuh"
#include "ahx_resident.cuh"

namespace ahx {

int fail(ahx_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  set_global_error(msg);
  return code;
}
int cuda_fail(ahx_ctx *ctx, cudaError_t e, const char *what) {
  return fail(ctx, AHX_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}
int enter(ahx_ctx *ctx, bool need_loaded) {
  if (!ctx) return fail(nullptr, AHX_ERR_ARG, "null context");
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
  if (need_loaded && !ctx->loaded) return fail(ctx, AHX_ERR_STATE, "no linkage group loaded (call ahx_load first)");
  return AHX_OK;
}

// ---------------------------------------------------------------------------
// Tour.Evaluate, contig-space kernel.
//   grid = ceil(P / T) CTAs, kBlock threads.  Dynamic shared memory:
//     x    double[N*T]   mid-point of contig c in tour t at x[c*T+t] (NaN = inactive)
//     tig  int[L]        staging of one tour
//   XGLOBAL: x lives in a per-CTA slice of global memory instead (N too large
//   for shared memory); T must be 1.
// ---------------------------------------------------------------------------
template <int T, typename Coo, typename IdxT, bool XGLOBAL>
__global__ void __launch_bounds__(kBlock)
eval_m_sparse_kernel(const IdxT *__restrict__ tours, const int *__restrict__ rows, int P, int L, int N,
                     const long long *__restrict__ sizes, const typename Coo::Elem *__restrict__ coo, int E,
                     double *__restrict__ xglobal, double *__restrict__ out, int *__restrict__ errflag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long scan_scratch[kWarps];
  __shared__ double red_scratch[kWarps];
  const int tid = threadIdx.x;
  double *x; int *tig;
  if constexpr (XGLOBAL) {
    x = xglobal + static_cast<size_t>(blockIdx.x) * N * T;
    tig = reinterpret_cast<int *>(smem_raw);
  } else {
    x = reinterpret_cast<double *>(smem_raw);
    tig = reinterpret_cast<int *>(smem_raw + sizeof(double) * static_cast<size_t>(N) * T);
  }
  const int p0 = blockIdx.x * T;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (int i = tid; i < N * T; i += kBlock) x[i] = qnan;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < T; ++t) {
    if (p0 + t >= P) break;
    const int row = rows ? rows[p0 + t] : p0 + t;
    const IdxT *src = tours + static_cast<size_t>(row) * L;
    for (int i = tid; i < L; i += kBlock) {
      int v = static_cast<int>(src[i]);
      if (static_cast<unsigned>(v) >= static_cast<unsigned>(N)) { *errflag = 1; v = 0; }  // reported by the host wrapper
      tig[i] = v;
    }
    __syncthreads();
    scatter_mids<T>(tig, L, sizes, x, t, scan_scratch, tid);
    __syncthreads();
  }
  double acc[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.0;
  accumulate_pairs<T, Coo>(x, coo, E, acc, tid);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const double s = block_sum(acc[t], red_scratch, tid);
    if (tid == 0 && p0 + t < P) {
      const int row = rows ? rows[p0 + t] : p0 + t;
      out[row] = 0.0 - s;  // score starts at +0.0 and only subtracts (evaluate.go:128,140)
    }
  }
}

// ---------------------------------------------------------------------------
// Tour.Evaluate, contig-space kernel with 32-bit coordinates (the product path).
//   grid = ceil(P / T) CTAs, kBlock threads.  Dynamic shared memory:
//     tiles  Elem[kStages*kTile]  pair-table tiles filled by TMA bulk copies
//     X      uint32[N*T]          2*mid of contig c in tour t at X[c*T+t]
//     tig    int[L]               staging of one tour
//   The pair table is pre-ordered on the host so that the 8 lanes of every
//   quarter-warp hit 8 different 16-byte bank groups (see order_pairs()).
// ---------------------------------------------------------------------------
template <int T, typename Coo, typename IdxT>
__global__ void __launch_bounds__(kBlock, T == 4 ? 3 : 4)
eval_m_x32_kernel(const IdxT *__restrict__ tours, const int *__restrict__ rows, int P, int L, int N,
                  const uint32_t *__restrict__ sizes32, const typename Coo::Elem *__restrict__ coo, int E_pad,
                  double *__restrict__ out, int *__restrict__ errflag) {
  using Elem = typename Coo::Elem;
  extern __shared__ __align__(128) unsigned char smem_x32[];
  __shared__ long long scan_scratch[kWarps * T];
  __shared__ double red_scratch[kWarps];
  __shared__ __align__(8) uint64_t bars[2 * kStages];
  Elem *tiles = reinterpret_cast<Elem *>(smem_x32);
  Elem *queues = tiles + kStages * kTile;
  uint32_t *X = reinterpret_cast<uint32_t *>(queues + kWarps * kQueue);
  uint32_t *S = X + static_cast<size_t>(N) * T;
  const int tid = threadIdx.x;
  const int p0 = blockIdx.x * T;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars[s], 1); mbar_init(&bars[kStages + s], kWarps); }
    mbar_init_fence();
    prefetch_pairs<Elem>(coo, E_pad, tiles, bars);
  }
  {
    uint4 *X4 = reinterpret_cast<uint4 *>(X);
    const uint4 inact = make_uint4(kInactiveX, kInactiveX, kInactiveX, kInactiveX);
    for (int i = tid; i < (N * T) / 4; i += kBlock) X4[i] = inact;
    for (int i = (N * T) / 4 * 4 + tid; i < N * T; i += kBlock) X[i] = kInactiveX;
  }
  for (int i = tid; i < N; i += kBlock) S[i] = __ldg(sizes32 + i);
  DirectTours<T, IdxT> src;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int p = p0 + t;
    src.ptr[t] = p < P ? tours + static_cast<size_t>(rows ? rows[p] : p) * L : nullptr;
  }
  __syncthreads();
  scatter_tours_x32<T>(src, L, N, S, X, scan_scratch, errflag, tid);
  __syncthreads();
  PairSink<T, Coo> sink;
  sink.init(queues + (tid >> 5) * kQueue, X, tid & 31);
  stream_pairs_x32<T, Coo>(coo, E_pad, tiles, bars, sink, tid);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const double s = block_sum(sink.acc[t], red_scratch, tid);
    if (tid == 0 && p0 + t < P) {
      const int row = rows ? rows[p0 + t] : p0 + t;
      out[row] = 0.0 - s;
    }
  }
}

// ---------------------------------------------------------------------------
// Tour.Evaluate, position-space kernel over the dense int32 M (literal loop of
// evaluate.go:129-141).  One CTA per tour; warp w owns rows i = w, w+8, ...;
// lanes sweep j = i+1.. and stop at the first j beyond LIMIT.
// Shared memory: tig int[L], mid double[L].
// ---------------------------------------------------------------------------
template <typename IdxT>
__global__ void __launch_bounds__(kBlock)
eval_m_dense_kernel(const IdxT *__restrict__ tours, int P, int L, int N, const long long *__restrict__ sizes,
                    const int *__restrict__ M, double *__restrict__ out, int *__restrict__ errflag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long scan_scratch[kWarps];
  __shared__ double red_scratch[kWarps];
  double *mid = reinterpret_cast<double *>(smem_raw);
  int *tig = reinterpret_cast<int *>(smem_raw + sizeof(double) * static_cast<size_t>(L));
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const IdxT *src = tours + static_cast<size_t>(blockIdx.x) * L;
  for (int i = tid; i < L; i += kBlock) {
    int v = static_cast<int>(src[i]);
    if (static_cast<unsigned>(v) >= static_cast<unsigned>(N)) { *errflag = 1; v = 0; }
    tig[i] = v;
  }
  __syncthreads();
  {
    const int per = (L + kBlock - 1) / kBlock;
    const int beg = min(L, tid * per), end = min(L, beg + per);
    long long local = 0;
    for (int i = beg; i < end; ++i) local += __ldg(sizes + tig[i]);
    long long cum = block_excl_scan(local, scan_scratch, tid);
    for (int i = beg; i < end; ++i) {
      const long long s = __ldg(sizes + tig[i]);
      mid[i] = static_cast<double>(2 * cum + s) * 0.5;
      cum += s;
    }
  }
  __syncthreads();
  double acc = 0.0;
  for (int i = wid; i < L - 1; i += kWarps) {
    const int *row = M + static_cast<size_t>(tig[i]) * N;
    const double mi = mid[i];
    for (int j = i + 1 + lane; j < L; j += 32) {
      const double d = mid[j] - mi;
      if (d > kLimit) break;
      const int n = __ldg(row + tig[j]);
      if (n != 0) acc += static_cast<double>(n) / d;
    }
  }
  const double s = block_sum(acc, red_scratch, tid);
  if (tid == 0) out[blockIdx.x] = 0.0 - s;
}

__global__ void build_dense_kernel(const int *__restrict__ pa, const int *__restrict__ pb, const int *__restrict__ nl,
                                   long long E, int N, int *__restrict__ M) {
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < E;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int a = pa[e], b = pb[e], n = nl[e];
    M[static_cast<size_t>(a) * N + b] = n;   // clm.go:443-444
    M[static_cast<size_t>(b) * N + a] = n;
  }
}

// ---------------------------------------------------------------------------
// CLM.EvaluateQ (orientation.go:159-193).  One CTA per individual.
// Shared memory: cum long long[L] | pos int[N] | tig int[L] | sg uint8[N].
// For a stored pair {a<b}: (i,j) = sorted tour positions; the contig at i is the
// row index of Q and the one at j the column, so the GArray looked up is
// orientedContacts[(first, second, Signs[first], Signs[second])]
// (orientation.go:145-150) = slots[e][dir][2*[sf=='-'] + [ss=='-']].
// The sentinel `continue` and the LIMIT `break` of the reference reduce to
// per-pair masks because dist is monotone in j for fixed i.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
eval_q_kernel(const int *__restrict__ tours, int shared_tour, int P, int L, int N, const long long *__restrict__ sizes,
              const unsigned char *__restrict__ signs, const int *__restrict__ pa, const int *__restrict__ pb,
              const int *__restrict__ slots, const int *__restrict__ hist, int E, double *__restrict__ out,
              unsigned char *gscratch, size_t gstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long scan_scratch[kWarps];
  __shared__ double red_scratch[kWarps];
  // the per-individual arrays live in shared memory, or -- groups beyond ~13 000 contigs -- in this CTA's slice of a
  // global scratch buffer (L2-resident; the CTA then walks its individuals p = blockIdx.x, + gridDim.x, ...)
  unsigned char *base = gscratch ? gscratch + static_cast<size_t>(blockIdx.x) * gstride : smem_raw;
  long long *cum = reinterpret_cast<long long *>(base);
  int *pos = reinterpret_cast<int *>(base + sizeof(long long) * static_cast<size_t>(L));
  int *tig = pos + N;
  unsigned char *sg = reinterpret_cast<unsigned char *>(tig + L);
  const int tid = threadIdx.x;
  for (int p = blockIdx.x; p < P; p += gridDim.x) {
  const int *src = tours + (shared_tour ? 0 : static_cast<size_t>(p) * L);
  const unsigned char *sgsrc = signs + static_cast<size_t>(p) * N;
  for (int i = tid; i < N; i += kBlock) { pos[i] = -1; sg[i] = sgsrc[i]; }
  for (int i = tid; i < L; i += kBlock) tig[i] = src[i];
  __syncthreads();
  {
    const int per = (L + kBlock - 1) / kBlock;
    const int beg = min(L, tid * per), end = min(L, beg + per);
    long long local = 0;
    for (int i = beg; i < end; ++i) local += __ldg(sizes + tig[i]);
    long long c = block_excl_scan(local, scan_scratch, tid);
    for (int i = beg; i < end; ++i) {
      cum[i] = c;                                  // cumsize[i], orientation.go:167-170
      pos[tig[i]] = i;
      c += __ldg(sizes + tig[i]);
    }
  }
  __syncthreads();
  double acc = 0.0;
  for (int e = tid; e < E; e += kBlock) {
    const int a = __ldg(pa + e), b = __ldg(pb + e);
    const int px = pos[a], py = pos[b];
    if (px < 0 || py < 0) continue;
    const int i = min(px, py), j = max(px, py);
    const long long dist = cum[j - 1] - cum[i];    // orientation.go:181 (sic)
    if (dist > kLimitI) continue;
    const int dir = px < py ? 0 : 1;
    const unsigned char sf = sg[dir ? b : a], ss = sg[dir ? a : b];
    if ((sf != '+' && sf != '-') || (ss != '+' && ss != '-')) continue;
    const int h = __ldg(slots + static_cast<size_t>(e) * 8 + dir * 4 + 2 * (sf == '-') + (ss == '-'));
    if (h < 0) continue;                           // Q[a][b][0] == -1
    const int4 *hp = reinterpret_cast<const int4 *>(hist + static_cast<size_t>(h) * 12);
    const int4 q0 = __ldg(hp), q1 = __ldg(hp + 1), q2 = __ldg(hp + 2);
    const int q[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
    for (int k = 0; k < 12; ++k)
      if (q[k] != 0) acc += static_cast<double>(q[k]) * log(static_cast<double>(c_GR[k] + dist));
  }
  const double s = block_sum(acc, red_scratch, tid);
  if (tid == 0) out[p] = 0.0 - s;
  __syncthreads();   // the arrays are reused by the CTA's next individual
  }
}

// ---------------------------------------------------------------------------
// Roofline probes.
// ---------------------------------------------------------------------------
__global__ void probe_dfma_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void probe_ddiv_kernel(double *out, int iters) {
  double d0 = 3.0 + threadIdx.x, d1 = d0 + 0.5, d2 = d0 + 0.25, d3 = d0 + 0.125;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int i = 0; i < iters; ++i) {
    s0 += 7.0 / d0; s1 += 7.0 / d1; s2 += 7.0 / d2; s3 += 7.0 / d3;
    d0 += 1.0; d1 += 1.0; d2 += 1.0; d3 += 1.0;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}
__global__ void probe_read_kernel(const int4 *__restrict__ in, size_t n, int reps, int4 *out) {
  int4 acc = make_int4(0, 0, 0, 0);
  for (int r = 0; r < reps; ++r)
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const int4 v = __ldcg(in + i);
      acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
  if (acc.x == 0x7fffffff) out[0] = acc;
}
__global__ void fill_kernel(int4 *p, size_t n, int v) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[i] = make_int4(v, v, v, v);
}

// ---------------------------------------------------------------------------
// Launch helpers.
// ---------------------------------------------------------------------------
size_t eval_m_smem_bytes(int T, int32_t N, int32_t L) {
  return sizeof(double) * static_cast<size_t>(N) * T + sizeof(int) * static_cast<size_t>(L);
}

static int pick_T(const ahx_ctx *ctx, int32_t P, int32_t L) {
  if (const char *env = getenv("AHX_EVAL_T")) {
    const int t = atoi(env);
    if ((t == 1 || t == 2) && eval_m_smem_bytes(t, ctx->N, L) <= ctx->smem_optin - 1024) return t;
  }
  // Amortise the COO stream over several tours while keeping >= 2 CTAs per SM
  // and enough CTAs to fill the machine.
  for (int t : {2}) {
    const size_t bytes = eval_m_smem_bytes(t, ctx->N, L);
    if (bytes * 2 + 4096 <= ctx->smem_optin && (P + t - 1) / t >= 4 * ctx->sm_count) return t;
  }
  return 1;
}

template <int T, typename Coo, typename IdxT, bool XG>
static cudaError_t launch_sparse_t(ahx_ctx *ctx, const IdxT *tours, const int32_t *rows, int32_t P, int32_t L,
                                   double *out, cudaStream_t s) {
  auto kern = eval_m_sparse_kernel<T, Coo, IdxT, XG>;
  const size_t smem = XG ? sizeof(int) * static_cast<size_t>(L) : eval_m_smem_bytes(T, ctx->N, L);
  cudaError_t e = opt_in_smem(kern, ctx->smem_optin);
  if (e != cudaSuccess) return e;
  const typename Coo::Elem *coo;
  if constexpr (sizeof(typename Coo::Elem) == 8) coo = reinterpret_cast<const typename Coo::Elem *>(ctx->d_coo16.p);
  else coo = reinterpret_cast<const typename Coo::Elem *>(ctx->d_coo32.p);
  const int grid = (P + T - 1) / T;
  double *xg = nullptr;
  if (XG) {
    e = ctx->d_xglobal.reserve(static_cast<size_t>(grid) * ctx->N * T);
    if (e != cudaSuccess) return e;
    xg = ctx->d_xglobal.p;
  }
  kern<<<grid, kBlock, smem, s>>>(tours, rows, P, L, ctx->N, ctx->d_sizes.p, coo, static_cast<int>(ctx->E), xg, out, ctx->d_errflag.p);
  ctx->launches++;
  return cudaGetLastError();
}

template <typename Coo, typename IdxT>
static cudaError_t launch_sparse_c(ahx_ctx *ctx, const IdxT *tours, const int32_t *rows, int32_t P, int32_t L,
                                   double *out, cudaStream_t s) {
  if (eval_m_smem_bytes(1, ctx->N, L) + 1024 > ctx->smem_optin)
    return launch_sparse_t<1, Coo, IdxT, true>(ctx, tours, rows, P, L, out, s);
  switch (pick_T(ctx, P, L)) {
    case 2: return launch_sparse_t<2, Coo, IdxT, false>(ctx, tours, rows, P, L, out, s);
    default: return launch_sparse_t<1, Coo, IdxT, false>(ctx, tours, rows, P, L, out, s);
  }
}

size_t eval_x32_smem_bytes(int T, bool coo16, int32_t N, int32_t L) {
  (void)L;
  return (coo16 ? sizeof(uint2) : sizeof(int4)) * (static_cast<size_t>(kStages) * kTile + kWarps * kQueue) +
         sizeof(uint32_t) * static_cast<size_t>(N) * (T + 1);
}

static int pick_T32(const ahx_ctx *ctx, int32_t P, int32_t L) {
  auto fits = [&](int t, int ctas) { return (eval_x32_smem_bytes(t, ctx->coo16, ctx->N, L) + 1536) * ctas <= ctx->smem_optin + 1024 * (ctas - 1) && eval_x32_smem_bytes(t, ctx->coo16, ctx->N, L) + 2048 <= ctx->smem_optin; };
  if (const char *env = getenv("AHX_EVAL_T")) {
    const int t = atoi(env);
    if ((t == 1 || t == 2 || t == 4) && fits(t, 1)) return t;
  }
  // several tours per CTA amortise the pair-table stream and make each shared-memory
  // gather serve T tours; keep >= 3 CTAs per SM and enough CTAs for two waves
  for (int t : {4, 2})
    if (fits(t, 3) && (P + t - 1) / t >= ctx->sm_count) return t;
  return 1;
}

template <int T, typename Coo, typename IdxT>
static cudaError_t launch_x32_t(ahx_ctx *ctx, const IdxT *tours, const int32_t *rows, int32_t P, int32_t L, double *out, cudaStream_t s) {
  auto kern = eval_m_x32_kernel<T, Coo, IdxT>;
  const size_t smem = eval_x32_smem_bytes(T, ctx->coo16, ctx->N, L);
  cudaError_t e = opt_in_smem(kern, ctx->smem_optin);
  if (e != cudaSuccess) return e;
  const typename Coo::Elem *coo;
  if constexpr (sizeof(typename Coo::Elem) == 8) coo = reinterpret_cast<const typename Coo::Elem *>(ctx->d_coo16o.p);
  else coo = reinterpret_cast<const typename Coo::Elem *>(ctx->d_coo32o.p);
  kern<<<(P + T - 1) / T, kBlock, smem, s>>>(tours, rows, P, L, ctx->N, ctx->d_sizes32.p, coo, static_cast<int>(ctx->E_pad), out, ctx->d_errflag.p);
  ctx->launches++;
  return cudaGetLastError();
}

template <typename Coo, typename IdxT>
static cudaError_t launch_x32_c(ahx_ctx *ctx, const IdxT *tours, const int32_t *rows, int32_t P, int32_t L, double *out, cudaStream_t s) {
  switch (pick_T32(ctx, P, L)) {
    case 4: return launch_x32_t<4, Coo, IdxT>(ctx, tours, rows, P, L, out, s);
    case 2: return launch_x32_t<2, Coo, IdxT>(ctx, tours, rows, P, L, out, s);
    default: return launch_x32_t<1, Coo, IdxT>(ctx, tours, rows, P, L, out, s);
  }
}

// ---- resident-table kernel ------------------------------------------------------
// Usable when the coordinates fit 32 bits, contig indices fit 16 bits and the
// packed table + two X buffers fit one SM's shared memory.  T: as many tours per
// batch as still leave one batch per SM.
// Which form of the persistent kernel can score P tours: T tours per batch, table resident
// in shared memory or streamed (T = 0: neither -- fall back to the x32 / fp64 kernels).
ResPlan resident_plan(const ahx_ctx *ctx, int32_t P) {
  ResPlan pl;
  if (const char *env = getenv("AHX_EVAL_RES")) if (!atoi(env)) return pl;
  if (const char *env = getenv("AHX_EVAL_FP64")) if (atoi(env)) return pl;
  if (!ctx->x32_ok || !ctx->coo16 || ctx->E_res[0] <= 0) return pl;
  auto fits = [&](int t, bool stream, bool sizes) {
    if (!stream && static_cast<uint64_t>(ctx->N) * 4u * t > 65535u) return false;
    return res_smem_bytes(t, ctx->N, ctx->E_res[res_ti(t)], stream, sizes) + 1024 <= ctx->smem_optin;
  };
  int forced = 0;
  if (const char *env = getenv("AHX_EVAL_T")) { const int t = atoi(env); if (t == 1 || t == 2 || t == 4) forced = t; }
  const bool no_stream = getenv("AHX_EVAL_STREAM") && !atoi(getenv("AHX_EVAL_STREAM"));
  const bool only_stream = getenv("AHX_EVAL_STREAM") && atoi(getenv("AHX_EVAL_STREAM")) == 1;
  const bool no_sizes = getenv("AHX_EVAL_SIZES") && !atoi(getenv("AHX_EVAL_SIZES"));
  // preference: the table resident, then the largest T; the contig sizes staged in shared memory whenever
  // they fit next to the rest (the producers' size gathers from global memory cost as many LSU cycles as
  // all of the consumers' shared-memory traffic) -- for a streamed table even at the price of a smaller T
  for (int stream = 0; stream < 2; ++stream) {
    if (stream ? no_stream : only_stream) continue;
    for (int pass = 0; pass < 2; ++pass)          // streamed: first insist on staged sizes, then take any T
      for (int t : {4, 2, 1}) {
        if (forced && t != forced) continue;
        if (!forced && t > 1 && (P + t - 1) / t < ctx->sm_count) continue;   // keep one batch per SM
        const bool sz = !no_sizes && fits(t, stream != 0, true);
        if (stream && pass == 0 && !sz && !no_sizes) continue;
        if (sz || fits(t, stream != 0, false)) { pl.T = t; pl.stream = stream != 0; pl.sizes = sz; return pl; }
      }
  }
  return pl;
}
int resident_T(const ahx_ctx *ctx, int32_t P) { return resident_plan(ctx, P).T; }

template <int T, bool STREAM, typename IdxT>
static cudaError_t launch_res_t(ahx_ctx *ctx, bool sizes, const IdxT *tours, const int32_t *rows, int32_t P, int32_t L, double *out, cudaStream_t s) {
  auto kern = L < ctx->N ? eval_m_res_kernel<T, true, STREAM, ResDirect<T, IdxT>> : eval_m_res_kernel<T, false, STREAM, ResDirect<T, IdxT>>;
  const size_t smem = res_smem_bytes(T, ctx->N, ctx->E_res[res_ti(T)], STREAM, sizes);
  cudaError_t e = opt_in_smem(kern, ctx->smem_optin);
  if (e != cudaSuccess) return e;
  const bool scaled = static_cast<uint64_t>(ctx->N) * 4u * T <= 65535u;   // the table of byte offsets exists (ahx_load)
  ResArgs g{(STREAM && !scaled ? ctx->d_rraw : ctx->d_rpairs)[res_ti(T)].p, ctx->d_rnl8[res_ti(T)].p, ctx->d_sizes32.p, static_cast<int>(ctx->E_res[res_ti(T)]), ctx->N, L, ctx->d_errflag.p, sizes ? 1 : 0, scaled ? 1 : 0};
  ResDirect<T, IdxT> src{};
  src.tours = tours; src.rows = rows; src.P = P; src.L = L; src.out = out;
  const int nb = (P + T - 1) / T;
  kern<<<std::min(nb, ctx->sm_count), kRThreads, smem, s>>>(g, src);
  ctx->launches++;
  return cudaGetLastError();
}

template <typename IdxT>
static cudaError_t launch_res(ahx_ctx *ctx, const ResPlan &pl, const IdxT *tours, const int32_t *rows, int32_t P, int32_t L, double *out, cudaStream_t s) {
  if (pl.stream) {
    switch (pl.T) {
      case 4: return launch_res_t<4, true, IdxT>(ctx, pl.sizes, tours, rows, P, L, out, s);
      case 2: return launch_res_t<2, true, IdxT>(ctx, pl.sizes, tours, rows, P, L, out, s);
      default: return launch_res_t<1, true, IdxT>(ctx, pl.sizes, tours, rows, P, L, out, s);
    }
  }
  switch (pl.T) {
    case 4: return launch_res_t<4, false, IdxT>(ctx, pl.sizes, tours, rows, P, L, out, s);
    case 2: return launch_res_t<2, false, IdxT>(ctx, pl.sizes, tours, rows, P, L, out, s);
    default: return launch_res_t<1, false, IdxT>(ctx, pl.sizes, tours, rows, P, L, out, s);
  }
}

// Which formulation can run: 32-bit coordinates need 2*sum(sizes) + 2*LIMIT + 2 < 2^32
// and N*4 bytes of shared memory; otherwise the fp64 kernel (x in shared or global memory).
bool use_x32(const ahx_ctx *ctx, int32_t L) {
  if (const char *env = getenv("AHX_EVAL_FP64")) if (atoi(env)) return false;
  return ctx->x32_ok && eval_x32_smem_bytes(1, ctx->coo16, ctx->N, L) + 2048 <= ctx->smem_optin;
}

int launch_eval_m_sparse(ahx_ctx *ctx, const int32_t *d_tours, const uint16_t *d_tours16, const int32_t *d_rows,
                         int32_t P, int32_t L, double *d_out, cudaStream_t s) {
  if (P == 0) return AHX_OK;
  cudaError_t e;
  if (const ResPlan pl = resident_plan(ctx, P); pl.T) {
    e = d_tours16 ? launch_res<uint16_t>(ctx, pl, d_tours16, d_rows, P, L, d_out, s) : launch_res<int32_t>(ctx, pl, d_tours, d_rows, P, L, d_out, s);
  } else if (use_x32(ctx, L)) {
    if (d_tours16) e = ctx->coo16 ? launch_x32_c<Coo16, uint16_t>(ctx, d_tours16, d_rows, P, L, d_out, s) : launch_x32_c<Coo32, uint16_t>(ctx, d_tours16, d_rows, P, L, d_out, s);
    else e = ctx->coo16 ? launch_x32_c<Coo16, int32_t>(ctx, d_tours, d_rows, P, L, d_out, s) : launch_x32_c<Coo32, int32_t>(ctx, d_tours, d_rows, P, L, d_out, s);
  } else if (d_tours16) {
    e = ctx->coo16 ? launch_sparse_c<Coo16, uint16_t>(ctx, d_tours16, d_rows, P, L, d_out, s)
                   : launch_sparse_c<Coo32, uint16_t>(ctx, d_tours16, d_rows, P, L, d_out, s);
  } else {
    e = ctx->coo16 ? launch_sparse_c<Coo16, int32_t>(ctx, d_tours, d_rows, P, L, d_out, s)
                   : launch_sparse_c<Coo32, int32_t>(ctx, d_tours, d_rows, P, L, d_out, s);
  }
  if (e != cudaSuccess) return cuda_fail(ctx, e, "eval_m kernel launch");
  return AHX_OK;
}

static int ensure_dense(ahx_ctx *ctx) {
  if (ctx->dense_ready) return AHX_OK;
  const size_t nn = static_cast<size_t>(ctx->N) * ctx->N;
  AHX_CUDA(ctx, ctx->d_M.reserve(nn));
  AHX_CUDA(ctx, cudaMemsetAsync(ctx->d_M.p, 0, nn * sizeof(int32_t), ctx->stream));  // Make2DSlice zeros, base.go:279
  if (ctx->E > 0) {
    ahx::DevBuf<int32_t> d_nl;
    AHX_CUDA(ctx, d_nl.reserve(ctx->E));
    AHX_CUDA(ctx, cudaMemcpyAsync(d_nl.p, ctx->h_nl.data(), ctx->E * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = static_cast<int>(std::min<int64_t>((ctx->E + 255) / 256, 4096));
    build_dense_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_pa.p, ctx->d_pb.p, d_nl.p, ctx->E, ctx->N, ctx->d_M.p);
    ctx->launches++;
    AHX_CUDA(ctx, cudaGetLastError());
    AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    d_nl.release();
  }
  ctx->dense_ready = true;
  return AHX_OK;
}

int launch_eval_m_dense(ahx_ctx *ctx, const int32_t *d_tours, int32_t P, int32_t L, double *d_out, cudaStream_t s) {
  if (P == 0) return AHX_OK;
  int rc = ensure_dense(ctx);
  if (rc != AHX_OK) return rc;
  const size_t smem = (sizeof(double) + sizeof(int)) * static_cast<size_t>(L);
  if (smem + 1024 > ctx->smem_optin) return fail(ctx, AHX_ERR_ARG, "ahx_eval_m_dense: tour too long for the dense kernel's shared memory");
  AHX_CUDA(ctx, opt_in_smem(eval_m_dense_kernel<int32_t>, ctx->smem_optin));
  eval_m_dense_kernel<int32_t><<<P, kBlock, smem, s>>>(d_tours, P, L, ctx->N, ctx->d_sizes.p, ctx->d_M.p, d_out, ctx->d_errflag.p);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  return AHX_OK;
}

int launch_eval_q(ahx_ctx *ctx, const int32_t *d_tours, int shared_tour, const uint8_t *d_signs, int32_t P, int32_t L,
                  double *d_out, cudaStream_t s) {
  if (P == 0) return AHX_OK;
  const size_t smem = sizeof(long long) * static_cast<size_t>(L) + sizeof(int) * (static_cast<size_t>(ctx->N) + L) + ctx->N;
  AHX_CUDA(ctx, opt_in_smem(eval_q_kernel, ctx->smem_optin));
  const bool force_global = getenv("AHX_EVALQ_GLOBAL") && atoi(getenv("AHX_EVALQ_GLOBAL"));   // test knob
  if (smem + 1024 > ctx->smem_optin || force_global) {
    // 17 N bytes per individual do not fit an SM's shared memory: one slice of a global scratch buffer per CTA
    const size_t stride = (smem + 255) & ~static_cast<size_t>(255);
    const int grid = std::min<int>(P, 4 * ctx->sm_count);
    AHX_CUDA(ctx, ctx->d_qscratch.reserve(stride * grid));
    eval_q_kernel<<<grid, kBlock, 0, s>>>(d_tours, shared_tour, P, L, ctx->N, ctx->d_sizes.p, d_signs, ctx->d_pa.p, ctx->d_pb.p,
                                           ctx->d_slots.p, ctx->d_hist.p, static_cast<int>(ctx->E), d_out, ctx->d_qscratch.p, stride);
  } else {
    eval_q_kernel<<<P, kBlock, smem, s>>>(d_tours, shared_tour, P, L, ctx->N, ctx->d_sizes.p, d_signs, ctx->d_pa.p, ctx->d_pb.p,
                                          ctx->d_slots.p, ctx->d_hist.p, static_cast<int>(ctx->E), d_out, nullptr, 0);
  }
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  return AHX_OK;
}

// Waits for the context stream and turns the device-side range-check flag into an error.
int finish_checked(ahx_ctx *ctx, const char *who) {
  int flag = 0;
  AHX_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_errflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (flag) {
    AHX_CUDA(ctx, cudaMemsetAsync(ctx->d_errflag.p, 0, sizeof(int), ctx->stream));
    return fail(ctx, AHX_ERR_ARG, std::string(who) + ": contig index out of range in a tour");
  }
  return AHX_OK;
}

// Tour.Tigs must hold distinct contig indices (they come from tigToIdx).
int validate_tour(ahx_ctx *ctx, int32_t L, const int32_t *tour, const char *who) {
  if (L < 0 || L > ctx->N || (L > 0 && !tour)) return fail(ctx, AHX_ERR_ARG, std::string(who) + ": bad tour length");
  std::vector<uint8_t> seen(ctx->N, 0);
  for (int32_t i = 0; i < L; ++i) {
    const int32_t c = tour[i];
    if (c < 0 || c >= ctx->N || seen[c]) return fail(ctx, AHX_ERR_ARG, std::string(who) + ": tour is not a set of distinct contig indices");
    seen[c] = 1;
  }
  return AHX_OK;
}

template <typename IdxT>
static int validate_tours(ahx_ctx *ctx, int32_t P, int32_t L, const IdxT *tours, const char *who) {
  if (P < 0 || L < 0 || L > ctx->N || (static_cast<int64_t>(P) * L > 0 && !tours)) return fail(ctx, AHX_ERR_ARG, std::string(who) + ": bad population shape");
  // Contig indices are range-checked on the device (errflag); a duplicated index
  // inside a tour is the caller's bug, as it is in the reference.
  return AHX_OK;
}


// Bank-conflict-free ordering of the pair table for the x32 kernel.  With T = 4
// tours per CTA a contig's coordinates are one 16-byte shared-memory word and a
// 128-bit load is served one quarter-warp (8 lanes) at a time across 8 bank
// groups, group = contig mod 8.  Pairs are therefore emitted in blocks of 8 in
// which all `a mod 8` and all `b mod 8` are distinct (greedy over the 64
// (a mod 8, b mod 8) buckets, largest bucket first); lanes of a block that cannot be
// filled that way take conflicting pairs rather than padding.
// The order of the terms only changes the (fixed) summation order.
// W: lanes that are served together by one shared-memory wavefront of the coordinate gather =
// 32 / T (a contig's T coordinates are 4T bytes: 8 lanes x 16 B, 16 x 8 B or 32 x 4 B per 128 B);
// within every aligned group of W entries all `a mod W` and all `b mod W` are distinct.
void order_pairs(int64_t E, const int32_t *pa, const int32_t *pb, std::vector<int64_t> *order /* -1 = padding */, int W) {
  std::vector<std::vector<int64_t>> bucket(static_cast<size_t>(W) * W);
  for (int64_t e = E - 1; e >= 0; --e) bucket[static_cast<size_t>(pa[e] % W) * W + (pb[e] % W)].push_back(e);   // pop_back yields ascending e
  order->clear();
  order->reserve(static_cast<size_t>(E + E / 8 + 512));
  int64_t left = E;
  int rot = 0;
  std::vector<char> used_b(W);
  while (left > 0) {
    std::fill(used_b.begin(), used_b.end(), 0);
    int filled = 0;
    for (int i = 0; i < W; ++i) {
      const int r = (i + rot) % W;
      int best = -1; size_t best_n = 0;
      for (int cc = 0; cc < W; ++cc) {
        const size_t n = bucket[static_cast<size_t>(r) * W + cc].size();
        if (!used_b[cc] && n > best_n) { best = cc; best_n = n; }
      }
      if (best >= 0) {
        order->push_back(bucket[static_cast<size_t>(r) * W + best].back());
        bucket[static_cast<size_t>(r) * W + best].pop_back();
        used_b[best] = 1; --left; ++filled;
      }
    }
    // lanes without a conflict-free candidate take the pair that collides with the fewest lanes of
    // the block (ties: fullest bucket): a bank conflict costs one extra wavefront of one gather, a
    // padding entry a whole term per tour (the padded tables were 9 % longer than the pair lists at
    // config3 and config5, 33 % at config2)
    if (filled < W && left > 0) {
      std::vector<int> na(W, 0), nb(W, 0);
      for (size_t i = order->size() - filled; i < order->size(); ++i) { ++na[pa[(*order)[i]] % W]; ++nb[pb[(*order)[i]] % W]; }
      for (; filled < W && left > 0; ++filled, --left) {
        int best = -1, best_cost = 0; size_t best_n = 0;
        for (int r = 0; r < W; ++r)
          for (int cc = 0; cc < W; ++cc) {
            const size_t n = bucket[static_cast<size_t>(r) * W + cc].size();
            if (!n) continue;
            const int cost = na[r] + nb[cc];
            if (best < 0 || cost < best_cost || (cost == best_cost && n > best_n)) { best = r * W + cc; best_cost = cost; best_n = n; }
          }
        order->push_back(bucket[best].back());
        bucket[best].pop_back();
        ++na[best / W]; ++nb[best % W];
      }
    }
    for (; filled < W; ++filled) order->push_back(-1);
    rot = (rot + 1) % W;
  }
  while (order->size() % kBlock) order->push_back(-1);
}
}  // namespace ahx

using namespace ahx;

// ===========================================================================
// C ABI
// ===========================================================================
// ---- batched Tour.Evaluate through host buffers --------------------------------
// Chunks the population so the H2D copy of chunk k+1 overlaps the kernel of chunk k.
template <typename IdxT>
static int eval_m_host(ahx_ctx *ctx, int32_t P, int32_t L, const IdxT *tours, double *out, int kind, const char *who) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if ((rc = validate_tours(ctx, P, L, tours, who)) != AHX_OK) return rc;
  if (P > 0 && !out) return fail(ctx, AHX_ERR_ARG, std::string(who) + ": null out");
  if (P == 0) return AHX_OK;
  constexpr bool k16 = sizeof(IdxT) == 2;
  if (k16 && ctx->N > 65535) return fail(ctx, AHX_ERR_ARG, "ahx_eval_m_u16: n_tigs exceeds 65535");
  const size_t total = static_cast<size_t>(P) * L;
  IdxT *dt;
  if constexpr (k16) { AHX_CUDA(ctx, ctx->d_tours16.reserve(total)); dt = ctx->d_tours16.p; }
  else { AHX_CUDA(ctx, ctx->d_tours.reserve(total)); dt = ctx->d_tours.p; }
  AHX_CUDA(ctx, ctx->d_scores.reserve(P));
  // ~4 MB per chunk, at most 16 chunks, chunk sizes multiple of 4 tours
  int nchunk = static_cast<int>(std::min<size_t>(16, std::max<size_t>(1, total * sizeof(IdxT) / (4u << 20))));
  if (const char *env = getenv("AHX_EVAL_CHUNKS")) nchunk = std::max(1, std::min(64, atoi(env)));
  int32_t per = ((P + nchunk - 1) / nchunk + 3) & ~3;
  nchunk = (P + per - 1) / per;
  while (static_cast<int>(ctx->chunk_ev.size()) < nchunk) {
    cudaEvent_t ev;
    AHX_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ctx->chunk_ev.push_back(ev);
  }
  // the copy stream must not overwrite a buffer an earlier call is still using
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  AHX_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_a, 0));
  for (int c = 0; c < nchunk; ++c) {
    const int32_t p0 = c * per, pc = std::min(per, P - p0);
    AHX_CUDA(ctx, cudaMemcpyAsync(dt + static_cast<size_t>(p0) * L, tours + static_cast<size_t>(p0) * L,
                                  sizeof(IdxT) * static_cast<size_t>(pc) * L, cudaMemcpyHostToDevice, ctx->copy_stream));
    AHX_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream));
    AHX_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0));
    if (kind == 0) {
      if constexpr (k16) rc = launch_eval_m_sparse(ctx, nullptr, dt + static_cast<size_t>(p0) * L, nullptr, pc, L, ctx->d_scores.p + p0, ctx->stream);
      else rc = launch_eval_m_sparse(ctx, dt + static_cast<size_t>(p0) * L, nullptr, nullptr, pc, L, ctx->d_scores.p + p0, ctx->stream);
    } else {
      if constexpr (k16) return fail(ctx, AHX_ERR_ARG, "dense kernel takes int32 tours");
      else rc = launch_eval_m_dense(ctx, dt + static_cast<size_t>(p0) * L, pc, L, ctx->d_scores.p + p0, ctx->stream);
    }
    if (rc != AHX_OK) return rc;
  }
  AHX_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_scores.p, sizeof(double) * P, cudaMemcpyDeviceToHost, ctx->stream));
  return finish_checked(ctx, who);
}


extern "C" {

int ahx_version(void) { return AHX_VERSION; }
const char *ahx_last_error(const ahx_ctx *ctx) { return ctx ? ctx->err.c_str() : global_error(); }

int ahx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int ahx_create(int device, ahx_ctx **out) {
  if (!out) return fail(nullptr, AHX_ERR_ARG, "ahx_create: null out");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, AHX_ERR_NODEVICE, "ahx_create: no CUDA device available; libahx has no CPU path");
  }
  if (device < 0 || device >= n) return fail(nullptr, AHX_ERR_ARG, "ahx_create: device index out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major < 10) return fail(nullptr, AHX_ERR_NODEVICE, "ahx_create: libahx is built for sm_100a (Blackwell B200) only");
  auto *ctx = new ahx_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->l2_bytes = static_cast<size_t>(prop.l2CacheSize);
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev_a)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev_b)) != cudaSuccess ||
      (e = cudaHostAlloc(&ctx->h_pin, 4096, cudaHostAllocDefault)) != cudaSuccess) {
    int rc = cuda_fail(nullptr, e, "ahx_create");
    delete ctx;
    return rc;
  }
  if ((e = ctx->d_errflag.reserve(1)) != cudaSuccess || (e = cudaMemset(ctx->d_errflag.p, 0, sizeof(int))) != cudaSuccess) {
    int rc = cuda_fail(nullptr, e, "ahx_create");
    delete ctx;
    return rc;
  }
  *out = ctx;
  return AHX_OK;
}

void ahx_destroy(ahx_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);          // this context's work only: other contexts of the process keep running
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  for (auto &kv : ctx->isl_opened) cudaIpcCloseMemHandle(kv.second);
  ctx->d_mbox.release(); ctx->d_isl.release(); ctx->d_qscratch.release();
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  ctx->d_sizes.release(); ctx->d_sizes32.release(); ctx->d_coo16.release(); ctx->d_coo32.release(); ctx->d_coo16o.release(); ctx->d_coo32o.release(); for (auto &b : ctx->d_rpairs) b.release(); for (auto &b : ctx->d_rraw) b.release(); for (auto &b : ctx->d_rnl8) b.release(); ctx->d_pa.release(); ctx->d_pb.release();
  ctx->d_slots.release(); ctx->d_hist.release(); ctx->d_M.release(); ctx->d_adj_ptr.release(); ctx->d_adj_e.release();
  ctx->d_tours.release(); ctx->d_res_tours.release(); ctx->d_res_scores.release(); ctx->d_tours16.release(); ctx->d_scores.release(); ctx->d_signs.release(); ctx->d_xglobal.release();
  ctx->d_flush.release(); ctx->d_errflag.release();
  for (int i = 0; i < 2; ++i) { ctx->d_pop[i].release(); ctx->d_fit[i].release(); }
  ctx->d_hof.release(); ctx->d_sel_idx.release(); ctx->d_sel_tours.release(); ctx->d_sel_fit.release(); ctx->d_taken.release();
  ctx->d_state.release(); ctx->d_plan.release(); ctx->d_pos.release(); ctx->d_cum.release(); ctx->d_S4.release(); ctx->d_trace.release();
  ctx->d_pairinfo.release(); ctx->d_stepacc.release(); ctx->d_counts.release();
  ctx->d_rowof[0].release(); ctx->d_rowof[1].release(); ctx->d_sel_rows.release(); ctx->d_row_fit.release(); ctx->d_gbar.release(); ctx->d_used.release(); ctx->d_gmin.release(); ctx->d_xpool.release(); ctx->d_gmask.release(); ctx->d_step_off.release(); ctx->d_rec_o.release(); ctx->d_rec_d.release();
  for (auto ev : ctx->chunk_ev) cudaEventDestroy(ev);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

int ahx_host_alloc(void **ptr, uint64_t bytes) {
  if (!ptr) return fail(nullptr, AHX_ERR_ARG, "ahx_host_alloc: null");
  cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc");
  return AHX_OK;
}
int ahx_host_free(void *ptr) {
  if (ptr && cudaFreeHost(ptr) != cudaSuccess) return fail(nullptr, AHX_ERR_CUDA, "cudaFreeHost failed");
  return AHX_OK;
}

int ahx_load(ahx_ctx *ctx, int32_t N, const int64_t *sizes, int64_t E, const int32_t *pa, const int32_t *pb,
             const int32_t *nl, const int8_t *strand, const int32_t *slots, int64_t H, const int32_t *hist) {
  int rc = enter(ctx, false);
  if (rc != AHX_OK) return rc;
  if (ctx->ga_active) return fail(ctx, AHX_ERR_STATE, "ahx_load: a GA run is active (call ahx_ga_end first)");
  if (N <= 0 || !sizes || E < 0 || H < 0 || E > INT32_MAX / 8 || (E > 0 && (!pa || !pb || !nl || !slots)) || (H > 0 && !hist))
    return fail(ctx, AHX_ERR_ARG, "ahx_load: bad argument");
  for (int32_t i = 0; i < N; ++i)
    if (sizes[i] < 1) return fail(ctx, AHX_ERR_ARG, "ahx_load: contig sizes must be >= 1 (a zero-length contig makes the reference divide by zero)");
  for (int64_t e = 0; e < E; ++e) {
    if (pa[e] < 0 || pa[e] >= pb[e] || pb[e] >= N) return fail(ctx, AHX_ERR_ARG, "ahx_load: pairs must satisfy 0 <= a < b < n_tigs");
    if (e > 0 && (pa[e] < pa[e - 1] || (pa[e] == pa[e - 1] && pb[e] <= pb[e - 1])))
      return fail(ctx, AHX_ERR_ARG, "ahx_load: pairs must be unique and sorted by (a, b)");
    if (nl[e] < 0) return fail(ctx, AHX_ERR_ARG, "ahx_load: negative nlinks");
    for (int k = 0; k < 8; ++k)
      if (slots[e * 8 + k] < -1 || slots[e * 8 + k] >= H) return fail(ctx, AHX_ERR_ARG, "ahx_load: slot index out of range");
  }
  ctx->loaded = false; ctx->dense_ready = false;
  ctx->N = N; ctx->E = E; ctx->H = H;
  ctx->h_sizes.assign(sizes, sizes + N);
  ctx->h_pa.assign(pa, pa + E); ctx->h_pb.assign(pb, pb + E); ctx->h_nl.assign(nl, nl + E);
  if (strand) ctx->h_strand.assign(strand, strand + E); else ctx->h_strand.assign(E, 1);
  ctx->coo16 = N <= 65535;
  AHX_CUDA(ctx, ctx->d_sizes.reserve(N));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_sizes.p, sizes, sizeof(int64_t) * N, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, ctx->d_pa.reserve(E)); AHX_CUDA(ctx, ctx->d_pb.reserve(E));
  AHX_CUDA(ctx, ctx->d_slots.reserve(E * 8)); AHX_CUDA(ctx, ctx->d_hist.reserve(H * 12));
  std::vector<uint2> c16; std::vector<int4> c32;
  if (ctx->coo16) {
    c16.resize(E);
    for (int64_t e = 0; e < E; ++e) c16[e] = make_uint2(static_cast<uint32_t>(pa[e]) | (static_cast<uint32_t>(pb[e]) << 16), static_cast<uint32_t>(nl[e]));
    AHX_CUDA(ctx, ctx->d_coo16.reserve(E));
    if (E) AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_coo16.p, c16.data(), sizeof(uint2) * E, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    c32.resize(E);
    for (int64_t e = 0; e < E; ++e) c32[e] = make_int4(pa[e], pb[e], nl[e], 0);
    AHX_CUDA(ctx, ctx->d_coo32.reserve(E));
    if (E) AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_coo32.p, c32.data(), sizeof(int4) * E, cudaMemcpyHostToDevice, ctx->stream));
  }
  {
    long long total = 0;
    for (int32_t i = 0; i < N; ++i) total += sizes[i];
    ctx->x32_ok = 2 * total + 2LL * AHX_LIMIT + 2 < 0xFFFFFFFFLL;
    std::vector<uint32_t> s32(static_cast<size_t>(N) + 1, 0u);   // entry N = 0: the "skip" contig of res_build_x
    for (int32_t i = 0; i < N; ++i) s32[i] = static_cast<uint32_t>(std::min<int64_t>(sizes[i], 0xFFFFFFFFLL));
    AHX_CUDA(ctx, ctx->d_sizes32.reserve(static_cast<size_t>(N) + 1));
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_sizes32.p, s32.data(), sizeof(uint32_t) * (static_cast<size_t>(N) + 1), cudaMemcpyHostToDevice, ctx->stream));
    AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> ord;
    order_pairs(E, pa, pb, &ord, 8);
    ctx->E_pad = static_cast<int64_t>(ord.size());
    if (ctx->coo16) {
      std::vector<uint2> o16(ord.size());
      for (size_t i = 0; i < ord.size(); ++i) o16[i] = ord[i] < 0 ? make_uint2(0, 0) : c16[ord[i]];
      AHX_CUDA(ctx, ctx->d_coo16o.reserve(ord.size()));
      if (!ord.empty()) AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_coo16o.p, o16.data(), sizeof(uint2) * ord.size(), cudaMemcpyHostToDevice, ctx->stream));
      AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
      std::vector<int4> o32(ord.size());
      for (size_t i = 0; i < ord.size(); ++i) o32[i] = ord[i] < 0 ? make_int4(0, 0, 0, 0) : c32[ord[i]];
      AHX_CUDA(ctx, ctx->d_coo32o.reserve(ord.size()));
      if (!ord.empty()) AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_coo32o.p, o32.data(), sizeof(int4) * ord.size(), cudaMemcpyHostToDevice, ctx->stream));
      AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // resident-table form (ahx_resident.cuh): 8-bit link counts, so a pair with more
    // than 255 links becomes several entries; ordered and padded like the list above
    for (auto &er : ctx->E_res) er = 0;
    if (ctx->coo16 && E > 0) {
      std::vector<int32_t> xa, xb, xn;
      xa.reserve(E); xb.reserve(E); xn.reserve(E);
      for (int64_t e = 0; e < E; ++e)
        for (int32_t left = nl[e]; left > 0; left -= 255) { xa.push_back(pa[e]); xb.push_back(pb[e]); xn.push_back(std::min(left, 255)); }
      // storage order: consumer thread tid reads its kRUnroll entries of iteration I as one vector, so the
      // logical entry I*kRPad + j*kRCT + tid lives at (I*kRCT + tid)*kRUnroll + j
      auto phys = [](size_t i) { const size_t I = i / kRPad, r = i % kRPad; return (I * kRCT + r % kRCT) * kRUnroll + r / kRCT; };
      for (int ti = 0; ti < 3 && xa.size() <= static_cast<size_t>(8) * E + 4096; ++ti) {   // (absurd link counts: leave it to the fallback kernels)
        // one ordering per tours-per-batch T = 1, 2, 4: conflict-free groups of 32 / T lanes
        const uint32_t T = 1u << ti, stride = 4u * T;
        std::vector<int64_t> rord;
        order_pairs(static_cast<int64_t>(xa.size()), xa.data(), xb.data(), &rord, static_cast<int>(32u / T));
        const size_t er = (rord.size() + kRPad - 1) / kRPad * kRPad;
        std::vector<uint8_t> r8p(er + 16, 0);
        std::vector<uint32_t> ra(er, 0u), rb(er, N > 1 ? 1u : 0u);   // padding: contigs (0, 1) with 0 links
        for (size_t i = 0; i < rord.size(); ++i) {
          if (rord[i] < 0) continue;
          const int64_t x = rord[i];
          ra[i] = static_cast<uint32_t>(xa[x]); rb[i] = static_cast<uint32_t>(xb[x]);
          r8p[phys(i)] = static_cast<uint8_t>(xn[x]);
        }
        AHX_CUDA(ctx, ctx->d_rnl8[ti].reserve(er + 16));
        AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_rnl8[ti].p, r8p.data(), er + 16, cudaMemcpyHostToDevice, ctx->stream));
        std::vector<uint32_t> rp(er);
        if (static_cast<uint64_t>(N) * stride <= 65535u) {   // resident mode: entries are byte offsets into X
          for (size_t i = 0; i < er; ++i) rp[phys(i)] = (ra[i] * stride) | ((rb[i] * stride) << 16);
          AHX_CUDA(ctx, ctx->d_rpairs[ti].reserve(er));
          AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_rpairs[ti].p, rp.data(), 4 * er, cudaMemcpyHostToDevice, ctx->stream));
          AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // rp is reused
        }
        // streamed mode: plain contig indices a | b << 16 (N * 4T may exceed 16 bits)
        for (size_t i = 0; i < er; ++i) rp[phys(i)] = ra[i] | (rb[i] << 16);
        AHX_CUDA(ctx, ctx->d_rraw[ti].reserve(er));
        AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_rraw[ti].p, rp.data(), 4 * er, cudaMemcpyHostToDevice, ctx->stream));
        AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->E_res[ti] = static_cast<int64_t>(er);
      }
    }
  }
  // CSR of pairs touching each contig (for flipOne's O(deg) deltas)
  std::vector<int32_t> ptr(N + 1, 0), adj(2 * E);
  for (int64_t e = 0; e < E; ++e) { ptr[pa[e] + 1]++; ptr[pb[e] + 1]++; }
  ctx->h_deg.assign(ptr.begin() + 1, ptr.end());
  for (int32_t i = 0; i < N; ++i) ptr[i + 1] += ptr[i];
  {
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < E; ++e) { adj[fill[pa[e]]++] = static_cast<int32_t>(e); adj[fill[pb[e]]++] = static_cast<int32_t>(e); }
  }
  AHX_CUDA(ctx, ctx->d_adj_ptr.reserve(N + 1)); AHX_CUDA(ctx, ctx->d_adj_e.reserve(2 * E));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_adj_ptr.p, ptr.data(), sizeof(int32_t) * (N + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (E) {
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_adj_e.p, adj.data(), sizeof(int32_t) * 2 * E, cudaMemcpyHostToDevice, ctx->stream));
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_pa.p, pa, 4 * E, cudaMemcpyHostToDevice, ctx->stream));
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_pb.p, pb, 4 * E, cudaMemcpyHostToDevice, ctx->stream));
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_slots.p, slots, 4 * 8 * E, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (H) AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_hist.p, hist, 4 * 12 * H, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->loaded = true;
  ctx->pop_P = ctx->pop_L = 0;
  ctx->h_diag.clear();
  return AHX_OK;
}

int ahx_load_clm(ahx_ctx *ctx, const ahx_clm *clm) {
  if (!ctx || !clm) return fail(ctx, AHX_ERR_ARG, "ahx_load_clm: null argument");
  const ClmView v = view_of(clm);
  if (!v.finalized) return fail(ctx, AHX_ERR_STATE, "ahx_load_clm: call ahx_clm_finalize first");
  const int rc = ahx_load(ctx, v.n, v.sizes, v.E, v.pa, v.pb, v.nl, v.strand, v.slots, v.H, v.hist);
  if (rc == AHX_OK && v.diag) ctx->h_diag.assign(v.diag, v.diag + v.n);   // self-pair lines: the diagonal of O (flipAll)
  return rc;
}

int ahx_eval_m(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tours, double *out) { return eval_m_host<int32_t>(ctx, P, L, tours, out, 0, "ahx_eval_m"); }
int ahx_eval_m_u16(ahx_ctx *ctx, int32_t P, int32_t L, const uint16_t *tours, double *out) { return eval_m_host<uint16_t>(ctx, P, L, tours, out, 0, "ahx_eval_m_u16"); }
int ahx_eval_m_dense(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tours, double *out) { return eval_m_host<int32_t>(ctx, P, L, tours, out, 1, "ahx_eval_m_dense"); }

int ahx_eval_q(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tours, int32_t shared_tour, const uint8_t *signs, double *out) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  const int32_t PT = shared_tour ? (P > 0 ? 1 : 0) : P;
  if ((rc = validate_tours(ctx, PT, L, tours, "ahx_eval_q")) != AHX_OK) return rc;
  if (P > 0 && (!signs || !out)) return fail(ctx, AHX_ERR_ARG, "ahx_eval_q: null signs/out");
  if (P == 0) return AHX_OK;
  // pos[] is an inverse permutation: a duplicated contig would silently drop terms
  for (int32_t p = 0; p < PT; ++p)
    if ((rc = validate_tour(ctx, L, tours + static_cast<size_t>(p) * L, "ahx_eval_q")) != AHX_OK) return rc;
  AHX_CUDA(ctx, ctx->d_tours.reserve(static_cast<size_t>(PT) * L));
  AHX_CUDA(ctx, ctx->d_signs.reserve(static_cast<size_t>(P) * ctx->N));
  AHX_CUDA(ctx, ctx->d_scores.reserve(P));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_tours.p, tours, sizeof(int32_t) * static_cast<size_t>(PT) * L, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_signs.p, signs, static_cast<size_t>(P) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  if (shared_tour && static_cast<size_t>(ctx->N) + 1024 <= ctx->smem_optin) {
    // one tour, many sign vectors (flipWhole / flipAll / batched orientation search): logs once, then lookups
    if ((rc = eval_q_shared_tour(ctx, P, L, tours, ctx->d_signs.p, ctx->d_scores.p)) != AHX_OK) return rc;   // brackets its kernels with ev_a / ev_b
  } else {
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    if ((rc = launch_eval_q(ctx, ctx->d_tours.p, shared_tour, ctx->d_signs.p, P, L, ctx->d_scores.p, ctx->stream)) != AHX_OK) return rc;
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  }
  AHX_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_scores.p, sizeof(double) * P, cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  AHX_CUDA(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev_a, ctx->ev_b));
  return AHX_OK;
}

// ---- resident population ----------------------------------------------------------
int ahx_pop_set(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tours) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if ((rc = validate_tours(ctx, P, L, tours, "ahx_pop_set")) != AHX_OK) return rc;
  AHX_CUDA(ctx, ctx->d_res_tours.reserve(static_cast<size_t>(P) * L));
  AHX_CUDA(ctx, ctx->d_res_scores.reserve(P));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_res_tours.p, tours, sizeof(int32_t) * static_cast<size_t>(P) * L, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->pop_P = P; ctx->pop_L = L;
  return AHX_OK;
}

int ahx_pop_eval_m(ahx_ctx *ctx, int32_t kind, int32_t repeat, float *elapsed_ms) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (ctx->pop_P <= 0) return fail(ctx, AHX_ERR_STATE, "ahx_pop_eval_m: no resident population (call ahx_pop_set)");
  if (repeat < 1 || (kind != 0 && kind != 1)) return fail(ctx, AHX_ERR_ARG, "ahx_pop_eval_m: bad argument");
  if (kind == 1 && (rc = ensure_dense(ctx)) != AHX_OK) return rc;
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  for (int r = 0; r < repeat; ++r) {
    rc = kind == 0 ? launch_eval_m_sparse(ctx, ctx->d_res_tours.p, nullptr, nullptr, ctx->pop_P, ctx->pop_L, ctx->d_res_scores.p, ctx->stream)
                   : launch_eval_m_dense(ctx, ctx->d_res_tours.p, ctx->pop_P, ctx->pop_L, ctx->d_res_scores.p, ctx->stream);
    if (rc != AHX_OK) return rc;
  }
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  AHX_CUDA(ctx, cudaEventSynchronize(ctx->ev_b));
  float ms = 0.f;
  AHX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  if (elapsed_ms) *elapsed_ms = ms;
  return finish_checked(ctx, "ahx_pop_eval_m");
}

int ahx_pop_scores(ahx_ctx *ctx, double *out) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (ctx->pop_P <= 0 || !out) return fail(ctx, AHX_ERR_STATE, "ahx_pop_scores: no resident population");
  AHX_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_res_scores.p, sizeof(double) * ctx->pop_P, cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AHX_OK;
}

int ahx_eval_m_kernel(const ahx_ctx *ctx, int32_t P, int32_t L, int32_t *tours_per_batch) {
  if (!ctx || !ctx->loaded) return AHX_ERR_STATE;
  int kind = 0, T = 1;
  if (const ResPlan pl = resident_plan(ctx, P); pl.T) { kind = pl.stream ? 3 : 2; T = pl.T; }
  else if (use_x32(ctx, L)) { kind = 1; T = pick_T32(ctx, P, L); }
  else if (eval_m_smem_bytes(1, ctx->N, L) + 1024 <= ctx->smem_optin) T = pick_T(ctx, P, L);
  if (tours_per_batch) *tours_per_batch = T;
  return kind;
}

int64_t ahx_pair_table_entries(const ahx_ctx *ctx, int32_t tours_per_batch) {
  if (!ctx || !ctx->loaded) return AHX_ERR_STATE;
  if (tours_per_batch != 1 && tours_per_batch != 2 && tours_per_batch != 4) return AHX_ERR_ARG;
  return ctx->E_res[res_ti(tours_per_batch)];
}

int ahx_flush_l2(ahx_ctx *ctx) {
  int rc = enter(ctx, false);
  if (rc != AHX_OK) return rc;
  const size_t bytes = std::max<size_t>(ctx->l2_bytes * 2, 256u << 20);
  AHX_CUDA(ctx, ctx->d_flush.reserve(bytes));
  fill_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(reinterpret_cast<int4 *>(ctx->d_flush.p), bytes / 16, 1);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AHX_OK;
}

// ---- probes -------------------------------------------------------------------------
int ahx_probe_fp64(ahx_ctx *ctx, double *dfma_tflops, double *ddiv_gops) {
  int rc = enter(ctx, false);
  if (rc != AHX_OK) return rc;
  const int grid = ctx->sm_count * 8, block = 256, iters = 1 << 14;
  DevBuf<double> buf;
  AHX_CUDA(ctx, buf.reserve(static_cast<size_t>(grid) * block));
  float ms = 0.f;
  for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    probe_dfma_kernel<<<grid, block, 0, ctx->stream>>>(buf.p, iters);
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    AHX_CUDA(ctx, cudaEventSynchronize(ctx->ev_b));
    AHX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  }
  if (dfma_tflops) *dfma_tflops = 2.0 * 8.0 * iters * static_cast<double>(grid) * block / (ms * 1e-3) / 1e12;
  const int diters = 1 << 11;
  for (int pass = 0; pass < 2; ++pass) {
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    probe_ddiv_kernel<<<grid, block, 0, ctx->stream>>>(buf.p, diters);
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    AHX_CUDA(ctx, cudaEventSynchronize(ctx->ev_b));
    AHX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  }
  if (ddiv_gops) *ddiv_gops = 4.0 * diters * static_cast<double>(grid) * block / (ms * 1e-3) / 1e9;
  ctx->launches += 4;
  buf.release();
  return AHX_OK;
}

int ahx_probe_l2(ahx_ctx *ctx, uint64_t bytes, double *gbs) {
  int rc = enter(ctx, false);
  if (rc != AHX_OK) return rc;
  if (bytes < 4096) return fail(ctx, AHX_ERR_ARG, "ahx_probe_l2: buffer too small");
  DevBuf<char> buf;
  AHX_CUDA(ctx, buf.reserve(bytes + 16));
  const size_t n = bytes / 16;
  fill_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(reinterpret_cast<int4 *>(buf.p), n, 3);
  const int reps = static_cast<int>(std::max<uint64_t>(4, (8ull << 30) / bytes));
  float ms = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    probe_read_kernel<<<ctx->sm_count * 8, 512, 0, ctx->stream>>>(reinterpret_cast<const int4 *>(buf.p), n, reps, reinterpret_cast<int4 *>(buf.p) + n);
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    AHX_CUDA(ctx, cudaEventSynchronize(ctx->ev_b));
    AHX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  }
  if (gbs) *gbs = static_cast<double>(n) * 16.0 * reps / (ms * 1e-3) / 1e9;
  ctx->launches += 3;
  buf.release();
  return AHX_OK;
}

int64_t ahx_launch_count(const ahx_ctx *ctx) { return ctx ? ctx->launches : 0; }
double ahx_last_device_ms(const ahx_ctx *ctx) { return ctx ? static_cast<double>(ctx->last_ms) : 0.0; }
int ahx_sm_count(const ahx_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
