// ahx_resident.cuh -- Tour.Evaluate (evaluate.go:116-143) as a persistent,
// warp-specialised kernel with the pair table RESIDENT in shared memory.
//
// One CTA per SM (kRThreads = 640 threads, up to ~200 KB of shared memory):
//   table   uint32[E_res]   nonzero pairs of M packed a | b << 16, in the
//                           bank-conflict-free order of order_pairs(), padded
//                           with null pairs (a == b) to a multiple of
//                           kRCT * kRUnroll; loaded ONCE per CTA by TMA bulk copies
//   nl8     uint8[E_res]    min(nlinks, 255); 255 = look nlinks up in global memory
//   X[2]    uint32[N*T]     2*mid of contig c in tour t at X[c*T + t], double buffered
//   queues  uint32[16][160] per-consumer-warp queue of hot pair indices
// Roles:
//   4 producer warps  build X[buf] for batch k+1 (T tours: read the rows, prefix-sum
//                     the sizes, scatter 2*cum+size to contig space) while the
//                     consumers work on batch k; they also fold the 16 per-warp
//                     partial sums of a finished batch in a fixed order and write
//                     the scores.  full[buf] / done[buf] mbarriers hand the buffers over.
//   16 consumer warps each owns 1/16 of the pair table: window test on the integer
//                     pipe for the T tours of the batch, hot pairs (inside the window
//                     of at least one tour) compacted into the warp's queue, 32 at a
//                     time evaluated branch-free on the fp64 pipe (PairSink logic).
// Batches are assigned round-robin (batch b -> CTA b mod gridDim.x) and every
// reduction has a fixed shape, so a score depends only on the tours that share
// its batch, never on scheduling: bit-reproducible run to run.
#ifndef AHX_RESIDENT_CUH_
#define AHX_RESIDENT_CUH_

#include "ahx_device.cuh"

namespace ahx {

constexpr int kRC = 16;                       // consumer warps
constexpr int kRP = 4;                        // producer warps
constexpr int kRCT = kRC * 32;                // consumer threads
constexpr int kRPT = kRP * 32;                // producer threads
constexpr int kRThreads = kRCT + kRPT;
constexpr int kRUnroll = 4;                   // pairs per consumer lane per iteration
constexpr int kRQueue = 32 + 32 * kRUnroll;   // queue entries per consumer warp
constexpr int kRPad = kRCT * kRUnroll;        // E_res is a multiple of this
constexpr int kRChunk = 16;                   // tour positions a producer thread keeps in registers

struct ResArgs {
  const uint32_t *pairs;      // [E_res]
  const uint8_t *nl8;         // [E_res]
  const int32_t *nl32;        // [E_res]
  const uint32_t *sizes32;    // [N]
  int E_res, N, L;
  double *out;
  int *errflag;
};

__host__ __device__ inline size_t res_align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }
// dynamic shared memory of the resident kernel
inline size_t res_smem_bytes(int T, int N, long long E_res) {
  return res_align16(4ull * E_res) + res_align16(static_cast<size_t>(E_res)) + 2 * res_align16(4ull * N * T) +
         4ull * kRC * kRQueue + 8ull * 2 * kRC * T + 4ull * kRP * T + 8 * 8;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- batch sources -------------------------------------------------------------
// A source maps (batch, slot t) to a tour: n_batches(), bind(batch), live(t),
// get(t, i), put(t, i, c) and out_row(batch, t) (-1: nothing to write).
template <int T, typename IdxT>
struct ResDirect {
  const IdxT *tours; const int *rows; int P, L;
  const IdxT *ptr[T];
  __device__ __forceinline__ int n_batches() const { return (P + T - 1) / T; }
  __device__ __forceinline__ int out_row(int batch, int t) const {
    const int p = batch * T + t;
    return p < P ? (rows ? rows[p] : p) : -1;
  }
  __device__ __forceinline__ void bind(int batch) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int r = out_row(batch, t);
      ptr[t] = r >= 0 ? tours + static_cast<size_t>(r) * L : nullptr;
    }
  }
  __device__ __forceinline__ bool live(int t) const { return ptr[t] != nullptr; }
  __device__ __forceinline__ unsigned get(int t, int i) const { return static_cast<unsigned>(ptr[t][i]); }
  __device__ __forceinline__ void put(int, int, unsigned) const {}
};

// ---- producer: X[buf] for one batch ----------------------------------------------
// Thread ptid owns positions [ptid*per, ptid*per + per) of every tour of the batch.
// 2*sum(sizes) < 2^32 in x32 mode, so the prefix sums run in 32 bits.
template <int T, typename Src>
__device__ __forceinline__ void res_build_x(const Src &src, int L, int N, const uint32_t *__restrict__ sizes32, uint32_t *X,
                                            uint32_t *scan_scratch /* kRP*T */, int *errflag, int ptid) {
  const int lane = ptid & 31, wid = ptid >> 5;
  const int per = (L + kRPT - 1) / kRPT;
  const int beg = min(L, ptid * per), end = min(L, beg + per);
  bool bad = false;
  if (L < N) {   // sub-tour: contigs outside the tour must fail every window test
    for (int i = ptid; i < N * T; i += kRPT) X[i] = kInactiveX;
    named_bar_sync(1, kRPT);
  }
  uint32_t cum[T];
  if (per <= kRChunk) {
    unsigned c[T][kRChunk];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int i = 0; i < kRChunk; ++i) c[t][i] = (src.live(t) && beg + i < end) ? src.get(t, beg + i) : 0xFFFFFFFFu;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      uint32_t local = 0;
#pragma unroll
      for (int i = 0; i < kRChunk; ++i) {
        const bool on = src.live(t) && beg + i < end;
        if (on) src.put(t, beg + i, c[t][i]);
        if (on && c[t][i] >= static_cast<unsigned>(N)) { bad = true; c[t][i] = 0xFFFFFFFFu; }
        local += c[t][i] != 0xFFFFFFFFu ? __ldg(sizes32 + c[t][i]) : 0u;
      }
      cum[t] = local;
    }
    // exclusive scan over the kRPT producer threads
    uint32_t incl[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      uint32_t v = cum[t];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += n; }
      incl[t] = v;
      if (lane == 31) scan_scratch[wid * T + t] = v;
    }
    named_bar_sync(1, kRPT);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      uint32_t base = 0;
#pragma unroll
      for (int w = 0; w < kRP; ++w) base += (w < wid) ? scan_scratch[w * T + t] : 0u;
      uint32_t c2 = 2u * (base + incl[t] - cum[t]);
#pragma unroll
      for (int i = 0; i < kRChunk; ++i) {
        if (c[t][i] == 0xFFFFFFFFu) continue;
        const uint32_t sz = __ldg(sizes32 + c[t][i]);               // L1 hit: read a moment ago
        X[static_cast<size_t>(c[t][i]) * T + t] = c2 + sz;          // 2*cum + size = 2*mid, evaluate.go:120-126
        c2 += 2u * sz;
      }
    }
  } else {
    uint32_t incl[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      uint32_t local = 0;
      if (src.live(t))
        for (int i = beg; i < end; ++i) {
          const unsigned c = src.get(t, i);
          src.put(t, i, c);
          if (c >= static_cast<unsigned>(N)) bad = true; else local += __ldg(sizes32 + c);
        }
      cum[t] = local;
      uint32_t v = local;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += n; }
      incl[t] = v;
      if (lane == 31) scan_scratch[wid * T + t] = v;
    }
    named_bar_sync(1, kRPT);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (!src.live(t)) continue;
      uint32_t base = 0;
#pragma unroll
      for (int w = 0; w < kRP; ++w) base += (w < wid) ? scan_scratch[w * T + t] : 0u;
      uint32_t c2 = 2u * (base + incl[t] - cum[t]);
      for (int i = beg; i < end; ++i) {
        const unsigned c = src.get(t, i);
        if (c >= static_cast<unsigned>(N)) continue;
        const uint32_t s = __ldg(sizes32 + c);
        X[static_cast<size_t>(c) * T + t] = c2 + s;
        c2 += 2u * s;
      }
    }
  }
  if (bad) *errflag = 1;
  named_bar_sync(1, kRPT);   // scan_scratch may be rewritten by the next batch
}

// ---- consumer: window test + hot-pair queue + fp64 drain ------------------------
template <int T>
struct ResSink {
  using V = typename XVec<T>::type;
  const uint32_t *table; const uint8_t *nl8; const int32_t *nl32;
  uint32_t *q;
  const uint32_t *X;
  int cnt, lane;
  double acc[T];
  __device__ __forceinline__ bool test(uint32_t p) const {
    uint32_t xa[T], xb[T];
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(p & 0xffffu) * T), xa);
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(p >> 16) * T), xb);
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int t = 0; t < T; ++t) m = min(m, __usad(xa[t], xb[t], 0u) - 1u);
    return m < kLimit2;                                   // 0 < |dX| <= 2*LIMIT for some tour
  }
  __device__ __forceinline__ void evaluate(uint32_t e) {
    const uint32_t p = table[e];
    uint32_t nl = nl8[e];
    if (nl == 255u) nl = static_cast<uint32_t>(__ldg(nl32 + e));
    uint32_t xa[T], xb[T];
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(p & 0xffffu) * T), xa);
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(p >> 16) * T), xb);
    const double n2 = 2.0 * u32_to_double(nl);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const uint32_t d = __usad(xa[t], xb[t], 0u);
      const bool in = d - 1u < kLimit2;
      const double term = div_pos(n2, u32_to_double(in ? d : 1u));   // nlinks / dist, evaluate.go:140
      acc[t] += in ? term : 0.0;
    }
  }
};

template <int T, typename Src>
__global__ void __launch_bounds__(kRThreads, 1) eval_m_res_kernel(ResArgs g, Src src) {
  extern __shared__ __align__(128) unsigned char smem_res[];
  unsigned char *sp = smem_res;
  uint32_t *table = reinterpret_cast<uint32_t *>(sp); sp += res_align16(4ull * g.E_res);
  uint8_t *nl8 = sp; sp += res_align16(static_cast<size_t>(g.E_res));
  uint32_t *X0 = reinterpret_cast<uint32_t *>(sp);
  const size_t xstride = res_align16(4ull * g.N * T) / 4;   // words between the two X buffers
  sp += 2 * res_align16(4ull * g.N * T);
  uint32_t *queues = reinterpret_cast<uint32_t *>(sp); sp += 4ull * kRC * kRQueue;
  double *partial = reinterpret_cast<double *>(sp); sp += 8ull * 2 * kRC * T;      // [buf][warp][t]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sp); sp += 8 * 8;                  // full[2], done[2], tab
  uint32_t *scan_scratch = reinterpret_cast<uint32_t *>(sp);
  uint64_t *full = bars, *done = bars + 2, *tab = bars + 4;
  const int tid = threadIdx.x;
  const int nb = src.n_batches();
  if (tid == 0) {
    mbar_init(&full[0], kRPT); mbar_init(&full[1], kRPT);
    mbar_init(&done[0], kRC); mbar_init(&done[1], kRC);
    mbar_init(tab, 1);
    mbar_init_fence();
    if (static_cast<int>(blockIdx.x) < nb && g.E_res > 0) {
      // the pair table streams in while the producers build the first batch
      const uint32_t tb = 4u * static_cast<uint32_t>(g.E_res), nbytes = static_cast<uint32_t>(res_align16(static_cast<size_t>(g.E_res)));
      mbar_expect_tx(tab, tb + nbytes);
      for (uint32_t off = 0; off < tb; off += 32768u)
        bulk_copy_g2s(reinterpret_cast<unsigned char *>(table) + off, reinterpret_cast<const unsigned char *>(g.pairs) + off, min(32768u, tb - off), tab);
      for (uint32_t off = 0; off < nbytes; off += 32768u)
        bulk_copy_g2s(nl8 + off, g.nl8 + off, min(32768u, nbytes - off), tab);
    }
  }
  __syncthreads();
  if (tid >= kRCT) {
    // ------------------------------------------------------------ producers
    const int ptid = tid - kRCT;
    auto finalize = [&](int k, int batch) {
      const int buf = k & 1;
      mbar_wait(&done[buf], (k >> 1) & 1);
      if (ptid < T) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kRC; ++w) s += partial[(buf * kRC + w) * T + ptid];
        const int row = src.out_row(batch, ptid);
        if (row >= 0) g.out[row] = 0.0 - s;   // score starts at +0.0 and only subtracts (evaluate.go:128,140)
      }
    };
    int k = 0;
    for (int b = blockIdx.x; b < nb; b += gridDim.x, ++k) {
      if (k >= 2) finalize(k - 2, b - 2 * static_cast<int>(gridDim.x));
      src.bind(b);
      res_build_x<T>(src, g.L, g.N, g.sizes32, X0 + (k & 1) * xstride, scan_scratch, g.errflag, ptid);
      mbar_arrive(&full[k & 1]);
    }
    for (int kk = max(0, k - 2); kk < k; ++kk) finalize(kk, static_cast<int>(blockIdx.x) + kk * static_cast<int>(gridDim.x));
  } else {
    // ------------------------------------------------------------ consumers
    const int lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    ResSink<T> sink;
    sink.table = table; sink.nl8 = nl8; sink.nl32 = g.nl32; sink.q = queues + wid * kRQueue; sink.lane = lane;
    bool table_ready = g.E_res == 0;
    int k = 0;
    for (int b = blockIdx.x; b < nb; b += gridDim.x, ++k) {
      const int buf = k & 1;
      if (!table_ready) { mbar_wait(tab, 0); table_ready = true; }
      mbar_wait(&full[buf], (k >> 1) & 1);
      sink.X = X0 + buf * xstride;
      sink.cnt = 0;
#pragma unroll
      for (int t = 0; t < T; ++t) sink.acc[t] = 0.0;
      for (int i = tid; i < g.E_res; i += kRPad) {
        uint32_t p[kRUnroll]; bool hot[kRUnroll]; unsigned bal[kRUnroll];
#pragma unroll
        for (int j = 0; j < kRUnroll; ++j) p[j] = table[i + j * kRCT];
#pragma unroll
        for (int j = 0; j < kRUnroll; ++j) hot[j] = sink.test(p[j]);
#pragma unroll
        for (int j = 0; j < kRUnroll; ++j) bal[j] = __ballot_sync(0xffffffffu, hot[j]);
        int base = sink.cnt;
#pragma unroll
        for (int j = 0; j < kRUnroll; ++j) {
          if (hot[j]) sink.q[base + __popc(bal[j] & lt)] = static_cast<uint32_t>(i + j * kRCT);
          base += __popc(bal[j]);
        }
        sink.cnt = base;
        while (sink.cnt >= 32) {
          __syncwarp();
          sink.cnt -= 32;
          sink.evaluate(sink.q[sink.cnt + lane]);
          __syncwarp();
        }
      }
      __syncwarp();
      if (lane < sink.cnt) sink.evaluate(sink.q[lane]);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const double s = warp_sum(sink.acc[t]);
        if (lane == 0) partial[(buf * kRC + wid) * T + t] = s;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[buf]);
    }
  }
}

}  // namespace ahx

#endif  // AHX_RESIDENT_CUH_
