// ahx_device.cuh -- device-side building blocks shared by the libahx kernels.
//
// Data layout in HBM for one linkage group:
//   sizes   int64[N]                 contig lengths (CLM.Tigs[i].Size)
//   coo     uint2[E] (N<=65535: x = a | b<<16, y = nlinks) or int4[E] (a,b,nlinks,0)
//           the nonzero upper triangle of M (clm.go:437-447), sorted by (a,b)
//   M       int32[N*N]               dense M, only built for the dense kernel
//   slots   int32[E*8], hist int32[H*12]   oriented GArrays (orientation.go:137-153)
// Tours are int32[P*L] (or uint16 on the u16 entry point), one row per individual.
#ifndef AHX_DEVICE_CUH_
#define AHX_DEVICE_CUH_

#include <cstdint>
#include <cuda_runtime.h>

namespace ahx {

constexpr int kBlock = 256;              // threads per CTA for the evaluate kernels
constexpr int kWarps = kBlock / 32;
constexpr double kLimit = 10000000.0;    // evaluate.go:22 LIMIT
constexpr long long kLimitI = 10000000;  // orientation.go:182 (integer compare)

static __constant__ long long c_GR[12] = {5778, 9349, 15127, 24476, 39603, 64079,
                                   103682, 167761, 271443, 439204, 710647, 1149851};  // base.go:101-103

// ---- block-wide helpers (fixed shape => bit-reproducible) -------------------
__device__ __forceinline__ long long warp_incl_scan(long long v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    long long n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}

// Exclusive prefix sum over the kBlock per-thread values; `scratch` holds kWarps
// long longs.  All threads must call.
__device__ __forceinline__ long long block_excl_scan(long long v, long long *scratch, int tid) {
  const int lane = tid & 31, wid = tid >> 5;
  const long long incl = warp_incl_scan(v, lane);
  if (lane == 31) scratch[wid] = incl;
  __syncthreads();
  long long base = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) base += (w < wid) ? scratch[w] : 0;
  __syncthreads();
  return base + incl - v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;  // valid in lane 0
}

// Sum over the block in a fixed tree; result valid in thread 0.  `scratch`: kWarps doubles.
__device__ __forceinline__ double block_sum(double v, double *scratch, int tid) {
  const int lane = tid & 31, wid = tid >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double s = 0.0;
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += scratch[w];
  }
  __syncthreads();
  return s;
}

// ---- COO accessors ----------------------------------------------------------
struct Coo16 {
  using Elem = uint2;
  __device__ static __forceinline__ void get(const Elem *p, int e, int &a, int &b, double &n) {
    const uint2 v = __ldg(p + e);
    a = static_cast<int>(v.x & 0xffffu); b = static_cast<int>(v.x >> 16); n = static_cast<double>(v.y);
  }
  __device__ static __forceinline__ void unpack(const Elem v, int &a, int &b, uint32_t &n) {
    a = static_cast<int>(v.x & 0xffffu); b = static_cast<int>(v.x >> 16); n = v.y;
  }
};
struct Coo32 {
  using Elem = int4;
  __device__ static __forceinline__ void get(const Elem *p, int e, int &a, int &b, double &n) {
    const int4 v = __ldg(p + e);
    a = v.x; b = v.y; n = static_cast<double>(v.z);
  }
  __device__ static __forceinline__ void unpack(const Elem v, int &a, int &b, uint32_t &n) {
    a = v.x; b = v.y; n = static_cast<uint32_t>(v.z);
  }
};

// ---- Tour.Evaluate building blocks (evaluate.go:116-143) ---------------------
// Phase 1: mid-points of one tour scattered to contig space.
//   mid_i = cumSum_i + size_i/2 (evaluate.go:120-126); every value is an exact
//   half-integer below 2^53, so (2*cum + size) * 0.5 in integers is bit-identical.
// tig: the tour in shared memory (L ints); x: contig-space mid-points, element
// (c, t) at x[c*T + t].  Inactive contigs keep NaN so every pair touching them
// fails the `d <= LIMIT` test.
template <int T>
__device__ __forceinline__ void scatter_mids(const int *tig, int L, const long long *__restrict__ sizes,
                                             double *x, int t, long long *scan_scratch, int tid) {
  const int per = (L + kBlock - 1) / kBlock;
  const int beg = min(L, tid * per), end = min(L, beg + per);
  long long local = 0;
  for (int i = beg; i < end; ++i) local += __ldg(sizes + tig[i]);
  long long cum = block_excl_scan(local, scan_scratch, tid);
  for (int i = beg; i < end; ++i) {
    const int c = tig[i];
    const long long s = __ldg(sizes + c);
    x[static_cast<size_t>(c) * T + t] = static_cast<double>(2 * cum + s) * 0.5;
    cum += s;
  }
}

// Phase 2: stream the nonzero pairs of M once for the T tours held by this CTA.
//   score -= float64(nlinks) / dist   for dist <= LIMIT   (evaluate.go:135-140)
// The break at the first j beyond LIMIT is a mask because mid is monotone in j.
template <int T, typename Coo>
__device__ __forceinline__ void accumulate_pairs(const double *x, const typename Coo::Elem *__restrict__ coo, int E,
                                                 double (&acc)[T], int tid) {
#pragma unroll 2
  for (int e = tid; e < E; e += kBlock) {
    int a, b; double n;
    Coo::get(coo, e, a, b, n);
    const double *xa = x + static_cast<size_t>(a) * T, *xb = x + static_cast<size_t>(b) * T;
    if constexpr (T % 2 == 0) {
#pragma unroll
      for (int t = 0; t < T; t += 2) {
        const double2 va = *reinterpret_cast<const double2 *>(xa + t);
        const double2 vb = *reinterpret_cast<const double2 *>(xb + t);
        const double d0 = fabs(va.x - vb.x), d1 = fabs(va.y - vb.y);
        if (d0 <= kLimit) acc[t] += n / d0;
        if (d1 <= kLimit) acc[t + 1] += n / d1;
      }
    } else {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const double d = fabs(xa[t] - xb[t]);
        if (d <= kLimit) acc[t] += n / d;
      }
    }
  }
}

// ---- Tour.Evaluate, 32-bit coordinate form -----------------------------------
// X_c = 2*mid_c = 2*cumSum + size is an exact integer; when 2*sum(sizes) +
// 2*LIMIT + 2 < 2^32 it fits a uint32 and
//     dist = |X_a - X_b| / 2,   dist <= LIMIT  <=>  0 < |X_a - X_b| <= 2*LIMIT,
//     nlinks / dist = (2*nlinks) / |X_a - X_b|     (scaling by 2 is exact in fp64)
// so the window test runs on the integer pipe and only in-window pairs touch the
// fp64 pipe.  Inactive contigs hold 0xFFFFFFFF: against an active contig the
// difference exceeds 2*LIMIT, against each other it is 0, both fail the test
// (d - 1 < 2*LIMIT, unsigned).  Element (c, t) of the T tours of a CTA lives at
// X[c*T + t], so one 4*T-byte shared-memory load fetches contig c for all T tours.
constexpr uint32_t kLimit2 = 20000000u;      // 2 * LIMIT
constexpr uint32_t kInactiveX = 0xFFFFFFFFu;

template <int T>
__device__ __forceinline__ void scatter_x32(const int *tig, int L, const long long *__restrict__ sizes,
                                            uint32_t *X, int t, long long *scan_scratch, int tid) {
  const int per = (L + kBlock - 1) / kBlock;
  const int beg = min(L, tid * per), end = min(L, beg + per);
  long long local = 0;
  for (int i = beg; i < end; ++i) local += __ldg(sizes + tig[i]);
  long long cum = block_excl_scan(local, scan_scratch, tid);
  for (int i = beg; i < end; ++i) {
    const int c = tig[i];
    const long long s = __ldg(sizes + c);
    X[static_cast<size_t>(c) * T + t] = static_cast<uint32_t>(2 * cum + s);
    cum += s;
  }
}

template <int T> struct XVec;
template <> struct XVec<1> { using type = uint32_t; };
template <> struct XVec<2> { using type = uint2; };
template <> struct XVec<4> { using type = uint4; };

__device__ __forceinline__ void x_unpack(uint32_t v, uint32_t (&o)[1]) { o[0] = v; }
__device__ __forceinline__ void x_unpack(uint2 v, uint32_t (&o)[2]) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void x_unpack(uint4 v, uint32_t (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

// One stored pair (a, b, nlinks) against the T tours of this CTA.
template <int T>
__device__ __forceinline__ void pair_terms_x32(const uint32_t *X, int a, int b, uint32_t nlinks, double (&acc)[T]) {
  using V = typename XVec<T>::type;
  uint32_t xa[T], xb[T];
  x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(a) * T), xa);
  x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(b) * T), xb);
  const double n2 = 2.0 * static_cast<double>(nlinks);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const uint32_t d = xa[t] > xb[t] ? xa[t] - xb[t] : xb[t] - xa[t];
    if (d - 1u < kLimit2) acc[t] += n2 / static_cast<double>(d);   // evaluate.go:135-140
  }
}

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX) -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "AHX_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra AHX_DONE;\n"
      "bra AHX_WAIT;\n"
      "AHX_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// The pair table streamed through shared memory: kStages tiles of kTile entries,
// filled by TMA bulk copies (one elected thread issues, an mbarrier per stage
// signals arrival), consumed by all threads.  E_pad is a multiple of kBlock.
constexpr int kTile = 1024;
constexpr int kStages = 2;

template <int T, typename Coo>
__device__ __forceinline__ void stream_pairs_x32(const uint32_t *X, const typename Coo::Elem *__restrict__ coo, int E_pad,
                                                 typename Coo::Elem *tiles, uint64_t *bars, double (&acc)[T], int tid) {
  using Elem = typename Coo::Elem;
  const int ntiles = (E_pad + kTile - 1) / kTile;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      if (s < ntiles) {
        const int n = min(kTile, E_pad - s * kTile);
        mbar_expect_tx(&bars[s], n * static_cast<uint32_t>(sizeof(Elem)));
        bulk_copy_g2s(tiles + static_cast<size_t>(s) * kTile, coo + static_cast<size_t>(s) * kTile, n * static_cast<uint32_t>(sizeof(Elem)), &bars[s]);
      }
    }
  }
  for (int k = 0; k < ntiles; ++k) {
    const int s = k % kStages;
    mbar_wait(&bars[s], (k / kStages) & 1);
    const Elem *tile = tiles + static_cast<size_t>(s) * kTile;
    const int n = min(kTile, E_pad - k * kTile);
#pragma unroll
    for (int i = 0; i < kTile / kBlock; ++i) {
      const int e = tid + i * kBlock;
      if (e < n) {
        int a, b; uint32_t nl;
        Coo::unpack(tile[e], a, b, nl);
        pair_terms_x32<T>(X, a, b, nl, acc);
      }
    }
    __syncthreads();   // every thread is done with stage s before it is refilled
    if (tid == 0 && k + kStages < ntiles) {
      const int kn = k + kStages;
      const int nn = min(kTile, E_pad - kn * kTile);
      mbar_expect_tx(&bars[s], nn * static_cast<uint32_t>(sizeof(Elem)));
      bulk_copy_g2s(tiles + static_cast<size_t>(s) * kTile, coo + static_cast<size_t>(kn) * kTile, nn * static_cast<uint32_t>(sizeof(Elem)), &bars[s]);
    }
  }
}

// ---- Philox4x32-10 (counter-based RNG, Salmon et al. SC'11) -------------------
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
    const uint64_t p = static_cast<uint64_t>(a) * b;
    hi = static_cast<uint32_t>(p >> 32); lo = static_cast<uint32_t>(p);
  }
  __host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t h0, l0, h1, l1;
      mulhilo(0xD2511F53u, c0, h0, l0);
      mulhilo(0xCD9E8D57u, c2, h1, l1);
      const uint32_t n0 = h1 ^ c1 ^ ka, n1 = l1, n2 = h0 ^ c3 ^ kb, n3 = l0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};
// rng.Intn(n): multiply-shift (bias < n / 2^32)
__host__ __device__ inline int rand_below(uint32_t r, int n) { return static_cast<int>((static_cast<uint64_t>(r) * static_cast<uint32_t>(n)) >> 32); }
// rng.Float64(): 53 random bits in [0,1)
__host__ __device__ inline double rand_unit(uint32_t hi, uint32_t lo) {
  return static_cast<double>(((static_cast<uint64_t>(hi) << 32 | lo) >> 11)) * (1.0 / 9007199254740992.0);
}

// ---- mutation operators as index remaps (evaluate.go:157-218) -----------------
// new[i] = old[mut_source(...)].  op: 0 MutPermute, 1 MutSplice (k = p),
// 2 MutInsertion (dir != 0: pop q, insert at p), 3 MutInversion, <0: identity.
__host__ __device__ inline int mut_source(int op, int p, int q, int dir, int L, int i) {
  switch (op) {
    case 0: return i == p ? q : (i == q ? p : i);                       // evaluate.go:200-208
    case 1: { const int s = i + p; return s >= L ? s - L : s; }        // evaluate.go:212-218 (b ++ a)
    case 2:                                                             // evaluate.go:172-197
      if (p == q || i < p || i > q) return i;
      if (dir) return i == p ? q : i - 1;
      return i == q ? p : i + 1;
    case 3: return (p != q && i >= p && i <= q) ? p + q - i : i;        // evaluate.go:157-169
    default: return i;
  }
}

}  // namespace ahx

#endif  // AHX_DEVICE_CUH_
