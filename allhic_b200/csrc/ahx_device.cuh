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

template <int T> struct XVec;
template <> struct XVec<1> { using type = uint32_t; };
template <> struct XVec<2> { using type = uint2; };
template <> struct XVec<4> { using type = uint4; };

__device__ __forceinline__ void x_unpack(uint32_t v, uint32_t (&o)[1]) { o[0] = v; }
__device__ __forceinline__ void x_unpack(uint2 v, uint32_t (&o)[2]) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void x_unpack(uint4 v, uint32_t (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

// Exclusive block scan of T values per thread with one pair of barriers.
// scratch: kWarps * T long longs.
template <int T>
__device__ __forceinline__ void block_excl_scan_vec(long long (&v)[T], long long *scratch, int tid) {
  const int lane = tid & 31, wid = tid >> 5;
  long long incl[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    incl[t] = warp_incl_scan(v[t], lane);
    if (lane == 31) scratch[wid * T + t] = incl[t];
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < T; ++t) {
    long long base = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) base += (w < wid) ? scratch[w * T + t] : 0;
    v[t] = base + incl[t] - v[t];
  }
  __syncthreads();
}

// Prologue of the x32 kernels: coordinates of up to T tours scattered to contig
// space in one pass.  tour_ptr[t] = row of tour t (nullptr: slot unused);
// S = contig sizes staged in shared memory (uint32: x32 mode implies sizes < 2^31).
// Thread `tid` owns the contiguous chunk [tid*per, tid*per+per) of every tour and
// keeps it in registers between the two passes (kChunk elements per round), so the
// global loads of a round are independent and issued back to back.
constexpr int kChunk = 8;

// Source of tour elements for the prologue: get(t, i) = contig at position i of
// tour slot t (only called when live(t)); put(t, i, c) lets the GA write the
// offspring row while it is being read.
template <int T, typename IdxT>
struct DirectTours {
  const IdxT *ptr[T];
  __device__ __forceinline__ bool live(int t) const { return ptr[t] != nullptr; }
  __device__ __forceinline__ unsigned get(int t, int i) const { return static_cast<unsigned>(ptr[t][i]); }
  __device__ __forceinline__ void put(int, int, unsigned) const {}
};

template <int T, typename Src>
__device__ __forceinline__ void scatter_tours_x32(const Src &src, int L, int N, const uint32_t *S, uint32_t *X,
                                                  long long *scan_scratch, int *errflag, int tid) {
  const int per = (L + kBlock - 1) / kBlock;
  const int beg = min(L, tid * per), end = min(L, beg + per);
  bool bad = false;
  if (per <= kChunk) {
    unsigned c[T][kChunk];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int i = 0; i < kChunk; ++i)
        c[t][i] = (src.live(t) && beg + i < end) ? src.get(t, beg + i) : 0xFFFFFFFFu;
    long long cum[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      long long local = 0;
#pragma unroll
      for (int i = 0; i < kChunk; ++i) {
        const bool live = src.live(t) && beg + i < end;
        if (live) src.put(t, beg + i, c[t][i]);
        if (live && c[t][i] >= static_cast<unsigned>(N)) { bad = true; c[t][i] = 0xFFFFFFFFu; }
        if (c[t][i] != 0xFFFFFFFFu) local += S[c[t][i]];
      }
      cum[t] = local;
    }
    block_excl_scan_vec<T>(cum, scan_scratch, tid);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      uint32_t c2 = static_cast<uint32_t>(2 * cum[t]);
#pragma unroll
      for (int i = 0; i < kChunk; ++i) {
        if (c[t][i] == 0xFFFFFFFFu) continue;
        const uint32_t sz = S[c[t][i]];
        X[static_cast<size_t>(c[t][i]) * T + t] = c2 + sz;   // 2*cum + size
        c2 += 2u * sz;
      }
    }
  } else {
    long long cum[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      long long local = 0;
      if (src.live(t)) {
        for (int i = beg; i < end; ++i) {
          const unsigned c = src.get(t, i);
          src.put(t, i, c);
          if (c >= static_cast<unsigned>(N)) bad = true;
          else local += S[c];
        }
      }
      cum[t] = local;
    }
    block_excl_scan_vec<T>(cum, scan_scratch, tid);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (!src.live(t)) continue;
      uint32_t c2 = static_cast<uint32_t>(2 * cum[t]);
      for (int i = beg; i < end; ++i) {
        const unsigned c = src.get(t, i);
        if (c >= static_cast<unsigned>(N)) continue;
        const uint32_t sz = S[c];
        X[static_cast<size_t>(c) * T + t] = c2 + sz;
        c2 += 2u * sz;
      }
    }
  }
  if (bad) *errflag = 1;   // reported by the host wrapper
}

// Two-stage evaluation of the pair stream.  A random tour keeps only ~LIMIT/G
// of the stored pairs inside the window, and a divide issued for 3 live lanes
// costs as much as one issued for 32, so the fp64 work is separated from the
// window test:
//   filter  (integer pipe) every lane tests its pair against the T tours of the
//           CTA; pairs that are inside the window of at least one tour ("hot")
//           are compacted into a per-warp queue in shared memory (ballot/popc);
//   drain   (fp64 pipe) whenever the queue holds >= 32 pairs the warp pops 32
//           and every lane evaluates its pair for all T tours, branch-free, as T
//           independent divide chains.
// Which lane sums which term depends on the data of the T tours sharing the CTA
// but never on scheduling: the result is a deterministic function of (tours,
// batch shape), bit-reproducible run to run; the last bits of one tour's score
// may differ between batch shapes (different CTA mates), far inside 1e-12.
// Exact uint32 -> double without the (quarter-rate) I2F conversion: 2^52 + v is
// exactly representable, subtracting 2^52 leaves v.
__device__ __forceinline__ double u32_to_double(uint32_t v) {
  return __hiloint2double(0x43300000, static_cast<int>(v)) - 4503599627370496.0;
}
// n / d for 0 <= n < 2^34 and 1 <= d <= 2^32 (normal, finite operands): MUFU.RCP64H
// seed (~23 bits), two Newton steps, quotient, one Markstein correction.  The
// result is the correctly rounded quotient except for rare half-ulp ties of the
// uncorrected estimate, i.e. <= 1 ulp from IEEE division -- per-term error
// ~1e-16, four orders below the 1e-12 parity tolerance -- at a third of the
// instructions of the compiler's guarded division sequence.
__device__ __forceinline__ double div_pos(double n, double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  const double q = n * r;
  return fma(fma(-d, q, n), r, q);
}

constexpr int kQueue = 64;

template <int T, typename Coo>
struct PairSink {
  using Elem = typename Coo::Elem;
  using V = typename XVec<T>::type;
  Elem *q;             // this warp's queue
  const uint32_t *X;
  int cnt;             // warp-uniform fill level
  int lane;
  double acc[T];
  __device__ __forceinline__ void init(Elem *warp_q, const uint32_t *X_, int lane_) {
    q = warp_q; X = X_; lane = lane_; cnt = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0;
  }
  __device__ __forceinline__ void evaluate(const Elem c) {
    int a, b; uint32_t nl;
    Coo::unpack(c, a, b, nl);
    uint32_t xa[T], xb[T];
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(a) * T), xa);
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(b) * T), xb);
    const double n2 = 2.0 * u32_to_double(nl);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const uint32_t d = __usad(xa[t], xb[t], 0u);
      const bool in = d - 1u < kLimit2;                          // 0 < d <= 2*LIMIT
      const double term = div_pos(n2, u32_to_double(in ? d : 1u));   // nlinks / dist, evaluate.go:140
      acc[t] += in ? term : 0.0;
    }
  }
  // window test of one pair against the T tours (integer pipe, no warp-level ops)
  __device__ __forceinline__ bool test(const Elem c) const {
    int a, b; uint32_t nl;
    Coo::unpack(c, a, b, nl);
    uint32_t xa[T], xb[T];
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(a) * T), xa);
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(b) * T), xb);
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int t = 0; t < T; ++t) m = min(m, __usad(xa[t], xb[t], 0u) - 1u);
    return m < kLimit2;
  }
  // called by all 32 lanes of the warp, each with its own pair
  __device__ __forceinline__ void enqueue(const Elem c, bool hot) {
    const unsigned bal = __ballot_sync(0xffffffffu, hot);
    if (bal) {
      if (hot) q[cnt + __popc(bal & ((1u << lane) - 1u))] = c;
      cnt += __popc(bal);
      if (cnt >= 32) {
        __syncwarp();
        cnt -= 32;
        evaluate(q[cnt + lane]);
        __syncwarp();
      }
    }
  }
  __device__ __forceinline__ void flush() {
    __syncwarp();
    if (lane < cnt) evaluate(q[lane]);
    cnt = 0;
  }
};

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
// Same wait for a thread with slack (the resident kernel's producers run ahead of the
// consumers): a failed probe sleeps instead of re-issuing, so the spin does not take
// issue slots from the warps doing the arithmetic.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 2000;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(256);
  }
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The pair table streamed through shared memory as a kStages-deep ring of
// kTile-entry tiles: full[s] is completed by the TMA bulk copy, empty[s] by one
// arrival per warp once the warp is done with the tile.  Lane 0 of warp 0 is the
// producer; no block-wide barrier inside the loop.  E_pad is a multiple of
// kBlock so every warp runs the same trip counts.
constexpr int kTile = 512;
constexpr int kStages = 4;

template <typename Elem>
__device__ __forceinline__ void issue_tile(const Elem *__restrict__ coo, int E_pad, int k, Elem *tiles, uint64_t *full) {
  const int s = k % kStages;
  const uint32_t bytes = static_cast<uint32_t>(min(kTile, E_pad - k * kTile)) * static_cast<uint32_t>(sizeof(Elem));
  mbar_expect_tx(&full[s], bytes);
  bulk_copy_g2s(tiles + static_cast<size_t>(s) * kTile, coo + static_cast<size_t>(k) * kTile, bytes, &full[s]);
}

// Called by thread 0 right after the barriers are initialised: the first tiles
// stream in while the CTA is still scattering its tours.
template <typename Elem>
__device__ __forceinline__ void prefetch_pairs(const Elem *__restrict__ coo, int E_pad, Elem *tiles, uint64_t *bars) {
  const int ntiles = (E_pad + kTile - 1) / kTile;
  for (int k = 0; k < kStages && k < ntiles; ++k) issue_tile<Elem>(coo, E_pad, k, tiles, bars);
}

// bars: full[kStages] then empty[kStages]; initialised by the caller (full: 1 arrival, empty: kWarps).
template <int T, typename Coo>
__device__ __forceinline__ void stream_pairs_x32(const typename Coo::Elem *__restrict__ coo, int E_pad,
                                                 typename Coo::Elem *tiles, uint64_t *bars, PairSink<T, Coo> &sink, int tid) {
  using Elem = typename Coo::Elem;
  uint64_t *full = bars, *empty = bars + kStages;
  const int ntiles = (E_pad + kTile - 1) / kTile;
  for (int k = 0; k < ntiles; ++k) {
    const int s = k % kStages;
    const uint32_t parity = (k / kStages) & 1;
    mbar_wait(&full[s], parity);
    const Elem *tile = tiles + static_cast<size_t>(s) * kTile;
    const int n = min(kTile, E_pad - k * kTile);
    // loads and window tests of the whole tile slice first (independent, pipelined),
    // then the warp-synchronous compaction
    constexpr int U = kTile / kBlock;
    Elem c[U]; bool hot[U];
#pragma unroll
    for (int i = 0; i < U; ++i) c[i] = tile[min(tid + i * kBlock, kTile - 1)];
#pragma unroll
    for (int i = 0; i < U; ++i) hot[i] = (tid + i * kBlock < n) && sink.test(c[i]);
#pragma unroll
    for (int i = 0; i < U; ++i)
      if (i * kBlock < n) sink.enqueue(c[i], hot[i]);
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    if (tid == 0 && k + kStages < ntiles) {
      mbar_wait(&empty[s], parity);          // all warps are done with stage s
      issue_tile<Elem>(coo, E_pad, k + kStages, tiles, full);
    }
  }
  sink.flush();
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
