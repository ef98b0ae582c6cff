// ahx_resident.cuh -- Tour.Evaluate (evaluate.go:116-143) as a persistent,
// warp-specialised kernel; the building blocks are shared with the persistent GA
// kernel (ahx_ga_res.cuh).
//
// One CTA per SM (kRThreads = 768 threads, up to ~215 KB of shared memory):
//   table   uint32[E_res]   nonzero pairs of M as the byte offsets of the two contigs in X,
//                           (a*4T) | (b*4T) << 16 (N*4T < 64 KB whenever X fits), in the
//                           (nearly) bank-conflict-free order of order_pairs() for this T,
//                           rounded up with (0, 1, nlinks 0) entries to a multiple of
//                           kRCT * kRUnroll and stored so that a thread's 4 entries of an iteration are one
//                           LDS.128; loaded ONCE per CTA by TMA bulk copies -- or, STREAM
//                           mode, streamed for every batch through a ring of kRStages tiles
//   nl8     uint8[E_res]    nlinks of the entry (a pair with more than 255 links is split
//                           over several entries: sum_i n_i / d == n / d up to rounding)
//   X[2]    uint32[N*T]     2*mid of contig c in tour t at X[c*T + t], double buffered
//   sizes   uint32[N + 1]   contig sizes, staged here when they fit (else read from global); entry N = 0
// Roles:
//   8 producer warps  build X[buf] for batch k+1 (T tours: vector loads of the rows,
//                     two-pass prefix sum of the sizes, scatter 2*cum+size to contig
//                     space) while the consumers work on batch k; they also fold the 16
//                     per-warp partial sums of a finished batch in a fixed order and write
//                     the scores.  full[buf] / done[buf] mbarriers hand the buffers over.
//                     (STREAM: the last of them is the table loader.)
//   16 consumer warps each owns 1/16 of the pair table and evaluates EVERY stored pair
//                     for the T tours of the batch, branch-free: the window test only
//                     changes the exponent of the converted distance (res_pair_terms),
//                     the reciprocal is MUFU.RCP64H + one cubic correction (<= 1 ulp).
//                     The cost is data independent -- random and converged tours run at
//                     the same rate -- and bound by instruction issue (an fp64 instruction
//                     takes 2 issue cycles; profiles/microbench/term_probe.cu); filtering
//                     in-window pairs first (the r1a/r1b design: ballot compaction into
//                     per-warp queues, re-gather, divide) cost more than the divides it saved.
// Batches are assigned round-robin (batch b -> CTA b mod gridDim.x); every tour has its
// own accumulators and every reduction a fixed shape, so a score depends neither on the
// tours that share its batch nor on scheduling: bit-reproducible.
#ifndef AHX_RESIDENT_CUH_
#define AHX_RESIDENT_CUH_

#include <type_traits>

#include "ahx_device.cuh"

namespace ahx {

constexpr int kRC = 16;                       // consumer warps
constexpr int kRP = 8;                        // producer warps
constexpr int kRCT = kRC * 32;                // consumer threads
constexpr int kRPT = kRP * 32;                // producer threads
constexpr int kRThreads = kRCT + kRPT;
constexpr int kRUnroll = 4;                   // pairs per consumer lane per iteration (one uint4 of the table)
constexpr int kRPad = kRCT * kRUnroll;        // E_res is a multiple of this
// STREAM mode (the pair table does not fit next to the coordinates): the table is streamed
// from L2 for every batch through a ring of kRStages tiles of kRPad entries (4 B offsets
// + 1 B link counts each) filled by TMA bulk copies; the last producer warp becomes the
// loader, the other kRP - 1 build coordinates.
constexpr double kResFlush = 1e-30;          // sums below this consist of masked-out terms only: written as +0
constexpr int kRStages = 4;
constexpr int kRTileBytes = 5 * kRPad;
template <bool STREAM> struct ResRoles { static constexpr int NB = STREAM ? kRP - 1 : kRP; };   // warps that build X

struct ResArgs {
  const uint32_t *pairs;      // [E_res]
  const uint8_t *nl8;         // [E_res]
  const uint32_t *sizes32;    // [N + 1], entry N = 0
  int E_res, N, L;
  int *errflag;
  int sizes_in_smem;          // the contig sizes are staged in shared memory (4N bytes) instead of gathered from global memory
  int scaled;                 // STREAM: `pairs` holds byte offsets (a*4T) | (b*4T) << 16 like the resident table, not contig indices
};

__host__ __device__ inline size_t res_align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }
// dynamic shared memory of the resident kernel
inline size_t res_smem_bytes(int T, int N, long long E_res, bool stream = false, bool sizes = false) {
  return (sizes ? res_align16(4ull * (N + 1)) : 0) + (stream ? static_cast<size_t>(kRStages) * kRTileBytes : res_align16(4ull * E_res) + res_align16(static_cast<size_t>(E_res))) +
         2 * res_align16(4ull * N * T) + (stream ? 8ull * 2 * kRStages : 0) +
         8ull * 2 * kRC * T + res_align16(4ull * (kRC + kRP) * T) + 8 * 8;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// cp.async.bulk.prefetch.L2: bytes rounded down to a multiple of 16 from a 16-byte aligned start
__device__ __forceinline__ void l2_prefetch(const void *p, uint32_t bytes) {
  const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 15) & ~static_cast<uintptr_t>(15);
  const uintptr_t e = (reinterpret_cast<uintptr_t>(p) + bytes) & ~static_cast<uintptr_t>(15);
  if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(static_cast<uint32_t>(e - a)) : "memory");
}

template <int T, int NW, int BAR, typename Src>
__device__ __forceinline__ void res_build_x(const Src &src, int batch, int L, int N, const uint32_t *sizes32, uint32_t *X,
                                            uint32_t *scan_scratch, int *errflag, int ptid);

// ---- batch sources -------------------------------------------------------------
// A source turns a batch into coordinates and takes the scores back:
//   build<NW, BAR>(batch, ..., X, ...)  called by NW warps: fill X[c*T + t] for the T tours of the batch
//   prefetch(batch, t)                   optional L2 prefetch of a later batch
//   store(batch, t, score)               called once per slot with the finished score
// ResDirect reads tours from a P x L array and prefix-sums the sizes (res_build_x, through
// cursor(batch, t): live(), get4(i, end, c) = contigs at positions i..i+3); the GA's source
// (ahx_ga_res.cuh) transforms the parent's stored coordinates instead.
template <int T, typename IdxT>
struct ResDirect {
  const IdxT *tours; const int *rows; int P, L;
  double *out;
  struct Cursor {
    const IdxT *ptr;
    __device__ __forceinline__ bool live() const { return ptr != nullptr; }
    __device__ __forceinline__ unsigned get(int i) const { return static_cast<unsigned>(ptr[i]); }
    // positions i .. i+3 (i % 4 == 0), the ones at or beyond `end` as 0xFFFFFFFF; one vector load when the row allows it
    __device__ __forceinline__ void get4(int i, int end, unsigned (&c)[4]) const {
      if (i + 3 < end && (reinterpret_cast<uintptr_t>(ptr) & (4 * sizeof(IdxT) - 1)) == 0) {
        if constexpr (sizeof(IdxT) == 4) {
          const uint4 v = *reinterpret_cast<const uint4 *>(ptr + i);
          c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
        } else {
          const uint2 v = *reinterpret_cast<const uint2 *>(ptr + i);
          c[0] = v.x & 0xffffu; c[1] = v.x >> 16; c[2] = v.y & 0xffffu; c[3] = v.y >> 16;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) c[e] = i + e < end ? static_cast<unsigned>(ptr[i + e]) : 0xFFFFFFFFu;
      }
    }
  };
  __device__ __forceinline__ int n_batches() const { return (P + T - 1) / T; }
  static constexpr bool kRaggedBatches = false;   // at most the last batch of a call is ragged: not worth a specialised loop
  __device__ __forceinline__ int live(int batch) const { return min(T, P - batch * T); }   // live tours: the first slots of the batch
  __device__ __forceinline__ int out_row(int batch, int t) const {
    const int p = batch * T + t;
    return p < P ? (rows ? rows[p] : p) : -1;
  }
  __device__ __forceinline__ Cursor cursor(int batch, int t) const {
    const int r = out_row(batch, t);
    return Cursor{r >= 0 ? tours + static_cast<size_t>(r) * L : nullptr};
  }
  __device__ __forceinline__ void store(int batch, int t, double score) const {
    const int r = out_row(batch, t);
    if (r >= 0) out[r] = score;
  }
  template <int NW, int BAR>
  __device__ __forceinline__ void build(int batch, int L, int N, const uint32_t *sizes32, uint32_t *X, uint32_t *scratch,
                                        int *errflag, int gtid) const {
    res_build_x<T, NW, BAR>(*this, batch, L, N, sizes32, X, scratch, errflag, gtid);
  }
  // asks the L2 for the rows of a later batch (called by T lanes, one row each)
  __device__ __forceinline__ void prefetch(int batch, int t) const {
    const int r = batch < n_batches() ? out_row(batch, t) : -1;
    if (r >= 0) l2_prefetch(tours + static_cast<size_t>(r) * L, static_cast<uint32_t>(sizeof(IdxT)) * L);
  }
};

// ---- producer: X[buf] for one batch ----------------------------------------------
// Lane l of a producer warp works on tour t = l mod T, on 4 consecutive positions at
// offset 4 * (l / T): a warp instruction covers 4 * 32/T consecutive positions of each
// of the T tours (vector loads, sector-dense), and the scatter X[c*T + t] spreads over
// all 32 banks.  Producer warp w owns positions [w*Q, (w+1)*Q).  Pass 1 sums the sizes of
// the warp's range per tour, a named barrier publishes the kRP x T totals, pass 2 re-reads
// the range (L2/L1 hits), prefix-sums 4 positions in the lane and the lane totals with a
// segmented warp scan per tour, and scatters 2*cum + size = 2*mid (evaluate.go:120-126).
// 2*sum(sizes) < 2^32 in x32 mode, so all prefix sums run in 32 bits.
// NW warps build one batch together (the kRP producer warps in steady state, all kRC + kRP
// warps for the first batch of a call, when nothing else can run yet); BAR: named barrier.
// size of contig c, c in [0, N]: entry N of the sizes array is 0 (the "skip" contig of positions
// beyond the end and of invalid entries).  SM: the array is the shared-memory copy (32-bit address).
template <bool SM>
__device__ __forceinline__ uint32_t res_size_at(const uint32_t *sizes32, uint32_t sbase, uint32_t c) {
  if constexpr (SM) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + 4u * c));
    return v;
  } else {
    return __ldg(sizes32 + c);
  }
}

template <int T, int NW, int BAR, bool SM, typename Src>
__device__ __forceinline__ void res_build_x_impl(const Src &src, int batch, int L, int N, const uint32_t *sizes32, uint32_t *X,
                                                 uint32_t *scan_scratch /* NW*T */, int *errflag, int ptid) {
  constexpr int PPI = 4 * (32 / T);              // positions of one tour per warp instruction group
  constexpr int kU = 4;                          // groups in flight per lane (16 positions)
  const int lane = ptid & 31, wid = ptid >> 5;
  const int t = lane % T, sub = 4 * (lane / T);
  const auto cur = src.cursor(batch, t);
  const int Q = ((L + NW - 1) / NW + PPI - 1) / PPI * PPI;
  const int beg = min(L, wid * Q), end = cur.live() ? min(L, beg + Q) : beg;   // dead slot: every position is "beyond the end"
  const int wend = min(L, beg + Q);                                            // warp-uniform loop bound
  // iterations whose 16 positions per lane are all inside the range, for every lane of the warp, take
  // the short path (no per-position range tests); the last, ragged one (and dead slots) the general one
  const int fend = __all_sync(0xffffffffu, cur.live()) ? wend : beg;
  const uint32_t un = static_cast<uint32_t>(N);
  const uint32_t sbase = SM ? static_cast<uint32_t>(__cvta_generic_to_shared(sizes32)) : 0u;
  bool bad = false;
  if (L < N) {   // sub-tour: contigs outside the tour must fail every window test
    for (int i = ptid; i < N * T; i += NW * 32) X[i] = kInactiveX;
  }
  // contigs of the lane's 16 positions from p0 on, sanitised to [0, N]: N = skip (beyond the end, or an
  // invalid entry, which also raises `bad`)
  auto fetch = [&](int p0, unsigned (&c)[kU][4], auto ragged) {
#pragma unroll
    for (int u = 0; u < kU; ++u) cur.get4(p0 + u * PPI, end, c[u]);
#pragma unroll
    for (int u = 0; u < kU; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if constexpr (decltype(ragged)::value) bad |= (p0 + u * PPI + e < end) && c[u][e] >= un;
        else bad |= c[u][e] >= un;
        c[u][e] = min(c[u][e], un);
      }
  };
  // pass 1: sum of sizes over this warp's range, per tour
  uint32_t local = 0;
  auto pass1 = [&](int p0, auto ragged) {
    unsigned c[kU][4];
    fetch(p0, c, ragged);
#pragma unroll
    for (int u = 0; u < kU; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) local += res_size_at<SM>(sizes32, sbase, c[u][e]);
  };
  {
    int p0 = beg + sub;
    for (; p0 - sub + PPI * kU <= fend; p0 += PPI * kU) pass1(p0, std::false_type{});
    for (; p0 - sub < wend; p0 += PPI * kU) pass1(p0, std::true_type{});
  }
#pragma unroll
  for (int d = T; d < 32; d <<= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if (lane < T) scan_scratch[wid * T + lane] = local;
  named_bar_sync(BAR, NW * 32);
  uint32_t carry = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) carry += (w < wid) ? scan_scratch[w * T + t] : 0u;
  // pass 2: prefix sums + scatter
  auto pass2 = [&](int p0, auto ragged) {
    unsigned c[kU][4]; uint32_t sz[kU][4];
    fetch(p0, c, ragged);
#pragma unroll
    for (int u = 0; u < kU; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e) sz[u][e] = res_size_at<SM>(sizes32, sbase, c[u][e]);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t tot = sz[u][0] + sz[u][1] + sz[u][2] + sz[u][3];
      uint32_t v = tot;
#pragma unroll
      for (int d = T; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += n; }
      uint32_t cum = carry + v - tot;                                           // sizes before this lane's 4 positions
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (c[u][e] != un) X[c[u][e] * T + t] = 2u * cum + sz[u][e];            // 2*cum + size = 2*mid (evaluate.go:120-126)
        cum += sz[u][e];
      }
      carry += __shfl_sync(0xffffffffu, v, 32 - T + t);
    }
  };
  {
    int p0 = beg + sub;
    for (; p0 - sub + PPI * kU <= fend; p0 += PPI * kU) pass2(p0, std::false_type{});
    for (; p0 - sub < wend; p0 += PPI * kU) pass2(p0, std::true_type{});
  }
  if (bad) *errflag = 1;
  named_bar_sync(BAR, NW * 32);   // scan_scratch may be rewritten by the next batch; X[buf] is complete
}

template <int T, int NW, int BAR, typename Src>
__device__ __forceinline__ void res_build_x(const Src &src, int batch, int L, int N, const uint32_t *sizes32, uint32_t *X,
                                            uint32_t *scan_scratch /* NW*T */, int *errflag, int ptid) {
  if (__isShared(sizes32)) res_build_x_impl<T, NW, BAR, true>(src, batch, L, N, sizes32, X, scan_scratch, errflag, ptid);
  else res_build_x_impl<T, NW, BAR, false>(src, batch, L, N, sizes32, X, scan_scratch, errflag, ptid);
}

// ---- consumer: every stored pair x T tours, branch-free ----------------------------
// 1/x for a positive normal double: MUFU.RCP64H seed (relative error 2^-20 on
// [1, 2^25], measured) and one cubic step r(1 + e + e^2), e = 1 - x r  => 2^-60
// before rounding, i.e. <= 1 ulp of the IEEE quotient (rcp_probe.cu, V1).
__device__ __forceinline__ double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}

// acc[t] += nlinks / dist for one table entry:
//   p  = byte offsets of the two contigs' coordinate groups in X: (a*4T) | (b*4T) << 16
//   n  = nlinks of the entry as a double (res_n); the sum is doubled once at the end
//   d  = |X_a - X_b| = 2 dist;  inside the window  <=>  0 < d <= 2 LIMIT  (evaluate.go:135-137)
//   term = (2 nlinks) / d = 2 (n / d)                                      (evaluate.go:140)
// The window mask rides on the uint32 -> double conversion: the exact conversion is
// (2^52 + d) - 2^52, i.e. a double with high word 0x43300000 and low word d, minus 2^52;
// outside the window the high word is 0x4B300000 instead (one SEL replaces the constant's
// MOV), which makes x ~ 2^180: the term becomes ~1e-52 instead of exactly 0, no inf/NaN can
// arise (d = 0 included), and no operand needs a 64-bit select.  Every genuine term is
// >= 2 / (2 LIMIT) = 1e-7, so the masked terms (< 1e-46 in total) can never change a
// non-empty sum, and a sum below kResFlush is an empty one and is written as exactly +0
// (Go's score starts from +0.0, evaluate.go:128).
// SUB = false (full tours, L == N): all X are distinct and padding entries are
//   (0, 1, n = 0), so d >= 1 and the test is d <= 2 LIMIT.
// SUB = true  (sub-tour): inactive contigs share X = 0xFFFFFFFF, d can be 0: the test is
//   d - 1 < 2 LIMIT (unsigned).
// Issue-slot model measured with profiles/microbench/term_probe.cu: an fp64 instruction
// costs 2 issue cycles, anything else 1.
// Numerator: byte K of a word of four 8-bit link counts as a double -- one I2F.F64.U8 with a
// byte selector (XU pipe, no shared-memory access; the 256-entry 2n table it replaces was 20 %
// of the kernel's shared-memory wavefronts, r1s).  The factor 2 of (2 nlinks) / d is applied
// once per tour to the finished sum (exact).
template <int K>
__device__ __forceinline__ double res_n(uint32_t n8x4) {
  double v;
  asm("{.reg .b16 h; cvt.u16.u32 h, %1; cvt.rn.f64.u8 %0, h;}" : "=d"(v) : "r"(n8x4 >> (8 * K)));
  return v;
}

// SCALED: p holds byte offsets (resident tables); otherwise contig indices a | b << 16 (the
// streamed table of large groups, where N * 4T can exceed 16 bits).
// TL: tours of the batch that are live (the first TL slots; a ragged last batch): the others cost nothing.
template <int T, bool SUB, bool SCALED, int TL = T>
__device__ __forceinline__ void res_pair_terms(const uint32_t *X, uint32_t p, double n, double (&acc)[T]) {
  using V = typename XVec<T>::type;
  uint32_t xa[T], xb[T];
  const unsigned char *Xb = reinterpret_cast<const unsigned char *>(X);
  if constexpr (SCALED) {
    x_unpack(*reinterpret_cast<const V *>(Xb + (p & 0xffffu)), xa);
    x_unpack(*reinterpret_cast<const V *>(Xb + (p >> 16)), xb);
  } else {
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(p & 0xffffu) * T), xa);
    x_unpack(*reinterpret_cast<const V *>(X + static_cast<size_t>(p >> 16) * T), xb);
  }
#pragma unroll
  for (int t = 0; t < TL; ++t) {
    const uint32_t d = __usad(xa[t], xb[t], 0u);
    const bool in = SUB ? (d - 1u < kLimit2) : (d <= kLimit2);
    // exact uint32 -> double is (2^52 + d) - 2^52 with the high word 0x43300000; outside the window the
    // high word is 0x4B300000 instead, so x ~ 2^180 and the term is ~1e-52 (see kResFlush)
    const double x = __hiloint2double(in ? 0x43300000 : 0x4B300000, static_cast<int>(d)) - 4503599627370496.0;
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));                 // MUFU.RCP64H: high word, relative error 2^-20
    r0 = __hiloint2double(__double2hiint(r0), static_cast<int>(d));         // any low word will do (2^-27): saves zeroing it
    double e = fma(-x, r0, 1.0);
    e = fma(e, e, e);
    acc[t] = fma(n, fma(r0, e, r0), acc[t]);                               // r0 (1 + e + e^2): <= 1 ulp of 1/x (rcp_probe.cu V1)
  }
}

// ---- shared-memory layout and the batch loop, shared by the evaluate kernel and the
// ---- persistent GA kernel (ahx_ga_res.cuh) ------------------------------------------
struct ResSmem {
  uint32_t *table; uint8_t *nl8; uint32_t *X0; size_t xstride;
  double *partial; uint64_t *full, *done, *tab, *tfull, *tempty; uint32_t *scan_scratch;
  unsigned char *ring;       // STREAM: kRStages tiles
  const uint32_t *sizes;     // contig sizes: shared-memory copy, or the global array
  unsigned char *extra;      // first byte after the evaluate layout (GA scratch)
};

template <int T, bool STREAM = false>
__device__ __forceinline__ ResSmem res_carve(unsigned char *sp, const ResArgs &g) {
  ResSmem m;
  m.table = nullptr; m.nl8 = nullptr; m.ring = nullptr; m.tfull = m.tempty = nullptr;
  if constexpr (STREAM) {
    m.ring = sp; sp += static_cast<size_t>(kRStages) * kRTileBytes;
  } else {
    m.table = reinterpret_cast<uint32_t *>(sp); sp += res_align16(4ull * g.E_res);
    m.nl8 = sp; sp += res_align16(static_cast<size_t>(g.E_res));
  }
  m.X0 = reinterpret_cast<uint32_t *>(sp);
  m.xstride = res_align16(4ull * g.N * T) / 4;   // words between the two X buffers
  sp += 2 * res_align16(4ull * g.N * T);
  m.partial = reinterpret_cast<double *>(sp); sp += 8ull * 2 * kRC * T;      // [buf][warp][t]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sp); sp += 8 * 8;            // full[2], done[2], tab
  m.full = bars; m.done = bars + 2; m.tab = bars + 4;
  if constexpr (STREAM) { m.tfull = reinterpret_cast<uint64_t *>(sp); m.tempty = m.tfull + kRStages; sp += 8ull * 2 * kRStages; }
  m.scan_scratch = reinterpret_cast<uint32_t *>(sp); sp += res_align16(4ull * (kRC + kRP) * T);
  m.sizes = g.sizes32;
  if (g.sizes_in_smem) { m.sizes = reinterpret_cast<const uint32_t *>(sp); sp += res_align16(4ull * (g.N + 1)); }
  m.extra = sp;
  return m;
}

// Called by all threads once, before the first res_run_batches: barriers, the numerator
// table, and the TMA bulk copies of the pair table (they land while the first batch's
// coordinates are being built).  Ends with a block-wide barrier.
template <bool STREAM = false>
__device__ __forceinline__ void res_init(const ResSmem &m, const ResArgs &g, bool load_table) {
  const int tid = threadIdx.x;
  if (g.sizes_in_smem)
    for (int i = tid; i <= g.N; i += kRThreads) const_cast<uint32_t *>(m.sizes)[i] = __ldg(g.sizes32 + i);   // entry N = 0: the skip contig
  if (tid == 0) {
    mbar_init(&m.full[0], ResRoles<STREAM>::NB * 32); mbar_init(&m.full[1], ResRoles<STREAM>::NB * 32);
    mbar_init(&m.done[0], kRC); mbar_init(&m.done[1], kRC);
    mbar_init(m.tab, 1);
    if constexpr (STREAM)
      for (int i = 0; i < kRStages; ++i) { mbar_init(&m.tfull[i], 1); mbar_init(&m.tempty[i], kRC); }
    mbar_init_fence();
    if (!STREAM && load_table && g.E_res > 0) {
      const uint32_t tb = 4u * static_cast<uint32_t>(g.E_res), nbytes = static_cast<uint32_t>(res_align16(static_cast<size_t>(g.E_res)));
      mbar_expect_tx(m.tab, tb + nbytes);
      for (uint32_t off = 0; off < tb; off += 32768u)
        bulk_copy_g2s(reinterpret_cast<unsigned char *>(m.table) + off, reinterpret_cast<const unsigned char *>(g.pairs) + off, min(32768u, tb - off), m.tab);
      for (uint32_t off = 0; off < nbytes; off += 32768u)
        bulk_copy_g2s(m.nl8 + off, g.nl8 + off, min(32768u, nbytes - off), m.tab);
    }
  }
  __syncthreads();
}

// Scores batches first, first + stride, ... < nb of `src`.  Called by all kRThreads
// threads; `kdone` counts the batches this CTA has pushed through the X buffers since
// res_init (it carries the mbarrier phases from one call to the next) and must be the
// same in every thread.  Returns with all scores stored and both X buffers free.
//   src.cursor(batch, t)          tour in slot t of the batch (producer side)
//   src.prefetch(batch, t)        optional L2 prefetch of a later batch
//   src.store(batch, t, score)    called once per slot with the finished score
template <int T, bool SUB, bool STREAM, typename Src>
__device__ __forceinline__ void res_run_batches(const ResArgs &g, const ResSmem &m, const Src &src, int first, int stride, int nb,
                                                int &kdone, bool &table_ready) {
  constexpr int NB = ResRoles<STREAM>::NB;
  const int tid = threadIdx.x;
  const int k0 = kdone;
  int k = k0;
  const bool loader = STREAM && tid >= kRCT + NB * 32;
  const int ntiles = g.E_res / kRPad;
  // first batch of the call: nobody has anything to score yet, so all consumer and build warps fill its
  // coordinates together (a third of the latency of the build warps alone); both X buffers are free here.
  // Build thread p of the steady state is thread kRCT + p; in the joint build the thread index is just tid.
  if (first < nb && !loader) src.template build<kRC + NB, 2>(first, g.L, g.N, m.sizes, m.X0 + (k & 1) * m.xstride, m.scan_scratch, g.errflag, tid);
  if (loader) {
    // ------------------------------------------------------------ table loader (STREAM)
    if (tid == kRCT + NB * 32) {
      for (int b = first; b < nb; b += stride, ++k)
        for (int tile = 0; tile < ntiles; ++tile) {
          const long long c = static_cast<long long>(k) * ntiles + tile;          // tiles pushed through the ring since res_init
          const int st = static_cast<int>(c % kRStages);
          if (c >= kRStages) mbar_wait(&m.tempty[st], static_cast<uint32_t>((c / kRStages - 1) & 1));
          unsigned char *dst = m.ring + static_cast<size_t>(st) * kRTileBytes;
          mbar_expect_tx(&m.tfull[st], kRTileBytes);
          bulk_copy_g2s(dst, reinterpret_cast<const unsigned char *>(g.pairs) + static_cast<size_t>(tile) * 4 * kRPad, 4 * kRPad, &m.tfull[st]);
          bulk_copy_g2s(dst + 4 * kRPad, g.nl8 + static_cast<size_t>(tile) * kRPad, kRPad, &m.tfull[st]);
        }
    } else {
      for (int b = first; b < nb; b += stride) ++k;
    }
  } else if (tid >= kRCT) {
    // ------------------------------------------------------------ producers
    const int ptid = tid - kRCT;
    auto finalize = [&](int kk, int batch) {
      const int buf = kk & 1;
      mbar_wait_relaxed(&m.done[buf], (kk >> 1) & 1);
      if (ptid < T) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kRC; ++w) s += m.partial[(buf * kRC + w) * T + ptid];
        src.store(batch, ptid, 0.0 - (s < kResFlush ? 0.0 : 2.0 * s));   // score starts at +0.0 and only subtracts (evaluate.go:128,140)
      }
    };
    for (int b = first; b < nb; b += stride, ++k) {
      if (k - k0 >= 2) {
        // X[k & 1] is free once the consumers are done with batch k - 2: build warp 0 polls the mbarrier
        // (and folds that batch's partial sums), the others block on a named barrier -- polling from all
        // build warps took 10 % of the SM's issue slots (r1s)
        if (ptid < 32) finalize(k - 2, b - 2 * stride);
        named_bar_sync(3, NB * 32);
      }
      if (ptid < T) src.prefetch(b + stride, ptid);   // next batch's rows: HBM -> L2 now
      if (k > k0) src.template build<NB, 1>(b, g.L, g.N, m.sizes, m.X0 + (k & 1) * m.xstride, m.scan_scratch, g.errflag, ptid);
      mbar_arrive(&m.full[k & 1]);
    }
    if (ptid < 32)
      for (int kk = max(k0, k - 2); kk < k; ++kk) finalize(kk, first + (kk - k0) * stride);
  } else {
    // ------------------------------------------------------------ consumers
    const int lane = tid & 31, wid = tid >> 5;
    for (int b = first; b < nb; b += stride, ++k) {
      const int buf = k & 1;
      if (!STREAM && !table_ready) { mbar_wait(m.tab, 0); table_ready = true; }
      mbar_wait(&m.full[buf], (k >> 1) & 1);
      const uint32_t *X = m.X0 + buf * m.xstride;
      double acc[T];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = 0.0;
      // entry j of thread tid in iteration I is stored at [(I*kRCT + tid)*4 + j]: one LDS.128 + one LDS.32
      // fetch the thread's four entries and link counts (logical order: order_pairs(), ahx_core.cu)
      // UI table vectors (4 entries each) per thread are in flight at once, so that 16 independent
      // reciprocal chains hide the MUFU/fp64 latency whatever T is
      constexpr int UI = T >= 4 ? 1 : (T == 2 ? 2 : 4);
      if constexpr (!STREAM) {
        const uint4 *table4 = reinterpret_cast<const uint4 *>(m.table);
        const uint32_t *nl8x4 = reinterpret_cast<const uint32_t *>(m.nl8);
        const int nvec = g.E_res / kRUnroll;
        // software-pipelined: the entries of iteration I + 1 are fetched while iteration I computes
        uint4 pn[UI]; uint32_t nn[UI];
        auto fetch = [&](int i) {
#pragma unroll
          for (int u = 0; u < UI; ++u) {
            const int iu = min(i + u * kRCT, nvec - 1);
            pn[u] = table4[iu]; nn[u] = i + u * kRCT < nvec ? nl8x4[iu] : 0u;   // beyond the end: zero link counts
          }
        };
        // a batch with fewer live tours than slots (the last batch of a population; in the GA every CTA's last
        // batch, since the mutants are dealt out evenly) skips the dead tours' arithmetic: T = 4 only
        auto sweep_table = [&](auto live_tours) {
          constexpr int TL = decltype(live_tours)::value;
          fetch(tid);
          for (int i = tid; i < nvec; i += UI * kRCT) {
            uint4 p[UI]; uint32_t n[UI];
#pragma unroll
            for (int u = 0; u < UI; ++u) { p[u] = pn[u]; n[u] = nn[u]; }
            if (i + UI * kRCT < nvec) fetch(i + UI * kRCT);
#pragma unroll
            for (int u = 0; u < UI; ++u) {
              res_pair_terms<T, SUB, true, TL>(X, p[u].x, res_n<0>(n[u]), acc);
              res_pair_terms<T, SUB, true, TL>(X, p[u].y, res_n<1>(n[u]), acc);
              res_pair_terms<T, SUB, true, TL>(X, p[u].z, res_n<2>(n[u]), acc);
              res_pair_terms<T, SUB, true, TL>(X, p[u].w, res_n<3>(n[u]), acc);
            }
          }
        };
        if constexpr (T == 4 && Src::kRaggedBatches) {
          switch (src.live(b)) {
            case 1: sweep_table(std::integral_constant<int, 1>{}); break;
            case 2: sweep_table(std::integral_constant<int, 2>{}); break;
            case 3: sweep_table(std::integral_constant<int, 3>{}); break;
            default: sweep_table(std::integral_constant<int, 4>{}); break;
          }
        } else {
          sweep_table(std::integral_constant<int, T>{});
        }
      } else {
        constexpr int US = UI > 2 ? 2 : UI;   // at most half of the ring's stages in use, the other half loading
        // ring position of this batch's first tile; stage and phase then advance incrementally
        const long long c0 = static_cast<long long>(k) * ntiles;                  // tiles pushed through the ring since res_init
        int st = static_cast<int>(c0 % kRStages);
        uint32_t ph = static_cast<uint32_t>((c0 / kRStages) & 1);
        auto run = [&](auto scaled, auto count) {
          constexpr bool SC = decltype(scaled)::value;
          constexpr int NT = decltype(count)::value;                              // tiles of this step (US, or 1 for the odd last one)
          uint4 p[NT]; uint32_t n[NT]; int used[NT];
#pragma unroll
          for (int u = 0; u < NT; ++u) {
            mbar_wait(&m.tfull[st], ph);
            const unsigned char *src_tile = m.ring + static_cast<size_t>(st) * kRTileBytes;
            p[u] = reinterpret_cast<const uint4 *>(src_tile)[tid];
            n[u] = reinterpret_cast<const uint32_t *>(src_tile + 4 * kRPad)[tid];
            used[u] = st;
            if (++st == kRStages) { st = 0; ph ^= 1u; }
          }
#pragma unroll
          for (int u = 0; u < NT; ++u) {
            res_pair_terms<T, SUB, SC>(X, p[u].x, res_n<0>(n[u]), acc);
            res_pair_terms<T, SUB, SC>(X, p[u].y, res_n<1>(n[u]), acc);
            res_pair_terms<T, SUB, SC>(X, p[u].z, res_n<2>(n[u]), acc);
            res_pair_terms<T, SUB, SC>(X, p[u].w, res_n<3>(n[u]), acc);
          }
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int u = 0; u < NT; ++u) mbar_arrive(&m.tempty[used[u]]);         // done with the stage
          }
        };
        auto sweep = [&](auto scaled) {
          int tile = 0;
          for (; tile + US <= ntiles; tile += US) run(scaled, std::integral_constant<int, US>{});
          for (; tile < ntiles; ++tile) run(scaled, std::integral_constant<int, 1>{});
        };
        // entries are byte offsets into X when N * 4T fits 16 bits (one instruction less per gather), else contig indices
        if (g.scaled) sweep(std::true_type{}); else sweep(std::false_type{});
      }
      // warp totals of the T accumulators by a transposing butterfly: at the first log2(T) levels a lane
      // keeps half of its tours and hands the other half to its partner, so T tours cost
      // T - 1 + 5 - log2(T) exchanges instead of 5 T (fixed shape: deterministic)
      if constexpr (T == 4) {
        const bool b4 = lane & 16, b3 = lane & 8;
        double k0 = b4 ? acc[2] : acc[0], k1 = b4 ? acc[3] : acc[1];
        k0 += __shfl_xor_sync(0xffffffffu, b4 ? acc[0] : acc[2], 16);
        k1 += __shfl_xor_sync(0xffffffffu, b4 ? acc[1] : acc[3], 16);
        double s = b3 ? k1 : k0;
        s += __shfl_xor_sync(0xffffffffu, b3 ? k0 : k1, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if ((lane & 7) == 0) m.partial[(buf * kRC + wid) * T + (lane >> 3)] = s;      // lanes 0, 8, 16, 24: tours 0..3
      } else if constexpr (T == 2) {
        const bool b4 = lane & 16;
        double s = b4 ? acc[1] : acc[0];
        s += __shfl_xor_sync(0xffffffffu, b4 ? acc[0] : acc[1], 16);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if ((lane & 15) == 0) m.partial[(buf * kRC + wid) * T + (lane >> 4)] = s;     // lanes 0, 16: tours 0, 1
      } else {
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const double s = warp_sum(acc[t]);
          if (lane == 0) m.partial[(buf * kRC + wid) * T + t] = s;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&m.done[buf]);
    }
  }
  kdone = k;
}

template <int T, bool SUB, bool STREAM, typename Src>
__global__ void __launch_bounds__(kRThreads, 1) eval_m_res_kernel(ResArgs g, Src src) {
  extern __shared__ __align__(128) unsigned char smem_res[];
  const ResSmem m = res_carve<T, STREAM>(smem_res, g);
  const int nb = src.n_batches();
  res_init<STREAM>(m, g, static_cast<int>(blockIdx.x) < nb);
  int kdone = 0;
  bool table_ready = g.E_res == 0;
  res_run_batches<T, SUB, STREAM>(g, m, src, blockIdx.x, gridDim.x, nb, kdone, table_ready);
}

}  // namespace ahx

#endif  // AHX_RESIDENT_CUH_
