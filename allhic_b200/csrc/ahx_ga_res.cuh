// ahx_ga_res.cuh -- CLM.GARun (evaluate.go:258-315) as ONE persistent cooperative
// kernel: many generations per launch, the pair table resident in shared memory for
// the whole run, one grid barrier per generation.  Included by ahx_ga.cu after the
// Philox / tournament / mutation-plan helpers.
//
// Population storage.  eaopt clones every selected individual and mutates a fraction
// MutRate of the clones.  A clone that is not mutated is an exact copy
// of its parent, so individuals are stored by reference: rowof[g & 1][i] is the
// physical row of individual i of generation g in a pool of 2P rows.  An un-mutated
// offspring inherits its parent's row id and fitness (no copy); a mutated offspring
// gets a fresh row from the rows generation g does not reference (at least P of the 2P;
// bit maps in global memory, set with atomicOr while a generation is built).
// The k-th mutated offspring (ascending index) takes the k-th free row, so the layout
// is a pure function of the seed.
//
// One generation g -> g + 1, every CTA (G = gridDim.x, 24 warps).  State that is ready in shared memory when it
// starts: the mutation mask of g (+ word prefix), the bit map of the pool rows g references (+ prefix of the free ones).
//   1. front half (started right behind the previous grid barrier, before that generation's bookkeeping):
//      for the mutated offspring this CTA scores -- the M mutants are dealt out evenly, CTA c gets a contiguous
//      run of them -- offspring index (select on the mask), mutation plan, tournament parent, the operator's two
//      maps (remap_of: needs the parent's tour and coordinate rows); in other threads this CTA's share of the
//      un-mutated offspring (o = c, c + G, ...): tournament, inherit row and fitness
//   2. bookkeeping of generation g (evaluate.go:287-306) from its 128-bit minimum; the last CTA copies the
//      winner's row to the hall of fame, nobody waits
//   3. the TMA bulk copies issued when the barrier opened have landed: mask of g + 1 and row map of g, one block
//      scan over both; back half: the k-th mutant takes the k-th free row
//   4. housekeeping: this CTA's words of the mask of g + 2 (first Philox draw of each offspring; independent of
//      fitness), clear the row map of g - 1, reset the minimum of g + 2
//   5. res_run_batches: the coordinates of T offspring at a time are built straight from the parent rows through
//      the mutation's maps (the child rows are written on the way), consumers score them (ahx_resident.cuh; a
//      ragged last batch only pays for its live tours); an individual that beats the hall of fame enters the
//      128-bit atomic minimum (fitness key, offspring index, pool row)
//   6. grid barrier; the polling thread starts the bulk copies of step 3 for the next generation
// Islands: at exchange generations bookkeeping + ga_island_exchange come first and the next generation starts
// plainly (no speculation, maps reloaded).
// Scores do not depend on which tours share a batch (each tour has its own accumulators
// and a fixed summation order), tournaments and plans are functions of (seed, generation,
// offspring): the run is bit-identical for any grid size and any T.
#ifndef AHX_GA_RES_CUH_
#define AHX_GA_RES_CUH_

namespace ahx {

struct GaRes {
  int P, L, k, nbmax;
  double mut_prob;
  uint32_t key0, key1;
  long long ngen, max_gens;
  int *pool;            // [2P][L]   tours
  uint32_t *xpool;      // [2P][N]   coordinates 2*mid of every contig in that tour (N == L)
  const uint32_t *sizes32;
  uint32_t total2;      // 2 * sum(sizes of the tour's contigs)
  int *rowof[2];        // [P]
  double *fit[2];       // [P]
  int *hof;             // [L]
  GaState *st;
  unsigned int *gbar;   // grid barrier counter, zero at launch
  uint32_t *used;       // [3][ga_wr_pad(P)] bit maps: bit r of used[g % 3] is set iff generation g references pool row r
  ulonglong2 *gmin;     // [3] best individual evaluated while building generation g (index g % 3) that beats the hall of fame: .y = order-preserving
                        // key of its fitness, .x = offspring index << 32 | pool row; 128-bit minimum, so the lowest index wins among equal fitness values
  uint32_t *gmask;      // [3][ga_wm_pad(P)] mutation masks of generations g (index g % 3), filled two generations ahead
  // islands (include/ahx.h "Islands"); n_islands <= 1: single population
  int n_islands, island, migrate_every, n_elite;
  GaIslands *isl;
};

// words of a mutation mask / of a row map, padded to 16 bytes (they travel by TMA bulk copy)
__host__ __device__ inline int ga_wm_pad(int P) { return ((P + 31) / 32 + 3) & ~3; }
__host__ __device__ inline int ga_wr_pad(int P) { return ((2 * P + 31) / 32 + 3) & ~3; }

// One mutated offspring: where it comes from, where it goes, and its mutation in two
// branch-free forms (the four operators of Tour.Mutate, evaluate.go:157-218, fit both --
// no divergence between the tours of a batch):
//   tour:         new[i] = parent[src(i)],
//                 src = i; if (i - lo <u len) src = a*i + b; if (i == sp) src = sv; if (i == sp2) src = sv2;
//                 if (src >= L) src -= L
//   coordinates:  X = 2*mid (evaluate.go:120-126) of contig c in the parent, uint32 arithmetic mod 2^32,
//                 X' = (X - xlo <u xlen) ? xa*X + xb : X + xb0;  if (c == m1) X' = v1;  if (c == m2) X' = v2
// because every operator moves whole segments: contigs between the two cut points shift by
// a constant (swap, insertion, rotation) or are mirrored (inversion), the one or two moved
// contigs get explicit values.  So a child's coordinates are an elementwise map of its
// parent's stored coordinates: no prefix sum, no size lookups, no scatter.
struct GaSlot {
  int o, parent_row, child_row;
  int lo, len, a, b, sp, sv, sp2, sv2;
  uint32_t xlo, xlen, xa, xb, xb0, v1, v2;
  int m1, m2;
};

// pt / px: the parent's tour row and coordinate row; s32: contig sizes; total2 = 2 * sum(sizes)
__device__ __forceinline__ void remap_of(const MutPlan &mp, int L, const int *pt, const uint32_t *px, const uint32_t *__restrict__ s32,
                                         uint32_t total2, GaSlot *s) {
  s->lo = 0; s->len = 0; s->a = 1; s->b = 0; s->sp = -1; s->sv = 0; s->sp2 = -1; s->sv2 = 0;
  s->xlo = 0; s->xlen = 0; s->xa = 1; s->xb = 0; s->xb0 = 0; s->m1 = -1; s->v1 = 0; s->m2 = -1; s->v2 = 0;
  const int p = mp.p, q = mp.q;
  if (mp.op < 0 || (mp.op != 1 && p == q)) return;                 // randomTwoInts may return p == q: every operator is then a no-op
  const int cp = __ldcg(pt + p), cq = __ldcg(pt + (mp.op == 1 ? p : q));
  const uint32_t Xp = __ldcg(px + cp), Xq = __ldcg(px + cq), zp = __ldg(s32 + cp), zq = __ldg(s32 + cq);
  switch (mp.op) {
    case 0:                                                         // MutPermute: swap positions p < q
      s->sp = p; s->sv = q; s->sp2 = q; s->sv2 = p;
      s->xlo = Xp + 1; s->xlen = Xq - Xp - 1; s->xb = 2u * (zq - zp);            // everything in between shifts by size_q - size_p
      s->m1 = cq; s->v1 = Xp - zp + zq;                                           // q's contig now starts where p's did
      s->m2 = cp; s->v2 = Xq + zq - zp;                                           // p's contig now ends where q's did
      break;
    case 1:                                                         // MutSplice: new = old[k:] ++ old[:k], k = p
      s->lo = 0; s->len = L; s->b = p;
      s->xlo = Xp; s->xlen = 0u - Xp; s->xb = 0u - (Xp - zp); s->xb0 = total2 - (Xp - zp);   // tail moves to the front, head behind it
      break;
    case 2:                                                         // MutInsertion
      if (mp.dir) {                                                 // pop q, insert at p: [p, q) shifts right by size_q
        s->lo = p + 1; s->len = q - p; s->b = -1; s->sp = p; s->sv = q;
        s->xlo = Xp; s->xlen = Xq - Xp; s->xb = 2u * zq; s->m1 = cq; s->v1 = Xp - zp + zq;
      } else {                                                      // pop p, insert at q: (p, q] shifts left by size_p
        s->lo = p; s->len = q - p; s->b = 1; s->sp = q; s->sv = p;
        s->xlo = Xp + 1; s->xlen = Xq - Xp; s->xb = 0u - 2u * zp; s->m1 = cp; s->v1 = Xq + zq - zp;
      }
      break;
    case 3:                                                         // MutInversion: reverse [p, q] -> mirror image of the segment
      s->lo = p; s->len = q - p + 1; s->a = -1; s->b = p + q;
      s->xlo = Xp; s->xlen = Xq - Xp + 1; s->xa = 0xFFFFFFFFu; s->xb = (Xp - zp) + (Xq + zq);
      break;
    default: break;
  }
}


constexpr int kIslMaxMigrants = 64;          // n_elite * (n_islands - 1) at most
// island scratch in shared memory: per-warp (key, index) pairs of the block-wide selection, the last
// winner, the selected individuals (elites, then worst) and the migrants' pool rows
constexpr size_t kIslScratchBytes = 2 * 16 * (kRThreads / 32) + 32 + 4 * (3 * kIslMaxMigrants) + 16 + 8 * kIslMaxMigrants;
// The two elementwise maps of a slot (the operator code the persistent kernel runs; ahx_mutate_remap
// applies exactly these to one tour for the bit-exact operator tests).
// coordinate of contig c in the child, from its coordinate x in the parent:
// SUB: the tour covers a subset of the contigs; the others keep the "inactive" coordinate.
template <bool SUB = false>
__device__ __forceinline__ uint32_t slot_coord(const GaSlot &s, int c, uint32_t x) {
  uint32_t y = (x - s.xlo < s.xlen) ? s.xa * x + s.xb : x + s.xb0;
  y = c == s.m1 ? s.v1 : y;
  y = c == s.m2 ? s.v2 : y;
  if (SUB) y = x == kInactiveX ? x : y;
  return y;
}
// parent position that feeds child position i:
__device__ __forceinline__ int slot_source(const GaSlot &s, int i, int L) {
  int src = static_cast<unsigned>(i - s.lo) < static_cast<unsigned>(s.len) ? s.a * i + s.b : i;
  src = i == s.sp ? s.sv : src;
  src = i == s.sp2 ? s.sv2 : src;
  src -= src >= L ? L : 0;
  return src;
}

// The M mutated offspring of a generation are dealt out evenly: CTA c scores the mutants first .. first + n - 1
// (n = M / G or M / G + 1), T at a time; its last batch is then usually ragged and only pays for its live tours
// (res_run_batches).  Batch-granular dealing (batch b -> CTA b mod G) left a third of the CTAs one whole batch short.
__device__ __forceinline__ void ga_deal(uint32_t M, int cta, int G, uint32_t *first, int *n) {
  const uint32_t qn = M / static_cast<uint32_t>(G), r = M % static_cast<uint32_t>(G), c = static_cast<uint32_t>(cta);
  *first = c * qn + min(c, r);
  *n = static_cast<int>(qn + (c < r ? 1u : 0u));
}

// first slot of the CTA's slot table that thread `tid` handles (then every kRThreads-th)
__device__ __forceinline__ int slot_of_thread(int tid) {
  constexpr int NWARP = kRThreads / 32;
  return (NWARP - 1 - (tid >> 5)) + NWARP * (tid & 31);
}

inline size_t ga_res_extra_bytes(int P, int T, int nbmax) {
  const size_t wm = static_cast<size_t>(ga_wm_pad(P)), wr = static_cast<size_t>(ga_wr_pad(P));
  return res_align16(4 * (4 * wm + 2 * wr)) + 2 * res_align16(sizeof(GaSlot) * static_cast<size_t>(nbmax) * T) + res_align16(4 * (kRThreads / 32 + 1)) +
         res_align16(kIslScratchBytes) + 32 + 64;
}

// order-preserving map double -> uint64 (for atomicMin on fitness values) and back
__device__ __forceinline__ unsigned long long fit_key(double f) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(f));
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double key_fit(unsigned long long k) {
  return __longlong_as_double(static_cast<long long>((k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k));
}

// 128-bit minimum of (hi, lo) at addr (.y = hi, .x = lo): atom.cas.b128 (sm_90+).  The first guess is the reset value.
__device__ __forceinline__ void atomic_min_u128(ulonglong2 *addr, unsigned long long hi, unsigned long long lo) {
  unsigned long long chi = ~0ULL, clo = ~0ULL;
  for (;;) {
    if (hi > chi || (hi == chi && lo >= clo)) return;
    unsigned long long ohi, olo;
    asm volatile("{\n\t.reg .b128 cmp, swp, old;\n\tmov.b128 cmp, {%2, %3};\n\tmov.b128 swp, {%4, %5};\n\tatom.relaxed.gpu.global.cas.b128 old, [%6], cmp, swp;\n\tmov.b128 {%0, %1}, old;\n\t}"
                 : "=l"(olo), "=l"(ohi) : "l"(clo), "l"(chi), "l"(lo), "l"(hi), "l"(addr) : "memory");
    if (ohi == chi && olo == clo) return;
    chi = ohi; clo = olo;
  }
}

template <int T, bool SUB = false>
struct GaSrc {
  const int *pool; int *pool_w; const uint32_t *xpool; uint32_t *xpool_w; const GaSlot *slots; double *fitn; ulonglong2 *gmin; double hof_fit; int nlive;   // nlive: this CTA's mutated offspring (slots 0 .. nlive-1)
  // Lane l works on tour t = l mod T and on contig (then position) base + l / T: 32/T consecutive
  // elements of each tour per warp instruction (one 32-byte sector per tour), and the coordinate
  // store X[c*T + t] of a warp is 32 consecutive words.
  template <int NW, int BAR>
  __device__ __forceinline__ void build(int j, int L, int N, const uint32_t *__restrict__, uint32_t *X, uint32_t *, int *, int gtid) const {
    constexpr int CPI = 32 / T, kU = 8;
    const int lane = gtid & 31, wid = gtid >> 5, t = lane % T, sub = lane / T;
    const GaSlot s = slots[j * T + t];
    const bool live = s.o >= 0;
    const uint32_t *px = xpool + static_cast<size_t>(s.parent_row) * N;
    uint32_t *cx = xpool_w + static_cast<size_t>(s.child_row) * N;
    for (int c0 = wid * CPI + sub; c0 - sub < N; c0 += NW * CPI * kU) {
      uint32_t x[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) { const int c = c0 + u * NW * CPI; x[u] = (live && c < N) ? __ldcg(px + c) : 0u; }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int c = c0 + u * NW * CPI;
        const uint32_t y = slot_coord<SUB>(s, c, x[u]);
        if (c < N) { X[static_cast<size_t>(c) * T + t] = y; if (live) cx[c] = y; }
      }
    }
    // Tour.Clone + Tour.Mutate: the child row is the parent row through the operator's index remap
    const int *pt = pool + static_cast<size_t>(s.parent_row) * L;
    int *ct = pool_w + static_cast<size_t>(s.child_row) * L;
    if (live)
      for (int i0 = wid * CPI + sub; i0 < L; i0 += NW * CPI * kU) {
        int v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int i = i0 + u * NW * CPI;
          v[u] = i < L ? __ldcg(pt + slot_source(s, i, L)) : 0;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) { const int i = i0 + u * NW * CPI; if (i < L) ct[i] = v[u]; }
      }
    named_bar_sync(BAR, NW * 32);   // X[buf] is complete
  }
  __device__ __forceinline__ void prefetch(int, int) const {}
  static constexpr bool kRaggedBatches = true;    // every CTA's last batch of every generation
  __device__ __forceinline__ int live(int j) const { return min(T, max(0, nlive - j * T)); }
  __device__ __forceinline__ void store(int j, int t, double score) const {
    const GaSlot &sl = slots[j * T + t];
    if (sl.o >= 0) {
      fitn[sl.o] = score;
      // only an individual that beats the hall of fame matters (evaluate.go:291); rare, so the 128-bit CAS loop is cold
      if (score < hof_fit) atomic_min_u128(gmin, fit_key(score), (static_cast<unsigned long long>(sl.o) << 32) | static_cast<unsigned int>(sl.child_row));
    }
  }
};

// Exclusive prefix sum of one value per thread over the kRThreads threads; *total gets
// the sum.  scratch: kRThreads / 32 + 1 words.  Ends with a block barrier.
__device__ __forceinline__ uint32_t block_scan_u32(uint32_t v, uint32_t *scratch, uint32_t *total) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += n; }
  if (lane == 31) scratch[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < kRThreads / 32 ? scratch[lane] : 0u, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += n; }
    if (lane < kRThreads / 32) scratch[lane] = wi - w;
    if (lane == 31) scratch[kRThreads / 32] = wi;
  }
  __syncthreads();
  const uint32_t r = scratch[wid] + incl - v;
  *total = scratch[kRThreads / 32];
  __syncthreads();
  return r;
}

// In-place exclusive prefix sum of n words in shared memory (any n: chunks of kRThreads with a
// running carry); returns the total.  Ends with a block barrier.
__device__ __forceinline__ uint32_t block_scan_array(uint32_t *vals, int n, uint32_t *scratch) {
  uint32_t carry = 0;
  for (int base = 0; base < n; base += kRThreads) {
    const int i = base + static_cast<int>(threadIdx.x);
    uint32_t tot = 0;
    const uint32_t e = block_scan_u32(i < n ? vals[i] : 0u, scratch, &tot);
    if (i < n) vals[i] = carry + e;
    carry += tot;
  }
  __syncthreads();
  return carry;
}

// index of the m-th (0-based) set bit of a mask given the exclusive word prefix.  Kept compact on purpose: the
// per-generation set-up code runs once per 40 us from a cold instruction cache, and an unrolled 8-ary search with
// a popcount-based select measured SLOWER than this loop + __fns (44.5 against 42.8 us per generation).
__device__ __forceinline__ int mask_select(const uint32_t *mask, const uint32_t *pref, int nwords, uint32_t m, bool invert, int nbits) {
  int lo = 0, hi = nwords - 1;
  while (lo < hi) {   // last word whose prefix is <= m
    const int mid = (lo + hi + 1) >> 1;
    if (pref[mid] <= m) lo = mid; else hi = mid - 1;
  }
  uint32_t w = invert ? ~mask[lo] : mask[lo];
  if (invert && lo == nwords - 1 && (nbits & 31)) w &= (1u << (nbits & 31)) - 1u;
  return lo * 32 + static_cast<int>(__fns(w, 0, static_cast<int>(m - pref[lo]) + 1));
}

// Arrival is a release reduction (no return value to wait for; the release is cumulative over the CTA's writes
// ordered before it by the block barrier), the wait an acquire poll by one thread.
__device__ __forceinline__ void grid_barrier(unsigned int *gbar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gbar) : "memory");
    unsigned int v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gbar) : "memory"); } while (v < target);
  }
  __syncthreads();
}

// The same barrier; the polling thread also fetches 16 bytes that every thread needs right behind it (the
// generation's 128-bit minimum) into shared memory -- one request per CTA to that one line instead of one per warp.
template <typename F>
__device__ __forceinline__ void grid_barrier_fetch(unsigned int *gbar, unsigned int target, const ulonglong2 *src, ulonglong2 *dst_smem, F &&after) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gbar) : "memory");
    unsigned int v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gbar) : "memory"); } while (v < target);
    if (!after()) *dst_smem = __ldcg(src);                // after(): starts the bulk copies of what the next generation needs (the 16 bytes included), or returns false
  }
  __syncthreads();
}

// Mutation flags of generation `gen` (first Philox draw of every offspring: eaopt's
// rng.Float64() < MutRate) as a bit mask.  They do not depend on fitness, so they are
// produced one generation ahead: CTA c computes the words w = c, c + G, ... of the next
// generation's mask while it builds the current one (ga_publish_mask, before the grid
// barrier) and every CTA reads the whole mask back after the barrier (ga_load_mask).
__device__ __forceinline__ void ga_publish_mask(const GaRes &q, const Philox &ph, long long gen, uint32_t *gmask, int cta, int G) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, WM = (q.P + 31) / 32;
  const uint32_t glo = static_cast<uint32_t>(gen), ghi = static_cast<uint32_t>(gen >> 32);
  for (int w = cta + warp * G; w < WM; w += (kRThreads / 32) * G) {   // one warp per word
    const int o = w * 32 + lane;
    bool mut = false;
    if (o < q.P) {
      uint32_t r[4];
      ph(glo, ghi, static_cast<uint32_t>(o), kStreamMutate, r);
      mut = rand_unit(r[0], r[1]) < q.mut_prob;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, mut);
    if (lane == 0) gmask[w] = word;
  }
}
// returns the number of mutated offspring; mflag / mpref: the mask and its exclusive word prefix
__device__ __forceinline__ uint32_t ga_load_mask(const GaRes &q, const uint32_t *gmask, uint32_t *mflag, uint32_t *mpref, uint32_t *scan_tmp) {
  const int WM = (q.P + 31) / 32;
  for (int w = threadIdx.x; w < WM; w += kRThreads) {
    const uint32_t v = __ldcg(gmask + w);
    mflag[w] = v; mpref[w] = __popc(v);
  }
  __syncthreads();
  return block_scan_array(mpref, WM, scan_tmp);
}

// Row bit map of a generation (from global memory, filled while that generation was built) and the exclusive
// word prefix of the FREE rows, for mask_select.
__device__ __forceinline__ void ga_free_prefix(const GaRes &q, const uint32_t *rused, uint32_t *rpref, uint32_t *scan_tmp) {
  const int tid = threadIdx.x, WR = (2 * q.P + 31) / 32;
  for (int wi = tid; wi < WR; wi += kRThreads) {
    uint32_t w = ~rused[wi];
    const int nvalid = min(32, 2 * q.P - wi * 32);
    if (nvalid < 32) w &= (1u << nvalid) - 1u;
    rpref[wi] = __popc(w);
  }
  __syncthreads();
  block_scan_array(rpref, WR, scan_tmp);   // at least P of the 2P rows are free
  __syncthreads();
}
__device__ __forceinline__ void ga_row_map(const GaRes &q, const uint32_t *used_cur, uint32_t *rused, uint32_t *rpref, uint32_t *scan_tmp) {
  const int tid = threadIdx.x, WR = (2 * q.P + 31) / 32;
  for (int wi = tid; wi < WR; wi += kRThreads) rused[wi] = __ldcg(used_cur + wi);
  __syncthreads();
  ga_free_prefix(q, rused, rpref, scan_tmp);
}

// The mutation mask of the generation after next and the row map of the next generation are both complete in
// global memory once the grid barrier that ends a generation has been passed.  The thread that polled the barrier
// then starts two TMA bulk copies (ga_issue_mask_rows) that land them in shared memory -- no registers, nobody
// waits -- while the dependent loads of the next generation's front half run; ga_finish_mask_rows waits on the
// mbarrier and turns them into the select structures (popcount + ONE block scan over both when WM + WR <= kRThreads).
__device__ __forceinline__ void ga_issue_mask_rows(const GaRes &q, const uint32_t *gmask, const uint32_t *used_cur, const ulonglong2 *gmin, uint32_t *mflag, uint32_t *rused,
                                                   ulonglong2 *gm_smem, uint64_t *bar) {
  const uint32_t mb = 4u * static_cast<uint32_t>(ga_wm_pad(q.P)), rb = 4u * static_cast<uint32_t>(ga_wr_pad(q.P));
  asm volatile("fence.proxy.async;" ::: "memory");       // the other CTAs' (generic proxy) stores, acquired at the barrier, before the async-proxy reads
  mbar_expect_tx(bar, mb + rb + 16u);
  bulk_copy_g2s(gm_smem, gmin, 16u, bar);
  bulk_copy_g2s(mflag, gmask, mb, bar);
  bulk_copy_g2s(rused, used_cur, rb, bar);
}
// One thread polls the mbarrier; the others sleep in the block barrier instead of spinning on it (23 polling warps
// took the issue slots the front half's warps needed).
__device__ __forceinline__ void ga_wait_pending(uint64_t *bar, uint32_t parity) {
  if (threadIdx.x == 0) mbar_wait(bar, parity);
  __syncthreads();
}
// returns the number of mutated offspring of the mask; ends with a block barrier
__device__ __forceinline__ uint32_t ga_finish_mask_rows(const GaRes &q, uint64_t *bar, uint32_t parity, const uint32_t *mflag, uint32_t *mpref, const uint32_t *rused,
                                                        uint32_t *rpref, uint32_t *scan_tmp) {
  const int tid = threadIdx.x, WM = (q.P + 31) / 32, WR = (2 * q.P + 31) / 32, wi = tid - WM;
  (void)bar; (void)parity;                                // the caller has waited for the copies (ga_wait_pending)
  if (WM + WR <= 32) {
    // small populations (the CLI's default 100: 4 + 7 words): one warp scans both maps with shuffles, one block
    // barrier instead of five -- at 10 us per generation every barrier of 24 warps shows
    if (tid < 32) {
      uint32_t cnt = 0;
      if (tid < WM) cnt = __popc(mflag[tid]);
      else if (wi < WR) {
        uint32_t w = ~rused[wi];
        const int nvalid = min(32, 2 * q.P - wi * 32);
        if (nvalid < 32) w &= (1u << nvalid) - 1u;
        cnt = __popc(w);
      }
      uint32_t incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += n; }
      const uint32_t excl = incl - cnt, M = __shfl_sync(0xffffffffu, excl, WM);   // lane WM's exclusive prefix = mutants of the mask
      if (tid < WM) mpref[tid] = excl;
      else if (wi < WR) rpref[wi] = excl - M;
      if (tid == 0) scan_tmp[0] = M;
    }
    __syncthreads();
    return scan_tmp[0];
  }
  if (WM + WR > kRThreads) {
    for (int w = tid; w < WM; w += kRThreads) mpref[w] = __popc(mflag[w]);
    __syncthreads();
    const uint32_t M = block_scan_array(mpref, WM, scan_tmp);
    ga_free_prefix(q, rused, rpref, scan_tmp);
    return M;
  }
  uint32_t cnt = 0;
  if (tid < WM) cnt = __popc(mflag[tid]);
  else if (wi < WR) {
    uint32_t w = ~rused[wi];
    const int nvalid = min(32, 2 * q.P - wi * 32);
    if (nvalid < 32) w &= (1u << nvalid) - 1u;
    cnt = __popc(w);
  }
  uint32_t total = 0;
  const uint32_t excl = block_scan_u32(cnt, scan_tmp, &total);
  if (tid < WM) mpref[tid] = excl;
  if (tid == WM) scan_tmp[0] = excl;                      // = number of mutated offspring (sum of the mask segment)
  __syncthreads();
  const uint32_t M = scan_tmp[0];
  if (tid >= WM && wi < WR) rpref[wi] = excl - M;
  __syncthreads();
  return M;
}

// ---- islands: the exchange step, inside the kernel ------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct IslScratch {
  unsigned long long *wkey;   // [2][warps] (double buffered by selection round: one block barrier per round)
  int *widx;                  // [2][warps]
  unsigned long long *last;   // [1] key of the previous winner (large populations' path)
  int *last_idx;              // [1] (and [2]: time-out flag)
  int *sel;                   // [2 * kIslMaxMigrants] selected individuals
  int *rows;                  // [kIslMaxMigrants] pool rows of the migrants
  double *mfit;               // [kIslMaxMigrants] fitness of the migrants
};
__device__ __forceinline__ IslScratch isl_scratch(unsigned char *p) {
  IslScratch s;
  s.wkey = reinterpret_cast<unsigned long long *>(p); p += 2 * 8 * (kRThreads / 32);
  s.last = reinterpret_cast<unsigned long long *>(p); p += 8;
  s.mfit = reinterpret_cast<double *>(p); p += 8 * kIslMaxMigrants;
  s.widx = reinterpret_cast<int *>(p); p += 2 * 4 * (kRThreads / 32);
  s.last_idx = reinterpret_cast<int *>(p); p += 16;
  s.sel = reinterpret_cast<int *>(p); p += 4 * 2 * kIslMaxMigrants;
  s.rows = reinterpret_cast<int *>(p);
  return s;
}

// The k individuals with the smallest (key, index), ascending, into out[0..k): key = fit_key(fitness)
// (best first, lowest index on ties) or its complement (worst first, lowest index on ties).  Whole block.
// Populations up to kSelReg * kRThreads: every thread reads its share of the fitness array ONCE; a round is a
// warp arg-min by shuffles, one block barrier, and the same arg-min over the warps' results done by every warp
// for itself; the winner's owner retires it.  (r1: k sweeps of the array from L2, two barriers and a serial
// merge per round -- 45 us for the 14 migrants of 8 islands, measured with AHX_GA_TIMING.)
__device__ __forceinline__ void isl_select_k(const double *fit, int P, int k, bool worst, int *out, const IslScratch &sc) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int kSelReg = 12, NW = kRThreads / 32;
  constexpr unsigned long long kNoKey = ~0ULL; constexpr int kNoIdx = 0x7fffffff;
  auto better = [](unsigned long long ka, int ia, unsigned long long kb, int ib) { return ka < kb || (ka == kb && ia < ib); };
  auto warp_argmin = [&](unsigned long long &bk, int &bi) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, d);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
      if (better(ok, oi, bk, bi)) { bk = ok; bi = oi; }
    }
  };
  if (P <= kSelReg * kRThreads) {
    unsigned long long keys[kSelReg];
#pragma unroll
    for (int j = 0; j < kSelReg; ++j) {
      const int i = tid + j * kRThreads;
      unsigned long long key = i < P ? fit_key(__ldcg(fit + i)) : 0ULL;
      if (worst) key = ~key;
      keys[j] = i < P ? key : kNoKey;
    }
    // lane-local and warp-local minima are cached: a round only recomputes them in the thread / warp that owned
    // the previous winner, every other warp just redoes the 24-entry arg-min over the warps' results
    auto local_min = [&](unsigned long long &bk, int &bi) {
      bk = kNoKey; bi = kNoIdx;
#pragma unroll
      for (int j = 0; j < kSelReg; ++j) {
        const int i = keys[j] == kNoKey ? kNoIdx : tid + j * kRThreads;      // retired / beyond the end
        if (better(keys[j], i, bk, bi)) { bk = keys[j]; bi = i; }
      }
    };
    unsigned long long mk; int mi;                                           // this thread's best candidate
    local_min(mk, mi);
    bool dirty = true;                                                       // this warp's entry needs recomputing
    for (int r = 0; r < k; ++r) {
      unsigned long long *wk = sc.wkey + (r & 1) * NW; int *wi = sc.widx + (r & 1) * NW;
      const unsigned long long *wk_prev = sc.wkey + ((r & 1) ^ 1) * NW; const int *wi_prev = sc.widx + ((r & 1) ^ 1) * NW;
      unsigned long long bk = mk; int bi = mi;
      if (dirty) warp_argmin(bk, bi);                                        // warp-uniform condition
      if (lane == 0) {
        if (dirty) { wk[wid] = bk; wi[wid] = bi; }
        else { wk[wid] = wk_prev[wid]; wi[wid] = wi_prev[wid]; }
      }
      __syncthreads();
      bk = lane < NW ? wk[lane] : kNoKey; bi = lane < NW ? wi[lane] : kNoIdx;
      warp_argmin(bk, bi);                                                   // every warp: the block's winner
      if (tid == 0) out[r] = bi;
      dirty = bi != kNoIdx && ((bi % kRThreads) >> 5) == wid;                // the winner lives in this warp
      if (bi != kNoIdx && bi % kRThreads == tid) {                           // its owner retires it
#pragma unroll
        for (int j = 0; j < kSelReg; ++j)
          if (tid + j * kRThreads == bi) keys[j] = kNoKey;
        local_min(mk, mi);
      }
    }
    __syncthreads();
    return;
  }
  unsigned long long lk = 0; int li = -1;                       // previous winner: candidates must be strictly greater
  for (int r = 0; r < k; ++r) {
    unsigned long long bk = kNoKey; int bi = kNoIdx;
    for (int i = tid; i < P; i += kRThreads) {
      unsigned long long key = fit_key(__ldcg(fit + i));
      if (worst) key = ~key;
      const bool after = r == 0 || key > lk || (key == lk && i > li);
      if (after && better(key, i, bk, bi)) { bk = key; bi = i; }
    }
    warp_argmin(bk, bi);
    if (lane == 0) { sc.wkey[wid] = bk; sc.widx[wid] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NW; ++w)
        if (better(sc.wkey[w], sc.widx[w], bk, bi)) { bk = sc.wkey[w]; bi = sc.widx[w]; }
      out[r] = bi; *sc.last = bk; *sc.last_idx = bi;
    }
    __syncthreads();
    lk = *sc.last; li = *sc.last_idx;
  }
}

// n 4-byte words from src to dst by one warp: 16-byte vectors when both are aligned, four in flight per lane
__device__ __forceinline__ void warp_copy_words(void *dst, const void *src, int n, int lane) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
    const int nv = n >> 2;
    const int4 *s4 = reinterpret_cast<const int4 *>(src); int4 *d4 = reinterpret_cast<int4 *>(dst);
    int i = lane;
    for (; i + 96 < nv; i += 128) {
      const int4 a = __ldcg(s4 + i), b = __ldcg(s4 + i + 32), c = __ldcg(s4 + i + 64), d = __ldcg(s4 + i + 96);
      d4[i] = a; d4[i + 32] = b; d4[i + 64] = c; d4[i + 96] = d;
    }
    for (; i < nv; i += 32) d4[i] = __ldcg(s4 + i);
    for (int j = (nv << 2) + lane; j < n; j += 32) reinterpret_cast<int *>(dst)[j] = __ldcg(reinterpret_cast<const int *>(src) + j);
  } else {
    for (int i = lane; i < n; i += 32) reinterpret_cast<int *>(dst)[i] = __ldcg(reinterpret_cast<const int *>(src) + i);
  }
}

#ifdef AHX_GA_TIMING
__device__ long long g_isl_ticks[8];   // CTA 0, thread 0: clock64 totals of the exchange's phases
#define AHX_ISL_TICK(i) do { if (cta == 0 && tid == 0) { const long long n_ = clock64(); g_isl_ticks[i] += n_ - tisl; tisl = n_; } } while (0)
#else
#define AHX_ISL_TICK(i) do { } while (0)
#endif
// One exchange at generation `gen` (a multiple of migrate_every; `cur` = gen & 1 is the population just
// built).  Called by every thread of every CTA after the generation's bookkeeping.  Export: CTA c writes
// this island's elites into the mailbox of island (island + 1 + c) mod W.  Import: CTA 0 waits for the
// other islands' flags of this epoch, replaces the worst individuals, refreshes the hall of fame and the
// row map and takes the collective stop decision; one grid barrier publishes it.
__device__ __forceinline__ void ga_island_exchange(const GaRes &q, int N, long long gen, double &hof_fit, long long &updated, int &stop,
                                                   unsigned int &nsync, uint32_t *rused, uint32_t *rpref, uint32_t *scan_tmp, unsigned char *isl_smem) {
  const int tid = threadIdx.x, G = gridDim.x, cta = blockIdx.x, W = q.n_islands, ne = q.n_elite, L = q.L;
  const int cur = static_cast<int>(gen & 1);
  int *rowc = q.rowof[cur]; double *fitc = q.fit[cur];
  GaIslands *isl = q.isl;
  const unsigned long long epoch = static_cast<unsigned long long>(gen / q.migrate_every);
  const int par = static_cast<int>(epoch & 1);
  const size_t flags_b = isl_flags_bytes(W), head_b = isl_head_bytes(ne), slot_b = isl_slot_bytes(ne, L, N);
  const IslScratch sc = isl_scratch(isl_smem);
  const bool exporter = cta < W - 1;
#ifdef AHX_GA_TIMING
  long long tisl = clock64();
#endif
  if (exporter || cta == 0) isl_select_k(fitc, q.P, ne, false, sc.sel, sc);          // elites: the same list in every CTA that needs it
  AHX_ISL_TICK(0);   // elite selection
  for (int c = cta; c < W - 1; c += G) {
    const int dst = (q.island + 1 + c) % W;
    unsigned char *slot = isl->mbox[dst] + flags_b + (static_cast<size_t>(q.island) * 2 + par) * slot_b;
    double *head = reinterpret_cast<double *>(slot);
    if (tid == 0) head[0] = hof_fit;
    if (tid < ne) head[1 + tid] = __ldcg(fitc + sc.sel[tid]);
    int *dt = reinterpret_cast<int *>(slot + head_b);
    uint32_t *dx = reinterpret_cast<uint32_t *>(dt + static_cast<size_t>(ne) * L);
    for (int w = tid >> 5; w < 2 * ne; w += kRThreads / 32) {          // one warp per (elite, tour row | coordinate row)
      const int e = w >> 1, row = __ldcg(rowc + sc.sel[e]);
      if (w & 1) warp_copy_words(dx + static_cast<size_t>(e) * N, q.xpool + static_cast<size_t>(row) * N, N, tid & 31);
      else warp_copy_words(dt + static_cast<size_t>(e) * L, q.pool + static_cast<size_t>(row) * L, L, tid & 31);
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) st_release_sys_u64(reinterpret_cast<unsigned long long *>(isl->mbox[dst] + 32 * static_cast<size_t>(q.island)), epoch);
  }
  AHX_ISL_TICK(1);   // export
  if (cta == 0) {
    const unsigned char *mine = isl->mbox[q.island];
    if (tid == 0) sc.last_idx[2] = 0;
    __syncthreads();
    if (tid < W && tid != q.island) {
      const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(mine + 32 * static_cast<size_t>(tid));
      const long long t0 = global_timer_ns();
      unsigned int spins = 0;
      while (ld_acquire_sys_u64(flag) < epoch) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > isl->timeout_ns) { sc.last_idx[2] = 1; break; }
      }
    }
    __syncthreads();
    const bool timed_out = sc.last_idx[2] != 0;
    AHX_ISL_TICK(2);   // wait for the peers' flags
    const int m = ne * (W - 1);
    double g_best = isl->g_best; long long g_updated = isl->g_updated;
    int win = -1;                                                                 // thread 0: the migrant that beats the hall of fame
    if (!timed_out) {
      isl_select_k(fitc, q.P, m, true, sc.sel + kIslMaxMigrants, sc);             // worst first
      ga_row_map(q, q.used + static_cast<size_t>(gen % 3) * ga_wr_pad(q.P), rused, rpref, scan_tmp);
      const int WR = (2 * q.P + 31) / 32;
      if (tid < m) sc.rows[tid] = mask_select(rused, rpref, WR, static_cast<uint32_t>(tid), true, 2 * q.P);
      __syncthreads();
      AHX_ISL_TICK(3);   // worst selection + free rows
      // migrant j: elite j % ne of the (j / ne)-th OTHER island, in island order.  One warp per (migrant, row) copies,
      // thread j adopts migrant j (row id + fitness of the individual it replaces), then thread 0 walks the m
      // fitness values in order (first of equals wins the hall of fame).
      auto slot_of = [&](int j) {
        const int s0 = j / ne, src = s0 < q.island ? s0 : s0 + 1;
        return mine + flags_b + (static_cast<size_t>(src) * 2 + par) * slot_b;
      };
      for (int w = tid >> 5; w < 2 * m; w += kRThreads / 32) {
        const int j = w >> 1, e = j % ne;
        const int *st = reinterpret_cast<const int *>(slot_of(j) + head_b);
        const uint32_t *sx = reinterpret_cast<const uint32_t *>(st + static_cast<size_t>(ne) * L);
        if (w & 1) warp_copy_words(q.xpool + static_cast<size_t>(sc.rows[j]) * N, sx + static_cast<size_t>(e) * N, N, tid & 31);
        else warp_copy_words(q.pool + static_cast<size_t>(sc.rows[j]) * L, st + static_cast<size_t>(e) * L, L, tid & 31);
      }
      if (tid < m) {
        const double f = __ldcg(reinterpret_cast<const double *>(slot_of(tid)) + 1 + tid % ne);
        const int ind = sc.sel[kIslMaxMigrants + tid];
        rowc[ind] = sc.rows[tid]; fitc[ind] = f;
        sc.mfit[tid] = f;
      }
      __syncthreads();
      if (tid == 0)
        for (int j = 0; j < m; ++j)
          if (sc.mfit[j] < hof_fit) { hof_fit = sc.mfit[j]; updated = gen; win = j; }   // a migrant beats the hall of fame (first of equals)
      if (tid == 0) {
        double nb = hof_fit;
        for (int s = 0; s < W; ++s)
          if (s != q.island) nb = fmin(nb, __ldcg(reinterpret_cast<const double *>(mine + flags_b + (static_cast<size_t>(s) * 2 + par) * slot_b)));
        if (nb < g_best) { g_best = nb; g_updated = gen; }
      }
    }
    if (tid == 0) {
      sc.last_idx[1] = win;
      isl->g_best = g_best; isl->g_updated = g_updated; isl->epochs = epoch;
      if (timed_out) { isl->err = 1; stop = 1; }
      else if (gen - g_updated > q.ngen || gen >= q.max_gens) stop = 1;
      isl->x_stop = stop;
    }
    __syncthreads();
    if (!timed_out) {
      if (sc.last_idx[1] >= 0) {                                                      // hall of fame <- the winning migrant's row
        const int *best_row = q.pool + static_cast<size_t>(sc.rows[sc.last_idx[1]]) * L;
        for (int i = tid; i < L; i += kRThreads) q.hof[i] = best_row[i];
      }
      // exact row map of the edited generation (fresh rows are allocated from its complement)
      uint32_t *used_cur = q.used + static_cast<size_t>(gen % 3) * ga_wr_pad(q.P);
      for (int i = tid; i < ga_wr_pad(q.P); i += kRThreads) used_cur[i] = 0u;
      __syncthreads();
      for (int i0 = tid; i0 < q.P; i0 += 8 * kRThreads) {                           // loads first, then the (fire-and-forget) atomics
        int rr[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) rr[u] = i0 + u * kRThreads < q.P ? __ldcg(rowc + i0 + u * kRThreads) : -1;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (rr[u] >= 0) atomicOr(used_cur + (rr[u] >> 5), 1u << (rr[u] & 31));
      }
    }
    if (tid == 0) { isl->x_hof_fit = hof_fit; isl->x_updated = updated; }
    AHX_ISL_TICK(4);   // copies + bookkeeping + row map
  }
  ++nsync;
  grid_barrier(q.gbar, nsync * static_cast<unsigned int>(G));
  hof_fit = __ldcg(&isl->x_hof_fit); updated = __ldcg(&isl->x_updated); stop = __ldcg(&isl->x_stop);
  AHX_ISL_TICK(5);   // closing barrier
}

// SUB: the tours cover a subset of the group's contigs (L < N: a resumed run whose tour file lists fewer
// contigs, clm.go:400-417); the evaluate side then uses the sub-tour window test (res_pair_terms).
template <int T, bool STREAM, bool SUB>
__global__ void __launch_bounds__(kRThreads, 1) ga_persist_kernel(ResArgs g, GaRes q, long long n_gens) {
  extern __shared__ __align__(128) unsigned char smem_res[];
  const ResSmem m = res_carve<T, STREAM>(smem_res, g);
  const int tid = threadIdx.x, G = gridDim.x, cta = blockIdx.x;
  const int WM = (q.P + 31) / 32, WR = (2 * q.P + 31) / 32;
  // mutation mask + exclusive word prefix of generations g (buffer g & 1) and g + 1, row bit map + prefix of the free rows
  const int WMp = ga_wm_pad(q.P), WRp = ga_wr_pad(q.P);                 // padded to 16 bytes: the words arrive by TMA bulk copy
  uint32_t *mbase = reinterpret_cast<uint32_t *>(m.extra), *rused = mbase + 4 * WMp, *rpref = rused + WRp;
  const size_t slot_bytes = res_align16(sizeof(GaSlot) * static_cast<size_t>(q.nbmax) * T);
  unsigned char *slot_base = m.extra + res_align16(4ull * (4 * WMp + 2 * WRp));   // two slot tables (generation g & 1)
  uint32_t *scan_tmp = reinterpret_cast<uint32_t *>(slot_base + 2 * slot_bytes);   // kRThreads/32 + 1 words
  unsigned char *isl_smem = reinterpret_cast<unsigned char *>(scan_tmp) + res_align16(4 * (kRThreads / 32 + 1));
  res_init<STREAM>(m, g, true);
  int kdone = 0;
  bool table_ready = g.E_res == 0;
  const Philox ph{q.key0, q.key1};
  // generation state: identical in every thread of every CTA
  long long gen = q.st->gen, updated = q.st->updated, evals = q.st->evals;
  double hof_fit = q.st->hof_fit;
  int stop = q.st->stop;
  unsigned int nsync = 0;
  uint32_t Mb[2] = {0u, 0u};                                           // mutated offspring of the masks in the two buffers
  auto mflag_of = [&](long long gg) { return mbase + static_cast<size_t>(gg & 1) * 2 * WMp; };
  auto gmask_of = [&](long long gg) { return q.gmask + static_cast<size_t>(gg % 3) * WMp; };   // three in rotation: a launch's first CTAs publish generation g + 2 while the last still read g
  auto used_of = [&](long long gg) { return q.used + static_cast<size_t>(gg % 3) * WRp; };
  // Masks are published TWO generations ahead (they do not depend on fitness), so that the mask of generation g + 1
  // is already scanned in shared memory when generation g's barrier falls.  The first two of this launch: here.
  if (n_gens > 0 && !stop) {
    ga_publish_mask(q, ph, gen, gmask_of(gen), cta, G);
    ga_publish_mask(q, ph, gen + 1, gmask_of(gen + 1), cta, G);
  }
  ++nsync;
  grid_barrier(q.gbar, nsync * static_cast<unsigned int>(G));
  if (n_gens > 0 && !stop) {
    Mb[gen & 1] = ga_load_mask(q, gmask_of(gen), mflag_of(gen), mflag_of(gen) + WMp, scan_tmp);
    Mb[(gen + 1) & 1] = ga_load_mask(q, gmask_of(gen + 1), mflag_of(gen + 1), mflag_of(gen + 1) + WMp, scan_tmp);
    ga_row_map(q, used_of(gen), rused, rpref, scan_tmp);
  }
#ifdef AHX_GA_TIMING   // make EXTRA_NVFLAGS=-DAHX_GA_TIMING: per-phase clock64 totals of the first and last CTA, printed at exit
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define AHX_TICK(i) do { if (tid == 0) { const long long now_ = clock64(); tph[i] += now_ - tprev; tprev = now_; } } while (0)
  long long tq[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tqprev = clock64();   // the same for the first slot thread
#define AHX_TICKQ(i) do { if (tid == kRThreads - 32) { const long long now_ = clock64(); tq[i] += now_ - tqprev; tqprev = now_; } } while (0)
#else
#define AHX_TICK(i) do { } while (0)
#define AHX_TICKQ(i) do { } while (0)
#endif
  ulonglong2 *bcast = reinterpret_cast<ulonglong2 *>(isl_smem + res_align16(kIslScratchBytes));
  uint64_t *pbar = reinterpret_cast<uint64_t *>(bcast + 1);               // mbarrier of the bulk copies issued behind each grid barrier
  if (tid == 0) { mbar_init(pbar, 1); mbar_init_fence(); }                 // visible to the block by the barriers below
  // Selection of offspring o of generation gg + 1 (parents: generation gg): the winner, its pool row and fitness.
  // (Fetching both rounds' candidates and their row ids in one round trip -- 18 loads instead of 3 + 3 + 1 + 1 --
  // measured slower: 44.3 against 42.8 us per generation at config3.)
  auto select_parent = [&](long long gg, int o, int *row, double *fit) {
    const int c = static_cast<int>(gg & 1);
    const uint32_t glo = static_cast<uint32_t>(gg), ghi = static_cast<uint32_t>(gg >> 32);
    const int parent = tournament_pair(q.fit[c], q.P, q.k, ph, glo, ghi, static_cast<uint32_t>(o >> 1), o & 1);
    *row = __ldcg(q.rowof[c] + parent); *fit = __ldcg(q.fit[c] + parent);
    return parent;
  };
  // Front half of building generation gg + 1: (1) for the mutated offspring this CTA scores (batch b -> CTA b mod G)
  // the offspring index, mutation plan, tournament parent and the operator's two maps -- a chain of dependent loads
  // (fitness + row id -> tour row -> coordinate row); (2) this CTA's share of the un-mutated offspring, which inherit
  // row and fitness (eaopt keeps Evaluated = true on clones).  The two run in different threads, and nothing here
  // reads what the rest of the set-up (free rows, next mask) produces, so the front half starts right behind the
  // previous generation's barrier, while those loads are in flight and before that generation's bookkeeping.  It
  // writes only generation gg + 1's buffers: if the run turns out to stop at gg they are never read.
  auto front_half = [&](long long gg) {
    const int c = static_cast<int>(gg & 1);
    int *rown = q.rowof[c ^ 1]; double *fitn = q.fit[c ^ 1];
    const uint32_t glo = static_cast<uint32_t>(gg), ghi = static_cast<uint32_t>(gg >> 32);
    const uint32_t *mf = mflag_of(gg), *mp_ = mf + WMp;
    const uint32_t Mg = Mb[gg & 1];
    uint32_t first; int nmine;
    ga_deal(Mg, cta, G, &first, &nmine);
    const int nbl = (nmine + T - 1) / T;
    GaSlot *sl_tab = reinterpret_cast<GaSlot *>(slot_base + static_cast<size_t>(gg & 1) * slot_bytes);
    uint32_t *used_nx = used_of(gg + 1);
    // slot s -> lane s / 24 of warp 23 - s % 24: one slot per WARP while there are at most 24 (their code paths differ:
    // operator, tournament round, so slots sharing a warp would run one after the other); the clones start at thread 0
    for (int s = slot_of_thread(tid); s < nbl * T; s += kRThreads) {
      const uint32_t mi = first + static_cast<uint32_t>(s);
      GaSlot sl{};
      sl.o = -1;
      if (s < nmine) {
        AHX_TICKQ(0);
        const int o = mask_select(mf, mp_, WM, mi, false, q.P);
        const MutPlan mp = plan_mutation(ph, glo, ghi, static_cast<uint32_t>(o), q.mut_prob, q.L);
        AHX_TICKQ(1);   // select + plan
        double pf;
        select_parent(gg, o, &sl.parent_row, &pf);
        AHX_TICKQ(2);   // tournament + row
        remap_of(mp, q.L, q.pool + static_cast<size_t>(sl.parent_row) * q.L, q.xpool + static_cast<size_t>(sl.parent_row) * g.N, q.sizes32, q.total2, &sl);
        AHX_TICKQ(3);   // remap_of
        sl.o = o;
      }
      sl_tab[s] = sl;
    }
    for (int o = cta + tid * G; o < q.P; o += kRThreads * G) {
      if ((mf[o >> 5] >> (o & 31)) & 1u) continue;
      int row; double pf;
      select_parent(gg, o, &row, &pf);
      rown[o] = row;
      atomicOr(used_nx + (row >> 5), 1u << (row & 31));
      fitn[o] = pf;
    }
  };
  // Bookkeeping of a finished generation `gnow` (evaluate.go:287-306) from the 128-bit minimum of its evaluations.
  // Clones are copies of individuals that already exist, so only a mutated offspring can beat the hall of fame;
  // the minimum carries the winner's offspring index and pool row, so there is nothing to agree on: the last CTA
  // copies the row (it stays referenced by generation gnow while gnow + 1 is built) and nobody waits for it.
  ulonglong2 gm = make_ulonglong2(0ULL, ~0ULL);
  uint32_t Mlast = 0;
  auto bookkeeping = [&](long long gnow) {
    const bool improved = gm.y < fit_key(hof_fit);          // updateHallOfFame keeps the best ever; currentBest > *best
    if (improved) {
      hof_fit = key_fit(gm.y); updated = gnow;
      if (cta == G - 1) {
        const int *best_row = q.pool + static_cast<size_t>(gm.x & 0xffffffffULL) * q.L;
        for (int i = tid; i < q.L; i += kRThreads) q.hof[i] = __ldcg(best_row + i);
        if (tid == 0) q.st->best_idx = static_cast<int>(gm.x >> 32);
      }
    }
    evals += Mlast;
    // EarlyStop (evaluate.go:304-306); with islands it is a collective decision taken at the exchange points
    if ((q.n_islands <= 1 && gnow - updated > q.ngen) || gnow >= q.max_gens) stop = 1;
    return improved;
  };
  bool book = false;                                   // generation `gen` exists, its bookkeeping is still to be done
  bool pend = false;                                   // bulk copies of the next mask / this generation's row map are in flight
  uint32_t pphase = 0;
  long long it = 0;
  while (it < n_gens && !stop) {
    // ---- an exchange is due at this generation: it changes the population, so bookkeeping and exchange come first
    if (book && q.n_islands > 1 && gen % q.migrate_every == 0) {
      const bool improved = bookkeeping(gen);
      book = false;
      if (improved) { ++nsync; grid_barrier(q.gbar, nsync * static_cast<unsigned int>(G)); }   // CTA 0 may replace the hall-of-fame tour by a migrant's
      ga_island_exchange(q, g.N, gen, hof_fit, updated, stop, nsync, rused, rpref, scan_tmp, isl_smem);
      AHX_TICK(5);   // island exchange
      if (stop) break;
      // mask of gen + 1 and the (edited) row map of gen, plainly
      Mb[(gen + 1) & 1] = ga_load_mask(q, gmask_of(gen + 1), mflag_of(gen + 1), mflag_of(gen + 1) + WMp, scan_tmp);
      ga_row_map(q, used_of(gen), rused, rpref, scan_tmp);
    }
    front_half(gen);
    if (pend) { ga_wait_pending(pbar, pphase); pphase ^= 1u; gm = *bcast; }
    if (book) {
      bookkeeping(gen);
      book = false;
      if (stop) break;
    }
    AHX_TICK(4);   // front half + bookkeeping
    const int cur = static_cast<int>(gen & 1);
    int *rown = q.rowof[cur ^ 1];
    double *fitn = q.fit[cur ^ 1];
    uint32_t *used_new = used_of(gen + 1);        // marked while generation gen+1 is built
    uint32_t *used_old = used_of(gen + 2);        // generation gen-1's map: nobody reads it any more
    ulonglong2 *gmin_new = q.gmin + (gen + 1) % 3;
    GaSlot *slots = reinterpret_cast<GaSlot *>(slot_base + static_cast<size_t>(gen & 1) * slot_bytes);
    const uint32_t Mcur = Mb[gen & 1];
    uint32_t first_mi; int n_mine;
    ga_deal(Mcur, cta, G, &first_mi, &n_mine);
    const int nb_local = (n_mine + T - 1) / T;
    // ---- the loads issued behind the previous barrier have landed: free-row select structure of this generation,
    // mask of the next one
    if (pend) {
      Mb[(gen + 1) & 1] = ga_finish_mask_rows(q, pbar, pphase, mflag_of(gen + 1), mflag_of(gen + 1) + WMp, rused, rpref, scan_tmp);
      pend = false;
    }
    AHX_TICKQ(4);   // finish (scan)
    // ---- slot table, back half: the k-th mutated offspring takes the k-th free row (a thread finishes the slots it began)
    for (int s = slot_of_thread(tid); s < nb_local * T; s += kRThreads) {
      const int o = slots[s].o;
      if (o >= 0) {
        const int child = mask_select(rused, rpref, WR, first_mi + static_cast<uint32_t>(s), true, 2 * q.P);
        slots[s].child_row = child;
        rown[o] = child;
        atomicOr(used_new + (child >> 5), 1u << (child & 31));
      }
    }
    AHX_TICKQ(5);   // back half
    // ---- housekeeping for later generations: this CTA's words of the mutation mask of generation gen + 2,
    // clear generation gen-1's row map (this CTA's slice), reset its minimum
    ga_publish_mask(q, ph, gen + 2, gmask_of(gen + 2), cta, G);
    {
      for (int i = cta * kRThreads + tid; i < WRp; i += G * kRThreads) used_old[i] = 0u;
      if (cta == 0 && tid == 0) q.gmin[(gen + 2) % 3] = make_ulonglong2(~0ULL, ~0ULL);
    }
    AHX_TICKQ(6);   // publish + clear
    __syncthreads();
    AHX_TICKQ(7);   // sync
    AHX_TICK(1);   // rest of the set-up
    // ---- score this CTA's batches; an individual that beats the hall of fame goes to gmin_new
    GaSrc<T, SUB> src{q.pool, q.pool, q.xpool, q.xpool, slots, fitn, gmin_new, hof_fit, n_mine};
    res_run_batches<T, SUB, STREAM>(g, m, src, 0, 1, nb_local, kdone, table_ready);
    AHX_TICK(2);   // build + score
    ++nsync;
    // what the next generation needs from global memory starts to move the moment the barrier opens
    const bool more = it + 1 < n_gens && gen + 1 < q.max_gens;
    const bool issue = more && !(q.n_islands > 1 && (gen + 1) % q.migrate_every == 0);
    grid_barrier_fetch(q.gbar, nsync * static_cast<unsigned int>(G), gmin_new, bcast, [&]() {
      if (issue) ga_issue_mask_rows(q, gmask_of(gen + 2), used_new, gmin_new, mflag_of(gen + 2), rused, bcast, pbar);
      return issue;
    });
    AHX_TICK(3);   // grid barrier
    AHX_TICKQ(8);   // batches + barrier
    if (!issue) gm = *bcast;                              // else: read once the copies have landed
    Mlast = Mcur;
    ++it; ++gen;
    book = true;
    pend = issue;
  }
  if (book) {
    // the launch ends on a freshly built generation: its bookkeeping, and the exchange if one is due at it
    const bool improved = bookkeeping(gen);
    if (q.n_islands > 1 && gen % q.migrate_every == 0) {
      if (improved) { ++nsync; grid_barrier(q.gbar, nsync * static_cast<unsigned int>(G)); }
      ga_island_exchange(q, g.N, gen, hof_fit, updated, stop, nsync, rused, rpref, scan_tmp, isl_smem);
    }
  }
#ifdef AHX_GA_TIMING
  if (tid == kRThreads - 32 && cta == 0)
    printf("cta 0 slot thread: to-slot %lld  select+plan %lld  tournament+row %lld  remap_of %lld  finish %lld  back %lld  publish+clear %lld  sync %lld  batches+barrier %lld\n",
           tq[0], tq[1], tq[2], tq[3], tq[4], tq[5], tq[6], tq[7], tq[8]);
  if (tid == 0 && cta == 0 && q.n_islands > 1)
    printf("island %d exchange phases (cycles, totals): elites %lld  export %lld  wait-flags %lld  worst+rows %lld  import %lld  barrier %lld\n",
           q.island, g_isl_ticks[0], g_isl_ticks[1], g_isl_ticks[2], g_isl_ticks[3], g_isl_ticks[4], g_isl_ticks[5]);
  if (tid == 0 && (cta == 0 || cta == G - 1))
    printf("cta %d (%d batches/gen at the end): rowmap %lld  slots %lld  build+score %lld  barrier %lld  mask+hof %lld  exchange %lld cycles over %lld generations\n",
           cta, 0, tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], n_gens);
#endif
  if (cta == 0 && tid == 0) {
    q.st->gen = gen; q.st->updated = updated; q.st->evals = evals; q.st->hof_fit = hof_fit; q.st->stop = stop;
  }
}

}  // namespace ahx

#endif  // AHX_GA_RES_CUH_
