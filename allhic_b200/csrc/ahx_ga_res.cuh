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
// gets a fresh row from the rows generation g does not reference (at least P of the 2P).
// The k-th mutated offspring (ascending index) takes the k-th free row, so the layout
// is a pure function of the seed.
//
// One generation, every CTA (G = gridDim.x, 24 warps):
//   1. bit mask of the pool rows the current generation references (from a byte map in
//      global memory that the CTAs filled while they built that generation) + word prefix
//   2. slot table: for the mutated offspring this CTA scores (batch b -> CTA b mod G):
//      offspring index (select on the mutation mask), tournament parents, mutation plan,
//      fresh row (select on the free rows)
//   3. the CTA's share of un-mutated offspring (o = c, c + G, ...): tournament, inherit
//   4. res_run_batches: the coordinates of T offspring at a time are built straight from
//      the parent rows through the mutation's index remap (the child rows are written on the
//      way), consumers score them (ahx_resident.cuh); atomicMin of the new fitness values
//   5. (during 2-3) this CTA's words of the NEXT generation's mutation mask (first Philox draw
//      of each offspring; independent of fitness) to global memory
//   6. grid barrier; load the whole mask, word prefix
//   7. bookkeeping (evaluate.go:287-306): only a mutated offspring can beat the hall of fame,
//      so one load of the generation's minimum decides; an improvement (rare) costs a second
//      barrier to agree on the lowest offspring index and copy its row
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
  unsigned char *used;  // [3][pad32(2P)] byte maps: used[g % 3][r] = 1 iff generation g references pool row r
  unsigned long long *gmin;   // [3] order-preserving key of the best fitness evaluated while building generation g (index g % 3)
  int *gidx;            // lowest offspring index holding that fitness (0x7fffffff when idle)
  uint32_t *gmask;      // [2][ceil(P/32)] mutation masks of generations g (index g & 1), filled one generation ahead
  // islands (include/ahx.h "Islands"); n_islands <= 1: single population
  int n_islands, island, migrate_every, n_elite;
  GaIslands *isl;
};

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
constexpr size_t kIslScratchBytes = 16 * (kRThreads / 32) + 32 + 4 * (3 * kIslMaxMigrants) + 16;
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

inline size_t ga_res_extra_bytes(int P, int T, int nbmax) {
  const size_t wm = (static_cast<size_t>(P) + 31) / 32, wr = (2 * static_cast<size_t>(P) + 31) / 32;
  return res_align16(4 * (2 * wm + 2 * wr)) + res_align16(sizeof(GaSlot) * static_cast<size_t>(nbmax) * T) + res_align16(4 * (kRThreads / 32 + 1)) +
         res_align16(kIslScratchBytes) + 64;
}

// order-preserving map double -> uint64 (for atomicMin on fitness values) and back
__device__ __forceinline__ unsigned long long fit_key(double f) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(f));
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double key_fit(unsigned long long k) {
  return __longlong_as_double(static_cast<long long>((k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k));
}

template <int T, bool SUB = false>
struct GaSrc {
  const int *pool; int *pool_w; const uint32_t *xpool; uint32_t *xpool_w; const GaSlot *slots; double *fitn; unsigned long long *gmin;
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
  __device__ __forceinline__ void store(int j, int t, double score) const {
    const int o = slots[j * T + t].o;
    if (o >= 0) { fitn[o] = score; atomicMin(gmin, fit_key(score)); }
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

// index of the m-th (0-based) set bit of a mask given the exclusive word prefix
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

__device__ __forceinline__ void grid_barrier(unsigned int *gbar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(gbar, 1u);
    unsigned int v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gbar) : "memory"); } while (v < target);
    __threadfence();
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

// Bit mask of the pool rows a generation references (from its byte map in global memory, filled
// while that generation was built) and the exclusive word prefix of the FREE rows, for mask_select.
__device__ __forceinline__ void ga_row_map(const GaRes &q, const unsigned char *used_cur, uint32_t *rused, uint32_t *rpref, uint32_t *scan_tmp) {
  const int tid = threadIdx.x, WR = (2 * q.P + 31) / 32;
  for (int wi = tid; wi < WR; wi += kRThreads) {
    uint32_t w = 0;
    const uint4 *src = reinterpret_cast<const uint4 *>(used_cur + static_cast<size_t>(wi) * 32);
    const int nvalid = min(32, 2 * q.P - wi * 32);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (nvalid <= h * 16) break;
      const uint4 v = __ldcg(src + h);                       // the map is padded to a multiple of 32 bytes
      const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b) w |= ((x[i] >> (8 * b)) & 1u) << (h * 16 + i * 4 + b);
    }
    rused[wi] = w;
    w = ~w;
    if (nvalid < 32) w &= (1u << nvalid) - 1u;
    rpref[wi] = __popc(w);
  }
  __syncthreads();
  block_scan_array(rpref, WR, scan_tmp);   // at least P of the 2P rows are free
  __syncthreads();
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
  unsigned long long *wkey;   // [warps]
  int *widx;                  // [warps]
  unsigned long long *last;   // [1] key of the previous winner
  int *last_idx;              // [1] (and [2]: time-out flag)
  int *sel;                   // [2 * kIslMaxMigrants] selected individuals
  int *rows;                  // [kIslMaxMigrants] pool rows of the migrants
};
__device__ __forceinline__ IslScratch isl_scratch(unsigned char *p) {
  IslScratch s;
  s.wkey = reinterpret_cast<unsigned long long *>(p); p += 8 * (kRThreads / 32);
  s.last = reinterpret_cast<unsigned long long *>(p); p += 8;
  s.widx = reinterpret_cast<int *>(p); p += 4 * (kRThreads / 32);
  s.last_idx = reinterpret_cast<int *>(p); p += 16;
  s.sel = reinterpret_cast<int *>(p); p += 4 * 2 * kIslMaxMigrants;
  s.rows = reinterpret_cast<int *>(p);
  return s;
}

// The k individuals with the smallest (key, index), ascending, into out[0..k): key = fit_key(fitness)
// (best first, lowest index on ties) or its complement (worst first, lowest index on ties).  Whole block.
__device__ __forceinline__ void isl_select_k(const double *fit, int P, int k, bool worst, int *out, const IslScratch &sc) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned long long lk = 0; int li = -1;                       // previous winner: candidates must be strictly greater
  for (int r = 0; r < k; ++r) {
    unsigned long long bk = ~0ULL; int bi = 0x7fffffff;
    for (int i = tid; i < P; i += kRThreads) {
      unsigned long long key = fit_key(__ldcg(fit + i));
      if (worst) key = ~key;
      const bool after = r == 0 || key > lk || (key == lk && i > li);
      if (after && (key < bk || (key == bk && i < bi))) { bk = key; bi = i; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, d);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
      if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
    }
    if (lane == 0) { sc.wkey[wid] = bk; sc.widx[wid] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kRThreads / 32; ++w)
        if (sc.wkey[w] < bk || (sc.wkey[w] == bk && sc.widx[w] < bi)) { bk = sc.wkey[w]; bi = sc.widx[w]; }
      out[r] = bi; *sc.last = bk; *sc.last_idx = bi;
    }
    __syncthreads();
    lk = *sc.last; li = *sc.last_idx;
  }
}

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
  if (exporter || cta == 0) isl_select_k(fitc, q.P, ne, false, sc.sel, sc);          // elites: the same list in every CTA that needs it
  for (int c = cta; c < W - 1; c += G) {
    const int dst = (q.island + 1 + c) % W;
    unsigned char *slot = isl->mbox[dst] + flags_b + (static_cast<size_t>(q.island) * 2 + par) * slot_b;
    double *head = reinterpret_cast<double *>(slot);
    if (tid == 0) head[0] = hof_fit;
    if (tid < ne) head[1 + tid] = __ldcg(fitc + sc.sel[tid]);
    int *dt = reinterpret_cast<int *>(slot + head_b);
    uint32_t *dx = reinterpret_cast<uint32_t *>(dt + static_cast<size_t>(ne) * L);
    for (int e = 0; e < ne; ++e) {
      const int row = __ldcg(rowc + sc.sel[e]);
      const int *pt = q.pool + static_cast<size_t>(row) * L;
      const uint32_t *px = q.xpool + static_cast<size_t>(row) * N;
      for (int i = tid; i < L; i += kRThreads) dt[static_cast<size_t>(e) * L + i] = __ldcg(pt + i);
      for (int i = tid; i < N; i += kRThreads) dx[static_cast<size_t>(e) * N + i] = __ldcg(px + i);
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) st_release_sys_u64(reinterpret_cast<unsigned long long *>(isl->mbox[dst] + 32 * static_cast<size_t>(q.island)), epoch);
  }
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
    const int m = ne * (W - 1);
    double g_best = isl->g_best; long long g_updated = isl->g_updated;
    int win = -1;                                                                 // thread 0: the migrant that beats the hall of fame
    if (!timed_out) {
      isl_select_k(fitc, q.P, m, true, sc.sel + kIslMaxMigrants, sc);             // worst first
      ga_row_map(q, q.used + static_cast<size_t>(gen % 3) * ((2 * static_cast<size_t>(q.P) + 31) / 32 * 32), rused, rpref, scan_tmp);
      const int WR = (2 * q.P + 31) / 32;
      if (tid < m) sc.rows[tid] = mask_select(rused, rpref, WR, static_cast<uint32_t>(tid), true, 2 * q.P);
      __syncthreads();
      // migrant j: elite j % ne of the (j / ne)-th OTHER island, in island order
      for (int j = 0; j < m; ++j) {
        const int s0 = j / ne, src = s0 < q.island ? s0 : s0 + 1, e = j % ne;
        const unsigned char *slot = mine + flags_b + (static_cast<size_t>(src) * 2 + par) * slot_b;
        const int *st = reinterpret_cast<const int *>(slot + head_b);
        const uint32_t *sx = reinterpret_cast<const uint32_t *>(st + static_cast<size_t>(ne) * L);
        int *pt = q.pool + static_cast<size_t>(sc.rows[j]) * L;
        uint32_t *px = q.xpool + static_cast<size_t>(sc.rows[j]) * N;
        for (int i = tid; i < L; i += kRThreads) pt[i] = __ldcg(st + static_cast<size_t>(e) * L + i);
        for (int i = tid; i < N; i += kRThreads) px[i] = __ldcg(sx + static_cast<size_t>(e) * N + i);
        if (tid == 0) {
          const double f = __ldcg(reinterpret_cast<const double *>(slot) + 1 + e);
          const int ind = sc.sel[kIslMaxMigrants + j];
          rowc[ind] = sc.rows[j]; fitc[ind] = f;
          if (f < hof_fit) { hof_fit = f; updated = gen; win = j; }                 // a migrant beats the hall of fame (first of equals)
        }
      }
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
      unsigned char *used_cur = q.used + static_cast<size_t>(gen % 3) * ((2 * static_cast<size_t>(q.P) + 31) / 32 * 32);
      uint32_t *z = reinterpret_cast<uint32_t *>(used_cur);
      for (size_t i = tid; i < (2 * static_cast<size_t>(q.P) + 31) / 32 * 8; i += kRThreads) z[i] = 0u;
      __syncthreads();
      for (int i = tid; i < q.P; i += kRThreads) used_cur[__ldcg(rowc + i)] = 1;
    }
    if (tid == 0) { isl->x_hof_fit = hof_fit; isl->x_updated = updated; }
  }
  ++nsync;
  grid_barrier(q.gbar, nsync * static_cast<unsigned int>(G));
  hof_fit = __ldcg(&isl->x_hof_fit); updated = __ldcg(&isl->x_updated); stop = __ldcg(&isl->x_stop);
}

// SUB: the tours cover a subset of the group's contigs (L < N: a resumed run whose tour file lists fewer
// contigs, clm.go:400-417); the evaluate side then uses the sub-tour window test (res_pair_terms).
template <int T, bool STREAM, bool SUB>
__global__ void __launch_bounds__(kRThreads, 1) ga_persist_kernel(ResArgs g, GaRes q, long long n_gens) {
  extern __shared__ __align__(128) unsigned char smem_res[];
  const ResSmem m = res_carve<T, STREAM>(smem_res, g);
  const int tid = threadIdx.x, G = gridDim.x, cta = blockIdx.x;
  const int WM = (q.P + 31) / 32, WR = (2 * q.P + 31) / 32;
  uint32_t *mflag = reinterpret_cast<uint32_t *>(m.extra), *mpref = mflag + WM, *rused = mpref + WM, *rpref = rused + WR;
  GaSlot *slots = reinterpret_cast<GaSlot *>(m.extra + res_align16(4ull * (2 * WM + 2 * WR)));
  uint32_t *scan_tmp = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(slots) + res_align16(sizeof(GaSlot) * static_cast<size_t>(q.nbmax) * T));   // kRThreads/32 + 1 words
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
  const size_t ub = (2 * static_cast<size_t>(q.P) + 31) / 32 * 32;   // bytes per row-reference map (padded)
  // the first generation of this launch: every CTA publishes its share of the mask, one extra barrier
  if (n_gens > 0 && !stop) ga_publish_mask(q, ph, gen, q.gmask + static_cast<size_t>(gen & 1) * WM, cta, G);
  ++nsync;
  grid_barrier(q.gbar, nsync * static_cast<unsigned int>(G));
  uint32_t M = (n_gens > 0 && !stop) ? ga_load_mask(q, q.gmask + static_cast<size_t>(gen & 1) * WM, mflag, mpref, scan_tmp) : 0u;
#ifdef AHX_GA_TIMING   // make EXTRA_NVFLAGS=-DAHX_GA_TIMING: per-phase clock64 totals of the first and last CTA, printed at exit
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define AHX_TICK(i) do { if (tid == 0) { const long long now_ = clock64(); tph[i] += now_ - tprev; tprev = now_; } } while (0)
#else
#define AHX_TICK(i) do { } while (0)
#endif
  for (long long it = 0; it < n_gens && !stop; ++it) {
    const int cur = static_cast<int>(gen & 1);
    const int *rowc = q.rowof[cur]; int *rown = q.rowof[cur ^ 1];
    const double *fitc = q.fit[cur]; double *fitn = q.fit[cur ^ 1];
    const uint32_t glo = static_cast<uint32_t>(gen), ghi = static_cast<uint32_t>(gen >> 32);
    const unsigned char *used_cur = q.used + static_cast<size_t>(gen % 3) * ub;          // rows generation `gen` references
    unsigned char *used_new = q.used + static_cast<size_t>((gen + 1) % 3) * ub;        // marked while generation gen+1 is built
    unsigned char *used_old = q.used + static_cast<size_t>((gen + 2) % 3) * ub;        // generation gen-1's map: nobody reads it any more
    unsigned long long *gmin_new = q.gmin + (gen + 1) % 3;
    // ---- rows the current generation references -> free-row select structure
    ga_row_map(q, used_cur, rused, rpref, scan_tmp);
    AHX_TICK(0);   // row map -> free-row select
    // ---- slot table of the mutated offspring this CTA scores
    const uint32_t Mcur = M;
    const int nb_total = (static_cast<int>(Mcur) + T - 1) / T;
    const int nb_local = cta < nb_total ? (nb_total - cta + G - 1) / G : 0;
    for (int s = kRThreads - 1 - tid; s < nb_local * T; s += kRThreads) {   // threads from the top of the block; the clones below start at thread 0
      const int j = s / T, t = s % T;
      const uint32_t mi = static_cast<uint32_t>((cta + j * G) * T + t);
      GaSlot sl{};
      sl.o = -1;
      if (mi < Mcur) {
        const int o = mask_select(mflag, mpref, WM, mi, false, q.P);
        const MutPlan mp = plan_mutation(ph, glo, ghi, static_cast<uint32_t>(o), q.mut_prob, q.L);
        const int parent = tournament_pair(fitc, q.P, q.k, ph, glo, ghi, static_cast<uint32_t>(o >> 1), o & 1);
        sl.parent_row = __ldcg(rowc + parent);
        remap_of(mp, q.L, q.pool + static_cast<size_t>(sl.parent_row) * q.L, q.xpool + static_cast<size_t>(sl.parent_row) * g.N, q.sizes32, q.total2, &sl);
        sl.o = o;
        sl.child_row = mask_select(rused, rpref, WR, mi, true, 2 * q.P);
        rown[o] = sl.child_row;
        used_new[sl.child_row] = 1;
      }
      slots[s] = sl;
    }
    // ---- un-mutated offspring: inherit row and fitness (eaopt keeps Evaluated = true on clones)
    for (int o = cta + tid * G; o < q.P; o += kRThreads * G) {
      if ((mflag[o >> 5] >> (o & 31)) & 1u) continue;
      const int parent = tournament_pair(fitc, q.P, q.k, ph, glo, ghi, static_cast<uint32_t>(o >> 1), o & 1);
      const int row = __ldcg(rowc + parent);
      rown[o] = row;
      used_new[row] = 1;
      fitn[o] = __ldcg(fitc + parent);
    }
    // ---- housekeeping for later generations: this CTA's words of the next generation's mutation mask,
    // clear generation gen-1's row map (this CTA's slice), reset its minimum
    ga_publish_mask(q, ph, gen + 1, q.gmask + static_cast<size_t>((gen + 1) & 1) * WM, cta, G);
    {
      const size_t words = ub / 4;
      uint32_t *z = reinterpret_cast<uint32_t *>(used_old);
      for (size_t i = static_cast<size_t>(cta) * kRThreads + tid; i < words; i += static_cast<size_t>(G) * kRThreads) z[i] = 0u;
      if (cta == 0 && tid == 0) q.gmin[(gen + 2) % 3] = ~0ULL;
    }
    __syncthreads();
    AHX_TICK(1);   // slots + clones
    // ---- score this CTA's batches; the generation's best new fitness goes to gmin_new
    GaSrc<T, SUB> src{q.pool, q.pool, q.xpool, q.xpool, slots, fitn, gmin_new};
    res_run_batches<T, SUB, STREAM>(g, m, src, 0, 1, nb_local, kdone, table_ready);
    const long long ngen_now = gen + 1;
    __syncthreads();
    AHX_TICK(2);   // build + score
    ++nsync;
    grid_barrier(q.gbar, nsync * static_cast<unsigned int>(G));
    AHX_TICK(3);   // grid barrier
    const unsigned long long kmin = __ldcg(gmin_new);   // in flight together with the mask words (one L2 round trip, not two)
    M = ga_load_mask(q, q.gmask + static_cast<size_t>(ngen_now & 1) * WM, mflag, mpref, scan_tmp);
    // ---- bookkeeping (evaluate.go:287-306).  Clones are copies of individuals that already exist, so
    // only a mutated offspring can beat the hall of fame: compare the minimum over this generation's
    // evaluations; on a (rare) improvement the owners of that fitness agree on the lowest offspring index.
    if (kmin < fit_key(hof_fit)) {          // updateHallOfFame keeps the best ever; currentBest > *best
      hof_fit = key_fit(kmin); updated = ngen_now;
      for (int s = tid; s < nb_local * T; s += kRThreads) {
        const int o = slots[s].o;
        if (o >= 0 && fit_key(__ldcg(fitn + o)) == kmin) atomicMin(q.gidx, o);
      }
      ++nsync;
      grid_barrier(q.gbar, nsync * static_cast<unsigned int>(G));
      if (cta == 0) {
        const int bi = __ldcg(q.gidx);
        const int *best_row = q.pool + static_cast<size_t>(__ldcg(rown + bi)) * q.L;
        for (int i = tid; i < q.L; i += kRThreads) q.hof[i] = __ldcg(best_row + i);
        __syncthreads();
        if (tid == 0) { q.st->best_idx = bi; *q.gidx = 0x7fffffff; }   // next use is at least one grid barrier away
      }
    }
    evals += Mcur;
    // EarlyStop (evaluate.go:304-306); with islands it is a collective decision taken at the exchange points
    if ((q.n_islands <= 1 && ngen_now - updated > q.ngen) || ngen_now >= q.max_gens) stop = 1;
    gen = ngen_now;
    AHX_TICK(4);   // mask load + hall of fame
    if (q.n_islands > 1 && gen % q.migrate_every == 0) {
      ga_island_exchange(q, g.N, gen, hof_fit, updated, stop, nsync, rused, rpref, scan_tmp, isl_smem);
      AHX_TICK(5);   // island exchange
    }
  }
#ifdef AHX_GA_TIMING
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
