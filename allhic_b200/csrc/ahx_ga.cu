// ahx_ga.cu -- CLM.GARun (evaluate.go:258-315) as one kernel launch per generation.
//
// eaopt v0.4.2 semantics restated: ModGenerational with
// SelTournament{NContestants:k} (offspring generated two at a time from two
// distinct tournament winners, winners cloned with their fitness), CrossRate 0
// (Tour.Crossover is empty, evaluate.go:235), per-offspring mutation with
// probability MutRate by exactly one of the four operators of Tour.Mutate
// (evaluate.go:221-232), only mutated offspring re-scored, HallOfFame of size
// 1, EarlyStop when Generations - updated > NGen (evaluate.go:304-306).
//
// ga_generation_kernel: CTA o builds offspring o.  All random draws come from
// Philox4x32-10 keyed by (seed, island) with counter (generation, index,
// stream), so every CTA recomputes the decisions it needs and the population
// never leaves HBM:
//   1. tournament for offspring pair o/2 -> parent index
//   2. clone + mutate in one pass: new[i] = parent[mut_source(i)] (the four
//      operators are index remaps), written to the other population buffer
//   3. if mutated: Tour.Evaluate in contig space (same code as eval_m_sparse),
//      else inherit the parent's fitness (eaopt keeps Evaluated=true on clones)
//   4. the last CTA to finish (atomic ticket) takes the arg-min of the new
//      generation, updates the hall of fame / generation counters / stop flag.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "ahx_ctx.h"
#include "ahx_device.cuh"
#include "ahx_resident.cuh"

namespace ahx {

constexpr int kMaxContestants = 8;
constexpr uint32_t kStreamSelect = 0x100, kStreamMutate = 0x200;

struct GaDev {
  int P, L, N, E, k;
  double mut_prob;
  uint32_t key0, key1;
  long long ngen, max_gens;
  const long long *sizes;
  const void *coo;
  int *pop[2];
  double *fit[2];
  int *hof;
  GaState *st;
};

struct MutPlan { int mutated, op, p, q, dir; };

// Draws of eaopt's `rng.Float64() < MutRate` and Tour.Mutate (evaluate.go:221-232)
__host__ __device__ __noinline__ MutPlan plan_mutation(const Philox &ph, uint32_t glo, uint32_t ghi, uint32_t o, double mut_prob, int L) {
  uint32_t r[4], s[4];
  ph(glo, ghi, o, kStreamMutate, r);
  MutPlan m{0, -1, 0, 0, 0};
  if (!(rand_unit(r[0], r[1]) < mut_prob)) return m;
  m.mutated = 1;
  const double rd = rand_unit(r[2], r[3]);
  ph(glo, ghi, o, kStreamMutate + 1, s);
  if (rd >= 0.2 && rd < 0.4) {                 // MutSplice: k = Intn(L-1) + 1
    m.op = 1; m.p = rand_below(s[0], L - 1) + 1; m.q = 0;
    return m;
  }
  int p = rand_below(s[0], L), q = rand_below(s[1], L);   // randomTwoInts, evaluate.go:146-154
  if (p > q) { const int t = p; p = q; q = t; }
  m.p = p; m.q = q;
  if (rd < 0.2) m.op = 0;                      // MutPermute
  else if (rd < 0.7) { m.op = 2; m.dir = rand_unit(s[2], s[3]) < 0.5; }  // MutInsertion
  else m.op = 3;                               // MutInversion
  return m;
}

// SelTournament{k}.Apply(2, ...): two distinct winners; `which` picks one.
// The reference's k = 3 (evaluate.go:273) has its own out-of-line instance: loops unrolled, the
// sample in registers, the fitness loads issued together.  Out of line so that this cold, serial,
// register-hungry code does not weigh on the register allocation of the kernels' hot loops.
__device__ __noinline__ int tournament_pair_3(const double *__restrict__ fit, int P, const Philox &ph, uint32_t glo, uint32_t ghi,
                                              uint32_t pair, int which) {
  constexpr int K = 3;
  uint32_t r[8];
  {
    uint32_t o4[4];
    ph(glo, ghi, pair, kStreamSelect, o4);
    r[0] = o4[0]; r[1] = o4[1]; r[2] = o4[2]; r[3] = o4[3];
    ph(glo, ghi, pair, kStreamSelect + 1, o4);
    r[4] = o4[0]; r[5] = o4[1]; r[6] = o4[2]; r[7] = o4[3];
  }
  int banned = -1, winner = -1;
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    if (round > which) break;
    const int pool = P - (banned >= 0 ? 1 : 0);
    int chosen[K], ids[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      int v = rand_below(r[round * K + i], pool - i);   // v-th not yet sampled index
      int pos = 0;
#pragma unroll
      for (int m = 0; m < K; ++m) if (m < i && pos == m && chosen[m] <= v) { ++v; ++pos; }
#pragma unroll
      for (int m = K - 1; m > 0; --m) if (m <= i && m > pos) chosen[m] = chosen[m - 1];
#pragma unroll
      for (int m = 0; m < K; ++m) if (m == pos) chosen[m] = v;
      ids[i] = (banned >= 0 && v >= banned) ? v + 1 : v;
    }
    double f[K];
#pragma unroll
    for (int i = 0; i < K; ++i) f[i] = __ldcg(fit + ids[i]);   // independent loads; written by other CTAs in an earlier generation of the same launch
    int best = ids[0]; double bestfit = f[0];
#pragma unroll
    for (int i = 1; i < K; ++i) if (f[i] < bestfit) { best = ids[i]; bestfit = f[i]; }
    winner = best; banned = best;
  }
  return winner;
}

__device__ inline int tournament_pair(const double *__restrict__ fit, int P, int k, const Philox &ph, uint32_t glo,
                                      uint32_t ghi, uint32_t pair, int which) {
  if (k == 3) return tournament_pair_3(fit, P, ph, glo, ghi, pair, which);
  uint32_t r[2 * kMaxContestants];
#pragma unroll
  for (int c = 0; c < (2 * kMaxContestants) / 4; ++c) {
    if (4 * c >= 2 * k) break;
    uint32_t o4[4];
    ph(glo, ghi, pair, kStreamSelect + c, o4);
    r[4 * c] = o4[0]; r[4 * c + 1] = o4[1]; r[4 * c + 2] = o4[2]; r[4 * c + 3] = o4[3];
  }
  int banned = -1, winner = -1;
  for (int round = 0; round <= which; ++round) {
    const int pool = P - (banned >= 0 ? 1 : 0);
    int chosen[kMaxContestants];
    int best = -1; double bestfit = 0.0;
    for (int i = 0; i < k; ++i) {
      int v = rand_below(r[round * k + i], pool - i);   // v-th not yet sampled index
      int pos = 0;
      while (pos < i && chosen[pos] <= v) { ++v; ++pos; }
      for (int m = i; m > pos; --m) chosen[m] = chosen[m - 1];
      chosen[pos] = v;
      const int id = (banned >= 0 && v >= banned) ? v + 1 : v;
      const double f = __ldcg(fit + id);   // written by other CTAs in an earlier generation of the same launch
      if (best < 0 || f < bestfit) { best = id; bestfit = f; }
    }
    winner = best; banned = best;
  }
  return winner;
}

}  // namespace ahx
#include "ahx_ga_res.cuh"
namespace ahx {

template <typename Coo>
__global__ void __launch_bounds__(kBlock) ga_generation_kernel(GaDev g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long scan_scratch[kWarps];
  __shared__ double red_scratch[kWarps];
  __shared__ int red_idx[kWarps];
  __shared__ int s_flag;
  __shared__ int s_best;
  __shared__ double s_bestfit;
  double *x = reinterpret_cast<double *>(smem_raw);
  int *tig = reinterpret_cast<int *>(smem_raw + sizeof(double) * static_cast<size_t>(g.N));
  const int tid = threadIdx.x, o = blockIdx.x;
  GaState *st = g.st;
  if (*reinterpret_cast<volatile int *>(&st->stop)) return;
  const long long gen = *reinterpret_cast<volatile long long *>(&st->gen);
  const int cur = static_cast<int>(gen & 1);
  const int *popc = g.pop[cur]; int *popn = g.pop[cur ^ 1];
  const double *fitc = g.fit[cur]; double *fitn = g.fit[cur ^ 1];
  const Philox ph{g.key0, g.key1};
  const uint32_t glo = static_cast<uint32_t>(gen), ghi = static_cast<uint32_t>(gen >> 32);

  const int parent = tournament_pair(fitc, g.P, g.k, ph, glo, ghi, static_cast<uint32_t>(o >> 1), o & 1);
  const MutPlan mp = plan_mutation(ph, glo, ghi, static_cast<uint32_t>(o), g.mut_prob, g.L);

  const int *src = popc + static_cast<size_t>(parent) * g.L;
  int *dst = popn + static_cast<size_t>(o) * g.L;
  for (int i = tid; i < g.L; i += kBlock) {
    const int v = src[mut_source(mp.op, mp.p, mp.q, mp.dir, g.L, i)];
    tig[i] = v;
    dst[i] = v;
  }
  if (!mp.mutated) {
    if (tid == 0) fitn[o] = fitc[parent];
  } else {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = tid; i < g.N; i += kBlock) x[i] = qnan;
    __syncthreads();
    scatter_mids<1>(tig, g.L, g.sizes, x, 0, scan_scratch, tid);
    __syncthreads();
    double acc[1] = {0.0};
    accumulate_pairs<1, Coo>(x, reinterpret_cast<const typename Coo::Elem *>(g.coo), g.E, acc, tid);
    const double s = block_sum(acc[0], red_scratch, tid);
    if (tid == 0) {
      fitn[o] = 0.0 - s;
      atomicAdd(reinterpret_cast<unsigned long long *>(&st->evals), 1ULL);
    }
  }
  // ---- last CTA done: hall of fame + bookkeeping (evaluate.go:287-306)
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  double bf = __longlong_as_double(0x7ff0000000000000LL);  // +inf
  int bi = 0x7fffffff;
  for (int i = tid; i < g.P; i += kBlock) {
    const double f = __ldcg(fitn + i);
    if (f < bf) { bf = f; bi = i; }   // ascending i per thread => lowest index on ties
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double of = __shfl_down_sync(0xffffffffu, bf, d);
    const int oi = __shfl_down_sync(0xffffffffu, bi, d);
    if (of < bf || (of == bf && oi < bi)) { bf = of; bi = oi; }
  }
  if ((tid & 31) == 0) { red_scratch[tid >> 5] = bf; red_idx[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kWarps; ++w)
      if (red_scratch[w] < bf || (red_scratch[w] == bf && red_idx[w] < bi)) { bf = red_scratch[w]; bi = red_idx[w]; }
    s_best = bi; s_bestfit = bf;
    s_flag = bf < st->hof_fit;        // updateHallOfFame keeps the best ever
  }
  __syncthreads();
  if (s_flag) {
    const int *best_row = popn + static_cast<size_t>(s_best) * g.L;
    for (int i = tid; i < g.L; i += kBlock) g.hof[i] = __ldcg(best_row + i);
  }
  if (tid == 0) {
    const long long ngen_now = gen + 1;
    if (s_flag) { st->hof_fit = s_bestfit; st->updated = ngen_now; }   // currentBest > *best
    st->best_idx = s_best;
    st->ticket = 0;
    if (ngen_now - st->updated > g.ngen || ngen_now >= g.max_gens) st->stop = 1;
    __threadfence();
    st->gen = ngen_now;
  }
}

// ---------------------------------------------------------------------------
// Two-kernel generation for the x32 evaluate path.
//   ga_plan_kernel  (1 CTA)  tournaments + mutation plans for all P offspring, the
//                   inherited fitness of clones, and the ORDERED list of mutated
//                   offspring (block scan, so the grouping of tours into evaluate
//                   CTAs -- and with it every bit of every score -- is a pure
//                   function of the seed, never of scheduling)
//   ga_apply_kernel CTAs [0, ceil(P/T)): T mutated offspring each -- clone+mutate
//                   straight into the evaluate prologue, write the row, score it;
//                   remaining CTAs copy the unmutated clones; the last CTA to
//                   finish updates the hall of fame and the generation state.
// ---------------------------------------------------------------------------
struct GaPlan {
  int *parent;    // [P]
  int *info;      // [P] op | dir << 8 | mutated << 16   (op 0xff = identity)
  int *p, *q;     // [P]
  int *mlist;     // [P] mutated offspring, ascending
  int *nmut;      // [1]
};

__global__ void __launch_bounds__(256) ga_plan_kernel(GaDev g, GaPlan pl, unsigned int *plan_ticket) {
  __shared__ int warp_tot[8];
  __shared__ int s_last;
  GaState *st = g.st;
  if (*reinterpret_cast<volatile int *>(&st->stop)) return;
  const long long gen = *reinterpret_cast<volatile long long *>(&st->gen);
  const int cur = static_cast<int>(gen & 1);
  const double *fitc = g.fit[cur]; double *fitn = g.fit[cur ^ 1];
  const Philox ph{g.key0, g.key1};
  const uint32_t glo = static_cast<uint32_t>(gen), ghi = static_cast<uint32_t>(gen >> 32);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // one offspring per thread, all CTAs in parallel
  const int o = blockIdx.x * blockDim.x + tid;
  if (o < g.P) {
    const int parent = tournament_pair(fitc, g.P, g.k, ph, glo, ghi, static_cast<uint32_t>(o >> 1), o & 1);
    const MutPlan mp = plan_mutation(ph, glo, ghi, static_cast<uint32_t>(o), g.mut_prob, g.L);
    pl.parent[o] = parent;
    pl.info[o] = (mp.op & 0xff) | (mp.dir << 8) | (mp.mutated << 16);
    pl.p[o] = mp.p; pl.q[o] = mp.q;
    if (!mp.mutated) fitn[o] = fitc[parent];
  }
  // the last CTA to finish builds the ordered list of mutated offspring
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(plan_ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int per = (g.P + 255) / 256;
  const int beg = min(g.P, tid * per), end = min(g.P, beg + per);
  int cnt = 0;
  for (int i = beg; i < end; ++i) cnt += (__ldcg(pl.info + i) >> 16) & 1;
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += n; }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < wid; ++w) base += warp_tot[w];
  int pos = base + incl - cnt;
  for (int i = beg; i < end; ++i)
    if ((__ldcg(pl.info + i) >> 16) & 1) pl.mlist[pos++] = i;
  if (tid == 255) { *pl.nmut = base + incl; *plan_ticket = 0; }
}

template <int T>
struct GaTours {
  const int *popc; int *popn;
  int L;
  int o[T], parent[T], op[T], p[T], q[T], dir[T];
  __device__ __forceinline__ bool live(int t) const { return o[t] >= 0; }
  __device__ __forceinline__ unsigned get(int t, int i) const {
    return static_cast<unsigned>(popc[static_cast<size_t>(parent[t]) * L + mut_source(op[t], p[t], q[t], dir[t], L, i)]);
  }
  __device__ __forceinline__ void put(int t, int i, unsigned c) const { popn[static_cast<size_t>(o[t]) * L + i] = static_cast<int>(c); }
};

template <int T, typename Coo>
__global__ void __launch_bounds__(kBlock, T == 4 ? 3 : 4)
ga_apply_kernel(GaDev g, GaPlan pl, int n_eval_ctas, const uint32_t *__restrict__ sizes32, int E_pad, int *__restrict__ errflag) {
  using Elem = typename Coo::Elem;
  extern __shared__ __align__(128) unsigned char smem_ga[];
  __shared__ long long scan_scratch[kWarps * T];
  __shared__ double red_scratch[kWarps];
  __shared__ int red_idx[kWarps];
  __shared__ __align__(8) uint64_t bars[2 * kStages];
  __shared__ int s_flag, s_best;
  __shared__ double s_bestfit;
  const int tid = threadIdx.x;
  GaState *st = g.st;
  if (*reinterpret_cast<volatile int *>(&st->stop)) return;
  const long long gen = *reinterpret_cast<volatile long long *>(&st->gen);
  const int cur = static_cast<int>(gen & 1);
  const int *popc = g.pop[cur]; int *popn = g.pop[cur ^ 1];
  double *fitn = g.fit[cur ^ 1];
  const int nmut = *pl.nmut;
  if (static_cast<int>(blockIdx.x) < n_eval_ctas) {
    const int base = blockIdx.x * T;
    if (base < nmut) {
      Elem *tiles = reinterpret_cast<Elem *>(smem_ga);
      Elem *queues = tiles + kStages * kTile;
      uint32_t *X = reinterpret_cast<uint32_t *>(queues + kWarps * kQueue);
      uint32_t *S = X + static_cast<size_t>(g.N) * T;
      const Elem *coo = reinterpret_cast<const Elem *>(g.coo);
      if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(&bars[s], 1); mbar_init(&bars[kStages + s], kWarps); }
        mbar_init_fence();
        prefetch_pairs<Elem>(coo, E_pad, tiles, bars);
      }
      for (int i = tid; i < g.N * T; i += kBlock) X[i] = kInactiveX;
      for (int i = tid; i < g.N; i += kBlock) S[i] = __ldg(sizes32 + i);
      GaTours<T> src;
      src.popc = popc; src.popn = popn; src.L = g.L;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int o = base + t < nmut ? pl.mlist[base + t] : -1;
        src.o[t] = o;
        if (o >= 0) {
          const int info = pl.info[o];
          src.parent[t] = pl.parent[o]; src.op[t] = info & 0xff; src.dir[t] = (info >> 8) & 1; src.p[t] = pl.p[o]; src.q[t] = pl.q[o];
        } else { src.parent[t] = 0; src.op[t] = 0xff; src.dir[t] = 0; src.p[t] = 0; src.q[t] = 0; }
      }
      __syncthreads();
      scatter_tours_x32<T>(src, g.L, g.N, S, X, scan_scratch, errflag, tid);
      __syncthreads();
      PairSink<T, Coo> sink;
      sink.init(queues + (tid >> 5) * kQueue, X, tid & 31);
      stream_pairs_x32<T, Coo>(coo, E_pad, tiles, bars, sink, tid);
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const double s = block_sum(sink.acc[t], red_scratch, tid);
        if (tid == 0 && src.o[t] >= 0) fitn[src.o[t]] = 0.0 - s;
      }
    }
  } else {
    // clones: offspring row = parent row (eaopt Clone), fitness was inherited by the plan kernel
    const int ncopy = gridDim.x - n_eval_ctas;
    const bool vec = (g.L & 3) == 0;
    for (int o = blockIdx.x - n_eval_ctas; o < g.P; o += ncopy) {
      if ((pl.info[o] >> 16) & 1) continue;
      const int *src = popc + static_cast<size_t>(pl.parent[o]) * g.L;
      int *dst = popn + static_cast<size_t>(o) * g.L;
      if (vec) {
        const int4 *s4 = reinterpret_cast<const int4 *>(src); int4 *d4 = reinterpret_cast<int4 *>(dst);
        for (int i = tid; i < g.L / 4; i += kBlock) d4[i] = s4[i];
      } else {
        for (int i = tid; i < g.L; i += kBlock) dst[i] = src[i];
      }
    }
  }
  // ---- last CTA done: hall of fame + bookkeeping (evaluate.go:287-306)
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  double bf = __longlong_as_double(0x7ff0000000000000LL);
  int bi = 0x7fffffff;
  for (int i = tid; i < g.P; i += kBlock) {
    const double f = __ldcg(fitn + i);
    if (f < bf) { bf = f; bi = i; }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double of = __shfl_down_sync(0xffffffffu, bf, d);
    const int oi = __shfl_down_sync(0xffffffffu, bi, d);
    if (of < bf || (of == bf && oi < bi)) { bf = of; bi = oi; }
  }
  if ((tid & 31) == 0) { red_scratch[tid >> 5] = bf; red_idx[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kWarps; ++w)
      if (red_scratch[w] < bf || (red_scratch[w] == bf && red_idx[w] < bi)) { bf = red_scratch[w]; bi = red_idx[w]; }
    s_best = bi; s_bestfit = bf;
    s_flag = bf < st->hof_fit;
  }
  __syncthreads();
  if (s_flag) {
    const int *best_row = popn + static_cast<size_t>(s_best) * g.L;
    for (int i = tid; i < g.L; i += kBlock) g.hof[i] = __ldcg(best_row + i);
  }
  if (tid == 0) {
    const long long ngen_now = gen + 1;
    if (s_flag) { st->hof_fit = s_bestfit; st->updated = ngen_now; }
    st->best_idx = s_best;
    st->ticket = 0;
    st->evals += nmut;
    if (ngen_now - st->updated > g.ngen || ngen_now >= g.max_gens) st->stop = 1;
    __threadfence();
    st->gen = ngen_now;
  }
}

// ---------------------------------------------------------------------------
// Generation for groups too large for any kernel that keeps per-contig arrays in shared memory
// (more than ~26 000 contigs): ga_plan_kernel, then ga_offspring_kernel (Tour.Clone + Tour.Mutate
// row by row in global memory), then the evaluate launcher on the mutated rows (it falls back to
// coordinates in global memory by itself, eval_m_sparse_kernel<XGLOBAL>), then ga_finish_kernel
// (hall of fame + bookkeeping).  One small read-back per generation tells the host how many rows
// were mutated; a generation of such a group takes milliseconds, the read-back microseconds.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ga_offspring_kernel(GaDev g, GaPlan pl) {
  GaState *st = g.st;
  if (*reinterpret_cast<volatile int *>(&st->stop)) return;
  const long long gen = *reinterpret_cast<volatile long long *>(&st->gen);
  const int cur = static_cast<int>(gen & 1), o = blockIdx.x;
  const int info = pl.info[o], op = (info >> 16) & 1 ? static_cast<int>(static_cast<signed char>(info & 0xff)) : -1;
  const int dir = (info >> 8) & 1, p = pl.p[o], q = pl.q[o];
  const int *src = g.pop[cur] + static_cast<size_t>(pl.parent[o]) * g.L;
  int *dst = g.pop[cur ^ 1] + static_cast<size_t>(o) * g.L;
  for (int i = threadIdx.x; i < g.L; i += blockDim.x) dst[i] = src[mut_source(op, p, q, dir, g.L, i)];
}

__global__ void __launch_bounds__(1024) ga_finish_kernel(GaDev g, const int *__restrict__ nmut) {
  __shared__ double sf[32];
  __shared__ int si[32];
  __shared__ int s_best, s_flag;
  __shared__ double s_bestfit;
  GaState *st = g.st;
  if (*reinterpret_cast<volatile int *>(&st->stop)) return;
  const long long gen = st->gen;
  const int cur = static_cast<int>(gen & 1), tid = threadIdx.x;
  const double *fitn = g.fit[cur ^ 1];
  const int *popn = g.pop[cur ^ 1];
  double bf = __longlong_as_double(0x7ff0000000000000LL); int bi = 0x7fffffff;
  for (int i = tid; i < g.P; i += blockDim.x) { const double f = fitn[i]; if (f < bf) { bf = f; bi = i; } }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double of = __shfl_down_sync(0xffffffffu, bf, d);
    const int oi = __shfl_down_sync(0xffffffffu, bi, d);
    if (of < bf || (of == bf && oi < bi)) { bf = of; bi = oi; }
  }
  if ((tid & 31) == 0) { sf[tid >> 5] = bf; si[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
      if (sf[w] < bf || (sf[w] == bf && si[w] < bi)) { bf = sf[w]; bi = si[w]; }
    s_best = bi; s_bestfit = bf; s_flag = bf < st->hof_fit;        // updateHallOfFame keeps the best ever
  }
  __syncthreads();
  if (s_flag) {
    const int *best_row = popn + static_cast<size_t>(s_best) * g.L;
    for (int i = tid; i < g.L; i += blockDim.x) g.hof[i] = best_row[i];
  }
  __syncthreads();
  if (tid == 0) {
    const long long ngen_now = gen + 1;
    if (s_flag) { st->hof_fit = s_bestfit; st->updated = ngen_now; }
    st->best_idx = s_best;
    st->evals += *nmut;
    if (ngen_now - st->updated > g.ngen || ngen_now >= g.max_gens) st->stop = 1;   // evaluate.go:304-306
    __threadfence();
    st->gen = ngen_now;
  }
}

__global__ void ga_init_kernel(GaDev g, const int *__restrict__ init, const double *__restrict__ f0) {
  const size_t total = static_cast<size_t>(g.P) * g.L;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    g.pop[0][i] = init[i % g.L];   // MakeTour = r.Tour.Clone(): PopSize copies of one tour (evaluate.go:259-262)
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < static_cast<size_t>(g.P); i += static_cast<size_t>(gridDim.x) * blockDim.x)
    g.fit[0][i] = f0[0];
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < g.L; i += blockDim.x) g.hof[i] = init[i];
    if (threadIdx.x == 0) {
      GaState s{};
      s.gen = 0; s.updated = 0; s.evals = 1; s.hof_fit = f0[0];
      s.stop = (g.max_gens <= 0 || 0 - s.updated > g.ngen) ? 1 : 0;
      s.ticket = 0; s.best_idx = 0;
      *g.st = s;
    }
  }
}

// n rounds of arg-min (want_max = 0) or arg-max over fit[0..P) without replacement.
__global__ void __launch_bounds__(1024) select_extreme_kernel(const double *__restrict__ fit, int P, int n, int want_max,
                                                              unsigned char *taken, int *out_idx, double *out_fit) {
  __shared__ double sf[32];
  __shared__ int si[32];
  const int tid = threadIdx.x;
  for (int r = 0; r < n; ++r) {
    double bf = 0.0; int bi = 0x7fffffff;
    for (int i = tid; i < P; i += blockDim.x) {
      if (taken[i]) continue;
      const double f = want_max ? -fit[i] : fit[i];
      if (bi == 0x7fffffff || f < bf) { bf = f; bi = i; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double of = __shfl_down_sync(0xffffffffu, bf, d);
      const int oi = __shfl_down_sync(0xffffffffu, bi, d);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || of < bf || (of == bf && oi < bi))) { bf = of; bi = oi; }
    }
    if ((tid & 31) == 0) { sf[tid >> 5] = bf; si[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
        if (si[w] != 0x7fffffff && (bi == 0x7fffffff || sf[w] < bf || (sf[w] == bf && si[w] < bi))) { bf = sf[w]; bi = si[w]; }
      out_idx[r] = bi;
      out_fit[r] = want_max ? -bf : bf;
      taken[bi] = 1;
    }
    __syncthreads();
  }
}

__global__ void gather_rows_kernel(const int *__restrict__ pop, const int *__restrict__ idx, int L, int *__restrict__ out) {
  const int *src = pop + static_cast<size_t>(idx[blockIdx.x]) * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) out[static_cast<size_t>(blockIdx.x) * L + i] = src[i];
}
__global__ void scatter_rows_kernel(int *__restrict__ pop, const int *__restrict__ idx, int L, const int *__restrict__ in) {
  int *dst = pop + static_cast<size_t>(idx[blockIdx.x]) * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) dst[i] = in[static_cast<size_t>(blockIdx.x) * L + i];
}
// After migrants were written and scored: refresh the hall of fame from fit[0..P).
__global__ void __launch_bounds__(1024) hof_refresh_kernel(const double *__restrict__ fit, const int *__restrict__ pop, const int *__restrict__ rowof,
                                                           int P, int L, int *hof, GaState *st, int n_new_evals, long long max_gens) {
  __shared__ double sf[32];
  __shared__ int si[32];
  __shared__ int s_best, s_improved;
  const int tid = threadIdx.x;
  double bf = __longlong_as_double(0x7ff0000000000000LL); int bi = 0x7fffffff;
  for (int i = tid; i < P; i += blockDim.x) { const double f = fit[i]; if (f < bf) { bf = f; bi = i; } }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double of = __shfl_down_sync(0xffffffffu, bf, d);
    const int oi = __shfl_down_sync(0xffffffffu, bi, d);
    if (of < bf || (of == bf && oi < bi)) { bf = of; bi = oi; }
  }
  if ((tid & 31) == 0) { sf[tid >> 5] = bf; si[tid >> 5] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
      if (sf[w] < bf || (sf[w] == bf && si[w] < bi)) { bf = sf[w]; bi = si[w]; }
    s_best = bi; s_improved = bf < st->hof_fit;
    st->evals += n_new_evals;
    // an improvement restarts EarlyStop's patience (evaluate.go:304-306) -- but not a run that ended on max_gens
    if (s_improved) { st->hof_fit = bf; st->updated = st->gen; if (st->gen < max_gens) st->stop = 0; }
  }
  __syncthreads();
  if (s_improved) {
    const int row = rowof ? rowof[s_best] : s_best;
    for (int i = tid; i < L; i += blockDim.x) hof[i] = pop[static_cast<size_t>(row) * L + i];
  }
}

__global__ void mutate_kernel(const int *__restrict__ in, int L, int op, int p, int q, int dir, int *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) out[i] = in[mut_source(op, p, q, dir, L, i)];
}

// SplitMix-style whitening of (seed, island) into the Philox key
static Philox philox_key(uint64_t seed, int32_t island) {
  uint64_t z = seed + 0x9e3779b97f4a7c15ULL * (static_cast<uint64_t>(static_cast<uint32_t>(island)) + 1);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; z ^= z >> 31;
  return Philox{static_cast<uint32_t>(z), static_cast<uint32_t>(z >> 32)};
}

// ---- test hooks: the product kernels' operator code and random decisions on their own -------
// One mutation through remap_of + slot_source / slot_coord (what ga_persist_kernel runs).
__global__ void mutate_remap_kernel(const int *__restrict__ pt, const uint32_t *__restrict__ px, const uint32_t *__restrict__ s32, uint32_t total2,
                                    int L, int N, MutPlan mp, int *__restrict__ ct, uint32_t *__restrict__ cx) {
  __shared__ GaSlot s;
  if (threadIdx.x == 0) { s.o = 0; s.parent_row = 0; s.child_row = 0; remap_of(mp, L, pt, px, s32, total2, &s); }
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) ct[i] = pt[slot_source(s, i, L)];
  for (int c = threadIdx.x; c < N; c += blockDim.x) cx[c] = L < N ? slot_coord<true>(s, c, px[c]) : slot_coord<false>(s, c, px[c]);
}
// winners[2 * pair + which] = tournament_pair(...) for pairs 0 .. n_pairs-1 of generation gen
__global__ void tournaments_kernel(const double *__restrict__ fit, int P, int k, Philox ph, long long gen, int n_pairs, int *__restrict__ winners) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n_pairs)
    winners[i] = tournament_pair(fit, P, k, ph, static_cast<uint32_t>(gen), static_cast<uint32_t>(gen >> 32), static_cast<uint32_t>(i >> 1), i & 1);
}

static GaDev make_dev(ahx_ctx *ctx) {
  GaDev g{};
  const ahx_ga_params &p = ctx->ga_prm;
  g.P = p.npop; g.L = ctx->ga_L; g.N = ctx->N; g.E = static_cast<int>(ctx->E); g.k = p.tournament_k;
  g.mut_prob = p.mut_prob;
  const Philox key = philox_key(p.seed, p.island);
  g.key0 = key.k0; g.key1 = key.k1;
  g.ngen = p.ngen; g.max_gens = p.max_gens;
  g.sizes = ctx->d_sizes.p;
  g.coo = ctx->coo16 ? static_cast<const void *>(ctx->d_coo16.p) : static_cast<const void *>(ctx->d_coo32.p);
  g.pop[0] = ctx->d_pop[0].p; g.pop[1] = ctx->d_pop[1].p;
  g.fit[0] = ctx->d_fit[0].p; g.fit[1] = ctx->d_fit[1].p;
  g.hof = ctx->d_hof.p; g.st = ctx->d_state.p;
  return g;
}

static int ga_T(const ahx_ctx *ctx) {
  // tours per evaluate CTA: 4 when three CTAs fit an SM and a generation has enough mutants to fill the GPU
  const double expect = ctx->ga_prm.npop * ctx->ga_prm.mut_prob;
  for (int t : {4, 2})
    if ((eval_x32_smem_bytes(t, ctx->coo16, ctx->N, ctx->ga_L) + 1536) * 3 <= ctx->smem_optin + 2048 && expect / t >= ctx->sm_count) return t;
  return 1;
}

template <int T, typename Coo>
static cudaError_t launch_apply(ahx_ctx *ctx, const GaDev &g, const GaPlan &pl, cudaStream_t s) {
  auto kern = ga_apply_kernel<T, Coo>;
  const size_t smem = eval_x32_smem_bytes(T, ctx->coo16, ctx->N, ctx->ga_L);
  cudaError_t e = opt_in_smem(kern, ctx->smem_optin);
  if (e != cudaSuccess) return e;
  const int n_eval = (g.P + T - 1) / T;
  const int n_copy = std::min(g.P, ctx->sm_count * 4);
  kern<<<n_eval + n_copy, kBlock, smem, s>>>(g, pl, n_eval, ctx->d_sizes32.p, static_cast<int>(ctx->E_pad), ctx->d_errflag.p);
  return cudaGetLastError();
}

// ---- persistent resident GA (ahx_ga_res.cuh) ----------------------------------------
// Usable when the evaluate kernel's resident form is (32-bit coordinates, 16-bit contig
// indices, table + coordinates + GA scratch fit one SM's shared memory), the tour covers
// all contigs (L == N: distinct coordinates, see res_pair_terms) and the masks fit one
// word per thread.  T: tours per batch; G: CTAs (one per SM at most).
struct GaResCfg { int T = 0, G = 0, nbmax = 0; size_t smem = 0; bool stream = false; };

// X[c] = 2*cumSum + size = 2*mid of every contig of a full tour (evaluate.go:120-126), what the
// evaluate kernel's producers compute on the device; the GA keeps one such row per pool row.
static std::vector<uint32_t> tour_coordinates(const ahx_ctx *ctx, int32_t L, const int32_t *tour) {
  std::vector<uint32_t> x(ctx->N, kInactiveX);
  uint64_t cum = 0;
  for (int32_t i = 0; i < L; ++i) {
    const uint64_t sz = static_cast<uint64_t>(ctx->h_sizes[tour[i]]);
    x[tour[i]] = static_cast<uint32_t>(2 * cum + sz);
    cum += sz;
  }
  return x;
}

static GaResCfg ga_res_config(const ahx_ctx *ctx, int32_t P, int32_t L, double mut_prob) {
  GaResCfg c;
  if (const char *env = getenv("AHX_GA_RES")) if (!atoi(env)) return c;
  if (!ctx->x32_ok || !ctx->coo16 || ctx->E_res[0] <= 0 || L > ctx->N) return c;
  const double expect = std::max(1.0, P * mut_prob);
  const bool no_stream = getenv("AHX_EVAL_STREAM") && !atoi(getenv("AHX_EVAL_STREAM"));
  const bool only_stream = getenv("AHX_EVAL_STREAM") && atoi(getenv("AHX_EVAL_STREAM")) == 1;
  auto try_T = [&](int t) {
    // CTAs: enough for the expected mutants; the slot tables must hold the worst case (every offspring mutated) of a
    // CTA's share, so a large population with a tiny mutation rate takes more CTAs than its mutants need
    int G = static_cast<int>(std::min<double>(ctx->sm_count, std::max(1.0, std::ceil(1.2 * expect / t))));
    for (;; G = std::min(ctx->sm_count, 2 * G)) {
      const int nbmax = ((P + t - 1) / t + G - 1) / G;
      for (int stream = 0; stream < 2; ++stream) {   // table resident in shared memory if it fits, streamed otherwise
        if (stream ? no_stream : only_stream) continue;
        if (!stream && static_cast<uint64_t>(ctx->N) * 4u * t > 65535u) continue;
        const size_t smem = res_smem_bytes(t, ctx->N, ctx->E_res[res_ti(t)], stream != 0) + ga_res_extra_bytes(P, t, nbmax);
        if (smem + 1024 > ctx->smem_optin) continue;
        c.T = t; c.G = G; c.nbmax = nbmax; c.smem = smem; c.stream = stream != 0;
        return true;
      }
      if (G >= ctx->sm_count) return false;
    }
  };
  int want = expect >= 4.0 * ctx->sm_count ? 4 : (expect >= 2.0 * ctx->sm_count ? 2 : 1);
  if (const char *env = getenv("AHX_GA_T")) { const int t = atoi(env); if (t == 1 || t == 2 || t == 4) want = t; }
  for (int t = want; t >= 1; t >>= 1)
    if (try_T(t)) return c;
  return c;
}

__global__ void ga_res_init_kernel(GaRes q, const int *__restrict__ init, const double *__restrict__ f0) {
  // MakeTour = r.Tour.Clone(): PopSize copies of one tour (evaluate.go:259-262) = P references to row 0
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q.P; i += gridDim.x * blockDim.x) { q.rowof[0][i] = 0; q.fit[0][i] = f0[0]; }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * ga_wr_pad(q.P); i += gridDim.x * blockDim.x)
    q.used[i] = (i == 0) ? 1u : 0u;   // generation 0 references row 0 only
  if (blockIdx.x == 0 && threadIdx.x < 3) q.gmin[threadIdx.x] = make_ulonglong2(~0ULL, ~0ULL);
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < q.L; i += blockDim.x) { q.pool[i] = init[i]; q.hof[i] = init[i]; }
    if (threadIdx.x == 0) {
      GaState s{};
      s.gen = 0; s.updated = 0; s.evals = 1; s.hof_fit = f0[0];
      s.stop = (q.max_gens <= 0 || 0 - s.updated > q.ngen) ? 1 : 0;
      s.ticket = 0; s.best_idx = 0;
      *q.st = s;
      q.isl->g_best = f0[0]; q.isl->g_updated = 0; q.isl->epochs = 0; q.isl->err = 0;
      q.isl->x_hof_fit = f0[0]; q.isl->x_updated = 0; q.isl->x_stop = s.stop;
    }
  }
}

static GaRes make_res(ahx_ctx *ctx) {
  const GaDev d = make_dev(ctx);
  GaRes q{};
  q.P = d.P; q.L = d.L; q.k = d.k; q.nbmax = ctx->ga_nbmax;
  q.mut_prob = d.mut_prob; q.key0 = d.key0; q.key1 = d.key1; q.ngen = d.ngen; q.max_gens = d.max_gens;
  q.xpool = ctx->d_xpool.p; q.sizes32 = ctx->d_sizes32.p;
  { long long tot = 0; for (int32_t c : ctx->ga_active_set) tot += ctx->h_sizes[c]; q.total2 = static_cast<uint32_t>(2 * tot); }
  q.pool = ctx->d_pop[0].p; q.rowof[0] = ctx->d_rowof[0].p; q.rowof[1] = ctx->d_rowof[1].p;
  q.fit[0] = ctx->d_fit[0].p; q.fit[1] = ctx->d_fit[1].p; q.hof = ctx->d_hof.p; q.st = ctx->d_state.p;
  q.gbar = ctx->d_gbar.p;
  q.used = ctx->d_used.p; q.gmin = reinterpret_cast<ulonglong2 *>(ctx->d_gmin.p); q.gmask = ctx->d_gmask.p;
  const ahx_ga_params &p = ctx->ga_prm;
  q.n_islands = p.n_islands > 1 ? p.n_islands : 1; q.island = p.island; q.migrate_every = p.migrate_every; q.n_elite = p.n_elite;
  q.isl = ctx->d_isl.p;
  return q;
}

template <int T, bool STREAM, bool SUB>
static cudaError_t launch_persist_ts(ahx_ctx *ctx, long long n_gens, cudaStream_t s) {
  auto kern = ga_persist_kernel<T, STREAM, SUB>;
  cudaError_t e = opt_in_smem(kern, ctx->smem_optin);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ctx->d_gbar.p, 0, sizeof(unsigned int), s)) != cudaSuccess) return e;
  const bool scaled = static_cast<uint64_t>(ctx->N) * 4u * T <= 65535u;   // the table of byte offsets exists (ahx_load)
  ResArgs g{(STREAM && !scaled ? ctx->d_rraw : ctx->d_rpairs)[res_ti(T)].p, ctx->d_rnl8[res_ti(T)].p, ctx->d_sizes32.p, static_cast<int>(ctx->E_res[res_ti(T)]), ctx->N, ctx->ga_L, ctx->d_errflag.p, 0, scaled ? 1 : 0};
  GaRes q = make_res(ctx);
  void *args[] = {&g, &q, &n_gens};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kern), dim3(ctx->ga_G), dim3(kRThreads), args, ctx->ga_smem, s);
}

template <int T, bool STREAM>
static cudaError_t launch_persist_t(ahx_ctx *ctx, long long n_gens, cudaStream_t s) {
  return ctx->ga_L < ctx->N ? launch_persist_ts<T, STREAM, true>(ctx, n_gens, s) : launch_persist_ts<T, STREAM, false>(ctx, n_gens, s);
}

static int launch_persist(ahx_ctx *ctx, long long n_gens, cudaStream_t s) {
  cudaError_t e;
  if (ctx->ga_stream) {
    switch (ctx->ga_T) {
      case 4: e = launch_persist_t<4, true>(ctx, n_gens, s); break;
      case 2: e = launch_persist_t<2, true>(ctx, n_gens, s); break;
      default: e = launch_persist_t<1, true>(ctx, n_gens, s); break;
    }
  } else {
    switch (ctx->ga_T) {
      case 4: e = launch_persist_t<4, false>(ctx, n_gens, s); break;
      case 2: e = launch_persist_t<2, false>(ctx, n_gens, s); break;
      default: e = launch_persist_t<1, false>(ctx, n_gens, s); break;
    }
  }
  if (e != cudaSuccess) return cuda_fail(ctx, e, "ga_persist_kernel launch");
  ctx->launches++;
  return AHX_OK;
}

// rows[i] = rowof[idx[i]]: physical rows of the selected individuals
__global__ void rows_of_kernel(const int *__restrict__ rowof, const int *__restrict__ idx, int n, int *__restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rows[i] = rowof[idx[i]];
}
__global__ void scatter_fit_kernel(const double *__restrict__ by_row, const int *__restrict__ rows, const int *__restrict__ idx, int n, double *__restrict__ fit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fit[idx[i]] = by_row[rows[i]];
}
__global__ void set_rows_kernel(int *__restrict__ rowof, const int *__restrict__ idx, const int *__restrict__ rows, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rowof[idx[i]] = rows[i];
}

static int read_state(ahx_ctx *ctx, GaState *hs);

static int launch_generation(ahx_ctx *ctx, const GaDev &g, cudaStream_t s) {
  if (ctx->ga_global) {
    GaPlan pl{ctx->d_plan.p, ctx->d_plan.p + g.P, ctx->d_plan.p + 2 * static_cast<size_t>(g.P), ctx->d_plan.p + 3 * static_cast<size_t>(g.P),
              ctx->d_plan.p + 4 * static_cast<size_t>(g.P), ctx->d_plan.p + 5 * static_cast<size_t>(g.P)};
    GaState hs{};
    int rc = read_state(ctx, &hs);
    if (rc != AHX_OK || hs.stop) return rc;
    const int cur = static_cast<int>(hs.gen & 1);
    ga_plan_kernel<<<(g.P + 255) / 256, 256, 0, s>>>(g, pl, reinterpret_cast<unsigned int *>(ctx->d_plan.p + 5 * static_cast<size_t>(g.P) + 1));
    ga_offspring_kernel<<<g.P, 256, 0, s>>>(g, pl);
    ctx->launches += 2;
    AHX_CUDA(ctx, cudaGetLastError());
    int32_t *nmut_pin = reinterpret_cast<int32_t *>(static_cast<unsigned char *>(ctx->h_pin) + 2048);
    AHX_CUDA(ctx, cudaMemcpyAsync(nmut_pin, pl.nmut, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    AHX_CUDA(ctx, cudaStreamSynchronize(s));
    if ((rc = launch_eval_m_sparse(ctx, ctx->d_pop[cur ^ 1].p, nullptr, pl.mlist, *nmut_pin, g.L, ctx->d_fit[cur ^ 1].p, s)) != AHX_OK) return rc;
    ga_finish_kernel<<<1, 1024, 0, s>>>(g, pl.nmut);
    ctx->launches++;
    AHX_CUDA(ctx, cudaGetLastError());
    return AHX_OK;
  }
  if (ctx->ga_x32) {
    GaPlan pl{ctx->d_plan.p, ctx->d_plan.p + g.P, ctx->d_plan.p + 2 * static_cast<size_t>(g.P), ctx->d_plan.p + 3 * static_cast<size_t>(g.P),
              ctx->d_plan.p + 4 * static_cast<size_t>(g.P), ctx->d_plan.p + 5 * static_cast<size_t>(g.P)};
    GaDev g2 = g;
    g2.coo = ctx->coo16 ? static_cast<const void *>(ctx->d_coo16o.p) : static_cast<const void *>(ctx->d_coo32o.p);
    ga_plan_kernel<<<(g.P + 255) / 256, 256, 0, s>>>(g2, pl, reinterpret_cast<unsigned int *>(ctx->d_plan.p + 5 * static_cast<size_t>(g.P) + 1));
    AHX_CUDA(ctx, cudaGetLastError());
    cudaError_t e;
    switch (ctx->ga_T) {
      case 4: e = ctx->coo16 ? launch_apply<4, Coo16>(ctx, g2, pl, s) : launch_apply<4, Coo32>(ctx, g2, pl, s); break;
      case 2: e = ctx->coo16 ? launch_apply<2, Coo16>(ctx, g2, pl, s) : launch_apply<2, Coo32>(ctx, g2, pl, s); break;
      default: e = ctx->coo16 ? launch_apply<1, Coo16>(ctx, g2, pl, s) : launch_apply<1, Coo32>(ctx, g2, pl, s); break;
    }
    AHX_CUDA(ctx, e);
    ctx->launches += 2;
    return AHX_OK;
  }
  const size_t smem = eval_m_smem_bytes(1, ctx->N, ctx->ga_L);
  if (ctx->coo16) ga_generation_kernel<Coo16><<<g.P, kBlock, smem, s>>>(g);
  else ga_generation_kernel<Coo32><<<g.P, kBlock, smem, s>>>(g);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  return AHX_OK;
}

static int read_state(ahx_ctx *ctx, GaState *hs) {
  static_assert(sizeof(GaState) + sizeof(GaIslands) <= 4096, "pinned scratch");
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, ctx->d_state.p, sizeof(GaState), cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(hs, ctx->h_pin, sizeof(GaState));
  return AHX_OK;
}
// island bookkeeping of the active run (after the stream has drained)
static int read_islands(ahx_ctx *ctx, GaIslands *h) {
  void *pin = static_cast<unsigned char *>(ctx->h_pin) + sizeof(GaState);
  AHX_CUDA(ctx, cudaMemcpyAsync(pin, ctx->d_isl.p, sizeof(GaIslands), cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(h, pin, sizeof(GaIslands));
  return AHX_OK;
}

static void fill_stats(const ahx_ctx *ctx, const GaState &hs, ahx_ga_stats *stats) {
  if (!stats) return;
  stats->best_score = -hs.hof_fit;
  stats->generations = hs.gen;
  stats->updated = hs.updated;
  stats->evaluations = hs.evals;
  stats->stopped = hs.stop;
  stats->reserved = 0;
  stats->device_ms = ctx->ga_ms;
}

}  // namespace ahx

using namespace ahx;

extern "C" {

void ahx_ga_default_params(ahx_ga_params *p) {
  if (!p) return;
  p->npop = 100;          // base.go:67 Npop
  p->ngen = 5000;         // base.go:69 Ngen
  p->max_gens = 1000000;  // evaluate.go:270
  p->mut_prob = 0.2;      // base.go:71 MutaProb
  p->seed = 42;           // base.go:65 Seed
  p->tournament_k = 3;    // evaluate.go:273
  p->log_every = 500;     // evaluate.go:294
  p->phase = 1;
  p->island = 0;
  p->n_islands = 1;       // evaluate.go:269 ga.NPops
  p->migrate_every = 50;
  p->n_elite = 2;
  p->reserved = 0;
}

int ahx_ga_begin(ahx_ctx *ctx, const ahx_ga_params *prm, int32_t L, const int32_t *init_tour) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (ctx->ga_active) return fail(ctx, AHX_ERR_STATE, "ahx_ga_begin: a GA run is already active");
  if (!prm) return fail(ctx, AHX_ERR_ARG, "ahx_ga_begin: null params");
  if ((rc = validate_tour(ctx, L, init_tour, "ahx_ga_begin")) != AHX_OK) return rc;
  if (L < 2) return fail(ctx, AHX_ERR_ARG, "ahx_ga_begin: tour needs at least 2 contigs (MutSplice calls Intn(L-1), evaluate.go:214)");
  if (prm->tournament_k < 1 || prm->tournament_k > kMaxContestants) return fail(ctx, AHX_ERR_ARG, "ahx_ga_begin: tournament_k must be in [1,8]");
  // eaopt: "Not enough individuals to select 2 with NContestants = k"
  if (prm->npop < 2 || prm->npop - 2 < prm->tournament_k - 1) return fail(ctx, AHX_ERR_ARG, "ahx_ga_begin: npop too small for 2 tournaments of k contestants");
  if (!(prm->mut_prob >= 0.0 && prm->mut_prob <= 1.0) || prm->ngen < 0 || prm->max_gens < 0 || prm->log_every < 0)
    return fail(ctx, AHX_ERR_ARG, "ahx_ga_begin: bad parameter");
  const int W = prm->n_islands > 1 ? prm->n_islands : 1;
  if (W > 1) {
    const long long m = static_cast<long long>(prm->n_elite) * (W - 1);
    if (W > kMaxIslands || prm->island < 0 || prm->island >= W || prm->migrate_every < 1 || prm->n_elite < 1 || m > kIslMaxMigrants || 2 * m > prm->npop)
      return fail(ctx, AHX_ERR_ARG, "ahx_ga_begin: bad island parameters (need 0 <= island < n_islands <= 32, migrate_every >= 1, 1 <= n_elite, n_elite * (n_islands - 1) <= min(64, npop / 2))");
  }
  ctx->ga_prm = *prm; ctx->ga_L = L;
  ctx->isl_connected = false;
  {
    const GaResCfg rc = ga_res_config(ctx, prm->npop, L, prm->mut_prob);
    ctx->ga_res = rc.T > 0;
    if (ctx->ga_res) { ctx->ga_T = rc.T; ctx->ga_G = rc.G; ctx->ga_nbmax = rc.nbmax; ctx->ga_smem = rc.smem; ctx->ga_stream = rc.stream; }
  }
  if (W > 1 && !ctx->ga_res)
    return fail(ctx, AHX_ERR_ARG, "ahx_ga_begin: the in-kernel island exchange needs the persistent GA kernel (full tour, 32-bit coordinates, 16-bit contig indices); "
                                  "use ahx_ga_get_elites / ahx_ga_put_migrants for a host-mediated exchange on this group");
  if (ctx->ga_res) {
    AHX_CUDA(ctx, ctx->d_isl.reserve(1));
    {
      GaIslands h{};
      long long ms = 20000;
      if (const char *env = getenv("AHX_ISLAND_TIMEOUT_MS")) { const long long v = atoll(env); if (v > 0) ms = v; }
      h.timeout_ns = ms * 1000000LL;
      if (W > 1) {
        const size_t mb = isl_mbox_bytes(W, prm->n_elite, L, ctx->N);
        AHX_CUDA(ctx, ctx->d_mbox.reserve(mb));
        AHX_CUDA(ctx, cudaMemsetAsync(ctx->d_mbox.p, 0, mb, ctx->stream));
        h.mbox[prm->island] = ctx->d_mbox.p;
      }
      AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_isl.p, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
      AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // h is stack-owned
    }
    const size_t PL2 = 2 * static_cast<size_t>(prm->npop) * L;
    AHX_CUDA(ctx, ctx->d_pop[0].reserve(PL2));
    for (int i = 0; i < 2; ++i) { AHX_CUDA(ctx, ctx->d_fit[i].reserve(prm->npop)); AHX_CUDA(ctx, ctx->d_rowof[i].reserve(prm->npop)); }
    AHX_CUDA(ctx, ctx->d_hof.reserve(L)); AHX_CUDA(ctx, ctx->d_state.reserve(1)); AHX_CUDA(ctx, ctx->d_gbar.reserve(2));
    AHX_CUDA(ctx, ctx->d_used.reserve(3 * static_cast<size_t>(ga_wr_pad(prm->npop)))); AHX_CUDA(ctx, ctx->d_gmin.reserve(6)); AHX_CUDA(ctx, ctx->d_gmask.reserve(3 * static_cast<size_t>(ga_wm_pad(prm->npop))));
    AHX_CUDA(ctx, ctx->d_tours.reserve(L)); AHX_CUDA(ctx, ctx->d_scores.reserve(1));
    ctx->ga_ms = 0.0; ctx->ga_x32 = false; ctx->ga_global = false;
    ctx->ga_active_set.assign(init_tour, init_tour + L);
    std::sort(ctx->ga_active_set.begin(), ctx->ga_active_set.end());
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_tours.p, init_tour, sizeof(int32_t) * L, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_eval_m_sparse(ctx, ctx->d_tours.p, nullptr, nullptr, 1, L, ctx->d_scores.p, ctx->stream)) != AHX_OK) return rc;
    {
      AHX_CUDA(ctx, ctx->d_xpool.reserve(2 * static_cast<size_t>(prm->npop) * ctx->N));
      const std::vector<uint32_t> x0 = tour_coordinates(ctx, L, init_tour);
      AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_xpool.p, x0.data(), sizeof(uint32_t) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
      AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // x0 is stack-owned
    }
    ga_res_init_kernel<<<ctx->sm_count, 256, 0, ctx->stream>>>(make_res(ctx), ctx->d_tours.p, ctx->d_scores.p);
    ctx->launches++;
    AHX_CUDA(ctx, cudaGetLastError());
    AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->ga_active = true;
    return AHX_OK;
  }
  ctx->ga_x32 = use_x32(ctx, L);
  ctx->ga_T = ctx->ga_x32 ? ga_T(ctx) : 1;
  if (const char *env = getenv("AHX_GA_T")) { const int t = atoi(env); if (ctx->ga_x32 && (t == 1 || t == 2 || t == 4)) ctx->ga_T = t; }
  const size_t smem = eval_m_smem_bytes(1, ctx->N, L);
  // neither fallback kernel can keep its per-contig arrays in shared memory (more than ~26 000 contigs): rows and
  // coordinates stay in global memory (ga_offspring_kernel + the evaluate launcher + ga_finish_kernel)
  ctx->ga_global = (!ctx->ga_x32 && smem + 2048 > ctx->smem_optin) || (getenv("AHX_GA_GLOBAL") && atoi(getenv("AHX_GA_GLOBAL")));
  if (ctx->ga_global) ctx->ga_x32 = false;
  if (ctx->ga_x32 || ctx->ga_global) {
    AHX_CUDA(ctx, ctx->d_plan.reserve(5 * static_cast<size_t>(prm->npop) + 2));
    AHX_CUDA(ctx, cudaMemsetAsync(ctx->d_plan.p + 5 * static_cast<size_t>(prm->npop), 0, 2 * sizeof(int32_t), ctx->stream));
  }
  if (ctx->ga_x32 || ctx->ga_global) { /* attributes are set per launch */ }
  else if (ctx->coo16) AHX_CUDA(ctx, opt_in_smem(ga_generation_kernel<Coo16>, ctx->smem_optin));
  else AHX_CUDA(ctx, opt_in_smem(ga_generation_kernel<Coo32>, ctx->smem_optin));
  ctx->ga_ms = 0.0;
  const size_t PL = static_cast<size_t>(prm->npop) * L;
  for (int i = 0; i < 2; ++i) { AHX_CUDA(ctx, ctx->d_pop[i].reserve(PL)); AHX_CUDA(ctx, ctx->d_fit[i].reserve(prm->npop)); }
  AHX_CUDA(ctx, ctx->d_hof.reserve(L)); AHX_CUDA(ctx, ctx->d_state.reserve(1));
  AHX_CUDA(ctx, ctx->d_tours.reserve(L)); AHX_CUDA(ctx, ctx->d_scores.reserve(1));
  ctx->ga_active_set.assign(init_tour, init_tour + L);
  std::sort(ctx->ga_active_set.begin(), ctx->ga_active_set.end());
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_tours.p, init_tour, sizeof(int32_t) * L, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = launch_eval_m_sparse(ctx, ctx->d_tours.p, nullptr, nullptr, 1, L, ctx->d_scores.p, ctx->stream)) != AHX_OK) return rc;
  const GaDev g = make_dev(ctx);
  ga_init_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(g, ctx->d_tours.p, ctx->d_scores.p);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->ga_active = true;
  return AHX_OK;
}

// One launch of the persistent kernel for up to `batch` generations, asynchronous; persist_finish waits for it,
// reads the state back and accounts the time.  Split so that ahx_ga_run_islands can have all islands' kernels
// in flight before it waits for any of them (they wait for each other at the exchange points).
static int persist_start(ahx_ctx *ctx, int64_t batch) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  if ((rc = launch_persist(ctx, batch, ctx->stream)) != AHX_OK) return rc;
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  return AHX_OK;
}
static int persist_finish(ahx_ctx *ctx, GaState *hs) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if ((rc = read_state(ctx, hs)) != AHX_OK) return rc;
  float ms = 0.f;
  AHX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  ctx->ga_ms += ms;
  if (ctx->ga_prm.n_islands > 1) {
    GaIslands h{};
    if ((rc = read_islands(ctx, &h)) != AHX_OK) return rc;
    if (h.err) return fail(ctx, AHX_ERR_STATE, "island exchange timed out: a peer island did not reach the exchange point (AHX_ISLAND_TIMEOUT_MS)");
  }
  return AHX_OK;
}
constexpr int64_t kPersistChunk = 8192;   // generations per launch; the kernel leaves its loop by itself when EarlyStop fires

int ahx_ga_advance(ahx_ctx *ctx, int64_t n_gens, ahx_ga_stats *stats) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!ctx->ga_active) return fail(ctx, AHX_ERR_STATE, "ahx_ga_advance: no active GA run");
  if (n_gens < 0) return fail(ctx, AHX_ERR_ARG, "ahx_ga_advance: negative n_gens");
  if (ctx->ga_prm.n_islands > 1 && !ctx->isl_connected && n_gens > 0)
    return fail(ctx, AHX_ERR_STATE, "ahx_ga_advance: islands are not connected (ahx_ga_island_connect / _connect_local)");
  const GaDev g = make_dev(ctx);
  GaState hs{};
  if ((rc = read_state(ctx, &hs)) != AHX_OK) return rc;
  int64_t left = n_gens;
  while (ctx->ga_res && left > 0 && !hs.stop) {
    const int64_t batch = std::min<int64_t>(left, kPersistChunk);
    if ((rc = persist_start(ctx, batch)) != AHX_OK) return rc;
    if ((rc = persist_finish(ctx, &hs)) != AHX_OK) return rc;
    left -= batch;
  }
  while (!ctx->ga_res && left > 0 && !hs.stop) {
    const int64_t batch = std::min<int64_t>(left, 256);   // one host sync per batch to poll the stop flag
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    for (int64_t i = 0; i < batch; ++i)
      if ((rc = launch_generation(ctx, g, ctx->stream)) != AHX_OK) return rc;
    AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    if ((rc = read_state(ctx, &hs)) != AHX_OK) return rc;
    float ms = 0.f;
    AHX_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    ctx->ga_ms += ms;
    left -= batch;
  }
  fill_stats(ctx, hs, stats);
  return AHX_OK;
}

// ---- islands: wiring (include/ahx.h "Islands") ------------------------------------------
static int island_table(ahx_ctx *ctx, void *const *mbox) {
  const int W = ctx->ga_prm.n_islands;
  AHX_CUDA(ctx, cudaMemcpy(ctx->d_isl.p, mbox, sizeof(void *) * W, cudaMemcpyHostToDevice));   // GaIslands::mbox is the first member
  ctx->isl_connected = true;
  return AHX_OK;
}
static int island_ready(ahx_ctx *ctx, const char *who) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!ctx->ga_active || ctx->ga_prm.n_islands <= 1 || !ctx->ga_res)
    return fail(ctx, AHX_ERR_STATE, std::string(who) + ": no active island run (ahx_ga_begin with n_islands > 1 first)");
  return AHX_OK;
}

int ahx_ga_island_export(ahx_ctx *ctx, void *handle) {
  int rc = island_ready(ctx, "ahx_ga_island_export");
  if (rc != AHX_OK) return rc;
  if (!handle) return fail(ctx, AHX_ERR_ARG, "ahx_ga_island_export: null handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == AHX_ISLAND_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  AHX_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->d_mbox.p));
  memcpy(handle, &h, sizeof h);
  return AHX_OK;
}

int ahx_ga_island_connect(ahx_ctx *ctx, const void *handles) {
  int rc = island_ready(ctx, "ahx_ga_island_connect");
  if (rc != AHX_OK) return rc;
  if (!handles) return fail(ctx, AHX_ERR_ARG, "ahx_ga_island_connect: null handles");
  const int W = ctx->ga_prm.n_islands;
  void *mbox[kMaxIslands] = {};
  for (int i = 0; i < W; ++i) {
    if (i == ctx->ga_prm.island) { mbox[i] = ctx->d_mbox.p; continue; }
    const std::string key(static_cast<const char *>(handles) + static_cast<size_t>(i) * AHX_ISLAND_HANDLE_BYTES, AHX_ISLAND_HANDLE_BYTES);
    for (auto &kv : ctx->isl_opened) if (kv.first == key) mbox[i] = kv.second;     // a mapping of this allocation is already open
    if (!mbox[i]) {
      cudaIpcMemHandle_t h;
      memcpy(&h, key.data(), sizeof h);
      AHX_CUDA(ctx, cudaIpcOpenMemHandle(&mbox[i], h, cudaIpcMemLazyEnablePeerAccess));
      ctx->isl_opened.emplace_back(key, mbox[i]);
    }
  }
  return island_table(ctx, mbox);
}

int ahx_ga_island_connect_local(ahx_ctx **ctxs, int32_t n) {
  if (!ctxs || n < 1) { set_global_error("ahx_ga_island_connect_local: bad argument"); return AHX_ERR_ARG; }
  for (int i = 0; i < n; ++i) {
    int rc = island_ready(ctxs[i], "ahx_ga_island_connect_local");
    if (rc != AHX_OK) return rc;
    if (ctxs[i]->ga_prm.n_islands != n || ctxs[i]->ga_prm.island != i) return fail(ctxs[i], AHX_ERR_ARG, "ahx_ga_island_connect_local: ctxs[i] must run island i of n");
  }
  // islands that share a GPU must be able to run side by side: their cooperative grids wait for each other
  for (int i = 0; i < n; ++i) {
    int ctas = 0;
    for (int j = 0; j < n; ++j) if (ctxs[j]->device == ctxs[i]->device) ctas += ctxs[j]->ga_G;
    if (ctas > ctxs[i]->sm_count) return fail(ctxs[i], AHX_ERR_ARG, "ahx_ga_island_connect_local: the islands placed on one GPU need more CTAs than it has SMs (they must be co-resident)");
  }
  for (int i = 0; i < n; ++i) {
    ahx_ctx *ctx = ctxs[i];
    AHX_CUDA(ctx, cudaSetDevice(ctx->device));
    void *mbox[kMaxIslands] = {};
    for (int j = 0; j < n; ++j) {
      mbox[j] = ctxs[j]->d_mbox.p;
      if (ctxs[j]->device == ctx->device) continue;
      int can = 0;
      AHX_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[j]->device));
      if (!can) return fail(ctx, AHX_ERR_CUDA, "ahx_ga_island_connect_local: no peer access between the islands' GPUs");
      const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess");
    }
    int rc = island_table(ctx, mbox);
    if (rc != AHX_OK) return rc;
  }
  return AHX_OK;
}

int64_t ahx_ga_island_epochs(ahx_ctx *ctx) {
  if (!ctx || !ctx->ga_active || !ctx->ga_res || enter(ctx, true) != AHX_OK) return 0;
  GaIslands h{};
  if (read_islands(ctx, &h) != AHX_OK) return 0;
  return static_cast<int64_t>(h.epochs);
}

int ahx_ga_best(ahx_ctx *ctx, int32_t *best_tour, double *best_score) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!ctx->ga_active) return fail(ctx, AHX_ERR_STATE, "ahx_ga_best: no active GA run");
  GaState hs{};
  if (best_tour) AHX_CUDA(ctx, cudaMemcpyAsync(best_tour, ctx->d_hof.p, sizeof(int32_t) * ctx->ga_L, cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = read_state(ctx, &hs)) != AHX_OK) return rc;
  if (best_score) *best_score = -hs.hof_fit;
  return AHX_OK;
}

int ahx_ga_get_elites(ahx_ctx *ctx, int32_t n, int32_t *tours, double *fitness) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!ctx->ga_active) return fail(ctx, AHX_ERR_STATE, "ahx_ga_get_elites: no active GA run");
  if (n < 0 || n > ctx->ga_prm.npop || (n > 0 && (!tours || !fitness))) return fail(ctx, AHX_ERR_ARG, "ahx_ga_get_elites: bad argument");
  if (n == 0) return AHX_OK;
  GaState hs{};
  if ((rc = read_state(ctx, &hs)) != AHX_OK) return rc;
  const int cur = static_cast<int>(hs.gen & 1), P = ctx->ga_prm.npop, L = ctx->ga_L;
  AHX_CUDA(ctx, ctx->d_taken.reserve(P)); AHX_CUDA(ctx, ctx->d_sel_idx.reserve(n)); AHX_CUDA(ctx, ctx->d_sel_fit.reserve(n));
  AHX_CUDA(ctx, ctx->d_sel_tours.reserve(static_cast<size_t>(n) * L));
  AHX_CUDA(ctx, cudaMemsetAsync(ctx->d_taken.p, 0, P, ctx->stream));
  select_extreme_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_fit[cur].p, P, n, 0, ctx->d_taken.p, ctx->d_sel_idx.p, ctx->d_sel_fit.p);
  if (ctx->ga_res) {
    AHX_CUDA(ctx, ctx->d_sel_rows.reserve(n));
    rows_of_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_rowof[cur].p, ctx->d_sel_idx.p, n, ctx->d_sel_rows.p);
    gather_rows_kernel<<<n, 256, 0, ctx->stream>>>(ctx->d_pop[0].p, ctx->d_sel_rows.p, L, ctx->d_sel_tours.p);
    ctx->launches++;
  } else {
    gather_rows_kernel<<<n, 256, 0, ctx->stream>>>(ctx->d_pop[cur].p, ctx->d_sel_idx.p, L, ctx->d_sel_tours.p);
  }
  ctx->launches += 2;
  AHX_CUDA(ctx, cudaGetLastError());
  AHX_CUDA(ctx, cudaMemcpyAsync(tours, ctx->d_sel_tours.p, sizeof(int32_t) * static_cast<size_t>(n) * L, cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaMemcpyAsync(fitness, ctx->d_sel_fit.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AHX_OK;
}

int ahx_ga_put_migrants(ahx_ctx *ctx, int32_t n, const int32_t *tours) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!ctx->ga_active) return fail(ctx, AHX_ERR_STATE, "ahx_ga_put_migrants: no active GA run");
  if (n < 0 || n > ctx->ga_prm.npop || (n > 0 && !tours)) return fail(ctx, AHX_ERR_ARG, "ahx_ga_put_migrants: bad argument");
  if (n == 0) return AHX_OK;
  const int P = ctx->ga_prm.npop, L = ctx->ga_L;
  {   // bit-exact validity: a migrant must be a permutation of the active set (one marking pass per migrant)
    std::vector<int32_t> stamp(ctx->N, -2);
    for (int32_t c : ctx->ga_active_set) stamp[c] = -1;
    for (int32_t m = 0; m < n; ++m)
      for (int32_t i = 0; i < L; ++i) {
        const int32_t c = tours[static_cast<size_t>(m) * L + i];
        if (c < 0 || c >= ctx->N || stamp[c] == -2 || stamp[c] == m)
          return fail(ctx, AHX_ERR_ARG, "ahx_ga_put_migrants: migrant is not a permutation of the active contigs");
        stamp[c] = m;
      }
  }
  GaState hs{};
  if ((rc = read_state(ctx, &hs)) != AHX_OK) return rc;
  const int cur = static_cast<int>(hs.gen & 1);
  AHX_CUDA(ctx, ctx->d_taken.reserve(P)); AHX_CUDA(ctx, ctx->d_sel_idx.reserve(n)); AHX_CUDA(ctx, ctx->d_sel_fit.reserve(n));
  AHX_CUDA(ctx, ctx->d_sel_tours.reserve(static_cast<size_t>(n) * L));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_sel_tours.p, tours, sizeof(int32_t) * static_cast<size_t>(n) * L, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, cudaMemsetAsync(ctx->d_taken.p, 0, P, ctx->stream));
  select_extreme_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_fit[cur].p, P, n, 1, ctx->d_taken.p, ctx->d_sel_idx.p, ctx->d_sel_fit.p);
  if (ctx->ga_res) {
    // individuals are references to pool rows and several may share one: a migrant takes a row that
    // the current generation does not reference, then the replaced individual points at it
    std::vector<int32_t> rowof(P), freerows;
    AHX_CUDA(ctx, cudaMemcpyAsync(rowof.data(), ctx->d_rowof[cur].p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
    AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<uint8_t> used(2 * static_cast<size_t>(P), 0);
    for (int32_t i = 0; i < P; ++i) used[rowof[i]] = 1;
    for (int32_t r = 0; r < 2 * P && static_cast<int32_t>(freerows.size()) < n; ++r) if (!used[r]) freerows.push_back(r);
    AHX_CUDA(ctx, ctx->d_sel_rows.reserve(n)); AHX_CUDA(ctx, ctx->d_row_fit.reserve(2 * static_cast<size_t>(P)));
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_sel_rows.p, freerows.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    {
      // exact reference map of the edited generation (the kernel allocates fresh rows from its complement)
      std::vector<int32_t> worst(n);
      AHX_CUDA(ctx, cudaMemcpyAsync(worst.data(), ctx->d_sel_idx.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
      AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      for (int32_t i = 0; i < n; ++i) rowof[worst[i]] = freerows[i];
      const size_t wrp = static_cast<size_t>(ga_wr_pad(P));
      std::vector<uint32_t> map(wrp, 0u);
      for (int32_t i = 0; i < P; ++i) map[rowof[i] >> 5] |= 1u << (rowof[i] & 31);
      AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_used.p + static_cast<size_t>(hs.gen % 3) * wrp, map.data(), sizeof(uint32_t) * wrp, cudaMemcpyHostToDevice, ctx->stream));
      AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    scatter_rows_kernel<<<n, 256, 0, ctx->stream>>>(ctx->d_pop[0].p, ctx->d_sel_rows.p, L, ctx->d_sel_tours.p);
    std::vector<std::vector<uint32_t>> xrows(n);   // the migrants' coordinate rows (kept alive until the stream is drained below)
    for (int32_t m = 0; m < n; ++m) {
      xrows[m] = tour_coordinates(ctx, L, tours + static_cast<size_t>(m) * L);
      AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_xpool.p + static_cast<size_t>(freerows[m]) * ctx->N, xrows[m].data(), sizeof(uint32_t) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    }
    set_rows_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_rowof[cur].p, ctx->d_sel_idx.p, ctx->d_sel_rows.p, n);
    ctx->launches += 3;
    AHX_CUDA(ctx, cudaGetLastError());
    if ((rc = launch_eval_m_sparse(ctx, ctx->d_pop[0].p, nullptr, ctx->d_sel_rows.p, n, L, ctx->d_row_fit.p, ctx->stream)) != AHX_OK) return rc;
    scatter_fit_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_row_fit.p, ctx->d_sel_rows.p, ctx->d_sel_idx.p, n, ctx->d_fit[cur].p);
    hof_refresh_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_fit[cur].p, ctx->d_pop[0].p, ctx->d_rowof[cur].p, P, L, ctx->d_hof.p, ctx->d_state.p, n, ctx->ga_prm.max_gens);
    ctx->launches += 2;
    AHX_CUDA(ctx, cudaGetLastError());
    AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // freerows is stack-owned
    return AHX_OK;
  }
  scatter_rows_kernel<<<n, 256, 0, ctx->stream>>>(ctx->d_pop[cur].p, ctx->d_sel_idx.p, L, ctx->d_sel_tours.p);
  ctx->launches += 2;
  AHX_CUDA(ctx, cudaGetLastError());
  if ((rc = launch_eval_m_sparse(ctx, ctx->d_pop[cur].p, nullptr, ctx->d_sel_idx.p, n, L, ctx->d_fit[cur].p, ctx->stream)) != AHX_OK) return rc;
  hof_refresh_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_fit[cur].p, ctx->d_pop[cur].p, nullptr, P, L, ctx->d_hof.p, ctx->d_state.p, n, ctx->ga_prm.max_gens);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return AHX_OK;
}

int ahx_ga_kernel(ahx_ctx *ctx) {
  if (!ctx || !ctx->ga_active) return -1;
  if (ctx->ga_res) return ctx->ga_stream ? 3 : 2;
  return ctx->ga_x32 ? 1 : 0;
}

int ahx_ga_selfcheck(ahx_ctx *ctx, int64_t *n_bad) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!ctx->ga_active || !n_bad) return fail(ctx, AHX_ERR_STATE, "ahx_ga_selfcheck: no active GA run");
  *n_bad = 0;
  GaState hs{};
  if ((rc = read_state(ctx, &hs)) != AHX_OK) return rc;
  const int cur = static_cast<int>(hs.gen & 1), P = ctx->ga_prm.npop, L = ctx->ga_L, N = ctx->N;
  std::vector<int32_t> rowof(P), tour(L);
  std::vector<uint32_t> xrow(N);
  std::vector<double> fit(P);
  AHX_CUDA(ctx, cudaMemcpy(fit.data(), ctx->d_fit[cur].p, sizeof(double) * P, cudaMemcpyDeviceToHost));
  if (ctx->ga_res) AHX_CUDA(ctx, cudaMemcpy(rowof.data(), ctx->d_rowof[cur].p, sizeof(int32_t) * P, cudaMemcpyDeviceToHost));
  // re-score every individual's stored tour with the evaluate kernel (population laid out row by row)
  AHX_CUDA(ctx, ctx->d_tours.reserve(static_cast<size_t>(P) * L)); AHX_CUDA(ctx, ctx->d_scores.reserve(P));
  for (int i = 0; i < P; ++i) {
    const int32_t *src = ctx->ga_res ? ctx->d_pop[0].p + static_cast<size_t>(rowof[i]) * L : ctx->d_pop[cur].p + static_cast<size_t>(i) * L;
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_tours.p + static_cast<size_t>(i) * L, src, sizeof(int32_t) * L, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if ((rc = launch_eval_m_sparse(ctx, ctx->d_tours.p, nullptr, nullptr, P, L, ctx->d_scores.p, ctx->stream)) != AHX_OK) return rc;
  std::vector<double> again(P);
  AHX_CUDA(ctx, cudaMemcpyAsync(again.data(), ctx->d_scores.p, sizeof(double) * P, cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = finish_checked(ctx, "ahx_ga_selfcheck")) != AHX_OK) return rc;
  for (int i = 0; i < P; ++i) {
    bool bad = !(std::fabs(again[i] - fit[i]) <= 1e-12 * std::fabs(again[i]));
    AHX_CUDA(ctx, cudaMemcpy(tour.data(), ctx->d_tours.p + static_cast<size_t>(i) * L, sizeof(int32_t) * L, cudaMemcpyDeviceToHost));
    std::vector<int32_t> sorted(tour);
    std::sort(sorted.begin(), sorted.end());
    if (sorted != ctx->ga_active_set) bad = true;                       // bit-exact validity: a permutation of the active contigs
    if (ctx->ga_res && !bad) {
      AHX_CUDA(ctx, cudaMemcpy(xrow.data(), ctx->d_xpool.p + static_cast<size_t>(rowof[i]) * N, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost));
      if (xrow != tour_coordinates(ctx, L, tour.data())) bad = true;    // stored coordinates == coordinates of the stored tour
    }
    if (bad) ++*n_bad;
  }
  return AHX_OK;
}

int ahx_ga_end(ahx_ctx *ctx) {
  int rc = enter(ctx, false);
  if (rc != AHX_OK) return rc;
  ctx->ga_active = false;
  ctx->isl_connected = false;
  return AHX_OK;
}

int ahx_ga_run(ahx_ctx *ctx, const ahx_ga_params *prm, int32_t L, const int32_t *init_tour, int32_t *best_tour,
               ahx_ga_stats *stats, ahx_ga_callback cb, void *user) {
  int rc = ahx_ga_begin(ctx, prm, L, init_tour);
  if (rc != AHX_OK) return rc;
  std::vector<int32_t> hof(L);
  ahx_ga_stats st{};
  auto report = [&](int64_t gen) -> int {
    if (!cb) return AHX_OK;
    double score = 0.0;
    int r = ahx_ga_best(ctx, hof.data(), &score);
    if (r != AHX_OK) return r;
    cb(user, prm->phase, gen, score, hof.data(), L);
    return AHX_OK;
  };
  rc = ahx_ga_advance(ctx, 0, &st);
  if (rc == AHX_OK) rc = report(0);   // ga.init's callback at generation 0
  const int64_t period = prm->log_every > 0 ? prm->log_every : 500;
  while (rc == AHX_OK && !st.stopped) {
    const int64_t next = (st.generations / period + 1) * period;
    rc = ahx_ga_advance(ctx, next - st.generations, &st);
    if (rc == AHX_OK && prm->log_every > 0 && st.generations % period == 0) rc = report(st.generations);
  }
  if (rc == AHX_OK) {
    double score = 0.0;
    rc = ahx_ga_best(ctx, best_tour, &score);   // r.Tour = ga.HallOfFame[0] (evaluate.go:313)
    if (stats) *stats = st;
  }
  ahx_ga_end(ctx);
  return rc;
}

int ahx_mutate_remap(ahx_ctx *ctx, int32_t L, const int32_t *tour_in, int32_t op, int32_t p, int32_t q, int32_t dir, int32_t *tour_out,
                     uint32_t *coords_out) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!tour_in || !tour_out || op < 0 || op > 3) return fail(ctx, AHX_ERR_ARG, "ahx_mutate_remap: bad argument");
  if (!ctx->x32_ok || L < 2) return fail(ctx, AHX_ERR_ARG, "ahx_mutate_remap: needs a group with 32-bit coordinates (what the persistent GA kernel runs on) and L >= 2");
  if ((rc = validate_tour(ctx, L, tour_in, "ahx_mutate_remap")) != AHX_OK) return rc;
  if (op == 1 ? (p < 1 || p >= L) : (p < 0 || q < p || q >= L)) return fail(ctx, AHX_ERR_ARG, "ahx_mutate_remap: positions out of range");
  const std::vector<uint32_t> x0 = tour_coordinates(ctx, L, tour_in);
  long long tot = 0; for (int32_t i = 0; i < L; ++i) tot += ctx->h_sizes[tour_in[i]];
  const int32_t N = ctx->N;
  DevBuf<int32_t> a, b; DevBuf<uint32_t> xa, xb;
  AHX_CUDA(ctx, a.reserve(L)); AHX_CUDA(ctx, b.reserve(L)); AHX_CUDA(ctx, xa.reserve(N)); AHX_CUDA(ctx, xb.reserve(N));
  AHX_CUDA(ctx, cudaMemcpyAsync(a.p, tour_in, sizeof(int32_t) * L, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, cudaMemcpyAsync(xa.p, x0.data(), sizeof(uint32_t) * N, cudaMemcpyHostToDevice, ctx->stream));
  const MutPlan mp{1, op, p, op == 1 ? 0 : q, dir};
  mutate_remap_kernel<<<1, 256, 0, ctx->stream>>>(a.p, xa.p, ctx->d_sizes32.p, static_cast<uint32_t>(2 * tot), L, ctx->N, mp, b.p, xb.p);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  AHX_CUDA(ctx, cudaMemcpyAsync(tour_out, b.p, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, ctx->stream));
  if (coords_out) AHX_CUDA(ctx, cudaMemcpyAsync(coords_out, xb.p, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  a.release(); b.release(); xa.release(); xb.release();
  return AHX_OK;
}

int ahx_test_plan_mutations(uint64_t seed, int32_t island, int64_t gen, int32_t n, double mut_prob, int32_t L, int32_t *out) {
  if (n < 0 || L < 2 || !out) { set_global_error("ahx_test_plan_mutations: bad argument"); return AHX_ERR_ARG; }
  const Philox ph = philox_key(seed, island);
  for (int32_t o = 0; o < n; ++o) {
    const MutPlan m = plan_mutation(ph, static_cast<uint32_t>(gen), static_cast<uint32_t>(gen >> 32), static_cast<uint32_t>(o), mut_prob, L);
    out[5 * o] = m.mutated; out[5 * o + 1] = m.op; out[5 * o + 2] = m.p; out[5 * o + 3] = m.q; out[5 * o + 4] = m.dir;
  }
  return AHX_OK;
}

int ahx_test_tournaments(ahx_ctx *ctx, uint64_t seed, int32_t island, int64_t gen, int32_t P, const double *fit, int32_t k, int32_t n_pairs,
                         int32_t *winners) {
  int rc = enter(ctx, false);
  if (rc != AHX_OK) return rc;
  if (!fit || !winners || n_pairs < 0 || k < 1 || k > kMaxContestants || P < 2 || P - 2 < k - 1) return fail(ctx, AHX_ERR_ARG, "ahx_test_tournaments: bad argument");
  if (n_pairs == 0) return AHX_OK;
  DevBuf<double> f; DevBuf<int32_t> w;
  AHX_CUDA(ctx, f.reserve(P)); AHX_CUDA(ctx, w.reserve(2 * static_cast<size_t>(n_pairs)));
  AHX_CUDA(ctx, cudaMemcpyAsync(f.p, fit, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
  tournaments_kernel<<<(2 * n_pairs + 255) / 256, 256, 0, ctx->stream>>>(f.p, P, k, philox_key(seed, island), gen, n_pairs, w.p);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  AHX_CUDA(ctx, cudaMemcpyAsync(winners, w.p, sizeof(int32_t) * 2 * static_cast<size_t>(n_pairs), cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  f.release(); w.release();
  return AHX_OK;
}

int ahx_ga_run_islands(ahx_ctx **ctxs, int32_t n, const ahx_ga_params *prm, int32_t L, const int32_t *init_tour, int32_t *best_tour,
                       ahx_ga_stats *stats, ahx_ga_callback cb, void *user) {
  if (!ctxs || n < 1 || !prm) { set_global_error("ahx_ga_run_islands: bad argument"); return AHX_ERR_ARG; }
  if (n == 1) {
    ahx_ga_params p1 = *prm;
    p1.n_islands = 1; p1.island = 0;
    return ahx_ga_run(ctxs[0], &p1, L, init_tour, best_tour, stats, cb, user);
  }
  int rc = AHX_OK, begun = 0;
  for (int i = 0; i < n && rc == AHX_OK; ++i) {
    ahx_ga_params pi = *prm;
    pi.n_islands = n; pi.island = i;
    if ((rc = ahx_ga_begin(ctxs[i], &pi, L, init_tour)) == AHX_OK) ++begun;
    else ctxs[0]->err = ctxs[i]->err;
  }
  if (rc == AHX_OK && (rc = ahx_ga_island_connect_local(ctxs, n)) != AHX_OK)
    for (int i = 1; i < n; ++i) if (!ctxs[i]->err.empty() && ctxs[0]->err.empty()) ctxs[0]->err = ctxs[i]->err;
  std::vector<int32_t> hof(L);
  std::vector<GaState> hs(n);
  auto best_island = [&]() {   // the best hall of fame; lowest island on ties
    int b = 0;
    for (int i = 1; i < n; ++i) if (hs[i].hof_fit < hs[b].hof_fit) b = i;
    return b;
  };
  auto report = [&](int64_t gen) -> int {
    if (!cb) return AHX_OK;
    const int b = best_island();
    double score = 0.0;
    const int r = ahx_ga_best(ctxs[b], hof.data(), &score);
    if (r != AHX_OK) return r;
    cb(user, prm->phase, gen, score, hof.data(), L);
    return AHX_OK;
  };
  for (int i = 0; i < n && rc == AHX_OK; ++i) rc = read_state(ctxs[i], &hs[i]);
  if (rc == AHX_OK) rc = report(0);   // ga.init's callback at generation 0
  const int64_t period = prm->log_every > 0 ? prm->log_every : 500;
  while (rc == AHX_OK && !hs[0].stop) {
    const int64_t next = (hs[0].gen / period + 1) * period;
    int64_t left = next - hs[0].gen;
    while (rc == AHX_OK && left > 0 && !hs[0].stop) {
      const int64_t batch = std::min<int64_t>(left, kPersistChunk);
      for (int i = 0; i < n && rc == AHX_OK; ++i) rc = persist_start(ctxs[i], batch);     // every island's kernel in flight ...
      for (int i = 0; i < n; ++i) {                                                        // ... before waiting for any
        const int r = persist_finish(ctxs[i], &hs[i]);
        if (r != AHX_OK && rc == AHX_OK) { rc = r; ctxs[0]->err = ctxs[i]->err; }
      }
      left -= batch;
    }
    if (rc == AHX_OK && prm->log_every > 0 && hs[0].gen % period == 0) rc = report(hs[0].gen);
  }
  if (rc == AHX_OK) {
    const int b = best_island();
    double score = 0.0;
    rc = ahx_ga_best(ctxs[b], best_tour, &score);   // r.Tour = ga.HallOfFame[0] (evaluate.go:313), over all populations
    if (stats) {
      fill_stats(ctxs[b], hs[b], stats);
      stats->evaluations = 0; stats->device_ms = 0.0;
      for (int i = 0; i < n; ++i) { stats->evaluations += hs[i].evals; stats->device_ms = std::max(stats->device_ms, ctxs[i]->ga_ms); }
    }
  }
  for (int i = 0; i < begun; ++i) ahx_ga_end(ctxs[i]);
  return rc;
}

int ahx_mutate(ahx_ctx *ctx, int32_t L, const int32_t *tour_in, int32_t op, int32_t p, int32_t q, int32_t dir, int32_t *tour_out) {
  int rc = enter(ctx, false);
  if (rc != AHX_OK) return rc;
  if (L < 1 || !tour_in || !tour_out || op < 0 || op > 3) return fail(ctx, AHX_ERR_ARG, "ahx_mutate: bad argument");
  if (op == 1 ? (p < 1 || p >= L) : (p < 0 || q < p || q >= L)) return fail(ctx, AHX_ERR_ARG, "ahx_mutate: positions out of range");
  DevBuf<int32_t> a, b;
  AHX_CUDA(ctx, a.reserve(L)); AHX_CUDA(ctx, b.reserve(L));
  AHX_CUDA(ctx, cudaMemcpyAsync(a.p, tour_in, sizeof(int32_t) * L, cudaMemcpyHostToDevice, ctx->stream));
  mutate_kernel<<<std::min(64, (L + 255) / 256), 256, 0, ctx->stream>>>(a.p, L, op, p, q, dir, b.p);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  AHX_CUDA(ctx, cudaMemcpyAsync(tour_out, b.p, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  a.release(); b.release();
  return AHX_OK;
}

}  // extern "C"
