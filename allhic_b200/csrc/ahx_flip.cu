// ahx_flip.cu -- orientation heuristics flipAll / flipWhole / flipOne
// (orientation.go:31-120) on the device.
//
// EvaluateQ's terms factor as  sum_e S[e][slot_e(signs)]  with
//   S[e][o] = sum_k Q_o[k] * ln(GR[k] + dist_e),  dist_e a function of the tour only,
// so for a fixed tour the 12 logarithms per pair are computed ONCE
// (pair_scores_kernel) and every later sign change is a table lookup:
//   flipWhole  two full EvaluateQ sums (eval_q_kernel, P = 2)
//   flipOne    the reference re-scores the whole tour after each single flip
//              (1+L EvaluateQ calls, each rebuilding a 96*N^2-byte Q()); here
//              one warp walks the tour sequentially and only revisits the
//              pairs that touch the flipped contig (CSR adjacency), keeping
//              the reference's greedy order and its `newScore > score` rule.
//              The score after each step is score - delta instead of a fresh
//              full sum, so accept/reject can differ from the reference only
//              when |delta| is below the rounding noise of the full sum
//              (decision parity within 1e-12 relative ties).
#include <algorithm>
#include <cstring>

#include "ahx_ctx.h"
#include "ahx_device.cuh"

namespace ahx {

// info[e]: bit0 = dir (1: the higher-indexed contig of the pair comes first in
// the tour), bit1 = valid (both contigs in the tour and dist <= LIMIT).
__global__ void pair_scores_kernel(const int *__restrict__ pa, const int *__restrict__ pb, const int *__restrict__ slots,
                                   const int *__restrict__ hist, int E, const int *__restrict__ pos,
                                   const long long *__restrict__ cum, double *__restrict__ S4, unsigned char *__restrict__ info) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int px = pos[pa[e]], py = pos[pb[e]];
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  unsigned char inf = 0;
  if (px >= 0 && py >= 0) {
    const int i = min(px, py), j = max(px, py);
    const long long dist = cum[j - 1] - cum[i];   // orientation.go:181
    const int dir = px < py ? 0 : 1;
    inf = static_cast<unsigned char>(dir);
    if (dist <= kLimitI) {
      inf |= 2;
      double lg[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) lg[k] = log(static_cast<double>(c_GR[k] + dist));
      for (int o = 0; o < 4; ++o) {
        const int h = slots[static_cast<size_t>(e) * 8 + dir * 4 + o];
        if (h < 0) continue;
        const int *q = hist + static_cast<size_t>(h) * 12;
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 12; ++k) a += static_cast<double>(q[k]) * lg[k];
        s[o] = a;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) S4[static_cast<size_t>(e) * 4 + o] = s[o];
  info[e] = inf;
}

// sign codes: 0 '+', 1 '-', 2 anything else (matches no stored orientation)
__device__ __forceinline__ unsigned char sign_code(unsigned char b) { return b == '+' ? 0 : (b == '-' ? 1 : 2); }
__device__ __forceinline__ unsigned char rr_code(unsigned char c) { return c == 1 ? 0 : 1; }  // clm.go:160 rr

__device__ __forceinline__ double pair_term(const double *__restrict__ S4, const unsigned char *__restrict__ info, int e,
                                            unsigned char sa, unsigned char sb) {
  // sa / sb: sign codes of the lower / higher indexed contig of pair e
  const unsigned char inf = info[e];
  if (!(inf & 2)) return 0.0;
  const unsigned char sf = (inf & 1) ? sb : sa, ss = (inf & 1) ? sa : sb;
  if (sf > 1 || ss > 1) return 0.0;
  return S4[static_cast<size_t>(e) * 4 + 2 * sf + ss];
}

// One warp.  Shared memory: sign codes, N bytes.
__global__ void __launch_bounds__(32) flip_one_kernel(const int *__restrict__ tour, int L, int N, unsigned char *__restrict__ signs,
                                                      const int *__restrict__ pa, const int *__restrict__ pb, int E,
                                                      const int *__restrict__ adj_ptr, const int *__restrict__ adj_e,
                                                      const double *__restrict__ S4, const unsigned char *__restrict__ info,
                                                      double *__restrict__ trace, unsigned char *__restrict__ step_acc,
                                                      long long *__restrict__ counts, double *__restrict__ final_score) {
  extern __shared__ unsigned char sg[];
  const int lane = threadIdx.x;
  for (int i = lane; i < N; i += 32) sg[i] = sign_code(signs[i]);
  __syncwarp();
  // score := EvaluateQ() (orientation.go:91), fixed-order warp sum
  double sum = 0.0;
  for (int e = lane; e < E; e += 32) sum += pair_term(S4, info, e, sg[pa[e]], sg[pb[e]]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  double score = 0.0 - sum;
  long long n_acc = 0, n_rej = 0;
  for (int i = 0; i < L; ++i) {
    const int c = tour[i];
    const unsigned char oldc = sg[c], newc = rr_code(oldc);
    double delta = 0.0;   // change of the positive sum when c flips
    for (int k = adj_ptr[c] + lane; k < adj_ptr[c + 1]; k += 32) {
      const int e = adj_e[k];
      const int a = pa[e], b = pb[e];
      const unsigned char sa_old = sg[a], sb_old = sg[b];
      const unsigned char sa_new = (a == c) ? newc : sa_old, sb_new = (b == c) ? newc : sb_old;
      delta += pair_term(S4, info, e, sa_new, sb_new) - pair_term(S4, info, e, sa_old, sb_old);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, d);
    const double new_score = score - delta;
    const bool accept = new_score > score;        // orientation.go:97
    if (lane == 0) {
      if (trace) { trace[2 * i] = score; trace[2 * i + 1] = new_score; }
      if (step_acc) step_acc[i] = accept ? 1 : 0;
      // accepted: keep the flipped sign; rejected: the reference applies rr a second time
      sg[c] = accept ? newc : rr_code(newc);
    }
    if (accept) { score = new_score; ++n_acc; } else ++n_rej;
    __syncwarp();
  }
  for (int i = lane; i < N; i += 32) {
    const unsigned char orig = signs[i];
    const unsigned char code = sg[i];
    // untouched "other" bytes stay as they were
    signs[i] = (sign_code(orig) == 2 && code == 2) ? orig : (code == 1 ? '-' : '+');
  }
  if (lane == 0) { counts[0] = n_acc; counts[1] = n_rej; *final_score = score; }
}

// ---- flipOne, restructured ----------------------------------------------------------
// The greedy pass is sequential (each accept changes the baseline of the next contig), but
// almost nothing in a step depends on earlier decisions: the contig's own sign is still its
// initial one (every contig is visited once), and a partner's sign is either '+' or '-'.  So
// flip_records_kernel computes, for every (step, adjacent pair) IN PARALLEL, the change of
// the sum for both possible partner signs, laid out in step order; the sequential kernel
// then only streams these records (prefetched one step ahead), picks d[sign of partner] and
// reduces.  Same terms, same lane assignment and summation order as flip_one_kernel:
// bit-identical scores, a twentieth of the latency.
__global__ void flip_records_kernel(const int *__restrict__ tour, int L, const unsigned char *__restrict__ signs,
                                    const int *__restrict__ pa, const int *__restrict__ pb, const int *__restrict__ adj_ptr,
                                    const int *__restrict__ adj_e, const double *__restrict__ S4, const unsigned char *__restrict__ info,
                                    const int *__restrict__ step_off, double *__restrict__ rec_d0, double *__restrict__ rec_d1,
                                    int *__restrict__ rec_o) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= L) return;
  const int c = tour[i];
  const unsigned char oldc = sign_code(signs[c]), newc = rr_code(oldc);
  const int beg = adj_ptr[c], deg = adj_ptr[c + 1] - beg, off = step_off[i];
  for (int j = lane; j < deg; j += 32) {
    const int e = adj_e[beg + j];
    const int a = pa[e], b = pb[e];
    const bool c_is_a = a == c;
    double d[2];
#pragma unroll
    for (int so = 0; so < 2; ++so) {   // the partner's sign when this step runs
      const unsigned char s = static_cast<unsigned char>(so);
      d[so] = c_is_a ? pair_term(S4, info, e, newc, s) - pair_term(S4, info, e, oldc, s)
                     : pair_term(S4, info, e, s, newc) - pair_term(S4, info, e, s, oldc);
    }
    rec_d0[off + j] = d[0]; rec_d1[off + j] = d[1]; rec_o[off + j] = c_is_a ? b : a;
  }
}

// kBlock threads compute the starting score, then warp 0 walks the tour.  Shared memory:
// step_off (L + 1 ints), the tour (L ints), the current sign codes (N bytes, all 0 / 1).
__global__ void __launch_bounds__(kBlock) flip_one_stream_kernel(const int *__restrict__ tour_g, int L, int N, unsigned char *__restrict__ signs,
                                                                 const int *__restrict__ pa, const int *__restrict__ pb, int E,
                                                                 const double *__restrict__ S4, const unsigned char *__restrict__ info,
                                                                 const int *__restrict__ step_off_g, const double *__restrict__ rec_d0,
                                                                 const double *__restrict__ rec_d1, const int *__restrict__ rec_o,
                                                                 double *__restrict__ trace, unsigned char *__restrict__ step_acc,
                                                                 long long *__restrict__ counts, double *__restrict__ final_score) {
  extern __shared__ __align__(8) unsigned char smem_flip[];
  __shared__ double red_scratch[kWarps];
  __shared__ double s_score;
  int *step_off = reinterpret_cast<int *>(smem_flip);
  int *tour = step_off + L + 1;
  unsigned char *sg = reinterpret_cast<unsigned char *>(tour + L);
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i <= L; i += kBlock) step_off[i] = step_off_g[i];
  for (int i = tid; i < L; i += kBlock) tour[i] = tour_g[i];
  for (int i = tid; i < N; i += kBlock) sg[i] = sign_code(signs[i]);
  __syncthreads();
  double sum = 0.0;   // score := EvaluateQ() (orientation.go:91), fixed-shape block sum
  for (int e = tid; e < E; e += kBlock) sum += pair_term(S4, info, e, sg[pa[e]], sg[pb[e]]);
  sum = block_sum(sum, red_scratch, tid);
  if (tid == 0) s_score = 0.0 - sum;
  __syncthreads();
  if (tid >= 32) return;
  double score = s_score;
  long long n_acc = 0, n_rej = 0;
  // records of the next kD steps, the first 64 of each (two per lane; few contigs have more partners): the
  // loads of step i + kD are in flight while steps i .. i + kD - 1 are decided (a step is ~400 cycles, a
  // DRAM round trip ~1500)
  constexpr int kW = 2, kD = 4;
  double qd0[kD][kW], qd1[kD][kW]; int qo[kD][kW];
  auto fetch = [&](int s, double (&a0)[kW], double (&a1)[kW], int (&ao)[kW]) {
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      a0[w] = a1[w] = 0.0; ao[w] = -1;
      if (s < L && step_off[s] + w * 32 + lane < step_off[s + 1]) {
        const int k = step_off[s] + w * 32 + lane;
        a0[w] = rec_d0[k]; a1[w] = rec_d1[k]; ao[w] = rec_o[k];
      }
    }
  };
#pragma unroll
  for (int s = 0; s < kD; ++s) fetch(s, qd0[s], qd1[s], qo[s]);
  // the loop is unrolled by kD so that slot s is a fixed set of registers: step i uses slot i mod kD and
  // refills it for step i + kD right away (moving the slots along each step would wait for every load
  // one step after it was issued)
  for (int i0 = 0; i0 < L; i0 += kD) {
#pragma unroll
    for (int s = 0; s < kD; ++s) {
      const int i = i0 + s;
      if (i >= L) break;
      const int beg = step_off[i], end = step_off[i + 1];
      double d0[kW], d1[kW]; int o[kW];
#pragma unroll
      for (int w = 0; w < kW; ++w) { d0[w] = qd0[s][w]; d1[w] = qd1[s][w]; o[w] = qo[s][w]; }
      fetch(i + kD, qd0[s], qd1[s], qo[s]);
      double delta = 0.0;   // change of the positive sum when tour[i] flips (entries in adjacency order, lane-strided)
#pragma unroll
      for (int w = 0; w < kW; ++w) if (o[w] >= 0) delta += sg[o[w]] ? d1[w] : d0[w];
      for (int k = beg + kW * 32 + lane; k < end; k += 32) delta += sg[rec_o[k]] ? rec_d1[k] : rec_d0[k];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, d);
      const double new_score = score - delta;
      const bool accept = new_score > score;        // orientation.go:97
      if (lane == 0) {
        if (trace) { trace[2 * i] = score; trace[2 * i + 1] = new_score; }
        if (step_acc) step_acc[i] = accept ? 1 : 0;
        if (accept) { const int c = tour[i]; sg[c] = rr_code(sg[c]); }   // rejected: the reference flips back (orientation.go:101)
      }
      if (accept) { score = new_score; ++n_acc; } else ++n_rej;
      __syncwarp();
    }
  }
  for (int i = lane; i < N; i += 32) signs[i] = sg[i] == 1 ? '-' : '+';
  if (lane == 0) { counts[0] = n_acc; counts[1] = n_rej; *final_score = score; }
}

// EvaluateQ for many sign vectors on ONE tour: the 12 logarithms per pair depend on the
// tour only, so they are computed once (pair_scores_kernel) and each individual is a
// table lookup per stored pair.  One CTA per individual; shared memory: sign codes, N bytes.
__global__ void __launch_bounds__(kBlock) eval_q_lookup_kernel(int N, int E, const unsigned char *__restrict__ signs, const int *__restrict__ pa,
                                                               const int *__restrict__ pb, const double *__restrict__ S4,
                                                               const unsigned char *__restrict__ info, double *__restrict__ out) {
  extern __shared__ unsigned char sg[];
  __shared__ double red_scratch[kWarps];
  const int tid = threadIdx.x, p = blockIdx.x;
  for (int i = tid; i < N; i += kBlock) sg[i] = sign_code(signs[static_cast<size_t>(p) * N + i]);
  __syncthreads();
  double acc = 0.0;
  for (int e = tid; e < E; e += kBlock) acc += pair_term(S4, info, e, sg[pa[e]], sg[pb[e]]);
  const double s = block_sum(acc, red_scratch, tid);
  if (tid == 0) out[p] = 0.0 - s;
}

static int upload_tour_index(ahx_ctx *ctx, int32_t L, const int32_t *tour) {
  // pos = inverse permutation, cum = cumsize (orientation.go:165-170)
  std::vector<int32_t> pos(ctx->N, -1);
  std::vector<long long> cum(std::max<int32_t>(L, 1));
  long long c = 0;
  for (int32_t i = 0; i < L; ++i) { pos[tour[i]] = i; cum[i] = c; c += ctx->h_sizes[tour[i]]; }
  AHX_CUDA(ctx, ctx->d_pos.reserve(ctx->N)); AHX_CUDA(ctx, ctx->d_cum.reserve(std::max<int32_t>(L, 1)));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_pos.p, pos.data(), sizeof(int32_t) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_cum.p, cum.data(), sizeof(long long) * L, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // pos/cum are stack-owned
  return AHX_OK;
}

// d_signs: P x N sign bytes on the device; d_out: P scores.  Used by ahx_eval_q(shared_tour != 0).
int eval_q_shared_tour(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tour, const uint8_t *d_signs, double *d_out) {
  int rc = upload_tour_index(ctx, L, tour);
  if (rc != AHX_OK) return rc;
  const int E = static_cast<int>(ctx->E);
  AHX_CUDA(ctx, ctx->d_S4.reserve(static_cast<size_t>(std::max(E, 1)) * 4)); AHX_CUDA(ctx, ctx->d_pairinfo.reserve(std::max(E, 1)));
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  if (E > 0) {
    pair_scores_kernel<<<(E + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_pa.p, ctx->d_pb.p, ctx->d_slots.p, ctx->d_hist.p, E, ctx->d_pos.p,
                                                                 ctx->d_cum.p, ctx->d_S4.p, ctx->d_pairinfo.p);
    ctx->launches++;
  }
  AHX_CUDA(ctx, opt_in_smem(eval_q_lookup_kernel, ctx->smem_optin));
  eval_q_lookup_kernel<<<P, kBlock, ctx->N, ctx->stream>>>(ctx->N, E, d_signs, ctx->d_pa.p, ctx->d_pb.p, ctx->d_S4.p, ctx->d_pairinfo.p, d_out);
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  return AHX_OK;
}

// EvaluateQ for two sign vectors on one tour; out[0], out[1].
static int eval_q_two(ahx_ctx *ctx, int32_t L, const int32_t *tour, const uint8_t *signs_a, const uint8_t *signs_b, double *out) {
  std::vector<uint8_t> both(static_cast<size_t>(2) * ctx->N);
  std::memcpy(both.data(), signs_a, ctx->N);
  std::memcpy(both.data() + ctx->N, signs_b, ctx->N);
  return ahx_eval_q(ctx, 2, L, tour, 1, both.data(), out);
}

}  // namespace ahx

using namespace ahx;

extern "C" {

int ahx_flip_whole(ahx_ctx *ctx, int32_t L, const int32_t *tour, uint8_t *signs, double *old_score, double *new_score, int32_t *accepted) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!signs) return fail(ctx, AHX_ERR_ARG, "ahx_flip_whole: null signs");
  if ((rc = validate_tour(ctx, L, tour, "ahx_flip_whole")) != AHX_OK) return rc;
  std::vector<uint8_t> flipped(ctx->N);
  for (int32_t i = 0; i < ctx->N; ++i) flipped[i] = signs[i] == '-' ? '+' : '-';   // rr, orientation.go:73-75
  double sc[2];
  if ((rc = eval_q_two(ctx, L, tour, signs, flipped.data(), sc)) != AHX_OK) return rc;
  const bool accept = !(sc[1] <= sc[0]);                                          // orientation.go:78
  if (accept) std::memcpy(signs, flipped.data(), ctx->N);
  if (old_score) *old_score = sc[0];
  if (new_score) *new_score = sc[1];
  if (accepted) *accepted = accept ? 1 : 0;
  return AHX_OK;
}

int ahx_flip_one(ahx_ctx *ctx, int32_t L, const int32_t *tour, uint8_t *signs, double *final_score, int64_t *n_accepts,
                 int64_t *n_rejects, double *trace, uint8_t *step_accepted, int32_t *accepted) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!signs) return fail(ctx, AHX_ERR_ARG, "ahx_flip_one: null signs");
  if ((rc = validate_tour(ctx, L, tour, "ahx_flip_one")) != AHX_OK) return rc;
  // CLM.Signs only ever holds '+' / '-' (Activate writes '+', every change goes through rr, clm.go:160-165).
  // With any other byte the reference's greedy pass compares against a stale score (a rejected step
  // rewrites the byte to '+' without re-scoring): not worth imitating, so it is an argument error here.
  for (int32_t i = 0; i < ctx->N; ++i)
    if (signs[i] != '+' && signs[i] != '-') return fail(ctx, AHX_ERR_ARG, "ahx_flip_one: signs must be '+' or '-'");
  if (static_cast<size_t>(ctx->N) + 1024 > ctx->smem_optin) return fail(ctx, AHX_ERR_ARG, "ahx_flip_one: too many contigs");
  if ((rc = upload_tour_index(ctx, L, tour)) != AHX_OK) return rc;
  const int E = static_cast<int>(ctx->E), Lm = std::max<int32_t>(L, 1);
  AHX_CUDA(ctx, ctx->d_S4.reserve(static_cast<size_t>(std::max(E, 1)) * 4)); AHX_CUDA(ctx, ctx->d_pairinfo.reserve(std::max(E, 1)));
  AHX_CUDA(ctx, ctx->d_trace.reserve(static_cast<size_t>(Lm) * 2 + 2)); AHX_CUDA(ctx, ctx->d_stepacc.reserve(Lm));
  AHX_CUDA(ctx, ctx->d_counts.reserve(2)); AHX_CUDA(ctx, ctx->d_signs.reserve(ctx->N)); AHX_CUDA(ctx, ctx->d_tours.reserve(Lm));
  AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_signs.p, signs, ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  if (L) AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_tours.p, tour, sizeof(int32_t) * L, cudaMemcpyHostToDevice, ctx->stream));
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  if (E > 0) {
    pair_scores_kernel<<<(E + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_pa.p, ctx->d_pb.p, ctx->d_slots.p, ctx->d_hist.p, E, ctx->d_pos.p,
                                                                 ctx->d_cum.p, ctx->d_S4.p, ctx->d_pairinfo.p);
    ctx->launches++;
    AHX_CUDA(ctx, cudaGetLastError());
  }
  double *d_final = ctx->d_trace.p + static_cast<size_t>(Lm) * 2;
  const size_t smem_stream = sizeof(int) * (2 * static_cast<size_t>(L) + 1) + ctx->N;
  if (smem_stream + 1024 <= ctx->smem_optin) {
    std::vector<int32_t> off(static_cast<size_t>(L) + 1, 0);
    for (int32_t i = 0; i < L; ++i) off[i + 1] = off[i] + ctx->h_deg[tour[i]];
    const size_t nrec = static_cast<size_t>(std::max(off[L], 1));
    AHX_CUDA(ctx, ctx->d_step_off.reserve(static_cast<size_t>(L) + 1)); AHX_CUDA(ctx, ctx->d_rec_o.reserve(nrec));
    AHX_CUDA(ctx, ctx->d_rec_d.reserve(2 * nrec));
    AHX_CUDA(ctx, cudaMemcpyAsync(ctx->d_step_off.p, off.data(), sizeof(int32_t) * (L + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (L > 0) {
      flip_records_kernel<<<(L + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_tours.p, L, ctx->d_signs.p, ctx->d_pa.p, ctx->d_pb.p, ctx->d_adj_ptr.p, ctx->d_adj_e.p,
                                                               ctx->d_S4.p, ctx->d_pairinfo.p, ctx->d_step_off.p, ctx->d_rec_d.p, ctx->d_rec_d.p + nrec,
                                                               ctx->d_rec_o.p);
      ctx->launches++;
    }
    AHX_CUDA(ctx, opt_in_smem(flip_one_stream_kernel, ctx->smem_optin));
    flip_one_stream_kernel<<<1, kBlock, smem_stream, ctx->stream>>>(ctx->d_tours.p, L, ctx->N, ctx->d_signs.p, ctx->d_pa.p, ctx->d_pb.p, E, ctx->d_S4.p,
                                                                ctx->d_pairinfo.p, ctx->d_step_off.p, ctx->d_rec_d.p, ctx->d_rec_d.p + nrec, ctx->d_rec_o.p,
                                                                ctx->d_trace.p, ctx->d_stepacc.p, ctx->d_counts.p, d_final);
    AHX_CUDA(ctx, cudaGetLastError());
    AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `off` is stack-owned
  } else {
    AHX_CUDA(ctx, opt_in_smem(flip_one_kernel, ctx->smem_optin));
    flip_one_kernel<<<1, 32, ctx->N, ctx->stream>>>(ctx->d_tours.p, L, ctx->N, ctx->d_signs.p, ctx->d_pa.p, ctx->d_pb.p, E, ctx->d_adj_ptr.p,
                                                   ctx->d_adj_e.p, ctx->d_S4.p, ctx->d_pairinfo.p, ctx->d_trace.p, ctx->d_stepacc.p,
                                                   ctx->d_counts.p, d_final);
  }
  ctx->launches++;
  AHX_CUDA(ctx, cudaGetLastError());
  // The pass keeps a running score (score - delta per accepted step); the reference's final `score` is a full
  // EvaluateQ sum (orientation.go:96-99).  Re-synchronise: one full sum over the final signs, returned as the
  // final score, so rounding never accumulates from one pass into the next.
  if (static_cast<size_t>(ctx->N) + 1024 <= ctx->smem_optin) {
    AHX_CUDA(ctx, opt_in_smem(eval_q_lookup_kernel, ctx->smem_optin));
    eval_q_lookup_kernel<<<1, kBlock, ctx->N, ctx->stream>>>(ctx->N, E, ctx->d_signs.p, ctx->d_pa.p, ctx->d_pb.p, ctx->d_S4.p, ctx->d_pairinfo.p, d_final + 1);
    ctx->launches++;
    AHX_CUDA(ctx, cudaGetLastError());
  } else {
    AHX_CUDA(ctx, cudaMemcpyAsync(d_final + 1, d_final, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  AHX_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  long long counts[2] = {0, 0};
  double fs = 0.0;
  AHX_CUDA(ctx, cudaMemcpyAsync(signs, ctx->d_signs.p, ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaMemcpyAsync(counts, ctx->d_counts.p, sizeof(counts), cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaMemcpyAsync(&fs, d_final + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (trace && L) AHX_CUDA(ctx, cudaMemcpyAsync(trace, ctx->d_trace.p, sizeof(double) * 2 * L, cudaMemcpyDeviceToHost, ctx->stream));
  if (step_accepted && L) AHX_CUDA(ctx, cudaMemcpyAsync(step_accepted, ctx->d_stepacc.p, L, cudaMemcpyDeviceToHost, ctx->stream));
  AHX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  AHX_CUDA(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev_a, ctx->ev_b));
  if (final_score) *final_score = fs;
  if (n_accepts) *n_accepts = counts[0];
  if (n_rejects) *n_rejects = counts[1];
  if (accepted) *accepted = counts[0] > 0 ? 1 : 0;   // anyTagACCEPT, orientation.go:115-119
  return AHX_OK;
}

int ahx_flip_all(ahx_ctx *ctx, int32_t L, const int32_t *tour, uint8_t *signs, double *old_score, double *new_score, int32_t *accepted) {
  int rc = enter(ctx, true);
  if (rc != AHX_OK) return rc;
  if (!signs) return fail(ctx, AHX_ERR_ARG, "ahx_flip_all: null signs");
  if ((rc = validate_tour(ctx, L, tour, "ahx_flip_all")) != AHX_OK) return rc;
  // O[a][b] = strandedness * nlinks over ALL tigs (orientation.go:124-132); the
  // eigenvector of the largest eigenvalue gives the sign pattern (:42-52).
  std::vector<double> w(ctx->E), v(ctx->N);
  for (int64_t e = 0; e < ctx->E; ++e) w[e] = static_cast<double>(ctx->h_strand[e]) * static_cast<double>(ctx->h_nl[e]);
  top_eigenvector(ctx->N, ctx->E, ctx->h_pa.data(), ctx->h_pb.data(), w.data(), ctx->h_diag.empty() ? nullptr : ctx->h_diag.data(), v.data(), &ctx->eig);
  std::vector<uint8_t> cand(ctx->N);
  // `v < 0 => '-'` (orientation.go:47-51); components that are zero up to rounding (contigs without any contact,
  // the weaker components of a disconnected graph) are '+', as an exact zero is in the reference
  double vmax = 0.0;
  for (int32_t i = 0; i < ctx->N; ++i) vmax = std::max(vmax, std::fabs(v[i]));
  for (int32_t i = 0; i < ctx->N; ++i) cand[i] = v[i] < -1e-12 * vmax ? '-' : '+';
  double sc[2];
  if ((rc = eval_q_two(ctx, L, tour, signs, cand.data(), sc)) != AHX_OK) return rc;
  const bool accept = !(sc[1] < sc[0]);                                           // orientation.go:57
  if (accept) std::memcpy(signs, cand.data(), ctx->N);
  if (old_score) *old_score = sc[0];
  if (new_score) *new_score = sc[1];
  if (accepted) *accepted = accept ? 1 : 0;
  return AHX_OK;
}

int ahx_flip_all_info(const ahx_ctx *ctx, double *residual, int32_t *converged, int32_t *dense, int32_t *matvecs) {
  if (!ctx) return AHX_ERR_ARG;
  if (residual) *residual = ctx->eig.residual;
  if (converged) *converged = ctx->eig.converged ? 1 : 0;
  if (dense) *dense = ctx->eig.dense ? 1 : 0;
  if (matvecs) *matvecs = ctx->eig.matvecs;
  return AHX_OK;
}

}  // extern "C"
