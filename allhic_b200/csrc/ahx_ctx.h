// ahx_ctx.h -- the context object behind the opaque ahx_ctx handle.
#ifndef AHX_CTX_H_
#define AHX_CTX_H_

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "ahx_internal.h"

namespace ahx {

// A device allocation that only ever grows.
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Opts a kernel into the device's full shared-memory carve-out: the dynamic limit is the opt-in maximum minus
// the kernel's own static shared memory (the sum of the two is what the driver checks).  The value is a
// constant per (kernel, device), so concurrent contexts that launch the same kernel never fight over it.
template <typename K>
inline cudaError_t opt_in_smem(K kern, size_t smem_optin) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return e;
  const long long room = static_cast<long long>(smem_optin) - static_cast<long long>(fa.sharedSizeBytes);
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(room > 0 ? room : 0));
}

// Device-resident GA bookkeeping (one per context), evaluate.go:281-306.
struct GaState {
  long long gen;        // ga.Generations
  long long updated;    // *updated
  long long evals;      // Tour.Evaluate calls on the device
  double hof_fit;       // HallOfFame[0].Fitness (minimised)
  int stop;             // EarlyStop fired or max_gens reached
  unsigned int ticket;  // "last CTA done" counter
  int best_idx;         // argmin of the newest generation
  int pad;
};

// Island model (include/ahx.h "Islands"): what the persistent GA kernel needs to reach the
// other islands.  Lives in device memory; mbox[i] is island i's mailbox as mapped into THIS
// device's address space (peer access in one process, CUDA IPC across processes).
constexpr int kMaxIslands = 32;
struct GaIslands {
  unsigned char *mbox[kMaxIslands];
  double g_best;                 // best fitness over all islands as of the last exchange
  long long g_updated;           // generation (an exchange point) at which g_best last improved
  unsigned long long epochs;     // exchanges completed
  long long timeout_ns;          // how long an island waits for its peers before giving up
  int err;                       // 1: a peer did not show up in time
  int pad;
  // decisions of an exchange, written by CTA 0 and read by every CTA after the grid barrier
  double x_hof_fit; long long x_updated; int x_stop; int x_pad;
};
// Mailbox layout: one flag per source island (32 bytes apart), then [source][parity] slots of
//   double hof_fit; double fit[n_elite] (padded to 16 B); int tour[n_elite][L]; uint32 X[n_elite][N]
__host__ __device__ inline size_t isl_align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }
__host__ __device__ inline size_t isl_flags_bytes(int W) { return (32 * static_cast<size_t>(W) + 127) & ~static_cast<size_t>(127); }
__host__ __device__ inline size_t isl_head_bytes(int n_elite) { return isl_align16(8 + 8 * static_cast<size_t>(n_elite)); }
__host__ __device__ inline size_t isl_slot_bytes(int n_elite, int L, int N) {
  return isl_head_bytes(n_elite) + isl_align16(4 * static_cast<size_t>(n_elite) * (static_cast<size_t>(L) + N));
}
__host__ __device__ inline size_t isl_mbox_bytes(int W, int n_elite, int L, int N) {
  return isl_flags_bytes(W) + 2 * static_cast<size_t>(W) * isl_slot_bytes(n_elite, L, N);
}

}  // namespace ahx

struct ahx_ctx {
  int device = -1;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  std::string err;
  // 4 KB of pinned host memory for small read-backs: a copy to pageable memory keeps the calling thread
  // inside the driver until the stream reaches it, which stalls other threads' launches on this device
  // (islands of one process wait for each other's kernels)
  void *h_pin = nullptr;
  int sm_count = 0;
  size_t smem_optin = 0, l2_bytes = 0;
  int64_t launches = 0;
  float last_ms = 0.f;            // CUDA-event time of the kernels of the last ahx_eval_q / ahx_flip_one call (ahx_last_device_ms)

  // ---- loaded linkage group
  bool loaded = false;
  int32_t N = 0;
  int64_t E = 0, H = 0;
  bool coo16 = false;
  ahx::DevBuf<long long> d_sizes;
  ahx::DevBuf<uint32_t> d_sizes32;       // sizes as uint32 for the x32 kernels (valid when x32_ok)
  ahx::DevBuf<uint2> d_coo16, d_coo16o;   // sorted by (a,b) / bank-conflict-free order, padded
  ahx::DevBuf<int4> d_coo32, d_coo32o;
  int64_t E_pad = 0;                      // entries in the ordered table (multiple of 256)
  // resident-table kernel (ahx_resident.cuh): per entry the byte offsets of the two contigs in X,
  // (a*4T) | (b*4T) << 16, one table per T = 1, 2, 4, and an 8-bit link count (pairs with more than
  // 255 links are split over several entries); bank-conflict-free order, padded with (0, 1, 0)
  // index ti = log2(T): every T has its own bank-conflict-free ordering (groups of 32 / T lanes)
  ahx::DevBuf<uint32_t> d_rpairs[3];   // resident mode: byte offsets into X (only when N * 4T < 64 KB)
  ahx::DevBuf<uint32_t> d_rraw[3];     // streamed mode: contig indices a | b << 16
  ahx::DevBuf<uint8_t> d_rnl8[3];
  int64_t E_res[3] = {0, 0, 0};           // entries, multiple of kRPad (0 when E == 0)
  bool x32_ok = false;                    // 2*sum(sizes) + 2*LIMIT + 2 < 2^32
  ahx::DevBuf<int32_t> d_pa, d_pb, d_slots, d_hist, d_M;
  bool dense_ready = false;
  ahx::DevBuf<int32_t> d_adj_ptr, d_adj_e;   // CSR: pairs touching each contig
  std::vector<int64_t> h_sizes;
  std::vector<int32_t> h_pa, h_pb, h_nl;
  std::vector<int8_t> h_strand;
  std::vector<double> h_diag;              // O[a][a] of self-pair CLM lines (empty: none)
  ahx::EigInfo eig{};                      // how the last flipAll's eigenvector was obtained
  std::vector<int32_t> h_deg;              // pairs touching each contig (host copy of the CSR row lengths)

  // ---- scratch
  ahx::DevBuf<int32_t> d_tours;      // staging for ahx_eval_*
  ahx::DevBuf<int32_t> d_res_tours;  // resident population (ahx_pop_*)
  ahx::DevBuf<double> d_res_scores;
  ahx::DevBuf<uint16_t> d_tours16;
  ahx::DevBuf<double> d_scores;
  ahx::DevBuf<uint8_t> d_signs;
  ahx::DevBuf<double> d_xglobal;     // contig-space mid-points when N does not fit shared memory
  ahx::DevBuf<char> d_flush;
  ahx::DevBuf<int> d_errflag;        // set by kernels that meet an out-of-range contig index
  int32_t pop_P = 0, pop_L = 0;

  // ---- GA
  bool ga_active = false;
  ahx_ga_params ga_prm{};
  int32_t ga_L = 0;
  ahx::DevBuf<int32_t> d_pop[2];
  ahx::DevBuf<double> d_fit[2];
  ahx::DevBuf<int32_t> d_hof, d_sel_idx, d_sel_tours;
  ahx::DevBuf<double> d_sel_fit;
  ahx::DevBuf<uint8_t> d_taken;
  ahx::DevBuf<ahx::GaState> d_state;
  ahx::DevBuf<int32_t> d_plan;         // per-offspring parent/info/p/q, mutated list, count
  bool ga_x32 = false;
  bool ga_global = false;              // rows and coordinates in global memory (groups beyond the shared-memory kernels)
  int ga_T = 1;
  bool ga_res = false;                 // persistent resident-table kernel (ahx_ga_res.cuh): d_pop[0] is a pool of 2P rows
  bool ga_stream = false;              // ... with the pair table streamed through a ring
  int ga_G = 0, ga_nbmax = 0;
  size_t ga_smem = 0;
  ahx::DevBuf<int32_t> d_rowof[2], d_sel_rows;
  ahx::DevBuf<double> d_row_fit;
  ahx::DevBuf<unsigned int> d_gbar;           // [0] grid barrier counter, [1] hall-of-fame index scratch
  ahx::DevBuf<uint32_t> d_used;               // 3 bit maps of referenced pool rows
  ahx::DevBuf<unsigned long long> d_gmin;
  ahx::DevBuf<uint32_t> d_gmask;              // mutation masks of three generations (published two ahead)
  ahx::DevBuf<uint32_t> d_xpool;              // coordinate row of every pool row
  // islands: own mailbox, the table of peer mailboxes, IPC mappings opened so far (handle bytes -> pointer)
  ahx::DevBuf<unsigned char> d_qscratch;      // eval_q_kernel's per-CTA arrays for groups too large for shared memory
  ahx::DevBuf<unsigned char> d_mbox;
  ahx::DevBuf<ahx::GaIslands> d_isl;
  bool isl_connected = false;
  std::vector<std::pair<std::string, void *>> isl_opened;
  std::vector<int32_t> ga_active_set;  // sorted contig ids of the initial tour
  double ga_ms = 0.0;

  // ---- orientation scratch
  ahx::DevBuf<int32_t> d_pos;
  ahx::DevBuf<long long> d_cum;
  ahx::DevBuf<double> d_S4, d_trace;
  ahx::DevBuf<uint8_t> d_pairinfo, d_stepacc;
  ahx::DevBuf<long long> d_counts;
  ahx::DevBuf<int32_t> d_step_off, d_rec_o;   // flipOne: per-step records (ahx_flip.cu)
  ahx::DevBuf<double> d_rec_d;
};

namespace ahx {
int fail(ahx_ctx *ctx, int code, const std::string &msg);
int cuda_fail(ahx_ctx *ctx, cudaError_t e, const char *what);
// Selects the device; returns AHX_OK or an error code.
int enter(ahx_ctx *ctx, bool need_loaded);
// Launches the contig-space EvaluateM kernel over P tours already on the device.
// rows (optional): evaluate tours[rows[p]] and write out[rows[p]].
int launch_eval_m_sparse(ahx_ctx *ctx, const int32_t *d_tours, const uint16_t *d_tours16, const int32_t *d_rows,
                         int32_t P, int32_t L, double *d_out, cudaStream_t s);
int launch_eval_m_dense(ahx_ctx *ctx, const int32_t *d_tours, int32_t P, int32_t L, double *d_out, cudaStream_t s);
int launch_eval_q(ahx_ctx *ctx, const int32_t *d_tours, int shared_tour, const uint8_t *d_signs, int32_t P, int32_t L,
                  double *d_out, cudaStream_t s);
int eval_q_shared_tour(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tour, const uint8_t *d_signs, double *d_out);
int finish_checked(ahx_ctx *ctx, const char *who);
int validate_tour(ahx_ctx *ctx, int32_t L, const int32_t *tour, const char *who);
size_t eval_m_smem_bytes(int T, int32_t N, int32_t L);
size_t eval_x32_smem_bytes(int T, bool coo16, int32_t N, int32_t L);
bool use_x32(const ahx_ctx *ctx, int32_t L);
// Tours per batch of the persistent evaluate kernel for a population of P tours (0: not usable)
// and whether the pair table is streamed through a ring instead of resident in shared memory.
struct ResPlan { int T = 0; bool stream = false; bool sizes = false; };
ResPlan resident_plan(const ahx_ctx *ctx, int32_t P);
int resident_T(const ahx_ctx *ctx, int32_t P);
inline int res_ti(int T) { return T == 4 ? 2 : T - 1; }
}  // namespace ahx

#define AHX_CUDA(ctx, call)                                              \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) return ahx::cuda_fail((ctx), e_, #call);      \
  } while (0)

#endif  // AHX_CTX_H_
