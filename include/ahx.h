/*
 * ahx.h -- C ABI of libahx.so, the B200-native engine behind `allhic optimize`.
 *
 * This is the drop-in boundary for the reference's optimize hot path
 * (tanghaibao/allhic; file:line citations are relative to that repo).  The Go
 * host keeps its CLI, CLM/Tour/Optimizer types and file formats and calls
 * these entry points through cgo at the three call sites that own whole
 * batches of work (see INTEGRATION.md for the cgo stub):
 *
 *   NewCLM / Activate        clm.go:121, clm.go:392     -> ahx_clm_*, ahx_load*
 *   CLM.GARun                evaluate.go:258            -> ahx_ga_run (or begin/advance/end)
 *   Tour.Evaluate (batched)  evaluate.go:116            -> ahx_eval_m*
 *   CLM.EvaluateQ            orientation.go:159         -> ahx_eval_q
 *   flipAll/flipWhole/flipOne orientation.go:31,67,87   -> ahx_flip_all/_whole/_one
 *
 * Conventions
 *   - plain C types only; every buffer in a signature is HOST memory owned by
 *     the caller (pinned or pageable; pinned makes the copies asynchronous).
 *     The library owns all device memory and frees it in ahx_destroy.
 *   - every call returns 0 (AHX_OK) or a negative AHX_ERR_* code; the message
 *     is available from ahx_last_error().  Nothing exits, throws or calls back
 *     into the host from a CUDA thread.
 *   - a context is bound to one GPU and is NOT thread-safe: one context per
 *     goroutine/GPU.  Every entry point selects its device first.
 *   - there is no CPU fallback: without a CUDA device ahx_create fails with
 *     AHX_ERR_NODEVICE.
 *   - contigs are indexed 0..N-1 in REfile order (CLM.Tigs, clm.go:141-157);
 *     a tour is L distinct contig indices (Tour.Tigs[i].Idx); signs are the
 *     bytes '+' / '-' indexed by contig (CLM.Signs).
 */
#ifndef AHX_H_
#define AHX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AHX_VERSION 100            /* 0.1.0 */
#define AHX_BB 12                  /* base.go:34  BB: GoldenArray bins       */
#define AHX_LIMIT 10000000         /* evaluate.go:22 LIMIT                    */

#define AHX_OK 0
#define AHX_ERR_ARG (-1)           /* bad argument                            */
#define AHX_ERR_CUDA (-2)          /* CUDA runtime error                      */
#define AHX_ERR_STATE (-3)         /* call out of order (e.g. eval before load) */
#define AHX_ERR_NOMEM (-4)
#define AHX_ERR_NODEVICE (-5)      /* no usable CUDA device: there is no CPU path */
#define AHX_ERR_IO (-6)            /* file could not be read / malformed       */

typedef struct ahx_ctx ahx_ctx;    /* one GPU + one loaded linkage group       */
typedef struct ahx_clm ahx_clm;    /* host-side parsed CLM (no GPU needed)     */

int ahx_version(void);
/* Message of the last failed call on ctx (ctx may be NULL: last failure of a
 * call that had no context, e.g. ahx_create / ahx_clm_*). */
const char *ahx_last_error(const ahx_ctx *ctx);

/* ------------------------------------------------------------------------- *
 * CLM ingestion (host only).  Replaces NewCLM/readRE/readClmLines/readClm,
 * clm.go:121-240, with identical map semantics: per line GoldenArray
 * (base.go:156) and SumLog (base.go:138); contacts[pair] keeps the orientation
 * with the smallest SumLog (clm.go:229-236); orientedContacts[(a,b,ao,bo)] and
 * its mirror (b,a,rr(bo),rr(ao)) take the line's GArray, last line wins
 * (clm.go:237-238); lines naming a contig that is not in the REfile are
 * skipped (clm.go:211-218).
 * ------------------------------------------------------------------------- */
/* Multi-threaded parser for `allhic optimize <REfile> <clmfile>` inputs. */
int ahx_clm_parse(const char *clmfile, const char *refile, int nthreads, ahx_clm **out);
/* The same for all linkage groups of one genome in ONE pass over the CLM text: `pipeline` hands the
 * same genome-wide clmfile to every group's Optimizer (allhic.go:285-293) and each NewCLM filters it by
 * its own REfile (clm.go:211-218) -- k groups, k full parses.  Here the file is read and tokenised once,
 * a line's GoldenArray / SumLog computed once, and out[g] is exactly what ahx_clm_parse(clmfile,
 * refiles[g]) returns.  refiles: e.g. partition's <prefix>.counts_<RE>.<k>g<j>.txt (partition.go:205-215). */
int ahx_clm_parse_groups(const char *clmfile, const char *const *refiles, int32_t n_groups, int nthreads,
                         ahx_clm **out /* n_groups */);
/* Builder for a host that keeps its own parser (the Go readClm loop): create
 * with the REfile's sizes, then feed CLM lines in file order. */
int ahx_clm_new(int32_t n_tigs, const int64_t *sizes, ahx_clm **out);
int ahx_clm_add_line(ahx_clm *clm, int32_t ai, int32_t bi, uint8_t ao, uint8_t bo,
                     const int64_t *dists, int64_t n_dists);
/* Freezes the maps into the flat pair table below (idempotent). */
int ahx_clm_finalize(ahx_clm *clm);
void ahx_clm_free(ahx_clm *clm);

int32_t ahx_clm_ntigs(const ahx_clm *clm);
const char *ahx_clm_tig_name(const ahx_clm *clm, int32_t i); /* "" for a builder-made clm */
int64_t ahx_clm_tig_size(const ahx_clm *clm, int32_t i);
int32_t ahx_clm_tig_index(const ahx_clm *clm, const char *name); /* tigToIdx; -1 if absent */
/* Flat pair table (valid after finalize).  One row per unordered contig pair
 * {a<b} that has at least one CLM line:
 *   nlinks[e]  = M[a][b]                         (clm.go:437-447)
 *   strand[e]  = strandedness of the kept contact (O = strand*nlinks, orientation.go:124)
 *   slots[e*8 + dir*4 + 2*[sa=='-'] + [sb=='-']] = row of hist[] holding the
 *       GArray of orientedContacts[(first,second,sa,sb)] or -1 if absent, where
 *       dir 0 means (first,second)=(a,b) and dir 1 means (first,second)=(b,a)
 *   hist[h*12 + k]                                (base.go:98 GArray)            */
int64_t ahx_clm_npairs(const ahx_clm *clm);
int64_t ahx_clm_nhists(const ahx_clm *clm);
int ahx_clm_get_pairs(const ahx_clm *clm, int32_t *pair_a, int32_t *pair_b, int32_t *nlinks,
                      int8_t *strand, int32_t *slots /* E*8 */);
int ahx_clm_get_hists(const ahx_clm *clm, int32_t *hist /* H*12 */);

/* ------------------------------------------------------------------------- *
 * Context and upload ("build M/Q once on the host and upload them").
 * ------------------------------------------------------------------------- */
int ahx_device_count(void);
int ahx_create(int device, ahx_ctx **out);
void ahx_destroy(ahx_ctx *ctx);
/* Upload one linkage group from flat arrays (layout as ahx_clm_get_*).
 * Requires sizes[i] >= 1, 0 <= pair_a[e] < pair_b[e] < n_tigs, pairs unique. */
int ahx_load(ahx_ctx *ctx, int32_t n_tigs, const int64_t *sizes, int64_t n_pairs,
             const int32_t *pair_a, const int32_t *pair_b, const int32_t *nlinks,
             const int8_t *strand, const int32_t *slots, int64_t n_hists, const int32_t *hist);
int ahx_load_clm(ahx_ctx *ctx, const ahx_clm *clm);

/* Pinned host memory for callers that want asynchronous copies. */
int ahx_host_alloc(void **ptr, uint64_t bytes);
int ahx_host_free(void *ptr);

/* ------------------------------------------------------------------------- *
 * Tour.Evaluate ("EvaluateM"), evaluate.go:116-143, for a whole population:
 *   out[p] = - sum_{i<j, mid_j-mid_i <= LIMIT} M[t_i][t_j] / (mid_j - mid_i)
 * tours: P x L contig indices (row p = individual p); out: P doubles.
 * ahx_eval_m       contig-space kernel over the nonzero pairs of M (product path)
 * ahx_eval_m_u16   same, tours as uint16 (requires n_tigs <= 65535; halves the copy)
 * ahx_eval_m_dense position-space kernel gathering the dense int32 M (north_star's
 *                  literal formulation; builds the N x N matrix on first use)
 * Copies are chunked and overlapped with the kernels.
 * ------------------------------------------------------------------------- */
int ahx_eval_m(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tours, double *out);
int ahx_eval_m_u16(ahx_ctx *ctx, int32_t P, int32_t L, const uint16_t *tours, double *out);
int ahx_eval_m_dense(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tours, double *out);

/* CLM.EvaluateQ, orientation.go:159-193, for P individuals:
 *   out[p] = - sum_{i<j, entry for current signs, dist<=LIMIT} sum_k Q[k]*ln(GR[k]+dist),
 *   dist = cumsize[j-1]-cumsize[i]
 * tours: P x L, or one shared tour when shared_tour != 0; signs: P x n_tigs bytes. */
int ahx_eval_q(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tours, int32_t shared_tour,
               const uint8_t *signs, double *out);

/* Resident population, for measuring the kernels with inputs already in HBM. */
int ahx_pop_set(ahx_ctx *ctx, int32_t P, int32_t L, const int32_t *tours);
/* kind: 0 sparse (product), 1 dense.  Launches `repeat` times back to back on
 * the context stream, times them with CUDA events, returns ms for all repeats. */
int ahx_pop_eval_m(ahx_ctx *ctx, int32_t kind, int32_t repeat, float *elapsed_ms);
int ahx_pop_scores(ahx_ctx *ctx, double *out /* P */);
/* Which kernel ahx_eval_m / ahx_pop_eval_m(kind 0) would launch for P tours of length L:
 * 2 eval_m_res_kernel with the pair table resident in shared memory, 3 the same kernel with
 * the table streamed through a ring (large groups), 1 eval_m_x32_kernel (32-bit coordinates,
 * one CTA per T tours), 0 eval_m_sparse_kernel (fp64 coordinates).
 * *tours_per_batch (optional) receives the tours one CTA scores together. */
int ahx_eval_m_kernel(const ahx_ctx *ctx, int32_t P, int32_t L, int32_t *tours_per_batch);
/* Entries of eval_m_res_kernel's pair table for tours_per_batch = 1, 2 or 4: the stored pairs
 * (pairs with more than 255 links split) plus the padding of the bank-conflict-free order;
 * 0 if that table was not built.  Every entry costs one term per tour (roofline bookkeeping). */
int64_t ahx_pair_table_entries(const ahx_ctx *ctx, int32_t tours_per_batch);
/* Writes `bytes` of device memory (> L2) to evict the L2 between timed runs. */
int ahx_flush_l2(ahx_ctx *ctx);

/* ------------------------------------------------------------------------- *
 * CLM.GARun, evaluate.go:258-315 (eaopt ModGenerational + SelTournament{3} +
 * HallOfFame(1) + early stop), one whole generation per kernel launch, never
 * leaving HBM.  RNG is Philox4x32-10 keyed by (seed, generation, individual),
 * so the same (seed, npop, island) is bit-reproducible; it is not Go's
 * math/rand stream.
 * ------------------------------------------------------------------------- */
typedef struct ahx_ga_params {
  int32_t npop;          /* Optimizer.NPop   (--npop, base.go:67)                 */
  int32_t ngen;          /* Optimizer.NGen   (--ngen): generations w/o improvement */
  int64_t max_gens;      /* ga.NGenerations  (1000000, evaluate.go:270)           */
  double mut_prob;       /* Optimizer.MutProb (--mutapb, base.go:71)              */
  uint64_t seed;         /* Optimizer.Seed   (--seed)                             */
  int32_t tournament_k;  /* SelTournament.NContestants (3, evaluate.go:273)       */
  int32_t log_every;     /* callback period   (500, evaluate.go:294)              */
  int32_t phase;         /* label only        (GA%d, optimize.go:64)              */
  int32_t island;        /* this island's id = its RNG stream (0 for a single GPU) */
  /* Island model.  eaopt's multi-population switch, which the reference pins to one
   * population (ga.NPops = 1, evaluate.go:269): n_islands populations of npop individuals each, one
   * per GPU.  Every migrate_every generations each island sends its n_elite best individuals to every
   * other island, which replace that island's worst (see "Islands" below).  n_islands <= 1: single
   * population, the other two fields are ignored. */
  int32_t n_islands;     /* ga.NPops                                              */
  int32_t migrate_every; /* generations between exchanges (eaopt MigFrequency)    */
  int32_t n_elite;       /* individuals sent to each other island per exchange    */
  int32_t reserved;
} ahx_ga_params;

typedef struct ahx_ga_stats {
  double best_score;     /* -HallOfFame[0].Fitness                                */
  int64_t generations;   /* ga.Generations                                        */
  int64_t updated;       /* generation of the last improvement                    */
  int64_t evaluations;   /* Tour.Evaluate calls made on the device                */
  int32_t stopped;       /* 1 once EarlyStop / max_gens fired                     */
  int32_t reserved;
  double device_ms;      /* CUDA-event time spent in generation kernels           */
} ahx_ga_stats;

/* Invoked on the calling thread at generation 0 and every log_every
 * generations (evaluate.go:287-301) with the hall-of-fame tour. */
typedef void (*ahx_ga_callback)(void *user, int32_t phase, int64_t gen, double best_score,
                                const int32_t *best_tour, int32_t L);

void ahx_ga_default_params(ahx_ga_params *p);
int ahx_ga_run(ahx_ctx *ctx, const ahx_ga_params *params, int32_t L, const int32_t *init_tour,
               int32_t *best_tour, ahx_ga_stats *stats, ahx_ga_callback cb, void *user);
/* The same loop in pieces, for the island model (exchange between advances). */
int ahx_ga_begin(ahx_ctx *ctx, const ahx_ga_params *params, int32_t L, const int32_t *init_tour);
int ahx_ga_advance(ahx_ctx *ctx, int64_t n_gens, ahx_ga_stats *stats);
int ahx_ga_best(ahx_ctx *ctx, int32_t *best_tour, double *best_score);
/* n best individuals of the current population (fitness = -score, ascending). */
int ahx_ga_get_elites(ahx_ctx *ctx, int32_t n, int32_t *tours /* n*L */, double *fitness /* n */);
/* Replace the n worst individuals by migrants (tours validated as permutations
 * of the active set; fitness re-evaluated on the device). */
int ahx_ga_put_migrants(ahx_ctx *ctx, int32_t n, const int32_t *tours /* n*L */);
/* Test hook: re-derives every individual of the current generation from its stored tour
 * (permutation of the active contigs, fitness within 1e-12 of a fresh Tour.Evaluate, stored
 * coordinates equal to the tour's) and counts the individuals that fail. */
int ahx_ga_selfcheck(ahx_ctx *ctx, int64_t *n_bad);
/* ------------------------------------------------------------------------- *
 * Islands: the exchange step runs INSIDE the persistent GA kernel.  Every island owns a
 * mailbox in its GPU's memory; at generation g = e * migrate_every each island's kernel
 * stores its n_elite best individuals (tour row, coordinate row, fitness) and its best-ever
 * fitness straight into every other island's mailbox over NVLink peer memory, publishes a
 * release flag, waits for the other islands' flags of the same epoch e and replaces its
 * worst individuals (worst first) by the migrants in island order -- no host round trip, no
 * re-scoring (a score is a pure function of the tour), no kernel relaunch.  EarlyStop
 * (evaluate.go:304-306) becomes a collective decision taken at the exchange points from the
 * exchanged best-ever values: stop when the best over all islands has not improved for more
 * than ngen generations; every island takes the same decision at the same generation.
 * Same (seed, npop, n_islands, migrate_every, n_elite) => bit-identical run.
 *
 * Wiring, after ahx_ga_begin on every island (params.n_islands > 1) and before the first
 * ahx_ga_advance:
 *   one process per GPU:  ahx_ga_island_export on every island, all-gather the 64-byte
 *                         handles on the host (any transport), ahx_ga_island_connect with all
 *                         n_islands handles in island order (CUDA IPC).  The all-gather is
 *                         also the barrier that orders one run's mailbox after the previous one's.
 *   one process, n GPUs:  ahx_ga_island_connect_local (peer access), or simply
 *                         ahx_ga_run_islands, which does begin / connect / advance / end for
 *                         all islands from the calling thread.
 * All islands must then be advanced by the same number of generations (their kernels wait
 * for each other at every exchange point; a peer that never shows up is an AHX_ERR_STATE
 * after AHX_ISLAND_TIMEOUT_MS, default 20000).  Requires the persistent GA kernel
 * (ahx_ga_kernel() >= 2); ahx_ga_get_elites / ahx_ga_put_migrants remain for a
 * host-mediated exchange on the other paths.
 * ------------------------------------------------------------------------- */
#define AHX_ISLAND_HANDLE_BYTES 64
int ahx_ga_island_export(ahx_ctx *ctx, void *handle /* AHX_ISLAND_HANDLE_BYTES */);
int ahx_ga_island_connect(ahx_ctx *ctx, const void *handles /* n_islands x AHX_ISLAND_HANDLE_BYTES */);
int ahx_ga_island_connect_local(ahx_ctx **ctxs /* n_islands contexts of this process, island order */, int32_t n);
/* CLM.GARun as an island model on n GPUs driven by the calling thread: ctxs[i] (one per GPU, the
 * same group loaded in each) runs island i with params->npop individuals.  The callback gets the
 * best tour over all islands; best_tour / stats describe the best island's hall of fame
 * (stats->evaluations and device_ms: summed / max over islands). */
int ahx_ga_run_islands(ahx_ctx **ctxs, int32_t n, const ahx_ga_params *params, int32_t L, const int32_t *init_tour,
                       int32_t *best_tour, ahx_ga_stats *stats, ahx_ga_callback cb, void *user);
/* Exchanges completed by the active run's kernel so far (0 without islands). */
int64_t ahx_ga_island_epochs(ahx_ctx *ctx);

/* Which generation loop the active run uses: 3 persistent kernel with the pair table streamed,
 * 2 persistent kernel with the table in shared memory, 1 two kernels per generation (32-bit
 * coordinates), 0 the fp64 kernels (per-contig arrays in shared memory, or -- groups beyond ~26 000
 * contigs -- tour rows and coordinates in global memory: any group size runs); -1 when no run is active. */
int ahx_ga_kernel(ahx_ctx *ctx);
int ahx_ga_end(ahx_ctx *ctx);

/* One mutation operator applied to one tour on the device (test hook for
 * evaluate.go:157-218): op 0 MutPermute(p,q), 1 MutSplice(k=p), 2 MutInsertion
 * (p,q; dir!=0 moves q to p), 3 MutInversion(p,q). */
int ahx_mutate(ahx_ctx *ctx, int32_t L, const int32_t *tour_in, int32_t op, int32_t p, int32_t q,
               int32_t dir, int32_t *tour_out);

/* Test hooks for the code the persistent GA kernel actually runs:
 * ahx_mutate_remap     the same mutation through the kernel's operator code -- remap_of() turns the
 *                      operator into a branch-free index map for the tour row and an elementwise map
 *                      of the parent's stored coordinates (no prefix sum); coords_out (optional,
 *                      n_tigs words) receives 2*mid (evaluate.go:120-126) of every contig of the
 *                      child as derived that way.  Needs a loaded group and a full tour (L = n_tigs).
 * ahx_test_plan_mutations  host-side evaluation of the kernels' mutation plan for offspring 0..n-1 of
 *                      generation gen (eaopt's rng.Float64() < MutRate, then Tour.Mutate's draw,
 *                      evaluate.go:221-232, randomTwoInts, evaluate.go:146-154):
 *                      out[5*o..] = mutated, op (-1 none), p, q, dir.
 * ahx_test_tournaments winners[2*pair + which] of the kernels' SelTournament{k} for pairs
 *                      0..n_pairs-1 of generation gen over the fitness array fit[P] (device). */
int ahx_mutate_remap(ahx_ctx *ctx, int32_t L, const int32_t *tour_in, int32_t op, int32_t p, int32_t q,
                     int32_t dir, int32_t *tour_out, uint32_t *coords_out);
int ahx_test_plan_mutations(uint64_t seed, int32_t island, int64_t gen, int32_t n, double mut_prob, int32_t L,
                            int32_t *out /* n x 5 */);
int ahx_test_tournaments(ahx_ctx *ctx, uint64_t seed, int32_t island, int64_t gen, int32_t P, const double *fit,
                         int32_t k, int32_t n_pairs, int32_t *winners /* n_pairs x 2 */);

/* ------------------------------------------------------------------------- *
 * Orientation heuristics, orientation.go:31-120.  signs (n_tigs bytes) is
 * updated in place; *accepted = 1 for ACCEPT, 0 for REJECT.  ahx_flip_one
 * requires every byte to be '+' or '-' (all CLM.Signs ever holds).
 * ------------------------------------------------------------------------- */
int ahx_flip_whole(ahx_ctx *ctx, int32_t L, const int32_t *tour, uint8_t *signs,
                   double *old_score, double *new_score, int32_t *accepted);
/* trace (optional, L x 2): score before / score after each single flip;
 * step_accepted (optional, L bytes). */
int ahx_flip_one(ahx_ctx *ctx, int32_t L, const int32_t *tour, uint8_t *signs, double *final_score,
                 int64_t *n_accepts, int64_t *n_rejects, double *trace, uint8_t *step_accepted,
                 int32_t *accepted);
int ahx_flip_all(ahx_ctx *ctx, int32_t L, const int32_t *tour, uint8_t *signs,
                 double *old_score, double *new_score, int32_t *accepted);

/* How the last ahx_flip_all on this context found the top eigenvector of O (orientation.go:33-44): relative
 * residual |O v - lambda v| / |lambda| of the returned vector (0 for the dense decomposition), *dense = 1
 * when the dense Jacobi path ran (groups of <= 48 contigs, or Lanczos not converged and <= 1500 contigs),
 * *converged = 0 when Lanczos stopped at its restart limit (near-degenerate top eigenvalues on a large group:
 * the sign pattern may then differ from a dense decomposition's; the host logs it). */
int ahx_flip_all_info(const ahx_ctx *ctx, double *residual, int32_t *converged, int32_t *dense, int32_t *matvecs);

/* ------------------------------------------------------------------------- *
 * Roofline probes (measured live by bench.py).
 * ------------------------------------------------------------------------- */
int ahx_probe_fp64(ahx_ctx *ctx, double *dfma_tflops, double *ddiv_gops);
int ahx_probe_l2(ahx_ctx *ctx, uint64_t bytes, double *gbs);
/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
int64_t ahx_launch_count(const ahx_ctx *ctx);
int ahx_sm_count(const ahx_ctx *ctx);
/* CUDA-event time (ms) of the kernels of the last ahx_eval_q / ahx_flip_one call on this context:
 * host<->device copies excluded (bench.py's roofline legs of the orientation phase). */
double ahx_last_device_ms(const ahx_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* AHX_H_ */
