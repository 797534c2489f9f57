/*
 * ngu_epi.h — C ABI of the B200 episodic-novelty engine (libngu_epi.so).
 *
 * This is the drop-in boundary for the per-step episodic-novelty hot path of
 * minoring/NGU.  The reference has no FFI of its own (it is pure Python over
 * torch eager ops), so each entry point cites the reference Python interface
 * it replaces; paths are relative to the reference checkout.
 *
 * Conventions
 *   - plain C, no torch / C++ types in any signature;
 *   - every function returns NGU_OK (0) or a negative NGU_E* code;
 *     ngu_epi_last_error() gives the text for the last failure on a handle
 *     (or, with a NULL handle, for the last failed ngu_epi_create);
 *   - pointers named *_dev are device pointers on the handle's GPU, *_host are
 *     host pointers; `stream` is a cudaStream_t passed as void* (NULL = the
 *     legacy default stream).  Calls that take a stream only enqueue work and
 *     return without synchronising; *_host calls are synchronous;
 *   - one handle per GPU / per shard of actors; a handle is not thread-safe;
 *   - every call runs on the handle's GPU and restores the caller's current CUDA device on return;
 *   - stream contract: *_host calls run on a private non-blocking stream of the handle.  They are
 *     ordered after whatever the caller enqueued THROUGH THIS HANDLE on the stream it last passed
 *     (append, device-mask reset, device-pointer steps).  Work enqueued on a stream by other means
 *     (e.g. a torch copy into the memory view) must be synchronised by the caller first;
 *   - sticky error: the k-NN merge and the multi-GPU exchange wait for data with a (generous) time
 *     budget.  If one ever gives up, the results would be silently wrong, so the library fails loudly
 *     instead: ngu_epi_state.exchange_timeouts becomes non-zero and stays so, every reward produced
 *     from then on is NaN, and ngu_epi_scan / _finish / _step* return NGU_ESTATE.  Recreate the handle;
 *   - there is no CPU fallback: without a CUDA device every call fails.
 *
 * Data layout in HBM (owned by the handle unless `memory_dev` is supplied):
 *   episodic memory  float [n_actors][capacity][32]   (actor-major; each
 *   actor's ring is one contiguous stream of 128-byte rows).  The reference's
 *   tensor is slot-major (capacity, n_actors, 32)
 *   (episodic_novelty.py:32-33); ngu_epi_memory() returns the strides that
 *   present this buffer as that (capacity, n_actors, 32) view.
 */
#ifndef NGU_EPI_H
#define NGU_EPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGU_EPI_ABI_VERSION 5
#define NGU_EPI_DIM 32          /* controllable_state_dim, episodic_novelty.py:23 */
#define NGU_EPI_MAX_K 32        /* k-NN list lives in one warp */

enum {
    NGU_OK = 0,
    NGU_EINVAL = -1,            /* bad argument / unsupported configuration */
    NGU_ECUDA = -2,             /* a CUDA runtime / driver call failed */
    NGU_ENOMEM = -3,
    NGU_ESTATE = -4             /* call order violated (e.g. finish before scan), or the sticky error above */
};

enum {
    NGU_MODE_REFERENCE = 0,     /* what minoring/NGU computes: k FARTHEST slots, Euclidean
                                   (sqrt) distance, per-rank running mean */
    NGU_MODE_PAPER = 1          /* NGU paper: k NEAREST slots, squared distance */
};

typedef struct ngu_epi ngu_epi;

/* Mirrors the arguments of EpisodicNovelty.__init__ (episodic_novelty.py:9-35)
 * and the model_hypr keys it reads (models/model_hypr.py:52-57). */
typedef struct ngu_epi_cfg {
    int32_t n_actors;           /* actors held by THIS handle (local shard)          */
    int32_t capacity;           /* episodic_memory_capacity (N slots per actor)      */
    int32_t dim;                /* must be NGU_EPI_DIM                               */
    int32_t k;                  /* num_neighbors, 1..NGU_EPI_MAX_K, k <= capacity    */
    int32_t mode;               /* NGU_MODE_*                                        */
    int32_t device;             /* CUDA device ordinal                               */
    float kernel_epsilon;       /* model_hypr['kernel_epsilon']                      */
    float cluster_distance;     /* model_hypr['cluster_distance']                    */
    float pseudo_counts;        /* model_hypr['kernel_pseudo_counts_constant']       */
    float max_similarity;       /* model_hypr['kernel_maximum_similarity']           */
    float max_reward_scale;     /* IntrinsicNovelty.max_rew = L (intrinsic_novelty.py:16) */
    float reserved0;
    double prior_count;         /* initial RunningMeanStd.count of ed_rms; the reference's
                                   RunningMeanStd((k,)) makes this k (mpi_util.py:38-42) */
    int32_t scan_ctas_per_sm;   /* 0 = default; tuning knobs, see DESIGN.md          */
    int32_t scan_variant;       /* 0 = default                                       */
} ngu_epi_cfg;

/* Running statistics + ring cursor; checkpoint / parity injection.
 * ed_* mirror EpisodicNovelty.ed_rms (episodic_novelty.py:28), epi_* mirror
 * epi_novel_rms (:29), intr_* mirror IntrinsicNovelty.intrinsic_novel_rms
 * (intrinsic_novelty.py:17); cursor mirrors memory_idx (episodic_novelty.py:34). */
typedef struct ngu_epi_state {
    double ed_mean[NGU_EPI_MAX_K];
    double ed_var[NGU_EPI_MAX_K];
    double ed_count;
    double epi_mean, epi_var, epi_count;
    double intr_mean, intr_var, intr_count;
    int64_t cursor;
    int64_t steps;              /* finished steps since create / set_state */
    int64_t exchange_timeouts;  /* read-only, sticky: k-NN merges / p2p exchange rounds that gave up waiting (0 = healthy);
                                   ngu_epi_set_state ignores it */
    double ll_mean, ll_var, ll_count;   /* LifelongNovelty.ll_rms (lifelong_novelty.py:26), used by ngu_epi_step_rnd */
} ngu_epi_state;

/* Number of doubles exchanged between shards per step:
 * [0,k) sum_b dist_j (f32-exact), [k,2k) sum_b dist_j^2, [2k] actor count. */
#define NGU_EPI_RANK_SUMS(k) (2 * (k) + 1)
/* Log-only sums: {sum r_epi, sum r_epi^2, sum r_int, sum r_int^2, count}. */
#define NGU_EPI_LOG_SUMS 5

int ngu_epi_abi_version(void);

/* EpisodicNovelty.__init__ (episodic_novelty.py:9-35): allocates the zeroed
 * memory (unless memory_dev != NULL: caller-owned float[n_actors*capacity*32],
 * 128-byte aligned, must already be zeroed or hold valid rows), cursor 0,
 * running statistics at their priors. */
int ngu_epi_create(const ngu_epi_cfg* cfg, float* memory_dev, ngu_epi** out);
void ngu_epi_destroy(ngu_epi* h);
const char* ngu_epi_last_error(const ngu_epi* h);

/* EpisodicNovelty.episodic_memory (episodic_novelty.py:32-33, read by
 * intrinsic_novelty.py:39-40): base pointer + element strides of the
 * (capacity, n_actors, 32) view. */
int ngu_epi_memory(ngu_epi* h, float** memory_dev, int64_t strides[3]);

/* episodic_novelty.py:51-54 — distance scan over all `capacity` slots of every
 * actor + top-k (kernel A + merge).  q_dev: float[n_actors][32], the controllable
 * states returned by Embedding (:47).  Leaves the per-actor k-NN (d^2, slot)
 * lists and this shard's rank sums in the handle. */
int ngu_epi_scan(ngu_epi* h, const float* q_dev, void* stream);

/* Device pointer to this shard's double[NGU_EPI_RANK_SUMS(k)], valid after
 * ngu_epi_scan on the same stream.  This is the buffer the reference reduces
 * with MPI.Allreduce (mpi_util.py:7-18 via episodic_novelty.py:56); with several
 * shards, sum it across shards (NCCL) and pass the result to ngu_epi_finish. */
int ngu_epi_rank_sums(ngu_epi* h, double** sums_dev);

/* episodic_novelty.py:56-78 + intrinsic_novelty.py:26-31 — running-mean update
 * from the (global) rank sums, normalise, cluster clamp, inverse kernel,
 * similarity, reward, RND multiplier clipped to [1, L].
 *   sums_global_dev  NULL = use this handle's own rank sums (single shard)
 *   alpha_dev        float[n_actors] RND modulator (lifelong_novelty.py:57), NULL = 1
 *   r_int_dev        float[n_actors] out: r_episodic * clip(alpha, 1, L)
 *   r_epi_dev        float[n_actors] out or NULL: r_episodic (what
 *                    compute_episodic_novelty returns, episodic_novelty.py:84)
 *   log_sums_dev     double[NGU_EPI_LOG_SUMS] out or NULL: this shard's batch sums
 *                    for the log-only statistics (see ngu_epi_log_update). */
int ngu_epi_finish(ngu_epi* h, const double* sums_global_dev, const float* alpha_dev,
                   float* r_int_dev, float* r_epi_dev, double* log_sums_dev, void* stream);

/* episodic_novelty.py:82 + intrinsic_novelty.py:32 — log-only RunningMeanStd
 * updates (epi_novel_rms, intrinsic_novel_rms) from (globally summed) log sums. */
int ngu_epi_log_update(ngu_epi* h, const double* log_sums_global_dev, void* stream);

/* EpisodicNovelty.insert_state (episodic_novelty.py:86-89): memory[cursor,:,:] = q;
 * cursor = (cursor+1) % capacity.  The cursor lives on the device so the step
 * can be replayed from a CUDA graph. */
int ngu_epi_append(ngu_epi* h, const float* q_dev, void* stream);

/* IntrinsicNovelty.reset_memory_if_done (intrinsic_novelty.py:36-40):
 * zero every slot of each actor b with done_dev[b] != 0; cursor untouched. */
int ngu_epi_reset(ngu_epi* h, const uint8_t* done_dev, void* stream);

/* One whole step in two launches: the scan kernel (distance scan + top-k + merge), then ONE epilogue
 * kernel = rank sums -> [p2p exchange over NVLink when connected, see below] -> running-mean update ->
 * rewards x clip(alpha, 1, L) -> append; the log-only trackers absorb the step's rewards at the start of
 * the NEXT fused launch (or on ngu_epi_get_state / ngu_epi_log_flush).  I.e. one call of
 * EpisodicNovelty.compute_episodic_novelty followed by the multiplier of
 * IntrinsicNovelty.compute_intrinsic_novelty (episodic_novelty.py:37-84, intrinsic_novelty.py:20-34),
 * all on `stream`.  r_int is float32: the reference's product of a float32 reward and numpy's float64
 * alpha is float64 under numpy >= 2 (intrinsic_novelty.py:31); the values agree to float32 rounding. */
int ngu_epi_step(ngu_epi* h, const float* q_dev, const float* alpha_dev,
                 float* r_int_dev, float* r_epi_dev, void* stream);

/* ngu_epi_step for callers that enqueue steps BACK TO BACK with inputs that are already on the device (a sequence of
 * precomputed controllable states; bench.py's device-resident loop): the step is software-pipelined with the previous
 * one.  Step s+1 depends on step s only through the one slot per actor that step s appends to
 * (episodic_novelty.py:80,86-89 - the query at :51-54 reads the memory, the append writes one row), so this call's
 * scan kernel does not wait for the previous step's epilogue: it streams while that step's k-NN merge tail, its
 * reward epilogue and (connected ranks) its NVLink exchange are still running, and takes the content of the slot
 * being appended from the previous call's q_dev.  Results are identical to ngu_epi_step's, bit for bit.
 * The caller promises, for a call that directly follows another ngu_epi_step / ngu_epi_step_pipelined on the same
 * stream (any other call on the handle in between, a different stream, or a stream capture makes this call behave
 * exactly like ngu_epi_step):
 *   - nothing else was enqueued on `stream` since that previous call;
 *   - q_dev / alpha_dev of THIS call were complete before the previous call was made (or are written by work that
 *     was enqueued on `stream` before it) - they may be read as soon as the previous step starts;
 *   - the previous call's q_dev stays intact until this step has completed;
 *   - r_int_dev / r_epi_dev are written in step order, as with ngu_epi_step. */
int ngu_epi_step_pipelined(ngu_epi* h, const float* q_dev, const float* alpha_dev,
                           float* r_int_dev, float* r_epi_dev, void* stream);

/* ngu_epi_step with the scalar tail of LifelongNovelty.compute_lifelong_curiosity fused in
 * (lifelong_novelty.py:53-57): instead of a ready-made modulator the caller passes the RND
 * predictor / target features float[n_actors][feat_dim] (feat_dim must be 128,
 * rnd_prediction.py:27); the library computes err = mean((pred-targ)^2, axis=1), updates ll_rms
 * (use_mpi=False: local batch mean/std, mpi_util.py:51-52), alpha = 1 + (err - mean)/sqrt(var),
 * and continues as ngu_epi_step.  alpha_out_dev (float[n_actors], may be NULL) receives alpha.
 * No device->host sync is needed between the RND forward and the reward. */
int ngu_epi_step_rnd(ngu_epi* h, const float* q_dev, const float* pred_dev, const float* targ_dev, int32_t feat_dim,
                     float* r_int_dev, float* r_epi_dev, float* alpha_out_dev, void* stream);

/* Same, with HOST buffers (what the Python drop-in calls with CPU tensors):
 * H2D of q/alpha (from a pinned staging buffer: a PDL-chained kernel pulls it over PCIe on large
 * shapes, the copy engine on small ones), the step, D2H of the rewards (written by the epilogue
 * straight into mapped host memory), synchronous. */
int ngu_epi_step_host(ngu_epi* h, const float* q_host, const float* alpha_host,
                      float* r_int_host, float* r_epi_host);
int ngu_epi_reset_host(ngu_epi* h, const uint8_t* done_host);

/* Last scan's k-NN lists, best first: d2_host float[n_actors][k] (squared
 * distance, the exact bits of the reference's ((mem-q)**2).sum(-1)),
 * idx_host int32[n_actors][k] (slot index; ties -> lowest slot).  Synchronous.
 * The reference discards the indices (episodic_novelty.py:54-56); they are part
 * of this library's contract. */
int ngu_epi_topk_host(ngu_epi* h, float* d2_host, int32_t* idx_host);

/* Checkpoint / parity injection of the running statistics and the ring cursor.  Synchronous.
 * The log-only trackers (epi_*, intr_*: episodic_novelty.py:82, intrinsic_novelty.py:32 - never read by
 * the hot path) are kept off the step's critical path: a fused step (ngu_epi_step*, ngu_epi_step_host)
 * stashes its rewards on the device and the next fused step folds them in while its scan runs.
 * ngu_epi_get_state folds an outstanding stash in first, so what it returns is always up to date
 * (single GPU; connected ranks call ngu_epi_log_flush first, below).
 * ngu_epi_set_state replaces the PUBLIC fields only (an outstanding stash is folded in first, i.e. into the
 * values being replaced).  The library's bookkeeping - log-only sums waiting for the next exchange round, the
 * exchange sequence numbers, the sticky exchange_timeouts counter - is never touched, so a get/modify/set
 * round trip loses nothing and a set_state on one rank cannot desynchronise the exchange.  With connected
 * ranks the public fields are replicated state: set the same values on every rank.
 * ngu_epi_set_cursor moves only the ring cursor (EpisodicNovelty.memory_idx, episodic_novelty.py:34). */
int ngu_epi_get_state(ngu_epi* h, ngu_epi_state* out);
int ngu_epi_set_state(ngu_epi* h, const ngu_epi_state* in);
int ngu_epi_set_cursor(ngu_epi* h, int64_t cursor);

/* Bulk load / dump of the episodic memory in the REFERENCE layout
 * float[capacity][n_actors][32] on the host (tests, checkpoints).  Synchronous. */
int ngu_epi_load_memory_host(ngu_epi* h, const float* slot_major_host);
int ngu_epi_dump_memory_host(ngu_epi* h, float* slot_major_host);

/* Multi-GPU, one process per GPU: replace the caller's all-reduce by an exchange over NVLink peer
 * memory fused into the finish kernel (what MPI.Allreduce does at mpi_util.py:17).
 *   1. every rank:  ngu_epi_p2p_export(h, handle [64 bytes out], world)
 *   2. all-gather the 64-byte handles in rank order (any transport)
 *   3. every rank:  ngu_epi_p2p_connect(h, rank, world, handles [world*64 bytes])
 * Afterwards ngu_epi_step / ngu_epi_step_rnd / ngu_epi_step_host sum the rank sums over all ranks
 * inside the epilogue kernel (one exchange per step); all ranks must call them in lockstep.  Ranks
 * must sit on DISTINCT GPUs (the kernel polls words written by the peers).  The log-only sums
 * (epi_novel_rms, intrinsic_novel_rms; never read by the hot path) ride on the NEXT step's
 * exchange, so those two trackers lag one step; ngu_epi_log_flush (collective: every rank, same
 * point of its stream) brings them up to date before they are read.  Unconnected, it folds the
 * outstanding reward stash into the trackers on `stream` (what ngu_epi_get_state does synchronously). */
int ngu_epi_p2p_export(ngu_epi* h, void* ipc_handle_out, int32_t world);
int ngu_epi_p2p_connect(ngu_epi* h, int32_t rank, int32_t world, const void* ipc_handles);
int ngu_epi_log_flush(ngu_epi* h, void* stream);

/* Introspection for bench.py / DESIGN.md: kernels launched by the handle so far,
 * and the scan launch geometry {grid, block, dyn_smem, static tiles per warp, stages, tile_rows,
 * dynamic chunk tiles (0 = static schedule), chunks in the schedule}. */
int64_t ngu_epi_launch_count(const ngu_epi* h);
int ngu_epi_scan_geometry(const ngu_epi* h, int32_t out[8]);
/* Test aid: tile range [out[0], out[1]) of chunk `idx` (0 <= idx < chunks) of the scan kernel's work
 * schedule, in units of tiles of tile_rows memory slots over the linear space actor * tiles_per_actor +
 * tile; the ranges of all chunks partition that space (tests/test_gpu_parity.py). */
int ngu_epi_debug_chunk(const ngu_epi* h, int32_t idx, int32_t out[2]);
/* Test aid: make the handle look as if a merge / exchange round had timed out (sets the sticky counter on the
 * device), so that the failure path can be exercised: NaN rewards, NGU_ESTATE from the next step. */
int ngu_epi_debug_inject_timeout(ngu_epi* h);
/* The same plan WITHOUT a handle or a GPU (host logic only; the CPU test suite checks the partition for
 * many shapes): geo[8] as ngu_epi_scan_geometry for a device with `sm_count` SMs; chunks (may be NULL)
 * receives 2 * geo[7] ints, [first tile, end tile) per chunk; max_chunks = capacity of `chunks` in chunks. */
int ngu_epi_plan_host(const ngu_epi_cfg* cfg, int32_t sm_count, int32_t geo[8], int32_t* chunks, int32_t max_chunks);

/* Measurement aid (bench.py "roofline"): bracket the next `max_calls` launches of
 * the scan kernel (kernel A alone, not the merge) with CUDA events on the launching
 * stream; _end synchronises and returns the summed device time and the number of
 * launches timed. */
/* Debug aid (tools/scan_timeline.py; handle created with NGU_EPI_DEBUG_TIMELINE=1 in the environment):
 * per-warp records of 8 words {kernel start, first tile landed, scan done, flush/merge done (%globaltimer, ns),
 * SM id, 3 unused} of the last scan launch (out_host: max_warps * 8 words); returns the number of warps
 * copied (<= max_warps) or a negative error. */
int ngu_epi_debug_timeline(ngu_epi* h, uint64_t* out_host, int32_t max_warps);
int ngu_epi_scan_timing_begin(ngu_epi* h, int32_t max_calls);
int ngu_epi_scan_timing_end(ngu_epi* h, double* total_ms, int32_t* calls);

#ifdef __cplusplus
}
#endif
#endif /* NGU_EPI_H */
