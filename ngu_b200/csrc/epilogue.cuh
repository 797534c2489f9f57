// epilogue.cuh — kernel B (rank sums, running-mean update, reward epilogue, append in
// one single-CTA launch), plus the standalone append / reset / layout kernels.
//
// Reference lines replaced (ngu/models/intrinsic_novelty/...):
//   episodic_novelty.py:54      topk (final merge of the per-warp lists)
//   episodic_novelty.py:52      .sqrt() — applied to the k winners only
//   episodic_novelty.py:56      ed_rms.update(...)          -> rank sums + finish
//   ngu/utils/mpi_util.py:7-73  mpi_mean / mpi_moments / RunningMeanStd
//   episodic_novelty.py:58-78   normalise, cluster clamp, inverse kernel, similarity, reward
//   intrinsic_novelty.py:26-31  RND multiplier clipped to [1, L], product
//   episodic_novelty.py:86-89   insert_state
//   intrinsic_novelty.py:36-40  reset_memory_if_done
#pragma once
#include <cstdint>

#include "p2p_exchange.cuh"
#include "warp_topk.cuh"

namespace ngu {

// Device-resident running statistics (ngu_epi_state without the host mirror).
struct DevState {
    double ed_mean[32];
    double ed_var[32];
    double ed_count;
    double epi_mean, epi_var, epi_count;
    double intr_mean, intr_var, intr_count;
    long long steps;     // the kStagedWords words from ed_count to stash_open are staged together by the epilogue
    double ll_mean, ll_var, ll_count;   // LifelongNovelty.ll_rms (lifelong_novelty.py:26,54), RND fused path only
    // ---- internal from here on: ngu_epi_set_state never touches these ----
    double pending_log[5];              // p2p mode: this shard's log-only sums of the last step, not yet exchanged
    long long xchg_seq;                 // p2p mode: sequence number (tag) of the next per-step exchange round; NOT the public
                                        // `steps`, so that a set_state on one rank cannot desynchronise the tags
    long long cursor;                   // (public, but staged with the block)
    long long exchange_timeouts;   // merges / p2p exchange rounds that gave up waiting (must stay 0; sticky error, see below)
    long long stash_open;          // 1: log_stash holds the rewards of a fused step not yet folded into the trackers
    long long flush_seq;                // sequence number of on-demand log flush rounds
};
constexpr int kStagedWords = 20;   // ed_count .. stash_open
// indices into the staged block
enum : int { SC_ED_COUNT = 0, SC_EPI = 1, SC_INTR = 4, SC_STEPS = 7, SC_LL = 8, SC_PENDING = 11, SC_XCHG_SEQ = 16, SC_CURSOR = 17,
             SC_TIMEOUTS = 18, SC_STASH_OPEN = 19 };

struct EpiConsts {
    float eps, cluster, pseudo, s_max, max_rew;
    int32_t k, n_actors, capacity, largest, sqrt_dist;
};

// ---------------------------------------------------------------------------
// This kernel runs ONE CTA once per step; its cost is dependent-instruction latency, not
// throughput and not instruction fetch (measured with %globaltimer stamps, tools/epilogue_phases.py:
// a dry run of the whole post-wait body right before the real one does not make the real one
// faster, and out-of-line IEEE helpers were ~0.25 us slower than inlined ones).  So: the chain
// to the rewards is kept short (log-only work runs in other warps or in the next launch, see
// log_absorb), loads are issued as one batch, and the helpers below are inlined.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float sqdiff_rn(float a, float b) { const float d = __fsub_rn(a, b); return __fmul_rn(d, d); }
__device__ __forceinline__ float f32_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float f32_rcp(float a) { return __frcp_rn(a); }
__device__ __forceinline__ float f32_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double f64_div(double a, double b) { return a / b; }

// ---------------------------------------------------------------------------
// numpy's float32 pairwise summation (x.sum(axis=0), mpi_util.py:11) over one
// column v[b*stride], b = 0..n-1, evaluated cooperatively by an aligned group of
// 8 lanes (lane8 = lane & 7).  numpy: leaves of <= 128 elements use eight strided
// accumulators r[0..7] combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and a
// sequential tail; larger n splits at n/2 rounded down to a multiple of 8,
// recursively.  The recursion is flattened on the host into a postfix plan of
// (offset, length) leaves and (-1, 0) "add the two top values" entries
// (build_pairwise_plan in ngu_epi.cu).  Lane a owns accumulator r[a]; the combine is
// three xor-shuffles (float add commutes, so every lane ends with the same bits).
// All 32 lanes of a warp must call this together.
// ---------------------------------------------------------------------------
__device__ __noinline__ float np_pairwise_leaf8(const float* v, int stride, int n, int lane8) {
    float r;
    int i;
    if (n < 8) {
        r = -0.0f;
        i = 0;
    } else {
        const int lim = n - (n % 8);
        r = v[lane8 * stride];
#pragma unroll 4
        for (i = 8; i < lim; i += 8) r = __fadd_rn(r, v[(i + lane8) * stride]);
        r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 4));
    }
#pragma unroll 1
    for (; i < n; ++i) r = __fadd_rn(r, v[i * stride]);
    return r;
}

constexpr int kPlanMaxDepth = 24;
constexpr int kPlanMaxLeaves = 256;                     // n <= 128 * 256 rows per column sum
constexpr int kPlanMaxEntries = 2 * kPlanMaxLeaves - 1; // postfix entries; the leaf list follows them in the plan buffer

__device__ __forceinline__ float np_pairwise_sum8(const float* v, int stride, const int2* plan, int plan_len,
                                                  int lane8, float* stack /* [kPlanMaxDepth] per thread */) {
    int sp = 0;
#pragma unroll 1
    for (int i = 0; i < plan_len; ++i) {
        const int2 e = plan[i];
        if (e.y > 0) {
            stack[sp++] = np_pairwise_leaf8(v + e.x * stride, stride, e.y, lane8);
        } else {
            --sp;
            stack[sp - 1] = __fadd_rn(stack[sp - 1], stack[sp]);
        }
    }
    return stack[0];
}

// ---------------------------------------------------------------------------
// Kernel B (epilogue): ONE CTA, launched with programmatic dependent launch right
// behind the scan kernel, so it is already resident when the scan drains.
// Phases (bitmask, so the same code serves the single-shard step and the split
// scan | exchange | finish sequence of a multi-GPU step):
//   RANKSUMS  column sums of this shard's (B,k) k-NN distance matrix
//             = mpi_mean's localsum (mpi_util.py:11-15)
//   FINISH    RunningMeanStd update from the (global) sums (mpi_util.py:48-73),
//             per-actor reward (episodic_novelty.py:58-78), RND multiplier clip and
//             product (intrinsic_novelty.py:26-31), log-only batch sums
//   APPEND    insert_state (episodic_novelty.py:86-89)
// ---------------------------------------------------------------------------
//   EXCHANGE  (multi-GPU, p2p mode) sum the rank sums - and, after the rewards, the log-only
//             sums - over all shards through NVLink peer stores (p2p_exchange.cuh)
//   RND       (with FINISH) the scalar tail of LifelongNovelty.compute_lifelong_curiosity
//             (lifelong_novelty.py:53-57): err = mean_128((pred-targ)^2), ll_rms update, alpha =
//             1 + (err - mean)/sqrt(var) - replaces the alpha input; the per-actor error is
//             computed while the scan is still running
enum : int { PH_RANKSUMS = 1, PH_FINISH = 2, PH_APPEND = 4, PH_LOGFUSE = 8, PH_EXCHANGE = 16, PH_RND = 32 };

struct EpilogueParams {
    const float* dist;        // [B][k]  sqrt(d2) (reference) or d2 (paper), written by the scan kernel
    double* rank_sums;        // [2k+1]  out (RANKSUMS)
    const double* sums_in;    // [2k+1]  global sums for FINISH; null = the sums computed in this launch
    const float* alpha;       // [B] or null
    float* r_int;             // [B]
    float* r_epi;             // [B] or null
    double* log_sums;         // [5] or null
    const float* q;           // [B][32] (APPEND)
    float* memory;            // [B][N][32] (APPEND)
    DevState* state;
    const int2* plan;         // pairwise-sum plan for n = B: plan_len postfix entries, then the n_leaves leaves in order
    int32_t plan_len;
    int32_t phases;
    int32_t stage_q;          // 1: q fits the dynamic shared memory of this launch
    int32_t n_leaves;
    const float* rnd_pred;    // [B][128] RND predictor features (PH_RND)
    const float* rnd_targ;    // [B][128] RND target features
    float* alpha_out;         // [B] or null: the modulator computed by the RND phase
    P2PView p2p;              // used when phases & PH_EXCHANGE
    float* log_stash;         // [2][B] r_epi, r_int of the last fused step (PH_LOGFUSE), see log_absorb
    const unsigned int* arm_go;       // pre-armed launch (ngu_epi_step_host): device word, arm_seq << 1 | cancelled once the stage
    unsigned int arm_seq;             // kernel has decided (cancelled = nobody rang the doorbell in time); 0 = not pre-armed
    unsigned int* epi_done;   // device word: receives `seq` after this launch's last write (pipelined steps poll it)
    unsigned int seq;         // host sequence number of this fused step; 0 = none (split sequence, graph capture)
    unsigned int wait_prev;   // pipelined step: sequence number of the PREVIOUS fused step, whose epilogue may still be running
                              // when this CTA starts - its stash is absorbed only once that epilogue has completed; 0 = the
                              // previous epilogue is complete by construction (ordinary launches, see log_absorb)
    unsigned int* host_err;   // mapped host word: set to 1 when a merge / exchange timed out (sticky error, never cleared)
    unsigned int* host_flag;  // mapped host word or null: receives host_seq after the kernel's last write
    unsigned int host_seq;
    unsigned long long* dbg_stamps;   // debug (NGU_EPI_DEBUG_EPILOGUE=1): %globaltimer at phase boundaries, thread 0
};

// Chan/Welford merge of (mean, var, count) with a batch (mpi_util.py:59-73), all f64.
__device__ __noinline__ void rms_merge(double* mean, double* var, double* count, double b_mean, double b_var,
                                       double b_count) {
    // every product and sum separately rounded (numpy does not contract): no FMA
    const double delta = __dsub_rn(b_mean, *mean);
    const double tot = __dadd_rn(*count, b_count);
    const double m_a = __dmul_rn(*var, *count), m_b = __dmul_rn(b_var, b_count);
    const double cross = f64_div(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), *count), b_count), tot);
    const double m2 = __dadd_rn(__dadd_rn(m_a, m_b), cross);
    *mean = __dadd_rn(*mean, f64_div(__dmul_rn(delta, b_count), tot));
    *var = f64_div(m2, tot);
    *count = tot;
}

// sc = {epi_mean, epi_var, epi_count, intr_mean, intr_var, intr_count} (the DevState order)
__device__ __noinline__ void log_rms_apply(double* sc, const double* ls) {
    const double n = ls[4];
    if (n > 0.0) {
        const double m1 = f64_div(ls[0], n), v1 = fmax(__dsub_rn(f64_div(ls[1], n), __dmul_rn(m1, m1)), 0.0);
        rms_merge(&sc[0], &sc[1], &sc[2], m1, v1, n);
        const double m2 = f64_div(ls[2], n), v2 = fmax(__dsub_rn(f64_div(ls[3], n), __dmul_rn(m2, m2)), 0.0);
        rms_merge(&sc[3], &sc[4], &sc[5], m2, v2, n);
    }
}
__device__ __forceinline__ long long st_steps_plus_one(double raw) { return __double_as_longlong(raw) + 1; }

// Tracker `which` (0: epi_novel_rms, 1: intrinsic_novel_rms) absorbs the batch described by
// ls = {sum r_epi, sum r_epi^2, sum r_int, sum r_int^2, count}; sc = the staged scalar block.
__device__ __forceinline__ void apply_log_sums(DevState* st, double* s_scal, const double* ls, int which) {
    const double n = ls[4];
    if (n > 0.0) {
        const int o = 2 * which;
        const double m = f64_div(ls[o], n), v = fmax(__dsub_rn(f64_div(ls[o + 1], n), __dmul_rn(m, m)), 0.0);
        double* sc = s_scal + 1 + 3 * which;
        rms_merge(&sc[0], &sc[1], &sc[2], m, v, n);
        for (int q = 0; q < 3; ++q) (&st->ed_count)[1 + 3 * which + q] = sc[q];
    }
}

// Fixed-order sum over the CTA of four f64 quantities: per-thread partials (thread t owns actors t,
// t+THREADS, ...) -> xor-shuffle tree per warp -> xor-shuffle tree over the warp partials.
// s_log[0..3] = the sums, s_log[4] = B.  All threads call it; ends with a barrier.
template <int THREADS>
__device__ __forceinline__ void log_tree_reduce(const double (&acc4)[4], double (*red)[THREADS / 32], double* s_log,
                                                int B) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double v = acc4[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double v = threadIdx.x < THREADS / 32 ? red[q][threadIdx.x] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if (threadIdx.x == 0) s_log[q] = v;
        }
        if (threadIdx.x == 0) s_log[4] = static_cast<double>(B);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// Deferred log-only trackers.  epi_novel_rms / intrinsic_novel_rms (episodic_novelty.py:82,
// intrinsic_novelty.py:32) are never read by the hot path, yet their batch sums and Chan merges
// (f64 shuffles and divisions, ~2 us of dependent latency) used to sit at the end of every
// step's critical path.  A PH_LOGFUSE launch now only stashes its per-actor rewards and raises
// DevState::stash_open; the NEXT fused launch folds the stash into the trackers before its
// griddepcontrol.wait, i.e. while its scan is still streaming (the CTA is resident 15-60 us
// early).  That is safe because a fused epilogue always follows its scan kernel, and the scan
// signals launch_dependents only AFTER its own griddepcontrol.wait: when this CTA starts, the
// previous epilogue (and everything else earlier in the stream) has completed.  Everything is
// device state, so a captured CUDA graph replays correctly.  ngu_epi_get_state and
// ngu_epi_log_flush run the same routine from log_absorb_kernel.  The sums are formed in the
// same fixed order as before (log_tree_reduce): the trackers hold the same bits, one launch later.
// All threads of the CTA call it.
// ---------------------------------------------------------------------------
template <int THREADS>
__device__ __noinline__ void log_absorb(DevState* st, const float* stash, int B, bool to_pending,
                                        double (*red)[THREADS / 32], double* s_log) {
    double acc4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int b = threadIdx.x; b < B; b += THREADS) {
        const float r = __ldcg(stash + b), ri = __ldcg(stash + B + b);
        acc4[0] += r; acc4[1] += static_cast<double>(r) * r;
        acc4[2] += ri; acc4[3] += static_cast<double>(ri) * ri;
    }
    log_tree_reduce<THREADS>(acc4, red, s_log, B);
    if (to_pending) {            // p2p mode: this shard's sums wait for the next exchange round
        if (threadIdx.x < 5) st->pending_log[threadIdx.x] = s_log[threadIdx.x];
    } else if (threadIdx.x < 2) {
        double sc[7];                                       // laid out like the staged scalar block
        for (int q = 0; q < 3; ++q) sc[1 + 3 * threadIdx.x + q] = __ldcg(&st->ed_count + 1 + 3 * threadIdx.x + q);
        apply_log_sums(st, sc, s_log, threadIdx.x);
    }
    if (threadIdx.x == 32) st->stash_open = 0;
    __syncthreads();
}

// torch.max(x, 0) semantics (NaN propagates), episodic_novelty.py:60-65
__device__ __forceinline__ float relu_nanprop(float x) { return (x != x) ? x : (x > 0.0f ? x : 0.0f); }

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Dynamic shared memory of the epilogue: [dist B*k][alpha B][tmp B][q B*32] floats.
__host__ __device__ inline size_t epilogue_q_offset(int B, int k) {   // floats; keeps q 16-byte aligned
    return (static_cast<size_t>(B) * (k + 2) + 3) & ~static_cast<size_t>(3);
}
__host__ __device__ inline size_t epilogue_smem_bytes(int B, int k, bool with_q) {
    return (epilogue_q_offset(B, k) + (with_q ? static_cast<size_t>(B) * 32 : 0)) * sizeof(float);
}

template <int THREADS>   // THREADS >= 8 * NGU_EPI_MAX_K (8 lanes per column in the rank-sum pass)
__global__ void __launch_bounds__(THREADS) epilogue_kernel(const EpilogueParams p, const EpiConsts c) {
    __shared__ double s_sums[2 * 32 + 1 + 5 + 2];   // rank sums (+ piggybacked log sums in p2p mode)
    __shared__ double s_mean[32], s_var[32];
    __shared__ double s_scal[kStagedWords];   // DevState from ed_count on: see the SC_* indices
    __shared__ int s_failed;                  // this launch saw (or inherited) a timed-out merge / exchange: rewards become NaN
    __shared__ double s_recv[kMaxWorld * 72];   // p2p receive scratch
    __shared__ double s_ll[2];
    __shared__ float mean32[32];
    __shared__ double red[4][THREADS / 32];
    __shared__ long long s_cursor;
    __shared__ double s_log[8];
    __shared__ int2 s_plan[kPlanMaxEntries + kPlanMaxLeaves];
    __shared__ float s_stack[256][kPlanMaxDepth + 1];   // value stacks of the rank-sum threads (padded)
    extern __shared__ float s_dyn[];
    const int k = c.k;
    const int B = c.n_actors;
    float* s_dist = s_dyn;                 // [B][k]
    float* s_alpha = s_dist + B * k;       // [B]
    float* s_tmp = s_alpha + B;            // [B]
    float4* s_q = reinterpret_cast<float4*>(s_dyn + epilogue_q_offset(B, k));   // [B][8]
    DevState* st = p.state;
    const bool finish = (p.phases & PH_FINISH) != 0, append = (p.phases & PH_APPEND) != 0;

#define EPI_STAMP(i) do { if (p.dbg_stamps && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg_stamps[i] = t_; } } while (0)
    EPI_STAMP(8);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // next step's scan may start its prologue
    if (threadIdx.x == 0) s_failed = 0;   // (read after the load phase's barrier at the earliest)

    // While the scan is still streaming: pull this kernel's small inputs into L2 (the scan's TMA
    // traffic is tagged evict-first, so they stay there).  The previous epilogue has completed by now
    // (the scan signals launch_dependents after its own griddepcontrol.wait).
    if (threadIdx.x < 5) prefetch_l2(reinterpret_cast<const char*>(st) + 128 * threadIdx.x);
    if (threadIdx.x == 5) prefetch_l2(p.plan);
    if (append && !p.stage_q)
        for (int i = threadIdx.x; i < B; i += THREADS) prefetch_l2(p.q + static_cast<int64_t>(i) * 32);

    const bool rnd = finish && (p.phases & PH_RND) != 0;
    if (rnd) {
        // err[b] = mean over the 128 features of (pred - targ)^2 in numpy's f32 pairwise order
        // (n = 128: eight strided accumulators, lifelong_novelty.py:53).  The features were written by
        // kernels that precede the scan in stream order, so this runs while the scan streams.
#pragma unroll 1
        for (int b = threadIdx.x; b < B; b += THREADS) {
            const float4* P = reinterpret_cast<const float4*>(p.rnd_pred + static_cast<int64_t>(b) * 128);
            const float4* T = reinterpret_cast<const float4*>(p.rnd_targ + static_cast<int64_t>(b) * 128);
            float r[8];
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const float4 p0 = __ldg(P + 2 * i), p1 = __ldg(P + 2 * i + 1);
                const float4 t0 = __ldg(T + 2 * i), t1 = __ldg(T + 2 * i + 1);
                const float e[8] = {sqdiff_rn(p0.x, t0.x), sqdiff_rn(p0.y, t0.y), sqdiff_rn(p0.z, t0.z), sqdiff_rn(p0.w, t0.w),
                                    sqdiff_rn(p1.x, t1.x), sqdiff_rn(p1.y, t1.y), sqdiff_rn(p1.z, t1.z), sqdiff_rn(p1.w, t1.w)};
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = (i == 0) ? e[j] : __fadd_rn(r[j], e[j]);
            }
            const float sum = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
            s_alpha[b] = __fmul_rn(sum, 1.0f / 128.0f);   // exact: mean = sum / 128
        }
    }

    // The step's inputs (controllable states for the append, modulators) were complete before the scan
    // kernel could start, so they are staged on chip NOW, while the scan streams.
    if (finish && !rnd) {
#pragma unroll 1
        for (int i = threadIdx.x; i < B; i += THREADS) s_alpha[i] = p.alpha ? __ldcg(p.alpha + i) : 1.0f;
    }
    if (append && p.stage_q) {
#pragma unroll 1
        for (int i = threadIdx.x; i < B * 8; i += THREADS) s_q[i] = __ldcg(reinterpret_cast<const float4*>(p.q) + i);
    }

    // The previous fused step's reward stash -> log-only trackers, while the scan drains (see log_absorb).
    const bool logfuse = finish && (p.phases & PH_LOGFUSE) != 0;
    if (logfuse) {
        if (p.wait_prev != 0u && threadIdx.x == 0) {
            // Pipelined step: this CTA became resident without the scan in front of it having waited for the previous
            // epilogue, which writes the stash and the trackers.  Wait for it here (it is resident or complete; this
            // CTA has nothing else to do until its own scan drains).
            unsigned int v;
            for (int polls = 0;; ++polls) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.epi_done) : "memory");
                if (static_cast<int>(v - p.wait_prev) >= 0) break;
                if (polls > (1 << 22)) { atomicAdd(reinterpret_cast<unsigned long long*>(&st->exchange_timeouts), 1ull); break; }
            }
        }
        if (threadIdx.x == 0) s_cursor = __ldcg(&st->stash_open);      // s_cursor: scratch until the load phase
        __syncthreads();
        const bool open = s_cursor != 0;
        __syncthreads();
        if (open) log_absorb<THREADS>(st, p.log_stash, B, (p.phases & PH_EXCHANGE) != 0, red, s_log);
    }

    EPI_STAMP(9);
    asm volatile("griddepcontrol.wait;" ::: "memory");               // scan kernel complete, its writes visible
    EPI_STAMP(0);
    // A pre-armed step whose doorbell never rang: its scan did nothing, and neither does this launch.  (The
    // stash absorbed above is the same operation ngu_epi_get_state performs: harmless whenever it happens.)
    if (p.arm_seq != 0u && __ldcg(p.arm_go) == ((p.arm_seq << 1) | 1u)) return;

    // ---- one batch of independent loads: everything the phases below need goes on chip ----
    // dist goes global -> shared as 16-byte cp.async (L2 only): every chunk is in flight at once instead of
    // one L2 round trip per loop iteration
    {
        const int n16 = (B * k) >> 2;
        const unsigned s_base = static_cast<unsigned>(__cvta_generic_to_shared(s_dist));
#pragma unroll 1
        for (int i = threadIdx.x; i < n16; i += THREADS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_base + 16u * i), "l"(p.dist + 4 * i) : "memory");
        const int t = (n16 << 2) + threadIdx.x;
        if (t < B * k) s_dist[t] = __ldcg(p.dist + t);
    }
    if (finish) {
        if (p.sums_in && threadIdx.x < 2 * k + 1) s_sums[threadIdx.x] = __ldcg(p.sums_in + threadIdx.x);
        if (threadIdx.x < k) {
            s_mean[threadIdx.x] = __ldcg(&st->ed_mean[threadIdx.x]);
            s_var[threadIdx.x] = __ldcg(&st->ed_var[threadIdx.x]);
        }
    }
    // staged even without FINISH: cursor (APPEND) and the sticky timeout counter live in the block
    if (threadIdx.x >= 32 && threadIdx.x < 32 + kStagedWords) s_scal[threadIdx.x - 32] = __ldcg(&st->ed_count + (threadIdx.x - 32));
#pragma unroll 1
    for (int i = threadIdx.x; i < p.plan_len + p.n_leaves; i += THREADS) s_plan[i] = __ldcg(p.plan + i);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // Sticky error (include/ngu_epi.h): a merge of this or an earlier scan, or an earlier exchange, gave up waiting.
    // Whatever would be computed from here on is wrong, so say so loudly: NaN rewards and the mapped error word.
    const bool failed_before = __double_as_longlong(s_scal[SC_TIMEOUTS]) != 0;
    EPI_STAMP(1);

    if (rnd) {
        // ll_rms.update(err) with use_mpi=False (mpi_util.py:51-52: np.mean / np.std over the local batch,
        // float32 pairwise sums), then alpha = 1 + (err - mean)/sqrt(var) in float64 (lifelong_novelty.py:54-57)
        if (threadIdx.x < 32) {
            const float sm = np_pairwise_sum8(s_alpha, 1, s_plan, p.plan_len, threadIdx.x & 7, s_stack[threadIdx.x]);
            if (threadIdx.x == 0) s_ll[0] = static_cast<double>(f32_div(sm, static_cast<float>(B)));
        }
        __syncthreads();
        const float bmean = static_cast<float>(s_ll[0]);
#pragma unroll 1
        for (int b = threadIdx.x; b < B; b += THREADS) {
            const float d = __fsub_rn(s_alpha[b], bmean);
            s_tmp[b] = __fmul_rn(d, d);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const float sv = np_pairwise_sum8(s_tmp, 1, s_plan, p.plan_len, threadIdx.x & 7, s_stack[threadIdx.x]);
            if (threadIdx.x == 0) {
                const float nB = static_cast<float>(B);
                const float b_std = f32_sqrt(f32_div(sv, nB));
                const float b_var = __fmul_rn(b_std, b_std);                       // np.square(batch_std)
                const double m_b = static_cast<double>(__fmul_rn(b_var, nB));      // batch_var * batch_count stays f32
                const double n_b = static_cast<double>(B);
                const double mean = s_scal[SC_LL], var = s_scal[SC_LL + 1], count = s_scal[SC_LL + 2];
                const double delta = static_cast<double>(bmean) - mean;
                const double totc = count + n_b;
                // every product and sum rounded separately, like numpy (no FMA contraction): mpi_util.py:59-73
                const double new_mean = __dadd_rn(mean, f64_div(__dmul_rn(delta, n_b), totc));
                const double m2 = __dadd_rn(__dadd_rn(__dmul_rn(var, count), m_b),
                                            f64_div(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), count), n_b), totc));
                const double new_var = f64_div(m2, totc);
                st->ll_mean = new_mean; st->ll_var = new_var; st->ll_count = totc;
                s_ll[0] = new_mean;
                s_ll[1] = sqrt(new_var);
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int b = threadIdx.x; b < B; b += THREADS) {
            const float a = static_cast<float>(1.0 + f64_div(static_cast<double>(s_alpha[b]) - s_ll[0], s_ll[1]));
            s_alpha[b] = a;
            if (p.alpha_out) p.alpha_out[b] = a;
        }
        __syncthreads();
    }

    if (p.phases & PH_RANKSUMS) {
        float* s_leafsum = &s_stack[64][0];                        // [k][n_leaves]; stacks 0..31 serve the fold below
        if (threadIdx.x < 256) {
            // column sums in numpy's pairwise order.  Every (column, leaf) pair is an independent sum of
            // <= 128 rows: 8 lanes own numpy's 8 strided accumulators (whole warps: shuffles inside), the
            // 32 groups of 8 lanes take the pairs round robin - one round at 256 actors (10 x 2 pairs).
            const int group = threadIdx.x >> 3, lane8 = threadIdx.x & 7;
            const int L = p.n_leaves, items = k * L;
#pragma unroll 1
            for (int w0 = 0; w0 < items; w0 += 32) {
                const int w = w0 + group;
                const int wc = w < items ? w : items - 1;              // idle groups shadow the last pair
                const int j = wc / L, l = wc - j * L;
                const int2 e = s_plan[p.plan_len + l];
                const float s1 = np_pairwise_leaf8(s_dist + j + e.x * k, k, e.y, lane8);
                if (w < items && lane8 == 0) s_leafsum[wc] = s1;
            }
        } else {
            // meanwhile: f64 sums of squares (for the log-only variance), a full warp per column
            const int lane = threadIdx.x & 31;
#pragma unroll 1
            for (int j = (threadIdx.x >> 5) - 8; j < k; j += (THREADS - 256) / 32) {
                double s2 = 0.0;
#pragma unroll 4
                for (int bb = lane; bb < B; bb += 32) {
                    const double v = static_cast<double>(s_dist[bb * k + j]);
                    s2 += v * v;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(kFull, s2, o);
                if (lane == 0) s_sums[k + j] = s2;
            }
        }
        __syncthreads();
        if (threadIdx.x < k) {
            // fold the leaf sums of column j along numpy's recursion (the postfix plan)
            float* stack = s_stack[threadIdx.x];
            const float* leaf = s_leafsum + threadIdx.x * p.n_leaves;
            int sp = 0;
#pragma unroll 1
            for (int i = 0; i < p.plan_len; ++i) {
                if (s_plan[i].y > 0) {
                    stack[sp++] = *leaf++;
                } else {
                    --sp;
                    stack[sp - 1] = __fadd_rn(stack[sp - 1], stack[sp]);
                }
            }
            s_sums[threadIdx.x] = static_cast<double>(stack[0]);
        }
        if (threadIdx.x == 0) s_sums[2 * k] = static_cast<double>(B);
        __syncthreads();
        if (threadIdx.x < 2 * k + 1) p.rank_sums[threadIdx.x] = s_sums[threadIdx.x];
    }

    EPI_STAMP(2);
    const bool exchange = (p.phases & PH_EXCHANGE) != 0;
    if (exchange) {
        // ONE exchange per step: this step's rank sums (mpi_util.py:17) and, piggybacked, the log-only
        // sums of the PREVIOUS step (episodic_novelty.py:82, intrinsic_novelty.py:32) - the trackers they
        // feed are never read by the hot path, so they are brought up to date one step late (or on
        // demand: ngu_epi_log_flush) instead of paying a second rendezvous per step.
        const unsigned long long xseq = static_cast<unsigned long long>(__double_as_longlong(s_scal[SC_XCHG_SEQ]));
        if (threadIdx.x < 5) s_sums[2 * k + 1 + threadIdx.x] = s_scal[SC_PENDING + threadIdx.x];
        __syncthreads();
        if (!p2p_allreduce<THREADS>(p.p2p, s_sums, 2 * k + 6, 0, xseq, s_recv) && threadIdx.x == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(&st->exchange_timeouts), 1ull);
            s_failed = 1;
        }
        if (threadIdx.x == 32) st->xchg_seq = static_cast<long long>(xseq + 1ull);
        if (threadIdx.x < 2) apply_log_sums(st, s_scal, s_sums + 2 * k + 1, threadIdx.x);
        __syncthreads();
    }

    EPI_STAMP(3);
    if (finish) {
        // RunningMeanStd.update_from_moments (mpi_util.py:59-73).  Only the MEAN feeds the rewards
        // (episodic_novelty.py:58), so warp 0 runs the short chain to it while warp 1 runs the longer
        // variance chain (log-only) side by side.
        if ((threadIdx.x & 31) < k && threadIdx.x < 64) {
            const int j = threadIdx.x & 31;
            const double cnt = s_sums[2 * k];
            const float cnt32 = static_cast<float>(cnt);
            // mpi_mean: globalsum[:n] / globalsum[n] in float32 (mpi_util.py:18)
            const float b_mean = f32_div(static_cast<float>(s_sums[j]), cnt32);
            const double bm = static_cast<double>(b_mean);
            const double count = s_scal[SC_ED_COUNT];
            const double n_b = static_cast<double>(cnt32);
            const double delta = bm - s_mean[j];
            const double totc = count + n_b;
            if (threadIdx.x < 32) {
                const double new_mean = __dadd_rn(s_mean[j], f64_div(__dmul_rn(delta, n_b), totc));   // no FMA contraction (numpy)
                st->ed_mean[j] = new_mean;
                mean32[j] = static_cast<float>(new_mean);   // ptu.to_tensor(...).float(), pytorch_util.py:20
            } else {
                // mpi_moments' second pass (mpi_util.py:25-28) from the f64 sum of squares:
                // E[(x-m)^2] = E[x^2] - 2 m E[x] + m^2 with m the f32 batch mean.
                const double ex = f64_div(s_sums[j], cnt), ex2 = f64_div(s_sums[k + j], cnt);
                const float msd = static_cast<float>(fmax(ex2 - 2.0 * bm * ex + bm * bm, 0.0));
                const float b_std = f32_sqrt(msd);
                const float b_var = __fmul_rn(b_std, b_std);                      // np.square(batch_std), :56
                const double m_b = static_cast<double>(__fmul_rn(b_var, cnt32)); // batch_var * batch_count (f32), :65
                const double m2 = __dadd_rn(__dadd_rn(__dmul_rn(s_var[j], count), m_b),
                                            f64_div(__dmul_rn(__dmul_rn(__dmul_rn(delta, delta), count), n_b), totc));
                st->ed_var[j] = f64_div(m2, totc);
            }
        }
        if (threadIdx.x == 64) {
            st->ed_count = s_scal[SC_ED_COUNT] + static_cast<double>(static_cast<float>(s_sums[2 * k]));
            st->steps = st_steps_plus_one(s_scal[SC_STEPS]);
        }
        __syncthreads();
        EPI_STAMP(4);

        // (i) one thread per (actor, neighbour): the IEEE division and reciprocal each hide a branch to a
        // slow path, so ptxas will not interleave them - k of them in one thread are k serial chains
        const int j_step = THREADS % k;
        int j = threadIdx.x % k;
#pragma unroll 1
        for (int e = threadIdx.x; e < B * k; e += THREADS, j = j + j_step >= k ? j + j_step - k : j + j_step) {
            const float nd = __fdiv_rn(s_dist[e], mean32[j]);                   // :58
            const float cl = relu_nanprop(__fsub_rn(nd, c.cluster));            // :60-65
            s_dist[e] = __fmul_rn(__frcp_rn(__fadd_rn(cl, c.eps)), c.eps);      // :67-68 (reciprocal()*eps), in place
        }
        __syncthreads();
        // (ii) one thread per actor
        const bool failed = failed_before || s_failed != 0;
        double acc4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int b = threadIdx.x; b < B; b += THREADS) {
            float acc = s_dist[b * k];
#pragma unroll 8
            for (int jj = 1; jj < k; ++jj) acc = __fadd_rn(acc, s_dist[b * k + jj]);   // :71 sum(dim=0), sequential; loads hoisted
            const float sim = __fadd_rn(f32_sqrt(acc), c.pseudo);                   // :71-72
            float r = f32_rcp(sim);                                                 // :75
            if (sim > c.s_max) r = 0.0f;                                            // :76-78
            float a = s_alpha[b];
            if (a < 1.0f) a = 1.0f;                                                 // intrinsic_novelty.py:26-27
            if (a > c.max_rew) a = c.max_rew;                                       // :28-29
            float ri = __fmul_rn(r, a);                                             // :31
            if (failed) { r = __int_as_float(0x7fc00000); ri = r; }                 // sticky error: never a plausible number
            if (p.r_epi) p.r_epi[b] = r;
            p.r_int[b] = ri;
            if (logfuse) {       // log-only statistics: stashed, absorbed by the next launch (log_absorb)
                p.log_stash[b] = r;
                p.log_stash[B + b] = ri;
            } else if (p.log_sums) {
                acc4[0] += r; acc4[1] += static_cast<double>(r) * r;
                acc4[2] += ri; acc4[3] += static_cast<double>(ri) * ri;
            }
        }
        if (logfuse && threadIdx.x == 96) st->stash_open = 1;
        EPI_STAMP(5);
        if (p.log_sums && !logfuse) {   // split sequence: the caller reduces the log-only sums over the shards itself
            log_tree_reduce<THREADS>(acc4, red, s_log, B);
            if (threadIdx.x < 5) p.log_sums[threadIdx.x] = s_log[threadIdx.x];
        }
    }

    EPI_STAMP(6);
    if (append) {
        const long long cur = __double_as_longlong(s_scal[SC_CURSOR]);
        const float4* q4 = reinterpret_cast<const float4*>(p.q);
#pragma unroll 1
        for (int i = threadIdx.x; i < B * 8; i += THREADS) {
            const int b = i >> 3, ch = i & 7;
            float4* dst = reinterpret_cast<float4*>(p.memory + (static_cast<int64_t>(b) * c.capacity + cur) * 32) + ch;
            *dst = p.stage_q ? s_q[i] : __ldcg(q4 + i);
        }
        if (threadIdx.x == 64) st->cursor = (cur + 1) % c.capacity;
    }

    EPI_STAMP(7);
    if (threadIdx.x == 0 && p.host_err && (failed_before || s_failed != 0)) *reinterpret_cast<volatile unsigned int*>(p.host_err) = 1u;
    if (p.host_flag || p.seq != 0u) {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (p.seq != 0u) {   // every write of this launch precedes the word a pipelined successor polls
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.epi_done), "r"(p.seq) : "memory");
            }
            if (p.host_flag) {   // ngu_epi_step_host: tell the spinning host that every write of the step has been issued
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int*>(p.host_flag) = p.host_seq;
            }
        }
    }
}

// On-demand flush of the deferred log-only sums (p2p mode): one exchange round of 5 doubles,
// then the tracker update; all ranks must launch it together.
constexpr int kLogThreads = 512;   // = the epilogue's CTA size: log_tree_reduce must add in the same order

// Stand-alone absorb (ngu_epi_get_state, ngu_epi_log_flush outside p2p mode).
__global__ void __launch_bounds__(kLogThreads) log_absorb_kernel(DevState* st, const float* stash, int B, int to_pending) {
    __shared__ double red[4][kLogThreads / 32];
    __shared__ double s_log[8];
    if (__ldcg(&st->stash_open) == 0) return;              // CTA-uniform
    log_absorb<kLogThreads>(st, stash, B, to_pending != 0, red, s_log);
}

__global__ void __launch_bounds__(kLogThreads) log_flush_kernel(DevState* st, P2PView v, const float* stash, int B) {
    __shared__ double s_scal[kStagedWords];
    __shared__ double s_vals[8];
    __shared__ double s_recv[kMaxWorld * 8];
    __shared__ double red[4][kLogThreads / 32];
    __shared__ double s_log[8];
    if (__ldcg(&st->stash_open) != 0)                      // CTA-uniform
        log_absorb<kLogThreads>(st, stash, B, true, red, s_log);
    __syncthreads();
    if (threadIdx.x < kStagedWords) s_scal[threadIdx.x] = __ldcg(&st->ed_count + threadIdx.x);
    __syncthreads();
    if (threadIdx.x < 5) s_vals[threadIdx.x] = s_scal[SC_PENDING + threadIdx.x];
    __syncthreads();
    const unsigned long long seq = static_cast<unsigned long long>(st->flush_seq);
    if (!p2p_allreduce<kLogThreads>(v, s_vals, 5, kFlushElem, seq, s_recv) && threadIdx.x == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(&st->exchange_timeouts), 1ull);
    if (threadIdx.x < 2) apply_log_sums(st, s_scal, s_vals, threadIdx.x);
    if (threadIdx.x < 5) st->pending_log[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) st->flush_seq = static_cast<long long>(seq + 1);
}

__global__ void log_update_kernel(DevState* st, const double* log_sums) {
    if (threadIdx.x == 0 && blockIdx.x == 0) log_rms_apply(&st->epi_mean, log_sums);
}

// ---------------------------------------------------------------------------
// Append (insert_state, episodic_novelty.py:86-89).  One CTA: every thread reads
// the device cursor, writes its 16-byte piece of mem[b][cursor][:], then thread 0
// advances the cursor.  B*128 bytes in total.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) append_kernel(float* __restrict__ mem, const float* __restrict__ q,
                                                       DevState* st, int n_actors, int capacity) {
    const long long cur = st->cursor;
    const float4* q4 = reinterpret_cast<const float4*>(q);
    for (int i = threadIdx.x; i < n_actors * 8; i += blockDim.x) {
        const int b = i >> 3, ch = i & 7;
        float4* dst = reinterpret_cast<float4*>(mem + (static_cast<int64_t>(b) * capacity + cur) * 32) + ch;
        *dst = q4[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) st->cursor = (cur + 1) % capacity;
}

// ---------------------------------------------------------------------------
// Reset (reset_memory_if_done, intrinsic_novelty.py:36-40).  Most steps nobody is done, or one
// actor of hundreds is, so the grid must not grow with the actor count: grid (chunks, G), G small.
// Every CTA compacts the done mask (ballot + popc, 256 actors per pass); CTA (c, g) zero-fills
// chunk c (32 KB, 16-byte stores) of the done actors with ordinal g, g + G, ...
// 256 x 30k, per call through the Python engine (tools/reset_probe.py): one CTA row per actor cost 31 us
// of GPU time whoever was done; with G = 32 the call is host-bound (11.5 us) for 0-4 done actors,
// 18 us for 26 (5.5 TB/s) and 147 us when everybody is done (6.7 TB/s; 7.0 at G = 64, 5.7 at G = 8).
// ---------------------------------------------------------------------------
constexpr int kResetGroups = 32;
__global__ void __launch_bounds__(256) reset_kernel(float* __restrict__ mem, const uint8_t* __restrict__ done,
                                                     int n_actors, int capacity) {
    __shared__ int s_list[256];
    __shared__ int s_wcnt[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y, G = gridDim.y;
    const int64_t n4 = static_cast<int64_t>(capacity) * 8;  // float4 per actor
    const int64_t start = static_cast<int64_t>(blockIdx.x) * 2048;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    int seen = 0;                                            // done actors in the passes so far
#pragma unroll 1
    for (int b0 = 0; b0 < n_actors; b0 += 256) {
        const int b = b0 + threadIdx.x;
        const bool d = b < n_actors && done[b] != 0;
        const unsigned bal = __ballot_sync(kFull, d);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int c = s_wcnt[w];
            if (w < warp) off += c;
            tot += c;
        }
        if (d) s_list[off + __popc(bal & ((1u << lane) - 1u))] = b;
        __syncthreads();
        int first = (g - seen) % G;
        if (first < 0) first += G;
#pragma unroll 1
        for (int i = first; i < tot; i += G) {
            float4* base = reinterpret_cast<float4*>(mem) + static_cast<int64_t>(s_list[i]) * n4;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int64_t e = start + it * 256 + threadIdx.x;
                if (e < n4) base[e] = z;
            }
        }
        seen += tot;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// ngu_epi_step_host: the step's inputs (controllable states + modulators, B*132 bytes in pinned, mapped
// host memory) are pulled over PCIe by this kernel instead of by a copy-engine transfer.  A kernel can be
// chained with programmatic dependent launch, a memcpy cannot: the scan's CTAs become resident and issue
// their L2 prefetches while these loads are in flight, instead of being launched after the copy has
// completed.  One 16-byte load per thread, all in flight at once.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stage_inputs_kernel(float4* __restrict__ dst, const float4* src_host, int n16) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n16) {
        float4 v;
        asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src_host + i) : "memory");
        dst[i] = v;
    }
}

// ---------------------------------------------------------------------------
// Pre-armed host step (ngu_epi_step_host called back to back).  While step t runs, the host already enqueues the
// three launches of step t+1: this kernel, the scan and the epilogue, chained with programmatic dependent launch.
// This one-CTA kernel waits for the host's DOORBELL - a word in mapped host memory the host stores to after it
// has put the step's inputs into the pinned buffer - and then pulls those inputs over PCIe, exactly like
// stage_inputs_kernel.  What leaves the critical path of a call: three kernel launches (~9 us of enqueue) and the
// GPU-side launch latency; the scan's CTAs are already resident behind this kernel, their first tiles prefetched
// into L2, when the doorbell rings.
// The wait is BOUNDED (`window_ns`): a caller that does not come back (or synchronises the device) must not find
// the GPU occupied by a kernel that waits for it.  On timeout, or when the host rings with the cancel bit, the
// kernel records the cancellation in `go_dev` (the scan and epilogue behind it then exit at once without touching
// any state) and tells the host through `status_host`; the host then launches the step the ordinary way.
//   doorbell value : seq << 2 | cancel << 1 | has_alpha
//   status value   : seq << 1 | cancelled          (acknowledgement: written as soon as the doorbell has been seen)
//   go value       : seq << 1 | 1                  (device memory, written on cancellation only: read by the scan and the epilogue)
// ---------------------------------------------------------------------------
constexpr int kArmThreads = 1024;
__global__ void __launch_bounds__(kArmThreads) armed_stage_kernel(float4* __restrict__ dst, const float4* src_host, int n_actors,
                                                                  const unsigned int* doorbell_host, unsigned int seq,
                                                                  unsigned int* go_dev, unsigned int* status_host,
                                                                  long long window_ns) {
    __shared__ unsigned int s_db;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the scan becomes resident behind us and prefetches
    asm volatile("griddepcontrol.wait;" ::: "memory");                // previous epilogue complete (and, transitively, its scan):
                                                                      // whoever waits for THIS kernel inherits that order
    if (threadIdx.x == 0) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned int v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(doorbell_host) : "memory");   // the inputs were written before it
            if ((v >> 2) == seq) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (static_cast<long long>(t - t0) > window_ns) { v = (seq << 2) | 2u; break; }
        }
        *reinterpret_cast<volatile unsigned int*>(status_host) = (seq << 1) | ((v >> 1) & 1u);
        if (v & 2u) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go_dev), "r"((seq << 1) | 1u) : "memory");
        s_db = v;
    }
    __syncthreads();
    const unsigned int v = s_db;
    if (v & 2u) return;
    const int n16_q = n_actors * 8;                                   // float4 pieces of the controllable states
    const int n16_all = (n_actors * 33 + 3) / 4;                      // ... plus the modulators (padded buffers)
    const int n16 = (v & 1u) ? n16_all : n16_q;
    // every PCIe read of a batch is in flight before the first result is used: one round trip for up to 4 x 1024 pieces
    // (256 actors are 2 112 pieces)
#pragma unroll 1
    for (int i0 = threadIdx.x; i0 < n16; i0 += 4 * kArmThreads) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kArmThreads;
            if (i < n16)
                asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w) : "l"(src_host + i) : "memory");
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kArmThreads;
            if (i < n16) dst[i] = x[u];
        }
    }
    if (!(v & 1u)) {                                                  // no modulators: alpha = 1 (clip(1, 1, L) = 1, r * 1 = r)
        float* a = reinterpret_cast<float*>(dst) + static_cast<size_t>(n_actors) * 32;
        for (int i = threadIdx.x; i < n_actors; i += kArmThreads) a[i] = 1.0f;
    }
}

// Layout converters for load/dump in the reference's slot-major [N][B][32] order.
__global__ void transpose_rows_kernel(float4* __restrict__ dst, const float4* __restrict__ src, int outer,
                                      int inner) {
    // src [outer][inner][8 float4] -> dst [inner][outer][8 float4]
    const int64_t total = static_cast<int64_t>(outer) * inner * 8;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i & 7);
        const int64_t row = i >> 3;
        const int64_t o = row / inner, in = row - o * inner;
        dst[(in * outer + o) * 8 + ch] = src[i];
    }
}

}  // namespace ngu
