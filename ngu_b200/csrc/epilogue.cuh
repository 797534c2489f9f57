// epilogue.cuh — kernel B (merge partial lists -> per-actor k-NN, rank sums) and
// kernel C (running-mean update + reward epilogue), plus append / reset.
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

#include "warp_topk.cuh"

namespace ngu {

// Device-resident running statistics (ngu_epi_state without the host mirror).
struct DevState {
    double ed_mean[32];
    double ed_var[32];
    double ed_count;
    double epi_mean, epi_var, epi_count;
    double intr_mean, intr_var, intr_count;
    long long cursor;
    long long steps;
    unsigned int merge_ticket;   // last-block election for kernel B
    unsigned int pad;
};

struct EpiConsts {
    float eps, cluster, pseudo, s_max, max_rew;
    int32_t k, n_actors, capacity, largest, sqrt_dist;
};

// ---------------------------------------------------------------------------
// numpy's float32 pairwise summation (x.sum(axis=0), mpi_util.py:11) over one
// column v[b*stride], b = 0..n-1, evaluated cooperatively by an aligned group of
// 8 lanes (lane8 = lane & 7).  numpy: leaves of <= 128 elements use eight strided
// accumulators r[0..7] combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and a
// sequential tail; larger n splits at n/2 rounded down to a multiple of 8.
// Here lane a owns accumulator r[a]; the combine is three xor-shuffles (float
// add is commutative, so every lane of the group ends with the same bits).
// All 32 lanes of the warp must call this with the same n.
// ---------------------------------------------------------------------------
__device__ inline float np_pairwise_leaf8(const float* v, int stride, int n, int lane8) {
    if (n < 8) {
        float r = -0.0f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, __ldcg(v + static_cast<int64_t>(i) * stride));
        return r;
    }
    const int lim = n - (n % 8);
    float r = __ldcg(v + static_cast<int64_t>(lane8) * stride);
#pragma unroll 4
    for (int i = 8; i < lim; i += 8) r = __fadd_rn(r, __ldcg(v + static_cast<int64_t>(i + lane8) * stride));
    r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(kFull, r, 4));
    for (int i = lim; i < n; ++i) r = __fadd_rn(r, __ldcg(v + static_cast<int64_t>(i) * stride));
    return r;
}
__device__ inline float np_pairwise_sum8(const float* v, int stride, int n, int lane8) {
    if (n <= 128) return np_pairwise_leaf8(v, stride, n, lane8);
    int n2 = n / 2;
    n2 -= n2 % 8;
    const float lo = np_pairwise_sum8(v, stride, n2, lane8);
    const float hi = np_pairwise_sum8(v + static_cast<int64_t>(n2) * stride, stride, n - n2, lane8);
    return __fadd_rn(lo, hi);
}

// ---------------------------------------------------------------------------
// Kernel B: one warp per actor merges that actor's partial lists; the last CTA
// to finish reduces the (B,k) distance matrix to this shard's rank sums.
// ---------------------------------------------------------------------------
struct MergeParams {
    const u64* partial;      // [B][max_parts][k]
    float* topk_d2;          // [B][k]
    int32_t* topk_idx;       // [B][k]
    float* dist;             // [B][k]  sqrt(d2) (reference) or d2 (paper)
    double* rank_sums;       // [2k+1]
    DevState* state;
    int32_t tiles_per_actor, tiles_per_warp, max_parts, pad;
};

template <int WARPS>   // WARPS*32 >= 8*k is required (8 lanes per column in the rank-sum pass)
__global__ void __launch_bounds__(WARPS * 32) merge_topk_kernel(const MergeParams p, const EpiConsts c) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * WARPS + warp;
    const int k = c.k;
    const bool largest = c.largest != 0;

    if (b < c.n_actors) {
        const int64_t t0 = static_cast<int64_t>(b) * p.tiles_per_actor;
        const int first = static_cast<int>(t0 / p.tiles_per_warp);
        const int last = static_cast<int>((t0 + p.tiles_per_actor - 1) / p.tiles_per_warp);
        const int nkeys = (last - first + 1) * k;
        const u64* base = p.partial + static_cast<int64_t>(b) * p.max_parts * k;
        WarpTopK top;
        top.reset();
        for (int i0 = 0; i0 < nkeys; i0 += 32) {
            const int i = i0 + lane;
            const u64 key = i < nkeys ? base[i] : 0ull;
            top.offer(key, lane, k);
        }
        if (lane < k) {
            const float d2 = key_d2(top.lkey, largest);
            p.topk_d2[b * k + lane] = d2;
            p.topk_idx[b * k + lane] = key_slot(top.lkey);
            p.dist[b * k + lane] = c.sqrt_dist ? __fsqrt_rn(d2) : d2;
        }
    }

    // ---- last-CTA election (threadfence reduction) ----
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(&p.state->merge_ticket, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) p.state->merge_ticket = 0;  // re-arm for the next step

    // Rank sums (mpi_mean's localsum, mpi_util.py:11-15): column sums of the
    // (B,k) matrix in numpy's pairwise f32 order, then sum of squares in f64
    // for the (log-only) variance, then the row count.
    const int B = c.n_actors;
    const int group = threadIdx.x >> 3, lane8 = threadIdx.x & 7;   // 8 lanes per column
    const int j = group < k ? group : k - 1;                       // idle groups shadow column k-1
    const float s1 = np_pairwise_sum8(p.dist + j, k, B, lane8);
    double s2 = 0.0;
    for (int bb = lane8; bb < B; bb += 8) {
        const double v = static_cast<double>(__ldcg(p.dist + static_cast<int64_t>(bb) * k + j));
        s2 += v * v;
    }
    s2 += __shfl_xor_sync(kFull, s2, 1);
    s2 += __shfl_xor_sync(kFull, s2, 2);
    s2 += __shfl_xor_sync(kFull, s2, 4);
    if (group < k && lane8 == 0) {
        p.rank_sums[group] = static_cast<double>(s1);
        p.rank_sums[k + group] = s2;
    }
    if (threadIdx.x == 0) p.rank_sums[2 * k] = static_cast<double>(B);
}

// ---------------------------------------------------------------------------
// Kernel C: RunningMeanStd update (mpi_util.py:48-73) from the global rank sums,
// then the per-actor reward (episodic_novelty.py:58-78, intrinsic_novelty.py:26-31).
// Single CTA; B is at most a few thousand actors per shard.
// ---------------------------------------------------------------------------
struct FinishParams {
    const float* dist;       // [B][k]
    const double* sums;      // [2k+1] global
    const float* alpha;      // [B] or null
    float* r_int;            // [B]
    float* r_epi;            // [B] or null
    double* log_sums;        // [5] or null
    DevState* state;
    int32_t fuse_log_update; // 1: apply the log-only rms updates here (single shard)
    int32_t pad;
};

// torch.max(x, 0) semantics (NaN propagates), episodic_novelty.py:60-65
__device__ __forceinline__ float relu_nanprop(float x) { return (x != x) ? x : (x > 0.0f ? x : 0.0f); }

// Chan/Welford merge of (mean, var, count) with a batch (mpi_util.py:59-73), all f64.
__device__ inline void rms_merge(double& mean, double& var, double& count, double b_mean, double b_var,
                                 double b_count) {
    const double delta = b_mean - mean;
    const double tot = count + b_count;
    const double new_mean = mean + delta * b_count / tot;
    const double m2 = var * count + b_var * b_count + delta * delta * count * b_count / tot;
    mean = new_mean;
    var = m2 / tot;
    count = tot;
}

__device__ inline void log_rms_apply(DevState* st, const double* ls) {
    const double n = ls[4];
    if (n > 0.0) {
        const double m1 = ls[0] / n, v1 = fmax(ls[1] / n - m1 * m1, 0.0);
        rms_merge(st->epi_mean, st->epi_var, st->epi_count, m1, v1, n);
        const double m2 = ls[2] / n, v2 = fmax(ls[3] / n - m2 * m2, 0.0);
        rms_merge(st->intr_mean, st->intr_var, st->intr_count, m2, v2, n);
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) finish_kernel(const FinishParams p, const EpiConsts c) {
    __shared__ float mean32[32];
    __shared__ double red[4][THREADS / 32];
    __shared__ double tot[4];
    const int k = c.k;
    DevState* st = p.state;

    if (threadIdx.x < k) {
        const int j = threadIdx.x;
        const double cnt = p.sums[2 * k];
        const float cnt32 = static_cast<float>(cnt);
        // mpi_mean: globalsum[:n] / globalsum[n] in float32 (mpi_util.py:18)
        const float b_mean = __fdiv_rn(static_cast<float>(p.sums[j]), cnt32);
        // mpi_moments' second pass (mpi_util.py:25-28) from the f64 sum of squares:
        // E[(x-m)^2] = E[x^2] - 2 m E[x] + m^2 with m the f32 batch mean.
        const double ex = p.sums[j] / cnt, ex2 = p.sums[k + j] / cnt, bm = static_cast<double>(b_mean);
        const float msd = static_cast<float>(fmax(ex2 - 2.0 * bm * ex + bm * bm, 0.0));
        const float b_std = __fsqrt_rn(msd);
        const float b_var = __fmul_rn(b_std, b_std);                 // np.square(batch_std), :56
        const double m_b = static_cast<double>(__fmul_rn(b_var, cnt32));  // batch_var * batch_count (f32), :65
        const double count = st->ed_count;
        const double delta = bm - st->ed_mean[j];
        const double totc = count + static_cast<double>(cnt32);
        const double new_mean = st->ed_mean[j] + delta * static_cast<double>(cnt32) / totc;
        const double m2 = st->ed_var[j] * count + m_b +
                          delta * delta * count * static_cast<double>(cnt32) / totc;
        st->ed_mean[j] = new_mean;
        st->ed_var[j] = m2 / totc;
        mean32[j] = static_cast<float>(new_mean);                    // ptu.to_tensor(...).float(), pytorch_util.py:20
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->ed_count = st->ed_count + static_cast<double>(static_cast<float>(p.sums[2 * k]));
        st->steps += 1;
    }

    double s_e = 0.0, s_e2 = 0.0, s_i = 0.0, s_i2 = 0.0;
    for (int b = threadIdx.x; b < c.n_actors; b += THREADS) {
        float acc = 0.0f;
        for (int j = 0; j < k; ++j) {
            const float v = p.dist[b * k + j];
            const float nd = __fdiv_rn(v, mean32[j]);                          // :58
            const float cl = relu_nanprop(__fsub_rn(nd, c.cluster));           // :60-65
            const float kv = __fmul_rn(__frcp_rn(__fadd_rn(cl, c.eps)), c.eps); // :67-68 (reciprocal()*eps)
            acc = (j == 0) ? kv : __fadd_rn(acc, kv);                          // :71 sum(dim=0)
        }
        const float sim = __fadd_rn(__fsqrt_rn(acc), c.pseudo);                // :71-72
        float r = __frcp_rn(sim);                                              // :75
        if (sim > c.s_max) r = 0.0f;                                           // :76-78
        float a = p.alpha ? p.alpha[b] : 1.0f;
        if (a < 1.0f) a = 1.0f;                                                // intrinsic_novelty.py:26-27
        if (a > c.max_rew) a = c.max_rew;                                      // :28-29
        const float ri = __fmul_rn(r, a);                                      // :31
        if (p.r_epi) p.r_epi[b] = r;
        p.r_int[b] = ri;
        s_e += r; s_e2 += static_cast<double>(r) * r;
        s_i += ri; s_i2 += static_cast<double>(ri) * ri;
    }

    // log-only batch sums (epi_novel_rms / intrinsic_novel_rms), fixed-order tree
    if (p.log_sums || p.fuse_log_update) {
        double vals[4] = {s_e, s_e2, s_i, s_i2};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double v = vals[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double v = 0.0;
            for (int w = 0; w < THREADS / 32; ++w) v += red[threadIdx.x][w];
            tot[threadIdx.x] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ls[5] = {tot[0], tot[1], tot[2], tot[3], static_cast<double>(c.n_actors)};
            if (p.log_sums)
                for (int q = 0; q < 5; ++q) p.log_sums[q] = ls[q];
            if (p.fuse_log_update) log_rms_apply(st, ls);
        }
    }
}

__global__ void log_update_kernel(DevState* st, const double* log_sums) {
    if (threadIdx.x == 0 && blockIdx.x == 0) log_rms_apply(st, log_sums);
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
// Reset (reset_memory_if_done, intrinsic_novelty.py:36-40): grid (chunks, B);
// CTAs of actors that are not done exit at once; the others zero 32 KB each.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reset_kernel(float* __restrict__ mem, const uint8_t* __restrict__ done,
                                                     int capacity) {
    const int b = blockIdx.y;
    if (!done[b]) return;
    const int64_t n4 = static_cast<int64_t>(capacity) * 8;  // float4 per actor
    float4* base = reinterpret_cast<float4*>(mem) + static_cast<int64_t>(b) * n4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t start = static_cast<int64_t>(blockIdx.x) * 2048;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int64_t i = start + it * 256 + threadIdx.x;
        if (i < n4) base[i] = z;
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
