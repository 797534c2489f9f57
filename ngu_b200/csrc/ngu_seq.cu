// ngu_seq.cu — device-resident sequence collector (include/ngu_seq.h): the buffering half of
// NGUAgent.collect_sequence / push_collected (ngu/models/ngu/ngu_agent.py:56-94, :96-155) on the GPU.
//
// Four small launches per environment step, all on the caller's stream:
//   record   grid (B, chunks)  the step's transition -> each actor's local buffers at its own index   (:76-80)
//   plan     1 CTA             which actors have a full sequence (idx == L-1), their positions in the store ring (:98-102)
//   extract  grid (B, chunks)  collected actors: augmented / n-step rewards, first seq_len steps -> the store (:103-138)
//   advance  grid (B, chunks)  overlap carry (:141-155), idx = (idx + 1) % L (:82), idx = 0 where done (:349-350)
// Byte moves of the 28 KB observations are 16-byte vectorised and coalesced; the arithmetic (two roundings for the
// augmented reward, a sequential n-step sum) follows torch's float32 evaluation order, see the kernels.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/ngu_epi.h"   // NGU_OK / NGU_E*
#include "../../include/ngu_seq.h"

namespace {

thread_local std::string g_seq_create_error;

struct SeqDims {
    int B, L, seq_len, n_step, overlap, obs, cap;
};

std::string sfmt(const char* f, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

// flat float copy, this CTA's slice of [0, n): 16-byte pieces where both sides are aligned
__device__ __forceinline__ void copy_slice(float* __restrict__ dst, const float* __restrict__ src, int64_t n, int part, int parts) {
    const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0 && (n & 3) == 0;
    if (vec) {
        const int64_t n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int64_t i = static_cast<int64_t>(part) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(parts) * blockDim.x)
            d4[i] = s4[i];
    } else {
        for (int64_t i = static_cast<int64_t>(part) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(parts) * blockDim.x)
            dst[i] = src[i];
    }
}

// ngu_agent.py:76-80, per actor: buffers[b][idx[b]] = this step's values
__global__ void __launch_bounds__(256) seq_record_kernel(ngu_seq_views v, SeqDims d, const float* __restrict__ prev_obs,
                                                          const int64_t* __restrict__ prev_act, const int64_t* __restrict__ act,
                                                          const float* __restrict__ r_int, const float* __restrict__ r_ext) {
    const int b = blockIdx.x;
    const int64_t slot = v.local_mem_idx[b];
    const int64_t row = static_cast<int64_t>(b) * d.L + slot;
    copy_slice(v.state + row * d.obs, prev_obs + static_cast<int64_t>(b) * d.obs, d.obs, blockIdx.y, gridDim.y);
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        v.prev_action[row] = prev_act[b];
        v.curr_action[row] = act[b];
        v.intrinsic_reward[row] = r_int[b];
        v.extrinsic_reward[row] = r_ext[b];
    }
}

// ngu_agent.py:98-102: collected = local_mem_idx == mem_seq_len - 1, in actor order; store positions from the ring cursor
__global__ void __launch_bounds__(256) seq_plan_kernel(ngu_seq_views v, SeqDims d, int2* __restrict__ jobs, uint8_t* __restrict__ collected) {
    __shared__ int s_wcnt[8];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t cursor = v.counters[1];
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < d.B; b0 += 256) {
        const int b = b0 + threadIdx.x;
        const bool c = b < d.B && v.local_mem_idx[b] == d.L - 1;
        if (b < d.B) collected[b] = c ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, c);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_base, tot = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) off += s_wcnt[w];
            tot += s_wcnt[w];
        }
        if (c) {
            const int ord = off + __popc(bal & ((1u << lane) - 1u));
            const int pos = static_cast<int>((cursor + ord) % d.cap);
            jobs[ord] = make_int2(b, pos);
            v.last_pushed[1 + ord] = pos;
            v.st_actor[pos] = b;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int n = s_base;
        v.last_pushed[0] = n;
        v.counters[0] += n;
        v.counters[1] = (cursor + n) % d.cap;
    }
}

// ngu_agent.py:103-138 for collected actor `ord`: rewards (:111-112) and the first seq_len steps into the store
__global__ void __launch_bounds__(256) seq_extract_kernel(ngu_seq_views v, SeqDims d, const int2* __restrict__ jobs,
                                                           const float* __restrict__ beta, const float* __restrict__ disc_pow) {
    extern __shared__ float s_amgt[];                        // [L]
    const int n_jobs = v.last_pushed[0];
#pragma unroll 1
    for (int ord = blockIdx.x; ord < n_jobs; ord += gridDim.x) {
    __syncthreads();                                         // (s_amgt is reused by the next job)
    const int2 job = jobs[ord];
    const int b = job.x;
    const int64_t pos = job.y;
    const int64_t lrow = static_cast<int64_t>(b) * d.L, srow = pos * d.seq_len;
    copy_slice(v.st_state + srow * d.obs, v.state + lrow * d.obs, static_cast<int64_t>(d.seq_len) * d.obs, blockIdx.y, gridDim.y);
    if (blockIdx.y != 0) continue;
    const float bt = beta[b];
    for (int t = threadIdx.x; t < d.L; t += blockDim.x)     // amgt = extr + beta * intr: product and sum rounded separately (:111)
        s_amgt[t] = __fadd_rn(v.extrinsic_reward[lrow + t], __fmul_rn(bt, v.intrinsic_reward[lrow + t]));
    __syncthreads();
    for (int t = threadIdx.x; t < d.seq_len; t += blockDim.x) {
        v.st_prev_action[srow + t] = v.prev_action[lrow + t];
        v.st_curr_action[srow + t] = v.curr_action[lrow + t];
        v.st_intrinsic_reward[srow + t] = v.intrinsic_reward[lrow + t];
        v.st_extrinsic_reward[srow + t] = v.extrinsic_reward[lrow + t];
        v.st_augmented_reward[srow + t] = s_amgt[t];
        float acc = 0.0f;                                     // r_nstep[:, t] += (discount ** i) * rews[:, t + i], i ascending
        for (int i = 0; i < d.n_step; ++i)                    // (r2d2_actor.py:74-77)
            acc = __fadd_rn(acc, __fmul_rn(disc_pow[b * d.n_step + i], s_amgt[t + i]));
        v.st_nstep_reward[srow + t] = acc;
    }
    }
}

// ngu_agent.py:141-155 (overlap carry, idx = overlap), :82 (advance), :349-350 (idx = 0 where done)
__global__ void __launch_bounds__(256) seq_advance_kernel(ngu_seq_views v, SeqDims d, const uint8_t* __restrict__ collected,
                                                           const uint8_t* __restrict__ done) {
    const int b = blockIdx.x;
    const bool c = collected[b] != 0;
    const int64_t lrow = static_cast<int64_t>(b) * d.L;
    if (c) {
        const int64_t src = lrow + d.L - d.overlap;
        copy_slice(v.state + lrow * d.obs, v.state + src * d.obs, static_cast<int64_t>(d.overlap) * d.obs, blockIdx.y, gridDim.y);
        if (blockIdx.y == 0)
            for (int t = threadIdx.x; t < d.overlap; t += blockDim.x) {
                v.prev_action[lrow + t] = v.prev_action[src + t];
                v.curr_action[lrow + t] = v.curr_action[src + t];
                v.intrinsic_reward[lrow + t] = v.intrinsic_reward[src + t];
                v.extrinsic_reward[lrow + t] = v.extrinsic_reward[src + t];
            }
    }
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        int64_t i = c ? d.overlap : v.local_mem_idx[b];
        i = (i + 1) % d.L;
        if (done[b]) i = 0;
        v.local_mem_idx[b] = i;
    }
}

}  // namespace

struct ngu_seq {
    ngu_seq_cfg cfg{};
    SeqDims dims{};
    ngu_seq_views v{};
    int2* jobs = nullptr;
    uint8_t* collected = nullptr;
    float* beta = nullptr;
    float* disc_pow = nullptr;
    std::string err;
};

namespace {
int sfail(ngu_seq* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_seq_create_error = msg;
    return code;
}
struct SeqDeviceGuard {
    int prev = -1;
    explicit SeqDeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess || prev == dev || cudaSetDevice(dev) != cudaSuccess) prev = -1;
    }
    ~SeqDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
template <class T>
cudaError_t zalloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc(p, n * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
}
}  // namespace

extern "C" {

const char* ngu_seq_last_error(const ngu_seq* h) { return h ? h->err.c_str() : g_seq_create_error.c_str(); }

int ngu_seq_create(const ngu_seq_cfg* cfg, ngu_seq** out) {
    if (!cfg || !out) return sfail(nullptr, NGU_EINVAL, "null cfg/out");
    *out = nullptr;
    if (cfg->n_actors < 1 || cfg->seq_len < 1 || cfg->n_step < 1 || cfg->overlap < 0 || cfg->obs_elems < 1 || cfg->store_capacity < 1)
        return sfail(nullptr, NGU_EINVAL, "n_actors, seq_len, n_step, obs_elems, store_capacity must be >= 1 and overlap >= 0");
    const int L = cfg->seq_len + cfg->n_step;
    if (2 * cfg->overlap > L) return sfail(nullptr, NGU_EINVAL, "2 * overlap must not exceed seq_len + n_step (the carried steps would overlap their source)");
    if (L > 12000) return sfail(nullptr, NGU_EINVAL, "seq_len + n_step too large for the on-chip reward buffer");
    if (cfg->store_capacity < cfg->n_actors) return sfail(nullptr, NGU_EINVAL, "store_capacity must hold one sequence per actor (they can all complete in one step)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return sfail(nullptr, NGU_ECUDA, "no CUDA device; the sequence collector has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return sfail(nullptr, NGU_EINVAL, "device ordinal out of range");
    ngu_seq* h = new (std::nothrow) ngu_seq();
    if (!h) return sfail(nullptr, NGU_ENOMEM, "host allocation failed");
    h->cfg = *cfg;
    h->dims = SeqDims{cfg->n_actors, L, cfg->seq_len, cfg->n_step, cfg->overlap, cfg->obs_elems, cfg->store_capacity};
    SeqDeviceGuard guard(cfg->device);
    const size_t B = cfg->n_actors, C = cfg->store_capacity, S = cfg->seq_len, O = cfg->obs_elems;
    ngu_seq_views& v = h->v;
    cudaError_t e = cudaSuccess;
#define SEQ_ALLOC(ptr, n) if (e == cudaSuccess) e = zalloc(&(ptr), (n))
    SEQ_ALLOC(v.state, B * L * O); SEQ_ALLOC(v.prev_action, B * L); SEQ_ALLOC(v.curr_action, B * L);
    SEQ_ALLOC(v.intrinsic_reward, B * L); SEQ_ALLOC(v.extrinsic_reward, B * L); SEQ_ALLOC(v.local_mem_idx, B);
    SEQ_ALLOC(v.st_state, C * S * O); SEQ_ALLOC(v.st_prev_action, C * S); SEQ_ALLOC(v.st_curr_action, C * S);
    SEQ_ALLOC(v.st_intrinsic_reward, C * S); SEQ_ALLOC(v.st_extrinsic_reward, C * S); SEQ_ALLOC(v.st_augmented_reward, C * S);
    SEQ_ALLOC(v.st_nstep_reward, C * S); SEQ_ALLOC(v.st_actor, C); SEQ_ALLOC(v.counters, 2); SEQ_ALLOC(v.last_pushed, 1 + B);
    SEQ_ALLOC(h->jobs, B); SEQ_ALLOC(h->collected, B); SEQ_ALLOC(h->beta, B); SEQ_ALLOC(h->disc_pow, B * cfg->n_step);
#undef SEQ_ALLOC
    if (e == cudaSuccess) e = cudaFuncSetAttribute(reinterpret_cast<const void*>(&seq_extract_kernel),
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, L * static_cast<int>(sizeof(float)));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        g_seq_create_error = sfmt("allocation failed: %s", cudaGetErrorString(e));
        ngu_seq_destroy(h);
        return NGU_ECUDA;
    }
    *out = h;
    return NGU_OK;
}

void ngu_seq_destroy(ngu_seq* h) {
    if (!h) return;
    SeqDeviceGuard guard(h->cfg.device);
    ngu_seq_views& v = h->v;
    cudaFree(v.state); cudaFree(v.prev_action); cudaFree(v.curr_action); cudaFree(v.intrinsic_reward); cudaFree(v.extrinsic_reward);
    cudaFree(v.local_mem_idx); cudaFree(v.st_state); cudaFree(v.st_prev_action); cudaFree(v.st_curr_action);
    cudaFree(v.st_intrinsic_reward); cudaFree(v.st_extrinsic_reward); cudaFree(v.st_augmented_reward); cudaFree(v.st_nstep_reward);
    cudaFree(v.st_actor); cudaFree(v.counters); cudaFree(v.last_pushed);
    cudaFree(h->jobs); cudaFree(h->collected); cudaFree(h->beta); cudaFree(h->disc_pow);
    delete h;
}

int ngu_seq_views_get(ngu_seq* h, ngu_seq_views* out) {
    if (!h || !out) return sfail(h, NGU_EINVAL, "null argument");
    *out = h->v;
    return NGU_OK;
}

int ngu_seq_set_factors(ngu_seq* h, const float* beta_host, const float* disc_pow_host) {
    if (!h || !beta_host || !disc_pow_host) return sfail(h, NGU_EINVAL, "null argument");
    SeqDeviceGuard guard(h->cfg.device);
    cudaError_t e = cudaMemcpy(h->beta, beta_host, sizeof(float) * h->cfg.n_actors, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->disc_pow, disc_pow_host, sizeof(float) * h->cfg.n_actors * h->cfg.n_step, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return sfail(h, NGU_ECUDA, sfmt("set_factors: %s", cudaGetErrorString(e)));
    return NGU_OK;
}

int ngu_seq_step(ngu_seq* h, const float* prev_obs_dev, const int64_t* prev_act_dev, const int64_t* act_dev,
                 const float* r_int_dev, const float* r_ext_dev, const uint8_t* done_dev, void* stream) {
    if (!h || !prev_obs_dev || !prev_act_dev || !act_dev || !r_int_dev || !r_ext_dev || !done_dev) return sfail(h, NGU_EINVAL, "null argument");
    SeqDeviceGuard guard(h->cfg.device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const SeqDims d = h->dims;
    const int obs_chunks = static_cast<int>((static_cast<int64_t>(d.obs) / 4 + 255) / 256) > 0 ? static_cast<int>((static_cast<int64_t>(d.obs) / 4 + 255) / 256) : 1;
    const int seq_chunks = obs_chunks * 2 < 16 ? obs_chunks * 2 : 16;     // a whole sequence of observations per collected actor
    const int job_ctas = d.B < 64 ? d.B : 64;                             // collected actors are dealt round robin over these
    seq_record_kernel<<<dim3(d.B, obs_chunks), 256, 0, s>>>(h->v, d, prev_obs_dev, prev_act_dev, act_dev, r_int_dev, r_ext_dev);
    seq_plan_kernel<<<1, 256, 0, s>>>(h->v, d, h->jobs, h->collected);
    seq_extract_kernel<<<dim3(job_ctas, seq_chunks), 256, d.L * sizeof(float), s>>>(h->v, d, h->jobs, h->beta, h->disc_pow);
    seq_advance_kernel<<<dim3(d.B, obs_chunks < 8 ? obs_chunks : 8), 256, 0, s>>>(h->v, d, h->collected, done_dev);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sfail(h, NGU_ECUDA, sfmt("ngu_seq_step: %s", cudaGetErrorString(e)));
    return NGU_OK;
}

}  // extern "C"
