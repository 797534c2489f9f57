// ngu_ar.cu — learner-gradient all-reduce over NVLink peer memory (include/ngu_ar.h): one kernel per rank,
// reduce-scatter and all-gather fused into a single pass over the rank's slice.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/ngu_ar.h"
#include "../../include/ngu_epi.h"   // NGU_OK / NGU_E*

namespace {

constexpr int kArMaxWorld = 8;
constexpr int kArThreads = 512;
constexpr int kFlagStride = 16;                       // u64 words between the two phases' flag rows: separate 128-byte lines
constexpr long long kArSpinBudget = 4000000000ll;   // ~2 s

thread_local std::string g_ar_create_error;

struct ArView {
    float* bucket[kArMaxWorld];                 // rank g's bucket as mapped in this process
    unsigned long long* flags[kArMaxWorld];     // rank g's flag words [2 phases][kArMaxWorld]
    int rank, world;
    long long n4;                               // float4 elements per slice
};

__device__ __forceinline__ void ar_st_flag(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ar_ld_flag(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ar_ld4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ar_st4(float4* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// threads 0 .. world-1 each wait for one peer's flag of `phase`; returns (to all threads) after a CTA barrier
__device__ __forceinline__ void ar_wait_all(const ArView& v, int phase, unsigned long long seq, unsigned long long* timeouts) {
    if (threadIdx.x < v.world) {
        const unsigned long long* f = v.flags[v.rank] + phase * kFlagStride + threadIdx.x;
        const long long t0 = clock64();
        while (ar_ld_flag(f) < seq) {
            if (clock64() - t0 > kArSpinBudget) { atomicAdd(timeouts, 1ull); break; }
        }
    }
    __syncthreads();
}

template <int WORLD>
__global__ void __launch_bounds__(kArThreads) ar_allreduce_kernel(ArView v, unsigned long long seq, unsigned int* cta_ctr,
                                                                  unsigned long long* timeouts) {
    // ---- ready barrier: my gradients (written by earlier work on this stream) may be read, and nobody still reads or
    //      writes the previous round (its done barrier).  Only CTA 0 talks to the peers; the other CTAs wait for a LOCAL
    //      word it sets (hundreds of CTAs polling the line the peers' flag stores have to reach delay those stores) ----
    unsigned long long* go = reinterpret_cast<unsigned long long*>(cta_ctr) + 2;     // own 8 bytes of the counter block
    if (blockIdx.x == 0) {
        if (threadIdx.x < WORLD) ar_st_flag(v.flags[threadIdx.x] + 0 * kFlagStride + v.rank, seq);
        ar_wait_all(v, 0, seq, timeouts);
        if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(go), "l"(seq) : "memory");
    } else {
        if (threadIdx.x == 0) {
            unsigned long long g;
            const long long t0 = clock64();
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(g) : "l"(go) : "memory");
            } while (g < seq && clock64() - t0 < kArSpinBudget);
        }
        __syncthreads();
    }
    // ---- slice `rank`: sum over ranks in rank order, average, store into every rank's bucket ----
    // U elements per thread and trip so that small worlds also keep ~8 remote 16-byte loads in flight per thread
    constexpr int U = WORLD <= 2 ? 4 : (WORLD <= 4 ? 2 : 1);
    const long long base = static_cast<long long>(v.rank) * v.n4;
    const float inv = 1.0f / static_cast<float>(WORLD);
    const long long stride = static_cast<long long>(gridDim.x) * kArThreads;
    for (long long i0 = static_cast<long long>(blockIdx.x) * kArThreads + threadIdx.x; i0 < v.n4; i0 += stride * U) {
        float4 x[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < v.n4) {
#pragma unroll
                for (int g = 0; g < WORLD; ++g) x[u][g] = ar_ld4(reinterpret_cast<const float4*>(v.bucket[g]) + base + i);   // all loads in flight
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < v.n4) {
                float4 a = x[u][0];
#pragma unroll
                for (int g = 1; g < WORLD; ++g) { a.x += x[u][g].x; a.y += x[u][g].y; a.z += x[u][g].z; a.w += x[u][g].w; }
                a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
#pragma unroll
                for (int g = 0; g < WORLD; ++g) ar_st4(reinterpret_cast<float4*>(v.bucket[g]) + base + i, a);
            }
        }
    }
    // ---- done barrier: the last CTA of this rank tells the peers that its slice has landed everywhere, then waits
    //      for theirs: when the kernel completes the local bucket holds the whole average ----
    __syncthreads();
    __shared__ unsigned int s_last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(cta_ctr, 1u) == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *cta_ctr = 0u;
    if (threadIdx.x < WORLD) ar_st_flag(v.flags[threadIdx.x] + 1 * kFlagStride + v.rank, seq);
    ar_wait_all(v, 1, seq, timeouts);
}

std::string arfmt(const char* f, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

}  // namespace

struct ngu_ar {
    int device = 0;
    long long n_floats = 0, n_padded = 0;
    float* bucket = nullptr;
    unsigned long long* flags = nullptr;        // [2 phases][kFlagStride] + the timeout counter behind them
    unsigned int* cta_ctr = nullptr;
    ArView view{};
    bool connected = false;
    unsigned long long seq = 0;
    int sm_count = 0;
    void* peer_bucket[kArMaxWorld] = {};
    void* peer_flags[kArMaxWorld] = {};
    std::string err;
};

namespace {
int arfail(ngu_ar* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_ar_create_error = msg;
    return code;
}
struct ArDeviceGuard {
    int prev = -1;
    explicit ArDeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess || prev == dev || cudaSetDevice(dev) != cudaSuccess) prev = -1;
    }
    ~ArDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" {

const char* ngu_ar_last_error(const ngu_ar* h) { return h ? h->err.c_str() : g_ar_create_error.c_str(); }

int ngu_ar_create(int64_t n_floats, int32_t device, ngu_ar** out) {
    if (!out || n_floats < 1) return arfail(nullptr, NGU_EINVAL, "bad argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return arfail(nullptr, NGU_ECUDA, "no CUDA device; no CPU fallback");
    if (device < 0 || device >= ndev) return arfail(nullptr, NGU_EINVAL, "device ordinal out of range");
    ngu_ar* h = new (std::nothrow) ngu_ar();
    if (!h) return arfail(nullptr, NGU_ENOMEM, "host allocation failed");
    h->device = device;
    h->n_floats = n_floats;
    const long long quantum = 4ll * 840;                 // whole float4s per slice for EVERY world size 2..8 (lcm = 840)
    h->n_padded = (n_floats + quantum - 1) / quantum * quantum;
    ArDeviceGuard guard(device);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    if (e == cudaSuccess) e = cudaMalloc(&h->bucket, h->n_padded * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(h->bucket, 0, h->n_padded * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&h->flags, (2 * kFlagStride + 2) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(h->flags, 0, (2 * kFlagStride + 2) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&h->cta_ctr, 64);
    if (e == cudaSuccess) e = cudaMemset(h->cta_ctr, 0, 64);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        g_ar_create_error = arfmt("ngu_ar_create: %s", cudaGetErrorString(e));
        ngu_ar_destroy(h);
        return NGU_ECUDA;
    }
    *out = h;
    return NGU_OK;
}

void ngu_ar_destroy(ngu_ar* h) {
    if (!h) return;
    ArDeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    if (h->connected)
        for (int g = 0; g < h->view.world; ++g)
            if (g != h->view.rank) {
                if (h->peer_bucket[g]) cudaIpcCloseMemHandle(h->peer_bucket[g]);
                if (h->peer_flags[g]) cudaIpcCloseMemHandle(h->peer_flags[g]);
            }
    cudaFree(h->bucket); cudaFree(h->flags); cudaFree(h->cta_ctr);
    delete h;
}

int ngu_ar_buffer(ngu_ar* h, float** bucket_dev) {
    if (!h || !bucket_dev) return arfail(h, NGU_EINVAL, "null argument");
    *bucket_dev = h->bucket;
    return NGU_OK;
}

int ngu_ar_export(ngu_ar* h, void* handle_out) {
    if (!h || !handle_out) return arfail(h, NGU_EINVAL, "null argument");
    ArDeviceGuard guard(h->device);
    cudaIpcMemHandle_t a, b;
    static_assert(2 * sizeof(cudaIpcMemHandle_t) == NGU_AR_HANDLE_BYTES, "two 64-byte IPC handles");
    cudaError_t e = cudaIpcGetMemHandle(&a, h->bucket);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&b, h->flags);
    if (e != cudaSuccess) return arfail(h, NGU_ECUDA, arfmt("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)));
    std::memcpy(handle_out, &a, sizeof a);
    std::memcpy(static_cast<char*>(handle_out) + sizeof a, &b, sizeof b);
    return NGU_OK;
}

int ngu_ar_connect(ngu_ar* h, int32_t rank, int32_t world, const void* handles) {
    if (!h || !handles) return arfail(h, NGU_EINVAL, "null argument");
    if (world < 2 || world > kArMaxWorld || rank < 0 || rank >= world) return arfail(h, NGU_EINVAL, "bad rank / world (2..8 ranks)");
    if (h->connected) return arfail(h, NGU_ESTATE, "already connected");
    ArDeviceGuard guard(h->device);
    ArView& v = h->view;
    v = ArView{};
    v.rank = rank; v.world = world;
    v.n4 = h->n_padded / 4 / world;
    for (int g = 0; g < world; ++g) {
        if (g == rank) { v.bucket[g] = h->bucket; v.flags[g] = h->flags; continue; }
        cudaIpcMemHandle_t a, b;
        std::memcpy(&a, static_cast<const char*>(handles) + static_cast<size_t>(g) * NGU_AR_HANDLE_BYTES, sizeof a);
        std::memcpy(&b, static_cast<const char*>(handles) + static_cast<size_t>(g) * NGU_AR_HANDLE_BYTES + sizeof a, sizeof b);
        cudaError_t e = cudaIpcOpenMemHandle(&h->peer_bucket[g], a, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&h->peer_flags[g], b, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return arfail(h, NGU_ECUDA, arfmt("cudaIpcOpenMemHandle(rank %d): %s", g, cudaGetErrorString(e)));
        v.bucket[g] = static_cast<float*>(h->peer_bucket[g]);
        v.flags[g] = static_cast<unsigned long long*>(h->peer_flags[g]);
    }
    h->connected = true;
    return NGU_OK;
}

int ngu_ar_allreduce_avg(ngu_ar* h, void* stream) {
    if (!h) return arfail(h, NGU_EINVAL, "null argument");
    if (!h->connected) return arfail(h, NGU_ESTATE, "call ngu_ar_connect first");
    ArDeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned long long seq = ++h->seq;
    long long ctas = (h->view.n4 + kArThreads - 1) / kArThreads;
    const long long max_ctas = 2ll * h->sm_count;          // all co-resident: the done barrier's last CTA waits on peers only
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    unsigned long long* timeouts = h->flags + 2 * kFlagStride;
    const dim3 grid(static_cast<unsigned>(ctas)), block(kArThreads);
    switch (h->view.world) {
#define AR_CASE(W) case W: ar_allreduce_kernel<W><<<grid, block, 0, s>>>(h->view, seq, h->cta_ctr, timeouts); break;
        AR_CASE(2) AR_CASE(3) AR_CASE(4) AR_CASE(5) AR_CASE(6) AR_CASE(7) AR_CASE(8)
#undef AR_CASE
        default: return arfail(h, NGU_EINVAL, "unsupported world size");
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return arfail(h, NGU_ECUDA, arfmt("ngu_ar_allreduce_avg: %s", cudaGetErrorString(e)));
    return NGU_OK;
}

int64_t ngu_ar_timeouts(ngu_ar* h) {
    if (!h) return -1;
    ArDeviceGuard guard(h->device);
    unsigned long long t = 0;
    if (cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(&t, h->flags + 2 * kFlagStride, sizeof t, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return static_cast<int64_t>(t);
}

}  // extern "C"
