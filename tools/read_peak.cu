// read_peak.cu — what can a READ-ONLY stream reach on this B200?  (measurement aid, not product)
//
// The roofline denominator in MEASURED_PEAKS.json is a copy (read + write).  The scan kernel only
// reads, so this probes the ceiling for its access pattern: a 983 MB buffer (= 256 actors x 30 000
// slots x 128 B) streamed once per launch by
//   ldg      : grid-stride ld.global.nc.L1::no_allocate.v4, UNROLL independent loads per thread
//   bulk/C   : per-warp rings of cp.async.bulk (1-D TMA) tiles, each warp owns a CONTIGUOUS chunk (the scan's layout)
//   bulk/I   : same rings, tiles interleaved over all warps (the whole grid reads one moving window)
// plus cudaMemcpy D2D and cudaMemset for reference.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int UNROLL>
__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ src, size_t n16, unsigned* out) {
    unsigned acc = 0;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += stride) { const uint4 v = src[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *out = acc;
}

// per-warp ring of 1-D bulk copies; lane 0 produces, the warp consumes one word per lane per 128 B
template <int WARPS, int STAGES, int TILE_BYTES, bool INTERLEAVED, bool EVICT_FIRST>
__global__ void __launch_bounds__(WARPS * 32) bulk_kernel(const char* __restrict__ src, long long n_tiles,
                                                           long long tiles_per_warp, unsigned* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tiles0 = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t my_tiles = tiles0 + warp * STAGES * TILE_BYTES;
    const uint32_t my_bars = tiles0 + WARPS * STAGES * TILE_BYTES + warp * STAGES * 8;
    const long long gw = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const long long n_warps = static_cast<long long>(gridDim.x) * WARPS;
    long long t, t_end, t_step;
    if (INTERLEAVED) { t = gw; t_end = n_tiles; t_step = n_warps; }
    else { t = gw * tiles_per_warp; t_end = t + tiles_per_warp; if (t_end > n_tiles) t_end = n_tiles; t_step = 1; }
    if (t >= t_end) return;
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(my_bars + 8 * s), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    auto issue = [&](long long tile, int s) {
        const uint32_t bar = my_bars + 8 * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TILE_BYTES) : "memory");
        if (EVICT_FIRST)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(my_tiles + s * TILE_BYTES), "l"(src + tile * TILE_BYTES), "r"(TILE_BYTES), "r"(bar), "l"(pol) : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(my_tiles + s * TILE_BYTES), "l"(src + tile * TILE_BYTES), "r"(TILE_BYTES), "r"(bar) : "memory");
    };
    long long pt = t;
    if (lane == 0)
        for (int s = 0; s < STAGES && pt < t_end; ++s, pt += t_step) issue(pt, s);
    pt = t + STAGES * t_step;
    int stage = 0; uint32_t parity = 0; unsigned acc = 0;
    for (; t < t_end; t += t_step) {
        const uint32_t bar = my_bars + 8 * stage;
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                     ::"r"(bar), "r"(parity) : "memory");
        const uint32_t tile = my_tiles + stage * TILE_BYTES;
#pragma unroll
        for (int r = 0; r < TILE_BYTES / 128 / 32; ++r) {
            unsigned v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tile + (r * 32 + lane) * 128 + (lane & 7) * 16));
            acc ^= v;
        }
        __syncwarp();
        if (lane == 0 && pt < t_end) issue(pt, stage);
        pt += t_step;
        if (++stage == STAGES) { stage = 0; parity ^= 1u; }
    }
    if (acc == 0x12345678u) *out = acc;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    template <class F> double best_ms(F f, int reps = 20) {
        double best = 1e30, sum = 0;
        for (int i = 0; i < 3; ++i) f();
        for (int i = 0; i < reps; ++i) {
            cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
            sum += ms;
        }
        last_avg = sum / reps;
        return best;
    }
    double last_avg = 0;
};


// Dynamic scheduling: chunk idx -> tile range by a 2-phase closed form (n1 chunks of s1 tiles handed out
// statically, one per warp, then chunks of c2 tiles from an atomic counter).  Lane 0 is the producer and
// records what it put into each ring stage in a per-stage descriptor; -1 = no more work.
struct DynSched { long long n_tiles; int n1, s1, c2; long long n_chunks; };
__device__ __forceinline__ void chunk_range(const DynSched& d, long long idx, long long& t, long long& te) {
    if (idx < d.n1) { t = idx * d.s1; te = t + d.s1; }
    else { t = static_cast<long long>(d.n1) * d.s1 + (idx - d.n1) * d.c2; te = t + d.c2; }
    if (te > d.n_tiles) te = d.n_tiles;
}
template <int WARPS, int STAGES, int TILE_BYTES>
__global__ void __launch_bounds__(WARPS * 32) bulk_dyn_kernel(const char* __restrict__ src, const DynSched d,
                                                               unsigned* counter, unsigned* done, unsigned* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tiles0 = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t my_tiles = tiles0 + warp * STAGES * TILE_BYTES;
    const uint32_t my_bars = tiles0 + WARPS * STAGES * TILE_BYTES + warp * STAGES * 8;
    const uint32_t my_desc = tiles0 + WARPS * STAGES * TILE_BYTES + WARPS * STAGES * 8 + warp * STAGES * 8;
    const long long gw = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const long long n_warps = static_cast<long long>(gridDim.x) * WARPS;
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(my_bars + 8 * s), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    // producer state (lane 0)
    long long t = 0, te = 0, pending = d.n_chunks;
    bool live = gw < d.n_chunks;
    if (lane == 0 && live) {
        chunk_range(d, gw, t, te);
        pending = n_warps + atomicAdd(counter, 1u);
    }
    auto produce = [&](int s) {   // lane 0
        long long tile = -1;
        if (live) {
            if (t == te) {
                if (pending < d.n_chunks) { chunk_range(d, pending, t, te); pending = n_warps + atomicAdd(counter, 1u); }
                else live = false;
            }
            if (live) tile = t++;
        }
        asm volatile("st.shared.s64 [%0], %1;" ::"r"(my_desc + 8 * s), "l"(tile) : "memory");
        if (tile >= 0) {
            const uint32_t bar = my_bars + 8 * s;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TILE_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(my_tiles + s * TILE_BYTES), "l"(src + tile * TILE_BYTES), "r"(TILE_BYTES), "r"(bar), "l"(pol) : "memory");
        }
    };
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) produce(s);
    __syncwarp();
    int stage = 0; uint32_t parity = 0; unsigned acc = 0;
    for (;;) {
        long long tile;
        asm volatile("ld.shared.s64 %0, [%1];" : "=l"(tile) : "r"(my_desc + 8 * stage) : "memory");
        if (tile < 0) break;
        const uint32_t bar = my_bars + 8 * stage;
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                     ::"r"(bar), "r"(parity) : "memory");
        const uint32_t tb = my_tiles + stage * TILE_BYTES;
#pragma unroll
        for (int r = 0; r < TILE_BYTES / 128 / 32; ++r) {
            unsigned v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tb + (r * 32 + lane) * 128 + (lane & 7) * 16));
            acc ^= v;
        }
        __syncwarp();
        if (lane == 0) produce(stage);
        __syncwarp();
        if (++stage == STAGES) { stage = 0; parity ^= 1u; }
    }
    if (lane == 0) {
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == n_warps - 1) { *counter = 0; *done = 0; }
    }
    if (acc == 0x12345678u) *out = acc;
}

template <int W, int S, int TB>
void run_dyn(const char* name, const char* src, size_t bytes, int ctas_per_sm, int sms, double static_frac, int c2,
             unsigned* ctr, unsigned* out, Timer& tm) {
    DynSched d;
    d.n_tiles = bytes / TB;
    const int grid = sms * ctas_per_sm;
    const long long warps = static_cast<long long>(grid) * W;
    d.n1 = static_cast<int>(warps);
    d.s1 = static_cast<int>(static_frac * d.n_tiles / warps);
    d.c2 = c2;
    if (d.s1 < 1) { d.n1 = 0; d.s1 = 1; }
    const long long rest = d.n_tiles - static_cast<long long>(d.n1) * d.s1;
    d.n_chunks = d.n1 + (rest + c2 - 1) / c2;
    const int smem = W * S * TB + 1024 + W * S * 16;
    CK(cudaFuncSetAttribute(bulk_dyn_kernel<W, S, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const double ms = tm.best_ms([&] { bulk_dyn_kernel<W, S, TB><<<grid, W * 32, smem>>>(src, d, ctr, ctr + 1, out); });
    CK(cudaGetLastError());
    printf("%-36s static %.2f (s1=%3d) c2=%2d  grid %4d  best %.1f us  %.0f GB/s  (avg %.0f GB/s)\n", name, static_frac, d.s1, c2, grid,
           ms * 1e3, bytes / ms / 1e6, bytes / tm.last_avg / 1e6);
}

template <int W, int S, int TB, bool I, bool E>
void run_bulk(const char* name, const char* src, size_t bytes, int ctas_per_sm, int sms, unsigned* out, Timer& tm) {
    const long long n_tiles = bytes / TB;
    const int grid = sms * ctas_per_sm;
    const long long warps = static_cast<long long>(grid) * W;
    const long long tpw = (n_tiles + warps - 1) / warps;
    const int smem = W * S * TB + 1024 + W * S * 8;
    CK(cudaFuncSetAttribute(bulk_kernel<W, S, TB, I, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const double ms = tm.best_ms([&] { bulk_kernel<W, S, TB, I, E><<<grid, W * 32, smem>>>(src, n_tiles, tpw, out); });
    CK(cudaGetLastError());
    printf("%-44s grid %4d x %3d thr  smem %6d  best %.1f us  %.0f GB/s  (avg %.0f GB/s)\n", name, grid, W * 32, smem, ms * 1e3,
           bytes / ms / 1e6, bytes / tm.last_avg / 1e6);
}

int main() {
    const size_t bytes = 256ull * 30000 * 128;
    char *src, *dst; unsigned* out;
    CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(src, 1, bytes));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    Timer tm;
    double ms = tm.best_ms([&] { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice); });
    printf("%-44s best %.1f us  %.0f GB/s (read+write)  (avg %.0f)\n", "cudaMemcpy D2D", ms * 1e3, 2.0 * bytes / ms / 1e6, 2.0 * bytes / tm.last_avg / 1e6);
    ms = tm.best_ms([&] { cudaMemsetAsync(dst, 0, bytes); });
    printf("%-44s best %.1f us  %.0f GB/s (write)\n", "cudaMemset", ms * 1e3, bytes / ms / 1e6);
    const size_t n16 = bytes / 16;
    for (int cps : {2, 4}) {
        ms = tm.best_ms([&] { ldg_kernel<4><<<sms * cps, 512>>>(reinterpret_cast<const uint4*>(src), n16, out); });
        printf("ldg v4 unroll 4, %d CTAs/SM x 512                best %.1f us  %.0f GB/s  (avg %.0f)\n", cps, ms * 1e3, bytes / ms / 1e6, bytes / tm.last_avg / 1e6);
        ms = tm.best_ms([&] { ldg_kernel<8><<<sms * cps, 512>>>(reinterpret_cast<const uint4*>(src), n16, out); });
        printf("ldg v4 unroll 8, %d CTAs/SM x 512                best %.1f us  %.0f GB/s  (avg %.0f)\n", cps, ms * 1e3, bytes / ms / 1e6, bytes / tm.last_avg / 1e6);
    }
    CK(cudaGetLastError());
    run_bulk<8, 3, 4096, false, true>("bulk 4K x3 x8w, contiguous/warp, evict_first", src, bytes, 2, sms, out, tm);
    run_bulk<8, 3, 4096, false, false>("bulk 4K x3 x8w, contiguous/warp, default", src, bytes, 2, sms, out, tm);
    run_bulk<8, 3, 4096, true, true>("bulk 4K x3 x8w, interleaved, evict_first", src, bytes, 2, sms, out, tm);
    run_bulk<8, 3, 4096, true, false>("bulk 4K x3 x8w, interleaved, default", src, bytes, 2, sms, out, tm);
    run_bulk<8, 2, 4096, false, true>("bulk 4K x2 x8w (128K/SM), contiguous", src, bytes, 2, sms, out, tm);
    run_bulk<8, 6, 4096, false, true>("bulk 4K x6 x8w 1CTA (192K/SM), contiguous", src, bytes, 1, sms, out, tm);
    run_bulk<4, 6, 8192, false, true>("bulk 8K x6 x4w 1CTA (192K/SM), contiguous", src, bytes, 1, sms, out, tm);
    run_bulk<4, 3, 16384, false, true>("bulk 16K x3 x4w 1CTA (192K/SM), contiguous", src, bytes, 1, sms, out, tm);
    run_bulk<4, 3, 16384, true, true>("bulk 16K x3 x4w 1CTA (192K/SM), interleaved", src, bytes, 1, sms, out, tm);
    run_bulk<2, 3, 32768, true, true>("bulk 32K x3 x2w 1CTA (192K/SM), interleaved", src, bytes, 1, sms, out, tm);
    run_bulk<8, 2, 2048, false, true>("bulk 2K x2 x8w 2CTA (64K/SM), contiguous", src, bytes, 2, sms, out, tm);
    run_bulk<8, 4, 2048, false, true>("bulk 2K x4 x8w 2CTA (128K/SM), contiguous", src, bytes, 2, sms, out, tm);
    run_bulk<8, 2, 4096, false, true>("bulk 4K x2 x8w 1CTA (64K/SM), contiguous", src, bytes, 1, sms, out, tm);

    unsigned* ctr; CK(cudaMalloc(&ctr, 8)); CK(cudaMemset(ctr, 0, 8));
    for (double sf : {0.0, 0.5, 0.65, 0.8, 1.0})
        for (int c2 : {4, 8, 16}) {
            if (sf == 1.0 && c2 != 4) continue;
            run_dyn<8, 3, 4096>("dyn 4K x3 x8w 2CTA", src, bytes, 2, sms, sf, c2, ctr, out, tm);
        }
    for (double sf : {0.0, 0.65})
        for (int c2 : {4, 8, 16}) {
            run_dyn<8, 2, 4096>("dyn 4K x2 x8w 2CTA (128K/SM)", src, bytes, 2, sms, sf, c2, ctr, out, tm);
            run_dyn<8, 2, 4096>("dyn 4K x2 x8w 3CTA (192K/SM)", src, bytes, 3, sms, sf, c2, ctr, out, tm);
        }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
