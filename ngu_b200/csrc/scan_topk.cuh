// scan_topk.cuh — kernel A: fused distance scan + per-warp top-k over the episodic memory.
//
// Replaces, in one pass over HBM, the reference's
//     euc_dist = ((episodic_memory - controllable_state.unsqueeze(0))**2).sum(dim=-1).sqrt()
//     k_neighbor = euc_dist.topk(k, dim=0)
// (ngu/models/intrinsic_novelty/episodic_novelty/episodic_novelty.py:51-54), which
// materialises two N*B*32 temporaries and one N*B distance matrix.
//
// Structure (B200 / sm_100a):
//   * memory is actor-major [B][N][32] f32: each actor's ring is a contiguous
//     stream of 128-byte rows; the whole buffer is one 2-D TMA tensor
//     [B*N rows][32 floats] with SWIZZLE_128B.
//   * every warp is an independent producer+consumer pipeline: lane 0 issues
//     `cp.async.bulk.tensor.2d` (TMA) for TILE_ROWS x 128 B tiles into the
//     warp's private ring of STAGES shared-memory slots, completion is signalled
//     on one mbarrier per slot.  No __syncthreads in the main loop, no empty
//     barriers (the warp that reads a slot is the one that refills it).
//   * thread-per-row consumption: lane r reads row r of the tile with eight
//     LDS.128; the 128B hardware swizzle (chunk ^ (row & 7)) makes every quarter
//     warp hit 32 distinct banks.  Owning a whole row is what lets one thread
//     reproduce torch-CPU's exact add order (below) without any shuffles.
//   * the query (32 floats) stays in registers for the whole actor segment;
//     the running k-th-best threshold is a warp-uniform register (WarpTopK).
//   * work = B * ceil(N / TILE_ROWS) tiles, cut into equal contiguous chunks of
//     `tiles_per_warp`; a warp that crosses an actor boundary flushes its list
//     and starts a new one.  Partial lists go to partial[B][max_parts][k]; the
//     LAST warp to flush a given actor (per-actor arrival counter) merges that
//     actor's partial lists in-kernel and writes the final k-NN (d^2, slot, dist),
//     so no separate merge launch sits on the step's critical path.
//   * programmatic dependent launch: the kernel lets its dependent (the epilogue
//     kernel) get resident immediately and itself waits on its predecessor only
//     right before the first TMA load (griddepcontrol).
//
// Arithmetic order (bit-exact with torch 2.11 CPU):
//   diff = m - q ; s = diff*diff (separately rounded, no FMA);
//   a[l] = ((s[l] + s[l+8]) + s[l+16]) + s[l+24], l = 0..7 ;
//   d2 = ((((((a0+a1)+a2)+a3)+a4)+a5)+a6)+a7.
#pragma once
#include <cuda.h>
#include <cstdint>

#include "warp_topk.cuh"

namespace ngu {

struct ScanParams {
    const float* q;          // [B][32]
    u64* partial;            // [B][max_parts][k]
    int32_t n_actors;        // B
    int32_t capacity;        // N
    int32_t k;
    int32_t largest;         // 1 = reference mode (farthest), 0 = paper (nearest)
    int32_t tiles_per_actor; // ceil(N / TILE_ROWS)
    int32_t tiles_per_warp;  // chunk length
    int32_t max_parts;       // partial lists reserved per actor
    int32_t sqrt_dist;       // 1: dist = sqrt(d2) (reference), 0: dist = d2 (paper)
    int64_t total_tiles;     // B * tiles_per_actor
    unsigned int* actor_arrivals;  // [B] zero between launches
    float* topk_d2;          // [B][k]
    int32_t* topk_idx;       // [B][k]
    float* dist;             // [B][k]
    unsigned long long* timeline;  // debug (tools/scan_timeline.py): per warp {start, first tile, scan end, flush end} ns, or null
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Release-only arrival.  Deliberately not __threadfence() + atomicAdd: __threadfence() compiles to
// MEMBAR.SC.GPU + CCTL.IVALL (a sequentially consistent fence plus an L1 invalidate); the release
// form is MEMBAR.ALL.GPU + ATOMG and is all the writer side needs.
__device__ __forceinline__ unsigned atom_add_release_gpu(unsigned* addr, unsigned v) {
    unsigned old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;"
                 : "=r"(old) : "l"(__cvta_generic_to_global(addr)), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Final k-NN of one actor from the warp-resident list.
__device__ __forceinline__ void write_knn(const ScanParams& p, int actor, const WarpTopK& top, int lane) {
    if (lane < p.k) {
        const bool largest = p.largest != 0;
        const float d2 = key_d2(top.lkey, largest);
        const int64_t o = static_cast<int64_t>(actor) * p.k + lane;
        p.topk_d2[o] = d2;
        p.topk_idx[o] = key_slot(top.lkey);
        p.dist[o] = p.sqrt_dist ? __fsqrt_rn(d2) : d2;   // episodic_novelty.py:52 .sqrt(), winners only
    }
}

// Called by a whole warp when it has finished its share of `actor`.
__device__ __forceinline__ void flush_actor(const ScanParams& p, int actor, int64_t gw, WarpTopK& top, int lane) {
    const int64_t t0 = static_cast<int64_t>(actor) * p.tiles_per_actor;
    const int first = static_cast<int>(t0 / p.tiles_per_warp);
    const int last = static_cast<int>((t0 + p.tiles_per_actor - 1) / p.tiles_per_warp);
    const int nparts = last - first + 1;
    if (nparts == 1) {           // this warp scanned the whole actor
        write_knn(p, actor, top, lane);
        return;
    }
    const int k = p.k;
    u64* base = p.partial + static_cast<int64_t>(actor) * p.max_parts * k;
    if (lane < k) base[(gw - first) * k + lane] = top.lkey;
    __syncwarp();    // orders every lane's list store before lane 0's release below
    unsigned prev = 0;
    if (lane == 0) prev = atom_add_release_gpu(p.actor_arrivals + actor, 1u);
    prev = __shfl_sync(kFull, prev, 0);
    if (prev != static_cast<unsigned>(nparts - 1)) return;
    // last arrival: every partial list of this actor is in L2 (acquire side: one full fence, B times per launch)
    __threadfence();
    if (lane == 0) p.actor_arrivals[actor] = 0;   // re-arm for the next launch
    // Pass 1: every partial list that is full proves "k keys >= its k-th entry exist", so the largest
    // of those k-th entries T is a lower bound of the final k-th best key.  Seeding the merge with it
    // (floor = T-1: keys are integers and T itself must stay eligible) leaves only a few dozen real
    // candidates however many lists there are, instead of sorting all nparts*k keys.
    u64 T = 0ull;
    for (int l = lane; l < nparts; l += 32) T = umax64(T, __ldcg(base + l * k + (k - 1)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) T = umax64(T, __shfl_xor_sync(kFull, T, o));
    WarpTopK m;
    m.reset(T ? T - 1ull : 0ull);
    // Pass 2: offer all keys, four independent loads in flight per lane.
    const int nkeys = nparts * k;
    for (int i0 = 0; i0 < nkeys; i0 += 128) {
        u64 key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u + lane;
            key[u] = i < nkeys ? __ldcg(base + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) m.offer(key[u], lane, k);
    }
    write_knn(p, actor, m, lane);
}

// ---- PTX wrappers ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// The memory stream is read exactly once per step and is ~8x the L2: tag it evict-first so
// it does not flush the small hot data (queries, partial lists, running statistics, the
// epilogue's inputs) out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0,
                                            int32_t c1, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}

__device__ __forceinline__ float sqdiff(float m, float q) {
    const float d = __fsub_rn(m, q);
    return __fmul_rn(d, d);
}

// d^2 of one swizzled 128-byte row against the register-resident query.
// row_addr = shared address of the row start, sw = row & 7.
__device__ __forceinline__ float row_d2(uint32_t row_addr, uint32_t sw, const float (&q)[32]) {
    float a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 m = lds128(row_addr + ((static_cast<uint32_t>(c) ^ sw) << 4));
        const int l0 = 4 * (c & 1);
        const float s0 = sqdiff(m.x, q[4 * c + 0]);
        const float s1 = sqdiff(m.y, q[4 * c + 1]);
        const float s2 = sqdiff(m.z, q[4 * c + 2]);
        const float s3 = sqdiff(m.w, q[4 * c + 3]);
        if (c < 2) {  // first contribution to each lane accumulator: 0 + s == s
            a[l0 + 0] = s0; a[l0 + 1] = s1; a[l0 + 2] = s2; a[l0 + 3] = s3;
        } else {
            a[l0 + 0] = __fadd_rn(a[l0 + 0], s0);
            a[l0 + 1] = __fadd_rn(a[l0 + 1], s1);
            a[l0 + 2] = __fadd_rn(a[l0 + 2], s2);
            a[l0 + 3] = __fadd_rn(a[l0 + 3], s3);
        }
    }
    float h = a[0];
#pragma unroll
    for (int l = 1; l < 8; ++l) h = __fadd_rn(h, a[l]);
    return h;
}

template <int WARPS, int STAGES, int RPL>
struct ScanCfg {
    static constexpr int kWarps = WARPS;
    static constexpr int kStages = STAGES;
    static constexpr int kTileRows = 32 * RPL;
    static constexpr int kTileBytes = kTileRows * 128;
    static constexpr int kThreads = WARPS * 32;
    // tiles + 1024 B alignment slack + barriers
    static constexpr int kSmemBytes = WARPS * STAGES * kTileBytes + 1024 + WARPS * STAGES * 8;
};

template <int WARPS, int STAGES, int RPL>
__global__ void __launch_bounds__(WARPS * 32) scan_topk_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                const ScanParams p) {
    using Cfg = ScanCfg<WARPS, STAGES, RPL>;
    extern __shared__ uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // 128B-swizzle atoms are 1024 B: align the tile area.
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles0 = (raw + 1023u) & ~1023u;
    const uint32_t bars0 = tiles0 + WARPS * STAGES * Cfg::kTileBytes;
    const uint32_t my_tiles = tiles0 + warp * STAGES * Cfg::kTileBytes;
    const uint32_t my_bars = bars0 + warp * STAGES * 8;

    pdl_launch_dependents();   // the epilogue kernel may become resident now; it waits on us
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
    const unsigned long long tl0 = p.timeline ? globaltimer_ns() : 0ull;
    unsigned long long tl1 = 0ull;
    const int64_t t_begin = gw * p.tiles_per_warp;
    int64_t t_end = t_begin + p.tiles_per_warp;
    if (t_end > p.total_tiles) t_end = p.total_tiles;
    if (t_begin >= t_end) return;  // whole warp exits together; no CTA-wide barriers below

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(my_bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    pdl_wait();   // predecessor (previous step's epilogue: append, readers of dist/topk) is complete

    const int tpa = p.tiles_per_actor;
    const int N = p.capacity;
    int b = static_cast<int>(t_begin / tpa);          // actor of the current tile
    int ti = static_cast<int>(t_begin - static_cast<int64_t>(b) * tpa);  // tile index inside the actor

    // Prologue: fill the ring.  (Producer cursor runs STAGES tiles ahead.)
    const uint64_t policy = l2_evict_first_policy();
    int pb = b, pti = ti;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            if (t_begin + s < t_end) {
                const uint32_t bar = my_bars + 8 * s;
                mbar_expect_tx(bar, Cfg::kTileBytes);
                tma_load_2d(my_tiles + s * Cfg::kTileBytes, &tmap, 0,
                            pb * N + pti * Cfg::kTileRows, bar, policy);
                if (++pti == tpa) { pti = 0; ++pb; }
            }
        }
    }

    const bool largest = p.largest != 0;
    const int k = p.k;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    float q[32];
    WarpTopK top;
    int cur = -1;    // actor whose list is being built
    int stage = 0;
    uint32_t parity = 0;

    for (int64_t t = t_begin; t < t_end; ++t) {
        if (b != cur) {
            if (cur >= 0) flush_actor(p, cur, gw, top, lane);
            cur = b;
            top.reset();
            const float4* qv = reinterpret_cast<const float4*>(p.q + static_cast<int64_t>(b) * 32);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = __ldg(qv + c);
                q[4 * c + 0] = v.x; q[4 * c + 1] = v.y; q[4 * c + 2] = v.z; q[4 * c + 3] = v.w;
            }
        }

        mbar_wait(my_bars + 8 * stage, parity);
        if (p.timeline && t == t_begin) tl1 = globaltimer_ns();
        const uint32_t tile = my_tiles + stage * Cfg::kTileBytes;
        const int row0 = ti * Cfg::kTileRows;
        float d2[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i)
            d2[i] = row_d2(tile + static_cast<uint32_t>(lane + 32 * i) * 128u, sw, q);
        __syncwarp();  // every lane is done reading this slot

        // Refill the slot with the tile STAGES ahead.
        if (lane == 0 && t + STAGES < t_end) {
            const uint32_t bar = my_bars + 8 * stage;
            mbar_expect_tx(bar, Cfg::kTileBytes);
            tma_load_2d(tile, &tmap, 0, pb * N + pti * Cfg::kTileRows, bar, policy);
            if (++pti == tpa) { pti = 0; ++pb; }
        }

#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int slot = row0 + lane + 32 * i;
            const u64 key = slot < N ? make_key(d2[i], static_cast<uint32_t>(slot), largest) : 0ull;
            top.offer(key, lane, k);
        }

        if (++stage == STAGES) { stage = 0; parity ^= 1u; }
        if (++ti == tpa) { ti = 0; ++b; }
    }

    const unsigned long long tl2 = p.timeline ? globaltimer_ns() : 0ull;
    flush_actor(p, cur, gw, top, lane);
    if (p.timeline && lane == 0) {
        unsigned long long* o = p.timeline + gw * 4;
        o[0] = tl0; o[1] = tl1; o[2] = tl2; o[3] = globaltimer_ns();
    }
}

}  // namespace ngu
