// p2p_exchange.cuh — the per-step rank-sum exchange as peer-memory stores over NVLink,
// executed INSIDE the epilogue kernel (no NCCL launch, no extra kernel).
//
// Replaces the reference's comm.Allreduce (ngu/utils/mpi_util.py:17, reached from
// episodic_novelty.py:56 and, for the log-only trackers, :82 and intrinsic_novelty.py:32).
//
// Protocol: low-latency flagged words (the idea of NCCL's LL protocol).  A double is sent as two
// 8-byte words, each carrying 32 bits of payload and a 32-bit tag = step+1; an aligned 8-byte
// store is single-copy atomic, so the receiver can poll the word itself: no separate flag, no
// __threadfence_system, one NVLink traversal per exchange.
//
// Every rank owns an inbox in its own HBM:  inbox[2 parities][world][kSlotWords] 8-byte words.
//   send:  word (parity, my rank, 2e / 2e+1) of EVERY rank's inbox  <-  {tag, lo32} / {tag, hi32}
//   recv:  spin (bounded) on the local inbox until every word carries the tag, then sum the
//          `world` contributions in rank order — every rank adds the same f64 values in the same
//          order, so the replicated running statistics stay bit-identical across ranks.
// Parity double-buffering suffices: a rank can be at most one exchange ahead of its peers, and a
// slot is rewritten two exchanges later with a different tag.
// One process per GPU is required (a spin on words written by another kernel on the SAME GPU may
// never be co-scheduled); the single-GPU multi-shard test uses the collective path instead.
#pragma once
#include <cstdint>

namespace ngu {

constexpr int kSlotElems = 96;                  // doubles per (parity, rank) slot: [0,65) rank sums, [65,70) log sums
constexpr int kSlotWords = 2 * kSlotElems;      // 8-byte words
constexpr int kFlushElem = 80;                  // [80,85): the on-demand log flush round (own tag sequence)
constexpr int kMaxWorld = 16;
constexpr long long kSpinBudgetCycles = 4000000000ll;   // ~2 s at 1.9 GHz, then give up (counted in exchange_timeouts)

struct P2PView {
    unsigned long long* inbox;                  // local inbox, [2][world][kSlotWords]
    unsigned long long* peer[kMaxWorld];        // peer[g] = rank g's inbox as mapped in this process
    int32_t rank, world;
};

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All threads of the CTA call this.  vals[0..n) (shared memory) are this rank's contribution and
// are replaced by the sum over ranks.  `elem0` is the first slot element used, `seq` the exchange
// sequence number (same on all ranks).  recv = shared scratch of world*n doubles.
// Returns false if a peer did not show up within the spin budget.
template <int THREADS>
__device__ __forceinline__ bool p2p_allreduce(const P2PView& v, double* vals, int n, int elem0,
                                              unsigned long long seq, double* recv) {
    __shared__ int s_ok;
    const int parity = static_cast<int>(seq & 1ull);
    const unsigned long long tag = ((seq + 1ull) & 0xffffffffull) << 32;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    // send: two flagged words per value to every rank (own inbox included)
    for (int t = threadIdx.x; t < v.world * n; t += THREADS) {
        const int g = t / n, e = t - g * n;
        const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(vals[e]));
        unsigned long long* dst = v.peer[g] + (static_cast<size_t>(parity) * v.world + v.rank) * kSlotWords + 2 * (elem0 + e);
        st_relaxed_sys_u64(dst, tag | (bits & 0xffffffffull));
        st_relaxed_sys_u64(dst + 1, tag | (bits >> 32));
    }
    // receive: poll the words themselves
    for (int t = threadIdx.x; t < v.world * n; t += THREADS) {
        const int g = t / n, e = t - g * n;
        const unsigned long long* src = v.inbox + (static_cast<size_t>(parity) * v.world + g) * kSlotWords + 2 * (elem0 + e);
        const long long t0 = clock64();
        unsigned long long lo, hi;
        bool ok = true;
        while ((((lo = ld_relaxed_sys_u64(src)) ^ tag) >> 32) != 0ull) {
            if (clock64() - t0 > kSpinBudgetCycles) { ok = false; break; }
        }
        while (ok && (((hi = ld_relaxed_sys_u64(src + 1)) ^ tag) >> 32) != 0ull) {
            if (clock64() - t0 > kSpinBudgetCycles) { ok = false; break; }
        }
        if (!ok) { s_ok = 0; lo = hi = 0ull; }
        recv[g * n + e] = __longlong_as_double(static_cast<long long>((hi << 32) | (lo & 0xffffffffull)));
    }
    __syncthreads();
    // reduce in rank order
    for (int e = threadIdx.x; e < n; e += THREADS) {
        double acc = 0.0;
        for (int g = 0; g < v.world; ++g) acc += recv[g * n + e];
        vals[e] = acc;
    }
    __syncthreads();
    return s_ok != 0;
}

}  // namespace ngu
