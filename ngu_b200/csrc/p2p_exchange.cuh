// p2p_exchange.cuh — the per-step rank-sum exchange as peer-memory stores over NVLink,
// executed INSIDE the epilogue kernel (no NCCL launch, no extra kernel).
//
// Replaces the reference's comm.Allreduce (ngu/utils/mpi_util.py:17, reached from
// episodic_novelty.py:56 and, for the log-only trackers, :82 and intrinsic_novelty.py:32).
//
// Every rank owns an inbox in its own HBM:  inbox[2 parities][world][kSlotDoubles].
// One exchange round of step s on rank r:
//   1. store this rank's values into slot (s&1, r) of EVERY rank's inbox (remote stores
//      through the IPC-mapped peer pointers; own inbox included),
//   2. __threadfence_system, then publish flag = s+1 in each of those slots,
//   3. spin (bounded) until all `world` slots of the local inbox carry flag s+1,
//   4. sum the slots in rank order — every rank adds the same f64 values in the same
//      order, so the replicated running statistics stay bit-identical across ranks.
// Parity double-buffering is enough: a rank can only be one exchange ahead of its peers.
// One process per GPU is required (a spin on a flag written by another kernel on the
// SAME GPU may never be co-scheduled); the single-GPU multi-shard test uses the NCCL/gloo path.
#pragma once
#include <cstdint>

namespace ngu {

constexpr int kSlotDoubles = 128;      // [0,65) rank sums | [96,101) log sums | [126] flag round 1 | [127] flag round 2
constexpr int kSlotLog = 96;
constexpr int kSlotFlag1 = 126;
constexpr int kSlotFlag2 = 127;
constexpr int kMaxWorld = 16;
constexpr long long kSpinBudgetCycles = 4000000000ll;   // ~2 s at 1.9 GHz, then give up (sets exchange_timeouts)

struct P2PView {
    double* inbox;                      // local inbox, [2][world][kSlotDoubles]
    double* peer[kMaxWorld];            // peer[g] = rank g's inbox as mapped in this process
    int32_t rank, world;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// All threads of the CTA call this.  vals[0..n) (shared memory) are this rank's contribution and are
// replaced by the sum over ranks.  Returns false if a peer did not show up within the spin budget.
template <int THREADS>
__device__ __forceinline__ bool p2p_allreduce(const P2PView& v, double* vals, int n, int offset, int flag_idx,
                                              unsigned long long step) {
    __shared__ int s_ok;
    const int parity = static_cast<int>(step & 1ull);
    const unsigned long long want = step + 1ull;
    if (threadIdx.x == 0) s_ok = 1;
    // 1. scatter to every rank's inbox
    for (int t = threadIdx.x; t < v.world * n; t += THREADS) {
        const int g = t / n, i = t - g * n;
        v.peer[g][(parity * v.world + v.rank) * kSlotDoubles + offset + i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish
    if (threadIdx.x < v.world)
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(
                               v.peer[threadIdx.x] + (parity * v.world + v.rank) * kSlotDoubles + flag_idx),
                           want);
    // 3. wait for everyone (bounded)
    if (threadIdx.x < v.world) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(
            v.inbox + (parity * v.world + threadIdx.x) * kSlotDoubles + flag_idx);
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(f) != want) {
            if (clock64() - t0 > kSpinBudgetCycles) { s_ok = 0; break; }
        }
    }
    __syncthreads();
    // 4. reduce in rank order
    for (int i = threadIdx.x; i < n; i += THREADS) {
        double acc = 0.0;
        for (int g = 0; g < v.world; ++g)
            acc += ld_relaxed_sys_f64(v.inbox + (parity * v.world + g) * kSlotDoubles + offset + i);
        vals[i] = acc;
    }
    __syncthreads();
    return s_ok != 0;
}

}  // namespace ngu
