// warp_topk.cuh — warp-resident top-32 list of 64-bit selection keys.
//
// Selection key (one per memory slot):  hi = IEEE bits of d^2 (d^2 >= +0, so the
// unsigned bits are monotone; inverted for nearest-neighbour mode), lo =
// 0xFFFFFFFF - slot.  Larger key = better neighbour; keys of real slots are
// unique, so "ties go to the lowest slot index" is just integer order and the
// result does not depend on the order in which slots are offered.  0 is the
// sentinel (below every real key).
//
// The list lives in registers: lane j holds the j-th best key seen so far
// (descending).  A batch of 32 keys (one per lane) is offered at a time:
//   * fast path: nobody beats the current k-th best -> 2 compares + 1 vote;
//   * few candidates: serial shuffle-insert (ballot gives the rank);
//   * many candidates: bitonic sort of the batch + bitonic top-32 merge.
// Replaces torch's topk(k, dim=0) (reference episodic_novelty.py:54).
#pragma once
#include <cstdint>

namespace ngu {

using u64 = unsigned long long;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ u64 umax64(u64 a, u64 b) { return a > b ? a : b; }
__device__ __forceinline__ u64 umin64(u64 a, u64 b) { return a < b ? a : b; }

__device__ __forceinline__ u64 make_key(float d2, uint32_t slot, bool largest) {
    uint32_t hi = __float_as_uint(d2);
    if (!largest) hi = ~hi;
    return (static_cast<u64>(hi) << 32) | static_cast<u64>(0xFFFFFFFFu - slot);
}
__device__ __forceinline__ float key_d2(u64 key, bool largest) {
    uint32_t hi = static_cast<uint32_t>(key >> 32);
    if (!largest) hi = ~hi;
    return __uint_as_float(hi);
}
__device__ __forceinline__ int32_t key_slot(u64 key) {
    return static_cast<int32_t>(0xFFFFFFFFu - static_cast<uint32_t>(key));
}

struct WarpTopK {
    u64 lkey;   // lane j: j-th best key (descending over lanes)
    u64 thr;    // max(floor, key at lane k-1) (warp-uniform): a candidate must beat it
    u64 floor;  // externally known strict lower bound of every key worth keeping (0 = none)

    __device__ __forceinline__ void reset(u64 lower_bound = 0ull) { lkey = 0; thr = lower_bound; floor = lower_bound; }

    // Offer one key per lane.  `k` is warp-uniform, 1..32.  The common case (nobody
    // beats the k-th best) is 2 compares + 1 vote.  (The slow path stays inline: an
    // out-of-line call would force the list registers through local memory.)
    __device__ __forceinline__ void offer(u64 key, int lane, int k) {
        const unsigned m = __ballot_sync(kFull, key > thr);
        if (m != 0) insert(key, m, lane, k);
    }

    __device__ __forceinline__ void insert(u64 key, unsigned m, int lane, int k) {
        if (__popc(m) > 6) {
            merge_batch(((m >> lane) & 1u) ? key : 0ull, lane, k);
            return;
        }
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const u64 c = __shfl_sync(kFull, key, src);
            if (c > thr) {  // warp-uniform: thr may have risen since the ballot
                const unsigned better = __ballot_sync(kFull, lkey > c);  // prefix mask
                const int pos = __popc(better);
                const u64 up = __shfl_up_sync(kFull, lkey, 1);
                if (lane == pos) lkey = c;
                else if (lane > pos) lkey = up;
                thr = umax64(floor, __shfl_sync(kFull, lkey, k - 1));
            }
        }
    }

    // Full-width path: sort the batch descending (bitonic network over lanes),
    // then keep the 32 largest of (list U batch) with one bitonic merge.
    __device__ __forceinline__ void merge_batch(u64 x, int lane, int k) {
#pragma unroll 1
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll 1
            for (int j = size >> 1; j > 0; j >>= 1) {
                const u64 o = __shfl_xor_sync(kFull, x, j);
                const bool desc_block = (lane & size) == 0;
                const bool lower = (lane & j) == 0;
                x = (lower == desc_block) ? umax64(x, o) : umin64(x, o);
            }
        }
        // x descending.  max(list[i], x[31-i]) is bitonic and holds the top 32.
        const u64 y = __shfl_sync(kFull, x, 31 - lane);
        u64 z = umax64(lkey, y);
#pragma unroll 1
        for (int j = 16; j > 0; j >>= 1) {
            const u64 o = __shfl_xor_sync(kFull, z, j);
            z = ((lane & j) == 0) ? umax64(z, o) : umin64(z, o);
        }
        lkey = z;
        thr = umax64(floor, __shfl_sync(kFull, lkey, k - 1));
    }
};

}  // namespace ngu
