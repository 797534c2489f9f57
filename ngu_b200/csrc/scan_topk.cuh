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
//   * work = B * ceil(N / TILE_ROWS) tiles in one linear tile space, cut into chunks by a
//     closed-form schedule (ScanSched): every warp owns one big STATIC chunk (~60 % of its
//     fair share, no atomics), the rest is handed out DYNAMICALLY in small chunks from a
//     global counter.  Under saturation the HBM/L2 system serves the SMs unevenly (measured:
//     the same TPC pairs finish a fixed share 35 % earlier than others, tools/scan_timeline.py),
//     so equal static shares leave the kernel waiting for its slowest SMs while the rest idle;
//     with the dynamic tail every SM pulls until the memory is exhausted.  Lane 0 is the
//     producer: it walks the warp's chunks and records (actor, tile) of what it loaded into
//     each ring stage in a per-stage descriptor; the consumer flushes its list whenever the
//     actor changes.  A flush is ONE 64-bit atomic that reserves room in the actor's candidate
//     buffer and counts the tiles that have arrived; the candidates are published without a
//     fence as epoch-tagged words (see "partial results" below).  When a warp has streamed
//     its last tile it takes merge jobs (one actor each) and merges that actor's candidates
//     incrementally, as they are published, into the final k-NN (d^2, slot, dist): the merges
//     overlap the end of the stream, and no separate merge launch sits on the critical path.
//   * programmatic dependent launch: the kernel waits on its predecessor only right before
//     the first TMA load (griddepcontrol), after having prefetched its first tiles into L2, and
//     then lets its dependent (the epilogue kernel) become resident.
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

// Chunk index -> tile range, closed form (no table, no dependent load after the counter atomic):
//   chunks [0, n1)   s1 tiles each, tiles [0, n1*s1): chunk idx == global warp idx, taken statically;
//   chunks [n1, ...) the remaining tiles, cut into n2 "natural" chunks of c2 tiles that form a matrix
//                    of `rows` x `cols` (row ~ actor, cols ~ chunks per actor), handed out dynamically
//                    COLUMN BY COLUMN: consecutive handouts belong to different actors and every actor
//                    of the dynamic region is worked on during the whole dynamic phase, so later runs
//                    of an actor start from the floor its earlier runs have published (few candidates)
//                    and the merges of those actors proceed in parallel on different warps.
// Every chunk index below the number of launched warps is taken statically by that warp; the
// dynamic counter hands out the indices above.  Indices whose natural chunk is >= n2 are empty.
// (Measured on B200, 256 x 30k: s1 = 60 % of the fair share and c2 = 16 tiles; smaller chunks pay
// more in flushes than they win at the ragged end, and tapering the last chunks did not pay either.)
// A row of the matrix may taper (experiment knob NGU_EPI_TAPER, default off): its first cols_a chunks have c2
// tiles, the next cols_b have c2/2, the rest c2/4.  Handing out column by column, the narrow chunks are the last
// ones of the launch, so the warps run dry within one SHORT chunk of each other (a full chunk is ~22 us of one
// warp's time at 256 x 30k).  Measured: no gain at 256 x 30k - the extra run starts cost what the sharper end
// wins ("Measured and rejected").
struct ScanSched {
    int32_t total_tiles;     // B * tiles_per_actor
    int32_t n_chunks;        // n1 + rows * cols
    int32_t n1, s1;
    int32_t row_tiles, c2;   // tiles per matrix row = cols_a*c2 + cols_b*(c2/2) + (cols - cols_a - cols_b)*(c2/4)
    int32_t rows, cols;
    int32_t cols_a, cols_b;
};

__host__ __device__ __forceinline__ void chunk_range(const ScanSched& s, int idx, int& t, int& te) {
    if (idx < s.n1) {
        t = idx * s.s1;
        te = t + s.s1;
    } else {
        const int i = idx - s.n1;
        const int c = i / s.rows;
        const int r = i - c * s.rows;
        int off, len;
        if (c < s.cols_a) {
            off = c * s.c2; len = s.c2;
        } else if (c < s.cols_a + s.cols_b) {
            off = s.cols_a * s.c2 + (c - s.cols_a) * (s.c2 >> 1); len = s.c2 >> 1;
        } else {
            off = s.cols_a * s.c2 + s.cols_b * (s.c2 >> 1) + (c - s.cols_a - s.cols_b) * (s.c2 >> 2); len = s.c2 >> 2;
        }
        t = s.n1 * s.s1 + r * s.row_tiles + off;
        if (t >= s.total_tiles) { t = 0; te = 0; return; }
        te = t + len;
    }
    if (te > s.total_tiles) te = s.total_tiles;
}

struct ScanParams {
    const float* q;          // [B][32]
    u64* cand;               // [B][max_cand][2] epoch-tagged candidate keys (see "partial results")
    int32_t n_actors;        // B
    int32_t capacity;        // N
    int32_t k;
    int32_t largest;         // 1 = reference mode (farthest), 0 = paper (nearest)
    int32_t tiles_per_actor; // ceil(N / TILE_ROWS)
    int32_t max_cand;        // candidate entries reserved per actor (worst case: every run publishes k)
    int32_t sqrt_dist;       // 1: dist = sqrt(d2) (reference), 0: dist = d2 (paper)
    unsigned long long neg_zero2;   // (-0.0f, -0.0f): opaque addend of the packed squares, see f2_sq
    ScanSched sched;
    unsigned int* sched_ctr;       // [0] dynamic chunk counter, [1] finished warps, [3] merge-job counter, [4] leftover
                                   // merge jobs (all zero between launches), [2] launch epoch = tag of this launch's
                                   // candidates and floors (never 0)
    int32_t prefetch_tiles;        // tiles per warp prefetched into L2 before the grid dependency resolves
    int32_t cta_merge;             // static schedules: 1 = two-level merge (CTA combine in shared memory first), see static_cta_finish
    int32_t l2_mode;               // L2 eviction policy of the memory stream: 0 evict_first (read once per step, much larger
    float l2_frac;                 // than L2), 1 evict_normal, 2 evict_last for a fraction l2_frac of the lines (by address
                                   // hash: the same lines every step), rest evict_first.  Experiment knob (NGU_EPI_L2):
                                   // measured, no gain even for sets that fit the L2 ("Measured and rejected")
    int32_t* leftover;             // [B] actors whose merge job was given up (see merge_job), merged by the last warp
    long long poll_limit_cycles;   // a merge job that sees no progress for this long is given up (0: at once - tests)
    unsigned long long* timeouts;  // incremented if a merge gave up waiting for a candidate (must stay 0)
    u64* actor_state;              // [B][2]: {floor, ctr} (16 bytes, so that one bulk copy can fetch the floor)
                                   //   floor: epoch << 32 | hi word of a key known to be <= the final k-th best
                                   //   ctr  : hi32 tiles arrived, lo32 candidates reserved; zero between launches
    float* topk_d2;          // [B][k]
    int32_t* topk_idx;       // [B][k]
    float* dist;             // [B][k]
    // ---- pipelined step (ngu_epi_step_pipelined; see "software pipeline over steps" at the kernel's start) ----
    int32_t pipe;                     // 1: this launch does NOT wait for the previous step's epilogue before it streams
    uint32_t seq;                     // host sequence number of this fused step (never 0; 0 = split sequence / graph capture:
                                      // the launch takes no part in the cursor chain)
    uint32_t prev_seq;                // pipe: sequence number of the previous fused step (whose epilogue may still be running)
    uint32_t wait_epi;                // pipe: epilogue sequence number that must have completed before this launch touches
                                      // anything (the step before the previous one, if the previous one was pipelined too); 0 = none
    const unsigned int* epi_done;     // device word: sequence number of the last completed fused epilogue
    unsigned long long* scan_cursor;  // device word: seq << 32 | ring slot that step `seq` appends to, published by every scan
    const long long* cursor;          // DevState::cursor (exact once the predecessor has completed: ordinary launches)
    const float* q_prev;              // pipe: [B][32] the previous step's controllable states = what its epilogue is appending
    const unsigned int* arm_go;       // pre-armed launch (ngu_epi_step_host, see armed_stage_kernel): device word the stage kernel
    unsigned int arm_seq;             // sets to arm_seq << 1 | cancelled when the step's inputs are in place (or the step is
                                      // cancelled); arm_seq == 0: an ordinary launch
    unsigned long long* timeline;  // debug (tools/scan_timeline.py): per warp {start, first tile, scan end, flush end, smid} (8 words), or null
    int32_t debug;           // experiments only (NGU_EPI_DEBUG_SCAN): 1 = no top-k offers, 2 = no distance math, 4 = no flush,
                             // 8 = no merge, 16 = no floor seeding, 32 = no L2 prefetch before the grid dependency
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct KnnOut { float* d2; int32_t* idx; float* dist; };   // one actor's [k] outputs

// Final k-NN of one actor from a warp-resident list.  flags: 1 = largest, 2 = sqrt.
__device__ __forceinline__ void write_knn(KnnOut out, int k, int flags, u64 lkey, int lane) {
    if (lane < k) {
        const float d2 = key_d2(lkey, (flags & 1) != 0);
        out.d2[lane] = d2;
        out.idx[lane] = key_slot(lkey);
        out.dist[lane] = (flags & 2) ? __fsqrt_rn(d2) : d2;   // episodic_novelty.py:52 .sqrt(), winners only
    }
}
__device__ __forceinline__ KnnOut knn_out(const ScanParams& p, int actor) {
    const int64_t o = static_cast<int64_t>(actor) * p.k;
    return KnnOut{p.topk_d2 + o, p.topk_idx + o, p.dist + o};
}
__device__ __forceinline__ int knn_flags(const ScanParams& p) { return (p.largest ? 1 : 0) | (p.sqrt_dist ? 2 : 0); }
__device__ __forceinline__ u64* actor_floor(const ScanParams& p, int actor) { return p.actor_state + 2 * static_cast<int64_t>(actor); }
__device__ __forceinline__ u64* actor_ctr(const ScanParams& p, int actor) { return p.actor_state + 2 * static_cast<int64_t>(actor) + 1; }

// ---- partial results: fence-free, compacted, seeded ------------------------------------------------
// A run = the consecutive tiles of one actor that one warp scanned with one list.  With dynamic chunks
// an actor is covered by up to ~120 runs, so publishing and merging them must cost next to nothing:
//   * no fences.  A release fence (MEMBAR.ALL.GPU) in the middle of the stream waits for the warp's
//     in-flight TMA tiles and drains its pipeline (~9 us per flush, measured).  Instead every candidate
//     is written as two 8-byte words that each carry the launch epoch in their low half
//     (w0 = key.hi << 32 | tag, w1 = key.lo << 32 | tag; 8-byte stores are single-copy atomic), and the
//     merging warp polls the entries it is owed until both tags match - the LL protocol of NCCL.
//   * one atomic per flush.  actor_ctr[b] += tiles << 32 | n  both reserves n candidate slots (low
//     half of the return value = position) and counts arrived tiles (high half): when the tile count
//     is complete the low half is the number of candidates to merge.
//   * nobody waits for that atomic.  Its return value is first looked at after the NEXT tile has been
//     processed (complete_pending); the flushed list waits in two registers per lane meanwhile.
//   * floors.  A full list proves that its k-th key is <= the final k-th best; the hi word of the best
//     such proof is kept per actor (atomicMax, epoch in the high half so stale values lose).  A run
//     starts from it and keeps only keys above it, so late runs publish few or no candidates and the
//     merge reads a few hundred entries in ONE round trip instead of runs*k in many.
//   * nobody waits at a run start either.  The query row and the actor's floor are fetched by the
//     producer lane with two small bulk copies that complete on the SAME mbarrier as the run's first
//     tile (three tiles ahead of the consumer), instead of loads whose L2 latency under a saturated
//     memory system (1-2 us) the consumer would sit out.
__device__ __forceinline__ void st_ll(u64* dst, u64 key, uint32_t tag) {
    const u64 w0 = (key & 0xFFFFFFFF00000000ull) | tag;
    const u64 w1 = (key << 32) | tag;
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
}
// raw tagged words of one candidate; the tag check is left to the caller so that a batch of these loads
// can be issued back to back (a check right behind each load makes the warp wait out every round trip)
__device__ __forceinline__ void ld_ll_raw(const u64* src, u64& w0, u64& w1) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, u64& a, u64& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ bool ll_ready(u64 w0, u64 w1, uint32_t tag) {
    return static_cast<uint32_t>(w0) == tag && static_cast<uint32_t>(w1) == tag;
}
__device__ __forceinline__ u64 ld_volatile_u64(const u64* src) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
    return v;
}
constexpr int kMaxPrefetchTiles = 64;
constexpr int kLLSpinLimit = 1 << 22;   // ~ seconds; a healthy wait is about one tile time
// floor word -> the largest key that is certainly NOT needed (0 = no floor)
__device__ __forceinline__ u64 floor_key(u64 floor_word, uint32_t tag) {
    const uint32_t hi = static_cast<uint32_t>(floor_word);
    if (static_cast<uint32_t>(floor_word >> 32) != tag || hi == 0u) return 0ull;
    return (static_cast<u64>(hi) << 32) - 1ull;   // every key whose hi word is >= hi stays eligible
}

// A run of `run_tiles` tiles of `actor` ends (whole warp).  Returns true if a publication is pending:
// then pend_key (per lane) / pend_n hold the candidates and pend_ret (lane 0) the UNREAD result of the
// reserve+arrive atomic.
constexpr uint32_t kWrittenDirectly = 0xFFFFFFFFu;   // actor_ctr low word: the k-NN is already final
__device__ __forceinline__ bool flush_run(const ScanParams& p, int actor, int run_tiles, uint32_t tag, const WarpTopK& top,
                                          int lane, u64& pend_key, int& pend_n, u64& pend_ret) {
    const int k = p.k;
    if (run_tiles == p.tiles_per_actor) {           // this warp scanned the whole actor: no merge needed
        write_knn(knn_out(p, actor), k, knn_flags(p), top.lkey, lane);
        if (lane == 0)
            asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(actor_ctr(p, actor)),
                         "l"((static_cast<u64>(run_tiles) << 32) | kWrittenDirectly) : "memory");
        return false;
    }
    const bool have = lane < k && top.lkey != 0ull;
    pend_key = have ? top.lkey : 0ull;
    pend_n = __popc(__ballot_sync(kFull, have));    // the real keys are a prefix of the lanes
    if (pend_n == k && lane == k - 1)               // full list: its k-th key bounds the final k-th best from below
        atomicMax(actor_floor(p, actor), (static_cast<u64>(tag) << 32) | (top.lkey >> 32));
    if (lane == 0)
        pend_ret = atomicAdd(actor_ctr(p, actor), (static_cast<u64>(run_tiles) << 32) | static_cast<u64>(pend_n));
    return true;
}

// One tile later: the atomic has come back; publish the candidates at the reserved position.
__device__ __forceinline__ void complete_pending(const ScanParams& p, int actor, u64 pend_key, int pend_n, u64 pend_ret,
                                                 uint32_t tag, int lane) {
    const unsigned pos = static_cast<unsigned>(__shfl_sync(kFull, pend_ret, 0));
    u64* cand = p.cand + static_cast<int64_t>(actor) * p.max_cand * 2;
    if (pos + static_cast<unsigned>(pend_n) > static_cast<unsigned>(p.max_cand)) {   // cannot happen (max_cand is the
        if (lane == 0) atomicAdd(p.timeouts, 1ull);                                   // worst case); never write out of bounds
        return;
    }
    if (lane < pend_n) st_ll(cand + (pos + lane) * 2, pend_key, tag);
}

// Offer the candidates [from, to) of one actor to m.  They are staged through the warp's (now idle) tile
// buffers with 16-byte cp.async copies, sixteen batches of 32 in flight, and consumed from shared memory
// in a ROLLED loop.  Three simpler forms all cost ~0.45 us per batch (40 us for the 2 350 candidates of a
// B = 1 scan): a tag check right behind each load waits out every round trip; unrolled batches each run
// their own cold copy of the insert / bitonic code; a register ring cannot be rotated without reading
// registers whose loads are still in flight.
__device__ __forceinline__ void stage_candidates(const u64* cand, int from, int to, int lane, uint32_t stage) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        const int i = from + 32 * u + lane;
        if (i < to)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage + (32 * u + lane) * 16), "l"(cand + i * 2)
                         : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// staged entry e (global index base + e) -> key; polls global memory while its tags are not the epoch's
__device__ __forceinline__ u64 staged_key(const ScanParams& p, const u64* cand, int base, int e, uint32_t tag, uint32_t stage) {
    u64 w0, w1;
    lds_2x64(stage + e * 16, w0, w1);
    int polls = 0;
    while (!ll_ready(w0, w1, tag)) {                      // not yet published: poll it
        ld_ll_raw(cand + (base + e) * 2, w0, w1);
        if (++polls > kLLSpinLimit) { atomicAdd(p.timeouts, 1ull); return 0ull; }
    }
    return (w0 & 0xFFFFFFFF00000000ull) | (w1 >> 32);
}
__device__ __forceinline__ void consume_candidates(const ScanParams& p, const u64* cand, int from, int to, uint32_t tag, int lane,
                                                   uint32_t stage /* 8 KB smem */, WarpTopK& m) {
    const int k = p.k;
    for (int done = from; done < to; done += 512) {
        const int end = to - done < 512 ? to : done + 512;
        stage_candidates(cand, done, end, lane, stage);
#pragma unroll 1
        for (int e0 = 0; e0 < end - done; e0 += 32) {
            const int e = e0 + lane;
            m.offer(e < end - done ? staged_key(p, cand, done, e, tag, stage) : 0ull, lane, k);
        }
    }
}
// The same for `nlists` complete, sorted lists of k keys each (static schedules), offered RANK-MAJOR: the
// best key of every list first, then the second best, ...  The threshold is near its final value after
// the first rank, and as soon as a whole rank fails it the deeper (smaller) ranks cannot pass either: a
// merge touches 2-3 ranks and does ~15 insertions, where storage order does ~10 ln(rows / 10) of them
// (~200 cycles each, one warp) plus a threshold pass.
__device__ __forceinline__ void consume_lists_staged(const ScanParams& p, const u64* cand, int nlists, uint32_t tag, int lane,
                                                     uint32_t stage /* 8 KB smem */, WarpTopK& m) {
    const int k = p.k;
    const int per_round = 512 / k;                        // whole lists per staging round
    for (int l0 = 0; l0 < nlists; l0 += per_round) {
        const int nl = nlists - l0 < per_round ? nlists - l0 : per_round;
        stage_candidates(cand, l0 * k, (l0 + nl) * k, lane, stage);
#pragma unroll 1
        for (int r = 0; r < k; ++r) {
            unsigned passed = 0;
#pragma unroll 1
            for (int lb = 0; lb < nl; lb += 32) {
                const int l = lb + lane;
                const u64 key = l < nl ? staged_key(p, cand, l0 * k, l * k + r, tag, stage) : 0ull;
                passed |= __ballot_sync(kFull, key > m.thr);
                m.offer(key, lane, k);
            }
            if (!passed) break;
        }
    }
}

// MANY lists (an actor cut into hundreds of runs: 1 x 30 000 is 235 lists, 2 350 keys): staging everything
// costs a round trip per 51 lists and ~25 mostly useless batches (11 us, two fifths of such a step).
// Threshold first instead:
//   1. gather only the BEST key of every list (one round trip, lane l owns lists l, l+32, ...);
//   2. the k-th largest of the 32 lane maxima is a real key with >= k candidates at or above it: everything
//      below it is out (the list's floor) - one bitonic sort;
//   3. compact the survivors into a small buffer (ballot + popc), offer them 32 at a time: the threshold is
//      now the true k-th best of the rank-0 keys;
//   4. a list stays alive while its last key passed (lists are sorted): per rank, the alive lists' next keys
//      are gathered in one round trip; two or three ranks and a few dozen offered keys in all.
__device__ __forceinline__ u64 warp_sort_desc(u64 x, int lane) {
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
    return x;
}
__device__ __forceinline__ void consume_lists(const ScanParams& p, const u64* cand, int nlists, uint32_t tag, int lane,
                                              uint32_t stage /* 8 KB smem */, WarpTopK& m) {
    const int k = p.k;
    if (nlists * k <= 512 || nlists < 64) {               // everything fits one or two staging rounds
        consume_lists_staged(p, cand, nlists, tag, lane, stage, m);
        return;
    }
    constexpr int kRound = 384;                           // lists per round: 6 KB of 16-byte entries, 12 per lane
    constexpr int kBuf = 256;                             // compacted keys (2 KB) behind them
    const uint32_t buf = stage + kRound * 16;
    const unsigned lt = (1u << lane) - 1u;
    int nbuf = 0;                                         // warp-uniform
    auto drain = [&]() {
        __syncwarp();
#pragma unroll 1
        for (int e0 = 0; e0 < nbuf; e0 += 32) {
            u64 key = 0ull;
            if (e0 + lane < nbuf) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(key) : "r"(buf + (e0 + lane) * 8) : "memory");
            m.offer(key, lane, k);
        }
        nbuf = 0;
        __syncwarp();
    };
    auto push = [&](u64 key, bool q) {                    // all lanes; q: this lane's key goes in
        const unsigned bal = __ballot_sync(kFull, q);
        if (bal) {
            if (q) asm volatile("st.shared.u64 [%0], %1;" ::"r"(buf + (nbuf + __popc(bal & lt)) * 8), "l"(key) : "memory");
            nbuf += __popc(bal);
            if (nbuf + 32 > kBuf) drain();
        }
    };
#pragma unroll 1
    for (int l0 = 0; l0 < nlists; l0 += kRound) {
        const int nl = nlists - l0 < kRound ? nlists - l0 : kRound;
        const u64* base = cand + static_cast<int64_t>(l0) * k * 2;
        // rank 0 of every list of the round: one gather, every lane reads back only what it copied itself
#pragma unroll 1
        for (int i = lane; i < nl; i += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage + i * 16), "l"(base + static_cast<int64_t>(i) * k * 2)
                         : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        u64 mx = 0ull;
#pragma unroll 1
        for (int i = lane; i < nl; i += 32) {
            u64 w0, w1;
            lds_2x64(stage + i * 16, w0, w1);
            int polls = 0;
            while (!ll_ready(w0, w1, tag)) {
                ld_ll_raw(base + static_cast<int64_t>(i) * k * 2, w0, w1);
                if (++polls > kLLSpinLimit) { atomicAdd(p.timeouts, 1ull); w0 = 0ull; w1 = 0ull; break; }
            }
            const u64 key = (w0 & 0xFFFFFFFF00000000ull) | (w1 >> 32);
            asm volatile("st.shared.u64 [%0], %1;" ::"r"(stage + i * 16), "l"(key) : "memory");   // resolved key, in place
            mx = umax64(mx, key);
        }
        const u64 kth = __shfl_sync(kFull, warp_sort_desc(mx, lane), k - 1);
        if (kth > 1ull && kth - 1ull > m.thr) {           // >= k real keys are >= kth: nothing below it can make the top k
            m.floor = umax64(m.floor, kth - 1ull);
            m.thr = umax64(m.thr, m.floor);
        }
        unsigned alive = 0u;                              // bit j: list lane + 32 j of this round may still have a key that passes
#pragma unroll 1
        for (int j = 0; 32 * j < nl; ++j) {
            const int i = lane + 32 * j;
            u64 key = 0ull;
            if (i < nl) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(key) : "r"(stage + i * 16) : "memory");
            const bool q = key > m.thr;
            if (q) alive |= 1u << j;
            push(key, q);
        }
        drain();                                          // thr = the k-th best of the rank-0 keys
#pragma unroll 1
        for (int r = 1; r < k; ++r) {
            if (!__any_sync(kFull, alive != 0u)) break;
#pragma unroll 1
            for (int j = 0; 32 * j < nl; ++j)             // the alive lists' next keys: one gather
                if ((alive >> j) & 1u) {
                    const int i = lane + 32 * j;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage + i * 16),
                                 "l"(base + (static_cast<int64_t>(i) * k + r) * 2)
                                 : "memory");
                }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 1
            for (int j = 0; 32 * j < nl; ++j) {
                if (!__any_sync(kFull, (alive >> j) & 1u)) continue;
                const bool a = (alive >> j) & 1u;
                u64 key = 0ull;
                if (a) {
                    const int i = lane + 32 * j;
                    u64 w0, w1;
                    lds_2x64(stage + i * 16, w0, w1);
                    int polls = 0;
                    while (!ll_ready(w0, w1, tag)) {
                        ld_ll_raw(base + (static_cast<int64_t>(i) * k + r) * 2, w0, w1);
                        if (++polls > kLLSpinLimit) { atomicAdd(p.timeouts, 1ull); w0 = 0ull; w1 = 0ull; break; }
                    }
                    key = (w0 & 0xFFFFFFFF00000000ull) | (w1 >> 32);
                }
                const bool q = key > m.thr;
                if (a && !q) alive &= ~(1u << j);
                push(key, q);
            }
            drain();
        }
    }
}

// DYNAMIC schedules.  When a warp has streamed its last tile it takes merge jobs (one actor each) from a
// global counter until none are left.  Warps finish up to ~20 us apart, so the jobs are picked up by the
// early finishers while the rest still stream, and a job is worked INCREMENTALLY: the warp consumes the
// actor's candidates as they are published (the low half of the actor's counter says how many are
// reserved, the tags say which have landed) and keeps the running top-k in registers; when the tile
// count finally completes, only the last few candidates are left to read - one round trip after the
// last tile instead of a whole merge.  Nothing is merged inside the streaming loop, so the loop
// carries no merge registers.
// Waiting is BOUNDED.  A job polls for tiles that other warps still have to scan; should the grid ever not
// be fully co-resident (an SM limit imposed from outside, a long-running kernel of another stream holding
// shared memory), resident warps would otherwise sit polling for tiles of CTAs that cannot launch until
// they leave.  A job that sees no progress for poll_limit_cycles is therefore put on a leftover list and its
// warp moves on and exits; the LAST warp of the launch to finish - by then every tile has been scanned -
// merges the leftovers (`may_give_up` = false: nothing left to wait for but stores in flight).
// Returns false if the job was given up.
__device__ __forceinline__ bool merge_job(const ScanParams& p, int actor, uint32_t tag, int lane, uint32_t stage /* 8 KB smem */,
                                          bool may_give_up, unsigned long long* tl /* debug timeline record or null */) {
    const int k = p.k;
    u64* ctr = actor_ctr(p, actor);
    const u64* cand = p.cand + static_cast<int64_t>(actor) * p.max_cand * 2;
    WarpTopK m;
    m.reset();
    int done = 0, spins = 0;
    uint32_t n = 0;
    long long t_progress = clock64();
    const unsigned long long dbg_t0 = tl ? globaltimer_ns() : 0ull;
    for (;;) {
        // Any floor published during this launch is a valid lower bound of the final k-th best, and it
        // keeps rising while the runs flush: re-read it every round.
        const u64 fw = ld_volatile_u64(actor_floor(p, actor));
        const u64 c = ld_volatile_u64(ctr);
        const bool complete = static_cast<uint32_t>(c >> 32) == static_cast<uint32_t>(p.tiles_per_actor);
        n = static_cast<uint32_t>(c);
        if (n == kWrittenDirectly) break;                 // (only ever set together with the full tile count)
        m.floor = umax64(m.floor, floor_key(fw, tag));
        m.thr = umax64(m.thr, m.floor);
        const int avail = static_cast<int>(n);
        if (avail > done) t_progress = clock64();
        consume_candidates(p, cand, done, avail, tag, lane, stage, m);
        done = avail;
        if (complete) break;
        const int waited_too_long = __shfl_sync(kFull, clock64() - t_progress >= p.poll_limit_cycles ? 1 : 0, 0);   // lane 0 decides
        if (may_give_up && waited_too_long) {
            if (lane == 0) {
                p.leftover[atomicAdd(p.sched_ctr + 4, 1u)] = actor;
                __threadfence();                          // ... before this warp counts itself finished
            }
            return false;
        }
        if (++spins > kLLSpinLimit) { if (lane == 0) atomicAdd(p.timeouts, 1ull); break; }
    }
    __syncwarp();
    if (lane == 0) *ctr = 0ull;                           // re-arm for the next launch: every arrival has been performed
    if (tl && lane == 0) { tl[5] = dbg_t0; tl[6] = n; tl[7] = globaltimer_ns(); }   // merge job: start, candidates, end
    if (n == kWrittenDirectly || (p.debug & 8)) return true;
    write_knn(knn_out(p, actor), k, knn_flags(p), m.lkey, lane);
    return true;
}

// STATIC schedules (small shapes: every warp owns exactly one chunk of s1 tiles).  These are bound by
// latency, not bandwidth, so the protocol is the shortest one: a run's slot in the actor's candidate buffer
// follows from the chunk it belongs to, so its k keys are published at once (no reservation round trip);
// arrivals are counted in tiles; the run whose arrival completes the actor merges it - when the warp has
// finished streaming, so again nothing is merged inside the loop - and knows exactly when, no polling.
__device__ __forceinline__ bool flush_run_static(const ScanParams& p, int actor, int run_tiles, int run_first_ti, uint32_t tag,
                                                 const WarpTopK& top, int lane, u64& ret) {
    const int k = p.k;
    const int tpa = p.tiles_per_actor;
    if (run_tiles == tpa) {                         // this warp scanned the whole actor
        write_knn(knn_out(p, actor), k, knn_flags(p), top.lkey, lane);
        return false;
    }
    const int s1 = p.sched.s1, t0 = actor * tpa;
    const int slot = (t0 + run_first_ti) / s1 - t0 / s1;     // chunks of the actor before this run's chunk
    u64* cand = p.cand + static_cast<int64_t>(actor) * p.max_cand * 2;
    if ((slot + 1) * k > p.max_cand) {                       // cannot happen; never write out of bounds
        if (lane == 0) atomicAdd(p.timeouts, 1ull);
        return false;
    }
    if (lane < k) st_ll(cand + (slot * k + lane) * 2, top.lkey, tag);
    if (lane == 0) ret = atomicAdd(actor_ctr(p, actor), static_cast<u64>(run_tiles) << 32);
    return true;
}
__device__ __forceinline__ void finish_static(const ScanParams& p, int actor, int run_tiles, u64 ret, uint32_t tag, int lane,
                                              uint32_t stage, unsigned long long* tl) {
    const u64 prev = __shfl_sync(kFull, ret, 0);
    const int tpa = p.tiles_per_actor;
    if (static_cast<uint32_t>(prev >> 32) + static_cast<uint32_t>(run_tiles) != static_cast<uint32_t>(tpa)) return;
    // this run completed the actor; the other runs' keys are published or on their way (their tags are polled)
    if (lane == 0) *actor_ctr(p, actor) = 0ull;    // re-arm for the next launch
    if (p.debug & 8) return;
    const int k = p.k, s1 = p.sched.s1, t0 = actor * tpa;
    const int nparts = (t0 + tpa - 1) / s1 - t0 / s1 + 1;
    const u64* cand = p.cand + static_cast<int64_t>(actor) * p.max_cand * 2;
    WarpTopK m;
    m.reset();
    const unsigned long long dbg_t0 = tl ? globaltimer_ns() : 0ull;
    consume_lists(p, cand, nparts, tag, lane, stage, m);
    write_knn(knn_out(p, actor), k, knn_flags(p), m.lkey, lane);
    if (tl && lane == 0) { tl[5] = dbg_t0; tl[6] = static_cast<unsigned long long>(nparts * k); tl[7] = globaltimer_ns(); }
}

// TWO-LEVEL merge for static schedules whose actors are cut into MANY runs (ScanParams::cta_merge; chosen by the host
// when an actor has >= 96 runs: 1 x 30 000 has 235).  With few runs the single-level protocol above is shorter - the
// CTA-level combine is a serial phase of ~2 us in front of the same round trips (measured: 64 x 5 000, 32 runs per
// actor, 4.7 -> 10.7 us after the last tile; 1 x 30 000 9-11 -> 5.5 us).
// These are bound by latency, not
// bandwidth, so the merge is the shortest chain we know, in two levels:
//   1. inside the CTA, through shared memory.  The eight chunks of a CTA are consecutive in tile space, so its warps'
//      partial runs (a chunk's first and last run; the runs in between cover whole actors and are final) fall into a
//      few groups of one actor each.  Every warp leaves its <= 2 partial lists in shared memory and arrives on a CTA
//      counter; the LAST warp to arrive combines each group - the lists are sorted, so they are offered rank-major,
//      32 / L ranks of the L lists per batch: one bitonic batch merge and a few inserts, ~1 us;
//   2. between the CTAs of an actor, through global memory as before: the CTA publishes ONE list per actor at the
//      slot that follows from its block index (k epoch-tagged entries, no reservation round trip), counts its tiles
//      on the actor's counter, and the CTA whose arrival completes the actor merges the <= tiles_per_actor /
//      (WARPS * s1) + 2 lists - exact knowledge, no polling for arrivals, only for the entries' tags.
// 1 x 30 000: 235 per-warp lists merged by one warp (8-11 us after the last tile) became 30 CTA lists.
// A group that covers its whole actor inside one CTA is written directly.  All atomics of a CTA's groups are issued
// before any of their results is looked at (lane g keeps the result of group g).
template <int WARPS>
__device__ __forceinline__ void static_cta_finish(const ScanParams& p, uint32_t cta_area, int warp, int lane, uint32_t tag,
                                                  int n_active, int e0_actor, int e0_tiles, u64 e0_key, int e1_actor,
                                                  int e1_tiles, u64 e1_key, uint32_t stage, unsigned long long* tl) {
    const int k = p.k, tpa = p.tiles_per_actor;
    const uint32_t lists = cta_area;                           // [WARPS][2][32] u64
    const uint32_t meta = cta_area + WARPS * 2 * 32 * 8;       // [WARPS][2] {actor, tiles}
    const uint32_t ctr = meta + WARPS * 2 * 8;                 // arrivals (zeroed at kernel start)
    // ---- level 1: leave this warp's partial lists in shared memory, arrive ----
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(lists + ((warp * 2 + 0) * 32 + lane) * 8), "l"(lane < k ? e0_key : 0ull) : "memory");
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(lists + ((warp * 2 + 1) * 32 + lane) * 8), "l"(lane < k ? e1_key : 0ull) : "memory");
    if (lane == 0) {
        asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(meta + (warp * 2 + 0) * 8), "r"(e0_actor), "r"(e0_tiles) : "memory");
        asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(meta + (warp * 2 + 1) * 8), "r"(e1_actor), "r"(e1_tiles) : "memory");
    }
    __syncwarp();
    unsigned arrived = 0;
    if (lane == 0) {
        __threadfence_block();
        asm volatile("atom.shared.acq_rel.cta.add.u32 %0, [%1], 1;" : "=r"(arrived) : "r"(ctr) : "memory");
    }
    arrived = __shfl_sync(kFull, arrived, 0);
    if (static_cast<int>(arrived) != n_active - 1) {           // not the last warp of the CTA
        if (tl && lane == 0) tl[5] = 0ull;
        return;
    }
    __threadfence_block();
    // ---- the CTA's combiner: groups of consecutive entries with the same actor ----
    const int E = 2 * n_active;                                // <= 32 entries, one per lane
    int my_actor = -1, my_tiles = 0;
    if (lane < E) asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(my_actor), "=r"(my_tiles) : "r"(meta + lane * 8) : "memory");
    const int cs = WARPS * p.sched.s1;                         // tiles per CTA
    int g_actor = -1, g_tiles = 0;                             // lane g: group g (actor, tiles; tiles = 0: written directly)
    u64 g_ret = 0ull;                                          // lane g: UNREAD result of group g's arrive atomic
    int n_groups = 0;
    unsigned todo = __ballot_sync(kFull, my_actor >= 0);
    const unsigned long long dbg_t0 = tl ? globaltimer_ns() : 0ull;
#pragma unroll 1
    while (todo) {
        const int first = __ffs(todo) - 1;
        const int actor = __shfl_sync(kFull, my_actor, first);
        const bool mine = my_actor == actor;                   // entries of one actor are consecutive, but need not be
        const unsigned members = __ballot_sync(kFull, mine);   // (an absent first run leaves a hole): use the mask
        todo &= ~members;
        int tiles = mine ? my_tiles : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tiles += __shfl_xor_sync(kFull, tiles, o);
        const int L = __popc(members);
        // lane -> (list, rank): list = the (lane % L)-th member, ranks [b * R, (b + 1) * R) in batch b
        const int R = 32 / L;
        int nth = lane % L, entry = 0;
        {
            unsigned mm = members;
            for (int i = 0; i < nth; ++i) mm &= mm - 1;
            entry = __ffs(mm) - 1;
        }
        WarpTopK m;
        m.reset();
#pragma unroll 1
        for (int r0 = 0; r0 < k; r0 += R) {
            const int r = r0 + lane / L;
            u64 key = 0ull;
            if (lane / L < R && r < k) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(key) : "r"(lists + (entry * 32 + r) * 8) : "memory");
            const unsigned passed = __ballot_sync(kFull, key > m.thr);
            if (!passed) break;                                // sorted lists: no deeper rank can pass either
            m.insert(key, passed, lane, k);
        }
        if (tiles == tpa) {                                    // the CTA covered the whole actor
            write_knn(knn_out(p, actor), k, knn_flags(p), m.lkey, lane);
        } else {
            const int t0 = actor * tpa;
            const int slot = static_cast<int>(blockIdx.x) - t0 / cs;
            u64* cand = p.cand + static_cast<int64_t>(actor) * p.max_cand * 2;
            if (slot < 0 || (slot + 1) * k > p.max_cand) {     // cannot happen; never write out of bounds
                if (lane == 0) atomicAdd(p.timeouts, 1ull);
            } else {
                if (lane < k) st_ll(cand + (slot * k + lane) * 2, m.lkey, tag);
                if (lane == n_groups) {
                    g_actor = actor; g_tiles = tiles;
                    g_ret = atomicAdd(actor_ctr(p, actor), static_cast<u64>(tiles) << 32);
                }
                ++n_groups;
            }
        }
    }
    // ---- level 2: the groups whose arrival completed their actor ----
    if (tl && lane == 0) { tl[5] = dbg_t0; tl[6] = globaltimer_ns(); tl[7] = tl[6]; }   // combiner: start, level 1 done, level 2 done
#pragma unroll 1
    for (int g = 0; g < n_groups; ++g) {
        const int actor = __shfl_sync(kFull, g_actor, g);
        const int tiles = __shfl_sync(kFull, g_tiles, g);
        const u64 prev = __shfl_sync(kFull, g_ret, g);
        if (static_cast<uint32_t>(prev >> 32) + static_cast<uint32_t>(tiles) != static_cast<uint32_t>(tpa)) continue;
        if (lane == 0) *actor_ctr(p, actor) = 0ull;            // re-arm for the next launch
        if (p.debug & 8) continue;
        const int t0 = actor * tpa;
        const int nparts = (t0 + tpa - 1) / cs - t0 / cs + 1;  // CTAs that overlap the actor
        const u64* cand = p.cand + static_cast<int64_t>(actor) * p.max_cand * 2;
        WarpTopK m;
        m.reset();
        consume_lists(p, cand, nparts, tag, lane, stage, m);
        write_knn(knn_out(p, actor), k, knn_flags(p), m.lkey, lane);
        if (tl && lane == 0) tl[7] = globaltimer_ns();
    }
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
// Experiment knob: working sets that (partly) fit the 126 MB L2 - a strong-scaling shard of 32 actors x 30k is 123 MB -
// could stay resident across steps: evict_normal, or a FIXED fraction of the lines evict_last (picked by address, so
// the same lines every step).  Measured: no gain (tools/l2_sweep.py; DESIGN.md, "Measured and rejected").
__device__ __forceinline__ uint64_t l2_stream_policy(int mode, float frac) {
    uint64_t pol;
    if (mode == 1) {
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    } else if (mode == 2) {
        asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(frac));
    } else {
        pol = l2_evict_first_policy();
    }
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
// plain (1-D) bulk copy global -> shared, completing on an mbarrier; 16-byte granularity
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// TMA prefetch of one tile into L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}

// ---- packed fp32 arithmetic (sm_100: FADD2 / FFMA2 work on two fp32 values in a 64-bit register pair,
// each half IEEE round-to-nearest) -----------------------------------------------------------------
// The distance of one row is 32 subtractions, 32 squares and 31 additions in a FIXED order with every
// result rounded separately (that is what makes d^2 bit-identical to torch-CPU); done as scalar FADD /
// FMUL they are 95 of the ~270 instructions a warp issues per tile, and the warps are busy enough on
// that path for it to show in the kernel time.  Packed, they are 16 + 16 + 12 (+ 7 scalar adds).
// The square is written fma(d, d, z) with z = (-0.0, -0.0) passed as a KERNEL PARAMETER:
//   * ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (single rounding -> different bits) even
//     with --fmad=false, and it folds a literal -0.0 addend back into a multiply first;
//   * an addend it cannot see through keeps the product a separately rounded FFMA2 result, and
//     x*x + (-0.0) == x*x for every x (a square is never -0.0), so the bits are those of mul.rn.
using f2 = unsigned long long;
constexpr f2 kNegZero2 = 0x8000000080000000ull;
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_sq(f2 d, f2 z) { f2 r; asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(d), "l"(z)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// d^2 of one swizzled 128-byte row against the register-resident query (16 packed pairs).
// row_addr = shared address of the row start, sw = row & 7, z = kNegZero2 from the parameters.
// Rows are 128-byte aligned (tile rows; the q_prev rows of the aux area), so the chunk offset (c ^ sw) << 4 can be
// XORed into the address: one LOP3 with an immediate per chunk from the base row_addr ^ (sw << 4), instead of a LOP3 and
// an add per chunk.
__device__ __forceinline__ float row_d2(uint32_t row_addr, uint32_t sw, const f2 (&q)[16], f2 z) {
    f2 a[4];   // lane accumulators (a0,a1) (a2,a3) (a4,a5) (a6,a7)
    const uint32_t base = row_addr ^ (sw << 4);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        f2 m01, m23;
        lds_2x64(base ^ (static_cast<uint32_t>(c) << 4), m01, m23);
        const f2 s01 = f2_sq(f2_sub(m01, q[2 * c]), z);
        const f2 s23 = f2_sq(f2_sub(m23, q[2 * c + 1]), z);
        const int i = 2 * (c & 1);   // elements 4c..4c+3 feed lane accumulators 4(c&1)..4(c&1)+3
        if (c < 2) {  // first contribution to each lane accumulator: 0 + s == s
            a[i] = s01; a[i + 1] = s23;
        } else {
            a[i] = f2_add(a[i], s01);
            a[i + 1] = f2_add(a[i + 1], s23);
        }
    }
    float lo, hi;
    f2_unpack(a[0], lo, hi);
    float h = __fadd_rn(lo, hi);
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        f2_unpack(a[i], lo, hi);
        h = __fadd_rn(__fadd_rn(h, lo), hi);
    }
    return h;
}

// Rare side loads of the producer lane (see produce()): query row + floor of a starting run, q_prev row of the slot
// being appended by the previous step.  Out of line on purpose.
__device__ __noinline__ void run_aux_loads(uint32_t aux, uint32_t qprev_row, const float* q, const float* q_prev, const u64* floor_word,
                                           int actor, uint32_t bar, bool run_start, bool hz) {
    if (run_start) {
        bulk_load(aux, q + static_cast<int64_t>(actor) * 32, 128, bar);
        bulk_load(aux + 128, floor_word, 16, bar);
    }
    if (hz) bulk_load(qprev_row, q_prev + static_cast<int64_t>(actor) * 32, 128, bar);
}

template <int WARPS, int STAGES, int RPL>
struct ScanCfg {
    static constexpr int kWarps = WARPS;
    static constexpr int kStages = STAGES;
    static constexpr int kTileRows = 32 * RPL;
    static constexpr int kTileBytes = kTileRows * 128;
    static constexpr int kThreads = WARPS * 32;
    // per warp: STAGES rows for the slot the previous step is appending (pipelined launches; 128-byte aligned like tile
    // rows: row_d2), then per stage the query row + {floor, ctr} of the actor whose run starts there
    static constexpr int kRunAuxBytes = 128 + 16;
    static constexpr int kWarpAuxBytes = (STAGES * (128 + kRunAuxBytes) + 127) / 128 * 128;
    static_assert((WARPS * STAGES * 16) % 128 == 0, "barriers + descriptors must keep the aux blocks 128-byte aligned");
    // tiles + 1024 B alignment slack + barriers + per-stage descriptors + per-stage run-start data
    static constexpr int kCtaMergeBytes = WARPS * (2 * 32 * 8 + 2 * 8) + 16;   // static schedules: the CTA-level combine
    static constexpr int kSmemBytes = WARPS * STAGES * kTileBytes + 1024 + WARPS * STAGES * 16 + WARPS * kWarpAuxBytes + kCtaMergeBytes;
    // CTAs per SM the shared memory allows (227 KB per SM): tells ptxas how many registers it may use.
    // Without it ptxas squeezes the kernel into 80 registers and spills producer state in the hot loop.
    static constexpr int kMaxCtasPerSm = (227 * 1024) / (kSmemBytes + 1024) > 0 ? (227 * 1024) / (kSmemBytes + 1024) : 1;
};

// DBG = false: the production flavour.  The experiment switches (ScanParams::debug, NGU_EPI_DEBUG_SCAN) and the per-warp
// timeline stamps are compiled out of the streaming loop - tested every tile they were ~20 of its ~255 instructions.
// The host launches the DBG = true flavour whenever one of them is set (ngu_epi_create).
template <int WARPS, int STAGES, int RPL, bool DBG>
__global__ void __launch_bounds__(WARPS * 32, ScanCfg<WARPS, STAGES, RPL>::kMaxCtasPerSm) scan_topk_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                const ScanParams p) {
    using Cfg = ScanCfg<WARPS, STAGES, RPL>;
    const int dbg = DBG ? p.debug : 0;
    unsigned long long* const timeline = DBG ? p.timeline : nullptr;
    extern __shared__ uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // 128B-swizzle atoms are 1024 B: align the tile area.
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles0 = (raw + 1023u) & ~1023u;
    const uint32_t bars0 = tiles0 + WARPS * STAGES * Cfg::kTileBytes;
    const uint32_t desc0 = bars0 + WARPS * STAGES * 8;
    const uint32_t my_tiles = tiles0 + warp * STAGES * Cfg::kTileBytes;
    const uint32_t my_bars = bars0 + warp * STAGES * 8;
    const uint32_t my_desc = desc0 + warp * STAGES * 8;   // per stage {actor, tile in actor}; actor < 0: no more work
    const uint32_t aux0 = desc0 + WARPS * STAGES * 8;     // 128-byte aligned (static_assert in ScanCfg)
    const uint32_t my_qprev = aux0 + warp * Cfg::kWarpAuxBytes;      // this warp's block: q_prev rows, one per stage,
    const uint32_t my_aux = my_qprev + STAGES * 128;                 // then query row + floor per stage

    const uint32_t cta_area = aux0 + WARPS * Cfg::kWarpAuxBytes;   // static schedules: lists, meta, arrival counter

    const int gw = blockIdx.x * WARPS + warp;
    const int n_warps = gridDim.x * WARPS;
    const unsigned long long tl0 = timeline ? globaltimer_ns() : 0ull;
    unsigned long long tl1 = 0ull;
    const ScanSched& sc = p.sched;
    if (threadIdx.x == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(cta_area + WARPS * (2 * 32 * 8 + 2 * 8)), "r"(0u) : "memory");
    __syncthreads();                // the only CTA-wide barrier: the arrival counter of static_cta_finish is zero
    if (gw >= sc.n_chunks) return;  // whole warp exits together; no CTA-wide barriers below

    const int tpa = p.tiles_per_actor;
    const int N = p.capacity;
    int pt, pte;                       // current chunk: next tile to load, end
    chunk_range(sc, gw, pt, pte);      // first chunk is static: no atomic before the first load
    int pb = pt / tpa;                 // actor / tile-in-actor of `pt`
    int pti = pt - pb * tpa;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(my_bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // While the predecessor (the previous step's epilogue) is still running, pull this warp's first tiles
        // into L2: the CTA is resident but may not touch the memory's CONTENT before the grid dependency
        // resolves (the epilogue is appending a row) - a prefetch does not, and L2 being the coherence point
        // the real loads below still see that row.  ~28 MB of HBM traffic leave the step's critical path.
        if (!(dbg & 32)) {
            int fb = pb, fti = pti;
#pragma unroll 1
            for (int s = 0; s < p.prefetch_tiles; ++s) {
                if (pt + s < pte) {
                    tma_prefetch_2d(&tmap, 0, fb * N + fti * Cfg::kTileRows);
                    if (++fti == tpa) { fti = 0; ++fb; }
                }
            }
        }
    }
    __syncwarp();
    // ---- software pipeline over steps (p.pipe, ngu_epi_step_pipelined) ------------------------------------------
    // Step s+1's scan depends on step s only through ONE row per actor: the slot c(s) that epilogue s appends q(s)
    // to.  A pipelined launch therefore does not wait for its predecessor at all: it streams while the previous
    // scan's merge tail, the previous epilogue and (multi-GPU) its NVLink exchange are still running, and takes the
    // content of row c(s) from q(s) itself (q_prev: an extra 128-byte bulk copy that rides on the tile's mbarrier and
    // replaces that row of the tile, whatever state the append is in).  c(s) comes from a word every scan publishes
    // before it lets its dependents start (scan_cursor = seq << 32 | slot): an ordinary launch reads the exact device
    // cursor after its grid dependency has resolved, a pipelined one advances the previous launch's value.
    // Two steps deep, no more: this launch works on the scan context (candidates, counters, k-NN outputs) that step
    // s-1 used, so it first waits for epilogue s-1 (wait_epi, normally long complete); all of its predecessors' CTAs
    // are resident by construction (a kernel only becomes resident after ALL CTAs of its primary have signalled), so
    // nothing it polls can be waiting for an SM.  Before a warp exits it executes griddepcontrol.wait, so that the
    // completion of this grid still implies the completion of everything before it (the epilogues stay ordered).
    int hz_tile = -1, hz_row = 0;      // tile-in-actor / row-in-tile of the slot being appended (none: -1)
    bool chained = false;
    if (p.pipe) {
        u64 w = 0ull;
        if (lane == 0) {
            // both words in one round trip: the cursor word (valid whatever the epilogues are doing) is requested first,
            // the acquire load behind it
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p.scan_cursor) : "memory");
            if (p.wait_epi != 0u) {
                unsigned int v;
                int polls = 0;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.epi_done) : "memory");
                    if (static_cast<int>(v - p.wait_epi) >= 0) break;
                    if (++polls > kLLSpinLimit) { atomicAdd(p.timeouts, 1ull); break; }
                }
            }
        }
        w = __shfl_sync(kFull, w, 0);
        const uint32_t ws = static_cast<uint32_t>(w >> 32);
        const int c = static_cast<int>(static_cast<uint32_t>(w));
        int row = -1;
        if (ws == p.prev_seq) row = c;                          // the previous step's slot
        else if (ws == p.seq) row = c == 0 ? N - 1 : c - 1;     // CTA 0 of this launch has already advanced the word
        if (row >= 0 && row < N) {
            chained = true;
            hz_tile = row / Cfg::kTileRows;
            hz_row = row - hz_tile * Cfg::kTileRows;
            if (gw == 0 && lane == 0 && ws == p.prev_seq) {
                const u64 nw = (static_cast<u64>(p.seq) << 32) | static_cast<u64>(row + 1 == N ? 0 : row + 1);
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p.scan_cursor), "l"(nw) : "memory");
                __threadfence();
            }
        }
    }
    if (!chained) {
        pdl_wait();   // predecessor (previous step's epilogue: append, readers of dist/topk) is complete
        if (p.seq != 0u && gw == 0 && lane == 0) {              // head of a chain: the exact cursor
            const u64 nw = (static_cast<u64>(p.seq) << 32) | static_cast<u64>(__ldcg(p.cursor));
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p.scan_cursor), "l"(nw) : "memory");
            __threadfence();
        }
    }
    // A pre-armed step (ngu_epi_step_host, armed_stage_kernel) whose doorbell never rang: leave without touching
    // anything (the exit releases the epilogue behind us, which does the same).
    // (Measured and rejected: polling this word for "go" instead of waiting for the stage kernel to exit.  With
    // 2 368 resident warps polling one line the release store queues behind them: +3 us at 256 x 30k, -1 us at 1 x 30k.)
    if (p.arm_seq != 0u && __ldcg(p.arm_go) == ((p.arm_seq << 1) | 1u)) return;
    // Only now may this step's epilogue become resident (it waits on us): signalled AFTER the wait, so that an
    // epilogue CTA never runs concurrently with the previous epilogue - its pre-wait section reads state the
    // previous one writes (epilogue.cuh, log_absorb).  Costs nothing: it could not become resident before one of
    // our CTAs exits (shared memory) or, on small grids, has tens of microseconds of streaming left to do so.
    pdl_launch_dependents();

    const uint64_t policy = l2_stream_policy(p.l2_mode, p.l2_frac);
    const uint32_t tag = __ldcg(p.sched_ctr + 2);   // launch epoch (first needed at the first flush)

    // ---- producer (lane 0): walks this warp's chunks STAGES tiles ahead of the consumer ----
    int pending = sc.n_chunks;         // next chunk index, fetched one chunk ahead
    int last_pb = -1;                  // actor of the tile issued last: a change = the consumer starts a new run there
    bool live = true;
    if (lane == 0 && n_warps < sc.n_chunks) pending = n_warps + static_cast<int>(atomicAdd(p.sched_ctr, 1u));
    auto produce = [&](int s) {        // lane 0 only
        while (live && pt == pte) {
            if (pending < sc.n_chunks) {
                chunk_range(sc, pending, pt, pte);
                pb = pt / tpa;
                pti = pt - pb * tpa;
                pending = n_warps + static_cast<int>(atomicAdd(p.sched_ctr, 1u));
            } else {
                live = false;
            }
        }
        const uint32_t dsc = my_desc + 8 * s;
        if (live) {
            asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(dsc), "r"(pb), "r"(pti) : "memory");
            const uint32_t bar = my_bars + 8 * s;
            const bool run_start = pb != last_pb;
            const bool hz = pti == hz_tile;   // pipelined launch: the tile holds the slot the previous step is appending to
            mbar_expect_tx(bar, Cfg::kTileBytes + (run_start ? 144 : 0) + (hz ? 128 : 0));
            tma_load_2d(my_tiles + s * Cfg::kTileBytes, &tmap, 0, pb * N + pti * Cfg::kTileRows, bar, policy);
            // the run's query row and floor travel with its first tile, and so does the final content of the slot the
            // previous step is appending to.  Both are rare (once per run / per actor) and sit behind a REAL branch:
            // predicated, each bulk copy costs its address arithmetic and uniform-register hand-over on every tile
            // (36 of the producer's 87 instructions per tile, ncu source page of r02f).
            if (run_start || hz) {
                run_aux_loads(my_aux + s * Cfg::kRunAuxBytes, my_qprev + s * 128, p.q, p.q_prev, actor_floor(p, pb), pb, bar, run_start, hz);
                last_pb = pb;
            }
            ++pt;
            if (++pti == tpa) { pti = 0; ++pb; }
        } else {
            asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(dsc), "r"(-1), "r"(0) : "memory");
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) produce(s);
    }
    __syncwarp();

    const bool largest = p.largest != 0;
    const int k = p.k;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    f2 q[16];
    WarpTopK top;
    int cur = -1;        // actor whose list is being built
    int run_tiles = 0;   // tiles of `cur` in the list
    bool run_start = false;     // the tile being consumed is the first of a run
    int pend_actor = -1, pend_n = 0;   // flushed run whose reserve/arrive atomic has not been read yet
    u64 pend_key = 0ull, pend_ret = 0ull;
    // static schedules: the warp's chunk can begin inside one actor and end inside another; the first of
    // these two partial runs waits here until the warp has finished streaming (pend_n = its tile count)
    const bool static_sched = sc.n_chunks == sc.n1;
    int run_first = 0;                 // tile-in-actor of the current run's first tile
    int stage = 0;
    uint32_t parity = 0;
    bool first = true;

    for (;;) {
        int b, ti;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(b), "=r"(ti) : "r"(my_desc + 8 * stage) : "memory");
        if (b < 0) break;
        if (b != cur) {
            if (cur >= 0 && !(dbg & 4)) {
                if (static_sched && p.cta_merge) {
                    if (run_tiles == tpa) {            // this warp scanned the whole actor: final
                        write_knn(knn_out(p, cur), k, knn_flags(p), top.lkey, lane);
                    } else {                           // the chunk's first run: kept in registers until the warp has streamed
                        if (pend_actor >= 0 && lane == 0) atomicAdd(p.timeouts, 1ull);   // (cannot happen: only a chunk's
                        pend_actor = cur;                                                 // first and last run are partial)
                        pend_n = run_tiles;
                        pend_key = top.lkey;
                    }
                } else if (static_sched) {
                    if (pend_actor >= 0)   // (cannot happen: only a chunk's first and last run are partial)
                        finish_static(p, pend_actor, pend_n, pend_ret, tag, lane, my_tiles, nullptr);
                    pend_actor = flush_run_static(p, cur, run_tiles, run_first, tag, top, lane, pend_ret) ? cur : -1;
                    pend_n = run_tiles;
                } else {
                    if (pend_actor >= 0) complete_pending(p, pend_actor, pend_key, pend_n, pend_ret, tag, lane);
                    pend_actor = flush_run(p, cur, run_tiles, tag, top, lane, pend_key, pend_n, pend_ret) ? cur : -1;
                }
            }
            cur = b;
            run_tiles = 0;
            run_first = ti;
            run_start = true;
        }

        mbar_wait(my_bars + 8 * stage, parity);
        if (timeline && first) { tl1 = globaltimer_ns(); first = false; }
        if (run_start) {   // query row and floor arrived with the tile (same mbarrier)
            const uint32_t aux = my_aux + stage * Cfg::kRunAuxBytes;
#pragma unroll
            for (int c = 0; c < 8; ++c) lds_2x64(aux + 16 * c, q[2 * c], q[2 * c + 1]);
            u64 fw;
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(fw) : "r"(aux + 128) : "memory");
            // keys at or below the actor's published floor cannot make the final top-k
            top.reset((dbg & 16) ? 0ull : floor_key(fw, tag));
            run_start = false;
        }
        const uint32_t tile = my_tiles + stage * Cfg::kTileBytes;
        const int row0 = ti * Cfg::kTileRows;
        float d2[RPL];
        if (dbg & 2) {
#pragma unroll
            for (int i = 0; i < RPL; ++i) d2[i] = lds128(tile + static_cast<uint32_t>(lane + 32 * i) * 128u + (sw << 4)).x;
        } else {
            const bool hz = ti == hz_tile;
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                uint32_t ra = tile + static_cast<uint32_t>(lane + 32 * i) * 128u, rs = sw;
                if (hz && lane + 32 * i == hz_row) {   // the slot being appended: its content from q_prev (linear, not swizzled)
                    ra = my_qprev + stage * 128;
                    rs = 0u;
                }
                d2[i] = row_d2(ra, rs, q, p.neg_zero2);
            }
        }
        __syncwarp();  // every lane is done reading this slot (and has read its descriptor)

        if (lane == 0) produce(stage);   // refill the slot with this warp's tile STAGES ahead
        __syncwarp();  // publishes the new descriptor

#pragma unroll
        for (int i = 0; i < RPL; ++i) {
            const int slot = row0 + lane + 32 * i;
            const u64 key = slot < N ? make_key(d2[i], static_cast<uint32_t>(slot), largest) : 0ull;
            if (dbg & 1) top.lkey ^= key; else
            top.offer(key, lane, k);
        }
        ++run_tiles;
        if (++stage == STAGES) { stage = 0; parity ^= 1u; }
        if (pend_actor >= 0 && !static_sched) {   // the previous run's atomic has had a whole tile to come back
            complete_pending(p, pend_actor, pend_key, pend_n, pend_ret, tag, lane);
            pend_actor = -1;
        }
    }

    const unsigned long long tl2 = timeline ? globaltimer_ns() : 0ull;
    const int workers = n_warps < sc.n_chunks ? n_warps : sc.n_chunks;
    unsigned long long* tl_rec = timeline ? timeline + static_cast<int64_t>(gw) * 8 : nullptr;
    if (cur >= 0 && !(dbg & 4) && static_sched && !p.cta_merge) {
        u64 ret = 0ull;
        const bool last_partial = flush_run_static(p, cur, run_tiles, run_first, tag, top, lane, ret);
        if (pend_actor >= 0) finish_static(p, pend_actor, pend_n, pend_ret, tag, lane, my_tiles, tl_rec);
        if (last_partial) finish_static(p, cur, run_tiles, ret, tag, lane, my_tiles, tl_rec);
    } else if (cur >= 0 && !(dbg & 4) && static_sched) {
        int last_actor = -1;
        if (run_tiles == tpa) write_knn(knn_out(p, cur), k, knn_flags(p), top.lkey, lane);
        else last_actor = cur;
        const int rest = sc.n_chunks - static_cast<int>(blockIdx.x) * WARPS;
        static_cta_finish<WARPS>(p, cta_area, warp, lane, tag, rest < WARPS ? rest : WARPS, pend_actor, pend_n, pend_key,
                                 last_actor, run_tiles, top.lkey, my_tiles, tl_rec);
    } else if (cur >= 0 && !(dbg & 4)) {
        if (pend_actor >= 0) complete_pending(p, pend_actor, pend_key, pend_n, pend_ret, tag, lane);
        if (flush_run(p, cur, run_tiles, tag, top, lane, pend_key, pend_n, pend_ret))
            complete_pending(p, cur, pend_key, pend_n, pend_ret, tag, lane);
        for (;;) {   // merge jobs: one actor each, first come first served
            int job = 0;
            if (lane == 0) job = static_cast<int>(atomicAdd(p.sched_ctr + 3, 1u));
            job = __shfl_sync(kFull, job, 0);
            if (job >= p.n_actors) break;
            merge_job(p, job, tag, lane, my_tiles, true, tl_rec);
        }
    }
    else if (top.lkey == 0x1234567ull) p.topk_d2[0] = 0.f;   // keep the experiment's loads alive
    if (chained) pdl_wait();   // pipelined launch: this grid must not complete before its predecessors have
    // The last working warp of the launch merges the jobs that were given up (normally none) and re-arms the
    // scheduler for the next launch (every other warp has consumed the result of its last counter atomic
    // before it got here).
    unsigned last = 0;
    if (lane == 0) last = atomicAdd(p.sched_ctr + 1, 1u) == static_cast<unsigned>(workers) - 1u ? 1u : 0u;
    last = __shfl_sync(kFull, last, 0);
    if (last) {
        __threadfence();
        const int n_left = static_cast<int>(__ldcg(p.sched_ctr + 4));
        for (int i = 0; i < n_left; ++i) merge_job(p, __ldcg(p.leftover + i), tag, lane, my_tiles, false, nullptr);
        if (lane == 0) {
            p.sched_ctr[0] = 0u; p.sched_ctr[1] = 0u; p.sched_ctr[3] = 0u; p.sched_ctr[4] = 0u;
            uint32_t next = tag + 1u;                      // next launch's epoch
            if (next == 0u) {                              // wrapped: 0 is never a valid tag, and floors of the old
                next = 1u;                                 // cycle would out-rank every new one in the atomicMax
                for (int a = 0; a < p.n_actors; ++a) *actor_floor(p, a) = 0ull;
            }
            p.sched_ctr[2] = next;
        }
    }
    if (timeline && lane == 0) {
        unsigned long long* o = timeline + static_cast<int64_t>(gw) * 8;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        o[0] = tl0; o[1] = tl1; o[2] = tl2; o[3] = globaltimer_ns(); o[4] = smid;
    }
}

}  // namespace ngu
