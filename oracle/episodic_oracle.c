/*
 * episodic_oracle.c -- plain-C restatement of the reference's distance scan + top-k
 * (TEST INFRASTRUCTURE ONLY: the checker for the CUDA path at full size and the
 * "port" CPU baseline timed by bench.py; never linked into the product).
 *
 * Follows ngu/models/intrinsic_novelty/episodic_novelty/episodic_novelty.py:51-54:
 *     euc_dist = ((episodic_memory - controllable_state.unsqueeze(0))**2).sum(dim=-1).sqrt()
 *     k_neighbor = euc_dist.topk(k, dim=0)
 * with torch-CPU's exact float32 arithmetic for the D=32 reduction:
 * separately rounded diff and square (no FMA; compile with -ffp-contract=off),
 * eight lane accumulators a[l] = ((s[l]+s[l+8])+s[l+16])+s[l+24], then
 * ((((((a0+a1)+a2)+a3)+a4)+a5)+a6)+a7.  Selection is on d^2 with the canonical
 * order (d^2 descending -- or ascending in paper mode --, slot ascending); sqrt is
 * applied by the caller to the k winners only.
 *
 * Parity pinning: the reference has no golden vectors; this file is pinned
 * against tests/golden/*.npz (outputs of the unmodified reference) and against
 * the numpy restatement in tests/test_oracle.py.
 *
 * Memory layout is the reference's: slot-major float[N][B][32].
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned long long u64;

static inline float d2_lane8(const float* m, const float* q) {
    float a[8];
    for (int l = 0; l < 8; ++l) { float d = m[l] - q[l]; a[l] = d * d; }
    for (int p = 1; p < 4; ++p)
        for (int l = 0; l < 8; ++l) { float d = m[8 * p + l] - q[8 * p + l]; a[l] = a[l] + d * d; }
    float h = a[0];
    for (int l = 1; l < 8; ++l) h = h + a[l];
    return h;
}

static inline u64 make_key(float d2, uint32_t slot, int largest) {
    uint32_t bits;
    memcpy(&bits, &d2, 4);
    if (!largest) bits = ~bits;
    return ((u64)bits << 32) | (u64)(0xFFFFFFFFu - slot);
}

/* insert into a descending list of k keys if it beats the last one */
static inline void offer(u64* list, int k, u64 key) {
    if (key <= list[k - 1]) return;
    int j = k - 1;
    while (j > 0 && list[j - 1] < key) { list[j] = list[j - 1]; --j; }
    list[j] = key;
}

/* mem [N][B][32], q [B][32]  ->  d2_out [k][B] (best first), idx_out [k][B].
 * nthreads <= 0: use all OpenMP threads.  Returns 0, or -1 on bad arguments. */
int oracle_scan_topk(const float* mem, const float* q, int B, int N, int k, int largest,
                     float* d2_out, int32_t* idx_out, int nthreads) {
    if (k < 1 || k > N || B < 1) return -1;
    int T = 1;
#ifdef _OPENMP
    T = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
    if (T > N) T = N;
    u64* lists = (u64*)calloc((size_t)T * B * k, sizeof(u64));   /* 0 = sentinel below every key */
    if (!lists) return -1;
#ifdef _OPENMP
#pragma omp parallel num_threads(T)
#endif
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        const int n0 = (int)((long long)N * t / T), n1 = (int)((long long)N * (t + 1) / T);
        u64* mine = lists + (size_t)t * B * k;
        for (int n = n0; n < n1; ++n) {
            const float* row = mem + (size_t)n * B * 32;
            for (int b = 0; b < B; ++b)
                offer(mine + (size_t)b * k, k, make_key(d2_lane8(row + b * 32, q + b * 32), (uint32_t)n, largest));
        }
    }
    for (int b = 0; b < B; ++b) {
        u64* dst = lists + (size_t)b * k;             /* thread 0's list collects the merge */
        for (int t = 1; t < T; ++t) {
            const u64* src = lists + ((size_t)t * B + b) * k;
            for (int j = 0; j < k; ++j) offer(dst, k, src[j]);
        }
        for (int j = 0; j < k; ++j) {
            uint32_t bits = (uint32_t)(dst[j] >> 32);
            if (!largest) bits = ~bits;
            float d2;
            memcpy(&d2, &bits, 4);
            d2_out[(size_t)j * B + b] = d2;
            idx_out[(size_t)j * B + b] = (int32_t)(0xFFFFFFFFu - (uint32_t)dst[j]);
        }
    }
    free(lists);
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
