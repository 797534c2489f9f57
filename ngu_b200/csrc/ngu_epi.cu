// ngu_epi.cu — C ABI (include/ngu_epi.h) over the sm_100a kernels.
//
// Host side of the B200 episodic-novelty engine: handle/state management, the
// TMA tensor map over the episodic memory, launch geometry, and the per-step
// kernel sequence  scan_topk (with in-kernel merge) -> epilogue (rank sums, [exchange],
// running-mean update, rewards, append), chained with programmatic dependent launch.
// No torch headers, no CPU fallback: every entry point needs a CUDA device.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/ngu_epi.h"
#include "epilogue.cuh"
#include "scan_topk.cuh"

using namespace ngu;

namespace {

thread_local std::string g_create_error;

typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                             const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::string fmt(const char* f, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

// Scan kernel variants (WARPS, STAGES, rows-per-lane); see DESIGN.md "kernel A".
struct Variant {
    int warps, stages, rpl, ctas_per_sm;
    const void* fn;       // launched
    const void* fn_dbg;   // flavour with the experiment switches / timeline stamps compiled in (NGU_EPI_DEBUG_SCAN, _TIMELINE)
    int smem;
};
// LEAN: also instantiate the production flavour (no debug code in the streaming loop).  Only the default geometry has
// one; the experiment geometries always run their debug-capable build.
template <int W, int S, int R, bool LEAN = false>
Variant make_variant(int ctas_per_sm) {
    return Variant{W, S, R, ctas_per_sm, reinterpret_cast<const void*>(&scan_topk_kernel<W, S, R, !LEAN>),
                   reinterpret_cast<const void*>(&scan_topk_kernel<W, S, R, true>), ScanCfg<W, S, R>::kSmemBytes};
}
const Variant kVariants[] = {
    make_variant<8, 3, 1, true>(2),   // 0 (default): 16 warps/SM, 3 x 4 KB ring per warp  = 192 KB in flight / SM
    make_variant<8, 6, 1>(1),   // 1:  8 warps/SM, 6 x 4 KB                         = 192 KB
    make_variant<8, 3, 2>(1),   // 2:  8 warps/SM, 3 x 8 KB                         = 192 KB
    make_variant<16, 3, 1>(1),  // 3: 16 warps/SM in one CTA, 3 x 4 KB              = 192 KB
    make_variant<8, 2, 1>(3),   // 4: 24 warps/SM, 2 x 4 KB                         = 192 KB
    make_variant<4, 4, 1>(3),   // 5: 12 warps/SM, 4 x 4 KB                         = 192 KB
    make_variant<8, 4, 1>(1),   // 6:  8 warps/SM, 4 x 4 KB                         = 128 KB
    make_variant<4, 6, 2>(1),   // 7:  4 warps/SM, 6 x 8 KB                         = 192 KB
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kEpilogueThreads = 512;
constexpr size_t kEpilogueStageBytes = 160 * 1024;   // (B,k) f32 staged in smem: B*k <= 40960 per handle

// Flatten numpy's pairwise-summation recursion for n elements into a postfix plan:
// (offset, length) = sum that leaf and push; (-1, 0) = pop two, push their f32 sum.
void build_pairwise_plan(int off, int n, std::vector<int2>& plan) {
    if (n <= 128) {
        plan.push_back(make_int2(off, n));
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    build_pairwise_plan(off, n2, plan);
    build_pairwise_plan(off + n2, n - n2, plan);
    plan.push_back(make_int2(-1, 0));
}
constexpr int kMinTilesPerWarp = 4;
constexpr int kDynMinAvgTilesDefault = 24;

}  // namespace

struct ngu_epi {
    ngu_epi_cfg cfg{};
    EpiConsts consts{};
    int sm_count = 0;
    bool owns_memory = false;
    float* memory = nullptr;         // [B][N][32]
    // Scan context: everything one scan launch writes and its epilogue reads.  There are two, used alternately, so
    // that a pipelined step's scan can run while the previous step's merge tail and epilogue are still using theirs.
    struct ScanCtx {
        u64* cand = nullptr;             // [B][max_cand][2] epoch-tagged candidate keys of the scan kernel
        float* topk_d2 = nullptr;        // [B][k]
        int32_t* topk_idx = nullptr;     // [B][k]
        float* dist = nullptr;           // [B][k]
        u64* actor_state = nullptr;      // [B][2] per actor {published floor, arrived tiles << 32 | reserved candidates}
        unsigned int* sched_ctr = nullptr;   // [8] dynamic chunk counter, finished-warp counter, launch epoch, merge-job counter, leftovers
        int32_t* leftover = nullptr;     // [B] merge jobs given up by their first taker (scan_topk.cuh: merge_job)
    } ctx[2];
    int cur = 0;                     // context of the scan launched last
    // fused-step chain (ngu_epi_step_pipelined)
    bool pipe_enabled = true;        // NGU_EPI_PIPE=0: every step is launched the ordinary way (A/B)
    bool chain_open = false;         // the last thing this handle enqueued was a fused step (eagerly, with a sequence number)
    cudaStream_t chain_stream = nullptr;
    unsigned int step_seq = 0;       // sequence number of the last eager fused step
    unsigned int chain_seq = 0, chain_prev_seq = 0;   // ... of the previous one / the one before it
    bool chain_prev_pipe = false;    // the previous fused step was launched pipelined
    const float* chain_q = nullptr;  // its controllable states (the row its epilogue appends)
    unsigned int* pipe_words = nullptr;   // device: [0] epi_done, [16] (64-byte offset) scan_cursor u64
    long pipe_launches = 0;
    double* rank_sums = nullptr;     // [2k+1]
    double* log_sums = nullptr;      // [5]
    float* log_stash = nullptr;      // [2][B] rewards of the last fused step, not yet in the log-only trackers (log_absorb)
    int prefetch_tiles = 4;                  // experiments: NGU_EPI_PREFETCH_TILES
    int cta_merge = 0;                       // static schedules: two-level merge (scan_topk.cuh: static_cta_finish); set by
                                             // plan_geometry when an actor is cut into many runs, NGU_EPI_CTA_MERGE=0|1 forces it
    int arm_prefetch_tiles = 4;              // ... of a pre-armed host step, whose CTAs sit behind the doorbell for the whole host
                                             // turnaround with HBM idle (NGU_EPI_ARM_PREFETCH_TILES)
    int l2_mode = 0; float l2_frac = 1.0f;   // L2 policy of the memory stream (ScanParams::l2_mode); NGU_EPI_L2=mode[,frac] overrides
    long long poll_limit_cycles = 4000000;   // ~2 ms without progress; NGU_EPI_DEBUG_POLL_LIMIT overrides (tests: 0)
    ScanSched sched{};
    double sched_static_frac = 0.6;          // share of the tiles handed out statically (experiments: NGU_EPI_SCHED=frac,c2)
    int sched_c2 = 16;                       // tiles per dynamic chunk
    double taper_half = 0.0, taper_quarter = 0.0;   // share of a matrix row cut into c2/2 and c2/4 chunks (NGU_EPI_TAPER=h,q)
    int sched_min_avg = kDynMinAvgTilesDefault;   // fair share (tiles per warp) below which the schedule is static only
    int sched_min_tiles = kMinTilesPerWarp;       // static-only schedules: at least this many tiles per warp
    int2* plan_dev = nullptr;        // pairwise-sum plan for n = B (epilogue rank sums): postfix entries, then the leaves
    int plan_len = 0, plan_leaves = 0;
    // p2p rank-sum exchange (multi-GPU): local inbox + IPC-mapped peer inboxes
    unsigned long long* inbox = nullptr;
    P2PView p2p{};
    bool p2p_connected = false;
    DevState* state = nullptr;
    bool use_pdl = true;
    unsigned long long* timeline = nullptr;   // NGU_EPI_DEBUG_TIMELINE=1: per-warp timestamps of the last scan
    int debug_phase_mask = ~0;       // NGU_EPI_DEBUG_PHASES: experiments only (tools/sweep.py)
    int debug_scan = 0;              // NGU_EPI_DEBUG_SCAN: experiments only (ScanParams::debug)
    unsigned long long* epi_stamps = nullptr;   // NGU_EPI_DEBUG_EPILOGUE=1: phase stamps of the last epilogue launch
    // staging for the *_host entry points
    cudaStream_t own_stream = nullptr;
    float* q_stage_dev = nullptr;    // 2 x {[B][32] q, [B] alpha}: one H2D transfer per step, buffers alternate with the step
                                     // parity (a pre-armed step stages its inputs while the previous epilogue may still read its own)
    size_t q_stage_stride = 0;       // floats between the two buffers
    // pre-armed host step (armed_stage_kernel, epilogue.cuh)
    bool arm_enabled = true;         // NGU_EPI_HOST_ARM=0 switches it off
    long long arm_window_ns = 150000;   // NGU_EPI_ARM_WINDOW_US: how long an armed step waits for the doorbell
    unsigned int arm_seq_next = 0;   // set while enqueue_armed builds its launches: they carry this sequence number
    unsigned int armed_seq = 0;      // host_seq value the enqueued, waiting launches belong to (0 = none)
    unsigned int* doorbell = nullptr;        // = flag_mapped + 16 (own 64-byte line): host -> GPU
    unsigned int* doorbell_dev = nullptr;
    unsigned int* arm_status = nullptr;      // = flag_mapped + 32: GPU -> host acknowledgement
    unsigned int* arm_status_dev = nullptr;
    unsigned int* arm_cancel_dev = nullptr;  // device word `go`: seq << 1 | cancelled, stored by armed_stage_kernel, polled by the armed scan
    std::chrono::steady_clock::time_point last_return{};   // when the previous ngu_epi_step_host returned
    bool have_last_return = false;
    long arm_hits = 0, arm_misses = 0;       // armed steps that ran / were cancelled (host profile)
    uint8_t* done_stage_dev = nullptr;
    float* pinned = nullptr;         // host, mapped: [B*32 q][B alpha] floats + B done bytes
    float* pinned_dev = nullptr;     // device alias of pinned (stage_inputs_kernel reads it over PCIe)
    bool stage_by_kernel = false;    // ngu_epi_step_host pulls its inputs with stage_inputs_kernel (large shapes) or the
                                     // copy engine (small ones); NGU_EPI_HOST_STAGE=copy|kernel forces one (tools/host_path_ab.py)
    float* r_mapped = nullptr;       // host, mapped: [2][B] rewards written by the epilogue over PCIe (zero-copy)
    float* r_mapped_dev = nullptr;   // device alias of r_mapped
    // completion of the *_host step: a mapped word the epilogue stores after its last write and the host
    // spins on (a stream synchronize costs a few us more of driver wake-up per step).  Measured and
    // rejected: (1) reading the inputs in place from mapped memory instead of the staged H2D copy - 1 us
    // better at 256 x 30k, 25 us worse on small shapes where the PCIe reads land on the critical path;
    // (2) replaying copy + scan + epilogue as one CUDA graph launch - the enqueue drops from 9.5 to 3.5 us
    // but the GPU side starts later and loses more (wait 162 -> 174 us at 256 x 30k).
    unsigned int* flag_mapped = nullptr;
    unsigned int* flag_mapped_dev = nullptr;
    unsigned int* err_mapped = nullptr;       // = flag_mapped + 8: sticky error word written by the epilogue (check_sticky)
    unsigned int* err_mapped_dev = nullptr;
    cudaStream_t user_stream = nullptr;       // last stream the caller passed to a stream-taking entry point
    bool user_stream_dirty = false;           // ... and it may still hold work the *_host calls must be ordered after
    cudaEvent_t order_ev = nullptr;
    unsigned int host_seq = 0;
    bool host_flag_wait = true;      // experiments: NGU_EPI_HOST_PATH=sync -> stream synchronize instead
    bool host_profile = false;       // NGU_EPI_HOST_PROFILE=1: wall-clock split of ngu_epi_step_host, printed at destroy
    double prof_us[4] = {0, 0, 0, 0};   // stage-in, enqueue, wait, copy-out
    long prof_calls = 0;
    // launch geometry
    Variant var{};
    CUtensorMap tmap{};
    int grid = 0, tiles_per_actor = 0, max_cand = 0;
    int64_t total_tiles = 0;
    bool scanned = false;
    int64_t launches = 0;
    // optional per-launch timing of kernel A (bench.py roofline)
    std::vector<cudaEvent_t> t_beg, t_end;
    size_t t_used = 0;
    std::string err;
};

namespace {

int fail(ngu_epi* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
#define CU_TRY(h, expr)                                                                         \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(h, NGU_ECUDA, fmt("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)

// Pre-armed host step: launches of the NEXT ngu_epi_step_host may be sitting in the private stream, waiting for the
// doorbell.  Every other entry point cancels them first (they exit without touching any state), so that nothing it
// enqueues or synchronises on has to sit out the doorbell window.  Returns when the cancellation is acknowledged.
void disarm(ngu_epi* h) {
    if (h->armed_seq == 0u) return;
    const unsigned int seq = h->armed_seq;
    h->armed_seq = 0u;
    __atomic_store_n(h->doorbell, (seq << 2) | 2u, __ATOMIC_RELEASE);
    volatile unsigned int* st = h->arm_status;
    for (unsigned spins = 1; (*st >> 1) != seq; ++spins) {
        if ((spins & 0xFFFFu) == 0u && cudaStreamQuery(h->own_stream) != cudaErrorNotReady) break;   // drained or failed
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    h->arm_misses += 1;
    h->cur ^= 1;          // the cancelled scan never ran: the context of the last real scan is current again
}
// Every entry point runs on the handle's GPU and leaves the caller's current device as it found it
// (a handle on cuda:1 must not silently switch a thread whose torch current device is cuda:0).
struct DeviceGuard {
    int prev = -1, rc = NGU_OK;
    explicit DeviceGuard(ngu_epi* h) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e == cudaSuccess && prev != h->cfg.device) e = cudaSetDevice(h->cfg.device); else if (e == cudaSuccess) prev = -1;
        if (e != cudaSuccess) { prev = -1; rc = fail(h, NGU_ECUDA, fmt("cudaSetDevice(%d): %s", h->cfg.device, cudaGetErrorString(e))); }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define BIND_DEVICE_KEEP_ARMED(h) DeviceGuard dev_guard_(h); if (dev_guard_.rc) return dev_guard_.rc
// ... and every entry point but the fused steps themselves ends a chain of pipelined steps (they restore the flag)
#define BIND_DEVICE(h) BIND_DEVICE_KEEP_ARMED(h); disarm(h); (h)->chain_open = false

// Sticky error: once a merge or an exchange round has timed out (DevState::exchange_timeouts != 0, mirrored by
// the epilogue into a mapped host word) every later result is wrong - the step entry points refuse to go on.
int check_sticky(ngu_epi* h) {
    if (h->err_mapped && *reinterpret_cast<volatile unsigned int*>(h->err_mapped) != 0u)
        return fail(h, NGU_ESTATE, "a k-NN merge or a p2p exchange round timed out earlier (ngu_epi_state.exchange_timeouts != 0): "
                                   "rewards since then are NaN and the replicated statistics may have diverged; recreate the handle");
    return NGU_OK;
}

// *_host calls run on the handle's private stream.  Work the caller enqueued through this handle on ITS stream
// (append, device-mask reset, device-pointer steps) must be ordered before them: remember that stream and make
// the private stream wait for it at the next *_host entry.
void note_user_stream(ngu_epi* h, cudaStream_t s) { h->user_stream = s; h->user_stream_dirty = true; }
int order_after_user_stream(ngu_epi* h) {
    if (!h->user_stream_dirty) return NGU_OK;
    CU_TRY(h, cudaEventRecord(h->order_ev, h->user_stream));
    CU_TRY(h, cudaStreamWaitEvent(h->own_stream, h->order_ev, 0));
    h->user_stream_dirty = false;
    return NGU_OK;
}

int build_tensor_map(ngu_epi* h) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
        return fail(h, NGU_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    auto encode = reinterpret_cast<PFN_tensorMapEncodeTiled>(fn);
    const cuuint64_t rows = static_cast<cuuint64_t>(h->cfg.n_actors) * h->cfg.capacity;
    const cuuint64_t gdim[2] = {32, rows};
    const cuuint64_t gstride[1] = {128};   // bytes between rows
    const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(32 * h->var.rpl)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->memory, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, NGU_ECUDA, fmt("cuTensorMapEncodeTiled failed: CUresult %d", (int)r));
    return NGU_OK;
}

// Launch geometry + chunk schedule of the scan kernel (ScanSched, scan_topk.cuh).
//   big shapes  : one full wave; every warp gets a static chunk of `sched_static_frac` of its fair
//                 share, the rest is handed out dynamically in chunks of c2 tiles, interleaved over the
//                 actors of the dynamic region;
//   small shapes: (fair share below kDynMinAvgTilesDefault tiles) static only, >= kMinTilesPerWarp tiles per
//                 warp, grid sized to the work - these are launch-latency bound, keep the atomics out.
void plan_geometry(ngu_epi* h) {
    const int tile_rows = 32 * h->var.rpl;
    h->tiles_per_actor = (h->cfg.capacity + tile_rows - 1) / tile_rows;
    h->total_tiles = static_cast<int64_t>(h->cfg.n_actors) * h->tiles_per_actor;
    const int64_t T = h->total_tiles;
    const int64_t max_warps = static_cast<int64_t>(h->sm_count) * h->var.ctas_per_sm * h->var.warps;
    ScanSched& sc = h->sched;
    sc = ScanSched{};
    sc.total_tiles = static_cast<int32_t>(T);
    sc.c2 = 1; sc.rows = 1; sc.cols = 0; sc.row_tiles = 1;
    int min_chunk;
    if (T < h->sched_min_avg * max_warps || h->sched_static_frac >= 1.0 || h->sched_c2 <= 0) {
        int64_t tpw = (T + max_warps - 1) / max_warps;
        if (tpw < h->sched_min_tiles) tpw = h->sched_min_tiles;
        sc.s1 = static_cast<int32_t>(tpw);
        sc.n1 = static_cast<int32_t>((T + tpw - 1) / tpw);
        sc.n_chunks = sc.n1;
        h->grid = (sc.n1 + h->var.warps - 1) / h->var.warps;
        min_chunk = sc.s1;
    } else {
        h->grid = h->sm_count * h->var.ctas_per_sm;
        const int64_t W = max_warps;
        int64_t s1 = static_cast<int64_t>(h->sched_static_frac * static_cast<double>(T) / static_cast<double>(W));
        const int c2 = h->sched_c2;
        sc.n1 = s1 >= 1 ? static_cast<int32_t>(W) : 0;
        sc.s1 = s1 >= 1 ? static_cast<int32_t>(s1) : 1;
        const int64_t rest = T - static_cast<int64_t>(sc.n1) * sc.s1;
        sc.c2 = c2;
        const int wide = h->tiles_per_actor / c2 > 0 ? h->tiles_per_actor / c2 : 1;   // ~ chunks per actor, untapered
        int cb = 0, cc = 0;
        if (c2 % 4 == 0 && wide >= 8) {      // taper: the last shares of a row in half and quarter chunks
            cb = static_cast<int>(h->taper_half * wide) * 2;
            cc = static_cast<int>(h->taper_quarter * wide) * 4;
            if (cb / 2 + cc / 4 > wide - 1) { cb = 0; cc = 0; }
        }
        sc.cols_a = wide - cb / 2 - cc / 4;
        sc.cols_b = cb;
        sc.cols = sc.cols_a + cb + cc;
        sc.row_tiles = sc.cols_a * c2 + cb * (c2 / 2) + cc * (c2 / 4);        // = wide * c2
        sc.rows = static_cast<int32_t>((rest + sc.row_tiles - 1) / sc.row_tiles);
        if (sc.rows < 1) sc.rows = 1;
        const int narrow = cc ? c2 / 4 : cb ? c2 / 2 : c2;
        min_chunk = static_cast<int>(sc.n1 > 0 && sc.s1 < narrow ? sc.s1 : narrow);
        sc.n_chunks = sc.n1 + sc.rows * sc.cols;
    }
    // static schedules: with many runs per actor the CTA-level combine pays (measured: >= 117 runs win 2-5 us after the
    // last tile, 32-72 runs lose 3-6 us)
    h->cta_merge = sc.n_chunks == sc.n1 && h->tiles_per_actor / sc.s1 >= 96 ? 1 : 0;
    if (const char* e = getenv("NGU_EPI_CTA_MERGE")) h->cta_merge = (atoi(e) != 0 && sc.n_chunks == sc.n1) ? 1 : 0;
    // worst case every run publishes k candidates; runs per actor <= chunks overlapping it
    h->max_cand = ((h->tiles_per_actor + min_chunk - 1) / min_chunk + 3) * h->cfg.k;
}

void apply_sched_env(ngu_epi* h) {
    if (const char* e = getenv("NGU_EPI_SCHED")) {   // experiments and tests: "static_frac,c2[,min_avg_tiles[,min_tiles_per_warp]]"
        double f = 0.6; int c2 = 16, mn = kDynMinAvgTilesDefault, mt = kMinTilesPerWarp;
        const int got = sscanf(e, "%lf,%d,%d,%d", &f, &c2, &mn, &mt);
        if (got >= 1 && f >= 0.0) h->sched_static_frac = f;
        if (got >= 2) h->sched_c2 = c2;
        if (got >= 3 && mn >= 0) h->sched_min_avg = mn;
        if (got >= 4 && mt >= 1) h->sched_min_tiles = mt;
    }
    if (const char* e = getenv("NGU_EPI_TAPER")) {   // experiments: "half_share,quarter_share" of every matrix row
        double a = 0.0, b = 0.0;
        const int got = sscanf(e, "%lf,%lf", &a, &b);
        if (got >= 1 && a >= 0.0 && a < 1.0) h->taper_half = a;
        if (got >= 2 && b >= 0.0 && b < 1.0) h->taper_quarter = b;
    }
}

int default_state(ngu_epi* h) {
    ngu_epi_state s;
    std::memset(&s, 0, sizeof s);
    for (int j = 0; j < NGU_EPI_MAX_K; ++j) s.ed_var[j] = 1.0;           // mpi_util.py:39-41
    s.ed_count = h->cfg.prior_count;                                      // mpi_util.py:42
    s.epi_var = 1.0;  s.epi_count = 1e-4;                                 // RunningMeanStd() defaults
    s.intr_var = 1.0; s.intr_count = 1e-4;
    s.ll_var = 1.0; s.ll_count = 1e-4;                                    // LifelongNovelty.ll_rms, lifelong_novelty.py:26
    return ngu_epi_set_state(h, &s);
}

}  // namespace

extern "C" {

int ngu_epi_abi_version(void) { return NGU_EPI_ABI_VERSION; }

const char* ngu_epi_last_error(const ngu_epi* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ngu_epi_create(const ngu_epi_cfg* cfg, float* memory_dev, ngu_epi** out) {
    if (!cfg || !out) return fail(nullptr, NGU_EINVAL, "null cfg/out");
    *out = nullptr;
    if (cfg->dim != NGU_EPI_DIM) return fail(nullptr, NGU_EINVAL, fmt("dim must be %d, got %d", NGU_EPI_DIM, cfg->dim));
    if (cfg->n_actors < 1 || cfg->capacity < 1) return fail(nullptr, NGU_EINVAL, "n_actors and capacity must be >= 1");
    if (cfg->k < 1 || cfg->k > NGU_EPI_MAX_K) return fail(nullptr, NGU_EINVAL, fmt("k must be in 1..%d, got %d", NGU_EPI_MAX_K, cfg->k));
    if (cfg->k > cfg->capacity)  // torch.topk raises "selected index k out of range" (episodic_novelty.py:54)
        return fail(nullptr, NGU_EINVAL, fmt("k=%d exceeds capacity=%d (the reference's topk would raise)", cfg->k, cfg->capacity));
    if (cfg->mode != NGU_MODE_REFERENCE && cfg->mode != NGU_MODE_PAPER) return fail(nullptr, NGU_EINVAL, "unknown mode");
    if (static_cast<int64_t>(cfg->n_actors) * cfg->capacity + 256 >= (1ll << 31))
        return fail(nullptr, NGU_EINVAL, "n_actors*capacity must stay below 2^31 rows per handle; shard the actors");
    if (cfg->scan_variant < 0 || cfg->scan_variant >= kNumVariants) return fail(nullptr, NGU_EINVAL, "unknown scan_variant");
    if (epilogue_smem_bytes(cfg->n_actors, cfg->k, false) > kEpilogueStageBytes)
        return fail(nullptr, NGU_EINVAL, "n_actors*(k+1) exceeds 40960 per handle (the epilogue stages the k-NN matrix on chip); shard the actors");
    if (memory_dev && (reinterpret_cast<uintptr_t>(memory_dev) & 127u))
        return fail(nullptr, NGU_EINVAL, "memory_dev must be 128-byte aligned (one row)");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, NGU_ECUDA, fmt("no CUDA device (%s); libngu_epi has no CPU fallback", cudaGetErrorString(e)));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, NGU_EINVAL, "device ordinal out of range");

    ngu_epi* h = new (std::nothrow) ngu_epi();
    if (!h) return fail(nullptr, NGU_ENOMEM, "host allocation failed");
    h->cfg = *cfg;
#define CREATE_TRY(expr)                                                                       \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            g_create_error = fmt("%s: %s", #expr, cudaGetErrorString(e_));                     \
            ngu_epi_destroy(h);                                                                \
            return NGU_ECUDA;                                                                  \
        }                                                                                      \
    } while (0)
    // the handle's GPU for the rest of this call; the caller's current device is put back on every way out
    struct RestoreDevice {
        int prev = -1;
        ~RestoreDevice() { if (prev >= 0) cudaSetDevice(prev); }
    } restore_device;
    if (cudaGetDevice(&restore_device.prev) != cudaSuccess || restore_device.prev == cfg->device) restore_device.prev = -1;
    CREATE_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CREATE_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) {
        g_create_error = fmt("device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
        ngu_epi_destroy(h);
        return NGU_EINVAL;
    }
    h->sm_count = prop.multiProcessorCount;

    h->var = kVariants[cfg->scan_variant];
    if (cfg->scan_ctas_per_sm > 0) h->var.ctas_per_sm = cfg->scan_ctas_per_sm;
    if (static_cast<size_t>(h->var.smem) * h->var.ctas_per_sm > prop.sharedMemPerMultiprocessor ||
        static_cast<size_t>(h->var.smem) > prop.sharedMemPerBlockOptin) {
        g_create_error = fmt("scan variant %d needs %d B smem x %d CTAs/SM; device offers %zu", cfg->scan_variant, h->var.smem,
                             h->var.ctas_per_sm, (size_t)prop.sharedMemPerMultiprocessor);
        ngu_epi_destroy(h);
        return NGU_EINVAL;
    }
    CREATE_TRY(cudaFuncSetAttribute(h->var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->var.smem));
    if (h->var.fn_dbg != h->var.fn) CREATE_TRY(cudaFuncSetAttribute(h->var.fn_dbg, cudaFuncAttributeMaxDynamicSharedMemorySize, h->var.smem));
    CREATE_TRY(cudaFuncSetAttribute(reinterpret_cast<const void*>(&epilogue_kernel<kEpilogueThreads>),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kEpilogueStageBytes)));
    if (const char* e = getenv("NGU_EPI_PDL")) h->use_pdl = atoi(e) != 0;
    if (const char* e = getenv("NGU_EPI_DEBUG_TIMELINE")) {
        if (atoi(e) != 0) {
            const size_t n = static_cast<size_t>(prop.multiProcessorCount) * 4 * 32 * 8;   // >= max warps * 8
            CREATE_TRY(cudaMalloc(&h->timeline, n * sizeof(unsigned long long)));
            CREATE_TRY(cudaMemset(h->timeline, 0, n * sizeof(unsigned long long)));
        }
    }
    if (const char* e = getenv("NGU_EPI_DEBUG_PHASES")) h->debug_phase_mask = atoi(e);
    if (const char* e = getenv("NGU_EPI_DEBUG_EPILOGUE")) {
        if (atoi(e) != 0) {
            CREATE_TRY(cudaMalloc(&h->epi_stamps, 16 * sizeof(unsigned long long)));
            CREATE_TRY(cudaMemset(h->epi_stamps, 0, 16 * sizeof(unsigned long long)));
        }
    }
    if (const char* e = getenv("NGU_EPI_DEBUG_SCAN")) h->debug_scan = atoi(e);
    if (h->debug_scan != 0 || h->timeline != nullptr) h->var.fn = h->var.fn_dbg;
    apply_sched_env(h);
    plan_geometry(h);

    const int B = cfg->n_actors, k = cfg->k;
    const size_t mem_bytes = static_cast<size_t>(B) * cfg->capacity * 32 * sizeof(float);
    if (memory_dev) {
        h->memory = memory_dev;
    } else {
        CREATE_TRY(cudaMalloc(&h->memory, mem_bytes));
        h->owns_memory = true;
        CREATE_TRY(cudaMemset(h->memory, 0, mem_bytes));                 // torch.zeros, episodic_novelty.py:32
    }
    for (ngu_epi::ScanCtx& cx : h->ctx) {
        // epoch-tagged candidates: zeroed once (tag 0 is never valid), 16 bytes per entry
        const size_t cbytes = static_cast<size_t>(B) * h->max_cand * 2 * sizeof(u64);
        CREATE_TRY(cudaMalloc(&cx.cand, cbytes));
        CREATE_TRY(cudaMemset(cx.cand, 0, cbytes));
        CREATE_TRY(cudaMalloc(&cx.topk_d2, static_cast<size_t>(B) * k * sizeof(float)));
        CREATE_TRY(cudaMalloc(&cx.topk_idx, static_cast<size_t>(B) * k * sizeof(int32_t)));
        CREATE_TRY(cudaMalloc(&cx.dist, static_cast<size_t>(B) * k * sizeof(float)));
        CREATE_TRY(cudaMalloc(&cx.actor_state, static_cast<size_t>(B) * 2 * sizeof(u64)));
        CREATE_TRY(cudaMemset(cx.actor_state, 0, static_cast<size_t>(B) * 2 * sizeof(u64)));
        CREATE_TRY(cudaMalloc(&cx.leftover, static_cast<size_t>(B) * sizeof(int32_t)));
        CREATE_TRY(cudaMalloc(&cx.sched_ctr, 8 * sizeof(unsigned int)));
        unsigned int init[8] = {0u, 0u, 1u, 0u, 0u, 0u, 0u, 0u};   // first launch epoch = 1
        if (const char* e = getenv("NGU_EPI_DEBUG_EPOCH")) init[2] = static_cast<unsigned int>(strtoul(e, nullptr, 0));   // tests: cross the wrap
        if (init[2] == 0u) init[2] = 1u;
        CREATE_TRY(cudaMemcpy(cx.sched_ctr, init, sizeof init, cudaMemcpyHostToDevice));
    }
    CREATE_TRY(cudaMalloc(&h->pipe_words, 256));
    CREATE_TRY(cudaMemset(h->pipe_words, 0, 256));
    if (const char* e = getenv("NGU_EPI_PIPE")) h->pipe_enabled = atoi(e) != 0;
    CREATE_TRY(cudaMalloc(&h->rank_sums, NGU_EPI_RANK_SUMS(NGU_EPI_MAX_K) * sizeof(double)));
    CREATE_TRY(cudaMalloc(&h->log_sums, 8 * sizeof(double)));
    CREATE_TRY(cudaMalloc(&h->log_stash, 2 * static_cast<size_t>(cfg->n_actors) * sizeof(float)));
    if (const char* e = getenv("NGU_EPI_DEBUG_POLL_LIMIT")) h->poll_limit_cycles = atoll(e);
    if (const char* e = getenv("NGU_EPI_L2")) {
        int m = 0; float f = 1.0f;
        const int got = sscanf(e, "%d,%f", &m, &f);
        if (got >= 1 && m >= 0 && m <= 2) h->l2_mode = m;
        if (got >= 2 && f > 0.0f && f <= 1.0f) h->l2_frac = f;
    }
    if (const char* e = getenv("NGU_EPI_PREFETCH_TILES")) h->prefetch_tiles = atoi(e) < 0 ? 0 : (atoi(e) > kMaxPrefetchTiles ? kMaxPrefetchTiles : atoi(e));
    h->arm_prefetch_tiles = h->prefetch_tiles;
    if (const char* e = getenv("NGU_EPI_ARM_PREFETCH_TILES")) h->arm_prefetch_tiles = atoi(e) < 0 ? 0 : (atoi(e) > kMaxPrefetchTiles ? kMaxPrefetchTiles : atoi(e));
    CREATE_TRY(cudaMalloc(&h->state, sizeof(DevState)));
    CREATE_TRY(cudaMemset(h->state, 0, sizeof(DevState)));
    {
        std::vector<int2> plan;
        build_pairwise_plan(0, B, plan);
        h->plan_len = static_cast<int>(plan.size());
        for (int i = 0; i < h->plan_len; ++i)
            if (plan[i].y > 0) plan.push_back(plan[i]);
        h->plan_leaves = static_cast<int>(plan.size()) - h->plan_len;
        if (h->plan_leaves > kPlanMaxLeaves || h->plan_len > kPlanMaxEntries) {
            g_create_error = fmt("n_actors=%d needs %d pairwise-sum leaves, the epilogue holds %d; shard the actors", B,
                                 h->plan_leaves, kPlanMaxLeaves);
            ngu_epi_destroy(h);
            return NGU_EINVAL;
        }
        CREATE_TRY(cudaMalloc(&h->plan_dev, plan.size() * sizeof(int2)));
        CREATE_TRY(cudaMemcpy(h->plan_dev, plan.data(), plan.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    CREATE_TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->q_stage_stride = (static_cast<size_t>(B) * 33 + 4 + 31) / 32 * 32;       // floats; keeps both buffers 128-byte aligned
    CREATE_TRY(cudaMalloc(&h->q_stage_dev, 2 * h->q_stage_stride * sizeof(float)));
    CREATE_TRY(cudaMemset(h->q_stage_dev, 0, 2 * h->q_stage_stride * sizeof(float)));
    CREATE_TRY(cudaMalloc(&h->arm_cancel_dev, 64));
    CREATE_TRY(cudaMemset(h->arm_cancel_dev, 0, 64));
    CREATE_TRY(cudaMalloc(&h->done_stage_dev, static_cast<size_t>(B)));
    CREATE_TRY(cudaHostAlloc(&h->pinned, static_cast<size_t>(B) * 33 * sizeof(float) + B + 16, cudaHostAllocMapped));
    CREATE_TRY(cudaHostGetDevicePointer(&h->pinned_dev, h->pinned, 0));
    // measured (tools/host_path_ab.py): the PDL-chained kernel wins 1.5 us per step at 256 x 30k and 3.7 us at
    // 1024 x 30k, the copy engine wins 1.2-1.8 us on steps of a few tens of microseconds
    h->stage_by_kernel = static_cast<double>(B) * cfg->capacity * 128.0 >= 512e6;
    if (const char* e = getenv("NGU_EPI_HOST_STAGE")) h->stage_by_kernel = std::strcmp(e, "copy") != 0;
    CREATE_TRY(cudaHostAlloc(&h->r_mapped, static_cast<size_t>(B) * 2 * sizeof(float), cudaHostAllocMapped));
    CREATE_TRY(cudaHostGetDevicePointer(&h->r_mapped_dev, h->r_mapped, 0));
    CREATE_TRY(cudaHostAlloc(&h->flag_mapped, 256, cudaHostAllocMapped));
    CREATE_TRY(cudaHostGetDevicePointer(&h->flag_mapped_dev, h->flag_mapped, 0));
    *h->flag_mapped = 0u;
    h->err_mapped = h->flag_mapped + 8; h->err_mapped_dev = h->flag_mapped_dev + 8;
    *h->err_mapped = 0u;
    h->doorbell = h->flag_mapped + 16; h->doorbell_dev = h->flag_mapped_dev + 16;
    h->arm_status = h->flag_mapped + 32; h->arm_status_dev = h->flag_mapped_dev + 32;
    *h->doorbell = 0u; *h->arm_status = 0u;
    if (const char* e = getenv("NGU_EPI_HOST_ARM")) h->arm_enabled = atoi(e) != 0;
    if (const char* e = getenv("NGU_EPI_ARM_WINDOW_US")) h->arm_window_ns = atoll(e) * 1000ll;
    CREATE_TRY(cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));

    if (const char* e = getenv("NGU_EPI_HOST_PATH")) h->host_flag_wait = std::strcmp(e, "sync") != 0;
    if (const char* e = getenv("NGU_EPI_HOST_PROFILE")) h->host_profile = atoi(e) != 0;
#undef CREATE_TRY

    EpiConsts& c = h->consts;
    c.eps = cfg->kernel_epsilon; c.cluster = cfg->cluster_distance; c.pseudo = cfg->pseudo_counts;
    c.s_max = cfg->max_similarity; c.max_rew = cfg->max_reward_scale;
    c.k = k; c.n_actors = B; c.capacity = cfg->capacity;
    c.largest = (cfg->mode == NGU_MODE_REFERENCE) ? 1 : 0;
    c.sqrt_dist = (cfg->mode == NGU_MODE_REFERENCE) ? 1 : 0;

    int rc = build_tensor_map(h);
    if (rc == NGU_OK) rc = default_state(h);
    if (rc != NGU_OK) {
        g_create_error = h->err;
        ngu_epi_destroy(h);
        return rc;
    }
    *out = h;
    return NGU_OK;
}

void ngu_epi_destroy(ngu_epi* h) {
    if (!h) return;
    if (h->host_profile && h->prof_calls > 0)
        fprintf(stderr, "[ngu_epi] step_host x%ld: stage-in %.2f us, enqueue %.2f us, wait %.2f us, copy-out %.2f us per call; "
                        "pre-armed steps run %ld, cancelled %ld\n",
                h->prof_calls, h->prof_us[0] / h->prof_calls, h->prof_us[1] / h->prof_calls, h->prof_us[2] / h->prof_calls,
                h->prof_us[3] / h->prof_calls, h->arm_hits, h->arm_misses);
    DeviceGuard dev_guard_(h);
    disarm(h);
    if (h->own_stream) { cudaStreamSynchronize(h->own_stream); cudaStreamDestroy(h->own_stream); }
    for (cudaEvent_t e : h->t_beg) cudaEventDestroy(e);
    for (cudaEvent_t e : h->t_end) cudaEventDestroy(e);
    if (h->order_ev) cudaEventDestroy(h->order_ev);
    if (h->p2p_connected)
        for (int g = 0; g < h->p2p.world; ++g)
            if (g != h->p2p.rank && h->p2p.peer[g]) cudaIpcCloseMemHandle(h->p2p.peer[g]);
    cudaFree(h->inbox);
    if (h->owns_memory) cudaFree(h->memory);
    for (ngu_epi::ScanCtx& cx : h->ctx) {
        cudaFree(cx.cand); cudaFree(cx.topk_d2); cudaFree(cx.topk_idx); cudaFree(cx.dist);
        cudaFree(cx.actor_state); cudaFree(cx.sched_ctr); cudaFree(cx.leftover);
    }
    cudaFree(h->pipe_words);
    cudaFree(h->rank_sums); cudaFree(h->log_sums); cudaFree(h->log_stash); cudaFree(h->state); cudaFree(h->plan_dev); cudaFree(h->timeline); cudaFree(h->epi_stamps);
    cudaFree(h->q_stage_dev); cudaFree(h->done_stage_dev); cudaFree(h->arm_cancel_dev);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->r_mapped) cudaFreeHost(h->r_mapped);
    if (h->flag_mapped) cudaFreeHost(h->flag_mapped);

    delete h;
}

int ngu_epi_memory(ngu_epi* h, float** memory_dev, int64_t strides[3]) {
    if (!h || !memory_dev || !strides) return fail(h, NGU_EINVAL, "null argument");
    *memory_dev = h->memory;
    strides[0] = 32;                                          // slot
    strides[1] = static_cast<int64_t>(h->cfg.capacity) * 32;  // actor
    strides[2] = 1;
    return NGU_OK;
}

// Launch with (optionally) the programmatic-stream-serialization attribute: the kernel may
// become resident before its predecessor in the stream has finished; both of our kernels
// order themselves with griddepcontrol.wait.
static cudaError_t launch_pdl(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                              void** args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelExC(&cfg, fn, args);
}

// Sequence numbers of a fused step's launches (all zero: a launch outside any chain - split sequence, RND step, capture).
struct ChainArgs { unsigned int seq = 0, prev_seq = 0, wait_epi = 0; bool pipe = false; const float* q_prev = nullptr; };

static int launch_scan(ngu_epi* h, const float* q_dev, cudaStream_t s, const ChainArgs& ca = ChainArgs{}) {
    if (reinterpret_cast<uintptr_t>(q_dev) & 15u)   // the scan fetches query rows with 16-byte bulk copies
        return fail(h, NGU_EINVAL, "q_dev must be 16-byte aligned");
    h->cur ^= 1;                                    // the other context: the previous step may still be using its own
    const ngu_epi::ScanCtx& cx = h->ctx[h->cur];
    ScanParams sp{};
    sp.pipe = ca.pipe ? 1 : 0; sp.seq = ca.seq; sp.prev_seq = ca.prev_seq; sp.wait_epi = ca.wait_epi; sp.q_prev = ca.q_prev;
    sp.epi_done = h->pipe_words; sp.scan_cursor = reinterpret_cast<unsigned long long*>(h->pipe_words + 16);
    sp.cursor = &h->state->cursor;
    sp.q = q_dev; sp.cand = cx.cand;
    sp.n_actors = h->cfg.n_actors; sp.capacity = h->cfg.capacity; sp.k = h->cfg.k;
    sp.largest = h->consts.largest; sp.sqrt_dist = h->consts.sqrt_dist; sp.neg_zero2 = kNegZero2;
    sp.tiles_per_actor = h->tiles_per_actor;
    sp.max_cand = h->max_cand; sp.sched = h->sched; sp.sched_ctr = cx.sched_ctr;
    sp.leftover = cx.leftover; sp.poll_limit_cycles = h->poll_limit_cycles;
    sp.prefetch_tiles = h->arm_seq_next != 0u ? h->arm_prefetch_tiles : h->prefetch_tiles;
    sp.cta_merge = h->cta_merge;
    sp.l2_mode = h->l2_mode; sp.l2_frac = h->l2_frac;
    sp.actor_state = cx.actor_state;
    sp.timeouts = reinterpret_cast<unsigned long long*>(&h->state->exchange_timeouts);
    sp.topk_d2 = cx.topk_d2; sp.topk_idx = cx.topk_idx; sp.dist = cx.dist;
    sp.timeline = h->timeline;
    sp.arm_go = h->arm_cancel_dev; sp.arm_seq = h->arm_seq_next;
    sp.debug = h->debug_scan;
    void* args[] = {&h->tmap, &sp};
    const bool timed = h->t_used < h->t_beg.size();
    if (timed) CU_TRY(h, cudaEventRecord(h->t_beg[h->t_used], s));
    CU_TRY(h, launch_pdl(h->var.fn, dim3(h->grid), dim3(h->var.warps * 32), h->var.smem, s, h->use_pdl && !timed, args));
    if (timed) CU_TRY(h, cudaEventRecord(h->t_end[h->t_used++], s));
    h->launches += 1;
    return NGU_OK;
}

struct RndArgs { const float* pred; const float* targ; float* alpha_out; };
struct HostFlag { unsigned int* flag; unsigned int seq; };

static int launch_epilogue(ngu_epi* h, int phases, const double* sums_in, const float* alpha, float* r_int,
                           float* r_epi, double* log_sums, const float* q, cudaStream_t s,
                           const RndArgs* rnd = nullptr, const HostFlag* hf = nullptr, const ChainArgs& ca = ChainArgs{}) {
    EpilogueParams ep{};
    ep.epi_done = h->pipe_words; ep.seq = ca.seq; ep.wait_prev = ca.pipe ? ca.prev_seq : 0u;
    if (hf) { ep.host_flag = hf->flag; ep.host_seq = hf->seq; }
    ep.host_err = h->err_mapped_dev;
    ep.arm_go = h->arm_cancel_dev; ep.arm_seq = h->arm_seq_next;
    ep.dbg_stamps = h->epi_stamps;
    if (rnd) { ep.rnd_pred = rnd->pred; ep.rnd_targ = rnd->targ; ep.alpha_out = rnd->alpha_out; }
    ep.dist = h->ctx[h->cur].dist; ep.rank_sums = h->rank_sums; ep.sums_in = sums_in; ep.alpha = alpha;
    ep.r_int = r_int; ep.r_epi = r_epi; ep.log_sums = log_sums; ep.q = q; ep.memory = h->memory;
    ep.state = h->state; ep.phases = phases & h->debug_phase_mask;
    if (h->p2p_connected && (phases & PH_FINISH) && (phases & PH_RANKSUMS)) {   // fused single-launch step only
        ep.phases |= PH_EXCHANGE;
        ep.p2p = h->p2p;
    }
    ep.plan = h->plan_dev; ep.plan_len = h->plan_len; ep.n_leaves = h->plan_leaves;
    ep.log_stash = h->log_stash;
    // dynamic smem: [dist B*k][alpha B][q B*32] when it all fits, else without q
    ep.stage_q = epilogue_smem_bytes(h->cfg.n_actors, h->cfg.k, true) <= kEpilogueStageBytes ? 1 : 0;
    const size_t dyn = epilogue_smem_bytes(h->cfg.n_actors, h->cfg.k, ep.stage_q != 0);
    void* args[] = {&ep, &h->consts};
    CU_TRY(h, launch_pdl(reinterpret_cast<const void*>(&epilogue_kernel<kEpilogueThreads>), dim3(1),
                         dim3(kEpilogueThreads), dyn, s, h->use_pdl, args));
    h->launches += 1;
    return NGU_OK;
}

int ngu_epi_scan(ngu_epi* h, const float* q_dev, void* stream) {
    if (!h || !q_dev) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    int rc = check_sticky(h);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    note_user_stream(h, s);
    rc = launch_scan(h, q_dev, s);
    if (rc) return rc;
    rc = launch_epilogue(h, PH_RANKSUMS, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s);
    if (rc) return rc;
    h->scanned = true;
    return NGU_OK;
}

int ngu_epi_rank_sums(ngu_epi* h, double** sums_dev) {
    if (!h || !sums_dev) return fail(h, NGU_EINVAL, "null argument");
    *sums_dev = h->rank_sums;
    return NGU_OK;
}

int ngu_epi_finish(ngu_epi* h, const double* sums_global_dev, const float* alpha_dev, float* r_int_dev,
                   float* r_epi_dev, double* log_sums_dev, void* stream) {
    if (!h || !r_int_dev) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    int rc = check_sticky(h);
    if (rc) return rc;
    if (!h->scanned) return fail(h, NGU_ESTATE, "ngu_epi_finish called before ngu_epi_scan");
    note_user_stream(h, static_cast<cudaStream_t>(stream));
    rc = launch_epilogue(h, PH_FINISH, sums_global_dev ? sums_global_dev : h->rank_sums, alpha_dev, r_int_dev,
                         r_epi_dev, log_sums_dev, nullptr, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    h->scanned = false;
    return NGU_OK;
}

int ngu_epi_log_update(ngu_epi* h, const double* log_sums_global_dev, void* stream) {
    if (!h || !log_sums_global_dev) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    note_user_stream(h, static_cast<cudaStream_t>(stream));
    log_update_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->state, log_sums_global_dev);
    CU_TRY(h, cudaGetLastError());
    h->launches += 1;
    return NGU_OK;
}

int ngu_epi_append(ngu_epi* h, const float* q_dev, void* stream) {
    if (!h || !q_dev) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    note_user_stream(h, static_cast<cudaStream_t>(stream));
    append_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(h->memory, q_dev, h->state, h->cfg.n_actors,
                                                                    h->cfg.capacity);
    CU_TRY(h, cudaGetLastError());
    h->launches += 1;
    return NGU_OK;
}

int ngu_epi_reset(ngu_epi* h, const uint8_t* done_dev, void* stream) {
    if (!h || !done_dev) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    if (static_cast<cudaStream_t>(stream) != h->own_stream) note_user_stream(h, static_cast<cudaStream_t>(stream));
    const int64_t n4 = static_cast<int64_t>(h->cfg.capacity) * 8;
    dim3 grid(static_cast<unsigned>((n4 + 2047) / 2048), h->cfg.n_actors < kResetGroups ? h->cfg.n_actors : kResetGroups);
    reset_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(h->memory, done_dev, h->cfg.n_actors, h->cfg.capacity);
    CU_TRY(h, cudaGetLastError());
    h->launches += 1;
    return NGU_OK;
}

// One fused step.  `pipelined`: the caller asked for ngu_epi_step_pipelined AND the chain was open when it entered, i.e.
// the previous thing this handle enqueued was a fused step; everything else is decided here.
static int step_impl(ngu_epi* h, const float* q_dev, const float* alpha_dev, float* r_int_dev, float* r_epi_dev,
                     cudaStream_t s, const HostFlag* hf, bool pipelined = false) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    const bool capturing = cap != cudaStreamCaptureStatusNone;   // a graph replays frozen arguments: no sequence numbers
    ChainArgs ca;
    if (!capturing) {
        ca.seq = ++h->step_seq;
        if (ca.seq == 0u) ca.seq = ++h->step_seq;
    }
    const bool timed = h->t_used < h->t_beg.size();
    if (pipelined && !capturing && !timed && h->pipe_enabled && h->use_pdl && h->chain_stream == s && h->chain_seq != 0u) {
        ca.pipe = true;
        ca.prev_seq = h->chain_seq;
        ca.wait_epi = h->chain_prev_pipe ? h->chain_prev_seq : 0u;
        ca.q_prev = h->chain_q;
        h->pipe_launches += 1;
    }
    int rc = launch_scan(h, q_dev, s, ca);
    if (rc) return rc;
    // one epilogue launch: rank sums -> running-mean update -> rewards (stashed for the log-only trackers, which the
    // next fused launch brings up to date: epilogue.cuh, log_absorb) -> append
    rc = launch_epilogue(h, PH_RANKSUMS | PH_FINISH | PH_APPEND | PH_LOGFUSE, nullptr, alpha_dev, r_int_dev,
                         r_epi_dev, nullptr, q_dev, s, nullptr, hf, ca);
    if (rc) return rc;
    h->chain_open = ca.seq != 0u;
    h->chain_stream = s;
    h->chain_prev_seq = h->chain_seq; h->chain_seq = ca.seq;
    h->chain_prev_pipe = ca.pipe;
    h->chain_q = q_dev;
    return NGU_OK;
}

int ngu_epi_step(ngu_epi* h, const float* q_dev, const float* alpha_dev, float* r_int_dev, float* r_epi_dev,
                 void* stream) {
    if (!h || !q_dev || !r_int_dev) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    const int rc = check_sticky(h);
    if (rc) return rc;
    note_user_stream(h, static_cast<cudaStream_t>(stream));
    return step_impl(h, q_dev, alpha_dev, r_int_dev, r_epi_dev, static_cast<cudaStream_t>(stream), nullptr);
}

int ngu_epi_step_pipelined(ngu_epi* h, const float* q_dev, const float* alpha_dev, float* r_int_dev, float* r_epi_dev,
                           void* stream) {
    if (!h || !q_dev || !r_int_dev) return fail(h, NGU_EINVAL, "null argument");
    const bool chain_was_open = h->chain_open;
    BIND_DEVICE(h);
    const int rc = check_sticky(h);
    if (rc) return rc;
    note_user_stream(h, static_cast<cudaStream_t>(stream));
    return step_impl(h, q_dev, alpha_dev, r_int_dev, r_epi_dev, static_cast<cudaStream_t>(stream), nullptr, chain_was_open);
}

int ngu_epi_step_rnd(ngu_epi* h, const float* q_dev, const float* pred_dev, const float* targ_dev, int32_t feat_dim,
                     float* r_int_dev, float* r_epi_dev, float* alpha_out_dev, void* stream) {
    if (!h || !q_dev || !pred_dev || !targ_dev || !r_int_dev) return fail(h, NGU_EINVAL, "null argument");
    if (feat_dim != 128) return fail(h, NGU_EINVAL, "the RND feature width must be 128 (rnd_prediction.py:27)");
    BIND_DEVICE(h);
    int rc = check_sticky(h);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    note_user_stream(h, s);
    rc = launch_scan(h, q_dev, s);
    if (rc) return rc;
    RndArgs ra{pred_dev, targ_dev, alpha_out_dev};
    return launch_epilogue(h, PH_RANKSUMS | PH_FINISH | PH_APPEND | PH_LOGFUSE | PH_RND, nullptr, nullptr, r_int_dev,
                           r_epi_dev, nullptr, q_dev, s, &ra);
}

// Spin on the mapped completion word the epilogue kernel stores after its last write (a stream
// synchronize costs 5-10 us of driver wake-up per step); the stream is queried now and then so that a
// failed launch cannot hang the caller.
static int wait_host_flag(ngu_epi* h, unsigned int seq, cudaStream_t s) {
    volatile unsigned int* f = h->flag_mapped;
    for (unsigned spins = 1;; ++spins) {
        if (*f == seq) return NGU_OK;
        if ((spins & 0xFFFu) == 0u) {
            const cudaError_t e = cudaStreamQuery(s);
            if (e == cudaSuccess) {                       // drained: every write of the step is visible by now
                if (*f == seq) return NGU_OK;
                return fail(h, NGU_ECUDA, "the stream drained without the step signalling completion");
            }
            if (e != cudaErrorNotReady) return fail(h, NGU_ECUDA, fmt("step failed: %s", cudaGetErrorString(e)));
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
}

// Enqueue the three launches of host step `seq` in pre-armed form: armed_stage_kernel (waits for the doorbell, pulls
// the inputs), scan, epilogue - the latter two exit at once if the stage kernel reports a cancellation.
static int enqueue_armed(ngu_epi* h, unsigned int seq) {
    const size_t B = h->cfg.n_actors;
    cudaStream_t s = h->own_stream;
    float* stage = h->q_stage_dev + (seq & 1u) * h->q_stage_stride;
    {
        float4* dst = reinterpret_cast<float4*>(stage);
        const float4* src = reinterpret_cast<const float4*>(h->pinned_dev);
        int n_actors = static_cast<int>(B);
        const unsigned int* db = h->doorbell_dev;
        unsigned int* cancel = h->arm_cancel_dev;
        unsigned int* status = h->arm_status_dev;
        long long window = h->arm_window_ns;
        void* args[] = {&dst, &src, &n_actors, &db, &seq, &cancel, &status, &window};
        CU_TRY(h, launch_pdl(reinterpret_cast<const void*>(&armed_stage_kernel), dim3(1), dim3(kArmThreads), 0, s, true, args));
        h->launches += 1;
    }
    const HostFlag hf{h->flag_mapped_dev, seq};
    h->arm_seq_next = seq;                       // picked up by launch_scan / launch_epilogue
    int rc = step_impl(h, stage, stage + B * 32, h->r_mapped_dev, h->r_mapped_dev + B, s, &hf);
    h->arm_seq_next = 0u;
    if (rc) return rc;
    h->armed_seq = seq;
    return NGU_OK;
}

int ngu_epi_step_host(ngu_epi* h, const float* q_host, const float* alpha_host, float* r_int_host,
                      float* r_epi_host) {
    if (!h || !q_host || !r_int_host) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE_KEEP_ARMED(h);
    h->chain_open = false;
    int rc = check_sticky(h);
    if (rc) return rc;
    using clk = std::chrono::steady_clock;
    const clk::time_point t0 = clk::now();
    if (h->user_stream_dirty) {            // (other entry points have already cancelled an armed step)
        rc = order_after_user_stream(h);
        if (rc) return rc;
    }
    const size_t B = h->cfg.n_actors;
    cudaStream_t s = h->own_stream;
    const bool prof = h->host_profile;
    clk::time_point t1, t2, t3;
    const unsigned int seq = ++h->host_seq;
    // one pinned buffer carries the step's controllable states and (optionally) the modulators
    std::memcpy(h->pinned, q_host, B * 32 * sizeof(float));
    if (alpha_host) std::memcpy(h->pinned + B * 32, alpha_host, B * sizeof(float));
    if (prof) t1 = clk::now();

    bool running = false;
    if (h->armed_seq == seq) {
        // The launches of this step are already in the stream, waiting: ring the doorbell.  The stage kernel
        // acknowledges at once (accepted, or cancelled if its window had just run out).
        h->armed_seq = 0u;
        __atomic_store_n(h->doorbell, (seq << 2) | (alpha_host ? 1u : 0u), __ATOMIC_RELEASE);
        volatile unsigned int* st = h->arm_status;
        unsigned int v;
        for (unsigned spins = 1; ((v = *st) >> 1) != seq; ++spins) {
            if ((spins & 0xFFFFu) == 0u) {
                const cudaError_t e = cudaStreamQuery(s);
                if (e != cudaErrorNotReady && e != cudaSuccess) return fail(h, NGU_ECUDA, fmt("step failed: %s", cudaGetErrorString(e)));
            }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        running = (v & 1u) == 0u;
        if (running) h->arm_hits += 1; else { h->arm_misses += 1; h->cur ^= 1; }   // (cancelled: its scan context was never used)
    } else if (h->armed_seq != 0u) {
        disarm(h);                               // (cannot happen: armed steps are always the next sequence number)
    }
    if (!running) {
        float* stage = h->q_stage_dev + (seq & 1u) * h->q_stage_stride;
        if (h->stage_by_kernel && h->use_pdl) {
            // whole 16-byte pieces: both buffers are padded, a few floats past the modulators ride along
            const int n16 = static_cast<int>((B * (alpha_host ? 33 : 32) + 3) / 4);
            stage_inputs_kernel<<<(n16 + 255) / 256, 256, 0, s>>>(reinterpret_cast<float4*>(stage),
                                                                  reinterpret_cast<const float4*>(h->pinned_dev), n16);
            CU_TRY(h, cudaGetLastError());
            h->launches += 1;
        } else {
            CU_TRY(h, cudaMemcpyAsync(stage, h->pinned, B * (alpha_host ? 33 : 32) * sizeof(float), cudaMemcpyHostToDevice, s));
        }
        const HostFlag hf{h->flag_mapped_dev, seq};
        // the rewards come back zero-copy: the epilogue writes them straight into mapped host memory
        rc = step_impl(h, stage, alpha_host ? stage + B * 32 : nullptr, h->r_mapped_dev, h->r_mapped_dev + B, s,
                       h->host_flag_wait ? &hf : nullptr);
        if (rc) return rc;
    }
    // While this step runs: enqueue the next one in pre-armed form - if the caller is in a tight loop (it came back
    // within the doorbell window last time); a sporadic caller never leaves anything waiting on the GPU.
    if (h->arm_enabled && h->use_pdl && h->host_flag_wait && h->have_last_return &&
        std::chrono::duration_cast<std::chrono::nanoseconds>(t0 - h->last_return).count() * 2 < h->arm_window_ns) {
        rc = enqueue_armed(h, seq + 1u);
        if (rc) return rc;
    }
    if (prof) t2 = clk::now();
    if (h->host_flag_wait) {
        rc = wait_host_flag(h, seq, s);
        if (rc) return rc;
    } else {
        CU_TRY(h, cudaStreamSynchronize(s));
    }
    if (prof) t3 = clk::now();
    std::memcpy(r_int_host, h->r_mapped, B * sizeof(float));
    if (r_epi_host) std::memcpy(r_epi_host, h->r_mapped + B, B * sizeof(float));
    rc = check_sticky(h);          // the step that just ran may be the one that timed out: its rewards are NaN
    if (rc) return rc;
    h->last_return = clk::now();
    h->have_last_return = true;
    if (prof) {
        const auto us = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        h->prof_us[0] += us(t0, t1); h->prof_us[1] += us(t1, t2); h->prof_us[2] += us(t2, t3); h->prof_us[3] += us(t3, h->last_return);
        h->prof_calls += 1;
    }
    return NGU_OK;
}

int ngu_epi_reset_host(ngu_epi* h, const uint8_t* done_host) {
    if (!h || !done_host) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    const int B = h->cfg.n_actors;
    bool any = false;
    for (int b = 0; b < B; ++b) any |= done_host[b] != 0;
    if (!any) return NGU_OK;   // nothing to zero: skip the launch (the common step)
    int rc = order_after_user_stream(h);
    if (rc) return rc;
    uint8_t* pd = reinterpret_cast<uint8_t*>(h->pinned + static_cast<size_t>(B) * 33);
    std::memcpy(pd, done_host, B);
    CU_TRY(h, cudaMemcpyAsync(h->done_stage_dev, pd, B, cudaMemcpyHostToDevice, h->own_stream));
    rc = ngu_epi_reset(h, h->done_stage_dev, h->own_stream);
    if (rc) return rc;
    CU_TRY(h, cudaStreamSynchronize(h->own_stream));
    return NGU_OK;
}

int ngu_epi_topk_host(ngu_epi* h, float* d2_host, int32_t* idx_host) {
    if (!h || !d2_host || !idx_host) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    CU_TRY(h, cudaDeviceSynchronize());
    const size_t n = static_cast<size_t>(h->cfg.n_actors) * h->cfg.k;
    CU_TRY(h, cudaMemcpy(d2_host, h->ctx[h->cur].topk_d2, n * sizeof(float), cudaMemcpyDeviceToHost));
    CU_TRY(h, cudaMemcpy(idx_host, h->ctx[h->cur].topk_idx, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return NGU_OK;
}

// Absorb an outstanding reward stash, then read the device state back (synchronous).
static int read_state(ngu_epi* h, DevState* d) {
    CU_TRY(h, cudaDeviceSynchronize());
    h->user_stream_dirty = false;
    log_absorb_kernel<<<1, kLogThreads>>>(h->state, h->log_stash, h->cfg.n_actors, h->p2p_connected ? 1 : 0);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaDeviceSynchronize());
    CU_TRY(h, cudaMemcpy(d, h->state, sizeof *d, cudaMemcpyDeviceToHost));
    return NGU_OK;
}

int ngu_epi_get_state(ngu_epi* h, ngu_epi_state* out) {
    if (!h || !out) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    DevState d;
    const int rc = read_state(h, &d);   // brings the log-only trackers up to date first (p2p mode: into pending_log)
    if (rc) return rc;
    std::memcpy(out->ed_mean, d.ed_mean, sizeof d.ed_mean);
    std::memcpy(out->ed_var, d.ed_var, sizeof d.ed_var);
    out->ed_count = d.ed_count;
    out->epi_mean = d.epi_mean; out->epi_var = d.epi_var; out->epi_count = d.epi_count;
    out->intr_mean = d.intr_mean; out->intr_var = d.intr_var; out->intr_count = d.intr_count;
    out->cursor = d.cursor; out->steps = d.steps; out->exchange_timeouts = d.exchange_timeouts;
    out->ll_mean = d.ll_mean; out->ll_var = d.ll_var; out->ll_count = d.ll_count;
    return NGU_OK;
}

int ngu_epi_set_state(ngu_epi* h, const ngu_epi_state* in) {
    if (!h || !in) return fail(h, NGU_EINVAL, "null argument");
    if (in->cursor < 0 || in->cursor >= h->cfg.capacity) return fail(h, NGU_EINVAL, "cursor out of range");
    BIND_DEVICE(h);
    // Only the PUBLIC fields are replaced.  The library's own bookkeeping survives: the log-only sums still waiting
    // for an exchange round (pending_log), the exchange / flush sequence numbers (tags must stay in step with
    // the peers) and the sticky timeout counter (the only evidence of a failed merge or exchange).
    DevState d;
    int rc = read_state(h, &d);
    if (rc) return rc;
    std::memcpy(d.ed_mean, in->ed_mean, sizeof d.ed_mean);
    std::memcpy(d.ed_var, in->ed_var, sizeof d.ed_var);
    d.ed_count = in->ed_count;
    d.epi_mean = in->epi_mean; d.epi_var = in->epi_var; d.epi_count = in->epi_count;
    d.intr_mean = in->intr_mean; d.intr_var = in->intr_var; d.intr_count = in->intr_count;
    d.cursor = in->cursor; d.steps = in->steps;
    d.ll_mean = in->ll_mean; d.ll_var = in->ll_var; d.ll_count = in->ll_count;
    CU_TRY(h, cudaMemcpy(h->state, &d, sizeof d, cudaMemcpyHostToDevice));
    return NGU_OK;
}

int ngu_epi_set_cursor(ngu_epi* h, int64_t cursor) {
    if (!h) return fail(h, NGU_EINVAL, "null argument");
    if (cursor < 0 || cursor >= h->cfg.capacity) return fail(h, NGU_EINVAL, "cursor out of range");
    BIND_DEVICE(h);
    CU_TRY(h, cudaDeviceSynchronize());
    h->user_stream_dirty = false;
    const long long c = cursor;
    CU_TRY(h, cudaMemcpy(&h->state->cursor, &c, sizeof c, cudaMemcpyHostToDevice));
    return NGU_OK;
}

static int transpose_copy(ngu_epi* h, const float* host_in, float* host_out) {
    BIND_DEVICE(h);
    const int B = h->cfg.n_actors, N = h->cfg.capacity;
    const size_t bytes = static_cast<size_t>(B) * N * 32 * sizeof(float);
    float* tmp = nullptr;
    CU_TRY(h, cudaDeviceSynchronize());
    CU_TRY(h, cudaMalloc(&tmp, bytes));
    cudaError_t e = cudaSuccess;
    if (host_in) {   // [N][B][32] host -> [B][N][32] device
        e = cudaMemcpy(tmp, host_in, bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            transpose_rows_kernel<<<1024, 256>>>(reinterpret_cast<float4*>(h->memory), reinterpret_cast<const float4*>(tmp), N, B);
            e = cudaDeviceSynchronize();
        }
    } else {         // [B][N][32] device -> [N][B][32] host
        transpose_rows_kernel<<<1024, 256>>>(reinterpret_cast<float4*>(tmp), reinterpret_cast<const float4*>(h->memory), B, N);
        e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaMemcpy(host_out, tmp, bytes, cudaMemcpyDeviceToHost);
    }
    cudaFree(tmp);
    h->launches += 1;
    if (e != cudaSuccess) return fail(h, NGU_ECUDA, fmt("memory load/dump: %s", cudaGetErrorString(e)));
    return NGU_OK;
}

int ngu_epi_load_memory_host(ngu_epi* h, const float* slot_major_host) {
    if (!h || !slot_major_host) return fail(h, NGU_EINVAL, "null argument");
    return transpose_copy(h, slot_major_host, nullptr);
}
int ngu_epi_dump_memory_host(ngu_epi* h, float* slot_major_host) {
    if (!h || !slot_major_host) return fail(h, NGU_EINVAL, "null argument");
    return transpose_copy(h, nullptr, slot_major_host);
}

int ngu_epi_p2p_export(ngu_epi* h, void* ipc_handle_out, int32_t world) {
    if (!h || !ipc_handle_out) return fail(h, NGU_EINVAL, "null argument");
    if (world < 2 || world > kMaxWorld) return fail(h, NGU_EINVAL, fmt("world must be 2..%d", kMaxWorld));
    BIND_DEVICE(h);
    if (!h->inbox) {
        const size_t bytes = static_cast<size_t>(2) * world * kSlotWords * sizeof(unsigned long long);
        CU_TRY(h, cudaMalloc(&h->inbox, bytes));
        CU_TRY(h, cudaMemset(h->inbox, 0, bytes));
        CU_TRY(h, cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t hd;
    CU_TRY(h, cudaIpcGetMemHandle(&hd, h->inbox));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(ipc_handle_out, &hd, sizeof hd);
    return NGU_OK;
}

int ngu_epi_p2p_connect(ngu_epi* h, int32_t rank, int32_t world, const void* ipc_handles) {
    if (!h || !ipc_handles) return fail(h, NGU_EINVAL, "null argument");
    if (!h->inbox) return fail(h, NGU_ESTATE, "call ngu_epi_p2p_export first");
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world) return fail(h, NGU_EINVAL, "bad rank/world");
    BIND_DEVICE(h);
    h->p2p = P2PView{};
    h->p2p.inbox = h->inbox; h->p2p.rank = rank; h->p2p.world = world;
    for (int g = 0; g < world; ++g) {
        if (g == rank) { h->p2p.peer[g] = h->inbox; continue; }
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, static_cast<const char*>(ipc_handles) + 64 * g, sizeof hd);
        void* ptr = nullptr;
        CU_TRY(h, cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
        h->p2p.peer[g] = static_cast<unsigned long long*>(ptr);
    }
    h->p2p_connected = true;
    return NGU_OK;
}

int ngu_epi_log_flush(ngu_epi* h, void* stream) {
    if (!h) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    if (!h->p2p_connected) {                  // only the reward stash of the last fused step is outstanding
        log_absorb_kernel<<<1, kLogThreads, 0, static_cast<cudaStream_t>(stream)>>>(h->state, h->log_stash, h->cfg.n_actors, 0);
    } else {
        log_flush_kernel<<<1, kLogThreads, 0, static_cast<cudaStream_t>(stream)>>>(h->state, h->p2p, h->log_stash,
                                                                                  h->cfg.n_actors);
    }
    CU_TRY(h, cudaGetLastError());
    h->launches += 1;
    return NGU_OK;
}

int ngu_epi_debug_timeline(ngu_epi* h, uint64_t* out_host, int32_t max_warps) {
    if (!h || !out_host) return fail(h, NGU_EINVAL, "null argument");
    if (max_warps < 0 && max_warps >= -16) {   // the epilogue's phase stamps instead (NGU_EPI_DEBUG_EPILOGUE=1)
        const int n = -max_warps;
        if (!h->epi_stamps) return fail(h, NGU_ESTATE, "create the handle with NGU_EPI_DEBUG_EPILOGUE=1");
        if (cudaDeviceSynchronize() != cudaSuccess ||
            cudaMemcpy(out_host, h->epi_stamps, n * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(h, NGU_ECUDA, "copying the epilogue stamps failed");
        return n;
    }
    if (!h->timeline) return fail(h, NGU_ESTATE, "create the handle with NGU_EPI_DEBUG_TIMELINE=1");
    BIND_DEVICE(h);
    int warps = h->grid * h->var.warps;
    if (warps > h->sched.n_chunks) warps = h->sched.n_chunks;
    const int n = warps < max_warps ? warps : max_warps;
    CU_TRY(h, cudaDeviceSynchronize());
    CU_TRY(h, cudaMemcpy(out_host, h->timeline, static_cast<size_t>(n) * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return n;
}

int ngu_epi_plan_host(const ngu_epi_cfg* cfg, int32_t sm_count, int32_t geo[8], int32_t* chunks, int32_t max_chunks) {
    if (!cfg || !geo || sm_count < 1) return NGU_EINVAL;
    if (cfg->n_actors < 1 || cfg->capacity < 1 || cfg->k < 1 || cfg->k > NGU_EPI_MAX_K || cfg->scan_variant < 0 ||
        cfg->scan_variant >= kNumVariants || static_cast<int64_t>(cfg->n_actors) * cfg->capacity + 256 >= (1ll << 31))
        return NGU_EINVAL;
    ngu_epi h;                        // host-side planning only: no CUDA call is made
    h.cfg = *cfg;
    h.sm_count = sm_count;
    h.var = kVariants[cfg->scan_variant];
    if (cfg->scan_ctas_per_sm > 0) h.var.ctas_per_sm = cfg->scan_ctas_per_sm;
    apply_sched_env(&h);
    plan_geometry(&h);
    ngu_epi_scan_geometry(&h, geo);
    if (chunks) {
        if (max_chunks < h.sched.n_chunks) return NGU_EINVAL;
        for (int i = 0; i < h.sched.n_chunks; ++i) {
            int t = 0, te = 0;
            chunk_range(h.sched, i, t, te);
            chunks[2 * i] = t; chunks[2 * i + 1] = te;
        }
    }
    return NGU_OK;
}

int ngu_epi_debug_chunk(const ngu_epi* h, int32_t idx, int32_t out[2]) {
    if (!h || !out) return NGU_EINVAL;
    if (idx < 0 || idx >= h->sched.n_chunks) return NGU_EINVAL;
    int t = 0, te = 0;
    chunk_range(h->sched, idx, t, te);
    out[0] = t; out[1] = te;
    return NGU_OK;
}

int ngu_epi_debug_inject_timeout(ngu_epi* h) {
    if (!h) return NGU_EINVAL;
    BIND_DEVICE(h);
    CU_TRY(h, cudaDeviceSynchronize());
    const long long one = 1;
    CU_TRY(h, cudaMemcpy(&h->state->exchange_timeouts, &one, sizeof one, cudaMemcpyHostToDevice));
    return NGU_OK;
}

int64_t ngu_epi_launch_count(const ngu_epi* h) { return h ? h->launches : 0; }

static void drop_timing(ngu_epi* h) {
    for (cudaEvent_t e : h->t_beg) cudaEventDestroy(e);
    for (cudaEvent_t e : h->t_end) cudaEventDestroy(e);
    h->t_beg.clear(); h->t_end.clear(); h->t_used = 0;
}

int ngu_epi_scan_timing_begin(ngu_epi* h, int32_t max_calls) {
    if (!h || max_calls < 1) return fail(h, NGU_EINVAL, "bad argument");
    BIND_DEVICE(h);
    drop_timing(h);
    for (int i = 0; i < max_calls; ++i) {
        cudaEvent_t a, b;
        CU_TRY(h, cudaEventCreate(&a));
        CU_TRY(h, cudaEventCreate(&b));
        h->t_beg.push_back(a); h->t_end.push_back(b);
    }
    return NGU_OK;
}

int ngu_epi_scan_timing_end(ngu_epi* h, double* total_ms, int32_t* calls) {
    if (!h || !total_ms || !calls) return fail(h, NGU_EINVAL, "null argument");
    BIND_DEVICE(h);
    CU_TRY(h, cudaDeviceSynchronize());
    double tot = 0.0;
    for (size_t i = 0; i < h->t_used; ++i) {
        float ms = 0.f;
        CU_TRY(h, cudaEventElapsedTime(&ms, h->t_beg[i], h->t_end[i]));
        tot += ms;
    }
    *total_ms = tot;
    *calls = static_cast<int32_t>(h->t_used);
    drop_timing(h);
    return NGU_OK;
}

int ngu_epi_scan_geometry(const ngu_epi* h, int32_t out[8]) {
    if (!h || !out) return NGU_EINVAL;
    out[0] = h->grid; out[1] = h->var.warps * 32; out[2] = h->var.smem;
    out[3] = h->sched.s1; out[4] = h->var.stages; out[5] = 32 * h->var.rpl;
    out[6] = h->sched.n_chunks > h->sched.n1 ? h->sched.c2 : 0; out[7] = h->sched.n_chunks;
    return NGU_OK;
}

}  // extern "C"
