// porechop_b200/csrc/engine.cu -- host engine + C-ABI (include/porechop_b200.h) of cpp_functions.so.
//
// Replaces, behind the same ctypes boundary, the reference's C shim + SeqAn DP + ScoredAlignment
// (porechop/src/adapter_align.cpp:11-44, porechop/src/alignment.cpp:6-121, seqan/align/dp_*.h).
// The host side only plans and pipelines: chunking, class selection by adapter length, host<->device
// copies on a ring of streams, kernel launches.  All alignment arithmetic runs in the sm_100a kernels of
// kernels.cuh; there is no CPU alignment path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/porechop_b200.h"
#include "kernels.cuh"

// NVTX ranges around submit / H2D / DP / D2H (SURVEY 5 "tracing"): header-only NVTX3, a no-op unless a profiler is attached.
// The host simulator (tests/sim) and -DPB_NO_NVTX build without it.
#if defined(__CUDACC__) && !defined(PB_NO_NVTX)
#include <nvtx3/nvToolsExt.h>
#define PB_NVTX 1
#endif

using namespace pb;

// hostpack.cpp (g++): ASCII -> two 4-bit Dna5 codes per byte on the host cores
extern "C" void pb_pack_nibbles(const uint8_t *in, int64_t n, uint8_t *out, int threads);

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
std::atomic<int> g_timing{0};

int fail(int code, const std::string &msg) { g_err = msg; return code; }

struct NvtxRange {          // RAII range on the calling host thread
#ifdef PB_NVTX
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
#else
    explicit NvtxRange(const char *) {}
#endif
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(PB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(PB200_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// pinned host staging memory (packed upload path)
struct HostBuf {
    uint8_t *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&p), want, cudaHostAllocDefault);
        if (e != cudaSuccess) { p = nullptr; return fail(PB200_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
};

struct Options {
    int64_t direct_max = 160;      // longest sequence aligned in a single (trace) pass.  Round 2 on B200: 150-column windows are 6 % faster in one
                                   // pass, 500-column reads 2.9x faster in two (the one-pass trace of a long window is 10.5 instead of 3.7
                                   // instructions per cell and its scratch, 129 KB per warp at 500 columns, leaves 2 blocks per SM)
    int64_t chunk_tasks = 1 << 17; // alignments per pipeline chunk (host-buffer API)
    int64_t device_chunk_tasks = 8 << 20; // alignments per launch group (device-resident API)
    int64_t chunk_bytes = 64ll << 20;   // sequence bytes per pipeline chunk (host-buffer API)
    int scratch_mb = 128;          // cap on the resident trace scratch (MB); 72 keeps it L2-resident at ~10% lower speed
    int profile = 1;               // 1: score pass fetches the substitution operands from a shared-memory query profile (same-read slots)
    int tight_window = 1;          // 1: second-pass windows sized per alignment from the end cell's row and score (window_cols)
    int h2d_pack = 0;              // 1: host-buffer API converts to 4-bit codes on the host cores and uploads half the bytes; 0 (default):
                                   // never; -1 = auto: for submits of >= 32 MB when the packer team has >= 12 threads (pack_wanted)
    int pack_threads = 0;          // host threads of the packer (default: hardware threads / ranks on the node, at most 32)
    int hbuf_mode = 0;             // 0 auto, 1 shared memory, 2 global scratch (staging of a slot's packed bases)
};
Options g_opt;
std::once_flag g_opt_once;
void load_env_options() {
    std::call_once(g_opt_once, [] {
        if (const char *v = getenv("PB200_DIRECT_MAX")) g_opt.direct_max = atoll(v);
        if (const char *v = getenv("PB200_CHUNK_TASKS")) g_opt.chunk_tasks = std::max(1ll, atoll(v));
        if (const char *v = getenv("PB200_SCRATCH_MB")) g_opt.scratch_mb = std::max(1, atoi(v));
        if (const char *v = getenv("PB200_TIGHT_WINDOW")) g_opt.tight_window = atoi(v);
        if (const char *v = getenv("PB200_H2D_PACK")) g_opt.h2d_pack = atoi(v);
        if (const char *v = getenv("PB200_PROFILE")) g_opt.profile = atoi(v);
        if (const char *v = getenv("PB200_PACK_THREADS")) g_opt.pack_threads = atoi(v);
        if (g_opt.pack_threads <= 0) {
            // packer threads: the host's hardware threads shared by the ranks of this node (torchrun exports
            // LOCAL_WORLD_SIZE, and OMP_NUM_THREADS=1 -- which would otherwise leave the packer single-threaded), at most 32
            // -- and no more than the cgroup CPU quota allows to run at once (round 2: the GPU boxes show 128 hardware threads
            // under a quota of 16 CPUs; a team larger than the quota is only throttled)
            unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
                char q[64]; long long period = 0;
                if (fscanf(f, "%63s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) {
                    const long long quota = atoll(q);
                    if (quota > 0) hw = (unsigned)std::min<long long>(hw, std::max<long long>(1, (quota + period - 1) / period));
                }
                fclose(f);
            }
            const int lw = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
            unsigned team = std::max<unsigned>(1u, hw / (unsigned)lw);
            if (team > 4) team -= 2;         // the submit thread and the driver's threads need CPUs of the same quota
            g_opt.pack_threads = (int)std::min<unsigned>(32u, team);
        }
        if (const char *v = getenv("PB200_HBUF")) g_opt.hbuf_mode = !strcmp(v, "smem") ? 1 : !strcmp(v, "global") ? 2 : 0;
    });
}

constexpr int NSTAGE = 3;
struct Stage {
    cudaStream_t stream = nullptr;
    DevBuf seq_raw, seq_codes, seq_off, tasks, tasks2, ends, out, order, bins, pair_seq, pair_ad, gtrace, misc;
    DevBuf dec_trim, dec_pairs, dec_top2;    // per-chunk outputs of decide_kernel
};

struct ClassPlan {
    int cls = 0;                      // row-capacity class (class_of); GENERIC_CLASS = int32 fallback
    std::vector<int32_t> ad_ids;      // adapters of the class, sorted by length
    int m_max = 0;
};

// timed DP launches by kind (pb200TimingReadKinds): which kernel dominates a step, and its per-launch duration
enum { TK_TRACE = 0, TK_TRACE_SCORE_ONLY = 1, TK_TRACE_WINDOW = 2, TK_SCORE = 3, TK_N = 4 };
struct TimedLaunch { cudaEvent_t a, b; int kind; };

struct Engine {
    int device = -1;
    int sm_count = 0;
    size_t smem_optin = 0;
    std::mutex mu;
    Stage st[NSTAGE];
    DevBuf gjobs, gscratch;
    static constexpr int MAX_DEC_JOBS = 8;
    DevBuf dec_cmin[MAX_DEC_JOBS], dec_cols[MAX_DEC_JOBS];   // threshold table + score columns of the decision jobs of a call
    // adapter plan cache (4 entries, LRU): repeated calls with the same adapters + scoring (the normal case: Porechop
    // alternates between its start-adapter and end-adapter lists) skip upload, encode and the host synchronisation
    struct PlanEntry {
        std::vector<uint8_t> ad;
        std::vector<int32_t> off;
        int sc[4] = {0, 0, 0, 0};
        bool valid = false;
        unsigned long long last_used = 0;
        DevBuf ad_raw, ad_codes, ad_off, cls_ad;
        std::shared_ptr<void> plan;
    };
    PlanEntry plans[4];
    unsigned long long plan_clock = 0;
    std::vector<TimedLaunch> timed;
    double timed_ms_acc[TK_N] = {0, 0, 0, 0};
    long long timed_n_acc[TK_N] = {0, 0, 0, 0};
    int next_trace_kind = TK_TRACE;      // set to TK_TRACE_WINDOW around the second-pass launch of a two-pass class
    DevBuf wcells;                       // device counter: DP cells of the windowed second passes (window_tasks_kernel)
    std::shared_ptr<void> packer;        // Packer (below): the thread that plans and packs chunks ahead of the submit loop
    // launch constants, queried once per (kernel, dynamic shared-memory size); all users hold `mu`.  The opt-in shared-memory
    // limit is a property of the KERNEL, not of a launch: it is only ever raised (a launch with less is always valid).
    std::map<std::pair<const void *, size_t>, int> bps_cache;
    std::map<const void *, size_t> smem_attr;
    int blocks_per_sm(const void *kern, int threads, size_t smem_bytes, int *bps) {
        const auto key = std::make_pair(kern, smem_bytes);
        auto it = bps_cache.find(key);
        if (it != bps_cache.end()) { *bps = it->second; return 0; }
        size_t &have = smem_attr[kern];
        if (smem_bytes > have) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            have = smem_bytes;
        }
        int b = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, threads, smem_bytes));
        if (b < 1) b = 1;
        bps_cache[key] = b;
        *bps = b;
        return 0;
    }
    bool init_done = false;
    int init() {
        if (init_done) return 0;
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            return fail(PB200_ERR_NO_DEVICE, "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                                 ", this library is built for sm_100a only");
        sm_count = prop.multiProcessorCount;
        smem_optin = prop.sharedMemPerBlockOptin;
        for (int i = 0; i < NSTAGE; ++i) CK(cudaStreamCreateWithFlags(&st[i].stream, cudaStreamNonBlocking));
        init_done = true;
        return 0;
    }
};

std::mutex g_engines_mu;
std::map<int, std::unique_ptr<Engine>> g_engines;

int get_engine(Engine **out) {
    load_env_options();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(PB200_ERR_NO_DEVICE, std::string("no CUDA device available: ") +
                                             (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_engines_mu);
    auto &slot = g_engines[dev];
    if (!slot) { slot.reset(new Engine()); slot->device = dev; }
    *out = slot.get();
    return 0;
}

// ---- int16 domain checks and the window bound ------------------------------------------------------------
struct SchemeInfo {
    bool int16_ok_base;   // sign conventions allow the packed kernels at all
    int A;                // max |score|
    bool bounded;         // window bound exists (both gap scores negative)
    int wnum, wden;       // W(m) = m + m*wnum/wden
};
SchemeInfo scheme_info(int ma, int mi, int go, int ge) {
    SchemeInfo s;
    auto ab = [](int x) { return x < 0 ? -(long long)x : (long long)x; };
    long long A = std::max(std::max(ab(ma), ab(mi)), std::max(ab(go), ab(ge)));
    s.A = (int)std::min<long long>(A, 1 << 30);
    s.int16_ok_base = (ma >= mi) && ((long long)ma - mi <= PB_MAX_SUBW) && go <= 0 && ge <= 0 && A <= PB_I16_LIMIT;
    s.bounded = (go < 0 && ge < 0);
    s.wnum = std::max(std::max(ma, mi), 0);
    s.wden = (int)std::min(ab(go), ab(ge));
    if (s.wden == 0) s.wden = 1;
    return s;
}
bool int16_ok(const SchemeInfo &s, int m) { return s.int16_ok_base && (long long)s.A * (m + 3) <= PB_I16_LIMIT; }
// Row capacity classes: G lanes x R rows per lane.  Class k: G = 4 << (k / 4), R = 5 + (k % 4)  -> capacities
// 20,24,28,32, 40,48,56,64, 80,96,112,128, 160,192,224,256.  Class 16 = generic int32 fallback.
constexpr int N_CLASSES = 17;
constexpr int GENERIC_CLASS = 16;
int class_G(int k) { return 4 << (k / 4); }
int class_R(int k) { return 5 + (k % 4); }
int class_of(const SchemeInfo &s, int m) {
    if (!int16_ok(s, m) || m > 256) return GENERIC_CLASS;
    for (int k = 0; k < GENERIC_CLASS; ++k)
        if (m <= class_G(k) * class_R(k)) return k;
    return GENERIC_CLASS;
}

// ---- kernel launch helpers -----------------------------------------------------------------------------
void timed_begin(Engine &E, cudaStream_t s, TimedLaunch &tl, bool &on, int kind) {
    on = g_timing.load() != 0;
    if (on) { tl.kind = kind; cudaEventCreate(&tl.a); cudaEventCreate(&tl.b); cudaEventRecord(tl.a, s); }
}
void timed_end(Engine &E, cudaStream_t s, TimedLaunch &tl, bool on) {
    if (on) { cudaEventRecord(tl.b, s); E.timed.push_back(tl); }
}

template <int G, int R, bool HS>
int launch_trace_variant(Engine &E, Stage &S, cudaStream_t stream, const TaskSrc &ts, int max_n,
                         const uint8_t *seq_codes, const uint8_t *ad_codes, const Scoring &sc, int32_t *out, int *status) {
    constexpr int SPW = 32 / G;
    constexpr int WPS = TraceWords<R>::value;
    const int max_steps = max_n + G - 1;
    const int wpb = PB_WARPS_PER_BLOCK;
    const size_t hb_words = HS ? (size_t)SPW * max_n : 0;
    const size_t smem_bytes = (size_t)wpb * (hb_words + PB_SCRATCH_WORDS) * 4;
    auto kern = trace_kernel<G, R, HS>;
    int bps = 0;
    if (int rc = E.blocks_per_sm(reinterpret_cast<const void *>(kern), wpb * 32, smem_bytes, &bps)) return rc;
    const size_t gwarp_bytes = ((size_t)((max_steps + PB_TCHUNK - 1) / PB_TCHUNK) * PB_TCHUNK * WPS * 32 +
                                (HS ? 0 : (((size_t)SPW * max_n + 31) & ~(size_t)31))) * 4;   // 128-byte lines per warp (kernels.cuh)
    // The trace scratch of the resident grid is rewritten slot after slot and mostly lives in L2.  `scratch_mb` caps
    // it (whole blocks per SM, never below 2): measured on B200 for 150x28 windows, 72 MB (3 blocks/SM) keeps the
    // trace L2-resident (DRAM traffic 1.9x the algorithmic bytes) at ~10 % lower kernel throughput than the default
    // 128 MB (5 blocks/SM), where ~1.1 GB of dead/dirty trace lines per launch are written back to HBM off the critical
    // path (8 % of the HBM bandwidth).  See DESIGN.md, "trace scratch".
    int64_t max_blocks = std::max<int64_t>(2 * E.sm_count, (int64_t)(((size_t)g_opt.scratch_mb << 20) / std::max<size_t>(gwarp_bytes * wpb, 1)));
    max_blocks -= max_blocks % E.sm_count;     // whole blocks per SM: the grid-stride loop gives every block the same work
    const int64_t n_tasks = ts.n_tasks;
    const int64_t n_slots = (n_tasks + 1) / 2;
    const int64_t n_wslots = (n_slots + SPW - 1) / SPW;
    int64_t blocks = (n_wslots + wpb - 1) / wpb;
    blocks = std::min<int64_t>(blocks, std::min<int64_t>((int64_t)bps * E.sm_count, max_blocks));
    if ((size_t)blocks * wpb * gwarp_bytes > (24ull << 30)) blocks = std::max<int64_t>(1, (int64_t)((24ull << 30) / (gwarp_bytes * wpb)));
    if (blocks <= 0) return 0;
    if (int rc = S.gtrace.ensure((size_t)blocks * wpb * gwarp_bytes)) return rc;
    TimedLaunch tl; bool on;
    timed_begin(E, stream, tl, on, E.next_trace_kind);
    kern<<<(unsigned)blocks, wpb * 32, smem_bytes, stream>>>(ts, seq_codes, ad_codes, sc, out, S.gtrace.as<uint32_t>(), max_steps,
                                                              max_n, status);
    timed_end(E, stream, tl, on);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int G, int R>
int launch_trace(Engine &E, Stage &S, cudaStream_t stream, const TaskSrc &ts, int max_n,
                 const uint8_t *seq_codes, const uint8_t *ad_codes, const Scoring &sc, int32_t *out, int *status) {
    constexpr int SPW = 32 / G;
    if (max_n < 1) max_n = 1;
    // packed read bases of a slot are staged in shared memory when they fit (<= 12 KB per warp), else in global scratch
    const bool hs = g_opt.hbuf_mode == 1 ? true : g_opt.hbuf_mode == 2 ? false : ((size_t)SPW * max_n * 4 <= 12288);
    if (hs && (size_t)PB_WARPS_PER_BLOCK * ((size_t)SPW * max_n + 4 + PB_SCRATCH_WORDS) * 4 <= E.smem_optin)
        return launch_trace_variant<G, R, true>(E, S, stream, ts, max_n, seq_codes, ad_codes, sc, out, status);
    return launch_trace_variant<G, R, false>(E, S, stream, ts, max_n, seq_codes, ad_codes, sc, out, status);
}

int launch_trace_class(Engine &E, Stage &S, cudaStream_t stream, int cls, const TaskSrc &ts, int max_n,
                       const uint8_t *seq_codes, const uint8_t *ad_codes, const Scoring &sc, int32_t *out, int *status) {
#define PB_CASE(K, GG, RR) case K: return launch_trace<GG, RR>(E, S, stream, ts, max_n, seq_codes, ad_codes, sc, out, status);
    switch (cls) {
        PB_CASE(0, 4, 5) PB_CASE(1, 4, 6) PB_CASE(2, 4, 7) PB_CASE(3, 4, 8)
        PB_CASE(4, 8, 5) PB_CASE(5, 8, 6) PB_CASE(6, 8, 7) PB_CASE(7, 8, 8)
        PB_CASE(8, 16, 5) PB_CASE(9, 16, 6) PB_CASE(10, 16, 7) PB_CASE(11, 16, 8)
        PB_CASE(12, 32, 5) PB_CASE(13, 32, 6) PB_CASE(14, 32, 7) PB_CASE(15, 32, 8)
    }
#undef PB_CASE
    return fail(PB200_ERR_INTERNAL, "bad class");
}

template <int G, int R, bool PROF>
int launch_score_variant(Engine &E, cudaStream_t stream, const TaskSrc &ts, unsigned long long *counter,
                         const uint8_t *seq_codes, const uint8_t *ad_codes, const Scoring &sc, EndCell *ends) {
    auto kern = score_kernel<G, R, PROF>;
    int bps = 0;
    if (int rc = E.blocks_per_sm(reinterpret_cast<const void *>(kern), PB_WARPS_PER_BLOCK * 32, 0, &bps)) return rc;
    constexpr int SPW = 32 / G;
    const int64_t n_tasks = ts.n_tasks;
    const int64_t n_slots = (n_tasks + 1) / 2;
    int64_t blocks = (n_slots + SPW * PB_WARPS_PER_BLOCK - 1) / (SPW * PB_WARPS_PER_BLOCK);
    blocks = std::min<int64_t>(blocks, (int64_t)bps * E.sm_count);
    if (blocks <= 0) return 0;
    CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    TimedLaunch tl; bool on;
    timed_begin(E, stream, tl, on, TK_SCORE);
    kern<<<(unsigned)blocks, PB_WARPS_PER_BLOCK * 32, 0, stream>>>(ts, counter, seq_codes, ad_codes, sc, ends);
    timed_end(E, stream, tl, on);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}
template <int G, int R>
int launch_score(Engine &E, cudaStream_t stream, const TaskSrc &ts, unsigned long long *counter, const uint8_t *seq_codes,
                 const uint8_t *ad_codes, const Scoring &sc, EndCell *ends) {
    // query profile: every slot is (one read, two adapters) -- cross mode with an even number of adapters in the class
    const bool prof = g_opt.profile != 0 && ts.tasks == nullptr && ts.n_cls_ad > 0 && (ts.n_cls_ad % 2) == 0;
    if (prof) return launch_score_variant<G, R, true>(E, stream, ts, counter, seq_codes, ad_codes, sc, ends);
    return launch_score_variant<G, R, false>(E, stream, ts, counter, seq_codes, ad_codes, sc, ends);
}
int launch_score_class(Engine &E, cudaStream_t stream, int cls, const TaskSrc &ts,
                       unsigned long long *counter, const uint8_t *seq_codes, const uint8_t *ad_codes, const Scoring &sc,
                       EndCell *ends) {
    switch (cls / 4) {
        case 0: return launch_score<4, 8>(E, stream, ts, counter, seq_codes, ad_codes, sc, ends);
        case 1: return launch_score<8, 8>(E, stream, ts, counter, seq_codes, ad_codes, sc, ends);
        case 2: return launch_score<16, 8>(E, stream, ts, counter, seq_codes, ad_codes, sc, ends);
        case 3: return launch_score<32, 8>(E, stream, ts, counter, seq_codes, ad_codes, sc, ends);
    }
    return fail(PB200_ERR_INTERNAL, "bad class");
}

int launch_unpack(cudaStream_t stream, const uint8_t *in, uint8_t *out, int64_t n, int sm_count) {
    if (n <= 0) return 0;
    int64_t blocks = std::min<int64_t>((n + 256 * 16 - 1) / (256 * 16), (int64_t)sm_count * 16);
    unpack_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, n);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

int launch_encode(cudaStream_t stream, const uint8_t *in, uint8_t *out, int64_t n, int sm_count) {
    if (n <= 0) return 0;
    int64_t blocks = std::min<int64_t>((n + 256 * 16 - 1) / (256 * 16), (int64_t)sm_count * 16);
    encode_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, n);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

// Run every task of one class; `ts` describes the tasks in slot order (explicit records or the cross product).
int run_class_tasks(Engine &E, Stage &S, cudaStream_t stream, int cls, int m_max, const TaskSrc &ts, int64_t max_n,
                    const uint8_t *seq_codes, const uint8_t *ad_codes, const Scoring &sc, const SchemeInfo &si,
                    int32_t *out, int *status, unsigned long long *counter) {
    const int64_t n_tasks = ts.n_tasks;
    if (n_tasks <= 0) return 0;
    if (g_opt.profile != 0 && ts.tasks == nullptr && ts.cls_ad != nullptr && ts.n_cls_ad >= 3 && (ts.n_cls_ad & 1) &&
        si.bounded && max_n > g_opt.direct_max) {
        // the score pass's query profile needs same-read slots: the paired adapters (one read, two adapters per slot) and the
        // odd last adapter (two reads per slot) run as two launch sequences; cross_task() indexes both exactly as it does
        // inside the whole class
        TaskSrc paired = ts, tail = ts;
        paired.n_cls_ad = ts.n_cls_ad - 1;
        paired.n_tasks = ts.n_seqs * (int64_t)paired.n_cls_ad;
        tail.cls_ad = ts.cls_ad + (ts.n_cls_ad - 1);
        tail.n_cls_ad = 1;
        tail.n_tasks = ts.n_seqs;
        if (int rc = run_class_tasks(E, S, stream, cls, m_max, paired, max_n, seq_codes, ad_codes, sc, si, out, status, counter)) return rc;
        return run_class_tasks(E, S, stream, cls, m_max, tail, max_n, seq_codes, ad_codes, sc, si, out, status, counter);
    }
    int64_t W = si.bounded ? (int64_t)m_max + ((int64_t)m_max * si.wnum) / si.wden : (int64_t)1 << 40;
    const bool two_pass = si.bounded && max_n > g_opt.direct_max && W + 1 < max_n;
    if (!two_pass) return launch_trace_class(E, S, stream, cls, ts, (int)max_n, seq_codes, ad_codes, sc, out, status);
    if (int rc = S.ends.ensure((size_t)n_tasks * sizeof(EndCell))) return rc;
    if (int rc = S.tasks2.ensure((size_t)n_tasks * sizeof(Task))) return rc;
    if (int rc = launch_score_class(E, stream, cls, ts, counter, seq_codes, ad_codes, sc, S.ends.as<EndCell>())) return rc;
    {
        int64_t blocks = (n_tasks + 255) / 256;
        if (!E.wcells.p) {
            if (int rc = E.wcells.ensure(8)) return rc;
            CK(cudaMemsetAsync(E.wcells.p, 0, 8, stream));
        }
        window_tasks_kernel<<<(unsigned)blocks, 256, 0, stream>>>(ts, S.ends.as<EndCell>(), S.tasks2.as<Task>(), si.wnum, si.wden, g_opt.tight_window,
                                                                   E.wcells.as<unsigned long long>());
        g_launches++;
        CK(cudaGetLastError());
    }
    TaskSrc t2 = ts;
    t2.tasks = S.tasks2.as<Task>();
    E.next_trace_kind = TK_TRACE_WINDOW;
    const int rc2 = launch_trace_class(E, S, stream, cls, t2, (int)W, seq_codes, ad_codes, sc, out, status);
    E.next_trace_kind = TK_TRACE;
    return rc2;
}

// Adapter-side planning shared by all entry points.
struct AdapterPlan {
    std::vector<ClassPlan> classes;   // non-empty classes only
    SchemeInfo si;
    Scoring sc;
    const uint8_t *d_ad_codes = nullptr;   // device copies owned by the engine's plan cache
    const int32_t *d_ad_off = nullptr;
    const int32_t *d_cls_ad = nullptr;
};
int plan_adapters(Engine &E, cudaStream_t stream, const uint8_t *adapters, const int32_t *ad_off, int32_t n_adapters,
                  int ma, int mi, int go, int ge, AdapterPlan &P) {
    const size_t nb = (size_t)ad_off[n_adapters];
    Engine::PlanEntry *slot = nullptr;
    for (auto &e : E.plans) {
        if (e.valid && e.plan && e.sc[0] == ma && e.sc[1] == mi && e.sc[2] == go && e.sc[3] == ge &&
            e.off.size() == (size_t)n_adapters + 1 && e.ad.size() == nb &&
            memcmp(e.off.data(), ad_off, e.off.size() * 4) == 0 && (nb == 0 || memcmp(e.ad.data(), adapters, nb) == 0)) {
            e.last_used = ++E.plan_clock;
            P = *static_cast<AdapterPlan *>(e.plan.get());
            return 0;
        }
    }
    for (auto &e : E.plans) if (!slot || !e.valid || (slot->valid && e.last_used < slot->last_used)) { slot = &e; if (!e.valid) break; }
    slot->valid = false;
    // the entry's device copies are about to change: everything queued by earlier calls must be done with them
    CK(cudaDeviceSynchronize());
    Engine::PlanEntry &PE = *slot;
    P.si = scheme_info(ma, mi, go, ge);
    P.sc = make_scoring(ma, mi, go, ge);
    std::vector<ClassPlan> cl(N_CLASSES);
    for (int c = 0; c < N_CLASSES; ++c) cl[c].cls = c;
    // Slots pair two alignments.  In cross mode a slot is (one read, two adapters): sort the adapters by length and
    // pair neighbours; the pair runs in the row-capacity class of its longer member, provided the shorter one would
    // not waste more than ~1/3 of the rows (otherwise it stays single and pairs two consecutive reads instead).
    std::vector<int32_t> order16;
    for (int a = 0; a < n_adapters; ++a) {
        int m = ad_off[a + 1] - ad_off[a];
        if (m < 0) return fail(PB200_ERR_ARG, "adapter offsets not monotone");
        int c = class_of(P.si, m);
        if (c == GENERIC_CLASS) { cl[c].ad_ids.push_back(a); cl[c].m_max = std::max(cl[c].m_max, m); }
        else order16.push_back(a);
    }
    auto len = [&](int a) { return ad_off[a + 1] - ad_off[a]; };
    std::stable_sort(order16.begin(), order16.end(), [&](int x, int y) { return len(x) > len(y); });
    std::vector<int32_t> single_of(N_CLASSES, -1);
    for (size_t i = 0; i < order16.size();) {
        const int a = order16[i], ca = class_of(P.si, len(a));
        const int capa = class_G(ca) * class_R(ca);
        if (i + 1 < order16.size()) {
            const int b = order16[i + 1], cb = class_of(P.si, len(b));
            const int capb = class_G(cb) * class_R(cb);
            if (3 * capa <= 4 * capb) {            // cap ratio <= 1.33: share a slot
                cl[ca].ad_ids.push_back(a); cl[ca].ad_ids.push_back(b);
                cl[ca].m_max = std::max(cl[ca].m_max, len(a));
                i += 2;
                continue;
            }
        }
        single_of[ca] = a;                          // at most one per class (it sits at the class's lower boundary)
        cl[ca].m_max = std::max(cl[ca].m_max, len(a));
        i += 1;
    }
    for (int c = 0; c < GENERIC_CLASS; ++c) if (single_of[c] >= 0) cl[c].ad_ids.push_back(single_of[c]);
    for (auto &c : cl) if (!c.ad_ids.empty()) P.classes.push_back(c);
    const size_t ad_bytes = (size_t)ad_off[n_adapters];
    if (int rc = PE.ad_raw.ensure(ad_bytes + 16)) return rc;
    if (int rc = PE.ad_codes.ensure(ad_bytes + 16)) return rc;
    if (int rc = PE.ad_off.ensure((size_t)(n_adapters + 1) * 4)) return rc;
    if (ad_bytes) CK(cudaMemcpyAsync(PE.ad_raw.p, adapters, ad_bytes, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(PE.ad_off.p, ad_off, (size_t)(n_adapters + 1) * 4, cudaMemcpyHostToDevice, stream));
    if (int rc = launch_encode(stream, PE.ad_raw.as<uint8_t>(), PE.ad_codes.as<uint8_t>(), (int64_t)ad_bytes, E.sm_count)) return rc;
    // class adapter-id lists, concatenated
    std::vector<int32_t> flat;
    for (auto &c : P.classes) flat.insert(flat.end(), c.ad_ids.begin(), c.ad_ids.end());
    if (int rc = PE.cls_ad.ensure(flat.size() * 4 + 16)) return rc;
    if (!flat.empty()) CK(cudaMemcpyAsync(PE.cls_ad.p, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, stream));
    // the host vectors above are pageable: the async copies have been staged by the driver before returning
    CK(cudaStreamSynchronize(stream));
    P.d_ad_codes = PE.ad_codes.as<uint8_t>(); P.d_ad_off = PE.ad_off.as<int32_t>(); P.d_cls_ad = PE.cls_ad.as<int32_t>();
    PE.ad.assign(adapters, adapters + ad_bytes);
    PE.off.assign(ad_off, ad_off + n_adapters + 1);
    PE.sc[0] = ma; PE.sc[1] = mi; PE.sc[2] = go; PE.sc[3] = ge;
    PE.plan = std::make_shared<AdapterPlan>(P);
    PE.last_used = ++E.plan_clock;
    PE.valid = true;
    return 0;
}

// Generic (int32) class: jobs are planned on the host, which needs the sequence lengths there.
int run_generic_cross(Engine &E, Stage &S, cudaStream_t stream, const ClassPlan &C, const int64_t *h_seq_off,
                      int64_t s0, int64_t cnt, int64_t base_off, const int32_t *h_ad_off, int32_t n_adapters,
                      const uint8_t *seq_codes, const uint8_t *ad_codes, const Scoring &sc, int32_t *out) {
    std::vector<GenericJob> jobs;
    const size_t budget = 2ull << 30;
    size_t used = 0;
    auto flush = [&]() -> int {
        if (jobs.empty()) return 0;
        if (int rc = E.gjobs.ensure(jobs.size() * sizeof(GenericJob))) return rc;
        if (int rc = E.gscratch.ensure(used + 64)) return rc;
        CK(cudaMemcpyAsync(E.gjobs.p, jobs.data(), jobs.size() * sizeof(GenericJob), cudaMemcpyHostToDevice, stream));
        int nb = (int)((jobs.size() + 63) / 64);
        generic_kernel<<<nb, 64, 0, stream>>>(E.gjobs.as<GenericJob>(), (int)jobs.size(), seq_codes, ad_codes, sc.ma, sc.mi,
                                              sc.go, sc.ge, E.gscratch.as<uint8_t>(), out);
        g_launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        jobs.clear(); used = 0;
        return 0;
    };
    for (int64_t s = s0; s < s0 + cnt; ++s) {
        for (int32_t a : C.ad_ids) {
            GenericJob j;
            j.seq_off = h_seq_off[s] - base_off;
            j.n = (int32_t)(h_seq_off[s + 1] - h_seq_off[s]);
            j.m = h_ad_off[a + 1] - h_ad_off[a];
            j.ad_off = h_ad_off[a];
            j.out_idx = (int32_t)((s - s0) * n_adapters + a);
            size_t tb = (((size_t)(j.n + 1) * (size_t)(j.m + 1) + 3) & ~(size_t)3) + (size_t)(j.m + 1) * 8;
            tb = (tb + 15) & ~(size_t)15;
            if (tb > (64ull << 30)) return fail(PB200_ERR_ARG, "alignment too large for the generic int32 path");
            if (used + tb > budget && !jobs.empty()) { if (int rc = flush()) return rc; }
            j.scratch_off = (int64_t)used;
            used += tb;
            jobs.push_back(j);
        }
    }
    return flush();
}

// Cross product of sequences [s0, s0+cnt) (already encoded on the device, offsets on the device) with all adapters.
// d_seq_off points at the offset of sequence s0 (cnt+1 entries), base_off is subtracted from every offset.
int run_cross_chunk(Engine &E, Stage &S, cudaStream_t stream, const AdapterPlan &P, const uint8_t *seq_codes,
                    const int64_t *d_seq_off, int64_t cnt, int64_t base_off, int64_t max_n, int32_t n_adapters,
                    int32_t *d_out, const int64_t *h_seq_off_abs, int64_t s0, const int32_t *h_ad_off) {
    if (int rc = S.misc.ensure(64)) return rc;
    int *status = S.misc.as<int>();
    unsigned long long *counter = reinterpret_cast<unsigned long long *>(S.misc.as<char>() + 16);
    // long reads (two-pass path): process sequences longest first
    const int32_t *d_order = nullptr;
    if (P.si.bounded && max_n > g_opt.direct_max && cnt > 1) {
        if (int rc = S.order.ensure((size_t)cnt * 4)) return rc;
        if (int rc = S.bins.ensure((size_t)PB_ORDER_BINS * 4)) return rc;
        CK(cudaMemsetAsync(S.bins.p, 0, (size_t)PB_ORDER_BINS * 4, stream));
        const unsigned nb = (unsigned)((cnt + 255) / 256);
        order_hist_kernel<<<nb, 256, 0, stream>>>(d_seq_off, cnt, max_n, S.bins.as<unsigned>());
        order_scan_kernel<<<1, 256, 0, stream>>>(S.bins.as<unsigned>());
        order_scatter_kernel<<<nb, 256, 0, stream>>>(d_seq_off, cnt, max_n, S.bins.as<unsigned>(), S.order.as<int32_t>());
        g_launches += 3;
        CK(cudaGetLastError());
        d_order = S.order.as<int32_t>();
    }
    size_t cls_pos = 0;
    for (const ClassPlan &C : P.classes) {
        const int32_t *d_cls = P.d_cls_ad + cls_pos;
        cls_pos += C.ad_ids.size();
        if (C.cls == GENERIC_CLASS) {
            if (!h_seq_off_abs) return fail(PB200_ERR_INTERNAL, "generic class needs host offsets");
            if (int rc = run_generic_cross(E, S, stream, C, h_seq_off_abs, s0, cnt, base_off, h_ad_off, n_adapters, seq_codes,
                                           P.d_ad_codes, P.sc, d_out)) return rc;
            continue;
        }
        TaskSrc ts;
        ts.tasks = nullptr;                                  // cross product, synthesised in the kernels
        ts.n_tasks = cnt * (int64_t)C.ad_ids.size();
        ts.cls_ad = d_cls; ts.n_cls_ad = (int32_t)C.ad_ids.size(); ts.n_adapters = n_adapters;
        ts.n_seqs = cnt; ts.seq_off = d_seq_off; ts.ad_off = P.d_ad_off;
        ts.seq_order = d_order;
        if (ts.n_tasks == 0) continue;
        (void)base_off;
        if (int rc = run_class_tasks(E, S, stream, C.cls, C.m_max, ts, max_n, seq_codes, P.d_ad_codes, P.sc,
                                     P.si, d_out, status, counter)) return rc;
    }
    return 0;
}

int check_status(Stage &S, cudaStream_t stream) {
    int st = 0;
    CK(cudaMemcpyAsync(&st, S.misc.p, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (st & 1) return fail(PB200_ERR_INTERNAL, "traceback left its window (window bound violated)");
    if (st & 2) return fail(PB200_ERR_INTERNAL, "decision kernel: value outside its table (len_aln >= table length or count > 65535)");
    return 0;
}

// seq_off rebasing kernel: offsets of a chunk relative to its first byte
__global__ void rebase_kernel(int64_t *off, int64_t n, int64_t base) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] -= base;
}
// ... and for a chunk of equally long sequences the offsets are made on the device instead of crossing PCIe (8 bytes per
// 150-byte window are 5 % of the upload of an end-trim step)
__global__ void stride_offsets_kernel(int64_t *off, int64_t n, int64_t stride) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] = i * stride;
}

// One pipeline chunk of the host-buffer API: sequences [s0, s1), longest sequence max_n.
struct HostChunk { int64_t s0, s1, max_n; bool uniform; };   // uniform: every sequence of the chunk has length max_n

// Pure host work done before the device is touched: argument validation (so a bad call fails the same way with or
// without a device).  Pair-list mode checks everything here; in cross mode the sequence offsets are checked chunk by chunk
// while the pipeline runs (next_chunk), so that the scan of chunk k+1 hides behind the device's work on chunk k.
int validate_args(const uint8_t *seqs, const int64_t *seq_off, int64_t n_seqs, const uint8_t *adapters,
                  const int32_t *ad_off, int32_t n_adapters, const int32_t *pair_seq, const int32_t *pair_adapter,
                  int64_t n_pairs, bool cross) {
    for (int32_t a = 0; a < n_adapters; ++a)
        if (ad_off[a + 1] < ad_off[a]) return fail(PB200_ERR_ARG, "adapter offsets not monotone");
    if (n_adapters > 0 && ad_off[0] < 0) return fail(PB200_ERR_ARG, "negative adapter offset");
    if (n_adapters > 0 && ad_off[n_adapters] > 0 && !adapters) return fail(PB200_ERR_ARG, "NULL adapter buffer");
    if (n_seqs > 0 && seq_off[n_seqs] > seq_off[0] && !seqs) return fail(PB200_ERR_ARG, "NULL sequence buffer");
    if (n_seqs > 0 && seq_off[0] < 0) return fail(PB200_ERR_ARG, "negative sequence offset");
    if (cross) return 0;
    if (n_pairs > 0x7fffffffll) return fail(PB200_ERR_ARG, "pair-list mode supports < 2^31 pairs per call");
    for (int64_t s = 0; s < n_seqs; ++s) {
        const int64_t len = seq_off[s + 1] - seq_off[s];
        if (len < 0) return fail(PB200_ERR_ARG, "sequence offsets not monotone");
        if (len > 0x7fff0000ll) return fail(PB200_ERR_ARG, "sequence longer than 2^31");
    }
    for (int64_t p = 0; p < n_pairs; ++p)
        if (pair_seq[p] < 0 || pair_seq[p] >= n_seqs || pair_adapter[p] < 0 || pair_adapter[p] >= n_adapters)
            return fail(PB200_ERR_ARG, "pair index out of range");
    return 0;
}

// The pipeline chunk that starts at sequence s0 of a cross-product job: chunk_tasks alignments (at most task_cap: the first
// chunks of a submit are smaller, see chunk_cap), at most chunk_bytes of sequence (but enough sequences to fill the GPU),
// its longest sequence, offsets checked.
int next_chunk(const int64_t *seq_off, int64_t n_seqs, int32_t n_adapters, int64_t s0, HostChunk &c, int64_t task_cap) {
    const int64_t max_cnt = std::max<int64_t>(1, std::min(g_opt.chunk_tasks, task_cap) / std::max<int32_t>(n_adapters, 1));
    int64_t s1 = std::min(n_seqs, s0 + max_cnt);
    while (s1 - s0 > 32768 && seq_off[s1] - seq_off[s0] > g_opt.chunk_bytes) s1 = s0 + std::max<int64_t>(32768, (s1 - s0) / 2);
    c = HostChunk{s0, s1, 0, false};
    int64_t mx = 0, mn = 0, shortest = INT64_MAX;
    for (int64_t s = s0; s < s1; ++s) {
        const int64_t len = seq_off[s + 1] - seq_off[s];
        mx = std::max(mx, len);
        mn = std::min(mn, len);
        shortest = std::min(shortest, len);
    }
    if (mn < 0) return fail(PB200_ERR_ARG, "sequence offsets not monotone");
    if (mx > 0x7fff0000ll) return fail(PB200_ERR_ARG, "sequence longer than 2^31");
    c.max_n = mx;
    c.uniform = s1 > s0 && shortest == mx;      // fixed-stride windows (the end windows of reads >= end_size): offsets are i * max_n
    return 0;
}

// One cross-product job of the host-buffer API (all sequences x all adapters of one call / one batch of a multi call).
struct CrossJob {
    const uint8_t *seqs; const int64_t *seq_off; int64_t n_seqs;
    const uint8_t *adapters; const int32_t *ad_off; int32_t n_adapters;
    int32_t *out;
    // decisions on the device (adapterEndDecisions): records are reduced per read before anything is copied back
    const pb200_end_batch_t *dec = nullptr;
    const int32_t *d_cmin = nullptr; int32_t cmin_len = 0;
    const int32_t *d_cols = nullptr;
    int64_t max_seq_len = 0;          // decision jobs: longest window the threshold table covers
};

// The first chunks of a submit ramp up (1/8, 1/4, 1/2 of chunk_tasks): the pipeline's fill -- pack + H2D of chunk 0 before
// any DP kernel can start -- costs an eighth of what a full chunk would; only when the submit is large enough to matter.
int64_t chunk_cap(size_t k, int64_t total_tasks) {
    if (total_tasks < 4 * g_opt.chunk_tasks || k >= 3) return g_opt.chunk_tasks;
    return std::max<int64_t>(4096, g_opt.chunk_tasks >> (3 - k));
}

// Packed upload or not for a submit of `total_bytes` of sequence.  Measured on B200 + Xeon 8562Y+ under a 16-CPU quota
// (round 2, 316 MB per step): the packed path is bound by the host conversion, ~51 GB/s with a 14-thread team = 6.2 ms per
// step against 6.8 ms for the plain upload (PCIe) and 4.2 ms of kernels; with 8 threads or fewer it is slower than PCIe.  So
// "auto" packs only large submits and only when the team is large enough (one rank per GPU on a box with few CPUs per rank --
// e.g. 8 ranks under that quota -- uploads the ASCII bytes as they are).  The gain depends on how many host cycles the
// process really gets: repeated runs on the same box gave 6.2-6.7 ms packed (once 8.2 ms inside a longer bench run) against
// a steady 6.8 ms plain, and non-temporal stores in the packer made it slower (7.06 vs 6.66 ms: the DMA engine then reads
// the codes from DRAM instead of the last-level cache) -- so the default stays the plain upload and packing is an option.
bool pack_wanted(int64_t total_bytes) {
    if (total_bytes <= 0 || g_opt.h2d_pack == 0) return false;
    if (g_opt.h2d_pack > 0) return true;
    return g_opt.pack_threads >= 12 && total_bytes >= (32ll << 20);
}

// One planned (and, with h2d_pack, packed) chunk on its way from the planner to the submit loop.
struct PackItem {
    size_t job = 0;
    HostChunk c{0, 0, 0, false};
    int buf = -1;                 // index of the pinned pack buffer holding the chunk's 4-bit codes (-1: not packed)
    int rc = 0; std::string err;  // planning error (bad offsets): the submit loop stops here
    bool end = false;             // no more chunks
};

// Planning of a submit's chunk sequence (the order of run_cross_jobs' loop), shared by the inline path and the packer thread.
struct ChunkPlanner {
    const std::vector<CrossJob> *jobs;
    size_t j = 0, k = 0;
    int64_t s0 = 0, total_tasks = 0;
    explicit ChunkPlanner(const std::vector<CrossJob> &J) : jobs(&J) {
        for (const CrossJob &x : J) total_tasks += x.n_seqs * (int64_t)x.n_adapters;
    }
    // next chunk -> item (rc / end set accordingly); errors come back as text because the planner may run on another thread
    void next(PackItem &it) {
        it = PackItem();
        while (j < jobs->size() && ((*jobs)[j].n_seqs <= 0 || (*jobs)[j].n_adapters <= 0 || s0 >= (*jobs)[j].n_seqs)) { ++j; s0 = 0; }
        if (j >= jobs->size()) { it.end = true; return; }
        const CrossJob &J = (*jobs)[j];
        it.job = j;
        it.rc = next_chunk(J.seq_off, J.n_seqs, J.n_adapters, s0, it.c, chunk_cap(k, total_tasks));
        if (!it.rc && J.dec && it.c.max_n > J.max_seq_len)
            it.rc = fail(PB200_ERR_ARG, "decision batches take windows of at most end_size bases");
        if (it.rc) { it.err = g_err; return; }
        s0 = it.c.s1;
        ++k;
    }
};

// Packer: a persistent host thread per engine that plans the chunks of a submit and converts them to 4-bit codes (option
// h2d_pack; hostpack.cpp, OpenMP team of pack_threads) AHEAD of the submit loop, into a ring of NPACK pinned buffers -- the
// conversion of chunk k+1.. runs while the submit loop enqueues chunk k and the device works on chunk k-1 (round 2: with
// the packer inside the submit loop the e2e step was bound by the host conversion, profiles/r2_options).
constexpr int NPACK = NSTAGE + 2;
struct Packer {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    const std::vector<CrossJob> *jobs = nullptr;   // request: non-null while a run is wanted
    bool abort = false, stop = false, busy = false;
    std::deque<PackItem> q;
    HostBuf buf[NPACK];
    cudaEvent_t ev[NPACK];
    enum { FREE = 0, QUEUED = 1, INFLIGHT = 2 };
    int state[NPACK] = {0, 0, 0, 0, 0};
    int device = 0;
    bool ev_made = false;

    void loop() {
        cudaSetDevice(device);
        for (;;) {
            const std::vector<CrossJob> *J;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || (jobs && !busy && q.empty()); });
                if (stop) return;
                J = jobs; busy = true;
            }
            ChunkPlanner plan(*J);
            for (size_t k = 0;; ++k) {
                PackItem it;
                bool stop_now;
                { std::lock_guard<std::mutex> lk(mu); stop_now = abort; }
                if (stop_now) it.end = true; else plan.next(it);
                const int b = (int)(k % NPACK);
                if (!it.end && !it.rc) {
                    bool inflight = false;
                    {   // the buffer's previous chunk must have been taken by the submit loop and its upload must be complete
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return abort || state[b] != QUEUED; });
                        if (abort) { it = PackItem(); it.end = true; }
                        inflight = state[b] == INFLIGHT;
                    }
                    if (!it.end) {
                        if (inflight) cudaEventSynchronize(ev[b]);
                        const CrossJob &X = (*J)[it.job];
                        const int64_t base = X.seq_off[it.c.s0], bytes = X.seq_off[it.c.s1] - base;
                        g_err.clear();
                        if (bytes > 0 && buf[b].ensure(((size_t)bytes + 1) / 2)) { it.rc = PB200_ERR_CUDA; it.err = g_err; }
                        else if (bytes > 0) { NvtxRange r("pb200:host_pack"); pb_pack_nibbles(X.seqs + base, bytes, buf[b].p, g_opt.pack_threads); }
                        it.buf = b;
                    }
                }
                const bool last = it.end || it.rc != 0;
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (it.buf >= 0) state[b] = QUEUED;
                    q.push_back(std::move(it));
                    if (last) { jobs = nullptr; busy = false; }
                }
                cv.notify_all();
                if (last) break;
            }
        }
    }
    ~Packer() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        if (th.joinable()) th.join();
        for (int i = 0; i < NPACK; ++i) if (buf[i].p) { cudaFreeHost(buf[i].p); buf[i].p = nullptr; }
    }
};

// Chunks of every job flow through ONE ring of NSTAGE streams (H2D / kernels / D2H of consecutive chunks overlap, also
// across job boundaries: no fill / drain bubble between the start-window and the end-window batch of an end-trim step).
// Caller holds E.mu and has run E.init().
int run_cross_jobs(Engine &E, std::vector<CrossJob> &jobs, int ma, int mi, int go, int ge) {
    NvtxRange submit_range("pb200:submit");
    for (int i = 0; i < NSTAGE; ++i) {
        if (int rc = E.st[i].misc.ensure(64)) return rc;
        CK(cudaMemsetAsync(E.st[i].misc.p, 0, 64, E.st[i].stream));
    }
    int64_t total_bytes = 0;
    for (const CrossJob &J : jobs) if (J.n_seqs > 0 && J.n_adapters > 0) total_bytes += J.seq_off[J.n_seqs] - J.seq_off[0];
    const bool packed = pack_wanted(total_bytes);
    Packer *PK = nullptr;
    if (packed) {
        if (!E.packer) {
            auto pk = std::make_shared<Packer>();
            pk->device = E.device;
            for (int i = 0; i < NPACK; ++i) CK(cudaEventCreateWithFlags(&pk->ev[i], cudaEventDisableTiming));
            pk->ev_made = true;
            pk->th = std::thread([p = pk.get()] { p->loop(); });
            E.packer = pk;
        }
        PK = static_cast<Packer *>(E.packer.get());
        {
            std::lock_guard<std::mutex> lk(PK->mu);
            PK->abort = false;
            PK->jobs = &jobs;
        }
        PK->cv.notify_all();
    }
    ChunkPlanner inline_plan(jobs);
    auto next_item = [&](PackItem &it) {
        if (!PK) { inline_plan.next(it); return; }
        std::unique_lock<std::mutex> lk(PK->mu);
        PK->cv.wait(lk, [&] { return !PK->q.empty(); });
        it = std::move(PK->q.front());
        PK->q.pop_front();
    };
    auto submit = [&](const CrossJob &J, const AdapterPlan &P, const PackItem &it, Stage &S) -> int {
        const HostChunk &c = it.c;
        const int64_t s0 = c.s0, cnt = c.s1 - c.s0;
        const int64_t base = J.seq_off[s0];
        const int64_t bytes = J.seq_off[c.s1] - base;
        cudaStream_t stream = S.stream;
        NvtxRange chunk_range("pb200:chunk");
        {
            NvtxRange r("pb200:stage_wait");
            CK(cudaStreamSynchronize(stream));   // previous use of this stage's buffers is complete
        }
        if (int rc = S.seq_raw.ensure((size_t)bytes + 16)) return rc;
        if (int rc = S.seq_codes.ensure((size_t)bytes + 16)) return rc;
        if (int rc = S.seq_off.ensure((size_t)(cnt + 1) * 8)) return rc;
        if (int rc = S.out.ensure((size_t)cnt * J.n_adapters * PB_REC * 4)) return rc;
        {
        NvtxRange r("pb200:h2d");
        if (it.buf >= 0) {
            // Dna5 conversion was done on the host cores by the packer thread, two codes per byte: half the bytes cross PCIe
            if (bytes) CK(cudaMemcpyAsync(S.seq_raw.p, PK->buf[it.buf].p, ((size_t)bytes + 1) / 2, cudaMemcpyHostToDevice, stream));
            CK(cudaEventRecord(PK->ev[it.buf], stream));
            { std::lock_guard<std::mutex> lk(PK->mu); PK->state[it.buf] = Packer::INFLIGHT; }
            PK->cv.notify_all();
        } else if (bytes) {
            CK(cudaMemcpyAsync(S.seq_raw.p, J.seqs + base, (size_t)bytes, cudaMemcpyHostToDevice, stream));
        }
        if (!c.uniform) CK(cudaMemcpyAsync(S.seq_off.p, J.seq_off + s0, (size_t)(cnt + 1) * 8, cudaMemcpyHostToDevice, stream));
        }
        NvtxRange dp_range("pb200:dp");
        if (c.uniform) stride_offsets_kernel<<<(unsigned)((cnt + 1 + 255) / 256), 256, 0, stream>>>(S.seq_off.as<int64_t>(), cnt + 1, c.max_n);
        else rebase_kernel<<<(unsigned)((cnt + 1 + 255) / 256), 256, 0, stream>>>(S.seq_off.as<int64_t>(), cnt + 1, base);
        g_launches++;
        if (it.buf >= 0) {
            if (int rc = launch_unpack(stream, S.seq_raw.as<uint8_t>(), S.seq_codes.as<uint8_t>(), bytes, E.sm_count)) return rc;
        } else {
            if (int rc = launch_encode(stream, S.seq_raw.as<uint8_t>(), S.seq_codes.as<uint8_t>(), bytes, E.sm_count)) return rc;
        }
        if (int rc = run_cross_chunk(E, S, stream, P, S.seq_codes.as<uint8_t>(), S.seq_off.as<int64_t>(), cnt, base, c.max_n,
                                     J.n_adapters, S.out.as<int32_t>(), J.seq_off, s0, J.ad_off)) return rc;
        if (J.dec) {
            const pb200_end_batch_t &D = *J.dec;
            if (int rc = S.dec_trim.ensure((size_t)cnt * 4)) return rc;
            if (int rc = S.dec_pairs.ensure((size_t)cnt * std::max<int32_t>(D.n_score_cols, 1) * 4)) return rc;
            DecideArgs a;
            a.records = S.out.as<int32_t>(); a.n = cnt; a.n_adapters = J.n_adapters;
            a.is_start = D.is_start; a.end_size = D.end_size; a.extra_trim = D.extra_trim_size; a.min_trim = D.min_trim_size;
            a.cmin = J.d_cmin; a.cmin_len = J.cmin_len; a.cols = J.d_cols; a.n_cols = D.n_score_cols;
            a.trim = S.dec_trim.as<int32_t>();
            a.pairs = (D.score_pairs && D.n_score_cols > 0) ? S.dec_pairs.as<uint32_t>() : nullptr;
            a.top2 = nullptr;
            if (D.top2) {
                if (int rc = S.dec_top2.ensure((size_t)cnt * 6 * 4)) return rc;
                a.top2 = S.dec_top2.as<int32_t>();
            }
            const int64_t blocks = std::min<int64_t>((cnt + 3) / 4, (int64_t)E.sm_count * 16);
            decide_kernel<<<(unsigned)blocks, 128, 0, stream>>>(a, S.misc.as<int>());
            g_launches++;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(D.trim + s0, S.dec_trim.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, stream));
            if (a.pairs)
                CK(cudaMemcpyAsync(D.score_pairs + (size_t)s0 * D.n_score_cols * 2, S.dec_pairs.p,
                                   (size_t)cnt * D.n_score_cols * 4, cudaMemcpyDeviceToHost, stream));
            if (a.top2)
                CK(cudaMemcpyAsync(D.top2 + (size_t)s0 * 6, S.dec_top2.p, (size_t)cnt * 6 * 4, cudaMemcpyDeviceToHost, stream));
        }
        if (J.out) {
            NvtxRange r("pb200:d2h");
            CK(cudaMemcpyAsync(J.out + (size_t)s0 * J.n_adapters * PB_REC, S.out.p, (size_t)cnt * J.n_adapters * PB_REC * 4,
                               cudaMemcpyDeviceToHost, stream));
        }
        return 0;
    };
    int rc_final = 0;
    std::string first_err;
    size_t k = 0, cur_job = (size_t)-1;
    AdapterPlan P;
    bool drained = false;
    while (!rc_final) {
        PackItem it;
        next_item(it);
        if (it.end) { drained = true; break; }
        if (it.rc) { rc_final = it.rc; first_err = it.err; drained = true; break; }     // planning error ends the planner's run too
        const CrossJob &J = jobs[it.job];
        if (it.job != cur_job) {
            // The adapter plan is made (or found in the 4-entry cache) right before the job's first chunk: a miss waits for the
            // device to go idle before it recycles an entry, so chunks of earlier jobs never lose their adapter copies.
            cur_job = it.job;
            P = AdapterPlan();
            rc_final = plan_adapters(E, E.st[k % NSTAGE].stream, J.adapters, J.ad_off, J.n_adapters, ma, mi, go, ge, P);
        }
        if (!rc_final) rc_final = submit(J, P, it, E.st[k % NSTAGE]);
        if (rc_final) first_err = g_err;
        ++k;
    }
    if (PK && !drained) {
        // an error on this side: stop the packer and take its remaining items so that it is idle (and its buffers free) again
        { std::lock_guard<std::mutex> lk(PK->mu); PK->abort = true; }
        PK->cv.notify_all();
        for (;;) {
            PackItem it;
            next_item(it);
            if (it.buf >= 0) { std::lock_guard<std::mutex> lk(PK->mu); PK->state[it.buf] = Packer::FREE; }
            PK->cv.notify_all();
            if (it.end || it.rc) break;
        }
    }
    // Whatever happened, nothing may still be writing into the caller's `out` (or reading `seqs`) when we return.
    for (int i = 0; i < NSTAGE; ++i) {
        if (!rc_final && E.st[i].misc.p) { rc_final = check_status(E.st[i], E.st[i].stream); if (rc_final) first_err = g_err; }
        cudaError_t e = cudaStreamSynchronize(E.st[i].stream);
        if (e != cudaSuccess && !rc_final) { rc_final = fail(PB200_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e)); first_err = g_err; }
    }
    if (PK) {        // all uploads are complete: the pack buffers are free for the next submit
        std::lock_guard<std::mutex> lk(PK->mu);
        for (int i = 0; i < NPACK; ++i) PK->state[i] = Packer::FREE;
    }
    if (rc_final && !first_err.empty()) g_err = first_err;
    return rc_final;
}

int batch_host(const uint8_t *seqs, const int64_t *seq_off, int64_t n_seqs, const uint8_t *adapters,
               const int32_t *ad_off, int32_t n_adapters, const int32_t *pair_seq, const int32_t *pair_adapter,
               int64_t n_pairs, int ma, int mi, int go, int ge, int32_t *out) {
    if (n_seqs < 0 || n_adapters < 0 || n_pairs < 0) return fail(PB200_ERR_ARG, "negative count");
    if ((pair_seq == nullptr) != (pair_adapter == nullptr)) return fail(PB200_ERR_ARG, "pair_seq/pair_adapter must both be given or both NULL");
    const bool cross = pair_seq == nullptr;
    if (cross && n_pairs != n_seqs * (int64_t)n_adapters) return fail(PB200_ERR_ARG, "cross mode: n_pairs != n_seqs*n_adapters");
    if (n_pairs == 0) return 0;
    if (!seq_off || !ad_off || !out) return fail(PB200_ERR_ARG, "NULL pointer");
    load_env_options();
    std::vector<CrossJob> jobs(1);
    jobs[0] = CrossJob{seqs, seq_off, n_seqs, adapters, ad_off, n_adapters, out};
    if (int rc = validate_args(seqs, seq_off, n_seqs, adapters, ad_off, n_adapters, pair_seq, pair_adapter, n_pairs, cross)) return rc;
    Engine *Ep = nullptr;
    if (int rc = get_engine(&Ep)) return rc;
    Engine &E = *Ep;
    std::lock_guard<std::mutex> lk(E.mu);
    if (int rc = E.init()) return rc;
    if (cross) return run_cross_jobs(E, jobs, ma, mi, go, ge);

    AdapterPlan P;
    if (int rc = plan_adapters(E, E.st[0].stream, adapters, ad_off, n_adapters, ma, mi, go, ge, P)) return rc;
    for (int i = 0; i < NSTAGE; ++i) {
        if (int rc = E.st[i].misc.ensure(64)) return rc;
        CK(cudaMemsetAsync(E.st[i].misc.p, 0, 64, E.st[i].stream));
    }

    // ---- pair-list mode: all sequences resident, pairs ordered per class on the host ----
    auto run_pairs = [&]() -> int {
    NvtxRange submit_range("pb200:submit_pairs");
    Stage &S = E.st[0];
    cudaStream_t stream = S.stream;
    const int64_t bytes = seq_off[n_seqs];
    if (int rc = S.seq_raw.ensure((size_t)bytes + 16)) return rc;
    if (int rc = S.seq_codes.ensure((size_t)bytes + 16)) return rc;
    if (int rc = S.seq_off.ensure((size_t)(n_seqs + 1) * 8)) return rc;
    if (int rc = S.out.ensure((size_t)n_pairs * PB_REC * 4)) return rc;
    if (int rc = S.pair_seq.ensure((size_t)n_pairs * 4)) return rc;
    if (int rc = S.pair_ad.ensure((size_t)n_pairs * 4)) return rc;
    if (int rc = S.order.ensure((size_t)n_pairs * 4)) return rc;
    if (int rc = S.misc.ensure(64)) return rc;
    CK(cudaMemsetAsync(S.misc.p, 0, 64, stream));
    if (bytes) CK(cudaMemcpyAsync(S.seq_raw.p, seqs, (size_t)bytes, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(S.seq_off.p, seq_off, (size_t)(n_seqs + 1) * 8, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(S.pair_seq.p, pair_seq, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(S.pair_ad.p, pair_adapter, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, stream));
    if (int rc = launch_encode(stream, S.seq_raw.as<uint8_t>(), S.seq_codes.as<uint8_t>(), bytes, E.sm_count)) return rc;
    // class of every adapter, then per-class ordering (adapter, length) so that slot halves have similar shapes
    std::vector<int> ad_class(n_adapters);
    for (int a = 0; a < n_adapters; ++a) ad_class[a] = class_of(P.si, ad_off[a + 1] - ad_off[a]);
    std::vector<std::vector<int32_t>> per_class(N_CLASSES);
    for (int64_t p = 0; p < n_pairs; ++p) per_class[ad_class[pair_adapter[p]]].push_back((int32_t)p);
    int *status = S.misc.as<int>();
    unsigned long long *counter = reinterpret_cast<unsigned long long *>(S.misc.as<char>() + 16);
    for (int c = 0; c < N_CLASSES; ++c) {
        auto &ord = per_class[c];
        if (ord.empty()) continue;
        int m_max = 0;
        int64_t max_n = 0;
        for (int32_t p : ord) {
            m_max = std::max(m_max, ad_off[pair_adapter[p] + 1] - ad_off[pair_adapter[p]]);
            max_n = std::max(max_n, seq_off[pair_seq[p] + 1] - seq_off[pair_seq[p]]);
        }
        if (max_n > 0x7fff0000ll) return fail(PB200_ERR_ARG, "sequence longer than 2^31");
        if (c == GENERIC_CLASS) {
            // generic: reuse the cross helper one pair at a time through a tiny job list
            std::vector<GenericJob> jobs;
            size_t used = 0;
            const size_t budget = 2ull << 30;
            auto flush = [&]() -> int {
                if (jobs.empty()) return 0;
                if (int rc = E.gjobs.ensure(jobs.size() * sizeof(GenericJob))) return rc;
                if (int rc = E.gscratch.ensure(used + 64)) return rc;
                CK(cudaMemcpyAsync(E.gjobs.p, jobs.data(), jobs.size() * sizeof(GenericJob), cudaMemcpyHostToDevice, stream));
                int nb = (int)((jobs.size() + 63) / 64);
                generic_kernel<<<nb, 64, 0, stream>>>(E.gjobs.as<GenericJob>(), (int)jobs.size(), S.seq_codes.as<uint8_t>(),
                                                      P.d_ad_codes, ma, mi, go, ge, E.gscratch.as<uint8_t>(),
                                                      S.out.as<int32_t>());
                g_launches++;
                CK(cudaGetLastError());
                CK(cudaStreamSynchronize(stream));
                jobs.clear(); used = 0;
                return 0;
            };
            for (int32_t p : ord) {
                GenericJob j;
                int64_t s = pair_seq[p]; int a = pair_adapter[p];
                j.seq_off = seq_off[s]; j.n = (int32_t)(seq_off[s + 1] - seq_off[s]);
                j.m = ad_off[a + 1] - ad_off[a]; j.ad_off = ad_off[a]; j.out_idx = p;
                size_t tb = (((size_t)(j.n + 1) * (size_t)(j.m + 1) + 3) & ~(size_t)3) + (size_t)(j.m + 1) * 8;
                tb = (tb + 15) & ~(size_t)15;
                if (tb > (64ull << 30)) return fail(PB200_ERR_ARG, "alignment too large for the generic int32 path");
                if (used + tb > budget && !jobs.empty()) { if (int rc = flush()) return rc; }
                j.scratch_off = (int64_t)used; used += tb;
                jobs.push_back(j);
            }
            if (int rc = flush()) return rc;
            continue;
        }
        std::sort(ord.begin(), ord.end(), [&](int32_t x, int32_t y) {
            int ax = pair_adapter[x], ay = pair_adapter[y];
            if (ax != ay) return ax < ay;
            int64_t nx = seq_off[pair_seq[x] + 1] - seq_off[pair_seq[x]], ny = seq_off[pair_seq[y] + 1] - seq_off[pair_seq[y]];
            if (nx != ny) return nx < ny;
            return x < y;
        });
        const int64_t n_tasks = (int64_t)ord.size();
        if (int rc = S.tasks.ensure((size_t)n_tasks * sizeof(Task))) return rc;
        CK(cudaMemcpyAsync(S.order.p, ord.data(), (size_t)n_tasks * 4, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));  // ord is pageable and reused per class
        build_tasks_pairs_kernel<<<(unsigned)((n_tasks + 255) / 256), 256, 0, stream>>>(
            S.tasks.as<Task>(), n_tasks, S.order.as<int32_t>(), S.pair_seq.as<int32_t>(), S.pair_ad.as<int32_t>(),
            S.seq_off.as<int64_t>(), P.d_ad_off);
        g_launches++;
        CK(cudaGetLastError());
        TaskSrc ts;
        ts.tasks = S.tasks.as<Task>(); ts.n_tasks = n_tasks;
        ts.cls_ad = nullptr; ts.n_cls_ad = 0; ts.n_adapters = n_adapters; ts.n_seqs = n_seqs;
        ts.seq_off = S.seq_off.as<int64_t>(); ts.ad_off = P.d_ad_off; ts.seq_order = nullptr;
        if (int rc = run_class_tasks(E, S, stream, c, m_max, ts, max_n, S.seq_codes.as<uint8_t>(), P.d_ad_codes,
                                     P.sc, P.si, S.out.as<int32_t>(), status, counter)) return rc;
    }
    CK(cudaMemcpyAsync(out, S.out.p, (size_t)n_pairs * PB_REC * 4, cudaMemcpyDeviceToHost, stream));
    return check_status(S, stream);
    };
    const int rc_pairs = run_pairs();
    if (rc_pairs) {                         // nothing may still be reading the caller's buffers when we return
        const std::string first_err = g_err;
        cudaStreamSynchronize(E.st[0].stream);
        g_err = first_err;
    }
    return rc_pairs;
}

int batch_host_multi(const pb200_batch_t *batches, int n_batches, int ma, int mi, int go, int ge) {
    if (n_batches < 0 || (n_batches > 0 && !batches)) return fail(PB200_ERR_ARG, "bad batch list");
    load_env_options();
    std::vector<CrossJob> jobs;
    for (int b = 0; b < n_batches; ++b) {
        const pb200_batch_t &B = batches[b];
        if (B.n_seqs < 0 || B.n_adapters < 0) return fail(PB200_ERR_ARG, "negative count");
        if (B.n_seqs == 0 || B.n_adapters == 0) continue;
        if (!B.seq_off || !B.ad_off || !B.out) return fail(PB200_ERR_ARG, "NULL pointer");
        jobs.push_back(CrossJob{B.seqs, B.seq_off, B.n_seqs, B.adapters, B.ad_off, B.n_adapters, B.out});
        if (int rc = validate_args(B.seqs, B.seq_off, B.n_seqs, B.adapters, B.ad_off, B.n_adapters, nullptr, nullptr,
                                   B.n_seqs * (int64_t)B.n_adapters, true)) return rc;
    }
    if (jobs.empty()) return 0;
    Engine *Ep = nullptr;
    if (int rc = get_engine(&Ep)) return rc;
    Engine &E = *Ep;
    std::lock_guard<std::mutex> lk(E.mu);
    if (int rc = E.init()) return rc;
    return run_cross_jobs(E, jobs, ma, mi, go, ge);
}

// float("%f" % (100.0*c/l)) exactly as the reference chain produces it: std::to_string(double) = sprintf("%f")
// (porechop/src/alignment.cpp:113-121), then Python's float() = strtod (nanopore_read.py:488-489)
double percent_exact(int32_t c, int32_t l) {
    char buf[64];
    volatile double cd = (double)c, ld = (double)l;
    snprintf(buf, sizeof buf, "%f", 100.0 * cd / ld);
    return strtod(buf, nullptr);
}
void trim_threshold_table(double thr, int32_t len, int32_t *cmin) {
    if (len > 0) cmin[0] = INT32_MAX;
    for (int32_t l = 1; l < len; ++l) {
        // percent_exact(., l) is non-decreasing: binary search for the first c with value > thr
        int32_t lo = 0, hi = l + 1;
        while (lo < hi) {
            const int32_t mid = lo + (hi - lo) / 2;
            if (percent_exact(mid, l) > thr) hi = mid; else lo = mid + 1;
        }
        cmin[l] = lo;
    }
}

// threshold tables are pure functions of (threshold, length): built once, reused by every later call
std::shared_ptr<const std::vector<int32_t>> cached_threshold_table(double thr, int32_t len) {
    static std::mutex mu;
    static std::map<std::pair<double, int32_t>, std::shared_ptr<const std::vector<int32_t>>> cache;
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_pair(thr, len);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto t = std::make_shared<std::vector<int32_t>>((size_t)len);
    trim_threshold_table(thr, len, t->data());
    if (cache.size() >= 64) cache.clear();                     // growth guard (thresholds are few in practice)
    cache[key] = t;
    return t;
}

int batch_end_decisions(const pb200_end_batch_t *batches, int n_batches, int ma, int mi, int go, int ge) {
    if (n_batches < 0 || (n_batches > 0 && !batches)) return fail(PB200_ERR_ARG, "bad batch list");
    load_env_options();
    std::vector<CrossJob> jobs;
    std::vector<std::shared_ptr<const std::vector<int32_t>>> tables;
    for (int b = 0; b < n_batches; ++b) {
        const pb200_end_batch_t &D = batches[b];
        const pb200_batch_t &B = D.batch;
        if (B.n_seqs < 0 || B.n_adapters < 0 || D.n_score_cols < 0) return fail(PB200_ERR_ARG, "negative count");
        if (B.n_seqs == 0) continue;
        if (!D.trim || (D.n_score_cols > 0 && ((!D.score_pairs && !D.top2) || !D.score_cols))) return fail(PB200_ERR_ARG, "NULL pointer");
        if (!(D.end_threshold >= 0.0)) return fail(PB200_ERR_ARG, "end_threshold must be >= 0 for the device decisions");
        for (int32_t k = 0; k < D.n_score_cols; ++k)
            if (D.score_cols[k] < 0 || D.score_cols[k] >= B.n_adapters) return fail(PB200_ERR_ARG, "score column out of range");
        if (B.n_adapters == 0) {                 // no adapters: nothing aligns, nothing is trimmed
            memset(D.trim, 0, (size_t)B.n_seqs * 4);
            if (D.top2) for (int64_t s = 0; s < B.n_seqs; ++s) { int32_t *o = D.top2 + s * 6; o[0] = o[3] = -1; o[1] = o[4] = 0; o[2] = o[5] = 1; }
            continue;
        }
        if (!B.seq_off || !B.ad_off) return fail(PB200_ERR_ARG, "NULL pointer");
        CrossJob J{B.seqs, B.seq_off, B.n_seqs, B.adapters, B.ad_off, B.n_adapters, B.out};
        if (int rc = validate_args(B.seqs, B.seq_off, B.n_seqs, B.adapters, B.ad_off, B.n_adapters, nullptr, nullptr,
                                   B.n_seqs * (int64_t)B.n_adapters, true)) return rc;
        // a window is seq[:end_size] / seq[-end_size:] (nanopore_read.py:172,194) and the aligned region is at most
        // window + adapter columns long: that sizes the threshold table (windows are checked chunk by chunk)
        if (D.end_size < 0) return fail(PB200_ERR_ARG, "negative end_size");
        int64_t m_max = 0;
        for (int32_t a = 0; a < B.n_adapters; ++a) m_max = std::max<int64_t>(m_max, B.ad_off[a + 1] - B.ad_off[a]);
        if ((int64_t)D.end_size + m_max + 2 > 65535) return fail(PB200_ERR_ARG, "windows too long for the device decisions (use the record API)");
        if (D.top2 && ((int64_t)D.end_size + m_max + 2 > 4095 || D.n_score_cols >= 0xFFFF))
            return fail(PB200_ERR_ARG, "windows too long / too many score columns for the device barcode ranking (use score_pairs)");
        J.dec = &D;
        J.max_seq_len = D.end_size;
        J.cmin_len = (int32_t)(D.end_size + m_max + 2);
        tables.push_back(cached_threshold_table(D.end_threshold, J.cmin_len));
        jobs.push_back(std::move(J));
    }
    if (jobs.empty()) return 0;
    if ((int)jobs.size() > Engine::MAX_DEC_JOBS) return fail(PB200_ERR_ARG, "too many decision batches in one call");
    Engine *Ep = nullptr;
    if (int rc = get_engine(&Ep)) return rc;
    Engine &E = *Ep;
    std::lock_guard<std::mutex> lk(E.mu);
    if (int rc = E.init()) return rc;
    cudaStream_t s0 = E.st[0].stream;
    for (size_t j = 0; j < jobs.size(); ++j) {
        const pb200_end_batch_t &D = *jobs[j].dec;
        if (int rc = E.dec_cmin[j].ensure((size_t)jobs[j].cmin_len * 4)) return rc;
        if (int rc = E.dec_cols[j].ensure((size_t)std::max<int32_t>(D.n_score_cols, 1) * 4)) return rc;
        CK(cudaMemcpyAsync(E.dec_cmin[j].p, tables[j]->data(), (size_t)jobs[j].cmin_len * 4, cudaMemcpyHostToDevice, s0));
        if (D.n_score_cols > 0)
            CK(cudaMemcpyAsync(E.dec_cols[j].p, D.score_cols, (size_t)D.n_score_cols * 4, cudaMemcpyHostToDevice, s0));
        jobs[j].d_cmin = E.dec_cmin[j].as<int32_t>();
        jobs[j].d_cols = E.dec_cols[j].as<int32_t>();
    }
    CK(cudaStreamSynchronize(s0));               // tables are in place before any stage's stream uses them
    return run_cross_jobs(E, jobs, ma, mi, go, ge);
}

int batch_device(const uint8_t *d_seqs, const int64_t *d_seq_off, int64_t n_seqs, int64_t total_seq_bytes,
                 int64_t max_seq_len, const uint8_t *adapters, const int32_t *ad_off, int32_t n_adapters, int ma, int mi,
                 int go, int ge, int32_t *d_out, void *user_stream) {
    if (n_seqs < 0 || n_adapters < 0 || total_seq_bytes < 0) return fail(PB200_ERR_ARG, "negative count");
    if (n_seqs == 0 || n_adapters == 0) return 0;
    if (!d_seq_off || !ad_off || !d_out) return fail(PB200_ERR_ARG, "NULL pointer");
    Engine *Ep = nullptr;
    if (int rc = get_engine(&Ep)) return rc;
    Engine &E = *Ep;
    std::lock_guard<std::mutex> lk(E.mu);
    if (int rc = E.init()) return rc;
    Stage &S = E.st[0];
    cudaStream_t stream = user_stream ? (cudaStream_t)user_stream : S.stream;
    NvtxRange submit_range("pb200:submit_device");
    AdapterPlan P;
    if (int rc = plan_adapters(E, stream, adapters, ad_off, n_adapters, ma, mi, go, ge, P)) return rc;
    if (int rc = S.seq_codes.ensure((size_t)total_seq_bytes + 16)) return rc;
    if (int rc = launch_encode(stream, d_seqs, S.seq_codes.as<uint8_t>(), total_seq_bytes, E.sm_count)) return rc;
    if (int rc = S.misc.ensure(64)) return rc;
    CK(cudaMemsetAsync(S.misc.as<char>() + 16, 0, 48, stream));   // status word (offset 0) is sticky until pb200Synchronize
    if (max_seq_len < 0) {
        unsigned long long *d_max = reinterpret_cast<unsigned long long *>(S.misc.as<char>() + 32);
        CK(cudaMemsetAsync(d_max, 0, 8, stream));
        int64_t blocks = std::min<int64_t>((n_seqs + 255) / 256, 1024);
        max_len_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_seq_off, n_seqs, d_max);
        g_launches++;
        unsigned long long h = 0;
        CK(cudaMemcpyAsync(&h, d_max, 8, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        max_seq_len = (int64_t)h;
    }
    if (max_seq_len > 0x7fff0000ll) return fail(PB200_ERR_ARG, "sequence longer than 2^31");
    std::vector<int64_t> h_off;   // only fetched when a generic class exists
    for (auto &c : P.classes) if (c.cls == GENERIC_CLASS) {
        h_off.resize((size_t)n_seqs + 1);
        CK(cudaMemcpyAsync(h_off.data(), d_seq_off, (size_t)(n_seqs + 1) * 8, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        break;
    }
    int64_t max_cnt = std::max<int64_t>(1, g_opt.device_chunk_tasks / std::max<int32_t>(n_adapters, 1));
    for (int64_t s0 = 0; s0 < n_seqs; s0 += max_cnt) {
        const int64_t cnt = std::min(max_cnt, n_seqs - s0);
        if (int rc = run_cross_chunk(E, S, stream, P, S.seq_codes.as<uint8_t>(), d_seq_off + s0, cnt, 0, max_seq_len,
                                     n_adapters, d_out + (size_t)s0 * n_adapters * PB_REC,
                                     h_off.empty() ? nullptr : h_off.data(), s0, ad_off)) return rc;
    }
    return 0;
}

}  // namespace

// =====================================================================================================
extern "C" {

int adapterAlignmentBatch(const uint8_t *seqs, const int64_t *seq_off, int64_t n_seqs, const uint8_t *adapters,
                          const int32_t *ad_off, int32_t n_adapters, const int32_t *pair_seq,
                          const int32_t *pair_adapter, int64_t n_pairs, int ma, int mi, int go, int ge, int32_t *out) {
    g_err.clear();
    return batch_host(seqs, seq_off, n_seqs, adapters, ad_off, n_adapters, pair_seq, pair_adapter, n_pairs, ma, mi, go, ge, out);
}

int adapterAlignmentBatchMulti(const pb200_batch_t *batches, int n_batches, int ma, int mi, int go, int ge) {
    g_err.clear();
    return batch_host_multi(batches, n_batches, ma, mi, go, ge);
}

int adapterEndDecisions(const pb200_end_batch_t *batches, int n_batches, int ma, int mi, int go, int ge) {
    g_err.clear();
    return batch_end_decisions(batches, n_batches, ma, mi, go, ge);
}

int pb200TrimThresholdTable(double end_threshold, int32_t len, int32_t *cmin) {
    if (len < 0 || (len > 0 && !cmin) || !(end_threshold >= 0.0)) return PB200_ERR_ARG;
    trim_threshold_table(end_threshold, len, cmin);
    return 0;
}

int adapterAlignmentBatchDevice(const uint8_t *d_seqs, const int64_t *d_seq_off, int64_t n_seqs, int64_t total_seq_bytes,
                                int64_t max_seq_len, const uint8_t *adapters, const int32_t *ad_off, int32_t n_adapters,
                                int ma, int mi, int go, int ge, int32_t *d_out, void *stream) {
    g_err.clear();
    return batch_device(d_seqs, d_seq_off, n_seqs, total_seq_bytes, max_seq_len, adapters, ad_off, n_adapters, ma, mi, go,
                        ge, d_out, stream);
}

int pb200FormatRecord(const int32_t *r, char *buf, int buflen) {
    int n;
    if (r[0] == -1 && r[4] == PB_SCORE_EMPTY) {
        n = snprintf(buf, (size_t)buflen, "-1,0,-1,0,-2147483648,0.000000,0.000000");
    } else {
        // the two divisions of porechop/src/alignment.cpp:82,90 in double; 0/0 prints "-nan" like the reference
        volatile double c1 = (double)r[5], l1 = (double)r[6], c2 = (double)r[7], l2 = (double)r[8];
        double p1 = 100.0 * c1 / l1, p2 = 100.0 * c2 / l2;
        n = snprintf(buf, (size_t)buflen, "%d,%d,%d,%d,%d,%f,%f", r[0], r[1], r[2], r[3], r[4], p1, p2);
    }
    return (n < 0 || n >= buflen) ? -1 : n;
}

int pb200PackNibbles(const uint8_t *ascii, int64_t n, uint8_t *packed, int threads) {
    if (n < 0 || (n > 0 && (!ascii || !packed))) return PB200_ERR_ARG;
    pb_pack_nibbles(ascii, n, packed, threads);
    return 0;
}

char *adapterAlignment(char *readSeq, char *adapterSeq, int ma, int mi, int go, int ge) {
    g_err.clear();
    const int64_t n = readSeq ? (int64_t)strlen(readSeq) : 0;
    const int64_t m = adapterSeq ? (int64_t)strlen(adapterSeq) : 0;
    int32_t rec[PB_REC];
    if (n == 0 || m == 0) {
        rec[0] = -1; rec[1] = 0; rec[2] = -1; rec[3] = 0; rec[4] = PB_SCORE_EMPTY; rec[5] = rec[6] = rec[7] = rec[8] = 0;
    } else {
        int64_t soff[2] = {0, n};
        int32_t aoff[2] = {0, (int32_t)m};
        int rc = batch_host((const uint8_t *)readSeq, soff, 1, (const uint8_t *)adapterSeq, aoff, 1, nullptr, nullptr, 1, ma, mi,
                            go, ge, rec);
        if (rc != 0) {
            fprintf(stderr, "porechop_b200: adapterAlignment failed (%d): %s\n", rc, g_err.c_str());
            return nullptr;
        }
    }
    char *buf = (char *)malloc(96);
    if (!buf) return nullptr;
    if (pb200FormatRecord(rec, buf, 96) < 0) { free(buf); return nullptr; }
    return buf;
}

void freeCString(char *p) { free(p); }

int pb200DeviceCount(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int pb200SetDevice(int device) {
    g_err.clear();
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(PB200_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return 0;
}
const char *pb200LastError(void) { return g_err.c_str(); }
long long pb200KernelLaunches(void) { return g_launches.load(); }
void pb200TimingEnable(int on) { g_timing.store(on ? 1 : 0); }

int pb200TimingReadKinds(double *ms, long long *launches, double *window_cells, int reset) {
    Engine *Ep = nullptr;
    if (int rc = get_engine(&Ep)) return rc;
    Engine &E = *Ep;
    std::lock_guard<std::mutex> lk(E.mu);
    for (auto &tl : E.timed) {
        float f = 0.f;
        if (cudaEventSynchronize(tl.b) == cudaSuccess && cudaEventElapsedTime(&f, tl.a, tl.b) == cudaSuccess) {
            E.timed_ms_acc[tl.kind] += f; E.timed_n_acc[tl.kind]++;
        }
        cudaEventDestroy(tl.a); cudaEventDestroy(tl.b);
    }
    E.timed.clear();
    for (int k = 0; k < TK_N; ++k) {
        if (ms) ms[k] = E.timed_ms_acc[k];
        if (launches) launches[k] = E.timed_n_acc[k];
    }
    unsigned long long wc = 0;
    if (E.wcells.p) {
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpyAsync(&wc, E.wcells.p, 8, cudaMemcpyDeviceToHost, E.st[0].stream));
        if (reset) CK(cudaMemsetAsync(E.wcells.p, 0, 8, E.st[0].stream));
        CK(cudaStreamSynchronize(E.st[0].stream));
    }
    if (window_cells) *window_cells = (double)wc;
    if (reset) for (int k = 0; k < TK_N; ++k) { E.timed_ms_acc[k] = 0; E.timed_n_acc[k] = 0; }
    return 0;
}

int pb200TimingRead(double *ms, long long *launches, double *cells, int reset) {
    double m[TK_N]; long long n[TK_N]; double wc = 0;
    if (int rc = pb200TimingReadKinds(m, n, &wc, reset)) return rc;
    if (ms) { *ms = 0; for (int k = 0; k < TK_N; ++k) *ms += m[k]; }
    if (launches) { *launches = 0; for (int k = 0; k < TK_N; ++k) *launches += n[k]; }
    if (cells) *cells = wc;
    return 0;
}

int pb200Synchronize(void) {
    g_err.clear();
    Engine *Ep = nullptr;
    if (int rc = get_engine(&Ep)) return rc;
    Engine &E = *Ep;
    std::lock_guard<std::mutex> lk(E.mu);
    if (!E.init_done) return 0;
    CK(cudaDeviceSynchronize());
    int rc_final = 0;
    for (int i = 0; i < NSTAGE; ++i) {
        if (!E.st[i].misc.p) continue;
        if (int rc = check_status(E.st[i], E.st[i].stream)) rc_final = rc;
        CK(cudaMemsetAsync(E.st[i].misc.p, 0, 4, E.st[i].stream));
        CK(cudaStreamSynchronize(E.st[i].stream));
    }
    return rc_final;
}

// Pinned staging memory for host callers (fastq.py gathers the trimmed reads of a chunk before the middle scan): `slot` 0..3,
// grown on demand, owned by the library, valid until the next call for the same slot.  NULL without a device (callers then
// use ordinary memory).  A pinned source makes the engine's uploads asynchronous DMA at PCIe speed and is reused chunk after
// chunk instead of page-faulting in a fresh gigabyte every time.
void *pb200HostBuffer(int slot, size_t bytes) {
    static HostBuf bufs[4];
    static std::mutex mu;
    if (slot < 0 || slot >= 4) return nullptr;
    Engine *Ep = nullptr;
    if (get_engine(&Ep)) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (bufs[slot].ensure(bytes ? bytes : 1)) return nullptr;
    return bufs[slot].p;
}

int pb200GetOption(const char *name) {
    load_env_options();
    if (!name) return -1;
    if (!strcmp(name, "h2d_pack")) return g_opt.h2d_pack;
    if (!strcmp(name, "h2d_pack_large_submit")) return pack_wanted(1ll << 40) ? 1 : 0;   // what "auto" resolves to for a large submit
    if (!strcmp(name, "pack_threads")) return g_opt.pack_threads;
    if (!strcmp(name, "tight_window")) return g_opt.tight_window;
    if (!strcmp(name, "profile")) return g_opt.profile;
    if (!strcmp(name, "scratch_mb")) return g_opt.scratch_mb;
    if (!strcmp(name, "hbuf")) return g_opt.hbuf_mode;
    if (!strcmp(name, "direct_max")) return (int)std::min<int64_t>(g_opt.direct_max, INT_MAX);
    if (!strcmp(name, "chunk_tasks")) return (int)std::min<int64_t>(g_opt.chunk_tasks, INT_MAX);
    return -1;
}

int pb200SetOption(const char *name, const char *value) {
    load_env_options();
    if (!name || !value) return PB200_ERR_ARG;
    if (!strcmp(name, "direct_max")) g_opt.direct_max = atoll(value);
    else if (!strcmp(name, "chunk_tasks")) g_opt.chunk_tasks = std::max(1ll, atoll(value));
    else if (!strcmp(name, "scratch_mb")) g_opt.scratch_mb = std::max(1, atoi(value));
    else if (!strcmp(name, "tight_window")) g_opt.tight_window = atoi(value);
    else if (!strcmp(name, "h2d_pack")) g_opt.h2d_pack = atoi(value);
    else if (!strcmp(name, "profile")) g_opt.profile = atoi(value);
    else if (!strcmp(name, "pack_threads")) { if (atoi(value) > 0) g_opt.pack_threads = atoi(value); }
    else if (!strcmp(name, "hbuf")) g_opt.hbuf_mode = !strcmp(value, "smem") ? 1 : !strcmp(value, "global") ? 2 : 0;
    else return PB200_ERR_ARG;
    return 0;
}

}  // extern "C"
