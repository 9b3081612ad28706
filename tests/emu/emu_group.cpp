// tests/emu/emu_group.cpp -- CPU emulation of one sub-warp group of the CUDA kernels (TEST ONLY).
//
// Runs the *product's own* device code of porechop_b200/csrc/dp_core.cuh (lane_init / lane_step /
// lane_track / scout_combine / traceback_stats compile as plain C++ here) lane by lane, replacing only the
// warp shuffle and the shared-memory trace addressing of kernels.cuh.  This lets the no-GPU test tier check
// the wavefront indexing, the 4-bit trace packing, the scout, the traceback and the statistics -- including
// the two-pass (score pass + bounded window) scheme -- against the oracle.  It is not a CPU fallback: the
// product never builds or loads this file.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#define PB_CHECK_RANGES 1
#include "../../porechop_b200/csrc/dp_core.cuh"

static long g_range_violations = 0;
namespace pb { void pb_range_violation() { ++g_range_violations; } }

using namespace pb;

namespace {

struct Half {
    std::vector<uint8_t> seq, ad;   // encoded
    Task t;
};

template <int R>
void run_trace_group(int G, const Half &A, const Half &B, const Scoring &sc, int32_t *recA, int32_t *recB, int *status) {
    const int WPS = TraceWords<R>::value;
    const Task &tA = A.t, &tB = B.t;
    HalfGeom gA = make_geom(tA.n, tA.m, G, R), gB = make_geom(tB.n, tB.m, G, R);
    const bool emptyA = tA.n <= 0 || tA.m <= 0, emptyB = tB.n <= 0 || tB.m <= 0;
    const int nmin = emptyA ? tB.n : (emptyB ? tA.n : std::min(tA.n, tB.n));
    const int nmax = std::max(tA.n, tB.n);
    const int T = nmax > 0 ? nmax + G - 1 : 0;
    const uint8_t *seqA = A.seq.data() + tA.seq_off, *seqB = B.seq.data() + tB.seq_off;
    const uint8_t *adA = A.ad.data() + tA.ad_off, *adB = B.ad.data() + tB.ad_off;
    std::vector<uint32_t> hbuf((size_t)std::max(nmax, 1));
    for (int c = 0; c < nmax; ++c) {
        uint32_t bA = c < tA.n ? seqA[c] : PB_PAD_H, bB = c < tB.n ? seqB[c] : PB_PAD_H;
        hbuf[c] = (bA << 8) | (bB << 24);
    }
    std::vector<Lane<R>> L((size_t)G);
    for (int g = 0; g < G; ++g)
        lane_init<R>(L[g], g, G, sc, adA, tA.m, (tA.flags & TASK_LEFT_INF) != 0, adB, tB.m, (tB.flags & TASK_LEFT_INF) != 0);
    std::vector<uint32_t> tr((size_t)std::max(T, 1) * WPS * G, 0u);
    std::vector<uint32_t> sS((size_t)G), sV((size_t)G);
    for (int t = 0; t < T; ++t) {
        for (int g = 0; g < G; ++g) { sS[g] = L[g].botX; sV[g] = L[g].botV; }   // shuffle snapshot
        for (int g = 0; g < G; ++g) {
            uint32_t recvS = g ? sS[g - 1] : sc.borderX2, recvV = g ? sV[g - 1] : sc.negb2;
            int j = t - g + 1;
            if (j >= 1 && j <= nmax) {
                uint32_t tw[WPS];
                if (j >= nmin) {
                    uint32_t vr[R];
                    lane_step<R, true, true>(L[g], recvS, recvV, hbuf[j - 1], sc, tw, vr);
                    lane_track_general<R>(L[g], g, j, gA, gB, vr, sc);
                } else {
                    lane_step<R, true, false>(L[g], recvS, recvV, hbuf[j - 1], sc, tw);
                    lane_track_lastrow<R>(L[g], j, sc);
                }
                for (int w = 0; w < WPS; ++w) tr[((size_t)t * WPS + w) * G + g] = tw[w];
            }
        }
    }
    for (int h = 0; h < 2; ++h) {
        const Task &tk = h ? tB : tA;
        const HalfGeom &gh = h ? gB : gA;
        int32_t *rec = h ? recB : recA;
        if (tk.out_idx < 0) continue;
        std::vector<ScoutCand> cand((size_t)G);
        for (int g = 0; g < G; ++g) cand[g] = make_cand<R>(L[g], h, sc);
        EndCell end;
        if (tk.flags & TASK_END_GIVEN) {
            end.j = tk.end_j; end.i = tk.end_i; end.score = tk.end_score; end.corr = tk.end_corr;
            if (tk.n_total <= 0 || tk.m <= 0) end.score = PB_SCORE_EMPTY;
        } else {
            end = scout_combine(cand.data(), G, gh);
        }
        const uint8_t *sq = h ? seqB : seqA;
        const uint8_t *ad = h ? adB : adA;
        auto nib = [&](int jl, int i) -> uint32_t {
            const int q = i + gh.pad - 1;
            const int gg = q / R, r = q % R;
            const int t = jl - 1 + gg;
            const uint32_t w = tr[((size_t)t * WPS + trace_word<R>(h, r)) * G + gg];
            return (w >> trace_shift<R>(h, r)) & 15u;
        };
        auto eq = [&](int jl, int i) -> bool { return sq[jl - 1] == ad[i - 1]; };
        int st = traceback_stats(nib, eq, end, sc.linear != 0, tk.col0, tk.n_total, tk.m, rec);
        if (st) *status |= 1;
    }
}

// PROF: the query-profile variant of score_kernel -- per lane a table [6 base codes][R rows] of substitution operands built
// with profile_word, the step fetches its R operands by the base code of the column (both halves read sequence A).
template <int R, bool PROF = false>
void run_score_group(int G, const Half &A, const Half &B, const Scoring &sc, EndCell *eA, EndCell *eB) {
    const Task &tA = A.t, &tB = B.t;
    HalfGeom gA = make_geom(tA.n, tA.m, G, R), gB = make_geom(tB.n, tB.m, G, R);
    const bool emptyA = tA.n <= 0 || tA.m <= 0, emptyB = tB.n <= 0 || tB.m <= 0;
    const int nmin = emptyA ? tB.n : (emptyB ? tA.n : std::min(tA.n, tB.n));
    const int nmax = std::max(tA.n, tB.n);
    const int T = nmax + G - 1;
    const uint8_t *seqA = A.seq.data() + tA.seq_off, *seqB = B.seq.data() + tB.seq_off;
    std::vector<Lane<R>> L((size_t)G);
    for (int g = 0; g < G; ++g)
        lane_init<R>(L[g], g, G, sc, A.ad.data() + tA.ad_off, tA.m, false, B.ad.data() + tB.ad_off, tB.m, false);
    std::vector<uint32_t> prof((size_t)G * 6 * R);           // [lane][base code][row]
    if (PROF)
        for (int g = 0; g < G; ++g)
            for (int b = 0; b < 6; ++b)
                for (int r = 0; r < R; ++r)
                    prof[((size_t)g * 6 + b) * R + r] = PB_PROF_ENCODE(profile_word(g * R + r + 1, (uint32_t)b, sc, A.ad.data() + tA.ad_off,
                                                                                    tA.m, gA.pad, B.ad.data() + tB.ad_off, tB.m, gB.pad));
    std::vector<uint32_t> sS((size_t)G), sV((size_t)G);
    for (int t = 0; t < T; ++t) {
        for (int g = 0; g < G; ++g) { sS[g] = L[g].botX; sV[g] = L[g].botV; }
        for (int g = 0; g < G; ++g) {
            uint32_t recvS = g ? sS[g - 1] : sc.borderX2, recvV = g ? sV[g - 1] : sc.negb2;
            int j = t - g + 1;
            if (j >= 1 && j <= nmax) {
                int ja = std::min(j, tA.n) - 1, jb = std::min(j, tB.n) - 1;
                uint32_t bA = ja >= 0 ? seqA[ja] : PB_PAD_H, bB = jb >= 0 ? seqB[jb] : PB_PAD_H;
                if (j > tA.n) bA = PB_PAD_H;
                if (j > tB.n) bB = PB_PAD_H;
                const uint32_t *subs = PROF ? &prof[((size_t)g * 6 + (bA >> 4)) * R] : nullptr;    // kernel: ring holds (bA >> 4) * ROWS
                if (j < nmin) {
                    lane_step<R, false, false, PROF>(L[g], recvS, recvV, (bA << 8) | (bB << 24), sc, nullptr, nullptr, subs);
                    lane_track_lastrow<R>(L[g], j, sc);
                } else {
                    uint32_t vr[R];
                    lane_step<R, false, true, PROF>(L[g], recvS, recvV, (bA << 8) | (bB << 24), sc, nullptr, vr, subs);
                    lane_track_general<R>(L[g], g, j, gA, gB, vr, sc);
                }
            }
        }
    }
    for (int h = 0; h < 2; ++h) {
        std::vector<ScoutCand> cand((size_t)G);
        for (int g = 0; g < G; ++g) cand[g] = make_cand<R>(L[g], h, sc);
        *(h ? eB : eA) = scout_combine(cand.data(), G, h ? gB : gA);
    }
}

Half make_half(const char *seq, int n, const char *ad, int m, int out_idx) {
    Half H;
    H.seq.resize((size_t)std::max(n, 1)); H.ad.resize((size_t)std::max(m, 1));
    for (int k = 0; k < n; ++k) H.seq[k] = (uint8_t)encode_byte((uint8_t)seq[k]);
    for (int k = 0; k < m; ++k) H.ad[k] = (uint8_t)encode_byte((uint8_t)ad[k]);
    Task &t = H.t;
    memset(&t, 0, sizeof t);
    t.seq_off = 0; t.n = n; t.m = m; t.ad_off = 0; t.out_idx = out_idx; t.n_total = n;
    return H;
}

// same transformation as window_tasks_kernel (kernels.cuh)
void to_window(Task &t, const EndCell &e, int wnum, int wden, bool tight) {
    if (t.n > 0 && t.m > 0) {
        int64_t W = window_cols(t.m, e.i, e.score, wnum, wden, tight);
        int64_t c0 = (int64_t)e.j - W;
        if (c0 < 0) c0 = 0;
        t.col0 = (int32_t)c0; t.seq_off += c0; t.n = e.j - (int32_t)c0;
        t.flags = TASK_END_GIVEN | (c0 > 0 ? TASK_LEFT_INF : 0);
        t.end_j = t.n; t.end_i = e.i; t.end_corr = e.corr; t.end_score = e.score;
    } else {
        t.n = 0; t.flags = TASK_END_GIVEN; t.end_j = 0; t.end_i = 0; t.end_corr = 0; t.end_score = PB_SCORE_EMPTY;
    }
}

}  // namespace

extern "C" {

// mode 0: single trace pass; mode 1: score pass + windowed trace pass (W = m + m*wnum/wden); mode | 4: the windows use the
// per-alignment bound (dp_core.cuh window_cols, tight); mode | 8: query-profile score pass (R = 8).
// G in {4,8,16,32}, R in {4..8}; pass nB < 0 to leave half B empty.  Returns the status bit (window violated).
int emu_align_slot(const char *seqA, int nA, const char *adA, int mA, const char *seqB, int nB, const char *adB, int mB,
                   int G, int R, int mode, int ma, int mi, int go, int ge, int wnum, int wden, int32_t *recA,
                   int32_t *recB) {
    Scoring sc = make_scoring(ma, mi, go, ge);
    Half A = make_half(seqA, nA, adA, mA, 0);
    Half B = nB >= 0 ? make_half(seqB, nB, adB, mB, 1) : make_half("", 0, "", 0, -1);
    int status = 0;
    const bool tight = (mode & 4) != 0;
    const bool prof = (mode & 8) != 0;            // query-profile score pass (R = 8, both halves the same read)
    mode &= 3;
    if (mode >= 1) {
        EndCell eA, eB;
        if (prof && R == 8) {
            run_score_group<8, true>(G, A, B, sc, &eA, &eB);
        } else {
            switch (R) {
                case 5: run_score_group<5>(G, A, B, sc, &eA, &eB); break;
                case 6: run_score_group<6>(G, A, B, sc, &eA, &eB); break;
                case 7: run_score_group<7>(G, A, B, sc, &eA, &eB); break;
                case 8: run_score_group<8>(G, A, B, sc, &eA, &eB); break;
                default: run_score_group<4>(G, A, B, sc, &eA, &eB); break;
            }
        }
        to_window(A.t, eA, wnum, wden, tight);
        if (nB >= 0) to_window(B.t, eB, wnum, wden, tight);
        // the trace pass may use a different (G,R) than the score pass, as in the engine: keep R, G as given
    }
    switch (R) {
        case 5: run_trace_group<5>(G, A, B, sc, recA, recB, &status); break;
        case 6: run_trace_group<6>(G, A, B, sc, recA, recB, &status); break;
        case 7: run_trace_group<7>(G, A, B, sc, recA, recB, &status); break;
        case 8: run_trace_group<8>(G, A, B, sc, recA, recB, &status); break;
        default: run_trace_group<4>(G, A, B, sc, recA, recB, &status); break;
    }
    return status | (g_range_violations ? 2 : 0);
}

long emu_range_violations() { return g_range_violations; }

// unpack_kernel (kernels.cuh) run serially: the same per-thread body (16 output bytes from 8 packed bytes through
// unpack_nibbles8, byte-wise tail through unpack_nibble1) over the whole buffer.
void emu_unpack(const uint8_t *in, uint8_t *out, int64_t n) {
    for (int64_t i = 0; i < n; i += 16) {
        if (i + 16 <= n) {
            uint32_t p[2], v[4];
            memcpy(p, in + (i >> 1), 8);
            unpack_nibbles8(p[0], v[0], v[1]);
            unpack_nibbles8(p[1], v[2], v[3]);
            memcpy(out + i, v, 16);
        } else {
            for (int64_t k = i; k < n && k < i + 16; ++k) out[k] = (uint8_t)unpack_nibble1(in[k >> 1], (int)(k & 1));
        }
    }
}

// decide_kernel (kernels.cuh) run serially: the same per-record core functions (dp_core.cuh end_trim_candidate /
// score_pair), max over the adapters of a read, pairs of the listed columns.  Returns the overflow flag.
int emu_decide(const int32_t *records, int64_t n, int32_t n_adapters, int is_start, int32_t end_size, int32_t extra_trim,
               int32_t min_trim, const int32_t *cmin, int32_t cmin_len, const int32_t *cols, int32_t n_cols, int32_t *trim,
               uint32_t *pairs) {
    int ovf = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t *rec = records + (size_t)i * n_adapters * PB_REC;
        int32_t best = 0;
        for (int32_t k = 0; k < n_adapters; ++k) {
            const int32_t t = end_trim_candidate(rec + (size_t)k * PB_REC, is_start, end_size, extra_trim, min_trim, cmin, cmin_len, &ovf);
            best = t > best ? t : best;
        }
        trim[i] = best;
        for (int32_t k = 0; k < n_cols; ++k) pairs[(size_t)i * n_cols + k] = score_pair(rec + (size_t)cols[k] * PB_REC, &ovf);
    }
    return ovf;
}

// encode_kernel's per-byte mapping, for comparison
void emu_encode(const uint8_t *in, uint8_t *out, int64_t n) {
    for (int64_t k = 0; k < n; ++k) out[k] = (uint8_t)encode_byte(in[k]);
}

}  // extern "C"
