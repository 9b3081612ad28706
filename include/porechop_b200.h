/*
 * include/porechop_b200.h -- C-ABI of the B200 adapter-alignment engine (cpp_functions.so).
 *
 * Drop-in boundary: the library is loaded by ctypes exactly like the reference's
 * porechop/cpp_functions.so (porechop/cpp_function_wrappers.py:21-39) and exports the reference's two
 * symbols with the same signatures, ownership and string format, plus a batched entry point so that the
 * host can submit every (read window, adapter) / (full read, adapter) pair in one call.
 *
 * Plain C types only (no torch / CUDA types): pointers, sizes, ints.  Every call runs on the CUDA
 * device that is current for the calling thread (cudaSetDevice / torch.cuda.set_device /
 * pb200SetDevice); there is NO CPU fallback -- without a usable sm_100 device the batch calls return
 * PB200_ERR_NO_DEVICE and adapterAlignment() returns NULL after printing the reason to stderr.
 */
#ifndef PORECHOP_B200_H
#define PORECHOP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reference ABI, unchanged -------------------------------------------------------------------
 * replaces porechop/include/adapter_align.h:13-14 (implemented in porechop/src/adapter_align.cpp:11-31).
 * readSeq / adapterSeq: NUL-terminated ASCII, borrowed for the call.  Returns a malloc'd NUL-terminated
 * string "readStart,readEnd,adapterStart,adapterEnd,rawScore,alignedRegion%ID,fullAdapter%ID"
 * (ints "%d", doubles "%f"; ends are inclusive 0-based; "-1,0,-1,0,-2147483648,0.000000,0.000000" for an
 * empty read or adapter).  Ownership passes to the caller, who releases it with freeCString().
 * Thread-safe (calls are serialised on the device). */
char *adapterAlignment(char *readSeq, char *adapterSeq, int matchScore, int mismatchScore, int gapOpenScore,
                       int gapExtensionScore);

/* replaces porechop/include/adapter_align.h:15 (porechop/src/adapter_align.cpp:34-36): free(p). */
void freeCString(char *p);

/* ---- batched entry points (new; SURVEY.md 8(b) "new batch export") --------------------------------
 * One record of 9 int32 per alignment:
 *   {readStart, readEnd, adapterStart, adapterEnd, rawScore, matchAligned, lenAligned, matchAdapter, lenAdapter}
 * alignedRegion%ID = 100.0*matchAligned/lenAligned and fullAdapter%ID = 100.0*matchAdapter/lenAdapter in
 * double, exactly the two divisions of porechop/src/alignment.cpp:82,90; an empty read or adapter gives
 * {-1,0,-1,0,INT32_MIN,0,0,0,0}.  pb200FormatRecord() turns a record into the reference string.
 *
 * seqs/seq_off   : concatenated ASCII reads (or read windows); sequence s is seqs[seq_off[s] .. seq_off[s+1])
 * adapters/ad_off: concatenated ASCII adapters, same convention (int32 offsets)
 * pair_seq/pair_adapter (n_pairs each): the alignments to compute; both NULL = the full cross product in
 *                  sequence-major order (n_pairs must equal n_seqs*n_adapters, record p = s*n_adapters + a)
 * out            : n_pairs * 9 int32, caller-owned
 * Returns 0 on success or a PB200_ERR_* code (pb200LastError() has the text).  Host pointers; pinned host
 * memory lets the internal host<->device copies overlap with the kernels. */
int adapterAlignmentBatch(const uint8_t *seqs, const int64_t *seq_off, int64_t n_seqs, const uint8_t *adapters,
                          const int32_t *ad_off, int32_t n_adapters, const int32_t *pair_seq,
                          const int32_t *pair_adapter, int64_t n_pairs, int matchScore, int mismatchScore,
                          int gapOpenScore, int gapExtensionScore, int32_t *out);

/* Several independent cross-product batches in ONE submit -- e.g. the two batches of Porechop's end-trim phase
 * (find_start_trim: start windows x start adapters, find_end_trim: end windows x end adapters,
 * porechop/nanopore_read.py:166-208) -- pipelined through the same ring of streams, so the upload of the second batch
 * overlaps the kernels of the first and the copy pipeline fills and drains once per submit instead of once per batch.
 * Each batch has the cross-product semantics of adapterAlignmentBatch (records in sequence-major order in its own
 * `out`); batches with no sequences or no adapters are skipped.  Same scoring scheme for all batches. */
typedef struct {
    const uint8_t *seqs; const int64_t *seq_off; int64_t n_seqs;        /* as in adapterAlignmentBatch */
    const uint8_t *adapters; const int32_t *ad_off; int32_t n_adapters;
    int32_t *out;                                                       /* n_seqs * n_adapters * 9 int32 */
} pb200_batch_t;
int adapterAlignmentBatchMulti(const pb200_batch_t *batches, int n_batches, int matchScore, int mismatchScore,
                               int gapOpenScore, int gapExtensionScore);

/* End-trim DECISIONS on the device.  Each batch is one cross product of read-end windows x
 * adapters (as in adapterAlignmentBatchMulti); the 9-int records stay on the device and a second kernel reduces them per
 * read to what Porechop's host logic consumes:
 *   trim[s]         find_start_trim / find_end_trim (porechop/nanopore_read.py:166-208): the largest trim amount over the
 *                   adapters whose alignment passes `aligned-region identity > end_threshold`, the end_size / 0 edge test
 *                   and `read_end - read_start >= min_trim_size`; 0 when none passes
 *   score_pairs     for every adapter index listed in score_cols (the barcode adapters, nanopore_read.py:181-183,
 *                   203-205): (matchAdapter, lenAdapter) as two uint16 -- full-adapter identity =
 *                   float("%f" % (100.0 * matchAdapter / lenAdapter)) on the host; a failed alignment gives (0, 1) = 0.0
 * so 4 + 4*n_score_cols bytes per read come back instead of 36 per alignment.  The identity test is exact: the host
 * builds, with the reference's own snprintf("%f") / strtod chain, the smallest passing match count per aligned length
 * (pb200TrimThresholdTable) and the kernel compares integers.  end_threshold must be >= 0.  batch.out may be NULL (records
 * not copied back) or a buffer for the full records. */
typedef struct {
    pb200_batch_t batch;
    int32_t is_start;                 /* 1 = find_start_trim rule (start windows), 0 = find_end_trim rule (end windows) */
    int32_t end_size;                 /* --end_size (porechop.py:130), the window length the edge tests refer to */
    int32_t extra_trim_size;          /* --extra_end_trim */
    int32_t min_trim_size;            /* --min_trim_size */
    double end_threshold;             /* --end_threshold */
    const int32_t *score_cols;        /* adapter indices whose full-adapter identity is wanted (may be NULL if none) */
    int32_t n_score_cols;
    int32_t *trim;                    /* out: n_seqs */
    uint16_t *score_pairs;            /* out: n_seqs * n_score_cols * 2; may be NULL when only top2 is wanted */
    int32_t *top2;                    /* out, optional (NULL = not wanted): n_seqs * 6 = {position, matchAdapter, lenAdapter}
                                       * of the best and of the second-best score column -- the first two entries of
                                       * determine_barcode's `sorted(scores.items(), reverse=True, key=score)`
                                       * (porechop/nanopore_read.py:404-415): highest full-adapter identity first, equal
                                       * identities in score_cols order; `position` indexes score_cols (-1, 0, 1 when
                                       * there are fewer columns).  The caller lists each barcode NAME once in score_cols
                                       * (the reference's dict keeps a repeated name's last value).  24 bytes per read
                                       * instead of 4*n_score_cols. */
} pb200_end_batch_t;
int adapterEndDecisions(const pb200_end_batch_t *batches, int n_batches, int matchScore, int mismatchScore,
                        int gapOpenScore, int gapExtensionScore);
/* cmin[l], l = 0 .. len-1: the smallest match count c for which float("%f" % (100.0*c/l)) > end_threshold, or l+1 if
 * none does (cmin[0] = INT32_MAX: 0/0 is NaN and never passes).  Pure host code. */
int pb200TrimThresholdTable(double end_threshold, int32_t len, int32_t *cmin);

/* Same, with the bulk data already resident in device memory (d_seqs, d_seq_off, d_out are device pointers on
 * the current device; adapters/ad_off stay host pointers -- a few KB).  Cross-product mode only.
 * max_seq_len: length of the longest sequence, or -1 to let the library compute it on the device.
 * The work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the library's own stream) and the
 * call returns after enqueueing everything unless the score pass needs a host decision; call
 * cudaStreamSynchronize / torch.cuda.synchronize before reading d_out. */
int adapterAlignmentBatchDevice(const uint8_t *d_seqs, const int64_t *d_seq_off, int64_t n_seqs,
                                int64_t total_seq_bytes, int64_t max_seq_len, const uint8_t *adapters,
                                const int32_t *ad_off, int32_t n_adapters, int matchScore, int mismatchScore,
                                int gapOpenScore, int gapExtensionScore, int32_t *d_out, void *stream);

/* Format one 9-int record as the reference string (alignment.cpp:113-121). Returns strlen, or -1 if buflen is
 * too small (64 bytes always suffice). */
int pb200FormatRecord(const int32_t *record, char *buf, int buflen);

/* Host-side Dna5 conversion of the packed upload path (option "h2d_pack"): n ASCII bases -> (n+1)/2 bytes, the code
 * of base 2k (A/a=0 C/c=1 G/g=2 T/t/U/u=3, anything else 4: seqan/basic/alphabet_residue_tabs.h:113-140) in the low
 * nibble of byte k and base 2k+1 in the high nibble.  Pure host code (AVX-512BW / AVX2 when the CPU has it, its own thread
 * team; threads <= 0 = the hardware threads the cgroup CPU quota allows); exported so that tests and hosts that already hold
 * packed reads can use the same packer. */
int pb200PackNibbles(const uint8_t *ascii, int64_t n, uint8_t *packed, int threads);

/* ---- device / diagnostics ------------------------------------------------------------------------------ */
int pb200DeviceCount(void);               /* number of CUDA devices visible (0 if none / no driver) */
int pb200SetDevice(int device);           /* cudaSetDevice for the calling thread */
int pb200Synchronize(void);               /* wait for everything this library enqueued on the current device and
                                             report deferred errors of adapterAlignmentBatchDevice calls */
const char *pb200LastError(void);         /* text of the last error on this thread ("" if none) */
long long pb200KernelLaunches(void);      /* kernels launched by this library since load (all threads) */
/* Kernel timing with CUDA events on the launching stream: enable, run, then read back the accumulated
 * duration and launch count of the DP kernels (trace_kernel + score_kernel) since the last reset. */
void pb200TimingEnable(int on);
int pb200TimingRead(double *dp_kernel_ms, long long *dp_kernel_launches, double *cells, int reset);
/* The same per kind of DP launch, arrays of PB200_TIMING_KINDS entries: 0 trace_kernel (windows, single pass),
 * 1 unused, 2 trace_kernel on the bounded windows of a two-pass class, 3 score_kernel (long reads).
 * window_cells = DP cells computed by the launches of kind 2 (the other kinds sweep every cell of their batch). */
#define PB200_TIMING_KINDS 4
int pb200TimingReadKinds(double *ms, long long *launches, double *window_cells, int reset);
/* Tunables (also read from the environment at first use, PB200_<NAME>); none of them changes a result:
 *   "direct_max"  longest sequence aligned in one pass (default 160; longer ones take score pass + bounded window)
 *   "chunk_tasks" alignments per pipeline chunk of the host-buffer API (default 131072)
 *   "scratch_mb"  cap on the resident trace scratch in MB (default 128; 72 keeps it L2-resident)
 *   "hbuf"        staging of a slot's packed bases: "auto" | "smem" | "global"
 *   "h2d_pack"    1 = the host-buffer calls convert the sequences to 4-bit codes on the host cores (a packer thread that
 *                 runs ahead of the submit loop) and upload half the bytes; 0 (default) = never; -1 = auto: submits of at
 *                 least 32 MB when the packer team has 12 or more threads; "pack_threads" = host threads of the packer
 *                 (default = the hardware threads the cgroup CPU quota allows / LOCAL_WORLD_SIZE, minus two for the submit
 *                 and driver threads, at most 32)
 *   "profile"     1 (default) = the long-read score pass fetches its substitution operands from a query profile in shared
 *                 memory when every slot is one read x two adapters (cross-product mode); 0 = always computed
 *   "tight_window" 1 (default) = second-pass windows sized per alignment from the end cell's row and score; 0 = the
 *                 per-adapter worst case
 * (round 2 measured and removed "short2p", "rowoff" and the trace-kernel profiles: profiles/r2_options) */
int pb200SetOption(const char *name, const char *value);
/* Current value of an integer tunable ("h2d_pack", "pack_threads", "tight_window", "profile", "scratch_mb", "direct_max",
 * "chunk_tasks", "hbuf" as 0/1/2), or "h2d_pack_large_submit": what h2d_pack = auto resolves to for a large submit on this
 * host (1 = packed).  -1 for an unknown name. */
int pb200GetOption(const char *name);
/* Pinned (page-locked) staging memory for host callers that assemble a batch before submitting it: slot 0..3, at least `bytes`
 * long, owned by the library and valid until the next call for the same slot; NULL when there is no device.  Uploads from it
 * are asynchronous DMA; the buffer is reused call after call. */
void *pb200HostBuffer(int slot, size_t bytes);

enum {
    PB200_OK = 0,
    PB200_ERR_NO_DEVICE = 100,   /* no CUDA device / driver, or device is not sm_100 */
    PB200_ERR_CUDA = 101,        /* a CUDA call failed (see pb200LastError) */
    PB200_ERR_ARG = 102,         /* invalid argument */
    PB200_ERR_INTERNAL = 103     /* internal invariant violated (e.g. window bound) */
};

#ifdef __cplusplus
}
#endif
#endif /* PORECHOP_B200_H */
