/*
 * porechop_b200_io.h -- host-side FASTQ ingest / emit helpers (plain C, OpenMP; no CUDA, no torch types).
 *
 * the reference keeps every read as Python objects (porechop/misc.py:151-168 load_fastq,
 * porechop/nanopore_read.py:23-55 NanoporeRead.__init__, :97-147 get_fasta/get_fastq).  These entry points do the same
 * byte work on flat buffers so that 10^6-10^7 reads reach the alignment engine (include/porechop_b200.h) and leave it
 * without per-read host objects.  Library: porechop_b200/libhostio.so (built by porechop_b200/build.py with gcc).
 * Every function is a pure function of its arguments; buffers are caller-owned.
 */
#ifndef PORECHOP_B200_IO_H
#define PORECHOP_B200_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBIO_OK 0
#define PBIO_ERR_RECORDS 1   /* not a whole number of 4-line records */
#define PBIO_ERR_HEADER 2    /* a record does not start with '@' (after stripping) */

/* worker threads of the functions below (n <= 0: leave unchanged); returns the current maximum.  torchrun exports
 * OMP_NUM_THREADS=1 to its workers, so a multi-process run sets cores / local_world_size here instead. */
int pbioSetThreads(int n);

/* number of lines of buf[0..n): '\n' terminated, plus one if the last byte is not '\n' (misc.py:158 iterates lines) */
int64_t pbioCountLines(const uint8_t *buf, int64_t n);

/* line_end[k] = index of the '\n' ending line k (or n for an unterminated last line); n_lines from pbioCountLines */
int pbioLineEnds(const uint8_t *buf, int64_t n, int64_t *line_end, int64_t n_lines);

/* str.strip() of every line: span_a[k], span_len[k] = the stripped extent of line k (misc.py:139, 160-166) */
void pbioLineSpans(const uint8_t *buf, const int64_t *line_end, int64_t n_lines, int64_t *span_a, int64_t *span_len);

/* 4-line records (misc.py:158-166): every line stripped like str.strip(); name = header minus its first character.
 * Writes the start and length of name / bases / qualities of each of the n_lines/4 records. */
int pbioFastqIndex(const uint8_t *buf, int64_t n, const int64_t *line_end, int64_t n_lines,
                   int64_t *name_a, int64_t *name_len, int64_t *seq_a, int64_t *seq_len,
                   int64_t *qual_a, int64_t *qual_len);

/* dst[dst_off[i] .. dst_off[i+1]) = src[src_a[i] .. src_a[i]+src_len[i]) followed by `fill` bytes (quality padding,
 * nanopore_read.py:34-36); src_len may be NULL = exactly the destination length. */
void pbioGather(uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_a,
                const int64_t *src_len, int fill, int64_t n);

/* NanoporeRead.__init__ (nanopore_read.py:26-31) on seq[off[i]..off[i+1]): upper-case; if count('U') > count('T')
 * the read is RNA: rna[i] = 1 and every U becomes T. */
void pbioNormalise(uint8_t *seq, const int64_t *off, int64_t n, uint8_t *rna);

/* get_fastq / get_fasta (nanopore_read.py:97-147) of n_rec output records into out[out_off[r] .. out_off[r+1]):
 *   fmt 0 (FASTQ): '@' name '\n' bases '\n+\n' qualities '\n'        length 1+nlen+1+slen+3+qlen+1
 *   fmt 1 (FASTA): '>' name '\n' bases wrapped at 70 columns, every line '\n' terminated (misc.py:327-338)
 *                                                                    length 1+nlen+1+slen+ceil(slen/70)
 * rna[r] != 0 writes T as U (nanopore_read.py:107,133). */
void pbioEmit(uint8_t *out, const int64_t *out_off, int64_t n_rec, int fmt,
              const uint8_t *names, const int64_t *name_a, const int64_t *name_len,
              const uint8_t *seq, const int64_t *seq_a, const int64_t *seq_len,
              const uint8_t *qual, const int64_t *qual_a, const int64_t *qual_len, const uint8_t *rna);

/* align_adapter()'s parse of the result string (nanopore_read.py:476-491) for n records {rs, re, as, ae, score,
 * match_aln, len_aln, match_ad, len_ad} (include/porechop_b200.h) without making the strings:
 *   full[i] = strtod(sprintf("%f", 100.0 * match_ad / len_ad)), part[i] likewise from match_aln / len_aln (NaN for 0/0,
 *   as float("-nan") is), read_start[i] = rs, read_end[i] = re + 1; a failed alignment (rs == -1, score == INT_MIN)
 *   gives 0.0, 0.0, -1, 0.  The (count, length) -> value table is built once per call for the pairs that occur. */
void pbioScores(const int32_t *records, int64_t n, double *full, double *part, int64_t *read_start, int64_t *read_end);

/* find_start_trim / find_end_trim (nanopore_read.py:166-208) for n reads x n_adapters records (read-major): per read
 * the largest trim amount over the adapters whose alignment passes `partial > end_threshold`, the end_size / 0 edge
 * test and `read_end - read_start >= min_trim_size`; 0 when none passes.  is_start selects the start or the end rule. */
void pbioEndTrim(const int32_t *records, int64_t n, int64_t n_adapters, int is_start, int64_t end_size,
                 int64_t extra_trim_size, double end_threshold, int64_t min_trim_size, int64_t *trim);

/* out[i, k] = full-adapter identity (as pbioScores) of record (read i, adapter cols[k]): the barcode score columns
 * (nanopore_read.py:181-183, 203-205) */
void pbioFullScores(const int32_t *records, int64_t n, int64_t n_adapters, const int64_t *cols, int64_t n_cols, double *out);

/* Parallel gzip for the .gz outputs (the reference pipes through `pigz -p threads` or `gzip`, porechop.py:640-650,
 * 683-690, 724-727): src is cut into `block`-byte pieces, each deflated as its own gzip member by one thread, members
 * concatenated in order -- a valid .gz whose decompression is src.  dst must hold pbioGzipBound(n, block) bytes.
 * Returns the compressed size, or -1 (also when the library was built without zlib: pbioGzipBound returns -1). */
int64_t pbioGzipBound(int64_t n, int64_t block);
int64_t pbioGzip(const uint8_t *src, int64_t n, int level, int64_t block, uint8_t *dst, int64_t cap);

/* bench / test workloads only: n iid uniform ACGT bytes, a pure function of (seed, n) for any number of threads
 * (SURVEY 8(d) synthetic reads: the bodies; adapter copies are implanted by porechop_b200/workloads.py) */
void pbioRandomBases(uint8_t *out, int64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
