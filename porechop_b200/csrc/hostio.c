/*
 * hostio.c -- flat-buffer FASTQ ingest / emit (include/porechop_b200_io.h).  Plain C + OpenMP, host only.
 * Behaviour follows the reference's Python loader and writers byte for byte (file:line in the header); the numpy
 * implementations in porechop_b200/fastq.py are the same functions in slow motion and are tested against these.
 */
#include "../../include/porechop_b200_io.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PBIO_BLOCK (1 << 20)

int pbioSetThreads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static inline int is_ws(uint8_t c)            /* what str.strip() removes from an ASCII line: 9-13, 28-31 and the space */
{
    return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31);
}

static int64_t count_nl(const uint8_t *p, int64_t n)
{
    int64_t c = 0;
    const uint8_t *e = p + n;
    while (p < e) {
        const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
        if (!q) break;
        ++c;
        p = q + 1;
    }
    return c;
}

int64_t pbioCountLines(const uint8_t *buf, int64_t n)
{
    if (n <= 0) return 0;
    int64_t blocks = (n + PBIO_BLOCK - 1) / PBIO_BLOCK, total = 0;
#pragma omp parallel for reduction(+ : total) schedule(static)
    for (int64_t b = 0; b < blocks; ++b) {
        int64_t lo = b * PBIO_BLOCK, hi = lo + PBIO_BLOCK < n ? lo + PBIO_BLOCK : n;
        total += count_nl(buf + lo, hi - lo);
    }
    return total + (buf[n - 1] != '\n');
}

int pbioLineEnds(const uint8_t *buf, int64_t n, int64_t *line_end, int64_t n_lines)
{
    /* one memchr pass (runs at memory bandwidth; the parallel work is in the per-record functions below) */
    int64_t k = 0;
    const uint8_t *p = buf, *e = buf + (n > 0 ? n : 0);
    while (p < e) {
        const uint8_t *q = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
        if (!q) break;
        if (k >= n_lines) return PBIO_ERR_RECORDS;
        line_end[k++] = (int64_t)(q - buf);
        p = q + 1;
    }
    if (n > 0 && buf[n - 1] != '\n') {
        if (k >= n_lines) return PBIO_ERR_RECORDS;
        line_end[k++] = n;
    }
    return k == n_lines ? PBIO_OK : PBIO_ERR_RECORDS;
}

static inline void strip(const uint8_t *buf, int64_t *a, int64_t *b)
{
    while (*b > *a && is_ws(buf[*b - 1])) --*b;
    while (*a < *b && is_ws(buf[*a])) ++*a;
}

void pbioLineSpans(const uint8_t *buf, const int64_t *line_end, int64_t n_lines, int64_t *span_a, int64_t *span_len)
{
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_lines; ++k) {
        int64_t a = k ? line_end[k - 1] + 1 : 0, b = line_end[k];
        strip(buf, &a, &b);
        span_a[k] = a;
        span_len[k] = b - a;
    }
}

int pbioFastqIndex(const uint8_t *buf, int64_t n, const int64_t *line_end, int64_t n_lines,
                   int64_t *name_a, int64_t *name_len, int64_t *seq_a, int64_t *seq_len,
                   int64_t *qual_a, int64_t *qual_len)
{
    (void)n;
    if (n_lines % 4 != 0) return PBIO_ERR_RECORDS;
    int64_t n_rec = n_lines / 4;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t r = 0; r < n_rec; ++r) {
        int64_t k = 4 * r;
        int64_t a0 = k ? line_end[k - 1] + 1 : 0, b0 = line_end[k];
        int64_t a1 = line_end[k] + 1, b1 = line_end[k + 1];
        int64_t a3 = line_end[k + 2] + 1, b3 = line_end[k + 3];
        strip(buf, &a0, &b0);
        strip(buf, &a1, &b1);
        strip(buf, &a3, &b3);
        if (b0 <= a0 || buf[a0] != '@') bad |= 1;
        name_a[r] = a0 + 1;
        name_len[r] = b0 > a0 ? b0 - a0 - 1 : 0;
        seq_a[r] = a1;
        seq_len[r] = b1 - a1;
        qual_a[r] = a3;
        qual_len[r] = b3 - a3;
    }
    return bad ? PBIO_ERR_HEADER : PBIO_OK;
}

void pbioGather(uint8_t *dst, const int64_t *dst_off, const uint8_t *src, const int64_t *src_a,
                const int64_t *src_len, int fill, int64_t n)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        int64_t d = dst_off[i], len = dst_off[i + 1] - d;
        int64_t cp = src_len ? (src_len[i] < len ? src_len[i] : len) : len;
        if (cp < 0) cp = 0;
        if (cp > 0) memcpy(dst + d, src + src_a[i], (size_t)cp);
        if (len > cp) memset(dst + d + cp, fill, (size_t)(len - cp));
    }
}

void pbioNormalise(uint8_t *seq, const int64_t *off, int64_t n, uint8_t *rna)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t *p = seq + off[i];
        int64_t len = off[i + 1] - off[i], n_u = 0, n_t = 0;
        for (int64_t k = 0; k < len; ++k) {
            uint8_t c = p[k];
            if (c >= 'a' && c <= 'z') p[k] = c = (uint8_t)(c - 32);
            n_u += c == 'U';
            n_t += c == 'T';
        }
        rna[i] = n_u > n_t;
        if (rna[i])
            for (int64_t k = 0; k < len; ++k)
                if (p[k] == 'U') p[k] = 'T';
    }
}

static inline void copy_bases(uint8_t *d, const uint8_t *s, int64_t len, int rna)
{
    if (!rna) {
        memcpy(d, s, (size_t)len);
        return;
    }
    for (int64_t k = 0; k < len; ++k) d[k] = s[k] == 'T' ? 'U' : s[k];
}

void pbioEmit(uint8_t *out, const int64_t *out_off, int64_t n_rec, int fmt,
              const uint8_t *names, const int64_t *name_a, const int64_t *name_len,
              const uint8_t *seq, const int64_t *seq_a, const int64_t *seq_len,
              const uint8_t *qual, const int64_t *qual_a, const int64_t *qual_len, const uint8_t *rna)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < n_rec; ++r) {
        uint8_t *p = out + out_off[r];
        int64_t nl = name_len[r], sl = seq_len[r];
        const uint8_t *s = seq + seq_a[r];
        int is_rna = rna && rna[r];
        *p++ = fmt == 0 ? '@' : '>';
        memcpy(p, names + name_a[r], (size_t)nl);
        p += nl;
        *p++ = '\n';
        if (fmt == 0) {
            copy_bases(p, s, sl, is_rna);
            p += sl;
            *p++ = '\n';
            *p++ = '+';
            *p++ = '\n';
            memcpy(p, qual + qual_a[r], (size_t)qual_len[r]);
            p += qual_len[r];
            *p++ = '\n';
        } else {
            for (int64_t k = 0; k < sl; k += 70) {
                int64_t w = sl - k < 70 ? sl - k : 70;
                copy_bases(p, s + k, w, is_rna);
                p += w;
                *p++ = '\n';
            }
        }
    }
}

/* float("%f" % (100.0 * c / l)) exactly as the reference chain produces it: std::to_string(double) = sprintf("%f")
 * (alignment.cpp:113-121) then Python's float() = strtod. */
static double percent_exact(int32_t c, int32_t l)
{
    if (l == 0) return NAN;
    char buf[64];
    volatile double cd = (double)c, ld = (double)l;
    snprintf(buf, sizeof buf, "%f", 100.0 * cd / ld);
    return strtod(buf, NULL);
}

#define PBIO_TABLE_MAX 2048

/* value table of percent_exact over the (count, length) pairs that occur in columns ci / li of the records;
 * t == NULL when the pairs do not fit a dense table (callers then format per record). */
typedef struct { double *t; int64_t L; } PercentTable;

static void table_build(const int32_t *records, int64_t n, int ci, int li, PercentTable *T)
{
    int32_t lmax = 0;
    int small = 1;
    T->t = NULL;
    T->L = 0;
#pragma omp parallel for schedule(static) reduction(max : lmax) reduction(& : small)
    for (int64_t i = 0; i < n; ++i) {
        const int32_t *r = records + 9 * i;
        if (r[ci] < 0 || r[li] < 0 || r[ci] > r[li]) small = 0;
        if (r[li] > lmax) lmax = r[li];
    }
    if (!small || lmax >= PBIO_TABLE_MAX) return;
    const int64_t L = (int64_t)lmax + 1;
    double *t = (double *)malloc((size_t)(L * L) * sizeof(double));
    uint8_t *present = (uint8_t *)calloc((size_t)(L * L), 1);
    if (!t || !present) { free(t); free(present); return; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const int32_t *r = records + 9 * i;
        present[(int64_t)r[ci] * L + r[li]] = 1;            /* benign race: every writer stores 1 */
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < L * L; ++k)
        if (present[k]) t[k] = percent_exact((int32_t)(k / L), (int32_t)(k % L));
    free(present);
    T->t = t;
    T->L = L;
}

static inline double table_get(const PercentTable *T, int32_t c, int32_t l)
{
    return T->t ? T->t[(int64_t)c * T->L + l] : percent_exact(c, l);
}

static inline int rec_failed(const int32_t *r) { return r[0] == -1 && r[4] == INT_MIN; }

void pbioScores(const int32_t *records, int64_t n, double *full, double *part, int64_t *read_start, int64_t *read_end)
{
    PercentTable P, F;
    table_build(records, n, 5, 6, &P);
    table_build(records, n, 7, 8, &F);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const int32_t *r = records + 9 * i;
        if (rec_failed(r)) {
            full[i] = 0.0; part[i] = 0.0; read_start[i] = -1; read_end[i] = 0;
            continue;
        }
        part[i] = table_get(&P, r[5], r[6]);
        full[i] = table_get(&F, r[7], r[8]);
        read_start[i] = r[0];
        read_end[i] = (int64_t)r[1] + 1;
    }
    free(P.t);
    free(F.t);
}

void pbioEndTrim(const int32_t *records, int64_t n, int64_t n_adapters, int is_start, int64_t end_size,
                 int64_t extra_trim_size, double end_threshold, int64_t min_trim_size, int64_t *trim)
{
    PercentTable P;
    table_build(records, n * n_adapters, 5, 6, &P);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int64_t best = 0;
        for (int64_t a = 0; a < n_adapters; ++a) {
            const int32_t *r = records + 9 * (i * n_adapters + a);
            const int failed = rec_failed(r);
            const double part = failed ? 0.0 : table_get(&P, r[5], r[6]);
            const int64_t rs = failed ? -1 : r[0], re = failed ? 0 : (int64_t)r[1] + 1;
            if (!(part > end_threshold) || re - rs < min_trim_size) continue;     /* NaN > x is false, as in Python */
            int64_t amount;
            if (is_start) {
                if (re == end_size) continue;
                amount = re + extra_trim_size;
            } else {
                if (rs == 0) continue;
                amount = (end_size - rs) + extra_trim_size;
            }
            if (amount > best) best = amount;
        }
        trim[i] = best;
    }
    free(P.t);
}

void pbioFullScores(const int32_t *records, int64_t n, int64_t n_adapters, const int64_t *cols, int64_t n_cols, double *out)
{
    PercentTable F;
    table_build(records, n * n_adapters, 7, 8, &F);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int64_t k = 0; k < n_cols; ++k) {
            const int32_t *r = records + 9 * (i * n_adapters + cols[k]);
            out[i * n_cols + k] = rec_failed(r) ? 0.0 : table_get(&F, r[7], r[8]);
        }
    free(F.t);
}

/* ---- parallel gzip: independent members of `block` input bytes each, concatenated (a valid .gz of the whole) ---- */
#ifdef PBIO_HAVE_ZLIB
#include <zlib.h>

int64_t pbioGzipBound(int64_t n, int64_t block)
{
    if (n <= 0) return 0;
    if (block < 1024) block = 1024;
    int64_t nb = (n + block - 1) / block;
    return nb * ((int64_t)compressBound((uLong)block) + 64);
}

int64_t pbioGzip(const uint8_t *src, int64_t n, int level, int64_t block, uint8_t *dst, int64_t cap)
{
    if (n <= 0) return 0;
    if (block < 1024) block = 1024;
    const int64_t nb = (n + block - 1) / block;
    const int64_t slot = (int64_t)compressBound((uLong)block) + 64;
    if (cap < nb * slot) return -1;
    int64_t *sizes = (int64_t *)malloc((size_t)nb * sizeof(int64_t));
    if (!sizes) return -1;
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
    for (int64_t b = 0; b < nb; ++b) {
        z_stream z;
        memset(&z, 0, sizeof z);
        sizes[b] = 0;
        if (deflateInit2(&z, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad |= 1; continue; }
        int64_t lo = b * block, len = (lo + block <= n) ? block : n - lo;
        z.next_in = (Bytef *)(src + lo);
        z.avail_in = (uInt)len;
        z.next_out = dst + b * slot;
        z.avail_out = (uInt)slot;
        if (deflate(&z, Z_FINISH) != Z_STREAM_END) bad |= 1;
        sizes[b] = (int64_t)z.total_out;
        deflateEnd(&z);
    }
    int64_t pos = 0;
    if (!bad)
        for (int64_t b = 0; b < nb; ++b) {              /* pack the members */
            if (pos != b * slot) memmove(dst + pos, dst + b * slot, (size_t)sizes[b]);
            pos += sizes[b];
        }
    free(sizes);
    return bad ? -1 : pos;
}
#else
int64_t pbioGzipBound(int64_t n, int64_t block) { (void)n; (void)block; return -1; }
int64_t pbioGzip(const uint8_t *src, int64_t n, int level, int64_t block, uint8_t *dst, int64_t cap)
{
    (void)src; (void)n; (void)level; (void)block; (void)dst; (void)cap;
    return -1;                                          /* built without zlib: callers use Python's gzip */
}
#endif


/* ---- synthetic workloads (bench.py / workloads.py): iid uniform ACGT, reproducible for any thread count -------------- */
static inline uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void pbioRandomBases(uint8_t *out, int64_t n, uint64_t seed)
{
    static const uint8_t acgt[4] = {'A', 'C', 'G', 'T'};
    if (n <= 0) return;
    const int64_t blocks = (n + PBIO_BLOCK - 1) / PBIO_BLOCK;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t b = 0; b < blocks; ++b) {
        /* one generator per 1 MiB block, keyed by (seed, block): the bytes do not depend on the thread count */
        uint64_t st = seed * 0xD1342543DE82EF95ull + (uint64_t)b * 0x2545F4914F6CDD1Dull + 1;
        int64_t lo = b * PBIO_BLOCK, hi = lo + PBIO_BLOCK < n ? lo + PBIO_BLOCK : n;
        int64_t i = lo;
        for (; i + 32 <= hi; i += 32) {
            uint64_t r = splitmix64(&st);
            for (int k = 0; k < 32; ++k) out[i + k] = acgt[(r >> (2 * k)) & 3];
        }
        if (i < hi) {
            uint64_t r = splitmix64(&st);
            for (int k = 0; i < hi; ++i, ++k) out[i] = acgt[(r >> (2 * k)) & 3];
        }
    }
}
