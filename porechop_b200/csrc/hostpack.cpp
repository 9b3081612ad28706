// porechop_b200/csrc/hostpack.cpp -- host half of the packed upload path of cpp_functions.so (option "h2d_pack").
//
// The sequences cross the reference's ABI as ASCII, one byte per base (porechop/src/adapter_align.cpp:15-16 turns them
// into Dna5 codes first thing: seqan/basic/alphabet_residue_tabs.h:113-140).  The end-to-end path of the engine is bound
// by those bytes crossing PCIe, not by the kernels, so this file does the Dna5 conversion on the
// host cores instead and packs two 4-bit codes per byte: half the bytes over the link, and the device unpacks
// (kernels.cuh unpack_kernel) into exactly the code bytes encode_kernel would have produced.
//
// Plain C++ (g++), linked into cpp_functions.so by build.py; AVX-512BW / AVX2 body selected at run time, scalar table otherwise.
// No alignment arithmetic here -- only the alphabet conversion.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

// same mapping as pb::encode_byte (dp_core.cuh): A/a=0 C/c=1 G/g=2 T/t/U/u=3, every other byte = 4
struct CodeTable {
    uint8_t t[256];
    CodeTable() {
        memset(t, 4, sizeof t);
        t[(int)'A'] = t[(int)'a'] = 0; t[(int)'C'] = t[(int)'c'] = 1; t[(int)'G'] = t[(int)'g'] = 2;
        t[(int)'T'] = t[(int)'t'] = t[(int)'U'] = t[(int)'u'] = 3;
    }
};
const CodeTable g_tab;

// bytes [i0, i1) of `in` -> nibbles; i0 is even, so every output byte belongs to exactly one caller
void pack_scalar(const uint8_t *in, int64_t i0, int64_t i1, int64_t n, uint8_t *out) {
    int64_t i = i0;
    for (; i + 1 < i1; i += 2) out[i >> 1] = (uint8_t)(g_tab.t[in[i]] | (g_tab.t[in[i + 1]] << 4));
    if (i < i1) out[i >> 1] = (uint8_t)(g_tab.t[in[i]] | ((i + 1 < n ? g_tab.t[in[i + 1]] : 0) << 4));
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_avx2(const uint8_t *in, int64_t i0, int64_t i1, int64_t n, uint8_t *out) {
    // per 128-bit lane tables indexed by the low nibble of the case-folded byte: the letter that nibble must be,
    // and its code ('A' 0x41, 'C' 0x43, 'T' 0x54, 'U' 0x55, 'G' 0x47)
    const __m256i tchar = _mm256_setr_epi8((char)0xFF, 0x41, (char)0xFF, 0x43, 0x54, 0x55, (char)0xFF, 0x47, (char)0xFF, (char)0xFF,
                                           (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF,
                                           (char)0xFF, 0x41, (char)0xFF, 0x43, 0x54, 0x55, (char)0xFF, 0x47, (char)0xFF, (char)0xFF,
                                           (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF);
    const __m256i tcode = _mm256_setr_epi8(4, 0, 4, 1, 3, 3, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 0, 4, 1, 3, 3, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4);
    const __m256i fold = _mm256_set1_epi8((char)0xDF), lo4 = _mm256_set1_epi8(0x0F), four = _mm256_set1_epi8(4);
    const __m256i w = _mm256_set1_epi16(0x1001);      // bytes (1, 16): even code + 16 * odd code
    const __m256i zero = _mm256_setzero_si256();
    int64_t i = i0;
    for (; i + 32 <= i1; i += 32) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(in + i));
        const __m256i f = _mm256_and_si256(x, fold);
        const __m256i nib = _mm256_and_si256(f, lo4);
        const __m256i ok = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(tchar, nib), f);
        const __m256i code = _mm256_blendv_epi8(four, _mm256_shuffle_epi8(tcode, nib), ok);
        const __m256i pr = _mm256_maddubs_epi16(code, w);                 // 16 x (c0 + 16*c1)
        const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi16(pr, zero), 0x08);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(out + (i >> 1)), _mm256_castsi256_si128(pk));
    }
    if (i < i1) pack_scalar(in, i, i1, n, out);
}
bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
// the same conversion 64 bases at a time (AVX-512BW: byte shuffles per 128-bit lane, compare into a mask, vpmovwb)
__attribute__((target("avx512f,avx512bw"))) void pack_avx512(const uint8_t *in, int64_t i0, int64_t i1, int64_t n, uint8_t *out) {
    const __m512i tchar = _mm512_broadcast_i32x4(_mm_setr_epi8((char)0xFF, 0x41, (char)0xFF, 0x43, 0x54, 0x55, (char)0xFF, 0x47, (char)0xFF,
                                                               (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF));
    const __m512i tcode = _mm512_broadcast_i32x4(_mm_setr_epi8(4, 0, 4, 1, 3, 3, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4));
    const __m512i fold = _mm512_set1_epi8((char)0xDF), lo4 = _mm512_set1_epi8(0x0F), four = _mm512_set1_epi8(4);
    const __m512i w = _mm512_set1_epi16(0x1001);
    int64_t i = i0;
    for (; i + 64 <= i1; i += 64) {
        const __m512i x = _mm512_loadu_si512(reinterpret_cast<const void *>(in + i));
        const __m512i f = _mm512_and_si512(x, fold);
        const __m512i nib = _mm512_and_si512(f, lo4);
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(tchar, nib), f);
        const __m512i code = _mm512_mask_blend_epi8(ok, four, _mm512_shuffle_epi8(tcode, nib));
        const __m512i pr = _mm512_maddubs_epi16(code, w);                  // 32 x (c0 + 16*c1)
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(out + (i >> 1)), _mm512_cvtepi16_epi8(pr));
    }
    if (i < i1) pack_avx2(in, i, i1, n, out);
}
bool have_avx512() {
    static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && !getenv("PB200_PACK_NO_AVX512");
    return v;
}
#endif

void pack_range(const uint8_t *in, int64_t i0, int64_t i1, int64_t n, uint8_t *out) {
#if defined(__x86_64__)
    if (have_avx512()) { pack_avx512(in, i0, i1, n, out); return; }
    if (have_avx2()) { pack_avx2(in, i0, i1, n, out); return; }
#endif
    pack_scalar(in, i0, i1, n, out);
}

}  // namespace

// n ASCII bytes -> (n+1)/2 bytes, base 2k in the low nibble and base 2k+1 in the high nibble of byte k (a missing last
// base packs as 0).  threads <= 0: the OpenMP default.  Safe to call concurrently on different buffers.
extern "C" void pb_pack_nibbles(const uint8_t *in, int64_t n, uint8_t *out, int threads) {
    if (n <= 0) return;
    const int64_t BLK = 1 << 16;                      // bytes per work item (even)
    const int64_t nblk = (n + BLK - 1) / BLK;
#if defined(_OPENMP)
    int nt = threads > 0 ? threads : omp_get_max_threads();
    if ((int64_t)nt > nblk) nt = (int)nblk;
    if (nt > 1) {
#pragma omp parallel for schedule(static) num_threads(nt)
        for (int64_t b = 0; b < nblk; ++b) pack_range(in, b * BLK, (b + 1) * BLK < n ? (b + 1) * BLK : n, n, out);
        return;
    }
#else
    (void)threads; (void)nblk;
#endif
    pack_range(in, 0, n, n, out);
}
