// hostpack.cpp — host-side half of the upload path of db200_sketch_batch: ASCII bases -> the packed genome store's
// 2-bit code words + validity bits (the very layout pack_kernel writes on the device, sketch.cuh), so that the PCIe link
// carries 0.375 bytes per base instead of 1 ("host packs ... or uploads ASCII and packs on device —
// measure both").  Compiled by g++ (not nvcc) with per-function target attributes; the widest ISA the CPU offers is picked
// once at run time.  This is a data-format conversion (alphabet.h:128's DNA4 table as bit tricks: A0 C1 G2 T3, anything but
// ACGTacgt invalid); k-mer hashing and register updates stay on the GPU.
//
//   codes[g]  (uint32, 16 bases): base j at bits [2j, 2j+1], code = ((c >> 1) ^ (c >> 2)) & 3
//   valid[g]  (uint16, 16 bases): bit j set iff base j is one of ACGTacgt
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

inline void pack16_scalar(const uint8_t *s, size_t n, uint32_t &code, uint16_t &val) {
    uint32_t c = 0, v = 0;
    for (size_t j = 0; j < n; ++j) {
        const uint32_t x = s[j], up = x & 0xDFu;
        c |= (((x >> 1) ^ (x >> 2)) & 3u) << (2 * j);
        v |= (uint32_t)(up == 0x41u || up == 0x43u || up == 0x47u || up == 0x54u) << j;
    }
    code = c; val = (uint16_t)v;
}

void pack_scalar(const uint8_t *ascii, size_t ngroups, uint32_t *codes, uint16_t *valid) {
    for (size_t g = 0; g < ngroups; ++g) pack16_scalar(ascii + 16 * g, 16, codes[g], valid[g]);
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_avx2(const uint8_t *ascii, size_t ngroups, uint32_t *codes, uint16_t *valid) {
    const __m256i df = _mm256_set1_epi8((char)0xDF), cA = _mm256_set1_epi8(0x41), cC = _mm256_set1_epi8(0x43), cG = _mm256_set1_epi8(0x47),
                  cT = _mm256_set1_epi8(0x54), three = _mm256_set1_epi8(3);
    const __m256i w14 = _mm256_set1_epi16(0x0401);          // bytes (1, 4): c0 + 4 c1
    const __m256i w116 = _mm256_set1_epi32(0x00100001);     // words (1, 16): lo + 16 hi
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    size_t g = 0;
    for (; g + 2 <= ngroups; g += 2) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(ascii + 16 * g));
        const __m256i up = _mm256_and_si256(x, df);
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(up, cA), _mm256_cmpeq_epi8(up, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(up, cG), _mm256_cmpeq_epi8(up, cT)));
        const uint32_t vm = (uint32_t)_mm256_movemask_epi8(ok);
        const __m256i c = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(x, 1), _mm256_srli_epi16(x, 2)), three);
        const __m256i b = _mm256_madd_epi16(_mm256_maddubs_epi16(c, w14), w116);       // one packed byte per 32-bit lane
        const __m256i q = _mm256_shuffle_epi8(b, gather);
        codes[g] = (uint32_t)_mm256_extract_epi32(q, 0);
        codes[g + 1] = (uint32_t)_mm256_extract_epi32(q, 4);
        valid[g] = (uint16_t)vm;
        valid[g + 1] = (uint16_t)(vm >> 16);
    }
    for (; g < ngroups; ++g) pack16_scalar(ascii + 16 * g, 16, codes[g], valid[g]);
}

__attribute__((target("avx512f,avx512bw"))) void pack_avx512(const uint8_t *ascii, size_t ngroups, uint32_t *codes, uint16_t *valid) {
    const __m512i df = _mm512_set1_epi8((char)0xDF), cA = _mm512_set1_epi8(0x41), cC = _mm512_set1_epi8(0x43), cG = _mm512_set1_epi8(0x47),
                  cT = _mm512_set1_epi8(0x54), three = _mm512_set1_epi8(3);
    const __m512i w14 = _mm512_set1_epi16(0x0401), w116 = _mm512_set1_epi32(0x00100001);
    size_t g = 0;
    for (; g + 4 <= ngroups; g += 4) {
        const __m512i x = _mm512_loadu_si512(ascii + 16 * g);
        const __m512i up = _mm512_and_si512(x, df);
        const __mmask64 vm = _mm512_cmpeq_epi8_mask(up, cA) | _mm512_cmpeq_epi8_mask(up, cC) | _mm512_cmpeq_epi8_mask(up, cG) | _mm512_cmpeq_epi8_mask(up, cT);
        const __m512i c = _mm512_and_si512(_mm512_xor_si512(_mm512_srli_epi16(x, 1), _mm512_srli_epi16(x, 2)), three);
        const __m512i b = _mm512_madd_epi16(_mm512_maddubs_epi16(c, w14), w116);       // one packed byte per 32-bit lane
        const __m128i q = _mm512_cvtepi32_epi8(b);                                      // 16 bytes = 4 code words
        _mm_storeu_si128(reinterpret_cast<__m128i *>(codes + g), q);
        const uint64_t v = (uint64_t)vm;
        std::memcpy(valid + g, &v, 8);
    }
    for (; g < ngroups; ++g) pack16_scalar(ascii + 16 * g, 16, codes[g], valid[g]);
}
#endif

using pack_fn = void (*)(const uint8_t *, size_t, uint32_t *, uint16_t *);

pack_fn pick() {
    // DB200_HOSTPACK_ISA=avx2|scalar caps the choice (tests exercise every variant on one machine)
    const char *cap = std::getenv("DB200_HOSTPACK_ISA");
    const bool no512 = cap && (!std::strcmp(cap, "avx2") || !std::strcmp(cap, "scalar")), no256 = cap && !std::strcmp(cap, "scalar");
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (!no512 && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f")) return pack_avx512;
    if (!no256 && __builtin_cpu_supports("avx2")) return pack_avx2;
#endif
    (void)no512; (void)no256;
    return pack_scalar;
}

} // namespace

// ascii[0, nbases) -> codes / valid for ceil(nbases / 16) groups; a trailing partial group is zero padded.
extern "C" __attribute__((visibility("default"))) void db200_hostpack(const uint8_t *ascii, size_t nbases, uint32_t *codes, uint16_t *valid) {
    static const pack_fn fn = pick();
    const size_t full = nbases / 16;
    fn(ascii, full, codes, valid);
    if (nbases % 16) pack16_scalar(ascii + 16 * full, nbases % 16, codes[full], valid[full]);
}

extern "C" __attribute__((visibility("default"))) const char *db200_hostpack_isa() {
    static const pack_fn fn = pick();
#if defined(__x86_64__)
    if (fn == pack_avx512) return "avx512bw";
    if (fn == pack_avx2) return "avx2";
#endif
    return "scalar";
}
