// mq_pack.cpp -- host side of the packed sequence format (include/mapquik_b200.h, "packed input").
//
// The reference copies every record once before it enters the hot path (`to_ascii_uppercase()`, closures.rs:63,106).
// A caller of this library makes that one pass with mq_pack / mq_pack_at instead: 2-bit codes (code = (byte >> 1) & 3,
// i.e. A0 C1 T2 G3), 16 bases per 32-bit word, plus what is needed to stay EXACTLY equivalent to the ASCII bytes:
//   flags  one bit per 64-base block that holds a byte other than A/C/G/T,
//   exc    the intervals of such bytes (start, len, byte), sorted by start.
// Host -> device traffic drops from 1 to 0.25 bytes per base and the scan kernel's per-byte digest disappears.
// No CUDA in this file.
#include "../../include/mapquik_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
inline uint8_t fold(uint8_t c, bool fold_case) { return (fold_case && c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

struct ExcSink {
    std::vector<mq_exc> v;
    void add(uint64_t pos, uint8_t byte) {
        if (!v.empty() && v.back().byte == byte && v.back().start + v.back().len == pos && v.back().len < 0xFFFFFFFFu) v.back().len++;
        else v.push_back(mq_exc{pos, 1u, byte});
    }
};

inline void or_word(uint32_t *p, uint32_t bits) { __atomic_fetch_or(p, bits, __ATOMIC_RELAXED); }

// bases [at, at+n) one at a time; words and flags are OR-ed in atomically (edges shared with a neighbouring range)
void pack_scalar(const uint8_t *src, uint64_t n, uint64_t at, uint32_t *words, uint32_t *flags, ExcSink &ex, bool fc) {
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t c = fold(src[i], fc);
        const uint64_t b = at + i;
        const uint32_t code = (c >> 1) & 3u;
        if (code) or_word(words + (b >> 4), code << (2 * (b & 15)));
        if (!is_acgt(c)) { ex.add(b, c); or_word(flags + (b >> 11), 1u << ((b >> 6) & 31)); }
    }
}

// Groups of 32 bases at a 32-aligned destination.  Packs whole groups until one holds a byte other than A/C/G/T (after
// folding); returns the number of clean groups written (the caller handles the offending group and calls again).
#if defined(__x86_64__)
__attribute__((target("avx2"))) size_t pack_groups_avx2(const uint8_t *src, size_t groups, uint32_t *wout, bool fc) {
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    const __m256i m3 = _mm256_set1_epi8(3), k0401 = _mm256_set1_epi16(0x0401), k1001 = _mm256_set1_epi32(0x00100001);
    const __m256i la = _mm256_set1_epi8('a' - 1), lz = _mm256_set1_epi8('z' + 1), k32 = _mm256_set1_epi8(32);
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    size_t g = 0;
    for (; g < groups; g++) {
        __m256i v = _mm256_loadu_si256((const __m256i *)(src + 32 * g));
        if (fc) {
            const __m256i lower = _mm256_and_si256(_mm256_cmpgt_epi8(v, la), _mm256_cmpgt_epi8(lz, v));
            v = _mm256_sub_epi8(v, _mm256_and_si256(lower, k32));
        }
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, cA), _mm256_cmpeq_epi8(v, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(v, cG), _mm256_cmpeq_epi8(v, cT)));
        if ((uint32_t)_mm256_movemask_epi8(ok) != 0xFFFFFFFFu) break;
        const __m256i codes = _mm256_and_si256(_mm256_srli_epi16(v, 1), m3);
        // c0 + 4 c1 per 16-bit lane, then (that) + 16 (next) per 32-bit lane: one byte of four codes per 32-bit lane
        const __m256i p32 = _mm256_madd_epi16(_mm256_maddubs_epi16(codes, k0401), k1001);
        const __m256i sh = _mm256_shuffle_epi8(p32, gather);        // low byte of the eight 32-bit lanes -> dword 0 of each half
        wout[2 * g] = (uint32_t)_mm256_cvtsi256_si32(sh);
        wout[2 * g + 1] = (uint32_t)_mm256_extract_epi32(sh, 4);
    }
    return g;
}
bool have_avx2() { static const bool h = __builtin_cpu_supports("avx2"); return h; }

// Groups of 64 bases (AVX-512 VBMI): one 128-entry byte table (two registers) maps a letter to its code or to 0x80, so the
// validity test, the case fold and the code extraction are ONE permute; bytes >= 0x80 are caught by OR-ing the input in.
// Four codes per byte with two multiply-adds, sixteen bytes out with one down-convert.  `stream`: the destination is
// 16-byte aligned and will not be read by this core again (pinned staging for a DMA) -> non-temporal store, no
// read-for-ownership of the destination lines.
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi"))) size_t pack_groups_avx512(const uint8_t *src, size_t groups, uint32_t *wout, bool fc, bool stream) {
    alignas(64) uint8_t lut[128];
    memset(lut, 0x80, sizeof lut);
    lut['A'] = 0; lut['C'] = 1; lut['T'] = 2; lut['G'] = 3;
    if (fc) { lut['a'] = 0; lut['c'] = 1; lut['t'] = 2; lut['g'] = 3; }
    const __m512i t0 = _mm512_load_si512(lut), t1 = _mm512_load_si512(lut + 64);
    const __m512i k0401 = _mm512_set1_epi16(0x0401), k1001 = _mm512_set1_epi32(0x00100001);
    size_t g = 0;
    for (; g < groups; g++) {
        const __m512i v = _mm512_loadu_si512(src + 64 * g);
        const __m512i c = _mm512_permutex2var_epi8(t0, v, t1);            // index = low 7 bits of the letter
        if (_mm512_movepi8_mask(_mm512_or_si512(c, v))) break;             // a byte >= 0x80, or one the table rejects
        const __m128i out = _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(c, k0401), k1001));
        if (stream) _mm_stream_si128((__m128i *)(wout + 4 * g), out);
        else _mm_storeu_si128((__m128i *)(wout + 4 * g), out);
    }
    if (stream) _mm_sfence();
    return g;
}
bool have_avx512() {
    static const bool h = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                          __builtin_cpu_supports("avx512vbmi") && !getenv("MQ_PACK_NO_AVX512");
    return h;
}
#else
bool have_avx2() { return false; }
bool have_avx512() { return false; }
#endif

inline uint64_t pack32_plain(const uint8_t *src, bool fc, bool *bad) {
    uint64_t w = 0; bool b = false;
    for (int i = 0; i < 32; i++) { const uint8_t c = fold(src[i], fc); w |= (uint64_t)((c >> 1) & 3u) << (2 * i); b |= !is_acgt(c); }
    *bad = b;
    return w;
}

// thread-safe against ranges that share its first / last word or flag word (those are OR-ed in; the destination must
// have been zeroed); full words in between are plain stores
void pack_range(const uint8_t *src, uint64_t n, uint64_t at, uint32_t *words, uint32_t *flags, ExcSink &ex, bool fc) {
    if (n == 0) return;
    uint64_t i = 0;
    const uint64_t head = std::min<uint64_t>(n, (32 - (at & 31)) & 31);      // up to a 32-base boundary (two whole words)
    pack_scalar(src, head, at, words, flags, ex, fc);
    i = head;
    const bool avx = have_avx2(), avx512 = have_avx512();
    while (i + 32 <= n) {
#if defined(__x86_64__)
        // code words are never shared with a neighbouring range (ranges start on 32-base boundaries here): plain stores
        if (avx512 && i + 64 <= n) {
            uint32_t *w = words + ((at + i) >> 4);
            const size_t done = pack_groups_avx512(src + i, (size_t)((n - i) / 64), w, fc, ((uintptr_t)w & 15) == 0);
            i += 64 * done;
            if (i + 32 > n) break;
            if (done) continue;                   // stopped at an exception: 32 bases the narrow way, then wide again
        }
        if (avx) {
            const size_t done = pack_groups_avx2(src + i, avx512 ? 1 : (size_t)((n - i) / 32), words + ((at + i) >> 4), fc);
            i += 32 * done;
            if (i + 32 > n) break;
            if (avx512 && done) continue;
        }
#endif
        // a group with a byte other than A/C/G/T (or no AVX2): one group the plain way
        bool bad;
        const uint64_t w = pack32_plain(src + i, fc, &bad);
        const uint64_t bb = at + i;
        words[bb >> 4] = (uint32_t)w; words[(bb >> 4) + 1] = (uint32_t)(w >> 32);
        if (bad) {
            for (int j = 0; j < 32; j++) { const uint8_t c = fold(src[i + j], fc); if (!is_acgt(c)) ex.add(bb + j, c); }
            or_word(flags + (bb >> 11), 1u << ((bb >> 6) & 31));
        }
        i += 32;
    }
    pack_scalar(src + i, n - i, at + i, words, flags, ex, fc);
}

}  // namespace

// Internal (not part of the C ABI): pack the 2048-base units [u0, u1) of a destination whose base 0 is ascii[0];
// n_bases = total bases of the destination.  Units are whole flag words, so concurrent callers on disjoint unit ranges
// share no word: nothing needs to be zeroed beforehand.  Used by mq_pack's threads and by the on-the-fly packer of
// mq_map_batch (mq_lib.cu).
void mq_pack_units(const uint8_t *ascii, uint64_t n_bases, uint64_t u0, uint64_t u1, uint32_t *words, uint32_t *flags,
                   std::vector<mq_exc> &exc, bool fold_case) {
    const uint64_t b0 = u0 * 2048, b1 = std::min(n_bases, u1 * 2048);
    memset(flags + u0, 0, (u1 - u0) * 4);
    if (b1 <= b0) return;
    // whole 32-base groups are written with plain stores; only the words of the trailing partial group are OR-ed into
    const uint64_t tail = b0 + ((b1 - b0) & ~31ull);
    memset(words + (tail >> 4), 0, ((b1 + 15) / 16 - (tail >> 4)) * 4);
    ExcSink ex;
    ex.v.swap(exc);
    pack_range(ascii + b0, b1 - b0, b0, words, flags, ex, fold_case);
    exc.swap(ex.v);
}

extern "C" {

uint64_t mq_packed_words(uint64_t n_bases) { return (n_bases + 15) / 16 + 64; }       // + 256 bytes of slack (word loads past the end)
uint64_t mq_packed_flag_words(uint64_t n_bases) { return (n_bases + 2047) / 2048 + 2; }

int mq_pack_at(const uint8_t *ascii, uint64_t n_bases, uint64_t at_base, uint32_t *words, uint32_t *flags, mq_exc *exc,
               uint64_t exc_cap, uint64_t *n_exc, int fold_case) {
    if ((!ascii && n_bases) || !words || !flags || !n_exc) return MQ_ERR_ARG;
    ExcSink ex;
    try {
        pack_range(ascii, n_bases, at_base, words, flags, ex, fold_case != 0);
    } catch (...) { return MQ_ERR_NOMEM; }
    *n_exc = ex.v.size();
    if (ex.v.size() > exc_cap) return MQ_ERR_RANGE;
    if (!ex.v.empty()) memcpy(exc, ex.v.data(), ex.v.size() * sizeof(mq_exc));
    return MQ_OK;
}

int mq_pack(const uint8_t *ascii, uint64_t n_bases, uint32_t *words, uint32_t *flags, mq_exc *exc, uint64_t exc_cap,
            uint64_t *n_exc, int n_threads, int fold_case) {
    if ((!ascii && n_bases) || !words || !flags || !n_exc) return MQ_ERR_ARG;
    *n_exc = 0;
    try {
        if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
        // pieces of whole 2048-base flag words: no word of the destination is shared between two threads
        const uint64_t units = (n_bases + 2047) / 2048;
        n_threads = (int)std::min<uint64_t>((uint64_t)n_threads, std::max<uint64_t>(1, units / 64));
        std::vector<std::vector<mq_exc>> sinks(n_threads);
        auto work = [&](int t) {
            mq_pack_units(ascii, n_bases, units * t / n_threads, units * (t + 1) / n_threads, words, flags, sinks[t], fold_case != 0);
        };
        if (n_threads == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
            for (auto &x : th) x.join();
        }
        // slack behind the last base reads as zero
        const uint64_t wend = (n_bases + 15) / 16;
        memset(words + wend, 0, (mq_packed_words(n_bases) - wend) * 4);
        memset(flags + units, 0, (mq_packed_flag_words(n_bases) - units) * 4);
        uint64_t tot = 0;
        for (auto &s : sinks) tot += s.size();
        *n_exc = tot;
        if (tot > exc_cap) return MQ_ERR_RANGE;
        uint64_t w = 0;
        for (auto &s : sinks) for (auto &e : s) {
            // an interval that continues the previous thread's last one is merged (cosmetic: either form is exact)
            if (w && exc[w - 1].byte == e.byte && exc[w - 1].start + exc[w - 1].len == e.start && (uint64_t)exc[w - 1].len + e.len <= 0xFFFFFFFFull) exc[w - 1].len += e.len;
            else exc[w++] = e;
        }
        *n_exc = w;
    } catch (...) { return MQ_ERR_NOMEM; }
    return MQ_OK;
}

// the inverse (tests, debugging): ASCII bytes of bases [first, first+n)
int mq_unpack(const uint32_t *words, const mq_exc *exc, uint64_t n_exc, uint64_t first, uint64_t n, uint8_t *ascii) {
    if (!words || (!ascii && n) || (n_exc && !exc)) return MQ_ERR_ARG;
    static const char L[4] = {'A', 'C', 'T', 'G'};
    for (uint64_t i = 0; i < n; i++) { const uint64_t b = first + i; ascii[i] = (uint8_t)L[(words[b >> 4] >> (2 * (b & 15))) & 3u]; }
    for (uint64_t e = 0; e < n_exc; e++) {
        const uint64_t s = std::max(exc[e].start, first), t = std::min(exc[e].start + exc[e].len, first + n);
        for (uint64_t b = s; b < t; b++) ascii[b - first] = (uint8_t)exc[e].byte;
    }
    return MQ_OK;
}

}  // extern "C"
