// mq_sim.cpp -- seeded genome / HiFi-like read simulator (host C++, no CUDA).
//
// Stands in for the tools the reference's experiments use but which are absent here
// (pbsim in example/simulate_pbsim.sh:7-14, experiments/simulate_chm13.sh,
// experiments/simulate_maize.sh).  Read names keep the truth-encoding format of the shipped
// fixture, `S1_<n>!<contig>!<start>!<end>!<strand>` (example/nearperfect-ecoli.100.fa:1), so the
// mapeval-style accuracy check (example/run_ecoli.sh:27) works on every synthetic config.
//
// PRNG: xoshiro256** seeded through splitmix64.  Every read / contig owns a generator derived
// from (seed, index), so output is independent of the number of threads.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {
struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    Rng(uint64_t seed, uint64_t stream) {
        uint64_t x = seed * 0xD1342543DE82EF95ULL + stream * 0x2545F4914F6CDD1DULL + 0x1234567ULL;
        for (int i = 0; i < 4; i++) s[i] = splitmix(x);
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    inline uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    inline double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint64_t below(uint64_t n) { return (uint64_t)(((__uint128_t)next() * n) >> 64); }
    inline double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};
const char BASES[4] = {'A', 'C', 'G', 'T'};
inline uint8_t comp(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}
inline uint8_t other_base(uint8_t c, Rng &r) {
    for (;;) { uint8_t b = (uint8_t)BASES[r.below(4)]; if (b != c) return b; }
}
// copy src[0..n) into dst[0..cap) with per-base divergence `div` (80 % substitutions, 10 % 1-base
// deletions, 10 % 1-base insertions); returns bytes written (<= cap)
uint64_t mutate_copy(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, double div, Rng &r) {
    uint64_t w = 0;
    for (uint64_t i = 0; i < n && w < cap; i++) {
        if (div > 0 && r.uni() < div) {
            double t = r.uni();
            if (t < 0.8) dst[w++] = other_base(src[i], r);
            else if (t < 0.9) { /* deletion */ }
            else { dst[w++] = (uint8_t)BASES[r.below(4)]; if (w < cap) dst[w++] = src[i]; }
        } else dst[w++] = src[i];
    }
    return w;
}
}  // namespace

extern "C" {

// Uniform random ACGT.  Blocks of 1 Mbase own a generator each, so the result does not depend on the thread count.
void mqsim_random_bases(uint64_t seed, uint64_t stream, uint8_t *out, uint64_t n) {
    const uint64_t BLK = 1u << 20;
    const int64_t nb = (int64_t)((n + BLK - 1) / BLK);
    #pragma omp parallel for schedule(dynamic, 4)
    for (int64_t b = 0; b < nb; b++) {
        Rng r(seed, (stream << 24) + (uint64_t)b + 1);
        uint64_t i = (uint64_t)b * BLK; const uint64_t end = std::min(n, i + BLK);
        while (i < end) {
            uint64_t x = r.next();
            for (int j = 0; j < 32 && i < end; j++, x >>= 2) out[i++] = (uint8_t)BASES[x & 3];
        }
    }
}

// Overlay tandem-satellite arrays on about `frac` of the contig: arrays of 20-400 kb, monomer
// 171-180 bp (alpha-satellite-like, half of the arrays) or 5-2000 bp, every copy diverged by
// 0.5-3 % from the array's monomer.
void mqsim_add_satellites(uint64_t seed, uint64_t stream, uint8_t *g, uint64_t n, double frac) {
    if (frac <= 0 || n < 100000) return;
    Rng r(seed ^ 0x5A7E111EULL, stream);
    uint64_t target = (uint64_t)(frac * (double)n), done = 0;
    std::vector<uint8_t> mono;
    while (done < target) {
        uint64_t alen = 20000 + r.below(380000);
        if (alen > n / 4) alen = n / 4;
        uint64_t at = r.below(n - alen);
        uint64_t p = (r.uni() < 0.5) ? 171 + r.below(10) : 5 + r.below(1996);
        mono.resize(p);
        for (auto &c : mono) c = (uint8_t)BASES[r.below(4)];
        double div = 0.005 + 0.025 * r.uni();
        uint64_t w = at, end = at + alen;
        while (w < end) {
            uint64_t got = mutate_copy(mono.data(), p, g + w, end - w, div, r);
            w += got ? got : 1;
        }
        done += alen;
    }
}

// Overlay segmental duplications on about `frac` of the genome: segments of 10-200 kb copied from a
// random source position (any contig) with 0.5-2 % divergence, random strand.
void mqsim_add_segdups(uint64_t seed, uint8_t *g, uint64_t n, double frac) {
    if (frac <= 0 || n < 1000000) return;
    Rng r(seed ^ 0x5E6D0B5ULL, 7);
    uint64_t target = (uint64_t)(frac * (double)n), done = 0;
    std::vector<uint8_t> tmp;
    while (done < target) {
        uint64_t len = 10000 + r.below(190000);
        uint64_t src = r.below(n - len), dst = r.below(n - len);
        tmp.assign(g + src, g + src + len);
        if (r.uni() < 0.5) {
            std::reverse(tmp.begin(), tmp.end());
            for (auto &c : tmp) c = comp(c);
        }
        double div = 0.005 + 0.015 * r.uni();
        uint64_t w = mutate_copy(tmp.data(), len, g + dst, len, div, r);
        (void)w;
        done += len;
    }
}

// Fill `frac` of the genome with copies of `n_fam` repeat families (consensus 5-12 kb, each copy
// diverged 0.1-10 % (skewed towards young copies) from its family consensus, random strand, 30 % truncated) -- the
// maize-like config (BASELINE.json configs[3]).
void mqsim_add_repeat_families(uint64_t seed, uint8_t *g, uint64_t n, double frac, uint32_t n_fam) {
    if (frac <= 0 || n_fam == 0 || n < 50000) return;
    Rng r(seed ^ 0xFA111E5ULL, 11);
    std::vector<std::vector<uint8_t>> fam(n_fam);
    for (auto &f : fam) {
        f.resize(5000 + r.below(7001));
        for (auto &c : f) c = (uint8_t)BASES[r.below(4)];
    }
    uint64_t target = (uint64_t)(frac * (double)n), done = 0;
    std::vector<uint8_t> tmp;
    while (done < target) {
        const auto &f = fam[r.below(n_fam)];
        uint64_t len = f.size(), off = 0;
        if (r.uni() < 0.3) { len = 500 + r.below(f.size() - 500); off = r.below(f.size() - len + 1); }
        if (len >= n) continue;
        tmp.assign(f.begin() + off, f.begin() + off + len);
        if (r.uni() < 0.5) {
            std::reverse(tmp.begin(), tmp.end());
            for (auto &c : tmp) c = comp(c);
        }
        double u = r.uni();
        double div = 0.001 + 0.099 * u * u * u;      // skewed young: median ~1.3 %, tail to 10 %
        uint64_t dst = r.below(n - len);
        mutate_copy(tmp.data(), len, g + dst, len, div, r);
        done += len;
    }
}

// ---- reads ----------------------------------------------------------------------------------------
struct mqsim_read_cfg {
    uint64_t seed;
    double mean_len, sd_len;
    uint64_t min_len;
    double error_rate;   // total, split 1:1:1 substitution / insertion / deletion
};

// Errors are placed by drawing the error-free stretch before each one from a geometric distribution (one draw per
// error instead of one per base); both passes replay the same stream.
static inline uint64_t clean_stretch(Rng &e, double inv_log1mp) {
    double u = e.uni();
    if (u < 1e-300) u = 1e-300;
    const double g = std::floor(std::log(u) * inv_log1mp);
    return g > 4.0e18 ? (uint64_t)4e18 : (uint64_t)g;
}
static inline void copy_template(uint8_t *dst, const uint8_t *src, uint64_t len, uint64_t j, uint64_t n, bool rc) {
    if (!rc) { memcpy(dst, src + j, n); return; }
    static const struct CT { uint8_t t[256]; CT() { for (int i = 0; i < 256; i++) t[i] = comp((uint8_t)i); } } ct;
    const uint8_t *p = src + (len - 1 - j);            // template base j on the reverse strand
    for (uint64_t i = 0; i < n; i++) dst[i] = ct.t[p[-(int64_t)i]];
}

// Pass 1: choose template (contig, start, len, strand) for reads [first, first+n) and compute the
// length of every simulated read.  contig_offs has n_contigs+1 entries into `genome`.
void mqsim_reads_plan(const mqsim_read_cfg *cfg, const uint64_t *contig_offs, uint32_t n_contigs,
                      uint64_t first, uint64_t n, uint32_t *t_contig, uint64_t *t_start, uint64_t *t_len,
                      uint8_t *t_strand, uint64_t *out_len) {
    const uint64_t total = contig_offs[n_contigs];
    const double p = cfg->error_rate, ilog = p > 0 && p < 1 ? 1.0 / std::log(1.0 - p) : 0.0;
    #pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < (int64_t)n; ii++) {
        uint64_t i = first + (uint64_t)ii;
        Rng r(cfg->seed, i * 2 + 1);
        uint32_t c; uint64_t clen;
        for (;;) {  // contig chosen proportionally to its length
            uint64_t x = r.below(total);
            c = (uint32_t)(std::upper_bound(contig_offs, contig_offs + n_contigs + 1, x) - contig_offs - 1);
            clen = contig_offs[c + 1] - contig_offs[c];
            if (clen >= cfg->min_len) break;
        }
        double L = cfg->mean_len + cfg->sd_len * r.normal();
        uint64_t len = L < (double)cfg->min_len ? cfg->min_len : (uint64_t)L;
        if (len > clen) len = clen;
        uint64_t start = r.below(clen - len + 1);
        uint8_t strand = (uint8_t)(r.next() & 1);
        t_contig[ii] = c; t_start[ii] = start; t_len[ii] = len; t_strand[ii] = strand;
        Rng e(cfg->seed, i * 2 + 2);   // error stream: replayed identically in pass 2
        uint64_t w = 0, j = 0;
        while (j < len) {
            const uint64_t run = p > 0 ? std::min(len - j, clean_stretch(e, ilog)) : len - j;
            w += run; j += run;
            if (j >= len) break;
            const uint64_t t = e.below(3);
            if (t == 0) { e.below(4); e.below(4); e.below(4); w++; }   // substitution draws (bounded replay)
            else if (t == 1) { e.below(4); w += 2; }                    // insertion
            /* t == 2: deletion */
            j++;
        }
        out_len[ii] = w;
    }
}

// Pass 2: write reads at out + out_offs[i] (out_offs = exclusive prefix of out_len).
void mqsim_reads_fill(const mqsim_read_cfg *cfg, const uint8_t *genome, const uint64_t *contig_offs,
                      uint64_t first, uint64_t n, const uint32_t *t_contig, const uint64_t *t_start,
                      const uint64_t *t_len, const uint8_t *t_strand, const uint64_t *out_offs, uint8_t *out) {
    const double p = cfg->error_rate, ilog = p > 0 && p < 1 ? 1.0 / std::log(1.0 - p) : 0.0;
    #pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < (int64_t)n; ii++) {
        uint64_t i = first + (uint64_t)ii;
        const uint8_t *src = genome + contig_offs[t_contig[ii]] + t_start[ii];
        const uint64_t len = t_len[ii];
        const bool rc = t_strand[ii] != 0;
        uint8_t *dst = out + out_offs[ii];
        Rng e(cfg->seed, i * 2 + 2);
        uint64_t w = 0, j = 0;
        while (j < len) {
            const uint64_t run = p > 0 ? std::min(len - j, clean_stretch(e, ilog)) : len - j;
            copy_template(dst + w, src, len, j, run, rc);
            w += run; j += run;
            if (j >= len) break;
            const uint8_t b = rc ? comp(src[len - 1 - j]) : src[j];
            const uint64_t t = e.below(3);
            if (t == 0) {   // substitution: three fixed draws pick a base != b
                uint64_t d0 = e.below(4), d1 = e.below(4), d2 = e.below(4);
                uint8_t nb = (uint8_t)BASES[d0];
                if (nb == b) nb = (uint8_t)BASES[d1];
                if (nb == b) nb = (uint8_t)BASES[d2];
                if (nb == b) nb = (uint8_t)BASES[(d2 + 1) & 3];
                dst[w++] = nb;
            } else if (t == 1) { dst[w++] = (uint8_t)BASES[e.below(4)]; dst[w++] = b; }
            j++;
        }
    }
}

}  // extern "C"
