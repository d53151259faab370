/*
 * mq_oracle.c -- CPU oracle (plain C + OpenMP) for the mapquik hot path.
 * TEST INFRASTRUCTURE ONLY -- see mq_oracle.h for who may load it and for the
 * parity status ("parity unpinned" for the external seeding crate).
 *
 * Reference citations are file:line into the upstream tree (read-only at
 * /root/reference in the build container).
 */
#include "mq_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* S1: ntHash-1 (Mohamadi et al. 2016), seeds as in the `nthash` crate that   */
/* rust-seq2kminmers builds on; N (and any other non-ACGT byte) hashes to 0.  */
/* ------------------------------------------------------------------------- */
#define SEED_A 0x3c8bfbb395c60474ULL
#define SEED_C 0x3193c18562a02b4cULL
#define SEED_G 0x20323ed082572324ULL
#define SEED_T 0x295549f54be24456ULL

static inline uint64_t rol64(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
static inline uint64_t ror64(uint64_t x, unsigned r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }

/* Bytes are taken as given: the caller upper-cases first, exactly like the reference's I/O layer
 * does before it calls in (closures.rs:63,106).  Anything that is not A/C/G/T hashes as 0. */

static inline uint64_t seed_f(uint8_t c) {
    switch (c) { case 'A': return SEED_A; case 'C': return SEED_C;
                 case 'G': return SEED_G; case 'T': return SEED_T; default: return 0; }
}
static inline uint64_t seed_r(uint8_t c) {   /* seed of the complement base */
    switch (c) { case 'A': return SEED_T; case 'C': return SEED_G;
                 case 'G': return SEED_C; case 'T': return SEED_A; default: return 0; }
}

uint64_t orc_nthash_fwd(const uint8_t *s, size_t l) {
    uint64_t v = 0;
    for (size_t i = 0; i < l; i++) v ^= rol64(seed_f(s[i]), (unsigned)((l - 1 - i) & 63));
    return v;
}
uint64_t orc_nthash_rev(const uint8_t *s, size_t l) {
    uint64_t v = 0;
    for (size_t i = 0; i < l; i++) v ^= rol64(seed_r(s[i]), (unsigned)(i & 63));
    return v;
}

/* hash_bound = (density as FH * H::MAX as FH) as H, Rust `as` saturates. */
uint64_t orc_hash_bound(double density) {
    double b = density * 18446744073709551615.0; /* literal rounds to 2^64 like u64::MAX as f64 */
    if (!(b > 0.0)) return 0;
    if (b >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)b;
}

/* Homopolymer compression: one symbol per maximal run of equal bytes; the symbol keeps the raw index of the run's first base. */
static size_t compress(const uint8_t *seq, size_t n, int use_hpc, uint8_t *sym, uint64_t *raw) {
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t c = seq[i];
        if (use_hpc && i > 0 && c == seq[i - 1]) continue;
        sym[m] = c; raw[m] = i; m++;
    }
    return m;
}

static size_t minimizers_impl(const uint8_t *seq, size_t n, const orc_params *p,
                              uint64_t *pos, uint64_t *hash, size_t cap, int slow) {
    const size_t l = p->l;
    if (n == 0 || l == 0) return 0;
    uint8_t  *sym = (uint8_t *)malloc(n);
    uint64_t *raw = (uint64_t *)malloc(n * sizeof(uint64_t));
    size_t m = compress(seq, n, (int)p->use_hpc, sym, raw);
    size_t cnt = 0;
    const uint64_t bound = orc_hash_bound(p->density);
    if (m >= l) {
        uint64_t f = orc_nthash_fwd(sym, l), r = orc_nthash_rev(sym, l);
        for (size_t i = 0;; i++) {
            uint64_t h = f < r ? f : r;       /* canonical = min(fwd, rev) */
            if (h < bound) {                  /* universe minimizer: hash below density*max */
                if (pos && cnt < cap) { pos[cnt] = raw[i]; hash[cnt] = h; }
                cnt++;
            }
            if (i + l >= m) break;
            if (slow) {
                f = orc_nthash_fwd(sym + i + 1, l); r = orc_nthash_rev(sym + i + 1, l);
            } else {
                uint8_t out = sym[i], in = sym[i + l];
                f = rol64(f, 1) ^ rol64(seed_f(out), (unsigned)(l & 63)) ^ seed_f(in);
                r = ror64(r, 1) ^ ror64(seed_r(out), 1) ^ rol64(seed_r(in), (unsigned)((l - 1) & 63));
            }
        }
    }
    free(sym); free(raw);
    return cnt;
}
/* streaming form of the same scan (ring buffer of the last l symbols): one pass, used by the
 * index / mapping paths so that the oracle is also a fair CPU baseline. */
typedef struct { uint64_t *pos, *hash; size_t n, cap; } mvec;
static void mvec_push(mvec *v, uint64_t pos, uint64_t hash) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 256;
        v->pos = (uint64_t *)realloc(v->pos, v->cap * 8); v->hash = (uint64_t *)realloc(v->hash, v->cap * 8);
    }
    v->pos[v->n] = pos; v->hash[v->n] = hash; v->n++;
}
static void minimizers_stream(const uint8_t *seq, size_t n, const orc_params *p, mvec *out) {
    const size_t l = p->l;
    if (n == 0 || l == 0) return;
    const uint64_t bound = orc_hash_bound(p->density);
    const int hpc = (int)p->use_hpc;
    /* per-byte seed tables for this l (same values the switch-based seed_f/seed_r give) */
    uint64_t t_in_f[256], t_out_f[256], t_in_r[256], t_out_r[256];
    for (int b = 0; b < 256; b++) {
        t_in_f[b] = seed_f((uint8_t)b);                 t_out_f[b] = rol64(seed_f((uint8_t)b), (unsigned)(l & 63));
        t_in_r[b] = rol64(seed_r((uint8_t)b), (unsigned)((l - 1) & 63)); t_out_r[b] = ror64(seed_r((uint8_t)b), 1);
    }
    uint8_t ring_sym[256]; uint64_t ring_raw[256];
    uint8_t *rs = ring_sym; uint64_t *rr = ring_raw;
    if (l > 256) { rs = (uint8_t *)malloc(l); rr = (uint64_t *)malloc(l * 8); }
    size_t m = 0, slot = 0;       /* symbols seen so far; slot = m % l */
    uint64_t f = 0, r = 0;
    int prev = -1;
    for (size_t i = 0; i < n; i++) {
        const uint8_t c = seq[i];
        if (hpc && (int)c == prev) continue;
        prev = c;
        if (m < l) {              /* still filling the first window */
            f ^= rol64(seed_f(c), (unsigned)((l - 1 - m) & 63));
            r ^= rol64(seed_r(c), (unsigned)(m & 63));
        } else {
            const uint8_t out_c = rs[slot];
            f = rol64(f, 1) ^ t_out_f[out_c] ^ t_in_f[c];
            r = ror64(r, 1) ^ t_out_r[out_c] ^ t_in_r[c];
        }
        rs[slot] = c; rr[slot] = i; m++;
        if (++slot == l) slot = 0;
        if (m >= l) {
            const uint64_t h = f < r ? f : r;
            if (h < bound) mvec_push(out, rr[slot], h);   /* oldest symbol in the ring = window start */
        }
    }
    if (l > 256) { free(rs); free(rr); }
}

size_t orc_minimizers(const uint8_t *seq, size_t n, const orc_params *p,
                      uint64_t *pos, uint64_t *hash, size_t cap) {
    mvec v = {0, 0, 0, 0};
    minimizers_stream(seq, n, p, &v);
    for (size_t i = 0; i < v.n && pos && i < cap; i++) { pos[i] = v.pos[i]; hash[i] = v.hash[i]; }
    free(v.pos); free(v.hash);
    return v.n;
}
/* materialised rolling form (kept as a cross-check of the streaming one) */
size_t orc_minimizers_rolling(const uint8_t *seq, size_t n, const orc_params *p,
                      uint64_t *pos, uint64_t *hash, size_t cap) {
    return minimizers_impl(seq, n, p, pos, hash, cap, 0);
}
size_t orc_minimizers_slow(const uint8_t *seq, size_t n, const orc_params *p,
                           uint64_t *pos, uint64_t *hash, size_t cap) {
    return minimizers_impl(seq, n, p, pos, hash, cap, 1);
}

/* ------------------------------------------------------------------------- */
/* S2: k-min-mer = k consecutive selected minimizers.                         */
/* ------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z) {   /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* canonical orientation = lexicographically smaller of (w, reverse(w)); a
 * palindromic vector is "forward".  hash folds the canonical vector. */
uint64_t orc_kminmer_hash(const uint64_t *w, uint32_t k, uint32_t *rev_out) {
    uint32_t rev = 0;
    for (uint32_t i = 0; i < k; i++) {
        uint64_t a = w[i], b = w[k - 1 - i];
        if (a < b) break;
        if (a > b) { rev = 1; break; }
    }
    uint64_t h = 0x9E3779B97F4A7C15ULL ^ (uint64_t)k;
    for (uint32_t i = 0; i < k; i++) h = mix64(h ^ (rev ? w[k - 1 - i] : w[i]));
    if (rev_out) *rev_out = rev;
    return h;
}

static size_t kminmers_alloc(const uint8_t *seq, size_t n, const orc_params *p, orc_kminmer **out) {
    const size_t k = p->k, l = p->l;
    *out = NULL;
    if (k == 0 || l == 0 || n < l + k - 1) return 0;            /* mers.rs:18,44 */
    mvec v = {0, 0, 0, 0};
    minimizers_stream(seq, n, p, &v);
    size_t q = v.n >= k ? v.n - k + 1 : 0;
    if (q) {
        orc_kminmer *km = (orc_kminmer *)malloc(q * sizeof(orc_kminmer));
        for (size_t j = 0; j < q; j++) {
            uint32_t rev;
            km[j].hash = orc_kminmer_hash(v.hash + j, (uint32_t)k, &rev);
            km[j].start = v.pos[j];
            km[j].end = v.pos[j + k - 1] + l;     /* exclusive; consumers use end-1 (chain.rs:166-167) */
            km[j].offset = j;                     /* 0-based ordinal in this sequence (index.rs:47) */
            km[j].rev = rev; km[j].pad_ = 0;
        }
        *out = km;
    }
    free(v.pos); free(v.hash);
    return q;
}
size_t orc_kminmers(const uint8_t *seq, size_t n, const orc_params *p, orc_kminmer *out, size_t cap) {
    orc_kminmer *km; size_t q = kminmers_alloc(seq, n, p, &km);
    for (size_t j = 0; j < q && out && j < cap; j++) out[j] = km[j];
    free(km);
    return q;
}

/* ------------------------------------------------------------------------- */
/* Index (index.rs): sharded open-addressing map, identity-hashed by KH.      */
/* ------------------------------------------------------------------------- */
typedef struct { uint64_t key, start, end, offset; uint32_t id; uint8_t rc, used; uint16_t pad_; } slot_t;
typedef struct { slot_t *slots; size_t cap, used; volatile int lock; } shard_t;
#define NSHARD_BITS 10
#define NSHARD (1u << NSHARD_BITS)
struct orc_index { shard_t sh[NSHARD]; };

static inline size_t shard_of(uint64_t h) { return (size_t)(h >> (64 - NSHARD_BITS)); }
static inline void lock(shard_t *s)   { while (__sync_lock_test_and_set(&s->lock, 1)) { while (s->lock) ; } }
static inline void unlock(shard_t *s) { __sync_lock_release(&s->lock); }

orc_index *orc_index_new(size_t hint) {
    orc_index *ix = (orc_index *)calloc(1, sizeof(orc_index));
    size_t per = hint / NSHARD * 2 + 16, cap = 16;
    while (cap < per) cap <<= 1;
    for (size_t i = 0; i < NSHARD; i++) {
        ix->sh[i].cap = cap; ix->sh[i].slots = (slot_t *)calloc(cap, sizeof(slot_t));
    }
    return ix;
}
void orc_index_free(orc_index *ix) {
    if (!ix) return;
    for (size_t i = 0; i < NSHARD; i++) free(ix->sh[i].slots);
    free(ix);
}
static slot_t *probe(slot_t *t, size_t cap, uint64_t key) {
    size_t i = (size_t)(key * 0x9E3779B97F4A7C15ULL >> 20) & (cap - 1);
    while (t[i].used && t[i].key != key) i = (i + 1) & (cap - 1);
    return &t[i];
}
static void grow(shard_t *s) {
    size_t nc = s->cap * 2; slot_t *nt = (slot_t *)calloc(nc, sizeof(slot_t));
    for (size_t i = 0; i < s->cap; i++) if (s->slots[i].used) *probe(nt, nc, s->slots[i].key) = s->slots[i];
    free(s->slots); s->slots = nt; s->cap = nc;
}
/* index.rs:100-104: old = insert(h, entry); if old.is_some() { insert(h, Entry::empty()) } */
void orc_index_add(orc_index *ix, uint64_t h, uint32_t id, uint64_t start, uint64_t end,
                   uint64_t offset, uint32_t rc) {
    shard_t *s = &ix->sh[shard_of(h)];
    lock(s);
    if ((s->used + 1) * 10 > s->cap * 7) grow(s);
    slot_t *e = probe(s->slots, s->cap, h);
    if (!e->used) {
        e->used = 1; e->key = h; e->id = id; e->start = start; e->end = end; e->offset = offset;
        e->rc = (uint8_t)rc; s->used++;
    } else {  /* Entry::empty(): id 0, start 0, end 0, offset 0, rc false */
        e->id = 0; e->start = 0; e->end = 0; e->offset = 0; e->rc = 0;
    }
    unlock(s);
}
uint64_t orc_index_count(const orc_index *ix) {   /* non-tombstone entries: !is_empty <=> end != 0 */
    uint64_t c = 0;
    for (size_t i = 0; i < NSHARD; i++)
        for (size_t j = 0; j < ix->sh[i].cap; j++)
            if (ix->sh[i].slots[j].used && ix->sh[i].slots[j].end != 0) c++;
    return c;
}
uint64_t orc_index_slots(const orc_index *ix) {
    uint64_t c = 0;
    for (size_t i = 0; i < NSHARD; i++) c += ix->sh[i].used;
    return c;
}
static const slot_t *index_get(const orc_index *ix, uint64_t h) {   /* index.rs:118-126 */
    const shard_t *s = &ix->sh[shard_of(h)];
    const slot_t *e = probe(s->slots, s->cap, h);
    if (e->used && e->end != 0) return e;
    return NULL;
}
int orc_index_get(const orc_index *ix, uint64_t h, uint32_t *id, uint64_t *start, uint64_t *end,
                  uint64_t *offset, uint32_t *rc) {
    const slot_t *e = index_get(ix, h);
    if (!e) return 0;
    if (id) *id = e->id;
    if (start) *start = e->start;
    if (end) *end = e->end;
    if (offset) *offset = e->offset;
    if (rc) *rc = e->rc;
    return 1;
}

uint64_t orc_ref_extract(orc_index *ix, uint32_t ref_idx, const uint8_t *seq, size_t n, const orc_params *p) {
    orc_kminmer *km; size_t q = kminmers_alloc(seq, n, p, &km);   /* mers.rs:15-38 */
    if (!q) return 0;
    for (size_t j = 0; j < q; j++)
        orc_index_add(ix, km[j].hash, ref_idx, km[j].start, km[j].end, km[j].offset, km[j].rev);
    free(km);
    return q;
}

/* ------------------------------------------------------------------------- */
/* Match (match.rs) + chain_matches (mers.rs:57-73)                           */
/* ------------------------------------------------------------------------- */
/* match.rs:39-43, parsed as Rust parses it: (A && B && (rc && D1)) || (!rc && D2) */
static int match_check(const orc_match *m, const orc_kminmer *q, const slot_t *r, const slot_t *p) {
    int32_t po = (int32_t)(uint32_t)p->offset, ro = (int32_t)(uint32_t)r->offset;  /* `as i32` truncation */
    int d1 = ((uint32_t)po - (uint32_t)ro) == 1u;       /* wrapping sub, release build */
    int d2 = ((uint32_t)ro - (uint32_t)po) == 1u;
    int A = r->id == p->id;
    int B = ((q->rev != (uint32_t)r->rc) ? 1 : 0) == (int)m->rc;
    return (A && B && (m->rc && d1)) || (!m->rc && d2);
}

/* core of chain_matches over an explicit k-min-mer list (also exported for unit tests that
 * craft k-min-mers by hand) */
static size_t chain_matches_kms(const orc_index *ix, const orc_kminmer *km, size_t q, orc_match **outp) {
    *outp = NULL;
    if (!q) return 0;
    orc_match *out = (orc_match *)malloc(q * sizeof(orc_match)); size_t cap = q;
    size_t nm = 0, i = 0;
    while (i < q) {                                   /* while let Some(q) = query_it.next() */
        const slot_t *r = index_get(ix, km[i].hash);
        const orc_kminmer *qq = &km[i];
        i++;
        if (!r) continue;
        orc_match m;                                  /* Match::new, match.rs:20-29 */
        m.q_start = qq->start; m.q_end = qq->end; m.r_start = r->start; m.r_end = r->end;
        m.count = 1; m.rc = (qq->rev != (uint32_t)r->rc); m.ref_id = r->id;   /* filed under head's id, mers.rs:68 */
        const slot_t *pe = r;
        for (;;) {                                    /* Match::extend, match.rs:45-58 (recursion unrolled) */
            if (i >= q) break;                        /* peek() == None */
            const slot_t *r2 = index_get(ix, km[i].hash);
            if (!r2) { i++; break; }                  /* miss: consume and stop */
            if (!match_check(&m, &km[i], r2, pe)) break;   /* hit but not extendable: NOT consumed */
            if (m.rc) m.r_start = r2->start; else m.r_end = r2->end;   /* update, match.rs:31-37 */
            m.q_end = km[i].end; m.count += 1;
            i++; pe = r2;
        }
        if (nm < cap) out[nm] = m;
        nm++;
    }
    *outp = out;
    return nm;
}
static size_t chain_matches_alloc(const orc_index *ix, const uint8_t *seq, size_t n, const orc_params *p,
                                  orc_match **outp) {
    *outp = NULL;
    orc_kminmer *km; size_t q = kminmers_alloc(seq, n, p, &km);
    size_t nm = chain_matches_kms(ix, km, q, outp);
    free(km);
    return nm;
}
size_t orc_chain_matches_kms(const orc_index *ix, const orc_kminmer *km, size_t q, orc_match *out, size_t cap) {
    orc_match *ms; size_t nm = chain_matches_kms(ix, km, q, &ms);
    for (size_t i = 0; i < nm && out && i < cap; i++) out[i] = ms[i];
    free(ms);
    return nm;
}
size_t orc_chain_matches(const orc_index *ix, const uint8_t *seq, size_t n, const orc_params *p,
                         orc_match *out, size_t cap) {
    orc_match *ms; size_t nm = chain_matches_alloc(ix, seq, n, p, &ms);
    for (size_t i = 0; i < nm && out && i < cap; i++) out[i] = ms[i];
    free(ms);
    return nm;
}

/* ------------------------------------------------------------------------- */
/* Chain (chain.rs)                                                           */
/* ------------------------------------------------------------------------- */
static int match_eq(const orc_match *a, const orc_match *b) {  /* derive(PartialEq), match.rs:8 */
    return a->q_start == b->q_start && a->q_end == b->q_end && a->r_start == b->r_start &&
           a->r_end == b->r_end && a->count == b->count && a->rc == b->rc;
}
/* chain.rs:132-142: every operand cast `usize as i32` before the subtraction; wraps in release */
static int gap_too_long(uint64_t a1, uint64_t a0, uint64_t b1, uint64_t b0, uint64_t g) {
    uint32_t g1 = (uint32_t)a1 - (uint32_t)a0;     /* a1 as i32 - a0 as i32 */
    uint32_t g2 = (uint32_t)b1 - (uint32_t)b0;
    int32_t d = (int32_t)(g1 - g2);
    int32_t ad = (d < 0) ? (int32_t)(0u - (uint32_t)d) : d;   /* i32::abs wraps at MIN in release */
    return (uint64_t)(int64_t)ad > g;                        /* `as usize` sign-extends */
}
static int compatible(const orc_match *h1, const orc_match *h2, uint64_t g) {  /* chain.rs:43-63 */
    if (match_eq(h1, h2)) return 1;
    if (h1->rc != h2->rc) return 0;
    const orc_match *u = (h1->q_start < h2->q_start) ? h1 : h2;
    const orc_match *v = (h1->q_start < h2->q_start) ? h2 : h1;
    if (u->rc) {
        if (u->r_start <= v->r_start || gap_too_long(v->q_start, u->q_end, u->r_start, v->r_end, g)) return 0;
    } else if (v->r_start <= u->r_start || gap_too_long(v->q_start, u->q_end, v->r_start, u->r_end, g)) {
        return 0;
    }
    return 1;
}

/* chain.rs:147-169 get_match on one reference's Vec<Match>; always Some for len>=1 */
static void get_match(orc_match *ms, size_t len, const orc_params *p, orc_hit *t) {
    if (len > 1) {                                    /* filter_matches_max, chain.rs:123-129 */
        size_t mx = 0; uint64_t mc = 0;
        for (size_t i = 0; i < len; i++) if (ms[i].count > mc) { mx = i; mc = ms[i].count; }
        orc_match h = ms[mx];
        size_t w = 0;
        for (size_t i = 0; i < len; i++) if (compatible(&h, &ms[i], p->g)) ms[w++] = ms[i];
        len = w;
    }
    uint64_t score = 0;
    for (size_t i = 0; i < len; i++) score += ms[i].count;
    int ok = (p->s != 0 && p->c != 0) && ((len >= p->c) || (score >= p->s));
    const orc_match *first = &ms[0], *last = &ms[len - 1];
    t->mapped = 1; t->rc = (uint8_t)first->rc; t->mapq = ok ? 60 : 0; t->pad_ = 0;
    t->q_start = first->q_start; t->q_end = last->q_end - 1; t->score = score;
    if (first->rc && len > 1) { t->r_start = last->r_start;  t->r_end = first->r_end - 1; }
    else                      { t->r_start = first->r_start; t->r_end = last->r_end - 1; }
}

/* mers.rs:131-179, usize arithmetic wraps in the release profile (Cargo.toml:42-49) */
void orc_find_coords(uint64_t q_len, uint64_t r_len, int rc, uint64_t q_start, uint64_t q_end,
                     uint64_t r_start, uint64_t r_end, uint64_t *fq_s, uint64_t *fq_e,
                     uint64_t *fr_s, uint64_t *fr_e) {
    uint64_t frs, fre, exs, exe;
    uint64_t tail = q_len - q_end - 1;
    if (!rc) {
        if (r_start >= q_start) { frs = r_start - q_start; exs = q_start; }
        else                    { frs = 0;                 exs = r_start; }
        if (r_end + tail <= r_len - 1) { fre = r_end + tail; exe = tail; }
        else                           { fre = r_len - 1;    exe = r_len - r_end - 1; }
    } else {
        if (r_end + q_start <= r_len - 1) { fre = r_end + q_start; exs = q_start; }
        else                              { fre = r_len - 1;       exs = r_len - r_end - 1; }
        if (r_start >= tail) { frs = r_start - tail; exe = tail; }
        else                 { frs = 0;              exe = r_start; }
    }
    *fq_s = q_start - exs; *fq_e = q_end + exe; *fr_s = frs; *fr_e = fre;
}

static int cmp_match_ref(const void *a, const void *b) {  /* stable by (ref_id, original order) */
    const orc_match *x = (const orc_match *)a, *y = (const orc_match *)b;
    if (x->ref_id != y->ref_id) return x->ref_id < y->ref_id ? -1 : 1;
    return 0;
}

/* mers.rs:80-92 on an explicit Match list (query order); q_len = read length */
static int best_of_matches(orc_match *ms, size_t nm, uint64_t q_len, const uint64_t *ref_lens, uint32_t n_refs,
                           const orc_params *p, orc_hit *out) {
    memset(out, 0, sizeof(*out));
    if (!nm) return 0;
    /* group per reference id keeping query order inside each group (HashMap<usize,Vec<Match>>):
     * stable insertion sort by ref id */
    for (size_t i = 1; i < nm; i++) {
        orc_match t = ms[i]; size_t j = i;
        while (j > 0 && cmp_match_ref(&ms[j - 1], &t) > 0) { ms[j] = ms[j - 1]; j--; }
        ms[j] = t;
    }
    /* determine_best_match / find_largest_two_chains (mers.rs:104-129): unique greatest score wins,
     * tie for the maximum => unmapped.  Independent of HashMap iteration order. */
    orc_hit best; memset(&best, 0, sizeof(best));
    uint64_t max_c = 0, second_c = 0; size_t groups = 0;
    for (size_t i = 0; i < nm;) {
        size_t j = i; while (j < nm && ms[j].ref_id == ms[i].ref_id) j++;
        orc_hit t; memset(&t, 0, sizeof(t));
        uint32_t rid = ms[i].ref_id;
        get_match(ms + i, j - i, p, &t);
        t.ref_idx = rid;
        groups++;
        if (t.score > max_c) { second_c = max_c; max_c = t.score; best = t; }
        else if (t.score > second_c) second_c = t.score;
        i = j;
    }
    if (groups == 0) return 0;
    if (groups > 1 && max_c == second_c) return 0;
    if (best.ref_idx >= n_refs) return 0;   /* ref_map.get().unwrap() would panic; unreachable */
    uint64_t a, b, c, d;
    orc_find_coords(q_len, ref_lens[best.ref_idx], best.rc, best.q_start, best.q_end,
                    best.r_start, best.r_end, &a, &b, &c, &d);
    best.q_start = a; best.q_end = b; best.r_start = c; best.r_end = d;
    *out = best;
    return 1;
}
int orc_find_matches(const orc_index *ix, const uint8_t *seq, size_t n, const uint64_t *ref_lens,
                     uint32_t n_refs, const orc_params *p, orc_hit *out) {
    orc_match *ms; size_t nm = chain_matches_alloc(ix, seq, n, p, &ms);
    int r = best_of_matches(ms, nm, (uint64_t)n, ref_lens, n_refs, p, out);
    free(ms);
    return r;
}
int orc_find_matches_kms(const orc_index *ix, const orc_kminmer *km, size_t q, uint64_t q_len,
                         const uint64_t *ref_lens, uint32_t n_refs, const orc_params *p, orc_hit *out) {
    orc_match *ms; size_t nm = chain_matches_kms(ix, km, q, &ms);
    int r = best_of_matches(ms, nm, q_len, ref_lens, n_refs, p, out);
    free(ms);
    return r;
}
int orc_best_of_matches(const orc_match *ms_in, size_t nm, uint64_t q_len, const uint64_t *ref_lens,
                        uint32_t n_refs, const orc_params *p, orc_hit *out) {
    orc_match *ms = (orc_match *)malloc((nm + 1) * sizeof(orc_match));
    memcpy(ms, ms_in, nm * sizeof(orc_match));
    int r = best_of_matches(ms, nm, q_len, ref_lens, n_refs, p, out);
    free(ms);
    return r;
}

void orc_index_add_batch(orc_index *ix, const uint8_t *seqs, const uint64_t *offs, uint32_t n,
                         uint32_t first_ref_idx, const orc_params *p, uint64_t *nb, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint64_t c = orc_ref_extract(ix, first_ref_idx + (uint32_t)i, seqs + offs[i],
                                     (size_t)(offs[i + 1] - offs[i]), p);
        if (nb) nb[i] = c;
    }
}
void orc_map_batch(const orc_index *ix, const uint8_t *seqs, const uint64_t *offs, uint32_t n,
                   const uint64_t *ref_lens, uint32_t n_refs, const orc_params *p, orc_hit *out,
                   int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    #pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)n; i++)
        orc_find_matches(ix, seqs + offs[i], (size_t)(offs[i + 1] - offs[i]), ref_lens, n_refs, p, &out[i]);
}

int orc_format_paf(char *buf, size_t cap, const char *q_id, uint64_t q_len, const char *r_id,
                   uint64_t r_len, const orc_hit *h) {  /* mers.rs:180-181 */
    int w = snprintf(buf, cap, "%s\t%llu\t%llu\t%llu\t%s\t%s\t%llu\t%llu\t%llu\t%llu\t%llu\t%u", q_id,
                     (unsigned long long)q_len, (unsigned long long)h->q_start, (unsigned long long)h->q_end,
                     h->rc ? "-" : "+", r_id, (unsigned long long)r_len, (unsigned long long)h->r_start,
                     (unsigned long long)h->r_end, (unsigned long long)h->score, (unsigned long long)r_len,
                     (unsigned)h->mapq);
    return (w < 0 || (size_t)w >= cap) ? -1 : w;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
