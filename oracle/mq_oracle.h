/*
 * mq_oracle.h -- CPU oracle for the mapquik seeding->chaining hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (mapquik_b200/, include/mapquik_b200.h) never links or calls it.
 *
 * What it restates (all citations into the upstream reference tree):
 *   - src/mers.rs:15-183   ref_extract / extract / chain_matches / find_matches /
 *                          determine_best_match / find_largest_two_chains / find_coords
 *   - src/index.rs:43-128  Entry, Index::add_with_mer (unique-or-tombstone), get_count,
 *                          ReadOnlyIndex::get
 *   - src/match.rs:10-58   Match::new / update / check (with its operator-precedence
 *                          behaviour) / extend
 *   - src/chain.rs:43-169  check_match_compatible, find_largest_match,
 *                          filter_matches_max, gap tests (i32 casts), get_match
 *
 * PARITY STATUS of the seeding stage (S1/S2): **parity unpinned**.  The l-mer
 * hashing / HPC / density sampling / k-min-mer canonicalisation+hash live in the
 * git dependency `rust-seq2kminmers` (Cargo.toml:30), which is neither vendored
 * nor pinned (no rev/tag; Cargo.lock is git-ignored), and no Rust toolchain
 * exists here.  The oracle therefore implements the seeding spec written down in
 * DESIGN.md section 2 (ntHash-1 64-bit canonical hash, published algorithm of
 * Mohamadi et al. 2016, pinned by the known-answer vectors of the `nthash` crate
 * in tests/test_oracle.py; everything around it is this repo's own documented
 * choice).  Everything downstream of the k-min-mer list follows the reference
 * sources line by line and is pinned by the reference's golden PAF line
 * (experiments/intersect_pafs.py:14).
 */
#ifndef MQ_ORACLE_H
#define MQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Params fields read by the hot path: mers.rs:16-26, chain.rs:152,158. */
typedef struct {
    uint32_t k;        /* k-min-mer length            main.rs:174 default 5  */
    uint32_t l;        /* l-mer (minimizer) length    main.rs:175 default 31 */
    double   density;  /* FH density                  main.rs:183 default 0.01 */
    uint32_t use_hpc;  /* HPC on by default           main.rs:185 */
    uint32_t c;        /* min chain length            main.rs:176 default 4  */
    uint32_t s;        /* min matching seeds          main.rs:177 default 11 */
    uint32_t g;        /* max gap difference          main.rs:178 default 2000 */
} orc_params;

/* One k-min-mer as yielded by KminmersIterator (fields named at index.rs:57-58,
 * match.rs:22-27). */
typedef struct {
    uint64_t start, end, offset, hash;
    uint32_t rev;
    uint32_t pad_;
} orc_kminmer;

/* Match (match.rs:10-17) plus the ref id of its head entry (mers.rs:68). */
typedef struct {
    uint64_t q_start, q_end, r_start, r_end, count;
    uint32_t rc;
    uint32_t ref_id;
} orc_match;

/* find_matches result: the numbers find_coords prints (mers.rs:131-183). */
typedef struct {
    uint8_t  mapped, rc, mapq, pad_;
    uint32_t ref_idx;
    uint64_t q_start, q_end, r_start, r_end, score;
} orc_hit;

typedef struct orc_index orc_index;

uint64_t orc_hash_bound(double density);
uint64_t orc_nthash_fwd(const uint8_t *s, size_t l);   /* definitional, O(l) */
uint64_t orc_nthash_rev(const uint8_t *s, size_t l);
uint64_t orc_kminmer_hash(const uint64_t *mers, uint32_t k, uint32_t *rev_out);

/* S1: selected minimizers (raw position, canonical hash).  Returns count; if
 * pos/hash are non-NULL they must hold `cap` entries (call once with NULL to size). */
size_t orc_minimizers(const uint8_t *seq, size_t n, const orc_params *p,
                      uint64_t *pos, uint64_t *hash, size_t cap);
/* the same scan over a materialised compressed sequence (rolling) -- cross-check of the streaming form */
size_t orc_minimizers_rolling(const uint8_t *seq, size_t n, const orc_params *p,
                              uint64_t *pos, uint64_t *hash, size_t cap);
/* same, but every l-mer hash is recomputed from the definition (no rolling) */
size_t orc_minimizers_slow(const uint8_t *seq, size_t n, const orc_params *p,
                           uint64_t *pos, uint64_t *hash, size_t cap);
/* S1+S2: k-min-mers of one sequence (returns 0 if n < l+k-1, mers.rs:18,44). */
size_t orc_kminmers(const uint8_t *seq, size_t n, const orc_params *p,
                    orc_kminmer *out, size_t cap);

orc_index *orc_index_new(size_t capacity_hint);
void       orc_index_free(orc_index *);
/* mers.rs:15 ref_extract: returns number of k-min-mers emitted */
uint64_t   orc_ref_extract(orc_index *, uint32_t ref_idx, const uint8_t *seq, size_t n,
                           const orc_params *p);
/* index.rs:100 add_with_mer on an explicit tuple (used by order-independence tests) */
void       orc_index_add(orc_index *, uint64_t hash, uint32_t id, uint64_t start,
                         uint64_t end, uint64_t offset, uint32_t rc);
uint64_t   orc_index_count(const orc_index *);          /* index.rs:90 get_count */
uint64_t   orc_index_slots(const orc_index *);          /* keys incl. tombstones */
/* index.rs:118 ReadOnlyIndex::get: 1 = present and not tombstone */
int        orc_index_get(const orc_index *, uint64_t hash, uint32_t *id, uint64_t *start,
                         uint64_t *end, uint64_t *offset, uint32_t *rc);

/* mers.rs:57 chain_matches, flattened in query order. Returns count. */
size_t orc_chain_matches(const orc_index *, const uint8_t *seq, size_t n,
                         const orc_params *p, orc_match *out, size_t cap);
/* mers.rs:77 find_matches; ref_lens[ref_idx] = reference sequence lengths. */
int    orc_find_matches(const orc_index *, const uint8_t *seq, size_t n,
                        const uint64_t *ref_lens, uint32_t n_refs,
                        const orc_params *p, orc_hit *out);

/* the same two steps on hand-crafted inputs (unit tests of the Match / Chain rules) */
size_t orc_chain_matches_kms(const orc_index *, const orc_kminmer *km, size_t q, orc_match *out, size_t cap);
int    orc_find_matches_kms(const orc_index *, const orc_kminmer *km, size_t q, uint64_t q_len,
                            const uint64_t *ref_lens, uint32_t n_refs, const orc_params *p, orc_hit *out);
int    orc_best_of_matches(const orc_match *ms, size_t nm, uint64_t q_len, const uint64_t *ref_lens,
                           uint32_t n_refs, const orc_params *p, orc_hit *out);

/* batch forms (OpenMP: one task per record like closures.rs:85,183) */
void orc_index_add_batch(orc_index *, const uint8_t *seqs, const uint64_t *offs, uint32_t n,
                         uint32_t first_ref_idx, const orc_params *p, uint64_t *nb_mers_out,
                         int threads);
void orc_map_batch(const orc_index *, const uint8_t *seqs, const uint64_t *offs, uint32_t n,
                   const uint64_t *ref_lens, uint32_t n_refs, const orc_params *p,
                   orc_hit *out, int threads);

/* mers.rs:181: the 12-column PAF line (no newline). Returns length or -1. */
int orc_format_paf(char *buf, size_t cap, const char *q_id, uint64_t q_len,
                   const char *r_id, uint64_t r_len, const orc_hit *h);
/* pure find_coords on a PseudoChainCoords tuple (mers.rs:131-179) */
void orc_find_coords(uint64_t q_len, uint64_t r_len, int rc, uint64_t q_start, uint64_t q_end,
                     uint64_t r_start, uint64_t r_end, uint64_t *fq_s, uint64_t *fq_e,
                     uint64_t *fr_s, uint64_t *fr_e);
int orc_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
