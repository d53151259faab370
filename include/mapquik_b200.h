/*
 * mapquik_b200.h -- C ABI of the B200-native mapquik seeding->chaining hot path.
 *
 * The upstream reference (ekimb/mapquik, Rust) has no FFI; its I/O layer crosses into the hot
 * path at three call sites, and this header is the batch-oriented, `extern "C"` restatement of
 * exactly those three calls (plain pointers and sizes, no torch / C++ types):
 *
 *   reference call site (file:line)                         replaced by
 *   ------------------------------------------------------  -------------------------------
 *   mers::ref_extract(ref_idx, seq, params, &index)         mq_index_add        (x n records)
 *       src/closures.rs:48  ->  src/mers.rs:15-38
 *   index.get_count() + ReadOnlyIndex::new(index.index)     mq_index_freeze
 *       src/closures.rs:92,94 -> src/index.rs:90-92,108-116
 *   mers::find_matches(q_id, q_len, q_str, ref_map,         mq_map_batch        (x n reads)
 *                      &index, params) -> Option<String>
 *       src/closures.rs:102 ->  src/mers.rs:77-92 (+ match.rs, chain.rs, find_coords)
 *
 * Contract shared with the reference:
 *   - sequences are ASCII, already upper-cased by the caller exactly as the reference does before
 *     it calls in (closures.rs:63,106 `to_ascii_uppercase`); bytes other than A/C/G/T hash as 0
 *     (ntHash "N") and still take part in homopolymer compression by byte equality;
 *   - a record shorter than l+k-1 yields nothing (mers.rs:18,44);
 *   - an index key seen more than once genome-wide is a tombstone (index.rs:100-104);
 *   - one mq_hit per read, in input order; `mapped == 0` <=> find_matches returned None;
 *     the numeric fields are the values find_coords prints (mers.rs:178-181), so the PAF line is
 *     mq_format_paf() of them -- coordinates are inclusive ends, column 11 repeats r_len.
 * Errors: the reference aborts (panic = "abort"); this library returns negative status codes and
 * never throws or aborts across the boundary.  There is NO CPU fallback: without a usable CUDA
 * device every entry point that computes returns MQ_ERR_CUDA.
 *
 * Threading: one host thread per mq_ctx at a time.  mq_create gives a context that drives one GPU; mq_create_multi one
 * that fans every call out to several GPUs of this process (closures.rs:85,183 fan out over threads at the same two
 * places): the reference is partitioned by base range, the per-GPU minimizer stores are exchanged GPU to GPU and every
 * GPU freezes the whole index; reads are sharded with no communication.  One process per GPU works too: mq_store_*
 * below is what the ranks all-gather before mq_index_freeze.
 *
 * Packed input: every entry point that takes upper-cased ASCII has a twin that takes the same sequences as 2-bit
 * codes (mq_packed).  The reference makes one pass over every record before the hot path sees it
 * (`to_ascii_uppercase`, closures.rs:63,106); a caller of this library makes that pass with mq_pack / mq_pack_at
 * instead and ships a quarter of the bytes to the GPU.  Results are identical by construction: codes + exception
 * intervals are exactly the ASCII bytes.
 */
#ifndef MAPQUIK_B200_H
#define MAPQUIK_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MQ_OK            0
#define MQ_ERR_ARG      -1   /* bad argument (k, l out of range, NULL pointer, ...)          */
#define MQ_ERR_CUDA     -2   /* CUDA runtime error / no device; see mq_last_error()          */
#define MQ_ERR_STATE    -3   /* call out of order (map before freeze, add after freeze, ...) */
#define MQ_ERR_NOMEM    -4   /* device or host allocation failed                             */
#define MQ_ERR_RANGE    -5   /* input exceeds a documented limit (record >= 2^32 bases, ...) */

#define MQ_MAX_L 32          /* 2 <= l <= 32 (the window lives in one 64-bit shift register)  */
#define MQ_MAX_K 32          /* 1 <= k <= 32                                                   */

/* Params fields the hot path reads: mers.rs:16-26 (k,l,density,use_hpc), chain.rs:152,158 (g,c,s).
 * Defaults main.rs:174-188: k=5 l=31 density=0.01 use_hpc=1 c=4 s=11 g=2000. */
typedef struct {
    uint32_t k;
    uint32_t l;
    double   density;
    uint32_t use_hpc;
    uint32_t c;
    uint32_t s;
    uint32_t g;
} mq_params;

/* The numbers of one PAF record (find_coords, mers.rs:131-183); 48 bytes. */
typedef struct {
    uint8_t  mapped;     /* 0 => read is unmapped (no PAF line)      */
    uint8_t  rc;         /* 1 => '-' strand                           */
    uint8_t  mapq;       /* 0 or 60 (chain.rs:158-161)                */
    uint8_t  pad_;
    uint32_t ref_idx;    /* index given to mq_index_add               */
    uint64_t q_start, q_end, r_start, r_end;   /* final, inclusive ends */
    uint64_t score;      /* sum of Match counts (chain.rs:157)        */
} mq_hit;

typedef struct mq_ctx mq_ctx;

/* ---- packed sequences --------------------------------------------------------------------------
 * Base i of a sequence array: code (words[i >> 4] >> 2*(i & 15)) & 3 with code = (byte >> 1) & 3 (A=0 C=1 T=2 G=3).
 * flags: bit (i >> 6) & 31 of flags[i >> 11] is set iff the 64-base block of i holds a byte other than A/C/G/T; exc lists
 * those bytes as sorted, disjoint intervals.  offs[] of the packed entry points index bases exactly as the ASCII ones
 * index bytes.  words needs mq_packed_words(n) entries (256 bytes of zeroed slack), flags mq_packed_flag_words(n). */
typedef struct { uint64_t start; uint32_t len; uint32_t byte; } mq_exc;      /* bases [start, start+len) are `byte` */
typedef struct {
    const uint32_t *words;
    const uint32_t *flags;
    const mq_exc   *exc;
    uint64_t        n_exc;
    uint64_t        n_bases;
} mq_packed;
uint64_t mq_packed_words(uint64_t n_bases);
uint64_t mq_packed_flag_words(uint64_t n_bases);
/* Pack a whole array with n_threads host threads (<= 0: all cores).  fold_case != 0 upper-cases first
 * (closures.rs:63,106).  Returns MQ_ERR_RANGE with *n_exc = the count needed when exc_cap is too small. */
int mq_pack(const uint8_t *ascii, uint64_t n_bases, uint32_t *words, uint32_t *flags, mq_exc *exc, uint64_t exc_cap,
            uint64_t *n_exc, int n_threads, int fold_case);
/* Pack n_bases bytes to base positions [at_base, at_base + n_bases) of a ZEROED destination.  Safe to call concurrently
 * for disjoint ranges (shared edge words are OR-ed in atomically) -- a FASTX parser's copy jobs call this instead of
 * their upper-casing memcpy.  exc receives this range's intervals (absolute positions); the caller concatenates the
 * lists of all ranges in position order.  Whole 32-base groups at 32-aligned positions are written with plain stores:
 * a caller that only ever passes such chunks (plus one ragged tail) needs zeros in the block bitmap and under the tail
 * only, and runs at full SIMD width -- the C++ CLI's parser stages wrapped FASTA lines into such chunks. */
int mq_pack_at(const uint8_t *ascii, uint64_t n_bases, uint64_t at_base, uint32_t *words, uint32_t *flags, mq_exc *exc,
               uint64_t exc_cap, uint64_t *n_exc, int fold_case);
int mq_unpack(const uint32_t *words, const mq_exc *exc, uint64_t n_exc, uint64_t first, uint64_t n, uint8_t *ascii);

/* ---- lifecycle ------------------------------------------------------------------------------- */
int  mq_create(mq_ctx **out, const mq_params *p, int device);
/* One context over n_devices GPUs of this process.  Entry points marked [multi] accept it. */
int  mq_create_multi(mq_ctx **out, const mq_params *p, const int *devices, int n_devices);
int  mq_device_count(const mq_ctx *);
void mq_destroy(mq_ctx *);
const char *mq_strerror(int code);
const char *mq_last_error(const mq_ctx *);       /* detail of the last failure on this ctx */
int  mq_abi_version(void);

/* Pinned host memory for the caller's sequence / hit buffers (optional but needed for full PCIe
 * speed; pageable buffers work and are staged through an internal pinned bounce buffer). */
void *mq_host_alloc(size_t bytes);
void  mq_host_free(void *);

/* ---- index build  (≙ ref_extract, closures.rs:48) ------------------------------------------- */
/* seqs: concatenated records; offs[n+1] byte offsets into seqs; record i gets id first_ref_idx+i.
 * nb_mers_out[i] (may be NULL) = number of k-min-mers the record emitted ("Indexed reference {}:
 * {} k-min-mers.", closures.rs:58). */
int mq_index_add(mq_ctx *, const uint8_t *seqs, const uint64_t *offs, uint32_t n,
                 uint32_t first_ref_idx, uint64_t *nb_mers_out);                                  /* [multi] */
int mq_index_add_packed(mq_ctx *, const mq_packed *seqs, const uint64_t *offs, uint32_t n,
                        uint32_t first_ref_idx, uint64_t *nb_mers_out);                           /* [multi] */

/* Scan one contiguous piece [seg_start, seg_start+own_len) of reference record `ref_idx` whose
 * total length is ref_len (multi-GPU partitioning by reference chunk).  `bytes` must start at
 * record position seg_start - (seg_start>0) (one byte of left context) and extend far enough to
 * the right to contain l-1 further homopolymer-run starts or the end of the record; n_bytes says
 * how far it goes.  Minimizers found go to this ctx's minimizer store. */
int mq_index_add_segment(mq_ctx *, const uint8_t *bytes, uint64_t n_bytes, uint32_t ref_idx,
                         uint64_t ref_len, uint64_t seg_start, uint64_t own_len);

/* Minimizer store (device SoA: raw position u32, canonical l-mer hash u64) + host directory of
 * (ref_idx, seg_start, count) triples in insertion order.  Export / import are how one-process-
 * per-GPU ranks exchange their share with an NCCL all-gather before mq_index_freeze. */
int mq_store_info(mq_ctx *, uint64_t *n_minimizers, uint32_t *n_segments);
int mq_store_export(mq_ctx *, void **d_pos_u32, void **d_hash_u64, uint64_t *dir_u64x3 /* n_segments*3 */);
int mq_store_import(mq_ctx *, const void *d_pos_u32, const void *d_hash_u64, uint64_t n_minimizers,
                    const uint64_t *dir_u64x3, uint32_t n_segments);
/* In-place variant for collectives: reserve room for the union (this rank's own entries stay at the front), let
 * ncclAllGather / ncclBroadcast write every rank's share into the returned arrays, then commit the merged directory. */
int mq_store_reserve(mq_ctx *, uint64_t n_minimizers, void **d_pos_u32, void **d_hash_u64);
int mq_store_commit(mq_ctx *, uint64_t n_minimizers, const uint64_t *dir_u64x3, uint32_t n_segments);

/* ≙ get_count + ReadOnlyIndex::new (closures.rs:92,94).  ref_lens[n_refs] are the record lengths
 * (the ref_map of closures.rs:29,49).  Builds the table from the store: k-min-mers are formed per
 * record over the position-ordered union of its segments, inserted with the unique-or-tombstone
 * rule.  n_unique (may be NULL) = get_count(); n_keys (may be NULL) = distinct keys incl.
 * tombstones. */
int mq_index_freeze(mq_ctx *, const uint64_t *ref_lens, uint32_t n_refs, uint64_t *n_unique,
                    uint64_t *n_keys);                                                            /* [multi] */
/* per-record k-min-mer counts of the frozen index (valid after freeze), nb[n_refs] */
int mq_index_nb_mers(mq_ctx *, uint64_t *nb, uint32_t n_refs);

/* ---- on-disk index (no counterpart upstream: the reference rebuilds its index on every run) ----
 * File = header (parameters, counts) + ref_lens + per-record k-min-mer counts + an opaque blob of the caller's
 * (the record names, e.g. NUL-separated) + the frozen table.  load needs a fresh context created with the same
 * k / l / density / hpc; afterwards the context behaves exactly as after mq_index_freeze. */
int mq_index_save(mq_ctx *, const char *path, const char *names_blob, uint64_t names_bytes);
int mq_index_load(mq_ctx *, const char *path, uint64_t *ref_lens_out, uint32_t ref_cap, uint32_t *n_refs_out,
                  char *names_out, uint64_t names_cap, uint64_t *names_bytes_out, uint64_t *n_unique_out);

/* ---- mapping  (≙ find_matches, closures.rs:102) --------------------------------------------- */
int mq_map_batch(mq_ctx *, const uint8_t *seqs, const uint64_t *offs, uint32_t n, mq_hit *out);            /* [multi] */
int mq_map_batch_packed(mq_ctx *, const mq_packed *seqs, const uint64_t *offs, uint32_t n, mq_hit *out);  /* [multi] */
/* The same with the sequences (and the hits) resident on this ctx's device: no sequence H2D / hit D2H inside.  offs is
 * a HOST array (the per-sub-batch tile tables are derived from it on the host).  d_seqs / d_seqs->words must be 16-byte
 * aligned with >= 256 readable bytes after the last record; for the packed variant words, flags and exc are device
 * pointers.  Every record < 2^31 bases. */
int mq_map_batch_device(mq_ctx *, const uint8_t *d_seqs, const uint64_t *offs, uint32_t n, mq_hit *d_out);
int mq_map_batch_packed_device(mq_ctx *, const mq_packed *d_seqs, const uint64_t *offs, uint32_t n, mq_hit *d_out);

/* mers.rs:181: 12-column PAF line, no trailing newline.  Returns its length, or <0 if cap is too
 * small.  Pure host formatting. */
int mq_format_paf(char *buf, size_t cap, const char *q_id, uint64_t q_len, const char *r_id,
                  uint64_t r_len, const mq_hit *h);

/* ---- introspection used by the parity tests and bench (stage outputs, timings) -------------- */
/* S1 only: minimizers of a batch.  seq_off[n+1] (u64) indexes pos/hash; pass NULL pos/hash to
 * size (total returned in *n_total). */
int mq_minimizers(mq_ctx *, const uint8_t *seqs, const uint64_t *offs, uint32_t n,
                  uint64_t *seq_off, uint32_t *pos, uint64_t *hash, uint64_t cap, uint64_t *n_total);
int mq_minimizers_packed(mq_ctx *, const mq_packed *seqs, const uint64_t *offs, uint32_t n,
                         uint64_t *seq_off, uint32_t *pos, uint64_t *hash, uint64_t cap, uint64_t *n_total);
/* S1+S2: k-min-mer tuples of a batch (start,end,offset<<1|rev as u32; hash u64) */
int mq_kminmers(mq_ctx *, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint64_t *seq_off,
                uint32_t *start, uint32_t *end, uint32_t *offrev, uint64_t *hash, uint64_t cap,
                uint64_t *n_total);
/* I5: probe the frozen index; found[i]=1 and the entry fields are filled when present & unique */
int mq_index_get(mq_ctx *, const uint64_t *hashes, uint64_t n, uint8_t *found, uint32_t *id,
                 uint32_t *start, uint32_t *end, uint32_t *offset, uint8_t *rc);
/* M2: the Match lists (query order) of a batch.  match_off[n+1]; fields per match:
 * q_start,q_end,r_start,r_end,count,(ref_id<<1|rc) as 6 x u32 */
int mq_matches(mq_ctx *, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint64_t *match_off,
               uint32_t *fields6, uint64_t cap, uint64_t *n_total);

/* device time of the stages of the last mq_index_* / mq_map_* call, CUDA events on the ctx stream.
 * names: "h2d","scan","scan_kernel" (k_scan_minimizers alone, also inside "scan"),"gather","insert","exchange","probe","chain","d2h","total".
 * Returns ms or <0; for a multi-GPU context the maximum over its devices. */
double mq_last_ms(mq_ctx *, const char *stage);
/* the same stages accumulated over all calls since mq_create */
double mq_total_ms(mq_ctx *, const char *stage);
/* number of kernels this library launched on this ctx since creation */
uint64_t mq_launch_count(mq_ctx *);
void *mq_stream(mq_ctx *);                /* cudaStream_t the kernels run on */
int mq_sync(mq_ctx *);
uint64_t mq_scan_kernel_launches(mq_ctx *);     /* launches of the dominant kernel (k_scan_minimizers) */
/* Host threads the mapping calls may use ([multi]: shared out over the devices).  The reference maps on `--threads` CPU
 * threads (main.rs:138-141,212, closures.rs:85,183); here they only ever touch the INPUT: with n > 0, mq_map_batch on ASCII host
 * buffers feeds the GPU from both ends of the batch -- ASCII sub-batches go over the PCIe link as they are (1 byte per
 * base) while the host threads pack sub-batches from the other end to 2 bits per base in pinned staging (the mq_pack
 * arithmetic) that cross the link at a quarter of the bytes; the split balances itself.  Results are identical either
 * way.  0 (the default) = never touch the input on the host. */
int mq_set_host_threads(mq_ctx *, int n);
/* counters of the last mapping call: "h2d_bytes" (bytes it uploaded), "sub_batches", "host_packed_sub_batches",
 * "host_packed_bases" (how much of the input took the packed route); [multi]: summed over the devices */
uint64_t mq_last_counter(mq_ctx *, const char *name);
uint64_t mq_minimizer_count(mq_ctx *, int reset); /* minimizers produced since the last reset */
/* device-memory helpers: keep inputs resident in HBM for mq_map_batch_device */
void *mq_dev_alloc(mq_ctx *, size_t bytes);
void  mq_dev_free(mq_ctx *, void *);
int   mq_dev_upload(mq_ctx *, void *dst, const void *src, size_t bytes);
int   mq_dev_download(mq_ctx *, void *dst, const void *src, size_t bytes);
int   mq_dev_memset(mq_ctx *, void *dst, int value, size_t bytes);
/* CUDA-event bracket on the ctx stream around any sequence of calls; end returns elapsed ms */
int    mq_region_begin(mq_ctx *);
double mq_region_end_ms(mq_ctx *);
uint64_t mq_table_bytes(mq_ctx *);
uint64_t mq_table_slots(mq_ctx *);

#ifdef __cplusplus
}
#endif
#endif
