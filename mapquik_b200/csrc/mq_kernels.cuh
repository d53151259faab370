// mq_kernels.cuh -- sm_100a kernels of the mapquik seeding->chaining hot path (everything but the S1 scan itself,
// which lives in mq_scan.cuh).
//
// Stage map (reference rows of SURVEY.md section 8a):
//   k_scan_minimizers   S1  (mq_scan.cuh) HPC + ntHash-1 canonical l-mer hash + universe-minimizer sampling
//                           (the KminmersIterator stage-1 the reference calls at mers.rs:27,53)
//   k_tile_prefix       S1  exclusive prefix of the per-tile minimizer counts (one launch, last block finishes)
//   k_gather_minimizers S1  ordered stream compaction of the per-tile event pools
//   k_insert_kminmers   S2+I3  window of k minimizers -> k-min-mer -> unique-or-tombstone insert
//                           (mers.rs:29-36, index.rs:100-104)
//   k_probe_match       S2+I5+M1+M2  per-read probe and Match segmentation (mers.rs:57-73,
//                           match.rs:20-58)
//   k_chain             C1-C4+P1  pseudo-chain, MAPQ, best reference, find_coords (chain.rs:43-169,
//                           mers.rs:77-183)
// All arithmetic is integer; nothing here is a contraction, so no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mq {

// ------------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------------
#ifndef MQ_CS_MAX
#define MQ_CS_MAX 128
#endif
constexpr uint32_t EV_CAP       = 2 * MQ_CS_MAX;   // staged events per tile before overflow pool (a tile holds 32 * MQ_CS_MAX bases)
constexpr uint64_t EMPTY_KEY    = 0xFFFFFFFFFFFFFFFFull;
constexpr int      MQ_MAX_K_    = 32;

constexpr uint64_t SEED_A = 0x3c8bfbb395c60474ull, SEED_C = 0x3193c18562a02b4cull,
                   SEED_G = 0x20323ed082572324ull, SEED_T = 0x295549f54be24456ull;

struct Slot {                  // 32 B = one DRAM sector
    uint64_t key;
    uint32_t id, start, end, offrc;   // offrc = offset<<1 | rc
    uint32_t count;            // number of inserts of this key; entry valid <=> count == 1
    uint32_t pad_;
};
static_assert(sizeof(Slot) == 32, "slot must be one sector");

struct MatchRec {              // 32 B, written by k_probe_match, read by k_chain
    uint32_t q_start, q_end, r_start, r_end;
    uint32_t head_j, last_j;   // count = last_j - head_j + 1
    uint32_t ref_rc;           // ref_id<<1 | rc
    uint32_t pad_;
};

struct HitRec {                // == mq_hit
    uint8_t mapped, rc, mapq, pad_;
    uint32_t ref_idx;
    uint64_t q_start, q_end, r_start, r_end, score;
};
static_assert(sizeof(HitRec) == 48, "mq_hit layout");

// Scalars of one batch, zeroed by ONE memset before its first kernel; the tail is read back by the host once per
// batch (there is no other device->host traffic between the upload of a batch and the download of its hits).
struct BatchScalars {
    uint32_t scan_ticket, ovf_count, probe_ticket, chain_ticket, big_count, prefix_done;
    uint32_t flags;            // BS_* below; any bit set => the kernels after k_tile_prefix do nothing
    uint32_t pad_;
    uint64_t n_minimizers;     // grand total of the tile counts
};
constexpr uint32_t BS_MINI_CAP = 1u, BS_OVF_CAP = 2u, BS_RANGE = 4u;

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t rol64(uint64_t x, unsigned r) { r &= 63; return (x << r) | (x >> ((64 - r) & 63)); }
__device__ __forceinline__ uint64_t rol1(uint64_t x) { return (x << 1) | (x >> 63); }
__device__ __forceinline__ uint64_t ror1(uint64_t x) { return (x >> 1) | (x << 63); }
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t *total) {
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane_id() >= (uint32_t)d) x += y; }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// ------------------------------------------------------------------------------------------------
// exclusive prefix of the per-tile counts, ONE launch.  Every block scans its 8192 counts in place and publishes its
// total; the block that finishes last scans the block totals.  The prefix of tile t is therefore read in two pieces,
// within[t] + blk[t / 8192] (tile_base below); entry n (one past the last tile) reads as the grand total.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_BLK = 1024, SCAN_ITEMS = 8;     // 8192 items per block
constexpr uint32_t PREFIX_SPAN = SCAN_BLK * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t *smem33, uint32_t *block_total) {
    uint32_t wt, e = warp_excl_scan(v, &wt);
    uint32_t w = threadIdx.x >> 5;
    if (lane_id() == 31) smem33[w] = wt;
    __syncthreads();
    if (w == 0) { uint32_t t, s = warp_excl_scan(smem33[lane_id()], &t); smem33[lane_id()] = s; if (lane_id() == 0) smem33[32] = t; }
    __syncthreads();
    e += smem33[w];
    *block_total = smem33[32];
    __syncthreads();
    return e;
}

struct TileBase { const uint32_t *within; const uint32_t *blk; };
__device__ __forceinline__ uint32_t tile_base(const TileBase &tb, uint32_t t) { return __ldg(tb.within + t) + __ldg(tb.blk + t / PREFIX_SPAN); }

// cnt[0..n) in, within[0..n] out (in place; cnt needs n+1 entries), blk[ceil((n+1)/8192)] out.  mini_cap / ovf_cap: the
// capacities of the buffers the later kernels write; exceeding one raises a flag instead of writing out of bounds.
__global__ void __launch_bounds__(SCAN_BLK) k_tile_prefix(uint32_t *cnt, uint32_t n, uint32_t *blk, BatchScalars *sc,
                                                          uint64_t mini_cap, uint32_t ovf_cap) {
    __shared__ uint32_t sm[33];
    __shared__ bool last;
    const uint32_t base = blockIdx.x * PREFIX_SPAN + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? cnt[base + i] : 0u; s += v[i]; }
    uint32_t tot, e = block_excl_scan_1024(s, sm, &tot);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i <= n) cnt[base + i] = e; e += v[i]; }
    if (threadIdx.x == 0) {
        blk[blockIdx.x] = tot;
        __threadfence();
        last = atomicAdd(&sc->prefix_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last block: exclusive scan of the block totals (64-bit carry: a batch may hold >= 2^32 minimizers, which
    // is reported, never wrapped)
    unsigned long long carry = 0;
    for (uint32_t b0 = 0; b0 < gridDim.x; b0 += SCAN_BLK) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t x = i < gridDim.x ? *(volatile uint32_t *)(blk + i) : 0u;
        uint32_t t2, e2 = block_excl_scan_1024(x, sm, &t2);
        if (i < gridDim.x) blk[i] = (uint32_t)(carry + e2);
        carry += t2;
        // partial sums above 2^32 - 1 would wrap inside one 1024-block round only if a single round held > 2^32
        // minimizers, i.e. > 2^32 / 8192 per tile: impossible (a tile holds <= 4096 bases)
    }
    if (threadIdx.x == 0) {
        sc->n_minimizers = carry;
        uint32_t f = 0;
        if (carry > mini_cap) f |= BS_MINI_CAP;
        if (carry >= (1ull << 32) - 64) f |= BS_RANGE;
        if (*(volatile uint32_t *)&sc->ovf_count > ovf_cap) f |= BS_OVF_CAP;
        if (f) sc->flags = f;
    }
}

// ------------------------------------------------------------------------------------------------
// tiles: a warp-tile covers up to TW_MAX raw bases of one record, on a 16-base aligned grid
// ------------------------------------------------------------------------------------------------
constexpr int CS_MAX   = MQ_CS_MAX;         // raw bases per lane chunk (<= 128, multiple of 16)
constexpr int TW_MAX   = 32 * CS_MAX;       // raw bases per warp tile
// tiles of record [gs, ge) (host and device use the same formula)
__host__ __device__ inline uint32_t tiles_of_record(uint64_t gs, uint64_t ge, uint32_t min_len) {
    const uint64_t len = ge - gs;
    if (len < min_len || len == 0) return 0;
    return (uint32_t)((ge - (gs & ~15ull) + TW_MAX - 1) / TW_MAX);
}
__device__ __forceinline__ void tile_geometry(uint64_t gs, uint64_t ge, uint32_t nt, uint32_t ti, uint32_t *Cs, uint64_t *tlo) {
    const uint64_t A = gs & ~15ull;
    const uint32_t span = (uint32_t)(ge - A), d = 32u * nt;     // records are < 2^31 bases: 32-bit division is exact
    uint32_t c = (span + d - 1) / d;
    c = (c + 15u) & ~15u;
    *Cs = c; *tlo = A + (uint64_t)ti * 32u * c;
}

// ordered compaction: event (lane, j) of tile t -> rank = excl(lane) + n_lane - 1 - j
struct GatherArgs {
    const uint64_t *ev_hash; const uint32_t *ev_meta; const uint16_t *lane_cnt;
    TileBase tb;                     // exclusive prefix of the per-tile totals
    const uint32_t *tile_seq; const uint32_t *first_tile; const uint64_t *offs;
    const uint32_t *pos_base;        // per record: position of its first byte inside its reference (or NULL)
    uint32_t n_tiles;
    const BatchScalars *sc;
    uint32_t *out_pos; uint64_t *out_hash;
};
// A warp handles one tile (typically ~50 events), so the kernel lives on memory-level parallelism: its loads are
// issued in two dependency levels -- everything addressable from the tile id first, then the record geometry and the
// first 64 events together -- instead of one dependent load after another.
__global__ void __launch_bounds__(256) k_gather_minimizers(GatherArgs g) {
    const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= g.n_tiles || g.sc->flags) return;
    const uint32_t lane = lane_id();
    // level 1
    const uint32_t base = tile_base(g.tb, tile), total = tile_base(g.tb, tile + 1) - base;
    const uint32_t n = __ldg(g.lane_cnt + (uint64_t)tile * 32 + lane);
    const uint32_t sq = __ldg(g.tile_seq + tile);
    if (total == 0) return;                                // warp-uniform
    // level 2: geometry of the record and the first two rounds of events
    const uint32_t staged = total < EV_CAP ? total : EV_CAP;
    const uint64_t eb = (uint64_t)tile * EV_CAP;
    uint32_t meta0 = 0, meta1 = 0; uint64_t h0 = 0, h1 = 0;
    if (lane < staged) { meta0 = g.ev_meta[eb + lane]; h0 = g.ev_hash[eb + lane]; }
    if (lane + 32 < staged) { meta1 = g.ev_meta[eb + lane + 32]; h1 = g.ev_hash[eb + lane + 32]; }
    const uint64_t gs = __ldg(g.offs + sq), ge = __ldg(g.offs + sq + 1);
    const uint32_t ft = __ldg(g.first_tile + sq), nt = __ldg(g.first_tile + sq + 1) - ft;
    const int64_t pb = g.pos_base ? (int64_t)__ldg(g.pos_base + sq) : 0;
    uint32_t tot;
    const uint32_t ex = warp_excl_scan(n, &tot);
    uint32_t Cs; uint64_t tlo;
    tile_geometry(gs, ge, nt, tile - ft, &Cs, &tlo);
    const int64_t org = (int64_t)tlo - (int64_t)gs + pb;
    for (uint32_t s0 = 0; s0 < staged; s0 += 32) {
        const uint32_t s = s0 + lane;
        uint32_t meta; uint64_t h;
        if (s0 == 0) { meta = meta0; h = h0; }
        else if (s0 == 32) { meta = meta1; h = h1; }
        else { meta = 0; h = 0; if (s < staged) { meta = g.ev_meta[eb + s]; h = g.ev_hash[eb + s]; } }
        const uint32_t ln = (meta >> 14) & 31, j = meta >> 19;
        const uint32_t e = __shfl_sync(0xffffffffu, ex, ln), c = __shfl_sync(0xffffffffu, n, ln);
        if (s < staged) {
            const uint32_t rank = e + c - 1 - j;
            g.out_pos[base + rank] = (uint32_t)((int64_t)(meta & 0x3FFF) + org);
            g.out_hash[base + rank] = h;
        }
    }
}
// events that did not fit their tile's pool (density close to 1): grid-stride over the overflow pool
__global__ void __launch_bounds__(256) k_gather_overflow(GatherArgs g, const uint32_t *ovf_tile, const uint32_t *ovf_meta, const uint64_t *ovf_hash) {
    if (g.sc->flags) return;
    const uint32_t n_ovf = g.sc->ovf_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ovf; i += gridDim.x * blockDim.x) {
        const uint32_t tile = ovf_tile[i], meta = ovf_meta[i];
        const uint32_t ln = (meta >> 14) & 31, j = meta >> 19;
        uint32_t e = 0;
        for (uint32_t q = 0; q < ln; q++) e += g.lane_cnt[(uint64_t)tile * 32 + q];
        const uint32_t c = g.lane_cnt[(uint64_t)tile * 32 + ln];
        const uint32_t sq = g.tile_seq[tile];
        const uint64_t gs = g.offs[sq], ge = g.offs[sq + 1];
        const uint32_t ft = g.first_tile[sq], nt = g.first_tile[sq + 1] - ft;
        uint32_t Cs; uint64_t tlo;
        tile_geometry(gs, ge, nt, tile - ft, &Cs, &tlo);
        const int64_t org = (int64_t)tlo - (int64_t)gs + (g.pos_base ? (int64_t)g.pos_base[sq] : 0);
        const uint32_t rank = e + c - 1 - j, base = tile_base(g.tb, tile);
        g.out_pos[base + rank] = (uint32_t)((int64_t)(meta & 0x3FFF) + org);
        g.out_hash[base + rank] = ovf_hash[i];
    }
}
// seq_off[i] = prefix of record i's first tile (i <= n): the minimizer range of each record, materialised for the
// callers that want it as a plain array (index build, introspection); the mapping kernels read it through tile_base
__global__ void k_seq_mini_off(const uint32_t *first_tile, TileBase tb, uint32_t n, uint32_t *seq_off) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) seq_off[i] = tile_base(tb, first_tile[i]);
}

// ------------------------------------------------------------------------------------------------
// S2: k-min-mer of the window starting at minimizer j (hashes h[j..j+k))
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t kminmer_hash(const uint64_t *h, uint32_t k, uint32_t *rev_out) {
    uint32_t rev = 0;
    for (uint32_t i = 0; i < k; i++) {
        uint64_t a = h[i], b = h[k - 1 - i];
        if (a < b) break;
        if (a > b) { rev = 1; break; }
    }
    uint64_t x = 0x9E3779B97F4A7C15ull ^ (uint64_t)k;
    for (uint32_t i = 0; i < k; i++) x = mix64(x ^ (rev ? h[k - 1 - i] : h[i]));
    *rev_out = rev;
    return x;
}

// the same with k known at compile time: the window stays in registers (no local-memory array)
template <int K> __device__ __forceinline__ uint64_t kminmer_hash_k(const uint64_t *__restrict__ g, uint32_t *rev_out) {
    uint64_t h[K];
#pragma unroll
    for (int i = 0; i < K; i++) h[i] = __ldg(g + i);
    uint32_t rev = 0; bool decided = false;
#pragma unroll
    for (int i = 0; i < K / 2; i++) {
        const uint64_t a = h[i], b = h[K - 1 - i];
        if (!decided && a != b) { rev = a > b; decided = true; }
    }
    uint64_t x = 0x9E3779B97F4A7C15ull ^ (uint64_t)K;
#pragma unroll
    for (int i = 0; i < K; i++) x = mix64(x ^ (rev ? h[K - 1 - i] : h[i]));
    *rev_out = rev;
    return x;
}
// k-min-mer hash of the window starting at g (global memory); k <= 8 takes the register path
__device__ __forceinline__ uint64_t kminmer_hash_at(const uint64_t *__restrict__ g, uint32_t k, uint32_t *rev_out) {
    switch (k) {
        case 1: return kminmer_hash_k<1>(g, rev_out);
        case 2: return kminmer_hash_k<2>(g, rev_out);
        case 3: return kminmer_hash_k<3>(g, rev_out);
        case 4: return kminmer_hash_k<4>(g, rev_out);
        case 5: return kminmer_hash_k<5>(g, rev_out);
        case 6: return kminmer_hash_k<6>(g, rev_out);
        case 7: return kminmer_hash_k<7>(g, rev_out);
        case 8: return kminmer_hash_k<8>(g, rev_out);
        default: {
            uint64_t h[MQ_MAX_K_];
            for (uint32_t i = 0; i < k; i++) h[i] = __ldg(g + i);
            return kminmer_hash(h, k, rev_out);
        }
    }
}

// records (for the index) are runs [rec_off[r], rec_off[r+1]) of the minimizer store
__device__ __forceinline__ uint32_t find_rec(const uint32_t *rec_off, uint32_t n_rec, uint32_t j) {
    uint32_t lo = 0, hi = n_rec;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (rec_off[mid] <= j) lo = mid; else hi = mid; }
    return lo;
}

struct Table { Slot *slots; uint64_t mask; };   // capacity = mask+1 (+1 spare slot for key == EMPTY_KEY)

__device__ __forceinline__ void table_insert(const Table &t, uint64_t key, uint32_t id, uint32_t start, uint32_t end, uint32_t offrc) {
    uint64_t i;
    if (key == EMPTY_KEY) i = t.mask + 1;
    else {
        i = key & t.mask;
        for (;;) {
            unsigned long long *kp = (unsigned long long *)&t.slots[i].key;
            unsigned long long cur = *(volatile unsigned long long *)kp;
            if (cur == EMPTY_KEY) cur = atomicCAS(kp, (unsigned long long)EMPTY_KEY, (unsigned long long)key);
            if (cur == EMPTY_KEY || cur == key) break;
            i = (i + 1) & t.mask;
        }
    }
    // index.rs:100-104 -- the final state only depends on how many times the key was inserted:
    // exactly once => the entry, more => tombstone.  Whoever draws ticket 0 stores the entry.
    uint32_t c = atomicAdd(&t.slots[i].count, 1u);
    if (c == 0) { t.slots[i].id = id; t.slots[i].start = start; t.slots[i].end = end; t.slots[i].offrc = offrc; }
}

struct Entry { uint32_t id, start, end, offrc; };
// index.rs:118-126: present and not a tombstone
__device__ __forceinline__ bool table_get(const Table &t, uint64_t key, Entry *e) {
    uint64_t i = key == EMPTY_KEY ? t.mask + 1 : (key & t.mask);
    for (;;) {
        // the whole 32-byte slot (= one DRAM sector) in ONE request: sm_100 has 256-bit global loads
        uint64_t k, w1, w2, w3;
        asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(k), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(&t.slots[i]));
        if (k == key) {
            if ((uint32_t)w3 != 1u) return false;           // count != 1: tombstone (or untouched spare slot)
            e->id = (uint32_t)w1; e->start = (uint32_t)(w1 >> 32); e->end = (uint32_t)w2; e->offrc = (uint32_t)(w2 >> 32);
            return true;
        }
        if (k == EMPTY_KEY) return false;
        i = (i + 1) & t.mask;
    }
}

// ------------------------------------------------------------------------------------------------
// Presence filter in front of the table.  Four out of five query k-min-mers of a 99.5 %-identity read are NOT in the
// index (one error anywhere in the ~335-base span of a k-min-mer changes its key), and with a multi-GB table every such
// miss costs a random DRAM access (ncu, 3.1 Gbp index: 158 B of DRAM per probe, L2 hit rate 10 %).  A blocked Bloom filter
// over the keys with a VALID entry (count == 1; tombstones answer "absent" anyway) -- one 64-bit word per key, three
// bits inside it, ~6 bits per key: 32 MB for a human genome -- fits the L2 (126 MB) and is pinned there with an access
// policy window, so most misses are answered without touching DRAM.  No false negatives, so results cannot change.
// ------------------------------------------------------------------------------------------------
struct Bloom { const unsigned long long *words; uint32_t wmask; };   // words == NULL: no filter
__device__ __forceinline__ uint32_t bloom_word(uint64_t key, uint32_t wmask) { return (uint32_t)(key >> 32) & wmask; }
__device__ __forceinline__ unsigned long long bloom_bits(uint64_t key) {
    return (1ull << (key & 63)) | (1ull << ((key >> 6) & 63)) | (1ull << ((key >> 12) & 63));
}
__device__ __forceinline__ bool bloom_maybe(const Bloom &b, uint64_t key) {
    if (!b.words) return true;
    const unsigned long long m = bloom_bits(key);
    return (__ldg(b.words + bloom_word(key, b.wmask)) & m) == m;
}
__global__ void __launch_bounds__(256) k_bloom_build(const Slot *s, uint64_t n, unsigned long long *words, uint32_t wmask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint4 b = __ldg((const uint4 *)&s[i] + 1);           // end, offrc, count, pad
        if (b.z == 1u) { const uint64_t key = s[i].key; atomicOr(words + bloom_word(key, wmask), bloom_bits(key)); }
    }
}

// thread per minimizer j of the store: if a full window of k fits inside its record -> insert.
// tuple outputs (optional, for introspection / tests) are indexed km_off[rec] + (j - rec_off[rec]).
struct KminmerArgs {
    const uint32_t *pos; const uint64_t *hash; uint32_t n_min;
    const uint32_t *rec_off; const uint32_t *rec_id; uint32_t n_rec;   // rec_id may be NULL (id = rec index)
    const uint32_t *km_off;            // exclusive prefix of max(0,cnt-k+1) per record (tuple outputs only)
    uint32_t k, l;
    uint32_t *t_start, *t_end, *t_offrev; uint64_t *t_hash;            // optional
};
__global__ void __launch_bounds__(256) k_insert_kminmers(KminmerArgs a, Table t, int do_insert) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_min) return;
    uint32_t r = find_rec(a.rec_off, a.n_rec, j);
    uint32_t r0 = a.rec_off[r], r1 = a.rec_off[r + 1];
    if (j + a.k > r1) return;
    uint32_t rev; uint64_t key = kminmer_hash_at(a.hash + j, a.k, &rev);
    uint32_t start = a.pos[j], end = a.pos[j + a.k - 1] + a.l, off = j - r0;
    if (do_insert) table_insert(t, key, a.rec_id ? a.rec_id[r] : r, start, end, (off << 1) | rev);
    if (a.t_hash) {
        uint32_t o = a.km_off[r] + off;
        a.t_start[o] = start; a.t_end[o] = end; a.t_offrev[o] = (off << 1) | rev; a.t_hash[o] = key;
    }
}

__global__ void __launch_bounds__(256) k_table_clear(Slot *s, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint4 a, b; a.x = 0xFFFFFFFFu; a.y = 0xFFFFFFFFu; a.z = 0; a.w = 0; b.x = b.y = b.z = b.w = 0;
        ((uint4 *)&s[i])[0] = a; ((uint4 *)&s[i])[1] = b;
    }
}
// get_count (index.rs:90-92): entries with count == 1; n_keys: slots with count >= 1
__global__ void __launch_bounds__(256) k_table_count(const Slot *s, uint64_t n, unsigned long long *out2) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t u = 0, kk = 0;
    for (; i < n; i += stride) { uint32_t c = s[i].count; u += (c == 1); kk += (c >= 1); }
    for (int d = 16; d; d >>= 1) { u += __shfl_down_sync(0xffffffffu, u, d); kk += __shfl_down_sync(0xffffffffu, kk, d); }
    if (lane_id() == 0) { if (u) atomicAdd(&out2[0], (unsigned long long)u); if (kk) atomicAdd(&out2[1], (unsigned long long)kk); }
}
__global__ void k_index_get(Table t, const uint64_t *keys, uint64_t n, uint8_t *found, uint32_t *id, uint32_t *start,
                            uint32_t *end, uint32_t *offset, uint8_t *rc) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Entry e; bool ok = table_get(t, keys[i], &e);
    found[i] = ok;
    if (ok) { id[i] = e.id; start[i] = e.start; end[i] = e.end; offset[i] = e.offrc >> 1; rc[i] = e.offrc & 1; }
}

// ------------------------------------------------------------------------------------------------
// per-read probe + Match segmentation (warp per read)
// ------------------------------------------------------------------------------------------------
struct ProbeArgs {
    const uint32_t *pos; const uint64_t *hash;     // minimizers of the batch
    const uint32_t *first_tile; TileBase tb;       // minimizer range of read r: tile_base(first_tile[r]) .. tile_base(first_tile[r+1])
    const BatchScalars *sc;
    uint32_t n_reads, k, l;
    MatchRec *matches;                             // read r writes at (its first minimizer) + ordinal
    uint32_t *n_matches;                           // per read
    uint32_t *read_ticket;
};

__global__ void __launch_bounds__(128) k_probe_match(ProbeArgs a, Table t, Bloom bf) {
    const uint32_t lane = lane_id();
    if (a.sc->flags) return;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(a.read_ticket, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.n_reads) break;
        const uint32_t m0 = tile_base(a.tb, __ldg(a.first_tile + r)), m1 = tile_base(a.tb, __ldg(a.first_tile + r + 1));
        const uint32_t M = m1 - m0, Q = M >= a.k ? M - a.k + 1 : 0;
        MatchRec *out = a.matches + m0;
        uint32_t n_heads = 0;                       // matches opened so far (warp-uniform)
        uint32_t s_in = 0;                          // automaton state after the previous block
        // carry of the last item of the previous block (only meaningful when it was a hit)
        uint32_t c_hit = 0, c_id = 0, c_off = 0, c_qend = 0, c_rstart = 0, c_rend = 0, c_j = 0;
        uint32_t c_mrc = 0;                         // rc of the Match open at the end of the previous block
        for (uint32_t j0 = 0; j0 < Q; j0 += 32) {
            const uint32_t j = j0 + lane;
            bool hit = false; Entry e; e.id = e.start = e.end = e.offrc = 0;
            uint32_t qstart = 0, qend = 0, qrev = 0;
            // the next block's minimizers: in flight while this block waits for the filter and the table
            if (j + 32 < M) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.hash + m0 + j + 32));
                if ((lane & 1u) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pos + m0 + j + 32));
            }
            if (j < Q) {
                uint64_t key = kminmer_hash_at(a.hash + m0 + j, a.k, &qrev);
                qstart = __ldg(a.pos + m0 + j); qend = __ldg(a.pos + m0 + j + a.k - 1) + a.l;
                hit = bloom_maybe(bf, key) && table_get(t, key, &e);
            }
            const uint32_t off = e.offrc >> 1, rc = hit ? (qrev ^ (e.offrc & 1)) : 0;   // Match::new rc = q.rev != r.rc
            // previous item (lane-1, or the carry for lane 0)
            uint32_t p_hit = __shfl_up_sync(0xffffffffu, (uint32_t)hit, 1), p_id = __shfl_up_sync(0xffffffffu, e.id, 1),
                     p_off = __shfl_up_sync(0xffffffffu, off, 1);
            if (lane == 0) { p_hit = c_hit; p_id = c_id; p_off = c_off; }
            // match.rs:39-43 with Rust precedence: rc Match: same ref && q.rev!=r.rc && p.off-r.off==1 ; fwd Match: r.off-p.off==1
            const bool fl = hit && p_hit && (off - p_off == 1u);
            const bool rl = hit && p_hit && (e.id == p_id) && rc && (p_off - off == 1u);
            // The automaton of match.rs:45-58 has three states {closed, open forward Match, open rc Match}.  Because an rc
            // Match only extends on rc hits (rl implies rc), the state after a hit is "forward" iff the hit is forward, or
            // it is an rc hit that extends a forward Match (offset + 1: the precedence quirk).  That is a generate /
            // propagate carry chain over the warp's ballots: G = forward hits, P = rc hits with offset + 1.
            const uint32_t HIT = __ballot_sync(0xffffffffu, hit), RCm = __ballot_sync(0xffffffffu, rc != 0),
                           FL = __ballot_sync(0xffffffffu, fl), RL = __ballot_sync(0xffffffffu, rl);
            const uint32_t G = HIT & ~RCm, P = RCm & FL;
            const uint32_t X = ((G << 1) | (s_in == 1u ? 1u : 0u)) & P;        // P-runs whose predecessor is in the forward state
            const uint32_t S1 = G | (P & ~(P + X));                            // the carry ripples through each such run
            const uint32_t S2 = HIT & ~S1;
            const uint32_t Em = ((((S1 << 1) | (s_in == 1u ? 1u : 0u)) & FL) | (((S2 << 1) | (s_in == 2u ? 1u : 0u)) & RL)) & HIT;
            const uint32_t Hm = HIT & ~Em;
            const bool head = (Hm >> lane) & 1u;
            const uint32_t s_after31 = (S1 >> 31) ? 1u : ((S2 >> 31) ? 2u : 0u);
            // close the Match carried in from the previous block if lane 0 does not extend it
            if (lane == 0 && c_hit && !(Em & 1u)) {
                MatchRec *mr = out + (n_heads - 1);
                mr->q_end = c_qend; mr->last_j = c_j;
                if (c_mrc) mr->r_start = c_rstart; else mr->r_end = c_rend;
            }
            const uint32_t below = Hm & ((2u << lane) - 1u);           // heads at or below me
            const uint32_t ord = n_heads + __popc(below) - 1;          // ordinal of the Match I belong to
            if (head) {
                MatchRec *mr = out + ord;
                mr->q_start = qstart; mr->head_j = j; mr->ref_rc = (e.id << 1) | rc; mr->pad_ = 0;
                if (rc) mr->r_end = e.end; else mr->r_start = e.start;
            }
            // I am the last item of my Match if the next item does not extend; lane 31 defers to the carry
            const bool nxt_ext = (Em >> 1 >> lane) & 1u;
            if (hit && lane < 31 && !nxt_ext) {
                const uint32_t mrc = below ? ((RCm >> (31 - __clz(below))) & 1u) : c_mrc;
                MatchRec *mr = out + ord;
                mr->q_end = qend; mr->last_j = j;
                if (mrc) mr->r_start = e.start; else mr->r_end = e.end;
            }
            // carry
            {
                const uint32_t below31 = Hm;   // heads at or below lane 31
                const uint32_t mrc31 = below31 ? ((RCm >> (31 - __clz(below31))) & 1u) : c_mrc;
                c_hit = __shfl_sync(0xffffffffu, (uint32_t)hit, 31); c_id = __shfl_sync(0xffffffffu, e.id, 31);
                c_off = __shfl_sync(0xffffffffu, off, 31); c_qend = __shfl_sync(0xffffffffu, qend, 31);
                c_rstart = __shfl_sync(0xffffffffu, e.start, 31); c_rend = __shfl_sync(0xffffffffu, e.end, 31);
                c_j = j0 + 31; c_mrc = mrc31;
                s_in = s_after31;
                n_heads += __popc(Hm);
            }
        }
        // flush the Match still open at the end of the read (last block's lane 31 was a hit)
        if (lane == 0 && c_hit && Q > 0 && (Q & 31u) == 0) {
            MatchRec *mr = out + (n_heads - 1);
            mr->q_end = c_qend; mr->last_j = c_j;
            if (c_mrc) mr->r_start = c_rstart; else mr->r_end = c_rend;
        }
        if (lane == 0) a.n_matches[r] = n_heads;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// chain + MAPQ + best reference + find_coords (warp per read)
// ------------------------------------------------------------------------------------------------
struct M6 { uint32_t qs, qe, rs, re, cnt, rc, ref; };
__device__ __forceinline__ M6 load_match(const MatchRec *p) {
    const uint4 *q = (const uint4 *)p; uint4 a = q[0], b = q[1];
    M6 m; m.qs = a.x; m.qe = a.y; m.rs = a.z; m.re = a.w; m.cnt = b.y - b.x + 1; m.rc = b.z & 1; m.ref = b.z >> 1;
    return m;
}
// chain.rs:132-142 (operands cast `as i32` before subtracting; everything wraps)
__device__ __forceinline__ bool gap_too_long(uint32_t a1, uint32_t a0, uint32_t b1, uint32_t b0, uint32_t g) {
    uint32_t g1 = a1 - a0, g2 = b1 - b0;
    int32_t d = (int32_t)(g1 - g2);
    int32_t ad = d < 0 ? (int32_t)(0u - (uint32_t)d) : d;
    return (uint64_t)(int64_t)ad > (uint64_t)g;
}
// chain.rs:43-63
__device__ __forceinline__ bool compatible(const M6 &h1, const M6 &h2, uint32_t g) {
    if (h1.qs == h2.qs && h1.qe == h2.qe && h1.rs == h2.rs && h1.re == h2.re && h1.cnt == h2.cnt && h1.rc == h2.rc) return true;
    if (h1.rc != h2.rc) return false;
    const bool first = h1.qs < h2.qs;
    const M6 &u = first ? h1 : h2; const M6 &v = first ? h2 : h1;
    if (u.rc) { if (u.rs <= v.rs || gap_too_long(v.qs, u.qe, u.rs, v.re, g)) return false; }
    else if (v.rs <= u.rs || gap_too_long(v.qs, u.qe, v.rs, u.re, g)) return false;
    return true;
}

struct ChainArgs {
    const MatchRec *matches; const uint32_t *n_matches; const uint32_t *first_tile; TileBase tb; const BatchScalars *sc;
    const uint64_t *offs;
    const uint64_t *ref_lens; uint32_t n_refs;
    uint32_t n_reads, c, s, g;
    HitRec *hits;
    uint32_t *read_ticket;
    uint32_t *big_list;        // reads with more than CHAIN_SMALL matches, filled by k_chain_small, drained by k_chain
    uint32_t *big_count;
};
constexpr uint32_t CHAIN_SMALL = 12;

// find_coords (mers.rs:131-179, wrapping usize arithmetic) + the mq_hit record
__device__ __forceinline__ void write_hit(const ChainArgs &a, uint32_t r, bool ok, uint32_t b_ref, uint32_t b_rc, uint32_t b_mapq,
                                          uint64_t b_qs, uint64_t b_qe, uint64_t b_rs, uint64_t b_re, uint64_t best_score) {
    HitRec h; h.mapped = 0; h.rc = 0; h.mapq = 0; h.pad_ = 0; h.ref_idx = 0; h.q_start = h.q_end = h.r_start = h.r_end = h.score = 0;
    if (ok && b_ref < a.n_refs) {
        const uint64_t q_len = a.offs[r + 1] - a.offs[r], r_len = a.ref_lens[b_ref];
        const uint64_t tail = q_len - b_qe - 1;
        uint64_t frs, fre, exs, exe;
        if (!b_rc) {
            if (b_rs >= b_qs) { frs = b_rs - b_qs; exs = b_qs; } else { frs = 0; exs = b_rs; }
            if (b_re + tail <= r_len - 1) { fre = b_re + tail; exe = tail; } else { fre = r_len - 1; exe = r_len - b_re - 1; }
        } else {
            if (b_re + b_qs <= r_len - 1) { fre = b_re + b_qs; exs = b_qs; } else { fre = r_len - 1; exs = r_len - b_re - 1; }
            if (b_rs >= tail) { frs = b_rs - tail; exe = tail; } else { frs = 0; exe = b_rs; }
        }
        h.mapped = 1; h.rc = (uint8_t)b_rc; h.mapq = (uint8_t)b_mapq; h.ref_idx = b_ref;
        h.q_start = b_qs - exs; h.q_end = b_qe + exe; h.r_start = frs; h.r_end = fre; h.score = best_score;
    }
    a.hits[r] = h;
}

// thread per read: almost every read has a handful of Matches, for which a warp per read wastes 31 lanes.
// Reads with more than CHAIN_SMALL Matches are queued for the warp-per-read kernel below.
// The body is instantiated for a compile-time bound NMAX on the number of Matches: for 1..4 (nearly every read) all loops
// unroll and the Matches stay in registers; a runtime-indexed array would live in local memory, and its dependent
// loads were what this kernel spent its time on (long-scoreboard stalls, 10 % issue utilisation).
template <int NMAX>
__device__ __forceinline__ void chain_thread(const ChainArgs &a, uint32_t r, uint32_t n, const MatchRec *ms) {
    M6 m[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; i++) if (i < (int)n) m[i] = load_match(ms + i); else { m[i] = M6{}; m[i].ref = 0xFFFFFFFFu; }
    uint64_t best_score = 0, second = 0; uint32_t groups = 0;
    uint32_t b_ref = 0, b_rc = 0, b_mapq = 0; uint64_t b_qs = 0, b_qe = 0, b_rs = 0, b_re = 0;
#pragma unroll
    for (int i = 0; i < NMAX; i++) {
        if (i >= (int)n) break;
        const uint32_t ref = m[i].ref;
        bool seen = false;
#pragma unroll
        for (int q = 0; q < i; q++) seen |= m[q].ref == ref;
        if (seen) continue;                                             // not the first Match of its reference
        uint32_t bc = 0, glen = 0; M6 mb = m[i];                        // C1: first Match with the strictly greatest count
#pragma unroll
        for (int q = i; q < NMAX; q++) if (q < (int)n && m[q].ref == ref) { glen++; if (m[q].cnt > bc) { bc = m[q].cnt; mb = m[q]; } }
        uint32_t lenf = 0; uint64_t score = 0; M6 mf = m[i], ml = m[i]; // C3: keep what is compatible with the largest
#pragma unroll
        for (int q = i; q < NMAX; q++)
            if (q < (int)n && m[q].ref == ref && (glen <= 1 || compatible(mb, m[q], a.g))) { if (!lenf) mf = m[q]; ml = m[q]; lenf++; score += m[q].cnt; }
        if (!lenf) continue;
        const uint32_t mapq = ((a.s != 0 && a.c != 0) && (lenf >= a.c || score >= a.s)) ? 60u : 0u;          // C4
        const uint32_t rc = mf.rc;
        const uint64_t qs = mf.qs, qe = (uint64_t)ml.qe - 1;
        uint64_t rs, re;
        if (rc && lenf > 1) { rs = ml.rs; re = (uint64_t)mf.re - 1; } else { rs = mf.rs; re = (uint64_t)ml.re - 1; }
        groups++;
        if (score > best_score) { second = best_score; best_score = score; b_ref = ref; b_rc = rc; b_mapq = mapq; b_qs = qs; b_qe = qe; b_rs = rs; b_re = re; }
        else if (score > second) second = score;
    }
    write_hit(a, r, groups == 1 || (groups > 1 && best_score != second), b_ref, b_rc, b_mapq, b_qs, b_qe, b_rs, b_re, best_score);
}

__global__ void __launch_bounds__(128) k_chain_small(ChainArgs a) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_reads || a.sc->flags) return;
    const uint32_t n = a.n_matches[r];
    if (n > CHAIN_SMALL) { a.big_list[atomicAdd(a.big_count, 1u)] = r; return; }
    const MatchRec *ms = a.matches + tile_base(a.tb, __ldg(a.first_tile + r));
    if (n <= 1) chain_thread<1>(a, r, n, ms);
    else if (n == 2) chain_thread<2>(a, r, n, ms);
    else if (n <= 4) chain_thread<4>(a, r, n, ms);
    else chain_thread<CHAIN_SMALL>(a, r, n, ms);
}


__global__ void __launch_bounds__(128) k_chain(ChainArgs a) {
    const uint32_t lane = lane_id();
    if (a.sc->flags) return;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(a.read_ticket, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (a.big_list) { if (r >= *a.big_count) break; r = a.big_list[r]; }      // only the reads k_chain_small queued
        else if (r >= a.n_reads) break;
        const uint32_t n = a.n_matches[r];
        const MatchRec *ms = a.matches + tile_base(a.tb, __ldg(a.first_tile + r));
        // best / second-best chain score over references (mers.rs:110-129)
        uint64_t best_score = 0, second = 0; uint32_t groups = 0;
        uint32_t b_ref = 0, b_rc = 0, b_mapq = 0; uint64_t b_qs = 0, b_qe = 0, b_rs = 0, b_re = 0;
        // leader = first Match of its reference.  Up to 32 Matches (nearly every read that gets here): one MATCH.ANY finds
        // them all and the loop visits leaders only; longer lists test every Match against the ones before it.
        const bool few = n <= 32;
        uint32_t leaders = 0;
        if (few) {
            const uint32_t myref = lane < n ? (ms[lane].ref_rc >> 1) : 0xFFFFFFFFu;      // reference ids are < 2^31
            const uint32_t same = __match_any_sync(0xffffffffu, myref);
            leaders = __ballot_sync(0xffffffffu, lane < n && (uint32_t)(__ffs(same) - 1) == lane);
        }
        for (uint32_t i = 0; i < n; i++) {
            if (few) {
                if (!leaders) break;
                i = (uint32_t)__ffs(leaders) - 1; leaders &= leaders - 1;
            }
            const uint32_t ref = ms[i].ref_rc >> 1;
            if (!few) {
                bool seen = false;
                for (uint32_t q = lane; q < i; q += 32) seen |= (ms[q].ref_rc >> 1) == ref;
                if (__any_sync(0xffffffffu, seen)) continue;
            }
            // C1: first match with the strictly greatest count (chain.rs:93-104)
            uint32_t bc = 0, bi = 0xFFFFFFFFu, glen = 0;
            for (uint32_t q = i + lane; q < n; q += 32) {
                const MatchRec mr = ms[q];
                if ((mr.ref_rc >> 1) == ref) { glen++; uint32_t c = mr.last_j - mr.head_j + 1; if (c > bc) { bc = c; bi = q; } }
            }
            for (int d = 16; d; d >>= 1) {
                uint32_t oc = __shfl_xor_sync(0xffffffffu, bc, d), oi = __shfl_xor_sync(0xffffffffu, bi, d);
                glen += __shfl_xor_sync(0xffffffffu, glen, d);
                if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
            }
            const M6 big = load_match(ms + bi);
            // C3: keep the matches compatible with the largest (only when the group has > 1)
            uint32_t lenf = 0, fi = 0xFFFFFFFFu, li = 0; uint64_t score = 0;
            for (uint32_t q = i + lane; q < n; q += 32) {
                const M6 m = load_match(ms + q);
                if (m.ref == ref && (glen <= 1 || compatible(big, m, a.g))) { lenf++; score += m.cnt; fi = min(fi, q); li = max(li, q); }
            }
            for (int d = 16; d; d >>= 1) {
                lenf += __shfl_xor_sync(0xffffffffu, lenf, d); score += __shfl_xor_sync(0xffffffffu, score, d);
                fi = min(fi, __shfl_xor_sync(0xffffffffu, fi, d)); li = max(li, __shfl_xor_sync(0xffffffffu, li, d));
            }
            if (lenf == 0) continue;                                       // cannot happen (largest is self-compatible)
            // C4: get_match (chain.rs:155-168)
            const M6 first = load_match(ms + fi), last = load_match(ms + li);
            const uint32_t mapq = ((a.s != 0 && a.c != 0) && (lenf >= a.c || score >= a.s)) ? 60u : 0u;
            const uint32_t rc = first.rc;
            uint64_t qs = first.qs, qe = (uint64_t)last.qe - 1, rs, re;
            if (rc && lenf > 1) { rs = last.rs; re = (uint64_t)first.re - 1; } else { rs = first.rs; re = (uint64_t)last.re - 1; }
            groups++;
            if (score > best_score) {
                second = best_score; best_score = score;
                b_ref = ref; b_rc = rc; b_mapq = mapq; b_qs = qs; b_qe = qe; b_rs = rs; b_re = re;
            } else if (score > second) second = score;
        }
        if (lane == 0)      // tie for the max => unmapped (mers.rs:106)
            write_hit(a, r, groups == 1 || (groups > 1 && best_score != second), b_ref, b_rc, b_mapq, b_qs, b_qe, b_rs, b_re, best_score);
        __syncwarp();
    }
}

}  // namespace mq
