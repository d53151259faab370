// mq_kernels.cuh -- sm_100a kernels of the mapquik seeding->chaining hot path.
//
// Stage map (reference rows of SURVEY.md section 8a):
//   k_scan_minimizers   S1  HPC + ntHash-1 canonical l-mer hash + universe-minimizer sampling
//                           (the KminmersIterator stage-1 the reference calls at mers.rs:27,53)
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
constexpr int      C_MAX        = 256;             // raw bases per lane chunk (max)
constexpr int      TW_MAX       = 32 * C_MAX;      // raw bases per warp tile (max)
constexpr int      LANE_PAD     = 4;               // smem bytes of padding between lane chunks
constexpr int      HALO_MAX     = 32;              // >= l-1 compressed symbols right of the tile
constexpr int      SCAN_WARPS   = 4;               // warps (= tiles in flight) per CTA
constexpr int      TILE_SMEM    = 33 * (C_MAX + LANE_PAD);   // 32 chunks + halo "chunk"
constexpr uint32_t EV_CAP       = 256;             // staged events per tile before overflow pool
constexpr uint32_t D_RUN = 8, D_N = 4;             // digest byte: bits0-1 code, bit2 non-ACGT, bit3 run start
constexpr uint64_t EMPTY_KEY    = 0xFFFFFFFFFFFFFFFFull;
constexpr int      MQ_MAX_K_    = 32;

constexpr uint64_t SEED_A = 0x3c8bfbb395c60474ull, SEED_C = 0x3193c18562a02b4cull,
                   SEED_G = 0x20323ed082572324ull, SEED_T = 0x295549f54be24456ull;

struct ScanTables {            // per-launch constants derived from l (host fills)
    uint64_t pairF[16];        // [in | out<<2] : rol(h(in), l-1) ^ ror(h(out), 1)
    uint64_t pairR[16];        // [in | out<<2] : hc(in) ^ rol(hc(out), l)
    uint64_t inF[4], outF[4], inR[4], outR[4];   // the single-symbol parts (N-aware path)
    uint64_t h[4], hc[4];      // base seeds by code (A,C,G,T) and of the complement
};

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
// generic exclusive scan (u32 -> u32), 3 kernels: block sums, scan of sums, apply.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_BLK = 1024, SCAN_ITEMS = 8;     // 8192 items per block

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

__global__ void __launch_bounds__(SCAN_BLK) k_scan_block_sums(const uint32_t *in, uint64_t n, uint32_t *block_sums) {
    __shared__ uint32_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLK * SCAN_ITEMS + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) s += in[base + i];
    uint32_t tot; block_excl_scan_1024(s, sm, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
// single block: exclusive scan of block_sums in place; total -> *total_out (u64)
__global__ void __launch_bounds__(SCAN_BLK) k_scan_sums(uint32_t *block_sums, uint32_t nb, uint64_t *total_out) {
    __shared__ uint32_t sm[33];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nb; base += SCAN_BLK) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nb ? block_sums[i] : 0, tot;
        uint32_t e = block_excl_scan_1024(v, sm, &tot);
        if (i < nb) block_sums[i] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0) *total_out = carry;
}
// out[i] = exclusive prefix; out may alias in. out has n+1 entries when write_total (out[n] = total)
__global__ void __launch_bounds__(SCAN_BLK) k_scan_apply(const uint32_t *in, uint64_t n, const uint32_t *block_sums,
                                                         uint32_t *out, int write_total) {
    __shared__ uint32_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLK * SCAN_ITEMS + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    uint32_t tot, e = block_excl_scan_1024(s, sm, &tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { if (base + i < n) out[base + i] = e; e += v[i]; }
    if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_BLK - 1) out[n] = block_sums[blockIdx.x] + tot;
}

// ------------------------------------------------------------------------------------------------
// tile setup
// ------------------------------------------------------------------------------------------------
// tiles of record i: span measured from the 4-byte-aligned address at or below its first byte.
__global__ void k_tiles_per_seq(const uint64_t *offs, uint32_t n, uint32_t min_len, uint32_t *tiles) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t gs = offs[i], ge = offs[i + 1], len = ge - gs;
    uint32_t t = 0;
    if (len >= min_len && len > 0) { uint64_t span = ge - (gs & ~3ull); t = (uint32_t)((span + TW_MAX - 1) / TW_MAX); }
    tiles[i] = t;
}
// tile -> record (binary search over first_tile[n+1])
__global__ void k_tile_seq(const uint32_t *first_tile, uint32_t n, uint32_t n_tiles, uint32_t *tile_seq) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    uint32_t lo = 0, hi = n;           // largest i with first_tile[i] <= t and first_tile[i+1] > t
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (first_tile[mid] <= t) lo = mid; else hi = mid; }
    tile_seq[t] = lo;
}

// ------------------------------------------------------------------------------------------------
// S1: the sequence scan
// ------------------------------------------------------------------------------------------------
// digest of one 4-byte word: per byte  code | non-ACGT<<2 | run-start<<3.
// pv = the same word shifted up by one byte with the preceding byte shifted in (for run starts).
__device__ __forceinline__ uint32_t digest_word(uint32_t u, uint32_t pv, bool use_hpc) {
    uint32_t t = (u >> 1) & 0x03030303u;
    uint32_t code = t ^ ((t >> 1) & 0x01010101u);                  // A0 C1 G2 T3
    // validity: rebuild the expected ASCII from the code with a byte-permute LUT and compare
    uint32_t sel = (code & 0x3u) | ((code >> 4) & 0x30u) | ((code >> 8) & 0x300u) | ((code >> 12) & 0x3000u);
    uint32_t expect = __byte_perm(0x54474341u /* 'A','C','G','T' */, 0u, sel);
    uint32_t diff = expect ^ u;
    uint32_t bad = (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;   // 0x80 where byte != expected
    uint32_t run = 0x80808080u;
    if (use_hpc) { uint32_t e = u ^ pv; run = (((e & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | e) & 0x80808080u; }
    return code | (bad >> 5) | (run >> 4);
}

struct ScanArgs {
    const uint8_t  *seqs;          // concatenated records (device), base 4-byte aligned, padded
    const uint64_t *offs;          // n+1
    const uint32_t *first_tile;    // n+1
    const uint32_t *tile_seq;      // n_tiles
    uint32_t n_tiles;
    uint32_t l;
    uint32_t use_hpc;
    uint64_t bound;
    // outputs
    uint64_t *ev_hash;             // n_tiles * EV_CAP
    uint32_t *ev_meta;             // n_tiles * EV_CAP   x' (14b) | lane<<14 (5b) | j<<19 (13b)
    uint16_t *lane_cnt;            // n_tiles * 32
    uint32_t *tile_cnt;            // n_tiles (total events of the tile, incl. overflowed)
    // overflow pool
    uint32_t *ovf_count;           // single counter
    uint32_t  ovf_cap;
    uint32_t *ovf_tile; uint32_t *ovf_meta; uint64_t *ovf_hash;
    uint32_t *tile_ticket;         // dynamic tile scheduler
    const uint32_t *emit_range;    // per record (or NULL): [lo, hi) record offsets; only l-mers STARTING inside are emitted
};
// emission window of a tile in x' coordinates (segment scans; the whole tile otherwise)
__device__ __forceinline__ void emit_window(const ScanArgs &a, uint32_t sq, uint64_t gs, uint64_t tlo, uint32_t *xlo, uint32_t *xhi) {
    *xlo = 0u; *xhi = 0x7FFFFFFFu;
    if (a.emit_range) {
        const int64_t sh = (int64_t)gs - (int64_t)tlo;
        const int64_t lo = (int64_t)a.emit_range[2 * sq] + sh, hi = (int64_t)a.emit_range[2 * sq + 1] + sh;
        *xlo = lo <= 0 ? 0u : (lo > 0x7FFFFFFF ? 0x7FFFFFFFu : (uint32_t)lo);
        *xhi = hi <= 0 ? 0u : (hi > 0x7FFFFFFF ? 0x7FFFFFFFu : (uint32_t)hi);
        if (*xhi < *xlo) *xhi = *xlo;
    }
}

struct LaneState { uint64_t F, R, W; uint32_t WN; };

// N-aware symbol step (prepend `d` on the left, drop the right-most symbol)
__device__ __forceinline__ void step_generic(LaneState &s, uint32_t d, const ScanTables &T, uint32_t l) {
    uint32_t in = d & 3, out = (uint32_t)s.W & 3;
    bool inN = d & D_N, outN = s.WN & 1;
    uint64_t tf = (inN ? 0 : T.inF[in]) ^ (outN ? 0 : T.outF[out]);
    uint64_t tr = (inN ? 0 : T.inR[in]) ^ (outN ? 0 : T.outR[out]);
    s.F = ror1(s.F) ^ tf; s.R = rol1(s.R) ^ tr;
    s.W = (s.W >> 2) | ((uint64_t)in << (2 * l - 2));
    s.WN = (s.WN >> 1) | ((inN ? 1u : 0u) << (l - 1));
}

template <bool HAS_N>
__device__ __forceinline__ void emit_event(uint32_t x, uint64_t h, uint32_t lane, uint32_t &nloc, uint32_t *tile_ev_smem,
                                           uint32_t tile, const ScanArgs &a) {
    uint32_t slot = atomicAdd(tile_ev_smem, 1u);
    uint32_t meta = x | (lane << 14) | (nloc << 19);
    nloc++;
    if (slot < EV_CAP) {
        a.ev_hash[(uint64_t)tile * EV_CAP + slot] = h;
        a.ev_meta[(uint64_t)tile * EV_CAP + slot] = meta;
    } else {
        uint32_t g = atomicAdd(a.ovf_count, 1u);
        if (g < a.ovf_cap) { a.ovf_tile[g] = tile; a.ovf_meta[g] = meta; a.ovf_hash[g] = h; }
    }
}

__global__ void __launch_bounds__(SCAN_WARPS * 32) k_scan_minimizers(ScanArgs a, ScanTables Tin) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ ScanTables T;
    __shared__ uint32_t ev_cnt[SCAN_WARPS];
    for (uint32_t i = threadIdx.x; i < sizeof(ScanTables) / 8; i += blockDim.x) ((uint64_t *)&T)[i] = ((const uint64_t *)&Tin)[i];
    __syncthreads();
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    uint8_t *S = smem_raw + (size_t)wid * TILE_SMEM;
    const uint32_t l = a.l;
    const bool hpc = a.use_hpc != 0;
    const uint32_t bound_hi = (uint32_t)(a.bound >> 32);

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) break;
        if (lane == 0) ev_cnt[wid] = 0;

        // ---- geometry -------------------------------------------------------------------------
        const uint32_t sq = a.tile_seq[tile];
        const uint64_t gs = a.offs[sq], ge = a.offs[sq + 1];
        const uint64_t A = gs & ~3ull;
        const uint32_t ft = a.first_tile[sq], nt = a.first_tile[sq + 1] - ft, ti = tile - ft;
        const uint32_t span = (uint32_t)(ge - A), dv = 32u * nt;
        uint32_t Cs = (span + dv - 1) / dv;
        Cs = (Cs + 7u) & ~7u;                                   // multiple of 8 => odd word stride with the pad
        const uint32_t TWs = 32u * Cs, stride = Cs + LANE_PAD;
        const uint64_t tlo = A + (uint64_t)ti * TWs;            // aligned address of x' = 0
        uint32_t nloc = 0;
        if (tlo >= ge) {                                        // empty tail tile
            a.lane_cnt[(uint64_t)tile * 32 + lane] = 0;
            if (lane == 0) a.tile_cnt[tile] = 0;
            continue;
        }
        const uint32_t own_lo = gs > tlo ? (uint32_t)(gs - tlo) : 0u;
        const uint32_t own_hi = (ge - tlo) < TWs ? (uint32_t)(ge - tlo) : TWs;      // exclusive
        const uint32_t magic = 0xFFFFFFFFu / Cs + 1u;          // x / Cs == umulhi(x, magic) for x < 2^16
        uint32_t xlo, xlim; emit_window(a, sq, gs, tlo, &xlo, &xlim);   // segment scans only

        // ---- stage + digest the tile ------------------------------------------------------------
        const uint32_t *gw = (const uint32_t *)(a.seqs + tlo);
        uint32_t nwords = (own_hi + 3) >> 2;
        uint32_t carry = 0;                                     // last byte of the previous word row
        if (tlo > 0 && lane == 0) carry = a.seqs[tlo - 1];
        carry = __shfl_sync(0xffffffffu, carry, 0);
        uint32_t anyN = 0;
        for (uint32_t w0 = 0; w0 < nwords; w0 += 32) {
            uint32_t w = w0 + lane;
            uint32_t u = (w < nwords) ? __ldg(gw + w) : 0u;
            uint32_t up = __shfl_up_sync(0xffffffffu, u, 1);
            uint32_t prevb = lane == 0 ? carry : (up >> 24);
            carry = __shfl_sync(0xffffffffu, u, 31) >> 24;
            uint32_t dg = digest_word(u, (u << 8) | prevb, hpc);
            uint32_t x = w << 2;
            if (x < own_lo || x + 4 > own_hi) {                 // partial word: blank bytes outside the record
                uint32_t m = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) if (x + b >= own_lo && x + b < own_hi) m |= 0xFFu << (8 * b);
                dg &= m;
            }
            if (x <= own_lo && own_lo < x + 4 && tlo + own_lo == gs) dg |= D_RUN << (8 * (own_lo - x));  // record start
            if (w < nwords) {
                anyN |= dg & 0x04040404u;
                uint32_t ch = __umulhi(x, magic);
                *(uint32_t *)(S + x + ch * LANE_PAD) = dg;
            }
        }
        // blank the rest of the tile window (words beyond the record end) so stale bytes are inert
        for (uint32_t w = nwords + lane; w < (TWs >> 2); w += 32) {
            uint32_t x = w << 2; uint32_t ch = __umulhi(x, magic);
            *(uint32_t *)(S + x + ch * LANE_PAD) = 0u;
        }

        // ---- halo: up to l-1 further run starts right of the tile -------------------------------
        uint32_t hcount = 0;
        uint8_t *H = S + 32u * stride;
        if (tlo + TWs < ge) {
            uint64_t haddr = tlo + TWs;                         // 4-aligned
            uint32_t hcarry = __shfl_sync(0xffffffffu, carry, 0);
            while (hcount < l - 1 && haddr < ge) {
                uint64_t wa = haddr + 4ull * lane;
                uint32_t u = (wa < ge) ? __ldg((const uint32_t *)(a.seqs + wa)) : 0u;
                uint32_t up = __shfl_up_sync(0xffffffffu, u, 1);
                uint32_t prevb = lane == 0 ? hcarry : (up >> 24);
                hcarry = __shfl_sync(0xffffffffu, u, 31) >> 24;
                uint32_t dg = digest_word(u, (u << 8) | prevb, hpc);
                uint32_t m = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) if (wa + b < ge) m |= 0xFFu << (8 * b);
                dg &= m;
                uint32_t runs = (dg >> 3) & 0x01010101u;
                uint32_t mine = __popc(runs), tot;
                uint32_t before = warp_excl_scan(mine, &tot);
                uint32_t r = hcount + before;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if ((dg >> (8 * b)) & D_RUN) { if (r < l - 1) { H[r] = (uint8_t)(dg >> (8 * b)); anyN |= (dg >> (8 * b)) & D_N; } r++; }
                }
                hcount = min(hcount + tot, l - 1);
                haddr += 128;
            }
        }
        anyN = __any_sync(0xffffffffu, anyN != 0);
        __syncwarp();
        // logical end of the symbols a right-walk may visit: a full tile continues into its halo, a
        // partial (record-final) tile ends with the record -- do not walk its blank tail
        const uint32_t data_end = own_hi < TWs ? own_hi : TWs + hcount;

        // ---- warm-up: window of the l-1 symbols right of my chunk (append mode) ------------------
        LaneState st; st.F = 0; st.R = 0; st.W = 0; st.WN = 0;
        const uint32_t lo = lane * Cs, hi = lo + Cs;
        uint32_t m = 0;                                        // real symbols in the window
        {
            uint32_t x = hi, p = (lane + 1) * stride, xo = 0;
            while (m < l - 1 && x < data_end) {
                uint32_t d = S[p];
                if (d & D_RUN) {
                    uint32_t c = d & 3;
                    if (!(d & D_N)) { st.F ^= rol64(T.h[c], l - 1 - m); st.R ^= rol64(T.hc[c], m); }
                    else st.WN |= 1u << (l - 1 - m);
                    st.W |= (uint64_t)c << (2 * (l - 1 - m));
                    m++;
                }
                x++; p++; xo++;
                if (xo == Cs && x <= TWs) { p += LANE_PAD; xo = 0; }
            }
        }
        uint32_t need = (l - 1) - m;                           // symbols to consume before a window is complete
        // Indices m..l-1 of the window are still empty (index l-1 always is: the warm-up collects l-1
        // symbols).  Fill them with phantom 'A's: every prepend drops index l-1, so each phantom is
        // XORed out again exactly when it leaves and never reaches an emitted hash.
        for (uint32_t i = m; i < l; i++) {
            st.F ^= rol64(T.h[0], l - 1 - i); st.R ^= rol64(T.hc[0], i);
        }

        // ---- main backward scan over my chunk --------------------------------------------------
        uint32_t *tev = &ev_cnt[wid];
        int x = (int)hi - 1;
        const uint8_t *P = S + lane * stride;                  // P[x - lo]
        // non-emitting steps until the window is complete (record-end lanes only)
        while (need > 0 && x >= (int)lo) {
            uint32_t d = P[x - (int)lo];
            if (d & D_RUN) { step_generic(st, d, T, l); need--; }
            x--;
        }
        if (anyN) {
            for (; x >= (int)lo; x--) {
                uint32_t d = P[x - (int)lo];
                if (d & D_RUN) {
                    step_generic(st, d, T, l);
                    uint32_t fh = (uint32_t)(st.F >> 32), rh = (uint32_t)(st.R >> 32);
                    if (min(fh, rh) <= bound_hi) {
                        uint64_t h = st.F < st.R ? st.F : st.R;
                        if (h < a.bound && (uint32_t)x - xlo < xlim - xlo) emit_event<true>((uint32_t)x, h, lane, nloc, tev, tile, a);
                    }
                }
            }
        } else {
            uint64_t F = st.F, R = st.R, W = st.W;
            const uint32_t sh = 2 * l - 2;
            for (; x >= (int)lo; x--) {
                uint32_t d = P[x - (int)lo];
                if (d & D_RUN) {
                    uint32_t in = d & 3;
                    uint32_t idx = in | (((uint32_t)W & 3u) << 2);
                    F = ror1(F) ^ T.pairF[idx];
                    R = rol1(R) ^ T.pairR[idx];
                    W = (W >> 2) | ((uint64_t)in << sh);
                    uint32_t fh = (uint32_t)(F >> 32), rh = (uint32_t)(R >> 32);
                    if (min(fh, rh) <= bound_hi) {
                        uint64_t h = F < R ? F : R;
                        if (h < a.bound && (uint32_t)x - xlo < xlim - xlo) emit_event<false>((uint32_t)x, h, lane, nloc, tev, tile, a);
                    }
                }
            }
        }
        __syncwarp();
        a.lane_cnt[(uint64_t)tile * 32 + lane] = (uint16_t)nloc;
        if (lane == 0) a.tile_cnt[tile] = *tev;
        __syncwarp();
    }
}

// ordered compaction: event (lane, j) of tile t -> rank = excl(lane) + n_lane - 1 - j
struct GatherArgs {
    const uint64_t *ev_hash; const uint32_t *ev_meta; const uint16_t *lane_cnt;
    const uint32_t *tile_base;       // exclusive scan of the per-tile totals, n_tiles+1 entries
    const uint32_t *tile_seq; const uint32_t *first_tile; const uint64_t *offs;
    const uint32_t *pos_base;        // per record: position of its first byte inside its reference (or NULL)
    uint32_t n_tiles;
    uint32_t grid_align;             // 4: k_scan_minimizers (v1) tiles, 16: k_scan_minimizers_v2 tiles
    uint32_t *out_pos; uint64_t *out_hash;
};
__device__ __forceinline__ void tile_origin(const GatherArgs &g, uint32_t tile, int64_t *x0_to_pos) {
    uint32_t sq = g.tile_seq[tile];
    const uint64_t am = (uint64_t)g.grid_align - 1;
    uint64_t gs = g.offs[sq], ge = g.offs[sq + 1], A = gs & ~am;
    uint32_t ft = g.first_tile[sq], nt = g.first_tile[sq + 1] - ft, ti = tile - ft;
    const uint32_t span = (uint32_t)(ge - A), dv = 32u * nt;
    uint32_t Cs = (span + dv - 1) / dv;
    const uint32_t cm = g.grid_align == 16 ? 15u : 7u;
    Cs = (Cs + cm) & ~cm;
    uint64_t tlo = A + (uint64_t)ti * 32u * Cs;
    *x0_to_pos = (int64_t)tlo - (int64_t)gs + (g.pos_base ? (int64_t)g.pos_base[sq] : 0);
}
// A warp handles GATHER_TPW tiles (typically ~50 events each; 1 is fastest: 0.138 ms vs 0.144 / 0.160 ms for 2 / 4 on the
// bench workload -- more warps beat more loads per warp), so the kernel lives on memory-level parallelism: the
// loads of all its tiles are issued in two dependency levels -- everything addressable from the tile id first, then the
// record geometry and the first 64 events together -- instead of one dependent load after another.
#ifndef MQ_GATHER_TPW
#define MQ_GATHER_TPW 1
#endif
constexpr int GATHER_TPW = MQ_GATHER_TPW;
__global__ void __launch_bounds__(256) k_gather_minimizers(GatherArgs g) {
    const uint32_t tile0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GATHER_TPW;
    if (tile0 >= g.n_tiles) return;
    const uint32_t lane = lane_id();
    uint32_t base[GATHER_TPW], total[GATHER_TPW], n[GATHER_TPW], sq[GATHER_TPW];
    // level 1
#pragma unroll
    for (int t = 0; t < GATHER_TPW; t++) {
        const uint32_t tile = tile0 + t;
        base[t] = 0; total[t] = 0; n[t] = 0; sq[t] = 0;
        if (tile < g.n_tiles) {
            base[t] = __ldg(g.tile_base + tile); total[t] = __ldg(g.tile_base + tile + 1);
            n[t] = __ldg(g.lane_cnt + (uint64_t)tile * 32 + lane);
            sq[t] = __ldg(g.tile_seq + tile);
        }
    }
    // level 2: geometry of the records and the first two rounds of events
    uint32_t meta0[GATHER_TPW], meta1[GATHER_TPW], staged[GATHER_TPW], ft[GATHER_TPW], nt[GATHER_TPW];
    uint64_t h0[GATHER_TPW], h1[GATHER_TPW], gs[GATHER_TPW], ge[GATHER_TPW]; int64_t pb[GATHER_TPW];
#pragma unroll
    for (int t = 0; t < GATHER_TPW; t++) {
        total[t] -= base[t];
        staged[t] = total[t] < EV_CAP ? total[t] : EV_CAP;
        const uint64_t eb = (uint64_t)(tile0 + t) * EV_CAP;
        meta0[t] = meta1[t] = 0; h0[t] = h1[t] = 0; gs[t] = ge[t] = 0; ft[t] = 0; nt[t] = 1; pb[t] = 0;
        if (total[t]) {
            if (lane < staged[t]) { meta0[t] = g.ev_meta[eb + lane]; h0[t] = g.ev_hash[eb + lane]; }
            if (lane + 32 < staged[t]) { meta1[t] = g.ev_meta[eb + lane + 32]; h1[t] = g.ev_hash[eb + lane + 32]; }
            gs[t] = __ldg(g.offs + sq[t]); ge[t] = __ldg(g.offs + sq[t] + 1);
            ft[t] = __ldg(g.first_tile + sq[t]); nt[t] = __ldg(g.first_tile + sq[t] + 1) - ft[t];
            pb[t] = g.pos_base ? (int64_t)__ldg(g.pos_base + sq[t]) : 0;
        }
    }
#pragma unroll
    for (int t = 0; t < GATHER_TPW; t++) {
        if (total[t] == 0) continue;                       // warp-uniform
        const uint32_t tile = tile0 + t;
        const uint64_t eb = (uint64_t)tile * EV_CAP;
        const uint64_t am = (uint64_t)g.grid_align - 1, A = gs[t] & ~am;
        const uint32_t ti = tile - ft[t];
        uint32_t tot;
        const uint32_t ex = warp_excl_scan(n[t], &tot);
        const uint32_t span = (uint32_t)(ge[t] - A), dv = 32u * nt[t];
        uint32_t Cs = (span + dv - 1) / dv;
        const uint32_t cm = g.grid_align == 16 ? 15u : 7u;
        Cs = (Cs + cm) & ~cm;
        const int64_t org = (int64_t)(A + (uint64_t)ti * 32u * Cs) - (int64_t)gs[t] + pb[t];
        for (uint32_t s0 = 0; s0 < staged[t]; s0 += 32) {
            const uint32_t s = s0 + lane;
            uint32_t meta; uint64_t h;
            if (s0 == 0) { meta = meta0[t]; h = h0[t]; }
            else if (s0 == 32) { meta = meta1[t]; h = h1[t]; }
            else { meta = 0; h = 0; if (s < staged[t]) { meta = g.ev_meta[eb + s]; h = g.ev_hash[eb + s]; } }
            const uint32_t ln = (meta >> 14) & 31, j = meta >> 19;
            const uint32_t e = __shfl_sync(0xffffffffu, ex, ln), c = __shfl_sync(0xffffffffu, n[t], ln);
            if (s < staged[t]) {
                const uint32_t rank = e + c - 1 - j;
                g.out_pos[base[t] + rank] = (uint32_t)((int64_t)(meta & 0x3FFF) + org);
                g.out_hash[base[t] + rank] = h;
            }
        }
    }
}
__global__ void k_gather_overflow(GatherArgs g, const uint32_t *ovf_tile, const uint32_t *ovf_meta, const uint64_t *ovf_hash,
                                  uint32_t n_ovf) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ovf) return;
    uint32_t tile = ovf_tile[i], meta = ovf_meta[i];
    uint32_t ln = (meta >> 14) & 31, j = meta >> 19, e = 0;
    for (uint32_t q = 0; q < ln; q++) e += g.lane_cnt[(uint64_t)tile * 32 + q];
    uint32_t c = g.lane_cnt[(uint64_t)tile * 32 + ln];
    int64_t org; tile_origin(g, tile, &org);
    uint32_t rank = e + c - 1 - j;
    g.out_pos[g.tile_base[tile] + rank] = (uint32_t)((int64_t)(meta & 0x3FFF) + org);
    g.out_hash[g.tile_base[tile] + rank] = ovf_hash[i];
}
// seq_off[i] = tile_base[first_tile[i]]  (i <= n; tile_base has n_tiles+1 entries)
__global__ void k_seq_mini_off(const uint32_t *first_tile, const uint32_t *tile_base, uint32_t n, uint32_t *seq_off) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) seq_off[i] = tile_base[first_tile[i]];
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
        const uint4 *p = (const uint4 *)&t.slots[i];
        uint4 a = __ldg(p), b = __ldg(p + 1);
        uint64_t k = (uint64_t)a.x | ((uint64_t)a.y << 32);
        if (k == key) {
            if (b.z != 1u) return false;                    // count != 1: tombstone (or untouched spare slot)
            e->id = a.z; e->start = a.w; e->end = b.x; e->offrc = b.y;
            return true;
        }
        if (k == EMPTY_KEY) return false;
        i = (i + 1) & t.mask;
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
    const uint32_t *seq_off;                       // n+1 : minimizer range of each read
    uint32_t n_reads, k, l;
    MatchRec *matches;                             // read r writes at seq_off[r] + ordinal
    uint32_t *n_matches;                           // per read
    uint32_t *read_ticket;
};

__global__ void __launch_bounds__(128) k_probe_match(ProbeArgs a, Table t) {
    const uint32_t lane = lane_id();
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(a.read_ticket, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.n_reads) break;
        const uint32_t m0 = a.seq_off[r], m1 = a.seq_off[r + 1];
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
            if (j < Q) {
                uint64_t key = kminmer_hash_at(a.hash + m0 + j, a.k, &qrev);
                qstart = __ldg(a.pos + m0 + j); qend = __ldg(a.pos + m0 + j + a.k - 1) + a.l;
                hit = table_get(t, key, &e);
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
    const MatchRec *matches; const uint32_t *n_matches; const uint32_t *seq_off; const uint64_t *offs;
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
    if (r >= a.n_reads) return;
    const uint32_t n = a.n_matches[r];
    if (n > CHAIN_SMALL) { a.big_list[atomicAdd(a.big_count, 1u)] = r; return; }
    const MatchRec *ms = a.matches + a.seq_off[r];
    if (n <= 1) chain_thread<1>(a, r, n, ms);
    else if (n == 2) chain_thread<2>(a, r, n, ms);
    else if (n <= 4) chain_thread<4>(a, r, n, ms);
    else chain_thread<CHAIN_SMALL>(a, r, n, ms);
}


__global__ void __launch_bounds__(128) k_chain(ChainArgs a) {
    const uint32_t lane = lane_id();
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(a.read_ticket, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (a.big_list) { if (r >= *a.big_count) break; r = a.big_list[r]; }      // only the reads k_chain_small queued
        else if (r >= a.n_reads) break;
        const uint32_t n = a.n_matches[r];
        const MatchRec *ms = a.matches + a.seq_off[r];
        // best / second-best chain score over references (mers.rs:110-129)
        uint64_t best_score = 0, second = 0; uint32_t groups = 0;
        uint32_t b_ref = 0, b_rc = 0, b_mapq = 0; uint64_t b_qs = 0, b_qe = 0, b_rs = 0, b_re = 0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t ref = ms[i].ref_rc >> 1;
            // leader = first match of its reference
            bool seen = false;
            for (uint32_t q = lane; q < i; q += 32) seen |= (ms[q].ref_rc >> 1) == ref;
            if (__any_sync(0xffffffffu, seen)) continue;
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
