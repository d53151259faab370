// mq_scan.cuh -- S1: homopolymer compression + ntHash-1 canonical l-mer hash + universe-minimizer sampling
// (the KminmersIterator stage 1 the reference calls at mers.rs:27,53; spec in DESIGN.md section 2).
//
// One kernel, k_scan_minimizers<HPC, PACKED>, two input formats:
//   PACKED = false  upper-cased ASCII, one byte per base (what closures.rs:63,106 hands to the hot path);
//   PACKED = true   2-bit codes, 16 bases per 32-bit word (code = (ascii >> 1) & 3: A0 C1 T2 G3), a bitmap with one bit
//                   per 64-base block that holds a byte other than A/C/G/T, and a sorted list of exception intervals
//                   (start, len, byte) carrying those bytes -- together exactly the ASCII sequence (mq_pack, mq_lib.cu).
//
// Work decomposition (unchanged from round 1's third-generation kernel): a warp owns a TILE of up to 4096 raw bases of
// one record, a lane owns a chunk of up to 128 raw bases of it.  Each lane
//   1. stages its chunk: finds the homopolymer-run starts and compacts them into a private stream of one byte per
//      SYMBOL (code << 3, bit 7 = not A/C/G/T) in shared memory.  Streams are LANE-INTERLEAVED (word w of lane L at
//      row w, column L), so lane L only ever touches bank L: one wavefront per LDS/STS whatever word each lane is at.
//      Compaction appends to a pending register and stores one aligned word when it fills.
//      ASCII: SWAR run detection + PRMT compaction per 4-byte word.  PACKED: run starts of 16 bases are
//      x ^ (x << 2 | carry), one PRMT compaction per 4 bases, no per-byte validity test (the block bitmap says
//      whether the tile may take this path at all).
//   2. appends the next l-1 symbols (from the streams to its right, or the tile's halo) as context;
//   3. walks its column from the top down: warm-up over the context, then one hash step per own symbol -- the
//      outgoing symbol of the window is the same stream read l symbols further, four symbols per shared-memory word,
//      `in | out << 2` pre-scaled so that it is the byte offset into the 16-entry pair tables;
//   4. parks l-mers that pass a 32-bit pre-filter (hash hi word <= bound hi word) in the dead rows of its own column
//      and resolves them (exact 64-bit test, symbol ordinal -> raw position via per-group run masks) afterwards.
// Tiles that hold a byte other than A/C/G/T (in the chunk, its left neighbour byte, the context or the halo) take the
// generic path: symbols carry a flag and hash as 0 (the `nthash` crate's h(N) = 0); it reads bytes through
// src_*(), which for PACKED input rebuilds them from codes + exception intervals.
// Outputs: per-tile event pools tagged (lane, ordinal), lane counts, tile totals; k_gather_minimizers (mq_kernels.cuh)
// turns them into the position-ordered minimizer list.
#pragma once
#include <cstddef>
#include <cstring>
#include "mq_kernels.cuh"

namespace mq {

#ifndef MQ_SCAN_WARPS
#define MQ_SCAN_WARPS 2
#endif
constexpr int GPL_MAX  = CS_MAX / 16;       // 16-base groups per lane
constexpr int STRIDE   = CS_MAX + 32 + 4;   // bytes per lane stream: symbols + context + zero
static_assert(CS_MAX <= 256 && CS_MAX % 64 == 0, "lane chunks: whole 16-base groups, four per cum[] word, ordinals in one byte");
constexpr int SCAN_WARPS = MQ_SCAN_WARPS;   // warps (= tiles in flight) per CTA
constexpr int SYM_SH = 3;                   // symbol byte = code << 3 (offset into the 8-byte pair-table rows)

struct ExcRec { uint64_t start; uint32_t len; uint32_t byte; };   // == mq_exc: bases [start, start+len) are `byte`

struct ScanTables {                         // byte offsets are used directly by the kernel
    uint64_t pairF[16], pairR[16];          // @0, @128 : [in + 4*out]: rol(h(in), l-1) ^ ror(h(out), 1) | hc(in) ^ rol(hc(out), l)
    uint64_t inF[4], outF[4], inR[4], outR[4];  // @256, @288, @320, @352: the single-symbol parts (generic path)
    uint64_t F0, R0;                        // @384: hash state of a window of l phantom 'A's
    uint32_t sel[16];                       // @400: PRMT selectors compacting the run-start bytes of a word (index = 4 run bits)
    uint32_t opq[8];                        // @464: 2^31, 2, bound_hi, &T -- read back through volatile loads (see hash_step)
    uint32_t sel55[86];                     // @496: indexed by run bits at even positions (mask 0x55): selector | run bits << 16 | 8 * count << 24
};
static_assert(offsetof(ScanTables, sel) == 400 && offsetof(ScanTables, opq) == 464 && offsetof(ScanTables, sel55) == 496, "the kernel addresses these fields by byte offset");

struct ScanArgs {
    const uint8_t  *seqs;          // ASCII input: concatenated records, 16-byte aligned, >= 64 readable bytes behind the end
    const uint32_t *packed;        // PACKED input: 2-bit codes; base i = (packed[i >> 4] >> 2*(i & 15)) & 3; same slack
    const uint32_t *flags;         // PACKED: bit (i >> 6) set <=> bases [64*(i>>6), +64) hold a byte other than A/C/G/T
    const ExcRec   *exc;           // PACKED: sorted, non-overlapping exception intervals
    uint32_t        n_exc;
    uint64_t        exc_base;      // PACKED: interval coordinates = base index + exc_base
    const uint64_t *offs;          // n+1 record boundaries (base indices)
    const uint32_t *first_tile;    // n+1
    const uint32_t *tile_seq;      // n_tiles
    uint32_t n_tiles;
    uint32_t l;
    uint64_t bound;
    // outputs
    uint64_t *ev_hash;             // n_tiles * EV_CAP
    uint32_t *ev_meta;             // n_tiles * EV_CAP   x' (14b) | lane<<14 (5b) | j<<19 (13b)
    uint16_t *lane_cnt;            // n_tiles * 32
    uint32_t *tile_cnt;            // n_tiles (total events of the tile, incl. overflowed)
    uint32_t *ovf_count;           // overflow pool: single counter
    uint32_t  ovf_cap;
    uint32_t *ovf_tile; uint32_t *ovf_meta; uint64_t *ovf_hash;
    uint32_t *tile_ticket;         // dynamic tile scheduler
    const uint32_t *emit_range;    // per record (or NULL): [lo, hi) record offsets; only l-mers STARTING inside are emitted
    const uint4    *warm;          // 256 x {F lo, F hi, R lo, R hi}: four warm-up steps at once (see warm_row)
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

// ---- small helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r;
}
// explicit shared-state-space accessors (32-bit shared addresses): keeps ptxas from re-deriving the
// generic->shared window base (S2R SR_CgaCtaId + LEA) inside the hot loops
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t lds64(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// One hash step: F = ror1(F) ^ tf, R = rol1(R) ^ tr, as funnel shifts on 32-bit halves (4 SHF + 4 LOP3).
struct Hash4 { uint32_t flo, fhi, rlo, rhi; };
__device__ __forceinline__ void hash_step(Hash4 &h, uint64_t tf, uint64_t tr) {
    const uint32_t tfl = (uint32_t)tf, tfh = (uint32_t)(tf >> 32), trl = (uint32_t)tr, trh = (uint32_t)(tr >> 32);
    { const uint32_t lo = h.flo, hi = h.fhi; h.flo = __funnelshift_r(lo, hi, 1) ^ tfl; h.fhi = __funnelshift_r(hi, lo, 1) ^ tfh; }
    { const uint32_t lo = h.rlo, hi = h.rhi; h.rlo = __funnelshift_l(hi, lo, 1) ^ trl; h.rhi = __funnelshift_l(lo, hi, 1) ^ trh; }
}
struct Hash2 { uint64_t F, R; };
// Four warm-up steps at once.  While the window still fills, the outgoing symbol is a phantom 'A', so a step is
// F <- ror1(F) ^ pairF[in]: linear in F.  Four of them are F <- ror4(F) ^ C[in3, in2, in1, in0] with C tabulated for all
// 256 symbol quadruples (host: fill_warm_table); x holds the four pre-scaled symbol bytes of one stream word.
__device__ __forceinline__ void warm_row(Hash4 &h, uint32_t x, const uint4 *__restrict__ warm) {
    const uint32_t y = (x >> SYM_SH) & 0x03030303u;
    const uint32_t idx = (y * 0x01041040u) >> 24;            // c0 | c1 << 2 | c2 << 4 | c3 << 6 (partial products never meet)
    const uint4 t = __ldg(warm + idx);
    { const uint32_t lo = h.flo, hi = h.fhi; h.flo = __funnelshift_r(lo, hi, 4) ^ t.x; h.fhi = __funnelshift_r(hi, lo, 4) ^ t.y; }
    { const uint32_t lo = h.rlo, hi = h.rhi; h.rlo = __funnelshift_l(hi, lo, 4) ^ t.z; h.rhi = __funnelshift_l(lo, hi, 4) ^ t.w; }
}

// ---- ASCII digest (SWAR over a 4-byte word) ----------------------------------------------------------------------
// 0 in every byte of u that is 'A', 'C', 'G' or 'T':  bits 1..2 are the code; the other six bits must read 0x41, or
// 0x50 when the code is 2 ('T' = 0x54) -- m marks code-2 bytes, m*0x0F + 0x41.. is the expected pattern
__device__ __forceinline__ uint32_t acgt_diff(uint32_t u) {
    const uint32_t m = (u >> 2) & ~(u >> 1) & 0x01010101u;
    return (u & 0xF9F9F9F9u) ^ (m * 0x0Fu + 0x41414141u);
}
// 0x80 in every byte of u that differs from the byte before it (pv = u shifted up one byte, previous byte shifted in)
template <bool HPC> __device__ __forceinline__ uint32_t run80(uint32_t u, uint32_t pv) {
    if (!HPC) return 0x80808080u;
    const uint32_t e = u ^ pv;
    return (((e & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | e) & 0x80808080u;
}
__device__ __forceinline__ uint32_t symw_of(uint32_t u) { return (u << (SYM_SH - 1)) & (0x03030303u << SYM_SH); }
__device__ __forceinline__ uint32_t bad80(uint32_t u) {
    const uint32_t diff = acgt_diff(u);
    return (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;
}

// ---- byte source of the generic paths ---------------------------------------------------------------------------
// ASCII letters of four 2-bit codes (low 8 bits of b): code order A C T G
__device__ __forceinline__ uint32_t decode4(uint32_t b) {
    const uint32_t sel = (b & 3u) | ((b & 0xCu) << 2) | ((b & 0x30u) << 4) | ((b & 0xC0u) << 6);
    return prmt(0x47544341u /* 'A','C','T','G' */, 0u, sel);
}
// patch the bytes of [x, x+4) that fall into an exception interval
__device__ __noinline__ uint32_t overlay4(const ScanArgs &a, uint64_t x, uint32_t u) {
    x += a.exc_base;
    uint32_t lo = 0, hi = a.n_exc;                     // first interval that ends behind x
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.exc[mid].start + a.exc[mid].len > x) hi = mid; else lo = mid + 1; }
    for (uint32_t i = lo; i < a.n_exc && a.exc[i].start < x + 4; i++) {
        const uint64_t s = a.exc[i].start, e = s + a.exc[i].len;
#pragma unroll
        for (int b = 0; b < 4; b++) if (x + b >= s && x + b < e) u = (u & ~(0xFFu << (8 * b))) | ((a.exc[i].byte & 0xFFu) << (8 * b));
    }
    return u;
}
__device__ __forceinline__ bool block_flag(const ScanArgs &a, uint64_t x) {
    if (a.n_exc == 0) return false;                  // no exception interval in this slice: the bitmap is all zero
    const uint64_t blk = x >> 6;
    return (__ldg(a.flags + (blk >> 5)) >> (blk & 31)) & 1u;
}
// The 16-base packed decoder is deliberately NOT inlined: it serves the rare paths only (record edges, tiles with a
// non-ACGT byte), and inlined copies at every call site made the packed kernel 70 % larger than the ASCII one -- enough
// to stall its warps on instruction fetch (ncu: 3.9 warps per issue waiting for "no instruction").
__device__ __forceinline__ uint32_t packed_u32(const ScanArgs &a, uint64_t x) {   // small, and inlined loads can overlap
    const uint32_t w = __ldg(a.packed + (x >> 4));
    uint32_t u = decode4(w >> (2u * ((uint32_t)x & 15u)));
    if (block_flag(a, x)) u = overlay4(a, x, u);
    return u;
}
__device__ __noinline__ uint4 packed_u128(const ScanArgs &a, uint64_t x) {
    const uint32_t w = __ldg(a.packed + (x >> 4));
    uint4 v = make_uint4(decode4(w), decode4(w >> 8), decode4(w >> 16), decode4(w >> 24));
    if (block_flag(a, x)) { v.x = overlay4(a, x, v.x); v.y = overlay4(a, x + 4, v.y); v.z = overlay4(a, x + 8, v.z); v.w = overlay4(a, x + 12, v.w); }
    return v;
}
// four bytes at the 4-aligned base index x
template <bool PACKED> __device__ __forceinline__ uint32_t src_u32(const ScanArgs &a, uint64_t x) {
    if (!PACKED) return __ldg((const uint32_t *)(a.seqs + x));
    return packed_u32(a, x);
}
template <bool PACKED> __device__ __forceinline__ uint4 src_u128(const ScanArgs &a, uint64_t x) {
    if (!PACKED) return __ldg((const uint4 *)(a.seqs + x));
    return packed_u128(a, x);
}
template <bool PACKED> __device__ __forceinline__ uint32_t src_u8(const ScanArgs &a, uint64_t x) {
    if (!PACKED) return a.seqs[x];
    return (src_u32<PACKED>(a, x & ~3ull) >> (8u * ((uint32_t)x & 3u))) & 0xFFu;
}
// PACKED: does any 64-base block touching [x0, x1) (clipped by the caller to the record) hold an exception?
__device__ __forceinline__ bool any_flag(const ScanArgs &a, uint64_t x0, uint64_t x1) {
    if (x1 <= x0) return false;
    for (uint64_t blk = x0 >> 6; blk <= ((x1 - 1) >> 6); blk++)
        if ((__ldg(a.flags + (blk >> 5)) >> (blk & 31)) & 1u) return true;
    return false;
}

// ---- per-warp shared-memory layout (byte offsets from the warp's base) --------------------------------------------
constexpr int ROWS      = STRIDE / 4 + 2;                            // words per lane column: symbols + context + zero, + 2 spare rows
constexpr int OFF_HALO  = ROWS * 128;                                // u8[48]  halo stream (contiguous; only lane 31 reads it)
constexpr int OFF_NSYM  = OFF_HALO + 48;                             // u32[33] symbols per stream (32 = halo)
constexpr int OFF_RUNM  = OFF_NSYM + 144;                            // u16 run masks: word j (groups 2j, 2j+1) of lane L at row j
constexpr int OFF_CUM   = OFF_RUNM + (GPL_MAX / 2) * 128;            // u8 symbol counts before each group, 4 per word, row-interleaved
constexpr int WARP_BYTES = (OFF_CUM + (GPL_MAX / 4) * 128 + 15) & ~15;
// Parked candidates live in the lane's OWN column, from the top row downwards, three rows each (hash lo, hash hi,
// ordinal): the scan walks its column from the top down, so the rows above the outgoing-symbol words are dead by the
// time candidates appear, and the list costs no shared memory of its own.
constexpr int CAND_TOP = (ROWS - 1) * 128;
constexpr uint32_t CUM_FILL = CS_MAX <= 128 ? 0x7F7F7F7Fu : 0xFFFFFFFFu;

// byte o of the stream whose column base is sb
__device__ __forceinline__ uint32_t col_baddr(uint32_t sb, uint32_t o) { return sb + o + (o >> 2) * 124u; }

// append cursor of a lane stream: P holds the bytes of the incomplete word (zero above them), n8 = 8 * symbols so far,
// wp = address of the incomplete word
struct Pend { uint32_t P, n8, wp; };
__device__ __forceinline__ void push(Pend &q, uint32_t comp, uint32_t c8) {       // comp: c8/8 bytes, zero above
    const uint32_t lo = q.P | __funnelshift_l(0u, comp, q.n8), hi = __funnelshift_l(comp, 0u, q.n8);   // (hi:lo) = comp << (n8 mod 32) | P
    const uint32_t n8n = q.n8 + c8;
    if ((n8n ^ q.n8) & 32u) { sts32(q.wp, lo); q.wp += 128u; q.P = hi; } else q.P = lo;
    q.n8 = n8n;
}

// position of the k-th (0-based) set bit of a 16-bit mask
__device__ __forceinline__ uint32_t select16(uint32_t m, uint32_t k) {
    uint32_t pos = 0, c;
    c = __popc(m & 0xFFu);        if (k >= c) { k -= c; pos += 8; m >>= 8; }
    c = __popc(m & 0xFu);         if (k >= c) { k -= c; pos += 4; m >>= 4; }
    c = __popc(m & 0x3u);         if (k >= c) { k -= c; pos += 2; m >>= 2; }
    c = m & 1u;                   if (k >= c) { pos += 1; }
    return pos;
}
// raw offset (inside the lane chunk) of the symbol with ordinal o: group = #(cum[g] <= o) - 1.
// cum[] holds the symbol count before each 16-base group; unused entries hold a sentinel no ordinal reaches (a chunk
// with unused groups has <= CS_MAX - 16 symbols).  With chunks of <= 128 bases every value is < 0x80, so "byte <= o" is
// one SWAR subtraction: bit 7 of (0x80|o) - byte; longer chunks use the byte-wise compare instruction.
__device__ __forceinline__ uint32_t raw_offset(uint32_t runm_l, uint32_t cum_l, uint32_t o) {
    uint32_t g = 0;
    const uint32_t ob = o * 0x01010101u;
#pragma unroll
    for (int w = 0; w < GPL_MAX / 4; w++) {
        const uint32_t cw = lds32(cum_l + 128 * w);
        if (CS_MAX <= 128) g += __popc(((ob | 0x80808080u) - cw) & 0x80808080u);
        else g += __popc(__vcmpleu4(cw, ob) & 0x01010101u);
    }
    g -= 1;
    const uint32_t base = lds8(cum_l + (g >> 2) * 128u + (g & 3u));
    return 16u * g + select16(lds16(runm_l + (g >> 1) * 128u + (g & 1u) * 2u), o - base);
}

// evh / evm: this tile's slice of the event pools (passed by value: the candidates are resolved in a non-inlined
// function, where reading them through `a` would be a generic load from parameter space per event)
__device__ __forceinline__ void emit_event(uint32_t x, uint64_t h, uint32_t lane, uint32_t j, uint32_t ev_a, uint32_t tile, const ScanArgs &a,
                                           uint64_t *evh, uint32_t *evm) {
    uint32_t slot;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(ev_a) : "memory");
    const uint32_t meta = x | (lane << 14) | (j << 19);
    if (slot < EV_CAP) {
        evh[slot] = h;
        evm[slot] = meta;
    } else {
        uint32_t g = atomicAdd(a.ovf_count, 1u);
        if (g < a.ovf_cap) { a.ovf_tile[g] = tile; a.ovf_meta[g] = meta; a.ovf_hash[g] = h; }
    }
}

// Resolve and emit the candidates parked in rows (cpz, top] of the lane's column, oldest first.  Deliberately few
// arguments: whatever can be re-derived from the warp's shared-memory base, the lane id and the kernel parameters is
// re-derived here, so that it does not occupy registers across the hot loop (64-register budget).
__device__ __noinline__ uint32_t flush_candidates(uint32_t ws_a, uint32_t cpz, uint32_t j0, uint32_t c_lo, uint32_t xlo, uint32_t xlim,
                                                  uint32_t ev_a, uint32_t tile, const ScanArgs &a, int o2) {
    const uint32_t lane = lane_id();
    const uint32_t top = ws_a + 4 * lane + CAND_TOP, runm_l = ws_a + OFF_RUNM + 4 * lane, cum_l = ws_a + OFF_CUM + 4 * lane;
    const uint64_t bound = a.bound;
    uint64_t *const evh = a.ev_hash + (uint64_t)tile * EV_CAP; uint32_t *const evm = a.ev_meta + (uint64_t)tile * EV_CAP;
    uint32_t j = j0;
    for (uint32_t p = top; p > cpz; p -= 384u) {
        const uint64_t h = ((uint64_t)lds32(p - 128u) << 32) | lds32(p);
        if (h >= bound) continue;                      // parked on the hi-word pre-filter only: exact test here
        const uint32_t o = lds32(p - 256u);
        if ((int)o > o2) continue;                     // a context symbol, or a window the record end leaves incomplete
        const uint32_t x = c_lo + raw_offset(runm_l, cum_l, o);
        if (x - xlo < xlim - xlo) { emit_event(x, h, lane, j, ev_a, tile, a, evh, evm); j++; }
    }
    return j;
}

// one ASCII word of a group that a record boundary cuts (or of a tile on the generic path): run bits masked to the
// bytes [lo_b, hi_b) of the word, a record start forces a run start
template <bool HPC, bool FLAG>
__device__ __forceinline__ void stage_cut_word(uint32_t u, uint32_t &prev, uint32_t x0, uint32_t own_lo, uint32_t own_hi, bool rec_start,
                                               uint32_t ta, Pend &q, uint32_t &rm, uint32_t &bad, int w) {
    const uint32_t r80 = run80<HPC>(u, (u << 8) | prev);
    prev = u >> 24;
    const uint32_t lo_b = own_lo > x0 ? min(own_lo - x0, 4u) : 0u, hi_b = own_hi > x0 ? min(own_hi - x0, 4u) : 0u;
    uint32_t p = (((r80 >> 7) * 0x01020408u) >> 24) & ((1u << hi_b) - 1u) & ~((1u << lo_b) - 1u);
    if (rec_start && own_lo >= x0 && own_lo < x0 + 4u && own_lo < own_hi) p |= 1u << (own_lo - x0);   // a record starts a run
    const uint32_t b80 = bad80(u);
    bad |= b80 & (((p * 0x00204081u) & 0x01010101u) << 7);      // flags of the selected bytes only
    const uint32_t symw = symw_of(u) | (FLAG ? b80 : 0u);
    const uint32_t comp = prmt(symw, 0u, lds32(ta + 400 + 4 * p));
    push(q, comp, __popc(p) << 3);
    rm |= p << (4 * w);
}

// What the packed fast path wants in flight BEFORE the block-bitmap vote decides that the tile may take it (otherwise the
// bitmap read, the context word and the first data word are three exposed memory latencies in a row): the word before
// the chunk (its top code is the run context) and the first word of the chunk's in-record groups.
__device__ __forceinline__ void packed_prefetch(const ScanArgs &a, uint64_t tlo, uint64_t gs, uint32_t c_lo, uint32_t gpl, uint32_t own_lo,
                                                uint32_t own_hi, uint32_t &w_prev, uint32_t &w0) {
    const uint32_t Cs = gpl << 4;
    const uint32_t *wp = a.packed + ((tlo + c_lo) >> 4);
    w_prev = 0; w0 = 0;
    if ((c_lo > own_lo && c_lo < own_hi) || (c_lo == own_lo && tlo + own_lo > gs)) w_prev = __ldg(wp - 1);
    const uint32_t lo_x = max(own_lo, c_lo), hi_x = min(own_hi, c_lo + Cs);
    if (lo_x < hi_x) {
        const uint32_t g0 = (lo_x - c_lo + 15) >> 4, g1 = (hi_x - c_lo) >> 4;
        if (g0 < g1) w0 = __ldg(wp + g0);
    }
}

// Stage + compact one lane chunk; leaves the append cursor in q (pending word NOT yet stored).
// Groups [g0, g1) lie completely inside the record: fast path.  The (at most two) groups cut by a record boundary go
// through stage_cut_word; groups outside the record hold no symbol.
// FLAG = false leaves the non-ACGT flag (bit 7) out of the symbol bytes; for ASCII input it reports in bad_out whether
// the chunk holds ANY byte other than A/C/G/T (the caller then stages the tile again with FLAG = true, which flags every
// symbol); PACKED tiles are dispatched by the block bitmap before staging.
template <bool HPC, bool PACKED, bool FLAG>
__device__ __forceinline__ void stage_chunk(const ScanArgs &a, uint64_t tlo, uint64_t gs, uint32_t c_lo, uint32_t gpl, uint32_t own_lo,
                                            uint32_t own_hi, uint32_t sb, uint32_t cum_l, uint32_t runm_l, uint32_t ta,
                                            Pend &q, uint32_t &bad_out, uint32_t pre_prev = 0, uint32_t pre_w0 = 0) {
    const uint32_t Cs = gpl << 4;
    uint32_t bad = 0;
    q.P = 0; q.n8 = 0; q.wp = sb;
    const uint64_t cx = tlo + c_lo;                      // base index of my first byte
    uint32_t prev = 0;                                   // byte before the next word (ASCII letter)
    constexpr bool FASTP = PACKED && !FLAG;              // packed fast path: 2-bit arithmetic, no byte reconstruction
    if ((c_lo > own_lo && c_lo < own_hi) || (c_lo == own_lo && tlo + own_lo > gs))     // byte before my chunk (same record)
        prev = FASTP ? prmt(0x47544341u, 0u, 0x4440u | (pre_prev >> 30)) : src_u8<PACKED>(a, cx - 1);   // fast path: its word came prefetched
    // a record that starts exactly at one of my group boundaries starts a run whatever the byte before it was
    const uint32_t gforce = (tlo + own_lo == gs && own_lo >= c_lo && own_lo < c_lo + Cs && ((own_lo - c_lo) & 15u) == 0u)
                                ? (own_lo - c_lo) >> 4 : 0xFFFFFFFFu;
#pragma unroll
    for (int r = 0; r < GPL_MAX / 4; r++) sts32(cum_l + 128 * r, CUM_FILL);
    const uint32_t lo_x = max(own_lo, c_lo), hi_x = min(own_hi, c_lo + Cs);
    uint32_t g0 = gpl, g1 = gpl;
    if (lo_x < hi_x) { g0 = (lo_x - c_lo + 15) >> 4; g1 = (hi_x - c_lo) >> 4; if (g1 < g0) g1 = g0; }
    // bookkeeping of group g: symbol count before it (cum), its 16 run bits (runm); both row-interleaved
    auto open_group = [&](uint32_t g) { sts8(cum_l + (g >> 2) * 128u + (g & 3u), q.n8 >> 3); };
    auto close_group = [&](uint32_t g, uint32_t rm) { sts16(runm_l + (g >> 1) * 128u + (g & 1u) * 2u, rm); };
    // a group outside the record holds no symbol; one cut by a record boundary goes word by word with masked run bits
    auto edge_group = [&](uint32_t g, bool after_packed_fast) {
        uint32_t rm = 0;
        open_group(g);
        const uint32_t xg = c_lo + 16 * g;
        if (xg < own_hi && xg + 16 > own_lo) {
            if (after_packed_fast) prev = prmt(0x47544341u, 0u, 0x4440u | ((prev >> 1) & 3u));   // code -> letter (one byte)
            const uint4 v = src_u128<PACKED>(a, cx + 16 * g);
            const uint32_t uw[4] = {v.x, v.y, v.z, v.w};
            const bool rec_start = tlo + own_lo == gs;
#pragma unroll
            for (int w = 0; w < 4; w++) stage_cut_word<HPC, FLAG>(uw[w], prev, xg + 4 * w, own_lo, own_hi, rec_start, ta, q, rm, bad, w);
        }
        close_group(g, rm);
    };
    if (FASTP) {
        const uint32_t *wp = a.packed + (cx >> 4);
        uint32_t nxw = pre_w0;                                 // one word (16 bases) ahead; the first one came prefetched
        for (uint32_t g = 0; g < gpl; g++) {
            if (g >= g0 && g < g1) {
                uint32_t rm = 0;
                open_group(g);
                const uint32_t x = nxw;
                if (g + 1 < g1) nxw = __ldg(wp + g + 1);
                if (g == gforce) prev = (~x & 3u) << 1;
                // run starts of 16 bases: a base starts a run iff its code differs from the code before it
                uint32_t m = 0x55555555u;
                if (HPC) { const uint32_t d = x ^ ((x << 2) | ((prev >> 1) & 3u)); m = (d | (d >> 1)) & 0x55555555u; }
                prev = (x >> 30) << 1;                    // only bits 1..2 of `prev` are read on this path
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const uint32_t b = (x >> (8 * w)) & 0xFFu;
                    uint32_t t = (b | (b << 12)) & 0x000F000Fu;                       // codes 0,1 | 2,3 into separate half-words
                    t = ((t | (t << 6)) & 0x03030303u) << SYM_SH;                     // one code per byte, pre-scaled
                    // table row of the byte's four run bits: PRMT selector | run bits << 16 | 8 * count << 24
                    const uint32_t e = lds32(ta + 496 + ((m >> (8 * w)) & 0x55u) * 4u);
                    push(q, prmt(t, 0u, e), e >> 24);                                 // run-start bytes first, zero fill
                    rm |= ((e >> 16) & 0xFu) << (4 * w);
                }
                close_group(g, rm);
            } else edge_group(g, g == g1 && g1 > g0);
        }
    } else {
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (g0 < g1) nxt = src_u128<PACKED>(a, cx + 16 * g0);   // prefetch: one group ahead
        for (uint32_t g = 0; g < gpl; g++) {
            if (g >= g0 && g < g1) {
                uint32_t rm = 0;
                open_group(g);
                const uint4 v = nxt;
                if (g + 1 < g1) nxt = src_u128<PACKED>(a, cx + 16 * (g + 1));
                if (g == gforce) prev = (v.x & 0xFFu) ^ 0xFFu;
                const uint32_t uw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const uint32_t u = uw[w];
                    const uint32_t r80 = run80<HPC>(u, (u << 8) | prev);
                    prev = u >> 24;
                    uint32_t symw = symw_of(u);
                    if (!FLAG) bad |= acgt_diff(u);
                    else { symw |= bad80(u); bad |= symw & r80; }
                    // r80 has bit 7 of byte i set for a run start: * 0x00204081 moves them to bits 28..31 (no two partial
                    // products meet, so nothing carries)
                    const uint32_t p = (r80 * 0x00204081u) >> 28;                     // the four run bits
                    const uint32_t comp = prmt(symw, 0u, lds32(p * 4u + (ta + 400)));  // run-start bytes first, zero fill
                    push(q, comp, __popc(p) << 3);
                    rm |= p << (4 * w);
                }
                close_group(g, rm);
            } else edge_group(g, false);
        }
    }
    bad_out = bad;
}

// generic (N-aware, bounds-checked) step at ordinal o of the lane's logical stream
__device__ __forceinline__ void step_generic(Hash2 &s, uint32_t sb, int o, int lim, uint32_t l, uint32_t ta) {
    const uint32_t in = lds8(col_baddr(sb, (uint32_t)o));
    const int oo = o + (int)l;
    const uint32_t out = oo < lim ? lds8(col_baddr(sb, (uint32_t)oo)) : 0u;
    const uint32_t io = ((in >> SYM_SH) & 3u) * 8u, oo8 = ((out >> SYM_SH) & 3u) * 8u;
    const uint64_t tf = ((in & 0x80u) ? 0ull : lds64(ta + 256 + io)) ^ ((out & 0x80u) ? 0ull : lds64(ta + 288 + oo8));
    const uint64_t tr = ((in & 0x80u) ? 0ull : lds64(ta + 320 + io)) ^ ((out & 0x80u) ? 0ull : lds64(ta + 352 + oo8));
    s.F = ror1(s.F) ^ tf; s.R = rol1(s.R) ^ tr;
}

// park: three stores down the column; flush when the next candidate would reach the rows the scan still reads.
// cq is the cursor minus ck = (l/4)*128 + 384, so that inside the word loop the test is simply cq <= wa (rows up to
// wa + (l/4)*128 + 128 are still read, and a candidate needs three rows).
#define MQ_CANDIDATE(ORD, LIVE)                                                                               \
    if (min(H.fhi, H.rhi) <= bound_hi) {                                                                      \
        const bool fmin = (((uint64_t)H.fhi << 32) | H.flo) < (((uint64_t)H.rhi << 32) | H.rlo);              \
        const uint32_t cpz = cq + ck;                                                                         \
        sts32(cpz, fmin ? H.flo : H.rlo); sts32(cpz - 128u, min(H.fhi, H.rhi)); sts32(cpz - 256u, (uint32_t)(ORD)); \
        cq -= 384u;                                                                                           \
        if (cq <= (LIVE)) { nloc = flush_candidates(ws_a, cq + ck, nloc, c_lo, xlo, xlim, ev_a, tile, a, o2); cq = ctop - ck; } \
    }

// HPC and the input format are compile-time flags (four instantiations)
template <bool HPC, bool PACKED>
__global__ void __launch_bounds__(SCAN_WARPS * 32, 32 / SCAN_WARPS) k_scan_minimizers(const __grid_constant__ ScanArgs a, const __grid_constant__ ScanTables Tin) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(256) ScanTables T;
    __shared__ uint32_t ev_cnt[SCAN_WARPS];
    for (uint32_t i = threadIdx.x; i < sizeof(ScanTables) / 4; i += blockDim.x) ((uint32_t *)&T)[i] = ((const uint32_t *)&Tin)[i];
    __syncthreads();
    if (threadIdx.x == 0) { T.opq[0] = 0x80000000u; T.opq[1] = 2u; T.opq[2] = (uint32_t)(a.bound >> 32); T.opq[3] = smem_addr(&T); }
    __syncthreads();
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    const uint32_t ws_a = smem_addr(smem_raw + (size_t)wid * WARP_BYTES);
    const uint32_t sb = ws_a + 4 * lane;                              // my stream: column `lane` of the rows
    const uint32_t ha = ws_a + OFF_HALO;
    const uint32_t nsym_a = ws_a + OFF_NSYM;
    const uint32_t runm_l = ws_a + OFF_RUNM + 4 * lane;
    const uint32_t cum_l = ws_a + OFF_CUM + 4 * lane;
    const uint32_t ctop = sb + CAND_TOP;                              // first candidate slot: the top row of my column
    // right neighbour's stream: column lane+1, or the contiguous halo for lane 31
    const uint32_t nb_a = lane < 31 ? sb + 4 : ha, nb_st = lane < 31 ? 128u : 4u;
    // scalars read back through volatile shared loads so that ptxas keeps them in registers instead of re-deriving
    // them (S2UR/ULEA/LDCU) inside the hot loop
    const uint32_t ta = lds32(smem_addr(&T) + 464 + 12);
    const uint32_t ev_a = smem_addr(&ev_cnt[wid]);
    const uint32_t l = a.l;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) break;
        if (lane == 0) sts32(ev_a, 0u);

        // ---- geometry -------------------------------------------------------------------------
        const uint32_t sq = a.tile_seq[tile];
        const uint64_t gs = a.offs[sq], ge = a.offs[sq + 1];
        const uint32_t ft = a.first_tile[sq], nt = a.first_tile[sq + 1] - ft, ti = tile - ft;
        uint32_t Cs; uint64_t tlo;
        tile_geometry(gs, ge, nt, ti, &Cs, &tlo);
        const uint32_t TWs = 32u * Cs, gpl = Cs >> 4;
        uint32_t nloc = 0;
        if (tlo >= ge) {
            a.lane_cnt[(uint64_t)tile * 32 + lane] = 0;
            if (lane == 0) a.tile_cnt[tile] = 0;
            continue;
        }
        const uint32_t own_lo = gs > tlo ? (uint32_t)(gs - tlo) : 0u;
        const uint32_t own_hi = (ge - tlo) < TWs ? (uint32_t)(ge - tlo) : TWs;
        uint32_t xlo, xlim; emit_window(a, sq, gs, tlo, &xlo, &xlim);

        // ---- stage + compact my chunk: one byte per homopolymer-run start ---------------------------
        const uint32_t c_lo = lane * Cs;                       // x' of my first byte
        Pend q; uint32_t bad = 0;
        bool generic;
        if (PACKED) {
            // the block bitmap decides: my chunk and the byte before it, clipped to the record
            const uint64_t r0 = max(tlo + c_lo, gs + 1) - 1, r1 = min(tlo + c_lo + Cs, ge);
            uint32_t w_prev, w0;
            packed_prefetch(a, tlo, gs, c_lo, gpl, own_lo, own_hi, w_prev, w0);       // in flight while the bitmap is read
            // a slice without a single exception interval (reads without N: the usual case) has an all-zero bitmap
            generic = a.n_exc != 0 && __any_sync(0xffffffffu, any_flag(a, r0, r1));
            if (!generic) stage_chunk<HPC, true, false>(a, tlo, gs, c_lo, gpl, own_lo, own_hi, sb, cum_l, runm_l, ta, q, bad, w_prev, w0);
        } else {
            stage_chunk<HPC, false, false>(a, tlo, gs, c_lo, gpl, own_lo, own_hi, sb, cum_l, runm_l, ta, q, bad);
            generic = __any_sync(0xffffffffu, bad != 0);       // some byte is not A/C/G/T: stage again with per-symbol flags
            if (generic) __syncwarp();
        }
        if (generic) stage_chunk<HPC, PACKED, true>(a, tlo, gs, c_lo, gpl, own_lo, own_hi, sb, cum_l, runm_l, ta, q, bad);
        const uint32_t n = q.n8 >> 3;
        // my incomplete word is published in the top row of my column (free until candidates are parked), NOT in
        // place: complete words below it are then immutable while neighbours read them, and the context append
        // below needs no second barrier
        sts32(ctop, q.P);
        sts32(nsym_a + 4 * lane, n);
        const uint32_t nz = __ballot_sync(0xffffffffu, n != 0);
        if (nz == 0) {
            // the whole tile lies inside one homopolymer run (or outside the record): no l-mer starts here,
            // so neither halo nor context is needed -- this keeps giant runs (N-gaps) linear instead of quadratic
            a.lane_cnt[(uint64_t)tile * 32 + lane] = 0;
            if (lane == 0) a.tile_cnt[tile] = 0;
            __syncwarp();
            continue;
        }

        // ---- halo stream: up to 32 (>= l-1) run-start symbols right of the tile -------------------------
        uint32_t hcount = 0;
        if (tlo + TWs < ge) {
            uint64_t haddr = tlo + TWs;
            uint32_t hcarry = src_u8<PACKED>(a, haddr - 1);
            while (hcount < 32u && haddr < ge) {
                const uint64_t wa = haddr + 4ull * lane;
                uint32_t u = (wa < ge) ? src_u32<PACKED>(a, wa) : 0u;
                uint32_t up = __shfl_up_sync(0xffffffffu, u, 1);
                uint32_t prevb = lane == 0 ? hcarry : (up >> 24);
                hcarry = __shfl_sync(0xffffffffu, u, 31) >> 24;
                const uint32_t r80 = run80<HPC>(u, (u << 8) | prevb), symw = symw_of(u) | bad80(u);
                uint32_t m = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) if (wa + b < ge) m |= 0x80u << (8 * b);
                const uint32_t run = r80 & m;
                uint32_t mine = __popc(run), tot;
                uint32_t r = hcount + warp_excl_scan(mine, &tot);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (run & (0x80u << (8 * b))) {
                        const uint32_t sy = (symw >> (8 * b)) & 0xFFu;
                        if (r < 32u) { sts8(ha + r, sy); if (r < l - 1) bad |= sy & 0x80u; }
                        r++;
                    }
                }
                hcount = min(hcount + tot, 32u);
                haddr += 128;
            }
        }
        if (lane == 0) sts32(nsym_a + 4 * 32, hcount);
        const bool anyN = __any_sync(0xffffffffu, bad != 0);
        __syncwarp();

        // ---- context: the next l-1 symbols after my chunk, appended in place from the streams to my right --------
        const uint32_t wp0 = q.wp, sh0 = q.n8 & 24u;            // where the context starts in my column
        uint32_t c = 0;
        if (n != 0) {
            if (lds32(nsym_a + 4 * (lane + 1)) >= 32u) {
                // common case: the stream to my right alone holds the l-1 (<= 31) symbols -- its first eight words, cut
                // after l-1 bytes (everything behind the context must read as 0, a phantom 'A')
                c = l - 1;
                uint32_t P = q.P;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int keep = (int)l - 1 - 4 * i;        // warp-uniform
                    const uint32_t m = keep >= 4 ? 0xFFFFFFFFu : (keep <= 0 ? 0u : ((1u << (8 * keep)) - 1u));
                    const uint32_t x = lds32(nb_a + i * nb_st) & m;
                    sts32(wp0 + 128 * i, P | (x << sh0));
                    P = __funnelshift_l(x, 0u, sh0);
                }
                sts32(wp0 + 128 * 8, P);
                sts32(wp0 + 128 * 9, 0u);
            } else {
                // record end or short streams: walk the non-empty streams to my right (the halo last), word by word
                uint32_t need = l - 1, rest = lane < 31 ? nz >> (lane + 1) : 0u, jb = lane + 1;
                bool halo_left = hcount != 0;
                while (need) {
                    uint32_t j, nj;
                    if (rest) { const uint32_t sk = __ffs(rest) - 1; j = jb + sk; jb = j + 1; rest = sk == 31 ? 0u : rest >> (sk + 1); nj = lds32(nsym_a + 4 * j); }
                    else if (halo_left) { j = 32; nj = hcount; halo_left = false; }
                    else break;
                    const uint32_t take = min(nj, need);
                    for (uint32_t w = 0; 4 * w < take; w++) {
                        uint32_t x = j == 32 ? lds32(ha + 4 * w) : (w < (nj >> 2) ? lds32(ws_a + 4 * j + 128 * w) : lds32(ws_a + 4 * j + CAND_TOP));
                        const uint32_t nb = min(take - 4 * w, 4u);
                        if (nb < 4) x &= (1u << (8 * nb)) - 1u;
                        push(q, x, 8 * nb);
                    }
                    need -= take; c += take;
                }
                sts32(q.wp, q.P);                               // the incomplete word, zero above its bytes
                for (uint32_t za = q.wp + 128; za <= wp0 + 128 * 9; za += 128) sts32(za, 0u);
            }
        }
        __syncwarp();

        // ---- phase 1 + 2: one pass down my column, row by row ------------------------------------------
        // Rows above the one holding my last own symbol are warm-up: the outgoing symbol of those steps lies beyond the
        // stream, i.e. it is a phantom 'A' (code 0), and the pair-table row for out == 0 is the whole step.  Zero bytes
        // behind the context are phantom 'A's entering a window of phantom 'A's, which leaves the state unchanged, so
        // whole rows are stepped.  From that row down every step may park a candidate; the up to three context symbols
        // sharing the row, and windows a record end leaves incomplete, are dropped at resolve time (ordinal > o2).
        Hash2 st; st.F = T.F0; st.R = T.R0;
        const int lim = (int)(n + c);                           // symbols available in my logical stream
        const int o2 = max(-1, min((int)n - 1, lim - (int)l));  // first ordinal whose window is complete and mine
        Hash4 H;
        H.flo = (uint32_t)st.F; H.fhi = (uint32_t)(st.F >> 32); H.rlo = (uint32_t)st.R; H.rhi = (uint32_t)(st.R >> 32);
        const uint32_t lr = (l >> 2) * 128u, ck = lr + 384u;
        uint32_t cq = ctop - ck;
        // read here, per tile, so that its live range does not span the staging code: with the 64-register budget ptxas
        // otherwise spills it and reloads it from local memory at every step of the hot loop (seen in the packed variant)
        const uint32_t bound_hi = lds32(ta + 464 + 8);
        if (anyN) {
            const uint32_t live0 = sb + 128u * ((uint32_t)lim >> 2) + 256u - ck;   // everything up to row lim/4 stays live
            int o = lim - 1;
            for (; o > o2; o--) step_generic(st, sb, o, lim, l, ta);
            for (; o >= 0; o--) {
                step_generic(st, sb, o, lim, l, ta);
                H.flo = (uint32_t)st.F; H.fhi = (uint32_t)(st.F >> 32); H.rlo = (uint32_t)st.R; H.rhi = (uint32_t)(st.R >> 32);
                MQ_CANDIDATE(o, live0)
            }
        } else if (n != 0) {
            int w = (lim - 1) >> 2;
            const int wm = (int)(n - 1) >> 2;
            uint32_t wa = sb + 128u * (uint32_t)w;
#ifndef MQ_WARM_TABLE
#define MQ_WARM_TABLE 1
#endif
            for (; w > wm; w--, wa -= 128u) {
#if MQ_WARM_TABLE
                warm_row(H, lds32(wa), a.warm);
#else
                const uint32_t x = lds32(wa);
                { const uint32_t o_ = x >> 24;           hash_step(H, lds64(ta + o_), lds64(ta + 128 + o_)); }
                { const uint32_t o_ = (x >> 16) & 0xFFu; hash_step(H, lds64(ta + o_), lds64(ta + 128 + o_)); }
                { const uint32_t o_ = (x >> 8) & 0xFFu;  hash_step(H, lds64(ta + o_), lds64(ta + 128 + o_)); }
                { const uint32_t o_ = x & 0xFFu;         hash_step(H, lds64(ta + o_), lds64(ta + 128 + o_)); }
#endif
            }
            const uint32_t ls = 8 * (l & 3);
            // software pipeline: the three stream words of the next iteration are loaded one iteration ahead,
            // and the table rows of an iteration are issued before its four dependent hash steps
            uint32_t inw = lds32(wa), ow0 = lds32(wa + lr), ow1 = lds32(wa + lr + 128);
            for (; w >= 0; w--) {
                const uint32_t comb = inw | (__funnelshift_r(ow0, ow1, ls) << 2);   // per byte: in + out * 4, pre-scaled (no N in this tile)
                const uint32_t o3 = prmt(comb, 0u, 0x4443u), o2b = prmt(comb, 0u, 0x4442u), o1b = prmt(comb, 0u, 0x4441u), o0b = prmt(comb, 0u, 0x4440u);
                const uint64_t f3 = lds64(ta + o3), r3 = lds64(ta + 128 + o3), f2 = lds64(ta + o2b), r2 = lds64(ta + 128 + o2b);
                const uint64_t f1 = lds64(ta + o1b), r1 = lds64(ta + 128 + o1b), f0 = lds64(ta + o0b), r0 = lds64(ta + 128 + o0b);
                if (w > 0) { wa -= 128; inw = lds32(wa); ow0 = lds32(wa + lr); ow1 = lds32(wa + lr + 128); }
                hash_step(H, f3, r3); MQ_CANDIDATE(4 * w + 3, wa)
                hash_step(H, f2, r2); MQ_CANDIDATE(4 * w + 2, wa)
                hash_step(H, f1, r1); MQ_CANDIDATE(4 * w + 1, wa)
                hash_step(H, f0, r0); MQ_CANDIDATE(4 * w, wa)
            }
        }
        if (cq != ctop - ck) nloc = flush_candidates(ws_a, cq + ck, nloc, c_lo, xlo, xlim, ev_a, tile, a, o2);
        __syncwarp();
        a.lane_cnt[(uint64_t)tile * 32 + lane] = (uint16_t)nloc;
        if (lane == 0) a.tile_cnt[tile] = lds32(ev_a);
        __syncwarp();
    }
}
#undef MQ_CANDIDATE

// ---- tile tables on the device (callers that only hold device-resident offsets) -----------------------------------
__global__ void k_tiles_per_seq(const uint64_t *offs, uint32_t n, uint32_t min_len, uint32_t *tiles) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tiles[i] = tiles_of_record(offs[i], offs[i + 1], min_len);
}

}  // namespace mq
