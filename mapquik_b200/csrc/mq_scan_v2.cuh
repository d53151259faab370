// mq_scan_v2.cuh -- S1 scan kernel, second generation.
//
// Same contract and outputs as k_scan_minimizers (mq_kernels.cuh): per-tile event pools with
// (lane, ordinal) tags, lane counts, tile totals.  What changed is the inside of a tile:
//
//   * every lane stages ITS OWN chunk (<= 256 raw bytes, sixteen 16-byte groups, LDG.128) and compacts
//     the homopolymer-run starts into a private byte stream in shared memory: one byte per SYMBOL
//     (code<<3 | nonACGT<<7), so the hot loop never sees a skipped base and run detection needs no
//     cross-lane traffic;
//   * the window's outgoing symbol is simply the same stream read l symbols further right, so the
//     64-bit shift register of v1 and its upkeep are gone; each lane's stream is followed by the l-1
//     context symbols of its right neighbours (or of the tile halo), then a zero byte;
//   * windows are warmed up by running the ordinary step over the context with a window full of
//     phantom 'A's (zero bytes) that cancel exactly when they drop out, so there is one step body;
//   * the hot loop handles four symbols per shared-memory word: in/out bytes are pre-scaled so that
//     `in | out<<2` is the byte offset into the 16-entry pair tables (one 128-byte bank row each);
//   * raw positions are recovered only for the ~1.5 % of l-mers that are selected, from per-group run
//     masks (find-n-th-set-bit), instead of being tracked per base.
#pragma once
#include <cstddef>
#include "mq_kernels.cuh"

namespace mq {

#ifndef MQ_V2_CS_MAX
#define MQ_V2_CS_MAX 128
#endif
#ifndef MQ_V2_WARPS
#define MQ_V2_WARPS 2
#endif
constexpr int V2_CS_MAX  = MQ_V2_CS_MAX;        // raw bytes per lane chunk (<= 256, multiple of 16)
constexpr int V2_TW_MAX  = 32 * V2_CS_MAX;      // raw bytes per warp tile
constexpr int V2_GPL_MAX = V2_CS_MAX / 16;      // 16-byte groups per lane
constexpr int V2_STRIDE  = V2_CS_MAX + 32 + 4;  // bytes per lane stream: symbols + context + zero; odd number of words
static_assert(((V2_STRIDE / 4) & 1) == 1 && V2_CS_MAX <= 256 && V2_CS_MAX % 16 == 0, "stream stride must be an odd word count");
constexpr int V2_WARPS   = MQ_V2_WARPS;         // warps (= tiles in flight) per CTA

struct ScanTablesV2 {                           // byte offsets are used directly by the kernel
    uint64_t pairF[16], pairR[16];              // @0, @128 : [in + 4*out]
    uint64_t inF[4], outF[4], inR[4], outR[4];  // @256, @288, @320, @352
    uint64_t F0, R0;                            // hash state of a window of l phantom 'A's
    uint32_t sel[16];                           // @400: PRMT selectors compacting the run-start bytes of a word
    uint32_t opq[4];                            // @464: 0x80000000, 2, bound_hi, shared address of this struct -- read back
                                                //       through volatile loads so they stay in registers (see v2_step)
};
static_assert(offsetof(ScanTablesV2, opq) == 464 && sizeof(ScanTablesV2) == 480, "the kernel addresses these fields by byte offset");

// tiles of record i on the 16-byte aligned grid
__global__ void k_tiles_per_seq_v2(const uint64_t *offs, uint32_t n, uint32_t min_len, uint32_t *tiles) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t gs = offs[i], ge = offs[i + 1], len = ge - gs;
    uint32_t t = 0;
    if (len >= min_len && len > 0) { uint64_t span = ge - (gs & ~15ull); t = (uint32_t)((span + V2_TW_MAX - 1) / V2_TW_MAX); }
    tiles[i] = t;
}

__device__ __forceinline__ void v2_geometry(uint64_t gs, uint64_t ge, uint32_t nt, uint32_t ti, uint32_t *Cs, uint64_t *tlo) {
    const uint64_t A = gs & ~15ull;
    const uint32_t span = (uint32_t)(ge - A), d = 32u * nt;     // records are < 2^31 bases: 32-bit division is exact
    uint32_t c = (span + d - 1) / d;
    c = (c + 15u) & ~15u;
    *Cs = c; *tlo = A + (uint64_t)ti * 32u * c;
}

struct V2Lane { uint64_t F, R; };

// How the two 64-bit rotate-by-one + xor updates of a hash step are issued.  The integer ALU pipe (LOP3/SHF/compare,
// one warp-instruction per two clocks per scheduler) is what bounds this kernel, while the FMA pipe (IMAD*) idles:
//   0  leave it to the compiler (64-bit shifts/ors)
//   1  funnel shifts (4 SHF + 4 LOP3, all ALU pipe)
//   2  both rotations through IMAD.WIDE: x * 2^31 = {x << 31, x >> 1}, x * 2 = {x << 1, x >> 31}; the halves are
//      recombined inside the xor (LOP3 (a|b)^c), so a step is 4 IMAD.WIDE (FMA pipe) + 4 LOP3 (ALU pipe)
//   3  forward strand through IMAD.WIDE, reverse strand through funnel shifts
//   4  reverse strand through IMAD.WIDE, forward strand through funnel shifts
#ifndef MQ_V2_ROT
#define MQ_V2_ROT 1
#endif
__device__ __forceinline__ void mulwide(uint32_t x, uint32_t k, uint32_t &lo, uint32_t &hi) {
    asm("{ .reg .b64 w; mul.wide.u32 w, %2, %3; mov.b64 {%0, %1}, w; }" : "=r"(lo), "=r"(hi) : "r"(x), "r"(k));
}
struct V2H { uint32_t flo, fhi, rlo, rhi; };
__device__ __forceinline__ void v2_step(V2H &h, uint64_t tf, uint64_t tr, uint32_t k31, uint32_t k2) {
    const uint32_t tfl = (uint32_t)tf, tfh = (uint32_t)(tf >> 32), trl = (uint32_t)tr, trh = (uint32_t)(tr >> 32);
    (void)k31; (void)k2; (void)tfl; (void)tfh; (void)trl; (void)trh;
#if MQ_V2_ROT == 0
    const uint64_t F = ror1(((uint64_t)h.fhi << 32) | h.flo) ^ tf, R = rol1(((uint64_t)h.rhi << 32) | h.rlo) ^ tr;
    h.flo = (uint32_t)F; h.fhi = (uint32_t)(F >> 32); h.rlo = (uint32_t)R; h.rhi = (uint32_t)(R >> 32);
#else
#if MQ_V2_ROT == 2 || MQ_V2_ROT == 3
    {   // F = ror1(F) ^ tf
        uint32_t l31, l1, h31, h1;
        mulwide(h.flo, k31, l31, l1); mulwide(h.fhi, k31, h31, h1);      // {x << 31, x >> 1}
        h.flo = (l1 | h31) ^ tfl; h.fhi = (h1 | l31) ^ tfh;
    }
#else
    { const uint32_t lo = h.flo, hi = h.fhi; h.flo = __funnelshift_r(lo, hi, 1) ^ tfl; h.fhi = __funnelshift_r(hi, lo, 1) ^ tfh; }
#endif
#if MQ_V2_ROT == 2 || MQ_V2_ROT == 4
    {   // R = rol1(R) ^ tr
        uint32_t ls, lc, hs, hc;
        mulwide(h.rlo, k2, ls, lc); mulwide(h.rhi, k2, hs, hc);          // {x << 1, x >> 31}
        h.rlo = (ls | hc) ^ trl; h.rhi = (hs | lc) ^ trh;
    }
#else
    { const uint32_t lo = h.rlo, hi = h.rhi; h.rlo = __funnelshift_l(hi, lo, 1) ^ trl; h.rhi = __funnelshift_l(lo, hi, 1) ^ trh; }
#endif
#endif
}

// v2 digest of one 4-byte word.  Symbol codes are the raw bits (c>>1)&3, i.e. A=0 C=1 T=2 G=3 (the tables are
// built in that order on the host), so no remap is needed.
//   symw : per byte  code<<3 | nonACGT<<7   (the byte that goes into the symbol stream)
//   run80: 0x80 in every byte that starts a homopolymer run (all bytes without HPC)
struct V2Dig { uint32_t symw, run80; };
// 0 in every byte of u that is 'A', 'C', 'G' or 'T':  bits 1..2 are the code; the other six bits must read 0x41, or
// 0x50 when the code is 2 ('T' = 0x54) -- m marks code-2 bytes, m*0x0F + 0x41.. is the expected pattern
__device__ __forceinline__ uint32_t v2_acgt_diff(uint32_t u) {
    const uint32_t m = (u >> 2) & ~(u >> 1) & 0x01010101u;
    return (u & 0xF9F9F9F9u) ^ (m * 0x0Fu + 0x41414141u);
}
__device__ __forceinline__ uint32_t v2_run80(uint32_t u, uint32_t pv, bool use_hpc) {
    if (!use_hpc) return 0x80808080u;
    const uint32_t e = u ^ pv;
    return (((e & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | e) & 0x80808080u;
}
__device__ __forceinline__ V2Dig v2_digest(uint32_t u, uint32_t pv, bool use_hpc) {
    const uint32_t diff = v2_acgt_diff(u);
    const uint32_t bad80 = (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;
    V2Dig d; d.symw = ((u << 2) & 0x18181818u) | bad80; d.run80 = v2_run80(u, pv, use_hpc);
    return d;
}
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
__device__ __forceinline__ void sts64(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// per-warp shared-memory layout (byte offsets from the warp's base)
constexpr int V2_CAND      = (V2_CS_MAX > 128 ? 12 : 8);                                  // candidates a lane can park before it must flush
constexpr int V2_OFF_NSYM  = 33 * V2_STRIDE;                       // u32[33]
constexpr int V2_OFF_RUNM  = V2_OFF_NSYM + 33 * 4;                 // u16[32][16]
constexpr int V2_OFF_CUM   = V2_OFF_RUNM + 32 * V2_GPL_MAX * 2;    // u8 [32][16] (always 16 per lane: vector compare)
constexpr int V2_OFF_CH    = (V2_OFF_CUM + 32 * 16 + 7) & ~7;   // u64[32][CAND+1]  candidate hashes
constexpr int V2_CH_STRIDE = (V2_CAND + 1) * 8;
constexpr int V2_OFF_CO    = V2_OFF_CH + 32 * V2_CH_STRIDE;        // u8 [32][16]       candidate ordinals
constexpr int V2_WARP_BYTES = (V2_OFF_CO + 32 * 16 + 15) & ~15;

// generic (N-aware, bounds-checked) step at ordinal o of the lane's logical stream
__device__ __forceinline__ void v2_step_generic(V2Lane &s, uint32_t sa, int o, int lim, uint32_t l, uint32_t ta) {
    const uint32_t in = lds8(sa + o);
    const int oo = o + (int)l;
    const uint32_t out = oo < lim ? lds8(sa + oo) : 0u;
    // single-symbol tables live behind the pair tables: inF @256, outF @288, inR @320, outR @352
    const uint64_t tf = ((in & 0x80u) ? 0ull : lds64(ta + 256 + (in & 0x18u))) ^ ((out & 0x80u) ? 0ull : lds64(ta + 288 + (out & 0x18u)));
    const uint64_t tr = ((in & 0x80u) ? 0ull : lds64(ta + 320 + (in & 0x18u))) ^ ((out & 0x80u) ? 0ull : lds64(ta + 352 + (out & 0x18u)));
    s.F = ror1(s.F) ^ tf; s.R = rol1(s.R) ^ tr;
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
// cum[] holds the symbol count before each 16-byte group; unused entries hold a sentinel no ordinal reaches.  With
// chunks of <= 128 bytes every value is < 0x80 (a chunk with fewer than 8 groups has <= 112 symbols), so "byte <= o"
// is one SWAR subtraction: bit 7 of (0x80|o) - byte.
constexpr uint32_t V2_CUM_FILL = V2_CS_MAX <= 128 ? 0x7F7F7F7Fu : 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t v2_raw_offset(uint32_t runm_a, uint32_t cum_a, uint32_t gpl, uint32_t o) {
    uint32_t g = 0;
    const uint32_t ob = o * 0x01010101u;
#pragma unroll
    for (int w = 0; w < V2_GPL_MAX / 4; w++) {
        const uint32_t cw = lds32(cum_a + 4 * w);
        if (V2_CS_MAX <= 128) g += __popc(((ob | 0x80808080u) - cw) & 0x80808080u);
        else g += __popc(__vcmpleu4(cw, ob) & 0x01010101u);
    }
    g -= 1;
    const uint32_t rm = lds16(runm_a + 2 * g), base = lds8(cum_a + g);
    (void)gpl;
    return 16u * g + select16(rm, o - base);
}

__device__ __forceinline__ void v2_emit(uint32_t x, uint64_t h, uint32_t lane, uint32_t j, uint32_t ev_a, uint32_t tile, const ScanArgs &a) {
    uint32_t slot;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(ev_a) : "memory");
    const uint32_t meta = x | (lane << 14) | (j << 19);
    if (slot < EV_CAP) {
        a.ev_hash[(uint64_t)tile * EV_CAP + slot] = h;
        a.ev_meta[(uint64_t)tile * EV_CAP + slot] = meta;
    } else {
        uint32_t g = atomicAdd(a.ovf_count, 1u);
        if (g < a.ovf_cap) { a.ovf_tile[g] = tile; a.ovf_meta[g] = meta; a.ovf_hash[g] = h; }
    }
}
// resolve and emit the candidates a lane has parked (ordinal -> raw position), oldest first
__device__ __noinline__ void v2_flush(uint32_t nc, uint32_t j0, uint32_t ch_a, uint32_t co_a, uint32_t runm_a, uint32_t cum_a, uint32_t gpl,
                                      uint32_t c_lo, uint32_t xlo, uint32_t xlim, uint32_t lane, uint32_t ev_a, uint32_t tile,
                                      const ScanArgs &a, uint64_t bound, uint32_t *nloc) {
    uint32_t j = j0;
    for (uint32_t i = 0; i < nc; i++) {
        const uint64_t h = lds64(ch_a + 8 * i);
        if (h >= bound) continue;                      // parked on the hi-word pre-filter only: exact test here
        const uint32_t o = lds8(co_a + i);
        const uint32_t x = c_lo + v2_raw_offset(runm_a, cum_a, gpl, o);
        if (x - xlo < xlim - xlo) { v2_emit(x, h, lane, j, ev_a, tile, a); j++; }
    }
    *nloc = j;
}

// Stage + compact one lane chunk: groups [g0, g1) lie completely inside the record (fast path, one PRMT compaction per
// word); the (at most two) groups cut by a record boundary go byte-wise; groups outside the record hold no symbol.
// FAST leaves the non-ACGT flag (bit 7) out of the symbol bytes and only reports whether the chunk holds ANY byte other
// than A/C/G/T (bad != 0) -- the caller then stages the tile again with FAST = false, which flags every symbol.
template <bool FAST>
__device__ __forceinline__ void v2_stage(const ScanArgs &a, uint64_t tlo, uint64_t gs, uint32_t c_lo, uint32_t gpl, uint32_t own_lo,
                                         uint32_t own_hi, uint32_t sa, uint32_t cum_a, uint32_t runm_a, uint32_t ta, bool hpc,
                                         uint32_t &n_out, uint32_t &bad_out) {
    const uint32_t Cs = gpl << 4;
    uint32_t n = 0, bad = 0;
    const uint8_t *cp = a.seqs + tlo + c_lo;
    uint32_t prev = 0;
    if (c_lo > own_lo && c_lo < own_hi) prev = cp[-1];  // byte before my chunk (same record)
    else if (c_lo == own_lo && tlo + own_lo > gs) prev = cp[-1];
    else if (c_lo == own_lo) prev = (uint32_t)cp[0] ^ 0xFFu;   // record starts exactly at my chunk: force a run start
    sts32(cum_a, V2_CUM_FILL); sts32(cum_a + 4, V2_CUM_FILL); sts32(cum_a + 8, V2_CUM_FILL); sts32(cum_a + 12, V2_CUM_FILL);
    const uint32_t lo_x = max(own_lo, c_lo), hi_x = min(own_hi, c_lo + Cs);
    uint32_t g0 = gpl, g1 = gpl;
    if (lo_x < hi_x) { g0 = (lo_x - c_lo + 15) >> 4; g1 = (hi_x - c_lo) >> 4; if (g1 < g0) g1 = g0; }
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (g0 < g1) nxt = __ldg((const uint4 *)(cp + 16 * g0));          // prefetch: one group ahead
    for (uint32_t g = 0; g < gpl; g++) {
        uint32_t rm = 0;
        sts8(cum_a + g, n);
        if (g >= g0 && g < g1) {
            const uint4 v = nxt;
            if (g + 1 < g1) nxt = __ldg((const uint4 *)(cp + 16 * (g + 1)));
            const uint32_t uw[4] = {v.x, v.y, v.z, v.w};
            // compact the run-start bytes of each word with one byte-permute (selector from a 16-entry
            // table), store all four bytes at the write cursor and advance the cursor only past the run
            // starts -- later stores overwrite the slack
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint32_t u = uw[w];
                const uint32_t run80 = v2_run80(u, (u << 8) | prev, hpc);
                prev = u >> 24;
                uint32_t symw;
                if (FAST) { symw = (u << 2) & 0x18181818u; bad |= v2_acgt_diff(u); }
                else { const V2Dig d = v2_digest(u, 0u, false); symw = d.symw; bad |= d.symw & run80; }
                const uint32_t p4 = ((run80 >> 7) * 0x04081020u) >> 24;            // 4 * (the four run bits)
                const uint32_t comp = prmt(symw, 0u, lds32(ta + 400 + p4));
                const uint32_t wa = sa + n;
                sts8(wa, comp); sts8(wa + 1, comp >> 8); sts8(wa + 2, comp >> 16); sts8(wa + 3, comp >> 24);
                n += __popc(p4);
                rm |= w ? (p4 << (4 * w - 2)) : (p4 >> 2);
            }
        } else {
            const uint32_t xg = c_lo + 16 * g;
            if (xg < own_hi && xg + 16 > own_lo) {       // cut by a record boundary: byte-wise
                const uint4 v = __ldg((const uint4 *)(cp + 16 * g));
                const uint32_t uw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const uint32_t u = uw[w];
                    const V2Dig d = v2_digest(u, (u << 8) | prev, hpc);
                    prev = u >> 24;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const uint32_t x = xg + 4 * w + b;
                        if (x < own_lo || x >= own_hi) continue;
                        const bool start = (x == own_lo && tlo + own_lo == gs) || ((d.run80 >> (8 * b)) & 0x80u);
                        if (!start) continue;
                        const uint32_t sb = (d.symw >> (8 * b)) & 0xFFu;
                        bad |= sb & 0x80u;
                        sts8(sa + n, FAST ? (sb & 0x7Fu) : sb); n++;
                        rm |= 1u << (4 * w + b);
                    }
                }
            }
        }
        sts16(runm_a + 2 * g, rm);
    }
    n_out = n; bad_out = bad;
}

#define V2_CANDIDATE(ORD)                                                                                     \
    if (min(H.fhi, H.rhi) <= bound_hi) {                                                                      \
        const uint64_t F_ = ((uint64_t)H.fhi << 32) | H.flo, R_ = ((uint64_t)H.rhi << 32) | H.rlo;            \
        sts64(ch_a + 8 * nc, F_ < R_ ? F_ : R_); sts8(co_a + nc, (uint32_t)(ORD)); nc++;                      \
        if (nc == V2_CAND) { v2_flush(nc, nloc, ch_a, co_a, runm_a, cum_a, gpl, c_lo, xlo, xlim, lane, ev_a, tile, a, bound, &nloc); nc = 0; } \
    }

__global__ void __launch_bounds__(V2_WARPS * 32) k_scan_minimizers_v2(ScanArgs a, ScanTablesV2 Tin) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(128) ScanTablesV2 T;
    __shared__ uint32_t ev_cnt[V2_WARPS];
    for (uint32_t i = threadIdx.x; i < sizeof(ScanTablesV2) / 4; i += blockDim.x) ((uint32_t *)&T)[i] = ((const uint32_t *)&Tin)[i];
    __syncthreads();
    if (threadIdx.x == 0) { T.opq[0] = 0x80000000u; T.opq[1] = 2u; T.opq[2] = (uint32_t)(a.bound >> 32); T.opq[3] = smem_addr(&T); }
    __syncthreads();
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    uint8_t *WS = smem_raw + (size_t)wid * V2_WARP_BYTES;
    const uint32_t ws_a = smem_addr(WS);
    const uint32_t sa = ws_a + lane * V2_STRIDE;                     // my stream
    const uint32_t nsym_a = ws_a + V2_OFF_NSYM;
    const uint32_t runm_a = ws_a + V2_OFF_RUNM + lane * V2_GPL_MAX * 2;
    const uint32_t cum_a = ws_a + V2_OFF_CUM + lane * 16;
    const uint32_t ch_a = ws_a + V2_OFF_CH + lane * V2_CH_STRIDE;
    const uint32_t co_a = ws_a + V2_OFF_CO + lane * 16;
    // pairF @0, pairR @128, single-symbol tables @256..; these four scalars come back through volatile shared loads so
    // that ptxas keeps them in registers instead of re-deriving them (S2UR/ULEA/LDCU) inside the hot loop
    const uint32_t ta = lds32(smem_addr(&T) + 464 + 12);
    const uint32_t k31 = lds32(ta + 464), k2 = lds32(ta + 464 + 4);
    const uint32_t bound_hi = lds32(ta + 464 + 8);
    const uint32_t ev_a = smem_addr(&ev_cnt[wid]);
    const uint32_t l = a.l;
    const bool hpc = a.use_hpc != 0;
    const uint64_t bound = a.bound;

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
        v2_geometry(gs, ge, nt, ti, &Cs, &tlo);
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
        uint32_t n, bad;
        v2_stage<true>(a, tlo, gs, c_lo, gpl, own_lo, own_hi, sa, cum_a, runm_a, ta, hpc, n, bad);
        if (__any_sync(0xffffffffu, bad != 0)) {               // some byte is not A/C/G/T: stage again with per-symbol flags
            __syncwarp();
            v2_stage<false>(a, tlo, gs, c_lo, gpl, own_lo, own_hi, sa, cum_a, runm_a, ta, hpc, n, bad);
        }
        sts32(nsym_a + 4 * lane, n);
        if (!__any_sync(0xffffffffu, n != 0)) {
            // the whole tile lies inside one homopolymer run (or outside the record): no l-mer starts here,
            // so neither halo nor context is needed -- this keeps giant runs (N-gaps) linear instead of quadratic
            a.lane_cnt[(uint64_t)tile * 32 + lane] = 0;
            if (lane == 0) a.tile_cnt[tile] = 0;
            __syncwarp();
            continue;
        }

        // ---- halo stream (slot 32): up to 32 (>= l-1) run-start symbols right of the tile -------------------
        uint32_t hcount = 0;
        if (tlo + TWs < ge) {
            const uint32_t ha = ws_a + 32 * V2_STRIDE;
            uint64_t haddr = tlo + TWs;
            uint32_t hcarry = a.seqs[haddr - 1];
            while (hcount < 32u && haddr < ge) {
                const uint64_t wa = haddr + 4ull * lane;
                uint32_t u = (wa < ge) ? __ldg((const uint32_t *)(a.seqs + wa)) : 0u;
                uint32_t up = __shfl_up_sync(0xffffffffu, u, 1);
                uint32_t prevb = lane == 0 ? hcarry : (up >> 24);
                hcarry = __shfl_sync(0xffffffffu, u, 31) >> 24;
                const V2Dig d = v2_digest(u, (u << 8) | prevb, hpc);
                uint32_t m = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) if (wa + b < ge) m |= 0x80u << (8 * b);
                const uint32_t run = d.run80 & m;
                uint32_t mine = __popc(run), tot;
                uint32_t r = hcount + warp_excl_scan(mine, &tot);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (run & (0x80u << (8 * b))) {
                        const uint32_t sb = (d.symw >> (8 * b)) & 0xFFu;
                        if (r < 32u) { sts8(ha + r, sb); if (r < l - 1) bad |= sb & 0x80u; }
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

        // ---- context: the next l-1 symbols after my chunk, from the streams to my right -----------------
        uint32_t c = 0;
        if (lds32(nsym_a + 4 * (lane + 1)) >= 32u) {
            // common case: the stream to my right alone holds the l-1 (<= 31) symbols -- copy eight words
            const uint32_t sj = ws_a + (lane + 1) * V2_STRIDE, wa = sa + n;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t x = lds32(sj + 4 * i);
                sts8(wa + 4 * i, x); sts8(wa + 4 * i + 1, x >> 8); sts8(wa + 4 * i + 2, x >> 16); sts8(wa + 4 * i + 3, x >> 24);
            }
            c = l - 1;
        } else {
            for (uint32_t j = lane + 1; j <= 32 && c < l - 1; j++) {
                const uint32_t nj = lds32(nsym_a + 4 * j);
                const uint32_t sj = ws_a + j * V2_STRIDE;
                for (uint32_t i = 0; i < nj && c < l - 1; i++, c++) sts8(sa + n + c, lds8(sj + i));
            }
        }
        sts8(sa + n + c, 0); sts8(sa + n + c + 1, 0); sts8(sa + n + c + 2, 0); sts8(sa + n + c + 3, 0);
        __syncwarp();

        // ---- phase 1: warm-up over context and record-final symbols, no emission ----------------------
        V2Lane st; st.F = T.F0; st.R = T.R0;
        const int lim = (int)(n + c);                           // symbols available in my logical stream
        int o = lim - 1;
        const int o2 = max(-1, min((int)n - 1, lim - (int)l));  // first ordinal whose window is complete and mine
        V2H H;
        if (anyN) {
            for (; o > o2; o--) v2_step_generic(st, sa, o, lim, l, ta);
            H.flo = (uint32_t)st.F; H.fhi = (uint32_t)(st.F >> 32); H.rlo = (uint32_t)st.R; H.rhi = (uint32_t)(st.R >> 32);
        } else {
            // o > o2 == lim-l (or -1): the outgoing ordinal o+l lies beyond the stream, i.e. it is a phantom
            // 'A' (code 0) -- the pair-table row for out == 0 is the whole step
            H.flo = (uint32_t)st.F; H.fhi = (uint32_t)(st.F >> 32); H.rlo = (uint32_t)st.R; H.rhi = (uint32_t)(st.R >> 32);
            for (; o > o2; o--) {
                const uint32_t off = lds8(sa + o);
                v2_step(H, lds64(ta + off), lds64(ta + 128 + off), k31, k2);
            }
        }

        // ---- phase 2: scan of my own symbols; selected l-mers are parked, positions resolved after --------
        uint32_t nc = 0;
        if (anyN) {
            for (; o >= 0; o--) {
                st.F = ((uint64_t)H.fhi << 32) | H.flo; st.R = ((uint64_t)H.rhi << 32) | H.rlo;
                v2_step_generic(st, sa, o, lim, l, ta);
                H.flo = (uint32_t)st.F; H.fhi = (uint32_t)(st.F >> 32); H.rlo = (uint32_t)st.R; H.rhi = (uint32_t)(st.R >> 32);
                V2_CANDIDATE(o)
            }
        } else {
            for (; o >= 0 && ((o + 1) & 3); o--) {                // bring o+1 to a multiple of 4
                const uint32_t off = lds8(sa + o) | (lds8(sa + o + (int)l) << 2);
                v2_step(H, lds64(ta + off), lds64(ta + 128 + off), k31, k2);
                V2_CANDIDATE(o)
            }
            const uint32_t lw4 = (l >> 2) * 4, ls = 8 * (l & 3);
            int w = ((o + 1) >> 2) - 1;
            // software pipeline: the three stream words of the next iteration are loaded one iteration ahead,
            // and the eight table loads of an iteration are issued before its four dependent hash steps
            uint32_t inw = 0, ow0 = 0, ow1 = 0;
            if (w >= 0) { inw = lds32(sa + 4 * w); ow0 = lds32(sa + 4 * w + lw4); ow1 = lds32(sa + 4 * w + lw4 + 4); }
            for (; w >= 0; w--) {
                const uint32_t comb = inw | (__funnelshift_r(ow0, ow1, ls) << 2);   // per byte: in*8 + out*32 (no N in this tile)
                const uint32_t o3 = comb >> 24, o2b = (comb >> 16) & 0xFFu, o1b = (comb >> 8) & 0xFFu, o0b = comb & 0xFFu;
                const uint64_t tf3 = lds64(ta + o3), tr3 = lds64(ta + 128 + o3);
                const uint64_t tf2 = lds64(ta + o2b), tr2 = lds64(ta + 128 + o2b);
                const uint64_t tf1 = lds64(ta + o1b), tr1 = lds64(ta + 128 + o1b);
                const uint64_t tf0 = lds64(ta + o0b), tr0 = lds64(ta + 128 + o0b);
                if (w > 0) { inw = lds32(sa + 4 * w - 4); ow0 = lds32(sa + 4 * w - 4 + lw4); ow1 = lds32(sa + 4 * w + lw4); }
                v2_step(H, tf3, tr3, k31, k2); V2_CANDIDATE(4 * w + 3)
                v2_step(H, tf2, tr2, k31, k2); V2_CANDIDATE(4 * w + 2)
                v2_step(H, tf1, tr1, k31, k2); V2_CANDIDATE(4 * w + 1)
                v2_step(H, tf0, tr0, k31, k2); V2_CANDIDATE(4 * w)
            }
        }
        if (nc) v2_flush(nc, nloc, ch_a, co_a, runm_a, cum_a, gpl, c_lo, xlo, xlim, lane, ev_a, tile, a, bound, &nloc);
        __syncwarp();
        a.lane_cnt[(uint64_t)tile * 32 + lane] = (uint16_t)nloc;
        if (lane == 0) a.tile_cnt[tile] = lds32(ev_a);
        __syncwarp();
    }
}
#undef V2_CANDIDATE

}  // namespace mq
