// mq_scan_v3.cuh -- S1 scan kernel, third generation: the v2 algorithm on a bank-conflict-free layout.
//
// ncu on v2 (profiles/r01_scan_v2_final_*): the shared-memory data pipe was the busiest unit of the SM (85 % of
// its wavefront slots) -- ahead of the ALU pipe (75 %) and the issue slots (73 %) -- and half of those wavefronts
// were bank conflicts: every lane kept its symbol stream contiguous (stride 164 bytes) and read / wrote it at its own
// cursor, so the 32 lanes of one LDS/STS hit pseudo-random banks (2.3-2.6 wavefronts per instruction), and
// compaction used four byte stores per input word.  Here
//   * the streams are LANE-INTERLEAVED: word w of lane L lives at row w, column L (byte (w*32 + L)*4), so whatever
//     word index each lane is at, lane L only ever touches bank L: one wavefront per LDS/STS;
//   * compaction appends to a pending register and stores one aligned word when it fills (<= 1 STS per input word);
//   * the context (the next l-1 symbols, taken from the right neighbour's first eight words) is appended with the
//     same funnel, word-wise;
//   * the per-group run masks and symbol counts are row-interleaved as well.
//   * the pair tables are interleaved ({F, R} per (in, out) pair: one LDS.128 per hash step), symbol bytes are
//     pre-scaled by 16 accordingly, and the warm-up runs on the context words while they are still in registers.
// Everything else (tiles, candidate parking, event pools, outputs) is v2's; see mq_scan_v2.cuh.
#pragma once
#include <cstring>
#include "mq_scan_v2.cuh"

namespace mq {

// MQ_V3_LDS128 = 1: the pair tables are interleaved ({F, R} per (in, out) pair, 16 bytes) and a hash step needs one
// LDS.128; symbol bytes are then code << 4.  = 0: v2's two 128-byte tables (two LDS.64 per step, code << 3), which
// never bank-conflict because each table is exactly one bank row.
#ifndef MQ_V3_LDS128
#define MQ_V3_LDS128 0
#endif
constexpr int V3_SH = MQ_V3_LDS128 ? 4 : 3;     // symbol byte = code << V3_SH
struct ScanTablesV3 {                           // same size and scalar offsets as ScanTablesV2
    uint64_t pair[32];                          // @0  : {F, R} for [in + 4*out], 16 bytes per entry (or pairF[16], pairR[16])
    uint64_t inF[4], outF[4], inR[4], outR[4];  // @256, @288, @320, @352
    uint64_t F0, R0;
    uint32_t sel[16];                           // @400
    uint32_t opq[8];                            // @464: 2^31, 2, bound_hi, &T, 15 << 30, 1, 4, -- (see MQ_V3_FMA)
};
static_assert(sizeof(ScanTablesV3) == sizeof(ScanTablesV2) + 16 && offsetof(ScanTablesV3, sel) == 400 && offsetof(ScanTablesV3, opq) == 464, "layout");
inline void fill_tables_v3(ScanTablesV3 &T, const ScanTablesV2 &S) {
    memset(&T, 0, sizeof(T));
    memcpy(&T, &S, sizeof(S));
    if (MQ_V3_LDS128) for (int i = 0; i < 16; i++) { T.pair[2 * i] = S.pairF[i]; T.pair[2 * i + 1] = S.pairR[i]; }
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
// table row for the pre-scaled (in, out) byte `off`
__device__ __forceinline__ uint4 v3_tab(uint32_t ta, uint32_t off) {
    if (MQ_V3_LDS128) return lds128(ta + off);
    const uint64_t f = lds64(ta + off), r = lds64(ta + 128 + off);
    return make_uint4((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)r, (uint32_t)(r >> 32));
}
__device__ __forceinline__ void v3_step(V2H &h, uint4 t, uint32_t k31, uint32_t k2) {
    v2_step(h, ((uint64_t)t.y << 32) | t.x, ((uint64_t)t.w << 32) | t.z, k31, k2);
}
// MQ_V3_PROBE_ALU / MQ_V3_PROBE_FMA (default 0): N extra LOP3 / IMAD instructions per hash step of the hot loop, on an
// accumulator of their own.  Only for the "which pipe binds" experiment in profiles/README.md: the slope of kernel time
// against N tells whether ALU-pipe slots, FMA-pipe slots or plain issue slots are the scarce resource.
#ifndef MQ_V3_PROBE_ALU
#define MQ_V3_PROBE_ALU 0
#endif
#ifndef MQ_V3_PROBE_FMA
#define MQ_V3_PROBE_FMA 0
#endif
#if MQ_V3_PROBE_ALU || MQ_V3_PROBE_FMA
#define V3_PROBE(T4)                                                                                          \
    { _Pragma("unroll") for (int i_ = 0; i_ < MQ_V3_PROBE_ALU; i_++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(probe_acc) : "r"((T4).y), "r"((T4).z)); \
      _Pragma("unroll") for (int i_ = 0; i_ < MQ_V3_PROBE_FMA; i_++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(probe_acc) : "r"((T4).y), "r"((T4).z)); }
#else
#define V3_PROBE(T4)
#endif
// MQ_V3_FMA = 1: additions / small multiplies whose result the integer ALU pipe would otherwise produce are issued as
// IMAD / IMAD.HI with multipliers read from shared memory (opaque to ptxas, which would strength-reduce literal
// multipliers back into shifts and adds): the ALU pipe is this kernel's bound, the FMA pipe is mostly idle.
#ifndef MQ_V3_FMA
#define MQ_V3_FMA 0
#endif
struct V3K { uint32_t c0, one, four; };          // 15 << 30, 1, 4
__device__ __forceinline__ uint32_t fma_add(uint32_t x, uint32_t c, const V3K &k) {           // x + c
#if MQ_V3_FMA
    uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(k.one), "r"(c)); return r;
#else
    (void)k; return x + c;
#endif
}
// 0 in every byte of u that is 'A', 'C', 'G' or 'T' (see v2_acgt_diff): m4 marks code-2 bytes at bit 2, and
// hi32(m4 * (15 << 30)) = (m4 * 15) >> 2 = 0x0F in those bytes, so the expected pattern is one IMAD.HI
__device__ __forceinline__ uint32_t v3_acgt_diff(uint32_t u, const V3K &k) {
#if MQ_V3_FMA
    const uint32_t m4 = u & ~(u << 1) & 0x04040404u;
    uint32_t e; asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(e) : "r"(m4), "r"(k.c0), "r"(0x41414141u));
    return (u & 0xF9F9F9F9u) ^ e;
#else
    (void)k; return v2_acgt_diff(u);
#endif
}
__device__ __forceinline__ uint32_t v3_run80(uint32_t u, uint32_t pv, bool use_hpc, const V3K &k) {
    if (!use_hpc) return 0x80808080u;
    const uint32_t e = u ^ pv;
    return (fma_add(e & 0x7F7F7F7Fu, 0x7F7F7F7Fu, k) | e) & 0x80808080u;
}
// symbol byte: code << V3_SH (code: A=0 C=1 T=2 G=3), bit 7 = not A/C/G/T
__device__ __forceinline__ uint32_t v3_symw(uint32_t u) { return (u << (V3_SH - 1)) & (0x03030303u << V3_SH); }
__device__ __forceinline__ uint32_t v3_bad80(uint32_t u) {
    const uint32_t diff = v2_acgt_diff(u);
    return (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;
}

constexpr int V3_ROWS     = V2_STRIDE / 4 + 2;                          // words per lane column: symbols + context + zero, + 2 spare rows
constexpr int V3_OFF_HALO = V3_ROWS * 128;                              // u8[48]  halo stream (contiguous; only lane 31 reads it)
constexpr int V3_OFF_NSYM = V3_OFF_HALO + 48;                           // u32[33] symbols per stream (32 = halo)
constexpr int V3_OFF_RUNM = V3_OFF_NSYM + 144;                          // u16 run masks: word j (groups 2j, 2j+1) of lane L at row j
constexpr int V3_OFF_CUM  = V3_OFF_RUNM + (V2_GPL_MAX / 2) * 128;       // u8 symbol counts before each group, 4 per word, row-interleaved
constexpr int V3_WARP_BYTES = (V3_OFF_CUM + (V2_GPL_MAX / 4) * 128 + 15) & ~15;
// Parked candidates live in the lane's OWN column, from the top row downwards, three rows each (hash lo, hash hi,
// ordinal): the scan walks its column from the top down, so the rows above the outgoing-symbol words are dead by the
// time candidates appear, and the list costs no shared memory of its own (6.5 KB per warp -> 32 warps per SM).
constexpr int V3_CAND_TOP = (V3_ROWS - 1) * 128;

// byte o of the stream whose column base is sb
__device__ __forceinline__ uint32_t v3_baddr(uint32_t sb, uint32_t o) { return sb + o + (o >> 2) * 124u; }

// append cursor of a lane stream: P holds the bytes of the incomplete word (zero above them), n8 = 8 * symbols so far,
// wp = address of the incomplete word
struct V3Pend { uint32_t P, n8, wp; };
#ifndef MQ_V3_PUSH_MAD
#define MQ_V3_PUSH_MAD 0
#endif
__device__ __forceinline__ void v3_push(V3Pend &q, uint32_t comp, uint32_t c8) {       // comp: c8/8 bytes, zero above
    const uint32_t f8 = q.n8 & 24u;
#if MQ_V3_PUSH_MAD
    uint32_t lo, hi;                                                                    // (hi:lo) = comp * 2^f8 + P
    asm("{ .reg .b64 w, p; mov.b64 p, {%3, %4}; mad.wide.u32 w, %2, %5, p; mov.b64 {%0, %1}, w; }"
        : "=r"(lo), "=r"(hi) : "r"(comp), "r"(q.P), "r"(0u), "r"(1u << f8));
#else
    (void)f8;                                                                           // the funnel shifter takes n8 mod 32
    const uint32_t lo = q.P | __funnelshift_l(0u, comp, q.n8), hi = __funnelshift_l(comp, 0u, q.n8);   // (hi:lo) = comp << f8 | P
#endif
    const uint32_t n8n = q.n8 + c8;
    if ((n8n ^ q.n8) & 32u) { sts32(q.wp, lo); q.wp += 128u; q.P = hi; } else q.P = lo;
    q.n8 = n8n;
}

__device__ __forceinline__ uint32_t v3_raw_offset(uint32_t runm_l, uint32_t cum_l, uint32_t o) {
    uint32_t g = 0;
    const uint32_t ob = o * 0x01010101u;
#pragma unroll
    for (int w = 0; w < V2_GPL_MAX / 4; w++) {
        const uint32_t cw = lds32(cum_l + 128 * w);
        if (V2_CS_MAX <= 128) g += __popc(((ob | 0x80808080u) - cw) & 0x80808080u);
        else g += __popc(__vcmpleu4(cw, ob) & 0x01010101u);
    }
    g -= 1;
    const uint32_t rm = lds16(runm_l + (g >> 1) * 128u + (g & 1u) * 2u), base = lds8(cum_l + (g >> 2) * 128u + (g & 3u));
    return 16u * g + select16(rm, o - base);
}

// resolve and emit the candidates parked in rows (cpz, top] of the lane's column, oldest first
__device__ __noinline__ uint32_t v3_flush(uint32_t top, uint32_t cpz, uint32_t j0, uint32_t runm_l, uint32_t cum_l,
                                          uint32_t c_lo, uint32_t xlo, uint32_t xlim, uint32_t lane, uint32_t ev_a, uint32_t tile,
                                          const ScanArgs &a, uint64_t bound, int o2) {
    uint32_t j = j0;
    for (uint32_t p = top; p > cpz; p -= 384u) {
        const uint64_t h = ((uint64_t)lds32(p - 128u) << 32) | lds32(p);
        if (h >= bound) continue;                      // parked on the hi-word pre-filter only: exact test here
        const uint32_t o = lds32(p - 256u);
        if ((int)o > o2) continue;                     // a context symbol, or a window the record end leaves incomplete
        const uint32_t x = c_lo + v3_raw_offset(runm_l, cum_l, o);
        if (x - xlo < xlim - xlo) { v2_emit(x, h, lane, j, ev_a, tile, a); j++; }
    }
    return j;
}

// v2_stage on the interleaved layout; leaves the append cursor in q (pending word NOT yet stored)
template <bool FAST>
__device__ __forceinline__ void v3_stage(const ScanArgs &a, uint64_t tlo, uint64_t gs, uint32_t c_lo, uint32_t gpl, uint32_t own_lo,
                                         uint32_t own_hi, uint32_t sb, uint32_t cum_l, uint32_t runm_l, uint32_t ta, bool hpc,
                                         const V3K &kk, V3Pend &q, uint32_t &bad_out) {
    const uint32_t Cs = gpl << 4;
    uint32_t bad = 0;
    q.P = 0; q.n8 = 0; q.wp = sb;
    const uint8_t *cp = a.seqs + tlo + c_lo;
    uint32_t prev = 0;
    if (c_lo > own_lo && c_lo < own_hi) prev = cp[-1];  // byte before my chunk (same record)
    else if (c_lo == own_lo && tlo + own_lo > gs) prev = cp[-1];
    else if (c_lo == own_lo) prev = (uint32_t)cp[0] ^ 0xFFu;   // record starts exactly at my chunk: force a run start
#pragma unroll
    for (int r = 0; r < V2_GPL_MAX / 4; r++) sts32(cum_l + 128 * r, V2_CUM_FILL);
    const uint32_t lo_x = max(own_lo, c_lo), hi_x = min(own_hi, c_lo + Cs);
    uint32_t g0 = gpl, g1 = gpl;
    if (lo_x < hi_x) { g0 = (lo_x - c_lo + 15) >> 4; g1 = (hi_x - c_lo) >> 4; if (g1 < g0) g1 = g0; }
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (g0 < g1) nxt = __ldg((const uint4 *)(cp + 16 * g0));          // prefetch: one group ahead
    uint32_t ca = cum_l, ra = runm_l;
    for (uint32_t g = 0; g < gpl; g++) {
        uint32_t rm = 0;
        sts8(ca, q.n8 >> 3);
        if (g >= g0 && g < g1) {
            const uint4 v = nxt;
            if (g + 1 < g1) nxt = __ldg((const uint4 *)(cp + 16 * (g + 1)));
            const uint32_t uw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint32_t u = uw[w];
                const uint32_t run80 = v3_run80(u, (u << 8) | prev, hpc, kk);
                prev = u >> 24;
                uint32_t symw = v3_symw(u);
                if (FAST) bad |= v3_acgt_diff(u, kk);
                else { symw |= v3_bad80(u); bad |= symw & run80; }
                // run80 has bit 7 of byte i set for a run start: * 0x00204081 moves them to bits 28..31 (no two partial
                // products meet, so nothing carries)
                const uint32_t p = (run80 * 0x00204081u) >> 28;                    // the four run bits
                const uint32_t comp = prmt(symw, 0u, lds32(p * 4u + (ta + 400)));   // run-start bytes first, zero fill
                v3_push(q, comp, __popc(p) << 3);
                rm |= p << (4 * w);
            }
        } else {
            const uint32_t xg = c_lo + 16 * g;
            if (xg < own_hi && xg + 16 > own_lo) {       // cut by a record boundary: same word-wise path, run bits masked
                const uint4 v = __ldg((const uint4 *)(cp + 16 * g));
                const uint32_t uw[4] = {v.x, v.y, v.z, v.w};
                const bool rec_start = tlo + own_lo == gs;
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const uint32_t u = uw[w], x0 = xg + 4 * w;
                    const uint32_t run80 = v2_run80(u, (u << 8) | prev, hpc);
                    prev = u >> 24;
                    const uint32_t lo_b = own_lo > x0 ? min(own_lo - x0, 4u) : 0u, hi_b = own_hi > x0 ? min(own_hi - x0, 4u) : 0u;
                    uint32_t p = (((run80 >> 7) * 0x01020408u) >> 24) & ((1u << hi_b) - 1u) & ~((1u << lo_b) - 1u);
                    if (rec_start && own_lo >= x0 && own_lo < x0 + 4u && own_lo < own_hi) p |= 1u << (own_lo - x0);   // a record starts a run
                    const uint32_t bad80 = v3_bad80(u);
                    bad |= bad80 & (((p * 0x00204081u) & 0x01010101u) << 7);      // flags of the selected bytes only
                    const uint32_t symw = v3_symw(u) | (FAST ? 0u : bad80);
                    const uint32_t comp = prmt(symw, 0u, lds32(ta + 400 + 4 * p));
                    v3_push(q, comp, __popc(p) << 3);
                    rm |= p << (4 * w);
                }
            }
        }
        sts16(ra, rm);
        ca += ((g & 3u) == 3u) ? 125u : 1u;
        ra += (g & 1u) ? 126u : 2u;
    }
    bad_out = bad;
}

// generic (N-aware, bounds-checked) step at ordinal o of the lane's logical stream
__device__ __forceinline__ void v3_step_generic(V2Lane &s, uint32_t sb, int o, int lim, uint32_t l, uint32_t ta) {
    const uint32_t in = lds8(v3_baddr(sb, (uint32_t)o));
    const int oo = o + (int)l;
    const uint32_t out = oo < lim ? lds8(v3_baddr(sb, (uint32_t)oo)) : 0u;
    const uint32_t io = ((in >> V3_SH) & 3u) * 8u, oo8 = ((out >> V3_SH) & 3u) * 8u;
    const uint64_t tf = ((in & 0x80u) ? 0ull : lds64(ta + 256 + io)) ^ ((out & 0x80u) ? 0ull : lds64(ta + 288 + oo8));
    const uint64_t tr = ((in & 0x80u) ? 0ull : lds64(ta + 320 + io)) ^ ((out & 0x80u) ? 0ull : lds64(ta + 352 + oo8));
    s.F = ror1(s.F) ^ tf; s.R = rol1(s.R) ^ tr;
}

// park: three stores down the column; flush when the next candidate would reach the rows the scan still reads.
// cq is the cursor minus ck = (l/4)*128 + 384, so that inside the word loop the test is simply cq <= wa (rows up to
// wa + (l/4)*128 + 128 are still read, and a candidate needs three rows).
#define V3_CANDIDATE(ORD, LIVE)                                                                               \
    if (min(H.fhi, H.rhi) <= bound_hi) {                                                                      \
        const bool fmin = (((uint64_t)H.fhi << 32) | H.flo) < (((uint64_t)H.rhi << 32) | H.rlo);              \
        const uint32_t cpz = fma_add(cq, ck, kk);                                                             \
        sts32(cpz, fmin ? H.flo : H.rlo); sts32(cpz - 128u, min(H.fhi, H.rhi)); sts32(cpz - 256u, (uint32_t)(ORD)); \
        cq = fma_add(cq, 0u - 384u, kk);                                                                      \
        if (cq <= (LIVE)) { nloc = v3_flush(ctop, cq + ck, nloc, runm_l, cum_l, c_lo, xlo, xlim, lane, ev_a, tile, a, bound, o2); cq = ctop - ck; } \
    }

// HPC is a compile-time flag (two instantiations): the run-start digest then has no per-word branch
template <bool HPC>
__global__ void __launch_bounds__(V2_WARPS * 32, 32 / V2_WARPS) k_scan_minimizers_v3(const __grid_constant__ ScanArgs a, const __grid_constant__ ScanTablesV3 Tin) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(256) ScanTablesV3 T;
    __shared__ uint32_t ev_cnt[V2_WARPS];
    for (uint32_t i = threadIdx.x; i < sizeof(ScanTablesV3) / 4; i += blockDim.x) ((uint32_t *)&T)[i] = ((const uint32_t *)&Tin)[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        T.opq[0] = 0x80000000u; T.opq[1] = 2u; T.opq[2] = (uint32_t)(a.bound >> 32); T.opq[3] = smem_addr(&T);
        T.opq[4] = 15u << 30; T.opq[5] = 1u; T.opq[6] = 4u;
    }
    __syncthreads();
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    const uint32_t ws_a = smem_addr(smem_raw + (size_t)wid * V3_WARP_BYTES);
    const uint32_t sb = ws_a + 4 * lane;                              // my stream: column `lane` of the rows
    const uint32_t ha = ws_a + V3_OFF_HALO;
    const uint32_t nsym_a = ws_a + V3_OFF_NSYM;
    const uint32_t runm_l = ws_a + V3_OFF_RUNM + 4 * lane;
    const uint32_t cum_l = ws_a + V3_OFF_CUM + 4 * lane;
    const uint32_t ctop = sb + V3_CAND_TOP;                           // first candidate slot: the top row of my column
    // right neighbour's stream: column lane+1, or the contiguous halo for lane 31
    const uint32_t nb_a = lane < 31 ? sb + 4 : ha, nb_st = lane < 31 ? 128u : 4u;
    // scalars read back through volatile shared loads so that ptxas keeps them in registers (see mq_scan_v2.cuh)
    const uint32_t ta = lds32(smem_addr(&T) + 464 + 12);
    const uint32_t k31 = lds32(ta + 464), k2 = lds32(ta + 464 + 4);
    V3K kk; kk.c0 = lds32(ta + 464 + 16); kk.one = lds32(ta + 464 + 20); kk.four = lds32(ta + 464 + 24);
    const uint32_t bound_hi = lds32(ta + 464 + 8);
    const uint32_t ev_a = smem_addr(&ev_cnt[wid]);
    const uint32_t l = a.l;
    constexpr bool hpc = HPC;
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
        V3Pend q; uint32_t bad;
        v3_stage<true>(a, tlo, gs, c_lo, gpl, own_lo, own_hi, sb, cum_l, runm_l, ta, hpc, kk, q, bad);
        if (__any_sync(0xffffffffu, bad != 0)) {               // some byte is not A/C/G/T: stage again with per-symbol flags
            __syncwarp();
            v3_stage<false>(a, tlo, gs, c_lo, gpl, own_lo, own_hi, sb, cum_l, runm_l, ta, hpc, kk, q, bad);
        }
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
            uint32_t hcarry = a.seqs[haddr - 1];
            while (hcount < 32u && haddr < ge) {
                const uint64_t wa = haddr + 4ull * lane;
                uint32_t u = (wa < ge) ? __ldg((const uint32_t *)(a.seqs + wa)) : 0u;
                uint32_t up = __shfl_up_sync(0xffffffffu, u, 1);
                uint32_t prevb = lane == 0 ? hcarry : (up >> 24);
                hcarry = __shfl_sync(0xffffffffu, u, 31) >> 24;
                V2Dig d; d.run80 = v2_run80(u, (u << 8) | prevb, hpc); d.symw = v3_symw(u) | v3_bad80(u);
                uint32_t m = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) if (wa + b < ge) m |= 0x80u << (8 * b);
                const uint32_t run = d.run80 & m;
                uint32_t mine = __popc(run), tot;
                uint32_t r = hcount + warp_excl_scan(mine, &tot);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (run & (0x80u << (8 * b))) {
                        const uint32_t sy = (d.symw >> (8 * b)) & 0xFFu;
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
                        uint32_t x = j == 32 ? lds32(ha + 4 * w) : (w < (nj >> 2) ? lds32(ws_a + 4 * j + 128 * w) : lds32(ws_a + 4 * j + V3_CAND_TOP));
                        const uint32_t nb = min(take - 4 * w, 4u);
                        if (nb < 4) x &= (1u << (8 * nb)) - 1u;
                        v3_push(q, x, 8 * nb);
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
        V2Lane st; st.F = T.F0; st.R = T.R0;
        const int lim = (int)(n + c);                           // symbols available in my logical stream
        const int o2 = max(-1, min((int)n - 1, lim - (int)l));  // first ordinal whose window is complete and mine
        V2H H;
        H.flo = (uint32_t)st.F; H.fhi = (uint32_t)(st.F >> 32); H.rlo = (uint32_t)st.R; H.rhi = (uint32_t)(st.R >> 32);
        const uint32_t lr = (l >> 2) * 128u, ck = lr + 384u;
        uint32_t cq = ctop - ck;
        if (anyN) {
            const uint32_t live0 = sb + 128u * ((uint32_t)lim >> 2) + 256u - ck;   // everything up to row lim/4 stays live
            int o = lim - 1;
            for (; o > o2; o--) v3_step_generic(st, sb, o, lim, l, ta);
            for (; o >= 0; o--) {
                v3_step_generic(st, sb, o, lim, l, ta);
                H.flo = (uint32_t)st.F; H.fhi = (uint32_t)(st.F >> 32); H.rlo = (uint32_t)st.R; H.rhi = (uint32_t)(st.R >> 32);
                V3_CANDIDATE(o, live0)
            }
        } else if (n != 0) {
            int w = (lim - 1) >> 2;
            const int wm = (int)(n - 1) >> 2;
            uint32_t wa = sb + 128u * (uint32_t)w;
            for (; w > wm; w--, wa -= 128u) {
                const uint32_t x = lds32(wa);
                v3_step(H, v3_tab(ta, (x >> 24)), k31, k2);
                v3_step(H, v3_tab(ta, ((x >> 16) & 0xFFu)), k31, k2);
                v3_step(H, v3_tab(ta, ((x >> 8) & 0xFFu)), k31, k2);
                v3_step(H, v3_tab(ta, (x & 0xFFu)), k31, k2);
            }
            const uint32_t ls = 8 * (l & 3);
            uint32_t probe_acc = lane; (void)probe_acc;
            // software pipeline: the three stream words of the next iteration are loaded one iteration ahead,
            // and the table rows of an iteration are issued before its four dependent hash steps
            uint32_t inw = lds32(wa), ow0 = lds32(wa + lr), ow1 = lds32(wa + lr + 128);
            for (; w >= 0; w--) {
                const uint32_t comb = inw | (__funnelshift_r(ow0, ow1, ls) << 2);   // per byte: in + out * 4, pre-scaled (no N in this tile)
                const uint4 t3 = v3_tab(ta, prmt(comb, 0u, 0x4443u)), t2 = v3_tab(ta, prmt(comb, 0u, 0x4442u));
                const uint4 t1 = v3_tab(ta, prmt(comb, 0u, 0x4441u)), t0 = v3_tab(ta, prmt(comb, 0u, 0x4440u));
                if (w > 0) { wa -= 128; inw = lds32(wa); ow0 = lds32(wa + lr); ow1 = lds32(wa + lr + 128); }
                v3_step(H, t3, k31, k2); V3_PROBE(t3) V3_CANDIDATE(4 * w + 3, wa)
                v3_step(H, t2, k31, k2); V3_PROBE(t2) V3_CANDIDATE(4 * w + 2, wa)
                v3_step(H, t1, k31, k2); V3_PROBE(t1) V3_CANDIDATE(4 * w + 1, wa)
                v3_step(H, t0, k31, k2); V3_PROBE(t0) V3_CANDIDATE(4 * w, wa)
            }
#if MQ_V3_PROBE_ALU || MQ_V3_PROBE_FMA
            if (probe_acc == 0x9E3779B9u && l == 77u) a.tile_cnt[tile] = probe_acc;   // never true (l <= 32): keeps the probe instructions
#endif
        }
        if (cq != ctop - ck) nloc = v3_flush(ctop, cq + ck, nloc, runm_l, cum_l, c_lo, xlo, xlim, lane, ev_a, tile, a, bound, o2);
        __syncwarp();
        a.lane_cnt[(uint64_t)tile * 32 + lane] = (uint16_t)nloc;
        if (lane == 0) a.tile_cnt[tile] = lds32(ev_a);
        __syncwarp();
    }
}
#undef V3_CANDIDATE

}  // namespace mq
