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
#include "mq_kernels.cuh"

namespace mq {

constexpr int V2_CS_MAX  = 256;                 // raw bytes per lane chunk
constexpr int V2_GPL_MAX = V2_CS_MAX / 16;      // 16-byte groups per lane
constexpr int V2_STRIDE  = 256 + 32 + 4;        // bytes per lane stream: symbols + context + zero; 73 words (odd)
constexpr int V2_WARP_SMEM = 33 * V2_STRIDE + 33 * 4 /*nsym*/ + 32 * V2_GPL_MAX * 2 /*run masks*/ + 32 * V2_GPL_MAX /*cum*/ + 12;
constexpr int V2_WARPS   = 4;

struct ScanTablesV2 {
    uint64_t pairF[16], pairR[16];              // [in + 4*out]
    uint64_t inF[4], outF[4], inR[4], outR[4];
    uint64_t F0, R0;                            // hash state of a window of l phantom 'A's
    uint32_t sel[16];                           // PRMT selectors compacting the run-start bytes of a word
};

// tiles of record i on the 16-byte aligned grid
__global__ void k_tiles_per_seq_v2(const uint64_t *offs, uint32_t n, uint32_t min_len, uint32_t *tiles) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t gs = offs[i], ge = offs[i + 1], len = ge - gs;
    uint32_t t = 0;
    if (len >= min_len && len > 0) { uint64_t span = ge - (gs & ~15ull); t = (uint32_t)((span + TW_MAX - 1) / TW_MAX); }
    tiles[i] = t;
}

__device__ __forceinline__ void v2_geometry(uint64_t gs, uint64_t ge, uint32_t nt, uint32_t ti, uint32_t *Cs, uint64_t *tlo) {
    const uint64_t A = gs & ~15ull, span = ge - A;
    uint32_t c = (uint32_t)((span + 32ull * nt - 1) / (32ull * nt));
    c = (c + 15u) & ~15u;
    *Cs = c; *tlo = A + (uint64_t)ti * 32u * c;
}

struct V2Lane { uint64_t F, R; };

// generic (N-aware, bounds-checked) step at ordinal o of the lane's logical stream
__device__ __forceinline__ void v2_step_generic(V2Lane &s, const uint8_t *S, int o, int lim, uint32_t l, const ScanTablesV2 &T) {
    const uint32_t in = S[o];
    const int oo = o + (int)l;
    const uint32_t out = oo < lim ? (uint32_t)S[oo] : 0u;
    const uint64_t tf = ((in & 0x80u) ? 0ull : T.inF[(in >> 3) & 3]) ^ ((out & 0x80u) ? 0ull : T.outF[(out >> 3) & 3]);
    const uint64_t tr = ((in & 0x80u) ? 0ull : T.inR[(in >> 3) & 3]) ^ ((out & 0x80u) ? 0ull : T.outR[(out >> 3) & 3]);
    s.F = ror1(s.F) ^ tf; s.R = rol1(s.R) ^ tr;
}

// raw offset (inside the lane chunk) of the symbol with ordinal o
__device__ __noinline__ uint32_t v2_raw_offset(const uint16_t *runm, const uint8_t *cum, uint32_t gpl, uint32_t o) {
    uint32_t g = 0;
    while (g + 1 < gpl && (uint32_t)cum[g + 1] <= o) g++;
    return 16u * g + __fns((uint32_t)runm[g], 0, (int)(o - cum[g]) + 1);
}

__device__ __forceinline__ void v2_emit(uint32_t x, uint64_t h, uint32_t lane, uint32_t &nloc, uint32_t *tile_ev_smem, uint32_t tile,
                                        const ScanArgs &a) {
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

__global__ void __launch_bounds__(V2_WARPS * 32) k_scan_minimizers_v2(ScanArgs a, ScanTablesV2 Tin) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(128) ScanTablesV2 T;
    __shared__ uint32_t ev_cnt[V2_WARPS];
    for (uint32_t i = threadIdx.x; i < sizeof(ScanTablesV2) / 4; i += blockDim.x) ((uint32_t *)&T)[i] = ((const uint32_t *)&Tin)[i];
    __syncthreads();
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    uint8_t *WS = smem_raw + (size_t)wid * ((V2_WARP_SMEM + 15) & ~15);
    uint8_t *S = WS + lane * V2_STRIDE;                    // my stream
    uint32_t *nsym = (uint32_t *)(WS + 33 * V2_STRIDE);    // symbols per stream (32 lanes + halo)
    uint16_t *runm = (uint16_t *)(nsym + 33) + lane * V2_GPL_MAX;
    uint8_t *cum = (uint8_t *)((uint16_t *)(nsym + 33) + 32 * V2_GPL_MAX) + lane * V2_GPL_MAX;
    const uint32_t l = a.l;
    const bool hpc = a.use_hpc != 0;
    const uint32_t bound_hi = (uint32_t)(a.bound >> 32);
    const uint64_t bound = a.bound;
    const uint8_t *TF = (const uint8_t *)T.pairF, *TR = (const uint8_t *)T.pairR;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_ticket, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) break;
        if (lane == 0) ev_cnt[wid] = 0;

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

        // ---- stage + compact my chunk -----------------------------------------------------------
        const uint32_t c_lo = lane * Cs;                       // x' of my first byte
        uint32_t n = 0, bad = 0;
        {
            const uint8_t *cp = a.seqs + tlo + c_lo;
            uint32_t prev = 0;
            if (c_lo > own_lo && c_lo < own_hi) prev = cp[-1];  // byte before my chunk (same record)
            else if (c_lo == own_lo && tlo + own_lo > gs) prev = cp[-1];
            uint64_t acc = 0; uint32_t fill = 0, wout = 0;
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (c_lo < own_hi && c_lo + 16 > own_lo) nxt = __ldg((const uint4 *)cp);
            for (uint32_t g = 0; g < gpl; g++) {
                const uint4 v = nxt;
                const uint32_t xg = c_lo + 16 * g;
                const bool live = xg < own_hi && xg + 16 > own_lo;
                const uint32_t xn = xg + 16;
                if (g + 1 < gpl && xn < own_hi && xn + 16 > own_lo) nxt = __ldg((const uint4 *)(cp + 16 * (g + 1)));
                uint32_t rm = 0;
                if (live) {
                    const bool partial = xg < own_lo || xg + 16 > own_hi;
                    const uint32_t uw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        const uint32_t u = uw[w];
                        uint32_t dg = digest_word(u, (u << 8) | prev, hpc);
                        prev = u >> 24;
                        if (partial) {
                            const uint32_t x = xg + 4 * w;
                            uint32_t m = 0;
#pragma unroll
                            for (int b = 0; b < 4; b++) if (x + b >= own_lo && x + b < own_hi) m |= 0xFFu << (8 * b);
                            dg &= m;
                            if (x <= own_lo && own_lo < x + 4 && tlo + own_lo == gs) dg |= D_RUN << (8 * (own_lo - x));   // record start
                        } else if (xg + 4 * w == own_lo && tlo + own_lo == gs) dg |= D_RUN;                             // aligned record start
                        const uint32_t p = (((dg >> 3) & 0x01010101u) * 0x01020408u) >> 24;      // 4 run bits
                        bad |= dg & 0x04040404u & ((dg & 0x08080808u) >> 1);           // non-ACGT among run starts
                        const uint32_t symw = ((dg & 0x03030303u) << 3) | ((dg & 0x04040404u) << 5);
                        const uint32_t comp = __byte_perm(symw, 0u, T.sel[p]);
                        const uint32_t cnt = __popc(p);
                        rm |= p << (4 * w);
                        acc |= (uint64_t)comp << (8 * fill);
                        fill += cnt;
                        if (fill >= 4) { *(uint32_t *)(S + 4 * wout) = (uint32_t)acc; acc >>= 32; fill -= 4; wout++; }
                    }
                }
                runm[g] = (uint16_t)rm; cum[g] = (uint8_t)n;
                n += __popc(rm);
            }
            *(uint32_t *)(S + 4 * wout) = (uint32_t)acc;     // tail (<= 3 symbols + zeros)
        }
        nsym[lane] = n;

        // ---- halo stream (slot 32): up to l-1 run-start symbols right of the tile -------------------
        uint32_t hcount = 0;
        if (tlo + TWs < ge) {
            uint8_t *H = WS + 32 * V2_STRIDE;
            uint64_t haddr = tlo + TWs;
            uint32_t hcarry = a.seqs[haddr - 1];
            while (hcount < l - 1 && haddr < ge) {
                const uint64_t wa = haddr + 4ull * lane;
                uint32_t u = (wa < ge) ? __ldg((const uint32_t *)(a.seqs + wa)) : 0u;
                uint32_t up = __shfl_up_sync(0xffffffffu, u, 1);
                uint32_t prevb = lane == 0 ? hcarry : (up >> 24);
                hcarry = __shfl_sync(0xffffffffu, u, 31) >> 24;
                uint32_t dg = digest_word(u, (u << 8) | prevb, hpc);
                uint32_t m = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) if (wa + b < ge) m |= 0xFFu << (8 * b);
                dg &= m;
                uint32_t mine = __popc((dg >> 3) & 0x01010101u), tot;
                uint32_t r = hcount + warp_excl_scan(mine, &tot);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t d = (dg >> (8 * b)) & 0xFu;
                    if (d & D_RUN) { if (r < l - 1) { H[r] = (uint8_t)(((d & 3u) << 3) | ((d & D_N) << 5)); bad |= d & D_N; } r++; }
                }
                hcount = min(hcount + tot, l - 1);
                haddr += 128;
            }
        }
        if (lane == 0) nsym[32] = hcount;
        const bool anyN = __any_sync(0xffffffffu, bad != 0);
        __syncwarp();

        // ---- context: the next l-1 symbols after my chunk, from the streams to my right -----------------
        uint32_t c = 0;
        for (uint32_t j = lane + 1; j <= 32 && c < l - 1; j++) {
            const uint32_t nj = nsym[j];
            const uint8_t *Sj = WS + j * V2_STRIDE;
            for (uint32_t i = 0; i < nj && c < l - 1; i++, c++) S[n + c] = Sj[i];
        }
        S[n + c] = 0; S[n + c + 1] = 0; S[n + c + 2] = 0; S[n + c + 3] = 0;
        __syncwarp();          // (streams are private from here on; the sync orders my context reads before later tiles)

        // ---- phase 1: warm-up over context and record-final symbols, no emission ----------------------
        V2Lane st; st.F = T.F0; st.R = T.R0;
        const int lim = (int)(n + c);                           // symbols available in my logical stream
        int o = lim - 1;
        const int o2 = max(-1, min((int)n - 1, lim - (int)l));  // first ordinal whose window is complete and mine
        for (; o > o2; o--) v2_step_generic(st, S, o, lim, l, T);

        // ---- phase 2: emitting scan of my own symbols -------------------------------------------------
        uint32_t *tev = &ev_cnt[wid];
        if (anyN) {
            for (; o >= 0; o--) {
                v2_step_generic(st, S, o, lim, l, T);
                if (min((uint32_t)(st.F >> 32), (uint32_t)(st.R >> 32)) <= bound_hi) {
                    const uint64_t h = st.F < st.R ? st.F : st.R;
                    if (h < bound) { const uint32_t x = c_lo + v2_raw_offset(runm, cum, gpl, (uint32_t)o); if (x - xlo < xlim - xlo) v2_emit(x, h, lane, nloc, tev, tile, a); }
                }
            }
        } else {
            uint64_t F = st.F, R = st.R;
            // bring o+1 to a multiple of 4
            for (; o >= 0 && ((o + 1) & 3); o--) {
                const uint32_t off = (uint32_t)S[o] | ((uint32_t)S[o + (int)l] << 2);
                F = ror1(F) ^ *(const uint64_t *)(TF + off); R = rol1(R) ^ *(const uint64_t *)(TR + off);
                if (min((uint32_t)(F >> 32), (uint32_t)(R >> 32)) <= bound_hi) {
                    const uint64_t h = F < R ? F : R;
                    if (h < bound) { const uint32_t x = c_lo + v2_raw_offset(runm, cum, gpl, (uint32_t)o); if (x - xlo < xlim - xlo) v2_emit(x, h, lane, nloc, tev, tile, a); }
                }
            }
            const uint32_t *S32 = (const uint32_t *)S;
            const uint32_t lw = l >> 2, ls = 8 * (l & 3);
            for (int w = ((o + 1) >> 2) - 1; w >= 0; w--) {
                const uint32_t inw = S32[w];
                const uint32_t outw = __funnelshift_r(S32[w + lw], S32[w + lw + 1], ls);
                const uint32_t comb = inw | (outw << 2);          // per byte: in*8 + out*32 (bit 7 clear: no N in this tile)
#pragma unroll
                for (int b = 3; b >= 0; b--) {
                    const uint32_t off = (comb >> (8 * b)) & 0xFFu;
                    F = ror1(F) ^ *(const uint64_t *)(TF + off); R = rol1(R) ^ *(const uint64_t *)(TR + off);
                    if (min((uint32_t)(F >> 32), (uint32_t)(R >> 32)) <= bound_hi) {
                        const uint64_t h = F < R ? F : R;
                        if (h < bound) {
                            const uint32_t x = c_lo + v2_raw_offset(runm, cum, gpl, (uint32_t)(4 * w + b));
                            if (x - xlo < xlim - xlo) v2_emit(x, h, lane, nloc, tev, tile, a);
                        }
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

}  // namespace mq
