// mq_lib.cu -- host side of libmapquik_b200.so: contexts, device memory, kernel launches and the extern "C" entry
// points declared in include/mapquik_b200.h.  No torch, no CPU fallback: every computing entry point fails with
// MQ_ERR_CUDA when no CUDA device is usable.
//
// Shape of the mapping path (mq_map_batch*): the batch is cut into sub-batches of <= SUB_BASES bases that flow
// through four slots -- uploads (copy stream) run ahead of the kernels (compute stream), which run ahead of the
// downloads of the hits (d2h stream).  The host never waits for the GPU between the upload of a
// sub-batch and the download of its hits: tile tables are computed on the host from the offsets it already has,
// device buffers are sized from upper bounds, and the one thing only the GPU knows (how many minimizers a sub-batch
// produced, whether a pool overflowed) comes back in a 48-byte status record next to the hits and is looked at when
// the slot is recycled.  A sub-batch whose status reports an overflow is redone with larger buffers (rare: the
// bounds assume at most ~3x the expected minimizer density).
#include "../../include/mapquik_b200.h"
#include "mq_kernels.cuh"
#include "mq_scan.cuh"

#include <algorithm>
#include <array>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

using namespace mq;

namespace {

struct DBuf {
    void *p = nullptr; size_t cap = 0;
    template <class T> T *as() const { return (T *)p; }
};
struct HBuf { void *p = nullptr; size_t cap = 0; };      // pinned host memory

#ifndef MQ_SUB_MBASES
#define MQ_SUB_MBASES 128
#endif
#ifndef MQ_SUB_MBASES_LIGHT
#define MQ_SUB_MBASES_LIGHT 512
#endif
constexpr uint64_t SUB_BASES = (uint64_t)MQ_SUB_MBASES << 20;             // bases per pipelined mapping sub-batch (ASCII from the host: 1 byte per base over PCIe)
constexpr uint64_t SUB_BASES_LIGHT = (uint64_t)MQ_SUB_MBASES_LIGHT << 20; // ... when the upload is light (packed input from the host): fewer, fuller launches
#ifndef MQ_SUB_MBASES_RESIDENT
#define MQ_SUB_MBASES_RESIDENT 2048
#endif
constexpr uint64_t SUB_BASES_RESIDENT = (uint64_t)MQ_SUB_MBASES_RESIDENT << 20;   // ... when there is no upload at all (inputs in HBM): nothing to overlap, only launch gaps to lose
constexpr uint64_t ADD_BASES = 256ull << 20;            // bases per pipelined index-build piece batch
constexpr size_t   PAD = 256;                          // zeroed slack behind sequence buffers (word loads past the end)

// One piece of work for the scan: bases [lo, hi) of the caller's sequence array form a "record" of the batch.
// Whole records: lo..hi = the record, emit everything.  Segments of a long reference record: the piece starts one
// base before seg_start (run context) and extends past its own range by the right halo; only l-mers STARTING in
// [emit_lo, emit_hi) (piece coordinates) are emitted and positions are shifted by pos_base.
struct Piece { uint64_t lo, hi; uint32_t ref_idx; uint32_t pos_base, emit_lo, emit_hi; uint64_t seg_start; };

struct BatchDev {              // device view of one staged sub-batch
    const uint8_t *seqs = nullptr; const uint32_t *packed = nullptr, *flags = nullptr; const ExcRec *exc = nullptr; uint32_t n_exc = 0;
    uint64_t exc_base = 0;     // exception intervals are in coordinates of (slice base index + exc_base)
    const uint64_t *offs = nullptr; const uint32_t *first_tile = nullptr, *tile_seq = nullptr, *pos_base = nullptr, *emit = nullptr;
    uint32_t n = 0, n_tiles = 0; uint64_t bases = 0;
    BatchScalars *sc = nullptr;
};

struct Slot2 {                 // resources of one in-flight sub-batch
    DBuf d_in;                 // ASCII bytes, or packed words
    DBuf d_meta;               // offs u64[n+1] | first_tile u32[n+1] | tile_seq u32[n_tiles] | pos_base u32[n] | emit u32[2n] | flags | exc
    HBuf h_meta;
    DBuf d_hits, d_sc;
    HBuf h_sc;                 // BatchScalars read back
    cudaEvent_t ev_copied = nullptr, ev_comp = nullptr, ev_done = nullptr;
    bool busy = false;
    int pack_buf = -1;         // staging buffer of the on-the-fly packer this sub-batch was uploaded from (released on retire)
    BatchDev bd;               // what the slot holds (device view; stays valid until the slot is staged again)
    uint32_t i0 = 0, i1 = 0;   // records [i0, i1) of the call
};

// pinned staging of one sub-batch packed on the fly by host threads (mq_set_host_threads)
struct PackBuf { HBuf words, flags; std::vector<mq_exc> exc; int state = 0; };   // 0 free, 1 being packed, 2 ready, 3 uploading
constexpr int N_PACK_BUFS = 8;
constexpr int N_SLOTS = 4;       // sub-batches in flight in the mapping pipeline (the index build uses two of them)

}  // namespace

// mq_pack.cpp
void mq_pack_units(const uint8_t *ascii, uint64_t n_bases, uint64_t u0, uint64_t u1, uint32_t *words, uint32_t *flags,
                   std::vector<mq_exc> &exc, bool fold_case);

struct mq_ctx {
    mq_params p{};
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr;
    uint64_t bound = 0;
    ScanTables tab{};
    DBuf d_warm;                  // 256 x uint4: the four-step warm-up table of the scan kernel
    int scan_ctas_per_sm[4] = {0, 0, 0, 0};
    std::string err;
    uint64_t launches = 0, scan_kernel_launches = 0;
    int n_sm = 148;
    cudaEvent_t region_a = nullptr, region_b = nullptr;
    uint64_t last_minimizers = 0;
    // workspace shared by the sub-batches (used on the compute stream only)
    DBuf d_ev_hash, d_ev_meta, d_lane_cnt, d_tile_cnt, d_blk, d_ovf_tile, d_ovf_meta, d_ovf_hash, d_pos, d_hash, d_seq_off,
         d_matches, d_nmatch, d_big_list, d_misc;
    uint32_t ovf_cap = 1u << 16;
    uint64_t mini_cap = 0;        // entries d_pos / d_hash / d_matches can hold
    double mini_rate = 0;         // learned upper estimate of minimizers per base (grows on overflow)
    Slot2 slot[N_SLOTS];
    // on-the-fly packing of ASCII host input (mq_set_host_threads)
    int host_threads = 0;
    uint64_t sub_bases = SUB_BASES, sub_bases_light = SUB_BASES_LIGHT, sub_bases_resident = SUB_BASES_RESIDENT;   // MQ_SUB_BASES (test knob, read at mq_create) sets all three
    PackBuf pack_buf[N_PACK_BUFS];
    uint64_t ctr_h2d_bytes = 0, ctr_host_packed_bases = 0, ctr_host_packed_subs = 0, ctr_subs = 0;   // of the last mapping call
    // minimizer store (reference side)
    DBuf st_pos, st_hash; uint64_t st_n = 0;
    std::vector<std::array<uint64_t, 3>> dir;          // (ref_idx, seg_start, count)
    // frozen index
    DBuf d_table, d_ref_lens; uint64_t tmask = 0; bool frozen = false; uint32_t n_refs = 0;
    DBuf d_bloom; uint32_t bloom_wmask = 0;      // presence filter in front of the table (mq_kernels.cuh), L2-resident
    std::vector<uint64_t> nb_mers; uint64_t n_unique = 0, n_keys = 0;
    HBuf h_pin;                   // bounce buffer (index save / load)
    // timings
    std::map<std::string, float> ms, ms_total;
    uint64_t timer_gen = 0;
    struct PendingTimer { const char *name; cudaEvent_t a, b; uint64_t gen; };
    std::vector<PendingTimer> pending;
    std::vector<cudaEvent_t> ev_pool;
    // multi-GPU parent: fans out to one child context per device (mq_create_multi)
    std::vector<mq_ctx *> kids;
    std::vector<std::array<uint64_t, 2>> kid_share;     // [first read, last read) mapped by each child in the last call
};

namespace {

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
            return MQ_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

int ensure(mq_ctx *c, DBuf &b, size_t bytes) {
    if (bytes <= b.cap) return MQ_OK;
    if (b.p) { cudaStreamSynchronize(c->stream); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
        if (e != cudaSuccess) { cudaGetLastError(); c->err = "cudaMalloc failed for " + std::to_string(bytes) + " bytes"; b.p = nullptr; return MQ_ERR_NOMEM; }
    }
    b.cap = want;
    return MQ_OK;
}
// grow preserving the first `keep` bytes
int ensure_keep(mq_ctx *c, DBuf &b, size_t bytes, size_t keep) {
    if (bytes <= b.cap) return MQ_OK;
    DBuf nb; size_t want = bytes + bytes / 2 + 256;
    cudaError_t e = cudaMalloc(&nb.p, want);
    if (e != cudaSuccess) { cudaGetLastError(); want = bytes; e = cudaMalloc(&nb.p, want); }
    if (e != cudaSuccess) { cudaGetLastError(); c->err = "cudaMalloc failed (store growth)"; return MQ_ERR_NOMEM; }
    nb.cap = want;
    if (keep && b.p) CK(cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, c->stream));
    if (b.p) { CK(cudaStreamSynchronize(c->stream)); cudaFree(b.p); }
    b = nb;
    return MQ_OK;
}
void dfree(DBuf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

int ensure_host(mq_ctx *c, HBuf &b, size_t bytes) {
    if (bytes <= b.cap) return MQ_OK;
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr; b.cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    CK(cudaMallocHost(&b.p, want));
    b.cap = want;
    return MQ_OK;
}
void hfree(HBuf &b) { if (b.p) cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }

// ---- stage timing (CUDA events on the stream the stage runs on) ----------------------------------
cudaEvent_t get_event(mq_ctx *c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct StageTimer {
    mq_ctx *c; const char *name; cudaEvent_t a, b; cudaStream_t st;
    StageTimer(mq_ctx *c_, const char *n, cudaStream_t s_ = nullptr) : c(c_), name(n), st(s_ ? s_ : c_->stream) {
        a = get_event(c); b = get_event(c); cudaEventRecord(a, st);
    }
    ~StageTimer() { cudaEventRecord(b, st); c->pending.push_back({name, a, b, c->timer_gen}); }
};
// fold every timer whose end event has completed into the totals (never blocks unless `wait`)
void timers_collect(mq_ctx *c, bool wait = true) {
    size_t w = 0;
    for (size_t i = 0; i < c->pending.size(); i++) {
        auto &pr = c->pending[i];
        if (!wait && cudaEventQuery(pr.b) != cudaSuccess) { cudaGetLastError(); c->pending[w++] = pr; continue; }
        float t = 0; cudaEventSynchronize(pr.b); cudaEventElapsedTime(&t, pr.a, pr.b);
        if (pr.gen == c->timer_gen) c->ms[pr.name] += t;
        c->ms_total[pr.name] += t;
        c->ev_pool.push_back(pr.a); c->ev_pool.push_back(pr.b);
    }
    c->pending.resize(w);
}
// a new call starts a new generation: per-call figures (mq_last_ms) restart, running totals (mq_total_ms) keep
// accumulating.  Finished timers of earlier calls are folded in here, so a caller that never asks for timings does
// not accumulate events.
void timers_reset(mq_ctx *c) { timers_collect(c, false); c->ms.clear(); c->timer_gen++; }

uint64_t hash_bound(double density) {   // (density as FH * H::MAX as FH) as H, saturating like Rust `as`
    double b = density * 18446744073709551615.0;
    if (!(b > 0.0)) return 0;
    if (b >= 18446744073709551616.0) return ~0ull;
    return (uint64_t)b;
}
uint64_t hrol(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }

// symbol codes are the raw bits (ascii >> 1) & 3: A=0 C=1 T=2 G=3
void fill_tables(ScanTables &T, uint32_t l) {
    memset(&T, 0, sizeof(T));
    const uint64_t h[4] = {SEED_A, SEED_C, SEED_T, SEED_G}, hc[4] = {SEED_T, SEED_G, SEED_A, SEED_C};
    for (int i = 0; i < 4; i++) {
        T.inF[i] = hrol(h[i], l - 1); T.outF[i] = hrol(h[i], 63);
        T.inR[i] = hc[i];             T.outR[i] = hrol(hc[i], l);
    }
    for (int i = 0; i < 4; i++) for (int o = 0; o < 4; o++) {
        T.pairF[i + 4 * o] = T.inF[i] ^ T.outF[o];
        T.pairR[i + 4 * o] = T.inR[i] ^ T.outR[o];
    }
    for (uint32_t i = 0; i < l; i++) { T.F0 ^= hrol(SEED_A, l - 1 - i); T.R0 ^= hrol(SEED_T, i); }   // window of l phantom 'A's
    auto selector = [](uint32_t p) {          // byte-permute selector: run-start bytes first, zero fill
        uint32_t sel = 0, j = 0;
        for (uint32_t b = 0; b < 4; b++) if (p & (1u << b)) sel |= b << (4 * j++);
        for (; j < 4; j++) sel |= 4u << (4 * j);
        return sel;
    };
    for (uint32_t p = 0; p < 16; p++) T.sel[p] = selector(p);
    for (uint32_t q = 0; q < 86; q++) {       // run bits at even positions -> selector | run bits << 16 | 8 * count << 24
        const uint32_t p = (q & 1u) | ((q >> 1) & 2u) | ((q >> 2) & 4u) | ((q >> 3) & 8u);
        T.sel55[q] = selector(p) | (p << 16) | ((uint32_t)__builtin_popcount(p) * 8u << 24);
    }
}

// C[idx] of warm_row (mq_scan.cuh): the state after the four steps in = c3, c2, c1, c0 (out = phantom 'A') from state 0
void fill_warm_table(const ScanTables &T, uint32_t out[256][4]) {
    for (uint32_t idx = 0; idx < 256; idx++) {
        uint64_t F = 0, R = 0;
        for (int b = 3; b >= 0; b--) {
            const uint32_t c = (idx >> (2 * b)) & 3u;
            F = ((F >> 1) | (F << 63)) ^ T.pairF[c];
            R = ((R << 1) | (R >> 63)) ^ T.pairR[c];
        }
        out[idx][0] = (uint32_t)F; out[idx][1] = (uint32_t)(F >> 32); out[idx][2] = (uint32_t)R; out[idx][3] = (uint32_t)(R >> 32);
    }
}

// ---- tile tables (host) ------------------------------------------------------------------------------
// layout of a slot's meta block; offsets in bytes, all 16-byte aligned
struct MetaLayout {
    size_t offs, first_tile, tile_seq, pos_base, emit, flags, exc, total;
    MetaLayout(uint32_t n, uint32_t n_tiles, bool pieces, size_t flag_words, size_t n_exc) {
        auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
        size_t o = 0;
        offs = o; o = al(o + ((size_t)n + 1) * 8);
        first_tile = o; o = al(o + ((size_t)n + 1) * 4);
        tile_seq = o; o = al(o + (size_t)n_tiles * 4);
        pos_base = o; o = al(o + (pieces ? (size_t)n * 4 : 0));
        emit = o; o = al(o + (pieces ? (size_t)n * 8 : 0));
        flags = o; o = al(o + flag_words * 4);
        exc = o; o = al(o + n_exc * sizeof(ExcRec));
        total = o;
    }
};

uint64_t mini_estimate(const mq_ctx *c, uint64_t bases) {
    double rate = c->mini_rate;
    if (rate <= 0) { rate = 2.0 * c->p.density * 1.5 + 0.004; }
    if (rate > 1.0) rate = 1.0;
    uint64_t est = (uint64_t)(rate * (double)bases) + 65536;
    return std::min<uint64_t>(est, bases + 64);
}

int ensure_workspace(mq_ctx *c, uint32_t n, uint32_t n_tiles, uint64_t bases, bool for_map) {
    int rc;
    if ((rc = ensure(c, c->d_ev_hash, (size_t)n_tiles * EV_CAP * 8 + 64))) return rc;
    if ((rc = ensure(c, c->d_ev_meta, (size_t)n_tiles * EV_CAP * 4 + 64))) return rc;
    if ((rc = ensure(c, c->d_lane_cnt, (size_t)n_tiles * 32 * 2 + 64))) return rc;
    if ((rc = ensure(c, c->d_tile_cnt, ((size_t)n_tiles + 2) * 4))) return rc;
    if ((rc = ensure(c, c->d_blk, ((size_t)n_tiles / PREFIX_SPAN + 2) * 4))) return rc;
    if ((rc = ensure(c, c->d_ovf_tile, (size_t)c->ovf_cap * 4))) return rc;
    if ((rc = ensure(c, c->d_ovf_meta, (size_t)c->ovf_cap * 4))) return rc;
    if ((rc = ensure(c, c->d_ovf_hash, (size_t)c->ovf_cap * 8))) return rc;
    const uint64_t want = mini_estimate(c, bases);
    if (want > c->mini_cap) {
        if ((rc = ensure(c, c->d_pos, (want + 64) * 4))) return rc;
        if ((rc = ensure(c, c->d_hash, (want + 64) * 8))) return rc;
        c->mini_cap = want;
    }
    if (for_map) {
        if ((rc = ensure(c, c->d_matches, (c->mini_cap + 64) * sizeof(MatchRec)))) return rc;
        if ((rc = ensure(c, c->d_nmatch, ((size_t)n + 1) * 4))) return rc;
        if ((rc = ensure(c, c->d_big_list, ((size_t)n + 2) * 4))) return rc;
    }
    return MQ_OK;
}

template <bool HPC, bool PACKED>
int launch_scan_t(mq_ctx *c, const ScanArgs &a) {
    const int vi = (HPC ? 1 : 0) | (PACKED ? 2 : 0);
    const size_t smem = (size_t)SCAN_WARPS * WARP_BYTES;
    auto kern = k_scan_minimizers<HPC, PACKED>;
    if (c->scan_ctas_per_sm[vi] == 0) {      // persistent grid = every CTA the chip can hold
        int nb = 0;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, SCAN_WARPS * 32, smem) != cudaSuccess || nb < 1) nb = 1;
        c->scan_ctas_per_sm[vi] = nb;
    }
    const uint32_t ctas_needed = (a.n_tiles + SCAN_WARPS - 1) / SCAN_WARPS;
    const uint32_t grid = std::min<uint32_t>(ctas_needed, (uint32_t)(c->n_sm * c->scan_ctas_per_sm[vi]));
    kern<<<grid, SCAN_WARPS * 32, smem, c->stream>>>(a, c->tab);
    return MQ_OK;
}

// S1 on a staged sub-batch: scalars memset, scan, tile prefix, gather.  Result: c->d_pos / c->d_hash; the minimizer
// range of record i is tile_base(first_tile[i]) .. tile_base(first_tile[i+1]).  Nothing here waits for the GPU.
int enqueue_scan(mq_ctx *c, const BatchDev &b) {
    CK(cudaMemsetAsync(b.sc, 0, sizeof(BatchScalars), c->stream));
    if (b.n_tiles == 0) return MQ_OK;
    ScanArgs a{};
    a.seqs = b.seqs; a.packed = b.packed; a.flags = b.flags; a.exc = b.exc; a.n_exc = b.n_exc; a.exc_base = b.exc_base;
    a.offs = b.offs; a.first_tile = b.first_tile; a.tile_seq = b.tile_seq; a.n_tiles = b.n_tiles; a.l = c->p.l; a.bound = c->bound;
    a.ev_hash = c->d_ev_hash.as<uint64_t>(); a.ev_meta = c->d_ev_meta.as<uint32_t>();
    a.lane_cnt = c->d_lane_cnt.as<uint16_t>(); a.tile_cnt = c->d_tile_cnt.as<uint32_t>();
    a.ovf_count = &b.sc->ovf_count; a.ovf_cap = c->ovf_cap;
    a.ovf_tile = c->d_ovf_tile.as<uint32_t>(); a.ovf_meta = c->d_ovf_meta.as<uint32_t>(); a.ovf_hash = c->d_ovf_hash.as<uint64_t>();
    a.tile_ticket = &b.sc->scan_ticket; a.emit_range = b.emit; a.warm = c->d_warm.as<uint4>();
    {
        StageTimer t(c, "scan");
        {
            StageTimer tk(c, "scan_kernel");   // the dominant kernel alone (roofline numerator)
            const bool packed = b.packed != nullptr;
            if (c->p.use_hpc) { if (packed) launch_scan_t<true, true>(c, a); else launch_scan_t<true, false>(c, a); }
            else { if (packed) launch_scan_t<false, true>(c, a); else launch_scan_t<false, false>(c, a); }
        }
        c->launches++; c->scan_kernel_launches++;
        CK(cudaGetLastError());
    }
    {
        StageTimer t(c, "gather");
        const uint32_t nb = (b.n_tiles + 1 + PREFIX_SPAN - 1) / PREFIX_SPAN;
        k_tile_prefix<<<nb, SCAN_BLK, 0, c->stream>>>(c->d_tile_cnt.as<uint32_t>(), b.n_tiles, c->d_blk.as<uint32_t>(), b.sc, c->mini_cap, c->ovf_cap);
        GatherArgs g{};
        g.ev_hash = c->d_ev_hash.as<uint64_t>(); g.ev_meta = c->d_ev_meta.as<uint32_t>(); g.lane_cnt = c->d_lane_cnt.as<uint16_t>();
        g.tb = TileBase{c->d_tile_cnt.as<uint32_t>(), c->d_blk.as<uint32_t>()};
        g.tile_seq = b.tile_seq; g.first_tile = b.first_tile; g.offs = b.offs; g.pos_base = b.pos_base; g.n_tiles = b.n_tiles; g.sc = b.sc;
        g.out_pos = c->d_pos.as<uint32_t>(); g.out_hash = c->d_hash.as<uint64_t>();
        k_gather_minimizers<<<(b.n_tiles + 7) / 8, 256, 0, c->stream>>>(g);
        k_gather_overflow<<<32, 256, 0, c->stream>>>(g, c->d_ovf_tile.as<uint32_t>(), c->d_ovf_meta.as<uint32_t>(), c->d_ovf_hash.as<uint64_t>());
        c->launches += 3;
        CK(cudaGetLastError());
    }
    return MQ_OK;
}

// probe / match / chain of a scanned sub-batch
int enqueue_probe_chain(mq_ctx *c, const BatchDev &b, HitRec *d_hits) {
    const uint32_t n = b.n;
    const TileBase tb{c->d_tile_cnt.as<uint32_t>(), c->d_blk.as<uint32_t>()};
    if (b.n_tiles == 0) {          // no record long enough: every read is unmapped
        CK(cudaMemsetAsync(d_hits, 0, (size_t)n * sizeof(HitRec), c->stream));
        return MQ_OK;
    }
    Table t{c->d_table.as<Slot>(), c->tmask};
    {
        StageTimer tm(c, "probe");
        ProbeArgs a{};
        a.pos = c->d_pos.as<uint32_t>(); a.hash = c->d_hash.as<uint64_t>(); a.first_tile = b.first_tile; a.tb = tb; a.sc = b.sc;
        a.n_reads = n; a.k = c->p.k; a.l = c->p.l; a.matches = c->d_matches.as<MatchRec>(); a.n_matches = c->d_nmatch.as<uint32_t>();
        a.read_ticket = &b.sc->probe_ticket;
        const uint32_t grid = std::min<uint32_t>((n + 3) / 4, (uint32_t)c->n_sm * 16);
        k_probe_match<<<grid, 128, 0, c->stream>>>(a, t, Bloom{c->d_bloom.as<unsigned long long>(), c->bloom_wmask});
        c->launches++;
        CK(cudaGetLastError());
    }
    {
        StageTimer tm(c, "chain");
        ChainArgs a{};
        a.matches = c->d_matches.as<MatchRec>(); a.n_matches = c->d_nmatch.as<uint32_t>(); a.first_tile = b.first_tile; a.tb = tb; a.sc = b.sc;
        a.offs = b.offs; a.ref_lens = c->d_ref_lens.as<uint64_t>(); a.n_refs = c->n_refs; a.n_reads = n;
        a.c = c->p.c; a.s = c->p.s; a.g = c->p.g; a.hits = d_hits; a.read_ticket = &b.sc->chain_ticket;
        // thread per read for the usual handful of Matches; reads with many Matches are queued for the warp kernel
        a.big_list = c->d_big_list.as<uint32_t>(); a.big_count = &b.sc->big_count;
        k_chain_small<<<(n + 127) / 128, 128, 0, c->stream>>>(a);
        const uint32_t grid = std::min<uint32_t>((n + 3) / 4, (uint32_t)c->n_sm * 8);
        k_chain<<<grid, 128, 0, c->stream>>>(a);
        c->launches += 2;
        CK(cudaGetLastError());
    }
    return MQ_OK;
}

// ---- input descriptions ---------------------------------------------------------------------------------
struct SeqInput {               // what a batch of sequences looks like to the staging code
    bool packed = false, resident = false;           // resident: pointers are device pointers of this ctx
    uint64_t base = 0;                               // packed host input: words / flags / exc describe bases from `base` on (multiple of 2048)
    const uint8_t *seqs = nullptr;
    const uint32_t *words = nullptr, *flags = nullptr; const mq_exc *exc = nullptr; uint64_t n_exc = 0;
};
static_assert(sizeof(mq_exc) == sizeof(ExcRec) && offsetof(mq_exc, len) == offsetof(ExcRec, len), "mq_exc layout");

int init_slot(mq_ctx *c, Slot2 &s) {
    if (s.ev_done) return MQ_OK;
    CK(cudaEventCreateWithFlags(&s.ev_copied, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    int rc;
    if ((rc = ensure(c, s.d_sc, sizeof(BatchScalars)))) return rc;
    if ((rc = ensure_host(c, s.h_sc, sizeof(BatchScalars)))) return rc;
    return MQ_OK;
}
int init_streams(mq_ctx *c) {
    if (c->copy_stream) return MQ_OK;
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    int rc;
    for (auto &s : c->slot) if ((rc = init_slot(c, s))) return rc;
    return MQ_OK;
}

// Stage pieces [i0, i1) into slot s: build offsets + tile tables on the host, upload sequence slice and tables on the
// copy stream.  The slice of the caller's array starts at `origin` (16-base aligned for ASCII, 2048 for packed), piece
// offsets are relative to it.  Fills bd.
// Records are either pieces pc[i0..i1) or, when pc is NULL, the plain records offs[i] .. offs[i+1] (the mapping path:
// no per-call piece array for millions of reads).
int stage_pieces(mq_ctx *c, Slot2 &s, const SeqInput &in, const Piece *pc, const uint64_t *offs, uint32_t i0, uint32_t i1, uint32_t min_len,
                 bool pieces, BatchDev &bd) {
    const uint32_t n = i1 - i0;
    auto rec_lo = [&](uint32_t i) { return pc ? pc[i].lo : offs[i]; };
    auto rec_hi = [&](uint32_t i) { return pc ? pc[i].hi : offs[i + 1]; };
    uint64_t lo = ~0ull, hi = 0;
    if (pc) for (uint32_t i = i0; i < i1; i++) { lo = std::min(lo, pc[i].lo); hi = std::max(hi, pc[i].hi); }
    else if (n) { lo = offs[i0]; hi = offs[i1]; }
    if (n == 0 || lo > hi) { lo = hi = 0; }
    const uint64_t origin = in.packed ? (lo & ~2047ull) : (lo & ~15ull);
    // exception intervals overlapping the slice
    uint64_t e0 = 0, e1 = 0;
    const uint64_t rel = origin - in.base;          // of the slice, in the coordinates of the packed arrays
    if (in.packed && in.n_exc && !in.resident) {
        e0 = std::partition_point(in.exc, in.exc + in.n_exc, [&](const mq_exc &e) { return e.start + e.len <= rel; }) - in.exc;
        e1 = std::partition_point(in.exc, in.exc + in.n_exc, [&](const mq_exc &e) { return e.start < hi - in.base; }) - in.exc;
        if (e1 < e0) e1 = e0;
    }
    const size_t flag_words = (in.packed && !in.resident) ? (size_t)((hi - origin + 2047) / 2048 + 1) : 0;
    // tile counts first (they size the meta block)
    uint64_t n_tiles64 = 0;
    for (uint32_t i = i0; i < i1; i++) n_tiles64 += tiles_of_record(rec_lo(i) - origin, rec_hi(i) - origin, min_len);
    if (n_tiles64 >= (1ull << 31)) { c->err = "batch too large (tile count)"; return MQ_ERR_RANGE; }
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    const MetaLayout ml(n, n_tiles, pieces, flag_words, (size_t)(e1 - e0));
    int rc;
    if ((rc = ensure_host(c, s.h_meta, ml.total))) return rc;
    if ((rc = ensure(c, s.d_meta, ml.total))) return rc;
    uint8_t *hm = (uint8_t *)s.h_meta.p;
    uint64_t *h_offs = (uint64_t *)(hm + ml.offs);
    uint32_t *h_ft = (uint32_t *)(hm + ml.first_tile), *h_ts = (uint32_t *)(hm + ml.tile_seq);
    uint32_t *h_pb = (uint32_t *)(hm + ml.pos_base), *h_em = (uint32_t *)(hm + ml.emit);
    // records of a batch are contiguous in the caller's array except for index segments, whose pieces may overlap
    // (context byte, halo): every piece is its own record [lo, hi), so offs is written as n pairs collapsed into
    // n+1 boundaries only when contiguous.  Segment batches therefore hold ONE piece each (see add_pieces).
    uint32_t t = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t qlo = rec_lo(i0 + i) - origin, qhi = rec_hi(i0 + i) - origin;
        h_offs[i] = qlo;
        h_ft[i] = t;
        const uint32_t nt = tiles_of_record(qlo, qhi, min_len);
        for (uint32_t k = 0; k < nt; k++) h_ts[t + k] = i;
        t += nt;
        if (pieces) { const Piece &q = pc[i0 + i]; h_pb[i] = q.pos_base; h_em[2 * i] = q.emit_lo; h_em[2 * i + 1] = q.emit_hi; }
    }
    h_offs[n] = n ? rec_hi(i1 - 1) - origin : 0;
    h_ft[n] = t;
    if (flag_words) memcpy(hm + ml.flags, in.flags + (rel >> 11), flag_words * 4);
    if (e1 > e0) {
        ExcRec *he = (ExcRec *)(hm + ml.exc);
        for (uint64_t e = e0; e < e1; e++) {     // slice coordinates
            const uint64_t st = std::max(in.exc[e].start, rel), en = in.exc[e].start + in.exc[e].len;
            he[e - e0] = ExcRec{st - rel, (uint32_t)(en - st), in.exc[e].byte};
        }
    }
    const uint64_t span = hi - origin;            // bases of the slice
    uint8_t *dm = (uint8_t *)s.d_meta.p;
    bd = BatchDev{};
    bd.n = n; bd.n_tiles = n_tiles; bd.bases = span; bd.sc = s.d_sc.as<BatchScalars>();
    bd.offs = (const uint64_t *)(dm + ml.offs); bd.first_tile = (const uint32_t *)(dm + ml.first_tile);
    bd.tile_seq = (const uint32_t *)(dm + ml.tile_seq);
    if (pieces) { bd.pos_base = (const uint32_t *)(dm + ml.pos_base); bd.emit = (const uint32_t *)(dm + ml.emit); }
    {
        StageTimer tm(c, "h2d", c->copy_stream);
        if (in.resident) {
            if (in.packed) {
                bd.packed = in.words + (origin >> 4); bd.flags = in.flags + (origin >> 11);
                bd.exc = (const ExcRec *)in.exc; bd.n_exc = (uint32_t)in.n_exc; bd.exc_base = origin;   // the whole list, absolute coordinates
            } else bd.seqs = in.seqs + origin;
        } else if (in.packed) {
            const size_t wbytes = (size_t)((span + 15) / 16) * 4;
            if ((rc = ensure(c, s.d_in, wbytes + PAD))) return rc;
            if (wbytes) CK(cudaMemcpyAsync(s.d_in.p, in.words + (rel >> 4), wbytes, cudaMemcpyHostToDevice, c->copy_stream));
            c->ctr_h2d_bytes += wbytes;
            CK(cudaMemsetAsync((uint8_t *)s.d_in.p + wbytes, 0, PAD, c->copy_stream));
            bd.packed = s.d_in.as<uint32_t>(); bd.flags = (const uint32_t *)(dm + ml.flags);
            bd.exc = (const ExcRec *)(dm + ml.exc); bd.n_exc = (uint32_t)(e1 - e0);
        } else {
            if ((rc = ensure(c, s.d_in, span + PAD))) return rc;
            if (span) CK(cudaMemcpyAsync(s.d_in.p, in.seqs + origin, span, cudaMemcpyHostToDevice, c->copy_stream));
            c->ctr_h2d_bytes += span;
            CK(cudaMemsetAsync((uint8_t *)s.d_in.p + span, 0, PAD, c->copy_stream));
            bd.seqs = s.d_in.as<uint8_t>();
        }
        CK(cudaMemcpyAsync(s.d_meta.p, s.h_meta.p, ml.total, cudaMemcpyHostToDevice, c->copy_stream));
        c->ctr_h2d_bytes += ml.total;
    }
    CK(cudaEventRecord(s.ev_copied, c->copy_stream));
    s.bd = bd; s.i0 = i0; s.i1 = i1;
    return MQ_OK;
}

// after a status record reported an overflow: make room for what it asked for
int grow_from_status(mq_ctx *c, const BatchScalars &sc, uint64_t bases) {
    if (sc.flags & BS_RANGE) { c->err = "sub-batch holds 2^32 minimizers or more"; return MQ_ERR_RANGE; }
    CK(cudaStreamSynchronize(c->stream));
    if (sc.flags & BS_OVF_CAP) {
        c->ovf_cap = sc.ovf_count + sc.ovf_count / 4 + 1024;
        dfree(c->d_ovf_tile); dfree(c->d_ovf_meta); dfree(c->d_ovf_hash);
    }
    if (sc.flags & BS_MINI_CAP) {
        const double rate = (double)sc.n_minimizers / (double)std::max<uint64_t>(bases, 1);
        c->mini_rate = std::max(c->mini_rate, rate * 1.25 + 0.001);
    }
    return MQ_OK;
}

int check_offs(mq_ctx *c, const uint64_t *offs, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        if (offs[i + 1] < offs[i]) { c->err = "offs not monotone"; return MQ_ERR_ARG; }
        if (offs[i + 1] - offs[i] >= (1ull << 31)) { c->err = "record of 2^31 bases or more"; return MQ_ERR_RANGE; }
    }
    return MQ_OK;
}

// ---- on-the-fly packing of ASCII host input -------------------------------------------------------------
// mq_map_batch on ASCII host buffers is bound by the PCIe link (1 byte per base).  With host threads at its disposal
// (mq_set_host_threads: the reference's --threads, main.rs:138-141) the pipeline feeds the GPU from BOTH ends of the batch:
// the main thread uploads ASCII sub-batches from the front as fast as the link takes them, while the host threads pack
// sub-batches from the back, 2 bits per base, into pinned staging buffers (a team: every thread packs 1/T of the
// sub-batch, so one is ready every couple of milliseconds) that are uploaded at a quarter of the bytes.  Whoever gets
// to a sub-batch first takes it; the two fronts meet wherever link and host threads balance.  Results do not depend
// on which route a read took (the packed format is exactly the ASCII sequence, tests/test_pack.py).
struct HostPacker {
    struct Job { uint32_t sub; int buf; int done = 0; uint64_t base = 0, n_bases = 0; std::vector<std::vector<mq_exc>> part_exc; };
    mq_ctx *c; const uint8_t *seqs; const uint64_t *offs; const std::vector<uint32_t> &cut;
    int T;
    std::mutex mu; std::condition_variable cv;
    std::vector<Job> jobs;                 // in claim order (reserved: references stay valid)
    std::vector<size_t> ready;             // job indices, oldest first
    size_t front = 0, back;                // unclaimed sub-batches: [front, back)
    size_t packing = 0;                    // jobs claimed and not yet ready
    bool stop = false, failed = false;
    std::vector<std::thread> th;

    HostPacker(mq_ctx *c_, const uint8_t *seqs_, const uint64_t *offs_, const std::vector<uint32_t> &cut_, int T_)
        : c(c_), seqs(seqs_), offs(offs_), cut(cut_), T(T_), back(cut_.size() - 1) { jobs.reserve(cut_.size()); }
    ~HostPacker() { finish(); }
    void start() { for (int t = 0; t < T; t++) th.emplace_back([this, t] { run(t); }); }
    void finish() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto &x : th) x.join();
        th.clear();
    }
    void run(int t) {
        for (size_t seq = 0;; seq++) {
            Job *job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                for (;;) {
                    if (stop || failed) return;
                    if (jobs.size() > seq) break;
                    if (front >= back) return;                     // nothing left to claim
                    int b = -1;
                    for (int i = 0; i < N_PACK_BUFS; i++) if (c->pack_buf[i].state == 0) { b = i; break; }
                    if (b >= 0) {
                        back--;
                        Job j; j.sub = (uint32_t)back; j.buf = b;
                        j.base = offs[cut[back]] & ~2047ull; j.n_bases = offs[cut[back + 1]] - j.base;
                        j.part_exc.resize(T);
                        jobs.push_back(std::move(j));
                        c->pack_buf[b].state = 1; packing++;
                        cv.notify_all();
                        break;
                    }
                    cv.wait(lk);                                   // every staging buffer is in use
                }
                job = &jobs[seq];
            }
            PackBuf &pb = c->pack_buf[job->buf];
            const uint64_t units = (job->n_bases + 2047) / 2048;
            bool ok = true;
            try {
                mq_pack_units(seqs + job->base, job->n_bases, units * t / T, units * (t + 1) / T, (uint32_t *)pb.words.p, (uint32_t *)pb.flags.p,
                              job->part_exc[t], false);
            } catch (...) { ok = false; }
            bool wake = false;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (!ok) { failed = true; wake = true; }
                if (++job->done == T) {                            // the last thread out hands the sub-batch over
                    pb.exc.clear();
                    for (auto &v : job->part_exc) pb.exc.insert(pb.exc.end(), v.begin(), v.end());
                    ((uint32_t *)pb.flags.p)[units] = 0; ((uint32_t *)pb.flags.p)[units + 1] = 0;     // the stager reads one word past the slice
                    pb.state = 2; packing--;
                    ready.push_back(seq);
                    wake = true;
                }
            }
            if (wake) cv.notify_all();
        }
    }
    // next sub-batch for the uploader: a packed one if one is ready, else the next ASCII one from the front.
    // Returns false when every sub-batch has been handed out.  *job_out = nullptr for an ASCII sub-batch.
    bool next(uint32_t *sub, const Job **job_out) {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            if (!ready.empty()) {
                const Job &j = jobs[ready.front()];
                ready.erase(ready.begin());
                c->pack_buf[j.buf].state = 3;
                *sub = j.sub; *job_out = &j;
                return true;
            }
            if (front < back) { *sub = (uint32_t)front++; *job_out = nullptr; return true; }
            if (packing == 0 || failed) return false;
            cv.wait(lk);
        }
    }
    void release(int buf) {
        { std::lock_guard<std::mutex> lk(mu); c->pack_buf[buf].state = 0; }
        cv.notify_all();
    }
};

// ---- mapping pipeline -------------------------------------------------------------------------------------
// reads [0, n) of `in` with host offsets `offs`; hits go to out (host) or d_out (device, resident path)
int map_pipeline(mq_ctx *c, const SeqInput &in, const uint64_t *offs, uint32_t n, mq_hit *out, mq_hit *d_out) {
    int rc;
    if ((rc = init_streams(c))) return rc;
    c->ctr_h2d_bytes = c->ctr_host_packed_bases = c->ctr_host_packed_subs = c->ctr_subs = 0;
    for (auto &s : c->slot) s.pack_buf = -1;         // (a call that failed half-way may have left one set)
    // sub-batch boundaries; from the host the first one is an eighth of the size so that the GPU starts while the next is
    // still being uploaded (resident input: nothing to wait for, full size from the start)
    std::vector<uint32_t> cut{0};
    const uint64_t sub = in.resident ? c->sub_bases_resident : (in.packed ? c->sub_bases_light : c->sub_bases);
    for (uint32_t i0 = 0; i0 < n;) {
        const uint64_t want = offs[i0] + ((i0 == 0 && !in.resident) ? sub / 8 : sub);
        uint32_t i1 = (uint32_t)(std::upper_bound(offs + i0 + 1, offs + n + 1, want) - offs) - 1;
        if (i1 <= i0) i1 = i0 + 1;
        cut.push_back(i1); i0 = i1;
    }
    const size_t ns = cut.size() - 1;
    const uint32_t min_len = c->p.l + c->p.k - 1;
    c->ctr_subs = ns;

    // host threads pack sub-batches from the back while the link carries ASCII ones from the front
    std::unique_ptr<HostPacker> hp;
    if (!in.packed && !in.resident && c->host_threads > 0 && ns >= 3) {
        uint64_t maxb = 0;
        for (size_t i = 0; i < ns; i++) maxb = std::max<uint64_t>(maxb, offs[cut[i + 1]] - (offs[cut[i]] & ~2047ull));
        for (auto &pb : c->pack_buf) {
            if ((rc = ensure_host(c, pb.words, (size_t)mq_packed_words(maxb) * 4))) return rc;
            if ((rc = ensure_host(c, pb.flags, (size_t)mq_packed_flag_words(maxb) * 4))) return rc;
            pb.state = 0;
        }
        hp.reset(new HostPacker(c, in.seqs, offs, cut, c->host_threads));
        hp->start();
    }

    // finish the sub-batch a slot holds: wait for its hits, look at its status, redo it if a buffer was too small
    auto retire = [&](Slot2 &s) -> int {
        if (!s.busy) return MQ_OK;
        if (hp) {
            // the packers want this core: poll with short sleeps instead of spinning inside cudaEventSynchronize (a
            // blocking-sync event would do, but its wake-up took milliseconds on the virtualised hosts this ran on)
            cudaError_t q;
            while ((q = cudaEventQuery(s.ev_done)) == cudaErrorNotReady) std::this_thread::sleep_for(std::chrono::microseconds(20));
            CK(q);
        } else CK(cudaEventSynchronize(s.ev_done));
        s.busy = false;
        const BatchScalars *hs = (const BatchScalars *)s.h_sc.p;
        c->last_minimizers += hs->n_minimizers;
        for (int attempt = 0; hs->flags; attempt++) {
            if (attempt == 3) { c->err = "sub-batch kept overflowing its buffers"; return MQ_ERR_NOMEM; }
            int r2;
            c->last_minimizers -= hs->n_minimizers;
            if ((r2 = grow_from_status(c, *hs, s.bd.bases))) return r2;
            // the slot still holds the inputs and its device view: run it again, synchronously
            const BatchDev &bd = s.bd;
            const uint32_t m = s.i1 - s.i0;
            if ((r2 = ensure_workspace(c, m, bd.n_tiles, bd.bases, true))) return r2;
            HitRec *dh = d_out ? (HitRec *)d_out + s.i0 : s.d_hits.as<HitRec>();
            if ((r2 = enqueue_scan(c, bd))) return r2;
            if ((r2 = enqueue_probe_chain(c, bd, dh))) return r2;
            if (!d_out) CK(cudaMemcpyAsync(out + s.i0, dh, (size_t)m * sizeof(HitRec), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(s.h_sc.p, s.d_sc.p, sizeof(BatchScalars), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            c->last_minimizers += hs->n_minimizers;
        }
        if (s.pack_buf >= 0 && hp) { hp->release(s.pack_buf); s.pack_buf = -1; }
        return MQ_OK;
    };
    // a staging buffer of the packers is free again as soon as its upload has finished (long before the slot retires)
    auto release_uploaded = [&]() {
        if (!hp) return;
        for (auto &s : c->slot)
            if (s.pack_buf >= 0 && cudaEventQuery(s.ev_copied) == cudaSuccess) { hp->release(s.pack_buf); s.pack_buf = -1; }
        cudaGetLastError();
    };

    for (size_t i = 0; i < ns; i++) {
        Slot2 &s = c->slot[i % N_SLOTS];
        if ((rc = retire(s))) return rc;
        release_uploaded();
        uint32_t j = (uint32_t)i;                    // the sub-batch this iteration stages
        const HostPacker::Job *job = nullptr;
        if (hp && !hp->next(&j, &job)) { c->err = "packing the input on the host failed (out of memory)"; return MQ_ERR_NOMEM; }
        const uint32_t m = cut[j + 1] - cut[j];
        BatchDev bd;
        if (job) {
            const PackBuf &pb = c->pack_buf[job->buf];
            SeqInput pin; pin.packed = true; pin.base = job->base;
            pin.words = (const uint32_t *)pb.words.p; pin.flags = (const uint32_t *)pb.flags.p; pin.exc = pb.exc.data(); pin.n_exc = pb.exc.size();
            s.pack_buf = job->buf;
            c->ctr_host_packed_bases += offs[cut[j + 1]] - offs[cut[j]]; c->ctr_host_packed_subs++;
            rc = stage_pieces(c, s, pin, nullptr, offs, cut[j], cut[j + 1], min_len, false, bd);
        } else rc = stage_pieces(c, s, in, nullptr, offs, cut[j], cut[j + 1], min_len, false, bd);
        if (rc) return rc;
        if ((rc = ensure_workspace(c, m, bd.n_tiles, bd.bases, true))) return rc;
        HitRec *dh;
        if (d_out) dh = (HitRec *)d_out + cut[j];
        else { if ((rc = ensure(c, s.d_hits, (size_t)m * sizeof(HitRec)))) return rc; dh = s.d_hits.as<HitRec>(); }
        CK(cudaStreamWaitEvent(c->stream, s.ev_copied, 0));
        if ((rc = enqueue_scan(c, bd))) return rc;
        if ((rc = enqueue_probe_chain(c, bd, dh))) return rc;
        CK(cudaEventRecord(s.ev_comp, c->stream));
        CK(cudaStreamWaitEvent(c->d2h_stream, s.ev_comp, 0));
        {
            StageTimer t(c, "d2h", c->d2h_stream);
            if (!d_out) CK(cudaMemcpyAsync(out + cut[j], dh, (size_t)m * sizeof(HitRec), cudaMemcpyDeviceToHost, c->d2h_stream));
            CK(cudaMemcpyAsync(s.h_sc.p, s.d_sc.p, sizeof(BatchScalars), cudaMemcpyDeviceToHost, c->d2h_stream));
        }
        CK(cudaEventRecord(s.ev_done, c->d2h_stream));
        s.busy = true;
    }
    for (auto &s : c->slot) if ((rc = retire(s))) return rc;
    return MQ_OK;
}

// ---- index build ---------------------------------------------------------------------------------------
int store_append(mq_ctx *c, uint64_t M) {
    int rc;
    if ((rc = ensure_keep(c, c->st_pos, (c->st_n + M + 64) * 4, c->st_n * 4))) return rc;
    if ((rc = ensure_keep(c, c->st_hash, (c->st_n + M + 64) * 8, c->st_n * 8))) return rc;
    if (M) {
        CK(cudaMemcpyAsync(c->st_pos.as<uint32_t>() + c->st_n, c->d_pos.p, M * 4, cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(c->st_hash.as<uint64_t>() + c->st_n, c->d_hash.p, M * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->st_n += M;
    return MQ_OK;
}

// scan one staged batch to completion (index build and introspection paths): returns M and, if wanted, the per-record
// minimizer offsets (n+1).  Waits for the GPU; redoes the batch if a buffer was too small.
int scan_sync(mq_ctx *c, Slot2 &s, const BatchDev &bd, uint64_t *M_out, std::vector<uint32_t> *seq_off) {
    int rc;
    BatchScalars *hs = (BatchScalars *)s.h_sc.p;
    for (int attempt = 0;; attempt++) {
        if ((rc = ensure_workspace(c, bd.n, bd.n_tiles, bd.bases, false))) return rc;
        if ((rc = ensure(c, c->d_seq_off, ((size_t)bd.n + 2) * 4))) return rc;
        CK(cudaStreamWaitEvent(c->stream, s.ev_copied, 0));
        if ((rc = enqueue_scan(c, bd))) return rc;
        if (bd.n_tiles) {
            k_seq_mini_off<<<(bd.n + 1 + 255) / 256, 256, 0, c->stream>>>(bd.first_tile, TileBase{c->d_tile_cnt.as<uint32_t>(), c->d_blk.as<uint32_t>()},
                                                                        bd.n, c->d_seq_off.as<uint32_t>());
            c->launches++;
        } else CK(cudaMemsetAsync(c->d_seq_off.p, 0, ((size_t)bd.n + 1) * 4, c->stream));
        CK(cudaMemcpyAsync(hs, bd.sc, sizeof(BatchScalars), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (!hs->flags) break;
        if (attempt == 3) { c->err = "batch kept overflowing its buffers"; return MQ_ERR_NOMEM; }
        if ((rc = grow_from_status(c, *hs, bd.bases))) return rc;
    }
    *M_out = hs->n_minimizers;
    c->last_minimizers += hs->n_minimizers;
    if (seq_off) {
        seq_off->resize((size_t)bd.n + 1);
        CK(cudaMemcpyAsync(seq_off->data(), c->d_seq_off.p, ((size_t)bd.n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return MQ_OK;
}

// number of bases after `from` needed to see `need` further run starts (or the end of the record)
uint64_t halo_end(const SeqInput &in, uint64_t rec_lo, uint64_t rec_hi, uint64_t from, uint32_t need, bool hpc) {
    if (!hpc) return std::min(rec_hi, from + need);
    auto byte_at = [&](uint64_t i) -> uint32_t {
        if (!in.packed) return in.seqs[i];
        uint32_t b = (in.words[i >> 4] >> (2 * (i & 15))) & 3u;
        if ((in.flags[i >> 11] >> ((i >> 6) & 31)) & 1u) {
            const mq_exc *e = std::partition_point(in.exc, in.exc + in.n_exc, [&](const mq_exc &x) { return x.start + x.len <= i; });
            if (e != in.exc + in.n_exc && e->start <= i) return 0x100u | e->byte;
        }
        return b;
    };
    uint64_t i = from;
    uint32_t prev = from > rec_lo ? byte_at(from - 1) : 0xFFFFu;
    while (i < rec_hi && need) { const uint32_t b = byte_at(i); if (b != prev) need--; prev = b; i++; }
    // round up generously: whole 64-base blocks, so the scan's halo loop never runs out before the record does
    return std::min<uint64_t>(rec_hi, (i + 63) & ~(uint64_t)63);
}

// Scan pieces of reference records into the minimizer store.  Long records are cut into sub-pieces of <= ADD_BASES
// bases so that uploads overlap scans; `ref_len_of(ref_idx)` is only used for validation by the callers.
int add_pieces(mq_ctx *c, const SeqInput &in, std::vector<Piece> &pcs, std::vector<uint64_t> *counts) {
    int rc;
    if ((rc = init_streams(c))) return rc;
    if (counts) counts->assign(pcs.size(), 0);
    // batches: consecutive whole-record pieces share a batch while contiguous and small; a segment piece is alone
    std::vector<std::array<size_t, 2>> bt;
    for (size_t i = 0; i < pcs.size();) {
        size_t j = i + 1;
        if (pcs[i].emit_hi == 0) while (j < pcs.size() && pcs[j].emit_hi == 0 && pcs[j].lo == pcs[j - 1].hi && pcs[j].hi - pcs[i].lo <= ADD_BASES && j - i < (1u << 20)) j++;
        bt.push_back({i, j}); i = j;
    }
    auto stage = [&](size_t b) -> int {
        const bool seg = pcs[bt[b][0]].emit_hi != 0;
        BatchDev bd;
        return stage_pieces(c, c->slot[b & 1], in, pcs.data(), nullptr, (uint32_t)bt[b][0], (uint32_t)bt[b][1], seg ? 0 : c->p.l + c->p.k - 1, seg, bd);
    };
    if (!bt.empty() && (rc = stage(0))) return rc;
    for (size_t b = 0; b < bt.size(); b++) {
        // upload of the next batch (copy stream) overlaps the scan of this one
        if (b + 1 < bt.size() && (rc = stage(b + 1))) return rc;
        Slot2 &s = c->slot[b & 1];
        uint64_t M = 0; std::vector<uint32_t> so;
        if ((rc = scan_sync(c, s, s.bd, &M, &so))) return rc;
        if ((rc = store_append(c, M))) return rc;
        for (size_t q = bt[b][0]; q < bt[b][1]; q++) {
            const uint64_t cnt = so[q - bt[b][0] + 1] - so[q - bt[b][0]];
            c->dir.push_back({(uint64_t)pcs[q].ref_idx, pcs[q].seg_start, cnt});
            if (counts) (*counts)[q] = cnt;
        }
    }
    CK(cudaStreamSynchronize(c->stream));
    return MQ_OK;
}

// cut record [lo, hi) (caller-array coordinates) of reference ref_idx, restricted to record positions
// [own_lo, own_hi), into pieces of <= ADD_BASES bases
void pieces_of_range(const mq_ctx *c, const SeqInput &in, uint64_t lo, uint64_t hi, uint32_t ref_idx, uint64_t own_lo, uint64_t own_hi,
                     std::vector<Piece> &out) {
    const uint64_t len = hi - lo;
    if (own_lo == 0 && own_hi == len && len <= ADD_BASES) { out.push_back(Piece{lo, hi, ref_idx, 0, 0, 0, 0}); return; }
    for (uint64_t s0 = own_lo; s0 < own_hi; s0 += ADD_BASES) {
        const uint64_t s1 = std::min(own_hi, s0 + ADD_BASES);
        const uint64_t ctx = s0 > 0 ? 1 : 0;
        const uint64_t he = halo_end(in, lo, hi, lo + s1, c->p.l - 1, c->p.use_hpc != 0);
        out.push_back(Piece{lo + s0 - ctx, he, ref_idx, (uint32_t)(s0 - ctx), (uint32_t)ctx, (uint32_t)(ctx + s1 - s0), s0});
    }
}

int freeze_local(mq_ctx *c, const uint64_t *ref_lens, uint32_t n_refs);

}  // namespace

// =================================================================================================
extern "C" {

int mq_abi_version(void) { return 2; }

const char *mq_strerror(int code) {
    switch (code) {
        case MQ_OK: return "ok";
        case MQ_ERR_ARG: return "bad argument";
        case MQ_ERR_CUDA: return "CUDA error or no usable device (there is no CPU fallback)";
        case MQ_ERR_STATE: return "call out of order";
        case MQ_ERR_NOMEM: return "out of memory";
        case MQ_ERR_RANGE: return "input exceeds a documented limit";
        default: return "unknown error";
    }
}
const char *mq_last_error(const mq_ctx *c) { return c ? c->err.c_str() : ""; }

static int create_one(mq_ctx **out, const mq_params *p, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) { cudaGetLastError(); return MQ_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return MQ_ERR_CUDA; }
    mq_ctx *c = new (std::nothrow) mq_ctx();
    if (!c) return MQ_ERR_NOMEM;
    c->p = *p; c->device = device; c->bound = hash_bound(p->density);
    if (const char *e = getenv("MQ_SUB_BASES")) {        // tests: many small sub-batches from small inputs
        const uint64_t v = strtoull(e, nullptr, 10);
        if (v >= 4096) c->sub_bases = c->sub_bases_light = c->sub_bases_resident = v;
    }
    if (const char *e = getenv("MQ_SUB_BASES_RESIDENT")) { const uint64_t v = strtoull(e, nullptr, 10); if (v >= 4096) c->sub_bases_resident = v; }   // tuning runs
    if (const char *e = getenv("MQ_SUB_BASES_PACKED")) { const uint64_t v = strtoull(e, nullptr, 10); if (v >= 4096) c->sub_bases_light = v; }
    fill_tables(c->tab, p->l);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->n_sm = prop.multiProcessorCount;
    // random 32-byte probes into a multi-GB table: fetch one sector per miss, not two
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32); cudaGetLastError();      // a hint; the device may keep its own
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MQ_ERR_CUDA; }
    uint32_t warm[256][4];
    fill_warm_table(c->tab, warm);
    if (cudaMalloc(&c->d_warm.p, sizeof warm) != cudaSuccess || cudaMemcpy(c->d_warm.p, warm, sizeof warm, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError(); cudaStreamDestroy(c->stream); delete c; return MQ_ERR_CUDA;
    }
    c->d_warm.cap = sizeof warm;
    *out = c;
    return MQ_OK;
}

int mq_create(mq_ctx **out, const mq_params *p, int device) {
    if (!out || !p) return MQ_ERR_ARG;
    *out = nullptr;
    if (p->l < 2 || p->l > MQ_MAX_L || p->k < 1 || p->k > MQ_MAX_K || !(p->density >= 0.0)) return MQ_ERR_ARG;
    try { return create_one(out, p, device); } catch (...) { return MQ_ERR_NOMEM; }
}

int mq_create_multi(mq_ctx **out, const mq_params *p, const int *devices, int n_devices) {
    if (!out || !p || !devices || n_devices < 1) return MQ_ERR_ARG;
    *out = nullptr;
    if (p->l < 2 || p->l > MQ_MAX_L || p->k < 1 || p->k > MQ_MAX_K || !(p->density >= 0.0)) return MQ_ERR_ARG;
    if (n_devices == 1) return mq_create(out, p, devices[0]);
    try {
        mq_ctx *par = new mq_ctx();
        par->p = *p; par->device = devices[0]; par->bound = hash_bound(p->density);
        for (int i = 0; i < n_devices; i++) {
            mq_ctx *k = nullptr;
            int rc = create_one(&k, p, devices[i]);
            if (rc) { mq_destroy(par); return rc; }
            par->kids.push_back(k);
        }
        // peer access lets the store exchange go GPU to GPU over NVLink; without it cudaMemcpyPeer stages through the host
        for (int i = 0; i < n_devices; i++) {
            cudaSetDevice(devices[i]);
            for (int j = 0; j < n_devices; j++) if (devices[i] != devices[j]) {     // (a device may be listed twice: two contexts on it)
                int can = 0; cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
                if (can) { cudaDeviceEnablePeerAccess(devices[j], 0); cudaGetLastError(); }
            }
        }
        *out = par;
        return MQ_OK;
    } catch (...) { return MQ_ERR_NOMEM; }
}
int mq_device_count(const mq_ctx *c) { return !c ? 0 : (c->kids.empty() ? 1 : (int)c->kids.size()); }

void mq_destroy(mq_ctx *c) {
    if (!c) return;
    for (mq_ctx *k : c->kids) mq_destroy(k);
    if (!c->stream) { delete c; return; }            // multi-GPU parent: owns no device state
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
    timers_collect(c);
    DBuf *bufs[] = {&c->d_ev_hash, &c->d_ev_meta, &c->d_lane_cnt, &c->d_tile_cnt, &c->d_blk, &c->d_ovf_tile, &c->d_ovf_meta, &c->d_ovf_hash,
                    &c->d_pos, &c->d_hash, &c->d_seq_off, &c->d_matches, &c->d_nmatch, &c->d_big_list, &c->d_misc, &c->st_pos, &c->st_hash,
                    &c->d_table, &c->d_ref_lens, &c->d_warm, &c->d_bloom};
    for (DBuf *b : bufs) dfree(*b);
    for (auto &s : c->slot) {
        dfree(s.d_in); dfree(s.d_meta); dfree(s.d_hits); dfree(s.d_sc); hfree(s.h_meta); hfree(s.h_sc);
        if (s.ev_copied) cudaEventDestroy(s.ev_copied);
        if (s.ev_comp) cudaEventDestroy(s.ev_comp);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
    }
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->region_a) { cudaEventDestroy(c->region_a); cudaEventDestroy(c->region_b); }
    hfree(c->h_pin);
    for (auto &pb : c->pack_buf) { hfree(pb.words); hfree(pb.flags); }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

void *mq_host_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void mq_host_free(void *p) { if (p) cudaFreeHost(p); }

#define FIRST(c) ((c)->kids.empty() ? (c) : (c)->kids[0])
void *mq_stream(mq_ctx *c) { return c ? (void *)FIRST(c)->stream : nullptr; }
int mq_sync(mq_ctx *c) {
    if (!c) return MQ_ERR_ARG;
    if (!c->kids.empty()) { for (mq_ctx *k : c->kids) { int rc = mq_sync(k); if (rc) { c->err = k->err; return rc; } } return MQ_OK; }
    cudaSetDevice(c->device); CK(cudaStreamSynchronize(c->stream)); timers_collect(c); return MQ_OK;
}
uint64_t mq_launch_count(mq_ctx *c) {
    if (!c) return 0;
    uint64_t s = c->launches; for (mq_ctx *k : c->kids) s += k->launches; return s;
}
double mq_total_ms(mq_ctx *c, const char *stage) {
    if (!c || !stage) return -1.0;
    if (!c->kids.empty()) { double m = 0; for (mq_ctx *k : c->kids) m = std::max(m, mq_total_ms(k, stage)); return m; }
    cudaSetDevice(c->device);
    timers_collect(c);
    auto it = c->ms_total.find(stage);
    return it == c->ms_total.end() ? 0.0 : it->second;
}
double mq_last_ms(mq_ctx *c, const char *stage) {
    if (!c || !stage) return -1.0;
    if (!c->kids.empty()) { double m = 0; for (mq_ctx *k : c->kids) m = std::max(m, mq_last_ms(k, stage)); return m; }
    cudaSetDevice(c->device);
    timers_collect(c);
    if (!strcmp(stage, "total")) { double s = 0; for (auto &kv : c->ms) if (kv.first != "scan_kernel") s += kv.second; return s; }
    auto it = c->ms.find(stage);
    return it == c->ms.end() ? 0.0 : it->second;
}
int mq_set_host_threads(mq_ctx *c, int n) {
    if (!c || n < 0) return MQ_ERR_ARG;
    if (n > 256) n = 256;
    c->host_threads = n;
    const int G = (int)c->kids.size();
    for (int g = 0; g < G; g++) c->kids[g]->host_threads = n / G + (g < n % G ? 1 : 0);      // a multi-GPU context shares them out
    return MQ_OK;
}
uint64_t mq_last_counter(mq_ctx *c, const char *name) {
    if (!c || !name) return 0;
    auto one = [&](const mq_ctx *k) -> uint64_t {
        if (!strcmp(name, "h2d_bytes")) return k->ctr_h2d_bytes;
        if (!strcmp(name, "host_packed_bases")) return k->ctr_host_packed_bases;
        if (!strcmp(name, "host_packed_sub_batches")) return k->ctr_host_packed_subs;
        if (!strcmp(name, "sub_batches")) return k->ctr_subs;
        return 0;
    };
    if (c->kids.empty()) return one(c);
    uint64_t s = 0;
    for (const mq_ctx *k : c->kids) s += one(k);
    return s;
}
uint64_t mq_scan_kernel_launches(mq_ctx *c) { if (!c) return 0; uint64_t s = c->scan_kernel_launches; for (mq_ctx *k : c->kids) s += k->scan_kernel_launches; return s; }
uint64_t mq_minimizer_count(mq_ctx *c, int reset) {
    if (!c) return 0;
    uint64_t v = c->last_minimizers; if (reset) c->last_minimizers = 0;
    for (mq_ctx *k : c->kids) v += mq_minimizer_count(k, reset);
    return v;
}

// device-memory helpers so that callers can keep inputs resident in HBM without another runtime (single-GPU contexts)
void *mq_dev_alloc(mq_ctx *c, size_t bytes) {
    if (!c || !c->kids.empty()) return nullptr;
    cudaSetDevice(c->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void mq_dev_free(mq_ctx *c, void *p) { if (c && p) { cudaSetDevice(c->device); cudaFree(p); } }
int mq_dev_upload(mq_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || !c->kids.empty() || (bytes && (!dst || !src))) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MQ_OK;
}
int mq_dev_download(mq_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || !c->kids.empty() || (bytes && (!dst || !src))) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MQ_OK;
}
int mq_dev_memset(mq_ctx *c, void *dst, int value, size_t bytes) {
    if (!c || !c->kids.empty() || (bytes && !dst)) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemsetAsync(dst, value, bytes, c->stream));
    return MQ_OK;
}
// CUDA-event bracket on the ctx stream around any sequence of calls
int mq_region_begin(mq_ctx *c) {
    if (!c || !c->kids.empty()) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    if (!c->region_a) { CK(cudaEventCreate(&c->region_a)); CK(cudaEventCreate(&c->region_b)); }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->region_a, c->stream));
    return MQ_OK;
}
double mq_region_end_ms(mq_ctx *c) {
    if (!c || !c->region_a) return -1.0;
    cudaSetDevice(c->device);
    if (cudaEventRecord(c->region_b, c->stream) != cudaSuccess) return -1.0;
    if (cudaEventSynchronize(c->region_b) != cudaSuccess) return -1.0;
    float ms = 0; cudaEventElapsedTime(&ms, c->region_a, c->region_b);
    return ms;
}

uint64_t mq_table_bytes(mq_ctx *c) { if (!c) return 0; c = FIRST(c); return c->frozen ? (c->tmask + 2) * sizeof(Slot) : 0; }
uint64_t mq_table_slots(mq_ctx *c) { if (!c) return 0; c = FIRST(c); return c->frozen ? c->tmask + 1 : 0; }

// ---- index build ---------------------------------------------------------------------------------
static int index_add_impl(mq_ctx *c, const SeqInput &in, const uint64_t *offs, uint32_t n, uint32_t first_ref_idx, uint64_t *nb_mers_out) {
    if (c->kids.empty() ? c->frozen : c->kids[0]->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    if (n == 0) return MQ_OK;
    const uint32_t min_len = c->p.l + c->p.k - 1;
    if (c->kids.empty()) {
        cudaSetDevice(c->device);
        timers_reset(c);
        std::vector<Piece> pcs; std::vector<uint32_t> owner;       // owner[piece] = record
        for (uint32_t i = 0; i < n; i++) {
            const size_t before = pcs.size();
            const uint64_t len = offs[i + 1] - offs[i];
            if (len < min_len) pcs.push_back(Piece{offs[i], offs[i + 1], first_ref_idx + i, 0, 0, 0, 0});   // mers.rs:18: yields nothing (no tiles)
            else pieces_of_range(c, in, offs[i], offs[i + 1], first_ref_idx + i, 0, len, pcs);
            for (size_t q = before; q < pcs.size(); q++) owner.push_back(i);
        }
        std::vector<uint64_t> counts;
        if ((rc = add_pieces(c, in, pcs, &counts))) return rc;
        if (nb_mers_out) {
            std::vector<uint64_t> per(n, 0);
            for (size_t q = 0; q < pcs.size(); q++) per[owner[q]] += counts[q];
            for (uint32_t i = 0; i < n; i++) nb_mers_out[i] = per[i] >= c->p.k ? per[i] - c->p.k + 1 : 0;
        }
        timers_collect(c);
        return MQ_OK;
    }
    // multi-GPU: every record is cut into one base range per device (closures.rs:85 parallelises over records, whose
    // sizes are too unequal for that to balance); the device threads scan concurrently
    const size_t G = c->kids.size();
    std::vector<std::vector<Piece>> pcs(G);
    std::vector<std::vector<uint32_t>> owner(G);
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t len = offs[i + 1] - offs[i];
        for (size_t g = 0; g < G; g++) {
            const size_t before = pcs[g].size();
            if (len < min_len) { if (g == 0) pcs[g].push_back(Piece{offs[i], offs[i + 1], first_ref_idx + i, 0, 0, 0, 0}); }
            else {
                const uint64_t a = len * g / G, b = len * (g + 1) / G;
                if (b > a) pieces_of_range(c, in, offs[i], offs[i + 1], first_ref_idx + i, a, b, pcs[g]);
            }
            for (size_t q = before; q < pcs[g].size(); q++) owner[g].push_back(i);
        }
    }
    std::vector<int> rcs(G, 0);
    std::vector<std::vector<uint64_t>> counts(G);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++) th.emplace_back([&, g]() {
        mq_ctx *k = c->kids[g];
        cudaSetDevice(k->device);
        timers_reset(k);
        rcs[g] = add_pieces(k, in, pcs[g], &counts[g]);
        timers_collect(k);
    });
    for (auto &t : th) t.join();
    for (size_t g = 0; g < G; g++) if (rcs[g]) { c->err = c->kids[g]->err; return rcs[g]; }
    if (nb_mers_out) {
        std::vector<uint64_t> per(n, 0);
        for (size_t g = 0; g < G; g++) for (size_t q = 0; q < pcs[g].size(); q++) per[owner[g][q]] += counts[g][q];
        for (uint32_t i = 0; i < n; i++) nb_mers_out[i] = per[i] >= c->p.k ? per[i] - c->p.k + 1 : 0;
    }
    return MQ_OK;
}

int mq_index_add(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint32_t first_ref_idx, uint64_t *nb_mers_out) {
    if (!c || (!seqs && n) || !offs) return MQ_ERR_ARG;
    try {
        SeqInput in; in.seqs = seqs;
        return index_add_impl(c, in, offs, n, first_ref_idx, nb_mers_out);
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}
int mq_index_add_packed(mq_ctx *c, const mq_packed *pk, const uint64_t *offs, uint32_t n, uint32_t first_ref_idx, uint64_t *nb_mers_out) {
    if (!c || !pk || !offs || (n && (!pk->words || !pk->flags)) || (pk->n_exc && !pk->exc)) return MQ_ERR_ARG;
    try {
        SeqInput in; in.packed = true; in.words = pk->words; in.flags = pk->flags; in.exc = pk->exc; in.n_exc = pk->n_exc;
        return index_add_impl(c, in, offs, n, first_ref_idx, nb_mers_out);
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

int mq_index_add_segment(mq_ctx *c, const uint8_t *bytes, uint64_t n_bytes, uint32_t ref_idx, uint64_t ref_len,
                         uint64_t seg_start, uint64_t own_len) {
    if (!c || !bytes || !c->kids.empty()) return MQ_ERR_ARG;
    if (c->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
    if (ref_len >= (1ull << 31) || seg_start + own_len > ref_len) { c->err = "segment outside its record / record too long"; return MQ_ERR_RANGE; }
    cudaSetDevice(c->device);
    timers_reset(c);
    if (ref_len < (uint64_t)c->p.l + c->p.k - 1 || own_len == 0) { c->dir.push_back({(uint64_t)ref_idx, seg_start, 0ull}); return MQ_OK; }   // mers.rs:18
    const uint64_t ctx = seg_start > 0 ? 1 : 0;
    if (n_bytes < ctx + own_len) { c->err = "segment buffer shorter than ctx+own_len"; return MQ_ERR_ARG; }
    try {
        // the record handed to the scan INCLUDES the context byte, so run starts are decided exactly as in a
        // whole-record scan; only l-mers starting in [ctx, ctx+own_len) are emitted
        SeqInput in; in.seqs = bytes;
        std::vector<Piece> pcs{Piece{0, n_bytes, ref_idx, (uint32_t)(seg_start - ctx), (uint32_t)ctx, (uint32_t)(ctx + own_len), seg_start}};
        int rc = add_pieces(c, in, pcs, nullptr);
        timers_collect(c);
        return rc;
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

int mq_store_info(mq_ctx *c, uint64_t *n_minimizers, uint32_t *n_segments) {
    if (!c || !c->kids.empty()) return MQ_ERR_ARG;
    if (n_minimizers) *n_minimizers = c->st_n;
    if (n_segments) *n_segments = (uint32_t)c->dir.size();
    return MQ_OK;
}
int mq_store_export(mq_ctx *c, void **d_pos, void **d_hash, uint64_t *dir) {
    if (!c || !c->kids.empty()) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    if (d_pos) *d_pos = c->st_pos.p;
    if (d_hash) *d_hash = c->st_hash.p;
    if (dir) for (size_t i = 0; i < c->dir.size(); i++) { dir[3 * i] = c->dir[i][0]; dir[3 * i + 1] = c->dir[i][1]; dir[3 * i + 2] = c->dir[i][2]; }
    return MQ_OK;
}
// make room for a store of n_minimizers entries and hand out its device arrays: a collective (ncclAllGather /
// broadcast of every rank's share) can then land directly in place; mq_store_commit says what arrived
int mq_store_reserve(mq_ctx *c, uint64_t n_minimizers, void **d_pos, void **d_hash) {
    if (!c || !c->kids.empty() || !d_pos || !d_hash) return MQ_ERR_ARG;
    if (c->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
    cudaSetDevice(c->device);
    int rc;
    if ((rc = ensure_keep(c, c->st_pos, (n_minimizers + 64) * 4, c->st_n * 4))) return rc;
    if ((rc = ensure_keep(c, c->st_hash, (n_minimizers + 64) * 8, c->st_n * 8))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    *d_pos = c->st_pos.p; *d_hash = c->st_hash.p;
    return MQ_OK;
}
int mq_store_commit(mq_ctx *c, uint64_t n_minimizers, const uint64_t *dir, uint32_t n_seg) {
    if (!c || !c->kids.empty() || (n_seg && !dir)) return MQ_ERR_ARG;
    if (c->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
    uint64_t tot = 0;
    for (uint32_t i = 0; i < n_seg; i++) tot += dir[3 * i + 2];
    if (tot != n_minimizers || (n_minimizers + 64) * 4 > c->st_pos.cap) { c->err = "directory counts do not add up to n_minimizers"; return MQ_ERR_ARG; }
    try {
        c->st_n = n_minimizers;
        c->dir.clear();
        for (uint32_t i = 0; i < n_seg; i++) c->dir.push_back({dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
    } catch (...) { return MQ_ERR_NOMEM; }
    return MQ_OK;
}
int mq_store_import(mq_ctx *c, const void *d_pos, const void *d_hash, uint64_t n_min, const uint64_t *dir, uint32_t n_seg) {
    if (!c || !c->kids.empty() || (n_min && (!d_pos || !d_hash)) || (n_seg && !dir)) return MQ_ERR_ARG;
    if (c->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
    cudaSetDevice(c->device);
    uint64_t tot = 0;
    for (uint32_t i = 0; i < n_seg; i++) tot += dir[3 * i + 2];
    if (tot != n_min) { c->err = "directory counts do not add up to n_minimizers"; return MQ_ERR_ARG; }
    int rc;
    DBuf np, nh;
    if ((rc = ensure(c, np, (n_min + 64) * 4))) return rc;
    if ((rc = ensure(c, nh, (n_min + 64) * 8))) { dfree(np); return rc; }
    if (n_min) {
        CK(cudaMemcpyAsync(np.p, d_pos, n_min * 4, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(nh.p, d_hash, n_min * 8, cudaMemcpyDefault, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    dfree(c->st_pos); dfree(c->st_hash);
    c->st_pos = np; c->st_hash = nh; c->st_n = n_min;
    try {
        c->dir.clear();
        for (uint32_t i = 0; i < n_seg; i++) c->dir.push_back({dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
    } catch (...) { return MQ_ERR_NOMEM; }
    return MQ_OK;
}

}  // extern "C"

namespace {

// Presence filter over the valid entries of the frozen table + an L2 access-policy window that keeps it resident.
// ~6 bits per key (power of two), at least 8 KB.  MQ_NO_BLOOM=1 in the environment leaves it out (A/B measurements).
int build_bloom(mq_ctx *c) {
    dfree(c->d_bloom); c->bloom_wmask = 0;
    if (const char *e = getenv("MQ_NO_BLOOM")) if (e[0] == '1') return MQ_OK;
    uint64_t bits = 1ull << 16, per_key = 4;
    if (const char *e = getenv("MQ_BLOOM_BITS")) { const uint64_t v = strtoull(e, nullptr, 10); if (v >= 1 && v <= 64) per_key = v; }   // tuning runs
    while (bits < per_key * c->n_unique && bits < (1ull << 33)) bits <<= 1;
    const uint64_t words = bits / 64;
    int rc;
    if ((rc = ensure(c, c->d_bloom, words * 8))) return rc;
    c->bloom_wmask = (uint32_t)(words - 1);
    CK(cudaMemsetAsync(c->d_bloom.p, 0, words * 8, c->stream));
    k_bloom_build<<<c->n_sm * 8, 256, 0, c->stream>>>(c->d_table.as<Slot>(), c->tmask + 2, c->d_bloom.as<unsigned long long>(), c->bloom_wmask);
    c->launches++;
    CK(cudaGetLastError());
    // keep it in L2 while reads stream through: persisting lines for the filter, streaming for everything else on this stream
    cudaDeviceProp prop;
    if (!getenv("MQ_BLOOM_NO_PERSIST") && cudaGetDeviceProperties(&prop, c->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        if (getenv("MQ_BLOOM_VERBOSE")) fprintf(stderr, "[bloom] %llu bytes, persistingL2CacheMaxSize %d, accessPolicyMaxWindowSize %d, l2CacheSize %d\n",
                                                (unsigned long long)(words * 8), prop.persistingL2CacheMaxSize, prop.accessPolicyMaxWindowSize, prop.l2CacheSize);
        const size_t want = std::min<size_t>(words * 8, (size_t)prop.persistingL2CacheMaxSize);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = c->d_bloom.p;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(words * 8, (size_t)prop.accessPolicyMaxWindowSize);
        av.accessPolicyWindow.hitRatio = want >= words * 8 ? 1.0f : (float)want / (float)(words * 8);
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    return MQ_OK;
}

// build the table of this context from its store (which must hold the minimizers of the WHOLE reference)
int freeze_local(mq_ctx *c, const uint64_t *ref_lens, uint32_t n_refs) {
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    const size_t ns = c->dir.size();
    // order segments by (ref_idx, seg_start); reorder the store if it was filled out of order
    std::vector<size_t> perm(ns); std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) {
        return c->dir[a][0] != c->dir[b][0] ? c->dir[a][0] < c->dir[b][0] : c->dir[a][1] < c->dir[b][1]; });
    bool identity = true;
    for (size_t i = 0; i < ns; i++) identity &= perm[i] == i;
    if (!identity && c->st_n) {
        std::vector<uint64_t> start(ns + 1, 0);
        for (size_t i = 0; i < ns; i++) start[i + 1] = start[i] + c->dir[i][2];
        DBuf np, nh;
        if ((rc = ensure(c, np, (c->st_n + 64) * 4))) return rc;
        if ((rc = ensure(c, nh, (c->st_n + 64) * 8))) { dfree(np); return rc; }
        uint64_t w = 0;
        for (size_t i = 0; i < ns; i++) {
            size_t s = perm[i]; uint64_t cnt = c->dir[s][2];
            if (cnt) {
                CK(cudaMemcpyAsync(np.as<uint32_t>() + w, c->st_pos.as<uint32_t>() + start[s], cnt * 4, cudaMemcpyDeviceToDevice, c->stream));
                CK(cudaMemcpyAsync(nh.as<uint64_t>() + w, c->st_hash.as<uint64_t>() + start[s], cnt * 8, cudaMemcpyDeviceToDevice, c->stream));
            }
            w += cnt;
        }
        CK(cudaStreamSynchronize(c->stream));
        dfree(c->st_pos); dfree(c->st_hash); c->st_pos = np; c->st_hash = nh;
    }
    std::vector<std::array<uint64_t, 3>> sd(ns);
    for (size_t i = 0; i < ns; i++) sd[i] = c->dir[perm[i]];
    // records = runs of equal ref_idx
    std::vector<uint32_t> rec_off, rec_id;
    uint64_t acc = 0, n_tuples = 0;
    c->nb_mers.assign(n_refs, 0);
    for (size_t i = 0; i < ns;) {
        size_t j = i; uint64_t cnt = 0;
        while (j < ns && sd[j][0] == sd[i][0]) { cnt += sd[j][2]; j++; }
        if (sd[i][0] >= n_refs) { c->err = "segment ref_idx >= n_refs"; return MQ_ERR_ARG; }
        rec_off.push_back((uint32_t)acc); rec_id.push_back((uint32_t)sd[i][0]);
        uint64_t q = cnt >= c->p.k ? cnt - c->p.k + 1 : 0;
        c->nb_mers[sd[i][0]] = q; n_tuples += q; acc += cnt;
        i = j;
    }
    rec_off.push_back((uint32_t)acc);
    if (acc != c->st_n) { c->err = "internal: directory/store mismatch"; return MQ_ERR_STATE; }
    if (c->st_n >= (1ull << 32) - 64) { c->err = "reference too large (2^32 minimizers)"; return MQ_ERR_RANGE; }
    const uint32_t n_rec = (uint32_t)rec_id.size();
    // table: power-of-two capacity >= 2 x tuples (load factor <= 0.5), one spare slot for key == EMPTY
    uint64_t cap = 1024;
    while (cap < 2 * n_tuples) cap <<= 1;
    if ((rc = ensure(c, c->d_table, (cap + 1) * sizeof(Slot)))) return rc;
    c->tmask = cap - 1;
    if ((rc = ensure(c, c->d_ref_lens, ((size_t)n_refs + 1) * 8))) return rc;
    if ((rc = ensure(c, c->d_misc, ((size_t)n_rec + 2) * 4 * 3 + 64))) return rc;
    uint32_t *d_rec_off = c->d_misc.as<uint32_t>(), *d_rec_id = d_rec_off + n_rec + 2;
    unsigned long long *d_cnt = (unsigned long long *)(d_rec_id + n_rec + 2);   // 2 * (n_rec + 2) words in: 8-byte aligned
    {
        StageTimer t(c, "insert");
        if (n_refs) CK(cudaMemcpyAsync(c->d_ref_lens.p, ref_lens, (size_t)n_refs * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d_rec_off, rec_off.data(), rec_off.size() * 4, cudaMemcpyHostToDevice, c->stream));
        if (n_rec) CK(cudaMemcpyAsync(d_rec_id, rec_id.data(), rec_id.size() * 4, cudaMemcpyHostToDevice, c->stream));
        k_table_clear<<<c->n_sm * 8, 256, 0, c->stream>>>(c->d_table.as<Slot>(), cap + 1);
        c->launches++;
        if (c->st_n && n_rec) {
            KminmerArgs a{};
            a.pos = c->st_pos.as<uint32_t>(); a.hash = c->st_hash.as<uint64_t>(); a.n_min = (uint32_t)c->st_n;
            a.rec_off = d_rec_off; a.rec_id = d_rec_id; a.n_rec = n_rec; a.km_off = nullptr; a.k = c->p.k; a.l = c->p.l;
            Table t2{c->d_table.as<Slot>(), c->tmask};
            k_insert_kminmers<<<(uint32_t)((c->st_n + 255) / 256), 256, 0, c->stream>>>(a, t2, 1);
            c->launches++;
        }
        CK(cudaMemsetAsync(d_cnt, 0, 16, c->stream));
        k_table_count<<<c->n_sm * 8, 256, 0, c->stream>>>(c->d_table.as<Slot>(), cap + 1, d_cnt);
        c->launches++;
        CK(cudaGetLastError());
        uint64_t hc[2];
        CK(cudaMemcpyAsync(hc, d_cnt, 16, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->n_unique = hc[0]; c->n_keys = hc[1];
        if ((rc = build_bloom(c))) return rc;
    }
    c->n_refs = n_refs; c->frozen = true;
    dfree(c->st_pos); dfree(c->st_hash); c->st_n = 0;
    timers_collect(c);
    return MQ_OK;
}

// multi-GPU store exchange: every child ends up with the union of all stores (an all-gather of device arrays,
// GPU to GPU: cudaMemcpyPeerAsync rides NVLink when peer access is on).  Every child pulls the parts of the others.
int exchange_stores(mq_ctx *par) {
    const size_t G = par->kids.size();
    std::vector<uint64_t> cnt(G), off(G + 1, 0);
    for (size_t g = 0; g < G; g++) { cnt[g] = par->kids[g]->st_n; off[g + 1] = off[g] + cnt[g]; }
    const uint64_t tot = off[G];
    std::vector<DBuf> np(G), nh(G);
    for (size_t g = 0; g < G; g++) {
        mq_ctx *c = par->kids[g];
        cudaSetDevice(c->device);
        int rc;
        if ((rc = ensure(c, np[g], (tot + 64) * 4))) { par->err = c->err; return rc; }
        if ((rc = ensure(c, nh[g], (tot + 64) * 8))) { par->err = c->err; return rc; }
    }
    for (size_t g = 0; g < G; g++) {
        mq_ctx *c = par->kids[g];
        cudaSetDevice(c->device);
        StageTimer t(c, "exchange");
        for (size_t q = 0; q < G; q++) {
            const size_t src = (g + q) % G;                  // start with my own part, then walk the ring
            mq_ctx *s = par->kids[src];
            if (!cnt[src]) continue;
            cudaError_t e1 = cudaMemcpyPeerAsync(np[g].as<uint32_t>() + off[src], c->device, s->st_pos.p, s->device, cnt[src] * 4, c->stream);
            cudaError_t e2 = cudaMemcpyPeerAsync(nh[g].as<uint64_t>() + off[src], c->device, s->st_hash.p, s->device, cnt[src] * 8, c->stream);
            if (e1 != cudaSuccess || e2 != cudaSuccess) { par->err = std::string("cudaMemcpyPeerAsync: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2); return MQ_ERR_CUDA; }
        }
    }
    for (size_t g = 0; g < G; g++) {
        mq_ctx *c = par->kids[g];
        cudaSetDevice(c->device);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { par->err = "store exchange failed"; return MQ_ERR_CUDA; }
    }
    std::vector<std::array<uint64_t, 3>> dir;
    for (size_t g = 0; g < G; g++) for (auto &d : par->kids[g]->dir) dir.push_back(d);
    for (size_t g = 0; g < G; g++) {
        mq_ctx *c = par->kids[g];
        cudaSetDevice(c->device);
        dfree(c->st_pos); dfree(c->st_hash);
        c->st_pos = np[g]; c->st_hash = nh[g]; c->st_n = tot; c->dir = dir;
    }
    return MQ_OK;
}

}  // namespace

extern "C" {

int mq_index_freeze(mq_ctx *c, const uint64_t *ref_lens, uint32_t n_refs, uint64_t *n_unique, uint64_t *n_keys) {
    if (!c || (n_refs && !ref_lens)) return MQ_ERR_ARG;
    try {
        if (c->kids.empty()) {
            if (c->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
            int rc = freeze_local(c, ref_lens, n_refs);
            if (rc) return rc;
            if (n_unique) *n_unique = c->n_unique;
            if (n_keys) *n_keys = c->n_keys;
            return MQ_OK;
        }
        if (c->kids[0]->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
        int rc = exchange_stores(c);
        if (rc) return rc;
        const size_t G = c->kids.size();
        std::vector<int> rcs(G, 0);
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back([&, g]() { rcs[g] = freeze_local(c->kids[g], ref_lens, n_refs); });
        for (auto &t : th) t.join();
        for (size_t g = 0; g < G; g++) if (rcs[g]) { c->err = c->kids[g]->err; return rcs[g]; }
        for (size_t g = 1; g < G; g++) if (c->kids[g]->n_unique != c->kids[0]->n_unique || c->kids[g]->n_keys != c->kids[0]->n_keys) { c->err = "replicated tables differ"; return MQ_ERR_STATE; }
        if (n_unique) *n_unique = c->kids[0]->n_unique;
        if (n_keys) *n_keys = c->kids[0]->n_keys;
        return MQ_OK;
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

int mq_index_nb_mers(mq_ctx *c, uint64_t *nb, uint32_t n_refs) {
    if (!c || !nb) return MQ_ERR_ARG;
    c = FIRST(c);
    if (!c->frozen) return MQ_ERR_STATE;
    for (uint32_t i = 0; i < n_refs; i++) nb[i] = i < c->nb_mers.size() ? c->nb_mers[i] : 0;
    return MQ_OK;
}

// ---- on-disk index (the reference rebuilds its index on every run) ----------------
namespace {
struct IndexFileHeader {
    char magic[8];             // "MQB200IX"
    uint32_t version, k, l, use_hpc;
    double density;
    uint64_t slots, n_unique, n_keys, n_refs, names_bytes;
};
}
int mq_index_save(mq_ctx *c, const char *path, const char *names_blob, uint64_t names_bytes) {
    if (!c || !path || (names_bytes && !names_blob)) return MQ_ERR_ARG;
    c = FIRST(c);
    if (!c->frozen) { c->err = "index not frozen"; return MQ_ERR_STATE; }
    cudaSetDevice(c->device);
    try {
        FILE *f = fopen(path, "wb");
        if (!f) { c->err = std::string("cannot create ") + path; return MQ_ERR_ARG; }
        IndexFileHeader h{};
        memcpy(h.magic, "MQB200IX", 8);
        h.version = 1; h.k = c->p.k; h.l = c->p.l; h.use_hpc = c->p.use_hpc; h.density = c->p.density;
        h.slots = c->tmask + 2; h.n_unique = c->n_unique; h.n_keys = c->n_keys; h.n_refs = c->n_refs; h.names_bytes = names_bytes;
        bool ok = fwrite(&h, sizeof h, 1, f) == 1;
        std::vector<uint64_t> lens(c->n_refs);
        if (c->n_refs) {
            if (cudaMemcpy(lens.data(), c->d_ref_lens.p, (size_t)c->n_refs * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { fclose(f); c->err = "D2H ref_lens"; return MQ_ERR_CUDA; }
            ok = ok && fwrite(lens.data(), 8, c->n_refs, f) == c->n_refs;
            std::vector<uint64_t> nb(c->n_refs, 0);
            for (uint32_t i = 0; i < c->n_refs && i < c->nb_mers.size(); i++) nb[i] = c->nb_mers[i];
            ok = ok && fwrite(nb.data(), 8, c->n_refs, f) == c->n_refs;
        }
        if (names_bytes) ok = ok && fwrite(names_blob, 1, names_bytes, f) == names_bytes;
        const size_t CH = 64u << 20;                       // table goes out in 64 MB pieces through the pinned bounce buffer
        int rc = ensure_host(c, c->h_pin, CH);
        if (rc) { fclose(f); return rc; }
        const size_t total = (size_t)h.slots * sizeof(Slot);
        for (size_t off = 0; off < total && ok; off += CH) {
            const size_t n = std::min(CH, total - off);
            if (cudaMemcpy(c->h_pin.p, (const uint8_t *)c->d_table.p + off, n, cudaMemcpyDeviceToHost) != cudaSuccess) { fclose(f); c->err = "D2H table"; return MQ_ERR_CUDA; }
            ok = fwrite(c->h_pin.p, 1, n, f) == n;
        }
        ok = (fclose(f) == 0) && ok;
        if (!ok) { c->err = std::string("short write to ") + path; return MQ_ERR_ARG; }
        return MQ_OK;
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

static int index_load_one(mq_ctx *c, const char *path, uint64_t *ref_lens_out, uint32_t ref_cap, uint32_t *n_refs_out, char *names_out,
                          uint64_t names_cap, uint64_t *names_bytes_out, uint64_t *n_unique_out) {
    if (c->frozen || c->st_n || !c->dir.empty()) { c->err = "load needs a fresh context"; return MQ_ERR_STATE; }
    cudaSetDevice(c->device);
    FILE *f = fopen(path, "rb");
    if (!f) { c->err = std::string("cannot open ") + path; return MQ_ERR_ARG; }
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
    struct stat st;
    if (fstat(fileno(f), &st) != 0) { c->err = "cannot stat index file"; return MQ_ERR_ARG; }
    const uint64_t fsize = (uint64_t)st.st_size;
    IndexFileHeader h{};
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "MQB200IX", 8) != 0 || h.version != 1) { c->err = "not a mapquik_b200 index file"; return MQ_ERR_ARG; }
    if (h.k != c->p.k || h.l != c->p.l || h.use_hpc != c->p.use_hpc || h.density != c->p.density) {
        c->err = "index was built with different k / l / density / hpc"; return MQ_ERR_ARG;
    }
    // the header is untrusted: every size must be consistent with the file before anything is allocated from it
    if (h.slots < 2 || ((h.slots - 1) & (h.slots - 2)) != 0 || h.slots > (1ull << 40) || h.n_refs >= (1ull << 31) || h.names_bytes > fsize ||
        sizeof h + h.n_refs * 16 + h.names_bytes + h.slots * sizeof(Slot) != fsize) { c->err = "corrupt or truncated index file"; return MQ_ERR_ARG; }
    if (n_refs_out) *n_refs_out = (uint32_t)h.n_refs;
    if (names_bytes_out) *names_bytes_out = h.names_bytes;
    if (n_unique_out) *n_unique_out = h.n_unique;
    std::vector<uint64_t> lens(h.n_refs), nb(h.n_refs);
    bool ok = true;
    if (h.n_refs) ok = fread(lens.data(), 8, h.n_refs, f) == h.n_refs && fread(nb.data(), 8, h.n_refs, f) == h.n_refs;
    if (ref_lens_out) { if (ref_cap < h.n_refs) { c->err = "ref_lens_out too small"; return MQ_ERR_ARG; } memcpy(ref_lens_out, lens.data(), h.n_refs * 8); }
    if (h.names_bytes) {
        if (names_out) { if (names_cap < h.names_bytes) { c->err = "names_out too small"; return MQ_ERR_ARG; } ok = ok && fread(names_out, 1, h.names_bytes, f) == h.names_bytes; }
        else ok = ok && fseek(f, (long)h.names_bytes, SEEK_CUR) == 0;
    }
    int rc;
    const size_t total = (size_t)h.slots * sizeof(Slot), CH = 64u << 20;
    if ((rc = ensure(c, c->d_table, total))) return rc;
    if ((rc = ensure(c, c->d_ref_lens, (h.n_refs + 1) * 8))) return rc;
    if ((rc = ensure_host(c, c->h_pin, CH))) return rc;
    for (size_t off = 0; off < total && ok; off += CH) {
        const size_t n = std::min(CH, total - off);
        ok = fread(c->h_pin.p, 1, n, f) == n;
        if (ok && cudaMemcpy((uint8_t *)c->d_table.p + off, c->h_pin.p, n, cudaMemcpyHostToDevice) != cudaSuccess) { c->err = "H2D table"; return MQ_ERR_CUDA; }
    }
    if (!ok) { c->err = "truncated index file"; return MQ_ERR_ARG; }
    if (h.n_refs && cudaMemcpy(c->d_ref_lens.p, lens.data(), h.n_refs * 8, cudaMemcpyHostToDevice) != cudaSuccess) { c->err = "H2D ref_lens"; return MQ_ERR_CUDA; }
    c->tmask = h.slots - 2; c->n_refs = (uint32_t)h.n_refs; c->n_unique = h.n_unique; c->n_keys = h.n_keys;
    c->nb_mers.assign(nb.begin(), nb.end());
    if ((rc = build_bloom(c))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    c->frozen = true;
    return MQ_OK;
}
int mq_index_load(mq_ctx *c, const char *path, uint64_t *ref_lens_out, uint32_t ref_cap, uint32_t *n_refs_out, char *names_out,
                  uint64_t names_cap, uint64_t *names_bytes_out, uint64_t *n_unique_out) {
    if (!c || !path) return MQ_ERR_ARG;
    try {
        if (c->kids.empty()) return index_load_one(c, path, ref_lens_out, ref_cap, n_refs_out, names_out, names_cap, names_bytes_out, n_unique_out);
        for (size_t g = 0; g < c->kids.size(); g++) {
            int rc = index_load_one(c->kids[g], path, g ? nullptr : ref_lens_out, ref_cap, n_refs_out, g ? nullptr : names_out, names_cap, names_bytes_out, n_unique_out);
            if (rc) { c->err = c->kids[g]->err; return rc; }
        }
        return MQ_OK;
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

// ---- mapping ---------------------------------------------------------------------------------------
static int map_impl(mq_ctx *c, const SeqInput &in, const uint64_t *offs, uint32_t n, mq_hit *out, mq_hit *d_out) {
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    if (c->kids.empty()) {
        if (!c->frozen) { c->err = "index not frozen"; return MQ_ERR_STATE; }
        cudaSetDevice(c->device);
        timers_reset(c);
        if (n == 0) return MQ_OK;
        rc = map_pipeline(c, in, offs, n, out, d_out);
        timers_collect(c, false);
        return rc;
    }
    // multi-GPU: reads are sharded in contiguous blocks of near-equal base counts, no inter-GPU communication
    // (find_matches is a pure function of the read and the frozen index, mers.rs:77)
    if (!c->kids[0]->frozen) { c->err = "index not frozen"; return MQ_ERR_STATE; }
    if (d_out || in.resident) { c->err = "device-resident inputs need a single-GPU context"; return MQ_ERR_ARG; }
    if (n == 0) return MQ_OK;
    const size_t G = c->kids.size();
    std::vector<uint32_t> cut(G + 1, n);
    cut[0] = 0;
    const uint64_t total = offs[n] - offs[0];
    for (size_t g = 1; g < G; g++) {
        const uint64_t want = offs[0] + total * g / G;
        cut[g] = (uint32_t)(std::lower_bound(offs, offs + n + 1, want) - offs);
        cut[g] = std::max(cut[g], cut[g - 1]);
    }
    c->kid_share.assign(G, {0, 0});
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++) {
        c->kid_share[g] = {cut[g], cut[g + 1]};
        if (cut[g + 1] == cut[g]) continue;
        th.emplace_back([&, g]() {
            mq_ctx *k = c->kids[g];
            cudaSetDevice(k->device);
            timers_reset(k);
            rcs[g] = map_pipeline(k, in, offs + cut[g], cut[g + 1] - cut[g], out + cut[g], nullptr);
            timers_collect(k, false);
        });
    }
    for (auto &t : th) t.join();
    for (size_t g = 0; g < G; g++) if (rcs[g]) { c->err = c->kids[g]->err; return rcs[g]; }
    return MQ_OK;
}

int mq_map_batch(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, mq_hit *out) {
    if (!c || !offs || (!seqs && n) || (!out && n)) return MQ_ERR_ARG;
    try { SeqInput in; in.seqs = seqs; return map_impl(c, in, offs, n, out, nullptr); }
    catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}
int mq_map_batch_packed(mq_ctx *c, const mq_packed *pk, const uint64_t *offs, uint32_t n, mq_hit *out) {
    if (!c || !pk || !offs || (n && (!pk->words || !pk->flags || !out)) || (pk->n_exc && !pk->exc)) return MQ_ERR_ARG;
    try {
        SeqInput in; in.packed = true; in.words = pk->words; in.flags = pk->flags; in.exc = pk->exc; in.n_exc = pk->n_exc;
        return map_impl(c, in, offs, n, out, nullptr);
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}
int mq_map_batch_device(mq_ctx *c, const uint8_t *d_seqs, const uint64_t *offs, uint32_t n, mq_hit *d_out) {
    if (!c || !offs || (n && (!d_seqs || !d_out))) return MQ_ERR_ARG;
    if (((uintptr_t)d_seqs & 15) != 0) { c->err = "device sequence buffer must be 16-byte aligned"; return MQ_ERR_ARG; }
    try { SeqInput in; in.seqs = d_seqs; in.resident = true; return map_impl(c, in, offs, n, nullptr, d_out); }
    catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}
int mq_map_batch_packed_device(mq_ctx *c, const mq_packed *d_pk, const uint64_t *offs, uint32_t n, mq_hit *d_out) {
    if (!c || !d_pk || !offs || (n && (!d_pk->words || !d_pk->flags || !d_out)) || (d_pk->n_exc && !d_pk->exc)) return MQ_ERR_ARG;
    if (((uintptr_t)d_pk->words & 15) != 0) { c->err = "device word buffer must be 16-byte aligned"; return MQ_ERR_ARG; }
    if (d_pk->n_exc >= (1ull << 32)) { c->err = "too many exception intervals"; return MQ_ERR_RANGE; }
    try {
        SeqInput in; in.packed = true; in.resident = true; in.words = d_pk->words; in.flags = d_pk->flags; in.exc = d_pk->exc; in.n_exc = d_pk->n_exc;
        return map_impl(c, in, offs, n, nullptr, d_out);
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

// mers.rs:181 -- hand-rolled (a CLI writes millions of these per second of GPU time; snprintf with twelve conversions was
// half a microsecond per line)
int mq_format_paf(char *buf, size_t cap, const char *q_id, uint64_t q_len, const char *r_id, uint64_t r_len, const mq_hit *h) {
    if (!buf || !q_id || !r_id || !h) return MQ_ERR_ARG;
    const size_t lq = strlen(q_id), lr = strlen(r_id);
    if (cap < lq + lr + 10 * 20 + 16) {                          // (ten numbers of <= 20 digits, strand, eleven tabs, NUL always fit above that)
        auto digits = [](uint64_t v) { size_t n = 1; while (v >= 10) { v /= 10; n++; } return n; };
        const size_t need = lq + lr + digits(q_len) + digits(h->q_start) + digits(h->q_end) + 2 * digits(r_len) + digits(h->r_start) +
                            digits(h->r_end) + digits(h->score) + digits(h->mapq) + 1 + 11 + 1;
        if (cap < need) return MQ_ERR_ARG;
    }
    char *w = buf;
    auto str = [&](const char *s_, size_t n) { memcpy(w, s_, n); w += n; *w++ = '\t'; };
    auto num = [&](uint64_t v, char sep) {
        char tmp[20]; int n = 0;
        do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (n) *w++ = tmp[--n];
        *w++ = sep;
    };
    str(q_id, lq); num(q_len, '\t'); num(h->q_start, '\t'); num(h->q_end, '\t');
    *w++ = h->rc ? '-' : '+'; *w++ = '\t';
    str(r_id, lr); num(r_len, '\t'); num(h->r_start, '\t'); num(h->r_end, '\t'); num(h->score, '\t'); num(r_len, '\t'); num(h->mapq, '\0');
    return (int)(w - buf) - 1;
}

// ---- introspection (single-GPU contexts; synchronous) ---------------------------------------------------------
// stage records [0, n) of an ASCII or packed host batch into slot 0 and scan them
static int scan_batch_sync(mq_ctx *c, const SeqInput &in, const uint64_t *offs, uint32_t n, uint32_t min_len, BatchDev &bd, uint64_t *M,
                           std::vector<uint32_t> *so) {
    int rc;
    if ((rc = init_streams(c))) return rc;
    if ((rc = stage_pieces(c, c->slot[0], in, nullptr, offs, 0, n, min_len, false, bd))) return rc;
    return scan_sync(c, c->slot[0], bd, M, so);
}

static int minimizers_impl(mq_ctx *c, const SeqInput &in, const uint64_t *offs, uint32_t n, uint64_t *seq_off, uint32_t *pos,
                           uint64_t *hash, uint64_t cap, uint64_t *n_total) {
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    uint64_t M = 0; std::vector<uint32_t> so; BatchDev bd;
    if (n && (rc = scan_batch_sync(c, in, offs, n, 0, bd, &M, &so))) return rc;
    if (n_total) *n_total = M;
    if (seq_off && n) for (uint32_t i = 0; i <= n; i++) seq_off[i] = so[i];
    if (pos && hash && M) {
        if (cap < M) { c->err = "output capacity too small"; return MQ_ERR_ARG; }
        CK(cudaMemcpyAsync(pos, c->d_pos.p, M * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(hash, c->d_hash.p, M * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    timers_collect(c);
    return MQ_OK;
}
int mq_minimizers(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint64_t *seq_off, uint32_t *pos,
                  uint64_t *hash, uint64_t cap, uint64_t *n_total) {
    if (!c || !c->kids.empty() || !offs || (!seqs && n)) return MQ_ERR_ARG;
    try { SeqInput in; in.seqs = seqs; return minimizers_impl(c, in, offs, n, seq_off, pos, hash, cap, n_total); }
    catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}
int mq_minimizers_packed(mq_ctx *c, const mq_packed *pk, const uint64_t *offs, uint32_t n, uint64_t *seq_off, uint32_t *pos,
                         uint64_t *hash, uint64_t cap, uint64_t *n_total) {
    if (!c || !c->kids.empty() || !pk || !offs || (n && (!pk->words || !pk->flags)) || (pk->n_exc && !pk->exc)) return MQ_ERR_ARG;
    try {
        SeqInput in; in.packed = true; in.words = pk->words; in.flags = pk->flags; in.exc = pk->exc; in.n_exc = pk->n_exc;
        return minimizers_impl(c, in, offs, n, seq_off, pos, hash, cap, n_total);
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

int mq_kminmers(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint64_t *seq_off, uint32_t *start,
                uint32_t *end, uint32_t *offrev, uint64_t *hash, uint64_t cap, uint64_t *n_total) {
    if (!c || !c->kids.empty() || !offs || (!seqs && n)) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    try {
        uint64_t M = 0, Q = 0;
        std::vector<uint32_t> so((size_t)n + 1, 0), ko((size_t)n + 1, 0);
        if (n) {
            SeqInput in; in.seqs = seqs; BatchDev bd;
            if ((rc = scan_batch_sync(c, in, offs, n, c->p.l + c->p.k - 1, bd, &M, &so))) return rc;
            for (uint32_t i = 0; i < n; i++) { uint32_t cnt = so[i + 1] - so[i]; ko[i + 1] = ko[i] + (cnt >= c->p.k ? cnt - c->p.k + 1 : 0); }
            Q = ko[n];
        }
        if (n_total) *n_total = Q;
        if (seq_off) for (uint32_t i = 0; i <= n; i++) seq_off[i] = ko[i];
        if (start && end && offrev && hash && Q) {
            if (cap < Q) { c->err = "output capacity too small"; return MQ_ERR_ARG; }
            if ((rc = ensure(c, c->d_misc, ((size_t)n + 2) * 4 + (Q + 16) * (4 * 3 + 8) + 64))) return rc;
            uint64_t *t_hash = c->d_misc.as<uint64_t>();
            uint32_t *t_start = (uint32_t *)(t_hash + Q + 1), *t_end = t_start + Q + 1, *t_off = t_end + Q + 1, *d_ko = t_off + Q + 1;
            CK(cudaMemcpyAsync(d_ko, ko.data(), ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, c->stream));
            KminmerArgs a{};
            a.pos = c->d_pos.as<uint32_t>(); a.hash = c->d_hash.as<uint64_t>(); a.n_min = (uint32_t)M;
            a.rec_off = c->d_seq_off.as<uint32_t>(); a.rec_id = nullptr; a.n_rec = n; a.km_off = d_ko; a.k = c->p.k; a.l = c->p.l;
            a.t_start = t_start; a.t_end = t_end; a.t_offrev = t_off; a.t_hash = t_hash;
            Table t{nullptr, 0};
            k_insert_kminmers<<<(uint32_t)((M + 255) / 256), 256, 0, c->stream>>>(a, t, 0);
            c->launches++;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(start, t_start, Q * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(end, t_end, Q * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(offrev, t_off, Q * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(hash, t_hash, Q * 8, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        timers_collect(c);
        return MQ_OK;
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

int mq_index_get(mq_ctx *c, const uint64_t *hashes, uint64_t n, uint8_t *found, uint32_t *id, uint32_t *start, uint32_t *end,
                 uint32_t *offset, uint8_t *rc_out) {
    if (!c || (n && (!hashes || !found || !id || !start || !end || !offset || !rc_out))) return MQ_ERR_ARG;
    c = FIRST(c);
    if (!c->frozen) return MQ_ERR_STATE;
    cudaSetDevice(c->device);
    if (n == 0) return MQ_OK;
    int rc;
    if ((rc = ensure(c, c->d_misc, n * (8 + 4 * 4 + 2) + 256))) return rc;
    uint64_t *d_k = c->d_misc.as<uint64_t>();
    uint32_t *d_id = (uint32_t *)(d_k + n), *d_s = d_id + n, *d_e = d_s + n, *d_o = d_e + n;
    uint8_t *d_f = (uint8_t *)(d_o + n), *d_r = d_f + n;
    CK(cudaMemcpyAsync(d_k, hashes, n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(d_id, 0, n * 18, c->stream));
    Table t{c->d_table.as<Slot>(), c->tmask};
    k_index_get<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(t, d_k, n, d_f, d_id, d_s, d_e, d_o, d_r);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(found, d_f, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(id, d_id, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(start, d_s, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(end, d_e, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(offset, d_o, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(rc_out, d_r, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MQ_OK;
}

int mq_matches(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint64_t *match_off, uint32_t *fields6,
               uint64_t cap, uint64_t *n_total) {
    if (!c || !c->kids.empty() || !offs || (!seqs && n)) return MQ_ERR_ARG;
    if (!c->frozen) return MQ_ERR_STATE;
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    if (n_total) *n_total = 0;
    if (n == 0) { if (match_off) match_off[0] = 0; return MQ_OK; }
    try {
        SeqInput in; in.seqs = seqs; BatchDev bd; uint64_t M = 0; std::vector<uint32_t> so;
        if ((rc = scan_batch_sync(c, in, offs, n, c->p.l + c->p.k - 1, bd, &M, &so))) return rc;
        if ((rc = ensure_workspace(c, n, bd.n_tiles, bd.bases, true))) return rc;
        if ((rc = ensure(c, c->slot[0].d_hits, (size_t)n * sizeof(HitRec)))) return rc;
        CK(cudaMemsetAsync(c->d_nmatch.p, 0, ((size_t)n + 1) * 4, c->stream));
        if ((rc = enqueue_probe_chain(c, bd, c->slot[0].d_hits.as<HitRec>()))) return rc;
        std::vector<uint32_t> nm(n);
        CK(cudaMemcpyAsync(nm.data(), c->d_nmatch.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        uint64_t tot = 0;
        for (uint32_t i = 0; i < n; i++) { if (match_off) match_off[i] = tot; tot += nm[i]; }
        if (match_off) match_off[n] = tot;
        if (n_total) *n_total = tot;
        if (fields6 && tot) {
            if (cap < tot) { c->err = "output capacity too small"; return MQ_ERR_ARG; }
            // only the first nm[i] records of a read's region are written by the kernel: copy exactly those
            std::vector<MatchRec> all(so[n]);
            for (uint32_t i = 0; i < n; i++)
                if (nm[i]) CK(cudaMemcpyAsync(all.data() + so[i], c->d_matches.as<MatchRec>() + so[i], (size_t)nm[i] * sizeof(MatchRec),
                                              cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            uint64_t w = 0;
            for (uint32_t i = 0; i < n; i++) for (uint32_t j = 0; j < nm[i]; j++, w++) {
                const MatchRec &m = all[so[i] + j];
                uint32_t *f = fields6 + 6 * w;
                f[0] = m.q_start; f[1] = m.q_end; f[2] = m.r_start; f[3] = m.r_end; f[4] = m.last_j - m.head_j + 1; f[5] = m.ref_rc;
            }
        }
        timers_collect(c);
        return MQ_OK;
    } catch (...) { c->err = "host allocation failed"; return MQ_ERR_NOMEM; }
}

}  // extern "C"
