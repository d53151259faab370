// mq_lib.cu -- host side of libmapquik_b200.so: context, device memory, kernel launches and the
// extern "C" entry points declared in include/mapquik_b200.h.  No torch, no CPU fallback: every
// computing entry point fails with MQ_ERR_CUDA when no CUDA device is usable.
#include "../../include/mapquik_b200.h"
#include "mq_kernels.cuh"
#include "mq_scan_v2.cuh"
#include "mq_scan_v3.cuh"
#include <cstdlib>
#include <cstddef>

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

using namespace mq;

namespace {

struct DBuf {
    void *p = nullptr; size_t cap = 0;
    template <class T> T *as() const { return (T *)p; }
};

constexpr uint64_t MAP_SUB_BATCH_BYTES = 128ull << 20;  // bases per pipelined mapping sub-batch (H2D of i+1 overlaps compute of i)
constexpr size_t   PAD = 256;                          // slack after sequence buffers (word loads)

}  // namespace

struct mq_ctx {
    mq_params p{};
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t bound = 0;
    ScanTables tab{};
    ScanTablesV2 tab2{};
    ScanTablesV3 tab3{};
    int v2_ctas_per_sm = 0;
    bool scan_v1 = false;          // MQ_SCAN_V1=1 selects the first-generation scan kernel (A/B, debugging)
    bool scan_v2 = false;          // MQ_SCAN_V2=1 selects the second-generation one (contiguous lane streams)
    std::string err;
    uint64_t launches = 0, scan_kernel_launches = 0;
    int n_sm = 148;
    cudaEvent_t region_a = nullptr, region_b = nullptr;
    uint64_t last_minimizers = 0;
    // per-batch workspace
    DBuf d_seqs, d_offs, d_first_tile, d_tile_seq, d_ev_hash, d_ev_meta, d_lane_cnt, d_tile_cnt, d_blocksums,
         d_scalars, d_ovf_tile, d_ovf_meta, d_ovf_hash, d_pos, d_hash, d_seq_off, d_matches, d_nmatch, d_hits,
         d_pos_base, d_emit_len, d_misc, d_big_list;
    uint32_t ovf_cap = 1u << 16;
    // minimizer store (reference side)
    DBuf st_pos, st_hash; uint64_t st_n = 0;
    std::vector<std::array<uint64_t, 3>> dir;          // (ref_idx, seg_start, count)
    // frozen index
    DBuf d_table, d_ref_lens; uint64_t tmask = 0; bool frozen = false; uint32_t n_refs = 0;
    std::vector<uint64_t> nb_mers; uint64_t n_unique = 0, n_keys = 0;
    // pinned bounce buffers
    void *h_pin = nullptr; size_t h_pin_cap = 0;
    // double-buffered upload path of mq_map_batch
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};
    DBuf d_seqs2[2], d_offs2[2];
    void *h_offs2[2] = {nullptr, nullptr}; size_t h_offs2_cap[2] = {0, 0};
    // timings
    std::map<std::string, float> ms, ms_total;
    uint64_t timer_gen = 0;
    struct PendingTimer { std::string name; cudaEvent_t a, b; uint64_t gen; };
    std::vector<PendingTimer> pending;
    std::vector<cudaEvent_t> ev_pool;
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
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
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

int ensure_pin(mq_ctx *c, size_t bytes) {
    if (bytes <= c->h_pin_cap) return MQ_OK;
    if (c->h_pin) cudaFreeHost(c->h_pin);
    c->h_pin = nullptr; c->h_pin_cap = 0;
    CK(cudaMallocHost(&c->h_pin, bytes + bytes / 4));
    c->h_pin_cap = bytes + bytes / 4;
    return MQ_OK;
}

// ---- stage timing (CUDA events on the ctx stream) ----------------------------------------------
cudaEvent_t get_event(mq_ctx *c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct StageTimer {
    mq_ctx *c; std::string name; cudaEvent_t a, b; cudaStream_t st;
    StageTimer(mq_ctx *c_, const char *n, cudaStream_t s_ = nullptr) : c(c_), name(n), st(s_ ? s_ : c_->stream) {
        a = get_event(c); b = get_event(c); cudaEventRecord(a, st);
    }
    ~StageTimer() { cudaEventRecord(b, st); c->pending.push_back({name, a, b, c->timer_gen}); }
};
// a new call starts a new generation: per-call figures (mq_last_ms) restart, running totals (mq_total_ms) keep
// accumulating; nothing is synchronised here
void timers_reset(mq_ctx *c) { c->ms.clear(); c->timer_gen++; }
void timers_collect(mq_ctx *c) {
    for (auto &pr : c->pending) {
        float t = 0; cudaEventSynchronize(pr.b); cudaEventElapsedTime(&t, pr.a, pr.b);
        if (pr.gen == c->timer_gen) c->ms[pr.name] += t;
        c->ms_total[pr.name] += t;
        c->ev_pool.push_back(pr.a); c->ev_pool.push_back(pr.b);
    }
    c->pending.clear();
}

uint64_t hash_bound(double density) {   // (density as FH * H::MAX as FH) as H, saturating like Rust `as`
    double b = density * 18446744073709551615.0;
    if (!(b > 0.0)) return 0;
    if (b >= 18446744073709551616.0) return ~0ull;
    return (uint64_t)b;
}
uint64_t hrol(uint64_t x, unsigned r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }

void fill_tables(ScanTables &T, uint32_t l) {
    const uint64_t h[4] = {SEED_A, SEED_C, SEED_G, SEED_T}, hc[4] = {SEED_T, SEED_G, SEED_C, SEED_A};
    for (int c = 0; c < 4; c++) {
        T.h[c] = h[c]; T.hc[c] = hc[c];
        T.inF[c] = hrol(h[c], l - 1); T.outF[c] = hrol(h[c], 63);
        T.inR[c] = hc[c];             T.outR[c] = hrol(hc[c], l);
    }
    for (int i = 0; i < 4; i++) for (int o = 0; o < 4; o++) {
        T.pairF[i | (o << 2)] = T.inF[i] ^ T.outF[o];
        T.pairR[i | (o << 2)] = T.inR[i] ^ T.outR[o];
    }
}
void fill_tables_v2(ScanTablesV2 &T, const ScanTables &S, uint32_t l) {
    // v2 symbol codes are the raw bits (c>>1)&3: A=0 C=1 T=2 G=3; S is in A,C,G,T order
    static const int v1_of[4] = {0, 1, 3, 2};
    static_assert(offsetof(ScanTablesV2, pairR) == 128 && offsetof(ScanTablesV2, inF) == 256 && offsetof(ScanTablesV2, outF) == 288 &&
                  offsetof(ScanTablesV2, inR) == 320 && offsetof(ScanTablesV2, outR) == 352 && offsetof(ScanTablesV2, sel) == 400,
                  "the kernel addresses these tables by byte offset");
    for (int i = 0; i < 4; i++) {
        T.inF[i] = S.inF[v1_of[i]]; T.outF[i] = S.outF[v1_of[i]]; T.inR[i] = S.inR[v1_of[i]]; T.outR[i] = S.outR[v1_of[i]];
        for (int o = 0; o < 4; o++) {
            T.pairF[i + 4 * o] = S.pairF[v1_of[i] | (v1_of[o] << 2)];
            T.pairR[i + 4 * o] = S.pairR[v1_of[i] | (v1_of[o] << 2)];
        }
    }
    T.F0 = 0; T.R0 = 0;                       // window of l phantom 'A's
    for (uint32_t i = 0; i < l; i++) { T.F0 ^= hrol(SEED_A, l - 1 - i); T.R0 ^= hrol(SEED_T, i); }
    for (uint32_t p = 0; p < 16; p++) {       // byte-permute selectors: run-start bytes first, zero fill
        uint32_t sel = 0, j = 0;
        for (uint32_t b = 0; b < 4; b++) if (p & (1u << b)) sel |= b << (4 * j++);
        for (; j < 4; j++) sel |= 4u << (4 * j);
        T.sel[p] = sel;
    }
}

// scalars block layout (u64 words): 0 scan total, 1 counts[2] .. ; u32 view used for tickets
enum { SC_TOTAL = 0, SC_COUNT0 = 1, SC_COUNT1 = 2, SC_TICKET = 3 /* u32[2] : ticket, ovf */, SC_WORDS = 8 };

// exclusive scan of n u32 values in place; out gets n+1 entries when write_total; total -> host
int excl_scan(mq_ctx *c, uint32_t *data, uint64_t n, bool write_total, uint64_t *total_host) {
    uint32_t nb = (uint32_t)((n + SCAN_BLK * SCAN_ITEMS - 1) / (SCAN_BLK * SCAN_ITEMS));
    if (nb == 0) nb = 1;
    int rc = ensure(c, c->d_blocksums, (size_t)nb * 4); if (rc) return rc;
    uint64_t *d_total = c->d_scalars.as<uint64_t>() + SC_TOTAL;
    k_scan_block_sums<<<nb, SCAN_BLK, 0, c->stream>>>(data, n, c->d_blocksums.as<uint32_t>());
    k_scan_sums<<<1, SCAN_BLK, 0, c->stream>>>(c->d_blocksums.as<uint32_t>(), nb, d_total);
    k_scan_apply<<<nb, SCAN_BLK, 0, c->stream>>>(data, n, c->d_blocksums.as<uint32_t>(), data, write_total ? 1 : 0);
    c->launches += 3;
    CK(cudaGetLastError());
    if (total_host) {
        CK(cudaMemcpyAsync(total_host, d_total, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return MQ_OK;
}

// S1 on a device-resident batch.  Result: c->d_pos / c->d_hash (M entries), c->d_seq_off (n+1).
int run_scan(mq_ctx *c, const uint8_t *d_seqs, const uint64_t *d_offs, uint32_t n, uint32_t min_len,
             const uint32_t *d_pos_base, const uint32_t *d_emit_range, uint64_t *M_out) {
    int rc;
    *M_out = 0;
    if ((rc = ensure(c, c->d_first_tile, ((size_t)n + 2) * 4))) return rc;
    if ((rc = ensure(c, c->d_seq_off, ((size_t)n + 2) * 4))) return rc;
    uint64_t n_tiles64 = 0;
    {
        StageTimer t(c, "scan");
        if (c->scan_v1) k_tiles_per_seq<<<(n + 255) / 256, 256, 0, c->stream>>>(d_offs, n, min_len, c->d_first_tile.as<uint32_t>());
        else k_tiles_per_seq_v2<<<(n + 255) / 256, 256, 0, c->stream>>>(d_offs, n, min_len, c->d_first_tile.as<uint32_t>());
        c->launches++;
        if ((rc = excl_scan(c, c->d_first_tile.as<uint32_t>(), n, true, &n_tiles64))) return rc;
    }
    if (n_tiles64 >= (1ull << 31)) { c->err = "batch too large (tile count)"; return MQ_ERR_RANGE; }
    const uint32_t n_tiles = (uint32_t)n_tiles64;
    if (n_tiles == 0) {
        CK(cudaMemsetAsync(c->d_seq_off.p, 0, ((size_t)n + 1) * 4, c->stream));
        return MQ_OK;
    }
    if ((rc = ensure(c, c->d_tile_seq, (size_t)n_tiles * 4))) return rc;
    if ((rc = ensure(c, c->d_ev_hash, (size_t)n_tiles * EV_CAP * 8))) return rc;
    if ((rc = ensure(c, c->d_ev_meta, (size_t)n_tiles * EV_CAP * 4))) return rc;
    if ((rc = ensure(c, c->d_lane_cnt, (size_t)n_tiles * 32 * 2))) return rc;
    if ((rc = ensure(c, c->d_tile_cnt, ((size_t)n_tiles + 2) * 4))) return rc;

    uint64_t M = 0; uint32_t n_ovf = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        if ((rc = ensure(c, c->d_ovf_tile, (size_t)c->ovf_cap * 4))) return rc;
        if ((rc = ensure(c, c->d_ovf_meta, (size_t)c->ovf_cap * 4))) return rc;
        if ((rc = ensure(c, c->d_ovf_hash, (size_t)c->ovf_cap * 8))) return rc;
        uint32_t *tickets = (uint32_t *)(c->d_scalars.as<uint64_t>() + SC_TICKET);
        {
            StageTimer t(c, "scan");
            CK(cudaMemsetAsync(tickets, 0, 8, c->stream));
            k_tile_seq<<<(n_tiles + 255) / 256, 256, 0, c->stream>>>(c->d_first_tile.as<uint32_t>(), n, n_tiles, c->d_tile_seq.as<uint32_t>());
            ScanArgs a{};
            a.seqs = d_seqs; a.offs = d_offs; a.first_tile = c->d_first_tile.as<uint32_t>(); a.tile_seq = c->d_tile_seq.as<uint32_t>();
            a.n_tiles = n_tiles; a.l = c->p.l; a.use_hpc = c->p.use_hpc; a.bound = c->bound;
            a.ev_hash = c->d_ev_hash.as<uint64_t>(); a.ev_meta = c->d_ev_meta.as<uint32_t>();
            a.lane_cnt = c->d_lane_cnt.as<uint16_t>(); a.tile_cnt = c->d_tile_cnt.as<uint32_t>();
            a.ovf_count = tickets + 1; a.ovf_cap = c->ovf_cap;
            a.ovf_tile = c->d_ovf_tile.as<uint32_t>(); a.ovf_meta = c->d_ovf_meta.as<uint32_t>(); a.ovf_hash = c->d_ovf_hash.as<uint64_t>();
            a.tile_ticket = tickets; a.emit_range = d_emit_range;
            {
                StageTimer tk(c, "scan_kernel");   // the dominant kernel alone (roofline numerator)
                if (c->scan_v1) {
                    const uint32_t ctas_needed = (n_tiles + SCAN_WARPS - 1) / SCAN_WARPS;
                    const uint32_t grid = std::min<uint32_t>(ctas_needed, (uint32_t)c->n_sm * 6);
                    k_scan_minimizers<<<grid, SCAN_WARPS * 32, SCAN_WARPS * TILE_SMEM, c->stream>>>(a, c->tab);
                } else {
                    const uint32_t ctas_needed = (n_tiles + V2_WARPS - 1) / V2_WARPS;
                    size_t smem = (size_t)V2_WARPS * (c->scan_v2 ? V2_WARP_BYTES : V3_WARP_BYTES);
                    { const char *e = getenv("MQ_SCAN_PAD"); if (e) smem += (size_t)atoi(e); }   // occupancy experiments
                    if (c->v2_ctas_per_sm == 0) {      // persistent grid = every CTA the chip can hold
                        int nb = 0;
                        const void *kern = c->scan_v2 ? (const void *)k_scan_minimizers_v2 : (c->p.use_hpc ? (const void *)k_scan_minimizers_v3<true> : (const void *)k_scan_minimizers_v3<false>);
                        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, V2_WARPS * 32, smem) != cudaSuccess || nb < 1) nb = 1;
                        c->v2_ctas_per_sm = nb;
                        if (getenv("MQ_DEBUG")) fprintf(stderr, "[mq] scan kernel: %d CTAs/SM x %d warps, %zu B dynamic smem per CTA\n", nb, V2_WARPS, smem);
                    }
                    const uint32_t grid = std::min<uint32_t>(ctas_needed, (uint32_t)(c->n_sm * c->v2_ctas_per_sm));
                    if (c->scan_v2) k_scan_minimizers_v2<<<grid, V2_WARPS * 32, smem, c->stream>>>(a, c->tab2);
                    else if (c->p.use_hpc) k_scan_minimizers_v3<true><<<grid, V2_WARPS * 32, smem, c->stream>>>(a, c->tab3);
                    else k_scan_minimizers_v3<false><<<grid, V2_WARPS * 32, smem, c->stream>>>(a, c->tab3);
                }
            }
            c->launches += 2;
            c->scan_kernel_launches++;
            CK(cudaGetLastError());
        }
        {
            StageTimer t(c, "gather");
            if ((rc = excl_scan(c, c->d_tile_cnt.as<uint32_t>(), n_tiles, true, &M))) return rc;
        }
        CK(cudaMemcpyAsync(&n_ovf, tickets + 1, 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (n_ovf <= c->ovf_cap) break;
        c->ovf_cap = n_ovf + n_ovf / 4 + 1024;    // pool too small: grow and redo the scan
        if (attempt == 2) { c->err = "overflow pool kept overflowing"; return MQ_ERR_NOMEM; }
    }
    if (M >= (1ull << 32) - 64) { c->err = "batch too large (minimizer count)"; return MQ_ERR_RANGE; }
    if ((rc = ensure(c, c->d_pos, (M + 64) * 4))) return rc;
    if ((rc = ensure(c, c->d_hash, (M + 64) * 8))) return rc;
    {
        StageTimer t(c, "gather");
        GatherArgs g{};
        g.ev_hash = c->d_ev_hash.as<uint64_t>(); g.ev_meta = c->d_ev_meta.as<uint32_t>(); g.lane_cnt = c->d_lane_cnt.as<uint16_t>();
        // tile_cnt was scanned in place: the per-tile totals are recovered as differences
        g.tile_base = c->d_tile_cnt.as<uint32_t>();
        g.tile_seq = c->d_tile_seq.as<uint32_t>(); g.first_tile = c->d_first_tile.as<uint32_t>(); g.offs = d_offs;
        g.pos_base = d_pos_base; g.n_tiles = n_tiles; g.grid_align = c->scan_v1 ? 4u : 16u; g.out_pos = c->d_pos.as<uint32_t>(); g.out_hash = c->d_hash.as<uint64_t>();
        k_gather_minimizers<<<(n_tiles + 8 * GATHER_TPW - 1) / (8 * GATHER_TPW), 256, 0, c->stream>>>(g);
        c->launches++;
        if (n_ovf) {
            k_gather_overflow<<<(n_ovf + 255) / 256, 256, 0, c->stream>>>(g, c->d_ovf_tile.as<uint32_t>(), c->d_ovf_meta.as<uint32_t>(),
                                                                        c->d_ovf_hash.as<uint64_t>(), n_ovf);
            c->launches++;
        }
        k_seq_mini_off<<<(n + 1 + 255) / 256, 256, 0, c->stream>>>(c->d_first_tile.as<uint32_t>(), c->d_tile_cnt.as<uint32_t>(), n,
                                                                c->d_seq_off.as<uint32_t>());
        c->launches++;
        CK(cudaGetLastError());
    }
    *M_out = M;
    c->last_minimizers += M;
    return MQ_OK;
}

int upload_batch(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t i0, uint32_t i1) {
    const uint64_t b0 = offs[i0], b1 = offs[i1], nb = b1 - b0; const uint32_t n = i1 - i0;
    int rc;
    if ((rc = ensure(c, c->d_seqs, nb + PAD))) return rc;
    if ((rc = ensure(c, c->d_offs, ((size_t)n + 1) * 8))) return rc;
    if ((rc = ensure_pin(c, ((size_t)n + 1) * 8))) return rc;
    uint64_t *ho = (uint64_t *)c->h_pin;
    for (uint32_t i = 0; i <= n; i++) ho[i] = offs[i0 + i] - b0;
    StageTimer t(c, "h2d");
    if (nb) CK(cudaMemcpyAsync(c->d_seqs.p, seqs + b0, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync((uint8_t *)c->d_seqs.p + nb, 0, PAD, c->stream));
    CK(cudaMemcpyAsync(c->d_offs.p, ho, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    return MQ_OK;
}

int check_offs(mq_ctx *c, const uint64_t *offs, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        if (offs[i + 1] < offs[i]) { c->err = "offs not monotone"; return MQ_ERR_ARG; }
        if (offs[i + 1] - offs[i] >= (1ull << 31)) { c->err = "record of 2^31 bases or more"; return MQ_ERR_RANGE; }
    }
    return MQ_OK;
}

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

// map a device-resident batch: S1 -> probe/match -> chain
int map_device(mq_ctx *c, const uint8_t *d_seqs, const uint64_t *d_offs, uint32_t n, HitRec *d_hits) {
    int rc; uint64_t M = 0;
    if ((rc = run_scan(c, d_seqs, d_offs, n, c->p.l + c->p.k - 1, nullptr, nullptr, &M))) return rc;
    if ((rc = ensure(c, c->d_matches, (M + 64) * sizeof(MatchRec)))) return rc;
    if ((rc = ensure(c, c->d_nmatch, ((size_t)n + 1) * 4))) return rc;
    if ((rc = ensure(c, c->d_big_list, ((size_t)n + 2) * 4))) return rc;
    uint32_t *tickets = (uint32_t *)(c->d_scalars.as<uint64_t>() + SC_TICKET);
    Table t{c->d_table.as<Slot>(), c->tmask};
    {
        StageTimer tm(c, "probe");
        CK(cudaMemsetAsync(tickets, 0, 8, c->stream));
        ProbeArgs a{};
        a.pos = c->d_pos.as<uint32_t>(); a.hash = c->d_hash.as<uint64_t>(); a.seq_off = c->d_seq_off.as<uint32_t>();
        a.n_reads = n; a.k = c->p.k; a.l = c->p.l; a.matches = c->d_matches.as<MatchRec>(); a.n_matches = c->d_nmatch.as<uint32_t>();
        a.read_ticket = tickets;
        const uint32_t grid = std::min<uint32_t>((n + 3) / 4, (uint32_t)c->n_sm * 16);
        k_probe_match<<<grid, 128, 0, c->stream>>>(a, t);
        c->launches++;
        CK(cudaGetLastError());
    }
    {
        StageTimer tm(c, "chain");
        ChainArgs a{};
        a.matches = c->d_matches.as<MatchRec>(); a.n_matches = c->d_nmatch.as<uint32_t>(); a.seq_off = c->d_seq_off.as<uint32_t>();
        a.offs = d_offs; a.ref_lens = c->d_ref_lens.as<uint64_t>(); a.n_refs = c->n_refs; a.n_reads = n;
        a.c = c->p.c; a.s = c->p.s; a.g = c->p.g; a.hits = d_hits; a.read_ticket = tickets + 1;
        // thread per read for the usual handful of Matches; reads with many Matches are queued for the warp kernel
        a.big_list = c->d_big_list.as<uint32_t>(); a.big_count = c->d_big_list.as<uint32_t>() + n;
        CK(cudaMemsetAsync(a.big_count, 0, 4, c->stream));
        k_chain_small<<<(n + 127) / 128, 128, 0, c->stream>>>(a);
        const uint32_t grid = std::min<uint32_t>((n + 3) / 4, (uint32_t)c->n_sm * 8);
        k_chain<<<grid, 128, 0, c->stream>>>(a);
        c->launches += 2;
        CK(cudaGetLastError());
    }
    return MQ_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int mq_abi_version(void) { return 1; }

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

int mq_create(mq_ctx **out, const mq_params *p, int device) {
    if (!out || !p) return MQ_ERR_ARG;
    *out = nullptr;
    if (p->l < 2 || p->l > MQ_MAX_L || p->k < 1 || p->k > MQ_MAX_K || !(p->density >= 0.0)) return MQ_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) { cudaGetLastError(); return MQ_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return MQ_ERR_CUDA; }
    mq_ctx *c = new mq_ctx();
    c->p = *p; c->device = device; c->bound = hash_bound(p->density);
    fill_tables(c->tab, p->l);
    fill_tables_v2(c->tab2, c->tab, p->l);
    fill_tables_v3(c->tab3, c->tab2);
    { const char *e = getenv("MQ_SCAN_V1"); c->scan_v1 = e && e[0] == '1'; }
    { const char *e = getenv("MQ_SCAN_V2"); c->scan_v2 = e && e[0] == '1'; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->n_sm = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MQ_ERR_CUDA; }
    if (cudaMalloc(&c->d_scalars.p, SC_WORDS * 8) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return MQ_ERR_CUDA; }
    c->d_scalars.cap = SC_WORDS * 8;
    cudaMemsetAsync(c->d_scalars.p, 0, SC_WORDS * 8, c->stream);
    *out = c;
    return MQ_OK;
}

void mq_destroy(mq_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    timers_collect(c);
    DBuf *bufs[] = {&c->d_seqs, &c->d_offs, &c->d_first_tile, &c->d_tile_seq, &c->d_ev_hash, &c->d_ev_meta, &c->d_lane_cnt,
                    &c->d_tile_cnt, &c->d_blocksums, &c->d_scalars, &c->d_ovf_tile, &c->d_ovf_meta, &c->d_ovf_hash, &c->d_pos,
                    &c->d_hash, &c->d_seq_off, &c->d_matches, &c->d_nmatch, &c->d_hits, &c->d_pos_base, &c->d_emit_len, &c->d_big_list,
                    &c->d_misc, &c->st_pos, &c->st_hash, &c->d_table, &c->d_ref_lens};
    for (DBuf *b : bufs) dfree(*b);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->region_a) { cudaEventDestroy(c->region_a); cudaEventDestroy(c->region_b); }
    if (c->h_pin) cudaFreeHost(c->h_pin);
    for (int b = 0; b < 2; b++) {
        dfree(c->d_seqs2[b]); dfree(c->d_offs2[b]);
        if (c->h_offs2[b]) cudaFreeHost(c->h_offs2[b]);
        if (c->ev_copied[b]) cudaEventDestroy(c->ev_copied[b]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

void *mq_host_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void mq_host_free(void *p) { if (p) cudaFreeHost(p); }

void *mq_stream(mq_ctx *c) { return c ? (void *)c->stream : nullptr; }
int mq_sync(mq_ctx *c) { if (!c) return MQ_ERR_ARG; cudaSetDevice(c->device); CK(cudaStreamSynchronize(c->stream)); timers_collect(c); return MQ_OK; }
uint64_t mq_launch_count(mq_ctx *c) { return c ? c->launches : 0; }
double mq_total_ms(mq_ctx *c, const char *stage) {
    if (!c || !stage) return -1.0;
    timers_collect(c);
    auto it = c->ms_total.find(stage);
    return it == c->ms_total.end() ? 0.0 : it->second;
}
double mq_last_ms(mq_ctx *c, const char *stage) {
    if (!c || !stage) return -1.0;
    timers_collect(c);
    if (!strcmp(stage, "total")) { double s = 0; for (auto &kv : c->ms) if (kv.first != "scan_kernel") s += kv.second; return s; }
    auto it = c->ms.find(stage);
    return it == c->ms.end() ? 0.0 : it->second;
}
uint64_t mq_scan_kernel_launches(mq_ctx *c) { return c ? c->scan_kernel_launches : 0; }
uint64_t mq_minimizer_count(mq_ctx *c, int reset) { if (!c) return 0; uint64_t v = c->last_minimizers; if (reset) c->last_minimizers = 0; return v; }

// device-memory helpers so that callers can keep inputs resident in HBM without another runtime
void *mq_dev_alloc(mq_ctx *c, size_t bytes) {
    if (!c) return nullptr;
    cudaSetDevice(c->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void mq_dev_free(mq_ctx *c, void *p) { if (c && p) { cudaSetDevice(c->device); cudaFree(p); } }
int mq_dev_upload(mq_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || (bytes && (!dst || !src))) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MQ_OK;
}
int mq_dev_download(mq_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || (bytes && (!dst || !src))) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MQ_OK;
}
int mq_dev_memset(mq_ctx *c, void *dst, int value, size_t bytes) {
    if (!c || (bytes && !dst)) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemsetAsync(dst, value, bytes, c->stream));
    return MQ_OK;
}
// CUDA-event bracket on the ctx stream around any sequence of calls
int mq_region_begin(mq_ctx *c) {
    if (!c) return MQ_ERR_ARG;
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

uint64_t mq_table_bytes(mq_ctx *c) { return c && c->frozen ? (c->tmask + 2) * sizeof(Slot) : 0; }
uint64_t mq_table_slots(mq_ctx *c) { return c && c->frozen ? c->tmask + 1 : 0; }

// ---- index build ---------------------------------------------------------------------------------
int mq_index_add(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint32_t first_ref_idx, uint64_t *nb_mers_out) {
    if (!c || (!seqs && n) || !offs) return MQ_ERR_ARG;
    if (c->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    if (n == 0) return MQ_OK;
    if ((rc = upload_batch(c, seqs, offs, 0, n))) return rc;
    uint64_t M = 0;
    if ((rc = run_scan(c, c->d_seqs.as<uint8_t>(), c->d_offs.as<uint64_t>(), n, c->p.l + c->p.k - 1, nullptr, nullptr, &M))) return rc;
    if ((rc = store_append(c, M))) return rc;
    std::vector<uint32_t> so((size_t)n + 1);
    CK(cudaMemcpyAsync(so.data(), c->d_seq_off.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < n; i++) {
        uint64_t cnt = so[i + 1] - so[i];
        c->dir.push_back({(uint64_t)first_ref_idx + i, 0ull, cnt});
        if (nb_mers_out) nb_mers_out[i] = cnt >= c->p.k ? cnt - c->p.k + 1 : 0;
    }
    timers_collect(c);
    return MQ_OK;
}

int mq_index_add_segment(mq_ctx *c, const uint8_t *bytes, uint64_t n_bytes, uint32_t ref_idx, uint64_t ref_len,
                         uint64_t seg_start, uint64_t own_len) {
    if (!c || !bytes) return MQ_ERR_ARG;
    if (c->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
    if (ref_len >= (1ull << 31) || seg_start + own_len > ref_len) { c->err = "segment outside its record / record too long"; return MQ_ERR_RANGE; }
    cudaSetDevice(c->device);
    timers_reset(c);
    if (ref_len < (uint64_t)c->p.l + c->p.k - 1 || own_len == 0) { c->dir.push_back({(uint64_t)ref_idx, seg_start, 0ull}); return MQ_OK; }   // mers.rs:18
    const uint64_t ctx = seg_start > 0 ? 1 : 0;
    if (n_bytes < ctx + own_len) { c->err = "segment buffer shorter than ctx+own_len"; return MQ_ERR_ARG; }
    int rc;
    if ((rc = ensure(c, c->d_seqs, n_bytes + PAD))) return rc;
    if ((rc = ensure(c, c->d_offs, 2 * 8))) return rc;
    if ((rc = ensure(c, c->d_pos_base, 4))) return rc;
    if ((rc = ensure(c, c->d_emit_len, 8))) return rc;
    if ((rc = ensure_pin(c, 64))) return rc;
    {
        StageTimer t(c, "h2d");
        // the record handed to the scan INCLUDES the context byte, so run starts are decided exactly as
        // in a whole-record scan; only l-mers starting in [ctx, ctx+own_len) are emitted
        uint64_t *ho = (uint64_t *)c->h_pin; ho[0] = 0; ho[1] = n_bytes;
        uint32_t *hu = (uint32_t *)(ho + 2); hu[0] = (uint32_t)(seg_start - ctx); hu[1] = (uint32_t)ctx; hu[2] = (uint32_t)(ctx + own_len);
        CK(cudaMemcpyAsync(c->d_seqs.p, bytes, n_bytes, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync((uint8_t *)c->d_seqs.p + n_bytes, 0, PAD, c->stream));
        CK(cudaMemcpyAsync(c->d_offs.p, ho, 16, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_pos_base.p, hu, 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_emit_len.p, hu + 1, 8, cudaMemcpyHostToDevice, c->stream));
    }
    uint64_t M = 0;
    if ((rc = run_scan(c, c->d_seqs.as<uint8_t>(), c->d_offs.as<uint64_t>(), 1, 0, c->d_pos_base.as<uint32_t>(),
                       c->d_emit_len.as<uint32_t>(), &M))) return rc;
    if ((rc = store_append(c, M))) return rc;
    c->dir.push_back({(uint64_t)ref_idx, seg_start, M});
    CK(cudaStreamSynchronize(c->stream));
    timers_collect(c);
    return MQ_OK;
}

int mq_store_info(mq_ctx *c, uint64_t *n_minimizers, uint32_t *n_segments) {
    if (!c) return MQ_ERR_ARG;
    if (n_minimizers) *n_minimizers = c->st_n;
    if (n_segments) *n_segments = (uint32_t)c->dir.size();
    return MQ_OK;
}
int mq_store_export(mq_ctx *c, void **d_pos, void **d_hash, uint64_t *dir) {
    if (!c) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    if (d_pos) *d_pos = c->st_pos.p;
    if (d_hash) *d_hash = c->st_hash.p;
    if (dir) for (size_t i = 0; i < c->dir.size(); i++) { dir[3 * i] = c->dir[i][0]; dir[3 * i + 1] = c->dir[i][1]; dir[3 * i + 2] = c->dir[i][2]; }
    return MQ_OK;
}
int mq_store_import(mq_ctx *c, const void *d_pos, const void *d_hash, uint64_t n_min, const uint64_t *dir, uint32_t n_seg) {
    if (!c || (n_min && (!d_pos || !d_hash)) || (n_seg && !dir)) return MQ_ERR_ARG;
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
    c->dir.clear();
    for (uint32_t i = 0; i < n_seg; i++) c->dir.push_back({dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]});
    return MQ_OK;
}

int mq_index_freeze(mq_ctx *c, const uint64_t *ref_lens, uint32_t n_refs, uint64_t *n_unique, uint64_t *n_keys) {
    if (!c || (n_refs && !ref_lens)) return MQ_ERR_ARG;
    if (c->frozen) { c->err = "index already frozen"; return MQ_ERR_STATE; }
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
    std::vector<uint32_t> rec_off, rec_id, km_off;
    uint64_t acc = 0, n_tuples = 0;
    c->nb_mers.assign(n_refs, 0);
    for (size_t i = 0; i < ns;) {
        size_t j = i; uint64_t cnt = 0;
        while (j < ns && sd[j][0] == sd[i][0]) { cnt += sd[j][2]; j++; }
        if (sd[i][0] >= n_refs) { c->err = "segment ref_idx >= n_refs"; return MQ_ERR_ARG; }
        rec_off.push_back((uint32_t)acc); rec_id.push_back((uint32_t)sd[i][0]); km_off.push_back((uint32_t)n_tuples);
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
    if ((rc = ensure(c, c->d_misc, ((size_t)n_rec + 2) * 4 * 3))) return rc;
    uint32_t *d_rec_off = c->d_misc.as<uint32_t>(), *d_rec_id = d_rec_off + n_rec + 2;
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
        unsigned long long *d_cnt = (unsigned long long *)(c->d_scalars.as<uint64_t>() + SC_COUNT0);
        CK(cudaMemsetAsync(d_cnt, 0, 16, c->stream));
        k_table_count<<<c->n_sm * 8, 256, 0, c->stream>>>(c->d_table.as<Slot>(), cap + 1, d_cnt);
        c->launches++;
        CK(cudaGetLastError());
        uint64_t hc[2];
        CK(cudaMemcpyAsync(hc, d_cnt, 16, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->n_unique = hc[0]; c->n_keys = hc[1];
    }
    if (n_unique) *n_unique = c->n_unique;
    if (n_keys) *n_keys = c->n_keys;
    c->n_refs = n_refs; c->frozen = true;
    dfree(c->st_pos); dfree(c->st_hash); c->st_n = 0;
    timers_collect(c);
    return MQ_OK;
}

int mq_index_nb_mers(mq_ctx *c, uint64_t *nb, uint32_t n_refs) {
    if (!c || !nb) return MQ_ERR_ARG;
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
    if (!c->frozen) { c->err = "index not frozen"; return MQ_ERR_STATE; }
    cudaSetDevice(c->device);
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
    int rc = ensure_pin(c, CH);
    if (rc) { fclose(f); return rc; }
    const size_t total = (size_t)h.slots * sizeof(Slot);
    for (size_t off = 0; off < total && ok; off += CH) {
        const size_t n = std::min(CH, total - off);
        if (cudaMemcpy(c->h_pin, (const uint8_t *)c->d_table.p + off, n, cudaMemcpyDeviceToHost) != cudaSuccess) { fclose(f); c->err = "D2H table"; return MQ_ERR_CUDA; }
        ok = fwrite(c->h_pin, 1, n, f) == n;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { c->err = std::string("short write to ") + path; return MQ_ERR_ARG; }
    return MQ_OK;
}
int mq_index_load(mq_ctx *c, const char *path, uint64_t *ref_lens_out, uint32_t ref_cap, uint32_t *n_refs_out, char *names_out,
                  uint64_t names_cap, uint64_t *names_bytes_out, uint64_t *n_unique_out) {
    if (!c || !path) return MQ_ERR_ARG;
    if (c->frozen || c->st_n || !c->dir.empty()) { c->err = "load needs a fresh context"; return MQ_ERR_STATE; }
    cudaSetDevice(c->device);
    FILE *f = fopen(path, "rb");
    if (!f) { c->err = std::string("cannot open ") + path; return MQ_ERR_ARG; }
    IndexFileHeader h{};
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "MQB200IX", 8) != 0 || h.version != 1) { fclose(f); c->err = "not a mapquik_b200 index file"; return MQ_ERR_ARG; }
    if (h.k != c->p.k || h.l != c->p.l || h.use_hpc != c->p.use_hpc || h.density != c->p.density) {
        fclose(f); c->err = "index was built with different k / l / density / hpc"; return MQ_ERR_ARG;
    }
    if (h.slots < 2 || ((h.slots - 1) & (h.slots - 2)) != 0) { fclose(f); c->err = "corrupt index header"; return MQ_ERR_ARG; }
    if (n_refs_out) *n_refs_out = (uint32_t)h.n_refs;
    if (names_bytes_out) *names_bytes_out = h.names_bytes;
    if (n_unique_out) *n_unique_out = h.n_unique;
    std::vector<uint64_t> lens(h.n_refs), nb(h.n_refs);
    bool ok = true;
    if (h.n_refs) ok = fread(lens.data(), 8, h.n_refs, f) == h.n_refs && fread(nb.data(), 8, h.n_refs, f) == h.n_refs;
    if (ref_lens_out) { if (ref_cap < h.n_refs) { fclose(f); c->err = "ref_lens_out too small"; return MQ_ERR_ARG; } memcpy(ref_lens_out, lens.data(), h.n_refs * 8); }
    if (h.names_bytes) {
        if (names_out) { if (names_cap < h.names_bytes) { fclose(f); c->err = "names_out too small"; return MQ_ERR_ARG; } ok = ok && fread(names_out, 1, h.names_bytes, f) == h.names_bytes; }
        else ok = ok && fseek(f, (long)h.names_bytes, SEEK_CUR) == 0;
    }
    int rc;
    const size_t total = (size_t)h.slots * sizeof(Slot), CH = 64u << 20;
    if ((rc = ensure(c, c->d_table, total))) { fclose(f); return rc; }
    if ((rc = ensure(c, c->d_ref_lens, (h.n_refs + 1) * 8))) { fclose(f); return rc; }
    if ((rc = ensure_pin(c, CH))) { fclose(f); return rc; }
    for (size_t off = 0; off < total && ok; off += CH) {
        const size_t n = std::min(CH, total - off);
        ok = fread(c->h_pin, 1, n, f) == n;
        if (ok && cudaMemcpy((uint8_t *)c->d_table.p + off, c->h_pin, n, cudaMemcpyHostToDevice) != cudaSuccess) { fclose(f); c->err = "H2D table"; return MQ_ERR_CUDA; }
    }
    fclose(f);
    if (!ok) { c->err = "truncated index file"; return MQ_ERR_ARG; }
    if (h.n_refs && cudaMemcpy(c->d_ref_lens.p, lens.data(), h.n_refs * 8, cudaMemcpyHostToDevice) != cudaSuccess) { c->err = "H2D ref_lens"; return MQ_ERR_CUDA; }
    c->tmask = h.slots - 2; c->n_refs = (uint32_t)h.n_refs; c->n_unique = h.n_unique; c->n_keys = h.n_keys;
    c->nb_mers.assign(nb.begin(), nb.end());
    c->frozen = true;
    return MQ_OK;
}

// ---- mapping ---------------------------------------------------------------------------------------
int mq_map_batch_device(mq_ctx *c, const uint8_t *d_seqs, const uint64_t *d_offs, uint32_t n, uint64_t total_bytes, mq_hit *d_out) {
    if (!c || !d_offs || !d_out || (!d_seqs && total_bytes)) return MQ_ERR_ARG;
    if (!c->frozen) { c->err = "index not frozen"; return MQ_ERR_STATE; }
    if (((uintptr_t)d_seqs & 15) != 0) { c->err = "device sequence buffer must be 16-byte aligned"; return MQ_ERR_ARG; }
    cudaSetDevice(c->device);
    timers_reset(c);
    if (n == 0) return MQ_OK;
    int rc = map_device(c, d_seqs, d_offs, n, (HitRec *)d_out);
    return rc;
}

// upload sub-batch [i0, i1) into double-buffer slot b on the copy stream
static int upload_async(mq_ctx *c, int b, const uint8_t *seqs, const uint64_t *offs, uint32_t i0, uint32_t i1) {
    const uint64_t b0 = offs[i0], nb = offs[i1] - b0; const uint32_t m = i1 - i0;
    int rc;
    if ((rc = ensure(c, c->d_seqs2[b], nb + PAD))) return rc;
    if ((rc = ensure(c, c->d_offs2[b], ((size_t)m + 1) * 8))) return rc;
    if (c->h_offs2_cap[b] < ((size_t)m + 1) * 8) {
        if (c->h_offs2[b]) cudaFreeHost(c->h_offs2[b]);
        c->h_offs2[b] = nullptr; c->h_offs2_cap[b] = 0;
        size_t want = ((size_t)m + 1) * 8 * 2;
        CK(cudaMallocHost(&c->h_offs2[b], want));
        c->h_offs2_cap[b] = want;
    }
    uint64_t *ho = (uint64_t *)c->h_offs2[b];
    for (uint32_t i = 0; i <= m; i++) ho[i] = offs[i0 + i] - b0;
    {
        StageTimer t(c, "h2d", c->copy_stream);
        if (nb) CK(cudaMemcpyAsync(c->d_seqs2[b].p, seqs + b0, nb, cudaMemcpyHostToDevice, c->copy_stream));
        CK(cudaMemsetAsync((uint8_t *)c->d_seqs2[b].p + nb, 0, PAD, c->copy_stream));
        CK(cudaMemcpyAsync(c->d_offs2[b].p, ho, ((size_t)m + 1) * 8, cudaMemcpyHostToDevice, c->copy_stream));
    }
    CK(cudaEventRecord(c->ev_copied[b], c->copy_stream));
    return MQ_OK;
}

int mq_map_batch(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, mq_hit *out) {
    if (!c || !offs || (!seqs && n) || (!out && n)) return MQ_ERR_ARG;
    if (!c->frozen) { c->err = "index not frozen"; return MQ_ERR_STATE; }
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    if (n == 0) return MQ_OK;
    if (!c->copy_stream) {
        CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_copied[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_copied[1], cudaEventDisableTiming));
    }
    // sub-batch boundaries: about an eighth of the batch, between 128 MB and 1 GB -- small enough that the first
    // (un-overlapped) upload is short, large enough that per-sub-batch launches and scalar read-backs stay negligible
    const uint64_t total_bytes = offs[n] - offs[0];
    const uint64_t sub_bytes = std::min<uint64_t>(std::max<uint64_t>(total_bytes / 8, MAP_SUB_BATCH_BYTES), 1ull << 30);
    std::vector<uint32_t> cut{0};
    for (uint32_t i0 = 0; i0 < n;) {
        uint32_t i1 = i0 + 1;
        while (i1 < n && offs[i1 + 1] - offs[i0] <= sub_bytes) i1++;
        cut.push_back(i1); i0 = i1;
    }
    const size_t ns = cut.size() - 1;
    // H2D of sub-batch i+1 runs on the copy stream while sub-batch i computes
    if ((rc = upload_async(c, 0, seqs, offs, cut[0], cut[1]))) return rc;
    for (size_t i = 0; i < ns; i++) {
        const int b = (int)(i & 1);
        const uint32_t m = cut[i + 1] - cut[i];
        if (i + 1 < ns && (rc = upload_async(c, 1 - b, seqs, offs, cut[i + 1], cut[i + 2]))) return rc;
        CK(cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
        if ((rc = ensure(c, c->d_hits, (size_t)m * sizeof(HitRec)))) return rc;
        if ((rc = map_device(c, c->d_seqs2[b].as<uint8_t>(), c->d_offs2[b].as<uint64_t>(), m, c->d_hits.as<HitRec>()))) return rc;
        {
            StageTimer t(c, "d2h");
            CK(cudaMemcpyAsync(out + cut[i], c->d_hits.p, (size_t)m * sizeof(HitRec), cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));      // slot b and d_hits are free again
    }
    CK(cudaStreamSynchronize(c->copy_stream));
    timers_collect(c);
    return MQ_OK;
}

int mq_format_paf(char *buf, size_t cap, const char *q_id, uint64_t q_len, const char *r_id, uint64_t r_len, const mq_hit *h) {
    if (!buf || !q_id || !r_id || !h) return MQ_ERR_ARG;
    int w = snprintf(buf, cap, "%s\t%llu\t%llu\t%llu\t%s\t%s\t%llu\t%llu\t%llu\t%llu\t%llu\t%u", q_id, (unsigned long long)q_len,
                     (unsigned long long)h->q_start, (unsigned long long)h->q_end, h->rc ? "-" : "+", r_id, (unsigned long long)r_len,
                     (unsigned long long)h->r_start, (unsigned long long)h->r_end, (unsigned long long)h->score,
                     (unsigned long long)r_len, (unsigned)h->mapq);
    return (w < 0 || (size_t)w >= cap) ? MQ_ERR_ARG : w;
}

// ---- introspection -----------------------------------------------------------------------------------
int mq_minimizers(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint64_t *seq_off, uint32_t *pos,
                  uint64_t *hash, uint64_t cap, uint64_t *n_total) {
    if (!c || !offs || (!seqs && n)) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    uint64_t M = 0;
    if (n) {
        if ((rc = upload_batch(c, seqs, offs, 0, n))) return rc;
        if ((rc = run_scan(c, c->d_seqs.as<uint8_t>(), c->d_offs.as<uint64_t>(), n, 0, nullptr, nullptr, &M))) return rc;
    }
    if (n_total) *n_total = M;
    if (seq_off && n) {
        std::vector<uint32_t> so((size_t)n + 1);
        CK(cudaMemcpyAsync(so.data(), c->d_seq_off.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (uint32_t i = 0; i <= n; i++) seq_off[i] = so[i];
    }
    if (pos && hash && M) {
        if (cap < M) { c->err = "output capacity too small"; return MQ_ERR_ARG; }
        CK(cudaMemcpyAsync(pos, c->d_pos.p, M * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(hash, c->d_hash.p, M * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    timers_collect(c);
    return MQ_OK;
}

int mq_kminmers(mq_ctx *c, const uint8_t *seqs, const uint64_t *offs, uint32_t n, uint64_t *seq_off, uint32_t *start,
                uint32_t *end, uint32_t *offrev, uint64_t *hash, uint64_t cap, uint64_t *n_total) {
    if (!c || !offs || (!seqs && n)) return MQ_ERR_ARG;
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    uint64_t M = 0, Q = 0;
    std::vector<uint32_t> so((size_t)n + 1, 0), ko((size_t)n + 1, 0);
    if (n) {
        if ((rc = upload_batch(c, seqs, offs, 0, n))) return rc;
        if ((rc = run_scan(c, c->d_seqs.as<uint8_t>(), c->d_offs.as<uint64_t>(), n, c->p.l + c->p.k - 1, nullptr, nullptr, &M))) return rc;
        CK(cudaMemcpyAsync(so.data(), c->d_seq_off.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
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
}

int mq_index_get(mq_ctx *c, const uint64_t *hashes, uint64_t n, uint8_t *found, uint32_t *id, uint32_t *start, uint32_t *end,
                 uint32_t *offset, uint8_t *rc_out) {
    if (!c || (n && (!hashes || !found || !id || !start || !end || !offset || !rc_out))) return MQ_ERR_ARG;
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
    if (!c || !offs || (!seqs && n)) return MQ_ERR_ARG;
    if (!c->frozen) return MQ_ERR_STATE;
    cudaSetDevice(c->device);
    timers_reset(c);
    int rc;
    if ((rc = check_offs(c, offs, n))) return rc;
    if (n_total) *n_total = 0;
    if (n == 0) { if (match_off) match_off[0] = 0; return MQ_OK; }
    if ((rc = upload_batch(c, seqs, offs, 0, n))) return rc;
    if ((rc = ensure(c, c->d_hits, (size_t)n * sizeof(HitRec)))) return rc;
    if ((rc = map_device(c, c->d_seqs.as<uint8_t>(), c->d_offs.as<uint64_t>(), n, c->d_hits.as<HitRec>()))) return rc;
    std::vector<uint32_t> so((size_t)n + 1), nm(n);
    CK(cudaMemcpyAsync(so.data(), c->d_seq_off.p, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
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
}

}  // extern "C"
