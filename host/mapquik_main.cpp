// mapquik (B200) -- C++ host CLI over the C ABI of libmapquik_b200.so.
//
// Mirrors the reference's command line, console lines and PAF output (src/main.rs:77-272,
// src/closures.rs:22-211) for the seeding->chaining path:
//     mapquik <reads.fa|fq[.gz|.lz4]> --reference <ref.fa[.gz|.lz4]> [-k K] [-l L] [-d D] [-c C] [-s S] [-g G]
//             [-p PREFIX] [--nohpc] [--threads N] [-b B] [-q Q] [--low-memory] [--nosimd]
//             [--parallelfastx] [--debug] [--gpu ID | --gpus N] [--ascii] [--save-index F] [--load-index F] [--rescue k,l,d]
// Host I/O: a reader thread (plus a worker pool, plus a decompressor thread for gz / lz4) parses records straight into
// pinned batch buffers while the main thread maps the previous batch on the GPU(s) and writes its PAF lines in input
// order (closures.rs:117-123).  Where the reference upper-cases a copy of every record (closures.rs:63,106), the
// block-parallel parser PACKS it (mq_pack_at: 2-bit codes + exception intervals, upper-casing folded in), so a quarter
// of the bytes cross PCIe; --ascii keeps one byte per base.
// All compute happens in the library.
#include "../include/mapquik_b200.h"

#include <fcntl.h>
#include <sys/resource.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <dlfcn.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Opt {
    std::string reads, reference, prefix, save_index, load_index, rescue, devices;
    bool has_prefix = false, debug = false, low_memory = false, nosimd = false, nohpc = false, pfx = false, parse_only = false;
    long k = -1, l = -1, c = -1, s = -1, g = -1, threads = -1, b = -1, q = -1;
    double density = -1;
    int gpu = 0, gpus = 0;
    bool ascii = false;
};

[[noreturn]] void die(const std::string &m) { fprintf(stderr, "%s\n", m.c_str()); exit(1); }

bool is_fasta_name(const std::string &f) {   // main.rs:196,202 (same substring test)
    auto ends = [&](const char *x) { size_t n = strlen(x); return f.size() >= n && f.compare(f.size() - n, n, x) == 0; };
    return f.find(".fasta.") != std::string::npos || ends(".fna") || f.find(".fna.") != std::string::npos ||
           f.find(".fa.") != std::string::npos || ends(".fa") || ends(".fasta");
}

// copy + closures.rs:63,106 `to_ascii_uppercase` in one pass (auto-vectorised)
inline void copy_upper(uint8_t *dst, const char *src, size_t n) {
    for (size_t i = 0; i < n; i++) { uint8_t c = (uint8_t)src[i]; dst[i] = (uint8_t)(c - ((c >= 'a' && c <= 'z') ? 32 : 0)); }
}

// growable byte buffer in pinned host memory (mq_host_alloc) so that mq_map_batch uploads at full PCIe speed
bool g_use_pinned = true;        // --parse-only (a host-only self-test) uses plain malloc instead
bool g_pack = true;              // block-parallel parser emits the packed input format (off: --ascii, --parse-only)

// ---- lz4 frame input (main.rs:68,71).  liblz4 ships without headers here, so the four entry points of its stable
// frame API are resolved at run time; a missing library is reported when an .lz4 file is actually opened.
struct Lz4Api {
    size_t (*create)(void **, unsigned) = nullptr;
    size_t (*free_ctx)(void *) = nullptr;
    size_t (*decompress)(void *, void *, size_t *, const void *, size_t *, const void *) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    bool ok = false;
    Lz4Api() {
        void *h = dlopen("liblz4.so.1", RTLD_NOW);
        if (!h) h = dlopen("liblz4.so", RTLD_NOW);
        if (!h) return;
        create = (decltype(create))dlsym(h, "LZ4F_createDecompressionContext");
        free_ctx = (decltype(free_ctx))dlsym(h, "LZ4F_freeDecompressionContext");
        decompress = (decltype(decompress))dlsym(h, "LZ4F_decompress");
        is_error = (decltype(is_error))dlsym(h, "LZ4F_isError");
        ok = create && free_ctx && decompress && is_error;
    }
};
struct Lz4Reader {
    static Lz4Api &api() { static Lz4Api a; return a; }
    int fd; void *ctx = nullptr; std::vector<char> in; size_t ipos = 0, ilen = 0; bool in_eof = false;
    explicit Lz4Reader(int fd_) : fd(fd_), in(4u << 20) {
        if (!api().ok) die("cannot read .lz4 input: liblz4 not found");
        if (api().is_error(api().create(&ctx, 100))) die("LZ4F_createDecompressionContext failed");
    }
    ~Lz4Reader() { if (ctx) api().free_ctx(ctx); }
    long read(char *dst, size_t cap) {          // decompressed bytes (0 at end of stream)
        size_t out = 0;
        while (out == 0) {
            if (ipos == ilen && !in_eof) {
                ssize_t r = ::read(fd, in.data(), in.size());
                if (r <= 0) in_eof = true; else { ipos = 0; ilen = (size_t)r; }
            }
            if (ipos == ilen && in_eof) return 0;
            size_t dn = cap, sn = ilen - ipos;
            const size_t rc = api().decompress(ctx, dst, &dn, in.data() + ipos, &sn, nullptr);
            if (api().is_error(rc)) die("corrupt lz4 stream");
            ipos += sn; out = dn;
        }
        return (long)out;
    }
};
struct PinnedBuf {
    uint8_t *p = nullptr; size_t size = 0, cap = 0;
    static uint8_t *alloc(size_t n) { return g_use_pinned ? (uint8_t *)mq_host_alloc(n) : (uint8_t *)malloc(n); }
    static void release(uint8_t *q) { if (g_use_pinned) mq_host_free(q); else free(q); }
    ~PinnedBuf() { if (p) release(p); }
    void reserve(size_t want) {
        if (want <= cap) return;
        // pinning costs ~0.5 s per GB (cudaHostAlloc), so a slot is sized to what it must hold, not to a growth schedule
        size_t nc = cap ? std::max(want, cap + cap / 2) : want;
        nc = (nc + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
        uint8_t *np = alloc(nc);
        if (!np) die("pinned host allocation failed");
        if (size) memcpy(np, p, size);
        if (p) release(p);
        p = np; cap = nc;
    }
    void clear() { size = 0; }
};

// An input file: plain (regular files are mapped), gzip (zlib) or lz4 frames (main.rs:68,71).  Everything that is not a
// plain regular file is a byte stream that a feeder thread decompresses ahead of the parser.
struct Input {
    gzFile f = nullptr; int fd = -1, bgzf_fd = -1; bool fasta;
    std::unique_ptr<Lz4Reader> lz4;
    size_t file_bytes = 0;                 // size of a plain regular file
    bool regular = false;                  // plain regular file with a known size: mmap
    Input(const std::string &path, bool fasta_) : fasta(fasta_) {
        unsigned char magic[2] = {0, 0};
        fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) die("Error opening file: " + path);
        ssize_t got = pread(fd, magic, 2, 0);
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {        // gzip: hand the descriptor to zlib
            bgzf_fd = dup(fd);                                          // (the feeder maps the file instead if it turns out to be BGZF)
            f = gzdopen(fd, "rb");
            if (!f) die("Error opening compressed file: " + path);
            gzbuffer(f, 1 << 20);
            fd = -1;
        } else if (path.size() > 4 && path.compare(path.size() - 4, 4, ".lz4") == 0) {   // main.rs:68
            lz4.reset(new Lz4Reader(fd));
        } else {
            posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
            struct stat st; if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) { file_bytes = (size_t)st.st_size; regular = true; }
        }
    }
    ~Input() { if (f) gzclose(f); if (fd >= 0) close(fd); if (bgzf_fd >= 0) close(bgzf_fd); }
    long read_some(char *dst, size_t cap) {      // next (decompressed) bytes of a stream; <= 0 at its end
        return f ? (long)gzread(f, dst, (unsigned)std::min<size_t>(cap, 1u << 30)) : lz4 ? lz4->read(dst, cap) : (long)read(fd, dst, cap);
    }
    static std::string id_of(const char *p, size_t n) {       // seq_io record.id(): the header up to the first SPACE
        size_t e = 1; while (e < n && p[e] != ' ') e++;
        return std::string(p + 1, e - 1);
    }
};

int g_parse_threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));

// persistent workers: a block is a few milliseconds of work per phase, thread creation would be a tenth of it.
// Never destroyed (die() may exit from inside a worker).
struct Pool {
    std::mutex m, run_m; std::condition_variable cv_go, cv_done;
    std::vector<std::thread> th;
    const std::function<void(int)> *fn = nullptr; int n = 0, next = 0, pending = 0;
    void worker() {
        for (;;) {
            int t;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_go.wait(lk, [&] { return next < n; });
                t = next++;
            }
            (*fn)(t);
            { std::lock_guard<std::mutex> lk(m); if (--pending == 0) cv_done.notify_all(); }
        }
    }
    void run(int n_, const std::function<void(int)> &f) {
        if (n_ <= 1) { if (n_ == 1) f(0); return; }
        std::lock_guard<std::mutex> one(run_m);
        {
            std::lock_guard<std::mutex> lk(m);
            while ((int)th.size() < n_ - 1) { th.emplace_back([this] { worker(); }); th.back().detach(); }
            fn = &f; next = 1; pending = n_ - 1; n = n_;
        }
        cv_go.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
        n = 0; next = 0;
    }
};
Pool &the_pool() { static Pool *pool = new Pool; return *pool; }
template <class Fn> void parallel_for(int n, Fn fn) {
    const std::function<void(int)> f = fn;
    the_pool().run(n, f);
}

// BGZF (bgzip, htslib): a multi-member gzip file whose members are <= 64 KB and carry their own compressed size in a
// 'BC' extra field, so members inflate independently -- here on the worker pool instead of zlib's one thread.
struct Bgzf {
    const uint8_t *base = nullptr; size_t size = 0, pos = 0;
    static bool header(const uint8_t *p, size_t avail, size_t *block_bytes, size_t *data_off) {
        if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
        const size_t xlen = p[10] | ((size_t)p[11] << 8);
        if (avail < 12 + xlen) return false;
        for (size_t o = 12; o + 4 <= 12 + xlen;) {
            const size_t slen = p[o + 2] | ((size_t)p[o + 3] << 8);
            if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) {
                *block_bytes = (size_t)(p[o + 4] | ((size_t)p[o + 5] << 8)) + 1;
                *data_off = 12 + xlen;
                return (p[3] & ~4) == 0 && *block_bytes >= *data_off + 8;      // no name / comment / header CRC fields (bgzip writes none)
            }
            o += 4 + slen;
        }
        return false;
    }
    bool open(int fd) {
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 28) return false;
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        size_t bb, off;
        if (!header((const uint8_t *)m, (size_t)st.st_size, &bb, &off)) { munmap(m, (size_t)st.st_size); return false; }
        base = (const uint8_t *)m; size = (size_t)st.st_size;
        madvise(m, size, MADV_SEQUENTIAL);
        return true;
    }
    ~Bgzf() { if (base) munmap((void *)base, size); }
    // inflate the next members into dst (at most cap bytes); returns the bytes written, *end = the file is exhausted
    size_t fill(char *dst, size_t cap, bool *end) {
        struct Mem { size_t src, csize, out, isize; uint32_t crc; };
        std::vector<Mem> ms; size_t out = 0;
        while (pos < size) {
            size_t bb, off;
            if (!header(base + pos, size - pos, &bb, &off) || pos + bb > size) die("corrupt BGZF member (or a plain gzip member inside a BGZF file)");
            const uint8_t *tr = base + pos + bb - 8;
            const uint32_t crc = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((uint32_t)tr[3] << 24);
            const size_t isize = tr[4] | (tr[5] << 8) | (tr[6] << 16) | ((size_t)tr[7] << 24);
            if (isize > cap) die("corrupt BGZF member (size)");
            if (out + isize > cap) break;
            ms.push_back({pos + off, bb - off - 8, out, isize, crc});
            out += isize; pos += bb;
        }
        *end = pos >= size;
        const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)g_parse_threads, ms.size() / 4 + 1));
        std::vector<char> bad(T, 0);
        parallel_for(T, [&](int t) {
            z_stream zs; memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { bad[t] = 1; return; }
            for (size_t i = ms.size() * (size_t)t / T, e = ms.size() * (size_t)(t + 1) / T; i < e; i++) {
                const Mem &m = ms[i];
                if (m.isize == 0) continue;
                inflateReset(&zs);
                zs.next_in = (Bytef *)(base + m.src); zs.avail_in = (uInt)m.csize;
                zs.next_out = (Bytef *)(dst + m.out); zs.avail_out = (uInt)m.isize;
                if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.avail_out != 0 ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)(dst + m.out), (uInt)m.isize) != m.crc) { bad[t] = 1; break; }
            }
            inflateEnd(&zs);
        });
        for (char b : bad) if (b) die("corrupt BGZF member (inflate / CRC)");
        return out;
    }
};

// decompresses a stream in chunks ahead of the parser (inflate is the slowest stage of a .gz input: it gets a thread of its own)
struct Feeder {
    static constexpr size_t CHUNK = 8u << 20; static constexpr int DEPTH = 12;
    Input *in; std::mutex m; std::condition_variable cv; std::thread th;
    struct Chunk { std::unique_ptr<char[]> p; size_t n = 0; };
    std::vector<Chunk> ring; size_t head = 0, tail = 0;      // [head, tail) are full
    bool done = false, quit = false;
    std::unique_ptr<Bgzf> bgzf;
    explicit Feeder(Input *in_) : in(in_), ring(DEPTH) {
        if (in->bgzf_fd >= 0 && !getenv("MQ_CLI_NO_BGZF")) { bgzf.reset(new Bgzf); if (!bgzf->open(in->bgzf_fd)) bgzf.reset(); }
        th = std::thread([this] { run(); });
    }
    ~Feeder() { { std::lock_guard<std::mutex> lk(m); quit = true; } cv.notify_all(); th.join(); }
    void run() {
        for (;;) {
            Chunk *c;
            { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return quit || tail - head < (size_t)DEPTH; }); if (quit) return; c = &ring[tail % DEPTH]; }
            if (!c->p) c->p.reset(new char[CHUNK]);
            size_t n = 0; bool end = false;
            if (bgzf) n = bgzf->fill(c->p.get(), CHUNK, &end);
            else {
                while (n < CHUNK) { const long r = in->read_some(c->p.get() + n, CHUNK - n); if (r <= 0) { end = true; break; } n += (size_t)r; }
            }
            c->n = n;
            { std::lock_guard<std::mutex> lk(m); if (n) tail++; if (end) done = true; }
            cv.notify_all();
            if (end) return;
        }
    }
    // the oldest full chunk (nullptr at the end of the stream); release() hands it back
    const Chunk *front() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return head < tail || done; });
        return head < tail ? &ring[head % DEPTH] : nullptr;
    }
    void release() { { std::lock_guard<std::mutex> lk(m); head++; } cv.notify_all(); }
};

// one batch of records; two of them rotate between the reader thread and the GPU
struct Batch {
    PinnedBuf seqs; std::vector<uint64_t> offs{0}; std::vector<std::string> ids; bool last = false;
    // packed form (block-parallel parser): 2-bit codes, block bitmap, exception intervals -- seqs is then unused
    bool packed = false; PinnedBuf words, flags; std::vector<mq_exc> exc; uint64_t n_bases = 0;
    void clear() { seqs.clear(); words.clear(); flags.clear(); exc.clear(); packed = false; n_bases = 0; offs.assign(1, 0); ids.clear(); last = false; }
    mq_packed view() const { mq_packed v; v.words = (const uint32_t *)words.p; v.flags = (const uint32_t *)flags.p; v.exc = exc.data(); v.n_exc = exc.size(); v.n_bases = n_bases; return v; }
    // ASCII bytes of record i (rescue pass, --parse-only)
    void record_bytes(size_t i, std::vector<uint8_t> &out) const {
        const size_t n = (size_t)(offs[i + 1] - offs[i]);
        out.resize(n);
        if (!packed) { memcpy(out.data(), seqs.p + offs[i], n); return; }
        mq_unpack((const uint32_t *)words.p, exc.data(), exc.size(), offs[i], n, out.data());
    }
};
struct BatchQueue {          // single producer / single consumer over two slots
    std::mutex m; std::condition_variable cv; Batch slot[2]; int filled[2] = {0, 0};
};
// ---- block-parallel parser for plain (uncompressed) files ---------------------------------------------------
// The file is consumed in blocks that are cut at record boundaries.  Per block:
//   (1) worker threads index the newlines of their slice of the block (and note which lines start with '>' and whether
//       any line ends in CR) -- the only pass besides (4) that reads the file bytes;
//   (2) worker threads turn line ends into a running base count per line (`cum`: bases of the block before line i;
//       header / '+' / quality lines count 0), so that any base of the batch can be located by binary search;
//   (3) one pass over the RECORDS (header lines only, not every line) fills ids and offsets;
//   (4) worker threads each own a range of the batch's bases, cut on 2,048-base units (one word of the packed format's
//       block bitmap, so no two threads share a word and nothing is zeroed or OR-ed), gather the line pieces of their
//       range through a small staging buffer and pack 64-base-aligned chunks with mq_pack_at at full SIMD width -- a
//       60-column reference FASTA packs as fast as single-line reads; with --ascii they upper-case the pieces in place.
// offsets (relative to `origin`) of the '\n' bytes of [p, e), appended to v; calls on_nl(q) for each (q = its address).
// Short lines (a 60-column FASTA has a newline in nearly every 64-byte block) make one memchr call per line expensive:
// with AVX2 the block's newline mask is computed once and its bits are walked.
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline uint64_t nl_mask64(const char *p) {
    const __m256i nl = _mm256_set1_epi8('\n');
    const uint32_t lo = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)p), nl));
    const uint32_t hi = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i *)(p + 32)), nl));
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
const bool g_avx2 = __builtin_cpu_supports("avx2") && !getenv("MQ_CLI_NO_AVX2");
#else
inline uint64_t nl_mask64(const char *) { return 0; }
const bool g_avx2 = false;
#endif
template <class OnNl> void scan_newlines(const char *p, const char *e, OnNl on_nl) {
    // 64 KB at a time; a chunk is walked by masks when the one before it averaged a newline per < 256 bytes, else by memchr
    // (which is faster than two compares per 64 bytes when lines are long)
    bool dense = false;
    while (p < e) {
        const char *ce = std::min(e, p + 65536);
        size_t found = 0;
        if (dense && g_avx2) {
            for (; p + 64 <= ce; p += 64) {
                uint64_t m = nl_mask64(p);
                while (m) { on_nl(p + __builtin_ctzll(m)); m &= m - 1; found++; }
            }
        }
        while (p < ce) {
            const char *q = (const char *)memchr(p, '\n', (size_t)(ce - p));
            if (!q) { p = ce; break; }
            on_nl(q); found++;
            p = q + 1;
        }
        dense = found > 256;
    }
}

struct BlockParser {
    // A plain file is mapped once; a block is a window [pos, pos + block_bytes) of the mapping that next_batch() advances
    // past the records it consumed -- no read() copy and no carry-over of an incomplete trailing record.  A stream
    // (gz / lz4 / pipe) is decompressed ahead by a Feeder; its block is a buffer that keeps the unconsumed tail.
    bool fasta; size_t block_bytes;
    const char *base = nullptr; size_t file_size = 0, pos = 0;             // mapped file
    Feeder *feed = nullptr; std::unique_ptr<char[]> sbuf; size_t s_cap = 0, s_have = 0, s_off = 0; bool s_end = false;   // stream
    const char *raw = nullptr; size_t fill = 0; bool eof = false;
    bool populate = getenv("MQ_CLI_POPULATE") != nullptr;     // measured on the 16-core box: 28 GB/s with, 29.5 without -- off unless asked for
    std::vector<std::vector<uint32_t>> nlv, hdv;     // per slice: newline offsets; local indices of lines FOLLOWED by a header line
    std::vector<uint32_t> nl, cum;            // per line: offset of its '\n' (or of the end of the block); bases before it
    BlockParser(int fd, bool fasta_, size_t block_bytes_, size_t file_size_) : fasta(fasta_), block_bytes(block_bytes_), file_size(file_size_) {
        if (file_size) {
            void *m = mmap(nullptr, file_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) die("mmap failed");
            base = (const char *)m;
            madvise(m, file_size, MADV_SEQUENTIAL);
        }
    }
    BlockParser(Feeder *feed_, bool fasta_, size_t block_bytes_) : fasta(fasta_), block_bytes(block_bytes_), feed(feed_) { populate = false; }
    ~BlockParser() { if (base) munmap((void *)base, file_size); }
    void top_up() {                                   // position the window
        if (!feed) {
            fill = std::min(block_bytes, file_size - pos);
            raw = base + pos;
            eof = pos + fill >= file_size;
            return;
        }
        if (s_cap < block_bytes) {                    // (first call, or the window was enlarged)
            std::unique_ptr<char[]> nb(new char[block_bytes]);
            if (s_have) memcpy(nb.get(), sbuf.get(), s_have);
            sbuf = std::move(nb); s_cap = block_bytes;
        }
        while (s_have < block_bytes && !s_end) {
            const Feeder::Chunk *c = feed->front();
            if (!c) { s_end = true; break; }
            const size_t take = std::min(c->n - s_off, block_bytes - s_have);
            memcpy(sbuf.get() + s_have, c->p.get() + s_off, take);
            s_have += take; s_off += take;
            if (s_off == c->n) { feed->release(); s_off = 0; }
        }
        raw = sbuf.get(); fill = s_have; eof = s_end;
    }
    void advance(size_t consumed) {                   // the block's first `consumed` bytes are done with
        if (!feed) {
            // the consumed part of the mapping is not needed again: drop it from this process's resident set
            const size_t pg = 4096, a0 = pos & ~(pg - 1), a1 = (pos + consumed) & ~(pg - 1);
            if (a1 > a0) madvise((void *)(base + a0), a1 - a0, MADV_DONTNEED);
            pos += consumed;
            return;
        }
        if (consumed < s_have) memmove(sbuf.get(), sbuf.get() + consumed, s_have - consumed);
        s_have -= consumed;
    }
    size_t line_start(size_t i) const { return i ? (size_t)nl[i - 1] + 1 : 0; }
    // pieces of the batch's bases [d0, d1): fn(source pointer, length, first base)
    template <class Fn> void for_pieces(size_t n_lines, uint64_t d0, uint64_t d1, Fn fn) const {
        if (d0 >= d1) return;
        size_t i = (size_t)(std::upper_bound(cum.begin(), cum.begin() + n_lines + 1, (uint32_t)d0) - cum.begin()) - 1;   // cum[i] <= d0 < cum[i+1]
        uint64_t d = d0;
        while (d < d1) {
            while (cum[i + 1] == cum[i]) i++;
            const uint64_t e = std::min<uint64_t>(cum[i + 1], d1);
            fn(raw + line_start(i) + (d - cum[i]), (size_t)(e - d), d);
            d = e; i++;
        }
    }
    // fills B with the records of the next block; false when the file is exhausted
    bool next_batch(Batch &B) {
        B.clear();
        static const bool timing = getenv("MQ_CLI_TIMING") != nullptr;
        auto T0 = std::chrono::steady_clock::now();
        auto lap = [&](const char *what) {
            if (!timing) return;
            auto t = std::chrono::steady_clock::now();
            fprintf(stderr, "[parser] %-8s %.1f ms\n", what, std::chrono::duration<double, std::milli>(t - T0).count());
            T0 = t;
        };
        for (;;) {
            top_up();
            if (fill == 0) return false;
            // (1) newline index; a line is a header iff it starts with '>' (FASTA)
            const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)g_parse_threads, fill / (2u << 20) + 1));
            if ((int)nlv.size() < T) { nlv.resize(T); hdv.resize(T); }      // (members: their capacity is reused from block to block)
            for (int t = 0; t < T; t++) { nlv[t].clear(); hdv[t].clear(); }
            std::vector<char> crv(T, 0);
            parallel_for(T, [&](int t) {
                size_t a = fill * (size_t)t / T, b = fill * (size_t)(t + 1) / T;
                const char *p = raw + a, *e = raw + b, *end = raw + fill;
#ifdef MADV_POPULATE_READ
                if (populate) {                        // map the slice's pages in one call instead of one fault per 16 pages (Linux >= 5.14; ignored elsewhere)
                    const uintptr_t pa = (uintptr_t)p & ~(uintptr_t)4095, pe = ((uintptr_t)e + 4095) & ~(uintptr_t)4095;
                    if (madvise((void *)pa, pe - pa, MADV_POPULATE_READ) != 0) populate = false;
                }
#endif
                auto &v = nlv[t]; auto &h = hdv[t];
                bool cr = false;
                scan_newlines(p, e, [&](const char *q) {
                    if (q > raw && q[-1] == '\r') cr = true;
                    if (fasta && q + 1 < end && q[1] == '>') h.push_back((uint32_t)v.size());
                    v.push_back((uint32_t)(q - raw));
                });
                crv[t] = cr;
            });
            std::vector<size_t> lbase(T + 1, 0);
            for (int t = 0; t < T; t++) lbase[t + 1] = lbase[t] + nlv[t].size();
            size_t n_lines = lbase[T];
            bool last_open = false;                   // last line without '\n' (only at the end of the file)
            {
                uint32_t last_nl = 0; bool any = false;
                for (int t = T - 1; t >= 0 && !any; t--) if (!nlv[t].empty()) { last_nl = nlv[t].back(); any = true; }
                last_open = eof && (!any || (size_t)last_nl != fill - 1);
            }
            if (nl.size() < n_lines + 2) nl.resize(n_lines + 2 + n_lines / 4);
            if (cum.size() < n_lines + 3) cum.resize(n_lines + 3 + n_lines / 4);
            parallel_for(T, [&](int t) { if (!nlv[t].empty()) memcpy(nl.data() + lbase[t], nlv[t].data(), nlv[t].size() * 4); });
            if (last_open) {
                if (fill > 0 && raw[fill - 1] == '\r') crv[T - 1] = 1;
                nl[n_lines++] = (uint32_t)fill;
            }
            bool any_cr = false; for (char c : crv) any_cr |= c != 0;
            // header lines, in order (global line indices)
            std::vector<uint32_t> hdr;
            if (fasta) {
                if (raw[0] == '>') hdr.push_back(0);
                for (int t = 0; t < T; t++) for (uint32_t li : hdv[t]) hdr.push_back((uint32_t)(lbase[t] + li + 1));
                while (!hdr.empty() && hdr.back() >= n_lines) hdr.pop_back();      // a '>' that is the very last byte of the block
            }
            lap("newlines");
            // stripped end of line i
            auto line_end = [&](size_t i) -> size_t {
                size_t e = nl[i]; const size_t a = line_start(i);
                if (any_cr && e > a && raw[e - 1] == '\r') e--;
                return e;
            };
            // which records does the block hold, and how many of its lines belong to them
            size_t n_rec = 0, used_lines = 0, consumed = 0;
            if (fasta) {
                n_rec = hdr.size();
                used_lines = n_lines;
                if (eof) consumed = fill;
                else if (n_rec >= 2) { n_rec--; used_lines = hdr[n_rec]; consumed = line_start(used_lines); }   // the last record may continue in the next block
                else n_rec = 0;
            } else {
                n_rec = n_lines / 4;
                used_lines = 4 * n_rec;
                consumed = eof ? fill : (n_rec ? (size_t)nl[used_lines - 1] + 1 : 0);
            }
            if (n_rec == 0 && !eof) {                     // not even one complete record in the window: enlarge it
                if (block_bytes >= 0xFFFFFFF0ull) die("record larger than 4 GB");        // line offsets are 32-bit
                block_bytes = (size_t)std::min<uint64_t>((uint64_t)block_bytes * 2, 0xFFFFFFF0ull);
                continue;
            }
            // (2) bases before every line: per-thread running sums over line ranges, then the thread bases
            const int T1 = (int)std::max<size_t>(1, std::min<size_t>((size_t)g_parse_threads, used_lines / 65536 + 1));
            std::vector<uint64_t> tsum(T1 + 1, 0);
            const size_t first_seq_line = fasta ? (hdr.empty() ? used_lines : (size_t)hdr[0]) : 0;   // FASTA: lines before the first header belong to no record
            auto line_bases = [&](size_t i, size_t &hp) -> uint32_t {      // hp: cursor into hdr (FASTA)
                if (fasta) {
                    if (i < first_seq_line) return 0;
                    while (hp < hdr.size() && hdr[hp] < i) hp++;
                    if (hp < hdr.size() && hdr[hp] == i) return 0;
                } else if ((i & 3) != 1) return 0;
                return (uint32_t)(line_end(i) - line_start(i));
            };
            parallel_for(T1, [&](int t) {
                const size_t l0 = used_lines * (size_t)t / T1, l1 = used_lines * (size_t)(t + 1) / T1;
                size_t hp = fasta ? (size_t)(std::lower_bound(hdr.begin(), hdr.end(), (uint32_t)l0) - hdr.begin()) : 0;
                uint64_t s = 0;
                for (size_t i = l0; i < l1; i++) { cum[i] = (uint32_t)s; s += line_bases(i, hp); }
                tsum[t + 1] = s;
            });
            for (int t = 0; t < T1; t++) tsum[t + 1] += tsum[t];
            const uint64_t dst = tsum[T1];
            if (dst >= 0xFFFFFFFFull) die("block holds 2^32 bases or more");
            parallel_for(T1, [&](int t) {
                if (t == 0 || tsum[t] == 0) return;
                const size_t l0 = used_lines * (size_t)t / T1, l1 = used_lines * (size_t)(t + 1) / T1;
                const uint32_t add = (uint32_t)tsum[t];
                for (size_t i = l0; i < l1; i++) cum[i] += add;
            });
            cum[used_lines] = (uint32_t)dst; cum[used_lines + 1] = (uint32_t)dst + 1;       // sentinel: the piece walk stops on a non-empty "line"
            lap("lines");
            // (3) records
            B.ids.reserve(n_rec); B.offs.reserve(n_rec + 1);
            if (fasta) {
                for (size_t r = 0; r < n_rec; r++) {
                    const size_t li = hdr[r];
                    B.ids.push_back(Input::id_of(raw + line_start(li), line_end(li) - line_start(li)));
                    B.offs.push_back(r + 1 < hdr.size() && hdr[r + 1] <= used_lines ? cum[hdr[r + 1]] : (uint32_t)dst);
                }
            } else {
                for (size_t r = 0; r < n_rec; r++) {
                    const size_t a = line_start(4 * r), n = line_end(4 * r) - a;
                    if (!n || raw[a] != '@') die("malformed FASTQ record");
                    B.ids.push_back(Input::id_of(raw + a, n));
                    B.offs.push_back(cum[4 * r + 2]);
                }
            }
            lap("records");
            // (4) sequence bytes into the pinned batch: packed (2 bits per base, upper-casing folded in) or upper-cased ASCII
            if (g_pack) {
                const size_t cap_bases = std::max<size_t>(dst, block_bytes) + 64;
                const size_t wbytes = (size_t)mq_packed_words(cap_bases) * 4, fbytes = (size_t)mq_packed_flag_words(cap_bases) * 4;
                B.words.reserve(wbytes); B.flags.reserve(fbytes);            // sized once per slot
                B.words.size = (size_t)mq_packed_words(dst) * 4; B.flags.size = (size_t)mq_packed_flag_words(dst) * 4;
                B.packed = true; B.n_bases = dst;
                lap("reserve");
                const uint64_t units = (dst + 2047) / 2048;
                const int T2 = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)g_parse_threads, units / 1024 + 1));
                std::vector<std::vector<mq_exc>> ex(T2);
                uint32_t *words = (uint32_t *)B.words.p, *flags = (uint32_t *)B.flags.p;
                const char *const raw_end = raw + fill;
                parallel_for(T2, [&](int t) {
                    const uint64_t u0 = units * (uint64_t)t / T2, u1 = units * (uint64_t)(t + 1) / T2;
                    const uint64_t d0 = u0 * 2048, d1 = std::min<uint64_t>(dst, u1 * 2048);
                    memset(flags + u0, 0, (size_t)(u1 - u0) * 4);
                    if (t + 1 == T2) {                 // the ragged tail is OR-ed in, and the slack behind the last base reads as zero
                        const uint64_t tail = dst & ~31ull;
                        memset(words + (tail >> 4), 0, B.words.size - (size_t)(tail >> 4) * 4);
                        memset(flags + units, 0, B.flags.size - (size_t)units * 4);
                    }
                    std::vector<mq_exc> tmp(256); auto &out = ex[t];
                    auto pack = [&](const uint8_t *src, uint64_t n, uint64_t at) {
                        if (!n) return;
                        uint64_t ne = 0;
                        int rc = mq_pack_at(src, n, at, words, flags, tmp.data(), tmp.size(), &ne, 1);
                        if (rc == MQ_ERR_RANGE) {      // more exception intervals than the scratch list holds: the codes are in place, list them again
                            tmp.resize((size_t)ne + 16);
                            rc = mq_pack_at(src, n, at, words, flags, tmp.data(), tmp.size(), &ne, 1);
                        }
                        if (rc != MQ_OK) die("mq_pack_at failed");
                        for (uint64_t k = 0; k < ne; k++) {       // an interval that continues the previous chunk's last one is merged
                            const mq_exc &e = tmp[k];
                            if (!out.empty() && out.back().byte == e.byte && out.back().start + out.back().len == e.start &&
                                (uint64_t)out.back().len + e.len <= 0xFFFFFFFFull) out.back().len += e.len;
                            else out.push_back(e);
                        }
                    };
                    // staging: bases [sb, sb + sn) of the batch, sb a multiple of 64
                    constexpr size_t STAGE = 8192;
                    alignas(64) uint8_t stage[STAGE + 64];
                    uint64_t sb = d0; size_t sn = 0;
                    auto flush = [&](bool all) {
                        const size_t n = all ? sn : (sn & ~(size_t)63);
                        pack(stage, n, sb);
                        if (n < sn) memmove(stage, stage + n, sn - n);
                        sb += n; sn -= n;
                    };
                    for_pieces(used_lines, d0, d1, [&](const char *src, size_t len, uint64_t) {
                        while (len) {
                            if (len >= 256) {              // long piece: top the staging buffer up to a 64-base boundary, then straight from the file
                                if (sn & 63) { const size_t take = 64 - (sn & 63); memcpy(stage + sn, src, take); sn += take; src += take; len -= take; }
                                flush(false);
                                const size_t n = len & ~(size_t)63;
                                pack((const uint8_t *)src, n, sb);
                                sb += n; src += n; len -= n;
                                continue;
                            }
                            if (len <= 64 && sn + 64 <= STAGE && src + 64 <= raw_end) {      // a line of a wrapped FASTA: one fixed-size copy (the surplus is overwritten)
                                memcpy(stage + sn, src, 64);
                                sn += len; len = 0;
                                if (sn == STAGE) flush(false);
                                continue;
                            }
                            const size_t take = std::min(len, STAGE - sn);
                            memcpy(stage + sn, src, take);
                            sn += take; src += take; len -= take;
                            if (sn == STAGE) flush(false);
                        }
                    });
                    flush(true);
                });
                for (auto &v : ex) B.exc.insert(B.exc.end(), v.begin(), v.end());     // thread order == position order
            } else {
                B.seqs.reserve(std::max<size_t>(dst, block_bytes) + 64); B.seqs.size = dst;     // sequence bytes never exceed the block: one allocation per slot
                lap("reserve");
                const int T2 = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)g_parse_threads, dst / (2u << 20) + 1));
                parallel_for(T2, [&](int t) {
                    for_pieces(used_lines, dst * (uint64_t)t / T2, dst * (uint64_t)(t + 1) / T2,
                               [&](const char *src, size_t len, uint64_t d) { copy_upper(B.seqs.p + d, src, len); });
                });
            }
            lap("copy");
            advance(consumed);
            B.last = eof && consumed == fill;
            return true;
        }
    }
};

void reader_thread(Input *in, BatchQueue *q, size_t batch_bytes) {
    int b = 0;
    size_t blk = in->regular ? std::min<size_t>(batch_bytes, in->file_bytes + 1) : batch_bytes;
    if (const char *e = getenv("MQ_CLI_BLOCK")) blk = (size_t)atol(e);      // tests: tiny blocks exercise the carry logic
    blk = std::max<size_t>(blk, 64);
    std::unique_ptr<Feeder> feed;
    std::unique_ptr<BlockParser> bp;
    if (in->regular) bp.reset(new BlockParser(in->fd, in->fasta, blk, in->file_bytes));
    else { feed.reset(new Feeder(in)); bp.reset(new BlockParser(feed.get(), in->fasta, blk)); }
    for (;;) {
        { std::unique_lock<std::mutex> lk(q->m); q->cv.wait(lk, [&] { return !q->filled[b]; }); }
        Batch &B = q->slot[b];
        bool more = true;
        if (!bp->next_batch(B)) { B.clear(); more = false; }
        else more = !B.last;
        B.last = !more;
        { std::lock_guard<std::mutex> lk(q->m); q->filled[b] = 1; }
        q->cv.notify_all();
        if (!more) return;
        b ^= 1;
    }
}
// run fn(batch) over every batch of the file; parsing of batch i+1 overlaps fn(batch i)
template <class Fn> void for_each_batch(const std::string &path, bool fasta, size_t batch_bytes, Fn fn) {
    Input in(path, fasta);
    BatchQueue q;
    std::thread th(reader_thread, &in, &q, batch_bytes);
    int b = 0;
    for (;;) {
        { std::unique_lock<std::mutex> lk(q.m); q.cv.wait(lk, [&] { return q.filled[b] != 0; }); }
        Batch &B = q.slot[b];
        const bool last = B.last;
        if (!B.ids.empty()) fn(B);
        { std::lock_guard<std::mutex> lk(q.m); q.filled[b] = 0; }
        q.cv.notify_all();
        if (last) break;
        b ^= 1;
    }
    th.join();
}

double secs(std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); }

void ck(mq_ctx *c, int rc, const char *what) {
    if (rc != MQ_OK) die(std::string(what) + ": " + mq_strerror(rc) + " (" + (c ? mq_last_error(c) : "") + ")");
}

}  // namespace

int main(int argc, char **argv) {
    auto t_start = std::chrono::steady_clock::now();
    Opt o;
    // structopt also takes `-k5` and `--density=0.01`: split those spellings into flag + value first
    std::vector<std::string> args;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        const size_t eq = a.find('=');
        if (a.size() > 2 && a[0] == '-' && a[1] == '-' && eq != std::string::npos) { args.push_back(a.substr(0, eq)); args.push_back(a.substr(eq + 1)); }
        else if (a.size() > 2 && a[0] == '-' && a[1] != '-' && strchr("kldcsgpbq", a[1]) && (isdigit((unsigned char)a[2]) || a[2] == '.' || a[1] == 'p')) {
            args.push_back(a.substr(0, 2)); args.push_back(a.substr(2));
        } else args.push_back(a);
    }
    for (size_t i = 0; i < args.size(); i++) {
        std::string a = args[i];
        auto val = [&](const char *name) -> std::string { if (i + 1 >= args.size()) die(std::string("missing value for ") + name); return args[++i]; };
        if (a == "--debug") o.debug = true;
        else if (a == "--low-memory") o.low_memory = true;
        else if (a == "--nosimd") o.nosimd = true;
        else if (a == "--nohpc") o.nohpc = true;
        else if (a == "--parallelfastx") o.pfx = true;
        else if (a == "-p" || a == "--prefix") { o.prefix = val("prefix"); o.has_prefix = true; }
        else if (a == "-k" || a == "--k") o.k = atol(val("k").c_str());
        else if (a == "-l" || a == "--l") o.l = atol(val("l").c_str());
        else if (a == "-d" || a == "--density") o.density = atof(val("density").c_str());
        else if (a == "-c" || a == "--chain") o.c = atol(val("chain").c_str());
        else if (a == "-s" || a == "--seed") o.s = atol(val("seed").c_str());
        else if (a == "-g" || a == "--gap-diff") o.g = atol(val("gap-diff").c_str());
        else if (a == "--reference") o.reference = val("reference");
        else if (a == "--threads") o.threads = atol(val("threads").c_str());
        else if (a == "-b" || a == "--b") o.b = atol(val("b").c_str());
        else if (a == "-q" || a == "--q") o.q = atol(val("q").c_str());
        else if (a == "--gpu") o.gpu = atoi(val("gpu").c_str());
        else if (a == "--gpus") o.gpus = atoi(val("gpus").c_str());           // extension: devices 0..N-1 in one run (mq_create_multi)
        else if (a == "--devices") o.devices = val("devices");               // extension: explicit device list "0,1,3" (a device may repeat)
        else if (a == "--ascii") o.ascii = true;                             // extension: one byte per base to the GPU instead of the packed format
        else if (a == "--save-index") o.save_index = val("save-index");      // extensions: the reference has no on-disk index
        else if (a == "--load-index") o.load_index = val("load-index");
        else if (a == "--parse-only") o.parse_only = true;                  // (testing) parse the reads, print a digest, exit
        else if (a == "--rescue") o.rescue = val("rescue");                  // "k,l,density": second pass over the unmapped reads
        else if (a == "-h" || a == "--help") { printf("mapquik <reads> --reference <ref> [-k -l -d -c -s -g -p --nohpc --threads --gpu ID | --gpus N --ascii]\n"); return 0; }
        else if (!a.empty() && a[0] == '-') die("unknown option " + a);
        else o.reads = a;
    }
    if (o.reads.empty()) die("Please specify an input file.");
    if (o.threads > 0) g_parse_threads = (int)std::min<long>(o.threads, 64);
    if (o.ascii) g_pack = false;
    if (o.parse_only) {
        g_use_pinned = false; g_pack = getenv("MQ_CLI_PACK") != nullptr;     // the digest is over the ASCII bytes either way
        // FNV-1a over "id\n" + sequence + "\n" of every record: identical for every container format / block size
        uint64_t h = 1469598103934665603ull, nrec = 0, nbase = 0;
        auto mixb = [&](const uint8_t *p, size_t n) { for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; } };
        std::vector<uint8_t> rec;
        static const bool no_digest = getenv("MQ_CLI_NODIGEST") != nullptr;     // parser throughput runs: count only
        const size_t bb = getenv("MQ_CLI_BATCH") ? (size_t)atol(getenv("MQ_CLI_BATCH")) : (size_t)(256u << 20);
        for_each_batch(o.reads, is_fasta_name(o.reads), bb, [&](Batch &B) {
            if (no_digest) { nrec += B.ids.size(); nbase += B.offs.back(); return; }
            for (size_t i = 0; i < B.ids.size(); i++) {
                mixb((const uint8_t *)B.ids[i].data(), B.ids[i].size()); mixb((const uint8_t *)"\n", 1);
                B.record_bytes(i, rec);
                mixb(rec.data(), rec.size()); mixb((const uint8_t *)"\n", 1);
                nrec++; nbase += B.offs[i + 1] - B.offs[i];
            }
        });
        printf("records %llu bases %llu digest %016llx\n", (unsigned long long)nrec, (unsigned long long)nbase, (unsigned long long)h);
        return 0;
    }
    if (o.reference.empty() && o.load_index.empty()) die("Please specify a reference file.");
    long k = 5, l = 31, c = 4, s = 11, g = 2000, b = 1, q = 200;
    double density = 0.01;
    const bool reads_fasta = is_fasta_name(o.reads), ref_fasta = o.reference.empty() || is_fasta_name(o.reference);
    if (reads_fasta) { printf("Input file: %s\nFormat: FASTA\n", o.reads.c_str()); }
    if (ref_fasta && !o.reference.empty()) { printf("Reference file: %s\nFormat: FASTA\n", o.reference.c_str()); }
    if (o.k >= 0) k = o.k; else printf("Warning: Using default k value (%ld).\n", k);                                  // main.rs:207-217
    if (o.l >= 0) l = o.l; else printf("Warning: Using default l value (%ld).\n", l);
    if (o.b >= 0) b = o.b; else printf("Warning: Using default buffer size (%ldX).\n", b);
    if (o.q >= 0) q = o.q; else printf("Warning: Using default queue length (%ld).\n", q);
    if (o.density >= 0) density = o.density; else printf("Warning: Using default density value (%g%%).\n", density * 100.0);
    if (o.threads < 0) printf("Warning: Using default number of threads (8).\n");
    if (o.c >= 0) c = o.c; else printf("Warning: Using default minimum chain length (%ld).\n", c);
    if (o.s >= 0) s = o.s; else printf("Warning: Using default minimum number of matching seeds (%ld).\n", s);
    if (o.g >= 0) g = o.g; else printf("Warning: Using default maximum seed gap difference (%ld).\n", g);
    char pre[256]; snprintf(pre, sizeof pre, "mapquik-k%ld-d%g-l%ld", k, density, l);
    std::string prefix = pre;
    if (o.has_prefix) prefix = o.prefix; else printf("Warning: Using default output prefix (%s).\n", prefix.c_str());
    printf(o.nohpc ? "Using regular ntHash (not HPC), CUDA sm_100a\n" : "Using HPC ntHash, CUDA sm_100a\n");
    (void)b; (void)q;

    mq_params p; p.k = (uint32_t)k; p.l = (uint32_t)l; p.density = density; p.use_hpc = o.nohpc ? 0 : 1;
    p.c = (uint32_t)c; p.s = (uint32_t)s; p.g = (uint32_t)g;
    mq_ctx *ctx = nullptr;
    std::vector<int> devs;
    for (int d = 0; d < o.gpus; d++) devs.push_back(d);
    if (!o.devices.empty()) { devs.clear(); for (const char *q = o.devices.c_str(); *q;) { devs.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q) q++; } }
    auto create = [&](mq_ctx **out, const mq_params *pp) { return devs.size() > 1 ? mq_create_multi(out, pp, devs.data(), (int)devs.size()) : mq_create(out, pp, devs.size() == 1 ? devs[0] : o.gpu); };
    ck(nullptr, create(&ctx, &p), "mq_create");

    std::string paf_name = prefix + ".paf";
    FILE *paf = fopen(paf_name.c_str(), "w");
    if (!paf) die("Couldn't create " + paf_name);
    std::vector<char> pafbuf(1 << 22); setvbuf(paf, pafbuf.data(), _IOFBF, pafbuf.size());

    // ---- reference ------------------------------------------------------------------------------------
    auto t_idx = std::chrono::steady_clock::now();
    std::vector<std::string> ref_names; std::vector<uint64_t> ref_lens;
    uint64_t n_unique = 0;
    if (!o.load_index.empty()) {
        // header layout of mq_index_save (include/mapquik_b200.h): magic[8] u32 version,k,l,hpc  f64 density  u64 slots,
        // n_unique,n_keys,n_refs,names_bytes
        struct { char magic[8]; uint32_t version, k, l, hpc; double density; uint64_t slots, n_unique, n_keys, n_refs, names_bytes; } h;
        FILE *f = fopen(o.load_index.c_str(), "rb");
        if (!f || fread(&h, sizeof h, 1, f) != 1) die("cannot read index file " + o.load_index);
        fclose(f);
        ref_lens.resize(h.n_refs);
        std::vector<char> names(h.names_bytes + 1);
        uint32_t n_refs = 0; uint64_t nb = 0;
        ck(ctx, mq_index_load(ctx, o.load_index.c_str(), ref_lens.data(), (uint32_t)ref_lens.size(), &n_refs, names.data(), h.names_bytes, &nb,
                              &n_unique), "mq_index_load");
        const char *q = names.data(), *e = names.data() + h.names_bytes;
        for (uint32_t i = 0; i < n_refs; i++) { ref_names.emplace_back(q < e ? q : ""); q += ref_names.back().size() + 1; }
        printf("Loaded index %s: %u references.\n", o.load_index.c_str(), n_refs);
    } else {
        for_each_batch(o.reference, ref_fasta, o.low_memory ? (64u << 20) : (128u << 20), [&](Batch &B) {
            std::vector<uint64_t> nb(B.ids.size());
            if (B.packed) { const mq_packed v = B.view(); ck(ctx, mq_index_add_packed(ctx, &v, B.offs.data(), (uint32_t)B.ids.size(), (uint32_t)ref_names.size(), nb.data()), "mq_index_add_packed"); }
            else ck(ctx, mq_index_add(ctx, B.seqs.p, B.offs.data(), (uint32_t)B.ids.size(), (uint32_t)ref_names.size(), nb.data()), "mq_index_add");
            for (size_t i = 0; i < B.ids.size(); i++) {
                printf("Indexed reference %s: %llu k-min-mers.\n", B.ids[i].c_str(), (unsigned long long)nb[i]);        // closures.rs:58
                ref_names.push_back(B.ids[i]); ref_lens.push_back(B.offs[i + 1] - B.offs[i]);
            }
        });
        ck(ctx, mq_index_freeze(ctx, ref_lens.data(), (uint32_t)ref_lens.size(), &n_unique, nullptr), "mq_index_freeze");
        if (!o.save_index.empty()) {
            std::string blob;
            for (auto &n : ref_names) { blob += n; blob.push_back('\0'); }
            ck(ctx, mq_index_save(ctx, o.save_index.c_str(), blob.data(), blob.size()), "mq_index_save");
        }
    }
    printf("Indexed %llu unique k-min-mers in %.6fs.\n", (unsigned long long)n_unique, secs(t_idx));                 // closures.rs:92

    // ---- reads ----------------------------------------------------------------------------------------
    auto t_map = std::chrono::steady_clock::now();
    PinnedBuf un_seqs; std::vector<uint64_t> un_offs{0}; std::vector<std::string> un_ids;      // unmapped reads (--rescue)
    {
        std::vector<mq_hit> hits; std::vector<char> line(1 << 16); std::vector<uint8_t> rec_tmp;
        // 64 MB batches: two pinned slots cost ~70 ms to allocate (256 MB ones cost 0.33 s, more than parsing 1 GB), and the
        // GPU side of a batch (H2D 1.3 ms + kernels) hides behind the parser either way
        for_each_batch(o.reads, reads_fasta, 64u << 20, [&](Batch &B) {
            hits.resize(B.ids.size());
            if (B.packed) { const mq_packed v = B.view(); ck(ctx, mq_map_batch_packed(ctx, &v, B.offs.data(), (uint32_t)B.ids.size(), hits.data()), "mq_map_batch_packed"); }
            else ck(ctx, mq_map_batch(ctx, B.seqs.p, B.offs.data(), (uint32_t)B.ids.size(), hits.data()), "mq_map_batch");
            for (size_t i = 0; i < B.ids.size(); i++) {                  // input order, closures.rs:117-123
                if (!hits[i].mapped) {
                    if (!o.rescue.empty()) {                             // keep the read for the second pass
                        const size_t n = (size_t)(B.offs[i + 1] - B.offs[i]);
                        B.record_bytes(i, rec_tmp);
                        un_seqs.reserve(un_seqs.size + n); memcpy(un_seqs.p + un_seqs.size, rec_tmp.data(), n); un_seqs.size += n;
                        un_offs.push_back(un_seqs.size); un_ids.push_back(B.ids[i]);
                    }
                    continue;
                }
                const std::string &rn = ref_names[hits[i].ref_idx];
                if (line.size() < B.ids[i].size() + rn.size() + 256) line.resize(B.ids[i].size() + rn.size() + 256);
                int n = mq_format_paf(line.data(), line.size(), B.ids[i].c_str(), B.offs[i + 1] - B.offs[i], rn.c_str(),
                                      ref_lens[hits[i].ref_idx], &hits[i]);
                if (n < 0) die("mq_format_paf failed");
                line[n] = '\n';
                fwrite(line.data(), 1, (size_t)n + 1, paf);
            }
        });
    }
    fclose(paf);
    printf("Mapped query sequences in %.6fs.\n", secs(t_map));                                                      // closures.rs:211
    mq_destroy(ctx);

    // ---- second pass over the unmapped reads with another (k, l, density) -------------------------------
    // (the reference does this with a second run of the binary on the unmapped set,
    //  experiments/chm13/run_chm13_mapquik_unmapped.sh:8-21; here both passes share one process)
    if (!o.rescue.empty()) {
        auto t_res = std::chrono::steady_clock::now();
        long k2 = 0, l2 = 0; double d2 = 0;
        if (sscanf(o.rescue.c_str(), "%ld,%ld,%lf", &k2, &l2, &d2) != 3) die("--rescue expects k,l,density");
        if (o.reference.empty()) die("--rescue needs --reference (the second index is built from it)");
        size_t rescued = 0;
        std::string res_name = prefix + ".rescue.paf";
        FILE *rp = fopen(res_name.c_str(), "w");
        if (!rp) die("Couldn't create " + res_name);
        if (!un_ids.empty()) {
            mq_params p2 = p; p2.k = (uint32_t)k2; p2.l = (uint32_t)l2; p2.density = d2;
            mq_ctx *c2 = nullptr;
            ck(nullptr, create(&c2, &p2), "mq_create (rescue)");
            std::vector<std::string> names2; std::vector<uint64_t> lens2;
            for_each_batch(o.reference, ref_fasta, o.low_memory ? (64u << 20) : (128u << 20), [&](Batch &B) {
                if (B.packed) { const mq_packed v = B.view(); ck(c2, mq_index_add_packed(c2, &v, B.offs.data(), (uint32_t)B.ids.size(), (uint32_t)names2.size(), nullptr), "mq_index_add_packed (rescue)"); }
                else ck(c2, mq_index_add(c2, B.seqs.p, B.offs.data(), (uint32_t)B.ids.size(), (uint32_t)names2.size(), nullptr), "mq_index_add (rescue)");
                for (size_t i = 0; i < B.ids.size(); i++) { names2.push_back(B.ids[i]); lens2.push_back(B.offs[i + 1] - B.offs[i]); }
            });
            ck(c2, mq_index_freeze(c2, lens2.data(), (uint32_t)lens2.size(), nullptr, nullptr), "mq_index_freeze (rescue)");
            std::vector<mq_hit> h2(un_ids.size()); std::vector<char> line(1 << 16);
            ck(c2, mq_map_batch(c2, un_seqs.p, un_offs.data(), (uint32_t)un_ids.size(), h2.data()), "mq_map_batch (rescue)");
            for (size_t i = 0; i < un_ids.size(); i++) {
                if (!h2[i].mapped) continue;
                const std::string &rn = names2[h2[i].ref_idx];
                if (line.size() < un_ids[i].size() + rn.size() + 256) line.resize(un_ids[i].size() + rn.size() + 256);
                int n = mq_format_paf(line.data(), line.size(), un_ids[i].c_str(), un_offs[i + 1] - un_offs[i], rn.c_str(), lens2[h2[i].ref_idx], &h2[i]);
                if (n < 0) die("mq_format_paf failed");
                line[n] = '\n'; fwrite(line.data(), 1, (size_t)n + 1, rp); rescued++;
            }
            mq_destroy(c2);
        }
        fclose(rp);
        printf("Rescued %zu of %zu unmapped reads with k=%ld l=%ld d=%g in %.6fs.\n", rescued, un_ids.size(), k2, l2, d2, secs(t_res));
    }
    printf("Total execution time: %.6fs\n", secs(t_start));                                                         // main.rs:270
    struct rusage ru; getrusage(RUSAGE_SELF, &ru);
    printf("Maximum RSS: %gGB\n", (double)ru.ru_maxrss * 1024.0 / 1024.0 / 1024.0 / 1024.0);                         // main.rs:271
    return 0;
}
