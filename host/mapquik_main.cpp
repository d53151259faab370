// mapquik (B200) -- C++ host CLI over the C ABI of libmapquik_b200.so.
//
// Mirrors the reference's command line, console lines and PAF output (src/main.rs:77-272,
// src/closures.rs:22-211) for the seeding->chaining path:
//     mapquik <reads.fa|fq[.gz]> --reference <ref.fa[.gz]> [-k K] [-l L] [-d D] [-c C] [-s S] [-g G]
//             [-p PREFIX] [--nohpc] [--threads N] [-b B] [-q Q] [--low-memory] [--nosimd]
//             [--parallelfastx] [--debug] [--gpu ID]
// Host I/O is deliberately simple (single reader thread, batches into pinned memory); the reference's
// seq_io / parallelfastx threading is out of scope.  All compute happens in the library.
#include "../include/mapquik_b200.h"

#include <sys/resource.h>
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Opt {
    std::string reads, reference, prefix;
    bool has_prefix = false, debug = false, low_memory = false, nosimd = false, nohpc = false, pfx = false;
    long k = -1, l = -1, c = -1, s = -1, g = -1, threads = -1, b = -1, q = -1;
    double density = -1;
    int gpu = 0;
};

[[noreturn]] void die(const std::string &m) { fprintf(stderr, "%s\n", m.c_str()); exit(1); }

bool is_fasta_name(const std::string &f) {   // main.rs:196,202 (same substring test)
    auto ends = [&](const char *x) { size_t n = strlen(x); return f.size() >= n && f.compare(f.size() - n, n, x) == 0; };
    return f.find(".fasta.") != std::string::npos || ends(".fna") || f.find(".fna.") != std::string::npos ||
           f.find(".fa.") != std::string::npos || ends(".fa") || ends(".fasta");
}

// FASTA / FASTQ reader (gz transparently via zlib; multi-line FASTA accepted)
struct Fastx {
    gzFile f = nullptr; bool fasta;
    std::vector<char> buf; size_t pos = 0, len = 0; bool eof = false;
    std::string pending;        // header line already consumed (FASTA)
    Fastx(const std::string &path, bool fasta_) : fasta(fasta_) {
        f = gzopen(path.c_str(), "rb");
        if (!f) die("Error opening file: " + path);
        gzbuffer(f, 1 << 20);
        buf.resize(1 << 22);
    }
    ~Fastx() { if (f) gzclose(f); }
    bool getline(std::string &out) {
        out.clear();
        for (;;) {
            if (pos == len) {
                if (eof) return !out.empty();
                int r = gzread(f, buf.data(), (unsigned)buf.size());
                if (r <= 0) { eof = true; return !out.empty(); }
                pos = 0; len = (size_t)r;
            }
            char *nl = (char *)memchr(buf.data() + pos, '\n', len - pos);
            if (nl) { out.append(buf.data() + pos, nl - (buf.data() + pos)); pos = (nl - buf.data()) + 1; break; }
            out.append(buf.data() + pos, len - pos); pos = len;
        }
        if (!out.empty() && out.back() == '\r') out.pop_back();
        return true;
    }
    // next record: id (up to first whitespace, record.id() in closures.rs:64,107) + sequence appended to seq
    bool next(std::string &id, std::vector<uint8_t> &seq) {
        std::string line;
        if (fasta) {
            std::string hdr;
            if (!pending.empty()) { hdr.swap(pending); }
            else { do { if (!getline(line)) return false; } while (line.empty() || line[0] != '>'); hdr = line; }
            id = hdr.substr(1, hdr.find_first_of(" \t") == std::string::npos ? std::string::npos : hdr.find_first_of(" \t") - 1);
            while (getline(line)) {
                if (!line.empty() && line[0] == '>') { pending = line; break; }
                seq.insert(seq.end(), line.begin(), line.end());
            }
            return true;
        }
        do { if (!getline(line)) return false; } while (line.empty());
        if (line[0] != '@') die("malformed FASTQ record: " + line);
        id = line.substr(1, line.find_first_of(" \t") == std::string::npos ? std::string::npos : line.find_first_of(" \t") - 1);
        if (!getline(line)) return false;
        seq.insert(seq.end(), line.begin(), line.end());
        std::string plus, qual;
        getline(plus); getline(qual);
        return true;
    }
};

inline void upper(uint8_t *p, size_t n) { for (size_t i = 0; i < n; i++) if (p[i] >= 'a' && p[i] <= 'z') p[i] -= 32; }  // closures.rs:63,106

double secs(std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); }

void ck(mq_ctx *c, int rc, const char *what) {
    if (rc != MQ_OK) die(std::string(what) + ": " + mq_strerror(rc) + " (" + (c ? mq_last_error(c) : "") + ")");
}

}  // namespace

int main(int argc, char **argv) {
    auto t_start = std::chrono::steady_clock::now();
    Opt o;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char *name) -> std::string { if (i + 1 >= argc) die(std::string("missing value for ") + name); return argv[++i]; };
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
        else if (a == "-h" || a == "--help") { printf("mapquik <reads> --reference <ref> [-k -l -d -c -s -g -p --nohpc --threads --gpu]\n"); return 0; }
        else if (!a.empty() && a[0] == '-') die("unknown option " + a);
        else o.reads = a;
    }
    if (o.reads.empty()) die("Please specify an input file.");
    if (o.reference.empty()) die("Please specify a reference file.");
    long k = 5, l = 31, c = 4, s = 11, g = 2000, b = 1, q = 200;
    double density = 0.01;
    const bool reads_fasta = is_fasta_name(o.reads), ref_fasta = is_fasta_name(o.reference);
    if (reads_fasta) { printf("Input file: %s\nFormat: FASTA\n", o.reads.c_str()); }
    if (ref_fasta) { printf("Reference file: %s\nFormat: FASTA\n", o.reference.c_str()); }
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
    ck(nullptr, mq_create(&ctx, &p, o.gpu), "mq_create");

    std::string paf_name = prefix + ".paf";
    FILE *paf = fopen(paf_name.c_str(), "w");
    if (!paf) die("Couldn't create " + paf_name);
    std::vector<char> pafbuf(1 << 22); setvbuf(paf, pafbuf.data(), _IOFBF, pafbuf.size());

    // ---- reference ------------------------------------------------------------------------------------
    auto t_idx = std::chrono::steady_clock::now();
    std::vector<std::string> ref_names; std::vector<uint64_t> ref_lens;
    {
        Fastx fx(o.reference, ref_fasta);
        const size_t BATCH = o.low_memory ? (64u << 20) : (512u << 20);
        std::vector<uint8_t> seqs; std::vector<uint64_t> offs{0}; std::vector<std::string> ids;
        auto flush = [&]() {
            if (ids.empty()) return;
            std::vector<uint64_t> nb(ids.size());
            ck(ctx, mq_index_add(ctx, seqs.data(), offs.data(), (uint32_t)ids.size(), (uint32_t)ref_names.size(), nb.data()), "mq_index_add");
            for (size_t i = 0; i < ids.size(); i++) {
                printf("Indexed reference %s: %llu k-min-mers.\n", ids[i].c_str(), (unsigned long long)nb[i]);   // closures.rs:58
                ref_names.push_back(ids[i]); ref_lens.push_back(offs[i + 1] - offs[i]);
            }
            seqs.clear(); offs.assign(1, 0); ids.clear();
        };
        std::string id;
        for (;;) {
            size_t before = seqs.size();
            if (!fx.next(id, seqs)) break;
            upper(seqs.data() + before, seqs.size() - before);
            offs.push_back(seqs.size()); ids.push_back(id);
            if (seqs.size() >= BATCH) flush();
        }
        flush();
    }
    uint64_t n_unique = 0;
    ck(ctx, mq_index_freeze(ctx, ref_lens.data(), (uint32_t)ref_lens.size(), &n_unique, nullptr), "mq_index_freeze");
    printf("Indexed %llu unique k-min-mers in %.6fs.\n", (unsigned long long)n_unique, secs(t_idx));                 // closures.rs:92

    // ---- reads ----------------------------------------------------------------------------------------
    auto t_map = std::chrono::steady_clock::now();
    {
        Fastx fx(o.reads, reads_fasta);
        const size_t BATCH = 256u << 20;
        std::vector<uint8_t> seqs; std::vector<uint64_t> offs{0}; std::vector<std::string> ids; std::vector<mq_hit> hits;
        std::vector<char> line(1 << 16);
        auto flush = [&]() {
            if (ids.empty()) return;
            hits.resize(ids.size());
            ck(ctx, mq_map_batch(ctx, seqs.data(), offs.data(), (uint32_t)ids.size(), hits.data()), "mq_map_batch");
            for (size_t i = 0; i < ids.size(); i++) {                    // input order, closures.rs:117-123
                if (!hits[i].mapped) continue;
                const std::string &rn = ref_names[hits[i].ref_idx];
                if (line.size() < ids[i].size() + rn.size() + 256) line.resize(ids[i].size() + rn.size() + 256);
                int n = mq_format_paf(line.data(), line.size(), ids[i].c_str(), offs[i + 1] - offs[i], rn.c_str(), ref_lens[hits[i].ref_idx], &hits[i]);
                if (n < 0) die("mq_format_paf failed");
                fwrite(line.data(), 1, (size_t)n, paf); fputc('\n', paf);
            }
            seqs.clear(); offs.assign(1, 0); ids.clear();
        };
        std::string id;
        for (;;) {
            size_t before = seqs.size();
            if (!fx.next(id, seqs)) break;
            upper(seqs.data() + before, seqs.size() - before);
            offs.push_back(seqs.size()); ids.push_back(id);
            if (seqs.size() >= BATCH) flush();
        }
        flush();
    }
    fclose(paf);
    printf("Mapped query sequences in %.6fs.\n", secs(t_map));                                                      // closures.rs:211
    mq_destroy(ctx);
    printf("Total execution time: %.6fs\n", secs(t_start));                                                         // main.rs:270
    struct rusage ru; getrusage(RUSAGE_SELF, &ru);
    printf("Maximum RSS: %gGB\n", (double)ru.ru_maxrss * 1024.0 / 1024.0 / 1024.0 / 1024.0);                         // main.rs:271
    return 0;
}
