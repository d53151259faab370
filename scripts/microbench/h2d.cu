// h2d.cu -- what limits the end-to-end read path when N GPUs pull reads from ONE host?
//
// Measures, with pinned host buffers and nothing else running:
//   (1) H2D bandwidth per GPU and in aggregate when 1, 2, 4, 8 GPUs copy concurrently (one stream per GPU, one host thread
//       per GPU, 1 GiB per copy, best of 5),
//   (2) the same with every GPU's pinned buffer first-touched on the NUMA node nearest to that GPU (nodes read from
//       /sys/bus/pci/devices/<bdf>/numa_node) versus all buffers on node 0,
//   (3) D2H for the hit records' direction,
//   (4) host memory read bandwidth of T threads (sum of 64-bit loads over 1 GiB each) -- the ceiling DMA reads share.
// Output: one JSON object.   nvcc -O3 -std=c++17 -o h2d h2d.cu -lpthread
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int gpu_numa_node(int dev) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, dev) != cudaSuccess) return -1;
    for (char *p = bdf; *p; p++) *p = (char)tolower(*p);
    std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/numa_node";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int n = -1; if (fscanf(f, "%d", &n) != 1) n = -1; fclose(f);
    return n;
}
static std::vector<int> cpus_of_node(int node) {
    std::vector<int> out;
    std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return out;
    char buf[4096]; if (!fgets(buf, sizeof buf, f)) { fclose(f); return out; } fclose(f);
    for (char *p = buf; *p && *p != '\n';) {
        int a = (int)strtol(p, &p, 10), b = a;
        if (*p == '-') b = (int)strtol(p + 1, &p, 10);
        for (int c = a; c <= b; c++) out.push_back(c);
        if (*p == ',') p++;
    }
    return out;
}
static void bind_to(const std::vector<int> &cpus) {
    if (cpus.empty()) return;
    cpu_set_t s; CPU_ZERO(&s);
    for (int c : cpus) CPU_SET(c, &s);
    sched_setaffinity(0, sizeof s, &s);
}

struct Lane { int dev; void *h = nullptr, *d = nullptr; cudaStream_t st; cudaEvent_t a, b; };

// n GPUs copy `bytes` concurrently; returns per-GPU GB/s (best of reps) and the aggregate of the best concurrent round
static void concurrent(std::vector<Lane> &L, int n, size_t bytes, bool h2d, int reps, std::vector<double> &per, double &agg) {
    per.assign(n, 0.0); agg = 0;
    for (int r = 0; r < reps; r++) {
        std::atomic<int> ready{0};
        std::vector<std::thread> th;
        std::vector<double> t(n);
        double t0 = 0, t1 = 0;
        std::atomic<int> go{0};
        for (int i = 0; i < n; i++) th.emplace_back([&, i]() {
            cudaSetDevice(L[i].dev);
            ready++;
            while (!go.load()) {}
            cudaEventRecord(L[i].a, L[i].st);
            if (h2d) cudaMemcpyAsync(L[i].d, L[i].h, bytes, cudaMemcpyHostToDevice, L[i].st);
            else cudaMemcpyAsync(L[i].h, L[i].d, bytes, cudaMemcpyDeviceToHost, L[i].st);
            cudaEventRecord(L[i].b, L[i].st);
            cudaEventSynchronize(L[i].b);
            float ms = 0; cudaEventElapsedTime(&ms, L[i].a, L[i].b);
            t[i] = ms / 1e3;
        });
        while (ready.load() < n) {}
        t0 = now(); go = 1;
        for (auto &x : th) x.join();
        t1 = now();
        for (int i = 0; i < n; i++) per[i] = std::max(per[i], bytes / t[i] / 1e9);
        agg = std::max(agg, n * (double)bytes / (t1 - t0) / 1e9);
    }
}

int main(int argc, char **argv) {
    size_t bytes = (size_t)1 << 30;
    if (argc > 1) bytes = (size_t)atoll(argv[1]) << 20;
    int ndev = 0; cudaGetDeviceCount(&ndev);
    if (ndev == 0) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t all; CPU_ZERO(&all); sched_getaffinity(0, sizeof all, &all);
    printf("{\"gpus\": %d, \"host_cpus\": %ld, \"copy_mib\": %zu, \"gpu_numa_node\": [", ndev, ncpu, bytes >> 20);
    std::vector<int> node(ndev);
    for (int d = 0; d < ndev; d++) { node[d] = gpu_numa_node(d); printf("%s%d", d ? ", " : "", node[d]); }
    printf("]");
    for (int placement = 0; placement < 2; placement++) {          // 0: every buffer first-touched wherever the main thread runs; 1: near its GPU
        std::vector<Lane> L(ndev);
        for (int d = 0; d < ndev; d++) {
            L[d].dev = d;
            cudaSetDevice(d);
            if (placement == 1 && node[d] >= 0) bind_to(cpus_of_node(node[d]));
            cudaMallocHost(&L[d].h, bytes); memset(L[d].h, 1, bytes);        // first touch decides the NUMA node
            sched_setaffinity(0, sizeof all, &all);
            cudaMalloc(&L[d].d, bytes);
            cudaStreamCreateWithFlags(&L[d].st, cudaStreamNonBlocking);
            cudaEventCreate(&L[d].a); cudaEventCreate(&L[d].b);
        }
        printf(", \"%s\": {", placement ? "numa_local_buffers" : "default_buffers");
        bool first = true;
        for (int n = 1; n <= ndev; n *= 2) {
            std::vector<double> per; double agg;
            concurrent(L, n, bytes, true, 5, per, agg);
            double mn = 1e30, mx = 0; for (double v : per) { mn = std::min(mn, v); mx = std::max(mx, v); }
            printf("%s\"h2d_%dgpu\": {\"per_gpu_min\": %.1f, \"per_gpu_max\": %.1f, \"aggregate\": %.1f}", first ? "" : ", ", n, mn, mx, agg);
            first = false;
            concurrent(L, n, bytes, false, 3, per, agg);
            mn = 1e30; mx = 0; for (double v : per) { mn = std::min(mn, v); mx = std::max(mx, v); }
            printf(", \"d2h_%dgpu\": {\"per_gpu_min\": %.1f, \"per_gpu_max\": %.1f, \"aggregate\": %.1f}", n, mn, mx, agg);
        }
        printf("}");
        for (int d = 0; d < ndev; d++) { cudaSetDevice(d); cudaFreeHost(L[d].h); cudaFree(L[d].d); }
    }
    // host memory read bandwidth
    printf(", \"host_read_gb_s\": {");
    bool first = true;
    for (int T : {1, 2, 4, 8, 16, 32, 64}) {
        if (T > ncpu) break;
        const size_t per = (size_t)256 << 20;
        std::vector<uint64_t *> bufs(T);
        for (int t = 0; t < T; t++) { bufs[t] = (uint64_t *)malloc(per); memset(bufs[t], t + 1, per); }
        std::vector<uint64_t> sink(T);
        double best = 0;
        for (int r = 0; r < 3; r++) {
            double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++) th.emplace_back([&, t]() { uint64_t s = 0; const uint64_t *p = bufs[t]; for (size_t i = 0; i < per / 8; i += 4) s += p[i] + p[i + 1] + p[i + 2] + p[i + 3]; sink[t] = s; });
            for (auto &x : th) x.join();
            best = std::max(best, T * (double)per / (now() - t0) / 1e9);
        }
        printf("%s\"%d_threads\": %.1f", first ? "" : ", ", T, best);
        first = false;
        for (int t = 0; t < T; t++) free(bufs[t]);
        if (sink[0] == 42) printf(" ");
    }
    printf("}}\n");
    return 0;
}
