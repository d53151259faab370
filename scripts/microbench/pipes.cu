// pipes.cu -- issue-rate microbenchmark for the integer instructions the S1 scan kernel is made of.
// Prints warp-instructions per clock per SM for each instruction class and for ALU+FMA mixes, so the
// hash-step formulation can be balanced across the two pipes.  Build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define ILP 8

template <int OP> __device__ __forceinline__ void op(uint32_t &a, uint32_t &b, uint32_t k) {
    if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));
    if (OP == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a) : "r"(b));
    if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(a) : "r"(b), "r"(k));
    if (OP == 3) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a), "r"(k)); a = (uint32_t)w; b = (uint32_t)(w >> 32); }
    if (OP == 4) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(k));
    if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(k));
    if (OP == 6) asm volatile("min.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 7) asm volatile("popc.b32 %0, %0;" : "+r"(a));
    if (OP == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 9) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));          // 1 ALU + 1 wide
                   uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(b), "r"(k)); b = (uint32_t)(w >> 32) ^ (uint32_t)w; }
    if (OP == 10) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));         // 1 ALU + 1 imad
                    asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(b) : "r"(a), "r"(k)); }
    if (OP == 11) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));         // 2 ALU + 1 wide
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a) : "r"(b));
                    uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(b), "r"(k)); b = (uint32_t)(w >> 32) ^ (uint32_t)w; }
    if (OP == 13) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));         // 2 ALU: LOP3 + SHF
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(b) : "r"(a)); }
    if (OP == 14) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));         // 2 ALU + 1 IMAD
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a) : "r"(b));
                    asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(b) : "r"(a), "r"(k)); }
    if (OP == 15) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));         // 1 ALU + 1 IMAD.HI
                    asm volatile("mad.hi.u32 %0, %0, %2, %1;" : "+r"(b) : "r"(a), "r"(k)); }
    if (OP == 16) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));         // 3 ALU + 1 IMAD
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(a) : "r"(b));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));
                    asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(b) : "r"(a), "r"(k)); }
    if (OP == 12) { uint64_t w; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(a), "r"(k), "l"(((uint64_t)b << 32) | a)); a = (uint32_t)w; b = (uint32_t)(w >> 32); }
}
static const char *NAMES[] = {"LOP3", "SHF.L.W", "IMAD", "IMAD.WIDE.U32", "IMAD.HI.U32", "PRMT", "VIMNMX", "POPC", "IADD",
                              "LOP3 + (IMAD.WIDE+LOP3)", "LOP3 + IMAD", "LOP3+SHF + (IMAD.WIDE+LOP3)", "IMAD.WIDE (mad)",
                              "LOP3 + SHF", "LOP3+SHF + IMAD", "LOP3 + IMAD.HI", "LOP3+SHF+LOP3 + IMAD"};
static const int NINSTR[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 3, 2, 4, 1, 2, 3, 2, 4};

template <int OP> __global__ void k(uint32_t *out, uint32_t k0, long long *cyc) {
    uint32_t a[ILP], b[ILP];
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 7 + i; b[i] = blockIdx.x + i * 3; }
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) op<OP>(a[i], b[i], k0);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < ILP; i++) s += a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

static int g_threads = 1024;
template <int OP> void run(uint32_t *d, long long *dc) {
    const int threads = g_threads, blocks = 148 * 2;
    k<OP><<<blocks, threads>>>(d, 0x80000000u, dc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d, 0x80000000u, dc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    // per SM: 2 blocks x 32 warps resident; warp-instrs per SM = 64 * ITER * ILP * NINSTR
    double wi = 2.0 * (threads / 32) * ITER * ILP * NINSTR[OP];
    // clock64() of one CTA and, independently, CUDA events around the whole launch (148 SMs, 2 CTAs each)
    printf("%-32s %6.3f warp-instr/clk64/SM  (%lld clk64)   %7.1f G warp-instr/s chip-wide by events (%.3f ms)\n", NAMES[OP], wi / (double)c, c,
           wi * 148.0 / (ms * 1e-3) / 1e9, ms);
}

int main() {
    uint32_t *d; long long *dc;
    cudaMalloc(&d, 148 * 2 * 1024 * 4); cudaMalloc(&dc, 8);
    run<0>(d, dc); run<1>(d, dc); run<2>(d, dc); run<3>(d, dc); run<4>(d, dc); run<5>(d, dc); run<6>(d, dc);
    run<7>(d, dc); run<8>(d, dc); run<9>(d, dc); run<10>(d, dc); run<11>(d, dc); run<12>(d, dc);
    run<13>(d, dc); run<14>(d, dc); run<15>(d, dc); run<16>(d, dc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
