// Pipe-throughput micro-benchmark for the search kernel's roofline denominator:
// measures warp-instructions per clock per SM for POPC, LOP3, IMAD, VIMNMX, REDUX and
// the mixes the Hamming search issues. Build: make -C tools. Run on the GPU box:
//   tools/microbench > gpurun_out/microbench.txt
#include <cstdio>
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int CHAINS = 8;

#define POPC(x) asm volatile("popc.b32 %0, %0;" : "+r"(x))
#define LOP(x, y) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x) : "r"(y))
#define IMAD(x, y) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y))
#define MINU(x, y) asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define IADD(x, y) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define REDUX(x) asm volatile("redux.sync.min.u32 %0, %0, 0xffffffff;" : "+r"(x))
#define SHFL(x) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x))
// floating-point / byte-permute ops of the refine kernel (values are bit patterns, not meaningful floats)
#define FFMA(x, y) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(x) : "r"(y))
#define FADD(x, y) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define FFMA2(x, y) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x) : "l"(y))
#define FMUL2(x, y) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y))
#define FADD2(x, y) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y))
#define PRMT(x, y) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x) : "r"(y))
#define IMADHI(x, y) asm volatile("mad.hi.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y))
#define IMADWIDE(x2, x, y) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x2) : "r"(x), "r"(y))
#define SHR(x, y) asm volatile("shr.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define HMIN2(x, y) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y))
#define DP4A(x, y) asm volatile("dp4a.u32.u32 %0, %0, %1, %0;" : "+r"(x) : "r"(y))

template<int MODE>
__global__ void bench(uint32_t* out, long long* cycles, int iters, uint32_t seed) {
    uint32_t x[CHAINS], y = seed | 3u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
        x[c] = seed * (c + 1) + threadIdx.x;
    unsigned long long x2[CHAINS], y2 = ((unsigned long long)(seed | 3u) << 32) | 0x3f800001u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
        x2[c] = ((unsigned long long)x[c] << 32) | x[c];
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) { POPC(x[c]); }
            if (MODE == 1) { LOP(x[c], y); }
            if (MODE == 2) { IMAD(x[c], y); }
            if (MODE == 3) { MINU(x[c], y); }
            if (MODE == 4) { IADD(x[c], y); }
            if (MODE == 5) { REDUX(x[c]); }
            if (MODE == 6) { SHFL(x[c]); }
            if (MODE == 7) { POPC(x[c]); LOP(x[c], y); }                       // 1 popc : 1 alu
            if (MODE == 8) { POPC(x[c]); LOP(x[c], y); LOP(x[c], y); }          // 1 : 2
            if (MODE == 9) { POPC(x[c]); LOP(x[c], y); LOP(x[c], y); LOP(x[c], y); } // 1 : 3
            if (MODE == 10) { POPC(x[c]); LOP(x[c], y); LOP(x[c], y); LOP(x[c], y); LOP(x[c], y); } // 1 : 4
            if (MODE == 11) { POPC(x[c]); IMAD(x[c], y); }                     // popc + fma pipe
            if (MODE == 12) { POPC(x[c]); IMAD(x[c], y); LOP(x[c], y); LOP(x[c], y); } // 1 popc, 1 fma, 2 alu
            if (MODE == 13) { POPC(x[c]); IMAD(x[c], y); LOP(x[c], y); LOP(x[c], y); LOP(x[c], y); } // 1,1,3
            if (MODE == 14) { LOP(x[c], y); IMAD(x[c], y); }                   // alu + fma co-issue
            if (MODE == 16) { FFMA(x[c], y); }
            if (MODE == 17) { FADD(x[c], y); }
            if (MODE == 18) { FFMA2(x2[c], y2); }
            if (MODE == 19) { FMUL2(x2[c], y2); }
            if (MODE == 20) { FADD2(x2[c], y2); }
            if (MODE == 21) { PRMT(x[c], y); }
            if (MODE == 22) { DP4A(x[c], y); }
            if (MODE == 23) { FFMA2(x2[c], y2); LOP(x[c], y); }                 // packed fp32 + alu co-issue
            if (MODE == 24) { FFMA(x[c], y); LOP(x[c], y); }
            if (MODE == 25) { IMADHI(x[c], y); }
            if (MODE == 26) { IMADWIDE(x2[c], x[c], y); }
            if (MODE == 27) { SHR(x[c], y); }
            if (MODE == 28) { IMADHI(x[c], y); LOP(x[c], y); }
            if (MODE == 29) { IMADWIDE(x2[c], x[c], y); LOP(x[c], y); }
            if (MODE == 30) { x[c] = __viaddmin_s16x2(x[c], y, x[(c + 3) % CHAINS]); }   // VIADDMNMX.S16x2: the elementwise fold of search_mma3
            if (MODE == 31) { x[c] = __vimin3_s16x2(x[c], x[(c + 1) % CHAINS], x[(c + 3) % CHAINS]); } // VIMNMX3.S16x2: the in-thread fold
            if (MODE == 32) { x[c] = __viaddmin_s16x2(x[c], y, x[(c + 3) % CHAINS]); IMAD(x[c], y); } // + FMA-pipe co-issue
            if (MODE == 33) { x[c] = __vimin3_s16x2(x[c], x[(c + 1) % CHAINS], x[(c + 3) % CHAINS]); LOP(x[c], y); }
            if (MODE == 34) { HMIN2(x[c], x[(c + 3) % CHAINS]); }                           // HMNMX2: which pipe?
            if (MODE == 35) { HMIN2(x[c], x[(c + 3) % CHAINS]); LOP(x[c], y); }
            if (MODE == 36) { x[c] = __viaddmin_s16x2(x[c], y, x[(c + 3) % CHAINS]); HMIN2(x[(c + 1) % CHAINS], x[(c + 5) % CHAINS]); }
            if (MODE == 15) { POPC(x[c]); POPC(x[c]); POPC(x[c]); LOP(x[c], y); LOP(x[c], y); LOP(x[c], y); LOP(x[c], y); LOP(x[c], y); LOP(x[c], y); MINU(x[c], y); MINU(x[c], y); IMAD(x[c], y); IMAD(x[c], y); IMAD(x[c], y); } // search-like
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
        acc ^= x[c] ^ (uint32_t)x2[c] ^ (uint32_t)(x2[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0)
        cycles[blockIdx.x] = t1 - t0;
}

template<int MODE>
int run(const char* name, int ops_per_chain_iter, int sms, uint32_t* out, long long* cyc) {
    const int threads = 256, blocks_per_sm = 8, iters = 2048;
    const int grid = sms * blocks_per_sm;
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a));
    CHECK(cudaEventCreate(&b));
    bench<MODE><<<grid, threads>>>(out, cyc, 16, 1);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(a));
    bench<MODE><<<grid, threads>>>(out, cyc, iters, 12345);
    CHECK(cudaEventRecord(b));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    long long* h = new long long[grid];
    CHECK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < grid; ++i)
        avg += (double)h[i];
    avg /= grid;
    delete[] h;
    const double warp_instr_per_sm = (double)blocks_per_sm * (threads / 32) * iters * CHAINS * ops_per_chain_iter;
    printf("%-28s %8.3f ms  %10.0f cyc/block  %6.3f warp-instr/clk/SM  (%5.2f thread-ops/clk/SM)  eff clock %.0f MHz\n",
           name, ms, avg, warp_instr_per_sm / avg, 32.0 * warp_instr_per_sm / avg, avg / (ms * 1e3));
    return 0;
}

// thread-level POPC per second, best of several long launches (clocks need ~100 ms to ramp up)
int popc_rate(int sms, uint32_t* out, long long* cyc) {
    const int threads = 256, blocks_per_sm = 8, iters = 8192;
    const int grid = sms * blocks_per_sm;
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a));
    CHECK(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 40; ++rep) {
        CHECK(cudaEventRecord(a));
        bench<0><<<grid, threads>>>(out, cyc, iters, 12345 + rep);
        CHECK(cudaEventRecord(b));
        CHECK(cudaDeviceSynchronize());
        float ms = 0;
        CHECK(cudaEventElapsedTime(&ms, a, b));
        const double ops = (double)grid * threads * iters * CHAINS;
        const double rate = ops / (ms * 1e-3);
        if (rate > best)
            best = rate;
    }
    printf("POPC_PER_S %.6e\n", best);
    return 0;
}

int main(int argc, char** argv) {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs, clock %.0f MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1e3);
    const int sms = p.multiProcessorCount;
    uint32_t* out;
    long long* cyc;
    CHECK(cudaMalloc(&out, sizeof(uint32_t) * sms * 8 * 256));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * sms * 8));
    if (argc > 1 && std::string(argv[1]) == "--popc")
        return popc_rate(sms, out, cyc);
    run<0>("POPC", 1, sms, out, cyc);
    run<1>("LOP3", 1, sms, out, cyc);
    run<2>("IMAD", 1, sms, out, cyc);
    run<3>("VIMNMX (min.u32)", 1, sms, out, cyc);
    run<4>("IADD", 1, sms, out, cyc);
    run<5>("REDUX.min", 1, sms, out, cyc);
    run<6>("SHFL.bfly", 1, sms, out, cyc);
    run<7>("POPC+1 LOP3", 2, sms, out, cyc);
    run<8>("POPC+2 LOP3", 3, sms, out, cyc);
    run<9>("POPC+3 LOP3", 4, sms, out, cyc);
    run<10>("POPC+4 LOP3", 5, sms, out, cyc);
    run<11>("POPC+IMAD", 2, sms, out, cyc);
    run<12>("POPC+IMAD+2 LOP3", 4, sms, out, cyc);
    run<13>("POPC+IMAD+3 LOP3", 5, sms, out, cyc);
    run<14>("LOP3+IMAD", 2, sms, out, cyc);
    run<15>("3POPC+6LOP+2MIN+3IMAD", 14, sms, out, cyc);
    run<16>("FFMA", 1, sms, out, cyc);
    run<17>("FADD", 1, sms, out, cyc);
    run<18>("FFMA2 (f32x2)", 1, sms, out, cyc);
    run<19>("FMUL2 (f32x2)", 1, sms, out, cyc);
    run<20>("FADD2 (f32x2)", 1, sms, out, cyc);
    run<21>("PRMT", 1, sms, out, cyc);
    run<22>("IDP4A", 1, sms, out, cyc);
    run<23>("FFMA2+LOP3", 2, sms, out, cyc);
    run<24>("FFMA+LOP3", 2, sms, out, cyc);
    run<25>("IMAD.HI", 1, sms, out, cyc);
    run<26>("IMAD.WIDE", 1, sms, out, cyc);
    run<27>("SHR", 1, sms, out, cyc);
    run<28>("IMAD.HI+LOP3", 2, sms, out, cyc);
    run<29>("IMAD.WIDE+LOP3", 2, sms, out, cyc);
    run<30>("VIADDMNMX.S16x2", 1, sms, out, cyc);
    run<31>("VIMNMX3.S16x2", 1, sms, out, cyc);
    run<32>("VIADDMNMX.S16x2+IMAD", 2, sms, out, cyc);
    run<33>("VIMNMX3.S16x2+LOP3", 2, sms, out, cyc);
    run<34>("HMNMX2 (min.f16x2)", 1, sms, out, cyc);
    run<35>("HMNMX2+LOP3", 2, sms, out, cyc);
    run<36>("VIADDMNMX.S16x2+HMNMX2", 2, sms, out, cyc);
    return 0;
}
