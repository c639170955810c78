// The interleaved fast-path square root / quotient of refine.cu (nxc_pair) against the library's __fsqrt_rn / __fdiv_rn
// on the GPU: random operands over the whole range nxc_pair takes the fast path for (log-uniform exponents, random
// mantissas, both signs of the covariance), plus operands shaped like the kernel's (sums of squares of half-integers
// minus means). Prints the number of differing results; exit status 1 if any.   tools/nxc_check [millions of pairs]
#include "../libbicos_b200/csrc/refine.cu"

#include <cstdio>
#include <cstdlib>

using namespace bicos_b200;

__device__ uint32_t mix(uint64_t& s) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t x = s;
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    return (uint32_t)x;
}

__global__ void check_kernel(unsigned long long* bad, unsigned long long* fast, int rounds, unsigned long long seed) {
    uint64_t s = seed + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    unsigned long long nbad = 0, nfast = 0;
    for (int i = 0; i < rounds; ++i) {
        float v0, v10, v11, c0, c1;
        if (i & 1) {
            // any float in the fast range: p = v0 * v1 with v0 = 1, exponent 87 .. 206; |a| exponent 87 .. 166
            v0 = 1.0f;
            v10 = __uint_as_float(((87u + mix(s) % 120u) << 23) | (mix(s) & 0x7FFFFFu));
            v11 = __uint_as_float(((87u + mix(s) % 120u) << 23) | (mix(s) & 0x7FFFFFu));
            c0 = __uint_as_float((mix(s) & 0x80000000u) | ((87u + mix(s) % 80u) << 23) | (mix(s) & 0x7FFFFFu));
            c1 = __uint_as_float((mix(s) & 0x80000000u) | ((87u + mix(s) % 80u) << 23) | (mix(s) & 0x7FFFFFu));
        } else {
            // like the kernel's: variances up to 33 * 255^2 with 2^-16 granularity, covariances of either sign
            v0 = (float)(mix(s) % 2145825u) + (float)(mix(s) & 0xFFFFu) * (1.0f / 65536.0f) + 0.5f;
            v10 = (float)(mix(s) % 2145825u) + (float)(mix(s) & 0xFFFFu) * (1.0f / 65536.0f) + 0.5f;
            v11 = (float)(mix(s) % 2145825u) + (float)(mix(s) & 0xFFFFu) * (1.0f / 65536.0f) + 0.5f;
            c0 = ((float)(mix(s) % 2145825u) + (float)(mix(s) & 0xFFFFu) * (1.0f / 65536.0f)) * ((mix(s) & 1) ? -1.f : 1.f);
            c1 = ((float)(mix(s) % 2145825u) + (float)(mix(s) & 0xFFFFu) * (1.0f / 65536.0f)) * ((mix(s) & 1) ? -1.f : 1.f);
        }
        const float p0 = __fmul_rn(v0, v10), p1 = __fmul_rn(v0, v11);
        nfast += nxc_fast_range(p0, c0) && nxc_fast_range(p1, c1);
        const NxcPair q = nxc_pair(c0, c1, v0, v10, v11);
        const float w0 = __fdiv_rn(c0, __fsqrt_rn(p0)), w1 = __fdiv_rn(c1, __fsqrt_rn(p1));
        nbad += (__float_as_uint(q.lo) != __float_as_uint(w0)) + (__float_as_uint(q.hi) != __float_as_uint(w1));
    }
    atomicAdd(bad, nbad);
    atomicAdd(fast, nfast);
}

int main(int argc, char** argv) {
    const int millions = argc > 1 ? atoi(argv[1]) : 2000;
    unsigned long long *d, h[2] = { 0, 0 };
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    const int threads = 256, blocks = 148 * 16;
    const int rounds = (int)((long long)millions * 1000000 / ((long long)threads * blocks));
    check_kernel<<<blocks, threads>>>(d, d + 1, rounds, 12345);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("CUDA error\n");
        return 2;
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const long long pairs = (long long)rounds * threads * blocks;
    printf("nxc_pair vs __fdiv_rn(c, __fsqrt_rn(v0 * v1)): %lld operand pairs (2 results each), %llu on the fast path, %llu results differ\n",
           pairs, h[1], h[0]);
    return h[0] ? 1 : 0;
}
