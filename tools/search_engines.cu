// Parity and timing of the two search engines on the same descriptors:
//   search_engines COLS ROWS K FLAGS [REPS] [POOL] [MMA_VARIANT] [COLTERM]
// Descriptors are drawn from a small pool with a few flipped bits, so that exact ties (the
// no-duplicates case) and near ties are frequent. The popcount engine (search.cu) is itself
// pinned to the oracle by tests/test_gpu_parity.py; here the tensor-core engine (search_mma.cu)
// must reproduce its four key arrays bit for bit. Exit status 1 on any mismatch.
#include "../libbicos_b200/csrc/kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

using namespace bicos_b200;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            std::printf("CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, #x); \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

int main(int argc, char** argv) {
    const int cols = argc > 1 ? std::atoi(argv[1]) : 2048;
    const int rows = argc > 2 ? std::atoi(argv[2]) : 64;
    const int K = argc > 3 ? std::atoi(argv[3]) : 4;
    const int flags = argc > 4 ? std::atoi(argv[4]) : 3;
    const int reps = argc > 5 ? std::atoi(argv[5]) : 5;
    const int pool = argc > 6 ? std::atoi(argv[6]) : 64;
    if (argc > 7)
        set_search_mma_variant(std::atoi(argv[7])); // else: default / BICOS_B200_MMA_VARIANT
    // COLTERM = 1: descriptors with a zero top bit (as the transform writes them) and the column-term kernels
    // COLTERM = 2: the top two bits are zero, which also admits the one-pass consistency kernel (variant 0 or 3)
    const int free_bits = argc > 8 ? std::atoi(argv[8]) : 0;
    const bool colterm = free_bits != 0;
    set_search_mma_colterm(colterm);

    const size_t pitch = ((size_t)cols * K + 3) / 4 * 4;
    std::mt19937 rng(1234u + cols + 7 * rows + 13 * K);
    std::vector<uint32_t> base((size_t)pool * K);
    for (auto& w: base)
        w = rng();
    auto fill = [&](std::vector<uint32_t>& d) {
        d.assign(pitch * rows, 0u);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                uint32_t* p = &d[(size_t)r * pitch + (size_t)c * K];
                const uint32_t* b = &base[(size_t)(rng() % pool) * K];
                for (int k = 0; k < K; ++k)
                    p[k] = b[k];
                const int flips = rng() % 4; // 0 flips: exact duplicates of a pool entry
                for (int f = 0; f < flips; ++f) {
                    const int bit = rng() % (32 * K);
                    p[bit / 32] ^= 1u << (bit % 32);
                }
                if (colterm)
                    p[K - 1] &= free_bits >= 2 ? 0x3FFFFFFFu : 0x7FFFFFFFu;
            }
    };
    std::vector<uint32_t> h0, h1;
    fill(h0);
    fill(h1);

    uint32_t *d0, *d1, *keys[2];
    const size_t px = (size_t)rows * cols;
    CK(cudaMalloc(&d0, h0.size() * 4));
    CK(cudaMalloc(&d1, h1.size() * 4));
    CK(cudaMalloc(&keys[0], px * 16));
    CK(cudaMalloc(&keys[1], px * 16));
    CK(cudaMemcpy(d0, h0.data(), h0.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d1, h1.data(), h1.size() * 4, cudaMemcpyHostToDevice));

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms[2] = { 0, 0 };
    for (int engine = 0; engine < 2; ++engine) {
        uint32_t* k = keys[engine];
        auto run = [&]() {
            return engine == 0
                ? launch_search_popc(d0, d1, K, rows, cols, pitch, flags, k, k + px, k + 2 * px, k + 3 * px, 0)
                : launch_search_mma(d0, d1, K, rows, cols, pitch, flags, k, k + px, k + 2 * px, k + 3 * px, 0, free_bits);
        };
        CK(cudaMemset(k, 0xFF, px * 16));
        CK(run());
        CK(cudaDeviceSynchronize());
        if (reps > 0) {
            // the popcount engine merges with atomicMin, so re-running on its own results is idempotent
            CK(cudaEventRecord(e0));
            for (int i = 0; i < reps; ++i)
                CK(run());
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms[engine], e0, e1));
            ms[engine] /= reps;
        }
    }

    std::vector<uint32_t> a(px * 4), b(px * 4);
    CK(cudaMemcpy(a.data(), keys[0], px * 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), keys[1], px * 16, cudaMemcpyDeviceToHost));
    const char* names[4] = { "fwd_first", "fwd_last", "rev_first", "rev_last" };
    const bool used[4] = { true, (flags & FLAG_NODUPES) != 0, (flags & FLAG_CONSISTENCY) != 0,
                           (flags & FLAG_NODUPES) && (flags & FLAG_CONSISTENCY) };
    long long bad_total = 0, ties = 0;
    for (int arr = 0; arr < 4; ++arr) {
        if (!used[arr])
            continue;
        long long bad = 0;
        for (size_t i = 0; i < px; ++i) {
            const uint32_t x = a[arr * px + i], y = b[arr * px + i];
            if (x != y) {
                if (bad < 6)
                    std::printf("  %s row %zu col %zu: popc cost %u col %u | mma cost %u col %u (raw %08x)\n", names[arr], i / cols,
                                i % cols, x >> 16, x & 0xFFFF, y >> 16, y & 0xFFFF, y);
                ++bad;
            }
        }
        std::printf("%s: %lld mismatches of %zu\n", names[arr], bad, px);
        bad_total += bad;
    }
    if (used[1])
        for (size_t i = 0; i < px; ++i)
            ties += (a[i] & 0xFFFF) != 65535u - (a[px + i] & 0xFFFF);
    const double pairs = (double)rows * cols * cols * ((flags & FLAG_CONSISTENCY) ? 1 : 1);
    std::printf(
        "cols %d rows %d K %d flags %d variant %d colterm %d [%s]: popc %.4f ms (%.3f Tpair/s), mma %.4f ms (%.3f Tpair/s), speed-up %.2fx, forward ties %lld, %s\n",
        cols, rows, K, flags, search_mma_variant(), free_bits, last_search_kernel(), ms[0], pairs / ms[0] * 1e-9, ms[1], pairs / ms[1] * 1e-9, ms[1] > 0 ? ms[0] / ms[1] : 0.0, ties,
        bad_total ? "MISMATCH" : "identical"
    );
    return bad_total ? 1 : 0;
}
