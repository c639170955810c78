// Development aid: times split / tile variants of the search kernel on random descriptors and
// checks that every variant produces the same key arrays. Not part of the product; the choices
// it led to are recorded in DESIGN.md. Build: make -C tools. Run: tools/search_tune [rows cols]
#include "../libbicos_b200/csrc/search.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace bicos_b200;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Buffers {
    uint32_t *d0, *d1, *keys, *ref;
    size_t px, pitch;
    int rows, cols;
};

static uint32_t rng_state = 12345;
static uint32_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 17;
    rng_state ^= rng_state << 5;
    return rng_state;
}

static std::vector<uint32_t> g_want, g_got; // shared by all instantiations of run()

template<int K, int FLAGS, int A, int NT, int UNROLL, int MIX = DEFAULT_MIX<K>>
void run(const char* name, Buffers& b, int schedule, bool is_reference = false) {
    const size_t key_bytes = b.px * 4 * sizeof(uint32_t);
    uint32_t *ff = b.keys, *fl = b.keys + b.px, *rf = b.keys + 2 * b.px, *rl = b.keys + 3 * b.px;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    float best = 1e9f, sum = 0;
    const int iters = 5;
    for (int it = 0; it < iters + 1; ++it) {
        CHECK(cudaMemset(b.keys, 0xFF, key_bytes));
        CHECK(cudaEventRecord(e0));
        CHECK((launch_one<K, FLAGS, A, NT, UNROLL, MIX>(b.d0, b.d1, b.rows, b.cols, b.pitch, ff, fl, rf, rl, nullptr, schedule)));
        CHECK(cudaEventRecord(e1));
        CHECK(cudaDeviceSynchronize());
        float ms;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0) {
            best = ms < best ? ms : best;
            sum += ms;
        }
    }
    // compare with the reference variant's keys (unused arrays stay 0xFF in both)
    std::vector<uint32_t>& want = g_want;
    std::vector<uint32_t>& got = g_got;
    got.resize(b.px * 4);
    CHECK(cudaMemcpy(got.data(), b.keys, key_bytes, cudaMemcpyDeviceToHost));
    size_t diff = 0;
    if (is_reference)
        want = got;
    else
        for (size_t i = 0; i < got.size() && i < want.size(); ++i)
            diff += got[i] != want[i];
    const double pairs = (double)b.cols * b.px;
    printf("%-44s %8.3f ms (min %7.3f)  %6.3f T pairs/s  %s\n", name, sum / iters, best, pairs / best / 1e9,
           is_reference ? "reference" : (diff ? "MISMATCH" : "same keys"));
    if (diff)
        printf("   !!! %zu keys differ\n", diff);
}

template<int K>
void fill(Buffers& b) {
    std::vector<uint32_t> h0(b.pitch * b.rows), h1(b.pitch * b.rows);
    for (auto& v: h1)
        v = rnd();
    // left = right shifted by 17 columns with a few flipped bits; low-entropy stripes create ties
    for (int r = 0; r < b.rows; ++r)
        for (int c = 0; c < b.cols; ++c)
            for (int k = 0; k < K; ++k) {
                const int src = (c + b.cols - 17) % b.cols;
                uint32_t v = h1[(size_t)r * b.pitch + (size_t)src * K + k];
                if ((rnd() & 7) == 0)
                    v ^= 1u << (rnd() & 31);
                if (r % 5 == 0)
                    v &= 0x3; // ties
                h0[(size_t)r * b.pitch + (size_t)c * K + k] = v;
            }
    for (int r = 0; r < b.rows; r += 5)
        for (size_t i = 0; i < (size_t)b.cols * K; ++i)
            h1[(size_t)r * b.pitch + i] &= 0x3;
    CHECK(cudaMemcpy(b.d0, h0.data(), h0.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(b.d1, h1.data(), h1.size() * 4, cudaMemcpyHostToDevice));
}

template<int K>
void suite(int rows, int cols) {
    Buffers b;
    b.rows = rows;
    b.cols = cols;
    b.px = (size_t)rows * cols;
    b.pitch = ((size_t)cols * K + 3) & ~(size_t)3;
    CHECK(cudaMalloc(&b.d0, b.pitch * rows * 4));
    CHECK(cudaMalloc(&b.d1, b.pitch * rows * 4));
    CHECK(cudaMalloc(&b.keys, b.px * 16));
    fill<K>(b);
    printf("---- K=%d  %d x %d ----\n", K, cols, rows);
    printf("CONSISTENCY\n");
    run<K, 2, 4, 128, 2>("A4 NT128 U2 x1 (one CTA per unit)", b, 1, true);
    run<K, 2, 4, 128, 2>("A4 NT128 U2 x2", b, 2);
    run<K, 2, 4, 128, 2>("A4 NT128 U2 x4", b, 4);
    run<K, 2, 4, 128, 2>("A4 NT128 U2 auto", b, 0);
    run<K, 2, 5, 128, 2>("A5 NT128 U2 auto", b, 0);
    run<K, 2, 3, 128, 2>("A3 NT128 U2 auto", b, 0);
    run<K, 2, 4, 128, 3>("A4 NT128 U3 x1", b, 1);
    if constexpr (K >= 8) { // pairs per thread with one adder less and one POPC more (ALU / XU pipe balance)
        run<K, 2, 4, 128, 2, 0>("A4 NT128 U2 auto MIX0", b, 0);
        run<K, 2, 4, 128, 2, 1>("A4 NT128 U2 auto MIX1", b, 0);
        run<K, 2, 4, 128, 2, 2>("A4 NT128 U2 auto MIX2", b, 0);
        run<K, 2, 4, 128, 2, 3>("A4 NT128 U2 auto MIX3", b, 0);
        run<K, 2, 5, 128, 2, 2>("A5 NT128 U2 auto MIX2", b, 0);
        run<K, 2, 3, 128, 2, 1>("A3 NT128 U2 auto MIX1", b, 0);
    }
    if constexpr (K <= 2) {
        run<K, 2, 8, 128, 2>("A8 NT128 U2 auto", b, 0);
        run<K, 2, 8, 128, 4>("A8 NT128 U4 auto", b, 0);
        run<K, 2, 4, 128, 4>("A4 NT128 U4 auto", b, 0);
        run<K, 2, 8, 64, 2>("A8 NT64 U2 auto", b, 0);
        run<K, 2, 16, 64, 2>("A16 NT64 U2 auto", b, 0);
    }
    printf("NODUPES\n");
    run<K, 1, 4, 128, 2>("A4 NT128 U2 x1 (one CTA per unit)", b, 1, true);
    run<K, 1, 4, 128, 2>("A4 NT128 U2 x2", b, 2);
    run<K, 1, 5, 128, 2>("A5 NT128 U2 auto", b, 0);
    run<K, 1, 3, 128, 2>("A3 NT128 U2 auto", b, 0);
    if constexpr (K >= 8) {
        run<K, 1, 4, 128, 2, 0>("A4 NT128 U2 auto MIX0", b, 0);
        run<K, 1, 4, 128, 2, 1>("A4 NT128 U2 auto MIX1", b, 0);
        run<K, 1, 4, 128, 2, 2>("A4 NT128 U2 auto MIX2", b, 0);
        run<K, 1, 5, 128, 2, 2>("A5 NT128 U2 auto MIX2", b, 0);
    }
    if constexpr (K <= 2) {
        run<K, 1, 8, 128, 2>("A8 NT128 U2 auto", b, 0);
        run<K, 1, 8, 128, 4>("A8 NT128 U4 auto", b, 0);
        run<K, 1, 16, 64, 2>("A16 NT64 U2 auto", b, 0);
    }
    printf("CONSISTENCY | NODUPES\n");
    run<K, 3, 4, 128, 2>("A4 NT128 U2 x1 (one CTA per unit)", b, 1, true);
    run<K, 3, 4, 128, 2>("A4 NT128 U2 x2", b, 2);
    run<K, 3, 5, 128, 2>("A5 NT128 U2 auto", b, 0);
    cudaFree(b.d0);
    cudaFree(b.d1);
    cudaFree(b.keys);
}

int main(int argc, char** argv) {
    const int rows = argc > 1 ? atoi(argv[1]) : 1536;
    const int cols = argc > 2 ? atoi(argv[2]) : 2048;
    const int k = argc > 3 ? atoi(argv[3]) : 4;
    if (k == 4)
        suite<4>(rows, cols);
    else if (k == 8)
        suite<8>(rows, cols);
    else if (k == 12)
        suite<12>(rows, cols);
    else if (k == 16)
        suite<16>(rows, cols);
    else if (k == 2)
        suite<2>(rows, cols);
    else
        suite<1>(rows, cols);
    return 0;
}
