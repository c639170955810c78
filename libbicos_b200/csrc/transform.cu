// Kernel 1 of the BICOS::match hot path: temporal binary descriptor per pixel.
//
// Replaces (behaviour, not structure):
//   reference include/impl/cpu/descriptor_transform.hpp:31-73   transform_limited
//   reference include/impl/cpu/descriptor_transform.hpp:75-123  transform_full
//   reference include/impl/cpu/bitfield.hpp:39-57               LSB-first bit append
//   reference include/impl/cuda/descriptor_transform.cuh:30-149 (CUDA twins)
//
// HBM-bound: n*b bytes in, 4K bytes out per pixel. Each thread owns 4 bytes worth of
// adjacent pixels (4 x u8 or 2 x u16) and reads them with one 32-bit load per image
// plane, so a warp covers 128 contiguous bytes of every plane; descriptors of adjacent
// pixels are stored back to back (16 B vector stores for K >= 4). The pixel stack lives
// in registers: all loops are unrolled over a compile-time bound NB (LIMITED: the largest
// n that fits K words; FULL: the exact n), with warp-uniform guards for the runtime n,
// so every bit position is a compile-time constant.
//
// The mean comparison `(float)p < fl(sum/n)` is evaluated as `p*n < sum` in integers:
// identical for n <= 65 and 16-bit pixels (gap 1/n exceeds half an ulp of any float
// below 65536; see DESIGN.md, checked by tests/test_transform.py against the oracle).

#include "kernels.cuh"

namespace bicos_b200 {
namespace {

constexpr int THREADS = 128;

template<int K>
struct Bits {
    uint32_t w[K];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < K; ++k)
            w[k] = 0u;
    }
};

template<typename TIn>
struct Px;
template<>
struct Px<uint8_t> {
    static constexpr int PER_THREAD = 4;
    __device__ static __forceinline__ int get(uint32_t raw, int q) {
        return (int)((raw >> (8 * q)) & 0xFFu);
    }
};
template<>
struct Px<uint16_t> {
    static constexpr int PER_THREAD = 2;
    __device__ static __forceinline__ int get(uint32_t raw, int q) {
        return (int)((raw >> (16 * q)) & 0xFFFFu);
    }
};

// 4 bytes of plane `t` starting at pixel `col` of `row`; falls back to element loads when
// the vector path is not allowed (unaligned pitch / base) or would cross the row end.
template<typename TIn>
__device__ __forceinline__ uint32_t load_group(
    const void* plane,
    size_t row_off,
    int col,
    int cols,
    bool vec_ok
) {
    constexpr int PT = Px<TIn>::PER_THREAD;
    const TIn* p = reinterpret_cast<const TIn*>(reinterpret_cast<const char*>(plane) + row_off) + col;
    if (vec_ok)
        return __ldg(reinterpret_cast<const uint32_t*>(p));
    uint32_t raw = 0;
#pragma unroll
    for (int q = 0; q < PT; ++q)
        if (col + q < cols)
            raw |= (uint32_t)__ldg(p + q) << (8 * sizeof(TIn) * q);
    return raw;
}

template<int K>
__device__ __forceinline__ void store_desc(uint32_t* dst, const Bits<K>& b) {
    if constexpr (K == 1) {
        dst[0] = b.w[0];
    } else if constexpr (K == 2) {
        *reinterpret_cast<uint2*>(dst) = make_uint2(b.w[0], b.w[1]);
    } else {
#pragma unroll
        for (int q = 0; q < K / 4; ++q)
            reinterpret_cast<uint4*>(dst)[q] =
                make_uint4(b.w[4 * q], b.w[4 * q + 1], b.w[4 * q + 2], b.w[4 * q + 3]);
    }
}

// LIMITED bit layout: t = 0,1 contribute 3 bits, t >= 2 contribute 4 bits.
__host__ __device__ constexpr int limited_base(int t) {
    return t < 2 ? 3 * t : 4 * t - 2;
}

// LIMITED descriptor of one pixel (descriptor_transform.hpp:31-73) without per-element
// branches: the per-t bits are computed for every t up to the compile-time capacity at
// compile-time positions (pixels past n are zero padding), the bits at and above the
// runtime tail position are masked off, and the four tail bits -- which involve p[n-2],
// p[n-1] and the pair sum two steps back -- are inserted at their runtime position.
template<typename TIn, int K>
__device__ __forceinline__ void describe_limited(
    const int (&p)[8 * K + 2],
    int n,
    int ta, // p[n-2]
    int tb, // p[n-1]
    int tprev, // p[n-4] + p[n-3], or -1 when n < 4
    Bits<K>& b
) {
    constexpr int NB = 8 * K + 1; // largest n with 4n-7 <= 32K
    int sum = 0;
#pragma unroll
    for (int t = 0; t < NB; ++t)
        sum += p[t]; // padding is zero
    b.clear();
#pragma unroll
    for (int t = 0; t < NB - 2; ++t) {
        // descriptor_transform.hpp:45-60
        const int base = limited_base(t);
        b.w[(base + 0) / 32] |= (uint32_t)(p[t] < p[t + 1]) << ((base + 0) % 32);
        b.w[(base + 1) / 32] |= (uint32_t)(p[t] < p[t + 2]) << ((base + 1) % 32);
        b.w[(base + 2) / 32] |= (uint32_t)(p[t] * n < sum) << ((base + 2) % 32);
        if (t >= 2)
            b.w[(base + 3) / 32] |= (uint32_t)(p[t >= 2 ? t - 2 : 0] + p[t >= 1 ? t - 1 : 0] < p[t] + p[t + 1]) << ((base + 3) % 32);
    }
    // descriptor_transform.hpp:62-69: tail for a = p[n-2], b = p[n-1]
    const int pos = n >= 4 ? 4 * n - 10 : 3 * (n - 2); // == limited_base(n - 2)
    const uint32_t nib = (uint32_t)(ta < tb) | ((uint32_t)(ta * n < sum) << 1) | ((uint32_t)(tb * n < sum) << 2)
        | ((uint32_t)(tprev < ta + tb) << 3);
    const unsigned long long ins = (unsigned long long)nib << (pos & 31);
    const int word = pos >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int keep = pos - 32 * k; // loop bits of this word that belong to t < n-2
        const uint32_t mask = keep <= 0 ? 0u : keep >= 32 ? 0xFFFFFFFFu : ((1u << keep) - 1u);
        uint32_t v = b.w[k] & mask;
        if (k == word)
            v |= (uint32_t)ins;
        if (k == word + 1)
            v |= (uint32_t)(ins >> 32);
        b.w[k] = v;
    }
}

// FULL layout for an exact compile-time n (descriptor_transform.hpp:75-123)
template<int N, int K>
__device__ __forceinline__ void describe_full(const int (&p)[N], Bits<K>& b) {
    int sum = 0;
#pragma unroll
    for (int t = 0; t < N; ++t)
        sum += p[t];
    int ps[N - 1];
#pragma unroll
    for (int t = 0; t < N - 1; ++t)
        ps[t] = p[t] + p[t + 1];
    b.clear();
    int pos = 0; // folded to constants after unrolling
#pragma unroll
    for (int t = 0; t < N - 2; ++t) {
        b.w[pos / 32] |= (uint32_t)(p[t] < p[t + 1]) << (pos % 32);
        ++pos;
        b.w[pos / 32] |= (uint32_t)(p[t] < p[t + 2]) << (pos % 32);
        ++pos;
        b.w[pos / 32] |= (uint32_t)(p[t] * N < sum) << (pos % 32);
        ++pos;
    }
    b.w[pos / 32] |= (uint32_t)(p[N - 2] < p[N - 1]) << (pos % 32);
    ++pos;
    b.w[pos / 32] |= (uint32_t)(p[N - 2] * N < sum) << (pos % 32);
    ++pos;
    b.w[pos / 32] |= (uint32_t)(p[N - 1] * N < sum) << (pos % 32);
    ++pos;
#pragma unroll
    for (int t = 0; t < N - 1; ++t) {
#pragma unroll
        for (int i = 0; i < N - 1; ++i) {
            if (i == t || i == t - 1 || i == t + 1)
                continue;
            b.w[pos / 32] |= (uint32_t)(ps[t] < ps[i]) << (pos % 32);
            ++pos;
        }
    }
}

template<typename TIn, int K>
__global__ void __launch_bounds__(THREADS) transform_limited_kernel(
    const PlaneTable planes,
    int n,
    int cols,
    size_t in_pitch,
    int vec_ok,
    uint32_t* __restrict__ desc,
    size_t desc_pitch_words
) {
    constexpr int PT = Px<TIn>::PER_THREAD;
    constexpr int NB = 8 * K + 1;
    const int row = blockIdx.y;
    const int col = (blockIdx.x * THREADS + threadIdx.x) * PT;
    if (col >= cols)
        return;
    const size_t row_off = (size_t)row * in_pitch;

    uint32_t raw[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t)
        raw[t] = t < n ? load_group<TIn>(planes.p[t], row_off, col, cols, vec_ok != 0) : 0u;
    // the planes the tail bits need, by runtime index (cache hits: just loaded above)
    const uint32_t raw_a = load_group<TIn>(planes.p[n - 2], row_off, col, cols, vec_ok != 0);
    const uint32_t raw_b = load_group<TIn>(planes.p[n - 1], row_off, col, cols, vec_ok != 0);
    const uint32_t raw_pa = n >= 4 ? load_group<TIn>(planes.p[n - 4], row_off, col, cols, vec_ok != 0) : 0u;
    const uint32_t raw_pb = n >= 4 ? load_group<TIn>(planes.p[n - 3], row_off, col, cols, vec_ok != 0) : 0u;

    uint32_t* out = desc + (size_t)row * desc_pitch_words + (size_t)col * K;
#pragma unroll
    for (int q = 0; q < PT; ++q) {
        if (col + q < cols) {
            int p[NB + 1];
#pragma unroll
            for (int t = 0; t < NB; ++t)
                p[t] = Px<TIn>::get(raw[t], q);
            p[NB] = 0;
            const int tprev = n >= 4 ? Px<TIn>::get(raw_pa, q) + Px<TIn>::get(raw_pb, q) : -1;
            Bits<K> b;
            describe_limited<TIn, K>(p, n, Px<TIn>::get(raw_a, q), Px<TIn>::get(raw_b, q), tprev, b);
            store_desc<K>(out + q * K, b);
        }
    }
}

template<typename TIn, int N, int K>
__global__ void __launch_bounds__(THREADS) transform_full_kernel(
    const PlaneTable planes,
    int cols,
    size_t in_pitch,
    int vec_ok,
    uint32_t* __restrict__ desc,
    size_t desc_pitch_words
) {
    constexpr int PT = Px<TIn>::PER_THREAD;
    const int row = blockIdx.y;
    const int col = (blockIdx.x * THREADS + threadIdx.x) * PT;
    if (col >= cols)
        return;
    const size_t row_off = (size_t)row * in_pitch;

    uint32_t raw[N];
#pragma unroll
    for (int t = 0; t < N; ++t)
        raw[t] = load_group<TIn>(planes.p[t], row_off, col, cols, vec_ok != 0);

    uint32_t* out = desc + (size_t)row * desc_pitch_words + (size_t)col * K;
#pragma unroll
    for (int q = 0; q < PT; ++q) {
        if (col + q < cols) {
            int p[N];
#pragma unroll
            for (int t = 0; t < N; ++t)
                p[t] = Px<TIn>::get(raw[t], q);
            Bits<K> b;
            describe_full<N, K>(p, b);
            store_desc<K>(out + q * K, b);
        }
    }
}

constexpr int full_words(int n) {
    const int bits = n * n - 2 * n + 3;
    return bits <= 32 ? 1 : bits <= 64 ? 2 : bits <= 128 ? 4 : 8;
}

template<typename TIn, int N>
cudaError_t launch_full_n(
    const PlaneTable& planes,
    int rows,
    int cols,
    size_t in_pitch,
    int vec_ok,
    int K,
    uint32_t* desc,
    size_t desc_pitch_words,
    cudaStream_t stream
) {
    constexpr int KK = full_words(N);
    if (K != KK)
        return cudaErrorInvalidValue;
    constexpr int PT = Px<TIn>::PER_THREAD;
    const dim3 grid((cols + THREADS * PT - 1) / (THREADS * PT), rows);
    transform_full_kernel<TIn, N, KK>
        <<<grid, THREADS, 0, stream>>>(planes, cols, in_pitch, vec_ok, desc, desc_pitch_words);
    return cudaGetLastError();
}

template<typename TIn>
cudaError_t launch_full(
    const PlaneTable& planes,
    int n,
    int rows,
    int cols,
    size_t in_pitch,
    int vec_ok,
    int K,
    uint32_t* desc,
    size_t desc_pitch_words,
    cudaStream_t stream
) {
#define FULL_CASE(N) \
    case N: \
        return launch_full_n<TIn, N>(planes, rows, cols, in_pitch, vec_ok, K, desc, desc_pitch_words, stream);
    switch (n) {
        FULL_CASE(2)
        FULL_CASE(3)
        FULL_CASE(4)
        FULL_CASE(5)
        FULL_CASE(6)
        FULL_CASE(7)
        FULL_CASE(8)
        FULL_CASE(9)
        FULL_CASE(10)
        FULL_CASE(11)
        FULL_CASE(12)
        FULL_CASE(13)
        FULL_CASE(14)
        FULL_CASE(15)
        FULL_CASE(16)
    }
#undef FULL_CASE
    return cudaErrorInvalidValue;
}

template<typename TIn, int K>
cudaError_t launch_limited_k(
    const PlaneTable& planes,
    int n,
    int rows,
    int cols,
    size_t in_pitch,
    int vec_ok,
    uint32_t* desc,
    size_t desc_pitch_words,
    cudaStream_t stream
) {
    constexpr int PT = Px<TIn>::PER_THREAD;
    if (n > 8 * K + 1)
        return cudaErrorInvalidValue;
    const dim3 grid((cols + THREADS * PT - 1) / (THREADS * PT), rows);
    transform_limited_kernel<TIn, K>
        <<<grid, THREADS, 0, stream>>>(planes, n, cols, in_pitch, vec_ok, desc, desc_pitch_words);
    return cudaGetLastError();
}

template<typename TIn>
cudaError_t launch_any(
    const PlaneTable& planes,
    int n,
    int rows,
    int cols,
    size_t in_pitch,
    int mode_full,
    int K,
    uint32_t* desc,
    size_t desc_pitch_words,
    cudaStream_t stream
) {
    // vector path: every plane 4 B aligned and the pitch a multiple of 4
    int vec_ok = (in_pitch % 4 == 0);
    for (int t = 0; t < n; ++t)
        if (reinterpret_cast<uintptr_t>(planes.p[t]) % 4 != 0)
            vec_ok = 0;
    if (mode_full)
        return launch_full<TIn>(planes, n, rows, cols, in_pitch, vec_ok, K, desc, desc_pitch_words, stream);
    switch (K) {
        case 1:
            return launch_limited_k<TIn, 1>(planes, n, rows, cols, in_pitch, vec_ok, desc, desc_pitch_words, stream);
        case 2:
            return launch_limited_k<TIn, 2>(planes, n, rows, cols, in_pitch, vec_ok, desc, desc_pitch_words, stream);
        case 4:
            return launch_limited_k<TIn, 4>(planes, n, rows, cols, in_pitch, vec_ok, desc, desc_pitch_words, stream);
        case 8:
            return launch_limited_k<TIn, 8>(planes, n, rows, cols, in_pitch, vec_ok, desc, desc_pitch_words, stream);
    }
    return cudaErrorInvalidValue;
}

} // namespace

cudaError_t launch_transform(
    const PlaneTable& planes,
    int n,
    int rows,
    int cols,
    size_t in_pitch,
    int is_u16,
    int mode_full,
    int K,
    uint32_t* desc,
    size_t desc_pitch_words,
    cudaStream_t stream
) {
    if (n < 2 || n > MAX_IMAGES || rows <= 0 || cols <= 0 || rows > 65535)
        return cudaErrorInvalidValue;
    if (is_u16)
        return launch_any<uint16_t>(planes, n, rows, cols, in_pitch, mode_full, K, desc, desc_pitch_words, stream);
    return launch_any<uint8_t>(planes, n, rows, cols, in_pitch, mode_full, K, desc, desc_pitch_words, stream);
}

} // namespace bicos_b200
