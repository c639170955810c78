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
// below 65536; see DESIGN.md, checked by tests/test_oracle.py::test_integer_mean_comparison_is_exact and tests/test_gpu_parity.py::test_transform_bit_exact).

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

// ---- uint8 LIMITED fast path: 4 pixels per register, byte-lane SIMD ---------------------
//
// ncu shows the scalar kernel above is ALU-bound (one compare + select + insert per bit per
// pixel), not HBM-bound. Here the four adjacent pixels a thread loads as one 32-bit word stay
// packed: an unsigned byte compare a < b of all four lanes is the carry out of b + ~a, i.e.
// maj(b, ~a, carry into bit 7), two integer instructions given the pre-masked low 7 bits of
// every plane word. Pair sums (9 bits) are compared in two 16-bit-lane words (even / odd
// pixels). The mean test p*n < sum becomes p < ceil(sum/n), a byte compare against a
// per-pixel threshold. Result bits are gathered "transposed": byte j of acc[m] holds
// descriptor bits 8m..8m+7 of pixel j; a final byte permute yields the K words per pixel.
// Runtime n: bits of t >= n-2 are masked off afterwards and the 4 tail bits are inserted at
// their runtime position, exactly as in describe_limited().

// x * m + y with m a kernel parameter equal to 1 (or -1): an integer add that ptxas must emit as
// IMAD, i.e. on the FMA pipe. ncu showed this kernel bound by the ALU pipe (LOP3 / SHF / IADD3 /
// PRMT, 91 % busy) with the FMA pipe at 10 %: steering the adds over halves the ALU work.
__device__ __forceinline__ uint32_t fma_add(uint32_t x, uint32_t m, uint32_t y) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(m), "r"(y));
    return r;
}

// bit 7 of every byte: a < b, all other bits 0. Operands pre-split per plane: lo = p & 0x7f..,
// nlo = ~p & 0x7f.., p7 = p & 0x80.., np7 = ~p & 0x80... The 7-bit sums cannot carry across lanes
// (<= 254), bit 7 of t is the carry into the top bit, and maj(b7, ~a7, carry) is the carry out of
// b + ~a = "a < b"; the pre-masked top bits keep every other result bit 0, so the caller can
// accumulate results with shifts and adds, without masking.
__device__ __forceinline__ uint32_t lt_u8x4(uint32_t a_nlo, uint32_t a_np7, uint32_t b_lo, uint32_t b_p7, uint32_t one) {
    uint32_t r;
    const uint32_t t = fma_add(b_lo, one, a_nlo);
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(b_p7), "r"(a_np7), "r"(t));
    return r;
}

__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// EXACT: the stack has exactly 8K+1 images (9 / 17 / 33 / 65: the largest stack of each descriptor
// width, and the common ones), so n is a compile-time constant: no load guards, no masking of
// unused steps, the tail bits sit at a constant position and reuse the planes already loaded.
template<int K, bool EXACT>
__global__ void __launch_bounds__(THREADS) transform_limited_u8x4_kernel(
    const PlaneTable planes,
    int n_runtime,
    int cols,
    size_t in_pitch,
    uint32_t* __restrict__ desc,
    size_t desc_pitch_words,
    uint32_t one, // 1 and -1 as run-time values, see fma_add()
    uint32_t minus_one
) {
    constexpr int NB = 8 * K + 1;
    constexpr uint32_t LO7 = 0x7F7F7F7Fu, MSB = 0x80808080u, H16 = 0x80008000u, B16 = 0x00FF00FFu, C15 = 0x7FFF7FFFu;
    const int n = EXACT ? NB : n_runtime;
    const int row = blockIdx.y;
    const int col = (blockIdx.x * THREADS + threadIdx.x) * 4;
    if (col >= cols)
        return;
    const size_t row_off = (size_t)row * in_pitch + col;

    uint32_t raw[NB + 1];
#pragma unroll
    for (int t = 0; t < NB; ++t)
        raw[t] = t < n ? __ldg(reinterpret_cast<const uint32_t*>(static_cast<const char*>(planes.p[t]) + row_off)) : 0u;
    raw[NB] = 0u;
    uint32_t ta, tb, tpa = 0u, tpb = 0u;
    if constexpr (EXACT) {
        ta = raw[NB - 2];
        tb = raw[NB - 1];
        tpa = raw[NB - 4];
        tpb = raw[NB - 3];
    } else {
        ta = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const char*>(planes.p[n - 2]) + row_off));
        tb = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const char*>(planes.p[n - 1]) + row_off));
        if (n >= 4) {
            tpa = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const char*>(planes.p[n - 4]) + row_off));
            tpb = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const char*>(planes.p[n - 3]) + row_off));
        }
    }

    // per-pixel sums: one IDP4A per pixel and plane (FMA pipe; the byte extraction it replaces was
    // three ALU instructions per plane); padding planes are 0
    uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u;
#pragma unroll
    for (int t = 0; t < NB; ++t) {
        s0 = __dp4a(raw[t], 0x00000001u, s0);
        s1 = __dp4a(raw[t], 0x00000100u, s1);
        s2 = __dp4a(raw[t], 0x00010000u, s2);
        s3 = __dp4a(raw[t], 0x01000000u, s3);
    }
    // thr = ceil(sum / n) per pixel: p*n < sum  <=>  p < thr. Exact reciprocal multiply:
    // floor(x/n) == (x*m) >> 24 with m = ceil(2^24/n) for x < 2^15, n <= 65.
    const uint32_t m = ((1u << 24) + (uint32_t)n - 1u) / (uint32_t)n;
    const uint32_t nm1 = (uint32_t)n - 1u;
    const uint32_t q0 = ((s0 + nm1) * m) >> 24, q1 = ((s1 + nm1) * m) >> 24;
    const uint32_t q2 = ((s2 + nm1) * m) >> 24, q3 = ((s3 + nm1) * m) >> 24;
    const uint32_t thr = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
    const uint32_t thr_lo = thr & LO7, thr_p7 = thr & MSB;

    uint32_t acc[4 * K];
#pragma unroll
    for (int i = 0; i < 4 * K; ++i)
        acc[i] = 0u;

    // pre-masked bit-7 result -> descriptor bit POS of the four pixels (byte j of acc[POS / 8])
    auto put8 = [&](int pos, uint32_t r) {
        acc[pos >> 3] += r >> (7 - (pos & 7));
    };
    // 16-bit-lane results (bit 15 SET means "less"; even word = pixels 0,2, odd word = pixels 1,3):
    // one PRMT in sign-replication mode gathers the four flags as 0x00 / 0xFF bytes in pixel order,
    // one LOP3 drops them into the bit position
    auto put16 = [&](int pos, uint32_t de, uint32_t dod) {
        const uint32_t g = prmt_b32(de, dod, 0xFBD9u);
        acc[pos >> 3] |= g & (0x01010101u << (pos & 7));
    };

    // rolling per-plane splits of p[t], p[t+1] (p[t+2] is split in the iteration)
    uint32_t lo_a = raw[0] & LO7, p7_a = raw[0] & MSB;
    uint32_t lo_b = raw[1] & LO7, p7_b = raw[1] & MSB;
    uint32_t np_prev2_e = 0u, np_prev2_o = 0u, np_prev1_e = 0u, np_prev1_o = 0u; // 0x7fff - ps(t-2), 0x7fff - ps(t-1), 16-bit lanes
    uint32_t e_cur = raw[0] & B16, o_cur = prmt_b32(raw[0], 0u, 0x4341u);
#pragma unroll
    for (int t = 0; t < NB - 2; ++t) {
        const int base = limited_base(t);
        const uint32_t b = raw[t + 1], c = raw[t + 2];
        const uint32_t lo_c = c & LO7;
        const uint32_t p7_c = fma_add(lo_c, minus_one, c); // c - (c & 0x7f..) == c & 0x80..
        const uint32_t nlo_a = lo_a ^ LO7;
        const uint32_t np7_a = fma_add(p7_a, minus_one, MSB); // 0x80.. - p7 == ~p & 0x80..
        put8(base + 0, lt_u8x4(nlo_a, np7_a, lo_b, p7_b, one)); // p[t] < p[t+1]
        put8(base + 1, lt_u8x4(nlo_a, np7_a, lo_c, p7_c, one)); // p[t] < p[t+2]
        put8(base + 2, lt_u8x4(nlo_a, np7_a, thr_lo, thr_p7, one)); // p[t]*n < sum
        const uint32_t e_nxt = b & B16, o_nxt = prmt_b32(b, 0u, 0x4341u);
        const uint32_t pe = fma_add(e_cur, one, e_nxt), po = fma_add(o_cur, one, o_nxt); // ps(t), 16-bit lanes, <= 510
        if (t >= 2) // ps(t-2) < ps(t): bit 15 of ps(t) + 0x7fff - ps(t-2) is set iff the difference is >= 1
            put16(base + 3, fma_add(pe, one, np_prev2_e), fma_add(po, one, np_prev2_o));
        np_prev2_e = np_prev1_e;
        np_prev2_o = np_prev1_o;
        np_prev1_e = fma_add(pe, minus_one, C15);
        np_prev1_o = fma_add(po, minus_one, C15);
        e_cur = e_nxt;
        o_cur = o_nxt;
        lo_a = lo_b;
        p7_a = p7_b;
        lo_b = lo_c;
        p7_b = p7_c;
    }

    // tail bits (descriptor_transform.hpp:62-69), per pixel lanes
    const uint32_t ta_nlo = ~ta & LO7, ta_np7 = ~ta & MSB;
    const uint32_t r0 = lt_u8x4(ta_nlo, ta_np7, tb & LO7, tb & MSB, one);
    const uint32_t r1 = lt_u8x4(ta_nlo, ta_np7, thr_lo, thr_p7, one);
    const uint32_t r2 = lt_u8x4(~tb & LO7, ~tb & MSB, thr_lo, thr_p7, one);
    const uint32_t ab_e = (ta & B16) + (tb & B16), ab_o = ((ta >> 8) & B16) + ((tb >> 8) & B16);
    const uint32_t pv_e = (tpa & B16) + (tpb & B16), pv_o = ((tpa >> 8) & B16) + ((tpb >> 8) & B16);
    // n < 4: the previous pair sum is -1, the bit is always set
    const uint32_t r3e = n >= 4 ? ~((pv_e | H16) - ab_e) : 0xFFFFFFFFu;
    const uint32_t r3o = n >= 4 ? ~((pv_o | H16) - ab_o) : 0xFFFFFFFFu;

    const int pos = n >= 4 ? 4 * n - 10 : 3 * (n - 2); // == limited_base(n - 2)
    const int word = pos >> 5;
    uint32_t keep_mask[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int keep = pos - 32 * k;
        keep_mask[k] = keep <= 0 ? 0u : keep >= 32 ? 0xFFFFFFFFu : ((1u << keep) - 1u);
    }

    // byte j of acc[4k .. 4k+3] -> word k of pixel j: a 4x4 byte transpose per word index, two
    // PRMT stages (8 instead of 12 permutes)
    uint32_t words[4][K]; // [pixel][word]
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t a0 = acc[4 * k], a1 = acc[4 * k + 1], a2 = acc[4 * k + 2], a3 = acc[4 * k + 3];
        const uint32_t t01l = prmt_b32(a0, a1, 0x5140u), t23l = prmt_b32(a2, a3, 0x5140u); // bytes 0,1 of each
        const uint32_t t01h = prmt_b32(a0, a1, 0x7362u), t23h = prmt_b32(a2, a3, 0x7362u); // bytes 2,3 of each
        words[0][k] = prmt_b32(t01l, t23l, 0x5410u);
        words[1][k] = prmt_b32(t01l, t23l, 0x7632u);
        words[2][k] = prmt_b32(t01h, t23h, 0x5410u);
        words[3][k] = prmt_b32(t01h, t23h, 0x7632u);
    }

    uint32_t* out = desc + (size_t)row * desc_pitch_words + (size_t)col * K;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (col + j < cols) {
            const int s16 = (j >> 1) * 16 + 15; // bit of the 16-bit-lane result of pixel j
            const uint32_t r3 = (j & 1) ? r3o : r3e;
            const uint32_t nib = ((r0 >> (8 * j + 7)) & 1u) | (((r1 >> (8 * j + 7)) & 1u) << 1)
                | (((r2 >> (8 * j + 7)) & 1u) << 2) | (((r3 >> s16) & 1u) << 3);
            const unsigned long long ins = (unsigned long long)nib << (pos & 31);
            Bits<K> d;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint32_t v = words[j][k] & keep_mask[k];
                if (k == word)
                    v |= (uint32_t)ins;
                if (k == word + 1)
                    v |= (uint32_t)(ins >> 32);
                d.w[k] = v;
            }
            store_desc<K>(out + j * K, d);
        }
    }
}

constexpr int full_words(int n) {
    const int bits = n * n - 2 * n + 3;
    // 12 and 16 words: the wide-descriptor extension (17..23 images), beyond the reference's 256 bits
    return bits <= 32 ? 1 : bits <= 64 ? 2 : bits <= 128 ? 4 : bits <= 256 ? 8 : bits <= 384 ? 12 : 16;
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
        FULL_CASE(17)
        FULL_CASE(18)
        FULL_CASE(19)
        FULL_CASE(20)
        FULL_CASE(21)
        FULL_CASE(22)
        FULL_CASE(23)
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
    if constexpr (sizeof(TIn) == 1) {
        if (vec_ok) {
            if (n == 8 * K + 1)
                transform_limited_u8x4_kernel<K, true><<<grid, THREADS, 0, stream>>>(planes, n, cols, in_pitch, desc, desc_pitch_words, 1u, 0xFFFFFFFFu);
            else
                transform_limited_u8x4_kernel<K, false><<<grid, THREADS, 0, stream>>>(planes, n, cols, in_pitch, desc, desc_pitch_words, 1u, 0xFFFFFFFFu);
            return cudaGetLastError();
        }
    }
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
