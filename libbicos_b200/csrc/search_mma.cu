// Kernel 2 of the BICOS::match hot path, tensor-core engine: the W x W Hamming matrix of one
// rectified row as an int8 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8,
// accumulators in TMEM), with the argmin as the epilogue. Same results, bit for bit, as the
// popcount engine in search.cu.
//
// Replaces (behaviour, not structure):
//   reference include/impl/cpu/bicos.hpp:29-76   ham / bicos_search
//   reference include/impl/cuda/bicos.cuh:50-176 bicos_search / bicos_kernel[_smem]
//
// Why a GEMM is exact here. With a the left and b the right descriptor,
//     ham(a, b) - popc(a) = #(a = 0, b = 1) - #(a = 1, b = 1) = sum_k b_k * (1 - 2 a_k).
// Bit 8i + s of a 32-bit descriptor word becomes, in the right operand, the unsigned byte b * 2^s
// (ONE instruction per four bytes: w & 0x01010101 << s; for s = 0 the byte is b * 128) and, in
// the left operand, the signed byte (1 - 2a) * 2^(7 - s) (+-1 for s = 0). Every product is
// +-128, so
//     acc(i, j) = 128 * (ham(i, j) - popc(a_i))                                   (int32, exact)
// and acc + u, u the column within the 128-column tile, orders the pairs of a tile by
// (ham, column): the minimum is the reference's first strict minimum (bicos.hpp:57-60); acc +
// (127 - u) finds the last column at the minimal cost, whose difference from the first is the
// no-duplicates test (bicos.hpp:62-71). Tiles are merged as 8192 * (ham - popc) + column, which
// is why rows of up to 8192 pixels are supported. For 128-bit descriptors |acc + u| < 2^15:
// tcgen05.ld packs the low halves of two accumulator columns into one register and one
// VIADDMNMX.S16x2 folds two pairs (the popcount engine needs 3 POPC + 6 LOP3 + 5 more per pair);
// wider descriptors use the 32-bit VIADDMNMX, one pair per instruction.
// The reverse search of the consistency check (bicos.hpp:99-106) is, in the first two kernels, the same
// kernel with the operands swapped (the second half of the work items): on the tensor cores the second
// W x W x KBITS product is cheaper than column-wise minima of the first TAKEN PER TILE; the third kernel
// takes them per item instead and computes the product once.
//
// Persistent CTAs walk contiguous ranges of work items (direction, row, M tile); within a CTA four
// roles are hand-shaken by mbarriers only:
//   loader     one thread: TMA bulk copies of the packed descriptors of the next 128-column tile
//              into a small ring, across item boundaries (no global loads in the producers: the
//              fence that publishes a tile to the async proxy waits for all loads of its thread)
//   producers  4 warps: one packed descriptor per thread -> a 128-byte row of the uint8 tile in
//              shared memory (128B-swizzled K-major, the layout a tensor-map TMA load would produce)
//   issuer     a whole warp with uniform control flow, one elected lane issuing KBITS/32 tcgen05.mma
//              (128 x 128 x 32, kind::i8) per tile and the commits to the stage / accumulator barriers
//   epilogue   4 warps per 128 left pixels (the four TMEM lane quadrants): tcgen05.ld the accumulator,
//              hand it back, fold it into the running minima; expand the next item's left tile in passing
// Three kernels: search_mma_kernel (every width; 2 CTAs per SM up to 256 bits; 128 left pixels per item,
// both operands in shared memory, 2 accumulators), search_mma2_kernel (128 / 256 bits, large images;
// 1 CTA per SM, 256 left pixels per item, left operand resident in TMEM, 3 accumulators in rotation,
// one issuer per half) and search_mma3_kernel (128 / 256 bits, Consistency without no_dupes: ONE product per
// row, the forward minima folded in-thread and the reverse minima elementwise across the row's tiles,
// see its own header below). DESIGN.md 3.2a / 3.2b have the measurements that led from one to the next.

#include "kernels.cuh"

#include <atomic>
#include <limits.h>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace bicos_b200 {
namespace {

constexpr int TM = 128; // left pixels per CTA = TMEM lanes
constexpr int TN = 128; // right pixels per accumulator
constexpr int NTHREADS = 320; // 4 epilogue warps, 4 producer warps, 1 MMA warp, 1 loader warp
constexpr int ATOM_BYTES = 128 * 128; // 128 pixels x 128 descriptor bits as int8: 128-byte rows, one swizzle atom wide
constexpr uint32_t TMEM_COLS = 2 * TN;
constexpr int COL_BITS = 13; // merged keys step by 8192 per unit of Hamming distance
constexpr int COL_MAX = (1 << COL_BITS) - 1;
#ifndef BICOS_MMA_COLTERM_DEFAULT
#define BICOS_MMA_COLTERM_DEFAULT 1 // validated on the B200 (tools/search_engines ... 1), 1.25 -> 1.05 ms on the metric configuration
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory"
    );
    return ok != 0;
}

// A wait that cannot hang the device and does not kill the context either. Fast path: one try_wait. Slow
// path (out of line): try_wait with __nanosleep back-off under a wall-clock bound read from %globaltimer
// (seconds, MmaArgs::timeout_ns; a bound in iterations would depend on how long try_wait suspends and could
// expire under a debugger, compute-sanitizer or time-slicing). When the bound expires the thread raises the
// CTA's abort word in shared memory and a flag in mapped host memory, every other role sees the abort word
// in its own slow path, all roles leave their loops and the kernel ends normally: the key arrays are then
// garbage, and the C ABI reports "search pipeline timeout" at its next synchronisation point
// (search_mma_take_timeout) instead of the process losing its CUDA context to a trap.
struct Watch {
    unsigned int* flag; // mapped host memory, see mma_timeout_flag()
    unsigned long long timeout_ns;
    unsigned int abort;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, Watch* watch) {
    const unsigned long long t0 = global_ns();
    unsigned int pause = 32;
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 32; ++i)
            if (mbar_try(bar, parity))
                return true;
        if (*(volatile unsigned int*)&watch->abort)
            return false;
        if (global_ns() - t0 > watch->timeout_ns) {
#ifdef BICOS_MMA_DEBUG
            printf("mbar timeout: block %d warp %d lane %d barrier +%u parity %u\n", blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31,
                   bar & 1023u, parity);
#endif
            *(volatile unsigned int*)&watch->abort = 1u;
            if (watch->flag)
                *(volatile unsigned int*)watch->flag = 1u + blockIdx.x;
            return false;
        }
        __nanosleep(pause);
        pause = pause < 1024 ? pause * 2 : 1024;
    }
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, Watch* watch) {
    if (mbar_try(bar, parity))
        return true;
    return mbar_wait_slow(bar, parity, watch);
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// TMA bulk copy global -> shared (contiguous bytes, 16-byte granular), completion counted on `bar`
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// One lane of a converged warp. The MMA issuer runs as a whole warp with uniform control flow and uniform
// operands and only the tcgen05 instructions themselves are elected: a `tid == first lane` branch around
// the loop makes ptxas wrap every uniform-register operand in an ELECT / BRA.U.ANY loop, and the issuing
// thread, not the tensor pipe, becomes the bound (ncu source page, r01).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(pred)
    );
    return pred != 0;
}

__device__ __forceinline__ uint32_t uniform(uint32_t v) {
    return __shfl_sync(0xFFFFFFFFu, v, 0);
}

__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M = 128, N = 128, K = 32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory"
    );
}

// 32 consecutive accumulator columns of this thread's lane. The load is asynchronous: the registers
// are valid after tc_load32_wait on the same array (which takes them as in/out operands, so that
// neither nvcc nor ptxas can move a use above the wait).
__device__ __forceinline__ void tc_load32_issue(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory"
    );
}

// 64 consecutive columns, the low 16 bits of columns 2r and 2r + 1 packed into register r
__device__ __forceinline__ void tc_load64_packed_issue(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory"
    );
}

__device__ __forceinline__ void tc_load32_wait(int (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory");
}

// 32 consecutive 32-bit columns of this thread's lane <- registers (asynchronous; tc_store_wait before the
// fence that publishes them)
__device__ __forceinline__ void tc_store32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory"
    );
}

__device__ __forceinline__ void tc_store_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T: the left operand read from tensor memory (lane = row, four int8 of K
// per 32-bit column), so that only the streamed operand costs shared-memory bandwidth
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
        "}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory"
    );
}

// Shared-memory matrix descriptor (sm_100 format): K-major rows of 128 bytes, 128-byte swizzle,
// groups of 8 rows 1024 bytes apart. `saddr` may point 32 * k bytes into the first row of a
// 1024-byte aligned tile (the k-th 32-byte slice of K); the swizzle acts on address bits.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4); // start address
    d |= (uint64_t)1 << 16; // leading byte offset: not used by swizzled K-major layouts
    d |= (uint64_t)(1024 >> 4) << 32; // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46; // descriptor version
    d |= (uint64_t)2 << 61; // SWIZZLE_128B
    return d;
}

// instruction descriptor: D = s32, A = signed int8, B = unsigned int8, both K-major, N = 128, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Bits s, 8 + s, 16 + s, 24 + s of a descriptor word -> four operand bytes (see the header).
template<bool RIGHT, int S>
__device__ __forceinline__ uint32_t expand_word(uint32_t w) {
    if constexpr (RIGHT) {
        if constexpr (S == 0)
            return (w << 7) & 0x80808080u; // b * 128
        else
            return w & (0x01010101u << S); // b * 2^S
    } else {
        constexpr int P = S == 0 ? 0 : 7 - S; // magnitude 2^P
        constexpr uint32_t MAG = 0x01010101u << P;
        constexpr uint32_t HIGH = 0x01010101u * ((0xFFu << (P + 1)) & 0xFFu); // -2^P = 2^P | HIGH, per byte
        const uint32_t neg = ((w >> S) & 0x01010101u) * 0xFFu; // 0xFF where the bit is set
        return (neg & HIGH) | MAG;
    }
}

// One pixel's descriptor (K / 4 uint4 in registers) -> row r of K/4 swizzled atoms starting at `tile`.
// Word wi of an atom fills the 16-byte chunks 2 wi (s = 0..3) and 2 wi + 1 (s = 4..7).
template<int K, bool RIGHT, bool CT = false>
__device__ __forceinline__ void expand_pixel(const uint4 (&d)[K / 4], uint32_t tile, int r) {
    const uint32_t row = tile + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const uint32_t w[4] = { d[q].x, d[q].y, d[q].z, d[q].w };
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            st_shared_v4(
                row + (uint32_t)q * ATOM_BYTES + ((((uint32_t)(2 * wi)) ^ sw) << 4),
                expand_word<RIGHT, 0>(w[wi]),
                expand_word<RIGHT, 1>(w[wi]),
                expand_word<RIGHT, 2>(w[wi]),
                expand_word<RIGHT, 3>(w[wi])
            );
            st_shared_v4(
                row + (uint32_t)q * ATOM_BYTES + ((((uint32_t)(2 * wi + 1)) ^ sw) << 4),
                expand_word<RIGHT, 4>(w[wi]),
                expand_word<RIGHT, 5>(w[wi]),
                expand_word<RIGHT, 6>(w[wi]),
                // CT: the byte of the (unused, zero) top descriptor bit carries this row's tile column
                (RIGHT && CT && q == K / 4 - 1 && wi == 3) ? (expand_word<RIGHT, 7>(w[wi]) | ((uint32_t)r << 24)) : expand_word<RIGHT, 7>(w[wi])
            );
        }
    }
}

template<int K>
__device__ __forceinline__ void load_pixel(const uint32_t* __restrict__ desc, uint4 (&d)[K / 4]) {
#pragma unroll
    for (int q = 0; q < K / 4; ++q)
        d[q] = __ldg(reinterpret_cast<const uint4*>(desc) + q);
}

// Running minima of one 128-column tile: acc + u (first) and acc + 127 - u (last), u the column
// within the tile, in four independent chains. Packed: two 16-bit lanes per register.
struct TileMin32 {
    int f[4], l[4];
    __device__ __forceinline__ TileMin32() {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            f[c] = l[c] = INT_MAX;
    }
};

struct TileMin16 {
    uint32_t f[8], l[8]; // eight chains: the epilogue warps are few, so the fold must not wait on itself
    __device__ __forceinline__ TileMin16() {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            f[c] = l[c] = 0x7FFF7FFFu;
    }
};

// CT ("column term"): the top descriptor bit (bit 32 K - 1) is unused in every descriptor the transform
// writes (4n - 6 and n^2 - 2n + 3 are never a multiple of 32). Its byte in the left operand is therefore
// always +1, and the producers put the tile column u into that byte of the right operand: the MMA itself
// delivers acc + u, the first-minimum fold needs no addition and becomes a three-input minimum
// (VIMNMX3[.S16x2]: two columns, or four packed ones, per instruction); the last-minimum key
// acc + 127 - u is (acc + u) + (127 - 2u).

// 32 accumulator columns starting at tile column U0
template<bool NODUPES, int U0, bool CT>
__device__ __forceinline__ void fold32(const int (&v)[32], TileMin32& m) {
    if constexpr (CT) {
#pragma unroll
        for (int u = 0; u < 32; u += 2)
            m.f[(u >> 1) & 3] = __vimin3_s32(m.f[(u >> 1) & 3], v[u], v[u + 1]);
    }
#pragma unroll
    for (int u = 0; u < 32; ++u) {
        if constexpr (!CT)
            m.f[u & 3] = min(m.f[u & 3], v[u] + (U0 + u));
        if constexpr (NODUPES)
            m.l[u & 3] = min(m.l[u & 3], v[u] + (CT ? 127 - 2 * (U0 + u) : 127 - U0 - u));
    }
}

// the same with a run-time bound: columns from `valid` on do not exist
template<bool NODUPES, bool CT>
__device__ __forceinline__ void fold32_guarded(const int (&v)[32], int u0, int valid, TileMin32& m) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
        if (u0 + u < valid) {
            m.f[u & 3] = min(m.f[u & 3], CT ? v[u] : v[u] + (u0 + u));
            if constexpr (NODUPES)
                m.l[u & 3] = min(m.l[u & 3], v[u] + (CT ? 127 - 2 * (u0 + u) : 127 - u0 - u));
        }
    }
}

// 64 accumulator columns starting at tile column U0, low halves packed two per register
template<bool NODUPES, int U0, bool CT>
__device__ __forceinline__ void fold64_packed(const int (&v)[32], TileMin16& m) {
    if constexpr (CT) {
#pragma unroll
        for (int r = 0; r < 32; r += 2)
            m.f[(r >> 1) & 7] = __vimin3_s16x2(m.f[(r >> 1) & 7], (uint32_t)v[r], (uint32_t)v[r + 1]);
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) {
        constexpr uint32_t ONE = 0x00010001u;
        const uint32_t uf = (uint32_t)(U0 + 2 * r) * ONE + 0x00010000u; // columns U0 + 2r | U0 + 2r + 1
        // 127 - column, or with the column already in the accumulator 127 - 2 column (16-bit lanes, two's complement)
        const uint32_t ul = CT ? ((uint32_t)(uint16_t)(int16_t)(127 - 2 * (U0 + 2 * r + 1)) << 16) | (uint32_t)(uint16_t)(int16_t)(127 - 2 * (U0 + 2 * r))
                               : (uint32_t)(127 - U0 - 2 * r) * ONE - 0x00010000u;
        if constexpr (!CT)
            m.f[r & 7] = __viaddmin_s16x2((uint32_t)v[r], uf, m.f[r & 7]);
        if constexpr (NODUPES)
            m.l[r & 7] = __viaddmin_s16x2((uint32_t)v[r], ul, m.l[r & 7]);
    }
}

// tile minimum t = 128 * (ham - popc) + u  ->  8192 * (ham - popc) + u
__device__ __forceinline__ int widen_key(int t) {
    return ((t & ~127) << 6) + (t & 127);
}

template<bool NODUPES>
__device__ __forceinline__ void merge_tile(const TileMin32& m, int tile0, int& m_first, int& m_last) {
    m_first = min(m_first, widen_key(min(min(m.f[0], m.f[1]), min(m.f[2], m.f[3]))) + tile0);
    if constexpr (NODUPES)
        m_last = min(m_last, widen_key(min(min(m.l[0], m.l[1]), min(m.l[2], m.l[3]))) + (COL_MAX - 127 - tile0));
}

__device__ __forceinline__ int min_of_lanes(const uint32_t (&c)[8]) {
    const uint32_t m = __vmins2(__vmins2(__vmins2(c[0], c[1]), __vmins2(c[2], c[3])), __vmins2(__vmins2(c[4], c[5]), __vmins2(c[6], c[7])));
    return min((int)(short)(m & 0xFFFFu), (int)(short)(m >> 16));
}

template<bool NODUPES>
__device__ __forceinline__ void merge_tile(const TileMin16& m, int tile0, int& m_first, int& m_last) {
    m_first = min(m_first, widen_key(min_of_lanes(m.f)) + tile0);
    if constexpr (NODUPES)
        m_last = min(m_last, widen_key(min_of_lanes(m.l)) + (COL_MAX - 127 - tile0));
}

struct MmaArgs {
    const uint32_t* left;
    const uint32_t* right;
    int rows, cols;
    size_t pitch_words;
    int mtiles; // ceil(cols / TM)
    int ntiles; // ceil(cols / TN)
    long long items; // directions x rows x mtiles
    uint32_t* fwd_first; // direction 0: per left pixel over the right row
    uint32_t* fwd_last;
    uint32_t* rev_first; // direction 1: per right pixel over the left row
    uint32_t* rev_last;
    unsigned int* timeout_flag; // mapped host word a timed-out wait raises (mbar_wait_slow)
    unsigned long long timeout_ns;
};

// right-tile stages and left-tile buffers in shared memory: what fits beside a second CTA (K <= 8) or alone
template<int K>
constexpr int STAGES = K == 4 ? 4 : K == 12 ? 3 : 2;
template<int K>
constexpr int LEFT_BUFFERS = K == 4 ? 2 : 1;
// ring of packed right tiles (TN descriptors as they lie in global memory), filled by TMA bulk copies
template<int K>
constexpr int PACKED_STAGES = K == 4 ? 4 : K == 8 ? 3 : 2;

// A work item = 128 pixels (M tile `mt`) of one row in one direction against the whole other row.
// Items are numbered direction-major, then row, then M tile; a CTA walks a contiguous range.
struct Item {
    int dir, row, mt;
    __device__ __forceinline__ void decode(long long item, int rows, int mtiles) {
        const long long per_dir = (long long)rows * mtiles;
        dir = (int)(item / per_dir);
        const int rem = (int)(item - dir * per_dir);
        row = rem / mtiles;
        mt = rem - row * mtiles;
    }
    __device__ __forceinline__ void next(int rows, int mtiles) {
        if (++mt == mtiles) {
            mt = 0;
            if (++row == rows) {
                row = 0;
                ++dir;
            }
        }
    }
};

template<int K, bool NODUPES, bool CT>
__global__ void __launch_bounds__(NTHREADS, (K <= 8) ? 2 : 1) search_mma_kernel(const MmaArgs p) {
    constexpr int KA = K / 4; // 128-bit atoms
    constexpr int NS = STAGES<K>;
    constexpr int NA = LEFT_BUFFERS<K>;
    extern __shared__ uint8_t smem_raw[];
    constexpr int NP = PACKED_STAGES<K>;
    constexpr int PACKED_BYTES = TN * K * 4;
    // stage full [NS], stage free [NS], accumulator full [2], accumulator drained [2], left tile full [NA],
    // packed full [NP], packed free [NP]
    __shared__ uint64_t bars[2 * NS + 4 + NA + 2 * NP];
    __shared__ uint32_t tmem_base_slot;
    __shared__ Watch s_watch;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int cols = p.cols;
    const int ntiles = p.ntiles;
    const long long item0 = p.items * blockIdx.x / gridDim.x;
    const int nitems = (int)(p.items * (blockIdx.x + 1) / gridDim.x - item0);

    const uint32_t s_a = (smem_u32(smem_raw) + 1023u) & ~1023u; // + buffer * KA * ATOM_BYTES
    const uint32_t s_b = s_a + NA * KA * ATOM_BYTES; // + stage * KA * ATOM_BYTES
    const uint32_t bar_stage_full = smem_u32(&bars[0]); // + 8 * stage
    const uint32_t bar_stage_free = bar_stage_full + 8 * NS;
    const uint32_t bar_acc_full = bar_stage_free + 8 * NS; // + 8 * accumulator
    const uint32_t bar_acc_drained = bar_acc_full + 16;
    const uint32_t bar_left_full = bar_acc_drained + 16; // + 8 * buffer
    const uint32_t bar_packed_full = bar_left_full + 8 * NA; // + 8 * packed stage
    const uint32_t bar_packed_free = bar_packed_full + 8 * NP;
    const uint32_t s_packed = s_b + NS * KA * ATOM_BYTES; // + packed stage * PACKED_BYTES

    if (tid == 0) {
        s_watch.flag = p.timeout_flag;
        s_watch.timeout_ns = p.timeout_ns;
        s_watch.abort = 0;
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_stage_full + 8 * s, TN);
            mbar_init(bar_stage_free + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_acc_full + 8 * a, 1);
            mbar_init(bar_acc_drained + 8 * a, TM);
        }
        for (int b = 0; b < NA; ++b)
            mbar_init(bar_left_full + 8 * b, TM);
        for (int s = 0; s < NP; ++s) {
            mbar_init(bar_packed_full + 8 * s, 1);
            mbar_init(bar_packed_free + 8 * s, TN);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;

    Item it;
    it.decode(item0, p.rows, p.mtiles);

    if (warp == 8) {
        // ---- MMA issuer: the whole warp walks the tiles, one elected lane issues. Stage / accumulator
        //      indices and barrier phases are carried incrementally and the descriptors differ from a
        //      per-stage base by constants: this loop's own latency bounds the kernel otherwise. ----
        {
            const uint32_t u_tmem = uniform(tmem);
            const uint64_t desc_a0 = smem_desc(uniform(s_a)), desc_b0 = smem_desc(uniform(s_b));
            constexpr uint32_t STAGE_STEP = (KA * ATOM_BYTES) >> 4; // descriptor start-address units (16 B)
            uint32_t s = 0, stage_phase = 0; // stage of tile g and the parity its "full" barrier completes with
            uint32_t a = 0; // accumulator g & 1
            uint32_t b = 0, left_phase = 0;
            int g = 0;
            for (int n = 0; n < nitems; ++n) {
                if (!mbar_wait(bar_left_full + 8 * b, left_phase, &s_watch))
                    goto teardown;
                const uint64_t desc_a = desc_a0 + b * STAGE_STEP;
                for (int t = 0; t < ntiles; ++t, ++g) {
                    if (!mbar_wait(bar_stage_full + 8 * s, stage_phase, &s_watch))
                        goto teardown;
                    if (g >= 2)
                        if (!mbar_wait(bar_acc_drained + 8 * a, ((g - 2) >> 1) & 1, &s_watch)) // the epilogue has read this accumulator
                            goto teardown;
                    tc_fence_after();
                    const uint64_t desc_b = desc_b0 + s * STAGE_STEP;
                    if (elect_one()) {
#pragma unroll
                        for (int q = 0; q < KA; ++q)
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                tc_mma_i8(
                                    u_tmem + a * TN,
                                    desc_a + (uint32_t)((q * ATOM_BYTES + kk * 32) >> 4),
                                    desc_b + (uint32_t)((q * ATOM_BYTES + kk * 32) >> 4),
                                    IDESC,
                                    (q | kk) != 0
                                );
                        tc_commit(bar_stage_free + 8 * s);
                        tc_commit(bar_acc_full + 8 * a);
                    }
                    __syncwarp();
                    if (++s == NS) {
                        s = 0;
                        stage_phase ^= 1;
                    }
                    a ^= 1;
                }
                if (++b == NA) {
                    b = 0;
                    left_phase ^= 1;
                }
            }
        }
    } else if (warp == 9) {
        // ---- loader: one thread streams the packed descriptors of the other image's rows, tile by tile
        //      and across item boundaries, into a small ring with TMA bulk copies ----
        if (tid == 9 * 32) {
            int f = 0;
            for (int n = 0; n < nitems; ++n) {
                const uint32_t* const row = (it.dir ? p.left : p.right) + (size_t)it.row * p.pitch_words;
                for (int t = 0; t < ntiles; ++t, ++f) {
                    const int s = f % NP;
                    if (f >= NP)
                        if (!mbar_wait(bar_packed_free + 8 * s, (f / NP - 1) & 1, &s_watch))
                            goto teardown;
                    const uint32_t bytes = (uint32_t)min(TN, cols - t * TN) * K * 4;
                    mbar_expect_tx(bar_packed_full + 8 * s, bytes);
                    bulk_copy_g2s(s_packed + (uint32_t)(s * PACKED_BYTES), row + (size_t)t * TN * K, bytes, bar_packed_full + 8 * s);
                }
                it.next(p.rows, p.mtiles);
            }
        }
    } else if (warp >= 4) {
        // ---- producers: one packed descriptor per thread -> a row of the uint8 tile in shared memory.
        //      No global loads here: the fence that publishes the tile to the tensor cores waits for
        //      every outstanding load of its thread, which would expose an L2 round trip per tile. ----
        const int r = tid - TM;
        const int total = nitems * ntiles;
        int t = 0;
        for (int g = 0; g < total; ++g) {
            const int ps = g % NP, s = g % NS;
            const int valid = min(TN, cols - t * TN); // the last tile of a row may be short: repeat its last pixel
            if (!mbar_wait(bar_packed_full + 8 * ps, (g / NP) & 1, &s_watch))
                goto teardown;
            uint4 d[KA];
            const uint32_t src = s_packed + (uint32_t)(ps * PACKED_BYTES) + (uint32_t)min(r, valid - 1) * (K * 4);
#pragma unroll
            for (int q = 0; q < KA; ++q)
                d[q] = ld_shared_v4(src + 16 * q);
            if (g >= NS)
                if (!mbar_wait(bar_stage_free + 8 * s, (g / NS - 1) & 1, &s_watch)) // the MMAs that read this stage are done
                    goto teardown;
            expand_pixel<K, true, CT>(d, s_b + (uint32_t)(s * KA * ATOM_BYTES), r);
            // only now: the stores above consumed the loaded registers, so the packed slot has been read
            // (an arrive right after the loads was observed to let the next bulk copy overtake them)
            mbar_arrive(bar_packed_free + 8 * ps);
            fence_async_smem();
            mbar_arrive(bar_stage_full + 8 * s);
            if (++t == ntiles)
                t = 0;
        }
    } else {
        // ---- epilogue: running minima of 8192 * (ham - popc) + column over the row ----
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        int pa_next = 0;
        // this thread's pixel of item `at` -> row `tid` of left buffer `b`; returns its popcount
        auto load_left = [&](const Item& at, uint4(&d)[KA]) {
            const uint32_t* const row = (at.dir ? p.right : p.left) + (size_t)at.row * p.pitch_words;
            load_pixel<K>(row + (size_t)min(at.mt * TM + tid, cols - 1) * K, d);
        };
        auto stage_left = [&](const uint4(&d)[KA], int n) {
            const int b = n % NA;
            expand_pixel<K, false>(d, s_a + (uint32_t)(b * KA * ATOM_BYTES), tid);
            fence_async_smem();
            mbar_arrive(bar_left_full + 8 * b);
            int pc = 0;
#pragma unroll
            for (int q = 0; q < KA; ++q)
                pc += __popc(d[q].x) + __popc(d[q].y) + __popc(d[q].z) + __popc(d[q].w);
            return pc;
        };
        uint4 dl[KA];
        if (nitems > 0) {
            load_left(it, dl);
            pa_next = stage_left(dl, 0);
        }
        int g = 0;
        for (int n = 0; n < nitems; ++n) {
            const int pa = pa_next; // popcount of this thread's own descriptor
            Item nx = it;
            nx.next(p.rows, p.mtiles);
            int m_first = INT_MAX, m_last = INT_MAX;
            if (n + 1 < nitems)
                load_left(nx, dl); // in flight over the first tiles of this item
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int a = g & 1;
                if (!mbar_wait(bar_acc_full + 8 * a, (g >> 1) & 1, &s_watch))
                    goto teardown;
                tc_fence_after();
                const uint32_t acc = lane_base + (uint32_t)(a * TN);
                const int tile0 = t * TN;
                int va[32], vb[32];
                if (tile0 + TN > cols) {
                    // ragged last tile
                    TileMin32 m;
#pragma unroll 1
                    for (int u0 = 0; u0 < TN && tile0 + u0 < cols; u0 += 32) {
                        tc_load32_issue(acc + (uint32_t)u0, va);
                        tc_load32_wait(va);
                        fold32_guarded<NODUPES, CT>(va, u0, cols - tile0, m);
                    }
                    tc_fence_before();
                    mbar_arrive(bar_acc_drained + 8 * a);
                    merge_tile<NODUPES>(m, tile0, m_first, m_last);
                } else if constexpr (K == 4) {
                    // 16-bit lanes: the whole accumulator goes to 64 registers and is handed back before the
                    // fold, so that the next MMAs into it overlap the fold
                    TileMin16 m;
                    tc_load64_packed_issue(acc, va);
                    tc_load64_packed_issue(acc + 64, vb);
                    tc_load32_wait(va);
                    tc_load32_wait(vb);
                    tc_fence_before();
                    mbar_arrive(bar_acc_drained + 8 * a);
                    fold64_packed<NODUPES, 0, CT>(va, m);
                    fold64_packed<NODUPES, 64, CT>(vb, m);
                    merge_tile<NODUPES>(m, tile0, m_first, m_last);
                } else {
                    TileMin32 m;
                    tc_load32_issue(acc, va);
                    tc_load32_wait(va);
                    tc_load32_issue(acc + 32, vb);
                    fold32<NODUPES, 0, CT>(va, m);
                    tc_load32_wait(vb);
                    tc_load32_issue(acc + 64, va);
                    fold32<NODUPES, 32, CT>(vb, m);
                    tc_load32_wait(va);
                    tc_load32_issue(acc + 96, vb);
                    fold32<NODUPES, 64, CT>(va, m);
                    tc_load32_wait(vb);
                    tc_fence_before();
                    mbar_arrive(bar_acc_drained + 8 * a);
                    fold32<NODUPES, 96, CT>(vb, m);
                    merge_tile<NODUPES>(m, tile0, m_first, m_last);
                }
                // The next item's left tile. With two buffers: early, the other buffer was last read by the
                // previous item, whose MMAs are all complete. With one: once this item's last MMAs are
                // complete, which the wait on its last accumulator has established. Its descriptor was requested at
                // the start of the item.
                if (n + 1 < nitems && t == (NA == 2 ? min(2, ntiles - 1) : ntiles - 1))
                    pa_next = stage_left(dl, n + 1);
            }
            const int i = it.mt * TM + tid;
            if (i < cols) {
                const size_t at = (size_t)it.row * cols + i;
                // m = 8192 * (cost - popc) + column: arithmetic shift = floor, the low bits are the column
                (it.dir ? p.rev_first : p.fwd_first)[at] = ((uint32_t)(pa + (m_first >> COL_BITS)) << 16) | ((uint32_t)m_first & COL_MAX);
                if constexpr (NODUPES)
                    (it.dir ? p.rev_last : p.fwd_last)[at] =
                        ((uint32_t)(pa + (m_last >> COL_BITS)) << 16) | ((65535u - COL_MAX) + ((uint32_t)m_last & COL_MAX));
            }
            it = nx;
        }
    }

teardown:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Variant for 128- and 256-bit descriptors: ONE CTA per SM, 256 left pixels per work item, the left
// operand resident in tensor memory. The kernel above is bound by the shared-memory pipe (MMA operand
// reads 32 KB + expansion stores 16 KB per 128 x 128 tile); here an MMA reads only the streamed operand
// from shared memory (16 KB) and every expanded right tile serves two MMA groups (8 KB of stores per
// 128 x 128 of output). TMEM: three 128-column accumulators in rotation (columns 0..383) + the left
// operand, 2 x 32 columns per 128 descriptor bits, double-buffered for 128 bits (columns 384..511).
//   warps 0-3 / 4-7  epilogue of the first / second 128 left pixels (the four lane quadrants each)
//   warps 8-11       producers        warps 12, 13  MMA issuers (one per half)        warp 14  loader   warp 15  idle
// Registers: only the epilogue needs many (the 64 accumulator registers of a tile). The kernel is compiled for
// V2_REGS_LAUNCH registers per thread and the roles re-balance them with setmaxnreg, which works on warpgroups of
// four warps (hence the idle sixteenth warp): the two epilogue warpgroups grow to V2_REGS_EPILOGUE, the producer
// warpgroup shrinks to V2_REGS_PRODUCER, the issuer / loader warpgroup to V2_REGS_ISSUER. The CTA then holds
// 512 x 80 = 40 960 of the SM's 65 536 registers instead of 480 x 128 = 61 440, so that a CTA of another kernel
// (the FP32-bound refine of the previous band or frame, 128 threads x 168 registers, 34 KB of shared memory) fits
// beside it and uses the issue slots and the FMA pipe this tensor-pipe-bound kernel leaves idle (cabi.cu, pipeline).
constexpr int V2_THREADS = 512;
template<int K>
constexpr int V2_REGS_LAUNCH = K == 4 ? 80 : 96; // 256-bit descriptors: 32-bit folds, twice the left operand
template<int K>
constexpr int V2_REGS_EPILOGUE = K == 4 ? 120 : 136; // 256 threads
template<int K>
constexpr int V2_REGS_PRODUCER = K == 4 ? 40 : 48; // 128 threads
template<int K>
constexpr int V2_REGS_ISSUER = 40; // 128 threads (K only keeps the three budgets uniform to use)
template<int K>
constexpr bool V2_REGS_FIT = 256 * V2_REGS_EPILOGUE<K> + 128 * V2_REGS_PRODUCER<K> + 128 * V2_REGS_ISSUER<K> <= V2_THREADS * V2_REGS_LAUNCH<K>;
static_assert(V2_REGS_FIT<4> && V2_REGS_FIT<8>, "the warpgroups cannot take more registers than the CTA was launched with");

template<int REGS>
__device__ __forceinline__ void regs_grow() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template<int REGS>
__device__ __forceinline__ void regs_shrink() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}
constexpr uint32_t V2_ACC_COLS = 3 * TN;
template<int K>
constexpr int V2_STAGES = K == 4 ? 8 : 4; // 128 KB of right tiles
constexpr int V2_PACKED = 4;
template<int K>
constexpr int V2_LEFT_BUFFERS = K == 4 ? 2 : 1;

// (1 - 2a) * 2^(7 - s) bytes of one descriptor word, the eight words s = 0..7 = TMEM columns 8 wi + s
template<int K>
__device__ __forceinline__ void expand_left_to_tmem(const uint4 (&d)[K / 4], uint32_t taddr) {
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const uint32_t w[4] = { d[q].x, d[q].y, d[q].z, d[q].w };
        uint32_t v[32];
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            v[8 * wi + 0] = expand_word<false, 0>(w[wi]);
            v[8 * wi + 1] = expand_word<false, 1>(w[wi]);
            v[8 * wi + 2] = expand_word<false, 2>(w[wi]);
            v[8 * wi + 3] = expand_word<false, 3>(w[wi]);
            v[8 * wi + 4] = expand_word<false, 4>(w[wi]);
            v[8 * wi + 5] = expand_word<false, 5>(w[wi]);
            v[8 * wi + 6] = expand_word<false, 6>(w[wi]);
            v[8 * wi + 7] = expand_word<false, 7>(w[wi]);
        }
        tc_store32(taddr + 32u * q, v);
    }
    tc_store_wait();
}

#define V2_ROLE_CONTEXT \
    uint32_t fresh; \
    asm volatile("mov.u32 %0, 0;" : "=r"(fresh)); \
    const uint32_t tmem = *(volatile uint32_t*)&tmem_base_slot + fresh; \
    const int cols = p.cols; \
    const int ntiles = p.ntiles; \
    const int mpairs = p.mtiles; \
    const long long item0 = p.items * (blockIdx.x + fresh) / gridDim.x; \
    const int nitems = (int)(p.items * (blockIdx.x + fresh + 1) / gridDim.x - item0); \
    const uint32_t s_b = ((smem_u32(smem_raw) + fresh) + 1023u) & ~1023u; \
    const uint32_t s_packed = s_b + NS * KA * ATOM_BYTES; \
    const uint32_t bar_stage_full = smem_u32(&bars[0]) + fresh; \
    const uint32_t bar_stage_free = bar_stage_full + 8 * NS; \
    const uint32_t bar_packed_full = bar_stage_free + 8 * NS; \
    const uint32_t bar_packed_free = bar_packed_full + 8 * NP; \
    const uint32_t bar_acc_full = bar_packed_free + 8 * NP; \
    const uint32_t bar_acc_drained = bar_acc_full + 48; \
    const uint32_t bar_left_full = bar_acc_drained + 48; \
    Item it; \
    it.decode(item0, p.rows, mpairs); \
    (void)tmem, (void)cols, (void)ntiles, (void)s_b, (void)s_packed, (void)bar_stage_full, (void)bar_stage_free, (void)bar_packed_full, \
        (void)bar_packed_free, (void)bar_acc_full, (void)bar_acc_drained, (void)bar_left_full;

template<int K, bool NODUPES, bool CT>
__global__ void __maxnreg__(V2_REGS_LAUNCH<K>) search_mma2_kernel(const MmaArgs p) {
    constexpr int KA = K / 4;
    constexpr int NS = V2_STAGES<K>;
    constexpr int NA = V2_LEFT_BUFFERS<K>;
    constexpr int NP = V2_PACKED;
    constexpr int PACKED_BYTES = TN * K * 4;
    constexpr uint32_t LEFT_COLS = 2 * 32 * KA; // both halves of one item
    extern __shared__ uint8_t smem_raw[];
    // stage full [NS], stage free [NS], packed full [NP], packed free [NP], accumulator full [3][2], drained [3][2],
    // left full [NA]. The accumulator barriers are per (accumulator, half): group q = 2 g + h uses accumulator
    // q % 3, so an accumulator alternates between the halves; with one barrier per accumulator each waiter would
    // see only every second phase, and a parity wait two phases ahead of the barrier succeeds at once.
    __shared__ uint64_t bars[2 * NS + 2 * NP + 12 + NA];
    __shared__ uint32_t tmem_base_slot;
    __shared__ Watch s_watch;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;

    if (tid == 0) {
        const uint32_t bar0 = smem_u32(&bars[0]);
        s_watch.flag = p.timeout_flag;
        s_watch.timeout_ns = p.timeout_ns;
        s_watch.abort = 0;
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar0 + 8 * s, TN); // stage full
            mbar_init(bar0 + 8 * (NS + s), 2); // stage free
        }
        for (int s = 0; s < NP; ++s) {
            mbar_init(bar0 + 8 * (2 * NS + s), 1); // packed full
            mbar_init(bar0 + 8 * (2 * NS + NP + s), TN); // packed free
        }
        for (int a = 0; a < 6; ++a) {
            mbar_init(bar0 + 8 * (2 * NS + 2 * NP + a), 1); // accumulator full
            mbar_init(bar0 + 8 * (2 * NS + 2 * NP + 6 + a), TM); // accumulator drained
        }
        for (int b = 0; b < NA; ++b)
            mbar_init(bar0 + 8 * (2 * NS + 2 * NP + 12 + b), 2 * TM); // left full
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // Each role re-balances the registers and only then derives what it needs, inside its own branch: ptxas
    // allocates a region after a join of branches with different setmaxnreg values for the smallest of them, and
    // it keeps a value in ONE register for its whole life, so anything computed before the split and used after
    // it would have to sit in the registers of the smallest role. `fresh` is an opaque zero that keeps the
    // compiler from hoisting the common expressions back above the split.
    if (warp >= 12) {
        regs_shrink<V2_REGS_ISSUER<K>>();
        V2_ROLE_CONTEXT
    if (warp == 15) {
        // the fourth warp of the issuer / loader warpgroup: only there because setmaxnreg works on warpgroups
    } else if (warp == 12 || warp == 13) {
        // ---- MMA issuers: warp 12 + h issues the MMAs of left half h (whole warp, one elected lane).
        //      Two issuers because the issue loop's own latency, not the tensor pipe, bounds a single one. ----
        {
            const int h = warp - 12;
            const uint32_t u_tmem = uniform(tmem);
            const uint64_t desc_b0 = smem_desc(uniform(s_b));
            constexpr uint32_t STAGE_STEP = (KA * ATOM_BYTES) >> 4;
            uint32_t s = 0, stage_phase = 0;
            uint32_t b = 0, left_phase = 0;
            int q = h; // MMA group 2 g + h
            for (int n = 0; n < nitems; ++n) {
                if (!mbar_wait(bar_left_full + 8 * b, left_phase, &s_watch))
                    goto teardown;
                const uint32_t left = u_tmem + V2_ACC_COLS + b * LEFT_COLS + (uint32_t)(h * 32 * KA);
                for (int t = 0; t < ntiles; ++t, q += 2) {
                    const uint32_t a = (uint32_t)q % 3u;
                    if (!mbar_wait(bar_stage_full + 8 * s, stage_phase, &s_watch))
                        goto teardown;
                    if (q >= 3) // the accumulator's previous use, by the other half, has been read
                        if (!mbar_wait(bar_acc_drained + 8 * (2 * a + (1 - h)), (((uint32_t)q - 3u) / 6u) & 1u, &s_watch))
                            goto teardown;
                    tc_fence_after();
                    const uint64_t desc_b = desc_b0 + s * STAGE_STEP;
                    if (elect_one()) {
#pragma unroll
                        for (int qa = 0; qa < KA; ++qa)
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                tc_mma_i8_ts(
                                    u_tmem + a * TN,
                                    left + (uint32_t)(qa * 32 + kk * 8),
                                    desc_b + (uint32_t)((qa * ATOM_BYTES + kk * 32) >> 4),
                                    IDESC,
                                    (qa | kk) != 0
                                );
                        tc_commit(bar_acc_full + 8 * (2 * a + h));
                        tc_commit(bar_stage_free + 8 * s); // counts 2: both halves have read the stage
                    }
                    __syncwarp();
                    if (++s == NS) {
                        s = 0;
                        stage_phase ^= 1;
                    }
                }
                if (++b == NA) {
                    b = 0;
                    left_phase ^= 1;
                }
            }
        }
    } else if (warp == 14) {
        // ---- loader ----
        if (tid == 14 * 32) {
            int f = 0;
            for (int n = 0; n < nitems; ++n) {
                const uint32_t* const row = (it.dir ? p.left : p.right) + (size_t)it.row * p.pitch_words;
                for (int t = 0; t < ntiles; ++t, ++f) {
                    const int s = f % NP;
                    if (f >= NP)
                        if (!mbar_wait(bar_packed_free + 8 * s, (f / NP - 1) & 1, &s_watch))
                            goto teardown;
                    const uint32_t bytes = (uint32_t)min(TN, cols - t * TN) * K * 4;
                    mbar_expect_tx(bar_packed_full + 8 * s, bytes);
                    bulk_copy_g2s(s_packed + (uint32_t)(s * PACKED_BYTES), row + (size_t)t * TN * K, bytes, bar_packed_full + 8 * s);
                }
                it.next(p.rows, mpairs);
            }
        }
    }
    } else if (warp >= 8) {
        // ---- producers ----
        regs_shrink<V2_REGS_PRODUCER<K>>();
        V2_ROLE_CONTEXT
        const int r = tid - 2 * TM;
        const int total = nitems * ntiles;
        int t = 0;
        for (int g = 0; g < total; ++g) {
            const int ps = g % NP, s = g % NS;
            const int valid = min(TN, cols - t * TN);
            if (!mbar_wait(bar_packed_full + 8 * ps, (g / NP) & 1, &s_watch))
                goto teardown;
            uint4 d[KA];
            const uint32_t src = s_packed + (uint32_t)(ps * PACKED_BYTES) + (uint32_t)min(r, valid - 1) * (K * 4);
#pragma unroll
            for (int q = 0; q < KA; ++q)
                d[q] = ld_shared_v4(src + 16 * q);
            if (g >= NS)
                if (!mbar_wait(bar_stage_free + 8 * s, (g / NS - 1) & 1, &s_watch))
                    goto teardown;
            expand_pixel<K, true, CT>(d, s_b + (uint32_t)(s * KA * ATOM_BYTES), r);
            mbar_arrive(bar_packed_free + 8 * ps); // after the stores that consumed the loaded registers
            fence_async_smem();
            mbar_arrive(bar_stage_full + 8 * s);
            if (++t == ntiles)
                t = 0;
        }
    } else {
        // ---- epilogue: half h = warps 4h..4h+3, thread = TMEM lane = left pixel 256 mp + 128 h + lane ----
        regs_grow<V2_REGS_EPILOGUE<K>>();
        V2_ROLE_CONTEXT
        const int h = warp >> 2;
        const int lane128 = tid & (TM - 1);
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        int pa_next = 0;
        auto load_left = [&](const Item& at, uint4(&d)[KA]) {
            const uint32_t* const row = (at.dir ? p.right : p.left) + (size_t)at.row * p.pitch_words;
            load_pixel<K>(row + (size_t)min(at.mt * 2 * TM + h * TM + lane128, cols - 1) * K, d);
        };
        auto stage_left = [&](const uint4(&d)[KA], int n) {
            const int b = n % NA;
            expand_left_to_tmem<K>(d, lane_base + V2_ACC_COLS + (uint32_t)(b * LEFT_COLS + h * 32 * KA));
            tc_fence_before();
            mbar_arrive(bar_left_full + 8 * b);
            int pc = 0;
#pragma unroll
            for (int q = 0; q < KA; ++q)
                pc += __popc(d[q].x) + __popc(d[q].y) + __popc(d[q].z) + __popc(d[q].w);
            return pc;
        };
        uint4 dl[KA];
        if (nitems > 0) {
            load_left(it, dl);
            pa_next = stage_left(dl, 0);
        }
        int g = 0;
        for (int n = 0; n < nitems; ++n) {
            const int pa = pa_next;
            Item nx = it;
            nx.next(p.rows, mpairs);
            int m_first = INT_MAX, m_last = INT_MAX;
            if (n + 1 < nitems)
                load_left(nx, dl);
            for (int t = 0; t < ntiles; ++t, ++g) {
                const int q = 2 * g + h, a = q % 3;
                if (!mbar_wait(bar_acc_full + 8 * (2 * a + h), ((uint32_t)q / 6u) & 1u, &s_watch))
                    goto teardown;
                tc_fence_after();
                const uint32_t acc = lane_base + (uint32_t)(a * TN);
                const int tile0 = t * TN;
                int va[32], vb[32];
                if (tile0 + TN > cols) {
                    TileMin32 m;
#pragma unroll 1
                    for (int u0 = 0; u0 < TN && tile0 + u0 < cols; u0 += 32) {
                        tc_load32_issue(acc + (uint32_t)u0, va);
                        tc_load32_wait(va);
                        fold32_guarded<NODUPES, CT>(va, u0, cols - tile0, m);
                    }
                    tc_fence_before();
                    mbar_arrive(bar_acc_drained + 8 * (2 * a + h));
                    merge_tile<NODUPES>(m, tile0, m_first, m_last);
                } else if constexpr (K == 4) {
                    TileMin16 m;
                    tc_load64_packed_issue(acc, va);
                    tc_load64_packed_issue(acc + 64, vb);
                    tc_load32_wait(va);
                    tc_load32_wait(vb);
                    tc_fence_before();
                    mbar_arrive(bar_acc_drained + 8 * (2 * a + h));
                    fold64_packed<NODUPES, 0, CT>(va, m);
                    fold64_packed<NODUPES, 64, CT>(vb, m);
                    merge_tile<NODUPES>(m, tile0, m_first, m_last);
                } else {
                    TileMin32 m;
                    tc_load32_issue(acc, va);
                    tc_load32_wait(va);
                    tc_load32_issue(acc + 32, vb);
                    fold32<NODUPES, 0, CT>(va, m);
                    tc_load32_wait(vb);
                    tc_load32_issue(acc + 64, va);
                    fold32<NODUPES, 32, CT>(vb, m);
                    tc_load32_wait(va);
                    tc_load32_issue(acc + 96, vb);
                    fold32<NODUPES, 64, CT>(va, m);
                    tc_load32_wait(vb);
                    tc_fence_before();
                    mbar_arrive(bar_acc_drained + 8 * (2 * a + h));
                    fold32<NODUPES, 96, CT>(vb, m);
                    merge_tile<NODUPES>(m, tile0, m_first, m_last);
                }
                // next item's left half: its TMEM columns are read only by this half's MMAs, all complete once
                // this half's accumulator of the buffer's previous user (two buffers: the previous item, seen
                // long ago; one: this item's last tile, just seen) has been committed
                if (n + 1 < nitems && t == (NA == 2 ? min(2, ntiles - 1) : ntiles - 1))
                    pa_next = stage_left(dl, n + 1);
            }
            const int i = it.mt * 2 * TM + h * TM + lane128;
            if (i < cols) {
                const size_t at = (size_t)it.row * cols + i;
                (it.dir ? p.rev_first : p.fwd_first)[at] = ((uint32_t)(pa + (m_first >> COL_BITS)) << 16) | ((uint32_t)m_first & COL_MAX);
                if constexpr (NODUPES)
                    (it.dir ? p.rev_last : p.fwd_last)[at] =
                        ((uint32_t)(pa + (m_last >> COL_BITS)) << 16) | ((65535u - COL_MAX) + ((uint32_t)m_last & COL_MAX));
            }
            it = nx;
        }
    }

teardown:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*(volatile uint32_t*)&tmem_base_slot), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// One-pass consistency search (128- and 256-bit descriptors, Consistency without no_dupes: the metric configuration).
// The two kernels above compute the W x W Hamming matrix of a row twice, once per direction, because a
// minimum ACROSS the TMEM lanes of one accumulator costs a cross-lane reduction per column. Here every pair
// is computed once (SURVEY 8d's count) and both minima are taken from the same accumulator:
//   work item   = (row, pair of blocks of 128 RIGHT pixels): each block is a B operand, expanded once per item
//                 into shared memory (signed bytes); the LEFT row streams past them as A operands, 128 pixels
//                 per tile, expanded by the producer warps straight into tensor memory (tcgen05.st; unsigned
//                 bytes, one LOP3 per four bytes), so an MMA reads only its 16 KB block from shared memory and
//                 every expanded tile serves two MMA groups (one per block)
//   accumulator = D[lane = left pixel of the tile][column = right pixel of the block]
//               = 128 * ham + column: both operands' top two descriptor bits are unused (4n-6 <= 32K-2,
//                 n^2-2n+3 mod 32 <= 27), and their operand bytes carry 1 x column (as the CT kernels) and
//                 128 x popc(right descriptor), which cancels the -128 popc of the signed operand
//   forward     (per left pixel, minimum over the right row): in-thread fold over the 128 columns
//                 (VIMNMX3.S16x2, four columns per instruction), merged over the blocks of the row by
//                 atomicMin on the key array (pre-filled by the launcher)
//   reverse     (per right pixel, minimum over the left row): ELEMENTWISE across the tiles of the item:
//                 R[column] = min(R[column], D + tile) with the tile index as the tie-breaker
//                 (VIADDMNMX.S16x2, two columns per instruction, no cross-lane traffic while the row streams);
//                 one cross-lane reduction per ITEM (through shared memory, with the lane as the last
//                 tie-breaker) instead of one per tile
// A repeated last pixel pads ragged tiles and blocks (and a missing second block): it ties with the real pixel
// and loses on the index.
//   warps 0-3 / 4-7  epilogue of block 0 / 1 (the four TMEM lane quadrants each)
//   warps 8-11       producers      warps 12, 13  MMA issuers (block 0 / 1)      warp 14  loader      warp 15  idle
// TMEM: three 128-column accumulators in rotation (columns 0..383, MMA group 2 g + h -> accumulator (2 g + h) % 3
// as in search_mma2_kernel), four A tiles of 32 columns (384..511).
// 256-bit descriptors (K = 8) run the same kernel: two 128-bit atoms per operand, two A tiles of 64 TMEM columns, and,
// because popc(right descriptor) no longer fits the signed operand byte, that byte carries popc - 128 and the
// accumulator is 128 ham + column - 16384 (still a 16-bit value: ham <= 254); the reverse fold adds the 16384 back
// together with the tile index, the forward fold's minimum adds it when the key is built.
constexpr int V3_THREADS = 512;
template<int K>
constexpr int V3_REGS_LAUNCH = K == 4 ? 104 : 112;
constexpr int V3_REGS_EPILOGUE = 168; // 64 running minima + 64 accumulator registers
template<int K>
constexpr int V3_REGS_PRODUCER = K == 4 ? 40 : 56;
template<int K>
constexpr int V3_REGS_ISSUER = K == 4 ? 40 : 56;
template<int K>
constexpr bool V3_REGS_FIT = 256 * V3_REGS_EPILOGUE + 128 * V3_REGS_PRODUCER<K> + 128 * V3_REGS_ISSUER<K> <= V3_THREADS * V3_REGS_LAUNCH<K>;
static_assert(V3_REGS_FIT<4> && V3_REGS_FIT<8>, "register budget");
template<int K>
constexpr int V3_ASLOTS = K == 4 ? 4 : 2; // A tiles in tensor memory: 128 columns beside the three accumulators
constexpr int V3_PACKED = 4;
constexpr uint32_t V3_A_COL0 = 3 * TN;
template<int K>
constexpr int V3_PACKED_BYTES = TN * K * 4;
template<int K>
constexpr int V3_OFFSET = K == 4 ? 0 : 16384; // 128 x the bias of the popcount byte
constexpr int V3_STATE_STRIDE = 132; // words per column-pair row of 128 lanes: conflict-free LDS.128
constexpr int V3_STATE_BYTES = 64 * V3_STATE_STRIDE * 4; // per epilogue group
constexpr int V3_FIN_BYTES = 2 * TN * 4; // per epilogue group
template<int K>
constexpr int V3_SMEM_BYTES = 4 * (K / 4) * ATOM_BYTES + V3_PACKED * V3_PACKED_BYTES<K> + 2 * V3_STATE_BYTES + 2 * V3_FIN_BYTES + 1024;
static_assert(V3_SMEM_BYTES<8> <= 227 * 1024, "shared memory");
// instruction descriptor: D = s32, A = unsigned int8 (the streamed left tile), B = signed int8 (the block)
constexpr uint32_t IDESC3 = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void tc_store8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// a value the compiler must keep in a register of its own (not re-derive from %tid inside a loop)
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// One left pixel -> its TMEM lane of A slot `taddr`: b * 2^s bytes, TMEM column 8 wi + s = bytes of bits s, 8 + s,
// 16 + s, 24 + s of word wi. The bytes of the two unused top bits (word 3, byte 3, s = 6 / 7) are 128 and 1.
template<int K>
__device__ __forceinline__ void expand_moving_to_tmem(const uint4 (&d)[K / 4], uint32_t taddr) {
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
    const uint32_t w[4] = { d[q].x, d[q].y, d[q].z, d[q].w };
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
        uint32_t v[8];
        v[0] = expand_word<true, 0>(w[wi]);
        v[1] = expand_word<true, 1>(w[wi]);
        v[2] = expand_word<true, 2>(w[wi]);
        v[3] = expand_word<true, 3>(w[wi]);
        v[4] = expand_word<true, 4>(w[wi]);
        v[5] = expand_word<true, 5>(w[wi]);
        v[6] = expand_word<true, 6>(w[wi]);
        v[7] = expand_word<true, 7>(w[wi]);
        if (q == K / 4 - 1 && wi == 3) {
            v[6] |= 128u << 24;
            v[7] |= 1u << 24;
        }
        tc_store8(taddr + 32u * q + 8u * wi, v);
    }
    }
    tc_store_wait();
}

// One right pixel -> row r of the block tile in shared memory: (1 - 2b) * 2^(7 - s) bytes (128B-swizzled K-major, as
// expand_pixel). The bytes of the two unused top bits carry popc(descriptor) (against the 128 of the left operand)
// and the block column r (against the 1).
template<int K>
__device__ __forceinline__ void expand_block_pixel(const uint4 (&d)[K / 4], uint32_t tile, int r) {
    const uint32_t row = tile + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    int pc = 0;
#pragma unroll
    for (int q = 0; q < K / 4; ++q)
        pc += __popc(d[q].x) + __popc(d[q].y) + __popc(d[q].z) + __popc(d[q].w);
    // 256 bits: up to 254 set bits, so the signed byte carries popc - 128 (see V3_OFFSET)
    const uint32_t pc_byte = (uint32_t)(pc - (K == 4 ? 0 : 128)) & 0xFFu;
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const uint32_t w[4] = { d[q].x, d[q].y, d[q].z, d[q].w };
        const uint32_t atom = row + (uint32_t)q * ATOM_BYTES;
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            st_shared_v4(atom + ((((uint32_t)(2 * wi)) ^ sw) << 4), expand_word<false, 0>(w[wi]), expand_word<false, 1>(w[wi]),
                         expand_word<false, 2>(w[wi]), expand_word<false, 3>(w[wi]));
            uint32_t s6 = expand_word<false, 6>(w[wi]), s7 = expand_word<false, 7>(w[wi]);
            if (q == K / 4 - 1 && wi == 3) {
                s6 = (s6 & 0x00FFFFFFu) | (pc_byte << 24);
                s7 = (s7 & 0x00FFFFFFu) | ((uint32_t)r << 24);
            }
            st_shared_v4(atom + ((((uint32_t)(2 * wi + 1)) ^ sw) << 4), expand_word<false, 4>(w[wi]), expand_word<false, 5>(w[wi]), s6, s7);
        }
    }
}

// The per-item cross-lane reduction, one thread's share: 64 lanes of one column pair's row of the dumped state ->
// the minima of (16-bit running minimum) << 16 | lane for the even (lo half) and the odd column. The odd column's key
// is one LOP3, (w & mask) | lane with the mask in a register; the even column's one IMAD on the otherwise idle FMA
// pipe, w * 65536 + lane with the factor in a register (ptxas would turn a literal shift into an ALU instruction).
template<int L>
__device__ __forceinline__ uint32_t key_of_hi(uint32_t w, uint32_t mask) {
    uint32_t k;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(k) : "r"(w), "r"(mask), "n"(L));
    return k;
}
template<int L4>
struct ReduceLanes {
    static __device__ __forceinline__ void run(uint32_t src, uint32_t mask, uint32_t shl16, uint32_t& mlo, uint32_t& mhi) {
        const uint4 w = ld_shared_v4(src + 16u * L4);
        mhi = min(mhi, min(key_of_hi<4 * L4>(w.x, mask), key_of_hi<4 * L4 + 1>(w.y, mask)));
        mlo = min(mlo, min(w.x * shl16 + (uint32_t)(4 * L4), w.y * shl16 + (uint32_t)(4 * L4 + 1)));
        mhi = min(mhi, min(key_of_hi<4 * L4 + 2>(w.z, mask), key_of_hi<4 * L4 + 3>(w.w, mask)));
        mlo = min(mlo, min(w.z * shl16 + (uint32_t)(4 * L4 + 2), w.w * shl16 + (uint32_t)(4 * L4 + 3)));
        ReduceLanes<L4 + 1>::run(src, mask, shl16, mlo, mhi);
    }
};
template<>
struct ReduceLanes<16> {
    static __device__ __forceinline__ void run(uint32_t, uint32_t, uint32_t, uint32_t&, uint32_t&) {}
};

#define V3_ROLE_CONTEXT \
    uint32_t fresh; \
    asm volatile("mov.u32 %0, 0;" : "=r"(fresh)); \
    const uint32_t tmem = *(volatile uint32_t*)&tmem_base_slot + fresh; \
    const int cols = p.cols; \
    const int ntiles = p.ntiles; /* left tiles per item */ \
    const int npairs = p.mtiles; /* pairs of right blocks per row */ \
    const long long item0 = p.items * (blockIdx.x + fresh) / gridDim.x; \
    const int nitems = (int)(p.items * (blockIdx.x + fresh + 1) / gridDim.x - item0); \
    const uint32_t s_blk = ((smem_u32(smem_raw) + fresh) + 1023u) & ~1023u; /* + (2 buffer + block) * BLK_BYTES */ \
    const uint32_t s_packed = s_blk + 4 * BLK_BYTES; \
    const uint32_t s_state = s_packed + NP * PACKED_BYTES; /* + group * V3_STATE_BYTES */ \
    const uint32_t s_fin = s_state + 2 * V3_STATE_BYTES; /* + group * V3_FIN_BYTES */ \
    const uint32_t bar_packed_full = smem_u32(&bars[0]) + fresh; \
    const uint32_t bar_packed_free = bar_packed_full + 8 * NP; \
    const uint32_t bar_a_full = bar_packed_free + 8 * NP; \
    const uint32_t bar_a_free = bar_a_full + 8 * NS; \
    const uint32_t bar_blk_full = bar_a_free + 8 * NS; /* + 8 * (2 buffer + block) */ \
    const uint32_t bar_blk_free = bar_blk_full + 32; \
    const uint32_t bar_acc_full = bar_blk_free + 32; \
    const uint32_t bar_acc_drained = bar_acc_full + 48; \
    const uint32_t bar_epi = bar_acc_drained + 48; /* + 8 * group */ \
    int row = (int)(item0 / npairs), bp = (int)(item0 - (long long)row * npairs); \
    (void)tmem, (void)cols, (void)ntiles, (void)s_blk, (void)s_packed, (void)s_state, (void)s_fin, (void)bar_packed_full, (void)bar_packed_free, \
        (void)bar_a_full, (void)bar_a_free, (void)bar_blk_full, (void)bar_blk_free, (void)bar_acc_full, (void)bar_acc_drained, (void)bar_epi, \
        (void)row, (void)bp;

template<int K>
__global__ void __maxnreg__(V3_REGS_LAUNCH<K>) search_mma3_kernel(const MmaArgs p) {
    constexpr int KA = K / 4; // 128-bit atoms per descriptor
    constexpr int NS = V3_ASLOTS<K>;
    constexpr int NP = V3_PACKED;
    constexpr int PACKED_BYTES = V3_PACKED_BYTES<K>;
    constexpr int BLK_BYTES = KA * ATOM_BYTES;
    constexpr uint32_t A_COLS = 32u * KA; // TMEM columns of one A tile
    constexpr int OFF = V3_OFFSET<K>;
    extern __shared__ uint8_t smem_raw[];
    // packed full [NP], packed free [NP], A full [NS], A free [NS], block full [2][2], block free [2][2],
    // accumulator full [3][2], drained [3][2] (per accumulator and block, see search_mma2_kernel), epilogue sync [2]
    __shared__ uint64_t bars[2 * NP + 2 * NS + 8 + 12 + 2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ Watch s_watch;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;

    if (tid == 0) {
        const uint32_t bar0 = smem_u32(&bars[0]);
        s_watch.flag = p.timeout_flag;
        s_watch.timeout_ns = p.timeout_ns;
        s_watch.abort = 0;
        for (int s = 0; s < NP; ++s) {
            mbar_init(bar0 + 8 * s, 1); // packed full
            mbar_init(bar0 + 8 * (NP + s), TM); // packed free
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar0 + 8 * (2 * NP + s), TM); // A full
            mbar_init(bar0 + 8 * (2 * NP + NS + s), 2); // A free: the MMAs of both blocks
        }
        for (int b = 0; b < 4; ++b) {
            mbar_init(bar0 + 8 * (2 * NP + 2 * NS + b), TN); // block full
            mbar_init(bar0 + 8 * (2 * NP + 2 * NS + 4 + b), 1); // block free
        }
        for (int a = 0; a < 6; ++a) {
            mbar_init(bar0 + 8 * (2 * NP + 2 * NS + 8 + a), 1); // accumulator full
            mbar_init(bar0 + 8 * (2 * NP + 2 * NS + 14 + a), TM); // accumulator drained
        }
        mbar_init(bar0 + 8 * (2 * NP + 2 * NS + 20), TM); // epilogue sync, block 0
        mbar_init(bar0 + 8 * (2 * NP + 2 * NS + 21), TM); // block 1
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= 12) {
        regs_shrink<V3_REGS_ISSUER<K>>();
        V3_ROLE_CONTEXT
    if (warp == 15) {
        // idle: setmaxnreg works on warpgroups
    } else if (warp == 12 || warp == 13) {
        // ---- MMA issuers: warp 12 + h issues every tile against block h (MMA group 2 g + h) ----
        {
            const int h = warp - 12;
            const uint32_t u_tmem = uniform(tmem);
            const uint64_t desc_b0 = smem_desc(uniform(s_blk)) + (uint32_t)h * (uint32_t)(BLK_BYTES >> 4);
            constexpr uint32_t BUFFER_STEP = (2 * BLK_BYTES) >> 4;
            uint32_t s = 0, a_phase = 0;
            // MMA group q = 2 g + h: accumulator a = q % 3; its previous use was group q - 3 of the other block, whose
            // "drained" barrier completes phase ((q - 3) / 6) & 1: both carried incrementally (d6 = (q - 3) % 6)
            uint32_t a = (uint32_t)h, d6 = (uint32_t)h + 3u, d_phase = 1u; // q = h: "q - 3 < 0", nothing to wait for until q >= 3
            int q = h;
            for (int n = 0; n < nitems; ++n) {
                const uint32_t b = (uint32_t)n & 1u;
                if (!mbar_wait(bar_blk_full + 8 * (2 * b + h), ((uint32_t)n >> 1) & 1u, &s_watch))
                    goto teardown;
                const uint64_t desc_b = desc_b0 + b * BUFFER_STEP;
                for (int t = 0; t < ntiles; ++t, q += 2) {
                    if (!mbar_wait(bar_a_full + 8 * s, a_phase, &s_watch))
                        goto teardown;
                    if (q >= 3) // the accumulator's previous use, by the other block, has been read
                        if (!mbar_wait(bar_acc_drained + 8 * (2 * a + (1 - h)), d_phase, &s_watch))
                            goto teardown;
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int qa = 0; qa < KA; ++qa)
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                tc_mma_i8_ts(u_tmem + a * TN, u_tmem + V3_A_COL0 + s * A_COLS + (uint32_t)(qa * 32 + kk * 8),
                                             desc_b + (uint32_t)((qa * ATOM_BYTES + kk * 32) >> 4), IDESC3, (qa | kk) != 0);
                        tc_commit(bar_acc_full + 8 * (2 * a + h));
                        tc_commit(bar_a_free + 8 * s); // counts 2: both blocks have read the tile
                    }
                    __syncwarp();
                    if (++s == NS) {
                        s = 0;
                        a_phase ^= 1u;
                    }
                    a = a == 0 ? 2u : a - 1u; // (a + 2) % 3
                    d6 += 2u;
                    if (d6 >= 6u) {
                        d6 -= 6u;
                        d_phase ^= 1u;
                    }
                }
                if (elect_one())
                    tc_commit(bar_blk_free + 8 * (2 * b + h)); // every MMA on this block is complete when it arrives
                __syncwarp();
            }
        }
    } else if (warp == 14) {
        // ---- loader: per item the two packed right blocks, then the packed left tiles of the row ----
        if (tid == 14 * 32) {
            // ring slot and the parity its "free" barrier completes with are carried incrementally; the first NP
            // entries find their slots free
            uint32_t slot = 0, free_phase = 1u;
            bool wrapped = false;
            const uint32_t tail_bytes = (uint32_t)(cols - (ntiles - 1) * TN) * (uint32_t)(K * 4); // the row's last (possibly short) tile
            auto put = [&](const uint32_t* src, uint32_t bytes) -> bool {
                if (wrapped)
                    if (!mbar_wait(bar_packed_free + 8 * slot, free_phase, &s_watch))
                        return false;
                mbar_expect_tx(bar_packed_full + 8 * slot, bytes);
                bulk_copy_g2s(s_packed + slot * (uint32_t)PACKED_BYTES, src, bytes, bar_packed_full + 8 * slot);
                if (++slot == (uint32_t)NP) {
                    slot = 0;
                    free_phase ^= 1u;
                    wrapped = true;
                }
                return true;
            };
            for (int n = 0; n < nitems; ++n) {
                const uint32_t* const lrow = p.left + (size_t)row * p.pitch_words;
                const uint32_t* const rrow = p.right + (size_t)row * p.pitch_words;
                for (int e = 0; e < 2; ++e) {
                    // a second block beyond the row: one copy of the row's last pixel, which the producers replicate
                    const int base = min((2 * bp + e) * TN, cols - 1);
                    if (!put(rrow + (size_t)base * K, (uint32_t)min(TN, cols - base) * (uint32_t)(K * 4)))
                        goto teardown;
                }
                for (int t = 0; t < ntiles - 1; ++t)
                    if (!put(lrow + (size_t)t * (TN * K), (uint32_t)PACKED_BYTES))
                        goto teardown;
                if (!put(lrow + (size_t)(ntiles - 1) * (TN * K), tail_bytes))
                    goto teardown;
                if (++bp == npairs) {
                    bp = 0;
                    ++row;
                }
            }
        }
    }
    } else if (warp >= 8) {
        // ---- producers: ring entries 0, 1 of an item = the right blocks -> shared memory (signed), entries 2.. = left
        //      tiles -> tensor memory (unsigned). Thread r = row r of a block = TMEM lane r of a tile. ----
        regs_shrink<V3_REGS_PRODUCER<K>>();
        V3_ROLE_CONTEXT
        const int r = (int)pinned((uint32_t)(tid - 2 * TM));
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        // everything a ring entry needs is carried incrementally: packed slot + parity, A slot + parity
        uint32_t ps = 0, full_phase = 0, as = 0, a_free_phase = 1u;
        bool a_wrapped = false;
        const uint32_t my_px = s_packed + (uint32_t)r * (uint32_t)(K * 4); // this thread's pixel of a full tile, + slot * PACKED_BYTES
        const uint32_t my_tail_px = s_packed + (uint32_t)min(r, cols - (ntiles - 1) * TN - 1) * (uint32_t)(K * 4); // of the row's last tile
        for (int n = 0; n < nitems; ++n) {
            for (int e = 0; e < 2; ++e) {
                const int valid = min(TN, cols - min((2 * bp + e) * TN, cols - 1));
                if (!mbar_wait(bar_packed_full + 8 * ps, full_phase, &s_watch))
                    goto teardown;
                uint4 d[KA];
                {
                    const uint32_t src = s_packed + ps * (uint32_t)PACKED_BYTES + (uint32_t)min(r, valid - 1) * (uint32_t)(K * 4);
#pragma unroll
                    for (int q = 0; q < KA; ++q)
                        d[q] = ld_shared_v4(src + 16u * q);
                }
                const uint32_t b = 2u * ((uint32_t)n & 1u) + (uint32_t)e;
                if (n >= 2) // the MMAs of the item that used this buffer are complete
                    if (!mbar_wait(bar_blk_free + 8 * b, (((uint32_t)n >> 1) - 1u) & 1u, &s_watch))
                        goto teardown;
                expand_block_pixel<K>(d, s_blk + b * BLK_BYTES, r);
                mbar_arrive(bar_packed_free + 8 * ps); // after the stores that consumed the loaded registers
                fence_async_smem();
                mbar_arrive(bar_blk_full + 8 * b);
                if (++ps == (uint32_t)NP) {
                    ps = 0;
                    full_phase ^= 1u;
                }
            }
            for (int t = 0; t < ntiles; ++t) {
                if (!mbar_wait(bar_packed_full + 8 * ps, full_phase, &s_watch))
                    goto teardown;
                uint4 d[KA];
                {
                    const uint32_t src = (t == ntiles - 1 ? my_tail_px : my_px) + ps * (uint32_t)PACKED_BYTES;
#pragma unroll
                    for (int q = 0; q < KA; ++q)
                        d[q] = ld_shared_v4(src + 16u * q);
                }
                if (a_wrapped)
                    if (!mbar_wait(bar_a_free + 8 * as, a_free_phase, &s_watch))
                        goto teardown;
                tc_fence_after();
                expand_moving_to_tmem<K>(d, lane_base + V3_A_COL0 + as * A_COLS);
                mbar_arrive(bar_packed_free + 8 * ps);
                tc_fence_before();
                mbar_arrive(bar_a_full + 8 * as);
                if (++ps == (uint32_t)NP) {
                    ps = 0;
                    full_phase ^= 1u;
                }
                if (++as == (uint32_t)NS) {
                    as = 0;
                    a_free_phase ^= 1u;
                    a_wrapped = true;
                }
            }
            if (++bp == npairs) {
                bp = 0;
                ++row;
            }
        }
    } else {
        // ---- epilogue: group h = warps 4h..4h+3 owns block h of the item; thread = TMEM lane = left pixel of the tile ----
        regs_grow<V3_REGS_EPILOGUE>();
        V3_ROLE_CONTEXT
        const int h = (int)pinned((uint32_t)(warp >> 2));
        const int lane128 = (int)pinned((uint32_t)(tid & (TM - 1)));
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t my_state = s_state + (uint32_t)(h * V3_STATE_BYTES);
        const uint32_t my_fin = s_fin + (uint32_t)(h * V3_FIN_BYTES);
        const uint32_t my_epi = bar_epi + 8u * (uint32_t)h;
        uint32_t R[64];
#pragma unroll
        for (int c = 0; c < 64; ++c)
            R[c] = 0x7FFF7FFFu;
        uint32_t epi_phase = 0;
        // MMA group q = 2 g + h of this block: accumulator q % 3, barrier phase (q / 6) & 1, both carried incrementally.
        // The accumulator is read in two halves of 64 columns, software-pipelined across tiles so that a TMEM load is
        // in flight during every fold: [lo(t) arrives] load hi(t) | fold lo(t) | [hi(t) arrives, accumulator handed
        // back] wait for tile t + 1, load lo(t + 1) | fold hi(t).
        uint32_t a = (uint32_t)h, acc_phase = 0;
        int lo[32], hi[32];
        int left = nitems * ntiles; // tiles still to load
        if (left > 0) {
            if (!mbar_wait(bar_acc_full + 8 * (2 * a + h), acc_phase, &s_watch))
                goto teardown;
            tc_fence_after();
            tc_load64_packed_issue(lane_base + a * TN, lo);
            --left;
        }
        for (int n = 0; n < nitems; ++n) {
            const size_t row_at = (size_t)row * cols;
            const int col0 = (2 * bp + h) * TN; // first right pixel of my block (may lie beyond the row: nothing is stored then)
            const int fwd_cols = col0 < cols ? cols : 0; // left pixels whose forward key this block may lower
            for (int t = 0; t < ntiles; ++t) {
                const uint32_t tile2 = (uint32_t)(t + OFF) * 0x00010001u; // + the popcount byte's bias (256 bits): the running minima are >= 0
                uint32_t f[8];
                tc_load32_wait(lo);
                tc_load64_packed_issue(lane_base + a * TN + 64u, hi);
                // forward (this left pixel's minimum of 128 ham + column over the block) and reverse (per column the
                // minimum of 128 ham + column + tile over the tiles), first 64 columns
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    f[c] = __vimin3_s16x2((uint32_t)lo[c], (uint32_t)lo[8 + c], (uint32_t)lo[16 + c]);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    f[c] = __vimin3_s16x2(f[c], (uint32_t)lo[24 + 2 * c], (uint32_t)lo[25 + 2 * c]);
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    R[c] = __viaddmin_s16x2((uint32_t)lo[c], tile2, R[c]);
                tc_load32_wait(hi);
                tc_fence_before();
                mbar_arrive(bar_acc_drained + 8 * (2 * a + h));
                // next tile of this block (the next item's first tile after the last one)
                a = a == 0 ? 2u : a - 1u; // (a + 2) % 3
                if (a == (uint32_t)h) // q = 2 g + h passes a multiple of 6 exactly when its accumulator index returns to h
                    acc_phase ^= 1u;
                if (left > 0) {
                    if (!mbar_wait(bar_acc_full + 8 * (2 * a + h), acc_phase, &s_watch))
                        goto teardown;
                    tc_fence_after();
                    tc_load64_packed_issue(lane_base + a * TN, lo);
                    --left;
                }
                // last 64 columns
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    f[c] = __vimin3_s16x2(f[c], (uint32_t)hi[c], (uint32_t)hi[8 + c]);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    f[c] = __vimin3_s16x2(f[c], (uint32_t)hi[16 + c], (uint32_t)hi[24 + c]);
                const uint32_t m2 = __vimin3_s16x2(__vimin3_s16x2(f[0], f[1], f[2]), __vimin3_s16x2(f[3], f[4], f[5]), __vmins2(f[6], f[7]));
                // m = 128 ham + column (>= 0 once the bias is added back); key = ham << 16 | col0 + column = 512 m - 511 (m & 127) + col0
                const uint32_t m = OFF ? (uint32_t)(min((int)(short)(m2 & 0xFFFFu), (int)(short)(m2 >> 16)) + OFF) : min(m2 & 0xFFFFu, m2 >> 16);
                const int i = t * TM + lane128;
                if (i < fwd_cols)
                    atomicMin(p.fwd_first + row_at + i, m * 512u + (uint32_t)col0 - (m & 127u) * 511u);
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    R[32 + c] = __viaddmin_s16x2((uint32_t)hi[c], tile2, R[32 + c]);
            }
            // ---- end of the item: the cross-lane reduction of the running minima ----
            {
                const uint32_t dst = my_state + (uint32_t)(lane128 * 4);
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    st_shared_u32(dst + (uint32_t)(c * V3_STATE_STRIDE * 4), R[c]);
                    R[c] = 0x7FFF7FFFu;
                }
                mbar_arrive(my_epi);
                if (!mbar_wait(my_epi, epi_phase, &s_watch))
                    goto teardown;
                epi_phase ^= 1u;
                // thread = (column pair cp, half lh of the lanes): 64 lanes of that pair's row
                const int cp = lane128 & 63, lh = lane128 >> 6;
                const uint32_t src = my_state + (uint32_t)((cp * V3_STATE_STRIDE + lh * 64) * 4);
                uint32_t mlo = 0xFFFFFFFFu, mhi = 0xFFFFFFFFu;
                ReduceLanes<0>::run(src, pinned(0xFFFF0000u), pinned(65536u), mlo, mhi);
                const uint32_t lane0 = (uint32_t)lh * 64u;
                st_shared_v2(my_fin + (uint32_t)((lh * TN + 2 * cp) * 4), mlo + lane0, mhi + lane0);
                mbar_arrive(my_epi);
                if (!mbar_wait(my_epi, epi_phase, &s_watch))
                    goto teardown;
                epi_phase ^= 1u;
                {
                    const int col = col0 + lane128;
                    const uint32_t k = min(ld_shared_u32(my_fin + (uint32_t)(lane128 * 4)), ld_shared_u32(my_fin + (uint32_t)((TN + lane128) * 4)));
                    // k = (128 ham + column + tile) << 16 | lane
                    const uint32_t vv = (k >> 16) - (uint32_t)lane128;
                    if (col < cols)
                        p.rev_first[row_at + col] = ((vv >> 7) << 16) | ((vv & 127u) * TM + (k & 0xFFFFu));
                }
            }
            if (++bp == npairs) {
                bp = 0;
                ++row;
            }
        }
    }

teardown:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*(volatile uint32_t*)&tmem_base_slot), "r"(512u) : "memory");
    }
}

int sm_count_of_current_device() {
    static thread_local int cached_dev = -1, cached_sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != cached_dev) {
        if (cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            cached_sms = 148;
        cached_dev = dev;
    }
    return cached_sms;
}

// Function attributes are per device and cost a driver call each: set once per (kernel, device, thread).
cudaError_t configure_once(void (*kernel)(const MmaArgs), int smem) {
    struct Entry {
        void (*kernel)(const MmaArgs);
        unsigned long long devices; // one bit per device ordinal
    };
    static thread_local Entry cache[16] = {};
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess)
        return err;
    Entry* e = nullptr;
    for (Entry& c: cache)
        if (c.kernel == kernel || c.kernel == nullptr) {
            e = &c;
            break;
        }
    if (e && e->kernel == kernel && dev < 64 && ((e->devices >> dev) & 1ull))
        return cudaSuccess;
    if ((err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
        return err;
    if ((err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess)
        return err;
    if (e && dev < 64) {
        e->kernel = kernel;
        e->devices |= 1ull << dev;
    }
    return cudaSuccess;
}

template<int K>
constexpr int v2_smem_bytes() {
    return V2_STAGES<K> * (K / 4) * ATOM_BYTES + V2_PACKED * TN * K * 4 + 1024;
}

template<int K, bool NODUPES, bool CT>
cudaError_t launch_k2(MmaArgs p, int dirs, cudaStream_t stream) {
    constexpr int smem = v2_smem_bytes<K>();
    auto kernel = search_mma2_kernel<K, NODUPES, CT>;
    cudaError_t err = configure_once(kernel, smem);
    if (err != cudaSuccess)
        return err;
    const int sms = sm_count_of_current_device();
    p.mtiles = (p.cols + 2 * TM - 1) / (2 * TM); // M pairs
    p.items = (long long)dirs * p.rows * p.mtiles;
    if (p.items > 0x7FFFFFFFLL)
        return cudaErrorInvalidConfiguration;
    const unsigned grid = (unsigned)(p.items < sms ? p.items : sms);
    kernel<<<grid, V2_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template<int K, bool NODUPES, bool CT>
cudaError_t launch_k(MmaArgs p, int dirs, cudaStream_t stream) {
    const int smem = search_mma_smem_bytes(K);
    auto kernel = search_mma_kernel<K, NODUPES, CT>;
    cudaError_t err = configure_once(kernel, smem);
    if (err != cudaSuccess)
        return err;
    // persistent CTAs: as many as are resident at once (__launch_bounds__ and the shared-memory budget
    // of search_mma_smem_bytes are laid out for 2 per SM up to 256 bits, 1 beyond; the occupancy API
    // reports 1 for the 2-CTA variants until the carve-out is raised, so it is not asked), each walking a
    // contiguous share of the items
    const int resident = sm_count_of_current_device() * (K <= 8 ? 2 : 1);
    p.items = (long long)dirs * p.rows * p.mtiles;
    if (p.items > 0x7FFFFFFFLL)
        return cudaErrorInvalidConfiguration;
    const unsigned grid = (unsigned)(p.items < resident ? p.items : resident);
    kernel<<<grid, NTHREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template<int K>
cudaError_t launch_k3(MmaArgs p, cudaStream_t stream) {
    auto kernel = search_mma3_kernel<K>;
    constexpr int smem = V3_SMEM_BYTES<K>;
    cudaError_t err = configure_once(kernel, smem);
    if (err != cudaSuccess)
        return err;
    p.mtiles = (p.cols + 2 * TN - 1) / (2 * TN); // pairs of right blocks per row
    p.items = (long long)p.rows * p.mtiles; // (row, pair of blocks of 128 right pixels)
    if (p.items > 0x7FFFFFFFLL)
        return cudaErrorInvalidConfiguration;
    // the forward keys of a row are merged over its blocks with atomicMin
    if ((err = cudaMemsetAsync(p.fwd_first, 0xFF, (size_t)p.rows * p.cols * sizeof(uint32_t), stream)) != cudaSuccess)
        return err;
    const int sms = sm_count_of_current_device();
    const unsigned grid = (unsigned)(p.items < sms ? p.items : sms);
    kernel<<<grid, V3_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

} // namespace

int search_mma_smem_bytes(int K) {
    const int stages = K == 4 ? STAGES<4> : K == 8 ? STAGES<8> : K == 12 ? STAGES<12> : STAGES<16>;
    const int lefts = K == 4 ? LEFT_BUFFERS<4> : 1;
    const int packed = K == 4 ? PACKED_STAGES<4> : K == 8 ? PACKED_STAGES<8> : PACKED_STAGES<16>;
    return (lefts + stages) * (K / 4) * ATOM_BYTES + packed * TN * K * 4 + 1024; // left tiles, right stages, packed ring, alignment slack
}

namespace {
std::atomic<int> g_variant { -1 };
std::atomic<int> g_colterm { -1 };
std::atomic<long long> g_timeout_ms { -1 };
std::atomic<unsigned int*> g_timeout_flag { nullptr };
std::mutex g_timeout_mutex;

// One word of mapped, portable host memory that a timed-out wait of either kernel raises (mbar_wait_slow):
// the host can read it at any time without touching the device. Null if the allocation fails (the kernels
// then only abort, and the corrupt keys go unreported -- better than refusing to search).
unsigned int* mma_timeout_flag() {
    unsigned int* f = g_timeout_flag.load(std::memory_order_acquire);
    if (f)
        return f;
    std::lock_guard<std::mutex> lk(g_timeout_mutex);
    f = g_timeout_flag.load(std::memory_order_acquire);
    if (!f) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, 64, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        memset(p, 0, 64);
        f = static_cast<unsigned int*>(p);
        g_timeout_flag.store(f, std::memory_order_release);
    }
    return f;
}

unsigned long long mma_timeout_ns() {
    long long ms = g_timeout_ms.load(std::memory_order_relaxed);
    if (ms < 0) {
        const char* v = getenv("BICOS_B200_MMA_TIMEOUT_MS");
        ms = v ? atoll(v) : 0;
        if (ms <= 0)
            ms = 10000; // a whole search of the largest configuration takes < 0.1 s
        g_timeout_ms.store(ms, std::memory_order_relaxed);
    }
    return (unsigned long long)ms * 1000000ull;
}
} // namespace

// Nonzero (1 + the CTA that gave up) if a wait of a tensor-core search kernel timed out since the last call;
// clears the flag. The C ABI asks at its synchronisation points and turns it into an error.
unsigned int search_mma_take_timeout() {
    unsigned int* f = g_timeout_flag.load(std::memory_order_acquire);
    if (!f)
        return 0;
    const unsigned int v = *(volatile unsigned int*)f;
    if (v)
        *(volatile unsigned int*)f = 0;
    return v;
}

// 1 = two CTAs per SM, both operands in shared memory; 2 = one CTA per SM, left operand in tensor memory
// (128 / 256 bits); 3 = the one-pass consistency kernel (128 / 256 bits, Consistency without no_dupes, two free top bits);
// 0 = automatic: 3 where it applies, else 2 where it applies and the image has an item for every SM (measured on
// the B200: 1.25 against 1.34 ms on the metric configuration, 2.35 against 2.63 ms for 256 bits x 4096
// columns; small images are served better by the finer items of variant 1). Environment
// BICOS_B200_MMA_VARIANT = 1 | 2 overrides for A/B timing.
int search_mma_variant() {
    int g = g_variant.load(std::memory_order_relaxed);
    if (g < 0) {
        const char* v = getenv("BICOS_B200_MMA_VARIANT");
        g = v && v[0] == '3' ? 3 : v && v[0] == '2' ? 2 : v && v[0] == '1' ? 1 : 0;
        g_variant.store(g, std::memory_order_relaxed);
    }
    return g;
}

void set_search_mma_variant(int v) {
    g_variant.store(v >= 1 && v <= 3 ? v : 0, std::memory_order_relaxed);
}

// Whether descriptors with a free top bit take the column-term kernels (fold32): environment
// BICOS_B200_MMA_COLTERM = 0 | 1, default BICOS_MMA_COLTERM_DEFAULT.
bool search_mma_colterm() {
    int g = g_colterm.load(std::memory_order_relaxed);
    if (g < 0) {
        const char* v = getenv("BICOS_B200_MMA_COLTERM");
        g = v ? (v[0] == '1') : BICOS_MMA_COLTERM_DEFAULT;
        g_colterm.store(g, std::memory_order_relaxed);
    }
    return g == 1;
}

void set_search_mma_colterm(bool on) {
    g_colterm.store(on ? 1 : 0, std::memory_order_relaxed);
}

// Whether launch_search_mma takes the one-pass consistency kernel (search_mma3_kernel) for this search
bool search_mma_onepass_applies(int K, int cols, int flags, int free_top_bits) {
    const int variant = search_mma_variant();
    return (K == 4 || K == 8) && cols >= 1 && cols <= COL_MAX + 1 && flags == FLAG_CONSISTENCY && free_top_bits >= 2
        && (variant == 3 || variant == 0) && search_mma_colterm();
}

bool search_mma_supports(int K, int cols) {
    return (K == 4 || K == 8 || K == 12 || K == 16) && cols >= 1 && cols <= COL_MAX + 1;
}

cudaError_t launch_search_mma(
    const uint32_t* desc0,
    const uint32_t* desc1,
    int K,
    int rows,
    int cols,
    size_t desc_pitch_words,
    int flags,
    uint32_t* fwd_first,
    uint32_t* fwd_last,
    uint32_t* rev_first,
    uint32_t* rev_last,
    cudaStream_t stream,
    int free_top_bits
) {
    const bool top_bit_free = free_top_bits >= 1;
    if (rows <= 0 || !search_mma_supports(K, cols))
        return cudaErrorInvalidValue;
    if ((((uintptr_t)desc0 | (uintptr_t)desc1) & 15) != 0 || (desc_pitch_words & 3) != 0)
        return cudaErrorMisalignedAddress; // the loader's bulk copies move 16-byte units
    MmaArgs p;
    p.left = desc0;
    p.right = desc1;
    p.rows = rows;
    p.cols = cols;
    p.pitch_words = desc_pitch_words;
    p.mtiles = (cols + TM - 1) / TM;
    p.ntiles = (cols + TN - 1) / TN;
    p.fwd_first = fwd_first;
    p.fwd_last = fwd_last;
    p.rev_first = rev_first;
    p.rev_last = rev_last;
    p.timeout_flag = mma_timeout_flag();
    p.timeout_ns = mma_timeout_ns();
    const int dirs = (flags & FLAG_CONSISTENCY) ? 2 : 1;
    const bool nodupes = (flags & FLAG_NODUPES) != 0;
    const long long pair_items = (long long)dirs * rows * ((cols + 2 * TM - 1) / (2 * TM));
    const int variant = search_mma_variant();
    // one product for both directions: 128- / 256-bit descriptors whose top TWO bits are free, Consistency without no_dupes
    if (search_mma_onepass_applies(K, cols, flags, free_top_bits)) {
        note_search_kernel("mma3<K=%d,nodupes=0,ct=2,onepass=1>", K);
        return K == 4 ? launch_k3<4>(p, stream) : launch_k3<8>(p, stream);
    }
    const bool v2 = (K == 4 || K == 8) && (variant == 2 || (variant == 0 && pair_items >= 2 * sm_count_of_current_device()));
    // column term through the MMA (see fold32): only where the caller vouches for the free top bit
    const bool ct = top_bit_free && (K == 4 || K == 8) && search_mma_colterm();
    note_search_kernel("mma%d<K=%d,nodupes=%d,ct=%d,dirs=%d>", v2 ? 2 : 1, K, (int)nodupes, (int)ct, dirs);
#define BICOS_MMA_DISPATCH(LAUNCH, KK)                                                                      \
    (ct ? (nodupes ? LAUNCH<KK, true, true>(p, dirs, stream) : LAUNCH<KK, false, true>(p, dirs, stream))      \
        : (nodupes ? LAUNCH<KK, true, false>(p, dirs, stream) : LAUNCH<KK, false, false>(p, dirs, stream)))
    if (v2 && K == 4)
        return BICOS_MMA_DISPATCH(launch_k2, 4);
    if (v2 && K == 8)
        return BICOS_MMA_DISPATCH(launch_k2, 8);
    switch (K) {
        case 4:
            return BICOS_MMA_DISPATCH(launch_k, 4);
        case 8:
            return BICOS_MMA_DISPATCH(launch_k, 8);
        case 12:
            return nodupes ? launch_k<12, true, false>(p, dirs, stream) : launch_k<12, false, false>(p, dirs, stream);
        case 16:
            return nodupes ? launch_k<16, true, false>(p, dirs, stream) : launch_k<16, false, false>(p, dirs, stream);
    }
#undef BICOS_MMA_DISPATCH
    return cudaErrorInvalidValue;
}

} // namespace bicos_b200
