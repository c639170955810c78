// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" harness around the *unmodified* reference CPU backend so that
// tests/ and bench.py (cpu_baseline / --impl reference) can call it through
// ctypes. It is compiled by oracle/Makefile together with
//   /root/reference/src/lib.cpp, src/impl/cpu.cpp, src/exception.cpp
// against oracle/shim/opencv2/core.hpp, into oracle/_ref/libbicos_ref.so.
// No reference source is copied into this repository; this file only *calls*
// the reference's public entry point BICOS::match (include/match.hpp:31-41) and
// its stage templates (include/impl/cpu/{descriptor_transform,bicos,agree}.hpp).
//
// All image arguments are dense planar arrays [n][rows][cols] (numpy C order).

#include "common.hpp"
#include "match.hpp"

#include "impl/cpu/agree.hpp"
#include "impl/cpu/bicos.hpp"
#include "impl/cpu/descriptor_transform.hpp"

#include <bitset>
#include <cstdint>
#include <cstring>
#include <string>

using namespace BICOS;
namespace rcpu = BICOS::impl::cpu;

namespace {

thread_local std::string g_error;

std::vector<cv::Mat> planes_of(const void* base, int n, int rows, int cols, int depth) {
    const size_t eb = cv::shim_depth_bytes(depth);
    std::vector<cv::Mat> v;
    v.reserve(n);
    for (int i = 0; i < n; ++i)
        v.emplace_back(
            rows,
            cols,
            CV_MAKETYPE(depth, 1),
            (void*)((const unsigned char*)base + eb * (size_t)rows * cols * i)
        );
    return v;
}

// descriptor <-> little-endian 32-bit words (bit i -> word i/32, bit i%32)
template<typename TDesc>
struct Words;
template<>
struct Words<uint32_t> {
    static constexpr int K = 1;
    static void put(uint32_t d, uint32_t* w) {
        w[0] = d;
    }
    static uint32_t get(const uint32_t* w) {
        return w[0];
    }
};
template<>
struct Words<uint64_t> {
    static constexpr int K = 2;
    static void put(uint64_t d, uint32_t* w) {
        w[0] = (uint32_t)d;
        w[1] = (uint32_t)(d >> 32);
    }
    static uint64_t get(const uint32_t* w) {
        return (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    }
};
template<>
struct Words<uint128_t> {
    static constexpr int K = 4;
    static void put(uint128_t d, uint32_t* w) {
        for (int i = 0; i < 4; ++i)
            w[i] = (uint32_t)(d >> (32 * i));
    }
    static uint128_t get(const uint32_t* w) {
        uint128_t d = 0;
        for (int i = 0; i < 4; ++i)
            d |= (uint128_t)w[i] << (32 * i);
        return d;
    }
};
// std::bitset<256> is what the reference dispatches to above 128 bits (src/impl/cpu.cpp:146-151);
// bitset<384> / bitset<512> feed the SAME stage templates for the wide-descriptor extension of this
// repository (FULL stacks of 17..23 images, which the reference's dispatch rejects at :153-155)
template<size_t NBITS>
struct Words<std::bitset<NBITS>> {
    static constexpr int K = NBITS / 32;
    static void put(const std::bitset<NBITS>& d, uint32_t* w) {
        for (int i = 0; i < K; ++i) {
            uint32_t x = 0;
            for (int b = 0; b < 32; ++b)
                x |= (uint32_t)d[32 * i + b] << b;
            w[i] = x;
        }
    }
    static std::bitset<NBITS> get(const uint32_t* w) {
        std::bitset<NBITS> d;
        for (size_t i = 0; i < NBITS; ++i)
            d[i] = (w[i / 32] >> (i % 32)) & 1u;
        return d;
    }
};

template<typename TInput, typename TDesc>
void descriptors_impl(const cv::Mat& merged, cv::Size sz, size_t n, int mode, uint32_t* out) {
    std::unique_ptr<impl::cpu::StepBuf<TDesc>> buf = mode
        ? rcpu::descriptor_transform<TInput, TDesc, rcpu::transform_full>(merged, sz, n)
        : rcpu::descriptor_transform<TInput, TDesc, rcpu::transform_limited>(merged, sz, n);
    constexpr int K = Words<TDesc>::K;
    for (int r = 0; r < sz.height; ++r)
        for (int c = 0; c < sz.width; ++c)
            Words<TDesc>::put(buf->row(r)[c], out + ((size_t)r * sz.width + c) * K);
}

template<typename TDesc>
void bicos_impl(
    const uint32_t* w0,
    const uint32_t* w1,
    int rows,
    int cols,
    int flags,
    int max_lr_diff,
    int16_t* out
) {
    constexpr int K = Words<TDesc>::K;
    cv::Size sz(cols, rows);
    auto d0 = std::make_unique<impl::cpu::StepBuf<TDesc>>(sz);
    auto d1 = std::make_unique<impl::cpu::StepBuf<TDesc>>(sz);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            d0->row(r)[c] = Words<TDesc>::get(w0 + ((size_t)r * cols + c) * K);
            d1->row(r)[c] = Words<TDesc>::get(w1 + ((size_t)r * cols + c) * K);
        }
    cv::Mat disp;
    using namespace BICOS::impl;
    switch (flags) {
        case BICOSFLAGS_NODUPES:
            rcpu::bicos<TDesc, BICOSFLAGS_NODUPES>(d0, d1, max_lr_diff, sz, disp);
            break;
        case BICOSFLAGS_CONSISTENCY:
            rcpu::bicos<TDesc, BICOSFLAGS_CONSISTENCY>(d0, d1, max_lr_diff, sz, disp);
            break;
        case BICOSFLAGS_NODUPES | BICOSFLAGS_CONSISTENCY:
            rcpu::bicos<TDesc, BICOSFLAGS_NODUPES | BICOSFLAGS_CONSISTENCY>(
                d0,
                d1,
                max_lr_diff,
                sz,
                disp
            );
            break;
        default:
            throw std::invalid_argument("bad flags");
    }
    for (int r = 0; r < rows; ++r)
        std::memcpy(out + (size_t)r * cols, disp.ptr<int16_t>(r), sizeof(int16_t) * cols);
}

} // namespace

extern "C" {

const char* ref_last_error() {
    return g_error.c_str();
}

void ref_set_threads(int n) {
    cv::shim_num_threads() = n;
}

int ref_hardware_threads() {
    return (int)std::thread::hardware_concurrency();
}

// Full path: BICOS::match. Negative threshold/step/min_variance mean "unset"
// (same convention as src/pybicos_c.cpp:59-69).
// Outputs: disparity written as int16 (disp_type 3) or float32 (disp_type 5) into
// disp_out (room for rows*cols*4 bytes); corrmap (float32, rows*cols) only if a
// threshold was given, else untouched. Returns 0, or -1 and sets ref_last_error().
int ref_match(
    const void* stack0,
    const void* stack1,
    int n,
    int rows,
    int cols,
    int depth,
    float nxcorr_threshold,
    float subpixel_step,
    float min_variance,
    int mode_full,
    int variant_consistency,
    int max_lr_diff,
    int no_dupes,
    void* disp_out,
    int* disp_type,
    float* corr_out
) {
    try {
        auto s0 = planes_of(stack0, n, rows, cols, depth);
        auto s1 = planes_of(stack1, n, rows, cols, depth);
        Config cfg;
        cfg.nxcorr_threshold =
            nxcorr_threshold >= 0 ? std::optional<float>(nxcorr_threshold) : std::nullopt;
        cfg.subpixel_step = subpixel_step >= 0 ? std::optional<float>(subpixel_step) : std::nullopt;
        cfg.min_variance = min_variance >= 0 ? std::optional<float>(min_variance) : std::nullopt;
        cfg.mode = mode_full ? TransformMode::FULL : TransformMode::LIMITED;
        if (variant_consistency)
            cfg.variant = Variant::Consistency { max_lr_diff, no_dupes != 0 };
        else
            cfg.variant = Variant::NoDuplicates {};

        cv::Mat disp, corr;
        BICOS::match(s0, s1, disp, cfg, &corr);

        *disp_type = disp.type();
        const size_t eb = disp.elemSize();
        for (int r = 0; r < rows; ++r)
            std::memcpy(
                (unsigned char*)disp_out + eb * (size_t)cols * r,
                disp.ptr<unsigned char>(r),
                eb * (size_t)cols
            );
        if (corr_out && corr.data)
            for (int r = 0; r < rows; ++r)
                std::memcpy(corr_out + (size_t)cols * r, corr.ptr<float>(r), sizeof(float) * cols);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

// Stage a2/a3: descriptors as K little-endian u32 words per pixel, K = 1/2/4/8
// chosen exactly like src/impl/cpu.cpp:122-156. Returns K, or -1.
int ref_descriptors(
    const void* stack,
    int n,
    int rows,
    int cols,
    int depth,
    int mode_full,
    uint32_t* out_words,
    int out_capacity_words_per_px
) {
    try {
        auto planes = planes_of(stack, n, rows, cols, depth);
        cv::Mat merged;
        cv::merge(planes, merged);
        // mode_full: bit 0 = TransformMode::FULL, bit 1 = allow the 384 / 512-bit extension
        const bool wide = (mode_full & 2) != 0;
        mode_full &= 1;
        const int bits = mode_full ? n * n - 2 * n + 3 : 4 * n - 7;
        const int K = bits <= 32 ? 1 : bits <= 64 ? 2 : bits <= 128 ? 4 : bits <= 256 ? 8
            : wide && bits <= 384 ? 12 : wide && bits <= 512 ? 16 : -1;
        if (K < 0)
            throw std::invalid_argument("too many bits: " + std::to_string(bits));
        if (K > out_capacity_words_per_px)
            throw std::invalid_argument("output capacity too small");
        cv::Size sz(cols, rows);
        const bool u8 = depth == CV_8U;
#define DISPATCH(TD) \
    (u8 ? descriptors_impl<uint8_t, TD>(merged, sz, n, mode_full, out_words) \
        : descriptors_impl<uint16_t, TD>(merged, sz, n, mode_full, out_words))
        switch (K) {
            case 1:
                DISPATCH(uint32_t);
                break;
            case 2:
                DISPATCH(uint64_t);
                break;
            case 4:
                DISPATCH(uint128_t);
                break;
            case 8:
                DISPATCH(std::bitset<256>);
                break;
            case 12:
                DISPATCH(std::bitset<384>);
                break;
            case 16:
                DISPATCH(std::bitset<512>);
                break;
        }
#undef DISPATCH
        return K;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

// Stage a6/a7: search + postfilter on caller-supplied descriptors.
// flags: 1 = NODUPES, 2 = CONSISTENCY, 3 = both (include/impl/common.hpp:46-47).
int ref_bicos(
    const uint32_t* desc0,
    const uint32_t* desc1,
    int K,
    int rows,
    int cols,
    int flags,
    int max_lr_diff,
    int16_t* out
) {
    try {
        switch (K) {
            case 1:
                bicos_impl<uint32_t>(desc0, desc1, rows, cols, flags, max_lr_diff, out);
                break;
            case 2:
                bicos_impl<uint64_t>(desc0, desc1, rows, cols, flags, max_lr_diff, out);
                break;
            case 4:
                bicos_impl<uint128_t>(desc0, desc1, rows, cols, flags, max_lr_diff, out);
                break;
            case 8:
                bicos_impl<std::bitset<256>>(desc0, desc1, rows, cols, flags, max_lr_diff, out);
                break;
            case 12:
                bicos_impl<std::bitset<384>>(desc0, desc1, rows, cols, flags, max_lr_diff, out);
                break;
            case 16:
                bicos_impl<std::bitset<512>>(desc0, desc1, rows, cols, flags, max_lr_diff, out);
                break;
            default:
                throw std::invalid_argument("bad K");
        }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

// Stage a9/a10: refine a caller-supplied raw int16 disparity.
// subpixel_step < 0: agree (raw_disp modified in place, copied to disp_i16_out);
// else agree_subpixel (float result in disp_f32_out). min_variance_times_n < 0 = unset
// (the caller multiplies by n, like src/impl/cpu.cpp:127).
int ref_agree(
    const int16_t* raw_disp,
    const void* stack0,
    const void* stack1,
    int n,
    int rows,
    int cols,
    int depth,
    float nxcorr_threshold,
    float subpixel_step,
    float min_variance_times_n,
    int16_t* disp_i16_out,
    float* disp_f32_out,
    float* corr_out
) {
    try {
        auto p0 = planes_of(stack0, n, rows, cols, depth);
        auto p1 = planes_of(stack1, n, rows, cols, depth);
        cv::Mat m0, m1;
        cv::merge(p0, m0);
        cv::merge(p1, m1);
        cv::Mat raw(rows, cols, CV_16SC1);
        std::memcpy(raw.data, raw_disp, sizeof(int16_t) * (size_t)rows * cols);
        cv::Mat corr(rows, cols, CV_32FC1);
        corr.setTo(std::numeric_limits<float>::quiet_NaN());
        std::optional<float> mv = min_variance_times_n >= 0
            ? std::optional<float>(min_variance_times_n)
            : std::nullopt;
        const bool u8 = depth == CV_8U;
        if (subpixel_step < 0) {
            if (u8)
                rcpu::agree<uint8_t>(raw, m0, m1, n, nxcorr_threshold, mv, &corr);
            else
                rcpu::agree<uint16_t>(raw, m0, m1, n, nxcorr_threshold, mv, &corr);
            std::memcpy(disp_i16_out, raw.data, sizeof(int16_t) * (size_t)rows * cols);
        } else {
            cv::Mat_<float> fd;
            if (u8)
                rcpu::agree_subpixel<uint8_t>(raw, m0, m1, n, nxcorr_threshold, subpixel_step, mv, fd, &corr);
            else
                rcpu::agree_subpixel<uint16_t>(raw, m0, m1, n, nxcorr_threshold, subpixel_step, mv, fd, &corr);
            std::memcpy(disp_f32_out, fd.data, sizeof(float) * (size_t)rows * cols);
        }
        if (corr_out)
            std::memcpy(corr_out, corr.data, sizeof(float) * (size_t)rows * cols);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

} // extern "C"
