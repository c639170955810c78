// Public types of the B200-native BICOS matcher. Same names, fields, defaults and
// meaning as the reference's include/common.hpp:34-90 (Config, TransformMode, Precision,
// Variant::NoDuplicates / Variant::Consistency, SearchVariant, INVALID_DISP, is_invalid,
// Exception), so that code written against libBICOS compiles against this header.
// This build is always a CUDA build (there is no CPU fallback), hence Precision exists.
#pragma once

#include <cmath>
#include <exception>
#include <limits>
#include <optional>
#include <string>
#include <type_traits>
#include <variant>

#if defined(BICOS_WITH_OPENCV)
    #include <opencv2/core.hpp>
    #include <opencv2/core/cuda.hpp>
#else
    #include "image.hpp"
#endif

#ifndef BICOS_CUDA
    #define BICOS_CUDA 1
#endif

namespace BICOS {

using uint128_t = __uint128_t;

// NaN for floating point disparities, the lowest value (-32768 for int16) for integers
template<typename T>
constexpr T INVALID_DISP = std::numeric_limits<T>::has_quiet_NaN
    ? std::numeric_limits<T>::quiet_NaN()
    : std::numeric_limits<T>::lowest();

template<typename T>
constexpr bool is_invalid(T disparity) {
    if constexpr (std::is_floating_point_v<T>)
        return disparity != disparity;
    else
        return disparity == INVALID_DISP<T>;
}

#if defined(BICOS_WITH_OPENCV)
using Image = cv::cuda::GpuMat;
#endif

enum class TransformMode { LIMITED, FULL };
enum class Precision { SINGLE, DOUBLE };

namespace Variant {
    struct NoDuplicates {};
    struct Consistency {
        int max_lr_diff = 1;
        bool no_dupes = false;
    };
} // namespace Variant

using SearchVariant = std::variant<Variant::NoDuplicates, Variant::Consistency>;

struct Config {
    std::optional<float> nxcorr_threshold = 0.5f;
    std::optional<float> subpixel_step = std::nullopt;
    std::optional<float> min_variance = std::nullopt;
    TransformMode mode = TransformMode::LIMITED;
    Precision precision = Precision::SINGLE;
    SearchVariant variant = Variant::NoDuplicates {};
};

class Exception: public std::exception {
    std::string message_;

public:
    explicit Exception(const std::string& message): message_(message) {}
    const char* what() const noexcept override {
        return message_.c_str();
    }
};

} // namespace BICOS
