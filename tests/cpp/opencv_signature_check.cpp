// Compile-only check (tests/test_abi.py): with -DBICOS_WITH_OPENCV, BICOS::match has the reference's
// signature (reference include/match.hpp:31-41): Image = cv::cuda::GpuMat and the sixth parameter is a
// cv::cuda::Stream&, defaulted to Stream::Null(). Compiled against the stand-in OpenCV headers under
// oracle/shim (test infrastructure); nothing here is linked or run.
#include <BICOS/match.hpp>

#include <type_traits>

static_assert(std::is_same_v<BICOS::Image, cv::cuda::GpuMat>, "Image must alias cv::cuda::GpuMat");

void reference_style_calls(const std::vector<cv::cuda::GpuMat>& s0, const std::vector<cv::cuda::GpuMat>& s1) {
    cv::cuda::GpuMat disparity, corrmap;
    BICOS::Config cfg;
    cfg.nxcorr_threshold = 0.9f;
    cfg.variant = BICOS::Variant::Consistency { 1, true };
    BICOS::match(s0, s1, disparity); // all defaults
    BICOS::match(s0, s1, disparity, cfg);
    BICOS::match(s0, s1, disparity, cfg, &corrmap);
    cv::cuda::Stream stream;
    BICOS::match(s0, s1, disparity, cfg, &corrmap, stream); // the reference's sixth parameter
    BICOS::match(s0, s1, disparity, cfg, &corrmap, static_cast<void*>(nullptr)); // raw cudaStream_t form
}
