// TEST / BASELINE INFRASTRUCTURE ONLY: stand-in for <opencv2/core/cuda/common.hpp> (see ../cuda.hpp).
#pragma once

namespace cv {
namespace cuda {
    namespace device {
        __host__ __device__ inline int divUp(int total, int grain) {
            return (total + grain - 1) / grain;
        }
    } // namespace device
} // namespace cuda
} // namespace cv
