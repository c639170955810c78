// TEST / BASELINE INFRASTRUCTURE ONLY: stand-in for <opencv2/core/cuda_stream_accessor.hpp>;
// cv::cuda::StreamAccessor lives in the cuda.hpp stand-in.
#pragma once
#include <opencv2/core/cuda.hpp>
