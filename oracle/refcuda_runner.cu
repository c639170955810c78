// BASELINE INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" harness around the *unmodified* reference CUDA backend
// (/root/reference/src/impl/cuda.cu + include/impl/cuda/*.cuh), compiled for sm_100a by
// oracle/Makefile (target `refcuda`) against the header stand-ins in oracle/shim, into
// oracle/_ref/libbicos_refcuda.so. It lets bench.py / tools report "the reference's own CUDA
// build on the same B200" next to the new kernels (BASELINE.json north_star) and lets the GPU
// tests cross-check integer outputs. No reference source is copied into this repository;
// this file only *calls* BICOS::match (reference include/match.hpp:31-41, CUDA signature).
//
// Inputs are dense planar host arrays [n][rows][cols]; they are uploaded once into pitched
// device matrices (what a cv::cuda::GpuMat user would hold) before anything is timed.

#include "common.hpp"
#include "match.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

using namespace BICOS;

// same layout as the reference's BicosConfig (src/pybicos_c.cpp:30-41, CUDA build)
struct RefConfig {
    float nxcorr_threshold, subpixel_step, min_variance; // negative = unset
    int mode, precision, variant_type, max_lr_diff, no_dupes;
};

namespace {

thread_local std::string g_error;

Config to_config(const RefConfig& c) {
    Config cfg;
    cfg.nxcorr_threshold = c.nxcorr_threshold >= 0 ? std::optional<float>(c.nxcorr_threshold) : std::nullopt;
    cfg.subpixel_step = c.subpixel_step >= 0 ? std::optional<float>(c.subpixel_step) : std::nullopt;
    cfg.min_variance = c.min_variance >= 0 ? std::optional<float>(c.min_variance) : std::nullopt;
    cfg.mode = c.mode ? TransformMode::FULL : TransformMode::LIMITED;
    cfg.precision = c.precision ? Precision::DOUBLE : Precision::SINGLE;
    if (c.variant_type)
        cfg.variant = Variant::Consistency { c.max_lr_diff, c.no_dupes != 0 };
    else
        cfg.variant = Variant::NoDuplicates {};
    return cfg;
}

std::vector<cv::cuda::GpuMat> upload(const void* host, int n, int rows, int cols, int depth) {
    const size_t eb = depth == CV_16U ? 2 : 1;
    std::vector<cv::cuda::GpuMat> v(n);
    for (int i = 0; i < n; ++i) {
        v[i].create(rows, cols, CV_MAKETYPE(depth, 1));
        const unsigned char* src = static_cast<const unsigned char*>(host) + eb * (size_t)rows * cols * i;
        if (cudaMemcpy2D(v[i].data, v[i].step, src, eb * cols, eb * cols, rows, cudaMemcpyHostToDevice) != cudaSuccess)
            throw std::runtime_error("upload failed");
    }
    return v;
}

void download(const cv::cuda::GpuMat& m, void* host) {
    if (!host || m.empty())
        return;
    const size_t row_bytes = (size_t)m.cols * m.elemSize();
    if (cudaMemcpy2D(host, row_bytes, m.data, m.step, row_bytes, m.rows, cudaMemcpyDeviceToHost) != cudaSuccess)
        throw std::runtime_error("download failed");
}

} // namespace

extern "C" {

const char* refcuda_last_error() {
    return g_error.c_str();
}

// One match; outputs copied to dense host buffers (each may be null). *disp_type / *corr_type
// receive the OpenCV type codes of what the reference produced (0 = not produced).
int refcuda_match(const void* left, const void* right, int n, int rows, int cols, int depth,
                  const RefConfig* c, void* disp_host, int* disp_type, void* corr_host, int* corr_type) {
    try {
        auto s0 = upload(left, n, rows, cols, depth);
        auto s1 = upload(right, n, rows, cols, depth);
        cv::cuda::GpuMat disp, corr;
        BICOS::match(s0, s1, disp, to_config(*c), &corr);
        if (cudaDeviceSynchronize() != cudaSuccess)
            throw std::runtime_error(cudaGetErrorString(cudaGetLastError()));
        if (disp_type)
            *disp_type = disp.empty() ? 0 : disp.type();
        if (corr_type)
            *corr_type = corr.empty() ? 0 : corr.type();
        download(disp, disp_host);
        download(corr, corr_host);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// Device-resident timing of BICOS::match as a caller sees it (its per-call allocations and
// host registrations included, as in the reference's own CLI timing, src/cli.cpp:177-205).
// Every call is timed on its own between two CUDA events (the reference's match() ends with
// cudaFree / cudaHostUnregister, i.e. it is synchronous); the per-call times vary a lot because
// of those allocations, so both the median and the minimum over `iters` calls are returned.
int refcuda_time(const void* left, const void* right, int n, int rows, int cols, int depth,
                 const RefConfig* c, int warmup, int iters, float* ms_median, float* ms_min) {
    try {
        auto s0 = upload(left, n, rows, cols, depth);
        auto s1 = upload(right, n, rows, cols, depth);
        const Config cfg = to_config(*c);
        cv::cuda::GpuMat disp, corr;
        for (int i = 0; i < warmup; ++i)
            BICOS::match(s0, s1, disp, cfg, &corr);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        std::vector<float> times;
        for (int i = 0; i < iters; ++i) {
            cudaEventRecord(e0, nullptr);
            BICOS::match(s0, s1, disp, cfg, &corr);
            cudaEventRecord(e1, nullptr);
            if (cudaEventSynchronize(e1) != cudaSuccess)
                throw std::runtime_error(cudaGetErrorString(cudaGetLastError()));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            times.push_back(ms);
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        std::sort(times.begin(), times.end());
        *ms_median = times[times.size() / 2];
        *ms_min = times.front();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

} // extern "C"
