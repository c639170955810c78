// Host side of the C++ API: BICOS::Image and BICOS::match on top of the C ABI
// (include/bicos_b200.h). Mirrors the validation and error behaviour of the reference's
// drivers (src/lib.cpp:31-49, src/impl/cpu.cpp:100-159, src/impl/cuda.cu:465-524).

#include "../../include/BICOS/match.hpp"
#include "../../include/bicos_b200.h"

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

namespace BICOS {

namespace {

void cuda_check(cudaError_t err, const char* what) {
    if (err != cudaSuccess)
        throw Exception(std::string(what) + ": " + cudaGetErrorString(err));
}

[[noreturn]] void rethrow_last(int status) {
    const std::string msg = bicos_b200_last_error();
    // the reference throws std::invalid_argument for ">256 bits" and BICOS::Exception otherwise
    if (status == BICOS_B200_ERR_INVALID && msg.rfind("input stacks too large", 0) == 0)
        throw std::invalid_argument(msg);
    throw Exception(msg);
}

// one workspace per (thread, device): BICOS::match is re-entrant across threads like the
// reference (which has no global mutable state), and repeated calls reuse their buffers.
struct HandleCache {
    std::map<int, bicos_b200_handle> handles;
    ~HandleCache() {
        for (auto& kv: handles)
            bicos_b200_destroy(kv.second);
    }
    bicos_b200_handle get(int device) {
        auto it = handles.find(device);
        if (it != handles.end())
            return it->second;
        bicos_b200_handle h = nullptr;
        const int rc = bicos_b200_create(&h, device);
        if (rc != 0)
            rethrow_last(rc);
        handles[device] = h;
        return h;
    }
};

thread_local HandleCache t_handles;

bicos_b200_config to_c_config(const Config& cfg) {
    bicos_b200_config c {};
    c.nxcorr_threshold = cfg.nxcorr_threshold.value_or(-1.f);
    c.negative_threshold_is_set = cfg.nxcorr_threshold.has_value() && !(*cfg.nxcorr_threshold >= 0) ? 1 : 0;
    c.subpixel_step = cfg.subpixel_step.value_or(-1.f);
    c.min_variance = cfg.min_variance.value_or(-1.f);
    c.mode = cfg.mode == TransformMode::FULL ? 1 : 0;
    c.precision = cfg.precision == Precision::DOUBLE ? 1 : 0;
    c.wide_descriptors = cfg.wide_descriptors ? 1 : 0;
    if (std::holds_alternative<Variant::Consistency>(cfg.variant)) {
        const auto& v = std::get<Variant::Consistency>(cfg.variant);
        c.variant_type = 1;
        c.max_lr_diff = v.max_lr_diff;
        c.no_dupes = v.no_dupes ? 1 : 0;
    } else {
        c.variant_type = 0;
        c.max_lr_diff = 1;
        c.no_dupes = 0;
    }
    return c;
}

struct StackView {
    std::vector<const void*> planes;
    int rows = 0, cols = 0, depth = 0;
    size_t step = 0;
};

StackView view_of(const std::vector<Image>& stack, const char* name) {
    StackView v;
    if (stack.size() < 2)
        throw Exception("need at least two images");
    const Image& first = stack.front();
    v.rows = first.rows;
    v.cols = first.cols;
    v.depth = first.depth();
    v.step = first.step;
    if (first.type() != IMG_8U && first.type() != IMG_16U)
        throw Exception("bad input depths, only CV_8UC1 and CV_16UC1 are supported");
    v.planes.reserve(stack.size());
    for (const Image& im: stack) {
        if (im.rows != v.rows || im.cols != v.cols || im.type() != first.type() || im.step != v.step || im.empty())
            throw Exception(std::string("images of ") + name + " differ in size, type or pitch");
        v.planes.push_back(im.data);
    }
    return v;
}

int device_of(const void* ptr) {
    cudaPointerAttributes attr {};
    cuda_check(cudaPointerGetAttributes(&attr, ptr), "cudaPointerGetAttributes");
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)
        throw Exception("input images must live in device memory");
    return attr.device;
}

} // namespace

size_t image_elem_size(int type) {
    switch (type & 7) {
        case IMG_8U:
        case 1:
            return 1;
        case IMG_16U:
        case IMG_16S:
            return 2;
        case 4:
        case IMG_32F:
            return 4;
        case IMG_64F:
            return 8;
    }
    throw Exception("unsupported image type");
}

Image::Image(int r, int c, int type, void* device_ptr, size_t stepb):
    rows(r),
    cols(c),
    step(stepb ? stepb : (size_t)c * image_elem_size(type)),
    data(static_cast<unsigned char*>(device_ptr)),
    type_(type) {}

void Image::create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_)
        return;
    release();
    if (r <= 0 || c <= 0)
        throw Exception("cannot create an empty image");
    void* ptr = nullptr;
    size_t pitch = 0;
    cuda_check(cudaMallocPitch(&ptr, &pitch, (size_t)c * image_elem_size(type), (size_t)r), "cudaMallocPitch");
    owner_ = std::shared_ptr<void>(ptr, [](void* p) { cudaFree(p); });
    data = static_cast<unsigned char*>(ptr);
    rows = r;
    cols = c;
    step = pitch;
    type_ = type;
}

void Image::release() {
    owner_.reset();
    data = nullptr;
    rows = cols = 0;
    step = 0;
}

void Image::upload(const HostImage& host, void* stream) {
    create(host.rows, host.cols, host.type());
    cuda_check(
        cudaMemcpy2DAsync(data, step, host.data, host.step, (size_t)cols * elemSize(), (size_t)rows,
                          cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)),
        "upload"
    );
    if (!stream)
        cuda_check(cudaStreamSynchronize(nullptr), "upload sync");
}

void Image::download(const HostImage& host, void* stream) const {
    if (host.rows != rows || host.cols != cols || host.type() != type_ || !host.data)
        throw Exception("download target does not match the image");
    cuda_check(
        cudaMemcpy2DAsync(host.data, host.step, data, step, (size_t)cols * elemSize(), (size_t)rows,
                          cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)),
        "download"
    );
    if (!stream)
        cuda_check(cudaStreamSynchronize(nullptr), "download sync");
}

void match(
    const std::vector<Image>& stack0,
    const std::vector<Image>& stack1,
    Image& disparity,
    Config cfg,
    Image* corrmap,
    void* stream
) {
    const StackView v0 = view_of(stack0, "stack0");
    const StackView v1 = view_of(stack1, "stack1");
    if (stack0.size() != stack1.size() || v0.rows != v1.rows || v0.cols != v1.cols || v0.depth != v1.depth || v0.step != v1.step)
        throw Exception("stack0 and stack1 differ in length, size, type or pitch");

    const bicos_b200_config c = to_c_config(cfg);
    const int device = device_of(v0.planes.front());
    bicos_b200_handle h = t_handles.get(device);

    int prev = 0;
    cuda_check(cudaGetDevice(&prev), "cudaGetDevice");
    if (prev != device)
        cuda_check(cudaSetDevice(device), "cudaSetDevice");
    try {
        disparity.create(v0.rows, v0.cols, bicos_b200_disparity_type(&c));
        const bool want_corr = corrmap && cfg.nxcorr_threshold.has_value(); // cpu.cpp:77-81
        if (want_corr)
            corrmap->create(v0.rows, v0.cols, bicos_b200_corrmap_type(&c));
        const int rc = bicos_b200_match(
            h, v0.planes.data(), v1.planes.data(), (int)stack0.size(), v0.rows, v0.cols, v0.step, v0.depth, &c,
            disparity.data, disparity.step, want_corr ? corrmap->data : nullptr, want_corr ? corrmap->step : 0, stream
        );
        if (rc != 0)
            rethrow_last(rc);
    } catch (...) {
        if (prev != device)
            cudaSetDevice(prev);
        throw;
    }
    if (prev != device)
        cudaSetDevice(prev);
}

void match_sharded(
    const std::vector<Image>& stack0,
    const std::vector<Image>& stack1,
    Image& disparity,
    const std::vector<int>& devices,
    Config cfg,
    Image* corrmap
) {
    if (devices.empty())
        throw Exception("no devices given");
    const StackView v0 = view_of(stack0, "stack0");
    const StackView v1 = view_of(stack1, "stack1");
    if (stack0.size() != stack1.size() || v0.rows != v1.rows || v0.cols != v1.cols || v0.depth != v1.depth || v0.step != v1.step)
        throw Exception("stack0 and stack1 differ in length, size, type or pitch");
    const int home = devices.front();
    if (device_of(v0.planes.front()) != home)
        throw Exception("inputs must be resident on devices[0]");

    const bicos_b200_config c = to_c_config(cfg);
    int prev = 0;
    cuda_check(cudaGetDevice(&prev), "cudaGetDevice");

    cuda_check(cudaSetDevice(home), "cudaSetDevice");
    disparity.create(v0.rows, v0.cols, bicos_b200_disparity_type(&c));
    const bool want_corr = corrmap && cfg.nxcorr_threshold.has_value();
    if (want_corr)
        corrmap->create(v0.rows, v0.cols, bicos_b200_corrmap_type(&c));
    cuda_check(cudaDeviceSynchronize(), "sync home device"); // inputs and outputs visible to the peers

    const int G = (int)devices.size();
    std::string error;
    for (int g = 0; g < G && error.empty(); ++g) {
        const int dev = devices[g];
        cuda_check(cudaSetDevice(dev), "cudaSetDevice");
        if (dev != home) {
            int can = 0;
            cuda_check(cudaDeviceCanAccessPeer(&can, dev, home), "cudaDeviceCanAccessPeer");
            if (!can)
                throw Exception("no peer access from device " + std::to_string(dev) + " to " + std::to_string(home));
            const cudaError_t e = cudaDeviceEnablePeerAccess(home, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                cuda_check(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        const int rb = (int)((long long)v0.rows * g / G), re = (int)((long long)v0.rows * (g + 1) / G);
        if (rb >= re)
            continue;
        bicos_b200_handle h = t_handles.get(dev);
        const int rc = bicos_b200_match_rows(
            h, v0.planes.data(), v1.planes.data(), (int)stack0.size(), v0.rows, v0.cols, v0.step, v0.depth, &c, rb, re,
            disparity.data, disparity.step, want_corr ? corrmap->data : nullptr, want_corr ? corrmap->step : 0, nullptr
        );
        if (rc != 0)
            error = bicos_b200_last_error();
    }
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(devices[g]);
        cudaDeviceSynchronize();
    }
    cudaSetDevice(prev);
    if (!error.empty())
        throw Exception(error);
}

} // namespace BICOS
