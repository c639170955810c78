// TEST / BASELINE INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for the slice of <opencv2/core/cuda.hpp> that the reference's CUDA
// backend touches (src/impl/cuda.cu, include/impl/cuda/*.cuh, include/stepbuf.hpp), so that
// those sources compile *unmodified* in an image without OpenCV (oracle/Makefile, target
// `refcuda`). Containers only: a pitched ref-counted device matrix, its kernel-side view,
// a stream wrapper and a fill. No arithmetic of the matching path lives here. Nothing under
// libbicos_b200/ includes this file.
#pragma once

#include <opencv2/core.hpp>

#include <cuda_runtime.h>

#include <memory>
#include <stdexcept>
#include <string>

namespace cv {
namespace cuda {

    class Stream {
    public:
        Stream() = default;
        explicit Stream(cudaStream_t s): s_(s) {}
        static Stream& Null() {
            static Stream null_stream;
            return null_stream;
        }
        cudaStream_t raw() const {
            return s_;
        }

    private:
        cudaStream_t s_ = nullptr;
    };

    struct StreamAccessor {
        static cudaStream_t getStream(const Stream& s) {
            return s.raw();
        }
    };

    template<typename T>
    struct PtrStepSz {
        T* data;
        size_t step; // bytes
        int cols, rows;

        __host__ __device__ T* ptr(int y = 0) {
            return (T*)((char*)data + (size_t)y * step);
        }
        __host__ __device__ const T* ptr(int y = 0) const {
            return (const T*)((const char*)data + (size_t)y * step);
        }
        __host__ __device__ T& operator()(int y, int x) {
            return ptr(y)[x];
        }
        __host__ __device__ const T& operator()(int y, int x) const {
            return ptr(y)[x];
        }
    };

#ifdef __CUDACC__
    namespace shim_detail {
        template<typename T>
        __global__ void fill_kernel(PtrStepSz<T> m, T v) {
            const int x = blockIdx.x * blockDim.x + threadIdx.x;
            const int y = blockIdx.y * blockDim.y + threadIdx.y;
            if (x < m.cols && y < m.rows)
                m(y, x) = v;
        }
    } // namespace shim_detail
#endif

    class GpuMat {
    public:
        int flags = 0;
        int rows = 0, cols = 0;
        size_t step = 0;
        unsigned char* data = nullptr;

        GpuMat() = default;
        GpuMat(int r, int c, int type) {
            create(r, c, type);
        }
        GpuMat(Size sz, int type) {
            create(sz, type);
        }
        // header over memory owned by the caller
        GpuMat(int r, int c, int type, void* external, size_t step_bytes):
            flags(type),
            rows(r),
            cols(c),
            step(step_bytes),
            data(static_cast<unsigned char*>(external)) {}

        int type() const {
            return flags;
        }
        int depth() const {
            return CV_MAT_DEPTH(flags);
        }
        int channels() const {
            return CV_MAT_CN(flags);
        }
        size_t elemSize() const {
            static const size_t sizes[] = { 1, 1, 2, 2, 4, 4, 8 };
            return sizes[depth()] * channels();
        }
        Size size() const {
            return Size(cols, rows);
        }
        bool empty() const {
            return data == nullptr;
        }

        void create(int r, int c, int type) {
            if (data && r == rows && c == cols && type == flags)
                return;
            release();
            if (r <= 0 || c <= 0)
                return;
            flags = type;
            rows = r;
            cols = c;
            void* p = nullptr;
            size_t pitch = 0;
            if (cudaMallocPitch(&p, &pitch, (size_t)c * elemSize(), (size_t)r) != cudaSuccess)
                throw std::runtime_error("shim GpuMat: cudaMallocPitch failed");
            owner_.reset(p, [](void* q) { cudaFree(q); });
            data = static_cast<unsigned char*>(p);
            step = pitch;
        }
        void create(Size sz, int type) {
            create(sz.height, sz.width, type);
        }
        void release() {
            owner_.reset();
            data = nullptr;
            rows = cols = 0;
            step = 0;
        }

#ifdef __CUDACC__
        template<typename V>
        GpuMat& setTo(V value, Stream& stream = Stream::Null()) {
            const dim3 block(32, 8), grid((cols + 31) / 32, (rows + 7) / 8);
            cudaStream_t s = stream.raw();
            switch (depth()) {
                case CV_16S:
                    shim_detail::fill_kernel<short><<<grid, block, 0, s>>>(*this, (short)value);
                    break;
                case CV_32F:
                    shim_detail::fill_kernel<float><<<grid, block, 0, s>>>(*this, (float)value);
                    break;
                case CV_64F:
                    shim_detail::fill_kernel<double><<<grid, block, 0, s>>>(*this, (double)value);
                    break;
                default:
                    throw std::runtime_error("shim GpuMat::setTo: unsupported depth");
            }
            return *this;
        }
#endif

        template<typename T>
        operator PtrStepSz<T>() const {
            return PtrStepSz<T> { reinterpret_cast<T*>(data), step, cols, rows };
        }

    private:
        std::shared_ptr<void> owner_;
    };

} // namespace cuda
} // namespace cv
