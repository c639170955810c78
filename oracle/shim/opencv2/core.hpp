// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for the slice of <opencv2/core.hpp> that the reference's CPU
// backend touches (src/impl/cpu.cpp, include/impl/cpu/*.hpp, include/stepbuf.hpp,
// include/common.hpp). It exists so that the reference sources can be compiled
// *unmodified* in a container without OpenCV (see oracle/Makefile). Nothing in
// libbicos_b200/ includes this file.
//
// Only containers, merge, convertTo(16S->32F), setTo and a row-parallel loop are
// provided; there is no arithmetic of the matching path in here.
#pragma once

#include <alloca.h>
#include <sys/types.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6

#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags)&CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)

#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

namespace cuda {
    class GpuMat; // real OpenCV forward-declares this in core.hpp as well
    class Stream;
} // namespace cuda

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h): width(w), height(h) {}
    int area() const {
        return width * height;
    }
    bool operator==(const Size& o) const {
        return width == o.width && height == o.height;
    }
};

struct Range {
    int start = 0, end = 0;
    Range() = default;
    Range(int s, int e): start(s), end(e) {}
    int size() const {
        return end - start;
    }
};

// static split of [start, end) over the host threads; the reference only needs
// "every index visited exactly once".
inline int& shim_num_threads() {
    static int n = 0; // 0 = hardware_concurrency
    return n;
}

template<typename F>
inline void parallel_for_(const Range& range, F&& body) {
    const int total = range.size();
    if (total <= 0)
        return;
    int nthreads = shim_num_threads() > 0 ? shim_num_threads()
                                          : (int)std::thread::hardware_concurrency();
    nthreads = std::max(1, std::min(nthreads, total));
    if (nthreads == 1) {
        body(range);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (int i = 0; i < nthreads; ++i) {
        const int lo = range.start + (int)((long long)total * i / nthreads);
        const int hi = range.start + (int)((long long)total * (i + 1) / nthreads);
        pool.emplace_back([&body, lo, hi] { body(Range(lo, hi)); });
    }
    for (auto& t: pool)
        t.join();
}

template<typename T>
struct DataType;
template<>
struct DataType<uint8_t> {
    enum { type = CV_8UC1 };
};
template<>
struct DataType<uint16_t> {
    enum { type = CV_16UC1 };
};
template<>
struct DataType<int16_t> {
    enum { type = CV_16SC1 };
};
template<>
struct DataType<int32_t> {
    enum { type = CV_MAKETYPE(CV_32S, 1) };
};
template<>
struct DataType<float> {
    enum { type = CV_32FC1 };
};
template<>
struct DataType<double> {
    enum { type = CV_64FC1 };
};

inline size_t shim_depth_bytes(int depth) {
    switch (depth) {
        case CV_8U:
        case CV_8S:
            return 1;
        case CV_16U:
        case CV_16S:
            return 2;
        case CV_32S:
        case CV_32F:
            return 4;
        case CV_64F:
            return 8;
    }
    throw std::invalid_argument("shim: bad depth");
}

class Mat {
public:
    int flags = 0; // type code only
    int rows = 0, cols = 0;
    size_t step = 0; // bytes per row
    unsigned char* data = nullptr;

    Mat() = default;
    Mat(int r, int c, int type) {
        create(r, c, type);
    }
    Mat(Size sz, int type) {
        create(sz, type);
    }
    // header over caller-owned memory (no copy, no ownership)
    Mat(int r, int c, int type, void* ext, size_t stepb = 0):
        flags(type),
        rows(r),
        cols(c),
        data((unsigned char*)ext) {
        step = stepb ? stepb : (size_t)c * elemSize();
    }

    int type() const {
        return flags;
    }
    int depth() const {
        return CV_MAT_DEPTH(flags);
    }
    int channels() const {
        return CV_MAT_CN(flags);
    }
    size_t elemSize1() const {
        return shim_depth_bytes(depth());
    }
    size_t elemSize() const {
        return elemSize1() * channels();
    }
    size_t total() const {
        return (size_t)rows * cols;
    }
    Size size() const {
        return Size(cols, rows);
    }
    bool empty() const {
        return data == nullptr || total() == 0;
    }

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == flags)
            return;
        flags = type;
        rows = r;
        cols = c;
        step = (size_t)c * elemSize();
        const size_t bytes = step * (size_t)r;
        owner_ = std::shared_ptr<unsigned char>(
            (unsigned char*)std::malloc(bytes ? bytes : 1),
            [](unsigned char* p) { std::free(p); }
        );
        data = owner_.get();
    }
    void create(Size sz, int type) {
        create(sz.height, sz.width, type);
    }

    template<typename T>
    T* ptr(int r = 0) {
        return (T*)(data + step * (size_t)r);
    }
    template<typename T>
    const T* ptr(int r = 0) const {
        return (const T*)(data + step * (size_t)r);
    }
    // (row, col) addresses the first channel of an interleaved pixel
    template<typename T>
    T* ptr(int r, int c) {
        return (T*)(data + step * (size_t)r + elemSize() * (size_t)c);
    }
    template<typename T>
    const T* ptr(int r, int c) const {
        return (const T*)(data + step * (size_t)r + elemSize() * (size_t)c);
    }

    template<typename T>
    T& at(int r, int c) {
        return ((T*)(data + step * (size_t)r))[c];
    }
    template<typename T>
    const T& at(int r, int c) const {
        return ((const T*)(data + step * (size_t)r))[c];
    }
    // linear index; the reference only uses it on single-row headers
    template<typename T>
    T& at(int i) {
        return rows == 1 ? ((T*)data)[i] : at<T>(i / cols, i % cols);
    }
    template<typename T>
    const T& at(int i) const {
        return rows == 1 ? ((const T*)data)[i] : at<T>(i / cols, i % cols);
    }

    Mat row(int r) const {
        Mat h;
        h.flags = flags;
        h.rows = 1;
        h.cols = cols;
        h.step = step;
        h.data = data + step * (size_t)r;
        h.owner_ = owner_;
        return h;
    }

    template<typename S>
    Mat& setTo(S value) {
        switch (depth()) {
            case CV_8U:
                fill_<uint8_t>(value);
                break;
            case CV_16U:
                fill_<uint16_t>(value);
                break;
            case CV_16S:
                fill_<int16_t>(value);
                break;
            case CV_32S:
                fill_<int32_t>(value);
                break;
            case CV_32F:
                fill_<float>(value);
                break;
            case CV_64F:
                fill_<double>(value);
                break;
            default:
                throw std::invalid_argument("shim: setTo depth");
        }
        return *this;
    }

    void convertTo(Mat& dst, int rtype) const {
        if (depth() != CV_16S || CV_MAT_DEPTH(rtype) != CV_32F || channels() != 1)
            throw std::invalid_argument("shim: convertTo only does 16S -> 32F");
        Mat out;
        out.create(rows, cols, CV_32FC1);
        for (int r = 0; r < rows; ++r) {
            const int16_t* s = ptr<int16_t>(r);
            float* d = out.ptr<float>(r);
            for (int c = 0; c < cols; ++c)
                d[c] = (float)s[c];
        }
        dst.assign_(out);
    }

protected:
    std::shared_ptr<unsigned char> owner_;

    void assign_(const Mat& o) {
        flags = o.flags;
        rows = o.rows;
        cols = o.cols;
        step = o.step;
        data = o.data;
        owner_ = o.owner_;
    }

    template<typename T, typename S>
    void fill_(S value) {
        const T v = (T)value;
        const size_t per_row = (size_t)cols * channels();
        for (int r = 0; r < rows; ++r) {
            T* p = ptr<T>(r);
            std::fill(p, p + per_row, v);
        }
    }
};

template<typename T>
class Mat_: public Mat {
public:
    Mat_() {
        flags = DataType<T>::type;
    }
    // shares the buffer when the element type already matches (all the reference needs)
    Mat_(const Mat& m) {
        if (m.data && m.type() != DataType<T>::type)
            throw std::invalid_argument("shim: Mat_ converting ctor with type change");
        Mat::operator=(m);
        flags = DataType<T>::type;
    }
    Mat_& operator=(const Mat& m) {
        if (m.data && m.type() != DataType<T>::type)
            throw std::invalid_argument("shim: Mat_ assignment with type change");
        Mat::operator=(m);
        flags = DataType<T>::type;
        return *this;
    }
    Mat_ row(int r) const {
        return Mat_(Mat::row(r));
    }
};

using Mat1s = Mat_<short>;
using Mat1f = Mat_<float>;

// planar single-channel images -> one interleaved H x W x n image
inline void merge(const std::vector<Mat>& planes, Mat& dst) {
    const int n = (int)planes.size();
    if (n == 0)
        throw std::invalid_argument("shim: merge of nothing");
    const Mat& first = planes.front();
    const size_t eb = first.elemSize1();
    dst.create(first.rows, first.cols, CV_MAKETYPE(first.depth(), n));
    for (int k = 0; k < n; ++k) {
        const Mat& p = planes[k];
        if (p.rows != first.rows || p.cols != first.cols || p.type() != first.type())
            throw std::invalid_argument("shim: merge of unequal planes");
        for (int r = 0; r < p.rows; ++r) {
            const unsigned char* s = p.data + p.step * (size_t)r;
            unsigned char* d = dst.data + dst.step * (size_t)r + eb * (size_t)k;
            for (int c = 0; c < p.cols; ++c)
                std::memcpy(d + eb * (size_t)n * c, s + eb * c, eb);
        }
    }
}

} // namespace cv
