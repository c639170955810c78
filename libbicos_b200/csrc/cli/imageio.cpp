#include "imageio.hpp"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace bicos_cli {
namespace {

std::vector<uint8_t> slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f)
        throw std::runtime_error("cannot open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

uint32_t be32(const uint8_t* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// ITU-R BT.601 luma with the 14-bit fixed-point weights OpenCV uses for 8-bit BGR2GRAY
inline uint32_t luma(uint32_t r, uint32_t g, uint32_t b) {
    return (r * 4899u + g * 9617u + b * 1868u + 8192u) >> 14;
}

GrayImage read_png(const std::vector<uint8_t>& file, const std::string& path, bool& was_colour) {
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0)
        throw std::runtime_error(path + ": not a PNG file");
    size_t pos = 8;
    uint32_t width = 0, height = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, palette;
    bool have_ihdr = false;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size())
            throw std::runtime_error(path + ": truncated PNG chunk");
        const uint8_t* body = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13)
                throw std::runtime_error(path + ": bad PNG IHDR length");
            width = be32(body);
            height = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
            have_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || width == 0 || height == 0)
        throw std::runtime_error(path + ": PNG without IHDR");
    // the library's own limits (include/bicos_b200.h: 32767 columns, 65535 rows): nothing larger is allocated
    if (width > 32767u || height > 65535u)
        throw std::runtime_error(path + ": PNG larger than 32767 x 65535");
    if (interlace != 0)
        throw std::runtime_error(path + ": interlaced PNG is not supported");
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: throw std::runtime_error(path + ": bad PNG colour type");
    }
    if (ctype == 3 && depth == 16)
        throw std::runtime_error(path + ": palette PNG with 16-bit indices");
    if (!(depth == 8 || depth == 16 || (ctype == 0 && (depth == 1 || depth == 2 || depth == 4))
          || (ctype == 3 && (depth == 1 || depth == 2 || depth == 4))))
        throw std::runtime_error(path + ": unsupported PNG bit depth");
    const size_t bpp_bits = (size_t)channels * depth;
    const size_t stride = (width * bpp_bits + 7) / 8;
    const size_t bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;
    std::vector<uint8_t> raw((stride + 1) * height);
    uLongf out_len = raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), idat.size()) != Z_OK || out_len != raw.size())
        throw std::runtime_error(path + ": PNG data does not inflate");
    // undo the scanline filters in place
    std::vector<uint8_t> prev(stride, 0);
    for (uint32_t y = 0; y < height; ++y) {
        uint8_t* line = &raw[(stride + 1) * y];
        const int filter = line[0];
        uint8_t* cur = line + 1;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = cur[i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: throw std::runtime_error(path + ": bad PNG filter");
            }
            cur[i] = (uint8_t)v;
        }
        std::memcpy(prev.data(), cur, stride);
    }
    GrayImage img;
    img.rows = (int)height;
    img.cols = (int)width;
    img.bits = depth == 16 ? 16 : 8;
    img.data.resize((size_t)width * height * (img.bits / 8));
    was_colour = ctype == 2 || ctype == 6 || ctype == 3;
    for (uint32_t y = 0; y < height; ++y) {
        const uint8_t* cur = &raw[(stride + 1) * y + 1];
        for (uint32_t x = 0; x < width; ++x) {
            uint32_t v;
            if (depth < 8) {
                const int per = 8 / depth;
                const int shift = (per - 1 - (int)(x % per)) * depth;
                const uint32_t idx = (cur[x / per] >> shift) & ((1u << depth) - 1);
                if (ctype == 3) {
                    if (3 * idx + 2 >= palette.size())
                        throw std::runtime_error(path + ": palette index out of range");
                    v = luma(palette[3 * idx], palette[3 * idx + 1], palette[3 * idx + 2]);
                } else {
                    v = idx * 255u / ((1u << depth) - 1);
                }
            } else if (depth == 8) {
                const uint8_t* p = cur + (size_t)x * channels;
                if (ctype == 3) {
                    if (3u * p[0] + 2 >= palette.size())
                        throw std::runtime_error(path + ": palette index out of range");
                    v = luma(palette[3 * p[0]], palette[3 * p[0] + 1], palette[3 * p[0] + 2]);
                } else {
                    v = channels >= 3 ? luma(p[0], p[1], p[2]) : p[0];
                }
            } else {
                const uint8_t* p = cur + (size_t)x * channels * 2;
                auto s = [&](int c) { return ((uint32_t)p[2 * c] << 8) | p[2 * c + 1]; };
                v = channels >= 3 ? (uint32_t)std::lround(0.299 * s(0) + 0.587 * s(1) + 0.114 * s(2)) : s(0);
            }
            if (img.bits == 16) {
                const uint16_t w = (uint16_t)v;
                std::memcpy(&img.data[((size_t)y * width + x) * 2], &w, 2);
            } else {
                img.data[(size_t)y * width + x] = (uint8_t)v;
            }
        }
    }
    return img;
}

GrayImage read_pgm(const std::vector<uint8_t>& file, const std::string& path) {
    size_t pos = 2;
    auto next_int = [&]() -> int {
        for (;;) {
            while (pos < file.size() && std::isspace(file[pos]))
                ++pos;
            if (pos < file.size() && file[pos] == '#') {
                while (pos < file.size() && file[pos] != '\n')
                    ++pos;
                continue;
            }
            break;
        }
        int v = 0;
        bool any = false;
        while (pos < file.size() && std::isdigit(file[pos])) {
            v = v * 10 + (file[pos++] - '0');
            any = true;
        }
        if (!any)
            throw std::runtime_error(path + ": bad PGM header");
        return v;
    };
    const int width = next_int(), height = next_int(), maxval = next_int();
    ++pos; // single whitespace after maxval
    if (width <= 0 || height <= 0 || maxval <= 0 || maxval > 65535)
        throw std::runtime_error(path + ": bad PGM header");
    GrayImage img;
    img.rows = height;
    img.cols = width;
    img.bits = maxval > 255 ? 16 : 8;
    const size_t bytes = (size_t)width * height * (img.bits / 8);
    if (pos + bytes > file.size())
        throw std::runtime_error(path + ": truncated PGM");
    img.data.assign(file.begin() + pos, file.begin() + pos + bytes);
    if (img.bits == 16) // PGM samples are big-endian
        for (size_t i = 0; i + 1 < img.data.size(); i += 2)
            std::swap(img.data[i], img.data[i + 1]);
    return img;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(x >> 24);
    v.push_back(x >> 16);
    v.push_back(x >> 8);
    v.push_back(x);
}

void png_chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& body) {
    put_be32(out, (uint32_t)body.size());
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), body.begin(), body.end());
    put_be32(out, (uint32_t)crc32(0, &out[start], (uInt)(out.size() - start)));
}

} // namespace

GrayImage read_image(const std::string& path, bool& was_colour) {
    was_colour = false;
    const std::vector<uint8_t> file = slurp(path);
    if (file.size() >= 2 && file[0] == 'P' && file[1] == '5')
        return read_pgm(file, path);
    return read_png(file, path, was_colour);
}

#include "colormaps.inc"

void colormap_lut(Colormap map, uint8_t lut[256][3]) {
    std::memcpy(lut, map == Colormap::TURBO ? LUT_TURBO : LUT_VIRIDIS, 256 * 3);
}

void write_png_rgb(const std::string& path, int rows, int cols, const std::vector<uint8_t>& rgb) {
    std::vector<uint8_t> raw((size_t)rows * (3 * (size_t)cols + 1));
    for (int y = 0; y < rows; ++y) {
        raw[(size_t)y * (3 * cols + 1)] = 0; // filter: none
        std::memcpy(&raw[(size_t)y * (3 * cols + 1) + 1], &rgb[(size_t)y * 3 * cols], 3 * (size_t)cols);
    }
    uLongf zlen = compressBound(raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), raw.size(), 6) != Z_OK)
        throw std::runtime_error("deflate failed for " + path);
    z.resize(zlen);
    std::vector<uint8_t> out = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)cols);
    put_be32(ihdr, (uint32_t)rows);
    ihdr.insert(ihdr.end(), { 8, 2, 0, 0, 0 }); // 8-bit RGB, deflate, adaptive filtering, no interlace
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    if (!f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size()))
        throw std::runtime_error("cannot write " + path);
}

void write_tiff(const std::string& path, int rows, int cols, int bits, int sample_format, const void* data) {
    // little-endian baseline TIFF, one uncompressed strip
    const uint32_t nbytes = (uint32_t)((size_t)rows * cols * (bits / 8));
    std::vector<uint8_t> out;
    auto u16 = [&](uint16_t v) {
        out.push_back(v & 0xFF);
        out.push_back(v >> 8);
    };
    auto u32 = [&](uint32_t v) {
        for (int i = 0; i < 4; ++i)
            out.push_back((v >> (8 * i)) & 0xFF);
    };
    out.insert(out.end(), { 'I', 'I' });
    u16(42);
    u32(8 + nbytes + (nbytes & 1)); // IFD after the pixel data
    out.insert(out.end(), static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + nbytes);
    if (nbytes & 1)
        out.push_back(0);
    struct Entry {
        uint16_t tag, type;
        uint32_t count, value;
    };
    const Entry entries[] = {
        { 256, 4, 1, (uint32_t)cols }, // ImageWidth
        { 257, 4, 1, (uint32_t)rows }, // ImageLength
        { 258, 3, 1, (uint32_t)bits }, // BitsPerSample
        { 259, 3, 1, 1 }, // Compression: none
        { 262, 3, 1, 1 }, // Photometric: BlackIsZero
        { 273, 4, 1, 8 }, // StripOffsets
        { 277, 3, 1, 1 }, // SamplesPerPixel
        { 278, 4, 1, (uint32_t)rows }, // RowsPerStrip
        { 279, 4, 1, nbytes }, // StripByteCounts
        { 339, 3, 1, (uint32_t)sample_format }, // SampleFormat
    };
    u16(sizeof entries / sizeof entries[0]);
    for (const Entry& e: entries) {
        u16(e.tag);
        u16(e.type);
        u32(e.count);
        if (e.type == 3) {
            u16((uint16_t)e.value);
            u16(0);
        } else {
            u32(e.value);
        }
    }
    u32(0);
    std::ofstream f(path, std::ios::binary);
    if (!f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size()))
        throw std::runtime_error("cannot write " + path);
}

std::vector<uint8_t> colorize(const void* image, int type, int rows, int cols, Colormap map) {
    const size_t n = (size_t)rows * cols;
    auto value = [&](size_t i, bool& valid) -> double {
        if (type == 3) {
            const int16_t v = static_cast<const int16_t*>(image)[i];
            valid = v != std::numeric_limits<int16_t>::lowest();
            return v;
        }
        const double v = type == 5 ? (double)static_cast<const float*>(image)[i] : static_cast<const double*>(image)[i];
        valid = v == v;
        return v;
    };
    double lo = std::numeric_limits<double>::infinity(), hi = -lo;
    for (size_t i = 0; i < n; ++i) {
        bool ok;
        const double v = value(i, ok);
        if (ok) {
            lo = std::fmin(lo, v);
            hi = std::fmax(hi, v);
        }
    }
    const double scale = hi > lo ? 255.0 / (hi - lo) : 0.0;
    uint8_t lut[256][3];
    colormap_lut(map, lut);
    std::vector<uint8_t> rgb(n * 3, 0);
    for (size_t i = 0; i < n; ++i) {
        bool ok;
        const double v = value(i, ok);
        if (!ok)
            continue; // invalid pixels stay black
        const int g = (int)std::fmin(255.0, std::fmax(0.0, std::nearbyint((v - lo) * scale)));
        std::memcpy(&rgb[3 * i], lut[g], 3);
    }
    return rgb;
}

bool read_q_matrix(const std::string& path, double q[16], std::string& error) {
    std::ifstream f(path);
    if (!f) {
        error = "cannot open " + path;
        return false;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    // YAML "Q: !!opencv-matrix ... data: [ ... ]" or XML "<Q ...> ... <data> ... </data>"
    size_t at = text.find("Q:");
    if (at == std::string::npos)
        at = text.find("<Q");
    if (at == std::string::npos) {
        error = "no matrix named Q in " + path;
        return false;
    }
    at = text.find("data", at);
    if (at == std::string::npos) {
        error = "matrix Q has no data in " + path;
        return false;
    }
    at += 4;
    int count = 0;
    const char* p = text.c_str() + at;
    while (*p && count < 16) {
        if (*p == ']' || (*p == '<' && p[1] == '/'))
            break;
        if (std::isdigit((unsigned char)*p) || *p == '-' || *p == '+' || (*p == '.' && std::isdigit((unsigned char)p[1]))) {
            char* end = nullptr;
            q[count++] = std::strtod(p, &end);
            p = end;
        } else {
            ++p;
        }
    }
    if (count != 16) {
        error = "matrix Q in " + path + " does not have 16 elements";
        return false;
    }
    return true;
}

} // namespace bicos_cli
