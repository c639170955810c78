// CPU-only driver for the CLI's dependency-free image I/O (libbicos_b200/csrc/cli/imageio.cpp),
// used by tests/test_cli.py to compare it with OpenCV's codecs without a GPU.
//   imageio_check read  <image> <out.raw>      int32 rows, cols, bits, was_colour + pixel data
//   imageio_check write <rows> <cols> <type> <in.raw> <prefix>   prefix.png (Turbo) + prefix.tiff
//   imageio_check q     <file>                 the 16 entries of matrix Q, one line
#include "../../libbicos_b200/csrc/cli/imageio.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

using namespace bicos_cli;

int main(int argc, char** argv) {
    try {
        if (argc == 4 && !std::strcmp(argv[1], "read")) {
            bool colour = false;
            const GrayImage img = read_image(argv[2], colour);
            std::ofstream f(argv[3], std::ios::binary);
            const int32_t hdr[4] = { img.rows, img.cols, img.bits, colour ? 1 : 0 };
            f.write(reinterpret_cast<const char*>(hdr), sizeof hdr);
            f.write(reinterpret_cast<const char*>(img.data.data()), (std::streamsize)img.data.size());
            return 0;
        }
        if (argc == 7 && !std::strcmp(argv[1], "write")) {
            const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]), type = std::atoi(argv[4]);
            const int bits = type == 3 ? 16 : type == 5 ? 32 : 64;
            std::ifstream f(argv[5], std::ios::binary);
            std::vector<char> data((size_t)rows * cols * (bits / 8));
            f.read(data.data(), (std::streamsize)data.size());
            const std::string prefix = argv[6];
            write_png_rgb(prefix + ".png", rows, cols, colorize(data.data(), type, rows, cols, Colormap::TURBO));
            write_tiff(prefix + ".tiff", rows, cols, bits, type == 3 ? 2 : 3, data.data());
            return 0;
        }
        if (argc == 3 && !std::strcmp(argv[1], "q")) {
            double q[16];
            std::string error;
            if (!read_q_matrix(argv[2], q, error)) {
                std::cerr << error << std::endl;
                return 1;
            }
            for (double v: q)
                std::printf("%.17g ", v);
            std::printf("\n");
            return 0;
        }
        std::cerr << "usage: imageio_check read|write|q ..." << std::endl;
        return 2;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
}
