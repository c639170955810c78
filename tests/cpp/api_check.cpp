// C++ consumer of the drop-in API (include/BICOS/match.hpp), built with plain g++ and linked
// against libbicos_b200.so: what a user of the reference's `BICOS::match` compiles. Driven by
// tests/test_cpp_api.py, which writes the inputs, runs this binary on the GPU box and compares
// the outputs with the oracle.
//
//   api_check <in.bin> <out.bin> [sharded]
//
// in.bin : int32 n, rows, cols, depth(0=8U, 2=16U), mode (bit 0 FULL, bit 1 wide descriptors), precision, variant, max_lr_diff, no_dupes;
//          float32 nxcorr_threshold, subpixel_step, min_variance (negative = unset);
//          then stack0 and stack1 as dense [n][rows][cols] arrays
// out.bin: int32 disparity type, corrmap type (0 = none), rows, cols; then both arrays, dense
// "sharded" runs BICOS::match_sharded over every visible GPU instead of BICOS::match.
#include <BICOS/match.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

using namespace BICOS;

static std::vector<char> read_file(const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        std::perror(path);
        std::exit(2);
    }
    std::fseek(f, 0, SEEK_END);
    const long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)size);
    if (std::fread(buf.data(), 1, buf.size(), f) != buf.size())
        std::exit(2);
    std::fclose(f);
    return buf;
}

static int expect_errors() {
    // reference behaviour: BICOS::Exception for fewer than two images (src/impl/cpu.cpp:110-111),
    // std::invalid_argument above 256 descriptor bits (src/impl/cpu.cpp:154-155)
    Image one(8, 8, IMG_8U), disp;
    int ok = 0;
    try {
        match({ one }, { one }, disp);
    } catch (const Exception&) {
        ok += 1;
    }
    std::vector<Image> many(20, one);
    Config full;
    full.mode = TransformMode::FULL;
    try {
        match(many, many, disp, full);
    } catch (const std::invalid_argument&) {
        ok += 1;
    }
    return ok == 2 ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: api_check <in.bin> <out.bin> [sharded]\n");
        return 2;
    }
    const bool sharded = argc > 3 && std::strcmp(argv[3], "sharded") == 0;
    const std::vector<char> in = read_file(argv[1]);
    int32_t hdr[9];
    float opt[3];
    std::memcpy(hdr, in.data(), sizeof hdr);
    std::memcpy(opt, in.data() + sizeof hdr, sizeof opt);
    const int n = hdr[0], rows = hdr[1], cols = hdr[2], depth = hdr[3];
    const size_t eb = depth == IMG_16U ? 2 : 1, plane = (size_t)rows * cols * eb;
    const char* base = in.data() + sizeof hdr + sizeof opt;

    Config cfg;
    cfg.nxcorr_threshold = opt[0] >= 0 ? std::optional<float>(opt[0]) : std::nullopt;
    cfg.subpixel_step = opt[1] >= 0 ? std::optional<float>(opt[1]) : std::nullopt;
    cfg.min_variance = opt[2] >= 0 ? std::optional<float>(opt[2]) : std::nullopt;
    cfg.mode = (hdr[4] & 1) ? TransformMode::FULL : TransformMode::LIMITED;
    cfg.wide_descriptors = (hdr[4] & 2) != 0; // extension: 384 / 512-bit descriptors
    cfg.precision = hdr[5] ? Precision::DOUBLE : Precision::SINGLE;
    if (hdr[6])
        cfg.variant = Variant::Consistency { hdr[7], hdr[8] != 0 };

    try {
        if (expect_errors() != 0) {
            std::fprintf(stderr, "error behaviour differs from the reference\n");
            return 1;
        }
        std::vector<Image> s0, s1;
        for (int i = 0; i < n; ++i) {
            s0.emplace_back(HostImage(rows, cols, depth, const_cast<char*>(base) + plane * i));
            s1.emplace_back(HostImage(rows, cols, depth, const_cast<char*>(base) + plane * (n + i)));
        }
        Image disp, corr;
        if (sharded) {
            std::vector<int> devices;
            const char* env = std::getenv("API_CHECK_DEVICES");
            const int count = env ? std::atoi(env) : 1;
            for (int d = 0; d < count; ++d)
                devices.push_back(d);
            match_sharded(s0, s1, disp, devices, cfg, &corr);
        } else {
            match(s0, s1, disp, cfg, &corr);
        }

        std::vector<char> dbuf((size_t)rows * cols * disp.elemSize());
        disp.download(HostImage(rows, cols, disp.type(), dbuf.data()));
        std::vector<char> cbuf;
        if (!corr.empty()) {
            cbuf.resize((size_t)rows * cols * corr.elemSize());
            corr.download(HostImage(rows, cols, corr.type(), cbuf.data()));
        }
        FILE* f = std::fopen(argv[2], "wb");
        if (!f)
            return 2;
        const int32_t oh[4] = { disp.type(), corr.empty() ? 0 : corr.type(), rows, cols };
        std::fwrite(oh, sizeof oh, 1, f);
        std::fwrite(dbuf.data(), 1, dbuf.size(), f);
        std::fwrite(cbuf.data(), 1, cbuf.size(), f);
        std::fclose(f);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
