// The Python-facing C ABI (include/pybicos_c.h): host images in, malloc'ed host results out.
// Behavioural twin of the reference's src/pybicos_c.cpp:92-209, on bicos_b200_match_host().

#include "../../include/pybicos_c.h"
#include "../../include/bicos_b200.h"

#include <cmath>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local std::string t_error;
std::mutex g_mutex; // one shared workspace per process; matches are serialised through it
bicos_b200_handle g_handle = nullptr;

BicosResult* fail(const std::string& msg) {
    t_error = msg;
    return nullptr;
}

} // namespace

// Result images are tens of megabytes: a fresh malloc is a fresh mmap whose pages fault in while
// the result is copied out, and free() unmaps them again. The Python side copies and frees every
// result right away (pybicos/__init__.py:237-243), so the last few buffers are kept for reuse.
namespace {
struct CachedBuffer {
    void* ptr;
    size_t bytes;
};
std::mutex g_cache_mutex;
std::vector<CachedBuffer> g_cache; // free buffers
std::vector<CachedBuffer> g_live; // handed out, with their sizes
constexpr size_t CACHE_SLOTS = 4;

void* result_alloc(size_t bytes) {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    void* p = nullptr;
    for (size_t i = 0; i < g_cache.size(); ++i)
        if (g_cache[i].bytes == bytes) {
            p = g_cache[i].ptr;
            g_cache.erase(g_cache.begin() + (long)i);
            break;
        }
    if (!p)
        p = std::malloc(bytes);
    if (p)
        g_live.push_back({ p, bytes });
    return p;
}

void result_free(void* p) {
    if (!p)
        return;
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    size_t bytes = 0;
    for (size_t i = 0; i < g_live.size(); ++i)
        if (g_live[i].ptr == p) {
            bytes = g_live[i].bytes;
            g_live.erase(g_live.begin() + (long)i);
            break;
        }
    if (bytes >= (1u << 20) && g_cache.size() < CACHE_SLOTS)
        g_cache.push_back({ p, bytes });
    else
        std::free(p);
}
} // namespace

extern "C" {

const char* BICOS_LastError(void) {
    return t_error.c_str();
}

BicosConfig* BICOS_CreateDefaultConfig(void) {
    BicosConfig* c = new (std::nothrow) BicosConfig();
    if (!c)
        return nullptr;
    c->nxcorr_threshold = 0.5f;
    c->subpixel_step = -1.0f;
    c->min_variance = -1.0f;
    c->mode = 0;
    c->precision = 0;
    c->variant_type = 0;
    c->max_lr_diff = 1;
    c->no_dupes = 0;
    return c;
}

void BICOS_FreeConfig(BicosConfig* config) {
    delete config;
}

void BICOS_FreeResult(BicosResult* result) {
    if (!result)
        return;
    result_free(result->disparity_data);
    result_free(result->corrmap_data);
    delete result;
}

BicosResult* BICOS_Match(void** stack0_data, int* stack0_rows, int* stack0_cols, int* stack0_types,
                         int stack0_size, void** stack1_data, int* stack1_rows, int* stack1_cols,
                         int* stack1_types, int stack1_size, BicosConfig* config) {
    try {
        if (!config || !stack0_data || !stack1_data || !stack0_rows || !stack0_cols || !stack0_types
            || !stack1_rows || !stack1_cols || !stack1_types)
            return fail("null argument");
        if (stack0_size < 2 || stack1_size < 2)
            return fail("need at least two images");
        if (stack0_size != stack1_size)
            return fail("stacks differ in length");
        const int rows = stack0_rows[0], cols = stack0_cols[0], depth = stack0_types[0] & 7;
        for (int i = 0; i < stack0_size; ++i)
            if (stack0_rows[i] != rows || stack0_cols[i] != cols || (stack0_types[i] & 7) != depth
                || stack1_rows[i] != rows || stack1_cols[i] != cols || (stack1_types[i] & 7) != depth)
                return fail("images differ in size or type");

        bicos_b200_config cfg {};
        cfg.nxcorr_threshold = config->nxcorr_threshold;
        cfg.subpixel_step = config->subpixel_step;
        cfg.min_variance = config->min_variance;
        // enums as the reference converts them (src/pybicos_c.cpp:56-89): any non-zero value is the second
        // enumerator. In particular bit 1 of bicos_b200_config::mode (BICOS_B200_MODE_WIDE, an extension the
        // reference does not have) cannot be reached through this ABI.
        cfg.mode = config->mode != 0 ? 1 : 0;
        cfg.precision = config->precision != 0 ? 1 : 0;
        cfg.variant_type = config->variant_type != 0 ? 1 : 0;
        cfg.max_lr_diff = config->max_lr_diff;
        cfg.no_dupes = config->no_dupes;

        const int disp_type = bicos_b200_disparity_type(&cfg);
        const int corr_type = bicos_b200_corrmap_type(&cfg);
        const size_t px = (size_t)rows * cols;
        const size_t disp_bytes = px * (disp_type == BICOS_B200_16S ? 2 : 4);
        const size_t corr_bytes = corr_type == 0 ? 0 : px * (corr_type == BICOS_B200_64F ? 8 : 4);

        BicosResult* res = new (std::nothrow) BicosResult();
        if (!res)
            return fail("out of memory");
        res->disparity_data = result_alloc(disp_bytes ? disp_bytes : 1);
        res->corrmap_data = corr_bytes ? result_alloc(corr_bytes) : nullptr;
        if (!res->disparity_data || (corr_bytes && !res->corrmap_data)) {
            BICOS_FreeResult(res);
            return fail("out of memory");
        }

        int rc;
        {
            std::lock_guard<std::mutex> lock(g_mutex);
            if (!g_handle) {
                rc = bicos_b200_create(&g_handle, -1);
                if (rc != 0) {
                    BICOS_FreeResult(res);
                    return fail(bicos_b200_last_error());
                }
            }
            rc = bicos_b200_match_host(g_handle, stack0_data, stack1_data, stack0_size, rows, cols, depth, &cfg,
                                       res->disparity_data, res->corrmap_data);
        }
        if (rc != 0) {
            BICOS_FreeResult(res);
            return fail(bicos_b200_last_error());
        }
        res->disparity_rows = rows;
        res->disparity_cols = cols;
        res->disparity_type = disp_type;
        // like the reference, an unset threshold yields an empty corrmap (type 0, no data)
        res->corrmap_rows = corr_type ? rows : 0;
        res->corrmap_cols = corr_type ? cols : 0;
        res->corrmap_type = corr_type;
        return res;
    } catch (const std::exception& e) {
        return fail(e.what());
    } catch (...) {
        return fail("unknown error");
    }
}

float BICOS_InvalidDisparityFloat(void) {
    return std::numeric_limits<float>::quiet_NaN();
}

int16_t BICOS_InvalidDisparityInt16(void) {
    return std::numeric_limits<int16_t>::lowest();
}

} // extern "C"
