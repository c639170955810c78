/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the
 * product path (libbicos_b200/). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Plain-C restatement of the reference's CPU implementation of BICOS::match
 * (descriptor transform -> row-wise Hamming search -> NXC refine -> postfilter).
 * Citations are relative to /root/reference. It is written from the algorithm, not
 * from the reference's code structure: pixels are widened to uint16 planes once and
 * every descriptor is an array of K little-endian 32-bit words.
 *
 * PARITY PINNING: float paths are pinned bit-for-bit against the unmodified reference
 * (oracle/_ref/libbicos_ref.so, built by oracle/Makefile from the reference sources)
 * by tests/test_oracle.py, and against tests/golden/ *.npz generated from that build.
 * The *_f64 paths restate the reference's CUDA-only nxcorrd (include/impl/cuda/agree.cuh:35-65)
 * inside the float path's (pinned) control flow; the reference has no CPU double
 * implementation to pin them against, so their arithmetic core is "parity unpinned".
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -pthread (no -mfma: fmaf/fma go through
 * libm, which is exactly-rounded; contraction would change subpixel results).
 */
#define _GNU_SOURCE
#include "bicos_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define CV_8U 0
#define CV_16U 2
#define CV_16S 3
#define CV_32F 5

#define FLAG_NODUPES 1      /* include/impl/common.hpp:46 */
#define FLAG_CONSISTENCY 2  /* include/impl/common.hpp:47 */
#define INVALID_I16 ((int16_t)-32768) /* include/common.hpp:34-37: lowest() for integers */
#define INVALID_COL INT_MIN           /* INVALID_DISP<int> */
#define MAXW 16 /* 8 words = the reference's 256 bits; 12 / 16 only with the wide extension */

static _Thread_local char g_error[256];
static int g_threads = 0;

const char* orc_last_error(void) {
    return g_error;
}

void orc_set_threads(int n) {
    g_threads = n;
}

int orc_hardware_threads(void) {
    const long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* row-parallel driver (the reference uses cv::parallel_for_ over rows in every stage):
 * rows are handed out one at a time from a shared counter to plain pthreads. */
typedef void (*row_fn)(int row, void* ctx);
typedef struct {
    row_fn fn;
    void* ctx;
    int rows;
    int next;
    pthread_mutex_t mu;
} row_job;

static void* row_worker(void* arg) {
    row_job* job = (row_job*)arg;
    for (;;) {
        pthread_mutex_lock(&job->mu);
        const int r = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (r >= job->rows)
            return NULL;
        job->fn(r, job->ctx);
    }
}

static void for_rows(int rows, row_fn fn, void* ctx) {
    int nt = g_threads > 0 ? g_threads : orc_hardware_threads();
    if (nt > rows)
        nt = rows;
    if (nt <= 1) {
        for (int r = 0; r < rows; ++r)
            fn(r, ctx);
        return;
    }
    if (nt > 256)
        nt = 256;
    row_job job = { fn, ctx, rows, 0, PTHREAD_MUTEX_INITIALIZER };
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nt; ++i)
        if (pthread_create(&th[started], NULL, row_worker, &job) == 0)
            started++;
    if (started == 0)
        row_worker(&job);
    for (int i = 0; i < started; ++i)
        pthread_join(th[i], NULL);
}

static int fail(const char* msg) {
    snprintf(g_error, sizeof g_error, "%s", msg);
    return -1;
}

/* planar [n][rows][cols] u8/u16 -> planar u16 (value-preserving) */
static uint16_t* widen(const void* stack, int n, int rows, int cols, int depth) {
    const size_t total = (size_t)n * rows * cols;
    uint16_t* out = (uint16_t*)malloc(total * sizeof(uint16_t) + 2);
    if (!out)
        return NULL;
    if (depth == CV_8U) {
        const uint8_t* s = (const uint8_t*)stack;
        for (size_t i = 0; i < total; ++i)
            out[i] = s[i];
    } else {
        memcpy(out, stack, total * sizeof(uint16_t));
    }
    return out;
}

/* src/impl/cpu.cpp:122-156: required_bits (LIMITED undercounts by one, harmless) -> word count.
 * `mode`: bit 0 = TransformMode::FULL; bit 1 = the wide-descriptor extension of this repository
 * (384 / 512 bits, FULL stacks of 17..23 images; the reference throws above 256 bits, :153-155). */
static int words_for(int n, int mode, int* bits_out) {
    const int mode_full = mode & 1;
    const int bits = mode_full ? n * n - 2 * n + 3 : 4 * n - 7;
    if (bits_out)
        *bits_out = bits;
    if (bits <= 32)
        return 1;
    if (bits <= 64)
        return 2;
    if (bits <= 128)
        return 4;
    if (bits <= 256)
        return 8;
    if ((mode & 2) && bits <= 384)
        return 12;
    if ((mode & 2) && bits <= 512)
        return 16;
    return -1;
}

/* ---------------------------------------------------------------- descriptors -- */

typedef struct {
    uint32_t w[MAXW];
    unsigned i;
} bitacc;

/* include/impl/cpu/bitfield.hpp:39-57: LSB-first append; bit i -> word i/32, bit i%32 */
static inline void put(bitacc* b, int value) {
    if (value && b->i < 32u * MAXW)
        b->w[b->i >> 5] |= 1u << (b->i & 31);
    b->i++;
}

/* include/impl/cpu/descriptor_transform.hpp:31-73 (transform_limited) */
static void describe_limited(const uint16_t* p, int n, bitacc* b) {
    float av = 0.0f;
    for (int t = 0; t < n; ++t)
        av += (float)p[t];
    av /= (float)n;

    for (int t = 0; t + 2 < n; ++t) {
        put(b, p[t] < p[t + 1]);
        put(b, p[t] < p[t + 2]);
        put(b, (float)p[t] < av);
        if (t >= 2) /* pair sum two steps back exists */
            put(b, (int)p[t - 2] + (int)p[t - 1] < (int)p[t] + (int)p[t + 1]);
    }
    const int a = p[n - 2], c = p[n - 1];
    put(b, a < c);
    put(b, (float)a < av);
    put(b, (float)c < av);
    /* previous pair sum of the same parity: ps(n-4) if it exists, else -1 */
    const int prev = n >= 4 ? (int)p[n - 4] + (int)p[n - 3] : -1;
    put(b, prev < a + c);
}

/* include/impl/cpu/descriptor_transform.hpp:75-123 (transform_full) */
static void describe_full(const uint16_t* p, int n, bitacc* b) {
    float av = 0.0f;
    for (int t = 0; t < n; ++t)
        av += (float)p[t];
    av /= (float)n;

    uint32_t ps[128];
    for (int t = 0; t + 2 < n; ++t) {
        put(b, p[t] < p[t + 1]);
        put(b, p[t] < p[t + 2]);
        put(b, (float)p[t] < av);
    }
    for (int t = 0; t + 1 < n; ++t)
        ps[t] = (uint32_t)p[t] + (uint32_t)p[t + 1];
    put(b, p[n - 2] < p[n - 1]);
    put(b, (float)p[n - 2] < av);
    put(b, (float)p[n - 1] < av);
    for (int t = 0; t + 1 < n; ++t)
        for (int i = 0; i + 1 < n; ++i) {
            if (i == t || i == t - 1 || i == t + 1)
                continue;
            put(b, ps[t] < ps[i]);
        }
}

/* include/impl/cpu/descriptor_transform.hpp:125-138: one descriptor per pixel, [rows][cols][K] */
typedef struct {
    const uint16_t* st;
    int n, rows, cols, mode_full, K;
    uint32_t* out;
} desc_ctx;

static void descriptors_row(int r, void* vctx) {
    const desc_ctx* x = (const desc_ctx*)vctx;
    const size_t plane = (size_t)x->rows * x->cols;
    uint16_t pix[128];
    for (int c = 0; c < x->cols; ++c) {
        const size_t at = (size_t)r * x->cols + c;
        for (int t = 0; t < x->n; ++t)
            pix[t] = x->st[plane * t + at];
        bitacc b;
        memset(&b, 0, sizeof b);
        if (x->mode_full)
            describe_full(pix, x->n, &b);
        else
            describe_limited(pix, x->n, &b);
        memcpy(x->out + at * x->K, b.w, sizeof(uint32_t) * x->K);
    }
}

static void descriptors_of(const uint16_t* st, int n, int rows, int cols, int mode_full, int K,
                           uint32_t* out) {
    desc_ctx x = { st, n, rows, cols, mode_full & 1, K, out }; /* bit 1 only widens the dispatch */
    for_rows(rows, descriptors_row, &x);
}

/* --------------------------------------------------------------------- search -- */

/* include/impl/cpu/bicos.hpp:29-48 (ham) */
static inline int hamming(const uint32_t* a, const uint32_t* b, int K) {
    int s = 0;
    for (int k = 0; k < K; ++k)
        s += __builtin_popcount(a[k] ^ b[k]);
    return s;
}

/* include/impl/cpu/bicos.hpp:50-76 (bicos_search): first strict minimum over the whole
 * row; with NODUPES a later tie with the current minimum invalidates the result. */
static int search_row(const uint32_t* d, const uint32_t* row, int cols, int K, int nodupes) {
    int best = INVALID_COL, min_cost = INT_MAX, ties = 0;
    for (int c = 0; c < cols; ++c) {
        const int cost = hamming(d, row + (size_t)c * K, K);
        if (cost < min_cost) {
            min_cost = cost;
            best = c;
            ties = 0;
        } else if (cost == min_cost) {
            ties++;
        }
    }
    if (nodupes && ties > 0)
        return INVALID_COL;
    return best;
}

/* include/impl/cpu/bicos.hpp:78-113 (bicos) */
typedef struct {
    const uint32_t *d0, *d1;
    int K, cols, flags, max_lr_diff;
    int16_t* out;
} bicos_ctx;

static void bicos_row(int r, void* vctx) {
    const bicos_ctx* x = (const bicos_ctx*)vctx;
    const int cols = x->cols, K = x->K;
    const int nodupes = (x->flags & FLAG_NODUPES) != 0;
    const uint32_t* row0 = x->d0 + (size_t)r * cols * K;
    const uint32_t* row1 = x->d1 + (size_t)r * cols * K;
    int16_t* o = x->out + (size_t)r * cols;
    for (int c0 = 0; c0 < cols; ++c0) {
        o[c0] = INVALID_I16;
        const int best = search_row(row0 + (size_t)c0 * K, row1, cols, K, nodupes);
        if (best == INVALID_COL)
            continue;
        if (x->flags & FLAG_CONSISTENCY) {
            const int rev = search_row(row1 + (size_t)best * K, row0, cols, K, nodupes);
            if (rev == INVALID_COL || abs(c0 - rev) > x->max_lr_diff)
                continue;
            o[c0] = (int16_t)((c0 + rev) / 2 - best);
        } else {
            o[c0] = (int16_t)(c0 - best);
        }
    }
}

static void bicos_rows(const uint32_t* d0, const uint32_t* d1, int K, int rows, int cols, int flags,
                       int max_lr_diff, int16_t* out) {
    bicos_ctx x = { d0, d1, K, cols, flags, max_lr_diff, out };
    for_rows(rows, bicos_row, &x);
}

/* --------------------------------------------------------------------- refine -- */

/* include/impl/cpu/agree.hpp:28-51 (nxcorr<T>): sequential float mean, fmaf chains,
 * min-variance test before the division. has_minvar == 0 <=> std::nullopt. */
static float nxcorr_f32(const uint16_t* p0, const uint16_t* p1, int n, int has_minvar, float minvar) {
    float mean0 = 0.f, mean1 = 0.f;
    for (int i = 0; i < n; ++i) {
        mean0 += (float)p0[i];
        mean1 += (float)p1[i];
    }
    mean0 /= (float)n;
    mean1 /= (float)n;
    float covar = 0.f, var0 = 0.f, var1 = 0.f;
    for (int i = 0; i < n; ++i) {
        const float diff0 = (float)p0[i] - mean0, diff1 = (float)p1[i] - mean1;
        covar = fmaf(diff0, diff1, covar);
        var0 = fmaf(diff0, diff0, var0);
        var1 = fmaf(diff1, diff1, var1);
    }
    if (has_minvar && (var0 < minvar || var1 < minvar))
        return -1.f;
    return covar / sqrtf(var0 * var1);
}

/* include/impl/cuda/agree.cuh:35-65 (nxcorrd): the same in double (CUDA backend only) */
static double nxcorr_f64(const uint16_t* p0, const uint16_t* p1, int n, int has_minvar, double minvar) {
    double mean0 = 0.0, mean1 = 0.0;
    for (int i = 0; i < n; ++i) {
        mean0 += (double)p0[i];
        mean1 += (double)p1[i];
    }
    mean0 /= (double)n;
    mean1 /= (double)n;
    double covar = 0.0, var0 = 0.0, var1 = 0.0;
    for (int i = 0; i < n; ++i) {
        const double diff0 = (double)p0[i] - mean0, diff1 = (double)p1[i] - mean1;
        covar = fma(diff0, diff1, covar);
        var0 = fma(diff0, diff0, var0);
        var1 = fma(diff1, diff1, var1);
    }
    if (has_minvar && (var0 < minvar || var1 < minvar))
        return -1.0;
    return covar / sqrt(var0 * var1);
}

static inline double nxc_any(const uint16_t* p0, const uint16_t* p1, int n, int has_minvar,
                             float minvar, int dbl) {
    /* double mode: minvar/threshold are float values widened (src/impl/cuda.cu.in:225) */
    return dbl ? nxcorr_f64(p0, p1, n, has_minvar, (double)minvar)
               : (double)nxcorr_f32(p0, p1, n, has_minvar, minvar);
}

static inline void gather(const uint16_t* st, size_t plane, int n, size_t at, uint16_t* out) {
    for (int t = 0; t < n; ++t)
        out[t] = st[plane * t + at];
}

static inline void store_corr(void* corr, int dbl, size_t at, double v) {
    if (!corr)
        return;
    if (dbl)
        ((double*)corr)[at] = v;
    else
        ((float*)corr)[at] = (float)v;
}

/* include/impl/cpu/agree.hpp:53-93 (agree): in place on the int16 disparity.
 * Comparison against the threshold happens in the precision of the NXC value
 * (float: agree.hpp:87; double: cuda/agree.cuh:157 `(TPrecision)min_nxc`). */
typedef struct {
    int16_t* disp_rw;
    const int16_t* disp;
    const uint16_t *s0, *s1;
    int n, rows, cols;
    float thr, step;
    int has_minvar;
    float minvar;
    unsigned wrap;
    float* out;
    void* corr;
    int dbl;
} agree_ctx;

static void agree_row(int r, void* vctx) {
    const agree_ctx* x = (const agree_ctx*)vctx;
    int16_t* disp = x->disp_rw;
    const uint16_t *s0 = x->s0, *s1 = x->s1;
    const int n = x->n, cols = x->cols, has_minvar = x->has_minvar, dbl = x->dbl;
    const float thr = x->thr, minvar = x->minvar;
    void* corr = x->corr;
    const size_t plane = (size_t)x->rows * cols;
    {
        uint16_t a[128], b[128];
        for (int c = 0; c < cols; ++c) {
            const size_t at = (size_t)r * cols + c;
            const int16_t d = disp[at];
            if (d == INVALID_I16)
                continue;
            const int c1 = c - d;
            if (c1 < 0 || cols <= c1) {
                disp[at] = INVALID_I16;
                continue;
            }
            gather(s0, plane, n, at, a);
            gather(s1, plane, n, (size_t)r * cols + c1, b);
            const double nxc = nxc_any(a, b, n, has_minvar, minvar, dbl);
            store_corr(corr, dbl, at, nxc);
            if (dbl ? (nxc < (double)thr) : ((float)nxc < thr))
                disp[at] = INVALID_I16;
        }
    }
}

static void agree_rows(int16_t* disp, const uint16_t* s0, const uint16_t* s1, int n, int rows,
                       int cols, float thr, int has_minvar, float minvar, void* corr, int dbl) {
    agree_ctx x = { disp, disp, s0, s1, n, rows, cols, thr, -1.f, has_minvar, minvar, 0, NULL, corr, dbl };
    for_rows(rows, agree_row, &x);
}

/* include/impl/cpu/agree.hpp:95-191 (agree_subpixel). `wrap` = 0xFF / 0xFFFF: the
 * interpolated value is converted to the input type modulo 2^bits (x86 cvttss2si +
 * truncation; same as `(TInput)__float2int_rn` in cuda/agree.cuh:235). */
static void agree_subpixel_row(int r, void* vctx) {
    const agree_ctx* x_ = (const agree_ctx*)vctx;
    const int16_t* disp = x_->disp;
    const uint16_t *s0 = x_->s0, *s1 = x_->s1;
    const int n = x_->n, cols = x_->cols, has_minvar = x_->has_minvar, dbl = x_->dbl;
    const float thr = x_->thr, minvar = x_->minvar, step = x_->step;
    const unsigned wrap = x_->wrap;
    float* out = x_->out;
    void* corr = x_->corr;
    const size_t plane = (size_t)x_->rows * cols;
    {
        uint16_t p0[128], y0[128], y1[128], y2[128], iv[128];
        float qa[128], qb[128], qc[128];
        for (int c = 0; c < cols; ++c) {
            const size_t at = (size_t)r * cols + c;
            out[at] = NAN; /* agree.hpp:109-110 */
            const int16_t d = disp[at];
            if (d == INVALID_I16)
                continue;
            const int c1 = c - d;
            if (c1 < 0 || cols <= c1)
                continue;
            gather(s0, plane, n, at, p0);
            gather(s1, plane, n, (size_t)r * cols + c1, y1);
            if (c1 == 0 || c1 == cols - 1) {
                const double nxc = nxc_any(p0, y1, n, has_minvar, minvar, dbl);
                store_corr(corr, dbl, at, nxc);
                if (dbl ? (nxc < (double)thr) : ((float)nxc < thr))
                    continue;
                out[at] = (float)d;
                continue;
            }
            gather(s1, plane, n, (size_t)r * cols + c1 - 1, y0);
            gather(s1, plane, n, (size_t)r * cols + c1 + 1, y2);
            for (int t = 0; t < n; ++t) {
                qa[t] = 0.5f * (((float)y0[t] - 2.0f * (float)y1[t]) + (float)y2[t]);
                qb[t] = 0.5f * (float)(-(int)y0[t] + (int)y2[t]);
                qc[t] = (float)y1[t];
            }
            float best_x = 0.f;
            double best = -1.0;
            for (float x = -1.f; x <= 1.f; x += step) {
                for (int t = 0; t < n; ++t) {
                    /* five separately rounded operations, left to right */
                    const float v = ((qa[t] * x) * x + qb[t] * x) + qc[t];
                    iv[t] = (uint16_t)((uint32_t)(int32_t)roundevenf(v) & wrap);
                }
                const double nxc = nxc_any(p0, iv, n, has_minvar, minvar, dbl);
                if (best < nxc) { /* strict: first maximum wins, NaN never wins */
                    best_x = x;
                    best = nxc;
                }
            }
            store_corr(corr, dbl, at, best);
            if (dbl ? (best < (double)thr) : ((float)best < thr))
                continue;
            out[at] = (float)d - best_x;
        }
    }
}

static void agree_subpixel_rows(const int16_t* disp, const uint16_t* s0, const uint16_t* s1, int n,
                                int rows, int cols, float thr, float step, int has_minvar,
                                float minvar, unsigned wrap, float* out, void* corr, int dbl) {
    agree_ctx x = { NULL, disp, s0, s1, n, rows, cols, thr, step, has_minvar, minvar, wrap, out, corr, dbl };
    for_rows(rows, agree_subpixel_row, &x);
}

/* ---------------------------------------------------------------- entry points -- */

int orc_descriptors(const void* stack, int n, int rows, int cols, int depth, int mode_full,
                    uint32_t* out_words, int cap) {
    if (n < 2 || n > 128)
        return fail("need 2..128 images");
    const int K = words_for(n, mode_full, NULL);
    if (K < 0)
        return fail("input stacks too large");
    if (K > cap)
        return fail("output capacity too small");
    uint16_t* st = widen(stack, n, rows, cols, depth);
    if (!st)
        return fail("out of memory");
    descriptors_of(st, n, rows, cols, mode_full, K, out_words);
    free(st);
    return K;
}

int orc_bicos(const uint32_t* desc0, const uint32_t* desc1, int K, int rows, int cols, int flags,
              int max_lr_diff, int16_t* out) {
    if (K != 1 && K != 2 && K != 4 && K != 8 && K != 12 && K != 16)
        return fail("bad K");
    if (flags < 1 || flags > 3)
        return fail("bad flags");
    bicos_rows(desc0, desc1, K, rows, cols, flags, max_lr_diff, out);
    return 0;
}

static int agree_any(const int16_t* raw, const void* stack0, const void* stack1, int n, int rows,
                     int cols, int depth, float thr, float step, float minvar_n, int16_t* di,
                     float* df, void* corr, int dbl) {
    if (n < 2 || n > 128)
        return fail("need 2..128 images");
    if (step == 0.0f)
        return fail("subpixel_step must be positive");
    uint16_t* s0 = widen(stack0, n, rows, cols, depth);
    uint16_t* s1 = widen(stack1, n, rows, cols, depth);
    if (!s0 || !s1) {
        free(s0);
        free(s1);
        return fail("out of memory");
    }
    const size_t px = (size_t)rows * cols;
    if (corr) {
        if (dbl)
            for (size_t i = 0; i < px; ++i)
                ((double*)corr)[i] = NAN;
        else
            for (size_t i = 0; i < px; ++i)
                ((float*)corr)[i] = NAN;
    }
    const int has_mv = minvar_n >= 0;
    if (step < 0) {
        memcpy(di, raw, px * sizeof(int16_t));
        agree_rows(di, s0, s1, n, rows, cols, thr, has_mv, minvar_n, corr, dbl);
    } else {
        agree_subpixel_rows(raw, s0, s1, n, rows, cols, thr, step, has_mv, minvar_n,
                            depth == CV_8U ? 0xFFu : 0xFFFFu, df, corr, dbl);
    }
    free(s0);
    free(s1);
    return 0;
}

int orc_agree(const int16_t* raw, const void* stack0, const void* stack1, int n, int rows, int cols,
              int depth, float thr, float step, float minvar_n, int16_t* di, float* df,
              float* corr) {
    return agree_any(raw, stack0, stack1, n, rows, cols, depth, thr, step, minvar_n, di, df, corr, 0);
}

int orc_agree_f64(const int16_t* raw, const void* stack0, const void* stack1, int n, int rows,
                  int cols, int depth, float thr, float step, float minvar_n, int16_t* di,
                  float* df, double* corr) {
    return agree_any(raw, stack0, stack1, n, rows, cols, depth, thr, step, minvar_n, di, df, corr, 1);
}

/* src/impl/cpu.cpp:100-159 (match) + :35-98 (match_impl) */
static int match_any(const void* stack0, const void* stack1, int n, int rows, int cols, int depth,
                     float thr, float step, float min_variance, int mode_full, int consistency,
                     int max_lr_diff, int no_dupes, void* disp_out, int* disp_type, void* corr,
                     int dbl) {
    if (n < 2)
        return fail("need at least two images");
    if (depth != CV_8U && depth != CV_16U)
        return fail("bad input depths, only CV_8UC1 and CV_16UC1 are supported");
    if (n > 128)
        return fail("input stacks too large");
    int bits;
    const int K = words_for(n, mode_full, &bits);
    if (K < 0) {
        snprintf(g_error, sizeof g_error, "input stacks too large, would require %d bits", bits);
        return -1;
    }
    if (step == 0.0f)
        return fail("subpixel_step must be positive");
    const size_t px = (size_t)rows * cols;
    uint16_t* s0 = widen(stack0, n, rows, cols, depth);
    uint16_t* s1 = widen(stack1, n, rows, cols, depth);
    uint32_t* d0 = (uint32_t*)malloc(px * K * sizeof(uint32_t));
    uint32_t* d1 = (uint32_t*)malloc(px * K * sizeof(uint32_t));
    int16_t* raw = (int16_t*)malloc(px * sizeof(int16_t));
    int rc = 0;
    if (!s0 || !s1 || !d0 || !d1 || !raw) {
        rc = fail("out of memory");
        goto done;
    }
    descriptors_of(s0, n, rows, cols, mode_full, K, d0);
    descriptors_of(s1, n, rows, cols, mode_full, K, d1);

    /* cpu.cpp:68-75 */
    const int flags = consistency ? (FLAG_CONSISTENCY | (no_dupes ? FLAG_NODUPES : 0)) : FLAG_NODUPES;
    bicos_rows(d0, d1, K, rows, cols, flags, consistency ? max_lr_diff : -1, raw);

    if (thr < 0) { /* no threshold: int16 result, corrmap untouched (cpu.cpp:77) */
        memcpy(disp_out, raw, px * sizeof(int16_t));
        *disp_type = CV_16S;
        goto done;
    }
    /* cpu.cpp:127: min_var = min_variance * n (float) */
    const int has_mv = min_variance >= 0;
    const float minvar_n = has_mv ? min_variance * (float)n : 0.f;
    if (corr) { /* cpu.cpp:78-81 */
        if (dbl)
            for (size_t i = 0; i < px; ++i)
                ((double*)corr)[i] = NAN;
        else
            for (size_t i = 0; i < px; ++i)
                ((float*)corr)[i] = NAN;
    }
    float* outf = (float*)disp_out;
    *disp_type = CV_32F;
    if (step >= 0) {
        agree_subpixel_rows(raw, s0, s1, n, rows, cols, thr, step, has_mv, minvar_n,
                            depth == CV_8U ? 0xFFu : 0xFFFFu, outf, corr, dbl);
    } else {
        agree_rows(raw, s0, s1, n, rows, cols, thr, has_mv, minvar_n, corr, dbl);
        for (size_t i = 0; i < px; ++i) /* cpu.cpp:88-94: convertTo keeps -32768 as -32768.0f */
            outf[i] = (float)raw[i];
    }
done:
    free(s0);
    free(s1);
    free(d0);
    free(d1);
    free(raw);
    return rc;
}

int orc_match(const void* stack0, const void* stack1, int n, int rows, int cols, int depth, float thr,
              float step, float min_variance, int mode_full, int consistency, int max_lr_diff,
              int no_dupes, void* disp_out, int* disp_type, float* corr_out) {
    return match_any(stack0, stack1, n, rows, cols, depth, thr, step, min_variance, mode_full,
                     consistency, max_lr_diff, no_dupes, disp_out, disp_type, corr_out, 0);
}

int orc_match_f64(const void* stack0, const void* stack1, int n, int rows, int cols, int depth,
                  float thr, float step, float min_variance, int mode_full, int consistency,
                  int max_lr_diff, int no_dupes, void* disp_out, int* disp_type, double* corr_out) {
    return match_any(stack0, stack1, n, rows, cols, depth, thr, step, min_variance, mode_full,
                     consistency, max_lr_diff, no_dupes, disp_out, disp_type, corr_out, 1);
}
