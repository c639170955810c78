/* bicos_b200.h -- thin C ABI over the sm_100a kernels of the BICOS::match hot path.
 *
 * This is the drop-in boundary underneath the reference's public entry point
 *     void BICOS::match(const std::vector<Image>&, const std::vector<Image>&,
 *                       Image& disparity, Config, Image* corrmap [, Stream&])
 *     (reference include/match.hpp:31-41, defined src/lib.cpp:31-49)
 * and underneath its Python FFI (reference src/pybicos_c.cpp:92-209, declared for this
 * repository in include/pybicos_c.h). A maintainer of the reference would replace the body
 * of impl::cuda::match (reference src/impl/cuda.cu:465-524) by one call to
 * bicos_b200_match(); INTEGRATION.md shows that stub.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative bicos_b200_status and records a message retrievable with
 * bicos_b200_last_error() (thread-local). No exceptions cross this boundary. All device
 * pointers must belong to the device the handle was created on. Work is enqueued on the
 * given stream (a cudaStream_t passed as void*; NULL = the legacy default stream) and is
 * NOT synchronised unless stated. There is no CPU fallback: without a CUDA device every
 * compute entry point fails with BICOS_B200_ERR_CUDA.
 */
#ifndef BICOS_B200_H
#define BICOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BICOS_B200_MAX_IMAGES 65 /* 4n-6 <= 256 descriptor bits; reference src/impl/cuda.cu:107 */

typedef enum {
    BICOS_B200_OK = 0,
    BICOS_B200_ERR_INVALID = -1, /* BICOS::Exception / std::invalid_argument in the reference */
    BICOS_B200_ERR_CUDA = -2, /* assertCudaSuccess, reference include/impl/cuda/cutil.cuh:32-41 */
    BICOS_B200_ERR_NOMEM = -3
} bicos_b200_status;

/* OpenCV depth codes, as the reference's C ABI passes them (pybicos/__init__.py:79-83) */
#define BICOS_B200_8U 0
#define BICOS_B200_16U 2
#define BICOS_B200_16S 3
#define BICOS_B200_32F 5
#define BICOS_B200_64F 6

/* reference include/impl/common.hpp:46-47 */
#define BICOS_B200_FLAG_NODUPES 1
#define BICOS_B200_FLAG_CONSISTENCY 2
/* bicos_b200_search only, OR-ed into `flags`: the caller vouches that bit 32 K - 1 of every descriptor is zero
 * (true for everything bicos_b200_transform writes: 4n-6 and n^2-2n+3 are never multiples of 32). The
 * tensor-core engine then carries the tile column through that bit's operand byte (search_mma.cu, "CT"),
 * as bicos_b200_match does; a set top bit with this flag gives wrong matches. */
#define BICOS_B200_FLAG_TOP_BIT_FREE 4
/* bicos_b200_search only: bits 32 K - 1 AND 32 K - 2 of every descriptor are zero (also true for everything
 * bicos_b200_transform writes: 4n-6 <= 32K-2, and n^2-2n+3 mod 32 <= 27). With 128- or 256-bit descriptors and
 * flags = CONSISTENCY the tensor-core engine then takes both directions from ONE product (search_mma.cu,
 * search_mma3_kernel), as bicos_b200_match does. Implies TOP_BIT_FREE. */
#define BICOS_B200_FLAG_TOP2_BITS_FREE 8

/* Same fields, order and "negative float = unset" convention as the reference's BicosConfig
 * (src/pybicos_c.cpp:30-41 with BICOS_CUDA defined; pybicos/__init__.py:41-51), plus one
 * trailing extension field. Zero-initialise the struct. */
typedef struct {
    float nxcorr_threshold; /* < 0: no NXC stage, int16 result (Config::nxcorr_threshold = nullopt) */
    float subpixel_step; /* < 0: integer disparities */
    float min_variance; /* < 0: no variance test */
    int mode; /* 0 = TransformMode::LIMITED, 1 = FULL */
    int precision; /* 0 = Precision::SINGLE, 1 = DOUBLE */
    int variant_type; /* 0 = Variant::NoDuplicates, 1 = Variant::Consistency */
    int max_lr_diff; /* Consistency only */
    int no_dupes; /* Consistency only */
    /* extension, after the reference's fields: nonzero = nxcorr_threshold is a real threshold even
     * though it is negative (BICOS::Config{.nxcorr_threshold = -1.0f}: "evaluate the NXC, reject
     * nothing", what the reference CLI uses for --corrmap without --threshold, cli.cpp:150-153) */
    int negative_threshold_is_set;
    /* extension: nonzero = descriptors of 384 and 512 bits (12 / 16 words) are allowed, i.e. FULL
     * stacks of 17..23 images, which the reference rejects ("input stacks too large",
     * src/impl/cpu.cpp:154-155, src/impl/cuda.cu:519-520). 0 = the reference's limit of 256 bits. */
    int wide_descriptors;
} bicos_b200_config;

/* `mode` arguments outside a config (bicos_b200_descriptor_words, bicos_b200_transform): bit 0 is
 * the TransformMode, OR-ing this in allows wide descriptors as bicos_b200_config::wide_descriptors does */
#define BICOS_B200_MODE_WIDE 2

typedef struct bicos_b200_handle_s* bicos_b200_handle;

const char* bicos_b200_last_error(void);

/* Number of CUDA devices visible, or a negative status. */
int bicos_b200_device_count(void);

/* A handle owns the per-device workspace (descriptor buffers, search keys, the subpixel x
 * table, pinned staging for the *_host entry point). One handle serves one match at a time;
 * use one handle per concurrent stream. device < 0 = current device. */
int bicos_b200_create(bicos_b200_handle* out, int device);
int bicos_b200_destroy(bicos_b200_handle h);

/* Words (uint32) per descriptor the reference's dispatch picks for n images
 * (src/impl/cpu.cpp:122-156): 1/2/4/8, or BICOS_B200_ERR_INVALID above 256 bits (with
 * BICOS_B200_MODE_WIDE in `mode`: 12 / 16 up to 512 bits). */
int bicos_b200_descriptor_words(int n, int mode);

/* Result type codes for a configuration: disparity BICOS_B200_16S (no threshold) or
 * BICOS_B200_32F; corrmap BICOS_B200_32F / BICOS_B200_64F, 0 when no threshold is set. */
int bicos_b200_disparity_type(const bicos_b200_config* cfg);
int bicos_b200_corrmap_type(const bicos_b200_config* cfg);

/* ---- the path, stage by stage (device memory) --------------------------------------- */

/* Stage 1, reference descriptor_transform<>() (include/impl/cpu/descriptor_transform.hpp:125-138).
 * planes: host array of n device pointers to single-channel images [rows][pitch_bytes].
 * desc: device [rows][desc_pitch_words] uint32, K words per pixel, rows 16-byte aligned. */
int bicos_b200_transform(bicos_b200_handle h, const void* const* planes, int n, int rows, int cols,
                         size_t pitch_bytes, int depth, int mode, uint32_t* desc,
                         size_t desc_pitch_words, void* stream);

/* Stage 2+4a, reference bicos<>() search part (include/impl/cpu/bicos.hpp:50-76, 78-97).
 * All results are device [rows][cols] uint32 minima of keys cost << 16 | column, filled by
 * this call:
 *   fwd_first  per left pixel: lowest Hamming cost and the FIRST right column attaining it
 *   fwd_last   (FLAG_NODUPES) same cost with 65535 - column: the LAST right column attaining
 *              it; the match is unique iff both name the same column
 *   rev_first / rev_last  (FLAG_CONSISTENCY; rev_last only with FLAG_NODUPES as well) the same
 *              per right column over the left row: the reverse search of the consistency check */
int bicos_b200_search(bicos_b200_handle h, const uint32_t* desc0, const uint32_t* desc1, int K,
                      int rows, int cols, size_t desc_pitch_words, int flags, uint32_t* fwd_first,
                      uint32_t* fwd_last, uint32_t* rev_first, uint32_t* rev_last, void* stream);

/* Stage 4b+3, reference bicos<>() postfilter (bicos.hpp:95-110) + agree / agree_subpixel
 * (include/impl/cpu/agree.hpp:53-191) on the keys of bicos_b200_search. raw_disp_out
 * (optional, dense int16 [rows][cols]) receives the postfilter result before the NXC test. */
int bicos_b200_refine(bicos_b200_handle h, const void* const* planes0, const void* const* planes1,
                      int n, int rows, int cols, size_t pitch_bytes, int depth,
                      const bicos_b200_config* cfg, const uint32_t* fwd_first, const uint32_t* fwd_last,
                      const uint32_t* rev_first, const uint32_t* rev_last, int16_t* raw_disp_out,
                      void* disparity, size_t disparity_pitch_bytes, void* corrmap,
                      size_t corrmap_pitch_bytes, void* stream);

/* ---- the whole path ------------------------------------------------------------------ */

/* Device-resident BICOS::match: transform x2 -> search -> postfilter+refine on `stream`.
 * disparity: int16 (no threshold) or float32 rows of disparity_pitch_bytes; integer mode with a
 * threshold keeps -32768.0f as the invalid marker, subpixel mode uses NaN (reference CPU
 * backend, src/impl/cpu.cpp:77-95). corrmap may be NULL; otherwise float32 (SINGLE) or
 * float64 (DOUBLE), NaN where no correlation was evaluated. */
int bicos_b200_match(bicos_b200_handle h, const void* const* planes0, const void* const* planes1,
                     int n, int rows, int cols, size_t pitch_bytes, int depth,
                     const bicos_b200_config* cfg, void* disparity, size_t disparity_pitch_bytes,
                     void* corrmap, size_t corrmap_pitch_bytes, void* stream);

/* Throughput mode: `count` independent stereo stacks of one shape, type and configuration (BASELINE.json's batch
 * configuration). planes0[f] / planes1[f] are the n plane pointers of frame f, disparity[f] / corrmap[f] its outputs
 * (corrmap may be NULL, or hold NULL entries). Same results as `count` calls of bicos_b200_match, but the frames
 * flow through two internal streams: the transform + search of frame f + 1 (tensor cores) runs beside the
 * postfilter + refine of frame f (FP32 pipe), which the register budget of the search kernel is laid out for.
 * `stream` is joined on both sides: earlier work on it completes before the first frame starts, and it continues
 * only after the last frame is complete. The reference has no batch entry point: its callers loop over
 * BICOS::match (src/cli.cpp:188-215 per stack). */
int bicos_b200_match_batch(bicos_b200_handle h, int count, const void* const* const* planes0,
                           const void* const* const* planes1, int n, int rows, int cols, size_t pitch_bytes,
                           int depth, const bicos_b200_config* cfg, void* const* disparity,
                           size_t disparity_pitch_bytes, void* const* corrmap, size_t corrmap_pitch_bytes,
                           void* stream);

/* Whether bicos_b200_match and bicos_b200_match_batch may run the search of one unit (row band, frame) beside the
 * refine of the previous one on the handle's two internal streams (default: yes). 0 = every kernel of a match one
 * after the other on the caller's stream, as in the reference (src/impl/cuda.cu:148-458): same results, used for
 * A/B timing and when a caller wants nothing enqueued outside its own stream. */
int bicos_b200_set_overlap(bicos_b200_handle h, int enabled);

/* Host-resident variant (what pybicos' BICOS_Match needs): dense host images in, dense host
 * results out; uploads, runs bicos_b200_match and downloads through the handle's pinned
 * staging buffers, then synchronises. */
int bicos_b200_match_host(bicos_b200_handle h, const void* const* host_planes0,
                          const void* const* host_planes1, int n, int rows, int cols, int depth,
                          const bicos_b200_config* cfg, void* host_disparity, void* host_corrmap);

/* The same in two halves, for callers that keep several frames in flight (one handle per
 * frame in flight): _begin validates, enqueues all copies and kernels on the handle's own
 * streams and returns; _end blocks until the host buffers hold the results. The host buffers
 * must stay valid, and should be pinned, between the two calls. A second _begin on a handle
 * whose previous host match has not been ended is an error. */
int bicos_b200_match_host_begin(bicos_b200_handle h, const void* const* host_planes0,
                                const void* const* host_planes1, int n, int rows, int cols, int depth,
                                const bicos_b200_config* cfg, void* host_disparity, void* host_corrmap);
int bicos_b200_match_host_end(bicos_b200_handle h);

/* Row-sharded variant for one process driving several GPUs: rows [row_begin, row_end) of the
 * same device-resident inputs are matched on this handle's device and written into the
 * matching rows of the (possibly peer-mapped) output buffers. Rows are independent
 * (SURVEY.md 8e), so no halo and no collective is involved. */
int bicos_b200_match_rows(bicos_b200_handle h, const void* const* planes0,
                          const void* const* planes1, int n, int rows, int cols, size_t pitch_bytes,
                          int depth, const bicos_b200_config* cfg, int row_begin, int row_end,
                          void* disparity, size_t disparity_pitch_bytes, void* corrmap,
                          size_t corrmap_pitch_bytes, void* stream);

/* ---- peer-memory output assembly for a row-sharded match, one process per GPU -------------
 * Rows are independent, so the only exchange of a row-sharded match is the assembly of the
 * output rows on one GPU. Instead of a gather after the kernels, the assembling rank allocates
 * the output images with bicos_b200_shared_alloc and publishes the 64-byte handles (over any
 * channel: torch.distributed, MPI, a pipe); every other rank maps them once with
 * bicos_b200_shared_open and passes mapped + row_begin * pitch as the disparity / corrmap
 * pointer of bicos_b200_match on ITS rows. Its refine kernel then stores its rows straight
 * into the assembling GPU's memory over NVLink; after a stream synchronisation and a barrier
 * the result is complete. Inside one process, BICOS::match_sharded does the same with
 * cudaDeviceEnablePeerAccess. */
#define BICOS_B200_IPC_HANDLE_BYTES 64
int bicos_b200_shared_alloc(int device, size_t bytes, void** dev_ptr, void* handle_out);
int bicos_b200_shared_open(int device, const void* handle, void** dev_ptr);
int bicos_b200_shared_close(int device, void* dev_ptr); /* for pointers from _shared_open */
int bicos_b200_shared_free(int device, void* dev_ptr); /* for pointers from _shared_alloc */

/* Blocks until all work enqueued through this handle's internal streams and `stream` is done. */
int bicos_b200_synchronize(bicos_b200_handle h, void* stream);

/* Per-stage device timing for benchmarks: while enabled, every match records CUDA events
 * around its three stages (0 = both descriptor transforms, 1 = search, 2 = postfilter+refine)
 * on the stream it runs on. bicos_b200_stage_times() waits for the device, then returns the
 * accumulated milliseconds per stage and the number of matches they cover. */
int bicos_b200_set_profiling(bicos_b200_handle h, int enabled);
int bicos_b200_stage_times(bicos_b200_handle h, double* ms_out3, long long* matches_out);

/* How many kernels of this library have been launched through the handle (bench bookkeeping). */
long long bicos_b200_kernel_launches(bicos_b200_handle h);

/* The search stage has two engines with identical results: BICOS_B200_SEARCH_TENSOR computes the
 * row's Hamming matrix as an int8 GEMM on the tensor cores (tcgen05, accumulators in TMEM, argmin as
 * epilogue; descriptors of 4/8/12/16 words, rows of up to 8192 pixels), BICOS_B200_SEARCH_POPC is the
 * XOR + POPC kernel on the integer pipes (any descriptor, rows of up to 32767 pixels). AUTO (the
 * default, also settable as environment BICOS_B200_SEARCH_ENGINE = auto | popc | mma) takes the
 * tensor-core engine wherever it applies. Process-wide; forcing TENSOR makes unsupported shapes
 * fail with BICOS_B200_ERR_CUDA instead of falling back. */
enum { BICOS_B200_SEARCH_AUTO = 0, BICOS_B200_SEARCH_POPC = 1, BICOS_B200_SEARCH_TENSOR = 2 };
int bicos_b200_set_search_engine(int engine);
int bicos_b200_get_search_engine(void);

/* Name of the kernel the calling thread's last search dispatched, e.g. "mma2<K=4,nodupes=0,ct=1,dirs=2>",
 * "mma1<...>" or "popc<K=4,flags=2>"; "" before the first search. For tests and benchmarks that must know
 * which of the engine's kernels they exercised. The string is thread-local and valid until the next search. */
const char* bicos_b200_last_search_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* BICOS_B200_H */
