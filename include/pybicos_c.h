/* pybicos_c.h -- the six symbols the reference's Python module binds with ctypes
 * (reference pybicos/__init__.py:12-77, implemented there in src/pybicos_c.cpp:92-209).
 * libbicos_b200/pybicos/pybicos_c.so exports exactly these, with the struct layouts the
 * Python side declares (pybicos/__init__.py:41-63), so the unmodified reference module
 * runs on top of the B200 kernels when that .so is placed next to its __init__.py.
 *
 * Differences to the reference's C file, all on its bug list:
 *  - `precision` is always part of BicosConfig (the reference drops it in CPU builds, which
 *    shifts every later field for the Python caller);
 *  - BICOS_FreeResult also frees the two result buffers (the reference leaks them);
 *  - no C++ exception escapes: any failure returns NULL, message in BICOS_LastError().
 */
#ifndef PYBICOS_C_H
#define PYBICOS_C_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/pybicos_c.cpp:30-41 (BICOS_CUDA layout) */
typedef struct {
    float nxcorr_threshold; /* < 0: unset */
    float subpixel_step; /* < 0: unset */
    float min_variance; /* < 0: unset */
    int mode; /* 0 = LIMITED, 1 = FULL */
    int precision; /* 0 = SINGLE, 1 = DOUBLE */
    int variant_type; /* 0 = NoDuplicates, 1 = Consistency */
    int max_lr_diff;
    int no_dupes;
} BicosConfig;

/* reference src/pybicos_c.cpp:44-53; types are OpenCV codes (CV_16S=3, CV_32F=5, CV_64F=6) */
typedef struct {
    void* disparity_data;
    int disparity_rows;
    int disparity_cols;
    int disparity_type;
    void* corrmap_data;
    int corrmap_rows;
    int corrmap_cols;
    int corrmap_type;
} BicosResult;

/* reference src/pybicos_c.cpp:92-108: defaults 0.5, -1, -1, LIMITED, SINGLE, NoDuplicates, 1, 0 */
BicosConfig* BICOS_CreateDefaultConfig(void);
/* reference src/pybicos_c.cpp:111-113 */
void BICOS_FreeConfig(BicosConfig* config);
/* reference src/pybicos_c.cpp:116-118 */
void BICOS_FreeResult(BicosResult* result);
/* reference src/pybicos_c.cpp:131-201: dense host images (data, rows, cols, depth code) per stack */
BicosResult* BICOS_Match(void** stack0_data, int* stack0_rows, int* stack0_cols, int* stack0_types,
                         int stack0_size, void** stack1_data, int* stack1_rows, int* stack1_cols,
                         int* stack1_types, int stack1_size, BicosConfig* config);
/* reference src/pybicos_c.cpp:203-209 */
float BICOS_InvalidDisparityFloat(void);
int16_t BICOS_InvalidDisparityInt16(void);

/* addition: message of the last failed BICOS_Match on this thread */
const char* BICOS_LastError(void);

#ifdef __cplusplus
}
#endif
#endif
