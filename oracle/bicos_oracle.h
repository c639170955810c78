/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's
 * BICOS::match hot path. See bicos_oracle.c for the per-function citations.
 * Same calling convention as oracle/ref_runner.cpp (prefix orc_ instead of ref_),
 * plus *_f64 entry points for Precision::DOUBLE, which the reference only has in
 * its CUDA backend (include/impl/cuda/agree.cuh:35-65). */
#ifndef BICOS_ORACLE_H
#define BICOS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* orc_last_error(void);
void orc_set_threads(int n);
int orc_hardware_threads(void);

int orc_match(const void* stack0, const void* stack1, int n, int rows, int cols, int depth,
              float nxcorr_threshold, float subpixel_step, float min_variance, int mode_full,
              int variant_consistency, int max_lr_diff, int no_dupes, void* disp_out,
              int* disp_type, float* corr_out);

int orc_match_f64(const void* stack0, const void* stack1, int n, int rows, int cols, int depth,
                  float nxcorr_threshold, float subpixel_step, float min_variance, int mode_full,
                  int variant_consistency, int max_lr_diff, int no_dupes, void* disp_out,
                  int* disp_type, double* corr_out);

int orc_descriptors(const void* stack, int n, int rows, int cols, int depth, int mode_full,
                    uint32_t* out_words, int out_capacity_words_per_px);

int orc_bicos(const uint32_t* desc0, const uint32_t* desc1, int K, int rows, int cols, int flags,
              int max_lr_diff, int16_t* out);

int orc_agree(const int16_t* raw_disp, const void* stack0, const void* stack1, int n, int rows,
              int cols, int depth, float nxcorr_threshold, float subpixel_step,
              float min_variance_times_n, int16_t* disp_i16_out, float* disp_f32_out,
              float* corr_out);

int orc_agree_f64(const int16_t* raw_disp, const void* stack0, const void* stack1, int n, int rows,
                  int cols, int depth, float nxcorr_threshold, float subpixel_step,
                  float min_variance_times_n, int16_t* disp_i16_out, float* disp_f32_out,
                  double* corr_out);

#ifdef __cplusplus
}
#endif
#endif
