// afa_internal.h -- what the translation units of libafa_sm100.so share (not part of the C ABI):
// the thread-local error string behind afa_last_error() and the launch counter behind afa_launch_count().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace afa_internal {
int set_error(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int cuda_error(cudaError_t e, const char* what);
void count_launch();

// afa_tc.cu: Activation1d forward with both FIR filters on the tensor cores (bf16 I/O, 16-byte aligned rows)
void tc_set_tuning(int enable, int ny, int rlog2);
void tc_set_mats(int mats);              // tap matrices per K slice, up * 10 + down: 22 (default), 21; harness also 12, 11
void tc_set_debug_window(int j0);        // harness: first block of the clock-stamp window
bool tc_eligible(const void* x, const void* y, int64_t batch, int64_t channels, int64_t T, int dtype);
void tc_plan(int64_t rows, int64_t T, int* rlog2_out, int* ny_out, int64_t* n_rgroups, int64_t* n_tstrips);
int tc_fwd_launch(const void* x, void* y, const float* alpha, const float* beta, const float* taps_up12,
                  const float* taps_down12, int64_t batch, int64_t channels, int64_t T, int flags, cudaStream_t st,
                  int debug, float* dbg, int64_t x_pitch = 0, int64_t y_pitch = 0);      // pitches in elements; 0 = dense (T)
bool tc_pitched_ok(const void* x, int64_t x_pitch, const void* y, int64_t y_pitch, int64_t batch, int64_t channels, int64_t T,
                   int dtype);
int tc_kernel_info(int32_t out[6]);
}  // namespace afa_internal
