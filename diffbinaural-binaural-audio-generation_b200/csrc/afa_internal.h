// afa_internal.h -- what the translation units of libafa_sm100.so share (not part of the C ABI):
// the thread-local error string behind afa_last_error() and the launch counter behind afa_launch_count().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

namespace afa_internal {
// Programmatic dependent launch (cudaLaunchAttributeProgrammaticStreamSerialization) of the tensor-core kernels (afa_tc_fwd_kernel,
// afa_tc_cl_fwd_kernel): they call griddepcontrol.launch_dependents at entry and griddepcontrol.wait before the first access to
// global memory a predecessor may have written, so the launch latency and the set-up of kernel N + 1 (barriers, tensor-memory
// allocation, tap matrices) run in the shadow of kernel N's tail: 1-2 us per launch.  Without the attribute both instructions
// are no-ops; afa_set_tuning(9, 0) launches without it (A/B measurements).  The register-walk kernels do NOT use it: measured
// same-box (profiles/r02_pdl_ab.log), their one-clip launches got 10-15 % slower with it, the large ones did not change.
bool pdl_enabled();
void pdl_set(int on);
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

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
void* tc_encode_tiled();                 // cuTensorMapEncodeTiled from the driver (nullptr: not available)
void tc_split_bf16(float v, uint16_t* hi, uint16_t* lo);
int tc_mats();
int tc_mode();                           // afa_set_tuning(5, mode, ...): 0 off / 1 heuristic / 2 whenever eligible

// afa_tc_cl.cu: the same kernel for channels-last [batch, T, channels] bf16 activations (the generator engine's layout)
void tc_cl_set_tuning(int enable, int ny);      // enable: 0 off / 1 heuristic / 2 whenever eligible; ny: forced blocks of 16 per CTA
bool tc_cl_eligible(const void* x, int64_t x_bs, const void* res, const void* y, int64_t y_bs, int64_t y_tpad, int64_t batch,
                    int64_t channels, int64_t T, int dtype);
int tc_cl_fwd_launch(const void* x, int64_t x_bs, const float* bias, void* y, int64_t y_bs, int64_t y_tpad, const float* alpha,
                     const float* beta, const float* taps_up12, const float* taps_down12, int64_t batch, int64_t channels,
                     int64_t T, int flags, cudaStream_t st);
int tc_cl_kernel_info(int32_t out[6]);
}  // namespace afa_internal
