// afa_internal.h -- what the translation units of libafa_sm100.so share (not part of the C ABI):
// the thread-local error string behind afa_last_error() and the launch counter behind afa_launch_count().
#pragma once
#include <cuda_runtime.h>

namespace afa_internal {
int set_error(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
int cuda_error(cudaError_t e, const char* what);
void count_launch();
}  // namespace afa_internal
