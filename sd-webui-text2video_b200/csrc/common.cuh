// Shared host-side helpers of the library.
#pragma once
#include <cuda_runtime.h>

namespace t2v {
void set_error(const char* fmt, ...);
int num_sms();
}  // namespace t2v
