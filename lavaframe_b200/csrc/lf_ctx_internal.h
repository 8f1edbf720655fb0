// lf_ctx_internal.h — what the multi-device group (lfcuda_group.cpp) needs to see of a context (lfcuda.cpp).  Not part of the C ABI.
#pragma once

#include <cuda_runtime.h>

#include "lfcuda.h"

namespace lf {

struct CtxView {
    int device;
    cudaStream_t stream;
    float* accum;               // W * H * 3 running sum of this context's frames
    size_t accum_floats;
    float* out_f;               // post-process outputs (W * H * 3)
    unsigned char* out_u8;
    int width, height;
    LfPostParams post;
};
bool ctx_view(lfcuda_ctx* ctx, CtxView* out);        // false until lfcuda_set_params has succeeded
void ctx_count_launch(lfcuda_ctx* ctx);
int  ctx_fail(lfcuda_ctx* ctx, int code, const char* msg);

}  // namespace lf
