// gtb_internal.h -- host-side plumbing shared by the .cu files of libgten_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/gten_b200.h"
#include "gtb_dev.cuh"

namespace gtb {

struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    int64_t mem = 0;
    std::string err;
};
Context& ctx();
int fail(int code, const char* fmt, ...);
int ensure_init();

#define GTB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return gtb::fail(GTB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define GTB_CHECK_INIT()                      \
    do {                                      \
        int _r = gtb::ensure_init();          \
        if (_r) return _r;                    \
        cudaSetDevice(gtb::ctx().device);     /* the current device is per host thread */ \
    } while (0)
#define GTB_ARG(cond)                                                                                \
    do {                                                                                             \
        if (!(cond)) return gtb::fail(GTB_ERR_ARG, "argument check failed: %s (%s:%d)", #cond, __FILE__, __LINE__); \
    } while (0)
#define GTB_LAUNCHED()                                                            \
    do {                                                                          \
        gtb::ctx().launches++;                                                    \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e != cudaSuccess)                                                    \
            return gtb::fail(GTB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

}  // namespace gtb

// Device layout of a weight matrix W[rows][cols] (one-time repack at upload, byte count unchanged).
//   Q4: data  = uint4[rows][cols/32], word l of a block = payload bytes (2l, 2l+1, 2l+8, 2l+9): its high
//               nibbles are elements {2l,2l+1,2l+8,2l+9}, its low nibbles the same +16 = reference lane l
//               (gten/ops.h:339-378);  scales = fp16[rows][cols/32].
//   Q8: data  = 2 x uint4 per block: words X0..X3 = bytes (2l,2l+1,2l+8,2l+9), Y0..Y3 = the same +16.
//   F16: data = uint4[rows][cols/64][8]: chunk c, lane l holds elements 64c + 8i + l, i = 0..7
//               (reference lane l of the 8-wide AVX accumulator, gten/ops.h:144-152).
struct gtb_weight {
    int dtype = 0, rows = 0, cols = 0;
    void* data = nullptr;
    uint16_t* scales = nullptr;
    size_t nbytes = 0;
    bool owns = true;          // false: data/scales are slices of a buffer owned by someone else (fused q|k|v, gate|up)
};

namespace gtb {
// repack a host payload into caller-provided device storage (slices of a fused buffer); *out is a non-owning view
int weight_upload_view(gtb_weight_t* out, const void* h_payload, int dtype, int rows, int cols, void* d_data, uint16_t* d_scales);
// repack a payload that already sits in DEVICE memory (gten layout); no synchronisation: the caller keeps d_payload alive until the
// stream has passed the repack kernel.  d_data == nullptr: the weight owns freshly allocated storage.
int weight_device_view(gtb_weight_t* out, const void* d_payload, int dtype, int rows, int cols, void* d_data, uint16_t* d_scales);
inline size_t weight_data_bytes(int dtype, int rows, int cols) {
    return dtype == GTB_F16 ? (size_t)rows * cols * 2 : (size_t)rows * cols / 32 * (dtype == GTB_Q4 ? 16 : 32);
}
inline size_t weight_scale_bytes(int dtype, int rows, int cols) { return dtype == GTB_F16 ? 0 : (size_t)rows * cols / 32 * 2; }
}
