// gtb_api.cu -- context, memory, weight upload/repack and row codecs of libgten_b200.so
#include <stdarg.h>

#include <vector>

#include "gtb_internal.h"

namespace gtb {

Context& ctx() {
    static Context c;
    return c;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    ctx().err = buf;
    return code;
}

int ensure_init() {
    if (ctx().ready) return GTB_OK;
    return gtb_init(0);
}

// ------------------------------------------------------------------ repack kernels (see gtb_internal.h)
__global__ void k_repack_q4(const uint8_t* __restrict__ raw, uint4* __restrict__ data, uint16_t* __restrict__ scales, size_t nblocks) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const uint8_t* b = raw + i * Q4_BYTES;
    scales[i] = (uint16_t)b[0] | ((uint16_t)b[1] << 8);
    uint32_t w[4];
#pragma unroll
    for (int l = 0; l < 4; l++)
        w[l] = (uint32_t)b[2 + 2 * l] | ((uint32_t)b[3 + 2 * l] << 8) | ((uint32_t)b[10 + 2 * l] << 16) | ((uint32_t)b[11 + 2 * l] << 24);
    data[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void k_repack_q8(const uint8_t* __restrict__ raw, uint4* __restrict__ data, uint16_t* __restrict__ scales, size_t nblocks) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const uint8_t* b = raw + i * Q8_BYTES;
    scales[i] = (uint16_t)b[0] | ((uint16_t)b[1] << 8);
    uint32_t w[8];
#pragma unroll
    for (int half = 0; half < 2; half++)
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const uint8_t* p = b + 2 + 16 * half + 2 * l;
            w[half * 4 + l] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[8] << 16) | ((uint32_t)p[9] << 24);
        }
    data[2 * i] = make_uint4(w[0], w[1], w[2], w[3]);
    data[2 * i + 1] = make_uint4(w[4], w[5], w[6], w[7]);
}

__global__ void k_repack_f16(const uint16_t* __restrict__ raw, uint4* __restrict__ data, int cols, size_t nvec) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row, chunk, lane)
    if (i >= nvec) return;
    const int cpr = cols / 64;
    const int l = (int)(i & 7);
    const size_t rc = i >> 3;
    const int c = (int)(rc % cpr);
    const size_t row = rc / cpr;
    const uint16_t* src = raw + row * cols + 64 * c + l;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; j++) w[j] = (uint32_t)src[8 * (2 * j)] | ((uint32_t)src[8 * (2 * j + 1)] << 16);
    data[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// natural element e (0..31) of a block from the device layout
__device__ __forceinline__ int q4_elem(const uint4& d, int e) {
    const int j = e & 15;                                 // payload byte index
    const int l = (j & 7) >> 1, pos = (j & 1) + 2 * (j >> 3);
    const uint32_t w = (l == 0) ? d.x : (l == 1) ? d.y : (l == 2) ? d.z : d.w;
    const uint32_t byte = (w >> (8 * pos)) & 0xffu;
    return (int)((e < 16) ? (byte >> 4) : (byte & 0x0fu)) - 7;
}
__device__ __forceinline__ int q8_elem(const uint4& dx, const uint4& dy, int e) {
    const uint4& d = (e < 16) ? dx : dy;
    const int j = e & 15;
    const int l = (j & 7) >> 1, pos = (j & 1) + 2 * (j >> 3);
    const uint32_t w = (l == 0) ? d.x : (l == 1) ? d.y : (l == 2) ? d.z : d.w;
    return (int)(int8_t)((w >> (8 * pos)) & 0xffu);
}

__global__ void k_dequant_weight(const void* __restrict__ data, const uint16_t* __restrict__ scales, int dtype, int cols,
                                 int row0, int nrows, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrows * cols) return;
    const int col = (int)(i % cols);
    const size_t row = row0 + i / cols;
    float v;
    if (dtype == DT_F16) {
        const int c = col >> 6, r = col & 63, l = r & 7, ii = r >> 3;
        const uint16_t* p = reinterpret_cast<const uint16_t*>(data) + ((row * (cols / 64) + c) * 8 + l) * 8 + ii;
        v = h2f(*p);
    } else {
        const size_t blk = row * (cols / 32) + (col >> 5);
        const float delta = h2f(scales[blk]);
        const uint4* d = reinterpret_cast<const uint4*>(data);
        const int q = (dtype == DT_Q4) ? q4_elem(d[blk], col & 31) : q8_elem(d[2 * blk], d[2 * blk + 1], col & 31);
        v = __fmul_rn((float)q, delta);                 // quants.h:74,86-87
    }
    out[i] = v;
}

// ------------------------------------------------------------------ row codecs on the reference layout
__global__ void k_write_rows(const float* __restrict__ in, uint8_t* __restrict__ out, int dtype, int rows, int n) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nb = (n + 31) / 32;
    if (warp >= rows * nb) return;
    const int row = warp / nb, b = warp % nb;
    const int e = b * 32 + lane;
    const float x = (e < n) ? in[(size_t)row * n + e] : 0.0f;
    uint8_t* orow = out + (size_t)row * row_nbytes(dtype, n);
    if (dtype == DT_Q8) {
        uint16_t dh;
        const int q = q8_encode_lane(x, &dh);
        uint8_t* blk = orow + (size_t)b * Q8_BYTES;
        if (lane == 0) { blk[0] = (uint8_t)(dh & 0xff); blk[1] = (uint8_t)(dh >> 8); }
        if (e < n) blk[2 + lane] = (uint8_t)(int8_t)q;
    } else if (dtype == DT_F16) {
        if (e < n) reinterpret_cast<uint16_t*>(orow)[e] = f2h(x);
    } else {
        if (e < n) reinterpret_cast<float*>(orow)[e] = x;
    }
}

__global__ void k_read_rows(const uint8_t* __restrict__ in, int dtype, float* __restrict__ out, int rows, int n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * n) return;
    const int row = (int)(i / n), e = (int)(i % n);
    out[i] = read_elem(in + (size_t)row * row_nbytes(dtype, n), dtype, e);
}

}  // namespace gtb

using namespace gtb;

extern "C" {

const char* gtb_last_error(void) { return ctx().err.c_str(); }
const char* gtb_version(void) { return "gten-b200 0.1 (sm_100a)"; }

int gtb_device_count(int* n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *n = 0; cudaGetLastError(); return GTB_OK; }
    *n = c;
    return GTB_OK;
}

int gtb_init(int device) {
    Context& c = ctx();
    if (c.ready && c.device == device) return GTB_OK;
    // one library instance <-> one GPU: the stream and every kernel attribute set so far belong to the first device
    if (c.ready) return fail(GTB_ERR_STATE, "libgten_b200 is bound to device %d; use one process per GPU (asked for device %d)", c.device, device);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(GTB_ERR_NO_DEVICE, "no CUDA device: libgten_b200 has no CPU fallback (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
    }
    GTB_ARG(device >= 0 && device < n);
    GTB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    GTB_CUDA(cudaGetDeviceProperties(&p, device));
    if (c.stream == nullptr) GTB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = device;
    c.sm_count = p.multiProcessorCount;
    c.ready = true;
    return GTB_OK;
}

int gtb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    GTB_CHECK_INIT();
    cudaDeviceProp p;
    GTB_CUDA(cudaGetDeviceProperties(&p, ctx().device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return GTB_OK;
}

int gtb_sync(void) {
    GTB_CHECK_INIT();
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    return GTB_OK;
}
void* gtb_stream(void) { return ctx().ready ? (void*)ctx().stream : nullptr; }
int64_t gtb_launch_count(void) { return ctx().launches; }
int64_t gtb_mem_allocated(void) { return ctx().mem; }

int gtb_malloc(void** d_ptr, size_t nbytes) {
    GTB_CHECK_INIT();
    GTB_ARG(d_ptr != nullptr);
    GTB_CUDA(cudaMalloc(d_ptr, nbytes ? nbytes : 1));
    ctx().mem += (int64_t)nbytes;
    return GTB_OK;
}
int gtb_free(void* d_ptr) {
    GTB_CHECK_INIT();
    if (d_ptr) { GTB_CUDA(cudaStreamSynchronize(ctx().stream)); GTB_CUDA(cudaFree(d_ptr)); }
    return GTB_OK;
}
int gtb_memset(void* d_ptr, int value, size_t nbytes) {
    GTB_CHECK_INIT();
    GTB_CUDA(cudaMemsetAsync(d_ptr, value, nbytes, ctx().stream));
    return GTB_OK;
}
int gtb_h2d(void* d_dst, const void* h_src, size_t nbytes) {
    GTB_CHECK_INIT();
    GTB_CUDA(cudaMemcpyAsync(d_dst, h_src, nbytes, cudaMemcpyHostToDevice, ctx().stream));
    return GTB_OK;
}
int gtb_d2h(void* h_dst, const void* d_src, size_t nbytes) {
    GTB_CHECK_INIT();
    GTB_CUDA(cudaMemcpyAsync(h_dst, d_src, nbytes, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    return GTB_OK;
}
int gtb_d2d(void* d_dst, const void* d_src, size_t nbytes) {
    GTB_CHECK_INIT();
    GTB_CUDA(cudaMemcpyAsync(d_dst, d_src, nbytes, cudaMemcpyDeviceToDevice, ctx().stream));
    return GTB_OK;
}
int gtb_host_alloc(void** h_ptr, size_t nbytes) {
    GTB_CHECK_INIT();
    GTB_CUDA(cudaMallocHost(h_ptr, nbytes ? nbytes : 1));
    return GTB_OK;
}
int gtb_host_free(void* h_ptr) {
    GTB_CHECK_INIT();
    if (h_ptr) GTB_CUDA(cudaFreeHost(h_ptr));
    return GTB_OK;
}

// ------------------------------------------------------------------ weights
static int weight_from_device_impl(gtb_weight_t* out, const void* d_payload, int dtype, int rows, int cols, void* d_data, uint16_t* d_scales) {
    GTB_CHECK_INIT();
    GTB_ARG(out && d_payload && rows > 0 && cols > 0);
    GTB_ARG(dtype == GTB_F16 || dtype == GTB_Q8 || dtype == GTB_Q4);
    GTB_ARG(dtype == GTB_F16 ? (cols % 64 == 0) : (cols % 32 == 0));
    auto* w = new gtb_weight();
    w->dtype = dtype; w->rows = rows; w->cols = cols;
    w->owns = (d_data == nullptr);
    cudaStream_t st = ctx().stream;
    const int T = 256;
    if (dtype == GTB_F16) {
        const size_t nvec = (size_t)rows * cols / 8;
        w->nbytes = nvec * 16;
        if (w->owns) GTB_CUDA(cudaMalloc(&w->data, w->nbytes)); else w->data = d_data;
        k_repack_f16<<<(unsigned)((nvec + T - 1) / T), T, 0, st>>>((const uint16_t*)d_payload, (uint4*)w->data, cols, nvec);
    } else {
        const size_t nblk = (size_t)rows * cols / 32;
        const size_t dbytes = nblk * (dtype == GTB_Q4 ? 16 : 32);
        w->nbytes = dbytes + nblk * 2;
        if (w->owns) {
            GTB_CUDA(cudaMalloc(&w->data, dbytes));
            GTB_CUDA(cudaMalloc((void**)&w->scales, nblk * 2));
        } else {
            w->data = d_data; w->scales = d_scales;
        }
        if (dtype == GTB_Q4) k_repack_q4<<<(unsigned)((nblk + T - 1) / T), T, 0, st>>>((const uint8_t*)d_payload, (uint4*)w->data, w->scales, nblk);
        else k_repack_q8<<<(unsigned)((nblk + T - 1) / T), T, 0, st>>>((const uint8_t*)d_payload, (uint4*)w->data, w->scales, nblk);
    }
    GTB_LAUNCHED();
    if (w->owns) ctx().mem += (int64_t)w->nbytes;
    *out = w;
    return GTB_OK;
}

int gtb_weight_from_device(gtb_weight_t* out, const void* d_payload, int dtype, int rows, int cols) {
    return weight_from_device_impl(out, d_payload, dtype, rows, cols, nullptr, nullptr);
}

static int weight_upload_impl(gtb_weight_t* out, const void* h_payload, int dtype, int rows, int cols, void* d_data, uint16_t* d_scales) {
    GTB_CHECK_INIT();
    GTB_ARG(out && h_payload && rows > 0 && cols > 0);
    const size_t nb = (size_t)rows * row_nbytes(dtype, cols);
    void* tmp = nullptr;
    GTB_CUDA(cudaMalloc(&tmp, nb));
    cudaError_t e = cudaMemcpyAsync(tmp, h_payload, nb, cudaMemcpyHostToDevice, ctx().stream);
    int r = (e == cudaSuccess) ? weight_from_device_impl(out, tmp, dtype, rows, cols, d_data, d_scales)
                               : fail(GTB_ERR_CUDA, "weight H2D failed: %s", cudaGetErrorString(e));
    cudaStreamSynchronize(ctx().stream);
    cudaFree(tmp);
    return r;
}

int gtb_weight_upload(gtb_weight_t* out, const void* h_payload, int dtype, int rows, int cols) {
    return weight_upload_impl(out, h_payload, dtype, rows, cols, nullptr, nullptr);
}

int gtb_weight_free(gtb_weight_t w) {
    if (!w) return GTB_OK;
    GTB_CHECK_INIT();
    cudaStreamSynchronize(ctx().stream);
    if (w->owns) {
        if (w->data) cudaFree(w->data);
        if (w->scales) cudaFree(w->scales);
        ctx().mem -= (int64_t)w->nbytes;
    }
    delete w;
    return GTB_OK;
}

int gtb_weight_nbytes(gtb_weight_t w, size_t* nbytes) {
    GTB_ARG(w && nbytes);
    *nbytes = w->nbytes;
    return GTB_OK;
}

int gtb_weight_dequant(gtb_weight_t w, int row0, int nrows, float* h_out) {
    GTB_CHECK_INIT();
    GTB_ARG(w && h_out && row0 >= 0 && nrows > 0 && row0 + nrows <= w->rows);
    const size_t n = (size_t)nrows * w->cols;
    float* d = nullptr;
    GTB_CUDA(cudaMalloc((void**)&d, n * 4));
    k_dequant_weight<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>(w->data, w->scales, w->dtype, w->cols, row0, nrows, d);
    ctx().launches++;
    cudaError_t e = cudaMemcpyAsync(h_out, d, n * 4, cudaMemcpyDeviceToHost, ctx().stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(GTB_ERR_CUDA, "weight dequant failed: %s", cudaGetErrorString(e));
    return GTB_OK;
}

// ------------------------------------------------------------------ row codecs
int gtb_write_rows_from_float(const float* d_in, void* d_out, int out_dtype, int rows, int n) {
    GTB_CHECK_INIT();
    GTB_ARG(d_in && d_out && rows > 0 && n > 0);
    GTB_ARG(out_dtype == GTB_Q8 || out_dtype == GTB_F16 || out_dtype == GTB_F32);     // no Q4 encoder exists (ops.h:73-96)
    const size_t warps = (size_t)rows * ((n + 31) / 32);
    k_write_rows<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ctx().stream>>>(d_in, (uint8_t*)d_out, out_dtype, rows, n);
    GTB_LAUNCHED();
    return GTB_OK;
}

int gtb_read_rows_to_float(const void* d_in, int in_dtype, float* d_out, int rows, int n) {
    GTB_CHECK_INIT();
    GTB_ARG(d_in && d_out && rows > 0 && n > 0);
    GTB_ARG(in_dtype == GTB_Q8 || in_dtype == GTB_Q4 || in_dtype == GTB_F16 || in_dtype == GTB_F32);
    GTB_ARG(in_dtype != GTB_Q4 || n % 32 == 0);
    const size_t tot = (size_t)rows * n;
    k_read_rows<<<(unsigned)((tot + 255) / 256), 256, 0, ctx().stream>>>((const uint8_t*)d_in, in_dtype, d_out, rows, n);
    GTB_LAUNCHED();
    return GTB_OK;
}

}  // extern "C"

namespace gtb {
int weight_device_view(gtb_weight_t* out, const void* d_payload, int dtype, int rows, int cols, void* d_data, uint16_t* d_scales) {
    return weight_from_device_impl(out, d_payload, dtype, rows, cols, d_data, d_scales);
}
int weight_upload_view(gtb_weight_t* out, const void* h_payload, int dtype, int rows, int cols, void* d_data, uint16_t* d_scales) {
    if (!d_data) return fail(GTB_ERR_ARG, "weight_upload_view: no destination");
    return weight_upload_impl(out, h_payload, dtype, rows, cols, d_data, d_scales);
}
}  // namespace gtb
