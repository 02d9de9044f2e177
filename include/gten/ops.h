// gten/ops.h -- the gten::ops entry points (reference gten/ops.h:554-1133) on device tensors.
// Same names, argument order and meaning; every call validates like the reference and then runs one CUDA kernel of
// libgten_b200.so on rows [start_pos, n_ctx).  There is no host implementation behind these.
#pragma once
#include <cmath>

#include "tensor.h"

namespace gten {
namespace ops {

inline int gdt(Dtype d) { return (int)d; }

/// out[i] = weight[tokens[i]] (a Q4 table row is dequantised and re-encoded as Q8).  reference ops.h:554
inline void token_embed(const Tensor& weight, const Tensor& tokens, Tensor& out, const int start_pos = 0) {
    GTEN_ASSERT(weight.is_2d() && tokens.is_1d() && tokens.dtype() == kInt32 && out.is_2d());
    GTEN_ASSERT(out.dimsize(0) == tokens.numel() && out.dimsize(1) == weight.dimsize(1));
    GTEN_CUDA_OK(gtb_token_embed(weight.weight_handle(), static_cast<const int32_t*>(tokens.device_in()), out.device_out(),
                                 gdt(out.dtype()), tokens.numel(), start_pos));
}

/// out[r, c] = dot(x[r, :], w[c, :]); a 1-D out receives the last row's fp32 products (the logits).  reference ops.h:651
static void matmul_2d(const Tensor& x, const Tensor& w, Tensor& out, const int start_pos = 0) {
    GTEN_ASSERT(x.is_2d() && w.is_2d() && x.dimsize(1) == w.dimsize(1));
    const int n_ctx = x.dimsize(0);
    if (out.is_1d()) {
        GTEN_ASSERT(out.dtype() == kFloat32 && out.numel() == w.dimsize(0));
        GTEN_CUDA_OK(gtb_matmul_2d(x.device_in(), gdt(x.dtype()), n_ctx, w.weight_handle(), out.device_out(), gdt(kFloat32), 1, n_ctx - 1));
    } else {
        GTEN_ASSERT(out.is_2d() && out.dimsize(0) == n_ctx && out.dimsize(1) == w.dimsize(0));
        GTEN_CUDA_OK(gtb_matmul_2d(x.device_in(), gdt(x.dtype()), n_ctx, w.weight_handle(), out.device_out(), gdt(out.dtype()), 0, start_pos));
    }
}

static void silu(const Tensor& inp, Tensor& out, const int start_pos = 0) {                       // ops.h:700
    GTEN_ASSERT(inp.is_2d() && out.is_2d() && inp.shape_eq(out.shape()) && inp.dtype() == out.dtype());
    GTEN_CUDA_OK(gtb_silu(inp.device_in(), gdt(inp.dtype()), inp.dimsize(0), inp.dimsize(1), out.device_out(), start_pos));
}
static void silu_inplace(Tensor& inp, const int start_pos = 0) {                                  // ops.h:708
    GTEN_ASSERT(inp.is_2d());
    void* p = inp.device_out();
    GTEN_CUDA_OK(gtb_silu(p, gdt(inp.dtype()), inp.dimsize(0), inp.dimsize(1), p, start_pos));
}
/// rotate-half RoPE per head of d_head, in place.  reference ops.h:757
static void rotary_emb(Tensor& inp, const int d_head, const int start_pos = 0) {
    GTEN_ASSERT(inp.is_2d() && inp.dimsize(1) % d_head == 0);
    GTEN_CUDA_OK(gtb_rotary_emb(inp.device_out(), gdt(inp.dtype()), inp.dimsize(0), inp.dimsize(1), d_head, start_pos));
}
static void rms_norm(const Tensor& inp, const Tensor& weight, Tensor& out, const int start_pos = 0) {   // ops.h:806
    GTEN_ASSERT(inp.is_2d() && weight.is_1d() && weight.dtype() == kFloat16 && out.is_2d());
    GTEN_ASSERT(inp.dimsize(1) == weight.numel() && inp.shape_eq(out.shape()) && inp.dtype() == out.dtype());
    GTEN_CUDA_OK(gtb_rms_norm(inp.device_in(), gdt(inp.dtype()), inp.dimsize(0), inp.dimsize(1), weight.device_in(), out.device_out(), start_pos));
}
static void mul(const Tensor& inp0, const Tensor& inp1, Tensor& out, const int start_pos = 0) {   // ops.h:853
    GTEN_ASSERT(inp0.is_2d() && inp0.shape_eq(inp1.shape()) && inp0.shape_eq(out.shape()) && inp0.dtype() == inp1.dtype());
    GTEN_CUDA_OK(gtb_mul(inp0.device_in(), inp1.device_in(), gdt(inp0.dtype()), inp0.dimsize(0), inp0.dimsize(1), out.device_out(), start_pos));
}
static void mul_inplace(Tensor& inp0, const Tensor& inp1, const int start_pos = 0) {              // ops.h:861
    GTEN_ASSERT(inp0.is_2d() && inp0.shape_eq(inp1.shape()) && inp0.dtype() == inp1.dtype());
    void* p = inp0.device_out();
    GTEN_CUDA_OK(gtb_mul(p, inp1.device_in(), gdt(inp0.dtype()), inp0.dimsize(0), inp0.dimsize(1), p, start_pos));
}
static void add(const Tensor& x0, const Tensor& x1, Tensor& out, const int start_pos = 0) {       // ops.h:900
    GTEN_ASSERT(x0.is_2d() && x0.shape_eq(x1.shape()) && x0.shape_eq(out.shape()) && x0.dtype() == x1.dtype() && x0.dtype() == out.dtype());
    GTEN_CUDA_OK(gtb_add(x0.device_in(), x1.device_in(), gdt(x0.dtype()), x0.dimsize(0), x0.dimsize(1), out.device_out(), start_pos));
}
/// causal GQA attention: softmax(q k^T / sqrt(d_head)) v with n_heads query heads sharing k/v heads.  reference ops.h:1118.
/// q: (n_ctx, n_heads*d_head), k, v: (n_ctx, n_kv*d_head), qkv: (n_ctx, n_heads*d_head).  `qk` is the reference's
/// n_heads x max_ctx x max_ctx score buffer; scores stay on chip here and the tensor is not touched.
static void qkv_attn(const Tensor& q, const Tensor& k, const Tensor& v, Tensor& qk, Tensor& qkv, const int max_ctx, const int start_pos = 0) {
    (void)qk;
    GTEN_ASSERT(q.is_2d() && k.is_2d() && v.is_2d() && qkv.is_2d());
    GTEN_ASSERT(q.dtype() == k.dtype() && k.dtype() == v.dtype() && v.dtype() == qkv.dtype());
    const int n_ctx = q.dimsize(0), d_head = 64;
    GTEN_ASSERT(q.dimsize(1) % d_head == 0 && k.dimsize(1) % d_head == 0 && k.shape_eq(v.shape()) && q.shape_eq(qkv.shape()));
    GTEN_CUDA_OK(gtb_qkv_attn(q.device_in(), k.device_in(), v.device_in(), nullptr, qkv.device_out(), gdt(q.dtype()), n_ctx,
                              q.dimsize(1) / d_head, k.dimsize(1) / d_head, d_head, max_ctx, start_pos));
}

/// dot product of two HOST rows in the given dtypes (reference ops.h:482-512): (Q8, Q8), (Q8, Q4), (F16, F16), (F32, F32);
/// evaluated on the device in the reference's AVX order (4 integer lanes per block / 8 float lanes), bit-identical
inline float vec_dot_product(const char* inp0, Dtype inp0_dtype, const char* inp1, Dtype inp1_dtype, int vecsize) {
    float r = 0.0f;
    GTEN_CUDA_OK(gtb_vec_dot_product(inp0, gdt(inp0_dtype), inp1, gdt(inp1_dtype), vecsize, &r));
    return r;
}

/// row codecs on host buffers (reference ops.h:40-96): decode / encode one row in the given dtype
inline void read_row_to_float(const char* inp, Dtype inp_dtype, float* out_buf, const int rowsize) {
    switch (inp_dtype) {
        case kQint4: q4_dequantize_row(reinterpret_cast<const Q4Block*>(inp), out_buf, rowsize); break;
        case kQint8: q8_dequantize_row(reinterpret_cast<const Q8Block*>(inp), out_buf, rowsize); break;
        case kFloat16: for (int i = 0; i < rowsize; i++) out_buf[i] = fp16_to_fp32(reinterpret_cast<const Float16*>(inp)[i]); break;
        case kFloat32: std::memcpy(out_buf, inp, (size_t)rowsize * 4); break;
        default: GTEN_ASSERT(false);
    }
}
inline void write_row_from_float(float* inp, char* out, Dtype out_dtype, int rowsize) {
    switch (out_dtype) {
        case kQint8: q8_quantize_row(inp, reinterpret_cast<Q8Block*>(out), rowsize); break;
        case kFloat16: for (int i = 0; i < rowsize; i++) reinterpret_cast<Float16*>(out)[i] = fp32_to_fp16(inp[i]); break;
        case kFloat32: std::memcpy(out, inp, (size_t)rowsize * 4); break;
        default: GTEN_ASSERT(false);                      // there is no Q4 encoder (ops.h:73-96)
    }
}

}  // namespace ops
}  // namespace gten
