// gtb_ops.cu -- the gten::ops entry points (gten/ops.h) on device tensors kept in the reference's own
// row layout (Q8Block / fp16 rows), row range [start_pos, n_ctx) like every reference op.
// These share the device building blocks of the engine (gtb_kernels.cuh); the engine is the fast path,
// this file is what the C++ drop-in modules in include/gten/ call one op at a time.
#include <math.h>

#include <map>
#include <vector>

#include "gtb_internal.h"
#include "gtb_kernels.cuh"

namespace gtb {

// CTA-wide: encode n floats (shared memory) as one output row in the reference layout.
__device__ void write_row_encoded(const float* srow, uint8_t* orow, int dtype, int n) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = wid; b < (n + 31) / 32; b += nw) {
        const int e = b * 32 + lane;
        const float x = (e < n) ? srow[e] : 0.0f;
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
}

// ---- token_embed (ops.h:514-564): F16/Q8 rows are copied, a Q4 row is dequantised and re-encoded as Q8
__global__ void __launch_bounds__(NT) k_op_embed(const void* __restrict__ wdata, const uint16_t* __restrict__ wsc, int wdt,
                                                  int n_embd, const int32_t* __restrict__ tokens, uint8_t* __restrict__ out,
                                                  int odt, int start_pos) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* srow = reinterpret_cast<float*>(smem);
    const int i = start_pos + blockIdx.x;
    const size_t row = (size_t)tokens[i];
    uint8_t* orow = out + (size_t)i * row_nbytes(odt, n_embd);
    for (int e = threadIdx.x; e < n_embd; e += blockDim.x) {
        if (wdt == DT_F16) {
            const int c = e >> 6, r = e & 63, l = r & 7, ii = r >> 3;
            reinterpret_cast<uint16_t*>(orow)[e] = reinterpret_cast<const uint16_t*>(wdata)[((row * (n_embd / 64) + c) * 8 + l) * 8 + ii];
        } else {
            const size_t blk = row * (n_embd / 32) + (e >> 5);
            const int le = e & 31;
            if (wdt == DT_Q8) {
                uint8_t* ob = orow + (size_t)(e >> 5) * Q8_BYTES;
                if (le == 0) { const uint16_t s = wsc[blk]; ob[0] = (uint8_t)(s & 0xff); ob[1] = (uint8_t)(s >> 8); }
                ob[2 + le] = reinterpret_cast<const uint8_t*>(wdata)[blk * 32 + perm_byte(le)];
            } else {
                const int j = le & 15;
                const int l = (j & 7) >> 1, pos = (j & 1) + 2 * (j >> 3);
                const uint8_t byte = reinterpret_cast<const uint8_t*>(wdata)[blk * 16 + l * 4 + pos];
                const int q = (int)((le < 16) ? (byte >> 4) : (byte & 0x0f)) - 7;
                srow[e] = __fmul_rn((float)q, h2f(wsc[blk]));
            }
        }
    }
    if (wdt == DT_Q4) {
        __syncthreads();
        write_row_encoded(srow, orow, DT_Q8, n_embd);
    }
}

// ---- matmul_2d (ops.h:613-670): blockIdx.y = activation row, blockIdx.x = share of the weight rows
template <int WT>
__global__ void __launch_bounds__(NT) k_op_matmul(const uint8_t* __restrict__ x, int K, const void* __restrict__ wdata,
                                                   const uint16_t* __restrict__ wsc, int n_out, float* __restrict__ tmp, int start_pos) {
    constexpr int AT = (WT == DT_F16) ? DT_F16 : DT_Q8;
    extern __shared__ __align__(16) unsigned char smem[];
    ActView av = act_carve(AT, K, smem);
    const size_t off = (act_bytes(AT, K) + 15) & ~(size_t)15;
    float* ps = reinterpret_cast<float*>(smem + off);
    const int r = start_pos + blockIdx.y;
    const uint8_t* xrow = x + (size_t)r * row_nbytes(AT, K);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b = wid; b < K / 32; b += NWARP) stage_encoded_block<AT>(av, b, lane, xrow);
    __syncthreads();
    float* my_ps = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ps) + (gemv_ps_bytes(WT, K) / NWARP) * wid);
    int done = 0;
    gemv_matrix<WT>(wdata, wsc, n_out, K, av, my_ps, tmp + (size_t)blockIdx.y * n_out, 0, &done);
}

// ---- row-wise elementwise ops: one CTA per row
enum { EW_SILU = 0, EW_MUL = 1, EW_ADD = 2, EW_ROPE = 3, EW_NORM = 4 };

__global__ void __launch_bounds__(NT) k_op_rowwise(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                    uint8_t* __restrict__ out, int dtype, int n, int start_pos, int d_head,
                                                    const float* __restrict__ rcos, const float* __restrict__ rsin,
                                                    const uint16_t* __restrict__ normw) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* x0 = reinterpret_cast<float*>(smem);
    float* y = x0 + n;
    ExactSumSmem& es = *reinterpret_cast<ExactSumSmem*>(smem + (((size_t)2 * n * 4 + 15) & ~(size_t)15));
    const int i = start_pos + blockIdx.x;
    const size_t rb = row_nbytes(dtype, n);
    const uint8_t* arow = a + (size_t)i * rb;
    for (int e = threadIdx.x; e < n; e += blockDim.x) x0[e] = read_elem(arow, dtype, e);
    __syncthreads();
    if (op == EW_SILU) {
        for (int e = threadIdx.x; e < n; e += blockDim.x) y[e] = silu_ref(x0[e]);
    } else if (op == EW_MUL || op == EW_ADD) {
        const uint8_t* brow = b + (size_t)i * rb;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const float v = read_elem(brow, dtype, e);
            y[e] = (op == EW_MUL) ? __fmul_rn(x0[e], v) : __fadd_rn(x0[e], v);
        }
    } else if (op == EW_ROPE) {
        const int dh = d_head / 2;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int j = e % d_head, base = e - j;
            const int jj = (j < dh) ? j : j - dh;
            const float v0 = x0[base + jj], v1 = x0[base + jj + dh];
            const float cs = rcos[(size_t)i * dh + jj], sn = rsin[(size_t)i * dh + jj];
            y[e] = (j < dh) ? __fsub_rn(__fmul_rn(v0, cs), __fmul_rn(v1, sn)) : __fadd_rn(__fmul_rn(v0, sn), __fmul_rn(v1, cs));
        }
    } else {   // EW_NORM (ops.h:762-778)
        const float sq = exact_sum_block([&](int k) { const float v = x0[k]; return __fmul_rn(v, v); }, n, es);
        const float denom = __fadd_rn(sqrtf(__fdiv_rn(sq, (float)n)), 1e-6f);
        for (int e = threadIdx.x; e < n; e += blockDim.x) y[e] = __fmul_rn(__fdiv_rn(x0[e], denom), h2f(normw[e]));
    }
    __syncthreads();
    write_row_encoded(y, out + (size_t)i * rb, dtype, n);
}

// ---- attention on reference-layout q/k/v
template <int AT>
__global__ void k_op_kv_convert(const uint8_t* __restrict__ k, const uint8_t* __restrict__ v, int n_ctx, int kv_dim,
                                uint8_t* kq, uint16_t* ks, uint8_t* vq, uint16_t* vs) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_ctx * kv_dim) return;
    const int i = (int)(idx / kv_dim), c = (int)(idx % kv_dim);
    const size_t rb = row_nbytes(AT, kv_dim);
    if (AT == DT_F16) {
        reinterpret_cast<uint16_t*>(kq)[idx] = reinterpret_cast<const uint16_t*>(k + i * rb)[c];
        reinterpret_cast<uint16_t*>(vq)[idx] = reinterpret_cast<const uint16_t*>(v + i * rb)[c];
    } else {
        const uint8_t* kb = k + i * rb + (size_t)(c >> 5) * Q8_BYTES;
        const uint8_t* vb = v + i * rb + (size_t)(c >> 5) * Q8_BYTES;
        kq[(size_t)i * kv_dim + (c & ~31) + perm_byte(c & 31)] = kb[2 + (c & 31)];
        vq[idx] = vb[2 + (c & 31)];
        if ((c & 31) == 0) {
            ks[(size_t)i * (kv_dim / 32) + (c >> 5)] = (uint16_t)kb[0] | ((uint16_t)kb[1] << 8);
            vs[(size_t)i * (kv_dim / 32) + (c >> 5)] = (uint16_t)vb[0] | ((uint16_t)vb[1] << 8);
        }
    }
}

template <int AT>
__global__ void __launch_bounds__(NT) k_op_attn(const uint8_t* __restrict__ q, int n_embd, KVCache kv, int gsz, int n_ctx,
                                                 int start_pos, float* __restrict__ tmp) {
    extern __shared__ __align__(16) unsigned char smem[];
    AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem);
    float* sc = reinterpret_cast<float*>(smem + ((sizeof(AttnSmem) + 15) & ~(size_t)15));
    const int h = blockIdx.x, row = start_pos + blockIdx.y, g = h / gsz;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint8_t* qrow = q + (size_t)row * row_nbytes(AT, n_embd);
    if (wid < 2) {
        const int e = h * 64 + wid * 32 + lane;
        if (AT == DT_F16) sm.qf[wid * 32 + lane] = h2f(reinterpret_cast<const uint16_t*>(qrow)[e]);
        else {
            const uint8_t* blk = qrow + (size_t)(e >> 5) * Q8_BYTES;
            reinterpret_cast<int8_t*>(sm.qw)[wid * 32 + perm_byte(lane)] = (int8_t)blk[2 + lane];
            if (lane == 0) sm.qd[wid] = h2f((uint16_t)blk[0] | ((uint16_t)blk[1] << 8));
        }
    }
    __syncthreads();
    attn_core<AT>(sm, sc, kv, g, row, n_ctx, false, tmp + ((size_t)blockIdx.y * n_embd + h * 64));
}

// host-built RoPE table shared by the op-level entry point (gten/ops.h:728-746 with the reference's libm calls)
struct RopeTable { float* cs = nullptr; float* sn = nullptr; int rows = 0; };
static std::map<int, RopeTable>& rope_tables() { static std::map<int, RopeTable> t; return t; }

static int rope_table(int d_head, int rows, const float** cs, const float** sn) {
    RopeTable& t = rope_tables()[d_head];
    if (t.rows < rows) {
        const int dh = d_head / 2;
        int newrows = rows < 256 ? 256 : rows * 2;
        std::vector<float> hc((size_t)newrows * dh), hs((size_t)newrows * dh);
        const float d = static_cast<float>(d_head);
        for (int p = 0; p < newrows; p++) {
            const float m = static_cast<float>(p);
            for (int j = 0; j < dh; j++) {
                const float m_theta_i = m * powf(10000.0f, -(2.0f * j / d));
                hc[(size_t)p * dh + j] = cosf(m_theta_i);
                hs[(size_t)p * dh + j] = sinf(m_theta_i);
            }
        }
        cudaStreamSynchronize(ctx().stream);
        if (t.cs) { cudaFree(t.cs); cudaFree(t.sn); }
        GTB_CUDA(cudaMalloc((void**)&t.cs, hc.size() * 4));
        GTB_CUDA(cudaMalloc((void**)&t.sn, hs.size() * 4));
        GTB_CUDA(cudaMemcpy(t.cs, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
        GTB_CUDA(cudaMemcpy(t.sn, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
        t.rows = newrows;
    }
    *cs = t.cs; *sn = t.sn;
    return GTB_OK;
}

static int rowwise(int op, const void* a, const void* b, void* out, int dtype, int n_ctx, int n, int start_pos, int d_head,
                   const float* rc, const float* rs, const void* normw) {
    GTB_ARG(a && out && n_ctx > 0 && n > 0 && start_pos >= 0 && start_pos <= n_ctx);
    GTB_ARG(dtype == GTB_Q8 || dtype == GTB_F16 || dtype == GTB_F32);
    if (start_pos == n_ctx) return GTB_OK;
    const size_t smem = (((size_t)2 * n * 4 + 15) & ~(size_t)15) + sizeof(ExactSumSmem) + 16;
    static bool attr = false;
    if (!attr) { GTB_CUDA(cudaFuncSetAttribute(k_op_rowwise, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    GTB_ARG(smem <= 200 * 1024);
    k_op_rowwise<<<n_ctx - start_pos, NT, smem, ctx().stream>>>(op, (const uint8_t*)a, (const uint8_t*)b, (uint8_t*)out, dtype, n,
                                                               start_pos, d_head, rc, rs, (const uint16_t*)normw);
    GTB_LAUNCHED();
    return GTB_OK;
}

template <int WT>
static int matmul_launch(const void* d_x, int n_ctx, gtb_weight_t w, float* tmp, int start_pos) {
    constexpr int AT = (WT == DT_F16) ? DT_F16 : DT_Q8;
    const int K = w->cols;
    const size_t smem = ((act_bytes(AT, K) + 15) & ~(size_t)15) + gemv_ps_bytes(WT, K) + 16;
    static bool attr = false;
    if (!attr) { GTB_CUDA(cudaFuncSetAttribute(k_op_matmul<WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    const int rows = n_ctx - start_pos;
    int gx = (w->rows + RPW * NWARP - 1) / (RPW * NWARP);
    const int cap = (ctx().sm_count * 2 + rows - 1) / rows;
    if (gx > cap) gx = cap < 1 ? 1 : cap;
    k_op_matmul<WT><<<dim3(gx, rows), NT, smem, ctx().stream>>>((const uint8_t*)d_x, K, w->data, w->scales, w->rows, tmp, start_pos);
    GTB_LAUNCHED();
    return GTB_OK;
}

}  // namespace gtb

using namespace gtb;


// ---------------------------------------------------------------- ops::vec_dot_product (gten/ops.h:482-512) on two rows
// One thread walks the reference's AVX evaluation order; rows are in the reference's own layout.
__global__ void k_vec_dot(const uint8_t* __restrict__ a, int adt, const uint8_t* __restrict__ b, int bdt, int n, float* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float r = 0.0f;
    if (adt == DT_Q8) {
        // ops.h:224-292 (Q8 x Q8) / 319-391 (Q8 x Q4): lane l = elements {2l, 2l+1, 2l+8, 2l+9} and the same + 16
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int blk = 0; blk < n / 32; blk++) {
            const uint8_t* ab = a + (size_t)blk * Q8_BYTES;
            const uint8_t* bb = b + (size_t)blk * ((bdt == DT_Q4) ? Q4_BYTES : Q8_BYTES);
            const float da = h2f((uint16_t)ab[0] | ((uint16_t)ab[1] << 8)), db = h2f((uint16_t)bb[0] | ((uint16_t)bb[1] << 8));
            const float s = __fmul_rn(da, db);
            for (int l = 0; l < 4; l++) {
                int sum = 0;
                for (int half = 0; half < 2; half++)
                    for (int j = 0; j < 4; j++) {
                        const int e = 16 * half + 2 * l + (j & 1) + 8 * (j >> 1);
                        const int qa = (int)(int8_t)ab[2 + e];
                        int qb;
                        if (bdt == DT_Q4) { const uint8_t by = bb[2 + (e & 15)]; qb = (int)((e < 16) ? (by >> 4) : (by & 0x0f)) - 7; }
                        else qb = (int)(int8_t)bb[2 + e];
                        sum += qa * qb;
                    }
                acc[l] = __fadd_rn(acc[l], __fmul_rn((float)sum, s));
            }
        }
        r = __fadd_rn(__fadd_rn(acc[0], acc[1]), __fadd_rn(acc[2], acc[3]));
    } else {
        // ops.h:140-160 (fp16) / 177-197 (fp32): eight lane accumulators over elements 8i + l, mul and add rounded separately
        // (simd_ops.h:59-61), lanes summed left to right (simd_ops.h:63-66), then the tail in order
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int n8 = (n / 8) * 8;
        auto ld = [&](const uint8_t* p, int dt, int i) { return dt == DT_F16 ? h2f(reinterpret_cast<const uint16_t*>(p)[i]) : reinterpret_cast<const float*>(p)[i]; };
        for (int i = 0; i < n8; i += 8)
            for (int l = 0; l < 8; l++) acc[l] = __fadd_rn(__fmul_rn(ld(a, adt, i + l), ld(b, bdt, i + l)), acc[l]);
        r = __fadd_rn(acc[0], acc[1]);
        for (int l = 2; l < 8; l++) r = __fadd_rn(r, acc[l]);
        for (int i = n8; i < n; i++) r = __fadd_rn(r, __fmul_rn(ld(a, adt, i), ld(b, bdt, i)));
    }
    *out = r;
}

extern "C" {

int gtb_token_embed(gtb_weight_t w, const int32_t* d_tokens, void* d_out, int out_dtype, int n_ctx, int start_pos) {
    GTB_CHECK_INIT();
    GTB_ARG(w && d_tokens && d_out && n_ctx > 0 && start_pos >= 0 && start_pos <= n_ctx);
    GTB_ARG(out_dtype == ((w->dtype == GTB_F16) ? GTB_F16 : GTB_Q8));      // ops.h:523
    if (start_pos == n_ctx) return GTB_OK;
    k_op_embed<<<n_ctx - start_pos, NT, (size_t)w->cols * 4 + 16, ctx().stream>>>(w->data, w->scales, w->dtype, w->cols, d_tokens,
                                                                                (uint8_t*)d_out, out_dtype, start_pos);
    GTB_LAUNCHED();
    return GTB_OK;
}

int gtb_matmul_2d(const void* d_x, int x_dtype, int n_ctx, gtb_weight_t w, void* d_out, int out_dtype, int out_is_1d, int start_pos) {
    GTB_CHECK_INIT();
    GTB_ARG(d_x && w && d_out && n_ctx > 0 && start_pos >= 0 && start_pos <= n_ctx);
    GTB_ARG(x_dtype == ((w->dtype == GTB_F16) ? GTB_F16 : GTB_Q8));
    GTB_ARG(out_dtype == GTB_Q8 || out_dtype == GTB_F16 || out_dtype == GTB_F32);
    GTB_ARG(!out_is_1d || n_ctx - start_pos == 1);                         // ops.h:660-662
    if (start_pos == n_ctx) return GTB_OK;
    const int rows = n_ctx - start_pos;
    float* tmp = nullptr;
    GTB_CUDA(cudaMallocAsync((void**)&tmp, (size_t)rows * w->rows * 4, ctx().stream));
    int r;
    switch (w->dtype) {
        case GTB_F16: r = matmul_launch<DT_F16>(d_x, n_ctx, w, tmp, start_pos); break;
        case GTB_Q8: r = matmul_launch<DT_Q8>(d_x, n_ctx, w, tmp, start_pos); break;
        default: r = matmul_launch<DT_Q4>(d_x, n_ctx, w, tmp, start_pos); break;
    }
    if (r == GTB_OK) {
        uint8_t* o = (uint8_t*)d_out + (out_is_1d ? 0 : (size_t)start_pos * row_nbytes(out_dtype, w->rows));
        r = gtb_write_rows_from_float(tmp, o, out_dtype, rows, w->rows);      // ops.h:645-646
    }
    cudaFreeAsync(tmp, ctx().stream);
    return r;
}

int gtb_rms_norm(const void* d_x, int dtype, int n_ctx, int n_embd, const void* d_weight_f16, void* d_out, int start_pos) {
    GTB_CHECK_INIT();
    GTB_ARG(d_weight_f16);
    return rowwise(EW_NORM, d_x, nullptr, d_out, dtype, n_ctx, n_embd, start_pos, 0, nullptr, nullptr, d_weight_f16);
}

int gtb_rotary_emb(void* d_x, int dtype, int n_ctx, int n_embd, int d_head, int start_pos) {
    GTB_CHECK_INIT();
    GTB_ARG(d_head > 0 && d_head % 2 == 0 && n_embd % d_head == 0);
    const float *rc, *rs;
    int r = rope_table(d_head, n_ctx, &rc, &rs);
    if (r) return r;
    return rowwise(EW_ROPE, d_x, nullptr, d_x, dtype, n_ctx, n_embd, start_pos, d_head, rc, rs, nullptr);
}

int gtb_silu(const void* d_x, int dtype, int n_ctx, int n_embd, void* d_out, int start_pos) {
    GTB_CHECK_INIT();
    return rowwise(EW_SILU, d_x, nullptr, d_out, dtype, n_ctx, n_embd, start_pos, 0, nullptr, nullptr, nullptr);
}

int gtb_mul(const void* d_a, const void* d_b, int dtype, int n_ctx, int n_embd, void* d_out, int start_pos) {
    GTB_CHECK_INIT();
    GTB_ARG(d_b);
    return rowwise(EW_MUL, d_a, d_b, d_out, dtype, n_ctx, n_embd, start_pos, 0, nullptr, nullptr, nullptr);
}

int gtb_add(const void* d_a, const void* d_b, int dtype, int n_ctx, int n_embd, void* d_out, int start_pos) {
    GTB_CHECK_INIT();
    GTB_ARG(d_b);
    return rowwise(EW_ADD, d_a, d_b, d_out, dtype, n_ctx, n_embd, start_pos, 0, nullptr, nullptr, nullptr);
}

int gtb_qkv_attn(const void* d_q, const void* d_k, const void* d_v, void* d_qk, void* d_out, int dtype,
                 int n_ctx, int n_heads, int n_kv_heads, int d_head, int max_ctx, int start_pos) {
    GTB_CHECK_INIT();
    (void)d_qk;                                             // scores stay on chip; the reference's qk buffer is scratch
    GTB_ARG(d_q && d_k && d_v && d_out && n_ctx > 0 && start_pos >= 0 && start_pos <= n_ctx);
    GTB_ARG(dtype == GTB_Q8 || dtype == GTB_F16);
    GTB_ARG(d_head == 64 && n_heads % n_kv_heads == 0 && max_ctx >= n_ctx);   // ops.h:1130
    if (start_pos == n_ctx) return GTB_OK;
    const int n_embd = n_heads * 64, kv_dim = n_kv_heads * 64, rows = n_ctx - start_pos;
    cudaStream_t st = ctx().stream;
    uint8_t *kq = nullptr, *vq = nullptr;
    uint16_t *ks = nullptr, *vs = nullptr;
    float* tmp = nullptr;
    const size_t cb = (size_t)n_ctx * kv_dim * (dtype == GTB_F16 ? 2 : 1);
    GTB_CUDA(cudaMallocAsync((void**)&kq, cb, st));
    GTB_CUDA(cudaMallocAsync((void**)&vq, cb, st));
    GTB_CUDA(cudaMallocAsync((void**)&ks, (size_t)n_ctx * (kv_dim / 32) * 2, st));
    GTB_CUDA(cudaMallocAsync((void**)&vs, (size_t)n_ctx * (kv_dim / 32) * 2, st));
    GTB_CUDA(cudaMallocAsync((void**)&tmp, (size_t)rows * n_embd * 4, st));
    const size_t tot = (size_t)n_ctx * kv_dim;
    const size_t smem = ((sizeof(AttnSmem) + 15) & ~(size_t)15) + (size_t)((n_ctx + 63) / 32 * 32) * 4;
    KVCache kv{kq, ks, vq, vs, kv_dim};
    if (dtype == GTB_F16) {
        k_op_kv_convert<DT_F16><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const uint8_t*)d_k, (const uint8_t*)d_v, n_ctx, kv_dim, kq, ks, vq, vs);
        static bool attr = false;
        if (!attr) { GTB_CUDA(cudaFuncSetAttribute(k_op_attn<DT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
        k_op_attn<DT_F16><<<dim3(n_heads, rows), NT, smem, st>>>((const uint8_t*)d_q, n_embd, kv, n_heads / n_kv_heads, n_ctx, start_pos, tmp);
    } else {
        k_op_kv_convert<DT_Q8><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const uint8_t*)d_k, (const uint8_t*)d_v, n_ctx, kv_dim, kq, ks, vq, vs);
        static bool attr = false;
        if (!attr) { GTB_CUDA(cudaFuncSetAttribute(k_op_attn<DT_Q8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr = true; }
        k_op_attn<DT_Q8><<<dim3(n_heads, rows), NT, smem, st>>>((const uint8_t*)d_q, n_embd, kv, n_heads / n_kv_heads, n_ctx, start_pos, tmp);
    }
    ctx().launches += 2;
    cudaError_t le = cudaGetLastError();
    int r = GTB_OK;
    if (le != cudaSuccess) r = fail(GTB_ERR_CUDA, "attention launch failed: %s", cudaGetErrorString(le));
    if (r == GTB_OK)
        r = gtb_write_rows_from_float(tmp, (uint8_t*)d_out + (size_t)start_pos * row_nbytes(dtype, n_embd), dtype, rows, n_embd);
    cudaFreeAsync(kq, st); cudaFreeAsync(vq, st); cudaFreeAsync(ks, st); cudaFreeAsync(vs, st); cudaFreeAsync(tmp, st);
    return r;
}

int gtb_vec_dot_product(const void* h_a, int a_dtype, const void* h_b, int b_dtype, int n, float* h_out) {
    GTB_CHECK_INIT();
    GTB_ARG(h_a && h_b && h_out && n > 0);
    const bool q = a_dtype == GTB_Q8 && (b_dtype == GTB_Q8 || b_dtype == GTB_Q4);
    const bool f = (a_dtype == GTB_F16 && b_dtype == GTB_F16) || (a_dtype == GTB_F32 && b_dtype == GTB_F32);
    if (!q && !f) return fail(GTB_ERR_ARG, "vec_dot_product: unsupported dtype pair (%d, %d)", a_dtype, b_dtype);   // ops.h:506-509 asserts
    if (q && n % 32 != 0) return fail(GTB_ERR_ARG, "vec_dot_product: %d is not a multiple of the block size", n);   // ops.h:229
    const size_t na = row_nbytes(a_dtype, n), nb = row_nbytes(b_dtype, n);
    uint8_t* d = nullptr;
    const size_t oa = 0, ob = (na + 15) & ~(size_t)15, oo = ob + ((nb + 15) & ~(size_t)15);
    GTB_CUDA(cudaMalloc((void**)&d, oo + 16));
    cudaStream_t st = ctx().stream;
    cudaError_t ce = cudaMemcpyAsync(d + oa, h_a, na, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d + ob, h_b, nb, cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) {
        k_vec_dot<<<1, 32, 0, st>>>(d + oa, a_dtype, d + ob, b_dtype, n, reinterpret_cast<float*>(d + oo));
        ctx().launches++;
        ce = cudaMemcpyAsync(h_out, d + oo, 4, cudaMemcpyDeviceToHost, st);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(d);
    if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "vec_dot_product failed: %s", cudaGetErrorString(ce));
    return GTB_OK;
}

}  // extern "C"
