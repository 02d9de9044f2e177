// gtb_kernels.cuh -- device building blocks of the decode path: activation staging (the reference's
// per-op re-encode points), order-exact GEMV warp passes, and the GQA attention core.
//
// Work decomposition of one decode row ("phase" = one dependent step, one kernel in the stream/graph):
//   P1  [x' = E(h + E(down)); n1 = E(rmsnorm(x'))]        -> q|k|v GEMV          -> raw fp32
//   P2  [E, RoPE, E on q,k; E on v; append K/V]            -> scores, softmax, P.V -> raw fp32
//   P3  [E(attn)]                                          -> o GEMV              -> raw fp32
//   P4  [h = E(x' + E(o)); n2 = E(rmsnorm(h))]             -> gate|up GEMV        -> raw fp32
//   P5  [E(E(silu(E(gate))) * E(up))]                      -> down GEMV           -> raw fp32
// E() is the activation re-encode (Q8 blocks or fp16) the reference applies after every op; the bracketed
// prologues are recomputed by every CTA (they are tiny) so that no extra grid-wide step is needed.
#pragma once
#include "gtb_dev.cuh"

namespace gtb {

struct DevState {
    int pos;        // row to process next
    int nctx_min;   // n_ctx of the current logits() call (P row length / P.V lane split), SURVEY App. A
    int stop;       // set when eos was generated
    int n_gen;
};

constexpr int NT = 256;          // threads per CTA for every phase kernel
constexpr int NWARP = NT / 32;
constexpr int RPW = 4;           // weight rows per warp pass (Q4/Q8)

__host__ __device__ inline int nb_pad_of(int nb) { return ((nb + 27) / 32) * 32 + 4; }   // == 4 (mod 32): conflict-free LDS.128 chains

// ---------------------------------------------------------------- staged activation vector (shared memory)
// Q8: per 32-block, eight words in the weight layout's order (X0..X3 = codes (2l,2l+1,2l+8,2l+9), Y = +16),
//     the fp32 value of the fp16 block scale, and (for Q4 weights) -7 * sum of the eight codes of each lane.
// F16: fp32 values of the fp16-rounded activations, transposed so that lane l of chunk c is contiguous.
struct ActView {
    uint32_t* aw;    // [nb][8]
    float* ad;       // [nb]
    int* ns7;        // [nb][4]
    float* xs;       // [K] (F16 only)
};

__host__ __device__ inline size_t act_bytes(int at, int K) {
    if (at == DT_F16) return (size_t)K * 4;
    const int nb = K / 32;
    return (size_t)nb * 32 + (size_t)nb * 4 + (size_t)nb * 16;
}
__device__ inline ActView act_carve(int at, int K, unsigned char* base) {
    ActView v{nullptr, nullptr, nullptr, nullptr};
    if (at == DT_F16) { v.xs = reinterpret_cast<float*>(base); return v; }
    const int nb = K / 32;
    v.aw = reinterpret_cast<uint32_t*>(base);
    v.ns7 = reinterpret_cast<int*>(base + (size_t)nb * 32);
    v.ad = reinterpret_cast<float*>(base + (size_t)nb * 32 + (size_t)nb * 16);
    return v;
}

__device__ __forceinline__ int perm_byte(int e) {       // natural element e of a block -> byte offset in the 32-byte permuted block
    const int half = e >> 4, j = e & 15;
    return half * 16 + ((j & 7) >> 1) * 4 + (j & 1) + 2 * (j >> 3);
}

// Encode block b (one element per lane, x = pre-encode value) into the staged vector; returns the decoded value.
template <int AT>
__device__ __forceinline__ float stage_block(const ActView& av, int b, int lane, float x) {
    if (AT == DT_F16) {
        const float d = f16_roundtrip(x);
        const int e = b * 32 + lane;
        av.xs[(((e >> 6) * 8) + (e & 7)) * 8 + ((e >> 3) & 7)] = d;
        return d;
    } else {
        uint16_t dh;
        const int q = q8_encode_lane(x, &dh);
        reinterpret_cast<int8_t*>(av.aw)[b * 32 + perm_byte(lane)] = (int8_t)q;
        int s = q + __shfl_xor_sync(0xffffffffu, q, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        const float delta = h2f(dh);
        if (lane < 8 && !(lane & 1)) av.ns7[b * 4 + (lane >> 1)] = -7 * s;
        if (lane == 0) av.ad[b] = delta;
        return __fmul_rn((float)q, delta);
    }
}

// Stage an ALREADY ENCODED block (reference row layout) -- used by the op-level matmul.
template <int AT>
__device__ __forceinline__ void stage_encoded_block(const ActView& av, int b, int lane, const uint8_t* row) {
    if (AT == DT_F16) {
        const int e = b * 32 + lane;
        av.xs[(((e >> 6) * 8) + (e & 7)) * 8 + ((e >> 3) & 7)] = h2f(reinterpret_cast<const uint16_t*>(row)[e]);
    } else {
        const uint8_t* blk = row + (size_t)b * Q8_BYTES;
        const int q = (int)(int8_t)blk[2 + lane];
        reinterpret_cast<int8_t*>(av.aw)[b * 32 + perm_byte(lane)] = (int8_t)q;
        int s = q + __shfl_xor_sync(0xffffffffu, q, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (lane < 8 && !(lane & 1)) av.ns7[b * 4 + (lane >> 1)] = -7 * s;
        if (lane == 0) av.ad[b] = h2f((uint16_t)blk[0] | ((uint16_t)blk[1] << 8));
    }
}

template <int AT>
__device__ __forceinline__ float roundtrip(float x) {     // E(): what the next op reads back
    return (AT == DT_F16) ? f16_roundtrip(x) : q8_roundtrip_lane(x);
}

// ---------------------------------------------------------------- prologues (whole CTA, n % 32 == 0)
// PRO_ENCODE: stage E(src0).
template <int AT>
__device__ void pro_encode(const ActView& av, const float* __restrict__ src, int n, float* cap) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int b = wid; b < n / 32; b += nwarp) {
        const float d = stage_block<AT>(av, b, lane, src[b * 32 + lane]);
        if (cap) cap[b * 32 + lane] = d;
    }
}

// PRO_NORM (gten/ops.h:762-804, 870-898): x = delta ? E(res + E(delta)) : res;  stage E(x / (rms(x) + 1e-6) * w).
// xbuf: n floats of scratch.  res_out (CTA 0 only) receives x; caps are optional decoded dumps.
template <int AT>
__device__ void pro_norm(const ActView& av, const float* __restrict__ res, const float* __restrict__ delta,
                         const uint16_t* __restrict__ normw, int n, float* xbuf, ExactSumSmem& es,
                         float* res_out, float* cap_delta, float* cap_res, float* cap_norm) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int b = wid; b < n / 32; b += nwarp) {
        const int e = b * 32 + lane;
        float x = res[e];
        if (delta) {
            const float d1 = roundtrip<AT>(delta[e]);
            if (cap_delta) cap_delta[e] = d1;
            x = roundtrip<AT>(__fadd_rn(x, d1));
        }
        xbuf[e] = x;
        if (res_out) res_out[e] = x;
        if (cap_res) cap_res[e] = x;
    }
    __syncthreads();
    const float sq_sum = exact_sum_block([&](int i) { const float v = xbuf[i]; return __fmul_rn(v, v); }, n, es);
    const float rms = sqrtf(__fdiv_rn(sq_sum, (float)n));
    const float denom = __fadd_rn(rms, 1e-6f);
    for (int b = wid; b < n / 32; b += nwarp) {
        const int e = b * 32 + lane;
        const float y = __fmul_rn(__fdiv_rn(xbuf[e], denom), h2f(normw[e]));
        const float d = stage_block<AT>(av, b, lane, y);
        if (cap_norm) cap_norm[e] = d;
    }
}

__device__ __forceinline__ float silu_ref(float x) {      // gten/ops.h:692
    return __fdiv_rn(x, __fadd_rn(1.0f, expf_glibc(-x)));
}

// PRO_SILU_MUL (gten/modules.cpp:238-247): stage E(E(silu(E(gate))) * E(up)).
template <int AT>
__device__ void pro_silu_mul(const ActView& av, const float* __restrict__ gate, const float* __restrict__ up, int n,
                             float* cap_gate, float* cap_up) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b = wid; b < n / 32; b += NWARP) {
        const int e = b * 32 + lane;
        const float g1 = roundtrip<AT>(gate[e]);
        const float u1 = roundtrip<AT>(up[e]);
        const float g2 = roundtrip<AT>(silu_ref(g1));
        const float g3 = stage_block<AT>(av, b, lane, __fmul_rn(g2, u1));
        if (cap_gate) cap_gate[e] = g3;
        if (cap_up) cap_up[e] = u1;
    }
}

// ---------------------------------------------------------------- order-exact GEMV warp passes
// Q4/Q8 weights (gten/ops.h:224-292, 319-391): per block the four integer lane sums are exact, then
// acc[l] = acc[l] + float(lane[l]) * (da*dw) in ascending block order, result (a0+a1)+(a2+a3).
// A warp takes RPW rows: every lane forms the products of whole (row, block) items from coalesced 128-bit
// loads, parks them in its warp's shared scratch, then lane (row, l) runs the ordered add chain.
template <int WT>
__device__ __forceinline__ void block_products(const uint4* __restrict__ wdata, const uint16_t* __restrict__ wsc,
                                               size_t blk, const ActView& av, int b, float p[4]) {
    const float dw = h2f(wsc[blk]);
    const uint4 ax = reinterpret_cast<const uint4*>(av.aw)[2 * b];
    const uint4 ay = reinterpret_cast<const uint4*>(av.aw)[2 * b + 1];
    const float s = __fmul_rn(av.ad[b], dw);
    int li[4];
    if (WT == DT_Q4) {
        const uint4 w = wdata[blk];
        const int4 n7 = reinterpret_cast<const int4*>(av.ns7)[b];
        li[0] = __dp4a((int)(w.x & 0x0f0f0f0fu), (int)ay.x, __dp4a((int)((w.x >> 4) & 0x0f0f0f0fu), (int)ax.x, n7.x));
        li[1] = __dp4a((int)(w.y & 0x0f0f0f0fu), (int)ay.y, __dp4a((int)((w.y >> 4) & 0x0f0f0f0fu), (int)ax.y, n7.y));
        li[2] = __dp4a((int)(w.z & 0x0f0f0f0fu), (int)ay.z, __dp4a((int)((w.z >> 4) & 0x0f0f0f0fu), (int)ax.z, n7.z));
        li[3] = __dp4a((int)(w.w & 0x0f0f0f0fu), (int)ay.w, __dp4a((int)((w.w >> 4) & 0x0f0f0f0fu), (int)ax.w, n7.w));
    } else {
        const uint4 wx = wdata[2 * blk], wy = wdata[2 * blk + 1];
        li[0] = __dp4a((int)wy.x, (int)ay.x, __dp4a((int)wx.x, (int)ax.x, 0));
        li[1] = __dp4a((int)wy.y, (int)ay.y, __dp4a((int)wx.y, (int)ax.y, 0));
        li[2] = __dp4a((int)wy.z, (int)ay.z, __dp4a((int)wx.z, (int)ax.z, 0));
        li[3] = __dp4a((int)wy.w, (int)ay.w, __dp4a((int)wx.w, (int)ax.w, 0));
    }
#pragma unroll
    for (int l = 0; l < 4; l++) p[l] = __fmul_rn((float)li[l], s);
}

template <int WT>
__device__ void gemv_q_warp_pass(const uint4* __restrict__ wdata, const uint16_t* __restrict__ wsc, int nb,
                                 int row0, int nrows, const ActView& av, float* ps, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int nbp = nb_pad_of(nb);
    const int items = nrows * nb;
    for (int it = lane; it < items; it += 32) {
        const int rl = it / nb, b = it - rl * nb;
        float p[4];
        block_products<WT>(wdata, wsc, (size_t)(row0 + rl) * nb + b, av, b, p);
#pragma unroll
        for (int l = 0; l < 4; l++) ps[(rl * 4 + l) * nbp + b] = p[l];
    }
    __syncwarp();
    float acc = 0.0f;
    if (lane < nrows * 4) {
        const float4* src = reinterpret_cast<const float4*>(ps + lane * nbp);
        for (int b4 = 0; b4 < nb / 4; b4++) {
            const float4 v = src[b4];
            acc = __fadd_rn(acc, v.x); acc = __fadd_rn(acc, v.y); acc = __fadd_rn(acc, v.z); acc = __fadd_rn(acc, v.w);
        }
        for (int b = (nb / 4) * 4; b < nb; b++) acc = __fadd_rn(acc, ps[lane * nbp + b]);
    }
    const float v1 = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    const float v2 = __fadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, 2));
    if ((lane & 3) == 0 && lane < nrows * 4) out[row0 + (lane >> 2)] = v2;
    __syncwarp();
}

// FP16 weights (gten/ops.h:140-160): lane l accumulates elements 8i+l in ascending i (products of fp16 values
// are exact in fp32, so one fused multiply-add == the reference's separate mul and add); lanes summed left to
// right.  A warp takes 4 rows x 8 lanes and streams its rows with 128-bit loads of the lane-major layout.
__device__ inline void gemv_f16_warp_pass(const uint4* __restrict__ wdata, int K, int row0, int nrows, const ActView& av,
                                   float* __restrict__ out) {
    const int lane = threadIdx.x & 31, rl = lane >> 3, l = lane & 7;
    const int cpr = K / 64;
    const bool active = rl < nrows;
    const uint4* src = wdata + ((size_t)(row0 + (active ? rl : 0)) * cpr) * 8 + l;
    float acc = 0.0f;
    constexpr int UN = 8;
    int c = 0;
    for (; c + UN <= cpr; c += UN) {
        uint4 w[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) w[u] = src[(size_t)(c + u) * 8];
#pragma unroll
        for (int u = 0; u < UN; u++) {
            const float4 x0 = reinterpret_cast<const float4*>(av.xs)[((c + u) * 8 + l) * 2];
            const float4 x1 = reinterpret_cast<const float4*>(av.xs)[((c + u) * 8 + l) * 2 + 1];
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w[u].x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&w[u].y));
            const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&w[u].z));
            const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&w[u].w));
            acc = fmaf(x0.x, a.x, acc); acc = fmaf(x0.y, a.y, acc); acc = fmaf(x0.z, b.x, acc); acc = fmaf(x0.w, b.y, acc);
            acc = fmaf(x1.x, d.x, acc); acc = fmaf(x1.y, d.y, acc); acc = fmaf(x1.z, e.x, acc); acc = fmaf(x1.w, e.y, acc);
        }
    }
    for (; c < cpr; c++) {
        const uint4 w = src[(size_t)c * 8];
        const float* x = av.xs + (c * 8 + l) * 8;
        const __half* hw = reinterpret_cast<const __half*>(&w);
#pragma unroll
        for (int i = 0; i < 8; i++) acc = fmaf(x[i], __half2float(hw[i]), acc);
    }
    float v = __shfl_sync(0xffffffffu, acc, lane & ~7);
#pragma unroll
    for (int j = 1; j < 8; j++) v = __fadd_rn(v, __shfl_sync(0xffffffffu, acc, (lane & ~7) + j));
    if (l == 0 && active) out[row0 + rl] = v;
}

// One CTA's share of a GEMV over a matrix: warp passes are dealt round-robin over all warps of the grid.
template <int WT>
__device__ void gemv_matrix(const void* __restrict__ wdata, const uint16_t* __restrict__ wsc, int rows, int K,
                            const ActView& av, float* ps_warp, float* __restrict__ out, int pass_offset, int* passes_done) {
    const int wid = threadIdx.x >> 5;
    const int gw = wid * gridDim.x + blockIdx.x;             // spread consecutive passes over SMs first
    const int total_warps = gridDim.x * NWARP;
    const int npass = (rows + RPW - 1) / RPW;
    // continue the round-robin where the previous matrix of this phase stopped
    int first = (gw - (pass_offset % total_warps) + total_warps) % total_warps;
    for (int p = first; p < npass; p += total_warps) {
        const int row0 = p * RPW;
        const int nrows = min(RPW, rows - row0);
        if (WT == DT_F16) gemv_f16_warp_pass(reinterpret_cast<const uint4*>(wdata), K, row0, nrows, av, out);
        else gemv_q_warp_pass<WT>(reinterpret_cast<const uint4*>(wdata), wsc, K / 32, row0, nrows, av, ps_warp, out);
    }
    *passes_done = pass_offset + npass;
}

__host__ __device__ inline size_t gemv_ps_bytes(int wt, int K) {
    if (wt == DT_F16) return 0;
    return (size_t)NWARP * RPW * 4 * nb_pad_of(K / 32) * 4;
}

// ---------------------------------------------------------------- attention core (one CTA = one query head of one row)
// K cache: per position, Q8: permuted words (same order as staged activations) + fp16 scales; F16: halves.
// V cache: natural order.  q staged in shared memory by the caller.
struct KVCache {
    const uint8_t* kq;       // Q8: int8 [max_ctx][kv_dim] permuted per block | F16: half [max_ctx][kv_dim]
    const uint16_t* ks;      // Q8: fp16 [max_ctx][kv_dim/32]
    const uint8_t* vq;       // Q8: int8 [max_ctx][kv_dim] natural            | F16: half [max_ctx][kv_dim]
    const uint16_t* vs;      // Q8: fp16 [max_ctx][kv_dim/32]
    int kv_dim;
};

struct AttnSmem {
    // staged q of this head: Q8 -> 16 words + 2 scales; F16 -> 64 floats
    uint32_t qw[16];
    float qd[2];
    float qf[64];
    // this row's own k / v for the head's group (position `pos` is not read back from the cache)
    uint32_t kw[16];
    float kd[2];
    float kf[64];
    float vf[64];            // decoded v of this row
    float part[8][64];
    float red[NWARP];
    float tmp[6][32];
    ExactSumSmem es;
};

template <int AT>
__device__ __forceinline__ float score_one(const AttnSmem& sm, const KVCache& kv, int g, int kcol, int pos) {
    if (AT == DT_F16) {
        float acc[8];
#pragma unroll
        for (int l = 0; l < 8; l++) acc[l] = 0.0f;
        if (kcol == pos) {
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int l = 0; l < 8; l++) acc[l] = fmaf(sm.qf[8 * i + l], sm.kf[8 * i + l], acc[l]);
        } else {
            const uint4* kp = reinterpret_cast<const uint4*>(kv.kq + ((size_t)kcol * kv.kv_dim + g * 64) * 2);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint4 w = kp[i];
                const __half* hw = reinterpret_cast<const __half*>(&w);
#pragma unroll
                for (int l = 0; l < 8; l++) acc[l] = fmaf(sm.qf[8 * i + l], __half2float(hw[l]), acc[l]);
            }
        }
        float d = __fadd_rn(acc[0], acc[1]);
#pragma unroll
        for (int l = 2; l < 8; l++) d = __fadd_rn(d, acc[l]);
        return d;
    } else {
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int bi = 0; bi < 2; bi++) {
            uint4 kx, ky;
            float kdv;
            if (kcol == pos) {
                kx = make_uint4(sm.kw[bi * 8 + 0], sm.kw[bi * 8 + 1], sm.kw[bi * 8 + 2], sm.kw[bi * 8 + 3]);
                ky = make_uint4(sm.kw[bi * 8 + 4], sm.kw[bi * 8 + 5], sm.kw[bi * 8 + 6], sm.kw[bi * 8 + 7]);
                kdv = sm.kd[bi];
            } else {
                const uint4* kp = reinterpret_cast<const uint4*>(kv.kq + (size_t)kcol * kv.kv_dim + g * 64 + bi * 32);
                kx = kp[0]; ky = kp[1];
                kdv = h2f(kv.ks[(size_t)kcol * (kv.kv_dim / 32) + g * 2 + bi]);
            }
            const float s = __fmul_rn(sm.qd[bi], kdv);
            const uint32_t* q = sm.qw + bi * 8;
            const int l0 = __dp4a((int)ky.x, (int)q[4], __dp4a((int)kx.x, (int)q[0], 0));
            const int l1 = __dp4a((int)ky.y, (int)q[5], __dp4a((int)kx.y, (int)q[1], 0));
            const int l2 = __dp4a((int)ky.z, (int)q[6], __dp4a((int)kx.z, (int)q[2], 0));
            const int l3 = __dp4a((int)ky.w, (int)q[7], __dp4a((int)kx.w, (int)q[3], 0));
            acc[0] = __fadd_rn(acc[0], __fmul_rn((float)l0, s));
            acc[1] = __fadd_rn(acc[1], __fmul_rn((float)l1, s));
            acc[2] = __fadd_rn(acc[2], __fmul_rn((float)l2, s));
            acc[3] = __fadd_rn(acc[3], __fmul_rn((float)l3, s));
        }
        return __fadd_rn(__fadd_rn(acc[0], acc[1]), __fadd_rn(acc[2], acc[3]));
    }
}

template <int AT>
__device__ __forceinline__ float v_at(const AttnSmem& sm, const KVCache& kv, int g, int i, int c, int pos) {
    if (i == pos) return sm.vf[c];
    if (AT == DT_F16) return h2f(reinterpret_cast<const uint16_t*>(kv.vq)[(size_t)i * kv.kv_dim + g * 64 + c]);
    const float delta = h2f(kv.vs[(size_t)i * (kv.kv_dim / 32) + g * 2 + (c >> 5)]);
    return __fmul_rn((float)(int8_t)kv.vq[(size_t)i * kv.kv_dim + g * 64 + c], delta);       // ops.h:1026
}

// sc: shared float [>= n_ctx rounded up to 32].  q (and, when own_kv, this row's k/v) already staged in sm.
// Writes the 64 raw fp32 outputs of head h to out64.  (gten/ops.h:930-1000, 1046-1087)
template <int AT>
__device__ void attn_core(AttnSmem& sm, float* sc, const KVCache& kv, int g, int pos, int n_ctx, bool own_kv,
                          float* __restrict__ out64) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int own = own_kv ? pos : -1;
    // scores, scaled by 1/sqrt(64) (exactly 0.125)
    float mx = -INFINITY;
    for (int k = tid; k <= pos; k += NT) {
        const float s = __fmul_rn(score_one<AT>(sm, kv, g, k, own), 0.125f);
        sc[k] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) sm.red[wid] = mx;
    __syncthreads();
    mx = sm.red[0];
#pragma unroll
    for (int w = 1; w < NWARP; w++) mx = fmaxf(mx, sm.red[w]);
    for (int k = tid; k <= pos; k += NT) sc[k] = expf_glibc(__fsub_rn(sc[k], mx));
    __syncthreads();
    const float sum = exact_sum_block([&](int i) { return sc[i]; }, pos + 1, sm.es);
    // probabilities, re-encoded as a row of length n_ctx (masked / non-existent entries are exact zeros)
    const int nblk = (pos + 32) / 32;                      // blocks that contain at least one unmasked entry
    for (int b = wid; b < nblk; b += NWARP) {
        const int i = b * 32 + lane;
        const float p = (i <= pos && i < n_ctx) ? __fdiv_rn(sc[i], sum) : 0.0f;
        const float ph = roundtrip<AT>(p);
        __syncwarp();
        if (i <= pos) sc[i] = ph;
    }
    __syncthreads();
    // P.V: eight position-lanes (i mod 8) per channel over [0, n8), lanes summed left to right, then the tail
    const int n8 = (n_ctx / 8) * 8;
    {
        const int l = wid;                                    // NWARP == 8 lanes
        float a0 = 0.0f, a1 = 0.0f;
        const int hi = min(n8, pos + 1);
        for (int i = l; i < hi; i += 8) {
            const float p = sc[i];
            a0 = __fadd_rn(__fmul_rn(p, v_at<AT>(sm, kv, g, i, lane, own)), a0);
            a1 = __fadd_rn(__fmul_rn(p, v_at<AT>(sm, kv, g, i, lane + 32, own)), a1);
        }
        sm.part[l][lane] = a0;
        sm.part[l][lane + 32] = a1;
    }
    __syncthreads();
    if (tid < 64) {
        float d = __fadd_rn(sm.part[0][tid], sm.part[1][tid]);
#pragma unroll
        for (int l = 2; l < 8; l++) d = __fadd_rn(d, sm.part[l][tid]);
        for (int i = n8; i < n_ctx && i <= pos; i++) d = __fadd_rn(d, __fmul_rn(sc[i], v_at<AT>(sm, kv, g, i, tid, own)));
        out64[tid] = d;
    }
}

}  // namespace gtb
