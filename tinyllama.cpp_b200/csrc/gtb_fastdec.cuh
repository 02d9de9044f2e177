// gtb_fastdec.cuh -- "fast" decode kernels: the per-row graph of gtb_engine.cu with the ORDER constraint dropped.
//
// The order-exact decode path (gtb_mega.cuh) is bound by serial work the reference's arithmetic implies (in-order sums,
// ordered accumulation chains).  These kernels keep the reference's operations and rounding points (Q8
// re-encode after every op, the same integer block dots, ops.h:224-391) but sum in whatever order the hardware likes.
// Results are therefore within the tolerance of the batched prefill, not bit-identical, and greedy tokens
// follow the reference only until the first near-tie.  Opt-in: gtb_engine_set_option(e, "fast_decode", 1).
// Q8-activation models (Q8 / Q4 weights).
//
// One row = 5 kernels per layer, chained with programmatic dependent launch (griddepcontrol): every kernel starts while
// its predecessor still runs, pulls the first rows of its weights into registers and pushes the weights of the GEMV
// `pf_ahead` steps later towards L2 (cp.async.bulk.prefetch.L2), and only then waits for the predecessor's results.  HBM
// streaming is thereby decoupled from the dependency chain of the row.
//
//   k_fd_gemv<NORM, RAW>     x = E(h + E(down)) | embedding row; RMSNorm; E -> staged;  q|k|v rows           (148 CTAs)
//   k_fd_attn                (head, position chunk): E/RoPE/E of q,k,v, K/V append, scores, chunk softmax, E(P), P.V;
//                            the last chunk of a head to finish combines the chunks and writes E(attention) staged
//   k_fd_gemv<CODES, RAW>    o rows
//   k_fd_gemv<NORM, SILU>    h = E(x + E(o)); RMSNorm; one CTA = one 32-channel block of the MLP: its 32 gate rows and 32 up
//                            rows, then E(E(silu(E(gate))) * E(up)) -> staged codes
//   k_fd_gemv<CODES, RAW>    down rows
//   head: k_fd_gemv<NORM, ARGMAX>  final norm, lm_head rows, first maximum; the last CTA appends the token
#pragma once
#include "gtb_kernels.cuh"
#include "gtb_mega.cuh"

namespace gtb {

constexpr int FD_NT = 256;            // threads per CTA (two CTAs of consecutive kernels share an SM)
constexpr int FD_NW = FD_NT / 32;
constexpr int FD_CHUNKS = 8;          // position chunks per head (flash-decoding split)
constexpr int FD_PART = 68;           // floats per (head, chunk) partial: 64 channels, chunk max, chunk sum, 2 pad

enum { FD_NORM = 0, FD_CODES = 1 };
enum { FD_RAW = 0, FD_SILU = 1, FD_ARGMAX = 2 };

// an activation vector already encoded and staged (codes permuted like the weight words, gtb_kernels.cuh ActView) in HBM
struct FdAct {
    uint8_t* codes;      // [K] int8, block b at b*32, permuted with perm_byte
    float* ad;           // [K/32] decoded fp16 block scale
    int* n7;             // [K/32] -7 * sum of the block's codes (Q4 weights: nibble - 7)
};

struct FdArgs {
    int K, n_rows;
    const uint4* w; const uint16_t* ws;
    float* out;                                               // RAW / ARGMAX: raw fp32 outputs
    // NORM prologue: x = src0 (+ E(src1)) or the embedding row of tokens[pos]
    const float* src0; const float* src1; const uint16_t* normw; float* res_out;
    const uint8_t* emb_w; const uint16_t* emb_s; const int32_t* tokens; int emb_dt;
    FdAct in;                                                 // CODES prologue
    FdAct act_out; int n_ffn;                                 // SILU epilogue
    float* arg_val; int* arg_idx; unsigned* counter; int32_t* tok_out; DevState* st; int eos_id;   // ARGMAX epilogue
    const void* pf[2]; size_t pf_bytes[2];                    // L2 look-ahead: data and scale planes of a later GEMV
};

__device__ __forceinline__ void fd_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fd_wait_prior() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float fd_block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < FD_NW; w++) t += red[w];
    __syncthreads();
    return t;
}

// Q8 encode (quants.h:52-66) of a block held as quads: 8 adjacent lanes x 4 consecutive elements.  Returns the decoded
// fp16 scale; q[] = codes.  All 32 lanes must call.
__device__ __forceinline__ float fd_quad_encode(const float (&x)[4], int (&q)[4]) {
    float m = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    const float d = __fdiv_rn(m, 127.0f);
    const float s = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = (int)roundf(__fmul_rn(x[i], s));
    return h2f(f2h(d));
}
__device__ __forceinline__ void fd_quad_roundtrip(float (&x)[4]) {
    int q[4];
    const float d = fd_quad_encode(x, q);
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = __fmul_rn((float)q[i], d);
}

struct FdStaged {            // shared-memory view of the staged GEMV input
    uint32_t* aw;            // [nb][8]
    float* ad;               // [nb]
    int* n7;                 // [nb]
};

// encode a quad of block b (elements 4j .. 4j+3, j = lane & 7) into the staged vector
__device__ __forceinline__ void fd_stage_quad(const FdStaged& s, int b, int j, const float (&x)[4], bool valid) {
    int q[4];
    const float d = fd_quad_encode(x, q);
    int sum = (q[0] + q[1]) + (q[2] + q[3]);
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    if (!valid) return;
    int8_t* dst = reinterpret_cast<int8_t*>(s.aw) + b * 32;
#pragma unroll
    for (int i = 0; i < 4; i++) dst[perm_byte(4 * j + i)] = (int8_t)q[i];
    if (j == 0) { s.ad[b] = d; s.n7[b] = -7 * sum; }
}

// integer dot of one 32-block (ops.h:282-287, 339-378 without the lane split: integer sums are exact in any order)
template <int WT>
__device__ __forceinline__ int fd_block_isum(const uint4& wa, const uint4& wb, const uint4& ax, const uint4& ay, int n7) {
    int s;
    if (WT == DT_Q4) {
        s = n7;
        s = __dp4a((int)((wa.x >> 4) & 0x0f0f0f0fu), (int)ax.x, s); s = __dp4a((int)(wa.x & 0x0f0f0f0fu), (int)ay.x, s);
        s = __dp4a((int)((wa.y >> 4) & 0x0f0f0f0fu), (int)ax.y, s); s = __dp4a((int)(wa.y & 0x0f0f0f0fu), (int)ay.y, s);
        s = __dp4a((int)((wa.z >> 4) & 0x0f0f0f0fu), (int)ax.z, s); s = __dp4a((int)(wa.z & 0x0f0f0f0fu), (int)ay.z, s);
        s = __dp4a((int)((wa.w >> 4) & 0x0f0f0f0fu), (int)ax.w, s); s = __dp4a((int)(wa.w & 0x0f0f0f0fu), (int)ay.w, s);
    } else {
        s = __dp4a((int)wa.x, (int)ax.x, 0); s = __dp4a((int)wb.x, (int)ay.x, s);
        s = __dp4a((int)wa.y, (int)ax.y, s); s = __dp4a((int)wb.y, (int)ay.y, s);
        s = __dp4a((int)wa.z, (int)ax.z, s); s = __dp4a((int)wb.z, (int)ay.z, s);
        s = __dp4a((int)wa.w, (int)ax.w, s); s = __dp4a((int)wb.w, (int)ay.w, s);
    }
    return s;
}

// weights of R rows x NBL blocks per lane (lane l holds blocks l, l+32, ...) in registers
template <int WT, int NBL, int R>
struct FdBatch {
    uint4 a[R][NBL];
    uint4 b[(WT == DT_Q8) ? R : 1][(WT == DT_Q8) ? NBL : 1];
    uint32_t sc[R][NBL];
};

template <int WT, int NBL, int R>
__device__ __forceinline__ void fd_load_batch(FdBatch<WT, NBL, R>& bt, const uint4* __restrict__ wd, const uint16_t* __restrict__ wsc,
                                              const int (&row)[R], int nb) {
    const int lane = threadIdx.x & 31;
    constexpr int WPB = (WT == DT_Q4) ? 1 : 2;
#pragma unroll
    for (int u = 0; u < R; u++) {
#pragma unroll
        for (int i = 0; i < NBL; i++) {
            const int b = lane + 32 * i;
            if (row[u] >= 0 && b < nb) {
                const size_t blk = (size_t)row[u] * nb + b;
                bt.a[u][i] = ldg_stream(wd + blk * WPB);
                if (WT == DT_Q8) bt.b[u][i] = ldg_stream(wd + blk * WPB + 1);
                bt.sc[u][i] = ldg_stream_u16(wsc + blk);
            } else {
                bt.a[u][i] = make_uint4(0, 0, 0, 0);
                if (WT == DT_Q8) bt.b[u][i] = make_uint4(0, 0, 0, 0);
                bt.sc[u][i] = 0;
            }
        }
    }
}

// dot products of the batch against the staged vector; every lane returns all R row sums
template <int WT, int NBL, int R>
__device__ __forceinline__ void fd_dot_batch(const FdBatch<WT, NBL, R>& bt, const FdStaged& s, int nb, float (&acc)[R]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < R; u++) acc[u] = 0.0f;
#pragma unroll
    for (int i = 0; i < NBL; i++) {
        const int b = lane + 32 * i;
        if (b < nb) {
            const uint4 ax = reinterpret_cast<const uint4*>(s.aw)[2 * b], ay = reinterpret_cast<const uint4*>(s.aw)[2 * b + 1];
            const int n7 = (WT == DT_Q4) ? s.n7[b] : 0;
            const float ad = s.ad[b];
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int is = fd_block_isum<WT>(bt.a[u][i], bt.b[(WT == DT_Q8) ? u : 0][(WT == DT_Q8) ? i : 0], ax, ay, n7);
                acc[u] = fmaf((float)is, __fmul_rn(ad, h2f((uint16_t)bt.sc[u][i])), acc[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < R; u++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
}

static __host__ __device__ inline size_t fd_gemv_smem(int K, bool norm) {
    const int nb = K / 32;
    size_t s = (size_t)nb * 32 + (size_t)nb * 8;        // codes, ad, n7
    s = (s + 15) & ~(size_t)15;
    if (norm) s += (size_t)K * 4;
    return s + 64 * 4 + 64;                             // reduction scratch / gate|up block
}

// NBL = blocks per lane (K <= 32 * 32 * NBL), R = rows per warp pass
template <int WT, int PRO, int EPI, int NBL, int R>
__global__ void __launch_bounds__(FD_NT, 2) k_fd_gemv(FdArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int K = a.K, nb = K / 32;
    FdStaged sv;
    sv.aw = reinterpret_cast<uint32_t*>(smem);
    sv.ad = reinterpret_cast<float*>(smem + (size_t)nb * 32);
    sv.n7 = reinterpret_cast<int*>(smem + (size_t)nb * 36);
    unsigned char* p = smem + ((((size_t)nb * 40) + 15) & ~(size_t)15);
    float* xbuf = reinterpret_cast<float*>(p);
    float* scratch = reinterpret_cast<float*>(p + ((PRO == FD_NORM) ? (size_t)K * 4 : 0));      // 64 floats
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    fd_launch_dependents();
    // ---- work that does not depend on the previous kernel: look-ahead L2 prefetch, first weight batch into registers
    if (tid < 2 && a.pf[tid]) {
        const size_t per = ((a.pf_bytes[tid] + gridDim.x - 1) / gridDim.x + 127) & ~(size_t)127;
        const size_t off = per * blockIdx.x;
        if (off < a.pf_bytes[tid]) l2_prefetch(reinterpret_cast<const unsigned char*>(a.pf[tid]) + off, min(per, a.pf_bytes[tid] - off));
    }
    // virtual rows of this CTA: RAW/ARGMAX: an even slice of the matrix; SILU: unit = 32 gate rows + 32 up rows
    int v0, v1, unit = blockIdx.x;
    if (EPI == FD_SILU) { v0 = 0; v1 = 64; }
    else { v0 = (int)(((long long)blockIdx.x * a.n_rows) / gridDim.x); v1 = (int)(((long long)(blockIdx.x + 1) * a.n_rows) / gridDim.x); }
    auto rows_of = [&](int v, int (&row)[R]) {
#pragma unroll
        for (int u = 0; u < R; u++) {
            const int vv = v + u;
            if (vv >= v1) row[u] = -1;
            else if (EPI == FD_SILU) row[u] = (vv < 32) ? unit * 32 + vv : a.n_ffn + unit * 32 + (vv - 32);
            else row[u] = vv;
        }
    };
    FdBatch<WT, NBL, R> bt;
    int row[R];
    int v = v0 + wid * R;
    rows_of(v, row);
    fd_load_batch<WT, NBL, R>(bt, a.w, a.ws, row, nb);

    fd_wait_prior();
    // ---- prologue: the staged input vector
    if (PRO == FD_NORM) {
        // x = E(res + E(delta)) (ops.h:870-898) or the embedding row (ops.h:514-564); y = E(x / (rms(x) + 1e-6) * w) (ops.h:762-804)
        const int nq = K / 4;
        float ssq = 0.0f;
        const size_t erow = a.emb_w ? (size_t)__ldcg(a.tokens + __ldcg(&a.st->pos)) : 0;
        for (int base = 0; base < nq; base += FD_NT) {
            const int qd = base + tid;
            const bool valid = qd < nq;
            const int e0 = 4 * (valid ? qd : 0);
            float x[4];
            if (a.emb_w) {
                const size_t blk = erow * nb + (e0 >> 5);
                const float delta = h2f(a.emb_s[blk]);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int el = (e0 & 31) + i;
                    if (a.emb_dt == DT_Q8) {
                        x[i] = __fmul_rn((float)(int8_t)a.emb_w[blk * 32 + perm_byte(el)], delta);
                    } else {
                        const int j = el & 15;
                        const uint8_t byte = a.emb_w[blk * 16 + ((j & 7) >> 1) * 4 + (j & 1) + 2 * (j >> 3)];
                        x[i] = __fmul_rn((float)((int)((el < 16) ? (byte >> 4) : (byte & 0x0f)) - 7), delta);
                    }
                }
                if (a.emb_dt != DT_Q8) fd_quad_roundtrip(x);          // Q4 row: dequantise, re-encode as Q8 (ops.h:522-528)
            } else {
                const float4 s0 = __ldcg(reinterpret_cast<const float4*>(a.src0 + e0));
                x[0] = s0.x; x[1] = s0.y; x[2] = s0.z; x[3] = s0.w;
                if (a.src1) {
                    const float4 s1 = __ldcg(reinterpret_cast<const float4*>(a.src1 + e0));
                    float t[4] = {s1.x, s1.y, s1.z, s1.w};
                    fd_quad_roundtrip(t);
#pragma unroll
                    for (int i = 0; i < 4; i++) x[i] = __fadd_rn(x[i], t[i]);
                    fd_quad_roundtrip(x);
                }
            }
            if (valid) {
                *reinterpret_cast<float4*>(xbuf + e0) = make_float4(x[0], x[1], x[2], x[3]);
                if (blockIdx.x == 0 && a.res_out) *reinterpret_cast<float4*>(a.res_out + e0) = make_float4(x[0], x[1], x[2], x[3]);
                ssq += (x[0] * x[0] + x[1] * x[1]) + (x[2] * x[2] + x[3] * x[3]);
            }
        }
        ssq = fd_block_sum(ssq, scratch);
        const float denom = sqrtf(ssq / (float)K) + 1e-6f;
        for (int base = 0; base < nq; base += FD_NT) {
            const int qd = base + tid;
            const bool valid = qd < nq;
            const int e0 = 4 * (valid ? qd : 0);
            const float4 xv = *reinterpret_cast<const float4*>(xbuf + e0);
            const uint2 wv = *reinterpret_cast<const uint2*>(a.normw + e0);
            float y[4];
            y[0] = __fmul_rn(__fdiv_rn(xv.x, denom), h2f((uint16_t)(wv.x & 0xffffu)));
            y[1] = __fmul_rn(__fdiv_rn(xv.y, denom), h2f((uint16_t)(wv.x >> 16)));
            y[2] = __fmul_rn(__fdiv_rn(xv.z, denom), h2f((uint16_t)(wv.y & 0xffffu)));
            y[3] = __fmul_rn(__fdiv_rn(xv.w, denom), h2f((uint16_t)(wv.y >> 16)));
            fd_stage_quad(sv, e0 >> 5, (e0 & 31) >> 2, y, valid);
        }
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(a.in.codes);
        for (int i = tid; i < K / 16; i += FD_NT) reinterpret_cast<uint4*>(sv.aw)[i] = __ldcg(src + i);
        for (int i = tid; i < nb; i += FD_NT) { sv.ad[i] = __ldcg(a.in.ad + i); sv.n7[i] = __ldcg(a.in.n7 + i); }
    }
    __syncthreads();

    // ---- rows
    float best = -INFINITY;
    int arg = 0x7fffffff;
    const int n_units = (EPI == FD_SILU) ? a.n_ffn / 32 : 1;
    for (;;) {
        for (; v < v1; ) {
            float acc[R];
            fd_dot_batch<WT, NBL, R>(bt, sv, nb, acc);
            const int vcur = v;
            int rcur[R];
#pragma unroll
            for (int u = 0; u < R; u++) rcur[u] = row[u];
            v += FD_NW * R;
            if (v < v1) { rows_of(v, row); fd_load_batch<WT, NBL, R>(bt, a.w, a.ws, row, nb); }
            if (EPI == FD_SILU) {
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < R; u++) if (rcur[u] >= 0) scratch[vcur + u] = acc[u];
                }
            } else {
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < R; u++) {
                        if (rcur[u] >= 0) {
                            a.out[rcur[u]] = acc[u];
                            if (EPI == FD_ARGMAX && acc[u] > best) { best = acc[u]; arg = rcur[u]; }
                        }
                    }
                }
            }
        }
        if (EPI != FD_SILU) break;
        // E(E(silu(E(gate))) * E(up)) (modules.cpp:238-247) of this unit's 32 channels -> staged codes for the down GEMV
        __syncthreads();
        const int next_unit = unit + gridDim.x;
        if (next_unit < n_units) {          // (grid smaller than the number of units) next unit's first batch while warp 0 encodes
            v = wid * R;
            const int keep = unit; unit = next_unit; rows_of(v, row); unit = keep;
            fd_load_batch<WT, NBL, R>(bt, a.w, a.ws, row, nb);
        }
        if (wid == 0) {
            const float g1 = q8_roundtrip_lane(scratch[lane]);
            const float u1 = q8_roundtrip_lane(scratch[32 + lane]);
            const float g2 = q8_roundtrip_lane(__fdividef(g1, 1.0f + __expf(-g1)));
            uint16_t dh;
            const int q = q8_encode_lane(__fmul_rn(g2, u1), &dh);
            int s = q;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            a.act_out.codes[unit * 32 + perm_byte(lane)] = (uint8_t)(int8_t)q;
            if (lane == 0) { a.act_out.ad[unit] = h2f(dh); a.act_out.n7[unit] = -7 * s; }
        }
        if (next_unit >= n_units) break;
        unit = next_unit;
        __syncthreads();
    }

    if (EPI == FD_ARGMAX) {
        // first maximum of this CTA's rows (tinyllama.cpp:416-424), then the last CTA to arrive reduces all of them
        float* sval = scratch;
        int* sidx = reinterpret_cast<int*>(scratch + 16);
        __shared__ bool is_last;
        if (lane == 0) { sval[wid] = best; sidx[wid] = arg; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < FD_NW; w++) if (sval[w] > best || (sval[w] == best && sidx[w] < arg)) { best = sval[w]; arg = sidx[w]; }
            a.arg_val[blockIdx.x] = best; a.arg_idx[blockIdx.x] = arg;
            __threadfence();
            const unsigned old = atomicAdd(a.counter, 1u);
            is_last = (old == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last && wid == 0) {
            __threadfence();
            best = -INFINITY; arg = 0x7fffffff;
            for (int c = lane; c < (int)gridDim.x; c += 32) {
                const float ov = __ldcg(a.arg_val + c);
                const int oi = __ldcg(a.arg_idx + c);
                if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
            }
            if (lane == 0) {
                if (arg == 0x7fffffff) arg = 0;
                const int pos = a.st->pos;
                a.tok_out[pos + 1] = arg;
                a.st->pos = pos + 1;
                a.st->n_gen += 1;
                if (arg == a.eos_id) a.st->stop = 1;
                *a.counter = 0u;
            }
        }
    }
}

// ---------------------------------------------------------------- attention: one CTA = (head, position chunk)
struct FdAttnArgs {
    const float* rqkv; int n_embd, kv_dim, gsz;
    uint8_t* kq; uint16_t* ks; uint8_t* vq; uint16_t* vs;
    const float* rope_cos; const float* rope_sin; const DevState* st;
    float* parts;             // [n_heads][FD_CHUNKS][FD_PART]
    unsigned* counters;       // [n_heads]
    FdAct out;                // E(attention output) staged for the o GEMV
    const void* pf[2]; size_t pf_bytes[2];
};

struct FdAttnSmem {
    uint32_t qw[16]; float qd[2];
    uint32_t kw[16]; float kd[2];
    float vf[64];
    float tmp[6][32];
    float red[FD_NW];
    float part[FD_NW][64];
    int last;
};

__global__ void __launch_bounds__(FD_NT, 2) k_fd_attn(FdAttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    FdAttnSmem& sm = *reinterpret_cast<FdAttnSmem*>(smem);
    float* sc = reinterpret_cast<float*>(smem + ((sizeof(FdAttnSmem) + 15) & ~(size_t)15));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int h = blockIdx.x / FD_CHUNKS, c = blockIdx.x % FD_CHUNKS, g = h / a.gsz;
    fd_launch_dependents();
    if (tid < 2 && a.pf[tid]) {
        const size_t per = ((a.pf_bytes[tid] + gridDim.x - 1) / gridDim.x + 127) & ~(size_t)127;
        const size_t off = per * blockIdx.x;
        if (off < a.pf_bytes[tid]) l2_prefetch(reinterpret_cast<const unsigned char*>(a.pf[tid]) + off, min(per, a.pf_bytes[tid] - off));
    }
    fd_wait_prior();
    const int pos = __ldcg(&a.st->pos);
    const bool writer = (h % a.gsz) == 0 && c == 0;
    // chunk of positions [lo, hi), aligned to the 32-position blocks of the probability row
    const int per = ((pos + 1 + FD_CHUNKS - 1) / FD_CHUNKS + 31) & ~31;
    const int lo = c * per, hi = min(pos + 1, lo + per);
    const int n = hi - lo;
    const int kvb = a.kv_dim / 32;
    // q, k, v of this row: Linear re-encode, RoPE, re-encode (ops.h:645-646, 733-753); K/V append by one CTA per group
    if (wid < 6) {
        const int which = wid >> 1, half = wid & 1;
        const float* src = (which == 0) ? a.rqkv + h * 64 : (which == 1) ? a.rqkv + a.n_embd + g * 64 : a.rqkv + a.n_embd + a.kv_dim + g * 64;
        const float x = __ldcg(src + half * 32 + lane);
        if (which < 2) {
            sm.tmp[wid][lane] = q8_roundtrip_lane(x);
        } else {
            uint16_t dh;
            const int q = q8_encode_lane(x, &dh);
            sm.vf[half * 32 + lane] = __fmul_rn((float)q, h2f(dh));
            if (writer) {
                a.vq[(size_t)pos * a.kv_dim + g * 64 + half * 32 + lane] = (uint8_t)(int8_t)q;
                if (lane == 0) a.vs[(size_t)pos * kvb + g * 2 + half] = dh;
            }
        }
    }
    __syncthreads();
    if (wid < 4) {
        const int which = wid >> 1, half = wid & 1;
        const float x0 = sm.tmp[which * 2][lane], x1 = sm.tmp[which * 2 + 1][lane];
        const float cs = a.rope_cos[(size_t)pos * 32 + lane], sn = a.rope_sin[(size_t)pos * 32 + lane];
        const float o = (half == 0) ? __fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn)) : __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs));
        uint16_t dh;
        const int q = q8_encode_lane(o, &dh);
        const int pb = perm_byte(lane);
        if (which == 0) {
            reinterpret_cast<int8_t*>(sm.qw)[half * 32 + pb] = (int8_t)q;
            if (lane == 0) sm.qd[half] = h2f(dh);
        } else {
            reinterpret_cast<int8_t*>(sm.kw)[half * 32 + pb] = (int8_t)q;
            if (lane == 0) sm.kd[half] = h2f(dh);
            if (writer) {
                a.kq[(size_t)pos * a.kv_dim + g * 64 + half * 32 + pb] = (uint8_t)(int8_t)q;
                if (lane == 0) a.ks[(size_t)pos * kvb + g * 2 + half] = dh;
            }
        }
    }
    __syncthreads();
    float* part = a.parts + ((size_t)h * FD_CHUNKS + c) * FD_PART;
    float mx = -INFINITY, lsum = 0.0f;
    if (n > 0) {
        // scores of the chunk, scaled by 1/sqrt(64); chunk maximum
        uint32_t qx[16];
#pragma unroll
        for (int i = 0; i < 16; i++) qx[i] = sm.qw[i];
        for (int k = lo + tid; k < hi; k += FD_NT) {
            float s = 0.0f;
#pragma unroll
            for (int bi = 0; bi < 2; bi++) {
                uint4 kx, ky;
                float kdv;
                if (k == pos) {
                    kx = make_uint4(sm.kw[bi * 8 + 0], sm.kw[bi * 8 + 1], sm.kw[bi * 8 + 2], sm.kw[bi * 8 + 3]);
                    ky = make_uint4(sm.kw[bi * 8 + 4], sm.kw[bi * 8 + 5], sm.kw[bi * 8 + 6], sm.kw[bi * 8 + 7]);
                    kdv = sm.kd[bi];
                } else {
                    const uint4* kp = reinterpret_cast<const uint4*>(a.kq + (size_t)k * a.kv_dim + g * 64 + bi * 32);
                    kx = __ldcg(kp); ky = __ldcg(kp + 1);
                    kdv = h2f(__ldcg(a.ks + (size_t)k * kvb + g * 2 + bi));
                }
                int is = __dp4a((int)kx.x, (int)qx[bi * 8 + 0], 0);
                is = __dp4a((int)kx.y, (int)qx[bi * 8 + 1], is); is = __dp4a((int)kx.z, (int)qx[bi * 8 + 2], is); is = __dp4a((int)kx.w, (int)qx[bi * 8 + 3], is);
                is = __dp4a((int)ky.x, (int)qx[bi * 8 + 4], is); is = __dp4a((int)ky.y, (int)qx[bi * 8 + 5], is);
                is = __dp4a((int)ky.z, (int)qx[bi * 8 + 6], is); is = __dp4a((int)ky.w, (int)qx[bi * 8 + 7], is);
                s = fmaf((float)is, __fmul_rn(sm.qd[bi], kdv), s);
            }
            s *= 0.125f;
            sc[k - lo] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        if (lane == 0) sm.red[wid] = mx;
        __syncthreads();
        mx = sm.red[0];
#pragma unroll
        for (int w = 1; w < FD_NW; w++) mx = fmaxf(mx, sm.red[w]);
        __syncthreads();
        // e = exp(s - chunk max); the Q8 re-encode of the probability row per 32 positions (ops.h:996): the codes depend only
        // on e / max(e of the block); the block scale stays relative to the chunk maximum and is normalised at the combine
        const int nblk = (n + 31) / 32;
        for (int b = wid; b < nblk; b += FD_NW) {
            const int i = b * 32 + lane;
            const float e = (i < n) ? __expf(sc[i] - mx) : 0.0f;
            lsum += e;
            const float emax = warp_max(e);
            const float code = (emax > 0.0f) ? floorf(e * __fdividef(127.0f, emax) + 0.5f) : 0.0f;
            sc[i] = code * (emax * (1.0f / 127.0f));
        }
        lsum = fd_block_sum(lsum, sm.red);       // (also orders the sc[] writes before the reads below)
        // P.V: thread = (position slot pp of 64, 16 channels cq); V rows as 128-bit loads
        const int pp = tid >> 2, cq = tid & 3;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; j++) acc[j] = 0.0f;
        for (int i0 = pp; i0 < n; i0 += 4 * 64) {
            uint4 vv[4];
            float wgt[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * 64;
                vv[u] = make_uint4(0, 0, 0, 0); wgt[u] = 0.0f;
                if (i < n && lo + i != pos) {
                    vv[u] = __ldcg(reinterpret_cast<const uint4*>(a.vq + (size_t)(lo + i) * a.kv_dim + g * 64 + cq * 16));
                    wgt[u] = sc[i] * h2f(__ldcg(a.vs + (size_t)(lo + i) * kvb + g * 2 + (cq >> 1)));      // ops.h:1026
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t wds[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] = fmaf((float)(int)(int8_t)(wds[j >> 2] >> (8 * (j & 3))), wgt[u], acc[j]);
            }
        }
        if (pos >= lo && pos < hi && ((pos - lo) & 63) == pp) {      // this row's own v is not read back from the cache
            const float p = sc[pos - lo];
#pragma unroll
            for (int j = 0; j < 16; j++) acc[j] = fmaf(sm.vf[cq * 16 + j], p, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; j++) {
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
        }
        if (lane < 4) {
#pragma unroll
            for (int j = 0; j < 16; j++) sm.part[wid][lane * 16 + j] = acc[j];
        }
        __syncthreads();
    }
    if (tid < 64) {
        float o = 0.0f;
        if (n > 0) {
#pragma unroll
            for (int w = 0; w < FD_NW; w++) o += sm.part[w][tid];
        }
        part[tid] = o;
    }
    if (tid == 64) part[64] = mx;
    if (tid == 65) part[65] = lsum;
    // ---- the last chunk of this head to finish combines: o = sum_c o_c * exp(m_c - m) / sum_c l_c * exp(m_c - m), E(o)
    __threadfence();
    __syncthreads();
    if (tid == 0) sm.last = (atomicAdd(a.counters + h, 1u) == FD_CHUNKS - 1);
    __syncthreads();
    if (!sm.last) return;
    __threadfence();
    if (wid < 2) {
        const float* p = a.parts + (size_t)h * FD_CHUNKS * FD_PART;
        float m = -INFINITY;
#pragma unroll
        for (int cc = 0; cc < FD_CHUNKS; cc++) m = fmaxf(m, __ldcg(p + cc * FD_PART + 64));
        float l = 0.0f, o = 0.0f;
#pragma unroll
        for (int cc = 0; cc < FD_CHUNKS; cc++) {
            const float mc = __ldcg(p + cc * FD_PART + 64);
            const float f = (mc == -INFINITY) ? 0.0f : __expf(mc - m);
            l = fmaf(__ldcg(p + cc * FD_PART + 65), f, l);
            o = fmaf(__ldcg(p + cc * FD_PART + tid), f, o);
        }
        uint16_t dh;
        const int q = q8_encode_lane(__fdividef(o, l), &dh);          // E(attention output), ops.h:1084
        int s = q;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) s += __shfl_xor_sync(0xffffffffu, s, ofs);
        const int b = 2 * h + wid;
        a.out.codes[b * 32 + perm_byte(lane)] = (uint8_t)(int8_t)q;
        if (lane == 0) { a.out.ad[b] = h2f(dh); a.out.n7[b] = -7 * s; }
    }
    if (tid == 0) a.counters[h] = 0u;
}

}  // namespace gtb
