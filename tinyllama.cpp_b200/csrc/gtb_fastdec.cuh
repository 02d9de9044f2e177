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
//   k_fd_attn                (KV group, position chunk), warp = query head: K/V rows of the chunk once in shared memory for the 8
//                            heads, E/RoPE/E of q (and k, v of the row), scores, block softmax, E(P), P.V; the last chunk of a
//                            group to finish combines the chunks and writes E(attention) staged
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
constexpr int FD_MAXCH = 32;          // position chunks per KV group at most (flash-decoding split)
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
    int n_seq, tok_stride;                                    // batched decode (k_fdb_*): sequences, tokens row stride
};

// debug timeline (option "fd_trace"): CTA 0 of every kernel of the chains stamps %globaltimer at entry, after the dependency wait
// and at its end; entry i = (tag << 48) | ns, tag = kernel kind * 4 + stage
// Compiled in with -DGTB_FD_TRACE only (python tinyllama.cpp_b200/build.py --trace): the pointer load alone costs ~0.5 us per kernel.
__device__ long long* g_fd_trace = nullptr;
#ifdef GTB_FD_TRACE
__device__ __forceinline__ void fd_trace(int tag) {
    if (g_fd_trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        const unsigned slot = atomicAdd(reinterpret_cast<unsigned*>(g_fd_trace), 1u);
        if (slot < 4000) g_fd_trace[1 + slot] = ((long long)tag << 48) | (gtimer() & 0xffffffffffffll);
    }
}
#else
__device__ __forceinline__ void fd_trace(int) {}
#endif

__device__ __forceinline__ void fd_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fd_wait_prior() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// the warp index through a shuffle: the compiler then knows it is warp-uniform, so branches and loops on it keep the warp
// converged and the shuffles inside them need no WARPSYNC / reconvergence wrappers
__device__ __forceinline__ int fd_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ float fd_block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = fd_warp_id();
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

// roundf (half away from zero) without the slow path: exact except for |v| within one ulp below a .5 boundary
__device__ __forceinline__ int fd_round_away(float v) { return (int)(v + copysignf(0.5f, v)); }

// Q8 encode (quants.h:52-66; fast division: the order-free path does not promise the last bit) of a block held as quads: 8 adjacent lanes x 4 consecutive elements.  Returns the decoded
// fp16 scale; q[] = codes.  All 32 lanes must call.
__device__ __forceinline__ float fd_quad_encode(const float (&x)[4], int (&q)[4]) {
    float m = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    const float s = (m > 0.0f) ? __fdividef(127.0f, m) : 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = fd_round_away(x[i] * s);
    return h2f(f2h(m * (1.0f / 127.0f)));
}
// the same for a block held one element per lane; returns the code, *d = decoded fp16 scale
__device__ __forceinline__ int fd_lane_encode(float x, float* d) {
    const float m = warp_max(fabsf(x));
    const float s = (m > 0.0f) ? __fdividef(127.0f, m) : 0.0f;
    *d = h2f(f2h(m * (1.0f / 127.0f)));
    return fd_round_away(x * s);
}
__device__ __forceinline__ float fd_lane_roundtrip(float x) {
    float d;
    const int q = fd_lane_encode(x, &d);
    return (float)q * d;
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
    (void)norm;
    return s + 64 * 4 + 64;                             // reduction scratch / gate|up block
}

// NORM prologue: x = E(res + E(delta)) or the embedding row, RMSNorm, E -> staged vector `sv` (shared memory).  Whole CTA.
template <int NQ>
__device__ __forceinline__ void fd_norm_stage(const FdArgs& a, const FdStaged& sv, float* scratch, const uint2 (&nw)[NQ], int pos, bool write_res) {
    const int K = a.K, nb = K / 32, tid = threadIdx.x;
    // x = E(res + E(delta)) (ops.h:870-898) or the embedding row (ops.h:514-564); y = E(x / (rms(x) + 1e-6) * w) (ops.h:762-804)
    // NQ quads per thread, all loads of the pass in flight at once, x stays in registers
    const int nq = K / 4;
    float x[NQ][4];
    float ssq = 0.0f;
    if (a.emb_w) {
        const size_t erow = (size_t)__ldcg(a.tokens + pos);
#pragma unroll
        for (int it = 0; it < NQ; it++) {
            const int qd = it * FD_NT + tid;
            const int e0 = 4 * ((qd < nq) ? qd : 0);
            const size_t blk = erow * nb + (e0 >> 5);
            const float delta = h2f(a.emb_s[blk]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int el = (e0 & 31) + i;
                if (a.emb_dt == DT_Q8) {
                    x[it][i] = __fmul_rn((float)(int8_t)a.emb_w[blk * 32 + perm_byte(el)], delta);
                } else {
                    const int j = el & 15;
                    const uint8_t byte = a.emb_w[blk * 16 + ((j & 7) >> 1) * 4 + (j & 1) + 2 * (j >> 3)];
                    x[it][i] = __fmul_rn((float)((int)((el < 16) ? (byte >> 4) : (byte & 0x0f)) - 7), delta);
                }
            }
            if (a.emb_dt != DT_Q8) fd_quad_roundtrip(x[it]);      // Q4 row: dequantise, re-encode as Q8 (ops.h:522-528)
        }
    } else {
        float4 s0[NQ], s1[NQ];
#pragma unroll
        for (int it = 0; it < NQ; it++) {
            const int qd = it * FD_NT + tid;
            const int e0 = 4 * ((qd < nq) ? qd : 0);
            s0[it] = __ldcg(reinterpret_cast<const float4*>(a.src0 + e0));
            if (a.src1) s1[it] = __ldcg(reinterpret_cast<const float4*>(a.src1 + e0));
        }
#pragma unroll
        for (int it = 0; it < NQ; it++) {
            x[it][0] = s0[it].x; x[it][1] = s0[it].y; x[it][2] = s0[it].z; x[it][3] = s0[it].w;
            if (a.src1) {
                float t[4] = {s1[it].x, s1[it].y, s1[it].z, s1[it].w};
                fd_quad_roundtrip(t);
#pragma unroll
                for (int i = 0; i < 4; i++) x[it][i] = __fadd_rn(x[it][i], t[i]);
                fd_quad_roundtrip(x[it]);
            }
        }
    }
#pragma unroll
    for (int it = 0; it < NQ; it++) {
        const int qd = it * FD_NT + tid;
        if (qd < nq) {
            if (write_res && a.res_out) *reinterpret_cast<float4*>(a.res_out + 4 * qd) = make_float4(x[it][0], x[it][1], x[it][2], x[it][3]);
            ssq += (x[it][0] * x[it][0] + x[it][1] * x[it][1]) + (x[it][2] * x[it][2] + x[it][3] * x[it][3]);
        }
    }
    ssq = fd_block_sum(ssq, scratch);
    const float rden = 1.0f / (sqrtf(ssq / (float)K) + 1e-6f);
#pragma unroll
    for (int it = 0; it < NQ; it++) {
        const int qd = it * FD_NT + tid;
        const bool valid = qd < nq;
        const int e0 = 4 * (valid ? qd : 0);
        float y[4];
        y[0] = __fmul_rn(x[it][0] * rden, h2f((uint16_t)(nw[it].x & 0xffffu)));
        y[1] = __fmul_rn(x[it][1] * rden, h2f((uint16_t)(nw[it].x >> 16)));
        y[2] = __fmul_rn(x[it][2] * rden, h2f((uint16_t)(nw[it].y & 0xffffu)));
        y[3] = __fmul_rn(x[it][3] * rden, h2f((uint16_t)(nw[it].y >> 16)));
        fd_stage_quad(sv, e0 >> 5, (e0 & 31) >> 2, y, valid);
    }
}
// NBL = blocks per lane (K <= 32 * 32 * NBL), R = rows per warp pass.
// `sync()` orders this phase after its producer: griddepcontrol.wait in the one-kernel-per-phase chain, a grid barrier in the
// persistent kernel.  Everything before it touches only weights.
template <int WT, int PRO, int EPI, int NBL, int R, typename Sync>
__device__ __forceinline__ void fd_gemv_phase(const FdArgs& a, unsigned char* smem, Sync& sync, int pos_in) {
    const int K = a.K, nb = K / 32;
    FdStaged sv;
    sv.aw = reinterpret_cast<uint32_t*>(smem);
    sv.ad = reinterpret_cast<float*>(smem + (size_t)nb * 32);
    sv.n7 = reinterpret_cast<int*>(smem + (size_t)nb * 36);
    float* scratch = reinterpret_cast<float*>(smem + ((((size_t)nb * 40) + 15) & ~(size_t)15));      // 64 floats
    const int tid = threadIdx.x, lane = tid & 31, wid = fd_warp_id();

    sync.arrive();
    sync.stamp();
    // ---- work that does not depend on the previous kernel: look-ahead L2 prefetch, first weight batch into registers
    if (tid < 2 && a.pf[tid]) {
        const size_t per = ((a.pf_bytes[tid] + gridDim.x - 1) / gridDim.x + 127) & ~(size_t)127;
        const size_t off = per * blockIdx.x;
        if (off < a.pf_bytes[tid]) l2_prefetch(reinterpret_cast<const unsigned char*>(a.pf[tid]) + off, min(per, a.pf_bytes[tid] - off));
    }
    // virtual rows of this CTA: RAW/ARGMAX: an even slice of the matrix; SILU: unit = 32 gate rows + 32 up rows
    int v0, v1, unit = blockIdx.x;
    const int n_units = (EPI == FD_SILU) ? a.n_ffn / 32 : 1;
    if (EPI == FD_SILU) { v0 = 0; v1 = (unit < n_units) ? 64 : 0; }          // (a grid larger than the number of units: idle CTAs)
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
    constexpr int NQ = (PRO == FD_NORM) ? NBL : 1;           // quads per thread of the NORM prologue (K <= 1024 * NBL)
    uint2 nw[NQ];                                            // norm weights are immutable: loaded before the wait
    if (PRO == FD_NORM) {
#pragma unroll
        for (int it = 0; it < NQ; it++) {
            const int qd = it * FD_NT + tid;
            nw[it] = *reinterpret_cast<const uint2*>(a.normw + 4 * ((qd < K / 4) ? qd : 0));
        }
    }

    sync.wait();
    sync.stamp();
    const int pos = (pos_in >= 0) ? pos_in : ((PRO == FD_NORM && a.emb_w) ? __ldcg(&a.st->pos) : 0);
    // ---- prologue: the staged input vector
    if (PRO == FD_NORM) {
        fd_norm_stage<NQ>(a, sv, scratch, nw, pos, blockIdx.x == 0);
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(a.in.codes);
        for (int i = tid; i < K / 16; i += FD_NT) reinterpret_cast<uint4*>(sv.aw)[i] = __ldcg(src + i);
        for (int i = tid; i < nb; i += FD_NT) { sv.ad[i] = __ldcg(a.in.ad + i); sv.n7[i] = __ldcg(a.in.n7 + i); }
    }
    __syncthreads();
    sync.stamp();

    // ---- rows
    float best = -INFINITY;
    int arg = 0x7fffffff;
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
        if (EPI != FD_SILU || unit >= n_units) break;
        // E(E(silu(E(gate))) * E(up)) (modules.cpp:238-247) of this unit's 32 channels -> staged codes for the down GEMV
        __syncthreads();
        const int next_unit = unit + gridDim.x;
        if (next_unit < n_units) {          // (grid smaller than the number of units) next unit's first batch while warp 0 encodes
            v = wid * R;
            const int keep = unit; unit = next_unit; rows_of(v, row); unit = keep;
            fd_load_batch<WT, NBL, R>(bt, a.w, a.ws, row, nb);
        }
        if (wid == 0) {
            const float g1 = fd_lane_roundtrip(scratch[lane]);
            const float u1 = fd_lane_roundtrip(scratch[32 + lane]);
            const float g2 = fd_lane_roundtrip(__fdividef(g1, 1.0f + __expf(-g1)));
            float dd;
            const int q = fd_lane_encode(g2 * u1, &dd);
            int s = q;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            a.act_out.codes[unit * 32 + perm_byte(lane)] = (uint8_t)(int8_t)q;
            if (lane == 0) { a.act_out.ad[unit] = dd; a.act_out.n7[unit] = -7 * s; }
        }
        if (next_unit >= n_units) break;
        unit = next_unit;
        __syncthreads();
    }

    sync.stamp();
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
                const int pcur = (pos_in >= 0) ? pos_in : __ldcg(&a.st->pos);
                a.tok_out[pcur + 1] = arg;
                a.st->pos = pcur + 1;
                a.st->n_gen = __ldcg(&a.st->n_gen) + 1;
                if (arg == a.eos_id) a.st->stop = 1;
                *a.counter = 0u;
            }
        }
    }
}

struct FdPdlSync {
    __device__ __forceinline__ void arrive() {}
    __device__ __forceinline__ void wait() { fd_wait_prior(); fd_trace(7 * 4 + 1); }
    __device__ __forceinline__ void stamp() {}
};

template <int WT, int PRO, int EPI, int NBL, int R>
__global__ void __launch_bounds__(FD_NT, 2) k_fd_gemv(FdArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    fd_trace((1 + EPI) * 4 + 0);
    fd_launch_dependents();
    FdPdlSync sync;
    fd_gemv_phase<WT, PRO, EPI, NBL, R>(a, smem, sync, -1);
    fd_trace((1 + EPI) * 4 + 2);
}

// ---------------------------------------------------------------- attention: one CTA = (KV group, position chunk), warp = query head
struct FdAttnArgs {
    const float* rqkv; int n_embd, kv_dim, gsz;
    uint8_t* kq; uint16_t* ks; uint8_t* vq; uint16_t* vs;
    const float* rope_cos; const float* rope_sin; const DevState* st;
    float* parts;             // [n_heads][FD_MAXCH][FD_PART]
    unsigned* counters;       // [n_groups] arrivals of a group's chunks
    FdAct out;                // E(attention output) staged for the o GEMV
    const void* pf[2]; size_t pf_bytes[2];
    size_t seq_kv_codes, seq_kv_scales;          // batched decode: per-sequence strides of the K/V cache planes
    int n_heads, n_groups;
    int chunk_len;            // positions per chunk: multiple of 32, FD_MAXCH chunks cover max_ctx (fd_chunk_len)
};

// `want` positions per chunk (64: lowest latency for one sequence; 128: fewer, fuller CTAs for a batch), raised until
// FD_MAXCH chunks cover max_ctx
static __host__ __device__ inline int fd_chunk_len(int max_ctx, int want) {
    const int cl = (((max_ctx + FD_MAXCH - 1) / FD_MAXCH) + 31) & ~31;
    return cl < want ? want : cl;
}
// shared memory of a chunk: K codes transposed [16 words][CL], K scales [2][CL], V rows [CL][64 B], V scales [CL][2],
// the staged q of the 8 heads, a flag
static __host__ __device__ inline size_t fd_attn_smem(int chunk_len) {
    return (size_t)chunk_len * 144 + FD_NW * 18 * 4 + 16;
}

// The 8 query heads of a KV group share one copy of the chunk's K/V rows in shared memory (GQA: 32 heads, 4 groups,
// tinyllama.cpp:12-20); warp w runs head g * gsz + w over the chunk's 32-position blocks (the re-encode unit of ops.h:996)
// with its own running (max, sum, 64 outputs).  The chunk that holds the row's own position also encodes and appends this
// row's k and v.  The last chunk of a group to finish (arrival counter) combines the chunks of its 8 heads.
template <typename Sync>
__device__ __forceinline__ void fd_attn_phase(const FdAttnArgs& a, unsigned char* smem, Sync& sync, int pos_in) {
    const int CL = a.chunk_len;
    uint32_t* kT = reinterpret_cast<uint32_t*>(smem);                       // [16][CL]
    float* kd = reinterpret_cast<float*>(smem + (size_t)CL * 64);           // [2][CL]
    uint4* vR = reinterpret_cast<uint4*>(smem + (size_t)CL * 72);           // [CL][4]
    float* vd = reinterpret_cast<float*>(smem + (size_t)CL * 136);          // [CL][2]
    uint32_t* qw = reinterpret_cast<uint32_t*>(smem + (size_t)CL * 144);    // [FD_NW][16]
    float* qd = reinterpret_cast<float*>(qw + FD_NW * 16);                  // [FD_NW][2]
    int* last = reinterpret_cast<int*>(qd + FD_NW * 2);
    const int tid = threadIdx.x, lane = tid & 31, wid = fd_warp_id();
    const int slot = lane >> 2, cq = lane & 3;          // P.V: lane = (position slot of 8, 16 channels)
    sync.arrive();
    sync.stamp();
    sync.wait();
    sync.stamp();
    const int pos = (pos_in >= 0) ? pos_in : __ldcg(&a.st->pos);
    const int nch = (pos + CL) / CL;                    // chunks in use follow the context
    const int kvb = a.kv_dim / 32;
    for (int unit = blockIdx.x; unit < a.n_groups * FD_MAXCH; unit += gridDim.x) {
        const int g = unit / FD_MAXCH, c = unit % FD_MAXCH;
        if (c >= nch) continue;
        const int lo = c * CL, hi = min(pos + 1, lo + CL), n = hi - lo, nblk = (n + 31) / 32;
        const bool own = (c == nch - 1);                // this chunk holds the row's own position
        const int h = g * a.gsz + wid;
        const bool head_warp = wid < a.gsz;
        // ---- the few words the prep waits for go first
        const float cs = a.rope_cos[(size_t)pos * 32 + lane], sn = a.rope_sin[(size_t)pos * 32 + lane];
        float xq0 = 0.0f, xq1 = 0.0f, xo0 = 0.0f, xo1 = 0.0f;
        if (head_warp) { xq0 = __ldcg(a.rqkv + h * 64 + lane); xq1 = __ldcg(a.rqkv + h * 64 + 32 + lane); }
        if (own && wid >= FD_NW - 2) {                  // warp 6: this row's k, warp 7: its v
            const float* src = a.rqkv + a.n_embd + (wid == FD_NW - 1 ? a.kv_dim : 0) + g * 64;
            xo0 = __ldcg(src + lane); xo1 = __ldcg(src + 32 + lane);
        }
        // ---- the chunk's K/V rows: in flight while the prep runs (first 512 uint4 of each; longer chunks loop below)
        uint4 kreg[2], vreg[2];
        uint32_t ksreg = 0, vsreg = 0;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int idx = tid + FD_NT * j, p = idx >> 2, q4 = idx & 3;
            if (p < n && lo + p != pos) {
                kreg[j] = __ldcg(reinterpret_cast<const uint4*>(a.kq + (size_t)(lo + p) * a.kv_dim + g * 64) + q4);
                vreg[j] = __ldcg(reinterpret_cast<const uint4*>(a.vq + (size_t)(lo + p) * a.kv_dim + g * 64) + q4);
            }
        }
        {
            const int p = tid >> 1, bsel = tid & 1;
            if (p < n && lo + p != pos) {
                ksreg = __ldcg(a.ks + (size_t)(lo + p) * kvb + g * 2 + bsel);
                vsreg = __ldcg(a.vs + (size_t)(lo + p) * kvb + g * 2 + bsel);
            }
        }
        // ---- q of this warp's head: Linear re-encode, RoPE, re-encode (ops.h:645-646, 733-753)
        auto rope_encode = [&](float x0, float x1, int (&q)[2], float (&d)[2]) {
            x0 = fd_lane_roundtrip(x0); x1 = fd_lane_roundtrip(x1);
            const float o0 = __fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn)), o1 = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs));
            q[0] = fd_lane_encode(o0, &d[0]); q[1] = fd_lane_encode(o1, &d[1]);
        };
        const int pb = perm_byte(lane);
        if (head_warp) {
            int q[2]; float d[2];
            rope_encode(xq0, xq1, q, d);
            reinterpret_cast<int8_t*>(qw + wid * 16)[pb] = (int8_t)q[0];
            reinterpret_cast<int8_t*>(qw + wid * 16)[32 + pb] = (int8_t)q[1];
            if (lane == 0) { qd[wid * 2] = d[0]; qd[wid * 2 + 1] = d[1]; }
        }
        if (own && wid == FD_NW - 2) {                  // k of this row: into the chunk and the cache (K rows are stored permuted)
            int q[2]; float d[2];
            rope_encode(xo0, xo1, q, d);
            const int pp = pos - lo;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int byte = half * 32 + pb;
                reinterpret_cast<int8_t*>(kT)[((size_t)(byte >> 2) * CL + pp) * 4 + (byte & 3)] = (int8_t)q[half];
                a.kq[(size_t)pos * a.kv_dim + g * 64 + byte] = (uint8_t)(int8_t)q[half];
                if (lane == 0) { kd[half * CL + pp] = d[half]; a.ks[(size_t)pos * kvb + g * 2 + half] = f2h(d[half]); }
            }
        }
        if (own && wid == FD_NW - 1) {                  // v of this row: Linear re-encode only
            const int pp = pos - lo;
            float d0, d1;
            const int q0 = fd_lane_encode(xo0, &d0), q1 = fd_lane_encode(xo1, &d1);
            reinterpret_cast<int8_t*>(vR)[(size_t)pp * 64 + lane] = (int8_t)q0;
            reinterpret_cast<int8_t*>(vR)[(size_t)pp * 64 + 32 + lane] = (int8_t)q1;
            a.vq[(size_t)pos * a.kv_dim + g * 64 + lane] = (uint8_t)(int8_t)q0;
            a.vq[(size_t)pos * a.kv_dim + g * 64 + 32 + lane] = (uint8_t)(int8_t)q1;
            if (lane == 0) {
                vd[pp * 2] = d0; vd[pp * 2 + 1] = d1;
                a.vs[(size_t)pos * kvb + g * 2] = f2h(d0); a.vs[(size_t)pos * kvb + g * 2 + 1] = f2h(d1);
            }
        }
        // ---- K/V rows into shared memory
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int idx = tid + FD_NT * j, p = idx >> 2, q4 = idx & 3;
            if (p < n && lo + p != pos) {
                kT[(q4 * 4 + 0) * CL + p] = kreg[j].x; kT[(q4 * 4 + 1) * CL + p] = kreg[j].y;
                kT[(q4 * 4 + 2) * CL + p] = kreg[j].z; kT[(q4 * 4 + 3) * CL + p] = kreg[j].w;
                vR[p * 4 + q4] = vreg[j];
            }
        }
        {
            const int p = tid >> 1, bsel = tid & 1;
            if (p < n && lo + p != pos) { kd[bsel * CL + p] = h2f((uint16_t)ksreg); vd[p * 2 + bsel] = h2f((uint16_t)vsreg); }
        }
        for (int idx = tid + 2 * FD_NT; idx < 4 * n; idx += FD_NT) {          // chunks longer than 128 positions
            const int p = idx >> 2, q4 = idx & 3;
            if (lo + p != pos) {
                const uint4 kk = __ldcg(reinterpret_cast<const uint4*>(a.kq + (size_t)(lo + p) * a.kv_dim + g * 64) + q4);
                kT[(q4 * 4 + 0) * CL + p] = kk.x; kT[(q4 * 4 + 1) * CL + p] = kk.y; kT[(q4 * 4 + 2) * CL + p] = kk.z; kT[(q4 * 4 + 3) * CL + p] = kk.w;
                vR[p * 4 + q4] = __ldcg(reinterpret_cast<const uint4*>(a.vq + (size_t)(lo + p) * a.kv_dim + g * 64) + q4);
            }
        }
        for (int idx = tid + FD_NT; idx < 2 * n; idx += FD_NT) {
            const int p = idx >> 1, bsel = idx & 1;
            if (lo + p != pos) {
                kd[bsel * CL + p] = h2f(__ldcg(a.ks + (size_t)(lo + p) * kvb + g * 2 + bsel));
                vd[p * 2 + bsel] = h2f(__ldcg(a.vs + (size_t)(lo + p) * kvb + g * 2 + bsel));
            }
        }
        __syncthreads();
        sync.stamp();
        // ---- this head's blocks: scores (x 1/sqrt(64)), block softmax numerators, E(P), P.V, merged into the running state
        if (head_warp) {
            float m_w = -INFINITY, l_w = 0.0f;
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; j++) acc[j] = 0.0f;
            uint32_t qx[16];
#pragma unroll
            for (int i = 0; i < 16; i++) qx[i] = qw[wid * 16 + i];
            const float qd0 = qd[wid * 2], qd1 = qd[wid * 2 + 1];
            for (int blk = 0; blk < nblk; blk++) {
                const int p = blk * 32 + lane;
                float s = -INFINITY;
                if (p < n) {
                    int i0 = 0, i1 = 0;
#pragma unroll
                    for (int w = 0; w < 8; w++) { i0 = __dp4a((int)kT[w * CL + p], (int)qx[w], i0); i1 = __dp4a((int)kT[(8 + w) * CL + p], (int)qx[8 + w], i1); }
                    s = 0.125f * fmaf((float)i1, qd1 * kd[CL + p], (float)i0 * (qd0 * kd[p]));
                }
                const float mb = warp_max(s);
                const float e = (p < n) ? __expf(s - mb) : 0.0f;
                float lb = e;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) lb += __shfl_xor_sync(0xffffffffu, lb, o);
                // Q8 re-encode of the probability row per 32 positions (ops.h:996): the block's largest e is exp(0) = 1, so the codes
                // are round(127 e); the block scale stays relative to the block maximum and is normalised at the combine
                const float pc = floorf(e * 127.0f + 0.5f) * (1.0f / 127.0f);
                const float m_new = fmaxf(m_w, mb);
                const float f_old = __expf(m_w - m_new), f_b = __expf(mb - m_new);       // m_w = -inf: f_old = 0
                l_w = fmaf(l_w, f_old, lb * f_b);
                m_w = m_new;
#pragma unroll
                for (int j = 0; j < 16; j++) acc[j] *= f_old;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = blk * 32 + slot + 8 * u;
                    const float pi = __shfl_sync(0xffffffffu, pc, slot + 8 * u) * f_b;
                    if (i < n) {
                        const float wgt = pi * vd[i * 2 + (cq >> 1)];          // ops.h:1026
                        const uint4 vw = vR[i * 4 + cq];
                        const uint32_t wds[4] = {vw.x, vw.y, vw.z, vw.w};
#pragma unroll
                        for (int j = 0; j < 16; j++) acc[j] = fmaf((float)(int)(int8_t)(wds[j >> 2] >> (8 * (j & 3))), wgt, acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {
                acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
                acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
                acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
            }
            float* part = a.parts + ((size_t)h * FD_MAXCH + c) * FD_PART;
            if (lane < 4) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(part + lane * 16 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            }
            if (lane == 4) { part[64] = m_w; part[65] = l_w; }
        }
        sync.stamp();
        // ---- the last chunk of this group to finish combines: o = sum_c o_c * exp(m_c - m) / sum_c l_c * exp(m_c - m), E(o)
        __threadfence();
        __syncthreads();
        if (tid == 0) *last = (atomicAdd(a.counters + g, 1u) == (unsigned)(nch - 1));
        __syncthreads();
        if (*last) {
            __threadfence();
            if (head_warp) {
                // lane cc holds chunk cc's (max, sum); all loads of the combine are in flight at once
                const float* p = a.parts + (size_t)h * FD_MAXCH * FD_PART;
                const float mc = (lane < nch) ? __ldcg(p + lane * FD_PART + 64) : -INFINITY;
                const float lc = (lane < nch) ? __ldcg(p + lane * FD_PART + 65) : 0.0f;
                const float m = warp_max(mc);
                const float fc = (lane < nch) ? __expf(mc - m) : 0.0f;
                float l = lc * fc;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
                float o0 = 0.0f, o1 = 0.0f;
                for (int c0 = 0; c0 < nch; c0 += 16) {          // 32 partial loads in flight per pass
                    float pv0[16], pv1[16];
#pragma unroll
                    for (int cc = 0; cc < 16; cc++) {
                        pv0[cc] = (c0 + cc < nch) ? __ldcg(p + (c0 + cc) * FD_PART + lane) : 0.0f;
                        pv1[cc] = (c0 + cc < nch) ? __ldcg(p + (c0 + cc) * FD_PART + 32 + lane) : 0.0f;
                    }
#pragma unroll
                    for (int cc = 0; cc < 16; cc++) {
                        const float f = __shfl_sync(0xffffffffu, fc, (c0 + cc) & 31);
                        o0 = fmaf(pv0[cc], f, o0);
                        o1 = fmaf(pv1[cc], f, o1);
                    }
                }
                const float inv = __fdividef(1.0f, l);
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    float dd;
                    const int q = fd_lane_encode((half ? o1 : o0) * inv, &dd);          // E(attention output), ops.h:1084
                    int sum = q;
#pragma unroll
                    for (int ofs = 16; ofs > 0; ofs >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, ofs);
                    const int b = 2 * h + half;
                    a.out.codes[b * 32 + pb] = (uint8_t)(int8_t)q;
                    if (lane == 0) { a.out.ad[b] = dd; a.out.n7[b] = -7 * sum; }
                }
            }
            if (tid == 0) a.counters[g] = 0u;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FD_NT, 2) k_fd_attn(FdAttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    fd_trace(4 * 4 + 0);
    fd_launch_dependents();
    FdPdlSync sync;
    fd_attn_phase(a, smem, sync, -1);
    fd_trace(4 * 4 + 2);
}

// ---------------------------------------------------------------- the same phases as ONE persistent cooperative kernel
// One CTA per SM; rows (prefill rows and greedy steps) loop inside; a phase boundary is a grid barrier (one relaxed atomic
// arrival per CTA + one polling thread per CTA) instead of a kernel boundary.  A CTA issues the register loads of its first
// weight rows of the NEXT phase before it arrives at the barrier, so weights never sit on the dependency chain.

struct FdMegaParams {
    const FdArgs* gemv;       // [4 * n_layers + 1]: per layer q|k|v, o, gate|up, down; the head last
    const FdAttnArgs* attn;   // [n_layers]
    int n_layers, n_heads;
    int n_body, n_head;       // rows without / with lm_head + argmax
    unsigned* bar;            // arrival counter (own 128-byte line), monotonic across launches
    unsigned bar_base;        // its value before this launch
    DevState* st;
    long long* prof; int prof_cta;
};

static inline size_t fd_mega_smem(int n_embd, int n_ffn, int max_ctx) {
    size_t s = fd_gemv_smem(n_embd, true);
    if (fd_gemv_smem(n_ffn, false) > s) s = fd_gemv_smem(n_ffn, false);
    if (fd_attn_smem(fd_chunk_len(max_ctx, 128)) > s) s = fd_attn_smem(fd_chunk_len(max_ctx, 128));
    return s;
}

struct FdGridSync {
    unsigned* bar;
    unsigned target;
    long long* prof;          // option "prof": SM-clock stamps of CTA `prof_cta` along the phases of the last row
    int idx;
    __device__ __forceinline__ void stamp() {
        if (prof && threadIdx.x == 0 && idx < 4000) prof[idx++] = clock64();
    }
    // arrive: this CTA's writes of the finished phase are published; wait: every CTA has arrived.  Weight loads of the
    // next phase are issued between the two (a fence after them would wait for them).
    __device__ __forceinline__ void arrive() {
        target += gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        }
    }
    __device__ __forceinline__ void wait() {
        if (threadIdx.x == 0) {
            unsigned v;
            long long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
                if ((int)(v - target) >= 0) break;
                if (++spins > (1ll << 26)) __trap();          // a lost CTA must not hang the GPU
            }
        }
        __syncthreads();
    }
};

// (A variant with two CTAs per SM -- 296 CTAs, 128 registers, every gate|up unit and 8 chunks per head in one wave -- measured
// slower: 762 vs 729 us per token; profiles/r01_04_fastdec.md.)
template <int WT>
__global__ void __launch_bounds__(FD_NT, 1) k_fd_mega(FdMegaParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int R2 = (WT == DT_Q4) ? 4 : 2, R6 = (WT == DT_Q4) ? 2 : 1, RS = 2 * R2;      // RS: a gate|up unit in one pass (Q4)
    FdGridSync sync{P.bar, P.bar_base, nullptr, 0};
    const int n_rows = P.n_body + P.n_head;
    int pos = __ldcg(&P.st->pos);          // rows advance it in lockstep in every CTA
    for (int r = 0; r < n_rows; r++, pos++) {
        if (r == n_rows - 1 && (int)blockIdx.x == P.prof_cta) sync.prof = P.prof;
        for (int li = 0; li < P.n_layers; li++) {
            fd_gemv_phase<WT, FD_NORM, FD_RAW, 2, R2>(P.gemv[4 * li + 0], smem, sync, pos);
            fd_attn_phase(P.attn[li], smem, sync, pos);
            fd_gemv_phase<WT, FD_CODES, FD_RAW, 2, R2>(P.gemv[4 * li + 1], smem, sync, pos);
            fd_gemv_phase<WT, FD_NORM, FD_SILU, 2, RS>(P.gemv[4 * li + 2], smem, sync, pos);
            fd_gemv_phase<WT, FD_CODES, FD_RAW, 6, R6>(P.gemv[4 * li + 3], smem, sync, pos);
        }
        if (r >= P.n_body) {
            fd_gemv_phase<WT, FD_NORM, FD_ARGMAX, 2, R2>(P.gemv[4 * P.n_layers], smem, sync, pos);
            sync.arrive(); sync.wait();                       // the new token, position and stop flag are visible to every CTA
            if (__ldcg(&P.st->stop)) break;
        } else {
            // nobody reads the position between the last attention phase and the next row's first barrier
            if (blockIdx.x == 0 && threadIdx.x == 0) P.st->pos = pos + 1;
        }
    }
}

// ---------------------------------------------------------------- batched decode: FDB_MAX sequences share every weight read
// SURVEY.md 8(f3).  The same phases with a sequence dimension: per-sequence buffers are [n_seq][...] arrays, every sequence
// has its own position (DevState), token row and K/V cache.  A weight block is loaded once and dotted against the staged
// vectors of all sequences; per sequence the arithmetic (block order, lane split, reductions) is exactly that of the
// single-sequence chain, so a sequence decoded in a batch gives the same bits as decoded alone (tests/test_fastdec_gpu.py).
// The redundant per-CTA NORM prologue would cost n_seq times as much, so the norm is its own small kernel here (one CTA per
// sequence) and every GEMV starts from staged codes.
constexpr int FDB_MAX = 16;         // sequences per batch at most; kernels are instantiated for 8 (R x 8 sums per lane) and 16 (R x 16)

__global__ void __launch_bounds__(FD_NT) k_fdb_norm(FdArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int K = a.K, nb = K / 32, s = blockIdx.x, tid = threadIdx.x;
    FdStaged sv;
    sv.aw = reinterpret_cast<uint32_t*>(smem);
    sv.ad = reinterpret_cast<float*>(smem + (size_t)nb * 32);
    sv.n7 = reinterpret_cast<int*>(smem + (size_t)nb * 36);
    float* scratch = reinterpret_cast<float*>(smem + ((((size_t)nb * 40) + 15) & ~(size_t)15));
    fd_trace(0 * 4 + 0);
    fd_launch_dependents();
    constexpr int NQ = 2;
    uint2 nw[NQ];
#pragma unroll
    for (int it = 0; it < NQ; it++) {
        const int qd = it * FD_NT + tid;
        nw[it] = *reinterpret_cast<const uint2*>(a.normw + 4 * ((qd < K / 4) ? qd : 0));
    }
    fd_wait_prior();
    fd_trace(0 * 4 + 1);
    FdArgs b = a;
    if (b.src0) b.src0 += (size_t)s * K;
    if (b.src1) b.src1 += (size_t)s * K;
    if (b.res_out) b.res_out += (size_t)s * K;
    if (b.tokens) b.tokens += (size_t)s * a.tok_stride;
    const int pos = a.emb_w ? __ldcg(&a.st[s].pos) : 0;
    fd_norm_stage<NQ>(b, sv, scratch, nw, pos, true);
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(a.act_out.codes + (size_t)s * K);
    for (int i = tid; i < K / 16; i += FD_NT) dst[i] = reinterpret_cast<const uint4*>(sv.aw)[i];
    for (int i = tid; i < nb; i += FD_NT) { a.act_out.ad[(size_t)s * nb + i] = sv.ad[i]; a.act_out.n7[(size_t)s * nb + i] = sv.n7[i]; }
    fd_trace(0 * 4 + 2);
}

static __host__ __device__ inline size_t fdb_gemv_smem(int K, int n_seq) {
    return (size_t)n_seq * (K / 32) * 40 + (size_t)FDB_MAX * 64 * 4 + 256;
}

// NSM = sequence slots the instantiation carries (8 or 16); R * NSM <= 32 row sums per lane
template <int WT, int EPI, int NBL, int R, int NSM>
__global__ void __launch_bounds__(FD_NT, 2) k_fdb_gemv(FdArgs a) {
    static_assert(R * NSM <= 32 && (R * NSM & (R * NSM - 1)) == 0, "the warp reduce-scatter hands one (row, sequence) sum to each lane");
    extern __shared__ __align__(16) unsigned char smem[];
    const int K = a.K, nb = K / 32, NS = a.n_seq;
    const int tid = threadIdx.x, lane = tid & 31, wid = fd_warp_id();
    // staged vectors of all sequences: [NS][nb][8] words, [NS][nb] scales, [NS][nb] -7 * code sums
    uint32_t* s_aw = reinterpret_cast<uint32_t*>(smem);
    float* s_ad = reinterpret_cast<float*>(smem + (size_t)NS * nb * 32);
    int* s_n7 = reinterpret_cast<int*>(smem + (size_t)NS * nb * 36);
    float* scratch = reinterpret_cast<float*>(smem + (((size_t)NS * nb * 40 + 15) & ~(size_t)15));          // [NSM][64]
    fd_trace((1 + EPI) * 4 + 0);
    fd_launch_dependents();
    if (tid < 2 && a.pf[tid]) {
        const size_t per = ((a.pf_bytes[tid] + gridDim.x - 1) / gridDim.x + 127) & ~(size_t)127;
        const size_t off = per * blockIdx.x;
        if (off < a.pf_bytes[tid]) l2_prefetch(reinterpret_cast<const unsigned char*>(a.pf[tid]) + off, min(per, a.pf_bytes[tid] - off));
    }
    int v0, v1, unit = blockIdx.x;
    const int n_units = (EPI == FD_SILU) ? a.n_ffn / 32 : 1;
    if (EPI == FD_SILU) { v0 = 0; v1 = (unit < n_units) ? 64 : 0; }
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
    fd_trace((1 + EPI) * 4 + 1);
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.in.codes);
        for (int i = tid; i < NS * (K / 16); i += FD_NT) reinterpret_cast<uint4*>(s_aw)[i] = __ldcg(src + i);
        for (int i = tid; i < NS * nb; i += FD_NT) { s_ad[i] = __ldcg(a.in.ad + i); s_n7[i] = __ldcg(a.in.n7 + i); }
    }
    __syncthreads();
    fd_trace((1 + EPI) * 4 + 3);
    float best = -INFINITY;          // ARGMAX: lane u * 8 + s tracks sequence s over its rows
    int arg = 0x7fffffff;
    for (; v < v1;) {
        float acc[R][NSM];
#pragma unroll
        for (int u = 0; u < R; u++)
#pragma unroll
            for (int s = 0; s < NSM; s++) acc[u][s] = 0.0f;
#pragma unroll
        for (int i = 0; i < NBL; i++) {
            const int b = lane + 32 * i;
            if (b < nb) {
                float dw[R];
#pragma unroll
                for (int u = 0; u < R; u++) dw[u] = h2f((uint16_t)bt.sc[u][i]);
#pragma unroll
                for (int s = 0; s < NSM; s++) {
                    if (s >= NS) break;          // a real (warp-uniform) branch: if-converted code would issue all 8 sequences
                    const uint4 ax = reinterpret_cast<const uint4*>(s_aw)[((size_t)s * nb + b) * 2];
                    const uint4 ay = reinterpret_cast<const uint4*>(s_aw)[((size_t)s * nb + b) * 2 + 1];
                    const int n7 = (WT == DT_Q4) ? s_n7[s * nb + b] : 0;
                    const float ad = s_ad[s * nb + b];
#pragma unroll
                    for (int u = 0; u < R; u++) {
                        const int is = fd_block_isum<WT>(bt.a[u][i], bt.b[(WT == DT_Q8) ? u : 0][(WT == DT_Q8) ? i : 0], ax, ay, n7);
                        acc[u][s] = fmaf((float)is, __fmul_rn(ad, dw[u]), acc[u][s]);
                    }
                }
            }
        }
        // warp reduce-scatter of the R x 8 sums: lane j ends up with the total of (row j / 8, sequence j % 8).  Offsets run
        // 16, 8, 4, 2, 1 like the butterfly of the single-sequence kernel, so every total is the same tree of additions.
        constexpr int NV = R * NSM;
        float vals[NV];
#pragma unroll
        for (int u = 0; u < R; u++)
#pragma unroll
            for (int s = 0; s < NSM; s++) vals[u * NSM + s] = acc[u][s];
        {
            int n = NV;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                if (n > off) {                       // halve: keep one half of the values, receive the partner's sums of it
                    const int half = n / 2;
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int k = 0; k < NV / 2; k++) {
                        if (k < half) {
                            const float send = up ? vals[k] : vals[k + half];
                            const float keep = up ? vals[k + half] : vals[k];
                            vals[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    n = half;
                } else {                             // fewer values than lanes: plain butterfly for the leading offsets
#pragma unroll
                    for (int k = 0; k < NV; k++) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], off);
                }
            }
        }
        const float mine = vals[0];
        const int vcur = v;
        int rcur[R];
#pragma unroll
        for (int u = 0; u < R; u++) rcur[u] = row[u];
        v += FD_NW * R;
        if (v < v1) { rows_of(v, row); fd_load_batch<WT, NBL, R>(bt, a.w, a.ws, row, nb); }
        const int mu = (lane & (NV - 1)) / NSM, ms = lane % NSM;
        int mrow = -1;
#pragma unroll
        for (int u = 0; u < R; u++) if (mu == u) mrow = rcur[u];
        if (lane < NV && ms < NS && mrow >= 0) {
            if (EPI == FD_SILU) {
                scratch[ms * 64 + vcur + mu] = mine;
            } else {
                a.out[(size_t)ms * a.n_rows + mrow] = mine;
                if (EPI == FD_ARGMAX && mine > best) { best = mine; arg = mrow; }
            }
        }
    }
    fd_trace((1 + EPI) * 4 + 2);
    if (EPI == FD_SILU && unit < n_units) {
        __syncthreads();
        for (int sq = wid; sq < NS; sq += FD_NW) {          // a warp per sequence: E(E(silu(E(gate))) * E(up)) (modules.cpp:238-247)
            const float g1 = fd_lane_roundtrip(scratch[sq * 64 + lane]);
            const float u1 = fd_lane_roundtrip(scratch[sq * 64 + 32 + lane]);
            const float g2 = fd_lane_roundtrip(__fdividef(g1, 1.0f + __expf(-g1)));
            float dd;
            const int q = fd_lane_encode(g2 * u1, &dd);
            int sum = q;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            a.act_out.codes[(size_t)sq * a.n_ffn + unit * 32 + perm_byte(lane)] = (uint8_t)(int8_t)q;
            if (lane == 0) { a.act_out.ad[(size_t)sq * n_units + unit] = dd; a.act_out.n7[(size_t)sq * n_units + unit] = -7 * sum; }
        }
    }
    if (EPI == FD_ARGMAX) {
        // rows of this warp for sequence ms: lanes ms, NSM + ms, ...; then the 8 warps; then the last CTA over all CTAs
        float* sval = scratch;                                          // [FD_NW][NSM]
        int* sidx = reinterpret_cast<int*>(scratch + FD_NW * NSM);      // [FD_NW][NSM]
        __shared__ bool is_last;
#pragma unroll
        for (int o = NSM; o <= 16; o <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
        }
        if (lane < NSM) { sval[wid * NSM + lane] = best; sidx[wid * NSM + lane] = arg; }
        __syncthreads();
        if (tid < NS) {
            best = sval[tid]; arg = sidx[tid];
            for (int w = 1; w < FD_NW; w++) {
                const float ov = sval[w * NSM + tid];
                const int oi = sidx[w * NSM + tid];
                if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
            }
            a.arg_val[(size_t)tid * gridDim.x + blockIdx.x] = best;
            a.arg_idx[(size_t)tid * gridDim.x + blockIdx.x] = arg;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
        __syncthreads();
        if (is_last) {
            __threadfence();
            for (int sq = wid; sq < NS; sq += FD_NW) {
                best = -INFINITY; arg = 0x7fffffff;
                for (int c = lane; c < (int)gridDim.x; c += 32) {
                    const float ov = __ldcg(a.arg_val + (size_t)sq * gridDim.x + c);
                    const int oi = __ldcg(a.arg_idx + (size_t)sq * gridDim.x + c);
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
                    DevState* st = a.st + sq;
                    const int pcur = __ldcg(&st->pos);
                    a.tok_out[(size_t)sq * a.tok_stride + pcur + 1] = arg;
                    st->pos = pcur + 1;
                    st->n_gen = __ldcg(&st->n_gen) + 1;
                }
            }
        }
        if (is_last && tid == 0) *a.counter = 0u;
    }
}

// attention of sequence blockIdx.y: the single-sequence phase on that sequence's slices
__global__ void __launch_bounds__(FD_NT, 2) k_fdb_attn(FdAttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    fd_trace(4 * 4 + 0);
    fd_launch_dependents();
    const int s = blockIdx.y;
    FdAttnArgs b = a;
    b.rqkv += (size_t)s * (a.n_embd + 2 * a.kv_dim);
    b.kq += s * a.seq_kv_codes; b.vq += s * a.seq_kv_codes; b.ks += s * a.seq_kv_scales; b.vs += s * a.seq_kv_scales;
    b.st += s;
    b.parts += (size_t)s * a.n_heads * FD_MAXCH * FD_PART;
    b.counters += (size_t)s * (a.n_heads + 1);
    b.out.codes += (size_t)s * a.n_embd; b.out.ad += (size_t)s * (a.n_embd / 32); b.out.n7 += (size_t)s * (a.n_embd / 32);
    FdPdlSync sync;
    fd_attn_phase(b, smem, sync, -1);
    fd_trace(4 * 4 + 2);
}

}  // namespace gtb
