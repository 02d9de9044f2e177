// gtb_engine.cu -- the whole TinyLlama::logits graph (tinyllama.cpp:45-61) resident on one GPU.
//
// One sequence row = 5 phase kernels per layer (see gtb_kernels.cuh) + embedding + lm_head + argmax, all on
// the library stream; position / token live in device memory so a captured CUDA graph replays every row
// and the greedy loop (tinyllama.cpp:395-440) never returns to the host between tokens.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <fstream>
#include <string>
#include <vector>

#include "gtb_internal.h"
#include "gtb_kernels.cuh"
#include "gtb_mega.cuh"
#include "gtb_prefill.h"
#include "gtb_fastdec.cuh"
#include "gtb_xrows.h"

namespace gtb {

enum { PRO_ENCODE = 0, PRO_NORM = 1, PRO_SILU_MUL = 2 };

struct GemvMat {
    const void* data;
    const uint16_t* scales;
    int rows;
    float* out;
};

struct PhaseArgs {
    int K;
    int n_mats;
    GemvMat mat[3];
    const float* src0;
    const float* src1;
    const uint16_t* normw;
    float* res_out;
    float* cap0;
    float* cap1;
    float* cap2;
};

template <int WT, int PRO>
__global__ void __launch_bounds__(NT) k_phase_gemv(PhaseArgs a) {
    constexpr int AT = (WT == DT_F16) ? DT_F16 : DT_Q8;
    extern __shared__ __align__(16) unsigned char smem[];
    const int K = a.K;
    // carve: [act][ps (per warp)][xbuf][ExactSumSmem]
    ActView av = act_carve(AT, K, smem);
    size_t off = (act_bytes(AT, K) + 15) & ~(size_t)15;
    float* ps = reinterpret_cast<float*>(smem + off);
    const size_t ps_warp = gemv_ps_bytes(WT, K) / NWARP;
    off += gemv_ps_bytes(WT, K);
    const bool c0 = blockIdx.x == 0;
    if (PRO == PRO_ENCODE) {
        pro_encode<AT>(av, a.src0, K, c0 ? a.cap0 : nullptr);
    } else if (PRO == PRO_NORM) {
        float* xbuf = reinterpret_cast<float*>(smem + off);
        off += (size_t)K * 4;
        off = (off + 15) & ~(size_t)15;
        ExactSumSmem& es = *reinterpret_cast<ExactSumSmem*>(smem + off);
        pro_norm<AT>(av, a.src0, a.src1, a.normw, K, xbuf, es, c0 ? a.res_out : nullptr,
                     c0 ? a.cap0 : nullptr, c0 ? a.cap1 : nullptr, c0 ? a.cap2 : nullptr);
    } else {
        pro_silu_mul<AT>(av, a.src0, a.src1, K, c0 ? a.cap0 : nullptr, c0 ? a.cap1 : nullptr);
    }
    __syncthreads();
    float* my_ps = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ps) + ps_warp * (threadIdx.x >> 5));
    int done = 0;
    for (int m = 0; m < a.n_mats; m++)
        gemv_matrix<WT>(a.mat[m].data, a.mat[m].scales, a.mat[m].rows, K, av, my_ps, a.mat[m].out, done, &done);
}

static size_t phase_smem(int wt, int pro, int K) {
    const int at = (wt == DT_F16) ? DT_F16 : DT_Q8;
    size_t s = (act_bytes(at, K) + 15) & ~(size_t)15;
    s += gemv_ps_bytes(wt, K);
    if (pro == PRO_NORM) { s += (size_t)K * 4; s = (s + 15) & ~(size_t)15; s += sizeof(ExactSumSmem); }
    return s + 16;
}

// ---------------------------------------------------------------- embedding (gten/ops.h:514-564)
template <int WT>
__global__ void __launch_bounds__(NT) k_embed(const void* __restrict__ wdata, const uint16_t* __restrict__ wsc, int n_embd,
                                               const int32_t* __restrict__ tokens, const DevState* __restrict__ st,
                                               float* __restrict__ xres, float* cap) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t row = (size_t)tokens[st->pos];
    for (int b = wid; b < n_embd / 32; b += NWARP) {
        const int e = b * 32 + lane;
        float v;
        if (WT == DT_F16) {
            const int c = e >> 6, r = e & 63, l = r & 7, ii = r >> 3;
            v = h2f(reinterpret_cast<const uint16_t*>(wdata)[((row * (n_embd / 64) + c) * 8 + l) * 8 + ii]);   // memcpy of the fp16 row
        } else {
            const size_t blk = row * (n_embd / 32) + b;
            const float delta = h2f(wsc[blk]);
            const int pb = perm_byte(lane);
            if (WT == DT_Q8) {
                const int8_t q = reinterpret_cast<const int8_t*>(wdata)[blk * 32 + pb];                      // memcpy of the Q8 row
                v = __fmul_rn((float)q, delta);
            } else {
                // Q4 row: dequantise, then re-encode as Q8 (ops.h:522-528)
                const int j = lane & 15;
                const int l = (j & 7) >> 1, pos = (j & 1) + 2 * (j >> 3);
                const uint8_t byte = reinterpret_cast<const uint8_t*>(wdata)[blk * 16 + l * 4 + pos];
                const int q = (int)((lane < 16) ? (byte >> 4) : (byte & 0x0f)) - 7;
                v = q8_roundtrip_lane(__fmul_rn((float)q, delta));
            }
        }
        xres[e] = v;
        if (cap) cap[e] = v;
    }
}

// ---------------------------------------------------------------- attention phase
struct AttnArgs {
    const float* rqkv;        // raw q | k | v of this row
    int n_embd, kv_dim, n_heads, gsz;
    uint8_t* kq; uint16_t* ks; uint8_t* vq; uint16_t* vs;
    const float* rope_cos; const float* rope_sin;    // [max_ctx][32], host-built with the reference's libm calls
    const DevState* st;
    float* out;               // raw attention output [n_embd]
    float* cap_q; float* cap_k; float* cap_v;
};

template <int AT>
__global__ void __launch_bounds__(NT) k_attn(AttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem);
    float* sc = reinterpret_cast<float*>(smem + ((sizeof(AttnSmem) + 15) & ~(size_t)15));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int h = blockIdx.x, g = h / a.gsz;
    const int pos = a.st->pos;
    const int n_ctx = max(a.st->nctx_min, pos + 1);
    const bool writer = (h % a.gsz) == 0;
    // first re-encode of the Linear outputs (q: warps 0,1; k: 2,3; v: 4,5), gten/ops.h:645-646
    if (wid < 6) {
        const int which = wid >> 1, half = wid & 1;
        const float* src = (which == 0) ? a.rqkv + h * 64 : (which == 1) ? a.rqkv + a.n_embd + g * 64 : a.rqkv + a.n_embd + a.kv_dim + g * 64;
        const float x = src[half * 32 + lane];
        if (which < 2) {
            sm.tmp[wid][lane] = roundtrip<AT>(x);
        } else if (AT == DT_F16) {
            const uint16_t hb = f2h(x);
            sm.vf[half * 32 + lane] = h2f(hb);
            if (writer) reinterpret_cast<uint16_t*>(a.vq)[(size_t)pos * a.kv_dim + g * 64 + half * 32 + lane] = hb;
            if (writer && a.cap_v) a.cap_v[g * 64 + half * 32 + lane] = h2f(hb);
        } else {
            uint16_t dh;
            const int q = q8_encode_lane(x, &dh);
            const float d = __fmul_rn((float)q, h2f(dh));
            sm.vf[half * 32 + lane] = d;
            if (writer) {
                a.vq[(size_t)pos * a.kv_dim + g * 64 + half * 32 + lane] = (uint8_t)(int8_t)q;
                if (lane == 0) a.vs[(size_t)pos * (a.kv_dim / 32) + g * 2 + half] = dh;
                if (a.cap_v) a.cap_v[g * 64 + half * 32 + lane] = d;
            }
        }
    }
    __syncthreads();
    // RoPE on q and k (gten/ops.h:733-751), then the second re-encode (ops.h:753)
    if (wid < 4) {
        const int which = wid >> 1, half = wid & 1;
        const float x0 = sm.tmp[which * 2][lane], x1 = sm.tmp[which * 2 + 1][lane];
        const float cs = a.rope_cos[(size_t)pos * 32 + lane], sn = a.rope_sin[(size_t)pos * 32 + lane];
        const float o = (half == 0) ? __fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn))
                                    : __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs));
        if (AT == DT_F16) {
            const uint16_t hb = f2h(o);
            const float d = h2f(hb);
            if (which == 0) { sm.qf[half * 32 + lane] = d; if (a.cap_q) a.cap_q[h * 64 + half * 32 + lane] = d; }
            else {
                sm.kf[half * 32 + lane] = d;
                if (writer) reinterpret_cast<uint16_t*>(a.kq)[(size_t)pos * a.kv_dim + g * 64 + half * 32 + lane] = hb;
                if (writer && a.cap_k) a.cap_k[g * 64 + half * 32 + lane] = d;
            }
        } else {
            uint16_t dh;
            const int q = q8_encode_lane(o, &dh);
            const float delta = h2f(dh);
            const int pb = perm_byte(lane);
            if (which == 0) {
                reinterpret_cast<int8_t*>(sm.qw)[half * 32 + pb] = (int8_t)q;
                if (lane == 0) sm.qd[half] = delta;
                if (a.cap_q) a.cap_q[h * 64 + half * 32 + lane] = __fmul_rn((float)q, delta);
            } else {
                reinterpret_cast<int8_t*>(sm.kw)[half * 32 + pb] = (int8_t)q;
                if (lane == 0) sm.kd[half] = delta;
                if (writer) {
                    a.kq[(size_t)pos * a.kv_dim + g * 64 + half * 32 + pb] = (uint8_t)(int8_t)q;
                    if (lane == 0) a.ks[(size_t)pos * (a.kv_dim / 32) + g * 2 + half] = dh;
                    if (a.cap_k) a.cap_k[g * 64 + half * 32 + lane] = __fmul_rn((float)q, delta);
                }
            }
        }
    }
    __syncthreads();
    KVCache kv{a.kq, a.ks, a.vq, a.vs, a.kv_dim};
    attn_core<AT>(sm, sc, kv, g, pos, n_ctx, true, a.out + h * 64);
}

// ---------------------------------------------------------------- argmax + bookkeeping (tinyllama.cpp:416-434)
__global__ void __launch_bounds__(1024) k_argmax_advance(const float* __restrict__ logits, int n, int32_t* tokens,
                                                          DevState* st, int eos_id) {
    __shared__ float sv[32];
    __shared__ int si[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int j = tid; j < n; j += blockDim.x) {
        const float v = logits[j];
        if (v > best) { best = v; arg = j; }        // ascending j per thread: first maximum wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
    }
    if (lane == 0) { sv[wid] = best; si[wid] = arg; }
    __syncthreads();
    if (wid == 0) {
        best = sv[lane]; arg = si[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
        }
        if (lane == 0) {
            if (arg == 0x7fffffff) arg = 0;          // all -inf/NaN: the reference's loop leaves max_index = 0
            const int pos = st->pos;
            tokens[pos + 1] = arg;
            st->pos = pos + 1;
            st->n_gen += 1;
            if (arg == eos_id) st->stop = 1;
        }
    }
}

__global__ void k_advance(DevState* st) { st->pos += 1; }
// rows [p0, p1) of a layer's natural K/V rows -> the chunk-major copies k_mega reads (one 16-byte chunk per thread; the thread
// of chunk c < 2 * n_groups also moves one scale of each cache)
__global__ void k_kv_transpose(const uint8_t* __restrict__ kq, const uint8_t* __restrict__ vq, const uint16_t* __restrict__ ks,
                               const uint16_t* __restrict__ vs, uint8_t* __restrict__ kqt, uint8_t* __restrict__ vqt,
                               uint16_t* __restrict__ kst, uint16_t* __restrict__ vst, int p0, int p1, int max_ctx, int n_groups) {
    const int cpr = n_groups * 4;                                  // 16-byte chunks per row
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)(p1 - p0) * cpr) return;
    const int pos = p0 + (int)(i / cpr), c = (int)(i % cpr);
    const size_t src = ((size_t)pos * cpr + c) * 16;
    *reinterpret_cast<uint4*>(kqt + ((size_t)c * max_ctx + pos) * 16) = *reinterpret_cast<const uint4*>(kq + src);
    *reinterpret_cast<uint4*>(vqt + ((size_t)c * max_ctx + pos) * 16) = *reinterpret_cast<const uint4*>(vq + src);
    if (c < 2 * n_groups) {
        kst[(size_t)c * max_ctx + pos] = ks[(size_t)pos * 2 * n_groups + c];
        vst[(size_t)c * max_ctx + pos] = vs[(size_t)pos * 2 * n_groups + c];
    }
}

__global__ void k_set_pos(DevState* st, int pos) { st->pos = pos; }     // stream-ordered, everything else untouched

// The k largest logits and their token ids, largest first, ties to the lower id (the candidate set of topk_sample,
// tinyllama.cpp:466-478): k rounds of a block-wide arg-max over the not-yet-taken entries.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) k_topk(const float* __restrict__ logits, int n, int k, float* __restrict__ out_v, int32_t* __restrict__ out_i) {
    __shared__ float sv[32];
    __shared__ int si[32];
    __shared__ int taken[64];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int r = 0; r < k; r++) {
        float best = -INFINITY;
        int arg = 0x7fffffff;
        for (int j = tid; j < n; j += 1024) {
            const float v = logits[j];
            bool skip = false;
            for (int q = 0; q < r; q++) skip |= (taken[q] == j);
            if (!skip && (v > best || (v == best && j < arg) || arg == 0x7fffffff)) { best = v; arg = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
            if (oi != 0x7fffffff && (arg == 0x7fffffff || ov > best || (ov == best && oi < arg))) { best = ov; arg = oi; }
        }
        if (lane == 0) { sv[wid] = best; si[wid] = arg; }
        __syncthreads();
        if (wid == 0) {
            best = sv[lane]; arg = si[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
                if (oi != 0x7fffffff && (arg == 0x7fffffff || ov > best || (ov == best && oi < arg))) { best = ov; arg = oi; }
            }
            if (lane == 0) { taken[r] = arg; out_v[r] = best; out_i[r] = arg; }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- host side
// self-test of the megakernel's exact in-order sum on arbitrary terms (one CTA of MT threads)
__global__ void __launch_bounds__(MT) k_selftest_exact_sum(const float* __restrict__ terms, int n, float* out) {
    __shared__ ExactSum2Smem es;
    float r = 0.0f;
    for (int rep = 0; rep < 4; rep++) {          // repetitions 1..3 run with warm instruction caches: out[1 + rep] = cycles
        __syncthreads();
        const long long t0 = clock64();
        r = exact_sum512([&](int i, float q[4]) {
#pragma unroll
            for (int u = 0; u < 4; u++) q[u] = (i + u < n) ? terms[i + u] : 0.0f;
        }, n, es);
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[1 + rep] = (float)(t1 - t0);
    }
    if (threadIdx.x == 0) *out = r;
}

// glibc-exact expf of the bit patterns [first, first + count) (exhaustive self-test against the host libm)
__global__ void k_selftest_expf(uint32_t first, uint32_t count, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = expf_glibc(__uint_as_float(first + (uint32_t)i));
}

struct LayerW {
    gtb_weight_t q = nullptr, k = nullptr, v = nullptr, o = nullptr, gate = nullptr, up = nullptr, down = nullptr;
    uint16_t* attn_norm = nullptr;
    uint16_t* ffn_norm = nullptr;
    uint8_t *kq = nullptr, *vq = nullptr;
    uint16_t *ks = nullptr, *vs = nullptr;
    uint8_t *kqt = nullptr, *vqt = nullptr;      // chunk-major copies of the Q8 K/V codes and scales for k_mega (gtb_mega.cuh: MegaLayer)
    uint16_t *kst = nullptr, *vst = nullptr;
    // q|k|v and gate|up live in one buffer each so that a GEMV phase of the persistent kernel streams ONE matrix
    uint8_t *qkv_data = nullptr, *gu_data = nullptr;
    uint16_t *qkv_sc = nullptr, *gu_sc = nullptr;
};

}  // namespace gtb

using namespace gtb;

struct gtb_engine {
    gtb_model_config cfg{};
    int adtype = 0, d_head = 0, kv_dim = 0, gsz = 0;
    gtb_weight_t embed = nullptr, lm_head = nullptr;
    uint16_t* final_norm = nullptr;
    std::vector<LayerW> L;
    float *rope_cos = nullptr, *rope_sin = nullptr;
    // per-row buffers
    float *xres = nullptr, *hres = nullptr, *rqkv = nullptr, *rattn = nullptr, *ro = nullptr, *rg = nullptr, *ru = nullptr,
          *rd = nullptr, *logits = nullptr, *xfinal = nullptr;
    int32_t* tokens = nullptr;
    DevState* st = nullptr;
    float* cap = nullptr;            // [n_layers][12][capw] + [2][capw]
    int capw = 0;
    bool capture = false;
    bool use_graph = true;
    cudaGraphExec_t g_body = nullptr, g_head = nullptr;
    int g_eos = -2;
    int grid = 0;
    int host_pos = 0;
    int kvt_pos = 0;                 // positions [0, kvt_pos) of the transposed K/V copies are current
    long long mega_rows = 0;         // rows run by k_mega since the exchange buffers were last cleared (epoch tags are 32-bit)
    size_t weight_bytes = 0;
    int launches_body = 0, launches_head = 0;
    // persistent megakernel (gtb_mega.cuh)
    bool use_mega = true;
    int pf_ahead = 4;
    bool prof = false;
    int prof_cta = 0;                // which CTA of k_mega writes the stamps
    int batch_eos = -1;              // exact batched decode: a slot that samples this id stops (tinyllama.cpp:426); -1: none
    MegaLayer* d_layers = nullptr;
    bool layers_valid = false;
    unsigned long long *x_qkv = nullptr, *x_sc = nullptr, *x_attn = nullptr, *x_o = nullptr, *x_gu = nullptr, *x_act = nullptr,
                       *x_down = nullptr, *x_arg = nullptr, *dbg = nullptr;
    unsigned int* epoch = nullptr;
    unsigned int* cnt = nullptr;
    long long* d_prof = nullptr;
    int sc_stride = 0;
    // batched prefill (gtb_prefill.cu): fp16 weight copies + activation workspace, built on first use
    gtb::PfPlan* pf = nullptr;
    float* pf_cap = nullptr;         // [n_layers*12 + 1][pf_cap_T][capw] when capture_acv is on
    int pf_cap_T = 0;
    // fast (order-free) decode kernels (gtb_fastdec.cuh): opt-in, tolerance-level parity
    bool fast = false;
    float* fd_parts = nullptr;       // [n_heads][FD_MAXCH][FD_PART] attention partials
    unsigned* fd_cnt = nullptr;      // [n_heads + 1] arrival counters (attention chunks per head; head CTAs)
    FdAct fd_attn{}, fd_act{};       // E(attention output) [n_embd], E(MLP activation) [n_ffn]: staged codes for the next GEMV
    float* fd_argv = nullptr;        // per-CTA maxima of the head kernel
    int* fd_argi = nullptr;
    int fd_ahead = 3;                // L2 look-ahead distance in GEMV steps
    // batched decode (gtb_engine_batch_*): per-sequence buffers [n][...] and K/V caches [n][max_ctx][kv_dim] per layer
    struct BatchBufs {
        int n = 0;
        float *rqkv = nullptr, *ro = nullptr, *rd = nullptr, *xres = nullptr, *hres = nullptr, *xfinal = nullptr, *logits = nullptr;
        float *parts = nullptr, *argv = nullptr;
        int* argi = nullptr;
        unsigned* cnt = nullptr;
        int32_t* tokens = nullptr;
        DevState* st = nullptr;
        FdAct norm_act{}, attn_act{}, mlp_act{};
        std::vector<uint8_t*> kq, vq;
        std::vector<uint16_t*> ks, vs;
        std::vector<int> host_pos;
        cudaGraphExec_t graph = nullptr;
        int graph_launches = 0;
        int graph_tcap = 0;          // exact mode: score-buffer capacity the captured graph was built for
        bool graph_exact = false;
    } bb;
    int fd_chunk = 64;               // positions per attention chunk of the single-sequence fast path (the batch uses 128)
    int fd_prof_cta = 0;             // which CTA writes the "prof" stamps of k_fd_mega
    bool fd_mega = false;            // fast decode as one persistent cooperative kernel (k_fd_mega, measured slower); false: PDL-chained kernels
    FdArgs* d_fd_gemv = nullptr;     // phase arguments of k_fd_mega
    FdAttnArgs* d_fd_attn = nullptr;
    unsigned* fd_bar = nullptr;      // grid-barrier arrival counter
    bool fd_args_valid = false, fd_args_head = false;
    int fd_args_eos = -2;
    int pf_layers = 0;               // debug: run only the first pf_layers layers (0 = all)
    int pf_fused = 1;                // RoPE/KV append and SiLU*up inside the GEMM epilogues
    int pf_2cta = 0;                 // CTA-pair GEMM kernel
    int pf_pdl = 1;                  // programmatic dependent launch inside the batched prefill
    int pf_attn2 = 0;                // 1: two-sweep attention (also reproduces the fp16 rounding of the P-row block scales)
    // order-exact multi-row path (gtb_xrows.cu): exact prefill in passes of up to 64 rows, exact batched decode
    gtb::XrPlan* xr = nullptr;
    bool use_xr = true;
    int xr_min_rows = 4;             // fewer rows than this stay on the row-at-a-time kernels
    int xr_rows = 512;               // rows per prefill pass (2024-token prompt: 159 ms at 64 rows per pass, 107 ms at 512, 103 ms at 1024)
    bool batch_exact = true;         // gtb_engine_batch_decode through the exact multi-row kernels (false: order-free kernels)
    std::vector<XrLayerW> xr_layers;
    std::vector<uint8_t*> xr_kq, xr_vq;
    std::vector<uint16_t*> xr_ks, xr_vs;
};

namespace {

float* capp(gtb_engine* e, int layer, int aid) {
    if (!e->capture) return nullptr;
    if (aid == GTB_A_EMB) return e->cap + (size_t)e->cfg.n_layers * 12 * e->capw;
    if (aid == GTB_A_FINAL_NORM) return e->cap + ((size_t)e->cfg.n_layers * 12 + 1) * e->capw;
    return e->cap + ((size_t)layer * 12 + (aid - GTB_A_ATTN_NORM)) * e->capw;
}

template <int WT, int PRO>
int launch_phase(gtb_engine* e, const PhaseArgs& a) {
    const size_t smem = phase_smem(WT, PRO, a.K);
    static bool attr_done = false;            // one per template instantiation
    if (!attr_done) {
        GTB_CUDA(cudaFuncSetAttribute(k_phase_gemv<WT, PRO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    k_phase_gemv<WT, PRO><<<e->grid, NT, smem, ctx().stream>>>(a);
    GTB_LAUNCHED();
    return GTB_OK;
}

GemvMat matof(gtb_weight_t w, float* out) { return GemvMat{w->data, w->scales, w->rows, out}; }

template <int WT>
int enqueue_row(gtb_engine* e, bool with_head, int eos_id) {
    constexpr int AT = (WT == DT_F16) ? DT_F16 : DT_Q8;
    const gtb_model_config& c = e->cfg;
    cudaStream_t st = ctx().stream;
    const int E = c.n_embd, F = c.n_ffn, KV = e->kv_dim;
    k_embed<WT><<<1, NT, 0, st>>>(e->embed->data, e->embed->scales, E, e->tokens, e->st, e->xres, capp(e, 0, GTB_A_EMB));
    GTB_LAUNCHED();
    static bool attn_attr = false;
    if (!attn_attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_attn<AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attn_attr = true;
    }
    const size_t attn_smem = ((sizeof(AttnSmem) + 15) & ~(size_t)15) + (size_t)((c.max_ctx + 63) / 32 * 32) * 4;
    for (int li = 0; li < c.n_layers; li++) {
        LayerW& l = e->L[li];
        {   // P1: residual/norm prologue + q|k|v
            PhaseArgs a{};
            a.K = E; a.n_mats = 3;
            a.mat[0] = matof(l.q, e->rqkv); a.mat[1] = matof(l.k, e->rqkv + E); a.mat[2] = matof(l.v, e->rqkv + E + KV);
            a.src0 = (li == 0) ? e->xres : e->hres;
            a.src1 = (li == 0) ? nullptr : e->rd;
            a.normw = l.attn_norm;
            a.res_out = e->xres;
            a.cap0 = (li == 0) ? nullptr : capp(e, li - 1, GTB_A_DOWN);
            a.cap1 = (li == 0) ? nullptr : capp(e, li - 1, GTB_A_ATTN_RES);
            a.cap2 = capp(e, li, GTB_A_ATTN_NORM);
            int r = launch_phase<WT, PRO_NORM>(e, a);
            if (r) return r;
        }
        {   // P2: attention
            AttnArgs a{};
            a.rqkv = e->rqkv; a.n_embd = E; a.kv_dim = KV; a.n_heads = c.n_heads; a.gsz = e->gsz;
            a.kq = l.kq; a.ks = l.ks; a.vq = l.vq; a.vs = l.vs;
            a.rope_cos = e->rope_cos; a.rope_sin = e->rope_sin; a.st = e->st; a.out = e->rattn;
            a.cap_q = capp(e, li, GTB_A_Q); a.cap_k = capp(e, li, GTB_A_K); a.cap_v = capp(e, li, GTB_A_V);
            k_attn<AT><<<c.n_heads, NT, attn_smem, st>>>(a);
            GTB_LAUNCHED();
        }
        {   // P3: o-proj
            PhaseArgs a{};
            a.K = E; a.n_mats = 1; a.mat[0] = matof(l.o, e->ro);
            a.src0 = e->rattn; a.cap0 = capp(e, li, GTB_A_ATTN_OUT);
            int r = launch_phase<WT, PRO_ENCODE>(e, a);
            if (r) return r;
        }
        {   // P4: residual/norm prologue + gate|up
            PhaseArgs a{};
            a.K = E; a.n_mats = 2; a.mat[0] = matof(l.gate, e->rg); a.mat[1] = matof(l.up, e->ru);
            a.src0 = e->xres; a.src1 = e->ro; a.normw = l.ffn_norm; a.res_out = e->hres;
            a.cap0 = capp(e, li, GTB_A_O); a.cap1 = capp(e, li, GTB_A_INP_RES); a.cap2 = capp(e, li, GTB_A_FFN_NORM);
            int r = launch_phase<WT, PRO_NORM>(e, a);
            if (r) return r;
        }
        {   // P5: SiLU*up prologue + down
            PhaseArgs a{};
            a.K = F; a.n_mats = 1; a.mat[0] = matof(l.down, e->rd);
            a.src0 = e->rg; a.src1 = e->ru; a.cap0 = capp(e, li, GTB_A_GATE); a.cap1 = capp(e, li, GTB_A_UP);
            int r = launch_phase<WT, PRO_SILU_MUL>(e, a);
            if (r) return r;
        }
    }
    if (with_head || e->capture) {   // final residual + norm + lm_head (tinyllama.cpp:57-58)
        PhaseArgs a{};
        a.K = E; a.n_mats = with_head ? 1 : 0;
        a.mat[0] = matof(e->lm_head, e->logits);
        a.src0 = e->hres; a.src1 = e->rd; a.normw = e->final_norm; a.res_out = e->xfinal;
        a.cap0 = capp(e, c.n_layers - 1, GTB_A_DOWN); a.cap1 = capp(e, c.n_layers - 1, GTB_A_ATTN_RES);
        a.cap2 = capp(e, 0, GTB_A_FINAL_NORM);
        int r = launch_phase<WT, PRO_NORM>(e, a);
        if (r) return r;
    }
    if (with_head) {
        k_argmax_advance<<<1, 1024, 0, st>>>(e->logits, c.n_vocab, e->tokens, e->st, eos_id);
        GTB_LAUNCHED();
    } else {
        k_advance<<<1, 1, 0, st>>>(e->st);
        GTB_LAUNCHED();
    }
    return GTB_OK;
}

// ---- order-free decode (gtb_fastdec.cuh): launches chained with programmatic dependent launch
template <typename... KArgs, typename... Args>
cudaError_t fd_launch(void (*kern)(KArgs...), int grid, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(FD_NT); cfg.dynamicSmemBytes = smem; cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <int WT, int PRO, int EPI, int NBL>
int launch_fd_k(const FdArgs& a, int grid) {
    constexpr int R = (WT == DT_Q4) ? ((NBL <= 2) ? 4 : 2) : ((NBL <= 2) ? 2 : 1);
    const size_t smem = fd_gemv_smem(a.K, PRO == FD_NORM);
    static bool attr_done = false;
    if (!attr_done) {
        GTB_CUDA(cudaFuncSetAttribute(k_fd_gemv<WT, PRO, EPI, NBL, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_done = true;
    }
    GTB_CUDA(fd_launch(k_fd_gemv<WT, PRO, EPI, NBL, R>, grid, smem, a));
    GTB_LAUNCHED();
    return GTB_OK;
}
template <int WT, int PRO, int EPI>
int launch_fd(const FdArgs& a, int grid) {
    const int nbl = (a.K / 32 + 31) / 32;
    if (nbl <= 2) return launch_fd_k<WT, PRO, EPI, 2>(a, grid);
    if (nbl <= 6) return launch_fd_k<WT, PRO, EPI, 6>(a, grid);
    return fail(GTB_ERR_STATE, "fast_decode: rows longer than 6144 elements are not instantiated");
}

// arguments of every phase of one row: gv = [4 * L (+ 1)] GEMV phases in launch order, av = [L] attention phases
int build_fast_args(gtb_engine* e, bool with_head, int eos_id, std::vector<FdArgs>& gv, std::vector<FdAttnArgs>& av) {
    const gtb_model_config& c = e->cfg;
    const int E = c.n_embd, F = c.n_ffn, KV = e->kv_dim, L = c.n_layers;
    const int wd = c.wdtype;
    if (e->gsz > FD_NW) return fail(GTB_ERR_STATE, "fast_decode: at most %d query heads per KV group", FD_NW);
    gv.clear(); av.clear();
    for (int li = 0; li < L; li++) {
        LayerW& l = e->L[li];
        FdArgs a{};
        a.K = E; a.n_rows = E + 2 * KV; a.w = (const uint4*)l.qkv_data; a.ws = l.qkv_sc; a.out = e->rqkv;
        a.normw = l.attn_norm; a.res_out = e->xres; a.st = e->st;
        if (li == 0) { a.emb_w = (const uint8_t*)e->embed->data; a.emb_s = e->embed->scales; a.tokens = e->tokens; a.emb_dt = wd == GTB_Q8 ? DT_Q8 : DT_Q4; }
        else { a.src0 = e->hres; a.src1 = e->rd; }
        gv.push_back(a);
        FdAttnArgs t{};
        t.rqkv = e->rqkv; t.n_embd = E; t.kv_dim = KV; t.gsz = e->gsz; t.kq = l.kq; t.ks = l.ks; t.vq = l.vq; t.vs = l.vs;
        t.rope_cos = e->rope_cos; t.rope_sin = e->rope_sin; t.st = e->st; t.parts = e->fd_parts; t.counters = e->fd_cnt;
        t.out = e->fd_attn; t.n_heads = c.n_heads; t.n_groups = c.n_groups; t.chunk_len = fd_chunk_len(c.max_ctx, e->fd_chunk);
        av.push_back(t);
        FdArgs o{};
        o.K = E; o.n_rows = E; o.w = (const uint4*)l.o->data; o.ws = l.o->scales; o.out = e->ro; o.in = e->fd_attn;
        gv.push_back(o);
        FdArgs g{};
        g.K = E; g.n_rows = 2 * F; g.w = (const uint4*)l.gu_data; g.ws = l.gu_sc; g.n_ffn = F; g.act_out = e->fd_act;
        g.src0 = e->xres; g.src1 = e->ro; g.normw = l.ffn_norm; g.res_out = e->hres;
        gv.push_back(g);
        FdArgs d{};
        d.K = F; d.n_rows = E; d.w = (const uint4*)l.down->data; d.ws = l.down->scales; d.out = e->rd; d.in = e->fd_act;
        gv.push_back(d);
    }
    if (with_head) {
        FdArgs hd{};
        hd.K = E; hd.n_rows = c.n_vocab; hd.w = (const uint4*)e->lm_head->data; hd.ws = e->lm_head->scales; hd.out = e->logits;
        hd.src0 = e->hres; hd.src1 = e->rd; hd.normw = e->final_norm; hd.res_out = e->xfinal;
        hd.arg_val = e->fd_argv; hd.arg_idx = e->fd_argi; hd.counter = e->fd_cnt + c.n_heads; hd.tok_out = e->tokens; hd.st = e->st; hd.eos_id = eos_id;
        gv.push_back(hd);
    }
    // L2 look-ahead: phase i pushes the weights of GEMV phase i + fd_ahead (cyclic: the next row starts over) towards L2
    const int ns = (int)gv.size();
    for (int i = 0; i < ns && e->fd_ahead > 0; i++) {
        const FdArgs& t = gv[(i + e->fd_ahead) % ns];
        gv[i].pf[0] = t.w; gv[i].pf_bytes[0] = weight_data_bytes(wd, t.n_rows, t.K);
        gv[i].pf[1] = t.ws; gv[i].pf_bytes[1] = weight_scale_bytes(wd, t.n_rows, t.K);
    }
    return GTB_OK;
}

// one row through the order-free kernels: 5 launches per layer (+ the head), chained with programmatic dependent launch
template <int WT>
int enqueue_row_fast(gtb_engine* e, bool with_head, int eos_id) {
    const gtb_model_config& c = e->cfg;
    const int L = c.n_layers, G = ctx().sm_count;
    std::vector<FdArgs> gv;
    std::vector<FdAttnArgs> av;
    int r = build_fast_args(e, with_head, eos_id, gv, av);
    if (r) return r;
    static bool attn_attr = false;
    if (!attn_attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_fd_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attn_attr = true;
    }
    const size_t attn_smem = fd_attn_smem(fd_chunk_len(c.max_ctx, e->fd_chunk));
    if (attn_smem > 64 * 1024) return fail(GTB_ERR_STATE, "fast_decode: max_ctx too large for the attention kernel's score buffer");
    for (int li = 0; li < L; li++) {
        if ((r = launch_fd<WT, FD_NORM, FD_RAW>(gv[4 * li + 0], G))) return r;
        GTB_CUDA(fd_launch(k_fd_attn, c.n_groups * FD_MAXCH, attn_smem, av[li]));
        GTB_LAUNCHED();
        if ((r = launch_fd<WT, FD_CODES, FD_RAW>(gv[4 * li + 1], G))) return r;
        if ((r = launch_fd<WT, FD_NORM, FD_SILU>(gv[4 * li + 2], c.n_ffn / 32))) return r;
        if ((r = launch_fd<WT, FD_CODES, FD_RAW>(gv[4 * li + 3], G))) return r;
    }
    if (with_head) {
        if ((r = launch_fd<WT, FD_NORM, FD_ARGMAX>(gv[4 * L], 2 * G))) return r;
    } else {
        k_advance<<<1, 1, 0, ctx().stream>>>(e->st);
        GTB_LAUNCHED();
    }
    return GTB_OK;
}

// the same phases as one persistent cooperative launch for all rows (k_fd_mega)
bool fast_mega_ok(const gtb_engine* e) {
    const gtb_model_config& c = e->cfg;
    return e->fast && e->fd_mega && !e->capture && c.wdtype != GTB_F16 && c.n_embd <= 2048 && c.n_ffn <= 6144 &&
           fd_mega_smem(c.n_embd, c.n_ffn, c.max_ctx) <= 100 * 1024;
}

template <int WT>
int launch_fast_mega(gtb_engine* e, FdMegaParams& p) {
    const gtb_model_config& c = e->cfg;
    const size_t smem = fd_mega_smem(c.n_embd, c.n_ffn, c.max_ctx);
    static bool attr_done = false;
    if (!attr_done) {
        GTB_CUDA(cudaFuncSetAttribute(k_fd_mega<WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        int nb = 0;
        GTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fd_mega<WT>, FD_NT, smem));
        if (nb < 1) return fail(GTB_ERR_CUDA, "k_fd_mega does not fit on an SM");
        attr_done = true;
    }
    void* args[] = {&p};
    GTB_CUDA(cudaLaunchCooperativeKernel((const void*)k_fd_mega<WT>, dim3(ctx().sm_count), dim3(FD_NT), args, smem, ctx().stream));
    GTB_LAUNCHED();
    return GTB_OK;
}

int run_rows_fast_mega(gtb_engine* e, int n_body, int n_head, int eos_id) {
    const gtb_model_config& c = e->cfg;
    const bool with_head = n_head > 0;
    if (!e->fd_args_valid || e->fd_args_head != with_head || e->fd_args_eos != eos_id) {
        std::vector<FdArgs> gv;
        std::vector<FdAttnArgs> av;
        int r = build_fast_args(e, with_head, eos_id, gv, av);
        if (r) return r;
        GTB_CUDA(cudaMemcpyAsync(e->d_fd_gemv, gv.data(), gv.size() * sizeof(FdArgs), cudaMemcpyHostToDevice, ctx().stream));
        GTB_CUDA(cudaMemcpyAsync(e->d_fd_attn, av.data(), av.size() * sizeof(FdAttnArgs), cudaMemcpyHostToDevice, ctx().stream));
        GTB_CUDA(cudaStreamSynchronize(ctx().stream));          // the vectors are locals
        e->fd_args_valid = true; e->fd_args_head = with_head; e->fd_args_eos = eos_id;
    }
    GTB_CUDA(cudaMemsetAsync(e->fd_bar, 0, 128, ctx().stream));
    FdMegaParams p{};
    p.gemv = e->d_fd_gemv; p.attn = e->d_fd_attn; p.n_layers = c.n_layers; p.n_heads = c.n_heads;
    p.n_body = n_body; p.n_head = n_head; p.bar = e->fd_bar; p.bar_base = 0; p.st = e->st;
    p.prof = e->prof ? e->d_prof : nullptr; p.prof_cta = e->fd_prof_cta;
    return (c.wdtype == GTB_Q8) ? launch_fast_mega<DT_Q8>(e, p) : launch_fast_mega<DT_Q4>(e, p);
}

// ---------------------------------------------------------------- batched order-free decode
void batch_free(gtb_engine* e) {
    auto& b = e->bb;
    if (b.graph) { cudaGraphExecDestroy(b.graph); b.graph = nullptr; }
    void* p[] = {b.rqkv, b.ro, b.rd, b.xres, b.hres, b.xfinal, b.logits, b.parts, b.argv, b.argi, b.cnt, b.tokens, b.st,
                 b.norm_act.codes, b.norm_act.ad, b.norm_act.n7, b.attn_act.codes, b.attn_act.ad, b.attn_act.n7,
                 b.mlp_act.codes, b.mlp_act.ad, b.mlp_act.n7};
    for (void* q : p) cudaFree(q);
    for (auto q : b.kq) cudaFree(q);
    for (auto q : b.vq) cudaFree(q);
    for (auto q : b.ks) cudaFree(q);
    for (auto q : b.vs) cudaFree(q);
    b = gtb_engine::BatchBufs{};
}

int batch_alloc(gtb_engine* e, int n) {
    batch_free(e);
    if (n <= 0) return GTB_OK;
    const gtb_model_config& c = e->cfg;
    const int E = c.n_embd, F = c.n_ffn, KV = e->kv_dim, MC = c.max_ctx;
    auto& b = e->bb;
    auto dalloc = [&](void** p, size_t bytes) -> int { GTB_CUDA(cudaMalloc(p, bytes)); GTB_CUDA(cudaMemsetAsync(*p, 0, bytes, ctx().stream)); ctx().mem += (int64_t)bytes; return GTB_OK; };
    int r = 0;
    const size_t N = (size_t)n;
    r |= dalloc((void**)&b.rqkv, N * (E + 2 * KV) * 4); r |= dalloc((void**)&b.ro, N * E * 4); r |= dalloc((void**)&b.rd, N * E * 4);
    r |= dalloc((void**)&b.xres, N * E * 4); r |= dalloc((void**)&b.hres, N * E * 4); r |= dalloc((void**)&b.xfinal, N * E * 4);
    r |= dalloc((void**)&b.logits, N * c.n_vocab * 4);
    r |= dalloc((void**)&b.parts, N * c.n_heads * FD_MAXCH * FD_PART * 4);
    r |= dalloc((void**)&b.argv, N * 1024 * 4); r |= dalloc((void**)&b.argi, N * 1024 * 4);
    r |= dalloc((void**)&b.cnt, (N * (c.n_heads + 1) + 1) * 4);
    r |= dalloc((void**)&b.tokens, N * (MC + 2) * 4); r |= dalloc((void**)&b.st, N * sizeof(DevState));
    FdAct* acts[3] = {&b.norm_act, &b.attn_act, &b.mlp_act};
    const int widths[3] = {E, E, F};
    for (int i = 0; i < 3; i++) {
        r |= dalloc((void**)&acts[i]->codes, N * widths[i]); r |= dalloc((void**)&acts[i]->ad, N * (widths[i] / 32) * 4);
        r |= dalloc((void**)&acts[i]->n7, N * (widths[i] / 32) * 4);
    }
    b.kq.assign(c.n_layers, nullptr); b.vq.assign(c.n_layers, nullptr); b.ks.assign(c.n_layers, nullptr); b.vs.assign(c.n_layers, nullptr);
    for (int li = 0; li < c.n_layers; li++) {
        const size_t esz = (e->adtype == GTB_F16) ? 2 : 1;           // fp16 K/V rows of an FP16 model, Q8 codes otherwise
        r |= dalloc((void**)&b.kq[li], N * MC * KV * esz); r |= dalloc((void**)&b.vq[li], N * MC * KV * esz);
        r |= dalloc((void**)&b.ks[li], N * MC * (KV / 32) * 2); r |= dalloc((void**)&b.vs[li], N * MC * (KV / 32) * 2);
    }
    if (r) { batch_free(e); return fail(GTB_ERR_CUDA, "batch buffers: allocation failed"); }
    b.n = n;
    b.host_pos.assign(n, 0);
    return GTB_OK;
}

template <int WT, int EPI, int NBL, int R, int NSM>
int launch_fdb_k(const FdArgs& a, int grid) {
    const size_t smem = fdb_gemv_smem(a.K, a.n_seq);
    static bool done = false;
    if (!done) { GTB_CUDA(cudaFuncSetAttribute(k_fdb_gemv<WT, EPI, NBL, R, NSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (NSM > 8 ? 128 : 64) * 1024)); done = true; }
    GTB_CUDA(fd_launch(k_fdb_gemv<WT, EPI, NBL, R, NSM>, grid, smem, a));
    GTB_LAUNCHED();
    return GTB_OK;
}
// rows per warp pass x sequence slots = 32 sums per lane at most: up to 8 sequences R = 4 (2 for long rows / Q8), up to 16 half of that
template <int WT, int EPI>
int launch_fdb(const FdArgs& a, int grid) {
    const bool wide = (a.K / 32 + 31) / 32 > 2, big = a.n_seq > 8;
    constexpr int R2 = (WT == DT_Q4) ? 4 : 2, R6 = (WT == DT_Q4) ? 2 : 1;
    if (!big) return wide ? launch_fdb_k<WT, EPI, 6, R6, 8>(a, grid) : launch_fdb_k<WT, EPI, 2, R2, 8>(a, grid);
    return wide ? launch_fdb_k<WT, EPI, 6, 1, 16>(a, grid) : launch_fdb_k<WT, EPI, 2, 2, 16>(a, grid);
}

// one row of every sequence of the batch: 7 launches per layer + 2 for the head
template <int WT>
int enqueue_batch_row(gtb_engine* e) {
    const gtb_model_config& c = e->cfg;
    auto& b = e->bb;
    const int E = c.n_embd, F = c.n_ffn, KV = e->kv_dim, L = c.n_layers, G = ctx().sm_count, NS = b.n, MC = c.max_ctx;
    const int wd = c.wdtype;
    std::vector<FdArgs> gv;          // GEMV steps in launch order (4 per layer + head), for the L2 look-ahead
    for (int li = 0; li < L; li++) {
        LayerW& l = e->L[li];
        FdArgs q{}; q.K = E; q.n_rows = E + 2 * KV; q.w = (const uint4*)l.qkv_data; q.ws = l.qkv_sc; q.out = b.rqkv; q.in = b.norm_act;
        FdArgs o{}; o.K = E; o.n_rows = E; o.w = (const uint4*)l.o->data; o.ws = l.o->scales; o.out = b.ro; o.in = b.attn_act;
        FdArgs g{}; g.K = E; g.n_rows = 2 * F; g.w = (const uint4*)l.gu_data; g.ws = l.gu_sc; g.n_ffn = F; g.in = b.norm_act; g.act_out = b.mlp_act;
        FdArgs d{}; d.K = F; d.n_rows = E; d.w = (const uint4*)l.down->data; d.ws = l.down->scales; d.out = b.rd; d.in = b.mlp_act;
        gv.push_back(q); gv.push_back(o); gv.push_back(g); gv.push_back(d);
    }
    {
        FdArgs hd{}; hd.K = E; hd.n_rows = c.n_vocab; hd.w = (const uint4*)e->lm_head->data; hd.ws = e->lm_head->scales; hd.out = b.logits;
        hd.in = b.norm_act; hd.arg_val = b.argv; hd.arg_idx = b.argi; hd.counter = b.cnt + (size_t)NS * (c.n_heads + 1);
        hd.tok_out = b.tokens; hd.st = b.st; hd.eos_id = -1;
        gv.push_back(hd);
    }
    const int ns = (int)gv.size();
    for (int i = 0; i < ns; i++) {
        gv[i].n_seq = NS; gv[i].tok_stride = MC + 2;
        if (e->fd_ahead > 0) {
            const FdArgs& t = gv[(i + e->fd_ahead) % ns];
            gv[i].pf[0] = t.w; gv[i].pf_bytes[0] = weight_data_bytes(wd, t.n_rows, t.K);
            gv[i].pf[1] = t.ws; gv[i].pf_bytes[1] = weight_scale_bytes(wd, t.n_rows, t.K);
        }
    }
    static bool attr = false;
    if (!attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_fdb_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        GTB_CUDA(cudaFuncSetAttribute(k_fdb_norm, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    const size_t attn_smem = fd_attn_smem(fd_chunk_len(MC, 128));
    if (attn_smem > 64 * 1024) return fail(GTB_ERR_STATE, "batched decode: max_ctx too large for the attention kernel's score buffer");
    const size_t norm_smem = fd_gemv_smem(E, true);
    auto norm = [&](const float* src0, const float* src1, const uint16_t* w, float* res_out, bool embed) -> int {
        FdArgs a{};
        a.K = E; a.src0 = src0; a.src1 = src1; a.normw = w; a.res_out = res_out; a.st = b.st; a.act_out = b.norm_act;
        a.n_seq = NS; a.tok_stride = MC + 2;
        if (embed) { a.emb_w = (const uint8_t*)e->embed->data; a.emb_s = e->embed->scales; a.tokens = b.tokens; a.emb_dt = wd == GTB_Q8 ? DT_Q8 : DT_Q4; }
        GTB_CUDA(fd_launch(k_fdb_norm, NS, norm_smem, a));
        GTB_LAUNCHED();
        return GTB_OK;
    };
    int r;
    for (int li = 0; li < L; li++) {
        LayerW& l = e->L[li];
        if ((r = norm(li ? b.hres : nullptr, li ? b.rd : nullptr, l.attn_norm, b.xres, li == 0))) return r;
        if ((r = launch_fdb<WT, FD_RAW>(gv[4 * li + 0], G))) return r;
        FdAttnArgs t{};
        t.rqkv = b.rqkv; t.n_embd = E; t.kv_dim = KV; t.gsz = e->gsz; t.kq = b.kq[li]; t.ks = b.ks[li]; t.vq = b.vq[li]; t.vs = b.vs[li];
        t.rope_cos = e->rope_cos; t.rope_sin = e->rope_sin; t.st = b.st; t.parts = b.parts; t.counters = b.cnt; t.out = b.attn_act;
        t.seq_kv_codes = (size_t)MC * KV; t.seq_kv_scales = (size_t)MC * (KV / 32); t.n_heads = c.n_heads; t.n_groups = c.n_groups;
        t.chunk_len = fd_chunk_len(MC, 128);
        {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(c.n_groups * FD_MAXCH, NS); cfg.blockDim = dim3(FD_NT); cfg.dynamicSmemBytes = attn_smem; cfg.stream = ctx().stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            GTB_CUDA(cudaLaunchKernelEx(&cfg, k_fdb_attn, t));
            GTB_LAUNCHED();
        }
        if ((r = launch_fdb<WT, FD_RAW>(gv[4 * li + 1], G))) return r;
        if ((r = norm(b.xres, b.ro, l.ffn_norm, b.hres, false))) return r;
        if ((r = launch_fdb<WT, FD_SILU>(gv[4 * li + 2], F / 32))) return r;
        if ((r = launch_fdb<WT, FD_RAW>(gv[4 * li + 3], G))) return r;
    }
    if ((r = norm(b.hres, b.rd, e->final_norm, b.xfinal, false))) return r;
    return launch_fdb<WT, FD_ARGMAX>(gv[4 * L], 2 * G);
}

int enqueue_row_dt(gtb_engine* e, bool with_head, int eos_id) {
    if (e->fast) {
        if (e->capture) return fail(GTB_ERR_STATE, "capture_acv records the order-exact kernels; switch fast_decode off");
        if (e->cfg.wdtype == GTB_Q8) return enqueue_row_fast<DT_Q8>(e, with_head, eos_id);
        if (e->cfg.wdtype == GTB_Q4) return enqueue_row_fast<DT_Q4>(e, with_head, eos_id);
        return fail(GTB_ERR_STATE, "fast_decode is built for Q8-activation models (Q8, Q4 weights)");
    }
    switch (e->cfg.wdtype) {
        case GTB_F16: return enqueue_row<DT_F16>(e, with_head, eos_id);
        case GTB_Q8: return enqueue_row<DT_Q8>(e, with_head, eos_id);
        default: return enqueue_row<DT_Q4>(e, with_head, eos_id);
    }
}

void drop_graphs(gtb_engine* e) {
    e->fd_args_valid = false;
    if (e->bb.graph) { cudaGraphExecDestroy(e->bb.graph); e->bb.graph = nullptr; }
    if (e->g_body) { cudaGraphExecDestroy(e->g_body); e->g_body = nullptr; }
    if (e->g_head) { cudaGraphExecDestroy(e->g_head); e->g_head = nullptr; }
}

int build_graph(gtb_engine* e, bool with_head, int eos_id, cudaGraphExec_t* out, int* n_launches) {
    cudaStream_t st = ctx().stream;
    // make sure every cudaFuncSetAttribute happened outside capture: run one row eagerly on scratch state? No --
    // attributes are set lazily inside launch_phase, which is legal during capture (not a stream operation).
    const int64_t before = ctx().launches;
    GTB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int r = enqueue_row_dt(e, with_head, eos_id);
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamEndCapture(st, &g);
    *n_launches = (int)(ctx().launches - before);
    ctx().launches = before;
    if (r) { if (g) cudaGraphDestroy(g); return r; }
    if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
    return GTB_OK;
}

constexpr int PROF_SLOTS = 4096;

bool mega_ok(const gtb_engine* e) {
    const gtb_model_config& c = e->cfg;
    // the persistent kernel is specialised for the TinyLlama dimensions (tinyllama.cpp:12-20); anything else takes
    // the one-kernel-per-phase path
    return e->use_mega && !e->fast && !e->capture && c.n_embd == ME && c.n_ffn == MF && c.n_heads == MH && c.n_groups * MGSZ == MH &&
           e->grid >= MH * 4 && e->grid <= 1024 && c.max_ctx <= 4 * MT && c.n_layers <= MEGA_MAX_LAYERS && attn_scratch_bytes(c.max_ctx) <= (size_t)PS_BYTES;
}

template <int WT, bool PROF>
int launch_mega_v(gtb_engine* e, MegaParams& p) {
    constexpr int AT = (WT == DT_F16) ? DT_F16 : DT_Q8;
    const size_t smem = mega_smem_bytes(AT);
    static bool attr_done = false;
    if (!attr_done) {
        GTB_CUDA(cudaFuncSetAttribute(k_mega<WT, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        GTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_mega<WT, PROF>, MT, smem));
        if (nb < 1) return fail(GTB_ERR_CUDA, "megakernel does not fit on an SM (%zu B shared memory)", smem);
        attr_done = true;
    }
    if (e->grid > ctx().sm_count) return fail(GTB_ERR_ARG, "cooperative grid (%d) exceeds the SM count (%d)", e->grid, ctx().sm_count);
    void* args[] = {&p};
    GTB_CUDA(cudaLaunchCooperativeKernel((const void*)k_mega<WT, PROF>, dim3(e->grid), dim3(MT), args, smem, ctx().stream));
    GTB_LAUNCHED();
    return GTB_OK;
}
template <int WT>
int launch_mega(gtb_engine* e, MegaParams& p) {
    return p.prof ? launch_mega_v<WT, true>(e, p) : launch_mega_v<WT, false>(e, p);     // the stamped build only under option "prof"
}

int run_rows_mega(gtb_engine* e, int n_body, int n_head, int eos_id, int start_pos) {
    const gtb_model_config& c = e->cfg;
    if (e->adtype == GTB_Q8) {
        // rows [kvt_pos, start_pos) were appended by other kernels (multi-row prefill, tensor-core prefill, ...): bring the copies up to date
        if (e->kvt_pos < start_pos) {
            const int n = start_pos - e->kvt_pos, cpr = c.n_groups * 4;
            for (auto& l : e->L) {
                k_kv_transpose<<<(unsigned)(((size_t)n * cpr + 255) / 256), 256, 0, ctx().stream>>>(l.kq, l.vq, l.ks, l.vs, l.kqt, l.vqt, l.kst, l.vst, e->kvt_pos, start_pos, c.max_ctx, c.n_groups);
                GTB_LAUNCHED();
            }
        }
        e->kvt_pos = start_pos + n_body + n_head;      // this launch maintains them for the rows it appends
    }
    // The exchange words carry a 32-bit epoch tag that advances < 200 per row; long before it could wrap (tag 0 = "never
    // written") the buffers and the epoch are cleared.  ~10^7 rows apart: hours of decoding.
    e->mega_rows += n_body + n_head;
    if (e->mega_rows > (1ll << 31) / 256) {
        cudaStream_t st = ctx().stream;
        const size_t E = c.n_embd, F = c.n_ffn, KV = e->kv_dim;
        GTB_CUDA(cudaStreamSynchronize(st));
        GTB_CUDA(cudaMemsetAsync(e->x_qkv, 0, (E + 2 * KV) * 8, st)); GTB_CUDA(cudaMemsetAsync(e->x_sc, 0, (size_t)c.n_heads * e->sc_stride * 8, st));
        GTB_CUDA(cudaMemsetAsync(e->x_attn, 0, E * 8, st)); GTB_CUDA(cudaMemsetAsync(e->x_o, 0, E * 8, st)); GTB_CUDA(cudaMemsetAsync(e->x_gu, 0, 2 * F * 8, st));
        GTB_CUDA(cudaMemsetAsync(e->x_act, 0, (F / 32) * 16 * 8, st)); GTB_CUDA(cudaMemsetAsync(e->x_down, 0, E * 8, st));
        GTB_CUDA(cudaMemsetAsync(e->x_arg, 0, (size_t)2 * 1024 * 8, st)); GTB_CUDA(cudaMemsetAsync(e->epoch, 0, 16, st));
        e->mega_rows = n_body + n_head;
    }
    if (!e->layers_valid) {
        std::vector<MegaLayer> h(c.n_layers);
        for (int i = 0; i < c.n_layers; i++) {
            LayerW& l = e->L[i];
            h[i].w[0] = (const uint4*)l.qkv_data; h[i].s[0] = l.qkv_sc;
            h[i].w[1] = (const uint4*)l.o->data; h[i].s[1] = l.o->scales;
            h[i].w[2] = (const uint4*)l.gu_data; h[i].s[2] = l.gu_sc;
            h[i].w[3] = (const uint4*)l.down->data; h[i].s[3] = l.down->scales;
            h[i].attn_norm = l.attn_norm; h[i].ffn_norm = l.ffn_norm;
            h[i].kq = l.kq; h[i].ks = l.ks; h[i].vq = l.vq; h[i].vs = l.vs; h[i].kqt = l.kqt; h[i].vqt = l.vqt; h[i].kst = l.kst; h[i].vst = l.vst;
        }
        GTB_CUDA(cudaMemcpyAsync(e->d_layers, h.data(), h.size() * sizeof(MegaLayer), cudaMemcpyHostToDevice, ctx().stream));
        GTB_CUDA(cudaStreamSynchronize(ctx().stream));
        e->layers_valid = true;
    }
    MegaParams p{};
    p.n_layers = c.n_layers; p.n_vocab = c.n_vocab; p.max_ctx = c.max_ctx; p.sc_stride = e->sc_stride;
    p.layers = e->d_layers;
    p.emb_w = e->embed->data; p.emb_s = e->embed->scales; p.head_w = (const uint4*)e->lm_head->data; p.head_s = e->lm_head->scales;
    p.final_norm = e->final_norm; p.rope_cos = e->rope_cos; p.rope_sin = e->rope_sin;
    p.x_qkv = e->x_qkv; p.x_sc = e->x_sc; p.x_attn = e->x_attn; p.x_o = e->x_o; p.x_gu = e->x_gu; p.x_act = e->x_act;
    p.x_down = e->x_down; p.x_arg = e->x_arg; p.cnt = e->cnt;
    p.logits = e->logits; p.tokens = e->tokens; p.st = e->st; p.epoch = e->epoch;
    p.n_body = n_body; p.n_head = n_head; p.eos_id = eos_id; p.pf_ahead = e->pf_ahead;
    p.dbg = e->dbg; p.prof = e->prof ? e->d_prof : nullptr; p.prof_cta = e->prof_cta;
    switch (c.wdtype) {
        case GTB_F16: return launch_mega<DT_F16>(e, p);
        case GTB_Q8: return launch_mega<DT_Q8>(e, p);
        default: return launch_mega<DT_Q4>(e, p);
    }
}

// ---- order-exact multi-row path (gtb_xrows.cu)
bool xr_ok(const gtb_engine* e) {
    return e->use_xr && !e->fast && !e->capture && xr_supported(e->cfg, e->gsz);
}

// host-side view of the weights / caches for the multi-row kernels
int xr_model(gtb_engine* e, XrModel& m) {
    const gtb_model_config& c = e->cfg;
    if (!e->xr) { int r = xr_create(&e->xr, c); if (r) return r; }
    e->xr_layers.resize(c.n_layers);
    e->xr_kq.resize(c.n_layers); e->xr_vq.resize(c.n_layers); e->xr_ks.resize(c.n_layers); e->xr_vs.resize(c.n_layers);
    for (int i = 0; i < c.n_layers; i++) {
        LayerW& l = e->L[i];
        XrLayerW& x = e->xr_layers[i];
        x.w[0] = (const uint4*)l.qkv_data; x.s[0] = l.qkv_sc;
        x.w[1] = (const uint4*)l.o->data; x.s[1] = l.o->scales;
        x.w[2] = (const uint4*)l.gu_data; x.s[2] = l.gu_sc;
        x.w[3] = (const uint4*)l.down->data; x.s[3] = l.down->scales;
        x.attn_norm = l.attn_norm; x.ffn_norm = l.ffn_norm;
        e->xr_kq[i] = l.kq; e->xr_vq[i] = l.vq; e->xr_ks[i] = l.ks; e->xr_vs[i] = l.vs;
    }
    m.cfg = c;
    m.emb_w = e->embed->data; m.emb_s = e->embed->scales;
    m.head_w = (const uint4*)e->lm_head->data; m.head_s = e->lm_head->scales;
    m.final_norm = e->final_norm; m.rope_cos = e->rope_cos; m.rope_sin = e->rope_sin;
    m.layers = e->xr_layers.data();
    return GTB_OK;
}

// rows [p0, p0 + n_rows) of the engine's own sequence in passes of up to xr_rows rows; the last row samples if with_head
int run_rows_xr(gtb_engine* e, int p0, int n_rows, int n_ctx, bool with_head, int eos_id) {
    XrModel m{};
    int r = xr_model(e, m);
    if (r) return r;
    XrKV kv{e->xr_kq.data(), e->xr_ks.data(), e->xr_vq.data(), e->xr_vs.data(), 0, 0};
    XrSeq sq{e->tokens, 0, e->st};
    for (int done = 0; done < n_rows;) {
        const int n = (n_rows - done < e->xr_rows) ? n_rows - done : e->xr_rows;
        const bool last = done + n == n_rows;
        r = xr_prefill_pass(e->xr, m, kv, sq, 0, p0 + done, n, n_ctx, with_head && last, eos_id, e->logits);
        if (r) return r;
        done += n;
    }
    return GTB_OK;
}

int set_pos(gtb_engine* e, int pos);

// rows to run: `n_body` rows without lm_head, then `n_head` rows with lm_head + argmax.  p0 >= 0: the position of the first
// row and the call's n_ctx are known on the host, so runs of rows can go through the exact multi-row kernels.
int run_rows(gtb_engine* e, int n_body, int n_head, int eos_id, int p0 = -1, int n_ctx = 0) {
    cudaStream_t st = ctx().stream;
    if (n_body + n_head <= 0) return GTB_OK;
    int start = (p0 >= 0) ? p0 : e->host_pos;              // first row of this call
    if (start < e->kvt_pos) e->kvt_pos = start;            // rows >= start are rewritten: k_mega's K/V copies end there
    if (p0 >= 0 && xr_ok(e) && n_body + (n_head > 0 ? 1 : 0) >= e->xr_min_rows) {
        // the prompt rows (and the first sampling row) side by side, bit-identical to the row-at-a-time kernels
        const int nx = n_body + (n_head > 0 ? 1 : 0);
        int r = run_rows_xr(e, p0, nx, n_ctx, n_head > 0, eos_id);
        if (r) return r;
        if (n_head == 0) return set_pos(e, p0 + n_body);       // no sampling row: only the position moves
        n_body = 0; n_head -= 1;
        start += nx;
        if (n_head == 0) return GTB_OK;
    }
    if (mega_ok(e)) return run_rows_mega(e, n_body, n_head, eos_id, start);
    if (fast_mega_ok(e)) return run_rows_fast_mega(e, n_body, n_head, eos_id);
    if (e->use_graph && !e->capture) {
        if (n_body > 0 && !e->g_body) { int r = build_graph(e, false, -1, &e->g_body, &e->launches_body); if (r) return r; }
        if (n_head > 0 && (!e->g_head || e->g_eos != eos_id)) {
            if (e->g_head) { cudaGraphExecDestroy(e->g_head); e->g_head = nullptr; }
            int r = build_graph(e, true, eos_id, &e->g_head, &e->launches_head);
            if (r) return r;
            e->g_eos = eos_id;
        }
        for (int i = 0; i < n_body; i++) { GTB_CUDA(cudaGraphLaunch(e->g_body, st)); ctx().launches += e->launches_body; }
        for (int i = 0; i < n_head; i++) { GTB_CUDA(cudaGraphLaunch(e->g_head, st)); ctx().launches += e->launches_head; }
    } else {
        for (int i = 0; i < n_body; i++) { int r = enqueue_row_dt(e, false, -1); if (r) return r; }
        for (int i = 0; i < n_head; i++) { int r = enqueue_row_dt(e, true, eos_id); if (r) return r; }
    }
    return GTB_OK;
}

int set_pos(gtb_engine* e, int pos) {
    k_set_pos<<<1, 1, 0, ctx().stream>>>(e->st, pos);
    GTB_LAUNCHED();
    return GTB_OK;
}

int set_state(gtb_engine* e, int pos, int nctx_min) {
    DevState s{pos, nctx_min, 0, 0};
    GTB_CUDA(cudaMemcpyAsync(e->st, &s, sizeof s, cudaMemcpyHostToDevice, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));      // `s` is a stack object
    return GTB_OK;
}

size_t expect_bytes(const gtb_engine* e, int tid, int* rows, int* cols, int* dt) {
    const gtb_model_config& c = e->cfg;
    int r = 0, k = c.n_embd, d = c.wdtype;
    switch (tid) {
        case GTB_T_EMBED: case GTB_T_LM_HEAD: r = c.n_vocab; break;
        case GTB_T_Q: case GTB_T_O: r = c.n_embd; break;
        case GTB_T_K: case GTB_T_V: r = e->kv_dim; break;
        case GTB_T_GATE: case GTB_T_UP: r = c.n_ffn; break;
        case GTB_T_DOWN: r = c.n_embd; k = c.n_ffn; break;
        case GTB_T_FINAL_NORM: case GTB_T_ATTN_NORM: case GTB_T_FFN_NORM: r = 1; d = GTB_F16; break;   // norm weights are fp16 (modules.cpp:84)
        default: return 0;
    }
    *rows = r; *cols = k; *dt = d;
    return (size_t)r * row_nbytes(d, k);
}

}  // namespace

extern "C" {

int gtb_engine_create(gtb_engine_t* out, const gtb_model_config* cfg) {
    GTB_CHECK_INIT();
    GTB_ARG(out && cfg);
    GTB_ARG(cfg->wdtype == GTB_F16 || cfg->wdtype == GTB_Q8 || cfg->wdtype == GTB_Q4);
    GTB_ARG(cfg->n_embd % 64 == 0 && cfg->n_ffn % 64 == 0 && cfg->n_heads > 0 && cfg->n_groups > 0);
    GTB_ARG(cfg->n_embd / cfg->n_heads == 64);               // d_head is a compile-time constant of the attention kernel
    GTB_ARG(cfg->n_heads % cfg->n_groups == 0 && cfg->max_ctx > 0 && cfg->n_layers > 0 && cfg->n_vocab > 0);
    GTB_ARG(cfg->n_embd <= NT * ES_EPT * 4);
    auto* e = new gtb_engine();
    e->cfg = *cfg;
    e->adtype = (cfg->wdtype == GTB_F16) ? GTB_F16 : GTB_Q8;  // tinyllama.cpp:258-265
    e->d_head = 64;
    e->kv_dim = 64 * cfg->n_groups;
    e->gsz = cfg->n_heads / cfg->n_groups;
    e->grid = ctx().sm_count;
    e->L.resize(cfg->n_layers);
    const int E = cfg->n_embd, F = cfg->n_ffn, KV = e->kv_dim, MC = cfg->max_ctx;
    int first_err = 0;
    auto dalloc = [&](void** p, size_t n) -> int {
        if (first_err) return first_err;          // after the first failure nothing else is attempted
        cudaError_t ce = cudaMalloc(p, n);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(*p, 0, n, ctx().stream);
        if (ce != cudaSuccess) { first_err = fail(GTB_ERR_CUDA, "engine buffers: %s", cudaGetErrorString(ce)); return first_err; }
        ctx().mem += (int64_t)n;
        return GTB_OK;
    };
    int r = 0;
    for (auto& l : e->L) {
        const size_t code_bytes = (size_t)MC * KV * (e->adtype == GTB_F16 ? 2 : 1);
        r |= dalloc((void**)&l.kq, code_bytes); r |= dalloc((void**)&l.vq, code_bytes);
        if (e->adtype == GTB_Q8) {
            r |= dalloc((void**)&l.ks, (size_t)MC * (KV / 32) * 2); r |= dalloc((void**)&l.vs, (size_t)MC * (KV / 32) * 2);
            r |= dalloc((void**)&l.kqt, code_bytes); r |= dalloc((void**)&l.vqt, code_bytes);
            r |= dalloc((void**)&l.kst, (size_t)MC * (KV / 32) * 2); r |= dalloc((void**)&l.vst, (size_t)MC * (KV / 32) * 2);
        }
        r |= dalloc((void**)&l.attn_norm, (size_t)E * 2); r |= dalloc((void**)&l.ffn_norm, (size_t)E * 2);
        r |= dalloc((void**)&l.qkv_data, weight_data_bytes(cfg->wdtype, E + 2 * KV, E));
        r |= dalloc((void**)&l.gu_data, weight_data_bytes(cfg->wdtype, 2 * F, E));
        if (cfg->wdtype != GTB_F16) {
            r |= dalloc((void**)&l.qkv_sc, weight_scale_bytes(cfg->wdtype, E + 2 * KV, E));
            r |= dalloc((void**)&l.gu_sc, weight_scale_bytes(cfg->wdtype, 2 * F, E));
        }
    }
    r |= dalloc((void**)&e->final_norm, (size_t)E * 2);
    r |= dalloc((void**)&e->xres, E * 4); r |= dalloc((void**)&e->hres, E * 4); r |= dalloc((void**)&e->xfinal, E * 4);
    r |= dalloc((void**)&e->rqkv, (size_t)(E + 2 * KV) * 4); r |= dalloc((void**)&e->rattn, E * 4); r |= dalloc((void**)&e->ro, E * 4);
    r |= dalloc((void**)&e->rg, F * 4); r |= dalloc((void**)&e->ru, F * 4); r |= dalloc((void**)&e->rd, E * 4);
    r |= dalloc((void**)&e->logits, (size_t)cfg->n_vocab * 4);
    r |= dalloc((void**)&e->fd_parts, (size_t)cfg->n_heads * FD_MAXCH * FD_PART * 4);
    r |= dalloc((void**)&e->fd_cnt, (size_t)(cfg->n_heads + 1) * 4);
    r |= dalloc((void**)&e->fd_attn.codes, E); r |= dalloc((void**)&e->fd_attn.ad, E / 32 * 4); r |= dalloc((void**)&e->fd_attn.n7, E / 32 * 4);
    r |= dalloc((void**)&e->fd_act.codes, F); r |= dalloc((void**)&e->fd_act.ad, F / 32 * 4); r |= dalloc((void**)&e->fd_act.n7, F / 32 * 4);
    r |= dalloc((void**)&e->fd_argv, 1024 * 4); r |= dalloc((void**)&e->fd_argi, 1024 * 4);
    r |= dalloc((void**)&e->d_fd_gemv, (size_t)(4 * cfg->n_layers + 1) * sizeof(FdArgs));
    r |= dalloc((void**)&e->d_fd_attn, (size_t)cfg->n_layers * sizeof(FdAttnArgs)); r |= dalloc((void**)&e->fd_bar, 128);
    r |= dalloc((void**)&e->tokens, (size_t)(MC + 2) * 4);
    r |= dalloc((void**)&e->st, sizeof(DevState));
    e->capw = (F > E) ? F : E;
    r |= dalloc((void**)&e->cap, ((size_t)cfg->n_layers * 12 + 2) * e->capw * 4);
    r |= dalloc((void**)&e->rope_cos, (size_t)MC * 32 * 4); r |= dalloc((void**)&e->rope_sin, (size_t)MC * 32 * 4);
    // exchange buffers of the persistent kernel: 64-bit {tag | payload} words, zero = never written (tags start at 1)
    e->sc_stride = (MC + 2) & ~1;
    r |= dalloc((void**)&e->x_qkv, (size_t)(E + 2 * KV) * 8); r |= dalloc((void**)&e->x_sc, (size_t)cfg->n_heads * e->sc_stride * 8);
    r |= dalloc((void**)&e->x_attn, (size_t)E * 8); r |= dalloc((void**)&e->x_o, (size_t)E * 8); r |= dalloc((void**)&e->x_gu, (size_t)2 * F * 8);
    r |= dalloc((void**)&e->x_act, (size_t)(F / 32) * 16 * 8); r |= dalloc((void**)&e->x_down, (size_t)E * 8);
    r |= dalloc((void**)&e->x_arg, (size_t)2 * 1024 * 8); r |= dalloc((void**)&e->dbg, 64); r |= dalloc((void**)&e->epoch, 16);
    r |= dalloc((void**)&e->cnt, (size_t)CNT_TOTAL * CNT_STRIDE * 4);
    r |= dalloc((void**)&e->d_prof, (size_t)PROF_SLOTS * 8); r |= dalloc((void**)&e->d_layers, (size_t)cfg->n_layers * sizeof(MegaLayer));
    if (r || first_err) { gtb_engine_destroy(e); return first_err ? first_err : GTB_ERR_CUDA; }
    {   // RoPE table with the reference's own expressions and libm (gten/ops.h:728-746; SURVEY §7 hard part 4)
        std::vector<float> cs((size_t)MC * 32), sn((size_t)MC * 32);
        const float d = 64.0f;
        for (int p = 0; p < MC; p++) {
            const float m = static_cast<float>(p);
            for (int j = 0; j < 32; j++) {
                const float m_theta_i = m * powf(10000.0f, -(2.0f * j / d));
                cs[(size_t)p * 32 + j] = cosf(m_theta_i);
                sn[(size_t)p * 32 + j] = sinf(m_theta_i);
            }
        }
        // on the library stream, i.e. AFTER the zero-fill of dalloc: a default-stream copy would overtake a memset that is still
        // queued behind other engines' work and the table would end up zeroed
        GTB_CUDA(cudaMemcpyAsync(e->rope_cos, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice, ctx().stream));
        GTB_CUDA(cudaMemcpyAsync(e->rope_sin, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice, ctx().stream));
        GTB_CUDA(cudaStreamSynchronize(ctx().stream));          // cs / sn are locals
    }
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    *out = e;
    return GTB_OK;
}

int gtb_engine_destroy(gtb_engine_t e) {
    if (!e) return GTB_OK;
    GTB_CHECK_INIT();
    cudaStreamSynchronize(ctx().stream);
    drop_graphs(e);
    for (auto& l : e->L) {
        gtb_weight_free(l.q); gtb_weight_free(l.k); gtb_weight_free(l.v); gtb_weight_free(l.o);
        gtb_weight_free(l.gate); gtb_weight_free(l.up); gtb_weight_free(l.down);
        cudaFree(l.attn_norm); cudaFree(l.ffn_norm); cudaFree(l.kq); cudaFree(l.vq); cudaFree(l.ks); cudaFree(l.vs); cudaFree(l.kqt); cudaFree(l.vqt); cudaFree(l.kst); cudaFree(l.vst);
        cudaFree(l.qkv_data); cudaFree(l.gu_data); cudaFree(l.qkv_sc); cudaFree(l.gu_sc);
    }
    gtb_weight_free(e->embed); gtb_weight_free(e->lm_head);
    void* bufs[] = {e->final_norm, e->xres, e->hres, e->xfinal, e->rqkv, e->rattn, e->ro, e->rg, e->ru, e->rd, e->logits,
                    e->tokens, e->st, e->cap, e->rope_cos, e->rope_sin, e->x_qkv, e->x_sc, e->x_attn, e->x_o, e->x_gu, e->x_act,
                    e->x_down, e->x_arg, e->dbg, e->epoch, e->cnt, e->d_prof, e->d_layers};
    for (void* b : bufs) cudaFree(b);
    if (e->pf) pf_destroy(e->pf);
    if (e->xr) xr_destroy(e->xr);
    cudaFree(e->pf_cap);
    cudaFree(e->fd_parts); cudaFree(e->fd_cnt); cudaFree(e->fd_argv); cudaFree(e->fd_argi);
    cudaFree(e->d_fd_gemv); cudaFree(e->d_fd_attn); cudaFree(e->fd_bar);
    batch_free(e);
    cudaFree(e->fd_attn.codes); cudaFree(e->fd_attn.ad); cudaFree(e->fd_attn.n7);
    cudaFree(e->fd_act.codes); cudaFree(e->fd_act.ad); cudaFree(e->fd_act.n7);
    delete e;
    return GTB_OK;
}

// src = host payload, or (on_device) a device buffer that holds it already: then nothing here synchronises and the caller
// keeps the buffer alive until the stream has passed the repack kernel (the pipelined loader below)
static int set_weight_from(gtb_engine_t e, int layer, int tensor_id, const void* h_payload, size_t nbytes, bool on_device) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_payload);
    int rows = 0, cols = 0, dt = 0;
    const size_t expect = expect_bytes(e, tensor_id, &rows, &cols, &dt);
    GTB_ARG(expect != 0);
    if (nbytes != expect)   // the reference's per-tensor check, tinyllama.cpp:316-319
        return fail(GTB_ERR_ARG, "Weight %d/%d data size: %zu does not match the expected size: %zu.", layer, tensor_id, nbytes, expect);
    const bool per_layer = tensor_id >= GTB_T_Q;
    GTB_ARG(!per_layer || (layer >= 0 && layer < e->cfg.n_layers));
    if (dt == GTB_F16 && rows == 1) {
        uint16_t* dst = (tensor_id == GTB_T_FINAL_NORM) ? e->final_norm : (tensor_id == GTB_T_ATTN_NORM) ? e->L[layer].attn_norm : e->L[layer].ffn_norm;
        // stream-ordered with the kernels that read it
        GTB_CUDA(cudaMemcpyAsync(dst, h_payload, nbytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx().stream));
        if (!on_device) GTB_CUDA(cudaStreamSynchronize(ctx().stream));
        return GTB_OK;
    }
    gtb_weight_t* slot = nullptr;
    switch (tensor_id) {
        case GTB_T_EMBED: slot = &e->embed; break;
        case GTB_T_LM_HEAD: slot = &e->lm_head; break;
        case GTB_T_Q: slot = &e->L[layer].q; break;
        case GTB_T_K: slot = &e->L[layer].k; break;
        case GTB_T_V: slot = &e->L[layer].v; break;
        case GTB_T_O: slot = &e->L[layer].o; break;
        case GTB_T_GATE: slot = &e->L[layer].gate; break;
        case GTB_T_UP: slot = &e->L[layer].up; break;
        case GTB_T_DOWN: slot = &e->L[layer].down; break;
        default: return fail(GTB_ERR_ARG, "bad tensor id %d", tensor_id);
    }
    if (*slot) { e->weight_bytes -= (*slot)->nbytes; gtb_weight_free(*slot); *slot = nullptr; drop_graphs(e); }
    e->layers_valid = false;
    if (e->pf) { pf_destroy(e->pf); e->pf = nullptr; }     // the fp16 copies are rebuilt on the next batched prefill
    int r;
    int row_off = -1;
    uint8_t* fdata = nullptr;
    uint16_t* fsc = nullptr;
    if (per_layer) {
        LayerW& l = e->L[layer];
        switch (tensor_id) {
            case GTB_T_Q: row_off = 0; fdata = l.qkv_data; fsc = l.qkv_sc; break;
            case GTB_T_K: row_off = e->cfg.n_embd; fdata = l.qkv_data; fsc = l.qkv_sc; break;
            case GTB_T_V: row_off = e->cfg.n_embd + e->kv_dim; fdata = l.qkv_data; fsc = l.qkv_sc; break;
            case GTB_T_GATE: row_off = 0; fdata = l.gu_data; fsc = l.gu_sc; break;
            case GTB_T_UP: row_off = e->cfg.n_ffn; fdata = l.gu_data; fsc = l.gu_sc; break;
            default: break;
        }
    }
    void* vd = (row_off >= 0) ? fdata + weight_data_bytes(dt, row_off, cols) : nullptr;
    uint16_t* vs = (row_off >= 0 && fsc) ? fsc + weight_scale_bytes(dt, row_off, cols) / 2 : nullptr;
    if (on_device) r = weight_device_view(slot, h_payload, dt, rows, cols, vd, vs);
    else if (row_off >= 0) r = weight_upload_view(slot, h_payload, dt, rows, cols, vd, vs);
    else r = gtb_weight_upload(slot, h_payload, dt, rows, cols);
    if (r == GTB_OK) e->weight_bytes += (*slot)->nbytes;
    return r;
}

int gtb_engine_set_weight(gtb_engine_t e, int layer, int tensor_id, const void* h_payload, size_t nbytes) {
    return set_weight_from(e, layer, tensor_id, h_payload, nbytes, false);
}

// load_from_ckpt (tinyllama.cpp:336-392) as a pipeline: the file is mapped, every payload travels through two pinned staging
// buffers (the memcpy out of the mapping is what pulls the pages from disk) into one of two device buffers in the gten layout,
// and the repack kernel (gtb_api.cu) writes the device layout from there -- disk read, H2D copy and repack of consecutive
// chunks / tensors overlap.  Record names are checked against the converter's order (tinyllama_to_gten.py:157-201).
int gtb_engine_load_gten(gtb_engine_t e, const char* path) {
    GTB_CHECK_INIT();
    GTB_ARG(e && path);
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(GTB_ERR_ARG, "cannot open %s", path);
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 8) { close(fd); return fail(GTB_ERR_ARG, "cannot stat %s", path); }
    const size_t fsize = (size_t)sb.st_size;
    const uint8_t* base = static_cast<const uint8_t*>(mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0));
    close(fd);
    if (base == MAP_FAILED) return fail(GTB_ERR_ARG, "cannot map %s", path);
    madvise(const_cast<uint8_t*>(base), fsize, MADV_SEQUENTIAL);
    constexpr size_t CHUNK = 16u << 20;
    void* pinned[2] = {nullptr, nullptr};
    uint8_t* dtmp[2] = {nullptr, nullptr};
    cudaEvent_t pev[2] = {nullptr, nullptr}, dev[2] = {nullptr, nullptr};
    size_t dcap = 0;
    cudaStream_t st = ctx().stream;
    int r = GTB_OK, pk = 0, dk = 0;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        for (int i = 0; i < 2; i++) {
            if (pinned[i]) cudaFreeHost(pinned[i]);
            if (dtmp[i]) cudaFree(dtmp[i]);
            if (pev[i]) cudaEventDestroy(pev[i]);
            if (dev[i]) cudaEventDestroy(dev[i]);
        }
        munmap(const_cast<uint8_t*>(base), fsize);
    };
    int64_t magic = 0;
    memcpy(&magic, base, 8);
    if (magic != 0x454c49464e455447LL) { cleanup(); return fail(GTB_ERR_ARG, "Magic number in the binary does not match the expected one."); }
    // the largest payload sizes the two device staging buffers
    {
        static const int tids[] = {GTB_T_EMBED, GTB_T_Q, GTB_T_K, GTB_T_GATE, GTB_T_DOWN, GTB_T_LM_HEAD, GTB_T_ATTN_NORM};
        for (int t : tids) { int a, b, c; const size_t n = expect_bytes(e, t, &a, &b, &c); if (n > dcap) dcap = n; }
    }
    for (int i = 0; i < 2 && r == GTB_OK; i++) {
        if (cudaMallocHost(&pinned[i], CHUNK) != cudaSuccess || cudaMalloc((void**)&dtmp[i], dcap) != cudaSuccess ||
            cudaEventCreateWithFlags(&pev[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&dev[i], cudaEventDisableTiming) != cudaSuccess)
            r = fail(GTB_ERR_CUDA, "loader staging buffers: %s", cudaGetErrorString(cudaGetLastError()));
    }
    size_t off = 8;
    auto rd_i32 = [&](int32_t* v) -> bool { if (off + 4 > fsize) return false; memcpy(v, base + off, 4); off += 4; return true; };
    auto one = [&](int layer, int tid, const std::string& expect_name) -> int {
        for (int rep = 0; rep < 2; rep++) {                       // layer header, then weight name (tinyllama.cpp:301-334): both carry the tensor name
            int32_t n = 0;
            if (!rd_i32(&n) || n < 0 || n > 4096 || off + (size_t)n > fsize) return fail(GTB_ERR_ARG, "corrupt .gten record header");
            if ((size_t)n != expect_name.size() || memcmp(base + off, expect_name.data(), (size_t)n) != 0)
                return fail(GTB_ERR_ARG, "unexpected record `%.*s` where `%s` belongs: not a TinyLlama .gten file of this shape, or records out of order",
                            n, reinterpret_cast<const char*>(base + off), expect_name.c_str());
            off += (size_t)n;
        }
        int32_t nb = 0;
        if (!rd_i32(&nb) || nb < 0) return fail(GTB_ERR_ARG, "corrupt .gten payload size");
        if (off + (size_t)nb > fsize) return fail(GTB_ERR_ARG, "truncated .gten file");
        int rows, cols, dt;
        const size_t expect = expect_bytes(e, tid, &rows, &cols, &dt);
        if ((size_t)nb != expect)   // the reference's per-tensor check, tinyllama.cpp:316-319
            return fail(GTB_ERR_ARG, "Weight `%s` data size: %d does not match the expected size: %zu.", expect_name.c_str(), nb, expect);
        uint8_t* d = dtmp[dk];
        GTB_CUDA(cudaEventSynchronize(dev[dk]));                // the repack that read this buffer two tensors ago is done
        for (size_t o = 0; o < (size_t)nb; o += CHUNK) {
            const size_t n = ((size_t)nb - o < CHUNK) ? (size_t)nb - o : CHUNK;
            GTB_CUDA(cudaEventSynchronize(pev[pk]));            // the copy out of this pinned buffer is done
            memcpy(pinned[pk], base + off + o, n);              // page-in from disk happens here, while the previous chunk is on the bus
            GTB_CUDA(cudaMemcpyAsync(d + o, pinned[pk], n, cudaMemcpyHostToDevice, st));
            GTB_CUDA(cudaEventRecord(pev[pk], st));
            pk ^= 1;
        }
        off += (size_t)nb;
        int rr = set_weight_from(e, layer, tid, d, (size_t)nb, true);
        if (rr) return rr;
        GTB_CUDA(cudaEventRecord(dev[dk], st));
        dk ^= 1;
        return GTB_OK;
    };
    if (r == GTB_OK) r = one(0, GTB_T_EMBED, "model.embed_tokens.weight");
    static const struct { int tid; const char* name; } order[] = {
        {GTB_T_Q, "self_attn.q_proj.weight"}, {GTB_T_K, "self_attn.k_proj.weight"}, {GTB_T_V, "self_attn.v_proj.weight"},
        {GTB_T_O, "self_attn.o_proj.weight"}, {GTB_T_GATE, "mlp.gate_proj.weight"}, {GTB_T_UP, "mlp.up_proj.weight"},
        {GTB_T_DOWN, "mlp.down_proj.weight"}, {GTB_T_ATTN_NORM, "input_layernorm.weight"}, {GTB_T_FFN_NORM, "post_attention_layernorm.weight"}};
    for (int li = 0; li < e->cfg.n_layers && !r; li++)
        for (const auto& t : order) { r = one(li, t.tid, "model.layers." + std::to_string(li) + "." + t.name); if (r) break; }
    if (!r) r = one(0, GTB_T_FINAL_NORM, "model.norm.weight");
    if (!r) r = one(0, GTB_T_LM_HEAD, "lm_head.weight");
    const std::string err = ctx().err;
    cleanup();
    if (r) ctx().err = err;
    return r;
}

static int check_loaded(gtb_engine_t e) {
    if (!e->embed || !e->lm_head) return fail(GTB_ERR_STATE, "engine weights are not loaded");
    for (auto& l : e->L)
        if (!l.q || !l.k || !l.v || !l.o || !l.gate || !l.up || !l.down) return fail(GTB_ERR_STATE, "engine weights are not loaded");
    return GTB_OK;
}

int gtb_engine_reset(gtb_engine_t e) {
    GTB_CHECK_INIT();
    GTB_ARG(e);
    e->host_pos = 0;
    e->kvt_pos = 0;
    return set_state(e, 0, 0);
}

int gtb_engine_logits(gtb_engine_t e, const int32_t* h_tokens, int n_tokens, int start_pos, float* h_logits) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_tokens && n_tokens > 0 && start_pos >= 0 && start_pos < n_tokens);
    if (n_tokens > e->cfg.max_ctx)   // tinyllama.cpp:46-49
        return fail(GTB_ERR_ARG, "Number of prompt tokens (%d) exceed provided maximum ctx size (%d)", n_tokens, e->cfg.max_ctx);
    int r = check_loaded(e);
    if (r) return r;
    GTB_CUDA(cudaMemcpyAsync(e->tokens + start_pos, h_tokens + start_pos, (size_t)(n_tokens - start_pos) * 4, cudaMemcpyHostToDevice, ctx().stream));
    r = set_state(e, start_pos, n_tokens);
    if (r) return r;
    r = run_rows(e, n_tokens - start_pos - 1, 1, -1, start_pos, n_tokens);
    if (r) return r;
    e->host_pos = n_tokens;
    if (h_logits) {
        GTB_CUDA(cudaMemcpyAsync(h_logits, e->logits, (size_t)e->cfg.n_vocab * 4, cudaMemcpyDeviceToHost, ctx().stream));
    }
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    return GTB_OK;
}

int gtb_engine_prefill(gtb_engine_t e, const int32_t* h_tokens, int n_tokens) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_tokens && n_tokens > 0 && n_tokens < e->cfg.max_ctx);
    int r = check_loaded(e);
    if (r) return r;
    GTB_CUDA(cudaMemcpyAsync(e->tokens, h_tokens, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, ctx().stream));
    r = set_state(e, 0, n_tokens);
    if (r) return r;
    // the last prompt row also produces logits and the first generated token (greedy_sample's i == 0 step)
    r = run_rows(e, n_tokens - 1, 1, -1, 0, n_tokens);
    if (r) return r;
    e->host_pos = n_tokens;
    return GTB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- batched prefill (gtb_prefill.cu)
template <int WT>
static int pf_head(gtb_engine* e) {
    // final residual + RMSNorm + lm_head of the last prompt row through the order-exact phase kernel (tinyllama.cpp:57-58)
    PhaseArgs a{};
    a.K = e->cfg.n_embd; a.n_mats = 1;
    a.mat[0] = matof(e->lm_head, e->logits);
    a.src0 = e->hres; a.src1 = e->rd; a.normw = e->final_norm; a.res_out = e->xfinal;
    int r = launch_phase<WT, PRO_NORM>(e, a);
    if (r) return r;
    k_argmax_advance<<<1, 1024, 0, ctx().stream>>>(e->logits, e->cfg.n_vocab, e->tokens, e->st, -1);
    GTB_LAUNCHED();
    return GTB_OK;
}

extern "C" {

int gtb_engine_prefill_fast(gtb_engine_t e, const int32_t* h_tokens, int n_tokens) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_tokens && n_tokens > 0 && n_tokens < e->cfg.max_ctx);
    const gtb_model_config& c = e->cfg;
    int r = check_loaded(e);
    if (r) return r;
    const int E = c.n_embd, F = c.n_ffn, KV = e->kv_dim;
    if (!e->pf) { r = pf_create(&e->pf, c); if (r) return r; }
    pf_set_fused(e->pf, e->pf_fused != 0);
    pf_set_two_cta(e->pf, e->pf_2cta != 0);
    pf_set_pdl(e->pf_pdl != 0);
    pf_set_attn_two_pass(e->pf, e->pf_attn2 != 0);
    if (!pf_weights_ready(e->pf)) {
        for (int li = 0; li < c.n_layers; li++) {
            LayerW& l = e->L[li];
            r = pf_set_weight(e->pf, li, 0, c.wdtype, l.qkv_data, l.qkv_sc, E + 2 * KV, E);
            if (!r) r = pf_set_weight(e->pf, li, 1, c.wdtype, l.o->data, l.o->scales, E, E);
            if (!r) r = pf_set_weight(e->pf, li, 2, c.wdtype, l.gu_data, l.gu_sc, 2 * F, E);
            if (!r) r = pf_set_weight(e->pf, li, 3, c.wdtype, l.down->data, l.down->scales, E, F);
            if (r) return r;
        }
    }
    if (e->capture && e->pf_cap_T != n_tokens) {
        cudaFree(e->pf_cap);
        e->pf_cap = nullptr;
        GTB_CUDA(cudaMalloc((void**)&e->pf_cap, ((size_t)c.n_layers * 12 + 1) * n_tokens * e->capw * 4));
        e->pf_cap_T = n_tokens;
    }
    GTB_CUDA(cudaMemcpyAsync(e->tokens, h_tokens, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, ctx().stream));
    e->kvt_pos = 0;                  // the tensor-core prefill rewrites the natural K/V rows only
    std::vector<PfLayerIO> io(c.n_layers);
    for (int li = 0; li < c.n_layers; li++) {
        LayerW& l = e->L[li];
        io[li] = PfLayerIO{l.attn_norm, l.ffn_norm, l.kq, l.ks, l.vq, l.vs};
    }
    PfRun run{};
    run.d_tokens = e->tokens; run.T = n_tokens; run.embed = e->embed; run.layers = io.data();
    run.rope_cos = e->rope_cos; run.rope_sin = e->rope_sin; run.last_res = e->hres; run.last_down = e->rd;
    run.cap = e->capture ? e->pf_cap : nullptr; run.capw = e->capw;
    run.n_layers_run = (e->pf_layers > 0 && e->pf_layers < c.n_layers) ? e->pf_layers : c.n_layers;
    if (e->capture && e->pf_cap) GTB_CUDA(cudaMemsetAsync(e->pf_cap, 0, ((size_t)c.n_layers * 12 + 1) * n_tokens * e->capw * 4, ctx().stream));
    r = pf_run(e->pf, run);
    if (r) return r;
    r = set_state(e, n_tokens - 1, n_tokens);
    if (r) return r;
    r = (c.wdtype == GTB_Q8) ? pf_head<DT_Q8>(e) : (c.wdtype == GTB_Q4) ? pf_head<DT_Q4>(e) : pf_head<DT_F16>(e);
    if (r) return r;
    e->host_pos = n_tokens;
    return GTB_OK;
}

int gtb_engine_pf_acv(gtb_engine_t e, int layer, int acv_id, int row, float* h_out, int* width) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_out);
    if (!e->capture || !e->pf_cap) return fail(GTB_ERR_STATE, "no batched-prefill capture: set capture_acv and call gtb_engine_prefill_fast");
    GTB_ARG(row >= 0 && row < e->pf_cap_T);
    int w = e->cfg.n_embd;
    if (acv_id == GTB_A_K || acv_id == GTB_A_V) w = e->kv_dim;
    if (acv_id == GTB_A_GATE || acv_id == GTB_A_UP) w = e->cfg.n_ffn;
    GTB_ARG(acv_id == GTB_A_EMB || (acv_id >= GTB_A_ATTN_NORM && acv_id <= GTB_A_ATTN_RES && layer >= 0 && layer < e->cfg.n_layers));
    const size_t T = (size_t)e->pf_cap_T;
    const size_t slot = (acv_id == GTB_A_EMB) ? (size_t)e->cfg.n_layers * 12 : (size_t)layer * 12 + (acv_id - GTB_A_ATTN_NORM);
    const float* src = e->pf_cap + (slot * T + row) * e->capw;
    GTB_CUDA(cudaMemcpyAsync(h_out, src, (size_t)w * 4, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    if (width) *width = w;
    return GTB_OK;
}

int gtb_pf_gemm_f32(const void* h_A16, const void* h_W16, int M, int N, int K, int bn, float* h_C) {
    GTB_CHECK_INIT();
    GTB_ARG(h_A16 && h_W16 && h_C && M > 0 && N > 0 && K > 0);
    void *dA = nullptr, *dW = nullptr;
    float* dC = nullptr;
    const int Ma = M < 128 ? 128 : M, Na = N < bn ? bn : N;
    GTB_CUDA(cudaMalloc(&dA, (size_t)Ma * K * 2));
    GTB_CUDA(cudaMalloc(&dW, (size_t)Na * K * 2));
    GTB_CUDA(cudaMalloc((void**)&dC, (size_t)M * N * 4));
    cudaMemsetAsync(dA, 0, (size_t)Ma * K * 2, ctx().stream);
    cudaMemsetAsync(dW, 0, (size_t)Na * K * 2, ctx().stream);
    cudaMemcpyAsync(dA, h_A16, (size_t)M * K * 2, cudaMemcpyHostToDevice, ctx().stream);
    cudaMemcpyAsync(dW, h_W16, (size_t)N * K * 2, cudaMemcpyHostToDevice, ctx().stream);
    int r = pf_gemm_f32(dA, dW, dC, M, N, K, bn);
    cudaError_t ce = cudaSuccess;
    if (!r) ce = cudaMemcpyAsync(h_C, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost, ctx().stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx().stream);
    cudaFree(dA); cudaFree(dW); cudaFree(dC);
    if (r) return r;
    if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "tcgen05 GEMM self-test failed: %s", cudaGetErrorString(ce));
    return GTB_OK;
}

int gtb_engine_decode(gtb_engine_t e, int n_steps) {
    GTB_CHECK_INIT();
    GTB_ARG(e && n_steps >= 0);
    if (e->host_pos + n_steps > e->cfg.max_ctx) return fail(GTB_ERR_ARG, "decode past max_ctx (%d + %d > %d)", e->host_pos, n_steps, e->cfg.max_ctx);
    int r = check_loaded(e);
    if (r) return r;
    r = run_rows(e, 0, n_steps, -1);
    if (r) return r;
    e->host_pos += n_steps;
    return GTB_OK;
}

int gtb_engine_generate(gtb_engine_t e, int32_t* h_tokens, int n_prompt, int n_new, int eos_id, int* n_generated) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_tokens && n_prompt > 0 && n_new > 0);
    if (n_prompt + n_new - 1 > e->cfg.max_ctx) return fail(GTB_ERR_ARG, "n_prompt + n_new - 1 exceeds max_ctx");
    int r = check_loaded(e);
    if (r) return r;
    GTB_CUDA(cudaMemcpyAsync(e->tokens, h_tokens, (size_t)n_prompt * 4, cudaMemcpyHostToDevice, ctx().stream));
    r = set_state(e, 0, n_prompt);
    if (r) return r;
    int produced = 1;
    if (mega_ok(e)) {
        // prefill rows, the first token and every further greedy step in ONE launch; an EOS ends the loop on the device
        r = run_rows(e, n_prompt - 1, n_new, eos_id, 0, n_prompt);
        if (r) return r;
        DevState s;
        GTB_CUDA(cudaMemcpyAsync(&s, e->st, sizeof s, cudaMemcpyDeviceToHost, ctx().stream));
        GTB_CUDA(cudaStreamSynchronize(ctx().stream));
        produced = s.n_gen;
        GTB_CUDA(cudaMemcpyAsync(h_tokens + n_prompt, e->tokens + n_prompt, (size_t)produced * 4, cudaMemcpyDeviceToHost, ctx().stream));
        GTB_CUDA(cudaStreamSynchronize(ctx().stream));
        e->host_pos = n_prompt + produced - 1;
        if (n_generated) *n_generated = produced;
        return GTB_OK;
    }
    r = run_rows(e, n_prompt - 1, 1, eos_id, 0, n_prompt);
    if (r) return r;
    if (eos_id < 0) {
        r = run_rows(e, 0, n_new - 1, eos_id);
        if (r) return r;
        produced = n_new;
    } else {
        // with an EOS stop the host looks at the flag every 16 tokens (the reference checks every token, :426)
        while (produced < n_new) {
            DevState s;
            GTB_CUDA(cudaMemcpyAsync(&s, e->st, sizeof s, cudaMemcpyDeviceToHost, ctx().stream));
            GTB_CUDA(cudaStreamSynchronize(ctx().stream));
            if (s.stop) break;
            const int chunk = (n_new - produced < 16) ? n_new - produced : 16;
            // run one token at a time inside the chunk so that nothing is generated after an EOS
            r = run_rows(e, 0, 1, eos_id);
            if (r) return r;
            produced += 1;
            (void)chunk;
        }
    }
    GTB_CUDA(cudaMemcpyAsync(h_tokens + n_prompt, e->tokens + n_prompt, (size_t)produced * 4, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    e->host_pos = n_prompt + produced - 1;
    if (n_generated) *n_generated = produced;
    return GTB_OK;
}

// ---- batched decode: n sequences advance together, one weight read per step for all of them
static bool batch_is_exact(const gtb_engine* e) { return e->batch_exact && xr_ok(e); }

int gtb_engine_batch_create(gtb_engine_t e, int n_seq) {
    GTB_CHECK_INIT();
    GTB_ARG(e && n_seq >= 0);
    if (e->cfg.wdtype == GTB_F16 && !batch_is_exact(e)) return fail(GTB_ERR_STATE, "the order-free batched decode is built for Q8-activation models (Q8, Q4 weights)");
    if (batch_is_exact(e)) {
        if (n_seq > XR_MAX_SLOTS) return fail(GTB_ERR_ARG, "batched decode: at most %d sequences", XR_MAX_SLOTS);
    } else {
        GTB_ARG(n_seq <= FDB_MAX);
        if (e->cfg.n_embd > 2048 || e->cfg.n_ffn > 6144) return fail(GTB_ERR_STATE, "batched decode: n_embd <= 2048 and n_ffn <= 6144");
        if (e->gsz > FD_NW) return fail(GTB_ERR_STATE, "batched decode: at most %d query heads per KV group", FD_NW);
    }
    return batch_alloc(e, n_seq);
}

// Exact prefill of `n_tokens` prompt ids straight into slot `seq` (multi-row kernels, bit-identical to gtb_engine_prefill
// followed by gtb_engine_batch_adopt): K/V land in the slot's cache, the first greedy token is appended.
int gtb_engine_batch_prefill(gtb_engine_t e, int seq, const int32_t* h_tokens, int n_tokens) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_tokens && seq >= 0 && seq < e->bb.n && n_tokens > 0 && n_tokens < e->cfg.max_ctx);
    int r = check_loaded(e);
    if (r) return r;
    if (!xr_ok(e)) return fail(GTB_ERR_STATE, "gtb_engine_batch_prefill needs the multi-row path (Q8/Q4 model, option xrows on)");
    auto& b = e->bb;
    const gtb_model_config& c = e->cfg;
    const size_t KV = e->kv_dim, MC = c.max_ctx;
    cudaStream_t st = ctx().stream;
    GTB_CUDA(cudaMemcpyAsync(b.tokens + (size_t)seq * (MC + 2), h_tokens, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, st));
    DevState ns{0, n_tokens, 0, 0};
    GTB_CUDA(cudaMemcpyAsync(b.st + seq, &ns, sizeof ns, cudaMemcpyHostToDevice, st));
    GTB_CUDA(cudaStreamSynchronize(st));                  // `ns` is a stack object
    XrModel m{};
    r = xr_model(e, m);
    if (r) return r;
    XrKV kv{b.kq.data(), b.ks.data(), b.vq.data(), b.vs.data(), MC * KV, MC * (KV / 32)};
    XrSeq sq{b.tokens, (int)MC + 2, b.st};
    for (int done = 0; done < n_tokens;) {
        const int n = (n_tokens - done < e->xr_rows) ? n_tokens - done : e->xr_rows;
        const bool last = done + n == n_tokens;
        r = xr_prefill_pass(e->xr, m, kv, sq, seq, done, n, n_tokens, last, e->batch_eos, b.logits + (size_t)seq * c.n_vocab);
        if (r) return r;
        done += n;
    }
    b.host_pos[seq] = n_tokens;
    return GTB_OK;
}

// slot `seq` <- the engine's current sequence (tokens, position, K/V cache), e.g. after gtb_engine_prefill[_fast]
int gtb_engine_batch_adopt(gtb_engine_t e, int seq) {
    GTB_CHECK_INIT();
    GTB_ARG(e && seq >= 0 && seq < e->bb.n);
    const gtb_model_config& c = e->cfg;
    auto& b = e->bb;
    cudaStream_t st = ctx().stream;
    DevState s;
    GTB_CUDA(cudaMemcpyAsync(&s, e->st, sizeof s, cudaMemcpyDeviceToHost, st));
    GTB_CUDA(cudaStreamSynchronize(st));
    const int pos = s.pos;
    GTB_ARG(pos >= 0 && pos < c.max_ctx);
    const size_t KV = e->kv_dim, MC = c.max_ctx, esz = (e->adtype == GTB_F16) ? 2 : 1;
    for (int li = 0; li < c.n_layers; li++) {
        LayerW& l = e->L[li];
        GTB_CUDA(cudaMemcpyAsync(b.kq[li] + seq * MC * KV * esz, l.kq, (size_t)pos * KV * esz, cudaMemcpyDeviceToDevice, st));
        GTB_CUDA(cudaMemcpyAsync(b.vq[li] + seq * MC * KV * esz, l.vq, (size_t)pos * KV * esz, cudaMemcpyDeviceToDevice, st));
        if (l.ks) {
            GTB_CUDA(cudaMemcpyAsync(b.ks[li] + seq * MC * (KV / 32), l.ks, (size_t)pos * (KV / 32) * 2, cudaMemcpyDeviceToDevice, st));
            GTB_CUDA(cudaMemcpyAsync(b.vs[li] + seq * MC * (KV / 32), l.vs, (size_t)pos * (KV / 32) * 2, cudaMemcpyDeviceToDevice, st));
        }
    }
    GTB_CUDA(cudaMemcpyAsync(b.tokens + (size_t)seq * (MC + 2), e->tokens, (size_t)(pos + 1) * 4, cudaMemcpyDeviceToDevice, st));
    DevState ns{pos, 0, 0, 0};
    GTB_CUDA(cudaMemcpyAsync(b.st + seq, &ns, sizeof ns, cudaMemcpyHostToDevice, st));
    GTB_CUDA(cudaStreamSynchronize(st));
    b.host_pos[seq] = pos;
    return GTB_OK;
}

int gtb_engine_batch_decode(gtb_engine_t e, int n_steps) {
    GTB_CHECK_INIT();
    GTB_ARG(e && n_steps > 0 && e->bb.n > 0);
    int r = check_loaded(e);
    if (r) return r;
    auto& b = e->bb;
    int max_pos = 0;
    for (int s = 0; s < b.n; s++) {
        if (b.host_pos[s] + n_steps > e->cfg.max_ctx) return fail(GTB_ERR_ARG, "batch decode past max_ctx (sequence %d: %d + %d > %d)", s, b.host_pos[s], n_steps, e->cfg.max_ctx);
        if (b.host_pos[s] > max_pos) max_pos = b.host_pos[s];
    }
    const bool exact = batch_is_exact(e);
    if (!exact && b.n > FDB_MAX) return fail(GTB_ERR_STATE, "the order-free batched decode takes at most %d sequences", FDB_MAX);
    cudaStream_t st = ctx().stream;
    // exact mode: rows of all slots through the multi-row kernels (gtb_xrows.cu); the attention score buffer is sized for the
    // longest context the call reaches, in steps of 256 positions
    XrModel m{};
    int t_cap = 0;
    if (exact) {
        r = xr_model(e, m);
        if (r) return r;
        t_cap = ((max_pos + n_steps + 255) / 256) * 256 + 32;
    }
    const size_t KV = e->kv_dim, MC = e->cfg.max_ctx;
    XrKV kv{b.kq.data(), b.ks.data(), b.vq.data(), b.vs.data(), MC * KV, MC * (KV / 32)};
    XrSeq sq{b.tokens, (int)MC + 2, b.st};
    auto enqueue = [&]() -> int {
        if (exact) return xr_decode_pass(e->xr, m, kv, sq, b.n, t_cap, e->batch_eos, b.logits);
        return (e->cfg.wdtype == GTB_Q8) ? enqueue_batch_row<DT_Q8>(e) : enqueue_batch_row<DT_Q4>(e);
    };
    if (e->use_graph) {
        if (b.graph && (b.graph_exact != exact || (exact && b.graph_tcap < t_cap))) { cudaGraphExecDestroy(b.graph); b.graph = nullptr; }
        if (!b.graph) {
            const int64_t before = ctx().launches;
            GTB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            r = enqueue();
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(st, &g);
            b.graph_launches = (int)(ctx().launches - before);
            ctx().launches = before;
            if (r) { if (g) cudaGraphDestroy(g); return r; }
            if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&b.graph, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
            b.graph_exact = exact; b.graph_tcap = t_cap;
        }
        for (int i = 0; i < n_steps; i++) { GTB_CUDA(cudaGraphLaunch(b.graph, st)); ctx().launches += b.graph_launches; }
    } else {
        for (int i = 0; i < n_steps; i++) {
            r = enqueue();
            if (r) return r;
        }
    }
    for (int s = 0; s < b.n; s++) b.host_pos[s] += n_steps;
    return GTB_OK;
}

int gtb_engine_batch_position(gtb_engine_t e, int seq, int* pos) {
    GTB_CHECK_INIT();
    GTB_ARG(e && pos && seq >= 0 && seq < e->bb.n);
    DevState s;
    GTB_CUDA(cudaMemcpyAsync(&s, e->bb.st + seq, sizeof s, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    *pos = s.pos;
    return GTB_OK;
}

int gtb_engine_batch_read_tokens(gtb_engine_t e, int seq, int32_t* h_tokens, int first, int count) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_tokens && seq >= 0 && seq < e->bb.n && first >= 0 && count > 0 && first + count <= e->cfg.max_ctx + 2);
    GTB_CUDA(cudaMemcpyAsync(h_tokens, e->bb.tokens + (size_t)seq * (e->cfg.max_ctx + 2) + first, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    return GTB_OK;
}

int gtb_engine_batch_read_logits(gtb_engine_t e, int seq, float* h_logits) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_logits && seq >= 0 && seq < e->bb.n);
    GTB_CUDA(cudaMemcpyAsync(h_logits, e->bb.logits + (size_t)seq * e->cfg.n_vocab, (size_t)e->cfg.n_vocab * 4, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    return GTB_OK;
}

int gtb_engine_position(gtb_engine_t e, int* pos) {
    GTB_CHECK_INIT();
    GTB_ARG(e && pos);
    DevState s;
    GTB_CUDA(cudaMemcpyAsync(&s, e->st, sizeof s, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    *pos = s.pos;
    return GTB_OK;
}

int gtb_engine_read_tokens(gtb_engine_t e, int32_t* h_tokens, int first, int count) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_tokens && first >= 0 && count > 0 && first + count <= e->cfg.max_ctx + 2);
    GTB_CUDA(cudaMemcpyAsync(h_tokens, e->tokens + first, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    return GTB_OK;
}

int gtb_engine_read_logits(gtb_engine_t e, float* h_logits) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_logits);
    GTB_CUDA(cudaMemcpyAsync(h_logits, e->logits, (size_t)e->cfg.n_vocab * 4, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    return GTB_OK;
}

int gtb_engine_topk(gtb_engine_t e, int k, float* h_values, int32_t* h_ids) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_values && h_ids && k > 0 && k <= 64 && k <= e->cfg.n_vocab);
    float* dv = nullptr;
    GTB_CUDA(cudaMalloc((void**)&dv, 64 * 8));
    int32_t* di = reinterpret_cast<int32_t*>(dv + 64);
    k_topk<<<1, 1024, 0, ctx().stream>>>(e->logits, e->cfg.n_vocab, k, dv, di);
    ctx().launches++;
    cudaError_t ce = cudaMemcpyAsync(h_values, dv, (size_t)k * 4, cudaMemcpyDeviceToHost, ctx().stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_ids, di, (size_t)k * 4, cudaMemcpyDeviceToHost, ctx().stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx().stream);
    cudaFree(dv);
    if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "top-k failed: %s", cudaGetErrorString(ce));
    return GTB_OK;
}

int gtb_engine_acv(gtb_engine_t e, int layer, int acv_id, float* h_out, int* width) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_out);
    if (!e->capture) return fail(GTB_ERR_STATE, "activation capture is off: gtb_engine_set_option(e, \"capture_acv\", 1)");
    int w = e->cfg.n_embd;
    if (acv_id == GTB_A_K || acv_id == GTB_A_V) w = e->kv_dim;
    if (acv_id == GTB_A_GATE || acv_id == GTB_A_UP) w = e->cfg.n_ffn;
    GTB_ARG(acv_id == GTB_A_EMB || acv_id == GTB_A_FINAL_NORM || (acv_id >= GTB_A_ATTN_NORM && acv_id <= GTB_A_ATTN_RES && layer >= 0 && layer < e->cfg.n_layers));
    const float* src = capp(e, layer, acv_id);
    GTB_CUDA(cudaMemcpyAsync(h_out, src, (size_t)w * 4, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    if (width) *width = w;
    return GTB_OK;
}

int gtb_engine_set_option(gtb_engine_t e, const char* name, int value) {
    GTB_ARG(e && name);
    if (!strcmp(name, "graph")) { e->use_graph = value != 0; return GTB_OK; }
    if (!strcmp(name, "capture_acv")) { e->capture = value != 0; drop_graphs(e); return GTB_OK; }
    if (!strcmp(name, "grid")) { GTB_ARG(value > 0); e->grid = value; drop_graphs(e); return GTB_OK; }
    if (!strcmp(name, "mega")) { e->use_mega = value != 0; return GTB_OK; }
    if (!strcmp(name, "pf_ahead")) { GTB_ARG(value >= 0 && value <= 64); e->pf_ahead = value; return GTB_OK; }
    if (!strcmp(name, "prof")) { e->prof = value != 0; return GTB_OK; }
    if (!strcmp(name, "prof_cta")) { GTB_ARG(value >= 0); e->prof_cta = value; return GTB_OK; }
    if (!strcmp(name, "fast_decode")) { e->fast = value != 0; drop_graphs(e); return GTB_OK; }
    if (!strcmp(name, "fd_ahead")) { GTB_ARG(value >= 0 && value <= 16); e->fd_ahead = value; drop_graphs(e); e->fd_args_valid = false; return GTB_OK; }
    if (!strcmp(name, "fd_chunk")) { GTB_ARG(value >= 32 && value <= 1024 && value % 32 == 0); e->fd_chunk = value; drop_graphs(e); return GTB_OK; }
    if (!strcmp(name, "fd_prof_cta")) { GTB_ARG(value >= 0); e->fd_prof_cta = value; return GTB_OK; }
    if (!strcmp(name, "fd_trace")) {          // debug timeline of the fast-decode chains into the "prof" buffer (gtb_engine_read_prof)
        long long* p = value ? e->d_prof : nullptr;
        GTB_CUDA(cudaMemsetAsync(e->d_prof, 0, 8, ctx().stream));
        GTB_CUDA(cudaMemcpyToSymbolAsync(g_fd_trace, &p, sizeof p, 0, cudaMemcpyHostToDevice, ctx().stream));
        GTB_CUDA(cudaStreamSynchronize(ctx().stream));
        return GTB_OK;
    }
    if (!strcmp(name, "fd_mega")) { e->fd_mega = value != 0; return GTB_OK; }
    if (!strcmp(name, "pf_layers")) { GTB_ARG(value >= 0); e->pf_layers = value; return GTB_OK; }
    if (!strcmp(name, "pf_fused")) { e->pf_fused = value != 0; return GTB_OK; }
    if (!strcmp(name, "pf_2cta")) { e->pf_2cta = value != 0; return GTB_OK; }
    if (!strcmp(name, "pf_pdl")) { e->pf_pdl = value != 0; return GTB_OK; }
    if (!strcmp(name, "pf_attn2")) { e->pf_attn2 = value != 0; return GTB_OK; }
    if (!strcmp(name, "xrows")) { e->use_xr = value != 0; return GTB_OK; }
    if (!strcmp(name, "xr_min_rows")) { GTB_ARG(value >= 1); e->xr_min_rows = value; return GTB_OK; }
    if (!strcmp(name, "xr_rows")) { GTB_ARG(value >= 1 && value <= XR_MAX_ROWS); e->xr_rows = value; return GTB_OK; }
    if (!strcmp(name, "batch_exact")) { e->batch_exact = value != 0; drop_graphs(e); return GTB_OK; }
    if (!strcmp(name, "batch_eos")) { GTB_ARG(value >= -1); e->batch_eos = value; drop_graphs(e); return GTB_OK; }
    if (!strcmp(name, "xr_tensor")) { xr_set_tensor(value != 0); drop_graphs(e); return GTB_OK; }
    if (!strcmp(name, "xr_trace")) {           // cycle counters of the tensor-core GEMM into the "prof" buffer (gtb_engine_read_prof)
        GTB_CUDA(cudaMemsetAsync(e->d_prof, 0, 64 * 8, ctx().stream));
        xr_set_trace(value ? e->d_prof : nullptr); drop_graphs(e);
        return GTB_OK;
    }
    if (!strcmp(name, "xr_variant")) { xr_set_variant(value); drop_graphs(e); return GTB_OK; }
    if (!strcmp(name, "xr_pdl")) { xr_set_pdl(value != 0); drop_graphs(e); return GTB_OK; }
    return fail(GTB_ERR_ARG, "unknown option %s", name);
}

int gtb_engine_uses_megakernel(gtb_engine_t e, int* yes) {
    GTB_ARG(e && yes);
    *yes = mega_ok(e) ? 1 : 0;
    return GTB_OK;
}

int gtb_engine_read_prof(gtb_engine_t e, long long* h_out, int count) {
    GTB_CHECK_INIT();
    GTB_ARG(e && h_out && count > 0 && count <= PROF_SLOTS);
    GTB_CUDA(cudaMemcpyAsync(h_out, e->d_prof, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx().stream));
    GTB_CUDA(cudaStreamSynchronize(ctx().stream));
    return GTB_OK;
}

int gtb_selftest_exact_sum(const float* h_terms, int n, float* h_out) {
    GTB_CHECK_INIT();
    GTB_ARG(h_terms && h_out && n > 0 && n <= (1 << 20));
    float *d = nullptr, *o = nullptr;
    GTB_CUDA(cudaMalloc((void**)&d, (size_t)n * 4));
    GTB_CUDA(cudaMalloc((void**)&o, 32));
    cudaError_t ce = cudaMemcpyAsync(d, h_terms, (size_t)n * 4, cudaMemcpyHostToDevice, ctx().stream);
    if (ce == cudaSuccess) {
        k_selftest_exact_sum<<<1, MT, 0, ctx().stream>>>(d, n, o);
        ctx().launches++;
        ce = cudaMemcpyAsync(h_out, o, 20, cudaMemcpyDeviceToHost, ctx().stream);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx().stream);
    cudaFree(d); cudaFree(o);
    if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "exact-sum self-test failed: %s", cudaGetErrorString(ce));
    return GTB_OK;
}

int gtb_selftest_expf(uint32_t first_bits, uint32_t count, float* h_out) {
    GTB_CHECK_INIT();
    GTB_ARG(h_out && count > 0 && count <= (1u << 26));
    float* d = nullptr;
    GTB_CUDA(cudaMalloc((void**)&d, (size_t)count * 4));
    k_selftest_expf<<<ctx().sm_count * 8, 256, 0, ctx().stream>>>(first_bits, count, d);
    ctx().launches++;
    cudaError_t ce = cudaMemcpyAsync(h_out, d, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx().stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx().stream);
    cudaFree(d);
    if (ce != cudaSuccess) return fail(GTB_ERR_CUDA, "expf self-test failed: %s", cudaGetErrorString(ce));
    return GTB_OK;
}

int gtb_engine_weight_bytes(gtb_engine_t e, size_t* nbytes) {
    GTB_ARG(e && nbytes);
    *nbytes = e->weight_bytes;
    return GTB_OK;
}

}  // extern "C"
