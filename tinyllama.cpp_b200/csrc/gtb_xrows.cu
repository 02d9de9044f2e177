// gtb_xrows.cu -- the order-exact MULTI-ROW path: R rows of TinyLlama::logits (tinyllama.cpp:45-61) side by side.
//
// Why it exists: the row-at-a-time kernels (gtb_mega.cuh) are bound by the reference's dependent chain (57 us per
// layer whatever the bandwidth); rows, however, are independent given their tokens -- gten/ops.h:632 loops over rows,
// tinyllama.cpp:395-440 over sequences -- so R ordered chains run side by side against ONE load of each weight block.
// Every (row, output column, lane) still runs exactly the reference's chain (ops.h:282-292): four integer lane sums per
// 32-block, `acc[l] += float(lane[l]) * (da * dw)` in ascending block order, `(a0+a1)+(a2+a3)` at the end; all
// re-encode points of SURVEY.md App. A sit in the epilogues.  The per-row in-order sums (ops.h:765-767, 982-988) are
// plain serial chains here: R of them run concurrently, which is what makes the simple form affordable.
//
// One pass over R <= 64 rows = 6 kernels per layer:
//   k_xr_norm               E(rmsnorm(x))                                  -> staged Q8 rows (XBlk records)
//   k_xr_gemm<EPI_QKV>      q|k|v Linear + E + RoPE + E, K/V append        -> staged q, K/V cache
//   k_xr_attn               scores, exact softmax, E(P), P.V, E            -> staged attention output
//   k_xr_gemm<EPI_RES>      o Linear;   x = E(x + E(o))                    -> residual stream (fp32 of Q8 values)
//   k_xr_norm
//   k_xr_gemm<EPI_SILU>     gate|up Linear; E(E(silu(E(gate))) * E(up))    -> staged MLP activation
//   k_xr_gemm<EPI_RES>      down Linear; x = E(x + E(down))
// plus final norm, k_xr_gemm<EPI_HEAD> (logits + per-tile first maximum) and k_xr_argmax for the rows that sample.
//
// The default GEMM is a SIMT kernel: the contract is integer lane sums (dp4a) followed by ORDERED fp32 adds; the ordered adds
// cannot go to a tensor core.  (The lane sums can, exactly: gtb_xtensor.cuh is that experiment -- bit-identical, not faster.)  CTA = 4 warps = 16 rows x 64 output columns; lane = column (2 per lane), warp = 4 rows;
// weights and staged rows stream through a 3-stage cp.async ring in chunks of 8 blocks (256 elements of K).
#include <algorithm>
#include <vector>

#include "gtb_xrows.h"
#include "gtb_mega.cuh"        // exact_sum512: the in-order sum emulated exactly by 512 threads

namespace gtb {

// ---------------------------------------------------------------- staged Q8 row: one 64-byte record per 32-block
// w[0..3] = codes of elements (2l, 2l+1, 2l+8, 2l+9), w[4..7] the same +16 (the weight layout's lane words,
// gtb_internal.h); nb[l] = XB_BIAS - 7 * (sum of the eight codes of lane l): the dp4a accumulator init for Q4 weights
// (value = nibble - 7) that also carries the int->float bias; d = fp32 value of the fp16 block scale.
struct __align__(16) XBlk { uint32_t w[8]; int32_t nb[4]; float d; uint32_t pad[3]; };
static_assert(sizeof(XBlk) == 64, "XBlk is one 64-byte record");
constexpr int XB_BIAS = 0x4b400000;          // bits of 1.5 * 2^23: (bits + l) is the float 12582912 + l for |l| < 2^22
constexpr float XB_M = 12582912.0f;

struct XrRow { int slot, pos, n_ctx, tok; };  // slot < 0: row not in use

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// programmatic dependent launch: the next kernel of the chain may start while this one drains; nothing that a predecessor wrote
// is touched before pdl_wait()
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// store one encoded block (one code per lane) as an XBlk record
__device__ __forceinline__ void xblk_store(XBlk* dst, int lane, int q, uint16_t dh) {
    const uint32_t qb = (uint32_t)q & 0xffu;
    const int w = lane & 7, e0 = 16 * (w >> 2) + 2 * (w & 3);
    const uint32_t b0 = __shfl_sync(0xffffffffu, qb, e0), b1 = __shfl_sync(0xffffffffu, qb, e0 + 1);
    const uint32_t b2 = __shfl_sync(0xffffffffu, qb, e0 + 8), b3 = __shfl_sync(0xffffffffu, qb, e0 + 9);
    int s = q + __shfl_xor_sync(0xffffffffu, q, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);                    // lanes 0, 2, 4, 6: reference lanes 0..3
    const int nbv = __shfl_sync(0xffffffffu, XB_BIAS - 7 * s, 2 * (lane & 3));
    if (lane < 8) dst->w[lane] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    else if (lane < 12) dst->nb[lane - 8] = nbv;
    else if (lane == 12) dst->d = h2f(dh);
}
// the 8 permuted code words of a block, assembled in lanes 0..7 (K cache rows use the same order)
__device__ __forceinline__ uint32_t perm_word(int lane, int q) {
    const uint32_t qb = (uint32_t)q & 0xffu;
    const int w = lane & 7, e0 = 16 * (w >> 2) + 2 * (w & 3);
    const uint32_t b0 = __shfl_sync(0xffffffffu, qb, e0), b1 = __shfl_sync(0xffffffffu, qb, e0 + 1);
    const uint32_t b2 = __shfl_sync(0xffffffffu, qb, e0 + 8), b3 = __shfl_sync(0xffffffffu, qb, e0 + 9);
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}
__device__ __forceinline__ uint32_t natural_word(int lane, int q) {       // bytes 4w .. 4w+3 in lanes 0..7
    const uint32_t qb = (uint32_t)q & 0xffu;
    const int e0 = 4 * (lane & 7);
    const uint32_t b0 = __shfl_sync(0xffffffffu, qb, e0), b1 = __shfl_sync(0xffffffffu, qb, e0 + 1);
    const uint32_t b2 = __shfl_sync(0xffffffffu, qb, e0 + 2), b3 = __shfl_sync(0xffffffffu, qb, e0 + 3);
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

// ---------------------------------------------------------------- pass plan
struct XrPlanArgs {
    int mode;                  // 0: rows = positions p0 + r of slot0;  1: row r = slot r at its own position
    int n_rows, slot0, p0, n_ctx, tok_stride;
    const int32_t* tokens;
    const DevState* st;
    XrRow* rows;
};
__global__ void k_xr_plan(XrPlanArgs a) {
    pdl_launch();
    pdl_wait();
    const int r = threadIdx.x;
    if (r >= XR_MAX_ROWS) return;
    XrRow row{-1, 0, 0, 0};
    if (r < a.n_rows) {
        if (a.mode == 0) {
            row.slot = a.slot0; row.pos = a.p0 + r; row.n_ctx = a.n_ctx;
        } else {
            row.slot = a.slot0 + r; row.pos = a.st[row.slot].pos; row.n_ctx = row.pos + 1;
            if (a.st[row.slot].stop) row.slot = -1;       // a sequence that sampled EOS has left the batch (tinyllama.cpp:426: break)
        }
        if (row.slot >= 0) row.tok = a.tokens[(size_t)row.slot * a.tok_stride + row.pos];
    }
    a.rows[r] = row;
}

// ---------------------------------------------------------------- RMSNorm rows (gten/ops.h:762-804), one CTA per row
struct XrNormArgs {
    const XrRow* rows;
    int row0;
    float* res;                // [rows][E]: the residual stream (decoded values of its Q8 encoding)
    const uint16_t* normw;
    XBlk* out;                 // [rows][E / 32]
    int E;
    const void* emb_w; const uint16_t* emb_s; int emb_dt;   // emb_w != null: the row is the token's embedding (ops.h:514-564)
};

constexpr int XN_NT = 512, XN_NW = XN_NT / 32;
__global__ void __launch_bounds__(XN_NT) k_xr_norm(XrNormArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* xbuf = reinterpret_cast<float*>(smem);                 // [E] the row
    float* sq = xbuf + a.E;                                       // [E] its squares
    __shared__ ExactSum2Smem es;
    static_assert(XN_NT == MT, "exact_sum512 runs on MT threads");
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nb = a.E / 32;
    const int row = a.row0 + blockIdx.x;
    // the norm weights do not depend on the previous kernel: fetch them before waiting for it
    float nw[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int b = wid + u * XN_NW; nw[u] = (b < nb) ? h2f(a.normw[b * 32 + lane]) : 0.0f; }
    pdl_wait();
    const XrRow rw = a.rows[row];
    if (rw.slot < 0) return;
    float* res = a.res + (size_t)row * a.E;
    for (int b = wid; b < nb; b += XN_NW) {
        const int e = b * 32 + lane;
        float v;
        if (a.emb_w) {
            const size_t blk = (size_t)rw.tok * nb + b;
            const float delta = h2f(a.emb_s[blk]);
            if (a.emb_dt == DT_Q8) {
                const int8_t q = reinterpret_cast<const int8_t*>(a.emb_w)[blk * 32 + perm_byte(lane)];        // the row is copied
                v = __fmul_rn((float)q, delta);
            } else {                                                                                         // Q4: dequantise, re-encode as Q8
                const int j = lane & 15;
                const int l = (j & 7) >> 1, bp = (j & 1) + 2 * (j >> 3);
                const uint8_t byte = reinterpret_cast<const uint8_t*>(a.emb_w)[blk * 16 + l * 4 + bp];
                const int q = (int)((lane < 16) ? (byte >> 4) : (byte & 0x0f)) - 7;
                v = q8_roundtrip_lane(__fmul_rn((float)q, delta));
            }
            res[e] = v;
        } else {
            v = res[e];
        }
        xbuf[e] = v;
        sq[e] = __fmul_rn(v, v);
    }
    __syncthreads();
    // ops.h:765-767: the sum of squares strictly in order, emulated exactly in parallel (gtb_mega.cuh: one integer add scan; 1.9 us
    // against 4.2 us for the plain 2048-step chain)
    const float sq_sum = exact_sum512([&](int i, float q[4]) {
        if (i < a.E) { const float4 v = *reinterpret_cast<const float4*>(sq + i); q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w; }
        else { q[0] = q[1] = q[2] = q[3] = 0.0f; }
    }, a.E, es);
    pdl_launch();
    const float denom = __fadd_rn(sqrtf(__fdiv_rn(sq_sum, (float)a.E)), 1e-6f);
    for (int b = wid, u = 0; b < nb; b += XN_NW, u++) {
        const int e = b * 32 + lane;
        const float y = __fmul_rn(__fdiv_rn(xbuf[e], denom), (u < 4) ? nw[u] : h2f(a.normw[e]));
        uint16_t dh;
        const int q = q8_encode_lane(y, &dh);
        xblk_store(a.out + (size_t)row * nb + b, lane, q, dh);
    }
}

// ---------------------------------------------------------------- the multi-row GEMM
constexpr int XG_BM = 16, XG_BN = 64, XG_KC = 8, XG_STAGES = 3;
enum { XEPI_QKV = 0, XEPI_RES = 1, XEPI_SILU = 2, XEPI_HEAD = 3 };

template <int WT>
struct XgStage {
    static constexpr int WB = (WT == DT_Q4) ? 1 : 2;           // 16-byte words per weight block
    uint4 wd[XG_BN][XG_KC * WB + 1];                            // +1: odd pitch, conflict-free 128-bit reads down a column
    uint4 ws[XG_BN];                                            // the chunk's 8 fp16 scales of each column
    XBlk act[XG_BM][XG_KC];
};

struct XrGemmArgs {
    const XrRow* rows;
    int row0, n_rows;                  // rows [row0, row0 + n_rows) of the pass; grid.y = ceil(n_rows / 16)
    const XBlk* act; int nb;           // staged input rows [row][nb], nb = K / 32 (a multiple of 8)
    const uint4* wd; const uint16_t* ws; int N;
    int up_off;                        // EPI_SILU: first row of `up` inside the fused gate|up matrix
    float* res; int E;                 // EPI_RES
    XBlk* out; int out_nb;             // EPI_SILU: staged MLP activation; EPI_QKV: staged q [row][n_heads * 2]
    uint8_t* kq; uint16_t* ks; uint8_t* vq; uint16_t* vs;
    size_t slot_codes, slot_scales;
    int kv_dim, n_heads, n_groups;
    const float* rope_cos; const float* rope_sin;
    float* logits; int ld_logits; float* arg_val; int* arg_idx; int n_tiles;   // EPI_HEAD
};

// Packed fp32 pairs (Blackwell f32x2 instructions): two independent IEEE round-to-nearest operations per instruction -- the same bits
// as two scalar operations, half the issue slots for the four lane chains of a (row, column)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// There is deliberately NO packed multiply here: ptxas 12.9 contracts `mul.rn.f32x2` followed by `add.rn.f32x2` into one FFMA2 (and
// does the same to `fma.rn.f32x2(a, b, -0.0)` + add), -fmad=false or not -- the product's rounding would be lost.  Products that feed
// an addition are scalar `__fmul_rn`; only the additions (and explicit FMAs) are packed.
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// The epilogue of one (row, 64-column tile n) for one warp: lane = column inside each 32-block (v[0]: column 64 n + lane, v[1]: column
// 64 n + 32 + lane; gate|up: gate channel 32 n + lane and up channel 32 n + lane), every re-encode is warp-local.  Shared by the SIMT
// GEMM (accumulators in registers) and the tensor-core GEMM's epilogue kernel (accumulated rows read back from HBM).
template <int EPI>
__device__ __forceinline__ void xr_epilogue(const XrGemmArgs& a, int row, const XrRow& rw, int n, int lane, const float (&v)[2], bool store_logits) {
    if (EPI == XEPI_RES) {
        // Residual (gten/ops.h:870-898): x = E(x + E(linear output))
#pragma unroll
        for (int c = 0; c < 2; c++) {
            float* px = a.res + (size_t)row * a.E + 64 * n + 32 * c + lane;
            const float d1 = q8_roundtrip_lane(v[c]);
            *px = q8_roundtrip_lane(__fadd_rn(*px, d1));
        }
    } else if (EPI == XEPI_SILU) {
        // gten/modules.cpp:238-247: E(E(silu(E(gate))) * E(up))
        const float g1 = q8_roundtrip_lane(v[0]);
        const float u1 = q8_roundtrip_lane(v[1]);
        const float g2 = q8_roundtrip_lane(silu_ref(g1));
        uint16_t dh;
        const int q = q8_encode_lane(__fmul_rn(g2, u1), &dh);
        xblk_store(a.out + (size_t)row * a.out_nb + n, lane, q, dh);
    } else if (EPI == XEPI_QKV) {
        const int pos = rw.pos;
        if (n < a.n_heads + a.n_groups) {
            // q / k head: E, RoPE on the pair (j, j + 32) (ops.h:733-751), E
            const float x0 = q8_roundtrip_lane(v[0]), x1 = q8_roundtrip_lane(v[1]);
            const float cs = __ldg(a.rope_cos + (size_t)pos * 32 + lane), sn = __ldg(a.rope_sin + (size_t)pos * 32 + lane);
            const float o0 = __fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn));
            const float o1 = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs));
            uint16_t dh0, dh1;
            const int q0 = q8_encode_lane(o0, &dh0), q1 = q8_encode_lane(o1, &dh1);
            if (n < a.n_heads) {
                XBlk* dst = a.out + ((size_t)row * a.n_heads + n) * 2;
                xblk_store(dst, lane, q0, dh0);
                xblk_store(dst + 1, lane, q1, dh1);
            } else {
                const int g = n - a.n_heads;
                const uint32_t w0 = perm_word(lane, q0), w1 = perm_word(lane, q1);
                uint32_t* kc = reinterpret_cast<uint32_t*>(a.kq + (size_t)rw.slot * a.slot_codes + (size_t)pos * a.kv_dim + g * 64);
                uint16_t* ks = a.ks + (size_t)rw.slot * a.slot_scales + (size_t)pos * (a.kv_dim / 32) + g * 2;
                if (lane < 8) { kc[lane] = w0; kc[8 + lane] = w1; }
                if (lane == 8) { ks[0] = dh0; ks[1] = dh1; }
            }
        } else {
            const int g = n - a.n_heads - a.n_groups;
            uint16_t dh0, dh1;
            const int q0 = q8_encode_lane(v[0], &dh0), q1 = q8_encode_lane(v[1], &dh1);
            const uint32_t w0 = natural_word(lane, q0), w1 = natural_word(lane, q1);
            uint32_t* vc = reinterpret_cast<uint32_t*>(a.vq + (size_t)rw.slot * a.slot_codes + (size_t)pos * a.kv_dim + g * 64);
            uint16_t* vs = a.vs + (size_t)rw.slot * a.slot_scales + (size_t)pos * (a.kv_dim / 32) + g * 2;
            if (lane < 8) { vc[lane] = w0; vc[8 + lane] = w1; }
            if (lane == 8) { vs[0] = dh0; vs[1] = dh1; }
        }
    } else {
        // lm_head (modules.cpp:70-81): fp32 logits + the tile's first maximum (tinyllama.cpp:416-424)
        const int c0 = 64 * n + lane, c1 = c0 + 32;
        float best = -INFINITY;
        int arg = 0x7fffffff;
        if (c0 < a.N) { if (store_logits) a.logits[(size_t)(row - a.row0) * a.ld_logits + c0] = v[0]; if (v[0] > best) { best = v[0]; arg = c0; } }
        if (c1 < a.N) { if (store_logits) a.logits[(size_t)(row - a.row0) * a.ld_logits + c1] = v[1]; if (v[1] > best) { best = v[1]; arg = c1; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
        }
        if (lane == 0) { a.arg_val[(size_t)row * a.n_tiles + n] = best; a.arg_idx[(size_t)row * a.n_tiles + n] = arg; }
    }
}

// NW warps per CTA, TR = 16 / NW rows per warp: (4, 4) amortises a weight block over 4 rows (fewest instructions; large N),
// (8, 2) doubles the warps per tile for the matrices with few column tiles (N <= 2560: q|k|v, o, down)
template <int WT, int EPI, int NW, int NS = XG_STAGES, int MINB = (NW == 4) ? 4 : 2>
__global__ void __launch_bounds__(NW * 32, MINB) k_xr_gemm(XrGemmArgs a) {
    constexpr int XG_NT = NW * 32, TR = XG_BM / NW;
    extern __shared__ __align__(16) unsigned char smem[];
    XgStage<WT>* stages = reinterpret_cast<XgStage<WT>*>(smem);
    constexpr int WB = XgStage<WT>::WB, WPC = XG_KC * WB;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = blockIdx.x;
    const int rbase = a.row0 + blockIdx.y * XG_BM, rend = a.row0 + a.n_rows;
    const int nb = a.nb, nchunk = nb / XG_KC;
    // weight row of local column c: plain tiles are 64 consecutive rows; gate|up tiles are 32 gate rows + the same 32 up rows
    auto wrow = [&](int c) -> int {
        int r = (EPI == XEPI_SILU) ? ((c < 32) ? 32 * n + c : a.up_off + 32 * n + c - 32) : 64 * n + c;
        return min(r, a.N - 1);
    };
    auto load_weights = [&](int s, int kc) {
        XgStage<WT>& st = stages[s];
#pragma unroll
        for (int i = tid; i < XG_BN * WPC; i += XG_NT) {
            const int c = i / WPC, j = i % WPC;
            cp_async16(&st.wd[c][j], a.wd + ((size_t)wrow(c) * nb + (size_t)kc * XG_KC) * WB + j, true);
        }
        if (tid < XG_BN) cp_async16(&st.ws[tid], a.ws + (size_t)wrow(tid) * nb + (size_t)kc * XG_KC, true);
    };
    auto load_act = [&](int s, int kc) {
        XgStage<WT>& st = stages[s];
#pragma unroll
        for (int i = tid; i < XG_BM * XG_KC * 4; i += XG_NT) {
            const int r = i >> 5, j = i & 31;
            const int gr = rbase + r;
            const bool ok = gr < rend;
            const uint4* src = reinterpret_cast<const uint4*>(a.act + ((size_t)(ok ? gr : a.row0) * nb + (size_t)kc * XG_KC)) + j;
            cp_async16(reinterpret_cast<uint4*>(&st.act[r][0]) + j, src, ok);
        }
    };
    f32x2 acc[TR][2][2];                                             // lanes (0, 1) and (2, 3) of each (row, column)
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) { acc[r][c][0] = pk2(0.0f, 0.0f); acc[r][c][1] = pk2(0.0f, 0.0f); }
    // weights never depend on the previous kernel of the chain: their first chunks are in flight before it has finished
#pragma unroll
    for (int s = 0; s < NS - 1; s++)
        if (s < nchunk) load_weights(s, s);
    pdl_wait();
#pragma unroll
    for (int s = 0; s < NS - 1; s++) {
        if (s < nchunk) load_act(s, s);
        cp_async_commit();
    }
    for (int kc = 0; kc < nchunk; kc++) {
        cp_async_wait<NS - 2>();
        __syncthreads();
        if (kc + NS - 1 < nchunk) {
            load_weights((kc + NS - 1) % NS, kc + NS - 1);
            load_act((kc + NS - 1) % NS, kc + NS - 1);
        }
        cp_async_commit();
        const XgStage<WT>& st = stages[kc % NS];
        const uint4 sA = st.ws[lane], sB = st.ws[lane + 32];
        const __half* hA = reinterpret_cast<const __half*>(&sA);
        const __half* hB = reinterpret_cast<const __half*>(&sB);
#pragma unroll
        for (int b = 0; b < XG_KC; b++) {
            // the two columns' blocks: wx = codes paired with the activation's X words, wy with its Y words
            uint32_t wx[2][4], wy[2][4];
            if (WT == DT_Q4) {
                const uint4 w0 = st.wd[lane][b], w1 = st.wd[lane + 32][b];
                const uint32_t r0[4] = {w0.x, w0.y, w0.z, w0.w}, r1[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    wx[0][l] = (r0[l] >> 4) & 0x0f0f0f0fu; wy[0][l] = r0[l] & 0x0f0f0f0fu;
                    wx[1][l] = (r1[l] >> 4) & 0x0f0f0f0fu; wy[1][l] = r1[l] & 0x0f0f0f0fu;
                }
            } else {
                const uint4 x0 = st.wd[lane][2 * b], y0 = st.wd[lane][2 * b + 1];
                const uint4 x1 = st.wd[lane + 32][2 * b], y1 = st.wd[lane + 32][2 * b + 1];
                wx[0][0] = x0.x; wx[0][1] = x0.y; wx[0][2] = x0.z; wx[0][3] = x0.w;
                wy[0][0] = y0.x; wy[0][1] = y0.y; wy[0][2] = y0.z; wy[0][3] = y0.w;
                wx[1][0] = x1.x; wx[1][1] = x1.y; wx[1][2] = x1.z; wx[1][3] = x1.w;
                wy[1][0] = y1.x; wy[1][1] = y1.y; wy[1][2] = y1.z; wy[1][3] = y1.w;
            }
            const float dw[2] = {__half2float(hA[b]), __half2float(hB[b])};
#pragma unroll
            for (int r = 0; r < TR; r++) {
                const XBlk& ab = st.act[wid * TR + r][b];
                const uint4 ax4 = *reinterpret_cast<const uint4*>(&ab.w[0]);
                const uint4 ay4 = *reinterpret_cast<const uint4*>(&ab.w[4]);
                const uint32_t ax[4] = {ax4.x, ax4.y, ax4.z, ax4.w}, ay[4] = {ay4.x, ay4.y, ay4.z, ay4.w};
                int ini[4] = {XB_BIAS, XB_BIAS, XB_BIAS, XB_BIAS};
                if (WT == DT_Q4) {
                    const int4 nbv = *reinterpret_cast<const int4*>(&ab.nb[0]);
                    ini[0] = nbv.x; ini[1] = nbv.y; ini[2] = nbv.z; ini[3] = nbv.w;
                }
                const float ad = ab.d;
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    // s = da * dw has <= 22 significant bits, so XB_M * s is exact and fmaf((M + l), s, -M s) is the
                    // single rounding of l * s: exactly `(float)lane * (da * dw)` of ops.h:282-287
                    const float s = __fmul_rn(ad, dw[c]);
                    const float ms = __fmul_rn(-XB_M, s);
                    const f32x2 ss = pk2(s, s), mm = pk2(ms, ms);
                    int t[4];
#pragma unroll
                    for (int l = 0; l < 4; l++) t[l] = __dp4a((int)wy[c][l], (int)ay[l], __dp4a((int)wx[c][l], (int)ax[l], ini[l]));
                    acc[r][c][0] = add2(acc[r][c][0], fma2(pk2(__int_as_float(t[0]), __int_as_float(t[1])), ss, mm));
                    acc[r][c][1] = add2(acc[r][c][1], fma2(pk2(__int_as_float(t[2]), __int_as_float(t[3])), ss, mm));
                }
            }
        }
    }
    cp_async_wait<0>();
    pdl_launch();            // the next kernel may become resident while this one runs its epilogue (not earlier: waiting CTAs hold SM resources)
    // ---------------- epilogue: lane = column inside each 32-block, warp = TR rows: every re-encode is warp-local
#pragma unroll
    for (int r = 0; r < TR; r++) {
        const int row = rbase + wid * TR + r;
        if (row >= rend) continue;                                   // warp-uniform
        const XrRow rw = a.rows[row];
        if (rw.slot < 0) continue;
        float v[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            float a0, a1, a2, a3;
            unpk2(acc[r][c][0], a0, a1); unpk2(acc[r][c][1], a2, a3);
            v[c] = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
        }
        xr_epilogue<EPI>(a, row, rw, n, lane, v, true);
    }
}

// ---------------------------------------------------------------- argmax over the tiles + bookkeeping (tinyllama.cpp:416-434)
#include "gtb_xtensor.cuh"

struct XrArgmaxArgs {
    const XrRow* rows;
    int row0;
    const float* arg_val; const int* arg_idx; int n_tiles;
    int32_t* tokens; int tok_stride;
    DevState* st;
    int eos_id;
};
__global__ void __launch_bounds__(128) k_xr_argmax(XrArgmaxArgs a) {
    __shared__ float sv[4];
    __shared__ int si[4];
    pdl_launch();
    pdl_wait();
    const int row = a.row0 + blockIdx.x;
    const XrRow rw = a.rows[row];
    if (rw.slot < 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int j = tid; j < a.n_tiles; j += 128) {
        const float ov = a.arg_val[(size_t)row * a.n_tiles + j];
        const int oi = a.arg_idx[(size_t)row * a.n_tiles + j];
        if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
    }
    if (lane == 0) { sv[wid] = best; si[wid] = arg; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 4; w++)
            if (sv[w] > best || (sv[w] == best && si[w] < arg)) { best = sv[w]; arg = si[w]; }
        if (arg == 0x7fffffff) arg = 0;                  // all -inf / NaN: the reference's loop leaves max_index = 0
        DevState& s = a.st[rw.slot];
        if (!s.stop) {
            a.tokens[(size_t)rw.slot * a.tok_stride + rw.pos + 1] = arg;
            s.pos = rw.pos + 1;
            s.n_gen += 1;
            if (arg == a.eos_id) s.stop = 1;
        }
    }
}

// ---------------------------------------------------------------- attention: CTA = (row, KV group), warp = query head
// gten/ops.h:930-1089 for one row: scores = Q8 dots with the group's cached K rows * 0.125, max, expf, the strictly
// in-order sum (ops.h:982-988), p = e / sum, the row re-encoded per 32 positions (ops.h:996), P.V with the eight
// position lanes of vec_dot_product_f32 (ops.h:181-197) and the tail after the lane sum, output re-encoded.
constexpr int XA_NT = 256, XA_KT = 128, XA_VT = 64, XA_VP = 68;     // V tile pitch in floats: 4 consecutive positions hit distinct banks

struct XrAttnArgs {
    const XrRow* rows;
    int row0;
    const XBlk* qst;           // [row][n_heads][2]
    const uint8_t* kq; const uint16_t* ks; const uint8_t* vq; const uint16_t* vs;
    size_t slot_codes, slot_scales;
    int kv_dim, n_heads;
    XBlk* out; int out_nb;     // staged attention output [row][n_embd / 32]
    int t_cap;                 // capacity of the score buffer in positions
};

__host__ __device__ inline size_t xr_attn_smem(int t_cap) {
    return (size_t)t_cap * 8 * 4 + 16384 + (size_t)XA_VT * XA_VP * 4 + 8 * 64 * 4 + 64;
}

__global__ void __launch_bounds__(XA_NT) k_xr_attn(XrAttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* sc = reinterpret_cast<float*>(smem);                                   // [position][8 heads]
    unsigned char* u = smem + (size_t)a.t_cap * 32;
    uint4 (*kt)[5] = reinterpret_cast<uint4 (*)[5]>(u);                           // K tile: 4 code words + scales, 80-byte pitch
    float (*part)[8][64] = reinterpret_cast<float (*)[8][64]>(u);                 // [lane][head][channel], after the K tiles are done
    float (*vt)[XA_VP] = reinterpret_cast<float (*)[XA_VP]>(u + 16384);
    float (*ob)[64] = reinterpret_cast<float (*)[64]>(u + 16384 + XA_VT * XA_VP * 4);
    float* sums = reinterpret_cast<float*>(u + 16384 + XA_VT * XA_VP * 4 + 8 * 64 * 4);
    const int g = blockIdx.x, row = a.row0 + blockIdx.y;
    pdl_wait();
    const XrRow rw = a.rows[row];
    if (rw.slot < 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t = rw.pos + 1, n_ctx = rw.n_ctx;
    const int nsc = a.kv_dim / 32;
    // ---- A: scores of head 8g + wid over positions [0, t)
    uint32_t q0[8], q1[8];
    float qd0, qd1;
    {
        const XBlk* qb = a.qst + ((size_t)row * a.n_heads + g * 8 + wid) * 2;
        const uint4 a0 = *reinterpret_cast<const uint4*>(&qb[0].w[0]), a1 = *reinterpret_cast<const uint4*>(&qb[0].w[4]);
        const uint4 b0 = *reinterpret_cast<const uint4*>(&qb[1].w[0]), b1 = *reinterpret_cast<const uint4*>(&qb[1].w[4]);
        q0[0] = a0.x; q0[1] = a0.y; q0[2] = a0.z; q0[3] = a0.w; q0[4] = a1.x; q0[5] = a1.y; q0[6] = a1.z; q0[7] = a1.w;
        q1[0] = b0.x; q1[1] = b0.y; q1[2] = b0.z; q1[3] = b0.w; q1[4] = b1.x; q1[5] = b1.y; q1[6] = b1.z; q1[7] = b1.w;
        qd0 = qb[0].d; qd1 = qb[1].d;
    }
    const uint8_t* kqb = a.kq + (size_t)rw.slot * a.slot_codes + g * 64;
    const uint16_t* ksb = a.ks + (size_t)rw.slot * a.slot_scales + g * 2;
    float mx = -INFINITY;
    for (int k0 = 0; k0 < t; k0 += XA_KT) {
        {
            const int p = tid >> 1, half = tid & 1, k = k0 + p;
            if (k < t) {
                const uint4* src = reinterpret_cast<const uint4*>(kqb + (size_t)k * a.kv_dim + half * 32);
                kt[p][half * 2] = __ldg(src);
                kt[p][half * 2 + 1] = __ldg(src + 1);
                if (half == 0) kt[p][4].x = __ldg(reinterpret_cast<const uint32_t*>(ksb + (size_t)k * nsc));
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < XA_KT / 32; j++) {
            const int p = lane + 32 * j, k = k0 + p;
            if (k < t) {
                const uint4 x0 = kt[p][0], y0 = kt[p][1], x1 = kt[p][2], y1 = kt[p][3];
                const uint32_t s2 = kt[p][4].x;
                const float s0 = __fmul_rn(qd0, h2f((uint16_t)(s2 & 0xffffu))), s1 = __fmul_rn(qd1, h2f((uint16_t)(s2 >> 16)));
                const float m0 = __fmul_rn(-XB_M, s0), m1 = __fmul_rn(-XB_M, s1);
                const uint32_t kx0[4] = {x0.x, x0.y, x0.z, x0.w}, ky0[4] = {y0.x, y0.y, y0.z, y0.w};
                const uint32_t kx1[4] = {x1.x, x1.y, x1.z, x1.w}, ky1[4] = {y1.x, y1.y, y1.z, y1.w};
                float al[4];
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    // gten/ops.h:224-292 over the two blocks of a 64-element head: acc = (0 + p0) + p1
                    const int t0 = __dp4a((int)ky0[l], (int)q0[4 + l], __dp4a((int)kx0[l], (int)q0[l], XB_BIAS));
                    const int t1 = __dp4a((int)ky1[l], (int)q1[4 + l], __dp4a((int)kx1[l], (int)q1[l], XB_BIAS));
                    al[l] = __fadd_rn(fmaf(__int_as_float(t0), s0, m0), fmaf(__int_as_float(t1), s1, m1));
                }
                const float s = __fmul_rn(__fadd_rn(__fadd_rn(al[0], al[1]), __fadd_rn(al[2], al[3])), 0.125f);
                sc[k * 8 + wid] = s;
                mx = fmaxf(mx, s);
            }
        }
        __syncthreads();
    }
    // ---- B: e = expf(s - max); the in-order sums of the 8 heads run side by side on 8 lanes of warp 0
    mx = warp_max(mx);
    for (int k = lane; k < t; k += 32) sc[k * 8 + wid] = expf_glibc(__fsub_rn(sc[k * 8 + wid], mx));
    __syncthreads();
    if (wid == 0 && lane < 8) {
        float s = 0.0f;
        int k = 0;
        for (; k + 8 <= t; k += 8) {
            float e[8];
#pragma unroll
            for (int uu = 0; uu < 8; uu++) e[uu] = sc[(k + uu) * 8 + lane];
#pragma unroll
            for (int uu = 0; uu < 8; uu++) s = __fadd_rn(s, e[uu]);
        }
        for (; k < t; k++) s = __fadd_rn(s, sc[k * 8 + lane]);
        sums[lane] = s;
    }
    __syncthreads();
    // ---- C: probabilities, re-encoded per 32 positions (masked entries are exact zeros and change nothing)
    {
        const float sum = sums[wid];
        const int nblk = (t + 31) / 32;
        int b = 0;
        for (; b + 2 <= nblk; b += 2) {             // two independent blocks in flight: the re-encode is a long dependent chain
            const int i = b * 32 + lane, j = i + 32;
            const float p0 = (i < t && i < n_ctx) ? __fdiv_rn(sc[i * 8 + wid], sum) : 0.0f;
            const float p1 = (j < t && j < n_ctx) ? __fdiv_rn(sc[j * 8 + wid], sum) : 0.0f;
            const float h0 = q8_roundtrip_lane(p0), h1 = q8_roundtrip_lane(p1);
            if (i < t) sc[i * 8 + wid] = h0;
            if (j < t) sc[j * 8 + wid] = h1;
        }
        for (; b < nblk; b++) {
            const int i = b * 32 + lane;
            const float p = (i < t && i < n_ctx) ? __fdiv_rn(sc[i * 8 + wid], sum) : 0.0f;
            const float ph = q8_roundtrip_lane(p);
            if (i < t) sc[i * 8 + wid] = ph;
        }
    }
    __syncthreads();
    // ---- D: P.V.  warp = 8 channels, lane = (head, position-lane pair {pl, pl + 4}): a V value is read once for 8 heads
    const int n8 = (n_ctx / 8) * 8, hi = min(n8, t);
    const int hh = lane & 7, pl = lane >> 3;
    f32x2 accv[2][4];                                     // channel pairs: packed fp32 (two independent rounded operations per instruction)
#pragma unroll
    for (int uu = 0; uu < 2; uu++)
#pragma unroll
        for (int c = 0; c < 4; c++) accv[uu][c] = pk2(0.0f, 0.0f);
    const uint8_t* vqb = a.vq + (size_t)rw.slot * a.slot_codes + g * 64;
    const uint16_t* vsb = a.vs + (size_t)rw.slot * a.slot_scales + g * 2;
    int i0_last = 0;
    for (int i0 = 0; i0 < t; i0 += XA_VT) {
        {
            const int p = tid >> 2, qt = tid & 3, i = i0 + p;
            if (i < t) {
                const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(vqb + (size_t)i * a.kv_dim + qt * 16));
                const float d = h2f(__ldg(vsb + (size_t)i * nsc + (qt >> 1)));
                const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
                float4* dst = reinterpret_cast<float4*>(&vt[p][qt * 16]);
#pragma unroll
                for (int uu = 0; uu < 4; uu++) {              // value = code * delta (ops.h:1026), exact
                    // int8 -> float without the conversion unit (a quarter-rate pipe): byte + 128 placed in the mantissa of 2^23
                    const uint32_t wv = cw[uu] ^ 0x80808080u;
                    const float f0 = __fsub_rn(__uint_as_float(__byte_perm(wv, 0x4b000000u, 0x7650)), 8388736.0f);
                    const float f1 = __fsub_rn(__uint_as_float(__byte_perm(wv, 0x4b000000u, 0x7651)), 8388736.0f);
                    const float f2 = __fsub_rn(__uint_as_float(__byte_perm(wv, 0x4b000000u, 0x7652)), 8388736.0f);
                    const float f3 = __fsub_rn(__uint_as_float(__byte_perm(wv, 0x4b000000u, 0x7653)), 8388736.0f);
                    dst[uu] = make_float4(__fmul_rn(f0, d), __fmul_rn(f1, d), __fmul_rn(f2, d), __fmul_rn(f3, d));
                }
            }
        }
        __syncthreads();
        const int lim = min(hi - i0, XA_VT);
        // one (position, 8 channels) step of this thread's chains: acc = (p * v) + acc per channel, both operations rounded (ops.h:181-197)
        auto pv_step = [&](int uu, float p, const float4& va, const float4& vb) {
            accv[uu][0] = add2(pk2(__fmul_rn(p, va.x), __fmul_rn(p, va.y)), accv[uu][0]);
            accv[uu][1] = add2(pk2(__fmul_rn(p, va.z), __fmul_rn(p, va.w)), accv[uu][1]);
            accv[uu][2] = add2(pk2(__fmul_rn(p, vb.x), __fmul_rn(p, vb.y)), accv[uu][2]);
            accv[uu][3] = add2(pk2(__fmul_rn(p, vb.z), __fmul_rn(p, vb.w)), accv[uu][3]);
        };
        int i8 = 0;
        // two octets per round, all twelve shared-memory loads issued before the arithmetic: with one or two CTAs per SM (decode batches)
        // nothing else hides their latency
        for (; i8 + 16 <= lim; i8 += 16) {
            float p[4];
            float4 va[4], vb[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int il = i8 + pl + 4 * q;               // q = 0, 1: first octet (lanes pl, pl + 4); q = 2, 3: second octet
                p[q] = sc[(i0 + il) * 8 + hh];
                va[q] = *reinterpret_cast<const float4*>(&vt[il][8 * wid]);
                vb[q] = *reinterpret_cast<const float4*>(&vt[il][8 * wid + 4]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) pv_step(q & 1, p[q], va[q], vb[q]);
        }
        for (; i8 < lim; i8 += 8) {
#pragma unroll
            for (int uu = 0; uu < 2; uu++) {
                const int il = i8 + pl + 4 * uu;
                if (il < lim) {
                    const float p = sc[(i0 + il) * 8 + hh];
                    const float4 va = *reinterpret_cast<const float4*>(&vt[il][8 * wid]);
                    const float4 vb = *reinterpret_cast<const float4*>(&vt[il][8 * wid + 4]);
                    pv_step(uu, p, va, vb);
                }
            }
        }
        i0_last = i0;
        if (i0 + XA_VT < t) __syncthreads();
    }
    pdl_launch();
#pragma unroll
    for (int uu = 0; uu < 2; uu++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float x0, x1;
            unpk2(accv[uu][c], x0, x1);
            part[pl + 4 * uu][hh][8 * wid + 2 * c] = x0; part[pl + 4 * uu][hh][8 * wid + 2 * c + 1] = x1;
        }
    __syncthreads();
    for (int o = tid; o < 512; o += XA_NT) {
        const int h2 = o >> 6, ch = o & 63;
        float d = __fadd_rn(part[0][h2][ch], part[1][h2][ch]);
#pragma unroll
        for (int l = 2; l < 8; l++) d = __fadd_rn(d, part[l][h2][ch]);
        for (int i = n8; i < t; i++) d = __fadd_rn(d, __fmul_rn(sc[i * 8 + h2], vt[i - i0_last][ch]));      // the tail sits in the last V tile
        ob[h2][ch] = d;
    }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint16_t dh;
        const int q = q8_encode_lane(ob[wid][half * 32 + lane], &dh);
        xblk_store(a.out + (size_t)row * a.out_nb + (g * 8 + wid) * 2 + half, lane, q, dh);
    }
}

// ---------------------------------------------------------------- attention, CTA = (row, query head): passes of few rows
// The group kernel above reads K/V once for 8 heads but gives a pass of R rows only 4 R CTAs; with R <= 32 rows at a long context
// that leaves most SMs idle behind one long CTA each.  This variant spreads the same arithmetic over 32 R CTAs (K/V are re-read
// per head from L2, harmless at this size): thread = position for the scores, (position lane, channel pair) for P.V.
struct XrHeadSmem {
    uint32_t qw[16]; float qd[2];
    float part[8][64];
    float red[8];
    float sum;
    float ob[64];
};
__host__ __device__ inline size_t xr_attn_head_smem(int t_cap) { return ((sizeof(XrHeadSmem) + 15) & ~(size_t)15) + (size_t)t_cap * 4; }

__global__ void __launch_bounds__(XA_NT) k_xr_attn_head(XrAttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    XrHeadSmem& sm = *reinterpret_cast<XrHeadSmem*>(smem);
    float* sc = reinterpret_cast<float*>(smem + ((sizeof(XrHeadSmem) + 15) & ~(size_t)15));
    const int h = blockIdx.x, g = h / 8, row = a.row0 + blockIdx.y;
    pdl_wait();
    const XrRow rw = a.rows[row];
    if (rw.slot < 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t = rw.pos + 1, n_ctx = rw.n_ctx;
    const int nsc = a.kv_dim / 32;
    if (tid < 18) {
        const XBlk* qb = a.qst + ((size_t)row * a.n_heads + h) * 2;
        if (tid < 16) sm.qw[tid] = qb[tid >> 3].w[tid & 7];
        else sm.qd[tid - 16] = qb[tid - 16].d;
    }
    __syncthreads();
    const uint8_t* kqb = a.kq + (size_t)rw.slot * a.slot_codes + g * 64;
    const uint16_t* ksb = a.ks + (size_t)rw.slot * a.slot_scales + g * 2;
    // ---- scores (gten/ops.h:224-292 over the two blocks of the head), scaled by 1/8
    float mx = -INFINITY;
    for (int k = tid; k < t; k += XA_NT) {
        const uint4* kp = reinterpret_cast<const uint4*>(kqb + (size_t)k * a.kv_dim);
        const uint4 x0 = __ldg(kp), y0 = __ldg(kp + 1), x1 = __ldg(kp + 2), y1 = __ldg(kp + 3);
        const uint32_t s2 = __ldg(reinterpret_cast<const uint32_t*>(ksb + (size_t)k * nsc));
        const float s0 = __fmul_rn(sm.qd[0], h2f((uint16_t)(s2 & 0xffffu))), s1 = __fmul_rn(sm.qd[1], h2f((uint16_t)(s2 >> 16)));
        const float m0 = __fmul_rn(-XB_M, s0), m1 = __fmul_rn(-XB_M, s1);
        const uint32_t kx0[4] = {x0.x, x0.y, x0.z, x0.w}, ky0[4] = {y0.x, y0.y, y0.z, y0.w};
        const uint32_t kx1[4] = {x1.x, x1.y, x1.z, x1.w}, ky1[4] = {y1.x, y1.y, y1.z, y1.w};
        float al[4];
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const int t0 = __dp4a((int)ky0[l], (int)sm.qw[4 + l], __dp4a((int)kx0[l], (int)sm.qw[l], XB_BIAS));
            const int t1 = __dp4a((int)ky1[l], (int)sm.qw[12 + l], __dp4a((int)kx1[l], (int)sm.qw[8 + l], XB_BIAS));
            al[l] = __fadd_rn(fmaf(__int_as_float(t0), s0, m0), fmaf(__int_as_float(t1), s1, m1));
        }
        const float s = __fmul_rn(__fadd_rn(__fadd_rn(al[0], al[1]), __fadd_rn(al[2], al[3])), 0.125f);
        sc[k] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) sm.red[wid] = mx;
    __syncthreads();
    mx = sm.red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, sm.red[w]);
    for (int k = tid; k < t; k += XA_NT) sc[k] = expf_glibc(__fsub_rn(sc[k], mx));
    __syncthreads();
    if (tid == 0) {                                        // ops.h:982-988: the sum strictly in order
        float s = 0.0f;
        int k = 0;
        for (; k + 8 <= t; k += 8) {
            const float4 e0 = *reinterpret_cast<const float4*>(sc + k), e1 = *reinterpret_cast<const float4*>(sc + k + 4);
            s = __fadd_rn(s, e0.x); s = __fadd_rn(s, e0.y); s = __fadd_rn(s, e0.z); s = __fadd_rn(s, e0.w);
            s = __fadd_rn(s, e1.x); s = __fadd_rn(s, e1.y); s = __fadd_rn(s, e1.z); s = __fadd_rn(s, e1.w);
        }
        for (; k < t; k++) s = __fadd_rn(s, sc[k]);
        sm.sum = s;
    }
    __syncthreads();
    {
        const float sum = sm.sum;
        const int nblk = (t + 31) / 32;
        for (int b = wid; b < nblk; b += 8) {
            const int i = b * 32 + lane;
            const float p = (i < t && i < n_ctx) ? __fdiv_rn(sc[i], sum) : 0.0f;
            const float ph = q8_roundtrip_lane(p);
            if (i < t) sc[i] = ph;
        }
    }
    __syncthreads();
    pdl_launch();
    // ---- P.V (ops.h:181-197, 1046-1087): warp = position lane (i mod 8), lane = channels (lane, lane + 32)
    const int n8 = (n_ctx / 8) * 8, hi = min(n8, t);
    const uint8_t* vqb = a.vq + (size_t)rw.slot * a.slot_codes + g * 64;
    const uint16_t* vsb = a.vs + (size_t)rw.slot * a.slot_scales + g * 2;
    {
        f32x2 a01 = pk2(0.0f, 0.0f);                       // the lane's two channels as one packed pair
        int i = wid;
        for (; i + 56 < hi; i += 64) {                     // eight positions of this lane per round: all loads first (L2 latency), then the chain
            uint32_t s2[8];
            int c0[8], c1[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int ii = i + 8 * u;
                s2[u] = __ldg(reinterpret_cast<const uint32_t*>(vsb + (size_t)ii * nsc));
                const uint8_t* vp = vqb + (size_t)ii * a.kv_dim;
                c0[u] = (int)(int8_t)__ldg(vp + lane);
                c1[u] = (int)(int8_t)__ldg(vp + 32 + lane);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float p = sc[i + 8 * u];
                // ops.h:1026; the int -> float step goes through the mantissa of 1.5 * 2^23 instead of the conversion unit
                const float v0 = __fmul_rn(__fsub_rn(__int_as_float(XB_BIAS + c0[u]), XB_M), h2f((uint16_t)(s2[u] & 0xffffu)));
                const float v1 = __fmul_rn(__fsub_rn(__int_as_float(XB_BIAS + c1[u]), XB_M), h2f((uint16_t)(s2[u] >> 16)));
                a01 = add2(pk2(__fmul_rn(p, v0), __fmul_rn(p, v1)), a01);
            }
        }
        for (; i < hi; i += 8) {
            const float p = sc[i];
            const uint32_t s2 = __ldg(reinterpret_cast<const uint32_t*>(vsb + (size_t)i * nsc));
            const uint8_t* vp = vqb + (size_t)i * a.kv_dim;
            const float v0 = __fmul_rn((float)(int8_t)__ldg(vp + lane), h2f((uint16_t)(s2 & 0xffffu)));
            const float v1 = __fmul_rn((float)(int8_t)__ldg(vp + 32 + lane), h2f((uint16_t)(s2 >> 16)));
            a01 = add2(pk2(__fmul_rn(p, v0), __fmul_rn(p, v1)), a01);
        }
        float a0, a1;
        unpk2(a01, a0, a1);
        sm.part[wid][lane] = a0;
        sm.part[wid][lane + 32] = a1;
    }
    __syncthreads();
    if (tid < 64) {
        float d = __fadd_rn(sm.part[0][tid], sm.part[1][tid]);
#pragma unroll
        for (int l = 2; l < 8; l++) d = __fadd_rn(d, sm.part[l][tid]);
        for (int i = n8; i < t; i++) {
            const float dl = h2f(__ldg(vsb + (size_t)i * nsc + (tid >> 5)));
            d = __fadd_rn(d, __fmul_rn(sc[i], __fmul_rn((float)(int8_t)__ldg(vqb + (size_t)i * a.kv_dim + tid), dl)));
        }
        sm.ob[tid] = d;
    }
    __syncthreads();
    if (wid < 2) {
        uint16_t dh;
        const int q = q8_encode_lane(sm.ob[wid * 32 + lane], &dh);
        xblk_store(a.out + (size_t)row * a.out_nb + h * 2 + wid, lane, q, dh);
    }
}

// ================================================================ FP16 models (FP16 weights, FP16 activations; tinyllama.cpp:258-260)
// Same pass structure; the arithmetic is gten/ops.h:140-160: eight lane accumulators over elements 8m + l (products of fp16 values are
// exact in fp32, so one fused multiply-add equals the reference's separate mul and add), lanes summed left to right, and every
// activation re-encode is a round-to-nearest-even to fp16.  Staged rows are the fp32 values of the fp16 activations in the weight
// layout's chunk order: element 64c + 8i + l at index (8c + l) * 8 + i, so the eight elements a lane needs are two 128-bit reads.
__device__ __forceinline__ int xf_index(int e) { return (((e >> 6) * 8) + (e & 7)) * 8 + ((e >> 3) & 7); }

struct XfNormArgs {
    const XrRow* rows;
    int row0;
    float* res;                // [rows][E]
    const uint16_t* normw;
    float* out;                // [rows][E] staged
    int E;
    const uint16_t* emb_w;     // device layout of the fp16 embedding table, or null
};

__global__ void __launch_bounds__(XN_NT) k_xf_norm(XfNormArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* xbuf = reinterpret_cast<float*>(smem);
    float* sq = xbuf + a.E;
    __shared__ ExactSum2Smem es;
    pdl_wait();
    const int row = a.row0 + blockIdx.x;
    const XrRow rw = a.rows[row];
    if (rw.slot < 0) return;
    float* res = a.res + (size_t)row * a.E;
    for (int e = threadIdx.x; e < a.E; e += XN_NT) {
        float v;
        if (a.emb_w) { v = h2f(a.emb_w[(size_t)rw.tok * a.E + xf_index(e)]); res[e] = v; }     // the fp16 row is copied (ops.h:514-564)
        else v = res[e];
        xbuf[e] = v;
        sq[e] = __fmul_rn(v, v);
    }
    __syncthreads();
    const float sq_sum = exact_sum512([&](int i, float q[4]) {
        if (i < a.E) { const float4 v = *reinterpret_cast<const float4*>(sq + i); q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w; }
        else { q[0] = q[1] = q[2] = q[3] = 0.0f; }
    }, a.E, es);
    pdl_launch();
    const float denom = __fadd_rn(sqrtf(__fdiv_rn(sq_sum, (float)a.E)), 1e-6f);
    float* out = a.out + (size_t)row * a.E;
    for (int e = threadIdx.x; e < a.E; e += XN_NT)
        out[xf_index(e)] = f16_roundtrip(__fmul_rn(__fdiv_rn(xbuf[e], denom), h2f(a.normw[e])));
}

constexpr int XF_KC = 2;                                        // 64-element chunks per pipeline stage
struct XfStage {
    uint4 wd[XG_BN][XF_KC * 8 + 1];                             // [column][chunk][lane]: 8 halves each; odd pitch
    float act[XG_BM][XF_KC * 64];
};

struct XfGemmArgs {
    const XrRow* rows;
    int row0, n_rows;
    const float* act; int K;           // staged input rows [row][K]
    const uint4* wd; int N;            // fp16 weights, device layout uint4[N][K/64][8]
    int up_off;
    float* res; int E;
    float* out; int out_ld;            // EPI_SILU: staged MLP rows [row][n_ffn]; EPI_QKV: q after RoPE [row][n_embd], natural order
    uint16_t* kq; uint16_t* vq; size_t slot_elems;
    int kv_dim, n_heads, n_groups;
    const float* rope_cos; const float* rope_sin;
    float* logits; int ld_logits; float* arg_val; int* arg_idx; int n_tiles;
};

template <int EPI>
__global__ void __launch_bounds__(256, 2) k_xf_gemm(XfGemmArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    XfStage* stages = reinterpret_cast<XfStage*>(smem);
    constexpr int NTH = 256, TR = 2;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = blockIdx.x;
    const int rbase = a.row0 + blockIdx.y * XG_BM, rend = a.row0 + a.n_rows;
    const int cpr = a.K / 64, nstage = cpr / XF_KC;
    auto wrow = [&](int c) -> int {
        int r = (EPI == XEPI_SILU) ? ((c < 32) ? 32 * n + c : a.up_off + 32 * n + c - 32) : 64 * n + c;
        return min(r, a.N - 1);
    };
    auto load_weights = [&](int s, int st_i) {
        XfStage& st = stages[s];
#pragma unroll
        for (int i = tid; i < XG_BN * XF_KC * 8; i += NTH) {
            const int c = i / (XF_KC * 8), j = i % (XF_KC * 8);
            cp_async16(&st.wd[c][j], a.wd + ((size_t)wrow(c) * cpr + (size_t)st_i * XF_KC) * 8 + j, true);
        }
    };
    auto load_act = [&](int s, int st_i) {
        XfStage& st = stages[s];
#pragma unroll
        for (int i = tid; i < XG_BM * XF_KC * 16; i += NTH) {
            const int r = i / (XF_KC * 16), j = i % (XF_KC * 16);
            const int gr = rbase + r;
            const bool ok = gr < rend;
            cp_async16(reinterpret_cast<uint4*>(&st.act[r][0]) + j,
                       reinterpret_cast<const uint4*>(a.act + (size_t)(ok ? gr : a.row0) * a.K + (size_t)st_i * XF_KC * 64) + j, ok);
        }
    };
    f32x2 acc[TR][8];                                       // per (row, lane): the two columns' chains as one packed pair
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
        for (int l = 0; l < 8; l++) acc[r][l] = pk2(0.0f, 0.0f);
#pragma unroll
    for (int s = 0; s < XG_STAGES - 1; s++)
        if (s < nstage) load_weights(s, s);
    pdl_wait();
#pragma unroll
    for (int s = 0; s < XG_STAGES - 1; s++) {
        if (s < nstage) load_act(s, s);
        cp_async_commit();
    }
    for (int si = 0; si < nstage; si++) {
        cp_async_wait<XG_STAGES - 2>();
        __syncthreads();
        if (si + XG_STAGES - 1 < nstage) {
            load_weights((si + XG_STAGES - 1) % XG_STAGES, si + XG_STAGES - 1);
            load_act((si + XG_STAGES - 1) % XG_STAGES, si + XG_STAGES - 1);
        }
        cp_async_commit();
        const XfStage& st = stages[si % XG_STAGES];
#pragma unroll
        for (int c = 0; c < XF_KC; c++) {
#pragma unroll
            for (int l = 0; l < 8; l++) {
                const uint4 w0 = st.wd[lane][c * 8 + l], w1 = st.wd[lane + 32][c * 8 + l];
                const __half2* h0 = reinterpret_cast<const __half2*>(&w0);
                const __half2* h1 = reinterpret_cast<const __half2*>(&w1);
                float f0[8], f1[8];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float2 p = __half22float2(h0[u]), q = __half22float2(h1[u]);
                    f0[2 * u] = p.x; f0[2 * u + 1] = p.y; f1[2 * u] = q.x; f1[2 * u + 1] = q.y;
                }
#pragma unroll
                for (int r = 0; r < TR; r++) {
                    const float4 xa = *reinterpret_cast<const float4*>(&st.act[wid * TR + r][(c * 8 + l) * 8]);
                    const float4 xb = *reinterpret_cast<const float4*>(&st.act[wid * TR + r][(c * 8 + l) * 8 + 4]);
                    const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                    for (int i = 0; i < 8; i++)                           // elements 64c + 8i + l, ascending i: the lane's own order;
                        acc[r][l] = fma2(pk2(x[i], x[i]), pk2(f0[i], f1[i]), acc[r][l]);   // one FFMA2 = the two columns' fused multiply-adds
                }
            }
        }
    }
    cp_async_wait<0>();
    pdl_launch();
#pragma unroll
    for (int r = 0; r < TR; r++) {
        const int row = rbase + wid * TR + r;
        if (row >= rend) continue;
        const XrRow rw = a.rows[row];
        if (rw.slot < 0) continue;
        float v[2];
        {
            f32x2 d = add2(acc[r][0], acc[r][1]);                       // simd_ops.h:63-66: lanes left to right (both columns at once)
#pragma unroll
            for (int l = 2; l < 8; l++) d = add2(d, acc[r][l]);
            unpk2(d, v[0], v[1]);
        }
        if (EPI == XEPI_RES) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                float* px = a.res + (size_t)row * a.E + 64 * n + 32 * c + lane;
                *px = f16_roundtrip(__fadd_rn(*px, f16_roundtrip(v[c])));
            }
        } else if (EPI == XEPI_SILU) {
            const float g1 = f16_roundtrip(v[0]), u1 = f16_roundtrip(v[1]);
            const float g2 = f16_roundtrip(silu_ref(g1));
            a.out[(size_t)row * a.out_ld + xf_index(32 * n + lane)] = f16_roundtrip(__fmul_rn(g2, u1));
        } else if (EPI == XEPI_QKV) {
            const int pos = rw.pos;
            if (n < a.n_heads + a.n_groups) {
                const float x0 = f16_roundtrip(v[0]), x1 = f16_roundtrip(v[1]);
                const float cs = __ldg(a.rope_cos + (size_t)pos * 32 + lane), sn = __ldg(a.rope_sin + (size_t)pos * 32 + lane);
                const uint16_t o0 = f2h(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn)));
                const uint16_t o1 = f2h(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs)));
                if (n < a.n_heads) {
                    float* q = a.out + (size_t)row * a.out_ld + n * 64;
                    q[lane] = h2f(o0); q[32 + lane] = h2f(o1);
                } else {
                    uint16_t* k = a.kq + (size_t)rw.slot * a.slot_elems + (size_t)pos * a.kv_dim + (n - a.n_heads) * 64;
                    k[lane] = o0; k[32 + lane] = o1;
                }
            } else {
                uint16_t* vv = a.vq + (size_t)rw.slot * a.slot_elems + (size_t)pos * a.kv_dim + (n - a.n_heads - a.n_groups) * 64;
                vv[lane] = f2h(v[0]); vv[32 + lane] = f2h(v[1]);
            }
        } else {
            const int c0 = 64 * n + lane, c1 = c0 + 32;
            float best = -INFINITY;
            int arg = 0x7fffffff;
            if (c0 < a.N) { a.logits[(size_t)(row - a.row0) * a.ld_logits + c0] = v[0]; if (v[0] > best) { best = v[0]; arg = c0; } }
            if (c1 < a.N) { a.logits[(size_t)(row - a.row0) * a.ld_logits + c1] = v[1]; if (v[1] > best) { best = v[1]; arg = c1; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
            }
            if (lane == 0) { a.arg_val[(size_t)row * a.n_tiles + n] = best; a.arg_idx[(size_t)row * a.n_tiles + n] = arg; }
        }
    }
}

struct XfAttnArgs {
    const XrRow* rows;
    int row0;
    const float* q;            // [row][n_embd] after RoPE, natural order
    const uint16_t* kq; const uint16_t* vq; size_t slot_elems;
    int kv_dim, n_heads;
    float* out; int out_ld;    // staged attention output [row][n_embd]
    int t_cap;
};
struct XfHeadSmem { float qf[64]; float part[8][64]; float red[8]; float sum; };
__host__ __device__ inline size_t xf_attn_smem(int t_cap) { return ((sizeof(XfHeadSmem) + 15) & ~(size_t)15) + (size_t)t_cap * 4; }

__global__ void __launch_bounds__(XA_NT) k_xf_attn_head(XfAttnArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    XfHeadSmem& sm = *reinterpret_cast<XfHeadSmem*>(smem);
    float* sc = reinterpret_cast<float*>(smem + ((sizeof(XfHeadSmem) + 15) & ~(size_t)15));
    const int h = blockIdx.x, g = h / 8, row = a.row0 + blockIdx.y;
    pdl_wait();
    const XrRow rw = a.rows[row];
    if (rw.slot < 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int t = rw.pos + 1, n_ctx = rw.n_ctx;
    if (tid < 64) sm.qf[tid] = a.q[(size_t)row * (a.n_heads * 64) + h * 64 + tid];
    __syncthreads();
    const uint16_t* kb = a.kq + (size_t)rw.slot * a.slot_elems + g * 64;
    const uint16_t* vb = a.vq + (size_t)rw.slot * a.slot_elems + g * 64;
    float mx = -INFINITY;
    for (int k = tid; k < t; k += XA_NT) {
        const uint4* kp = reinterpret_cast<const uint4*>(kb + (size_t)k * a.kv_dim);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 8; i++) {                          // ops.h:140-160: lane l accumulates elements 8 i + l
            const uint4 w = __ldg(kp + i);
            const __half2* hw = reinterpret_cast<const __half2*>(&w);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float2 f = __half22float2(hw[u]);
                acc[2 * u] = fmaf(sm.qf[8 * i + 2 * u], f.x, acc[2 * u]);
                acc[2 * u + 1] = fmaf(sm.qf[8 * i + 2 * u + 1], f.y, acc[2 * u + 1]);
            }
        }
        float d = __fadd_rn(acc[0], acc[1]);
#pragma unroll
        for (int l = 2; l < 8; l++) d = __fadd_rn(d, acc[l]);
        const float s = __fmul_rn(d, 0.125f);
        sc[k] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) sm.red[wid] = mx;
    __syncthreads();
    mx = sm.red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, sm.red[w]);
    for (int k = tid; k < t; k += XA_NT) sc[k] = expf_glibc(__fsub_rn(sc[k], mx));
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        for (int k = 0; k < t; k++) s = __fadd_rn(s, sc[k]);
        sm.sum = s;
    }
    __syncthreads();
    {
        const float sum = sm.sum;
        for (int k = tid; k < t; k += XA_NT) sc[k] = f16_roundtrip((k < n_ctx) ? __fdiv_rn(sc[k], sum) : 0.0f);
    }
    __syncthreads();
    pdl_launch();
    const int n8 = (n_ctx / 8) * 8, hi = min(n8, t);
    {
        float a0 = 0.0f, a1 = 0.0f;
        for (int i = wid; i < hi; i += 8) {
            const float p = sc[i];
            const uint16_t* vp = vb + (size_t)i * a.kv_dim;
            a0 = __fadd_rn(__fmul_rn(p, h2f(__ldg(vp + lane))), a0);
            a1 = __fadd_rn(__fmul_rn(p, h2f(__ldg(vp + 32 + lane))), a1);
        }
        sm.part[wid][lane] = a0;
        sm.part[wid][lane + 32] = a1;
    }
    __syncthreads();
    if (tid < 64) {
        float d = __fadd_rn(sm.part[0][tid], sm.part[1][tid]);
#pragma unroll
        for (int l = 2; l < 8; l++) d = __fadd_rn(d, sm.part[l][tid]);
        for (int i = n8; i < t; i++) d = __fadd_rn(d, __fmul_rn(sc[i], h2f(__ldg(vb + (size_t)i * a.kv_dim + tid))));
        a.out[(size_t)row * a.out_ld + xf_index(h * 64 + tid)] = f16_roundtrip(d);
    }
}

// ---------------------------------------------------------------- host side
// Programmatic dependent launch between the kernels of a pass.  Measured (tools/xrows_probe.py --pdl 0/1, same box): it helps the
// latency-bound small passes (8 rows: 2.66 vs 2.76 ms per step) and costs 17 % on full ones (64 rows: 4.58 vs 3.91 ms: the early
// resident CTAs of the next kernel take SM resources from the tail of the running one), so it is used for passes of <= 16 rows.
static bool g_xr_pdl = true;
static bool g_xr_pdl_now = false;
static int g_xr_variant = 0;        // experiment switch for the large-N GEMM configuration
static bool g_xr_tensor = false;    // opt-in experiment: Q4 / Q8 Linears on the tensor cores (gtb_xtensor.cuh).  Bit-identical, but end to end
                                    // slower than the SIMT kernel k_xr_gemm (profiles/r02_03_tensor_exact.md), which stays the default

struct XrPlan {
    gtb_model_config cfg{};
    XrRow* rows = nullptr;
    float* res = nullptr;
    XBlk *act_norm = nullptr, *act_attn = nullptr, *act_mlp = nullptr, *qst = nullptr;
    float* arg_val = nullptr;
    int* arg_idx = nullptr;
    float* logits = nullptr;           // [XR_MAX_ROWS][n_vocab] when the caller passes none
    float *f_norm = nullptr, *f_attn = nullptr, *f_mlp = nullptr, *f_q = nullptr;   // FP16 models: staged fp32 rows
    float* raw = nullptr; int ld_raw = 0;   // tensor-core GEMM: finished fp32 rows of one Linear [XR_MAX_ROWS][ld_raw]
    int n_tiles = 0;
    size_t bytes = 0;
};

void xr_set_pdl(bool on) { g_xr_pdl = on; }
void xr_set_variant(int v) { g_xr_variant = v; }
void xr_set_tensor(bool on) { g_xr_tensor = on; }
static long long* g_xr_trace = nullptr;
void xr_set_trace(long long* d_buf) { g_xr_trace = d_buf; }

bool xr_supported(const gtb_model_config& c, int gsz) {
    return (c.wdtype == GTB_Q8 || c.wdtype == GTB_Q4 || c.wdtype == GTB_F16) && gsz == 8 && c.n_embd % 256 == 0 && c.n_ffn % 256 == 0 && c.n_embd / c.n_heads == 64 &&
           c.n_embd % 16 == 0 && c.n_embd <= 4096;
}

int xr_create(XrPlan** out, const gtb_model_config& c) {
    auto* p = new XrPlan();
    p->cfg = c;
    p->n_tiles = (c.n_vocab + 63) / 64;
    const size_t R = XR_MAX_ROWS;
    auto dalloc = [&](void** q, size_t n) -> int {
        GTB_CUDA(cudaMalloc(q, n));
        GTB_CUDA(cudaMemsetAsync(*q, 0, n, ctx().stream));
        ctx().mem += (int64_t)n; p->bytes += n;
        return GTB_OK;
    };
    int r = 0;
    r |= dalloc((void**)&p->rows, R * sizeof(XrRow));
    r |= dalloc((void**)&p->res, R * c.n_embd * 4);
    if (c.wdtype == GTB_F16) {
        r |= dalloc((void**)&p->f_norm, R * c.n_embd * 4); r |= dalloc((void**)&p->f_attn, R * c.n_embd * 4);
        r |= dalloc((void**)&p->f_mlp, R * c.n_ffn * 4); r |= dalloc((void**)&p->f_q, R * c.n_embd * 4);
    } else {
        r |= dalloc((void**)&p->act_norm, R * (c.n_embd / 32) * sizeof(XBlk));
        r |= dalloc((void**)&p->act_attn, R * (c.n_embd / 32) * sizeof(XBlk));
        r |= dalloc((void**)&p->act_mlp, R * (c.n_ffn / 32) * sizeof(XBlk));
        r |= dalloc((void**)&p->qst, R * c.n_heads * 2 * sizeof(XBlk));
        p->ld_raw = std::max(2 * c.n_ffn, c.n_embd + 2 * 64 * c.n_groups);
        r |= dalloc((void**)&p->raw, R * (size_t)p->ld_raw * 4);
    }
    r |= dalloc((void**)&p->arg_val, R * p->n_tiles * 4);
    r |= dalloc((void**)&p->arg_idx, R * p->n_tiles * 4);
    r |= dalloc((void**)&p->logits, (size_t)XR_MAX_SLOTS * c.n_vocab * 4);      // only when a caller passes no logits buffer
    if (r) { xr_destroy(p); return fail(GTB_ERR_CUDA, "multi-row buffers: allocation failed"); }
    *out = p;
    return GTB_OK;
}

void xr_destroy(XrPlan* p) {
    if (!p) return;
    void* b[] = {p->rows, p->res, p->act_norm, p->act_attn, p->act_mlp, p->qst, p->arg_val, p->arg_idx, p->logits, p->f_norm, p->f_attn, p->f_mlp, p->f_q, p->raw};
    for (void* q : b) cudaFree(q);
    ctx().mem -= (int64_t)p->bytes;
    delete p;
}

namespace {

template <typename... KArgs, typename... Args>
cudaError_t xr_launch(void (*kern)(KArgs...), dim3 grid, int block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_xr_pdl_now ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <int WT, int EPI, int NW, int NS = XG_STAGES, int MINB = (NW == 4) ? 4 : 2>
int launch_gemm_nw(const XrGemmArgs& a, int n_tiles) {
    const size_t smem = sizeof(XgStage<WT>) * NS;
    static bool attr = false;
    if (!attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_xr_gemm<WT, EPI, NW, NS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid(n_tiles, (a.n_rows + XG_BM - 1) / XG_BM);
    GTB_CUDA(xr_launch(k_xr_gemm<WT, EPI, NW, NS, MINB>, grid, NW * 32, smem, a));
    GTB_LAUNCHED();
    return GTB_OK;
}
template <int WT, int EPI>
int launch_gemm(const XrGemmArgs& a, int n_tiles) {
    // few column tiles: 8 warps x 2 rows per tile keep two warps per scheduler busy; many tiles: 4 warps x 4 rows
    const int ctas = n_tiles * ((a.n_rows + XG_BM - 1) / XG_BM);
    if (EPI != XEPI_SILU && EPI != XEPI_HEAD && ctas <= 2 * ctx().sm_count) return launch_gemm_nw<WT, EPI, 8>(a, n_tiles);
    if (g_xr_variant == 1 && (EPI == XEPI_SILU || EPI == XEPI_HEAD)) return launch_gemm_nw<WT, EPI, 4, 2, 5>(a, n_tiles);   // 2 stages, 5 CTAs / SM
    return launch_gemm_nw<WT, EPI, 4>(a, n_tiles);
}

// ---- tensor-core Linear (gtb_xtensor.cuh): rows per thread chosen so that the (tile, row group) units fill the SMs
template <int WT, int RPT>
int launch_xt_rpt(const XtGemmArgs& a) {
    using Cfg = XtCfg<WT, RPT>;
    static bool attr = false;
    if (!attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_xt_gemm<WT, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr = true;
    }
    dim3 grid((a.N + XT_BM - 1) / XT_BM, (a.n_rows + Cfg::ROWS - 1) / Cfg::ROWS);
    GTB_CUDA(xr_launch(k_xt_gemm<WT, RPT>, grid, XT_NT, Cfg::SMEM, a));
    GTB_LAUNCHED();
    return GTB_OK;
}
template <int WT>
int launch_xt(const XtGemmArgs& a) {
    const int tiles = (a.N + XT_BM - 1) / XT_BM, sms = ctx().sm_count;
    int best = 8;
    long best_cost = -1;
    for (int rpt = 8; rpt >= 2; rpt >>= 1) {
        const long units = (long)tiles * ((a.n_rows + 4 * rpt - 1) / (4 * rpt));
        const long cost = ((units + sms - 1) / sms) * (90 + 60 * rpt);       // cycles per K block: fixed part + per row of a thread
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rpt; }
    }
    if (g_xr_variant >= 2) best = g_xr_variant;                               // tuning experiments: force the shape
    switch (best) {
        case 8: return launch_xt_rpt<WT, 8>(a);
        case 4: return launch_xt_rpt<WT, 4>(a);
        default: return launch_xt_rpt<WT, 2>(a);
    }
}
// one Linear of the pass: tensor-core GEMM + re-encode epilogue kernel, or the SIMT kernel with the fused epilogue
template <int WT, int EPI>
int xr_linear(XrPlan* p, const XrGemmArgs& a, int n_tiles) {
    if (!g_xr_tensor) return launch_gemm<WT, EPI>(a, n_tiles);
    XtGemmArgs t{};
    t.act = a.act; t.nb = a.nb; t.wd = a.wd; t.ws = a.ws; t.N = a.N; t.row0 = a.row0; t.n_rows = a.n_rows; t.dbg = g_xr_trace;
    if (EPI == XEPI_HEAD) { t.out = a.logits; t.ldo = a.ld_logits; t.out_row_sub = a.row0; }
    else { t.out = p->raw; t.ldo = p->ld_raw; t.out_row_sub = 0; }
    int r = launch_xt<WT>(t);
    if (r) return r;
    GTB_CUDA(xr_launch(k_xt_epi<EPI>, dim3(n_tiles, (a.n_rows + 7) / 8), 256, 0, a, (const float*)t.out, t.ldo));
    GTB_LAUNCHED();
    return GTB_OK;
}

template <int WT>
int run_pass(XrPlan* p, const XrModel& m, const XrKV& kv, const XrSeq& sq, const XrPlanArgs& plan, int t_cap, int head_row0,
             int head_rows, int eos_id, float* d_logits) {
    const gtb_model_config& c = m.cfg;
    const int E = c.n_embd, F = c.n_ffn, KVD = 64 * c.n_groups, R = plan.n_rows;
    g_xr_pdl_now = g_xr_pdl && R <= 16 && !g_xr_tensor;      // the tensor-core kernels do not carry the dependent-launch protocol
    GTB_CUDA(xr_launch(k_xr_plan, dim3(1), XR_MAX_ROWS, 0, plan));
    GTB_LAUNCHED();
    const size_t norm_smem = (size_t)E * 8;
    const size_t attn_smem = xr_attn_smem(t_cap);
    static bool attr = false;
    if (!attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_xr_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        GTB_CUDA(cudaFuncSetAttribute(k_xr_attn_head, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        GTB_CUDA(cudaFuncSetAttribute(k_xr_norm, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    if (attn_smem > 200 * 1024) return fail(GTB_ERR_STATE, "multi-row attention: context too long for the score buffer");
    auto norm = [&](const uint16_t* w, bool emb, int row0, int rows) -> int {
        XrNormArgs a{};
        a.rows = p->rows; a.row0 = row0; a.res = p->res; a.normw = w; a.out = p->act_norm; a.E = E;
        if (emb) { a.emb_w = m.emb_w; a.emb_s = m.emb_s; a.emb_dt = (c.wdtype == GTB_Q8) ? DT_Q8 : DT_Q4; }
        GTB_CUDA(xr_launch(k_xr_norm, dim3(rows), XN_NT, norm_smem, a));
        GTB_LAUNCHED();
        return GTB_OK;
    };
    XrGemmArgs base{};
    base.rows = p->rows; base.row0 = 0; base.n_rows = R; base.res = p->res; base.E = E;
    base.slot_codes = kv.slot_codes; base.slot_scales = kv.slot_scales; base.kv_dim = KVD; base.n_heads = c.n_heads; base.n_groups = c.n_groups;
    base.rope_cos = m.rope_cos; base.rope_sin = m.rope_sin;
    int r;
    for (int li = 0; li < c.n_layers; li++) {
        const XrLayerW& L = m.layers[li];
        if ((r = norm(L.attn_norm, li == 0, 0, R))) return r;
        {
            XrGemmArgs a = base;
            a.act = p->act_norm; a.nb = E / 32; a.wd = L.w[0]; a.ws = L.s[0]; a.N = E + 2 * KVD;
            a.out = p->qst; a.kq = kv.kq[li]; a.ks = kv.ks[li]; a.vq = kv.vq[li]; a.vs = kv.vs[li];
            if ((r = xr_linear<WT, XEPI_QKV>(p, a, a.N / 64))) return r;
        }
        {
            XrAttnArgs a{};
            a.rows = p->rows; a.row0 = 0; a.qst = p->qst; a.kq = kv.kq[li]; a.ks = kv.ks[li]; a.vq = kv.vq[li]; a.vs = kv.vs[li];
            a.slot_codes = kv.slot_codes; a.slot_scales = kv.slot_scales; a.kv_dim = KVD; a.n_heads = c.n_heads;
            a.out = p->act_attn; a.out_nb = E / 32; a.t_cap = t_cap;
            if (R <= 32) GTB_CUDA(xr_launch(k_xr_attn_head, dim3(c.n_heads, R), XA_NT, xr_attn_head_smem(t_cap), a));
            else GTB_CUDA(xr_launch(k_xr_attn, dim3(c.n_groups, R), XA_NT, attn_smem, a));
            GTB_LAUNCHED();
        }
        {
            XrGemmArgs a = base;
            a.act = p->act_attn; a.nb = E / 32; a.wd = L.w[1]; a.ws = L.s[1]; a.N = E;
            if ((r = xr_linear<WT, XEPI_RES>(p, a, E / 64))) return r;
        }
        if ((r = norm(L.ffn_norm, false, 0, R))) return r;
        {
            XrGemmArgs a = base;
            a.act = p->act_norm; a.nb = E / 32; a.wd = L.w[2]; a.ws = L.s[2]; a.N = 2 * F; a.up_off = F;
            a.out = p->act_mlp; a.out_nb = F / 32;
            if ((r = xr_linear<WT, XEPI_SILU>(p, a, F / 32))) return r;
        }
        {
            XrGemmArgs a = base;
            a.act = p->act_mlp; a.nb = F / 32; a.wd = L.w[3]; a.ws = L.s[3]; a.N = E;
            if ((r = xr_linear<WT, XEPI_RES>(p, a, E / 64))) return r;
        }
    }
    if (head_rows > 0) {
        if ((r = norm(m.final_norm, false, head_row0, head_rows))) return r;
        float* lg = d_logits ? d_logits : p->logits;
        XrGemmArgs a = base;
        a.row0 = head_row0; a.n_rows = head_rows;
        a.act = p->act_norm; a.nb = E / 32; a.wd = m.head_w; a.ws = m.head_s; a.N = c.n_vocab;
        a.logits = lg; a.ld_logits = c.n_vocab;          // logits of row (head_row0 + i) land in row i of the buffer
        a.arg_val = p->arg_val; a.arg_idx = p->arg_idx; a.n_tiles = p->n_tiles;
        if ((r = xr_linear<WT, XEPI_HEAD>(p, a, p->n_tiles))) return r;
        XrArgmaxArgs g{};
        g.rows = p->rows; g.row0 = head_row0; g.arg_val = p->arg_val; g.arg_idx = p->arg_idx; g.n_tiles = p->n_tiles;
        g.tokens = sq.tokens; g.tok_stride = sq.tok_stride; g.st = sq.st; g.eos_id = eos_id;
        GTB_CUDA(xr_launch(k_xr_argmax, dim3(head_rows), 128, 0, g));
        GTB_LAUNCHED();
    }
    return GTB_OK;
}

template <int EPI>
int launch_gemm_f16(const XfGemmArgs& a, int n_tiles) {
    const size_t smem = sizeof(XfStage) * XG_STAGES;
    static bool attr = false;
    if (!attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_xf_gemm<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid(n_tiles, (a.n_rows + XG_BM - 1) / XG_BM);
    GTB_CUDA(xr_launch(k_xf_gemm<EPI>, grid, 256, smem, a));
    GTB_LAUNCHED();
    return GTB_OK;
}

// one pass of an FP16 model (same structure as run_pass)
int run_pass_f16(XrPlan* p, const XrModel& m, const XrKV& kv, const XrSeq& sq, const XrPlanArgs& plan, int t_cap, int head_row0,
                 int head_rows, int eos_id, float* d_logits) {
    const gtb_model_config& c = m.cfg;
    const int E = c.n_embd, F = c.n_ffn, KVD = 64 * c.n_groups, R = plan.n_rows;
    g_xr_pdl_now = g_xr_pdl && R <= 16;
    GTB_CUDA(xr_launch(k_xr_plan, dim3(1), XR_MAX_ROWS, 0, plan));
    GTB_LAUNCHED();
    const size_t norm_smem = (size_t)E * 8;
    const size_t attn_smem = xf_attn_smem(t_cap);
    static bool attr = false;
    if (!attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_xf_attn_head, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        GTB_CUDA(cudaFuncSetAttribute(k_xf_norm, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    if (attn_smem > 64 * 1024) return fail(GTB_ERR_STATE, "multi-row attention: context too long for the score buffer");
    auto norm = [&](const uint16_t* w, bool emb, int row0, int rows) -> int {
        XfNormArgs a{};
        a.rows = p->rows; a.row0 = row0; a.res = p->res; a.normw = w; a.out = p->f_norm; a.E = E;
        a.emb_w = emb ? reinterpret_cast<const uint16_t*>(m.emb_w) : nullptr;
        GTB_CUDA(xr_launch(k_xf_norm, dim3(rows), XN_NT, norm_smem, a));
        GTB_LAUNCHED();
        return GTB_OK;
    };
    XfGemmArgs base{};
    base.rows = p->rows; base.row0 = 0; base.n_rows = R; base.res = p->res; base.E = E;
    base.slot_elems = kv.slot_codes; base.kv_dim = KVD; base.n_heads = c.n_heads; base.n_groups = c.n_groups;
    base.rope_cos = m.rope_cos; base.rope_sin = m.rope_sin;
    int r;
    for (int li = 0; li < c.n_layers; li++) {
        const XrLayerW& L = m.layers[li];
        if ((r = norm(L.attn_norm, li == 0, 0, R))) return r;
        {
            XfGemmArgs a = base;
            a.act = p->f_norm; a.K = E; a.wd = L.w[0]; a.N = E + 2 * KVD; a.out = p->f_q; a.out_ld = E;
            a.kq = reinterpret_cast<uint16_t*>(kv.kq[li]); a.vq = reinterpret_cast<uint16_t*>(kv.vq[li]);
            if ((r = launch_gemm_f16<XEPI_QKV>(a, a.N / 64))) return r;
        }
        {
            XfAttnArgs a{};
            a.rows = p->rows; a.row0 = 0; a.q = p->f_q; a.kq = reinterpret_cast<const uint16_t*>(kv.kq[li]);
            a.vq = reinterpret_cast<const uint16_t*>(kv.vq[li]); a.slot_elems = kv.slot_codes; a.kv_dim = KVD; a.n_heads = c.n_heads;
            a.out = p->f_attn; a.out_ld = E; a.t_cap = t_cap;
            GTB_CUDA(xr_launch(k_xf_attn_head, dim3(c.n_heads, R), XA_NT, attn_smem, a));
            GTB_LAUNCHED();
        }
        {
            XfGemmArgs a = base;
            a.act = p->f_attn; a.K = E; a.wd = L.w[1]; a.N = E;
            if ((r = launch_gemm_f16<XEPI_RES>(a, E / 64))) return r;
        }
        if ((r = norm(L.ffn_norm, false, 0, R))) return r;
        {
            XfGemmArgs a = base;
            a.act = p->f_norm; a.K = E; a.wd = L.w[2]; a.N = 2 * F; a.up_off = F; a.out = p->f_mlp; a.out_ld = F;
            if ((r = launch_gemm_f16<XEPI_SILU>(a, F / 32))) return r;
        }
        {
            XfGemmArgs a = base;
            a.act = p->f_mlp; a.K = F; a.wd = L.w[3]; a.N = E;
            if ((r = launch_gemm_f16<XEPI_RES>(a, E / 64))) return r;
        }
    }
    if (head_rows > 0) {
        if ((r = norm(m.final_norm, false, head_row0, head_rows))) return r;
        XfGemmArgs a = base;
        a.row0 = head_row0; a.n_rows = head_rows;
        a.act = p->f_norm; a.K = E; a.wd = m.head_w; a.N = c.n_vocab;
        a.logits = d_logits ? d_logits : p->logits; a.ld_logits = c.n_vocab;
        a.arg_val = p->arg_val; a.arg_idx = p->arg_idx; a.n_tiles = p->n_tiles;
        if ((r = launch_gemm_f16<XEPI_HEAD>(a, p->n_tiles))) return r;
        XrArgmaxArgs g{};
        g.rows = p->rows; g.row0 = head_row0; g.arg_val = p->arg_val; g.arg_idx = p->arg_idx; g.n_tiles = p->n_tiles;
        g.tokens = sq.tokens; g.tok_stride = sq.tok_stride; g.st = sq.st; g.eos_id = eos_id;
        GTB_CUDA(xr_launch(k_xr_argmax, dim3(head_rows), 128, 0, g));
        GTB_LAUNCHED();
    }
    return GTB_OK;
}

}  // namespace

int xr_prefill_pass(XrPlan* p, const XrModel& m, const XrKV& kv, const XrSeq& sq, int slot, int p0, int n_rows, int n_ctx,
                    bool with_head, int eos_id, float* d_logits) {
    if (n_rows <= 0 || n_rows > XR_MAX_ROWS) return fail(GTB_ERR_ARG, "multi-row pass: 1..%d rows", XR_MAX_ROWS);
    XrPlanArgs plan{};
    plan.mode = 0; plan.n_rows = n_rows; plan.slot0 = slot; plan.p0 = p0; plan.n_ctx = n_ctx; plan.tok_stride = sq.tok_stride;
    plan.tokens = sq.tokens; plan.st = sq.st; plan.rows = p->rows;
    const int t_cap = ((p0 + n_rows + 31) / 32) * 32 + 32;
    const int hr = with_head ? 1 : 0;
    if (m.cfg.wdtype == GTB_F16) return run_pass_f16(p, m, kv, sq, plan, t_cap, n_rows - 1, hr, eos_id, d_logits);
    return (m.cfg.wdtype == GTB_Q8) ? run_pass<DT_Q8>(p, m, kv, sq, plan, t_cap, n_rows - 1, hr, eos_id, d_logits)
                                    : run_pass<DT_Q4>(p, m, kv, sq, plan, t_cap, n_rows - 1, hr, eos_id, d_logits);
}

int xr_decode_pass(XrPlan* p, const XrModel& m, const XrKV& kv, const XrSeq& sq, int n_slots, int t_cap, int eos_id, float* d_logits) {
    if (n_slots <= 0 || n_slots > XR_MAX_SLOTS) return fail(GTB_ERR_ARG, "multi-row decode pass: 1..%d sequences", XR_MAX_SLOTS);
    XrPlanArgs plan{};
    plan.mode = 1; plan.n_rows = n_slots; plan.slot0 = 0; plan.tok_stride = sq.tok_stride;
    plan.tokens = sq.tokens; plan.st = sq.st; plan.rows = p->rows;
    if (m.cfg.wdtype == GTB_F16) return run_pass_f16(p, m, kv, sq, plan, t_cap, 0, n_slots, eos_id, d_logits);
    return (m.cfg.wdtype == GTB_Q8) ? run_pass<DT_Q8>(p, m, kv, sq, plan, t_cap, 0, n_slots, eos_id, d_logits)
                                    : run_pass<DT_Q4>(p, m, kv, sq, plan, t_cap, 0, n_slots, eos_id, d_logits);
}

}  // namespace gtb
