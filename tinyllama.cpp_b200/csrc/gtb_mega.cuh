// gtb_mega.cuh -- the per-token loop of TinyLlama::logits (tinyllama.cpp:45-61, 395-440) as ONE persistent
// cooperative kernel: one CTA per SM, every dependent step of a row ("phase") separated by a data exchange
// through L2 instead of a kernel boundary, rows and greedy steps looped inside the kernel.
//
// Why: batch-1 decode is ~155 dependent GEMVs per token; at 2-5 us per kernel boundary a graph of small
// kernels cannot approach the HBM roofline (SURVEY.md §7 hard part 2).  Here a phase boundary costs one L2
// round trip:
//   * exchange = "LL" words: every produced fp32 value travels as one 64-bit store {tag:32 | bits:32};
//     consumers poll the words themselves until the tag equals the phase's epoch, so data and flag arrive
//     together (no separate barrier, no fence on the critical path).  Tags are a monotone epoch counter.
//   * weights never depend on activations: each thread issues the 128-bit loads of its share of the NEXT
//     phase's weight blocks into registers before it starts polling, and one thread per CTA pushes the same
//     rows of a later phase into L2 with cp.async.bulk.prefetch -- HBM streaming is decoupled from the
//     dependency chain.
//   * the numeric contract is unchanged (gtb_dev.cuh / gtb_kernels.cuh): exact integer block dots, products
//     parked in shared memory, then one thread per (row, lane) runs the reference's ordered fp32 adds.
//
// Phases of one layer (all CTAs walk the same sequence; the tag of each exchange is the next epoch):
//   P1   [x' = E(h + E(down)); E(rmsnorm(x'))] -> q|k|v rows                      -> x_qkv   (raw fp32)
//   P2a  unit (head h, quarter j): E/RoPE/E of q_h,k_g,v_g; append K/V; scores of its position quarter -> x_sc
//   P2b  same unit: softmax over the whole row (exact in-order sum), P.V for its 16 channels          -> x_attn
//   P3   [E(attn)] -> o rows                                                       -> x_o
//   P4   [h = E(x' + E(o)); E(rmsnorm(h))] -> gate|up rows                         -> x_gu
//   P4b  block b of the FFN vector: E(E(silu(E(gate))) * E(up))                    -> x_act  (packed codes)
//   P5   -> down rows                                                              -> x_down
//   head [final residual + norm] -> logits rows, per-CTA first-maximum              -> x_arg -> next token
#pragma once
#include "gtb_kernels.cuh"

namespace gtb {

typedef unsigned long long ull;

constexpr int MT = 512;                    // threads per CTA
constexpr int MWARP = MT / 32;
constexpr int PS_BYTES = 84 * 1024;        // product staging (also the attention scratch)
constexpr int IT_Q4 = 10;                  // (row, block) items per thread per tile: MT*IT items in registers
constexpr int IT_Q8 = 5;
constexpr int SPIN_LIMIT = 1 << 24;        // ~ seconds; a stuck exchange traps instead of hanging the GPU

struct MegaLayer {
    const void* w[7];                      // q k v o gate up down (device layout, gtb_internal.h)
    const uint16_t* s[7];
    const uint16_t* attn_norm;
    const uint16_t* ffn_norm;
    uint8_t* kq; uint16_t* ks; uint8_t* vq; uint16_t* vs;
};

struct MegaParams {
    int E, F, KV, n_heads, gsz, n_layers, n_vocab, max_ctx, sc_stride;
    const MegaLayer* layers;
    const void* emb_w; const uint16_t* emb_s;
    const void* head_w; const uint16_t* head_s;
    const uint16_t* final_norm;
    const float* rope_cos; const float* rope_sin;
    ull *x_qkv, *x_sc, *x_attn, *x_o, *x_gu, *x_act, *x_down, *x_arg;
    float* logits;
    int32_t* tokens;
    DevState* st;
    unsigned int* epoch;
    int n_body, n_head, eos_id;
    int pf_ahead;                          // L2 prefetch distance in GEMV phases
    ull* dbg;                              // [8] watchdog diagnostics
    long long* prof;                       // optional: CTA 0 timestamps of the last row
};

// ---------------------------------------------------------------- LL words
__device__ __forceinline__ void ll_store(ull* p, uint32_t payload, uint32_t tag) {
    const ull v = ((ull)tag << 32) | (ull)payload;
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ll_load2(const ull* p, ull& a, ull& b) {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ ull ll_load1(const ull* p) {
    ull a;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
__device__ __noinline__ void ll_timeout(ull* dbg, uint32_t tag, const ull* p, ull seen) {
    if (dbg) {
        dbg[1] = tag; dbg[2] = (ull)p; dbg[3] = seen; dbg[4] = blockIdx.x; dbg[5] = threadIdx.x;
        __threadfence_system();
        dbg[0] = 0xdeadull;
        __threadfence_system();
    }
    __trap();
}
// wait for words [i, i+1] of src
__device__ __forceinline__ void ll_wait2(const ull* p, uint32_t tag, uint32_t& a, uint32_t& b, ull* dbg) {
    ull x, y;
    int spins = 0;
    while (true) {
        ll_load2(p, x, y);
        if ((uint32_t)(x >> 32) == tag && (uint32_t)(y >> 32) == tag) break;
        if (++spins > SPIN_LIMIT) ll_timeout(dbg, tag, p, x);
    }
    a = (uint32_t)x; b = (uint32_t)y;
}
__device__ __forceinline__ uint32_t ll_wait1(const ull* p, uint32_t tag, ull* dbg) {
    ull x;
    int spins = 0;
    while (true) {
        x = ll_load1(p);
        if ((uint32_t)(x >> 32) == tag) break;
        if (++spins > SPIN_LIMIT) ll_timeout(dbg, tag, p, x);
    }
    return (uint32_t)x;
}
// CTA-wide gather of n words (n even or odd, src 16-byte aligned); sink(index, payload)
template <typename Sink>
__device__ __forceinline__ void ll_gather(const ull* src, int n, uint32_t tag, ull* dbg, Sink sink) {
    const int n2 = n & ~1;
    for (int i = threadIdx.x * 2; i < n2; i += MT * 4) {
        // two pairs in flight per thread
        const int i2 = i + MT * 2;
        ull x0, y0, x1 = 0, y1 = 0;
        const bool second = i2 < n2;
        ll_load2(src + i, x0, y0);
        if (second) ll_load2(src + i2, x1, y1);
        int spins = 0;
        while ((uint32_t)(x0 >> 32) != tag || (uint32_t)(y0 >> 32) != tag) {
            ll_load2(src + i, x0, y0);
            if (++spins > SPIN_LIMIT) ll_timeout(dbg, tag, src + i, x0);
        }
        sink(i, (uint32_t)x0); sink(i + 1, (uint32_t)y0);
        if (second) {
            while ((uint32_t)(x1 >> 32) != tag || (uint32_t)(y1 >> 32) != tag) {
                ll_load2(src + i2, x1, y1);
                if (++spins > SPIN_LIMIT) ll_timeout(dbg, tag, src + i2, x1);
            }
            sink(i2, (uint32_t)x1); sink(i2 + 1, (uint32_t)y1);
        }
    }
    if ((n & 1) && threadIdx.x == MT - 1) sink(n - 1, ll_wait1(src + n - 1, tag, dbg));
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u16(const uint16_t* p) {
    unsigned short r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return (uint32_t)r;
}
__device__ __forceinline__ void l2_prefetch(const void* p, size_t bytes) {
    if (bytes == 0) return;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)15;
    const size_t n = ((reinterpret_cast<uintptr_t>(p) + bytes + 15) & ~(uintptr_t)15) - a;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((uint32_t)n) : "memory");
}
__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------- phase descriptors
struct PhaseDesc {
    const void* d[3];
    const uint16_t* s[3];
    int rows[3];
    int nm;        // matrices whose rows are concatenated into one row space
    int nb;        // K / 32
    int R;         // total rows
};

template <int WT> struct ItemsPerThread { static constexpr int v = (WT == DT_Q4) ? IT_Q4 : IT_Q8; };

struct MegaSm {
    ExactSumSmem es;
    float raw[192];
    uint32_t qw[16]; float qd[2]; float qf[64];
    uint32_t kw[16]; float kd[2]; float kf[64];
    float vf[64];
    float tmp[6][32];
    float red[MWARP];
    float part[8][16];
    float bestv[MWARP]; int besti[MWARP];
};

struct MCtx {
    float* res;            // residual stream of the current row (every CTA keeps its own copy)
    float* xbuf;           // scratch vector
    ActView av;            // staged GEMV input
    float* ps;             // product staging / attention scratch
    MegaSm* sm;
    unsigned int ep;       // epoch counter: identical sequence in every CTA
    int cta, G;
    int prof_i;
};

__host__ __device__ inline size_t mega_smem_bytes(int at, int E, int F) {
    size_t s = 0;
    s += (size_t)E * 4;                              // res
    s += (size_t)((F > E ? F : E) + 64) * 4;          // xbuf
    s += (act_bytes(at, F) + 15) & ~(size_t)15;      // act
    s += PS_BYTES;
    s += (sizeof(MegaSm) + 15) & ~(size_t)15;
    return s + 32;
}

template <int WT>
struct WRegs {
    static constexpr int IT = ItemsPerThread<WT>::v;
    uint4 a[IT];
    uint4 b[(WT == DT_Q8) ? IT : 1];
    uint32_t sc[IT];
};

template <int WT>
__device__ __forceinline__ int tile_rows(int nb) {
    int tr = MT / 4;                                       // 4 chain threads per row
    const int by_ps = PS_BYTES / (16 * (nb + 1));
    if (by_ps < tr) tr = by_ps;
    const int by_regs = (MT * ItemsPerThread<WT>::v) / nb;
    if (by_regs < tr) tr = by_regs;
    return tr;
}

// rows of the phase owned by this CTA: [r0, r1)
__device__ __forceinline__ void cta_rows(const PhaseDesc& pd, int cta, int G, int& r0, int& r1) {
    r0 = (int)(((long long)cta * pd.R) / G);
    r1 = (int)(((long long)(cta + 1) * pd.R) / G);
}

// issue the weight loads of one tile: rows [row_begin, row_begin + nrows) of the concatenated row space
template <int WT>
__device__ __forceinline__ void load_tile(const PhaseDesc& pd, int row_begin, int nrows, WRegs<WT>& w) {
    constexpr int IT = WRegs<WT>::IT;
    const int nb = pd.nb;
    int rl = threadIdx.x / nb, b = threadIdx.x - rl * nb;
    const int drl = MT / nb, db = MT - drl * nb;
    const int c1 = pd.rows[0], c2 = pd.rows[0] + pd.rows[1];
#pragma unroll
    for (int j = 0; j < IT; j++) {
        if (rl < nrows) {
            const int gr = row_begin + rl;
            const int m = (pd.nm > 1 && gr >= c1) ? ((pd.nm > 2 && gr >= c2) ? 2 : 1) : 0;
            const int lr = gr - (m == 0 ? 0 : (m == 1 ? c1 : c2));
            const size_t blk = (size_t)lr * nb + b;
            const uint4* dp = reinterpret_cast<const uint4*>(pd.d[m]);
            if (WT == DT_Q4) {
                w.a[j] = ldg_stream(dp + blk);
            } else {
                w.a[j] = ldg_stream(dp + 2 * blk);
                w.b[j] = ldg_stream(dp + 2 * blk + 1);
            }
            w.sc[j] = ldg_stream_u16(pd.s[m] + blk);
        }
        b += db; rl += drl;
        if (b >= nb) { b -= nb; rl++; }
    }
}

// exact integer lane sums of one block from registers, scaled: p[l] = float(lane[l]) * (da * dw)  (ops.h:282-287)
template <int WT>
__device__ __forceinline__ float4 block_products_r(const uint4& wa, const uint4& wb, uint32_t sc16, const ActView& av, int b) {
    const float dw = h2f((uint16_t)sc16);
    const uint4 ax = reinterpret_cast<const uint4*>(av.aw)[2 * b];
    const uint4 ay = reinterpret_cast<const uint4*>(av.aw)[2 * b + 1];
    const float s = __fmul_rn(av.ad[b], dw);
    int l0, l1, l2, l3;
    if (WT == DT_Q4) {
        const int4 n7 = reinterpret_cast<const int4*>(av.ns7)[b];
        l0 = __dp4a((int)(wa.x & 0x0f0f0f0fu), (int)ay.x, __dp4a((int)((wa.x >> 4) & 0x0f0f0f0fu), (int)ax.x, n7.x));
        l1 = __dp4a((int)(wa.y & 0x0f0f0f0fu), (int)ay.y, __dp4a((int)((wa.y >> 4) & 0x0f0f0f0fu), (int)ax.y, n7.y));
        l2 = __dp4a((int)(wa.z & 0x0f0f0f0fu), (int)ay.z, __dp4a((int)((wa.z >> 4) & 0x0f0f0f0fu), (int)ax.z, n7.z));
        l3 = __dp4a((int)(wa.w & 0x0f0f0f0fu), (int)ay.w, __dp4a((int)((wa.w >> 4) & 0x0f0f0f0fu), (int)ax.w, n7.w));
    } else {
        l0 = __dp4a((int)wb.x, (int)ay.x, __dp4a((int)wa.x, (int)ax.x, 0));
        l1 = __dp4a((int)wb.y, (int)ay.y, __dp4a((int)wa.y, (int)ax.y, 0));
        l2 = __dp4a((int)wb.z, (int)ay.z, __dp4a((int)wa.z, (int)ax.z, 0));
        l3 = __dp4a((int)wb.w, (int)ay.w, __dp4a((int)wa.w, (int)ax.w, 0));
    }
    return make_float4(__fmul_rn((float)l0, s), __fmul_rn((float)l1, s), __fmul_rn((float)l2, s), __fmul_rn((float)l3, s));
}

template <int WT>
__device__ __forceinline__ void tile_products(const PhaseDesc& pd, int nrows, const WRegs<WT>& w, const ActView& av, float* ps) {
    constexpr int IT = WRegs<WT>::IT;
    const int nb = pd.nb;
    int rl = threadIdx.x / nb, b = threadIdx.x - rl * nb;
    const int drl = MT / nb, db = MT - drl * nb;
    float4* ps4 = reinterpret_cast<float4*>(ps);
#pragma unroll
    for (int j = 0; j < IT; j++) {
        if (rl < nrows) ps4[rl * (nb + 1) + b] = block_products_r<WT>(w.a[j], w.b[(WT == DT_Q8) ? j : 0], w.sc[j], av, b);
        b += db; rl += drl;
        if (b >= nb) { b -= nb; rl++; }
    }
}

// ordered adds: thread (row, l) sums its lane's products in ascending block order, then (a0+a1)+(a2+a3).
// Returns the row result on the l == 0 thread of each row (valid when tid < nrows * 4).
__device__ __forceinline__ float tile_chain(int nb, int nrows, const float* ps) {
    const int ct = threadIdx.x;
    float acc = 0.0f;
    if (ct < nrows * 4) {
        const float* src = ps + (size_t)(ct >> 2) * (nb + 1) * 4 + (ct & 3);
        int b = 0;
        for (; b + 8 <= nb; b += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = src[(b + u) * 4];
#pragma unroll
            for (int u = 0; u < 8; u++) acc = __fadd_rn(acc, v[u]);
        }
        for (; b < nb; b++) acc = __fadd_rn(acc, src[b * 4]);
    }
    const float v1 = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    return __fadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, 2));
}

// L2 prefetch of this CTA's rows of a phase (one thread)
template <int WT>
__device__ __forceinline__ void prefetch_phase(const PhaseDesc& pd, int cta, int G) {
    int r0, r1;
    cta_rows(pd, cta, G, r0, r1);
    int base = 0;
    for (int m = 0; m < pd.nm; m++) {
        const int lo = max(r0, base) - base, hi = min(r1, base + pd.rows[m]) - base;
        if (hi > lo) {
            if (WT == DT_F16) {
                l2_prefetch(reinterpret_cast<const uint8_t*>(pd.d[m]) + (size_t)lo * pd.nb * 64, (size_t)(hi - lo) * pd.nb * 64);
            } else {
                const size_t bb = (WT == DT_Q4) ? 16 : 32;
                l2_prefetch(reinterpret_cast<const uint8_t*>(pd.d[m]) + (size_t)lo * pd.nb * bb, (size_t)(hi - lo) * pd.nb * bb);
                l2_prefetch(pd.s[m] + (size_t)lo * pd.nb, (size_t)(hi - lo) * pd.nb * 2);
            }
        }
        base += pd.rows[m];
    }
}

// ---------------------------------------------------------------- the GEMV sequence of one row
// index s: layer * 4 + {0: q|k|v, 1: o, 2: gate|up, 3: down}; s == 4 * n_layers: lm_head
__device__ __forceinline__ PhaseDesc phase_desc(const MegaParams& P, int s) {
    PhaseDesc pd;
    pd.d[1] = pd.d[2] = nullptr; pd.s[1] = pd.s[2] = nullptr; pd.rows[1] = pd.rows[2] = 0;
    if (s >= 4 * P.n_layers) {
        pd.d[0] = P.head_w; pd.s[0] = P.head_s; pd.rows[0] = P.n_vocab; pd.nm = 1; pd.nb = P.E / 32; pd.R = P.n_vocab;
        return pd;
    }
    const MegaLayer& L = P.layers[s >> 2];
    switch (s & 3) {
        case 0:
            pd.d[0] = L.w[0]; pd.d[1] = L.w[1]; pd.d[2] = L.w[2]; pd.s[0] = L.s[0]; pd.s[1] = L.s[1]; pd.s[2] = L.s[2];
            pd.rows[0] = P.E; pd.rows[1] = P.KV; pd.rows[2] = P.KV; pd.nm = 3; pd.nb = P.E / 32; pd.R = P.E + 2 * P.KV;
            break;
        case 1:
            pd.d[0] = L.w[3]; pd.s[0] = L.s[3]; pd.rows[0] = P.E; pd.nm = 1; pd.nb = P.E / 32; pd.R = P.E;
            break;
        case 2:
            pd.d[0] = L.w[4]; pd.d[1] = L.w[5]; pd.s[0] = L.s[4]; pd.s[1] = L.s[5];
            pd.rows[0] = P.F; pd.rows[1] = P.F; pd.nm = 2; pd.nb = P.E / 32; pd.R = 2 * P.F;
            break;
        default:
            pd.d[0] = L.w[6]; pd.s[0] = L.s[6]; pd.rows[0] = P.E; pd.nm = 1; pd.nb = P.F / 32; pd.R = P.E;
            break;
    }
    return pd;
}

// F16 weights: direct streaming, thread (row, lane l of 8), no staging (gten/ops.h:140-160)
__device__ __forceinline__ float f16_row_lane(const uint4* __restrict__ src, int cpr, int l, const ActView& av) {
    float acc = 0.0f;
    constexpr int UN = 8;
    int c = 0;
    for (; c + UN <= cpr; c += UN) {
        uint4 w[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) w[u] = ldg_stream(src + (size_t)(c + u) * 8);
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
        const uint4 w = ldg_stream(src + (size_t)c * 8);
        const float* x = av.xs + (c * 8 + l) * 8;
        const __half* hw = reinterpret_cast<const __half*>(&w);
#pragma unroll
        for (int i = 0; i < 8; i++) acc = fmaf(x[i], __half2float(hw[i]), acc);
    }
    return acc;
}

// One GEMV phase.  sink(row, value) is called by one thread per finished row (row = index in the phase's row space).
// For Q4/Q8 `w` holds the first tile's weights on entry; on exit it holds the first tile of `next` (if next_valid).
template <int WT, typename Sink>
__device__ __forceinline__ void gemv_phase(const MegaParams& P, MCtx& c, const PhaseDesc& pd, WRegs<WT>& w,
                                           const PhaseDesc& next, bool next_valid, Sink sink) {
    int r0, r1;
    cta_rows(pd, c.cta, c.G, r0, r1);
    if (WT == DT_F16) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, rl = lane >> 3, l = lane & 7;
        const int cpr = pd.nb / 2;                                   // 64-element chunks per row
        const int c1 = pd.rows[0], c2 = pd.rows[0] + pd.rows[1];
        for (int row0 = r0 + wid * 4; row0 < r1; row0 += MWARP * 4) {
            const int gr = row0 + rl;
            const bool active = gr < r1;
            const int grc = active ? gr : r0;
            const int m = (pd.nm > 1 && grc >= c1) ? ((pd.nm > 2 && grc >= c2) ? 2 : 1) : 0;
            const int lr = grc - (m == 0 ? 0 : (m == 1 ? c1 : c2));
            const float acc = f16_row_lane(reinterpret_cast<const uint4*>(pd.d[m]) + (size_t)lr * cpr * 8 + l, cpr, l, c.av);
            float v = __shfl_sync(0xffffffffu, acc, lane & ~7);
#pragma unroll
            for (int j = 1; j < 8; j++) v = __fadd_rn(v, __shfl_sync(0xffffffffu, acc, (lane & ~7) + j));
            if (l == 0 && active) sink(gr, v);
        }
        return;
    }
    const int nb = pd.nb;
    const int tr = tile_rows<WT>(nb);
    for (int t0 = r0; t0 < r1; t0 += tr) {
        const int nrows = min(tr, r1 - t0);
        tile_products<WT>(pd, nrows, w, c.av, c.ps);
        if (t0 + tr < r1) {
            load_tile<WT>(pd, t0 + tr, min(tr, r1 - t0 - tr), w);
        } else if (next_valid) {
            int n0, n1;
            cta_rows(next, c.cta, c.G, n0, n1);
            const int ntr = tile_rows<WT>(next.nb);
            load_tile<WT>(next, n0, min(ntr, n1 - n0), w);
        }
        __syncthreads();
        const float v = tile_chain(nb, nrows, c.ps);
        if ((threadIdx.x & 3) == 0 && threadIdx.x < nrows * 4) sink(t0 + (threadIdx.x >> 2), v);
        if (t0 + tr < r1) __syncthreads();
    }
    if (r1 <= r0 && next_valid) {                     // no rows of this phase here: still fetch the next phase's first tile
        int n0, n1;
        cta_rows(next, c.cta, c.G, n0, n1);
        load_tile<WT>(next, n0, min(tile_rows<WT>(next.nb), n1 - n0), w);
    }
}

// ---------------------------------------------------------------- attention, unit = (head h, quarter j)
template <int AT>
__device__ __forceinline__ float mega_score(const MegaSm& sm, const MegaLayer& L, int KV, int g, int kcol, int own) {
    if (AT == DT_F16) {
        float acc[8];
#pragma unroll
        for (int l = 0; l < 8; l++) acc[l] = 0.0f;
        if (kcol == own) {
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int l = 0; l < 8; l++) acc[l] = fmaf(sm.qf[8 * i + l], sm.kf[8 * i + l], acc[l]);
        } else {
            const uint4* kp = reinterpret_cast<const uint4*>(L.kq + ((size_t)kcol * KV + g * 64) * 2);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint4 w = __ldcg(kp + i);
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
            if (kcol == own) {
                kx = make_uint4(sm.kw[bi * 8 + 0], sm.kw[bi * 8 + 1], sm.kw[bi * 8 + 2], sm.kw[bi * 8 + 3]);
                ky = make_uint4(sm.kw[bi * 8 + 4], sm.kw[bi * 8 + 5], sm.kw[bi * 8 + 6], sm.kw[bi * 8 + 7]);
                kdv = sm.kd[bi];
            } else {
                const uint4* kp = reinterpret_cast<const uint4*>(L.kq + (size_t)kcol * KV + g * 64 + bi * 32);
                kx = __ldcg(kp); ky = __ldcg(kp + 1);
                kdv = h2f(__ldcg(L.ks + (size_t)kcol * (KV / 32) + g * 2 + bi));
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

// attention scratch carved from the product staging area
struct AttnScratch {
    float* sc;          // [max_ctx + 32] scores / probabilities
    uint8_t* vb;        // Q8: int8 [max_ctx][16]; F16: half [max_ctx][16]
    float* vd;          // Q8: block scale of the slice per position
};
__device__ __forceinline__ AttnScratch attn_scratch(int at, float* ps, int max_ctx) {
    AttnScratch a;
    a.sc = ps;
    const size_t o1 = (size_t)((max_ctx + 32 + 3) & ~3) * 4;
    a.vb = reinterpret_cast<uint8_t*>(ps) + o1;
    const size_t o2 = o1 + (size_t)max_ctx * (at == DT_F16 ? 32 : 16);
    a.vd = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ps) + o2);
    return a;
}
__host__ __device__ inline size_t attn_scratch_bytes(int at, int max_ctx) {
    return (size_t)((max_ctx + 32 + 3) & ~3) * 4 + (size_t)max_ctx * (at == DT_F16 ? 32 : 16) + (size_t)max_ctx * 4 + 64;
}

// P2a: re-encode / RoPE the unit's q, k, v (gten/ops.h:645-646, 733-753), append K/V, publish the scores of quarter j
template <int AT>
__device__ __forceinline__ void mega_attn_a(const MegaParams& P, const MegaLayer& L, MCtx& c, int pos, uint32_t tag_qkv, uint32_t tag_sc) {
    MegaSm& sm = *c.sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int h = c.cta >> 2, j = c.cta & 3, g = h / P.gsz;
    const bool writer = (h % P.gsz) == 0 && j == 0;
    const AttnScratch as = attn_scratch(AT, c.ps, P.max_ctx);
    // this unit's V slice (16 channels) of every cached position: independent of the exchange, issue first
    {
        const int ch0 = g * 64 + j * 16;
        for (int i = tid; i < pos; i += MT) {
            if (AT == DT_F16) {
                const uint4* vp = reinterpret_cast<const uint4*>(L.vq + ((size_t)i * P.KV + ch0) * 2);
                reinterpret_cast<uint4*>(as.vb)[i * 2] = __ldcg(vp);
                reinterpret_cast<uint4*>(as.vb)[i * 2 + 1] = __ldcg(vp + 1);
            } else {
                reinterpret_cast<uint4*>(as.vb)[i] = __ldcg(reinterpret_cast<const uint4*>(L.vq + (size_t)i * P.KV + ch0));
                as.vd[i] = h2f(__ldcg(L.vs + (size_t)i * (P.KV / 32) + g * 2 + (j >> 1)));
            }
        }
    }
    if (tid < 96) {
        const int seg = tid >> 5, o = (tid & 31) * 2;
        const ull* src = P.x_qkv + (seg == 0 ? h * 64 : (seg == 1 ? P.E + g * 64 : P.E + P.KV + g * 64)) + o;
        uint32_t a, b;
        ll_wait2(src, tag_qkv, a, b, P.dbg);
        sm.raw[seg * 64 + o] = __uint_as_float(a);
        sm.raw[seg * 64 + o + 1] = __uint_as_float(b);
    }
    __syncthreads();
    if (wid < 6) {
        const int which = wid >> 1, half = wid & 1;
        const float x = sm.raw[which * 64 + half * 32 + lane];
        const int ch = half * 32 + lane;                       // channel inside the head
        const int sl = ch - j * 16;                            // channel inside this unit's P.V slice
        if (which < 2) {
            sm.tmp[wid][lane] = roundtrip<AT>(x);
        } else if (AT == DT_F16) {
            const uint16_t hb = f2h(x);
            sm.vf[ch] = h2f(hb);
            if (sl >= 0 && sl < 16) reinterpret_cast<uint16_t*>(as.vb)[pos * 16 + sl] = hb;
            if (writer) reinterpret_cast<uint16_t*>(L.vq)[(size_t)pos * P.KV + g * 64 + ch] = hb;
        } else {
            uint16_t dh;
            const int q = q8_encode_lane(x, &dh);
            sm.vf[ch] = __fmul_rn((float)q, h2f(dh));
            if (sl >= 0 && sl < 16) as.vb[pos * 16 + sl] = (uint8_t)(int8_t)q;
            if (lane == 0 && half == (j >> 1)) as.vd[pos] = h2f(dh);
            if (writer) {
                L.vq[(size_t)pos * P.KV + g * 64 + ch] = (uint8_t)(int8_t)q;
                if (lane == 0) L.vs[(size_t)pos * (P.KV / 32) + g * 2 + half] = dh;
            }
        }
    }
    __syncthreads();
    if (wid < 4) {
        const int which = wid >> 1, half = wid & 1;
        const float x0 = sm.tmp[which * 2][lane], x1 = sm.tmp[which * 2 + 1][lane];
        const float cs = __ldg(P.rope_cos + (size_t)pos * 32 + lane), sn = __ldg(P.rope_sin + (size_t)pos * 32 + lane);
        const float o = (half == 0) ? __fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn))
                                    : __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs));
        if (AT == DT_F16) {
            const uint16_t hb = f2h(o);
            const float d = h2f(hb);
            if (which == 0) sm.qf[half * 32 + lane] = d;
            else {
                sm.kf[half * 32 + lane] = d;
                if (writer) reinterpret_cast<uint16_t*>(L.kq)[(size_t)pos * P.KV + g * 64 + half * 32 + lane] = hb;
            }
        } else {
            uint16_t dh;
            const int q = q8_encode_lane(o, &dh);
            const float delta = h2f(dh);
            const int pb = perm_byte(lane);
            if (which == 0) {
                reinterpret_cast<int8_t*>(sm.qw)[half * 32 + pb] = (int8_t)q;
                if (lane == 0) sm.qd[half] = delta;
            } else {
                reinterpret_cast<int8_t*>(sm.kw)[half * 32 + pb] = (int8_t)q;
                if (lane == 0) sm.kd[half] = delta;
                if (writer) {
                    L.kq[(size_t)pos * P.KV + g * 64 + half * 32 + pb] = (uint8_t)(int8_t)q;
                    if (lane == 0) L.ks[(size_t)pos * (P.KV / 32) + g * 2 + half] = dh;
                }
            }
        }
    }
    __syncthreads();
    // scores of this unit's quarter of the positions, scaled by 1/sqrt(64) (exactly 0.125)
    const int per = (pos + 4) >> 2;
    const int lo = j * per, hi = min(pos + 1, lo + per);
    ull* dst = P.x_sc + (size_t)h * P.sc_stride;
    for (int k = lo + tid; k < hi; k += MT) {
        const float s = __fmul_rn(mega_score<AT>(sm, L, P.KV, g, k, pos), 0.125f);
        ll_store(dst + k, __float_as_uint(s), tag_sc);
    }
}

// P2b: softmax over the whole row (gten/ops.h:967-996), then P.V for the unit's 16 channels (ops.h:1046-1087)
template <int AT>
__device__ __forceinline__ void mega_attn_b(const MegaParams& P, MCtx& c, int pos, int n_ctx, uint32_t tag_sc, uint32_t tag_attn) {
    MegaSm& sm = *c.sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int h = c.cta >> 2, j = c.cta & 3;
    const AttnScratch as = attn_scratch(AT, c.ps, P.max_ctx);
    float* sc = as.sc;
    ll_gather(P.x_sc + (size_t)h * P.sc_stride, pos + 1, tag_sc, P.dbg, [&](int i, uint32_t v) { sc[i] = __uint_as_float(v); });
    __syncthreads();
    float mx = -INFINITY;
    for (int k = tid; k <= pos; k += MT) mx = fmaxf(mx, sc[k]);
    mx = warp_max(mx);
    if (lane == 0) sm.red[wid] = mx;
    __syncthreads();
    mx = sm.red[0];
#pragma unroll
    for (int w = 1; w < MWARP; w++) mx = fmaxf(mx, sm.red[w]);
    for (int k = tid; k <= pos; k += MT) sc[k] = expf_glibc(__fsub_rn(sc[k], mx));
    __syncthreads();
    const float sum = exact_sum_block([&](int i) { return sc[i]; }, pos + 1, sm.es);
    const int nblk = (pos + 32) / 32;
    for (int b = wid; b < nblk; b += MWARP) {
        const int i = b * 32 + lane;
        const float p = (i <= pos) ? __fdiv_rn(sc[i], sum) : 0.0f;
        const float ph = roundtrip<AT>(p);
        __syncwarp();
        if (i <= pos) sc[i] = ph;
    }
    __syncthreads();
    const int n8 = (n_ctx / 8) * 8;
    auto v_of = [&](int i, int cc) -> float {
        if (AT == DT_F16) return h2f(reinterpret_cast<const uint16_t*>(as.vb)[i * 16 + cc]);
        return __fmul_rn((float)(int8_t)as.vb[i * 16 + cc], as.vd[i]);                       // ops.h:1026
    };
    if (tid < 128) {
        const int l = tid >> 4, cc = tid & 15;
        const int hi = min(n8, pos + 1);
        float a = 0.0f;
        int i = l;
        for (; i + 24 < hi; i += 32) {
            float p[4], v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { p[u] = sc[i + 8 * u]; v[u] = v_of(i + 8 * u, cc); }
#pragma unroll
            for (int u = 0; u < 4; u++) a = __fadd_rn(__fmul_rn(p[u], v[u]), a);
        }
        for (; i < hi; i += 8) a = __fadd_rn(__fmul_rn(sc[i], v_of(i, cc)), a);
        sm.part[l][cc] = a;
    }
    __syncthreads();
    if (tid < 16) {
        float d = __fadd_rn(sm.part[0][tid], sm.part[1][tid]);
#pragma unroll
        for (int l = 2; l < 8; l++) d = __fadd_rn(d, sm.part[l][tid]);
        for (int i = n8; i < n_ctx && i <= pos; i++) d = __fadd_rn(d, __fmul_rn(sc[i], v_of(i, tid)));
        ll_store(P.x_attn + h * 64 + j * 16 + tid, __float_as_uint(d), tag_attn);
    }
}

// P4b: one warp per block of 32 FFN channels: E(E(silu(E(gate))) * E(up)) (gten/modules.cpp:238-247), published as
// packed words in the staged layout (Q8: 8 code words + the fp16 scale; F16: 16 half2 words)
template <int AT>
__device__ __forceinline__ void mega_silu(const MegaParams& P, MCtx& c, uint32_t tag_gu, uint32_t tag_act) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nblk = P.F / 32;
    for (int b = c.cta + wid * c.G; b < nblk; b += c.G * MWARP) {
        const int e = b * 32 + lane;
        const float g0 = __uint_as_float(ll_wait1(P.x_gu + e, tag_gu, P.dbg));
        const float u0 = __uint_as_float(ll_wait1(P.x_gu + P.F + e, tag_gu, P.dbg));
        const float g1 = roundtrip<AT>(g0);
        const float u1 = roundtrip<AT>(u0);
        const float g2 = roundtrip<AT>(silu_ref(g1));
        const float y = __fmul_rn(g2, u1);
        if (AT == DT_F16) {
            const uint32_t hb = f2h(y);
            const uint32_t lo = __shfl_sync(0xffffffffu, hb, (2 * lane) & 31);
            const uint32_t hi = __shfl_sync(0xffffffffu, hb, (2 * lane + 1) & 31);
            if (lane < 16) ll_store(P.x_act + (size_t)b * 16 + lane, lo | (hi << 16), tag_act);
        } else {
            uint16_t dh;
            const uint32_t q = (uint32_t)q8_encode_lane(y, &dh) & 0xffu;
            // word w (0..7): half = w >> 2, l = w & 3: codes of elements 16*half + {2l, 2l+1, 2l+8, 2l+9}
            const int w = lane & 7, e0 = 16 * (w >> 2) + 2 * (w & 3);
            const uint32_t b0 = __shfl_sync(0xffffffffu, q, e0);
            const uint32_t b1 = __shfl_sync(0xffffffffu, q, e0 + 1);
            const uint32_t b2 = __shfl_sync(0xffffffffu, q, e0 + 8);
            const uint32_t b3 = __shfl_sync(0xffffffffu, q, e0 + 9);
            if (lane < 8) ll_store(P.x_act + (size_t)b * 9 + lane, b0 | (b1 << 8) | (b2 << 16) | (b3 << 24), tag_act);
            if (lane == 8) ll_store(P.x_act + (size_t)b * 9 + 8, (uint32_t)dh, tag_act);
        }
    }
}

// P5 prologue: the staged GEMV input straight from the packed words
template <int AT, int WT>
__device__ __forceinline__ void mega_gather_act(const MegaParams& P, MCtx& c, uint32_t tag_act) {
    const int nblk = P.F / 32;
    if (AT == DT_F16) {
        ll_gather(P.x_act, nblk * 16, tag_act, P.dbg, [&](int i, uint32_t v) {
            const int e = 2 * i;
            const float lo = h2f((uint16_t)(v & 0xffffu)), hi = h2f((uint16_t)(v >> 16));
            c.av.xs[(((e >> 6) * 8) + (e & 7)) * 8 + ((e >> 3) & 7)] = lo;
            c.av.xs[((((e + 1) >> 6) * 8) + ((e + 1) & 7)) * 8 + (((e + 1) >> 3) & 7)] = hi;
        });
    } else {
        ll_gather(P.x_act, nblk * 9, tag_act, P.dbg, [&](int i, uint32_t v) {
            const int b = i / 9, k = i - b * 9;
            if (k < 8) c.av.aw[b * 8 + k] = v;
            else c.av.ad[b] = h2f((uint16_t)v);
        });
        if (WT == DT_Q4) {
            __syncthreads();
            for (int i = threadIdx.x; i < nblk * 4; i += MT) {
                const int b = i >> 2, l = i & 3;
                const int s = __dp4a((int)c.av.aw[b * 8 + l], 0x01010101, __dp4a((int)c.av.aw[b * 8 + 4 + l], 0x01010101, 0));
                c.av.ns7[i] = -7 * s;
            }
        }
    }
}

// token embedding into the residual stream (gten/ops.h:514-564)
template <int WT>
__device__ __forceinline__ void mega_embed(const MegaParams& P, MCtx& c, int tok) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int E = P.E;
    const size_t row = (size_t)tok;
    for (int b = wid; b < E / 32; b += MWARP) {
        const int e = b * 32 + lane;
        float v;
        if (WT == DT_F16) {
            const int ch = e >> 6, r = e & 63, l = r & 7, ii = r >> 3;
            v = h2f(reinterpret_cast<const uint16_t*>(P.emb_w)[((row * (E / 64) + ch) * 8 + l) * 8 + ii]);
        } else {
            const size_t blk = row * (E / 32) + b;
            const float delta = h2f(P.emb_s[blk]);
            if (WT == DT_Q8) {
                v = __fmul_rn((float)reinterpret_cast<const int8_t*>(P.emb_w)[blk * 32 + perm_byte(lane)], delta);
            } else {
                const int jj = lane & 15;
                const int l = (jj & 7) >> 1, ps = (jj & 1) + 2 * (jj >> 3);
                const uint8_t byte = reinterpret_cast<const uint8_t*>(P.emb_w)[blk * 16 + l * 4 + ps];
                const int q = (int)((lane < 16) ? (byte >> 4) : (byte & 0x0f)) - 7;
                v = q8_roundtrip_lane(__fmul_rn((float)q, delta));
            }
        }
        c.res[e] = v;
    }
}

#define MEGA_PROF()                                                                   \
    do {                                                                              \
        if (P.prof && c.cta == 0 && threadIdx.x == 0) P.prof[c.prof_i++] = gtimer();  \
    } while (0)

template <int WT>
__global__ void __launch_bounds__(MT, 1) k_mega(const MegaParams P) {
    constexpr int AT = (WT == DT_F16) ? DT_F16 : DT_Q8;
    extern __shared__ __align__(16) unsigned char smem[];
    MCtx c;
    {
        size_t off = 0;
        c.res = reinterpret_cast<float*>(smem + off); off += (size_t)P.E * 4;
        c.xbuf = reinterpret_cast<float*>(smem + off); off += (size_t)((P.F > P.E ? P.F : P.E) + 64) * 4;
        c.av = act_carve(AT, P.F, smem + off); off += (act_bytes(AT, P.F) + 15) & ~(size_t)15;
        c.ps = reinterpret_cast<float*>(smem + off); off += PS_BYTES;
        c.sm = reinterpret_cast<MegaSm*>(smem + off);
    }
    c.cta = blockIdx.x; c.G = gridDim.x; c.prof_i = 0;
    c.ep = *reinterpret_cast<volatile unsigned int*>(P.epoch);
    MegaSm& sm = *c.sm;
    const int tid = threadIdx.x;
    const int n_units = P.n_heads * 4;
    const int n_seq = 4 * P.n_layers + 1;          // GEMV phases of a row that ends with the lm_head
    int pos = P.st->pos;
    const int nctx_min = P.st->nctx_min;
    const int n_rows = P.n_body + P.n_head;
    int n_gen = 0, stop = 0, next_tok = -1;
    WRegs<WT> w;
    PhaseDesc pd = phase_desc(P, 0);
    if (WT != DT_F16) {
        int r0, r1;
        cta_rows(pd, c.cta, c.G, r0, r1);
        const int tr = tile_rows<WT>(pd.nb);
        load_tile<WT>(pd, r0, min(tr, r1 - r0), w);
    }
    if (tid == 0) {
        for (int a = 1; a < P.pf_ahead && a < n_seq - 1; a++) prefetch_phase<WT>(phase_desc(P, a), c.cta, c.G);
    }
    for (int r = 0; r < n_rows; r++, pos++) {
        const bool with_head = r >= P.n_body;
        const bool last_row = (r == n_rows - 1);
        const int n_ctx = max(nctx_min, pos + 1);
        const int tok = (next_tok >= 0) ? next_tok : __ldcg(P.tokens + pos);
        // L2 prefetch of this CTA's rows of the GEMV phase pf_ahead steps after phase s (wrapping into the next row)
        auto prefetch_ahead = [&](int s) {
            if (tid != 0 || P.pf_ahead <= 0) return;
            int t = s + P.pf_ahead;
            const int nrow = with_head ? n_seq : n_seq - 1;
            if (t >= nrow) {
                if (last_row) return;
                t -= nrow;
                if (t >= n_seq - 1) return;
            }
            prefetch_phase<WT>(phase_desc(P, t), c.cta, c.G);
        };
        c.prof_i = 0;
        MEGA_PROF();
        mega_embed<WT>(P, c, tok);
        __syncthreads();
        uint32_t tag_down = 0;
        for (int li = 0; li < P.n_layers; li++) {
            const MegaLayer& L = P.layers[li];
            // ---------------- P1
            if (li > 0) {
                ll_gather(P.x_down, P.E, tag_down, P.dbg, [&](int i, uint32_t v) { c.xbuf[i] = __uint_as_float(v); });
                __syncthreads();
            }
            pro_norm<AT>(c.av, c.res, (li > 0) ? c.xbuf : nullptr, L.attn_norm, P.E, c.xbuf, sm.es, c.res, nullptr, nullptr, nullptr);
            __syncthreads();
            MEGA_PROF();
            const uint32_t tag_qkv = ++c.ep;
            prefetch_ahead(li * 4 + 0);
            PhaseDesc nx = phase_desc(P, li * 4 + 1);
            gemv_phase<WT>(P, c, pd, w, nx, true, [&](int row, float v) { ll_store(P.x_qkv + row, __float_as_uint(v), tag_qkv); });
            pd = nx;
            MEGA_PROF();
            // ---------------- P2
            const uint32_t tag_sc = ++c.ep;
            const uint32_t tag_attn = ++c.ep;
            if (c.cta < n_units) {
                __syncthreads();                                   // product staging is reused as attention scratch
                mega_attn_a<AT>(P, L, c, pos, tag_qkv, tag_sc);
                MEGA_PROF();
                mega_attn_b<AT>(P, c, pos, n_ctx, tag_sc, tag_attn);
                __threadfence();                                   // K/V appends visible before anything later is published
            }
            MEGA_PROF();
            // ---------------- P3
            ll_gather(P.x_attn, P.E, tag_attn, P.dbg, [&](int i, uint32_t v) { c.xbuf[i] = __uint_as_float(v); });
            __syncthreads();
            pro_encode<AT>(c.av, c.xbuf, P.E, nullptr);
            __syncthreads();
            MEGA_PROF();
            const uint32_t tag_o = ++c.ep;
            prefetch_ahead(li * 4 + 1);
            nx = phase_desc(P, li * 4 + 2);
            gemv_phase<WT>(P, c, pd, w, nx, true, [&](int row, float v) { ll_store(P.x_o + row, __float_as_uint(v), tag_o); });
            pd = nx;
            MEGA_PROF();
            // ---------------- P4
            ll_gather(P.x_o, P.E, tag_o, P.dbg, [&](int i, uint32_t v) { c.xbuf[i] = __uint_as_float(v); });
            __syncthreads();
            pro_norm<AT>(c.av, c.res, c.xbuf, L.ffn_norm, P.E, c.xbuf, sm.es, c.res, nullptr, nullptr, nullptr);
            __syncthreads();
            MEGA_PROF();
            const uint32_t tag_gu = ++c.ep;
            prefetch_ahead(li * 4 + 2);
            nx = phase_desc(P, li * 4 + 3);
            gemv_phase<WT>(P, c, pd, w, nx, true, [&](int row, float v) { ll_store(P.x_gu + row, __float_as_uint(v), tag_gu); });
            pd = nx;
            MEGA_PROF();
            // ---------------- P4b
            const uint32_t tag_act = ++c.ep;
            mega_silu<AT>(P, c, tag_gu, tag_act);
            MEGA_PROF();
            // ---------------- P5
            __syncthreads();
            mega_gather_act<AT, WT>(P, c, tag_act);
            __syncthreads();
            MEGA_PROF();
            tag_down = ++c.ep;
            prefetch_ahead(li * 4 + 3);
            const bool more = (li + 1 < P.n_layers) || with_head || !last_row;
            nx = (li + 1 < P.n_layers) ? phase_desc(P, li * 4 + 4) : (with_head ? phase_desc(P, 4 * P.n_layers) : phase_desc(P, 0));
            gemv_phase<WT>(P, c, pd, w, nx, more, [&](int row, float v) { ll_store(P.x_down + row, __float_as_uint(v), tag_down); });
            pd = nx;
            MEGA_PROF();
        }
        if (with_head) {
            // final residual + norm + lm_head (tinyllama.cpp:57-58), then argmax (tinyllama.cpp:416-424)
            ll_gather(P.x_down, P.E, tag_down, P.dbg, [&](int i, uint32_t v) { c.xbuf[i] = __uint_as_float(v); });
            __syncthreads();
            pro_norm<AT>(c.av, c.res, c.xbuf, P.final_norm, P.E, c.xbuf, sm.es, c.res, nullptr, nullptr, nullptr);
            __syncthreads();
            MEGA_PROF();
            const uint32_t tag_arg = ++c.ep;
            prefetch_ahead(4 * P.n_layers);
            float best = -INFINITY;
            int arg = 0x7fffffff;
            PhaseDesc nx = phase_desc(P, 0);
            gemv_phase<WT>(P, c, pd, w, nx, !last_row, [&](int row, float v) {
                P.logits[row] = v;
                if (v > best) { best = v; arg = row; }              // rows ascend per thread: first maximum wins
            });
            pd = nx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
            }
            if ((tid & 31) == 0) { sm.bestv[tid >> 5] = best; sm.besti[tid >> 5] = arg; }
            __syncthreads();
            if (tid == 0) {
                for (int q = 1; q < MWARP; q++) {
                    const float ov = sm.bestv[q];
                    const int oi = sm.besti[q];
                    if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
                }
                ll_store(P.x_arg + 2 * c.cta, __float_as_uint(best), tag_arg);
                ll_store(P.x_arg + 2 * c.cta + 1, (uint32_t)arg, tag_arg);
            }
            // every CTA reduces the per-CTA candidates the same way
            float* cv = c.xbuf;
            int* ci = reinterpret_cast<int*>(c.xbuf) + c.G;
            __syncthreads();
            ll_gather(P.x_arg, 2 * c.G, tag_arg, P.dbg, [&](int i, uint32_t v) {
                if (i & 1) ci[i >> 1] = (int)v; else cv[i >> 1] = __uint_as_float(v);
            });
            __syncthreads();
            if (tid < 32) {
                float bv = -INFINITY;
                int bi = 0x7fffffff;
                for (int q = tid; q < c.G; q += 32) {
                    const float ov = cv[q];
                    const int oi = ci[q];
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (tid == 0) sm.besti[0] = (bi == 0x7fffffff) ? 0 : bi;   // all -inf/NaN: the reference leaves index 0
            }
            __syncthreads();
            next_tok = sm.besti[0];
            __syncthreads();
            n_gen++;
            if (c.cta == 0 && tid == 0) P.tokens[pos + 1] = next_tok;
            if (next_tok == P.eos_id) { stop = 1; pos++; break; }
            MEGA_PROF();
        } else {
            next_tok = -1;
        }
    }
    if (c.cta == 0 && tid == 0) {
        P.st->pos = pos;
        P.st->n_gen += n_gen;
        if (stop) P.st->stop = 1;
        *P.epoch = c.ep;
    }
}

}  // namespace gtb
