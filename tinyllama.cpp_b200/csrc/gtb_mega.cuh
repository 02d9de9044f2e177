// gtb_mega.cuh -- the per-token loop of TinyLlama::logits (tinyllama.cpp:45-61, 395-440) as ONE persistent
// cooperative kernel: one CTA per SM, every dependent step of a row ("phase") separated by a data exchange
// through L2 instead of a kernel boundary, rows and greedy steps looped inside the kernel.
//
// Why: batch-1 decode is ~155 dependent GEMVs per token; at 2-5 us per kernel boundary a graph of small
// kernels cannot approach the HBM roofline (SURVEY.md §7 hard part 2).  Here a phase boundary costs a few L2
// round trips:
//   * exchange: every produced fp32 value travels as one 64-bit "LL" word {tag:32 | bits:32}; the tag is a
//     monotone epoch, so a reader can never mistake stale data for fresh data and no fence sits on the
//     critical path.  Readers do not spin on the data (148 CTAs polling 16 KB each saturate L2): one thread
//     per CTA spins on a per-exchange arrival counter (a relaxed RED per producer CTA -- a hint, not a
//     guarantee), the rest of the CTA sleeps in bar.sync, then every thread loads exactly the words it needs
//     and re-reads the rare word whose tag is not there yet.
//   * weights never depend on activations: each thread issues the 128-bit loads of its share of the NEXT
//     phase's weight blocks into registers before it waits, and one thread per CTA pushes the same rows of a
//     later phase into L2 with cp.async.bulk.prefetch -- HBM streaming is decoupled from the dependency chain.
//   * the numeric contract is unchanged (gtb_dev.cuh): exact integer block dots, products
//     parked in shared memory, then one thread per (row, lane) runs the reference's ordered fp32 adds.
//   * the model dimensions are compile-time constants (tinyllama.cpp:12-20); vectors of n_embd elements are
//     held 4 elements per thread ("quad": one staged 32-bit code word per thread, 8 threads per Q8 block), so
//     the per-op re-encodes (App. A) cost ~70 instructions per thread and the residual stream lives in registers.
//   * the kernel is as sensitive to its own size as to the arithmetic: the layer loop executes ~100 KB of code against a 32 KB
//     instruction cache level and sits at the 128-register limit, so profiling stamps are compiled in only for k_mega<WT, true>
//     and anything only one thread needs lives in shared memory (profiles/r02_02_megakernel.md).
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
constexpr int ME = 2048, MF = 5632, MKV = 256, MH = 32, MGSZ = 8;     // TinyLLamaParams, tinyllama.cpp:12-20
constexpr int NBE = ME / 32, NBF = MF / 32;
constexpr int PS_BYTES = 142 * 1024;       // product staging; during attention: probabilities + the unit's V slice as fp32
// (row, block) weight items per thread per tile: MT*IT items in registers.  5 covers q|k|v, o and down in one tile; gate|up runs as two.
// Round 1 held 10 Q4 items (one gate|up tile): 25 more live registers through every phase = 384 B of spills and 840 more instructions
// in a loop that overflows the instruction cache -- 0.968 vs 0.889 ms per token (profiles/r02_02_megakernel.md).
constexpr int IT_Q4 = 5;
constexpr int IT_Q8 = 5;
constexpr int MEGA_MAX_LAYERS = 32;        // the layer table is copied into shared memory (a phase descriptor is then 30 cycles away, not an L2 round trip)
constexpr int SPIN_LIMIT = 1 << 24;        // ~ seconds; a stuck exchange traps instead of hanging the GPU
static_assert(ME == MT * 4, "quad layout: 4 elements of an n_embd vector per thread");

// arrival counters (one 128-byte line each)
enum { CNT_ATTN = 0, CNT_O = 1, CNT_ACT = 2, CNT_DOWN = 3, CNT_ARG = 4, CNT_SC0 = 8, CNT_TOTAL = 8 + MH };
constexpr int CNT_STRIDE = 32;             // in 32-bit words

struct MegaLayer {
    const uint4* w[4];                     // q|k|v, o, gate|up, down (device layout, gtb_internal.h)
    const uint16_t* s[4];
    const uint16_t* attn_norm;
    const uint16_t* ffn_norm;
    uint8_t* kq; uint16_t* ks; uint8_t* vq; uint16_t* vs;
    // chunk-major copies of the Q8 K/V codes and scales, maintained for THIS kernel: kqt / vqt [group][4 chunks][max_ctx][16],
    // kst / vst [group][2 blocks][max_ctx].  Consecutive positions are contiguous there, so a warp's loads are coalesced (the
    // natural rows are 256 B apart: 32 cache lines per warp instruction, which made ISSUING the loads cost 4.4 us per layer)
    uint8_t* kqt; uint8_t* vqt; uint16_t* kst; uint16_t* vst;
};

struct MegaParams {
    int n_layers, n_vocab, max_ctx, sc_stride;
    const MegaLayer* layers;
    const void* emb_w; const uint16_t* emb_s;
    const uint4* head_w; const uint16_t* head_s;
    const uint16_t* final_norm;
    const float* rope_cos; const float* rope_sin;
    ull *x_qkv, *x_sc, *x_attn, *x_o, *x_gu, *x_act, *x_down, *x_arg;
    unsigned int* cnt;
    float* logits;
    int32_t* tokens;
    DevState* st;
    unsigned int* epoch;
    int n_body, n_head, eos_id;
    int pf_ahead;                          // L2 prefetch distance in GEMV phases
    ull* dbg;                              // [8] watchdog diagnostics
    long long* prof;                       // optional: timestamps of the last row, written by CTA prof_cta
    int prof_cta;
};

// ---------------------------------------------------------------- exchange primitives
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
static __device__ __noinline__ void ll_timeout(ull* dbg, uint32_t tag, const void* p, ull seen) {
    if (dbg) {
        dbg[1] = tag; dbg[2] = (ull)p; dbg[3] = seen; dbg[4] = blockIdx.x; dbg[5] = threadIdx.x;
        __threadfence_system();
        dbg[0] = 0xdeadull;
        __threadfence_system();
    }
    __trap();
}
// words [0, 1] at p (16-byte aligned)
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
__device__ __forceinline__ void cnt_add(unsigned int* c, unsigned int v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(c), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int cnt_load(const unsigned int* c) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
    return v;
}
// CTA-wide: (optionally) announce that this CTA's contribution to `sig` is on its way, then wait until counter
// `cnt` has reached `expected`.  One thread spins; everybody else is parked in the barrier.
__device__ __forceinline__ void xwait(unsigned int* sig, const unsigned int* cnt, unsigned int* expect_sm, unsigned int inc, ull* dbg) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // the running expectation lives in shared memory: only this thread ever needs it (six registers less in every thread)
        const unsigned int expected = *expect_sm + inc;
        *expect_sm = expected;
        if (sig) cnt_add(sig, 1u);
        int spins = 0;
        while ((int)(cnt_load(cnt) - expected) < 0) {
            if (++spins > SPIN_LIMIT) ll_timeout(dbg, expected, cnt, cnt_load(cnt));
        }
    }
    __syncthreads();
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

// ---------------------------------------------------------------- quad layout of an n_embd vector
// thread t: block b = t >> 3, word j = t & 7 (half hh = j >> 2, lane l = j & 3) holds the four elements
// e0 + {0, 1, 8, 9}, e0 = 32 b + 16 hh + 2 l: exactly the codes of staged word j of the block (gtb_kernels.cuh).
__device__ __forceinline__ int quad_e0(int t) { return (t >> 3) * 32 + ((t >> 2) & 1) * 16 + (t & 3) * 2; }

// roundf(): half away from zero, for |v| < 2^22 (quants.h:64)
__device__ __forceinline__ float round_away(float v) {
    const float M = 12582912.0f;
    float r = __fsub_rn(__fadd_rn(v, M), M);                    // nearest-even
    const float d = __fsub_rn(v, r);                            // exact
    if (fabsf(d) == 0.5f) r = __fadd_rn(v, copysignf(0.5f, v));
    return r;
}
// Q8 encode of a block held as 8 quads (quants.h:52-66): decoded values, the staged code word and the fp16 scale.
__device__ __forceinline__ void q8_encode_quad(const float x[4], float d[4], uint32_t& word, float& deltaf) {
    float am = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
    am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
    am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
    am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 4));
    const float delta = __fdiv_rn(am, 127.0f);
    deltaf = __half2float(__float2half_rn(delta));
    const float scale = (delta != 0.0f) ? __fdiv_rn(1.0f, delta) : 0.0f;     // from the UNROUNDED delta
    uint32_t cb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float r = round_away(__fmul_rn(x[i], scale));
        d[i] = __fmul_rn(r, deltaf);
        cb[i] = __float_as_uint(__fadd_rn(r, 12582912.0f));     // low byte = two's complement code
    }
    word = __byte_perm(__byte_perm(cb[0], cb[1], 0x0040), __byte_perm(cb[2], cb[3], 0x0040), 0x5410);
}
template <int AT>
__device__ __forceinline__ void roundtrip_quad(const float x[4], float d[4]) {
    if (AT == DT_F16) {
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = f16_roundtrip(x[i]);
    } else {
        uint32_t w; float df;
        q8_encode_quad(x, d, w, df);
    }
}
__device__ __forceinline__ int xs_index(int e) { return (((e >> 6) * 8) + (e & 7)) * 8 + ((e >> 3) & 7); }
// encode + write the staged GEMV input (the ActView layout of gtb_kernels.cuh)
template <int AT, int WT>
__device__ __forceinline__ void stage_quad(const ActView& av, const float x[4]) {
    const int t = threadIdx.x;
    if (AT == DT_F16) {
        const int e0 = quad_e0(t);
        av.xs[xs_index(e0)] = f16_roundtrip(x[0]);
        av.xs[xs_index(e0 + 1)] = f16_roundtrip(x[1]);
        av.xs[xs_index(e0 + 8)] = f16_roundtrip(x[2]);
        av.xs[xs_index(e0 + 9)] = f16_roundtrip(x[3]);
    } else {
        float d[4]; uint32_t w; float df;
        q8_encode_quad(x, d, w, df);
        av.aw[t] = w;                                           // t == b * 8 + j
        if ((t & 7) == 0) av.ad[t >> 3] = df;
        if (WT == DT_Q4) {
            int s = __dp4a((int)w, 0x01010101, 0);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if ((t & 4) == 0) av.ns7[(t >> 3) * 4 + (t & 3)] = -7 * s;
        }
    }
}

// ---------------------------------------------------------------- exact in-order sum, 512 threads x 4 elements per pass
constexpr int ES2_MAXEXP = 96;
struct ExactSum2Smem {
    float wsum[MWARP];
    uint32_t wint[MWARP];                 // per-warp totals of the integer increments
    int wcnt[MWARP];                      // per-warp counts of explicit elements
    uint2 item[ES2_MAXEXP + 1];           // {integer prefix at the element, term bits}, in element order
    float result;
};

struct NoSpill { __device__ __forceinline__ void operator()() const {} };

// In-order fp32 sum of non-negative terms, emulated exactly in parallel.
//
// While the running sum s stays inside one binade (ulp u), adding a term t is an integer step on the BIT PATTERN of s:
// bits(s) += floor(t / u) + (remainder > u/2).  Every term whose step is of that kind ("plain") contributes an integer that does
// not depend on s, so the plain terms between two special ones contribute the DIFFERENCE of one integer prefix sum -- a single
// 32-bit add scan over the CTA, no per-element state.  Special ("explicit") terms are applied as real float adds, in order, by
// one thread: those adjacent to a power-of-two crossing of the running sum (located by an approximate fp32 prefix sum with a
// proven allowance, see below), exact ties (remainder == u/2: round-to-even depends on s), terms larger than the sum, denormals.
// A 2048-term sum has 13-42 explicit terms; more than ES2_MAXEXP falls back to the plain serial chain.
//
// load4(i, q): the four terms starting at element i (i % 4 == 0; elements >= n come back as 0).  In the normal path every thread
// asks only for its own i = 4 * tid (+ pass offset); the serial fallback makes thread 0 ask for every i, so a caller that keeps
// its terms in registers passes `spill`, which writes them where load4 can find them.
template <typename LoadT, typename SpillT = NoSpill>
__device__ __forceinline__ float exact_sum512(LoadT load4, int n, ExactSum2Smem& sm, SpillT spill = SpillT()) {
    // No FP64 here: the prefix sums that locate the running sum's binade are plain fp32 scans.  Terms are non-negative, so a sum
    // taken through any addition tree of depth d is within d * 2^-24 (relative) of the exact one (d <= 16 here), and the strictly
    // sequential chain being emulated is within k * 2^-24 of exact after k terms; the allowance below, (4k + 64) * 2^-24, covers
    // both with room to spare: a plain term's running sum provably stays in the binade its increment was computed for.
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    float s_run = 0.0f;
    float carry = 0.0f;
    for (int p0 = 0; p0 < n; p0 += MT * 4) {
        const int i0 = p0 + tid * 4;
        float tv[4];
        load4(i0, tv);                                     // elements >= n must come back as 0
        const float loc = __fadd_rn(__fadd_rn(tv[0], tv[1]), __fadd_rn(tv[2], tv[3]));
        float inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = __fadd_rn(inc, v); }
        const float lane_pre = __shfl_up_sync(0xffffffffu, inc, 1);   // prefix of the lanes before this one
        if (lane == 31) sm.wsum[wid] = inc;
        __syncthreads();                                   // (1)
        float ws = (lane < MWARP) ? sm.wsum[lane] : 0.0f;
#pragma unroll
        for (int o = 1; o < MWARP; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws = __fadd_rn(ws, v); }
        const float wbase = __shfl_sync(0xffffffffu, ws, (wid + 31) & 31);       // inclusive total of warp wid-1
        const float total_f = __shfl_sync(0xffffffffu, ws, MWARP - 1);
        float A = __fadd_rn(__fadd_rn(carry, (wid > 0) ? wbase : 0.0f), (lane > 0) ? lane_pre : 0.0f);
        // classify: integer increment of a plain term, or explicit
        uint32_t pre[5];                                   // pre[j] = increments of this thread's elements before element j
        pre[0] = 0u;
        uint32_t exmask = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float An = __fadd_rn(A, tv[j]);
            uint32_t add = 0u;
            if (tv[j] != 0.0f) {
                const float rel = (float)(i0 + j + 16) * 0x1p-22f;
                const float lo = __fsub_rn(A, __fmul_rn(A, rel)), hi = __fadd_rn(An, __fmul_rn(An, rel));
                const int eL = (int)(__float_as_uint(lo) >> 23) - 127, eU = (int)(__float_as_uint(hi) >> 23) - 127;
                const uint32_t tb = __float_as_uint(tv[j]);
                const int te = (int)(tb >> 23) - 127;
                bool ex = !(lo > 0x1p-100f) || eL != eU || (tb >> 23) == 0u || te > eL;
                if (!ex) {
                    const int sh = min(eL - te, 25);
                    const uint32_t mt = (tb & 0x7fffffu) | 0x800000u;
                    const uint32_t rem2 = (mt & ((1u << sh) - 1u)) << 1, full = 1u << sh;
                    if (rem2 == full && sh > 0) ex = true;                     // exact tie: round-to-even depends on the running sum
                    else add = (mt >> sh) + (rem2 > full ? 1u : 0u);
                }
                if (ex) exmask |= 1u << j;
            }
            pre[j + 1] = pre[j] + add;
            A = An;
        }
        // one integer add scan over the CTA (wrap-around is harmless: only differences are used) + ranks of the explicit terms
        uint32_t pinc = pre[4];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, pinc, o); if (lane >= o) pinc += v; }
        const uint32_t b0 = __ballot_sync(0xffffffffu, exmask & 1u), b1 = __ballot_sync(0xffffffffu, exmask & 2u);
        const uint32_t b2 = __ballot_sync(0xffffffffu, exmask & 4u), b3 = __ballot_sync(0xffffffffu, exmask & 8u);
        const int rank_w = __popc(b0 & lt_mask) + __popc(b1 & lt_mask) + __popc(b2 & lt_mask) + __popc(b3 & lt_mask);
        if (lane == 31) { sm.wint[wid] = pinc; sm.wcnt[wid] = __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3); }
        __syncthreads();                                   // (2)
        uint32_t wi = (lane < MWARP) ? sm.wint[lane] : 0u;
        int wc = (lane < MWARP) ? sm.wcnt[lane] : 0;
#pragma unroll
        for (int o = 1; o < MWARP; o <<= 1) {
            const uint32_t vi = __shfl_up_sync(0xffffffffu, wi, o);
            const int vc = __shfl_up_sync(0xffffffffu, wc, o);
            if (lane >= o) { wi += vi; wc += vc; }
        }
        const uint32_t ibase = ((wid > 0) ? __shfl_sync(0xffffffffu, wi, (wid + 31) & 31) : 0u) + pinc - pre[4];
        const int cbase = ((wid > 0) ? __shfl_sync(0xffffffffu, wc, (wid + 31) & 31) : 0) + rank_w;
        const uint32_t itotal = __shfl_sync(0xffffffffu, wi, MWARP - 1);
        const int total = __shfl_sync(0xffffffffu, wc, MWARP - 1);
        if (total > ES2_MAXEXP) {                          // pathological input: plain serial chain (still exact)
            spill();
            __syncthreads();
            if (tid == 0) {
                float s = s_run;
                const int hi_i = min(n, p0 + MT * 4);
                for (int i = p0; i < hi_i; i += 4) {
                    float q[4];
                    load4(i, q);
#pragma unroll
                    for (int j = 0; j < 4; j++) s = __fadd_rn(s, q[j]);
                }
                sm.result = s;
            }
            __syncthreads();
            s_run = sm.result;
            carry = __fadd_rn(carry, total_f);
            __syncthreads();
            continue;
        }
        if (exmask) {
            int k = cbase;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (exmask & (1u << j)) sm.item[k++] = make_uint2(ibase + pre[j], __float_as_uint(tv[j]));
        }
        __syncthreads();                                   // (3)
        {
            // every thread walks the explicit items itself (shared-memory broadcasts): no result to publish, one barrier less.  The item
            // buffer is not written again before all threads have passed the next call's first two barriers.
            uint32_t sb = __float_as_uint(s_run), prev = 0u;
            int k = 0;
            for (; k + 4 <= total; k += 4) {
                uint2 it[4];
#pragma unroll
                for (int u = 0; u < 4; u++) it[u] = sm.item[k + u];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    sb += it[u].x - prev;                  // the plain terms since the previous explicit one
                    prev = it[u].x;
                    sb = __float_as_uint(__fadd_rn(__uint_as_float(sb), __uint_as_float(it[u].y)));
                }
            }
            for (; k < total; k++) {
                const uint2 it = sm.item[k];
                sb += it.x - prev;
                prev = it.x;
                sb = __float_as_uint(__fadd_rn(__uint_as_float(sb), __uint_as_float(it.y)));
            }
            sb += itotal - prev;
            s_run = __uint_as_float(sb);
        }
        carry = __fadd_rn(carry, total_f);
    }
    return s_run;
}

// ---------------------------------------------------------------- GEMV phases
struct PhaseDesc {
    const uint4* d;
    const uint16_t* s;
    int nb;        // K / 32
    int kind;      // 0 q|k|v, 1 o, 2 gate|up, 3 down, 4 lm_head
};

struct MegaSm {
    ExactSum2Smem es;
    int rr[5][2];              // this CTA's row range per phase kind
    float raw[192];
    uint32_t qw[16]; float qd[2]; float qf[64];
    uint32_t kw[16]; float kd[2]; float kf[64];
    float vf[64];
    float tmp[6][32];
    float red[MWARP];
    float part[8][16];
    float bestv[MWARP]; int besti[MWARP];
    unsigned int expect[6];    // arrival-counter expectations (xwait): attn, o, act, down, arg, scores of this CTA's head
    MegaLayer layers[MEGA_MAX_LAYERS];
};

__host__ __device__ inline size_t mega_smem_bytes(int at) {
    size_t s = 0;
    s += (size_t)(MF + 64) * 4;                       // xbuf
    s += (act_bytes(at, MF) + 15) & ~(size_t)15;      // staged GEMV input
    s += PS_BYTES;
    s += (sizeof(MegaSm) + 15) & ~(size_t)15;
    return s + 32;
}

template <int WT> struct ItemsPerThread { static constexpr int v = (WT == DT_Q4) ? IT_Q4 : IT_Q8; };

template <int WT>
struct WRegs {
    static constexpr int IT = ItemsPerThread<WT>::v;
    uint4 a[IT];
    uint4 b[(WT == DT_Q8) ? IT : 1];
    uint32_t sc[IT];
};

template <int WT>
__device__ __forceinline__ int tile_rows(int nb) {
    if (nb == NBE) {
        constexpr int by_ps = PS_BYTES / (16 * (NBE + 1));
        constexpr int by_regs = (MT * ItemsPerThread<WT>::v) / NBE;
        constexpr int a = (MT / 4 < by_ps) ? MT / 4 : by_ps;
        return (a < by_regs) ? a : by_regs;
    } else {
        constexpr int by_ps = PS_BYTES / (16 * (NBF + 1));
        constexpr int by_regs = (MT * ItemsPerThread<WT>::v) / NBF;
        constexpr int a = (MT / 4 < by_ps) ? MT / 4 : by_ps;
        return (a < by_regs) ? a : by_regs;
    }
}

__device__ __forceinline__ PhaseDesc phase_desc(const MegaParams& P, const MegaLayer* layers, int s) {
    PhaseDesc pd;
    if (s >= 4 * P.n_layers) {
        pd.d = P.head_w; pd.s = P.head_s; pd.nb = NBE; pd.kind = 4;
        return pd;
    }
    const MegaLayer& L = layers[s >> 2];
    const int k = s & 3;
    pd.d = L.w[k]; pd.s = L.s[k]; pd.nb = (k == 3) ? NBF : NBE; pd.kind = k;
    return pd;
}

// issue the weight loads of one tile: rows [row_begin, row_begin + nrows) -- contiguous blocks, no index math
template <int WT>
__device__ __forceinline__ void load_tile(const PhaseDesc& pd, int row_begin, int nrows, WRegs<WT>& w) {
    constexpr int IT = WRegs<WT>::IT;
    const int nitems = nrows * pd.nb;
    const size_t base = (size_t)row_begin * pd.nb;
    const uint4* dp = pd.d + ((WT == DT_Q8) ? 2 * base : base);
    const uint16_t* sp = pd.s + base;
#pragma unroll
    for (int j = 0; j < IT; j++) {
        const int it = j * MT + threadIdx.x;
        if (it < nitems) {
            if (WT == DT_Q4) {
                w.a[j] = ldg_stream(dp + it);
            } else {
                w.a[j] = ldg_stream(dp + 2 * it);
                w.b[j] = ldg_stream(dp + 2 * it + 1);
            }
            w.sc[j] = ldg_stream_u16(sp + it);
        }
    }
}

// exact integer lane sums of one block, scaled: p[l] = float(lane[l]) * (da * dw)  (gten/ops.h:282-287, 339-378)
template <int WT>
__device__ __forceinline__ float4 block_products_r(const uint4& wa, const uint4& wb, uint32_t sc16, const uint4& ax, const uint4& ay,
                                                   const int4& n7, float ad) {
    const float s = __fmul_rn(ad, h2f((uint16_t)sc16));
    int l0, l1, l2, l3;
    if (WT == DT_Q4) {
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
    // int -> float without the conversion pipe (|lane sum| < 2^22): bits(1.5 * 2^23) + l, minus 1.5 * 2^23
    const float f0 = __fsub_rn(__int_as_float(0x4b400000 + l0), 12582912.0f), f1 = __fsub_rn(__int_as_float(0x4b400000 + l1), 12582912.0f);
    const float f2 = __fsub_rn(__int_as_float(0x4b400000 + l2), 12582912.0f), f3 = __fsub_rn(__int_as_float(0x4b400000 + l3), 12582912.0f);
    return make_float4(__fmul_rn(f0, s), __fmul_rn(f1, s), __fmul_rn(f2, s), __fmul_rn(f3, s));
}

template <int WT>
__device__ __forceinline__ void tile_products(int nb, int nrows, const WRegs<WT>& w, const ActView& av, float* ps) {
    constexpr int IT = WRegs<WT>::IT;
    float4* ps4 = reinterpret_cast<float4*>(ps);
    const int nitems = nrows * nb;
    if (nb == NBE) {
        // K = 2048: the block index of a thread's items never changes -> its activation block stays in registers
        const int b = threadIdx.x & (NBE - 1), rl0 = threadIdx.x >> 6;
        const uint4 ax = reinterpret_cast<const uint4*>(av.aw)[2 * b];
        const uint4 ay = reinterpret_cast<const uint4*>(av.aw)[2 * b + 1];
        const int4 n7 = (WT == DT_Q4) ? reinterpret_cast<const int4*>(av.ns7)[b] : make_int4(0, 0, 0, 0);
        const float ad = av.ad[b];
#pragma unroll
        for (int j = 0; j < IT; j++) {
            if (j * MT + (int)threadIdx.x < nitems)
                ps4[(rl0 + j * (MT / NBE)) * (NBE + 1) + b] = block_products_r<WT>(w.a[j], w.b[(WT == DT_Q8) ? j : 0], w.sc[j], ax, ay, n7, ad);
        }
    } else {
        int rl = threadIdx.x / NBF, b = threadIdx.x - rl * NBF;
        constexpr int drl = MT / NBF, db = MT - drl * NBF;
#pragma unroll
        for (int j = 0; j < IT; j++) {
            if (j * MT + (int)threadIdx.x < nitems) {
                const uint4 ax = reinterpret_cast<const uint4*>(av.aw)[2 * b];
                const uint4 ay = reinterpret_cast<const uint4*>(av.aw)[2 * b + 1];
                const int4 n7 = (WT == DT_Q4) ? reinterpret_cast<const int4*>(av.ns7)[b] : make_int4(0, 0, 0, 0);
                ps4[rl * (NBF + 1) + b] = block_products_r<WT>(w.a[j], w.b[(WT == DT_Q8) ? j : 0], w.sc[j], ax, ay, n7, av.ad[b]);
            }
            b += db; rl += drl;
            if (b >= NBF) { b -= NBF; rl++; }
        }
    }
}

// ordered adds: thread (row, l) sums its lane's products in ascending block order, then (a0+a1)+(a2+a3).
__device__ __forceinline__ float tile_chain(int nb, int nrows, const float* ps) {
    const int ct = threadIdx.x;
    float acc = 0.0f;
    if (ct < nrows * 4) {
        const float* src = ps + (size_t)(ct >> 2) * (nb + 1) * 4 + (ct & 3);
        int b = 0;
        if (nb >= 8) {
            // the chain is one dependent FADD per block (4 cycles); the loads of the next eight blocks are in flight while it runs
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = src[u * 4];
            for (; b + 16 <= nb; b += 8) {
                float nx[8];
#pragma unroll
                for (int u = 0; u < 8; u++) nx[u] = src[(b + 8 + u) * 4];
#pragma unroll
                for (int u = 0; u < 8; u++) acc = __fadd_rn(acc, v[u]);
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = nx[u];
            }
#pragma unroll
            for (int u = 0; u < 8; u++) acc = __fadd_rn(acc, v[u]);
            b += 8;
        }
        for (; b < nb; b++) acc = __fadd_rn(acc, src[b * 4]);
    }
    const float v1 = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    return __fadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, 2));
}

// L2 prefetch of rows [r0, r1) of a phase (one thread)
template <int WT>
__device__ __forceinline__ void prefetch_rows(const PhaseDesc& pd, int r0, int r1) {
    if (r1 <= r0) return;
    if (WT == DT_F16) {
        l2_prefetch(reinterpret_cast<const uint8_t*>(pd.d) + (size_t)r0 * pd.nb * 64, (size_t)(r1 - r0) * pd.nb * 64);
    } else {
        const size_t bb = (WT == DT_Q4) ? 16 : 32;
        l2_prefetch(reinterpret_cast<const uint8_t*>(pd.d) + (size_t)r0 * pd.nb * bb, (size_t)(r1 - r0) * pd.nb * bb);
        l2_prefetch(pd.s + (size_t)r0 * pd.nb, (size_t)(r1 - r0) * pd.nb * 2);
    }
}

// F16 weights: direct streaming, thread (row, lane l of 8), no staging (gten/ops.h:140-160)
__device__ __forceinline__ float f16_row_lane(const uint4* __restrict__ src, int cpr, int l, const ActView& av) {
    float acc = 0.0f;
    constexpr int UN = 8;
    for (int c = 0; c < cpr; c += UN) {                        // cpr is 32 or 88: multiples of 8
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
    return acc;
}

// One GEMV phase over this CTA's rows [r0, r1).  sink(row, value) is called by one thread per finished row.
// For Q4/Q8 `w` holds the first tile's weights on entry; on exit it holds the first tile of `next` (if next_valid).
template <int WT, typename Sink, typename Bg>
__device__ __forceinline__ void gemv_phase(const PhaseDesc& pd, int r0, int r1, const ActView& av, float* ps, WRegs<WT>& w,
                                           const PhaseDesc& next, int n0, int n1, bool next_valid, Sink sink, Bg background,
                                           long long* prof = nullptr, int* prof_ip = nullptr, int prof_base = 0) {
#define GEMV_PROF(code) do { if (prof) { prof[(*prof_ip)++] = prof_base + (code); prof[(*prof_ip)++] = gtimer(); } } while (0)
    if (WT == DT_F16) {
        background();
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, rl = lane >> 3, l = lane & 7;
        const int cpr = pd.nb / 2;                                   // 64-element chunks per row
        for (int row0 = r0 + wid * 4; row0 < r1; row0 += MWARP * 4) {
            const int gr = row0 + rl;
            const bool active = gr < r1;
            const float acc = f16_row_lane(pd.d + (size_t)(active ? gr : r0) * cpr * 8 + l, cpr, l, av);
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
        tile_products<WT>(nb, nrows, w, av, ps);
        GEMV_PROF(0);
        if (t0 + tr < r1) load_tile<WT>(pd, t0 + tr, min(tr, r1 - t0 - tr), w);
        else if (next_valid) load_tile<WT>(next, n0, min(tile_rows<WT>(next.nb), n1 - n0), w);
        GEMV_PROF(1);
        __syncthreads();
        GEMV_PROF(2);
        if (t0 == r0) background();       // off the critical path: runs in a warp that owns no chain while the chains run
        const float v = tile_chain(nb, nrows, ps);
        GEMV_PROF(3);
        if ((threadIdx.x & 3) == 0 && threadIdx.x < nrows * 4) sink(t0 + (threadIdx.x >> 2), v);
        if (t0 + tr < r1) __syncthreads();
    }
    if (r1 <= r0) {
        background();
        if (next_valid) load_tile<WT>(next, n0, min(tile_rows<WT>(next.nb), n1 - n0), w);
    }
#undef GEMV_PROF
}

// ---------------------------------------------------------------- attention, unit = (head h, quarter j)
template <int AT>
__device__ __forceinline__ float mega_score(const MegaSm& sm, const MegaLayer& L, int g, int kcol, int own) {
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
            const uint4* kp = reinterpret_cast<const uint4*>(L.kq + ((size_t)kcol * MKV + g * 64) * 2);
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
                const uint4* kp = reinterpret_cast<const uint4*>(L.kq + (size_t)kcol * MKV + g * 64 + bi * 32);
                kx = __ldcg(kp); ky = __ldcg(kp + 1);
                kdv = h2f(__ldcg(L.ks + (size_t)kcol * (MKV / 32) + g * 2 + bi));
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

// attention scratch carved from the product staging area.  Both arrays are LANE-MAJOR: position i = 8 k + l lives at index k of lane
// l (the position lane of vec_dot_product_f32, ops.h:181-197), so a P.V chain reads four consecutive steps with one 128-bit load.
//   sc [8][kp]            probabilities of the row
//   vf [8][16][kp] (+4/l) the unit's V slice, decoded (exact: 7-bit code x fp16 scale, or fp16)
// kp = padded positions per lane, == 4 (mod 32): 128-bit reads down 8 (lane, channel) rows hit 8 distinct bank groups; every lane
// block is shifted by 4 floats so that the 8 lanes of one position (the staging writes) do too.
struct AttnScratch {
    float* sc;
    float* vf;
    int kp, lst;        // floats per (lane, channel) row; floats per lane block of vf
    __device__ __forceinline__ int sci(int i) const { return (i & 7) * kp + (i >> 3); }
    __device__ __forceinline__ int vfi(int i, int c) const { return (i & 7) * lst + c * kp + (i >> 3); }
};
__host__ __device__ inline int attn_kp(int max_ctx) { return (((max_ctx + 7) / 8 + 31) / 32) * 32 + 4; }
__device__ __forceinline__ AttnScratch attn_scratch(float* ps, int max_ctx) {
    AttnScratch a;
    a.kp = attn_kp(max_ctx);
    a.lst = 16 * a.kp + 4;
    a.sc = ps;
    a.vf = ps + 8 * a.kp;
    return a;
}
__host__ __device__ inline size_t attn_scratch_bytes(int max_ctx) {
    const size_t kp = (size_t)attn_kp(max_ctx);
    return (8 * kp + 8 * (16 * kp + 4)) * 4 + 64;
}

// K row of one cached position for group g, as loaded from the cache (Q8: 64 permuted codes + 2 scales)
struct KRow { uint4 x0, y0, x1, y1; uint32_t s0, s1; };

// P2a: re-encode / RoPE the unit's q, k, v (gten/ops.h:645-646, 733-753), append K/V, publish the scores of quarter j
template <int AT>
__device__ __forceinline__ void mega_attn_a(const MegaParams& P, const MegaLayer& L, MegaSm& sm, float* ps, int cta, int pos,
                                            uint32_t tag_qkv, uint32_t tag_sc, long long* prof, int& prof_i) {
#define SUB_PROF(code) do { if (prof) { prof[prof_i++] = (code); prof[prof_i++] = gtimer(); } } while (0)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int h = cta >> 2, j = cta & 3, g = h / MGSZ;
    const bool writer = (h % MGSZ) == 0 && j == 0;
    const AttnScratch as = attn_scratch(ps, P.max_ctx);
    const int per = (pos + 4) >> 2;
    const int lo = j * per, hi = min(pos + 1, lo + per);
    // Everything that does not depend on this row's q/k/v is issued before the exchange is waited for:
    // (1) the K row of this thread's score position (registers), (2) the unit's V slice, decoded into shared memory.
    const int k0 = lo + tid;
    KRow kr;
    if (AT != DT_F16 && k0 < hi && k0 != pos) {
        const size_t mc = (size_t)P.max_ctx;
        const uint4* kp = reinterpret_cast<const uint4*>(L.kqt + ((size_t)g * 4 * mc + k0) * 16);
        kr.x0 = __ldcg(kp); kr.y0 = __ldcg(kp + mc); kr.x1 = __ldcg(kp + 2 * mc); kr.y1 = __ldcg(kp + 3 * mc);
        kr.s0 = __ldcg(L.kst + (size_t)(g * 2) * mc + k0);
        kr.s1 = __ldcg(L.kst + (size_t)(g * 2 + 1) * mc + k0);
    }
    SUB_PROF(99);                                              // entry barrier + K row loads issued
    {
        const int ch0 = g * 64 + j * 16;
        if (AT == DT_F16) {
            for (int i = tid; i < pos; i += MT) {
                float* dst = as.vf + as.vfi(i, 0);
                const uint4* vp = reinterpret_cast<const uint4*>(L.vq + ((size_t)i * MKV + ch0) * 2);
                const uint4 a = __ldcg(vp), b = __ldcg(vp + 1);
                const __half* ha = reinterpret_cast<const __half*>(&a);
                const __half* hb = reinterpret_cast<const __half*>(&b);
#pragma unroll
                for (int u = 0; u < 8; u++) { dst[u * as.kp] = __half2float(ha[u]); dst[(8 + u) * as.kp] = __half2float(hb[u]); }
            }
        } else {
            // max_ctx <= 4 * MT: at most four positions per thread; ALL their loads are issued before the first is decoded, so the
            // cache misses (the K/V lines were last touched a token ago, 600 MB of weight streaming earlier) overlap instead of queueing
            uint4 cv[4];
            uint32_t dv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = tid + u * MT;
                if (i < pos) {
                    cv[u] = __ldcg(reinterpret_cast<const uint4*>(L.vqt + ((size_t)(g * 4 + j) * P.max_ctx + i) * 16));
                    dv[u] = __ldcg(L.vst + (size_t)(g * 2 + (j >> 1)) * P.max_ctx + i);
                }
            }
            SUB_PROF(100);                                     // V loads issued
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = tid + u * MT;
                if (i < pos) {
                    float* dst = as.vf + as.vfi(i, 0);
                    const float d = h2f((uint16_t)dv[u]);
                    const uint32_t cw[4] = {cv[u].x ^ 0x80808080u, cv[u].y ^ 0x80808080u, cv[u].z ^ 0x80808080u, cv[u].w ^ 0x80808080u};
#pragma unroll
                    for (int w4 = 0; w4 < 4; w4++) {                   // value = code * delta (ops.h:1026), exact in fp32
                        // int8 -> float without the conversion pipe: (2^23 + 128 + code) - (2^23 + 128)
                        const float f0 = __fsub_rn(__uint_as_float(__byte_perm(cw[w4], 0x4b000000u, 0x7650)), 8388736.0f);
                        const float f1 = __fsub_rn(__uint_as_float(__byte_perm(cw[w4], 0x4b000000u, 0x7651)), 8388736.0f);
                        const float f2 = __fsub_rn(__uint_as_float(__byte_perm(cw[w4], 0x4b000000u, 0x7652)), 8388736.0f);
                        const float f3 = __fsub_rn(__uint_as_float(__byte_perm(cw[w4], 0x4b000000u, 0x7653)), 8388736.0f);
                        dst[(4 * w4 + 0) * as.kp] = __fmul_rn(f0, d); dst[(4 * w4 + 1) * as.kp] = __fmul_rn(f1, d);
                        dst[(4 * w4 + 2) * as.kp] = __fmul_rn(f2, d); dst[(4 * w4 + 3) * as.kp] = __fmul_rn(f3, d);
                    }
                }
            }
        }
    }
    SUB_PROF(96);                                              // K row + V slice loads issued / decoded
    if (tid < 96) {
        const int seg = tid >> 5, o = (tid & 31) * 2;
        const ull* src = P.x_qkv + (seg == 0 ? h * 64 : (seg == 1 ? ME + g * 64 : ME + MKV + g * 64)) + o;
        uint32_t a, b;
        ll_wait2(src, tag_qkv, a, b, P.dbg);
        sm.raw[seg * 64 + o] = __uint_as_float(a);
        sm.raw[seg * 64 + o + 1] = __uint_as_float(b);
    }
    __syncthreads();
    SUB_PROF(97);                                              // q, k, v of this row arrived
    if (wid < 6) {
        const int which = wid >> 1, half = wid & 1;
        const float x = sm.raw[which * 64 + half * 32 + lane];
        const int ch = half * 32 + lane;                       // channel inside the head
        const int sl = ch - j * 16;                            // channel inside this unit's P.V slice
        if (which < 2) {
            sm.tmp[wid][lane] = roundtrip<AT>(x);
        } else {
            float d;
            if (AT == DT_F16) {
                const uint16_t hb = f2h(x);
                d = h2f(hb);
                if (writer) reinterpret_cast<uint16_t*>(L.vq)[(size_t)pos * MKV + g * 64 + ch] = hb;
            } else {
                uint16_t dh;
                const int q = q8_encode_lane(x, &dh);
                d = __fmul_rn((float)q, h2f(dh));
                if (writer) {
                    L.vq[(size_t)pos * MKV + g * 64 + ch] = (uint8_t)(int8_t)q;
                    L.vqt[((size_t)(g * 4 + (ch >> 4)) * P.max_ctx + pos) * 16 + (ch & 15)] = (uint8_t)(int8_t)q;
                    if (lane == 0) { L.vs[(size_t)pos * (MKV / 32) + g * 2 + half] = dh; L.vst[(size_t)(g * 2 + half) * P.max_ctx + pos] = dh; }
                }
            }
            if (sl >= 0 && sl < 16) as.vf[as.vfi(pos, sl)] = d;            // this row's own v is not read back from the cache
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
                if (writer) reinterpret_cast<uint16_t*>(L.kq)[(size_t)pos * MKV + g * 64 + half * 32 + lane] = hb;
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
                    L.kq[(size_t)pos * MKV + g * 64 + half * 32 + pb] = (uint8_t)(int8_t)q;
                    L.kqt[((size_t)(g * 4 + half * 2 + (pb >> 4)) * P.max_ctx + pos) * 16 + (pb & 15)] = (uint8_t)(int8_t)q;
                    if (lane == 0) { L.ks[(size_t)pos * (MKV / 32) + g * 2 + half] = dh; L.kst[(size_t)(g * 2 + half) * P.max_ctx + pos] = dh; }
                }
            }
        }
    }
    __syncthreads();
    SUB_PROF(98);                                              // E / RoPE / E done
    // scores of this unit's quarter of the positions, scaled by 1/sqrt(64) (exactly 0.125)
    ull* dst = P.x_sc + (size_t)h * P.sc_stride;
    for (int k = k0; k < hi; k += MT) {
        float d;
        if (AT != DT_F16 && k == k0 && k != pos) {
            // gten/ops.h:224-292 on the prefetched row
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int bi = 0; bi < 2; bi++) {
                const uint4 kx = bi ? kr.x1 : kr.x0, ky = bi ? kr.y1 : kr.y0;
                const float s = __fmul_rn(sm.qd[bi], h2f((uint16_t)(bi ? kr.s1 : kr.s0)));
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
            d = __fadd_rn(__fadd_rn(acc[0], acc[1]), __fadd_rn(acc[2], acc[3]));
        } else {
            d = mega_score<AT>(sm, L, g, k, pos);
        }
        ll_store(dst + k, __float_as_uint(__fmul_rn(d, 0.125f)), tag_sc);
    }
}

// P2b: softmax over the whole row (gten/ops.h:967-996), then P.V for the unit's 16 channels (ops.h:1046-1087).
// Thread t owns positions 4t .. 4t+3 (max_ctx <= 4 * MT): scores, exponentials and probabilities stay in registers.
template <int AT>
__device__ __forceinline__ void mega_attn_b(const MegaParams& P, MegaSm& sm, float* ps, float* xbuf, int cta, int pos, int n_ctx,
                                            uint32_t tag_sc, uint32_t tag_attn, long long* prof, int& prof_i) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int h = cta >> 2, j = cta & 3;
    const AttnScratch as = attn_scratch(ps, P.max_ctx);
    float* sc = as.sc;
    const int i0 = tid * 4;
    float s[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};            // masked / non-existent positions (ops.h:962-964)
    {
        const ull* src = P.x_sc + (size_t)h * P.sc_stride;
        uint32_t a, b;
        if (i0 + 1 <= pos) { ll_wait2(src + i0, tag_sc, a, b, P.dbg); s[0] = __uint_as_float(a); s[1] = __uint_as_float(b); }
        else if (i0 == pos) s[0] = __uint_as_float(ll_wait1(src + i0, tag_sc, P.dbg));
        if (i0 + 3 <= pos) { ll_wait2(src + i0 + 2, tag_sc, a, b, P.dbg); s[2] = __uint_as_float(a); s[3] = __uint_as_float(b); }
        else if (i0 + 2 == pos) s[2] = __uint_as_float(ll_wait1(src + i0 + 2, tag_sc, P.dbg));
    }
    SUB_PROF(80);                                              // scores loaded
    float mx = warp_max(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])));
    if (lane == 0) sm.red[wid] = mx;
    __syncthreads();
    mx = sm.red[0];
#pragma unroll
    for (int w = 1; w < MWARP; w++) mx = fmaxf(mx, sm.red[w]);
    SUB_PROF(81);                                              // max
    float e[4];
#pragma unroll
    for (int u = 0; u < 4; u++) e[u] = (i0 + u <= pos) ? expf_glibc(__fsub_rn(s[u], mx)) : 0.0f;
    SUB_PROF(82);                                              // expf
    const float sum = exact_sum512([&](int i, float q[4]) {
        if (i == i0) { q[0] = e[0]; q[1] = e[1]; q[2] = e[2]; q[3] = e[3]; }
        else {                                               // serial fallback: thread 0 walks every element
#pragma unroll
            for (int u = 0; u < 4; u++) q[u] = xbuf[i + u];
        }
    }, pos + 1, sm.es, [&]() { *reinterpret_cast<float4*>(xbuf + i0) = make_float4(e[0], e[1], e[2], e[3]); });
    SUB_PROF(83);                                              // exact sum
    // probabilities, re-encoded as a row (blocks of 32 positions = 8 consecutive threads); masked entries are exact zeros
    float p[4], ph[4];
#pragma unroll
    for (int u = 0; u < 4; u++) p[u] = (i0 + u <= pos) ? __fdiv_rn(e[u], sum) : 0.0f;
    roundtrip_quad<AT>(p, ph);
    if (i0 <= pos) {                                           // positions 4 t .. 4 t + 3: lanes 4 (t & 1) + u, step t >> 1
#pragma unroll
        for (int u = 0; u < 4; u++) sc[as.sci(i0 + u)] = ph[u];
    }
    __syncthreads();
    SUB_PROF(84);                                              // probabilities encoded
    // P.V: eight position-lanes (i mod 8) per channel over [0, n8), lanes summed left to right, then the tail
    const int n8 = (n_ctx / 8) * 8;
    const float* vf = as.vf;
    if (tid < 128) {
        const int l = tid >> 4, cc = tid & 15;
        const int hi = min(n8, pos + 1);
        const int nk = (hi > l) ? ((hi - l + 7) >> 3) : 0;     // steps of lane l: positions l, l + 8, ... < hi
        const float* pl = sc + l * as.kp;
        const float* vl = vf + l * as.lst + cc * as.kp;
        float a = 0.0f;
        int k = 0;
        if (nk >= 8) {
            // the chain is 4 cycles per step; the 128-bit loads of the NEXT eight steps are issued before the chain of the current eight
            float4 p0 = *reinterpret_cast<const float4*>(pl), p1 = *reinterpret_cast<const float4*>(pl + 4);
            float4 v0 = *reinterpret_cast<const float4*>(vl), v1 = *reinterpret_cast<const float4*>(vl + 4);
            for (; k + 16 <= nk; k += 8) {
                const float4 np0 = *reinterpret_cast<const float4*>(pl + k + 8), np1 = *reinterpret_cast<const float4*>(pl + k + 12);
                const float4 nv0 = *reinterpret_cast<const float4*>(vl + k + 8), nv1 = *reinterpret_cast<const float4*>(vl + k + 12);
                a = __fadd_rn(__fmul_rn(p0.x, v0.x), a); a = __fadd_rn(__fmul_rn(p0.y, v0.y), a);
                a = __fadd_rn(__fmul_rn(p0.z, v0.z), a); a = __fadd_rn(__fmul_rn(p0.w, v0.w), a);
                a = __fadd_rn(__fmul_rn(p1.x, v1.x), a); a = __fadd_rn(__fmul_rn(p1.y, v1.y), a);
                a = __fadd_rn(__fmul_rn(p1.z, v1.z), a); a = __fadd_rn(__fmul_rn(p1.w, v1.w), a);
                p0 = np0; p1 = np1; v0 = nv0; v1 = nv1;
            }
            a = __fadd_rn(__fmul_rn(p0.x, v0.x), a); a = __fadd_rn(__fmul_rn(p0.y, v0.y), a);
            a = __fadd_rn(__fmul_rn(p0.z, v0.z), a); a = __fadd_rn(__fmul_rn(p0.w, v0.w), a);
            a = __fadd_rn(__fmul_rn(p1.x, v1.x), a); a = __fadd_rn(__fmul_rn(p1.y, v1.y), a);
            a = __fadd_rn(__fmul_rn(p1.z, v1.z), a); a = __fadd_rn(__fmul_rn(p1.w, v1.w), a);
            k += 8;
        }
        for (; k < nk; k++) a = __fadd_rn(__fmul_rn(pl[k], vl[k]), a);
        sm.part[l][cc] = a;
    }
    __syncthreads();
    SUB_PROF(85);                                              // P.V lanes
    if (tid < 16) {
        float d = __fadd_rn(sm.part[0][tid], sm.part[1][tid]);
#pragma unroll
        for (int l = 2; l < 8; l++) d = __fadd_rn(d, sm.part[l][tid]);
        for (int i = n8; i < n_ctx && i <= pos; i++) d = __fadd_rn(d, __fmul_rn(sc[as.sci(i)], vf[as.vfi(i, tid)]));
        ll_store(P.x_attn + h * 64 + j * 16 + tid, __float_as_uint(d), tag_attn);
    }
}

#undef SUB_PROF
// P4b: one warp per block of 32 FFN channels: E(E(silu(E(gate))) * E(up)) (gten/modules.cpp:238-247), published as
// packed words in the staged layout (Q8: pairs (X_l, Y_l), the fp16 scale, a pad; F16: 16 half2 words)
template <int AT>
__device__ __forceinline__ void mega_silu(const MegaParams& P, int cta, int G, uint32_t tag_gu, uint32_t tag_act) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b = cta + wid * G; b < NBF; b += G * MWARP) {
        const int e = b * 32 + lane;
        const float g0 = __uint_as_float(ll_wait1(P.x_gu + e, tag_gu, P.dbg));
        const float u0 = __uint_as_float(ll_wait1(P.x_gu + MF + e, tag_gu, P.dbg));
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
            // lane w < 8 assembles staged word w: half = w >> 2, l = w & 3: codes of elements 16*half + {2l, 2l+1, 2l+8, 2l+9}
            const int w = lane & 7, e0 = 16 * (w >> 2) + 2 * (w & 3);
            const uint32_t b0 = __shfl_sync(0xffffffffu, q, e0);
            const uint32_t b1 = __shfl_sync(0xffffffffu, q, e0 + 1);
            const uint32_t b2 = __shfl_sync(0xffffffffu, q, e0 + 8);
            const uint32_t b3 = __shfl_sync(0xffffffffu, q, e0 + 9);
            if (lane < 8) ll_store(P.x_act + (size_t)b * 10 + (w & 3) * 2 + (w >> 2), b0 | (b1 << 8) | (b2 << 16) | (b3 << 24), tag_act);
            if (lane == 8) ll_store(P.x_act + (size_t)b * 10 + 8, (uint32_t)dh, tag_act);
            if (lane == 9) ll_store(P.x_act + (size_t)b * 10 + 9, 0u, tag_act);
        }
        __syncwarp();
        if (lane == 0) cnt_add(P.cnt + CNT_ACT * CNT_STRIDE, 1u);
    }
}

// P5 prologue: the staged GEMV input straight from the packed words
template <int AT, int WT>
__device__ __forceinline__ void mega_gather_act(const MegaParams& P, const ActView& av, uint32_t tag_act) {
    if (AT == DT_F16) {
        for (int p = threadIdx.x; p < NBF * 8; p += MT) {
            uint32_t a, b;
            ll_wait2(P.x_act + 2 * p, tag_act, a, b, P.dbg);
            const int e = 4 * p;
            av.xs[xs_index(e)] = h2f((uint16_t)(a & 0xffffu));
            av.xs[xs_index(e + 1)] = h2f((uint16_t)(a >> 16));
            av.xs[xs_index(e + 2)] = h2f((uint16_t)(b & 0xffffu));
            av.xs[xs_index(e + 3)] = h2f((uint16_t)(b >> 16));
        }
    } else {
        for (int p = threadIdx.x; p < NBF * 5; p += MT) {
            uint32_t a, b;
            ll_wait2(P.x_act + 2 * p, tag_act, a, b, P.dbg);
            const int blk = p / 5, k = p - blk * 5;
            if (k < 4) {
                av.aw[blk * 8 + k] = a;             // X_k
                av.aw[blk * 8 + 4 + k] = b;         // Y_k
                if (WT == DT_Q4) av.ns7[blk * 4 + k] = -7 * __dp4a((int)a, 0x01010101, __dp4a((int)b, 0x01010101, 0));
            } else {
                av.ad[blk] = h2f((uint16_t)a);
            }
        }
    }
}

// token embedding into the residual registers (gten/ops.h:514-564)
template <int WT>
__device__ __forceinline__ void mega_embed(const MegaParams& P, int tok, float res[4]) {
    const int t = threadIdx.x, b = t >> 3, j = t & 7;
    const size_t row = (size_t)tok;
    if (WT == DT_F16) {
        const int e0 = quad_e0(t);
        const uint16_t* src = reinterpret_cast<const uint16_t*>(P.emb_w) + row * ME;
        res[0] = h2f(src[xs_index(e0)]); res[1] = h2f(src[xs_index(e0 + 1)]);
        res[2] = h2f(src[xs_index(e0 + 8)]); res[3] = h2f(src[xs_index(e0 + 9)]);
    } else {
        const size_t blk = row * NBE + b;
        const float delta = h2f(P.emb_s[blk]);
        if (WT == DT_Q8) {
            const uint32_t w = reinterpret_cast<const uint32_t*>(P.emb_w)[blk * 8 + j];       // the row is copied: codes as they are
#pragma unroll
            for (int i = 0; i < 4; i++) res[i] = __fmul_rn((float)(int8_t)((w >> (8 * i)) & 0xffu), delta);
        } else {
            // Q4 row: dequantise, then re-encode as Q8 (ops.h:522-528)
            const uint32_t w = reinterpret_cast<const uint32_t*>(P.emb_w)[blk * 4 + (j & 3)];
            const uint32_t nib = (j < 4) ? ((w >> 4) & 0x0f0f0f0fu) : (w & 0x0f0f0f0fu);
            float x[4];
#pragma unroll
            for (int i = 0; i < 4; i++) x[i] = __fmul_rn((float)((int)((nib >> (8 * i)) & 0xffu) - 7), delta);
            uint32_t cw; float df;
            q8_encode_quad(x, res, cw, df);
        }
    }
}

// profiling stamps of one CTA (option "prof_cta", default 0): (code, globaltimer) pairs; code = phase kind * 16 + step
#define MEGA_PROF(code)                                                                                   \
    do {                                                                                                  \
        if (PROF && profp) { profp[prof_i++] = (code); profp[prof_i++] = gtimer(); }                     \
    } while (0)

// PROF: the stamps are compiled in only for the instantiation the option "prof" selects (they cost issue slots and instruction-cache
// footprint in a kernel whose layer loop already exceeds the instruction cache)
template <int WT, bool PROF>
__global__ void __launch_bounds__(MT, 1) k_mega(const MegaParams P) {
    constexpr int AT = (WT == DT_F16) ? DT_F16 : DT_Q8;
    extern __shared__ __align__(16) unsigned char smem[];
    float* xbuf = reinterpret_cast<float*>(smem);
    const ActView av = act_carve(AT, MF, smem + (size_t)(MF + 64) * 4);
    float* ps = reinterpret_cast<float*>(smem + (size_t)(MF + 64) * 4 + ((act_bytes(AT, MF) + 15) & ~(size_t)15));
    MegaSm& sm = *reinterpret_cast<MegaSm*>(reinterpret_cast<unsigned char*>(ps) + PS_BYTES);
    const int cta = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
    int prof_i = 0;
    long long* const profp = (PROF && P.prof && cta == P.prof_cta && tid == 0) ? P.prof : nullptr;
    unsigned int ep = *reinterpret_cast<volatile unsigned int*>(P.epoch);
    const int n_units = MH * 4;
    const int nL4 = 4 * P.n_layers;                // GEMV phases of a row without the lm_head
    static_assert(sizeof(MegaLayer) % 8 == 0, "layer table is copied as 64-bit words");
    for (int i = tid; i < P.n_layers * (int)(sizeof(MegaLayer) / 8); i += MT)
        reinterpret_cast<ull*>(sm.layers)[i] = reinterpret_cast<const ull*>(P.layers)[i];
    if (tid < 5) {
        const int R = (tid == 0) ? ME + 2 * MKV : ((tid == 2) ? 2 * MF : ((tid == 4) ? P.n_vocab : ME));
        sm.rr[tid][0] = (int)(((long long)cta * R) / G);
        sm.rr[tid][1] = (int)(((long long)(cta + 1) * R) / G);
    }
    // arrival counters only ever grow; what this launch waits for is relative to their value at its start
    if (tid == 0) {
        sm.expect[0] = cnt_load(P.cnt + CNT_ATTN * CNT_STRIDE); sm.expect[1] = cnt_load(P.cnt + CNT_O * CNT_STRIDE);
        sm.expect[2] = cnt_load(P.cnt + CNT_ACT * CNT_STRIDE); sm.expect[3] = cnt_load(P.cnt + CNT_DOWN * CNT_STRIDE);
        sm.expect[4] = cnt_load(P.cnt + CNT_ARG * CNT_STRIDE);
        sm.expect[5] = (cta < n_units) ? cnt_load(P.cnt + (CNT_SC0 + (cta >> 2)) * CNT_STRIDE) : 0u;
    }
    int pos = P.st->pos;
    const int nctx_min = P.st->nctx_min;
    const int n_rows = P.n_body + P.n_head;
    int n_gen = 0, stop = P.st->stop, next_tok = -1;      // an EOS sampled by the multi-row prefill's last row stops this launch too
    __syncthreads();
    WRegs<WT> w;
    {
        const PhaseDesc pd0 = phase_desc(P, sm.layers, 0);
        if (WT != DT_F16) load_tile<WT>(pd0, sm.rr[0][0], min(tile_rows<WT>(pd0.nb), sm.rr[0][1] - sm.rr[0][0]), w);
    }
    if (tid == 0) {
        for (int a = 1; a < P.pf_ahead && a < nL4; a++) prefetch_rows<WT>(phase_desc(P, sm.layers, a), sm.rr[a & 3][0], sm.rr[a & 3][1]);
    }
    float res[4] = {0.0f, 0.0f, 0.0f, 0.0f};      // this thread's quad of the residual stream
    const int e0 = quad_e0(tid);
    // The whole row is ONE loop over GEMV phases with a single copy of every building block: a phase's code runs once
    // per layer and is fetched through the instruction caches every time, so the loop body has to stay small.
    for (int r = 0; r < n_rows && !stop; r++, pos++) {
        const bool with_head = r >= P.n_body;
        const bool last_row = (r == n_rows - 1);
        const int n_ctx = max(nctx_min, pos + 1);
        const int tok = (next_tok >= 0) ? next_tok : __ldcg(P.tokens + pos);
        const int nphase = with_head ? nL4 + 1 : nL4;
        next_tok = -1;
        prof_i = 0;
        MEGA_PROF(255);
        mega_embed<WT>(P, tok, res);
        uint32_t tag_in = 0;                        // tag of the exchange the next prologue consumes
        for (int s = 0; s < nphase; s++) {
            const int kind = (s < nL4) ? (s & 3) : 4;              // the phase descriptor (two pointers) is read from shared memory where it is used
            const MegaLayer& L = sm.layers[min(s >> 2, P.n_layers - 1)];
            // ---------------- prologue: wait for the input vector and stage it as this phase's GEMV input
            if (kind == 3) {
                xwait(nullptr, P.cnt + CNT_ACT * CNT_STRIDE, &sm.expect[2], NBF, P.dbg);
                MEGA_PROF(kind * 16 + 0);
                mega_gather_act<AT, WT>(P, av, tag_in);
            } else {
                float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (s > 0) {
                    const ull* src;
                    if (kind == 1) {
                        xwait((cta < n_units) ? P.cnt + CNT_ATTN * CNT_STRIDE : nullptr, P.cnt + CNT_ATTN * CNT_STRIDE, &sm.expect[0], n_units, P.dbg);
                        src = P.x_attn;
                    } else if (kind == 2) {
                        xwait(P.cnt + CNT_O * CNT_STRIDE, P.cnt + CNT_O * CNT_STRIDE, &sm.expect[1], G, P.dbg);
                        src = P.x_o;
                    } else {
                        xwait(P.cnt + CNT_DOWN * CNT_STRIDE, P.cnt + CNT_DOWN * CNT_STRIDE, &sm.expect[3], G, P.dbg);
                        src = P.x_down;
                    }
                    MEGA_PROF(kind * 16 + 0);
                    uint32_t a0, a1, a2, a3;
                    ll_wait2(src + e0, tag_in, a0, a1, P.dbg);
                    ll_wait2(src + e0 + 8, tag_in, a2, a3, P.dbg);
                    x[0] = __uint_as_float(a0); x[1] = __uint_as_float(a1); x[2] = __uint_as_float(a2); x[3] = __uint_as_float(a3);
                }
                if (kind == 1) {
                    stage_quad<AT, WT>(av, x);                                   // E(attention output)
                } else {
                    // residual add (gten/ops.h:870-898) + RMSNorm (ops.h:762-804)
                    const uint16_t* nw = (kind == 0) ? L.attn_norm : ((kind == 2) ? L.ffn_norm : P.final_norm);
                    const uint32_t nw01 = __ldg(reinterpret_cast<const uint32_t*>(nw + e0));
                    const uint32_t nw23 = __ldg(reinterpret_cast<const uint32_t*>(nw + e0 + 8));
                    if (s > 0) {
                        float d1[4], y[4];
                        roundtrip_quad<AT>(x, d1);
#pragma unroll
                        for (int i = 0; i < 4; i++) y[i] = __fadd_rn(res[i], d1[i]);
                        roundtrip_quad<AT>(y, res);
                    }
                    *reinterpret_cast<float2*>(xbuf + e0) = make_float2(__fmul_rn(res[0], res[0]), __fmul_rn(res[1], res[1]));
                    *reinterpret_cast<float2*>(xbuf + e0 + 8) = make_float2(__fmul_rn(res[2], res[2]), __fmul_rn(res[3], res[3]));
                    __syncthreads();
                    MEGA_PROF(kind * 16 + 1);
                    const float sq_sum = exact_sum512([&](int i, float q[4]) {
                        const float4 v = *reinterpret_cast<const float4*>(xbuf + i);
                        q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
                    }, ME, sm.es);
                    MEGA_PROF(kind * 16 + 2);
                    const float denom = __fadd_rn(sqrtf(__fdiv_rn(sq_sum, (float)ME)), 1e-6f);
                    float y[4];
                    y[0] = __fmul_rn(__fdiv_rn(res[0], denom), h2f((uint16_t)(nw01 & 0xffffu)));
                    y[1] = __fmul_rn(__fdiv_rn(res[1], denom), h2f((uint16_t)(nw01 >> 16)));
                    y[2] = __fmul_rn(__fdiv_rn(res[2], denom), h2f((uint16_t)(nw23 & 0xffffu)));
                    y[3] = __fmul_rn(__fdiv_rn(res[3], denom), h2f((uint16_t)(nw23 >> 16)));
                    stage_quad<AT, WT>(av, y);
                }
            }
            __syncthreads();
            MEGA_PROF(kind * 16 + 3);
            // ---------------- GEMV
            const uint32_t tag = ++ep;
            const int r0 = sm.rr[kind][0], r1 = sm.rr[kind][1];
            // L2 prefetch of this CTA's rows, pf_ahead phases ahead (wraps into the next row).  Issuing the two bulk-prefetch instructions
            // blocks the issuing thread for ~0.45 us whatever their size: they go out from the last warp (it owns no ordered chain) after
            // the product barrier, while the other warps run the chains
            auto background = [&]() {
                if (tid == MT - 32 && P.pf_ahead > 0) {
                    int t = s + P.pf_ahead;
                    bool ok = true;
                    if (t >= nphase) { t -= nphase; ok = !last_row && t < nL4; }
                    if (ok) {
                        const PhaseDesc pf = phase_desc(P, sm.layers, t);
                        prefetch_rows<WT>(pf, sm.rr[pf.kind][0], sm.rr[pf.kind][1]);
                    }
                }
            };
            const bool more = (s + 1 < nphase) || !last_row;
            const PhaseDesc nx = phase_desc(P, sm.layers, (s + 1 < nphase) ? s + 1 : 0);
            ull* outp = (kind == 0) ? P.x_qkv : ((kind == 1) ? P.x_o : ((kind == 2) ? P.x_gu : P.x_down));
            float best = -INFINITY;
            int arg = 0x7fffffff;
            // q|k|v phase of a CTA with an attention unit: the next phase's weight tile (needed only at P3) is loaded AFTER the unit's
            // K/V loads have been issued -- the L1 load queue is in order, and the attention loads are the ones on the critical path
            const bool defer_tile = (WT != DT_F16) && kind == 0 && cta < n_units;
            const PhaseDesc pd = phase_desc(P, sm.layers, s);
            gemv_phase<WT>(pd, r0, r1, av, ps, w, nx, sm.rr[nx.kind][0], sm.rr[nx.kind][1], more && !defer_tile, [&](int row, float v) {
                if (kind == 4) {
                    P.logits[row] = v;
                    if (v > best) { best = v; arg = row; }          // rows ascend per thread: first maximum wins
                } else {
                    ll_store(outp + row, __float_as_uint(v), tag);
                }
            }, background, profp, &prof_i, 128 + kind * 8);
            tag_in = tag;
            MEGA_PROF(kind * 16 + 4);
            // ---------------- what follows the GEMV
            if (kind == 0) {
                const uint32_t tag_sc = ++ep;
                const uint32_t tag_attn = ++ep;
                if (cta < n_units) {
                    __syncthreads();                               // product staging is reused as attention scratch
                    mega_attn_a<AT>(P, L, sm, ps, cta, pos, tag, tag_sc, profp, prof_i);
                    if (defer_tile && more) load_tile<WT>(nx, sm.rr[nx.kind][0], min(tile_rows<WT>(nx.nb), sm.rr[nx.kind][1] - sm.rr[nx.kind][0]), w);
                    MEGA_PROF(kind * 16 + 5);
                    xwait(P.cnt + (CNT_SC0 + (cta >> 2)) * CNT_STRIDE, P.cnt + (CNT_SC0 + (cta >> 2)) * CNT_STRIDE, &sm.expect[5], 4, P.dbg);
                    MEGA_PROF(kind * 16 + 6);
                    mega_attn_b<AT>(P, sm, ps, xbuf, cta, pos, n_ctx, tag_sc, tag_attn, profp, prof_i);
                    __threadfence();                               // K/V appends visible before anything later is published
                } else if (tid == 0 && AT != DT_F16) {
                    // the CTAs without an attention unit pull the NEXT layer's K/V rows (last touched a token ago) towards L2
                    const int nl = (s >> 2) + 1;
                    const MegaLayer& LN = sm.layers[(nl < P.n_layers) ? nl : 0];     // after the last layer: layer 0 of the next row
                    const int nidle = G - n_units, me = cta - n_units;
                    const size_t rows0 = ((size_t)pos * me) / nidle, rows1 = ((size_t)pos * (me + 1)) / nidle;
                    if (rows1 > rows0 && (nl < P.n_layers || !last_row)) {
                        for (int q4 = 0; q4 < 4; q4++) {       // the four groups' K rows; their V slices
#pragma unroll
                            for (int c4 = 0; c4 < 4; c4++) {
                                l2_prefetch(LN.kqt + ((size_t)(q4 * 4 + c4) * P.max_ctx + rows0) * 16, (rows1 - rows0) * 16);
                                l2_prefetch(LN.vqt + ((size_t)(q4 * 4 + c4) * P.max_ctx + rows0) * 16, (rows1 - rows0) * 16);
                            }
                        }
                        for (int b8 = 0; b8 < MKV / 32; b8++) {
                            l2_prefetch(LN.kst + (size_t)b8 * P.max_ctx + rows0, (rows1 - rows0) * 2);
                            l2_prefetch(LN.vst + (size_t)b8 * P.max_ctx + rows0, (rows1 - rows0) * 2);
                        }
                    }
                }
                tag_in = tag_attn;
                MEGA_PROF(kind * 16 + 7);
            } else if (kind == 2) {
                const uint32_t tag_act = ++ep;
                mega_silu<AT>(P, cta, G, tag, tag_act);
                tag_in = tag_act;
                MEGA_PROF(kind * 16 + 8);
            } else if (kind == 4) {
                // argmax (tinyllama.cpp:416-424): per-CTA first maximum, exchanged, reduced identically by every CTA
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
                    if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
                }
                if ((tid & 31) == 0) { sm.bestv[tid >> 5] = best; sm.besti[tid >> 5] = arg; }
                __syncthreads();
                const uint32_t tag_arg = ++ep;
                if (tid == 0) {
                    for (int q = 1; q < MWARP; q++) {
                        const float ov = sm.bestv[q];
                        const int oi = sm.besti[q];
                        if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
                    }
                    ll_store(P.x_arg + 2 * cta, __float_as_uint(best), tag_arg);
                    ll_store(P.x_arg + 2 * cta + 1, (uint32_t)arg, tag_arg);
                }
                xwait(P.cnt + CNT_ARG * CNT_STRIDE, P.cnt + CNT_ARG * CNT_STRIDE, &sm.expect[4], G, P.dbg);
                if (tid < 32) {
                    float bv = -INFINITY;
                    int bi = 0x7fffffff;
                    for (int q = tid; q < G; q += 32) {
                        uint32_t a, b;
                        ll_wait2(P.x_arg + 2 * q, tag_arg, a, b, P.dbg);
                        const float ov = __uint_as_float(a);
                        const int oi = (int)b;
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
                if (cta == 0 && tid == 0) P.tokens[pos + 1] = next_tok;
                if (next_tok == P.eos_id) stop = 1;
                MEGA_PROF(kind * 16 + 9);
            }
        }
    }
    if (cta == 0 && tid == 0) {
        P.st->pos = pos;
        P.st->n_gen += n_gen;
        if (stop) P.st->stop = 1;
        *P.epoch = ep;
    }
}

}  // namespace gtb
