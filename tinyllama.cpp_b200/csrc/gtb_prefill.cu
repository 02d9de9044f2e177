// gtb_prefill.cu -- batched prefill: all T prompt rows of TinyLlama::logits (tinyllama.cpp:45-61) at once.
//
// Every Linear (gten/modules.cpp:44-62 -> ops.h:613-670) of the T rows is ONE GEMM
//     C[T][N] = A[T][K] . W[N][K]^T
// on the tcgen05 tensor cores: A and W tiles (fp16, K-major, 128-byte swizzle) are staged in shared memory by TMA
// through a 4-stage mbarrier ring, one elected thread issues tcgen05.mma (M128 x N256 x K16, fp32 accumulate) into a
// double-buffered TMEM accumulator, and four epilogue warps read it back with tcgen05.ld and apply the reference's
// output re-encode (Q8 blocks, quants.h:52-66) before anything reaches HBM.  The row-wise ops between the GEMMs
// (token_embed, add, rms_norm, rotary_emb, silu, mul: ops.h:514-910) work on "Q8 planar" activations
// (int8 codes [T][D] + fp16 block scales [T][D/32]) with the reference's arithmetic and rounding points, and the
// causal GQA attention (ops.h:930-1133) is a two-pass tensor-core kernel that reproduces the Q8 re-encode of the
// probability rows.  K and V land in the engine's cache layout so that decode continues from position T.
//
// fp16 operands are the dequantised Q8/Q4 values rounded to fp16 (relative error <= 2^-12 per element; FP16 models: exact) and the
// summation order of a dot differs from ops.h:224-391, so this path is tolerance-checked, not bit-checked.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "gtb_prefill.h"

namespace gtb {

// =================================================================================================== PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// A wait that cannot hang the GPU: a broken pipeline traps (the launch fails) after ~2 s instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, fp16 operands, fp32 accumulate; issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread `lane` of the warp receives row (lane_base + lane), columns col..col+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// Programmatic dependent launch: every kernel of the prefill chain lets its successor start launching as soon as all of its
// own CTAs are running (the successor's prologue -- barrier init, TMEM allocation, descriptor prefetch -- then overlaps this
// kernel's tail), and waits for its predecessor to complete and flush before touching global memory.
__device__ __forceinline__ void pdl_trigger_and_wait() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// =================================================================================================== block codecs
// One Q8 block (32 values) held by ONE thread: same operations as quants.h:52-66 / gtb_dev.cuh q8_encode_lane.
__device__ __forceinline__ uint16_t q8_encode32(const float (&x)[32], int (&q)[32]) {
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(x[i]));
    const float delta = __fdiv_rn(amax, 127.0f);
    const float scale = (delta != 0.0f) ? __fdiv_rn(1.0f, delta) : 0.0f;       // from the UNROUNDED delta
#pragma unroll
    for (int i = 0; i < 32; i++) q[i] = (int)roundf(__fmul_rn(x[i], scale));   // half away from zero
    return f2h(delta);
}
// four codes -> one word; only the low byte of each argument is used (two's complement code)
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
// natural order (planar activations, V cache)
__device__ __forceinline__ void store_codes32(int8_t* dst, const uint32_t (&q)[32]) {
    uint4 lo, hi;
    lo.x = pack4(q[0], q[1], q[2], q[3]);     lo.y = pack4(q[4], q[5], q[6], q[7]);
    lo.z = pack4(q[8], q[9], q[10], q[11]);   lo.w = pack4(q[12], q[13], q[14], q[15]);
    hi.x = pack4(q[16], q[17], q[18], q[19]); hi.y = pack4(q[20], q[21], q[22], q[23]);
    hi.z = pack4(q[24], q[25], q[26], q[27]); hi.w = pack4(q[28], q[29], q[30], q[31]);
    reinterpret_cast<uint4*>(dst)[0] = lo;
    reinterpret_cast<uint4*>(dst)[1] = hi;
}
// staged order of the K cache (gtb_kernels.cuh perm_byte): word l of a half = elements (2l, 2l+1, 2l+8, 2l+9)
__device__ __forceinline__ void store_codes32_perm(uint8_t* dst, const uint32_t (&q)[32]) {
    uint4 lo, hi;
    lo.x = pack4(q[0], q[1], q[8], q[9]);     lo.y = pack4(q[2], q[3], q[10], q[11]);
    lo.z = pack4(q[4], q[5], q[12], q[13]);   lo.w = pack4(q[6], q[7], q[14], q[15]);
    hi.x = pack4(q[16], q[17], q[24], q[25]); hi.y = pack4(q[18], q[19], q[26], q[27]);
    hi.z = pack4(q[20], q[21], q[28], q[29]); hi.w = pack4(q[22], q[23], q[30], q[31]);
    reinterpret_cast<uint4*>(dst)[0] = lo;
    reinterpret_cast<uint4*>(dst)[1] = hi;
}
__device__ __forceinline__ int sbyte(uint32_t w, int j) { return (int)(int8_t)((w >> (8 * j)) & 0xffu); }
__device__ __forceinline__ void load_codes32(const int8_t* src, int (&q)[32]) {
    const uint4 lo = reinterpret_cast<const uint4*>(src)[0], hi = reinterpret_cast<const uint4*>(src)[1];
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int i = 0; i < 32; i++) q[i] = sbyte(w[i >> 2], i & 3);
}
__device__ __forceinline__ void load_codes32(const int8_t* src, uint32_t (&q)[32]) {
    const uint4 lo = reinterpret_cast<const uint4*>(src)[0], hi = reinterpret_cast<const uint4*>(src)[1];
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int i = 0; i < 32; i++) q[i] = w[i >> 2] >> (8 * (i & 3));
}
// dequantised block of a planar activation: value = code * fp32(delta) (exact in fp32, quants.h:69-76)
__device__ __forceinline__ void load_deq32(const int8_t* q, const uint16_t* s, size_t row, int D, int b, float (&v)[32]) {
    int c[32];
    load_codes32(q + row * D + (size_t)b * 32, c);
    const float d = h2f(s[row * (D / 32) + b]);
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __fmul_rn((float)c[i], d);
}
__device__ __forceinline__ void load_half32(const __half* src, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint4 w4 = reinterpret_cast<const uint4*>(src)[i];
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
            v[8 * i + 2 * j] = f.x; v[8 * i + 2 * j + 1] = f.y;
        }
    }
}
// block b of row `row` of an activation matrix: Q8 planar (codes + scales) or fp16 [T][D] stored in the same buffer
__device__ __forceinline__ void load_act32(int at, const int8_t* q, const uint16_t* s, size_t row, int D, int b, float (&v)[32]);
__device__ __forceinline__ void store_half32(__half* dst, const float (&v)[32]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) reinterpret_cast<uint4*>(dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}
__device__ __forceinline__ void load_act32(int at, const int8_t* q, const uint16_t* s, size_t row, int D, int b, float (&v)[32]) {
    if (at == DT_F16) load_half32(reinterpret_cast<const __half*>(q) + row * D + (size_t)b * 32, v);
    else load_deq32(q, s, row, D, b, v);
}
__device__ __forceinline__ void store_f32x32(float* dst, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; i++) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}


// =================================================================================================== fused row-wise cores
// Arguments of the GEMM epilogues (and of the stand-alone row-wise kernels that share their arithmetic).
struct PfEpi {
    void* out0 = nullptr;            // EPI_F32: float C[M][N];  EPI_Q8: int8 codes [M][N]
    void* out1 = nullptr;            // EPI_Q8: fp16 block scales [M][N/32]
    // EPI_ROPE (q|k|v projection)
    const float* rope_cos = nullptr; const float* rope_sin = nullptr;
    __half *q16 = nullptr, *k16 = nullptr, *v16 = nullptr;
    uint8_t* kq = nullptr; uint16_t* ks = nullptr; uint8_t* vq = nullptr; uint16_t* vs = nullptr;
    int nh = 0, ng = 0;
    // EPI_SILU (gate|up projection, rows interleaved in groups of 32)
    __half* act16 = nullptr; int F = 0;
    float *cap0 = nullptr, *cap1 = nullptr, *cap2 = nullptr; int capw = 0;
    int at = DT_Q8;                  // activation dtype of the model: DT_Q8 (Q8/Q4 weights) or DT_F16 (tinyllama.cpp:258-265)
    long long* dbg_cycles = nullptr; // experiment: [cta][4] = {total, waiting for TMA data, waiting for a free accumulator, k-blocks} of the MMA thread
    int dbg_same_tile = 0;           // experiment: every TMA load fetches tile (0,0) (L2-resident operands, same shared-memory/MMA work)
};

// __expf and the fast division are each within 2 ulp of the reference's expf and '/': a 1e-7 relative change of silu(x), far below the
// fp16 operand rounding of this path (the exact path keeps glibc's algorithm, gtb_dev.cuh expf_glibc)
__device__ __forceinline__ float pf_silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// Q8 encode of a block held by one thread, decoded value left in place (write_row_from_float + read_row_to_float,
// ops.h:40-96; quants.h:52-66).  Rounding is to nearest via the 1.5 * 2^23 magic number: the code sits in the low byte of
// the biased sum.  This differs from the reference's roundf (half away from zero) only when x * scale lands exactly on
// k + 0.5; the bit-exact path (gtb_dev.cuh) keeps roundf.
template <int AT = DT_Q8>
__device__ __forceinline__ uint16_t pf_roundtrip32(float (&x)[32], uint32_t (&q)[32]) {
    if (AT == DT_F16) {                 // FP16 activations: every element is rounded to fp16 on its own (ops.h:83-90)
#pragma unroll
        for (int i = 0; i < 32; i++) x[i] = f16_roundtrip(x[i]);
        return 0;
    }
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) amax = fmaxf(amax, fabsf(x[i]));
    const float delta = __fdiv_rn(amax, 127.0f);
    const uint16_t dh = f2h(delta);
    const float d = h2f(dh);
    const float scale = (delta != 0.0f) ? __fdiv_rn(1.0f, delta) : 0.0f;       // from the UNROUNDED delta
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const float t = fmaf(x[i], scale, 12582912.0f);
        q[i] = __float_as_uint(t);
        x[i] = __fmul_rn(__fsub_rn(t, 12582912.0f), d);
    }
    return dh;
}

// One head slot of the q|k|v Linear output of row `row`: x0/x1 = the two decoded Q8 blocks, (q0,dh0)/(q1,dh1) their codes.
// q and k heads: RoPE (ops.h:714-760) and the second re-encode; k and v: append to the cache in the engine's layout.
template <int AT>
__device__ __forceinline__ void pf_rope_store(const PfEpi& ep, int row, int slot, float (&x0)[32], float (&x1)[32],
                                              uint32_t (&q0)[32], uint32_t (&q1)[32], uint16_t dh0, uint16_t dh1) {
    const int nh = ep.nh, ng = ep.ng, E = nh * 64, KV = ng * 64;
    const bool is_v = slot >= nh + ng;
    if (!is_v) {
        const float4* cp = reinterpret_cast<const float4*>(ep.rope_cos + (size_t)row * 32);
        const float4* sp = reinterpret_cast<const float4*>(ep.rope_sin + (size_t)row * 32);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 c4 = cp[k], s4 = sp[k];
            const float c[4] = {c4.x, c4.y, c4.z, c4.w}, s[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float a = x0[4 * k + j], b = x1[4 * k + j];
                x0[4 * k + j] = __fsub_rn(__fmul_rn(a, c[j]), __fmul_rn(b, s[j]));
                x1[4 * k + j] = __fadd_rn(__fmul_rn(a, s[j]), __fmul_rn(b, c[j]));
            }
        }
        dh0 = pf_roundtrip32<AT>(x0, q0);
        dh1 = pf_roundtrip32<AT>(x1, q1);
    }
    const int capw = ep.capw;
    constexpr bool f16 = AT == DT_F16;
    if (slot < nh) {
        store_half32(ep.q16 + (size_t)row * E + slot * 64, x0);
        store_half32(ep.q16 + (size_t)row * E + slot * 64 + 32, x1);
        if (ep.cap0) for (int i = 0; i < 32; i++) { ep.cap0[(size_t)row * capw + slot * 64 + i] = x0[i]; ep.cap0[(size_t)row * capw + slot * 64 + 32 + i] = x1[i]; }
    } else if (!is_v) {
        const int g = slot - nh;
        store_half32(ep.k16 + (size_t)row * KV + g * 64, x0);
        store_half32(ep.k16 + (size_t)row * KV + g * 64 + 32, x1);
        if (f16) {                                  // FP16 cache: halves in natural order
            store_half32(reinterpret_cast<__half*>(ep.kq) + (size_t)row * KV + g * 64, x0);
            store_half32(reinterpret_cast<__half*>(ep.kq) + (size_t)row * KV + g * 64 + 32, x1);
        } else {
            store_codes32_perm(ep.kq + (size_t)row * KV + g * 64, q0);
            store_codes32_perm(ep.kq + (size_t)row * KV + g * 64 + 32, q1);
            ep.ks[(size_t)row * (KV / 32) + g * 2] = dh0;
            ep.ks[(size_t)row * (KV / 32) + g * 2 + 1] = dh1;
        }
        if (ep.cap1) for (int i = 0; i < 32; i++) { ep.cap1[(size_t)row * capw + g * 64 + i] = x0[i]; ep.cap1[(size_t)row * capw + g * 64 + 32 + i] = x1[i]; }
    } else {
        const int g = slot - nh - ng;
        store_half32(ep.v16 + (size_t)row * KV + g * 64, x0);
        store_half32(ep.v16 + (size_t)row * KV + g * 64 + 32, x1);
        if (f16) {
            store_half32(reinterpret_cast<__half*>(ep.vq) + (size_t)row * KV + g * 64, x0);
            store_half32(reinterpret_cast<__half*>(ep.vq) + (size_t)row * KV + g * 64 + 32, x1);
        } else {
            store_codes32(reinterpret_cast<int8_t*>(ep.vq) + (size_t)row * KV + g * 64, q0);
            store_codes32(reinterpret_cast<int8_t*>(ep.vq) + (size_t)row * KV + g * 64 + 32, q1);
            ep.vs[(size_t)row * (KV / 32) + g * 2] = dh0;
            ep.vs[(size_t)row * (KV / 32) + g * 2 + 1] = dh1;
        }
        if (ep.cap2) for (int i = 0; i < 32; i++) { ep.cap2[(size_t)row * capw + g * 64 + i] = x0[i]; ep.cap2[(size_t)row * capw + g * 64 + 32 + i] = x1[i]; }
    }
}

// One 32-block of the FFN: g/u = decoded gate and up Linear outputs; SiLU (ops.h:673-711), re-encode, Multiply
// (ops.h:816-867), re-encode -> fp16 operand of the down projection
template <int AT>
__device__ __forceinline__ void pf_silu_store(const PfEpi& ep, int row, int b, float (&g)[32], const float (&u)[32]) {
#pragma unroll
    for (int k = 0; k < 32; k++) g[k] = pf_silu(g[k]);
    uint32_t q[32];
    pf_roundtrip32<AT>(g, q);
#pragma unroll
    for (int k = 0; k < 32; k++) g[k] = __fmul_rn(g[k], u[k]);
    pf_roundtrip32<AT>(g, q);
    store_half32(ep.act16 + (size_t)row * ep.F + (size_t)b * 32, g);
    if (ep.cap0) for (int k = 0; k < 32; k++) { ep.cap0[(size_t)row * ep.capw + b * 32 + k] = g[k]; ep.cap1[(size_t)row * ep.capw + b * 32 + k] = u[k]; }
}

// =================================================================================================== GEMM (tcgen05)
#ifdef GTB_PF_INSTRUMENT                // build with -DGTB_PF_INSTRUMENT to let GTB_PF_CYCLES=1 time the MMA thread (tools/gemm_probe.py)
constexpr bool PF_INSTRUMENT = true;
#else
constexpr bool PF_INSTRUMENT = false;
#endif
constexpr int PF_BM = 128;          // rows of A (prompt positions) per tile = UMMA M
constexpr int PF_BK = 64;           // fp16 elements per k-block = one 128-byte swizzle atom
constexpr int PF_THREADS = 320;     // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int PF_EPI_THREADS = 256; // two epilogue warps per TMEM lane quarter, each takes half of the tile's columns
enum { EPI_F32 = 0, EPI_Q8 = 1, EPI_ROPE = 2, EPI_SILU = 3 };

template <int BN> struct PfGemmCfg {
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr uint32_t A_BYTES = PF_BM * PF_BK * 2;
    static constexpr uint32_t B_BYTES = BN * PF_BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers + tmem slot*/;
    // kind::f16 instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = fp32 (bits 4-5 = 1),
    // A = B = fp16 (0), both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(PF_BM >> 4) << 24);
};

// shared-memory matrix descriptor of a K-major tile with 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B
// apart (SBO), LBO unused (1), descriptor version 1, layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// Epilogue of one accumulator tile for one thread: `tacc` = TMEM address of this thread's row (lane) at the tile's first
// column; the two epilogue warps of a lane quarter split the columns (chalf = 0/1).
template <int BN, int EPI, int AT>
__device__ __forceinline__ void pf_epilogue(const PfEpi& ep, uint32_t tacc, int row, bool row_ok, int col0, int N, int chalf) {
    if (EPI == EPI_F32 || EPI == EPI_Q8) {
#pragma unroll 1
        for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); c++) {
            const int col = col0 + c * 32;
            float v[32];
            tmem_ld32(tacc + (uint32_t)(c * 32), v);
            if (row_ok && col < N) {
                if (EPI == EPI_F32) {
                    store_f32x32(reinterpret_cast<float*>(ep.out0) + (size_t)row * N + col, v);
                } else {
                    uint32_t q[32];
                    const uint16_t dh = pf_roundtrip32<AT>(v, q);        // write_row_from_float, ops.h:645-646
                    if (AT == DT_F16) {
                        store_half32(reinterpret_cast<__half*>(ep.out0) + (size_t)row * N + col, v);
                    } else {
                        store_codes32(reinterpret_cast<int8_t*>(ep.out0) + (size_t)row * N + col, q);
                        reinterpret_cast<uint16_t*>(ep.out1)[(size_t)row * (N / 32) + (col >> 5)] = dh;
                    }
                }
            }
        }
    } else {
        // 64 output columns at a time: one head of q|k|v (EPI_ROPE) or one gate block + its up block (EPI_SILU)
#pragma unroll 1
        for (int c2 = chalf * (BN / 128); c2 < (chalf + 1) * (BN / 128); c2++) {
            const int col = col0 + c2 * 64;
            float a[32], b[32];
            tmem_ld32(tacc + (uint32_t)(c2 * 64), a);
            tmem_ld32(tacc + (uint32_t)(c2 * 64 + 32), b);
            if (row_ok && col < N) {
                uint32_t q0[32], q1[32];
                const uint16_t dh0 = pf_roundtrip32<AT>(a, q0), dh1 = pf_roundtrip32<AT>(b, q1);   // the Linear's own re-encode
                if (EPI == EPI_ROPE) pf_rope_store<AT>(ep, row, col >> 6, a, b, q0, q1, dh0, dh1);
                else pf_silu_store<AT>(ep, row, col >> 6, a, b);
            }
        }
    }
}

template <int BN, int EPI, int AT>
__global__ void __launch_bounds__(PF_THREADS, 1)
k_pf_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, const PfEpi ep) {
    using Cfg = PfGemmCfg<BN>;
    constexpr int ST = Cfg::STAGES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * Cfg::STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto bar_full = [&](int s) { return bar_base + 8u * s; };
    auto bar_empty = [&](int s) { return bar_base + 8u * (ST + s); };
    auto bar_tfull = [&](int a) { return bar_base + 8u * (2 * ST + a); };
    auto bar_tempty = [&](int a) { return bar_base + 8u * (2 * ST + 2 + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mt = (M + PF_BM - 1) / PF_BM, n_nt = (N + BN - 1) / BN;
    const int n_tiles = n_mt * n_nt, nkb = K / PF_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; s++) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), PF_EPI_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger_and_wait();

    if (warp == 0) {
        if (lane == 0) {        // ---------------- TMA producer
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            uint32_t s = 0, ph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int mt = tile % n_mt, nt = tile / n_mt;         // consecutive CTAs share one slice of W
                for (int kb = 0; kb < nkb; kb++) {
                    mbar_wait(bar_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_full(s), Cfg::STAGE_BYTES);
                    const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
                    const int zz = ep.dbg_same_tile ? 0 : 1;
                    tma_load_2d(sa, &tmA, bar_full(s), zz * kb * PF_BK, zz * mt * PF_BM);
                    tma_load_2d(sa + Cfg::A_BYTES, &tmB, bar_full(s), zz * kb * PF_BK, zz * nt * BN);
                    if (++s == ST) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {        // ---------------- MMA issuer
            uint32_t s = 0, ph = 0;
            int it = 0;
            const bool dbg = PF_INSTRUMENT && ep.dbg_cycles != nullptr;
            long long c_full = 0, c_acc = 0, c_n = 0;
            const long long c_t0 = dbg ? clock64() : 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
                const int as = it & 1;
                long long c0 = dbg ? clock64() : 0;
                mbar_wait(bar_tempty(as), ((it >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
                if (dbg) c_acc += clock64() - c0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < nkb; kb++) {
                    if (dbg) c0 = clock64();
                    mbar_wait(bar_full(s), ph);
                    if (dbg) { c_full += clock64() - c0; c_n++; }
                    tc_fence_after();
                    const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < PF_BK / 16; k++)          // UMMA K = 16 fp16 = 32 bytes inside the swizzle atom
                        umma_f16(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), Cfg::IDESC, (kb | k) != 0);
                    umma_commit(bar_empty(s));                      // frees the stage when these MMAs have read it
                    if (kb == nkb - 1) umma_commit(bar_tfull(as));  // accumulator complete
                    if (++s == ST) { s = 0; ph ^= 1; }
                }
            }
            if (dbg) {
                ep.dbg_cycles[blockIdx.x * 4 + 0] = clock64() - c_t0; ep.dbg_cycles[blockIdx.x * 4 + 1] = c_full;
                ep.dbg_cycles[blockIdx.x * 4 + 2] = c_acc; ep.dbg_cycles[blockIdx.x * 4 + 3] = c_n;
            }
        }
        __syncwarp();
    } else {                    // ---------------- epilogue warps: TMEM -> registers -> re-encode -> HBM
        const int lg = warp & 3;                                    // a warp may only touch TMEM lanes 32*(warp%4)..+31
        const int chalf = (warp - 2) >> 2;                          // which half of the tile's columns this warp re-encodes
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const int mt = tile % n_mt, nt = tile / n_mt;
            const int as = it & 1;
            mbar_wait(bar_tfull(as), (it >> 1) & 1);
            tc_fence_after();
            const int row = mt * PF_BM + lg * 32 + lane;
            const bool row_ok = row < M;
            pf_epilogue<BN, EPI, AT>(ep, tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN), row, row_ok, nt * BN, N, chalf);
            tc_fence_before();
            mbar_arrive(bar_tempty(as));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------------------------------------------------------
// The same GEMM on CTA PAIRS (tcgen05 cta_group::2).  A pair computes a 256 x 256 tile: each CTA stages its own 128 rows of A
// and HALF of the W tile (32 KB instead of 48 KB per k-block, so six stages fit), the leader CTA issues M256 MMAs that read
// both halves, and each CTA's TMEM receives its 128 rows.  Shared-memory traffic per SM drops by a third (l1tex throughput
// 62 % -> 44 %).  Measured: bit-exact on the integer tests and exactly as fast as k_pf_gemm on all four prefill shapes
// (profiles/r01_02_prefill.md) -- the single-CTA main loop is already at 90-94 % of the nominal tensor rate, what is lost is
// outside it -- so the engine keeps k_pf_gemm by default and this kernel behind the option "pf_2cta".
//   full[s]   (leader only, 2 arrivals): both producers arrive; both CTAs' TMA bytes complete_tx on the LEADER's barrier
//   empty[s]  (each CTA, 1 arrival)    : the leader's tcgen05.commit multicasts to both CTAs
//   tfull[a]  (each CTA, 1 arrival)    : commit multicast after the last k-block of a tile
//   tempty[a] (leader only, 512)       : the epilogue threads of BOTH CTAs arrive on the leader's barrier
constexpr int PF2_STAGES = 6;
constexpr uint32_t PF2_STAGE_BYTES = PF_BM * PF_BK * 2 + 128 * PF_BK * 2;
constexpr uint32_t PF2_SMEM = PF2_STAGES * PF2_STAGE_BYTES + 1024 + 256;

template <int EPI, int AT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PF_THREADS, 1)
k_pf_gemm2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, const PfEpi ep) {
    constexpr int BN = 256;
    constexpr int ST = PF2_STAGES;
    constexpr uint32_t A_BYTES = PF_BM * PF_BK * 2, STAGE_BYTES = PF2_STAGE_BYTES;
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);      // M = 256 across the pair
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto bar_full = [&](int s) { return bar_base + 8u * s; };
    auto bar_empty = [&](int s) { return bar_base + 8u * (ST + s); };
    auto bar_tfull = [&](int a) { return bar_base + 8u * (2 * ST + a); };
    auto bar_tempty = [&](int a) { return bar_base + 8u * (2 * ST + 2 + a); };

    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mt = (M + 255) / 256, n_nt = (N + BN - 1) / BN;
    const int n_tiles = n_mt * n_nt, nkb = K / PF_BK;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; s++) { mbar_init(bar_full(s), 2); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), 2 * PF_EPI_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger_and_wait();

    if (warp == 0) {
        if (lane == 0) {        // ---------------- TMA producer (both CTAs): own rows of A, own half of the W tile
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            uint32_t s = 0, ph = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
                const int mt = tile % n_mt, nt = tile / n_mt;
                for (int kb = 0; kb < nkb; kb++) {
                    mbar_wait(bar_empty(s), ph ^ 1);
                    const uint32_t lbar = bar_full(s) & 0xFEFFFFFFu;            // the leader CTA's barrier (peer bit cleared)
                    if (leader) {
                        mbar_expect_tx(bar_full(s), 2 * STAGE_BYTES);
                    } else {
                        asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                                     ::"r"(bar_full(s)) : "memory");
                    }
                    const uint32_t sa = smem_base + s * STAGE_BYTES;
                    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(sa), "l"(&tmA), "r"(lbar), "r"(kb * PF_BK), "r"(mt * 256 + (int)rank * PF_BM) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(sa + A_BYTES), "l"(&tmB), "r"(lbar), "r"(kb * PF_BK), "r"(nt * BN + (int)rank * (BN / 2)) : "memory");
                    if (++s == ST) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {   // ---------------- MMA issuer (leader CTA only)
            uint32_t s = 0, ph = 0;
            int it = 0;
            const bool dbg = PF_INSTRUMENT && ep.dbg_cycles != nullptr;
            long long c_full = 0, c_acc = 0, c_n = 0;
            const long long c_t0 = dbg ? clock64() : 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
                const int as = it & 1;
                long long c0 = dbg ? clock64() : 0;
                mbar_wait(bar_tempty(as), ((it >> 1) & 1) ^ 1);
                if (dbg) c_acc += clock64() - c0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < nkb; kb++) {
                    if (dbg) c0 = clock64();
                    mbar_wait(bar_full(s), ph);
                    if (dbg) { c_full += clock64() - c0; c_n++; }
                    tc_fence_after();
                    const uint32_t sa = smem_base + s * STAGE_BYTES;
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < PF_BK / 16; k++) {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(d_tmem), "l"(umma_desc_sw128(sa + k * 32)), "l"(umma_desc_sw128(sb + k * 32)), "r"(IDESC), "r"((uint32_t)((kb | k) != 0)) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(bar_empty(s)), "h"((uint16_t)3) : "memory");
                    if (kb == nkb - 1)
                        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                     ::"r"(bar_tfull(as)), "h"((uint16_t)3) : "memory");
                    if (++s == ST) { s = 0; ph ^= 1; }
                }
            }
            if (dbg) {
                ep.dbg_cycles[blockIdx.x * 4 + 0] = clock64() - c_t0; ep.dbg_cycles[blockIdx.x * 4 + 1] = c_full;
                ep.dbg_cycles[blockIdx.x * 4 + 2] = c_acc; ep.dbg_cycles[blockIdx.x * 4 + 3] = c_n;
            }
        }
        __syncwarp();
    } else {                    // ---------------- epilogue warps (both CTAs): own 128 rows
        const int lg = warp & 3;
        const int chalf = (warp - 2) >> 2;
        int it = 0;
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, it++) {
            const int mt = tile % n_mt, nt = tile / n_mt;
            const int as = it & 1;
            mbar_wait(bar_tfull(as), (it >> 1) & 1);
            tc_fence_after();
            const int row = mt * 256 + (int)rank * PF_BM + lg * 32 + lane;
            pf_epilogue<BN, EPI, AT>(ep, tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN), row, row < M, nt * BN, N, chalf);
            tc_fence_before();
            asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                         ::"r"(bar_tempty(as)) : "memory");
        }
    }
    tc_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
}

// =================================================================================================== row-wise kernels
// weights: device layout (gtb_internal.h) -> fp16 [rows][cols], value = fp16(code * delta)
// interleave_half > 0 (gate|up, = n_ffn): source row r of the first half goes to row (r/32)*64 + r%32, row r of the second
// half to (r/32)*64 + 32 + r%32, so that 64 consecutive output columns of the GEMM hold one gate block and its up block
__global__ void k_pf_w16(const void* __restrict__ data, const uint16_t* __restrict__ scales, int wdtype, size_t nblocks,
                         __half* __restrict__ out, int bpr, int interleave_half) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    size_t o = i;
    if (interleave_half > 0) {
        const size_t row = i / bpr, kb = i % bpr;
        const size_t r = row % interleave_half, up = row / interleave_half;
        o = ((r / 32) * 64 + up * 32 + (r % 32)) * bpr + kb;
    }
    const float d = h2f(scales[i]);
    float v[32];
    if (wdtype == DT_Q4) {
        const uint4 w4 = reinterpret_cast<const uint4*>(data)[i];
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int e = 0; e < 16; e++) {                    // payload byte j = e: word (j&7)>>1, position (j&1) + 2*(j>>3)
            const int l = (e & 7) >> 1, pos = (e & 1) + 2 * (e >> 3);
            const uint32_t byte = (w[l] >> (8 * pos)) & 0xffu;
            v[e] = __fmul_rn((float)((int)(byte >> 4) - 7), d);
            v[e + 16] = __fmul_rn((float)((int)(byte & 0x0fu) - 7), d);
        }
    } else {
        const uint4 x4 = reinterpret_cast<const uint4*>(data)[2 * i], y4 = reinterpret_cast<const uint4*>(data)[2 * i + 1];
        const uint32_t w[8] = {x4.x, x4.y, x4.z, x4.w, y4.x, y4.y, y4.z, y4.w};
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int half = e >> 4, j = e & 15, l = (j & 7) >> 1, pos = (j & 1) + 2 * (j >> 3);
            v[e] = __fmul_rn((float)sbyte(w[half * 4 + l], pos), d);
        }
    }
    store_half32(out + o * 32, v);
}

// FP16 weights: lane-major device layout (gtb_internal.h) -> row-major [rows][cols]; one thread per 8 consecutive elements
__global__ void k_pf_w16_f16(const __half* __restrict__ data, size_t ngroups, __half* __restrict__ out, int cols, int interleave_half) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;          // (row, chunk c, i8): elements 64c + 8*i8 .. +7
    if (i >= ngroups) return;
    const int gpr = cols / 8;
    const size_t row = i / gpr;
    const int gi = (int)(i % gpr), c = gi >> 3, i8 = gi & 7;
    const __half* src = data + (row * (cols / 64) + c) * 64;
    __half v[8];
#pragma unroll
    for (int l = 0; l < 8; l++) v[l] = src[l * 8 + i8];
    size_t orow = row;
    if (interleave_half > 0) {
        const size_t r = row % interleave_half, up = row / interleave_half;
        orow = (r / 32) * 64 + up * 32 + (r % 32);
    }
    *reinterpret_cast<uint4*>(out + orow * cols + 64 * c + 8 * i8) = *reinterpret_cast<const uint4*>(v);
}

// token_embed (ops.h:514-564): Q8 rows are copied, Q4 rows are dequantised and re-encoded as Q8
__global__ void k_pf_embed(const void* __restrict__ wdata, const uint16_t* __restrict__ wsc, int wdtype, const int32_t* __restrict__ tokens,
                           int T, int D, int8_t* __restrict__ xq, uint16_t* __restrict__ xs, float* cap, int capw) {
    pdl_trigger_and_wait();
    const int nb = D / 32;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)T * nb) return;
    const int row = (int)(i / nb), b = (int)(i % nb);
    if (wdtype == DT_F16) {                                  // memcpy of the fp16 row (ops.h:535-540), un-permuted from the lane-major layout
        const __half* src = reinterpret_cast<const __half*>(wdata) + ((size_t)tokens[row] * (D / 64) + (b >> 1)) * 64;
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int r = (b & 1) * 32 + e;
            v[e] = __half2float(src[(r & 7) * 8 + (r >> 3)]);
        }
        store_half32(reinterpret_cast<__half*>(xq) + (size_t)row * D + (size_t)b * 32, v);
        if (cap) for (int e = 0; e < 32; e++) cap[(size_t)row * capw + b * 32 + e] = v[e];
        return;
    }
    const size_t blk = (size_t)tokens[row] * nb + b;
    const float d = h2f(wsc[blk]);
    int q[32];
    uint16_t dh;
    if (wdtype == DT_Q4) {
        const uint4 w4 = reinterpret_cast<const uint4*>(wdata)[blk];
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
        float v[32];
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int l = (e & 7) >> 1, pos = (e & 1) + 2 * (e >> 3);
            const uint32_t byte = (w[l] >> (8 * pos)) & 0xffu;
            v[e] = __fmul_rn((float)((int)(byte >> 4) - 7), d);
            v[e + 16] = __fmul_rn((float)((int)(byte & 0x0fu) - 7), d);
        }
        dh = q8_encode32(v, q);
    } else {
        const uint4 x4 = reinterpret_cast<const uint4*>(wdata)[2 * blk], y4 = reinterpret_cast<const uint4*>(wdata)[2 * blk + 1];
        const uint32_t w[8] = {x4.x, x4.y, x4.z, x4.w, y4.x, y4.y, y4.z, y4.w};
#pragma unroll
        for (int e = 0; e < 32; e++) {
            const int half = e >> 4, j = e & 15, l = (j & 7) >> 1, pos = (j & 1) + 2 * (j >> 3);
            q[e] = sbyte(w[half * 4 + l], pos);
        }
        dh = wsc[blk];
    }
    uint32_t qb[32];
#pragma unroll
    for (int e = 0; e < 32; e++) qb[e] = (uint32_t)q[e];
    store_codes32(xq + (size_t)row * D + (size_t)b * 32, qb);
    xs[(size_t)row * nb + b] = dh;
    if (cap) {
        const float dd = h2f(dh);
        for (int e = 0; e < 32; e++) cap[(size_t)row * capw + b * 32 + e] = __fmul_rn((float)q[e], dd);
    }
}

// Residual (ops.h:870-910) + RMSNorm (ops.h:762-814), one 64-thread CTA per row, one 32-block per thread and pass:
//   x <- E(x + y)        (skipped when y == nullptr)
//   xn <- fp16( E( x / (rms(x) + 1e-6) * w ) )       the A operand of the next GEMM
// The sum of squares is a plain parallel fp32 sum (the reference sums in element order): its relative error of ~1e-7
// is three orders of magnitude below the fp16 operand rounding of the GEMMs of this path.
template <int AT>
__global__ void __launch_bounds__(64, 14) k_pf_add_norm(int8_t* __restrict__ xq, uint16_t* __restrict__ xs, const int8_t* __restrict__ yq,
                                                     const uint16_t* __restrict__ ys, const uint16_t* __restrict__ normw,
                                                     __half* __restrict__ xn16, int T, int D, float* cap_y, float* cap_x, float* cap_n, int capw) {
    pdl_trigger_and_wait();
    __shared__ float red[2];
    const int row = blockIdx.x, tid = threadIdx.x;
    const int nb = D / 32;
    const bool single = nb <= 64;                     // n_embd <= 2048: the block stays in registers between the passes
    float ssq = 0.0f;
    float v[32];
    for (int b = tid; b < nb; b += 64) {
        load_act32(AT, xq, xs, row, D, b, v);
        if (yq) {
            float y[32];
            load_act32(AT, yq, ys, row, D, b, y);
            if (cap_y) for (int i = 0; i < 32; i++) cap_y[(size_t)row * capw + b * 32 + i] = y[i];
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = __fadd_rn(v[i], y[i]);
            uint32_t q[32];
            const uint16_t dh = pf_roundtrip32<AT>(v, q);
            if (AT == DT_F16) {
                store_half32(reinterpret_cast<__half*>(xq) + (size_t)row * D + (size_t)b * 32, v);
            } else {
                store_codes32(xq + (size_t)row * D + (size_t)b * 32, q);
                xs[(size_t)row * nb + b] = dh;
            }
        }
        if (cap_x) for (int i = 0; i < 32; i++) cap_x[(size_t)row * capw + b * 32 + i] = v[i];
#pragma unroll
        for (int i = 0; i < 32; i++) ssq = __fadd_rn(ssq, __fmul_rn(v[i], v[i]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssq = __fadd_rn(ssq, __shfl_xor_sync(0xffffffffu, ssq, o));
    if ((tid & 31) == 0) red[tid >> 5] = ssq;
    __syncthreads();
    ssq = __fadd_rn(red[0], red[1]);
    const float rms = sqrtf(__fdiv_rn(ssq, (float)D));
    const float denom = __fadd_rn(rms, 1e-6f);
    for (int b = tid; b < nb; b += 64) {
        if (!single) load_act32(AT, xq, xs, row, D, b, v);
        const uint4* wp = reinterpret_cast<const uint4*>(normw + b * 32);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 w4 = wp[k];
            const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float wf = h2f((uint16_t)((w[j >> 1] >> (16 * (j & 1))) & 0xffffu));
                v[k * 8 + j] = __fmul_rn(__fdiv_rn(v[k * 8 + j], denom), wf);
            }
        }
        uint32_t q[32];
        pf_roundtrip32<AT>(v, q);
        if (cap_n) for (int i = 0; i < 32; i++) cap_n[(size_t)row * capw + b * 32 + i] = v[i];
        store_half32(xn16 + (size_t)row * D + (size_t)b * 32, v);
    }
}

// Unfused fallbacks of the two fused epilogues ("pf_fused" = 0): same arithmetic on the planar GEMM output.
// One thread per (row, head slot): slots [0,nh) = q heads, [nh, nh+ng) = k heads, [nh+ng, nh+2ng) = v heads.
__global__ void __launch_bounds__(128) k_pf_rope_kv(const int8_t* __restrict__ cq, const uint16_t* __restrict__ cs, int T, const PfEpi ep) {
    pdl_trigger_and_wait();
    const int nslots = ep.nh + 2 * ep.ng;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)T * nslots) return;
    const int row = (int)(idx / nslots), slot = (int)(idx % nslots);
    const int D = nslots * 64;
    float x0[32], x1[32];
    uint32_t q0[32], q1[32];
    load_deq32(cq, cs, row, D, slot * 2, x0);
    load_deq32(cq, cs, row, D, slot * 2 + 1, x1);
    load_codes32(cq + (size_t)row * D + (size_t)slot * 64, q0);
    load_codes32(cq + (size_t)row * D + (size_t)slot * 64 + 32, q1);
    pf_rope_store<DT_Q8>(ep, row, slot, x0, x1, q0, q1, cs[(size_t)row * (D / 32) + slot * 2], cs[(size_t)row * (D / 32) + slot * 2 + 1]);
}

// gate|up planar output with the rows of the two matrices interleaved in groups of 32: block 2b = gate block b, 2b+1 = up block b
__global__ void __launch_bounds__(128) k_pf_silu_mul(const int8_t* __restrict__ gq, const uint16_t* __restrict__ gs, int T, const PfEpi ep) {
    pdl_trigger_and_wait();
    const int nb = ep.F / 32;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)T * nb) return;
    const int row = (int)(i / nb), b = (int)(i % nb);
    float g[32], u[32];
    load_deq32(gq, gs, row, 2 * ep.F, 2 * b, g);
    load_deq32(gq, gs, row, 2 * ep.F, 2 * b + 1, u);
    pf_silu_store<DT_Q8>(ep, row, b, g, u);
}

// dequantised fp32 copies of row `row` for the exact final-norm + lm_head phase of the engine
__global__ void k_pf_tail(const int8_t* __restrict__ xq, const uint16_t* __restrict__ xs, const int8_t* __restrict__ dq, const uint16_t* __restrict__ ds,
                          int row, int D, float* __restrict__ res, float* __restrict__ down, float* cap_down, int T, int capw, int at) {
    pdl_trigger_and_wait();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D) return;
    auto val = [&](const int8_t* q, const uint16_t* sc, int r) -> float {
        if (at == DT_F16) return __half2float(reinterpret_cast<const __half*>(q)[(size_t)r * D + e]);
        return __fmul_rn((float)q[(size_t)r * D + e], h2f(sc[(size_t)r * (D / 32) + (e >> 5)]));
    };
    res[e] = val(xq, xs, row);
    down[e] = val(dq, ds, row);
    if (cap_down)
        for (int r = 0; r < T; r++) cap_down[(size_t)r * capw + e] = val(dq, ds, r);
}

// =================================================================================================== attention
// Causal GQA attention of all T rows (ops.h:930-1133).  One CTA = 64 query rows of one head, 4 warps x 16 rows,
// keys in tiles of 64, mma.sync m16n8k16 (fp16 in, fp32 accumulate).  Pass A: row maximum and sum of exp.  Pass B:
// p = exp(s - max) / sum, the Q8 re-encode of the probability row in 32-key blocks (ops.h:996), then the integer
// codes (exact in fp16) go through the tensor cores against V and the block's fp16 scale is applied to the partial
// sum, so the only operand rounding on this side is V's.  The output row is re-encoded per 32 channels (ops.h:1084).
constexpr int PA_LD = 72;           // shared-memory row stride in halves (64 + 8: conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rne_int(float x) {        // nearest integer for 0 <= x < 2^22 (ties to even)
    return __fsub_rn(__fadd_rn(x, 12582912.0f), 12582912.0f);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;                          // 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 64 x 64 fp16 tile (rows r0.., 64 columns starting at column c0 of a [T][ld] matrix) -> shared [64][PA_LD], asynchronously;
// rows >= T are zero
__device__ __forceinline__ void pa_load_tile_async(uint32_t dst, const __half* __restrict__ src, int r0, int T, int ld, int c0) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int idx = threadIdx.x + 128 * i, r = idx >> 3, ch = idx & 7;
        const bool ok = r0 + r < T;
        const int rr = ok ? r0 + r : T - 1;
        cp_async16(dst + (uint32_t)((r * PA_LD + ch * 8) * 2), src + (size_t)rr * ld + c0 + ch * 8, ok);
    }
}

// raw S[j][*] = Q_warp . K_tile^T for the 8 key octets j of the tile, -inf above the diagonal (ops.h:966-969)
__device__ __forceinline__ void pa_scores(float (&S)[8][4], const uint32_t (&qa)[4][4], uint32_t ks_addr, int lane, bool diag, int qrow0, int key0) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        S[j][0] = S[j][1] = S[j][2] = S[j][3] = 0.0f;
#pragma unroll
        for (int kp = 0; kp < 2; kp++) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(ks_addr + (uint32_t)(((8 * j + (lane & 7)) * PA_LD + 32 * kp + (lane >> 3) * 8) * 2), b0, b1, b2, b3);
            mma_16816(S[j], qa[2 * kp], b0, b1);
            mma_16816(S[j], qa[2 * kp + 1], b2, b3);
        }
    }
    if (diag) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int key = key0 + 8 * j + 2 * t + (c & 1), qrow = qrow0 + g + ((c >> 1) << 3);
                if (key > qrow) S[j][c] = -INFINITY;
            }
    }
}

// scores are kept unscaled; exp(0.125 * (s - max)) = 2^((s - max) * PA_C)    (scale = 1/sqrt(64), ops.h:1098)
#define PA_C (0.125f * 1.4426950408889634f)

// One kernel, two variants:
//  TWO_PASS = true : pass A computes each row's maximum and sum of exp, pass B the probabilities p = e / sum, the Q8 re-encode
//                    of the probability row per 32 keys INCLUDING the fp16 rounding of each block scale (ops.h:996), and P.V.
//  TWO_PASS = false: one sweep with a running maximum (the flash-attention recurrence).  The integer codes of a block depend
//                    only on e / max(e of the block), so they are the same; the block scale is applied unrounded
//                    (max_e / 127, rescaled with the running maximum, divided by the final sum at the end), i.e. the one
//                    rounding point this variant does not reproduce is the fp16 rounding of the P-row block scales
//                    (relative 2^-11 per block, averaging out over the blocks of a row -- the same size as the fp16 rounding
//                    of the V operand).  QK^T and the exps run once instead of twice; the decoded probabilities
//                    (code x scale) are the fp16 A operand of P.V, accumulated in one set of registers.
template <bool TWO_PASS, int AT>
__global__ void __launch_bounds__(128, TWO_PASS ? 3 : 4) k_pf_attn(const __half* __restrict__ q16, const __half* __restrict__ k16, const __half* __restrict__ v16,
                                                  __half* __restrict__ out16, int T, int n_heads, int gsz, float* cap, int capw) {
    pdl_trigger_and_wait();
    __shared__ __align__(16) __half Qs[64 * PA_LD];
    __shared__ __align__(16) __half Ks[2][64 * PA_LD];
    __shared__ __align__(16) __half Vs[2][64 * PA_LD];
    const int n_qt = (T + 63) / 64;
    const int qt = n_qt - 1 - (int)(blockIdx.x / n_heads);          // longest rows first
    const int h = blockIdx.x % n_heads, grp = h / gsz;
    const int E = n_heads * 64, KV = (n_heads / gsz) * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int q0 = qt * 64;
    const uint32_t ks_base = smem_u32(Ks[0]), vs_base = smem_u32(Vs[0]);
    auto ks_addr = [&](int b) { return ks_base + (uint32_t)(b * 64 * PA_LD * 2); };
    auto vs_addr = [&](int b) { return vs_base + (uint32_t)(b * 64 * PA_LD * 2); };

    pa_load_tile_async(smem_u32(Qs), q16, q0, T, E, h * 64);
    pa_load_tile_async(ks_addr(0), k16, 0, T, KV, grp * 64);
    if (!TWO_PASS) pa_load_tile_async(vs_addr(0), v16, 0, T, KV, grp * 64);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    uint32_t qa[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; kk++)
        ldsm_x4(smem_u32(Qs) + (uint32_t)(((16 * warp + (lane & 15)) * PA_LD + 16 * kk + (lane >> 4) * 8) * 2), qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3]);
    const int qrow0 = q0 + 16 * warp;

    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    if (TWO_PASS) {
        // ---- pass A: row maximum and sum of exp (ops.h:972-988); K tiles double-buffered with cp.async
        for (int kt = 0; kt <= qt; kt++) {
            const int cur = kt & 1;
            if (kt > 0) { cp_async_wait_all(); __syncthreads(); }
            const int nk = (kt < qt) ? kt + 1 : 0;                    // the last prefetch brings tile 0 of pass B (K again, plus V)
            pa_load_tile_async(ks_addr(cur ^ 1), k16, nk * 64, T, KV, grp * 64);
            if (kt == qt) pa_load_tile_async(vs_addr(cur ^ 1), v16, 0, T, KV, grp * 64);
            cp_async_commit();
            float S[8][4];
            pa_scores(S, qa, ks_addr(cur), lane, kt == qt, qrow0, kt * 64);
            float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; j++) { t0 = fmaxf(t0, fmaxf(S[j][0], S[j][1])); t1 = fmaxf(t1, fmaxf(S[j][2], S[j][3])); }
            t0 = quad_max(t0); t1 = quad_max(t1);
            const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);
            l0 *= ex2_approx((m0 - n0) * PA_C); l1 *= ex2_approx((m1 - n1) * PA_C);
            const float nc0 = -n0 * PA_C, nc1 = -n1 * PA_C;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                l0 += ex2_approx(fmaf(S[j][0], PA_C, nc0)) + ex2_approx(fmaf(S[j][1], PA_C, nc0));
                l1 += ex2_approx(fmaf(S[j][2], PA_C, nc1)) + ex2_approx(fmaf(S[j][3], PA_C, nc1));
            }
            m0 = n0; m1 = n1;
        }
        l0 = quad_sum(l0); l1 = quad_sum(l1);
    }
    const float pinv0 = TWO_PASS ? __fdiv_rn(1.0f, l0) : 1.0f, pinv1 = TWO_PASS ? __fdiv_rn(1.0f, l1) : 1.0f;

    // ---- main sweep: probabilities, Q8 re-encode per 32 keys, P.V
    float O[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) O[j][0] = O[j][1] = O[j][2] = O[j][3] = 0.0f;
    const int buf0 = TWO_PASS ? (qt + 1) & 1 : 0;                     // buffer holding tile 0 of this sweep
    for (int kt = 0; kt <= qt; kt++) {
        const int cur = (buf0 + kt) & 1;
        if (TWO_PASS || kt > 0) { cp_async_wait_all(); __syncthreads(); }
        if (kt < qt) {
            pa_load_tile_async(ks_addr(cur ^ 1), k16, (kt + 1) * 64, T, KV, grp * 64);
            pa_load_tile_async(vs_addr(cur ^ 1), v16, (kt + 1) * 64, T, KV, grp * 64);
            cp_async_commit();
        }
        float S[8][4];
        pa_scores(S, qa, ks_addr(cur), lane, kt == qt, qrow0, kt * 64);
        if (!TWO_PASS) {
            float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; j++) { t0 = fmaxf(t0, fmaxf(S[j][0], S[j][1])); t1 = fmaxf(t1, fmaxf(S[j][2], S[j][3])); }
            t0 = quad_max(t0); t1 = quad_max(t1);
            const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);
            if (__any_sync(0xffffffffu, n0 != m0 || n1 != m1)) {      // a new maximum: rescale what has been accumulated
                const float c0 = ex2_approx((m0 - n0) * PA_C), c1 = ex2_approx((m1 - n1) * PA_C);
                l0 *= c0; l1 *= c1;
#pragma unroll
                for (int j = 0; j < 8; j++) { O[j][0] *= c0; O[j][1] *= c0; O[j][2] *= c1; O[j][3] *= c1; }
                m0 = n0; m1 = n1;
            }
        }
        const float nc0 = -m0 * PA_C, nc1 = -m1 * PA_C;
#pragma unroll
        for (int j = 0; j < 8; j++) {                                 // e = exp(s - max), ops.h:982-988
            S[j][0] = ex2_approx(fmaf(S[j][0], PA_C, nc0)); S[j][1] = ex2_approx(fmaf(S[j][1], PA_C, nc0));
            S[j][2] = ex2_approx(fmaf(S[j][2], PA_C, nc1)); S[j][3] = ex2_approx(fmaf(S[j][3], PA_C, nc1));
            if (!TWO_PASS) { l0 += S[j][0] + S[j][1]; l1 += S[j][2] + S[j][3]; }
        }
        if (AT == DT_F16) {
            // FP16 activations: every probability is rounded to fp16 on its own (ops.h:996 with an fp16 row): p itself is the
            // A operand.  The single sweep rounds e = exp(s - running max) instead of e / sum.
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                uint32_t pa[4];
                if (TWO_PASS) {
                    pa[0] = pack_h2(__fdiv_rn(S[2 * kk][0], l0), __fdiv_rn(S[2 * kk][1], l0));
                    pa[1] = pack_h2(__fdiv_rn(S[2 * kk][2], l1), __fdiv_rn(S[2 * kk][3], l1));
                    pa[2] = pack_h2(__fdiv_rn(S[2 * kk + 1][0], l0), __fdiv_rn(S[2 * kk + 1][1], l0));
                    pa[3] = pack_h2(__fdiv_rn(S[2 * kk + 1][2], l1), __fdiv_rn(S[2 * kk + 1][3], l1));
                } else {
                    pa[0] = pack_h2(S[2 * kk][0], S[2 * kk][1]);
                    pa[1] = pack_h2(S[2 * kk][2], S[2 * kk][3]);
                    pa[2] = pack_h2(S[2 * kk + 1][0], S[2 * kk + 1][1]);
                    pa[3] = pack_h2(S[2 * kk + 1][2], S[2 * kk + 1][3]);
                }
#pragma unroll
                for (int jp = 0; jp < 4; jp++) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4_t(vs_addr(cur) + (uint32_t)(((16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8) * PA_LD + 16 * jp + (lane >> 4) * 8) * 2), b0, b1, b2, b3);
                    mma_16816(O[2 * jp], pa, b0, b1);
                    mma_16816(O[2 * jp + 1], pa, b2, b3);
                }
            }
        } else
#pragma unroll
        for (int bb = 0; bb < 2; bb++) {                              // one Q8 block = 32 keys = 4 octets (ops.h:996)
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int j = 4 * bb; j < 4 * bb + 4; j++) { a0 = fmaxf(a0, fmaxf(S[j][0], S[j][1])); a1 = fmaxf(a1, fmaxf(S[j][2], S[j][3])); }
            a0 = quad_max(a0); a1 = quad_max(a1);
            float f0, f1, dq0, dq1;
            if (TWO_PASS) {
                a0 *= pinv0; a1 *= pinv1;                             // largest probability of the block (p = e / sum, ops.h:991-994)
                const float de0 = __fdiv_rn(a0, 127.0f), de1 = __fdiv_rn(a1, 127.0f);
                f0 = (de0 != 0.0f) ? __fdiv_rn(1.0f, de0) * pinv0 : 0.0f;     // code = round(p / delta), p = e * pinv
                f1 = (de1 != 0.0f) ? __fdiv_rn(1.0f, de1) * pinv1 : 0.0f;
                dq0 = h2f(f2h(de0)); dq1 = h2f(f2h(de1));
            } else {
                f0 = (a0 != 0.0f) ? __fdiv_rn(127.0f, a0) : 0.0f; f1 = (a1 != 0.0f) ? __fdiv_rn(127.0f, a1) : 0.0f;
                dq0 = a0 * (1.0f / 127.0f); dq1 = a1 * (1.0f / 127.0f);
            }
            const float M = 12582912.0f;                              // (x + 1.5 * 2^23) - 1.5 * 2^23 = nearest integer
            if (TWO_PASS) {
                // the integer codes (exact in fp16) go through the tensor cores; the block scale multiplies the partial sum
                float Ob[8][4];
#pragma unroll
                for (int j = 0; j < 8; j++) Ob[j][0] = Ob[j][1] = Ob[j][2] = Ob[j][3] = 0.0f;
#pragma unroll
                for (int kk = 2 * bb; kk < 2 * bb + 2; kk++) {        // 16 keys per MMA k-step
                    uint32_t pa[4];
                    pa[0] = pack_h2(fmaf(S[2 * kk][0], f0, M) - M, fmaf(S[2 * kk][1], f0, M) - M);
                    pa[1] = pack_h2(fmaf(S[2 * kk][2], f1, M) - M, fmaf(S[2 * kk][3], f1, M) - M);
                    pa[2] = pack_h2(fmaf(S[2 * kk + 1][0], f0, M) - M, fmaf(S[2 * kk + 1][1], f0, M) - M);
                    pa[3] = pack_h2(fmaf(S[2 * kk + 1][2], f1, M) - M, fmaf(S[2 * kk + 1][3], f1, M) - M);
#pragma unroll
                    for (int jp = 0; jp < 4; jp++) {
                        uint32_t b0, b1, b2, b3;
                        ldsm_x4_t(vs_addr(cur) + (uint32_t)(((16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8) * PA_LD + 16 * jp + (lane >> 4) * 8) * 2), b0, b1, b2, b3);
                        mma_16816(Ob[2 * jp], pa, b0, b1);
                        mma_16816(Ob[2 * jp + 1], pa, b2, b3);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    O[j][0] = fmaf(dq0, Ob[j][0], O[j][0]); O[j][1] = fmaf(dq0, Ob[j][1], O[j][1]);
                    O[j][2] = fmaf(dq1, Ob[j][2], O[j][2]); O[j][3] = fmaf(dq1, Ob[j][3], O[j][3]);
                }
            } else {
                // decoded probabilities code * scale as the fp16 A operand (7-bit code x scale rounded to 11 bits, like every
                // other fp16 operand of this path): one accumulator, no per-block partial sums
#pragma unroll
                for (int kk = 2 * bb; kk < 2 * bb + 2; kk++) {
                    uint32_t pa[4];
                    pa[0] = pack_h2((fmaf(S[2 * kk][0], f0, M) - M) * dq0, (fmaf(S[2 * kk][1], f0, M) - M) * dq0);
                    pa[1] = pack_h2((fmaf(S[2 * kk][2], f1, M) - M) * dq1, (fmaf(S[2 * kk][3], f1, M) - M) * dq1);
                    pa[2] = pack_h2((fmaf(S[2 * kk + 1][0], f0, M) - M) * dq0, (fmaf(S[2 * kk + 1][1], f0, M) - M) * dq0);
                    pa[3] = pack_h2((fmaf(S[2 * kk + 1][2], f1, M) - M) * dq1, (fmaf(S[2 * kk + 1][3], f1, M) - M) * dq1);
#pragma unroll
                    for (int jp = 0; jp < 4; jp++) {
                        uint32_t b0, b1, b2, b3;
                        ldsm_x4_t(vs_addr(cur) + (uint32_t)(((16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8) * PA_LD + 16 * jp + (lane >> 4) * 8) * 2), b0, b1, b2, b3);
                        mma_16816(O[2 * jp], pa, b0, b1);
                        mma_16816(O[2 * jp + 1], pa, b2, b3);
                    }
                }
            }
        }
    }
    if (!TWO_PASS) {
        const float i0 = __fdiv_rn(1.0f, quad_sum(l0)), i1 = __fdiv_rn(1.0f, quad_sum(l1));
#pragma unroll
        for (int j = 0; j < 8; j++) { O[j][0] *= i0; O[j][1] *= i0; O[j][2] *= i1; O[j][3] *= i1; }
    }

    // ---- output row: re-encode (ops.h:1084), fp16 operand of the o-projection
    const int r0 = qrow0 + g, r1 = r0 + 8;
    if (AT == DT_F16) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int col = h * 64 + 8 * j + 2 * t;
            if (r0 < T) {
                *reinterpret_cast<uint32_t*>(out16 + (size_t)r0 * E + col) = pack_h2(O[j][0], O[j][1]);
                if (cap) { cap[(size_t)r0 * capw + col] = f16_roundtrip(O[j][0]); cap[(size_t)r0 * capw + col + 1] = f16_roundtrip(O[j][1]); }
            }
            if (r1 < T) {
                *reinterpret_cast<uint32_t*>(out16 + (size_t)r1 * E + col) = pack_h2(O[j][2], O[j][3]);
                if (cap) { cap[(size_t)r1 * capw + col] = f16_roundtrip(O[j][2]); cap[(size_t)r1 * capw + col + 1] = f16_roundtrip(O[j][3]); }
            }
        }
        return;
    }
#pragma unroll
    for (int cb = 0; cb < 2; cb++) {
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int j = 4 * cb; j < 4 * cb + 4; j++) { a0 = fmaxf(a0, fmaxf(fabsf(O[j][0]), fabsf(O[j][1]))); a1 = fmaxf(a1, fmaxf(fabsf(O[j][2]), fabsf(O[j][3]))); }
        a0 = quad_max(a0); a1 = quad_max(a1);
        const float de0 = __fdiv_rn(a0, 127.0f), de1 = __fdiv_rn(a1, 127.0f);
        const float sc0 = (de0 != 0.0f) ? __fdiv_rn(1.0f, de0) : 0.0f, sc1 = (de1 != 0.0f) ? __fdiv_rn(1.0f, de1) : 0.0f;
        const float dq0 = h2f(f2h(de0)), dq1 = h2f(f2h(de1));
#pragma unroll
        for (int j = 4 * cb; j < 4 * cb + 4; j++) {
            const int col = h * 64 + 8 * j + 2 * t;
            const float y0 = __fmul_rn(roundf(__fmul_rn(O[j][0], sc0)), dq0), y1 = __fmul_rn(roundf(__fmul_rn(O[j][1], sc0)), dq0);
            const float y2 = __fmul_rn(roundf(__fmul_rn(O[j][2], sc1)), dq1), y3 = __fmul_rn(roundf(__fmul_rn(O[j][3], sc1)), dq1);
            if (r0 < T) {
                *reinterpret_cast<uint32_t*>(out16 + (size_t)r0 * E + col) = pack_h2(y0, y1);
                if (cap) { cap[(size_t)r0 * capw + col] = y0; cap[(size_t)r0 * capw + col + 1] = y1; }
            }
            if (r1 < T) {
                *reinterpret_cast<uint32_t*>(out16 + (size_t)r1 * E + col) = pack_h2(y2, y3);
                if (cap) { cap[(size_t)r1 * capw + col] = y2; cap[(size_t)r1 * capw + col + 1] = y3; }
            }
        }
    }
}

// =================================================================================================== host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// fp16 matrix [rows][cols] (cols contiguous) -> tensor map with a [box_rows][64] box, 128-byte swizzle, zero fill out of bounds
static int make_tmap(CUtensorMap* m, const void* base, int rows, int cols, int box_rows) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return fail(GTB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)PF_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GTB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d matrix", (int)r, rows, cols);
    return GTB_OK;
}

static bool g_pdl = true;

// launch on the library stream; with `g_pdl` the launch carries the programmatic-stream-serialization attribute (the kernel
// itself orders its global accesses after the previous kernel with griddepcontrol.wait)
template <typename... KArgs, typename... Args>
static cudaError_t pf_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <int BN, int EPI, int AT>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const PfEpi& ep) {
    using Cfg = PfGemmCfg<BN>;
    static bool attr = false;
    if (!attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_pf_gemm<BN, EPI, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr = true;
    }
    const int n_tiles = ((M + PF_BM - 1) / PF_BM) * ((N + BN - 1) / BN);
    const int grid = n_tiles < ctx().sm_count ? n_tiles : ctx().sm_count;
    GTB_CUDA(pf_launch(k_pf_gemm<BN, EPI, AT>, dim3(grid), dim3(PF_THREADS), Cfg::SMEM, ta, tb, M, N, K, ep));
    GTB_LAUNCHED();
    return GTB_OK;
}

template <int EPI>
static int gemm_epi(const CUtensorMap& ta, const CUtensorMap& tb, int bn, int M, int N, int K, const PfEpi& ep) {
    if (ep.at == DT_F16) return bn == 256 ? launch_gemm<256, EPI, DT_F16>(ta, tb, M, N, K, ep) : launch_gemm<128, EPI, DT_F16>(ta, tb, M, N, K, ep);
    return bn == 256 ? launch_gemm<256, EPI, DT_Q8>(ta, tb, M, N, K, ep) : launch_gemm<128, EPI, DT_Q8>(ta, tb, M, N, K, ep);
}

// CTA-pair variant: 256 x 256 tiles; tb128 = tensor map of W with a 128-row box (each CTA stages half of the W tile)
template <int EPI>
static int gemm_epi2(const CUtensorMap& ta, const CUtensorMap& tb128, int M, int N, int K, const PfEpi& ep) {
    static bool attr = false;
    if (!attr) {
        GTB_CUDA(cudaFuncSetAttribute(k_pf_gemm2<EPI, DT_Q8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF2_SMEM));
        GTB_CUDA(cudaFuncSetAttribute(k_pf_gemm2<EPI, DT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF2_SMEM));
        attr = true;
    }
    const int n_tiles = ((M + 255) / 256) * ((N + 255) / 256);
    const int max_clusters = ctx().sm_count / 2;
    const int clusters = n_tiles < max_clusters ? n_tiles : max_clusters;
    if (ep.at == DT_F16) GTB_CUDA(pf_launch(k_pf_gemm2<EPI, DT_F16>, dim3(2 * clusters), dim3(PF_THREADS), PF2_SMEM, ta, tb128, M, N, K, ep));
    else GTB_CUDA(pf_launch(k_pf_gemm2<EPI, DT_Q8>, dim3(2 * clusters), dim3(PF_THREADS), PF2_SMEM, ta, tb128, M, N, K, ep));
    GTB_LAUNCHED();
    return GTB_OK;
}

// N-tile width.  cost = waves x tile width; a 128-wide tile runs the tensor pipe at about half rate (per K=16 step it moves
// 16 KB through shared memory in 67 clocks, twice what an SM delivers; measured 41.7 us vs 20.3 us on the q|k|v shape).
static int pick_bn(int M, int N) {
    const int sms = ctx().sm_count > 0 ? ctx().sm_count : 148;
    const int n_mt = (M + PF_BM - 1) / PF_BM;
    const long t256 = (long)n_mt * ((N + 255) / 256), t128 = (long)n_mt * ((N + 127) / 128);
    const double c256 = (double)((t256 + sms - 1) / sms) * 256.0, c128 = (double)((t128 + sms - 1) / sms) * 128.0 * 1.9;
    return c128 < c256 ? 128 : 256;
}

static void report_cycles(long long* dbg, const char* what, int M, int N, int K) {
    std::vector<long long> h(1024 * 4);
    cudaMemcpyAsync(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost, ctx().stream);
    cudaStreamSynchronize(ctx().stream);
    double tot = 0, full = 0, acc = 0, n = 0; int c = 0;
    for (int i = 0; i < 1024; i++) if (h[i * 4 + 3] > 0) { tot += h[i * 4]; full += h[i * 4 + 1]; acc += h[i * 4 + 2]; n += h[i * 4 + 3]; c++; }
    if (c) fprintf(stderr, "[pf_gemm %s M=%d N=%d K=%d] MMA threads: %d, cycles/k-block %.0f (ideal 540), waiting for TMA data %.1f %%, for a free accumulator %.1f %%\n",
                   what, M, N, K, c, tot / n, 100.0 * full / tot, 100.0 * acc / tot);
}

int pf_gemm_f32(const void* d_A16, const void* d_W16, float* d_C, int M, int N, int K, int bn) {
    GTB_ARG(M > 0 && N > 0 && K > 0 && K % PF_BK == 0 && N % 32 == 0 && (bn == 128 || bn == 256 || bn == 512));
    CUtensorMap ta, tb;
    int r = make_tmap(&ta, d_A16, M, K, PF_BM);
    if (r) return r;
    r = make_tmap(&tb, d_W16, N, K, bn == 512 ? 128 : bn);
    if (r) return r;
    PfEpi ep;
    ep.out0 = d_C;
    if (getenv("GTB_PF_SAME_TILE")) ep.dbg_same_tile = 1;
    long long* dbg = nullptr;
    if (getenv("GTB_PF_CYCLES")) {
        if (cudaMalloc((void**)&dbg, 1024 * 4 * 8) == cudaSuccess) { cudaMemsetAsync(dbg, 0, 1024 * 4 * 8, ctx().stream); ep.dbg_cycles = dbg; }
    }
    r = (bn == 512) ? gemm_epi2<EPI_F32>(ta, tb, M, N, K, ep)         // bn = 512 selects the CTA-pair kernel (256 x 256 tiles)
                    : gemm_epi<EPI_F32>(ta, tb, bn, M, N, K, ep);
    if (dbg) {
        report_cycles(dbg, bn == 512 ? "probe, CTA pairs" : (bn == 256 ? "probe, 256-wide" : "probe, 128-wide"), M, N, K);
        cudaFree(dbg);
    }
    return r;
}

struct PfLayerW {
    __half* w[4] = {nullptr, nullptr, nullptr, nullptr};      // q|k|v, o, gate|up (rows interleaved by 32), down as fp16 [N][K]
    CUtensorMap tm[4][2];                                     // [..][0]: 128-row box, [..][1]: 256-row box
    bool set[4] = {false, false, false, false};
};

struct PfPlan {
    gtb_model_config cfg{};
    int E = 0, F = 0, KV = 0, NQKV = 0, Tcap = 0;
    bool fused = true;               // RoPE/KV-append and SiLU*up inside the GEMM epilogues
    bool two_cta = false;            // CTA-pair GEMM (k_pf_gemm2) wherever the tile is 256 wide
    bool attn_two_pass = false;      // true: a second QK^T sweep so that the fp16 rounding of the P-row block scales is reproduced too
    std::vector<PfLayerW> L;
    // activations for up to Tcap rows
    int8_t *xq = nullptr, *qkvq = nullptr, *oq = nullptr, *guq = nullptr;
    uint16_t *xs = nullptr, *qkvs = nullptr, *os = nullptr, *gus = nullptr;
    __half *xn16 = nullptr, *q16 = nullptr, *k16 = nullptr, *v16 = nullptr, *attn16 = nullptr, *act16 = nullptr;
    size_t bytes = 0;
    int64_t launches_last = 0;
};

static int pf_alloc(PfPlan* p, void** ptr, size_t n) {
    GTB_CUDA(cudaMalloc(ptr, n));
    p->bytes += n;
    ctx().mem += (int64_t)n;
    return GTB_OK;
}

int pf_create(PfPlan** out, const gtb_model_config& cfg) {
    GTB_ARG(cfg.wdtype == GTB_Q8 || cfg.wdtype == GTB_Q4 || cfg.wdtype == GTB_F16);
    auto* p = new PfPlan();
    p->cfg = cfg;
    p->E = cfg.n_embd; p->F = cfg.n_ffn; p->KV = 64 * cfg.n_groups; p->NQKV = p->E + 2 * p->KV;
    p->Tcap = cfg.max_ctx < PF_BM ? PF_BM : cfg.max_ctx;      // TMA boxes are 128 rows tall
    p->L.resize(cfg.n_layers);
    const size_t T = (size_t)p->Tcap;
    int r = 0;
    const size_t eb = (cfg.wdtype == GTB_F16) ? 2 : 1;         // bytes per element of the residual-stream / o / down matrices
    r |= pf_alloc(p, (void**)&p->xq, T * p->E * eb); r |= pf_alloc(p, (void**)&p->xs, T * (p->E / 32) * 2);
    r |= pf_alloc(p, (void**)&p->qkvq, T * p->NQKV); r |= pf_alloc(p, (void**)&p->qkvs, T * (p->NQKV / 32) * 2);
    r |= pf_alloc(p, (void**)&p->oq, T * p->E * eb); r |= pf_alloc(p, (void**)&p->os, T * (p->E / 32) * 2);
    r |= pf_alloc(p, (void**)&p->guq, T * 2 * p->F); r |= pf_alloc(p, (void**)&p->gus, T * (2 * p->F / 32) * 2);
    r |= pf_alloc(p, (void**)&p->xn16, T * p->E * 2); r |= pf_alloc(p, (void**)&p->q16, T * p->E * 2);
    r |= pf_alloc(p, (void**)&p->k16, T * p->KV * 2); r |= pf_alloc(p, (void**)&p->v16, T * p->KV * 2);
    r |= pf_alloc(p, (void**)&p->attn16, T * p->E * 2); r |= pf_alloc(p, (void**)&p->act16, T * p->F * 2);
    if (r) { pf_destroy(p); return r; }
    // rows beyond T are read by the last row tile (their results are never stored): keep them finite
    cudaMemsetAsync(p->xn16, 0, T * p->E * 2, ctx().stream);
    cudaMemsetAsync(p->attn16, 0, T * p->E * 2, ctx().stream);
    cudaMemsetAsync(p->act16, 0, T * p->F * 2, ctx().stream);
    *out = p;
    return GTB_OK;
}

void pf_destroy(PfPlan* p) {
    if (!p) return;
    for (auto& l : p->L) for (int w = 0; w < 4; w++) cudaFree(l.w[w]);
    void* bufs[] = {p->xq, p->xs, p->qkvq, p->qkvs, p->oq, p->os, p->guq, p->gus, p->xn16, p->q16, p->k16, p->v16, p->attn16, p->act16};
    for (void* b : bufs) cudaFree(b);
    ctx().mem -= (int64_t)p->bytes;
    delete p;
}

int pf_set_weight(PfPlan* p, int layer, int which, int wdtype, const void* d_data, const uint16_t* d_scales, int rows, int cols) {
    GTB_ARG(p && layer >= 0 && layer < (int)p->L.size() && which >= 0 && which < 4 && (wdtype == GTB_Q8 || wdtype == GTB_Q4 || wdtype == GTB_F16));
    const int N[4] = {p->NQKV, p->E, 2 * p->F, p->E}, K[4] = {p->E, p->E, p->E, p->F};
    GTB_ARG(rows == N[which] && cols == K[which]);
    PfLayerW& l = p->L[layer];
    const size_t n = (size_t)rows * cols;
    if (!l.w[which]) { int r = pf_alloc(p, (void**)&l.w[which], n * 2); if (r) return r; }
    const size_t nblk = n / 32;
    if (wdtype == GTB_F16) {
        const size_t ng8 = n / 8;
        k_pf_w16_f16<<<(unsigned)((ng8 + 255) / 256), 256, 0, ctx().stream>>>(reinterpret_cast<const __half*>(d_data), ng8, l.w[which], cols, which == 2 ? p->F : 0);
    } else {
        k_pf_w16<<<(unsigned)((nblk + 255) / 256), 256, 0, ctx().stream>>>(d_data, d_scales, wdtype, nblk, l.w[which], cols / 32, which == 2 ? p->F : 0);
    }
    GTB_LAUNCHED();
    int r = make_tmap(&l.tm[which][0], l.w[which], rows, cols, 128);
    if (!r) r = make_tmap(&l.tm[which][1], l.w[which], rows, cols, 256);
    if (r) return r;
    l.set[which] = true;
    return GTB_OK;
}

bool pf_weights_ready(const PfPlan* p) {
    for (auto& l : p->L) for (int w = 0; w < 4; w++) if (!l.set[w]) return false;
    return true;
}
size_t pf_bytes(const PfPlan* p) { return p->bytes; }
int64_t pf_launches_last(const PfPlan* p) { return p->launches_last; }

void pf_set_fused(PfPlan* p, bool on) { p->fused = on; }
void pf_set_pdl(bool on) { g_pdl = on; }
void pf_set_attn_two_pass(PfPlan* p, bool on) { p->attn_two_pass = on; }
void pf_set_two_cta(PfPlan* p, bool on) { p->two_cta = on; }

template <int EPI>
static int gemm_any(bool two_cta, const CUtensorMap& ta, const CUtensorMap (&tb)[2], int bn, int M, int N, int K, const PfEpi& ep) {
    if (two_cta && bn == 256) return gemm_epi2<EPI>(ta, tb[0], M, N, K, ep);
    return gemm_epi<EPI>(ta, tb[bn == 256], bn, M, N, K, ep);
}

int pf_run(PfPlan* p, const PfRun& r) {
    GTB_ARG(p && r.T > 0 && r.T <= p->cfg.max_ctx && r.n_layers_run > 0 && r.n_layers_run <= p->cfg.n_layers);
    if (!pf_weights_ready(p)) return fail(GTB_ERR_STATE, "fast prefill: fp16 weight copies are not built");
    cudaStream_t st = ctx().stream;
    const int T = r.T, E = p->E, F = p->F, NQKV = p->NQKV;
    const int nh = p->cfg.n_heads, ng = p->cfg.n_groups, nl = p->cfg.n_layers;
    const int at = (p->cfg.wdtype == GTB_F16) ? DT_F16 : DT_Q8;
    const bool fused = p->fused || at == DT_F16;          // FP16 activations only have the fused epilogues
    const int64_t l0 = ctx().launches;
    CUtensorMap ta_xn, ta_attn, ta_act;
    const int Tm = T < PF_BM ? PF_BM : T;
    int rc = make_tmap(&ta_xn, p->xn16, Tm, E, PF_BM);
    if (!rc) rc = make_tmap(&ta_attn, p->attn16, Tm, E, PF_BM);
    if (!rc) rc = make_tmap(&ta_act, p->act16, Tm, F, PF_BM);
    if (rc) return rc;
    auto capp = [&](int layer, int aid) -> float* {
        if (!r.cap) return nullptr;
        if (aid == GTB_A_EMB) return r.cap + (size_t)nl * 12 * T * r.capw;
        return r.cap + ((size_t)layer * 12 + (aid - GTB_A_ATTN_NORM)) * T * r.capw;
    };
    const int nbE = E / 32;
    {
        const size_t n = (size_t)T * nbE;
        GTB_CUDA(pf_launch(k_pf_embed, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, r.embed->data, r.embed->scales, r.embed->dtype, r.d_tokens, T, E, p->xq, p->xs,
                                                                 capp(0, GTB_A_EMB), r.capw));
        GTB_LAUNCHED();
    }
    if (at == DT_F16) GTB_CUDA(pf_launch(k_pf_add_norm<DT_F16>, dim3(T), dim3(64), 0, p->xq, p->xs, nullptr, nullptr, r.layers[0].attn_norm, p->xn16, T, E, nullptr, nullptr,
                                            capp(0, GTB_A_ATTN_NORM), r.capw));
        else GTB_CUDA(pf_launch(k_pf_add_norm<DT_Q8>, dim3(T), dim3(64), 0, p->xq, p->xs, nullptr, nullptr, r.layers[0].attn_norm, p->xn16, T, E, nullptr, nullptr,
                                            capp(0, GTB_A_ATTN_NORM), r.capw));
    GTB_LAUNCHED();
    const int bn_qkv = pick_bn(T, NQKV), bn_o = pick_bn(T, E), bn_gu = pick_bn(T, 2 * F), bn_d = pick_bn(T, E);
    long long* dbgc = nullptr;                      // GTB_PF_CYCLES=1: where the MMA thread of each GEMM of layer 0 spends its cycles
    if (getenv("GTB_PF_CYCLES") && cudaMalloc((void**)&dbgc, 4 * 1024 * 4 * 8) == cudaSuccess) cudaMemsetAsync(dbgc, 0, 4 * 1024 * 4 * 8, st);
    for (int li = 0; li < r.n_layers_run; li++) {
        const PfLayerIO& io = r.layers[li];
        PfLayerW& w = p->L[li];
        PfEpi eq;                                   // q|k|v: Linear re-encode, RoPE, K/V append
        eq.at = at;
        eq.out0 = p->qkvq; eq.out1 = p->qkvs;
        eq.rope_cos = r.rope_cos; eq.rope_sin = r.rope_sin; eq.q16 = p->q16; eq.k16 = p->k16; eq.v16 = p->v16;
        eq.kq = io.kq; eq.ks = io.ks; eq.vq = io.vq; eq.vs = io.vs; eq.nh = nh; eq.ng = ng;
        eq.cap0 = capp(li, GTB_A_Q); eq.cap1 = capp(li, GTB_A_K); eq.cap2 = capp(li, GTB_A_V); eq.capw = r.capw;
        if (dbgc && li == 0) eq.dbg_cycles = dbgc;
        if (fused) {
            rc = gemm_any<EPI_ROPE>(p->two_cta, ta_xn, w.tm[0], bn_qkv, T, NQKV, E, eq);
            if (rc) return rc;
        } else {
            rc = gemm_any<EPI_Q8>(p->two_cta, ta_xn, w.tm[0], bn_qkv, T, NQKV, E, eq);
            if (rc) return rc;
            const size_t n = (size_t)T * (nh + 2 * ng);
            GTB_CUDA(pf_launch(k_pf_rope_kv, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, p->qkvq, p->qkvs, T, eq));
            GTB_LAUNCHED();
        }
        {
            const dim3 ag((unsigned)(((T + 63) / 64) * nh));
            float* ac = capp(li, GTB_A_ATTN_OUT);
            if (at == DT_F16) {
                if (p->attn_two_pass) GTB_CUDA(pf_launch(k_pf_attn<true, DT_F16>, ag, dim3(128), 0, p->q16, p->k16, p->v16, p->attn16, T, nh, nh / ng, ac, r.capw));
                else GTB_CUDA(pf_launch(k_pf_attn<false, DT_F16>, ag, dim3(128), 0, p->q16, p->k16, p->v16, p->attn16, T, nh, nh / ng, ac, r.capw));
            } else {
                if (p->attn_two_pass) GTB_CUDA(pf_launch(k_pf_attn<true, DT_Q8>, ag, dim3(128), 0, p->q16, p->k16, p->v16, p->attn16, T, nh, nh / ng, ac, r.capw));
                else GTB_CUDA(pf_launch(k_pf_attn<false, DT_Q8>, ag, dim3(128), 0, p->q16, p->k16, p->v16, p->attn16, T, nh, nh / ng, ac, r.capw));
            }
        }
        GTB_LAUNCHED();
        PfEpi eo;
        eo.at = at;
        eo.out0 = p->oq; eo.out1 = p->os;
        rc = gemm_any<EPI_Q8>(p->two_cta, ta_attn, w.tm[1], bn_o, T, E, E, eo);
        if (rc) return rc;
        if (at == DT_F16) GTB_CUDA(pf_launch(k_pf_add_norm<DT_F16>, dim3(T), dim3(64), 0, p->xq, p->xs, p->oq, p->os, io.ffn_norm, p->xn16, T, E, capp(li, GTB_A_O), capp(li, GTB_A_INP_RES),
                                        capp(li, GTB_A_FFN_NORM), r.capw));
        else GTB_CUDA(pf_launch(k_pf_add_norm<DT_Q8>, dim3(T), dim3(64), 0, p->xq, p->xs, p->oq, p->os, io.ffn_norm, p->xn16, T, E, capp(li, GTB_A_O), capp(li, GTB_A_INP_RES),
                                        capp(li, GTB_A_FFN_NORM), r.capw));
        GTB_LAUNCHED();
        PfEpi eg;                                   // gate|up: Linear re-encode, SiLU, Multiply
        eg.at = at;
        eg.out0 = p->guq; eg.out1 = p->gus; eg.act16 = p->act16; eg.F = F;
        eg.cap0 = capp(li, GTB_A_GATE); eg.cap1 = capp(li, GTB_A_UP); eg.capw = r.capw;
        if (dbgc && li == 0) eg.dbg_cycles = dbgc + 2 * 1024 * 4;
        if (fused) {
            rc = gemm_any<EPI_SILU>(p->two_cta, ta_xn, w.tm[2], bn_gu, T, 2 * F, E, eg);
            if (rc) return rc;
        } else {
            rc = gemm_any<EPI_Q8>(p->two_cta, ta_xn, w.tm[2], bn_gu, T, 2 * F, E, eg);
            if (rc) return rc;
            const size_t n = (size_t)T * (F / 32);
            GTB_CUDA(pf_launch(k_pf_silu_mul, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, p->guq, p->gus, T, eg));
            GTB_LAUNCHED();
        }
        rc = gemm_any<EPI_Q8>(p->two_cta, ta_act, w.tm[3], bn_d, T, E, F, eo);
        if (rc) return rc;
        if (li + 1 < r.n_layers_run) {
            if (at == DT_F16) GTB_CUDA(pf_launch(k_pf_add_norm<DT_F16>, dim3(T), dim3(64), 0, p->xq, p->xs, p->oq, p->os, r.layers[li + 1].attn_norm, p->xn16, T, E, capp(li, GTB_A_DOWN),
                                            capp(li, GTB_A_ATTN_RES), capp(li + 1, GTB_A_ATTN_NORM), r.capw));
        else GTB_CUDA(pf_launch(k_pf_add_norm<DT_Q8>, dim3(T), dim3(64), 0, p->xq, p->xs, p->oq, p->os, r.layers[li + 1].attn_norm, p->xn16, T, E, capp(li, GTB_A_DOWN),
                                            capp(li, GTB_A_ATTN_RES), capp(li + 1, GTB_A_ATTN_NORM), r.capw));
            GTB_LAUNCHED();
        } else {
            GTB_CUDA(pf_launch(k_pf_tail, dim3((E + 255) / 256), dim3(256), 0, p->xq, p->xs, p->oq, p->os, T - 1, E, r.last_res, r.last_down, capp(li, GTB_A_DOWN), T, r.capw, at));
            GTB_LAUNCHED();
        }
    }
    if (dbgc) {
        report_cycles(dbgc, "q|k|v + RoPE", T, NQKV, E);
        report_cycles(dbgc + 2 * 1024 * 4, "gate|up + SiLU", T, 2 * F, E);
        cudaFree(dbgc);
    }
    p->launches_last = ctx().launches - l0;
    return GTB_OK;
}

}  // namespace gtb
