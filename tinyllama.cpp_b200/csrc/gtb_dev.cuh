// gtb_dev.cuh -- device-side numeric contract of the gten forward path.
//
// Everything here reproduces, bit for bit, what the reference's `-O3 -fopenmp -mavx -mf16c` build
// computes: separate (unfused) multiply/add roundings, round-half-away Q8 codes, RNE fp16 stores,
// glibc-2.39 expf, and strictly in-order fp32 sums.  The library is compiled with -fmad=false so
// nvcc never contracts a*b+c; the few places that WANT a fused op call fmaf()/fma() explicitly.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gtb {

// dtype codes = gten_types.h:20-26
enum : int { DT_I32 = 0, DT_F16 = 1, DT_F32 = 2, DT_Q8 = 3, DT_Q4 = 4 };

constexpr int QBLK = 32;       // quants.h:13-14
constexpr int Q8_BYTES = 34;   // quants.h:17-23
constexpr int Q4_BYTES = 18;   // quants.h:25-31

__host__ __device__ inline size_t row_nbytes(int dt, int n) {
    switch (dt) {
        case DT_Q8: return (size_t)((n + QBLK - 1) / QBLK) * Q8_BYTES;
        case DT_Q4: return (size_t)(n / QBLK) * Q4_BYTES;
        case DT_F16: return (size_t)n * 2;
        default: return (size_t)n * 4;
    }
}

// ---------------------------------------------------------------- fp16 <-> fp32 (gten_types.h:79-149)
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t f2h(float f) {
    // RNE, overflow -> inf; the reference maps every NaN to sign|0x7E00
    if (f != f) return (uint16_t)(((__float_as_uint(f) >> 16) & 0x8000u) | 0x7E00u);
    return __half_as_ushort(__float2half_rn(f));
}

// one element of a row stored in the reference layout (gten/ops.h:40-70)
__device__ __forceinline__ float read_elem(const uint8_t* row, int dtype, int e) {
    switch (dtype) {
        case DT_Q8: {
            const uint8_t* blk = row + (size_t)(e >> 5) * Q8_BYTES;
            const float delta = h2f((uint16_t)blk[0] | ((uint16_t)blk[1] << 8));
            return __fmul_rn((float)(int8_t)blk[2 + (e & 31)], delta);
        }
        case DT_Q4: {
            const uint8_t* blk = row + (size_t)(e >> 5) * Q4_BYTES;
            const float delta = h2f((uint16_t)blk[0] | ((uint16_t)blk[1] << 8));
            const int j = e & 31;
            const uint8_t byte = blk[2 + (j & 15)];
            const int q = (int)((j < 16) ? (byte >> 4) : (byte & 0x0f)) - 7;
            return __fmul_rn((float)q, delta);
        }
        case DT_F16: return h2f(reinterpret_cast<const uint16_t*>(row)[e]);
        default: return reinterpret_cast<const float*>(row)[e];
    }
}


// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Q8 encode of one 32-element block held one element per lane (quants.h:52-66).
// Lanes >= n_valid must pass x = 0 (they do not exist in a partial block).  Returns the int8 code;
// *delta_h receives the fp16 bits of delta (same on every lane).
__device__ __forceinline__ int q8_encode_lane(float x, uint16_t* delta_h) {
    const float absmax = warp_max(fabsf(x));
    const float delta = __fdiv_rn(absmax, 127.0f);
    *delta_h = f2h(delta);
    const float scale = (delta != 0.0f) ? __fdiv_rn(1.0f, delta) : 0.0f;   // from the UNROUNDED delta
    return (int)roundf(__fmul_rn(x, scale));                                // half away from zero
}
// encode followed by decode: the value the next op reads back (quants.h:69-76)
__device__ __forceinline__ float q8_roundtrip_lane(float x) {
    uint16_t dh;
    const int q = q8_encode_lane(x, &dh);
    return __fmul_rn((float)q, h2f(dh));
}
__device__ __forceinline__ float f16_roundtrip(float x) { return h2f(f2h(x)); }

// ---------------------------------------------------------------- glibc 2.39 expf (sysdeps/ieee754/flt-32/e_expf.c,
// FMA ifunc variant).  Verified over all 2^32 inputs against the host libm by tests/test_expf_gpu.py (gtb_selftest_expf):
// the range reduction is r = fma(InvLn2N, x, -kd); the polynomial uses fused steps.
// In global memory, read through the L1 (ld.global.nc): the index differs per lane, and a __constant__ table would serialise the
// up to 32 distinct addresses of a warp (measured in k_mega: 4 expf per thread cost 1.5 us with the constant-bank table).
static __device__ const uint64_t c_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

__device__ __forceinline__ float expf_glibc(float x) {
    const uint32_t ux = __float_as_uint(x);
    const uint32_t abstop = (ux >> 20) & 0x7ffu;
    if (abstop >= 0x42bu) {                                   // |x| >= 88 or NaN
        if (ux == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8u) return x + x;
        if (x > 88.72283172607421875f) return __uint_as_float(0x7f800000u);   // 0x1.62e42ep6
        if (x < -103.97207641601562500f) return 0.0f;                         // -0x1.9fe368p6
    }
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    const double xd = (double)x;
    const double z = __dmul_rn(InvLn2N, xd);
    double kd = __dadd_rn(z, SHIFT);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dadd_rn(kd, -SHIFT);
    const double r = fma(InvLn2N, xd, -kd);
    uint64_t t = __ldg(&c_exp2f_tab[ki & 31u]);
    t += ki << 47;
    const double s = __longlong_as_double((long long)t);
    const double p = fma(C0, r, C1);
    const double r2 = __dmul_rn(r, r);
    double y = fma(C2, r, 1.0);
    y = fma(p, r2, y);
    y = __dmul_rn(y, s);
    return __double2float_rn(y);
}

// ---------------------------------------------------------------- exact in-order fp32 sum of non-negative terms
// s = (((0 + t0) + t1) + ...) with one rounding per add, computed by a whole CTA (SURVEY.md §7 hard part 3).
// While the running sum stays inside one binade, adding t is an integer step on the mantissa that depends
// on the state only through its parity ("add a if even, b if odd"); such maps compose associatively.
// A double-precision prefix sum bounds the true running sum tightly enough to know its binade for every
// element except the few adjacent to a power of two; those are applied as real float adds, in order.
struct PMap { uint32_t a, b; };
__device__ __forceinline__ PMap pmap_compose(PMap f, PMap g) {   // f first, then g
    PMap h;
    h.a = f.a + ((f.a & 1u) ? g.b : g.a);
    h.b = f.b + (((1u + f.b) & 1u) ? g.b : g.a);
    return h;
}
__device__ __forceinline__ int dexp(double d) { return (int)((__double2hiint(d) >> 20) & 0x7ff) - 1023; }

__device__ __forceinline__ PMap pmap_of(float tv, int e) {       // e = unbiased exponent of the running sum
    PMap m{0u, 0u};
    const uint32_t tb = __float_as_uint(tv);
    const int sh = e - ((int)(tb >> 23) - 127);
    if (sh >= 25) return m;
    const uint32_t mt = (tb & 0x7fffffu) | 0x800000u;
    if (sh == 0) { m.a = m.b = mt; return m; }
    const uint32_t k = (sh >= 24) ? 0u : (mt >> sh);
    const uint32_t rem = mt & ((1u << sh) - 1u);
    const uint32_t half = 1u << (sh - 1);
    if (rem < half) { m.a = m.b = k; }
    else if (rem > half) { m.a = m.b = k + 1u; }
    else { m.a = k + (k & 1u); m.b = k + ((k + 1u) & 1u); }
    return m;
}

constexpr int ES_EPT = 8;          // elements per thread per pass
constexpr int ES_MAXEXP = 96;      // explicit elements per pass before the serial fallback

struct ExactSumSmem {
    double wsum[32];
    PMap wtail[32];
    int wflag[32];
    int wcnt[32];
    uint32_t item[2 * ES_MAXEXP + 2][2];
    uint8_t itype[2 * ES_MAXEXP + 2];
    float result;
    double carry_A;
    int total;
};

// terms: shared or global pointer to n floats (all >= 0).  Every thread of the CTA must call; blockDim.x
// must be a multiple of 32 (<= 1024).  Returns the sum on every thread.
template <typename LoadT>
__device__ float exact_sum_block(LoadT load, int n, ExactSumSmem& sm) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const int per_pass = blockDim.x * ES_EPT;
    float s_run = 0.0f;
    if (tid == 0) sm.carry_A = 0.0;
    __syncthreads();
    for (int p0 = 0; p0 < n; p0 += per_pass) {
        const int i0 = p0 + tid * ES_EPT;
        float tv[ES_EPT];
        double loc = 0.0;
#pragma unroll
        for (int j = 0; j < ES_EPT; j++) { tv[j] = (i0 + j < n) ? load(i0 + j) : 0.0f; loc += (double)tv[j]; }
        // exclusive prefix of the per-thread sums (double)
        double inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) sm.wsum[wid] = inc;
        __syncthreads();
        double base = sm.carry_A;
        for (int w = 0; w < wid; w++) base += sm.wsum[w];
        double A = base + (inc - loc);
        // classify
        PMap em[ES_EPT];
        uint32_t exmask = 0;
#pragma unroll
        for (int j = 0; j < ES_EPT; j++) {
            const double An = A + (double)tv[j];
            em[j] = PMap{0u, 0u};
            if (tv[j] != 0.0f) {
                const double rel = (double)(i0 + j + 8) * 0x1p-22;     // >= 2x the worst-case drift of the float chain
                const double lo = A - A * rel, hi = An + An * rel;
                bool ex = !(lo > 0x1p-100);
                if (!ex) {
                    const int eL = dexp(lo), eU = dexp(hi);
                    const uint32_t tb = __float_as_uint(tv[j]);
                    if (eL != eU || (tb >> 23) == 0u || ((int)(tb >> 23) - 127) > eL) ex = true;
                    else em[j] = pmap_of(tv[j], eL);
                }
                if (ex) exmask |= 1u << j;
            }
            A = An;
        }
        const int nexp = __popc(exmask);
        // exclusive count of explicit elements
        int cinc = nexp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, cinc, o); if (lane >= o) cinc += v; }
        if (lane == 31) sm.wcnt[wid] = cinc;
        __syncthreads();
        int cbase = cinc - nexp;
        for (int w = 0; w < wid; w++) cbase += sm.wcnt[w];
        if (tid == blockDim.x - 1) sm.total = cbase + nexp;
        __syncthreads();
        const int total = sm.total;
        if (total > ES_MAXEXP) {                 // pathological input: plain serial chain (still exact)
            if (tid == 0) {
                float s = s_run;
                const int hi_i = min(n, p0 + per_pass);
                for (int i = p0; i < hi_i; i++) s = __fadd_rn(s, load(i));
                double a = sm.carry_A;
                for (int w = 0; w < nw; w++) a += sm.wsum[w];
                sm.carry_A = a;
                sm.result = s;
            }
            __syncthreads();
            s_run = sm.result;
            __syncthreads();
            continue;
        }
        // per-thread pieces
        PMap cur{0u, 0u}, head{0u, 0u};
        int ne = 0;
#pragma unroll
        for (int j = 0; j < ES_EPT; j++) {
            if (exmask & (1u << j)) {
                const int idx = cbase + ne;
                if (ne == 0) head = cur;
                else { sm.item[2 * idx][0] = cur.a; sm.item[2 * idx][1] = cur.b; sm.itype[2 * idx] = 0; }
                sm.item[2 * idx + 1][0] = __float_as_uint(tv[j]);
                sm.itype[2 * idx + 1] = 1;
                cur = PMap{0u, 0u};
                ne++;
            } else {
                cur = pmap_compose(cur, em[j]);
            }
        }
        // inclusive segmented scan of the tails across the CTA
        PMap sc = cur;
        int fl = (ne > 0) ? 1 : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t pa = __shfl_up_sync(0xffffffffu, sc.a, o);
            const uint32_t pb = __shfl_up_sync(0xffffffffu, sc.b, o);
            const int pf = __shfl_up_sync(0xffffffffu, fl, o);
            if (lane >= o) { if (!fl) sc = pmap_compose(PMap{pa, pb}, sc); fl |= pf; }
        }
        if (lane == 31) { sm.wtail[wid] = sc; sm.wflag[wid] = fl; }
        __syncthreads();
        // prefix from earlier warps, then the exclusive value for this thread
        PMap wpre{0u, 0u};
        for (int w = 0; w < wid; w++) { if (sm.wflag[w]) wpre = sm.wtail[w]; else wpre = pmap_compose(wpre, sm.wtail[w]); }
        PMap incl = fl ? sc : pmap_compose(wpre, sc);
        // exclusive = inclusive value of the previous thread
        uint32_t ea = __shfl_up_sync(0xffffffffu, incl.a, 1), eb = __shfl_up_sync(0xffffffffu, incl.b, 1);
        PMap excl = (lane == 0) ? wpre : PMap{ea, eb};
        if (ne > 0) {
            const PMap r = pmap_compose(excl, head);
            sm.item[2 * cbase][0] = r.a; sm.item[2 * cbase][1] = r.b; sm.itype[2 * cbase] = 0;
        }
        if (tid == blockDim.x - 1) { sm.item[2 * total][0] = incl.a; sm.item[2 * total][1] = incl.b; sm.itype[2 * total] = 0; }
        __syncthreads();
        if (tid == 0) {
            uint32_t sb = __float_as_uint(s_run);
            for (int q = 0; q <= 2 * total; q++) {
                const uint32_t a = sm.item[q][0], b = sm.item[q][1];
                if (sm.itype[q]) sb = __float_as_uint(__fadd_rn(__uint_as_float(sb), __uint_as_float(a)));
                else sb += (sb & 1u) ? b : a;
            }
            sm.result = __uint_as_float(sb);
            double a = sm.carry_A;
            for (int w = 0; w < nw; w++) a += sm.wsum[w];
            sm.carry_A = a;
        }
        __syncthreads();
        s_run = sm.result;
        __syncthreads();
    }
    return s_run;
}

}  // namespace gtb
