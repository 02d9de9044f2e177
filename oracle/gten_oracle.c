/* oracle/gten_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the tinyllama.cpp transformer-forward hot path, bit-for-bit as the
 * reference's `g++ -std=c++17 -O3 -fopenmp -mavx -mf16c` build evaluates it:
 * no FMA contraction, no reassociation, 4-lane block dots, 8-lane float dots, in-order
 * RMSNorm/softmax sums, glibc expf/powf/cosf/sinf.  Compile with -ffp-contract=off (Makefile).
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks every function here bit-for-bit against
 * oracle/_ref/libgten_ref.so (the unmodified reference compiled from /root/reference), and
 * tests/golden/ holds vectors generated from that library by tests/golden/make_golden.py.
 * The reference itself ships no tests or golden vectors for this path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 *
 * Deliberate difference: attention is the *correct* causal attention.  The reference addresses its
 * score buffer with element strides as byte offsets (gten/ops.h:946-947,996,1057-1058), which makes
 * multi-row calls overlap rows when 2*n_ctx > max_ctx (FP16) or ceil(n_ctx/32)*34 > max_ctx (Q8);
 * inside that domain (and for every single-row decode call) the two agree bit-for-bit.
 */
#include "gten_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define QBLK 32          /* gten/quants.h:13-14 */
#define Q8_BYTES 34      /* gten/quants.h:17-23  {fp16 delta; int8 data[32]} */
#define Q4_BYTES 18      /* gten/quants.h:25-31  {fp16 delta; uint8 data[16]} */

const char* orc_build_info(void) { return "plain-C restatement of the -mavx -mf16c evaluation order; -ffp-contract=off"; }

/* ------------------------------------------------------------------ fp16 <-> fp32 ----
 * gten/gten_types.h:79-119.  The reference uses float-multiply tricks; this is the same
 * function written with integer bit logic: exact widening; round-to-nearest-even narrowing,
 * overflow -> inf, NaN -> sign|0x7E00. */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

float orc_fp16_to_fp32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1f;
    uint32_t m = h & 0x3ffu;
    if (e == 0x1f) return u2f(sign | 0x7f800000u | (m << 13));
    if (e != 0) return u2f(sign | ((e + 112u) << 23) | (m << 13));
    if (m == 0) return u2f(sign);
    /* subnormal half: normalise */
    uint32_t ee = 113;
    while (!(m & 0x400u)) { m <<= 1; ee--; }
    return u2f(sign | (ee << 23) | ((m & 0x3ffu) << 13));
}

uint16_t orc_fp32_to_fp16(float f) {
    const uint32_t x = f2u(f);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    const uint32_t ax = x & 0x7fffffffu;
    if (ax > 0x7f800000u) return sign | 0x7e00u;
    if (ax >= 0x47800000u) return sign | 0x7c00u;
    if (ax >= 0x38800000u) {
        const uint32_t mant = ax & 0x7fffffu;
        uint32_t h = (((ax >> 23) - 112u) << 10) | (mant >> 13);
        const uint32_t rem = mant & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
        return sign | (uint16_t)h;
    }
    if (ax < 0x33000000u) return sign;
    {
        const uint32_t e = ax >> 23;
        const uint32_t mant = (ax & 0x7fffffu) | 0x800000u;
        const uint32_t shift = 126u - e; /* 14..24 */
        uint32_t h = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1u);
        const uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) h++;
        return sign | (uint16_t)h;
    }
}

static inline uint16_t ld16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline void st16(uint8_t* p, uint16_t v) { memcpy(p, &v, 2); }

/* ------------------------------------------------------------------ block codecs ---- */

/* gten/quants.h:52-66 */
static void q8_quantize_block(const float* x, uint8_t* blk, int n) {
    float absmax = 0.0f;
    for (int j = 0; j < n; j++) {
        const float a = fabsf(x[j]);
        if (absmax < a) absmax = a;
    }
    const float delta = absmax / 127.0f;
    st16(blk, orc_fp32_to_fp16(delta));
    const float scale = (delta != 0.0f) ? 1.0f / delta : 0.0f;   /* the UNROUNDED fp32 delta */
    int8_t* q = (int8_t*)(blk + 2);
    for (int i = 0; i < n; i++) q[i] = (int8_t)roundf(x[i] * scale);  /* half away from zero */
}

/* gten/quants.h:92-110 (partial last block allowed) */
void orc_q8_quantize_row(const float* inp, void* out, int n) {
    uint8_t* o = (uint8_t*)out;
    const int nb = n / QBLK;
    for (int i = 0; i < nb; i++) q8_quantize_block(inp + i * QBLK, o + (size_t)i * Q8_BYTES, QBLK);
    if (n % QBLK) q8_quantize_block(inp + nb * QBLK, o + (size_t)nb * Q8_BYTES, n % QBLK);
}

/* gten/quants.h:69-76,118-133 */
void orc_q8_dequantize_row(const void* inp, float* out, int n) {
    const uint8_t* b = (const uint8_t*)inp;
    for (int i = 0; i < n; i += QBLK, b += Q8_BYTES) {
        const float delta = orc_fp16_to_fp32(ld16(b));
        const int8_t* q = (const int8_t*)(b + 2);
        const int m = (n - i < QBLK) ? n - i : QBLK;
        for (int j = 0; j < m; j++) out[i + j] = (float)q[j] * delta;
    }
}

/* gten/quants.h:78-90,135-143: byte i = (elt i + 7) << 4 | (elt i+16 + 7) */
void orc_q4_dequantize_row(const void* inp, float* out, int n) {
    const uint8_t* b = (const uint8_t*)inp;
    for (int i = 0; i < n; i += QBLK, b += Q4_BYTES) {
        const float delta = orc_fp16_to_fp32(ld16(b));
        for (int j = 0; j < 16; j++) {
            const int hi = (int)(b[2 + j] >> 4) - 7;
            const int lo = (int)(b[2 + j] & 0x0f) - 7;
            out[i + j] = (float)hi * delta;
            out[i + j + 16] = (float)lo * delta;
        }
    }
}

/* gten/ops.h:40-70 */
void orc_read_row_to_float(const void* inp, int dtype, float* out, int n) {
    switch (dtype) {
        case ORC_Q4: orc_q4_dequantize_row(inp, out, n); break;
        case ORC_Q8: orc_q8_dequantize_row(inp, out, n); break;
        case ORC_F16: { const uint8_t* p = (const uint8_t*)inp; for (int i = 0; i < n; i++) out[i] = orc_fp16_to_fp32(ld16(p + 2 * i)); } break;
        case ORC_F32: memcpy(out, inp, (size_t)n * 4); break;
        default: fprintf(stderr, "orc_read_row_to_float: bad dtype\n"); abort();
    }
}

/* gten/ops.h:73-96 (no Q4 encoder exists) */
void orc_write_row_from_float(float* inp, void* out, int dtype, int n) {
    switch (dtype) {
        case ORC_Q8: orc_q8_quantize_row(inp, out, n); break;
        case ORC_F16: { uint8_t* p = (uint8_t*)out; for (int i = 0; i < n; i++) st16(p + 2 * i, orc_fp32_to_fp16(inp[i])); } break;
        case ORC_F32: memcpy(out, inp, (size_t)n * 4); break;
        default: fprintf(stderr, "orc_write_row_from_float: bad dtype\n"); abort();
    }
}

static size_t row_nbytes(int dtype, int n) {
    switch (dtype) {
        case ORC_Q8: return (size_t)((n + QBLK - 1) / QBLK) * Q8_BYTES;
        case ORC_Q4: return (size_t)(n / QBLK) * Q4_BYTES;
        case ORC_F16: return (size_t)n * 2;
        default: return (size_t)n * 4;
    }
}

/* ------------------------------------------------------------------ dot products ---- */

/* gten/ops.h:140-160 (AVX branch) + gten/simd_ops.h:59-66: eight lane accumulators, lane l takes
 * elements 8i+l in ascending i; product (exact for fp16 inputs) and add are rounded separately;
 * lanes are summed left to right; then a scalar tail. */
static float dot_f16(const uint8_t* a, const uint8_t* b, int n) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int n8 = (n / 8) * 8;
    for (int i = 0; i < n8; i += 8)
        for (int l = 0; l < 8; l++) {
            const float p = orc_fp16_to_fp32(ld16(a + 2 * (i + l))) * orc_fp16_to_fp32(ld16(b + 2 * (i + l)));
            acc[l] = p + acc[l];
        }
    float d = acc[0] + acc[1];
    for (int l = 2; l < 8; l++) d = d + acc[l];
    for (int i = n8; i < n; i++) d += orc_fp16_to_fp32(ld16(a + 2 * i)) * orc_fp16_to_fp32(ld16(b + 2 * i));
    return d;
}

/* gten/ops.h:177-197: as above on fp32 data (product rounded before the add). */
static float dot_f32(const float* a, const float* b, int n) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int n8 = (n / 8) * 8;
    for (int i = 0; i < n8; i += 8)
        for (int l = 0; l < 8; l++) {
            const float p = a[i + l] * b[i + l];
            acc[l] = p + acc[l];
        }
    float d = acc[0] + acc[1];
    for (int l = 2; l < 8; l++) d = d + acc[l];
    for (int i = n8; i < n; i++) d += a[i] * b[i];
    return d;
}

/* Integer lane sums of one 32-element block (gten/ops.h:252-280, 339-378): `_mm_madd_epi16` on
 * 8-wide halves puts element pairs (2l,2l+1)+{0,8,16,24} into lane l. */
static inline void lane_sums(const int8_t* a, const int* w, int lane[4]) {
    for (int l = 0; l < 4; l++) {
        int s = 0;
        for (int m = 0; m < 32; m += 8) s += (int)a[m + 2 * l] * w[m + 2 * l] + (int)a[m + 2 * l + 1] * w[m + 2 * l + 1];
        lane[l] = s;
    }
}

/* gten/ops.h:224-292 (AVX branch): acc[l] += float(lane[l]) * (fp32(da)*fp32(db)), blocks ascending,
 * result (a0+a1)+(a2+a3) from the two hadd's. */
static float dot_q8_q8(const uint8_t* a, const uint8_t* b, int n) {
    float acc[4] = {0, 0, 0, 0};
    for (int i = 0; i < n / QBLK; i++, a += Q8_BYTES, b += Q8_BYTES) {
        int w[32], lane[4];
        const int8_t* bq = (const int8_t*)(b + 2);
        for (int j = 0; j < 32; j++) w[j] = bq[j];
        lane_sums((const int8_t*)(a + 2), w, lane);
        const float s = orc_fp16_to_fp32(ld16(a)) * orc_fp16_to_fp32(ld16(b));
        for (int l = 0; l < 4; l++) { const float p = (float)lane[l] * s; acc[l] = acc[l] + p; }
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

/* gten/ops.h:319-391 (AVX branch): Q4 elements 0-15 are the high nibbles, 16-31 the low nibbles, minus 7. */
static float dot_q8_q4(const uint8_t* a, const uint8_t* b, int n) {
    float acc[4] = {0, 0, 0, 0};
    for (int i = 0; i < n / QBLK; i++, a += Q8_BYTES, b += Q4_BYTES) {
        int w[32], lane[4];
        for (int j = 0; j < 16; j++) { w[j] = (int)(b[2 + j] >> 4) - 7; w[j + 16] = (int)(b[2 + j] & 0x0f) - 7; }
        lane_sums((const int8_t*)(a + 2), w, lane);
        const float s = orc_fp16_to_fp32(ld16(a)) * orc_fp16_to_fp32(ld16(b));
        for (int l = 0; l < 4; l++) { const float p = (float)lane[l] * s; acc[l] = acc[l] + p; }
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

/* gten/ops.h:482-512 */
float orc_vec_dot_product(const void* a, int adt, const void* b, int bdt, int n) {
    switch (adt) {
        case ORC_Q8: return (bdt == ORC_Q4) ? dot_q8_q4((const uint8_t*)a, (const uint8_t*)b, n)
                                            : dot_q8_q8((const uint8_t*)a, (const uint8_t*)b, n);
        case ORC_F16: return dot_f16((const uint8_t*)a, (const uint8_t*)b, n);
        case ORC_F32: return dot_f32((const float*)a, (const float*)b, n);
        default: fprintf(stderr, "orc_vec_dot_product: bad dtype\n"); abort();
    }
}

/* ------------------------------------------------------------------ row-level ops ---- */

/* gten/ops.h:514-533: F16/Q8 rows are copied; a Q4 row is dequantised and re-encoded as Q8. */
static void embed_row(const uint8_t* w, int wdt, int n_embd, int token, uint8_t* out, int odt, float* buf) {
    const uint8_t* src = w + (size_t)token * row_nbytes(wdt, n_embd);
    if (wdt == ORC_Q4) {
        orc_q4_dequantize_row(src, buf, n_embd);
        orc_write_row_from_float(buf, out, ORC_Q8, n_embd);
    } else {
        memcpy(out, src, row_nbytes(wdt, n_embd));
    }
    (void)odt;
}

/* gten/ops.h:632-646: one activation row against every weight row, then one row encode. */
static void matmul_row(const uint8_t* x, int xdt, int k, const uint8_t* w, int wdt, int n_out,
                       uint8_t* out, int odt, float* buf) {
    const size_t wst = row_nbytes(wdt, k);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < n_out; c++) buf[c] = orc_vec_dot_product(x, xdt, w + (size_t)c * wst, wdt, k);
    orc_write_row_from_float(buf, out, odt, n_out);
}

/* gten/ops.h:762-778: in-order sum of squares; eps is added to the rms; divide then multiply. */
static void rms_norm_row(const uint8_t* x, int xdt, int n, const uint16_t* w, uint8_t* out, float* buf) {
    float* xin = buf;
    float* y = buf + n;
    orc_read_row_to_float(x, xdt, xin, n);
    float sq_sum = 0.0f;
    for (int i = 0; i < n; i++) sq_sum += xin[i] * xin[i];
    const float sq_mean = sq_sum / (float)n;
    const float rms = sqrtf(sq_mean);
    for (int i = 0; i < n; i++) y[i] = xin[i] / (rms + 1e-6f) * orc_fp16_to_fp32(w[i]);
    orc_write_row_from_float(y, out, xdt, n);
}

/* gten/ops.h:728-746: rotate-half pairs (j, j+d/2); angle from float pos and powf. */
void orc_rope_angles(int pos, int d_head, float* cos_out, float* sin_out) {
    const float d = (float)d_head;
    const float m = (float)pos;
    for (int j = 0; j < d_head / 2; j++) {
        const float th = m * powf(10000.0f, -(2.0f * j / d));
        cos_out[j] = cosf(th);
        sin_out[j] = sinf(th);
    }
}

/* gten/ops.h:729-754: decode, rotate every head, re-encode in place. */
static void rope_row(uint8_t* x, int xdt, int n_embd, int d_head, int pos, float* buf) {
    float cs[256], sn[256];
    orc_read_row_to_float(x, xdt, buf, n_embd);
    orc_rope_angles(pos, d_head, cs, sn);
    const int dh = d_head / 2;
    for (int h = 0; h < n_embd / d_head; h++) {
        float* v = buf + h * d_head;
        for (int j = 0; j < dh; j++) {
            const float x0 = v[j], x1 = v[j + dh];
            const float a = x0 * cs[j], b = x1 * sn[j], c = x0 * sn[j], e = x1 * cs[j];
            v[j] = a - b;
            v[j + dh] = c + e;
        }
    }
    orc_write_row_from_float(buf, x, xdt, n_embd);
}

float orc_expf(float x) { return expf(x); }
/* host libm expf over the bit patterns [first, first + count): the yardstick of the device restatement (tests/test_expf_gpu.py) */
void orc_expf_bits_range(uint32_t first, uint32_t count, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)count; i++) {
        union { uint32_t u; float f; } v;
        v.u = first + (uint32_t)i;
        out[i] = expf(v.f);
    }
}

/* gten/ops.h:687-696 */
static void silu_row(const uint8_t* x, int xdt, int n, uint8_t* out, float* buf) {
    orc_read_row_to_float(x, xdt, buf, n);
    for (int j = 0; j < n; j++) { const float v = buf[j]; buf[j] = v / (1.0f + expf(-v)); }
    orc_write_row_from_float(buf, out, xdt, n);
}

/* gten/ops.h:841-849 and 890-897 */
static void binary_row(const uint8_t* a, const uint8_t* b, int xdt, int n, uint8_t* out, float* buf, int is_mul) {
    float* x0 = buf; float* x1 = buf + n; float* y = buf + 2 * n;
    orc_read_row_to_float(a, xdt, x0, n);
    orc_read_row_to_float(b, xdt, x1, n);
    if (is_mul) for (int i = 0; i < n; i++) y[i] = x0[i] * x1[i];
    else        for (int i = 0; i < n; i++) y[i] = x0[i] + x1[i];
    orc_write_row_from_float(y, out, xdt, n);
}

/* One query row of gten/ops.h:930-1000 (scores, mask, softmax, encode) followed by gten/ops.h:1046-1087
 * (decode P, P.V, encode).  `row` is the query position, `n_ctx` the length of the call (it sets the
 * P row length and the lane/tail split of the P.V dots); K/V rows 0..row must be valid.
 * q: one encoded row [n_heads*d_head]; kc/vc: encoded rows [kv_dim] with stride kv_stride bytes. */
static void attn_row(const uint8_t* q, const uint8_t* kc, const uint8_t* vc, size_t kv_stride, int xdt,
                     int row, int n_ctx, int n_heads, int n_kv, int d_head, uint8_t* out, float* scratch) {
    const int grp = n_heads / n_kv;
    const int n_embd = n_heads * d_head;
    const float scale = 1.0f / sqrtf((float)d_head);
    float* obuf = scratch;                          /* n_embd */
    float* vt = obuf + n_embd;                      /* n_kv*d_head*n_ctx, transposed V (ops.h:1003-1044) */
    float* prow_all = vt + (size_t)n_kv * d_head * n_ctx; /* n_heads * n_ctx */
    uint8_t* penc_all = (uint8_t*)(prow_all + (size_t)n_heads * n_ctx);
    const size_t penc_stride = row_nbytes(ORC_F32, n_ctx) + 64;
    const size_t head_bytes = row_nbytes(xdt, d_head);

    /* dequantise + transpose V for positions <= row; later positions only ever meet exact-zero probabilities */
    for (size_t i = 0; i < (size_t)n_kv * d_head * n_ctx; i++) vt[i] = 0.0f;
    {
        float tmp[4096];
        for (int i = 0; i <= row; i++) {
            orc_read_row_to_float(vc + (size_t)i * kv_stride, xdt, tmp, n_kv * d_head);
            for (int c = 0; c < n_kv * d_head; c++) vt[(size_t)c * n_ctx + i] = tmp[c];
        }
    }
#pragma omp parallel for schedule(static)
    for (int h = 0; h < n_heads; h++) {
        float* p = prow_all + (size_t)h * n_ctx;
        uint8_t* penc = penc_all + (size_t)h * penc_stride;
        const uint8_t* qh = q + (size_t)h * head_bytes;
        for (int kcol = 0; kcol <= row; kcol++) {
            const uint8_t* kh = kc + (size_t)kcol * kv_stride + (size_t)(h / grp) * head_bytes;
            p[kcol] = orc_vec_dot_product(qh, xdt, kh, xdt, d_head) * scale;
        }
        for (int kcol = row + 1; kcol < n_ctx; kcol++) p[kcol] = -INFINITY;
        float mx = -INFINITY;
        for (int i = 0; i < n_ctx; i++) if (p[i] > mx) mx = p[i];
        float sum = 0.0f;
        for (int i = 0; i < n_ctx; i++) { const float e = expf(p[i] - mx); p[i] = e; sum += e; }
        for (int i = 0; i < n_ctx; i++) p[i] = p[i] / sum;
        orc_write_row_from_float(p, penc, xdt, n_ctx);
        orc_read_row_to_float(penc, xdt, p, n_ctx);
        for (int c = 0; c < d_head; c++)
            obuf[h * d_head + c] = dot_f32(p, vt + ((size_t)(h / grp) * d_head + c) * n_ctx, n_ctx);
    }
    orc_write_row_from_float(obuf, out, xdt, n_embd);
}

static size_t attn_scratch_floats(int n_ctx, int n_heads, int n_kv, int d_head) {
    return (size_t)n_heads * d_head + (size_t)n_kv * d_head * n_ctx + (size_t)n_heads * n_ctx
         + ((size_t)n_heads * (row_nbytes(ORC_F32, n_ctx) + 64)) / 4 + 64;
}

/* ------------------------------------------------------------------ tensor-level ops ---- */

static float* scratch_f(size_t n) { float* p = (float*)malloc(n * sizeof(float)); if (!p) abort(); return p; }

void orc_token_embed(const void* w, int wdt, int n_vocab, int n_embd, const int32_t* tokens, int n_ctx,
                     void* out, int odt, int start_pos) {
    (void)n_vocab;
    float* buf = scratch_f(n_embd);
    for (int i = start_pos; i < n_ctx; i++)
        embed_row((const uint8_t*)w, wdt, n_embd, tokens[i], (uint8_t*)out + (size_t)i * row_nbytes(odt, n_embd), odt, buf);
    free(buf);
}

void orc_matmul_2d(const void* x, int xdt, int n_ctx, int k, const void* w, int wdt, int n_out,
                   void* out, int odt, int out_1d, int start_pos) {
    float* buf = scratch_f(n_out);
    for (int r = start_pos; r < n_ctx; r++) {
        uint8_t* o = (uint8_t*)out + (out_1d ? 0 : (size_t)r * row_nbytes(odt, n_out));
        matmul_row((const uint8_t*)x + (size_t)r * row_nbytes(xdt, k), xdt, k, (const uint8_t*)w, wdt, n_out, o, odt, buf);
    }
    free(buf);
}

void orc_rms_norm(const void* x, int xdt, int n_ctx, int n_embd, const uint16_t* w, void* out, int start_pos) {
    float* buf = scratch_f(2 * n_embd);
    const size_t st = row_nbytes(xdt, n_embd);
    for (int i = start_pos; i < n_ctx; i++) rms_norm_row((const uint8_t*)x + i * st, xdt, n_embd, w, (uint8_t*)out + i * st, buf);
    free(buf);
}

void orc_rotary_emb(void* x, int xdt, int n_ctx, int n_embd, int d_head, int start_pos) {
    float* buf = scratch_f(n_embd);
    const size_t st = row_nbytes(xdt, n_embd);
    for (int i = start_pos; i < n_ctx; i++) rope_row((uint8_t*)x + i * st, xdt, n_embd, d_head, i, buf);
    free(buf);
}

void orc_silu(const void* x, int xdt, int n_ctx, int n_embd, void* out, int start_pos) {
    float* buf = scratch_f(n_embd);
    const size_t st = row_nbytes(xdt, n_embd);
    for (int i = start_pos; i < n_ctx; i++) silu_row((const uint8_t*)x + i * st, xdt, n_embd, (uint8_t*)out + i * st, buf);
    free(buf);
}

void orc_mul(const void* a, const void* b, int xdt, int n_ctx, int n_embd, void* out, int start_pos) {
    float* buf = scratch_f(3 * n_embd);
    const size_t st = row_nbytes(xdt, n_embd);
    for (int i = start_pos; i < n_ctx; i++)
        binary_row((const uint8_t*)a + i * st, (const uint8_t*)b + i * st, xdt, n_embd, (uint8_t*)out + i * st, buf, 1);
    free(buf);
}

void orc_add(const void* a, const void* b, int xdt, int n_ctx, int n_embd, void* out, int start_pos) {
    float* buf = scratch_f(3 * n_embd);
    const size_t st = row_nbytes(xdt, n_embd);
    for (int i = start_pos; i < n_ctx; i++)
        binary_row((const uint8_t*)a + i * st, (const uint8_t*)b + i * st, xdt, n_embd, (uint8_t*)out + i * st, buf, 0);
    free(buf);
}

/* gten/ops.h:1095-1133.  `qk` and `max_ctx` exist for signature parity only (see header comment). */
void orc_qkv_attn(const void* q, const void* k, const void* v, void* qk, void* out, int xdt,
                  int n_ctx, int n_heads, int n_kv_heads, int d_head, int max_ctx, int start_pos) {
    (void)qk; (void)max_ctx;
    const int n_embd = n_heads * d_head, kv_dim = n_kv_heads * d_head;
    float* scratch = scratch_f(attn_scratch_floats(n_ctx, n_heads, n_kv_heads, d_head));
    const size_t qst = row_nbytes(xdt, n_embd), kst = row_nbytes(xdt, kv_dim);
    for (int r = start_pos; r < n_ctx; r++)
        attn_row((const uint8_t*)q + r * qst, (const uint8_t*)k, (const uint8_t*)v, kst, xdt, r, n_ctx,
                 n_heads, n_kv_heads, d_head, (uint8_t*)out + r * qst, scratch);
    free(scratch);
}

/* ------------------------------------------------------------------ model ----
 * The graph of TinyLlama::logits (tinyllama.cpp:45-61) over AttentionBlock::forward
 * (gten/modules.cpp:193-254), evaluated one sequence row at a time (every op of the reference is
 * row-independent except attention, which only reads K/V rows <= its own). */

typedef struct {
    uint8_t *q, *k, *v, *o, *gate, *up, *down;
    uint16_t *attn_norm, *ffn_norm;
    uint8_t *kcache, *vcache;
} OrcLayer;

#define N_ACV 12
typedef struct {
    int n_vocab, n_embd, n_ffn, n_layers, n_heads, n_groups, max_ctx, wdt, adt;
    int d_head, kv_dim;
    uint8_t *embed, *lm_head;
    uint16_t* final_norm;
    OrcLayer* L;
    /* capture of one row's activations (encoded bytes) */
    int capture_row, captured_row;
    uint8_t* cap;        /* [n_layers][N_ACV][cap_stride] */
    uint8_t* cap_misc;   /* [2][cap_stride]: emb, final_norm */
    size_t cap_stride;
    float* fbuf;         /* scratch */
    float* attn_scratch;
    uint8_t* rows;       /* encoded row buffers */
} OrcModel;

static int acv_slot(int id) { return (id >= ORC_A_ATTN_NORM && id <= ORC_A_ATTN_RES) ? id - ORC_A_ATTN_NORM : -1; }
static int acv_width(const OrcModel* m, int id) {
    switch (id) {
        case ORC_A_K: case ORC_A_V: return m->kv_dim;
        case ORC_A_GATE: case ORC_A_UP: return m->n_ffn;
        default: return m->n_embd;
    }
}

static void* xmalloc(size_t n) { void* p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "oracle: out of memory (%zu)\n", n); abort(); } return p; }

void* orc_model_new(int n_vocab, int n_embd, int n_ffn, int n_layers, int n_heads, int n_groups,
                    int max_ctx, int wdtype) {
    OrcModel* m = (OrcModel*)calloc(1, sizeof(OrcModel));
    m->n_vocab = n_vocab; m->n_embd = n_embd; m->n_ffn = n_ffn; m->n_layers = n_layers;
    m->n_heads = n_heads; m->n_groups = n_groups; m->max_ctx = max_ctx; m->wdt = wdtype;
    m->adt = (wdtype == ORC_F16) ? ORC_F16 : ORC_Q8;          /* tinyllama.cpp:258-265 */
    m->d_head = n_embd / n_heads; m->kv_dim = m->d_head * n_groups;
    m->embed = xmalloc((size_t)n_vocab * row_nbytes(wdtype, n_embd));
    m->lm_head = xmalloc((size_t)n_vocab * row_nbytes(wdtype, n_embd));
    m->final_norm = xmalloc((size_t)n_embd * 2);
    m->L = (OrcLayer*)calloc(n_layers, sizeof(OrcLayer));
    for (int i = 0; i < n_layers; i++) {
        OrcLayer* l = &m->L[i];
        l->q = xmalloc((size_t)n_embd * row_nbytes(wdtype, n_embd));
        l->k = xmalloc((size_t)m->kv_dim * row_nbytes(wdtype, n_embd));
        l->v = xmalloc((size_t)m->kv_dim * row_nbytes(wdtype, n_embd));
        l->o = xmalloc((size_t)n_embd * row_nbytes(wdtype, n_embd));
        l->gate = xmalloc((size_t)n_ffn * row_nbytes(wdtype, n_embd));
        l->up = xmalloc((size_t)n_ffn * row_nbytes(wdtype, n_embd));
        l->down = xmalloc((size_t)n_embd * row_nbytes(wdtype, n_ffn));
        l->attn_norm = xmalloc((size_t)n_embd * 2);
        l->ffn_norm = xmalloc((size_t)n_embd * 2);
        l->kcache = xmalloc((size_t)max_ctx * row_nbytes(m->adt, m->kv_dim));
        l->vcache = xmalloc((size_t)max_ctx * row_nbytes(m->adt, m->kv_dim));
    }
    m->capture_row = -1; m->captured_row = -1;
    m->cap_stride = row_nbytes(ORC_F32, n_ffn > n_embd ? n_ffn : n_embd);
    m->cap = xmalloc((size_t)n_layers * N_ACV * m->cap_stride);
    m->cap_misc = xmalloc(2 * m->cap_stride);
    {
        size_t big = (size_t)(n_ffn > n_vocab ? n_ffn : n_vocab);
        if ((size_t)n_embd > big) big = n_embd;
        m->fbuf = scratch_f(4 * big + 64);
    }
    m->attn_scratch = scratch_f(attn_scratch_floats(max_ctx, n_heads, n_groups, m->d_head));
    m->rows = xmalloc(16 * m->cap_stride);
    return m;
}

void orc_model_free(void* h) {
    OrcModel* m = (OrcModel*)h;
    if (!m) return;
    for (int i = 0; i < m->n_layers; i++) {
        OrcLayer* l = &m->L[i];
        free(l->q); free(l->k); free(l->v); free(l->o); free(l->gate); free(l->up); free(l->down);
        free(l->attn_norm); free(l->ffn_norm); free(l->kcache); free(l->vcache);
    }
    free(m->L); free(m->embed); free(m->lm_head); free(m->final_norm);
    free(m->cap); free(m->cap_misc); free(m->fbuf); free(m->attn_scratch); free(m->rows); free(m);
}

void* orc_model_weight(void* h, int layer, int id, int64_t* nbytes) {
    OrcModel* m = (OrcModel*)h;
    const size_t we = row_nbytes(m->wdt, m->n_embd), wf = row_nbytes(m->wdt, m->n_ffn);
    switch (id) {
        case ORC_T_EMBED: *nbytes = (int64_t)(m->n_vocab * we); return m->embed;
        case ORC_T_LM_HEAD: *nbytes = (int64_t)(m->n_vocab * we); return m->lm_head;
        case ORC_T_FINAL_NORM: *nbytes = m->n_embd * 2; return m->final_norm;
        default: break;
    }
    OrcLayer* l = &m->L[layer];
    switch (id) {
        case ORC_T_Q: *nbytes = (int64_t)(m->n_embd * we); return l->q;
        case ORC_T_K: *nbytes = (int64_t)(m->kv_dim * we); return l->k;
        case ORC_T_V: *nbytes = (int64_t)(m->kv_dim * we); return l->v;
        case ORC_T_O: *nbytes = (int64_t)(m->n_embd * we); return l->o;
        case ORC_T_GATE: *nbytes = (int64_t)(m->n_ffn * we); return l->gate;
        case ORC_T_UP: *nbytes = (int64_t)(m->n_ffn * we); return l->up;
        case ORC_T_DOWN: *nbytes = (int64_t)(m->n_embd * wf); return l->down;
        case ORC_T_ATTN_NORM: *nbytes = m->n_embd * 2; return l->attn_norm;
        case ORC_T_FFN_NORM: *nbytes = m->n_embd * 2; return l->ffn_norm;
    }
    *nbytes = 0;
    return NULL;
}

void orc_model_capture_row(void* h, int row) { ((OrcModel*)h)->capture_row = row; }

static void capture(OrcModel* m, int layer, int id, const uint8_t* enc) {
    const int w = acv_width(m, id);
    if (id == ORC_A_EMB) memcpy(m->cap_misc, enc, row_nbytes(m->adt, w));
    else if (id == ORC_A_FINAL_NORM) memcpy(m->cap_misc + m->cap_stride, enc, row_nbytes(m->adt, w));
    else memcpy(m->cap + ((size_t)layer * N_ACV + acv_slot(id)) * m->cap_stride, enc, row_nbytes(m->adt, w));
}

static void forward_row(OrcModel* m, int token, int row, int n_ctx, float* logits_out) {
    const int A = m->adt, W = m->wdt, E = m->n_embd, F = m->n_ffn, KV = m->kv_dim;
    const int cap = (m->capture_row < 0) ? (row == n_ctx - 1) : (row == m->capture_row);
    const size_t S = m->cap_stride;
    uint8_t *x = m->rows, *a = x + S, *q = a + S, *ao = q + S, *o = ao + S, *hres = o + S, *f = hres + S,
            *g = f + S, *u = g + S, *d = u + S, *xn = d + S;
    float* fb = m->fbuf;
    const size_t kvst = row_nbytes(A, KV);

    embed_row(m->embed, W, E, token, x, A, fb);                                  /* modules.cpp:17-26 */
    if (cap) { capture(m, 0, ORC_A_EMB, x); m->captured_row = row; }
    for (int li = 0; li < m->n_layers; li++) {
        OrcLayer* l = &m->L[li];
        uint8_t* krow = l->kcache + (size_t)row * kvst;
        uint8_t* vrow = l->vcache + (size_t)row * kvst;
        rms_norm_row(x, A, E, l->attn_norm, a, fb);                              /* modules.cpp:88-100 */
        matmul_row(a, A, E, l->q, W, E, q, A, fb);                               /* modules.cpp:195 */
        matmul_row(a, A, E, l->k, W, KV, krow, A, fb);                           /* modules.cpp:196: key.acv IS the K cache */
        rope_row(q, A, E, m->d_head, row, fb);                                   /* modules.cpp:198-199 */
        rope_row(krow, A, KV, m->d_head, row, fb);
        matmul_row(a, A, E, l->v, W, KV, vrow, A, fb);                           /* modules.cpp:201 */
        attn_row(q, l->kcache, l->vcache, kvst, A, row, n_ctx, m->n_heads, m->n_groups, m->d_head, ao, m->attn_scratch);
        matmul_row(ao, A, E, l->o, W, E, o, A, fb);                              /* modules.cpp:204 */
        binary_row(x, o, A, E, hres, fb, 0);                                     /* modules.cpp:251 inp_res */
        rms_norm_row(hres, A, E, l->ffn_norm, f, fb);
        matmul_row(f, A, E, l->gate, W, F, g, A, fb);                            /* modules.cpp:238-247 */
        matmul_row(f, A, E, l->up, W, F, u, A, fb);
        silu_row(g, A, F, g, fb);
        binary_row(g, u, A, F, g, fb, 1);
        matmul_row(g, A, F, l->down, W, E, d, A, fb);
        binary_row(hres, d, A, E, xn, fb, 0);                                    /* modules.cpp:252 attn_res */
        if (cap) {
            capture(m, li, ORC_A_ATTN_NORM, a); capture(m, li, ORC_A_Q, q); capture(m, li, ORC_A_K, krow);
            capture(m, li, ORC_A_V, vrow); capture(m, li, ORC_A_ATTN_OUT, ao); capture(m, li, ORC_A_O, o);
            capture(m, li, ORC_A_INP_RES, hres); capture(m, li, ORC_A_FFN_NORM, f); capture(m, li, ORC_A_GATE, g);
            capture(m, li, ORC_A_UP, u); capture(m, li, ORC_A_DOWN, d); capture(m, li, ORC_A_ATTN_RES, xn);
        }
        memcpy(x, xn, row_nbytes(A, E));
    }
    rms_norm_row(x, A, E, m->final_norm, a, fb);                                 /* tinyllama.cpp:57 */
    if (cap) capture(m, 0, ORC_A_FINAL_NORM, a);
    if (logits_out) matmul_row(a, A, E, m->lm_head, W, m->n_vocab, (uint8_t*)logits_out, ORC_F32, fb); /* modules.cpp:70-81 */
}

void orc_model_logits(void* h, const int32_t* tokens, int n_tokens, int start_pos, float* out) {
    OrcModel* m = (OrcModel*)h;
    if (n_tokens > m->max_ctx) { fprintf(stderr, "oracle: %d tokens exceed max_ctx %d\n", n_tokens, m->max_ctx); abort(); }
    for (int r = start_pos; r < n_tokens; r++)
        forward_row(m, tokens[r], r, n_tokens, (r == n_tokens - 1) ? out : NULL);
}

/* greedy_sample (tinyllama.cpp:395-440) without tokenizer / EOS stop; strict '>' argmax. */
void orc_model_generate(void* h, int32_t* tokens, int n_prompt, int n_new, double* times, float* logits_out) {
    OrcModel* m = (OrcModel*)h;
    float* lg = scratch_f(m->n_vocab);
    int n = n_prompt;
    double tp = 0, td = 0;
    for (int i = 0; i < n_new; i++) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        orc_model_logits(m, tokens, n, (i == 0) ? 0 : n - 1, lg);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        if (i == 0) tp += dt; else td += dt;
        if (logits_out) memcpy(logits_out + (size_t)i * m->n_vocab, lg, sizeof(float) * m->n_vocab);
        float best = -INFINITY; int arg = 0;
        for (int j = 0; j < m->n_vocab; j++) if (lg[j] > best) { best = lg[j]; arg = j; }
        tokens[n++] = arg;
    }
    if (times) { times[0] = tp; times[1] = td; }
    free(lg);
}

static const uint8_t* cap_ptr(OrcModel* m, int layer, int id) {
    if (id == ORC_A_EMB) return m->cap_misc;
    if (id == ORC_A_FINAL_NORM) return m->cap_misc + m->cap_stride;
    if (acv_slot(id) < 0 || layer < 0 || layer >= m->n_layers) return NULL;
    return m->cap + ((size_t)layer * N_ACV + acv_slot(id)) * m->cap_stride;
}

int orc_model_acv(void* h, int layer, int id, int row, float* out) {
    OrcModel* m = (OrcModel*)h;
    const uint8_t* p = cap_ptr(m, layer, id);
    if (!p) return -1;
    if (row != m->captured_row) return -2;
    const int w = acv_width(m, id);
    orc_read_row_to_float(p, m->adt, out, w);
    return w;
}

int orc_model_acv_raw(void* h, int layer, int id, int row, void* out) {
    OrcModel* m = (OrcModel*)h;
    const uint8_t* p = cap_ptr(m, layer, id);
    if (!p) return -1;
    if (row != m->captured_row) return -2;
    const int nb = (int)row_nbytes(m->adt, acv_width(m, id));
    memcpy(out, p, nb);
    return nb;
}
