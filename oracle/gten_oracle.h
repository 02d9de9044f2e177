/* oracle/gten_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the tinyllama.cpp forward hot path as computed by the
 * reference's `-O3 -fopenmp -mavx -mf16c` build.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Pinned against oracle/_ref (the unmodified reference compiled here) by
 * tests/test_oracle_pin.py and against the committed vectors in tests/golden/.
 *
 * Dtype codes follow gten_types.h:20-26: 0 Int32, 1 Float16, 2 Float32, 3 Qint8, 4 Qint4.
 */
#ifndef GTEN_ORACLE_H
#define GTEN_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_I32 = 0, ORC_F16 = 1, ORC_F32 = 2, ORC_Q8 = 3, ORC_Q4 = 4 };

/* tensor ids (shared with ref_harness.cpp and include/gten_b200.h) */
enum {
    ORC_T_EMBED = 0, ORC_T_FINAL_NORM = 1, ORC_T_LM_HEAD = 2,
    ORC_T_Q = 10, ORC_T_K = 11, ORC_T_V = 12, ORC_T_O = 13, ORC_T_GATE = 14, ORC_T_UP = 15, ORC_T_DOWN = 16,
    ORC_T_ATTN_NORM = 17, ORC_T_FFN_NORM = 18
};
/* activation ids: the per-layer rounding points of SURVEY.md App. A */
enum {
    ORC_A_EMB = 0, ORC_A_FINAL_NORM = 1,
    ORC_A_ATTN_NORM = 10, ORC_A_Q = 11, ORC_A_K = 12, ORC_A_V = 13, ORC_A_ATTN_OUT = 14, ORC_A_O = 15,
    ORC_A_INP_RES = 16, ORC_A_FFN_NORM = 17, ORC_A_GATE = 18, ORC_A_UP = 19, ORC_A_DOWN = 20, ORC_A_ATTN_RES = 21
};

const char* orc_build_info(void);

/* scalar / row codecs */
uint16_t orc_fp32_to_fp16(float f);
float    orc_fp16_to_fp32(uint16_t h);
void  orc_q8_quantize_row(const float* inp, void* out, int n);
void  orc_q8_dequantize_row(const void* inp, float* out, int n);
void  orc_q4_dequantize_row(const void* inp, float* out, int n);
void  orc_read_row_to_float(const void* inp, int dtype, float* out, int n);
void  orc_write_row_from_float(float* inp, void* out, int dtype, int n);
float orc_vec_dot_product(const void* a, int adt, const void* b, int bdt, int n);

/* ops (same argument meaning as the ref_* wrappers in ref_harness.cpp) */
void orc_token_embed(const void* w, int wdt, int n_vocab, int n_embd, const int32_t* tokens, int n_ctx,
                     void* out, int odt, int start_pos);
void orc_matmul_2d(const void* x, int xdt, int n_ctx, int k, const void* w, int wdt, int n_out,
                   void* out, int odt, int out_1d, int start_pos);
void orc_rms_norm(const void* x, int xdt, int n_ctx, int n_embd, const uint16_t* w, void* out, int start_pos);
void orc_rotary_emb(void* x, int xdt, int n_ctx, int n_embd, int d_head, int start_pos);
void orc_silu(const void* x, int xdt, int n_ctx, int n_embd, void* out, int start_pos);
void orc_mul(const void* a, const void* b, int xdt, int n_ctx, int n_embd, void* out, int start_pos);
void orc_add(const void* a, const void* b, int xdt, int n_ctx, int n_embd, void* out, int start_pos);
void orc_qkv_attn(const void* q, const void* k, const void* v, void* qk, void* out, int xdt,
                  int n_ctx, int n_heads, int n_kv_heads, int d_head, int max_ctx, int start_pos);
float orc_expf(float x);
void orc_expf_bits_range(uint32_t first, uint32_t count, float* out);
void  orc_rope_angles(int pos, int d_head, float* cos_out, float* sin_out);

/* model */
void* orc_model_new(int n_vocab, int n_embd, int n_ffn, int n_layers, int n_heads, int n_groups,
                    int max_ctx, int wdtype);
void  orc_model_free(void* m);
void* orc_model_weight(void* m, int layer, int id, int64_t* nbytes);
void  orc_model_logits(void* m, const int32_t* tokens, int n_tokens, int start_pos, float* out);
void  orc_model_generate(void* m, int32_t* tokens, int n_prompt, int n_new, double* times, float* logits_out);
int   orc_model_acv(void* m, int layer, int id, int row, float* out);
int   orc_model_acv_raw(void* m, int layer, int id, int row, void* out);
/* choose which sequence row's activations orc_model_acv returns (default: last row processed) */
void  orc_model_capture_row(void* m, int row);

#ifdef __cplusplus
}
#endif
#endif
