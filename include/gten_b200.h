/* gten_b200.h -- C-ABI of libgten_b200.so: the B200 (sm_100a) implementation of tinyllama.cpp's
 * transformer-forward hot path.
 *
 * The reference has no FFI; its boundary is the C++ `gten` API.  This header is the thin
 * C layer the C++ drop-in headers in include/gten/ call, and what any other host language would bind
 * (ctypes stub: tinyllama.cpp_b200/capi.py).  Each entry point names the reference
 * interface it stands in for (paths relative to the reference repo).
 *
 * Conventions: every function returns 0 on success and a non-zero code on failure, with a message in
 * gtb_last_error(); nothing throws across the boundary.  The C++ wrappers turn non-zero into the
 * reference's assert-and-exit convention (gten/log.h:6-23).  One host thread <-> one CUDA stream <->
 * one GPU (the reference API is single-threaded and not re-entrant, gten/ops.h:37).
 * Pointers named d_* are device pointers obtained from gtb_malloc; h_* are host pointers.
 * There is NO CPU fallback: without a CUDA device every call fails with GTB_ERR_NO_DEVICE.
 */
#ifndef GTEN_B200_H
#define GTEN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dtype codes = enum class Dtype, gten/gten_types.h:20-26 */
enum { GTB_I32 = 0, GTB_F16 = 1, GTB_F32 = 2, GTB_Q8 = 3, GTB_Q4 = 4 };
enum { GTB_OK = 0, GTB_ERR_CUDA = 1, GTB_ERR_ARG = 2, GTB_ERR_NO_DEVICE = 3, GTB_ERR_STATE = 4 };

/* tensor ids of a TinyLlama checkpoint in file order (tinyllama.cpp:345-391) */
enum {
    GTB_T_EMBED = 0, GTB_T_FINAL_NORM = 1, GTB_T_LM_HEAD = 2,
    GTB_T_Q = 10, GTB_T_K = 11, GTB_T_V = 12, GTB_T_O = 13, GTB_T_GATE = 14, GTB_T_UP = 15, GTB_T_DOWN = 16,
    GTB_T_ATTN_NORM = 17, GTB_T_FFN_NORM = 18
};
/* activation ids: the module `acv` buffers of one AttentionBlock (gten/modules.h:147-167) */
enum {
    GTB_A_EMB = 0, GTB_A_FINAL_NORM = 1,
    GTB_A_ATTN_NORM = 10, GTB_A_Q = 11, GTB_A_K = 12, GTB_A_V = 13, GTB_A_ATTN_OUT = 14, GTB_A_O = 15,
    GTB_A_INP_RES = 16, GTB_A_FFN_NORM = 17, GTB_A_GATE = 18, GTB_A_UP = 19, GTB_A_DOWN = 20, GTB_A_ATTN_RES = 21
};

/* ---- context ------------------------------------------------------------------------------------- */
const char* gtb_last_error(void);
const char* gtb_version(void);
int gtb_init(int device);                       /* select device, create the stream; idempotent */
int gtb_device_count(int* n);
int gtb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
int gtb_sync(void);                             /* wait for the library stream */
void* gtb_stream(void);                         /* the cudaStream_t every launch uses (for CUDA-event timing) */
int64_t gtb_launch_count(void);                 /* kernels launched by this library since load */
int64_t gtb_mem_allocated(void);                /* replaces G_TensorMemAllocated, gten/tensor.h:17 */

/* ---- memory (Tensor storage, gten/tensor.cpp:27-67) ---------------------------------------------- */
int gtb_malloc(void** d_ptr, size_t nbytes);
int gtb_free(void* d_ptr);
int gtb_memset(void* d_ptr, int value, size_t nbytes);
int gtb_h2d(void* d_dst, const void* h_src, size_t nbytes);     /* async on the library stream */
int gtb_d2h(void* h_dst, const void* d_src, size_t nbytes);     /* returns after the copy completed */
int gtb_d2d(void* d_dst, const void* d_src, size_t nbytes);
int gtb_host_alloc(void** h_ptr, size_t nbytes);                /* pinned staging */
int gtb_host_free(void* h_ptr);

/* ---- weights: upload = one-time repack of a gten payload into the device layout ------------------ */
/* (payload layout = the bytes read_into_weight stores, tinyllama.cpp:301-321: Q4Block/Q8Block rows or fp16) */
typedef struct gtb_weight* gtb_weight_t;
int gtb_weight_upload(gtb_weight_t* out, const void* h_payload, int dtype, int rows, int cols);
int gtb_weight_from_device(gtb_weight_t* out, const void* d_payload, int dtype, int rows, int cols);
int gtb_weight_free(gtb_weight_t w);
int gtb_weight_nbytes(gtb_weight_t w, size_t* nbytes);
/* dequantise rows [row0,row0+nrows) from the DEVICE layout to fp32 (bit-exact vs quants.h:69-90) */
int gtb_weight_dequant(gtb_weight_t w, int row0, int nrows, float* h_out);

/* ---- row codecs (gten/quants.h:92-143, gten/ops.h:40-96); buffers hold `rows` rows, reference layout */
int gtb_write_rows_from_float(const float* d_in, void* d_out, int out_dtype, int rows, int n);
int gtb_read_rows_to_float(const void* d_in, int in_dtype, float* d_out, int rows, int n);

/* ops::vec_dot_product (gten/ops.h:482-512): dot of two HOST rows in the reference layout, evaluated on the device in the
 * reference's AVX order; dtype pairs (Q8, Q8), (Q8, Q4), (F16, F16), (F32, F32) like the reference's switch */
int gtb_vec_dot_product(const void* h_a, int a_dtype, const void* h_b, int b_dtype, int n, float* h_out);

/* ---- ops (gten/ops.h); activations are device buffers in the reference's own row layout ---------- */
int gtb_token_embed(gtb_weight_t w, const int32_t* d_tokens, void* d_out, int out_dtype, int n_ctx, int start_pos); /* ops.h:554 */
int gtb_matmul_2d(const void* d_x, int x_dtype, int n_ctx, gtb_weight_t w, void* d_out, int out_dtype,
                  int out_is_1d, int start_pos);                                                                    /* ops.h:651 */
int gtb_rms_norm(const void* d_x, int dtype, int n_ctx, int n_embd, const void* d_weight_f16, void* d_out, int start_pos); /* ops.h:806 */
int gtb_rotary_emb(void* d_x, int dtype, int n_ctx, int n_embd, int d_head, int start_pos);                        /* ops.h:757 */
int gtb_silu(const void* d_x, int dtype, int n_ctx, int n_embd, void* d_out, int start_pos);                       /* ops.h:700,708 */
int gtb_mul(const void* d_a, const void* d_b, int dtype, int n_ctx, int n_embd, void* d_out, int start_pos);       /* ops.h:853,861 */
int gtb_add(const void* d_a, const void* d_b, int dtype, int n_ctx, int n_embd, void* d_out, int start_pos);       /* ops.h:900 */
/* causal GQA attention over rows [start_pos, n_ctx); K/V rows 0..n_ctx-1 are read from d_k/d_v (the K/V
 * Linear outputs ARE the cache, gten/modules.cpp:196-201).  d_qk may be NULL: scores stay on chip. */
int gtb_qkv_attn(const void* d_q, const void* d_k, const void* d_v, void* d_qk, void* d_out, int dtype,
                 int n_ctx, int n_heads, int n_kv_heads, int d_head, int max_ctx, int start_pos);                  /* ops.h:1118 */

/* ---- engine: the whole TinyLlama::logits graph (tinyllama.cpp:45-61) resident on the GPU ---------- */
typedef struct gtb_engine* gtb_engine_t;
typedef struct {
    int n_vocab, n_embd, n_ffn, n_layers, n_heads, n_groups;   /* TinyLLamaParams, tinyllama.cpp:12-20 */
    int max_ctx;                                               /* TinyLlama{n_ctx, dtype} */
    int wdtype;                                                /* GTB_F16 / GTB_Q8 / GTB_Q4; activations follow tinyllama.cpp:258-265 */
} gtb_model_config;

int gtb_engine_create(gtb_engine_t* out, const gtb_model_config* cfg);
int gtb_engine_destroy(gtb_engine_t e);
/* load_from_ckpt's per-tensor step (tinyllama.cpp:301-321): payload size is checked like the reference does */
int gtb_engine_set_weight(gtb_engine_t e, int layer, int tensor_id, const void* h_payload, size_t nbytes);
int gtb_engine_load_gten(gtb_engine_t e, const char* path);              /* tinyllama.cpp:336-392 */
/* TinyLlama::logits(tokens, start_pos): tokens = ALL ids so far (host), rows [start_pos, n) are computed,
 * logits of the last row are written to h_logits (n_vocab floats).  Order-exact (greedy-identical) path. */
int gtb_engine_logits(gtb_engine_t e, const int32_t* h_tokens, int n_tokens, int start_pos, float* h_logits);
/* greedy_sample's loop (tinyllama.cpp:395-440) kept on the device: prefill rows [0,n_prompt), then n_new
 * argmax steps (strict '>', lowest index wins); h_tokens has room for n_prompt+n_new ids.  No EOS stop
 * unless eos_id >= 0.  Returns the number of generated ids in *n_generated. */
int gtb_engine_generate(gtb_engine_t e, int32_t* h_tokens, int n_prompt, int n_new, int eos_id, int* n_generated);
/* fine-grained control used by bench.py and the tests */
int gtb_engine_reset(gtb_engine_t e);
/* exact path, rows [0,n): up to 512 rows per pass through the multi-row kernels (gtb_xrows.cu), bit-identical to
 * the reference's row loop (gten/ops.h:632); the last row also samples the first new token */
int gtb_engine_prefill(gtb_engine_t e, const int32_t* h_tokens, int n_tokens);
int gtb_engine_decode(gtb_engine_t e, int n_steps);                                /* n greedy steps, device-resident */
/* Batched prefill of rows [0,n): every Linear of the n rows (ops.h:613-670) is one tcgen05/TMEM GEMM fed by TMA
 * (fp16 operands dequantised from the Q8/Q4 blocks), the reference's re-encode points are applied in the epilogues,
 * causal GQA attention runs on the tensor cores; K/V land in the cache and the last row's logits / first greedy
 * token come from the order-exact lm_head phase, so gtb_engine_decode continues from here.  Summation order differs
 * from ops.h:224-391: results match the reference within a tolerance (tests/test_prefill_gpu.py), not bit for bit.
 * All three formats: Q8 and Q4 weights with Q8 activations, FP16 weights with FP16 activations (tinyllama.cpp:258-265). */
int gtb_engine_prefill_fast(gtb_engine_t e, const int32_t* h_tokens, int n_tokens);
/* decoded fp32 row `row` of a module activation of the last gtb_engine_prefill_fast call ("capture_acv" on) */
int gtb_engine_pf_acv(gtb_engine_t e, int layer, int acv_id, int row, float* h_out, int* width);
/* self-test of the tcgen05 GEMM: C[M][N] = A[M][K] . W[N][K]^T, fp16 host inputs, fp32 host output; bn = 128 or 256 */
int gtb_pf_gemm_f32(const void* h_A16, const void* h_W16, int M, int N, int K, int bn, float* h_C);
int gtb_engine_position(gtb_engine_t e, int* pos);
int gtb_engine_read_tokens(gtb_engine_t e, int32_t* h_tokens, int first, int count);
int gtb_engine_read_logits(gtb_engine_t e, float* h_logits);
/* the k (<= 64) largest logits of the last processed row and their token ids, largest first, ties to the lower id: the
 * candidate set of topk_sample (tinyllama.cpp:466-478), so that 8k bytes instead of 128 KB return to the host per token;
 * temperature, softmax and the draw from std::discrete_distribution stay on the host with the caller's generator */
int gtb_engine_topk(gtb_engine_t e, int k, float* h_values, int32_t* h_ids);
/* decoded fp32 row of a module activation for the LAST processed row (debug/parity) */
int gtb_engine_acv(gtb_engine_t e, int layer, int acv_id, float* h_out, int* width);
/* options: "mega" (1: persistent cooperative kernel, default; 0: one kernel per phase), "graph" (CUDA-graph replay of
 * the per-phase path), "capture_acv", "grid", "pf_ahead" (L2 prefetch distance in GEMV phases), "prof",
 * "fast_decode" (1: rows run through the order-free kernels of gtb_fastdec.cuh -- 128-bit streaming GEMV with warp-shuffle
 * reductions, block-reduced RMSNorm, position-split attention with an online-softmax combine; same operations and re-encode
 * points, free summation order: tolerance-level parity like gtb_engine_prefill_fast, Q8/Q4 models; default 0),
 * "fd_mega" (fast_decode as one persistent cooperative kernel with grid barriers instead of the PDL-chained kernels, default 0),
 * "fd_ahead" (fast_decode: L2 look-ahead distance in GEMV steps, default 3), "fd_prof_cta" (which CTA writes the "prof" stamps),
 * "pf_layers" (debug: batched prefill stops after this many layers), "pf_fused" (RoPE/KV append and SiLU*up inside the
 * GEMM epilogues, default 1), "pf_pdl" (programmatic dependent launch, default 1), "pf_2cta" (CTA-pair tcgen05 GEMM, default 0),
 * "pf_attn2" (second attention sweep that also reproduces the fp16 rounding of the probability-row block scales, default 0),
 * "xrows" (1, default: runs of >= "xr_min_rows" (4) rows whose positions the host knows -- gtb_engine_prefill, _logits, _generate --
 * go through the order-exact multi-row kernels in passes of "xr_rows" (512, at most 1024) rows: same bits as the row-at-a-time kernels, one weight
 * read per pass; 0: every row through the persistent kernel), "batch_exact" (see gtb_engine_batch_*),
 * "xr_tensor" (process-wide experiment, default 0: the multi-row Linears of Q4 / Q8 models on the tensor cores -- kind::f16 tcgen05 MMAs
 * return the integer lane sums exactly, epilogue warps run the ordered fp32 chains out of TMEM; same bits, measured slower end to end),
 * "prof_cta" (which CTA of the persistent kernel writes the "prof" stamps; "prof" itself selects the stamped build of the kernel),
 * "batch_eos" (see gtb_engine_batch_*).  Tuning / debugging switches, not part of the contract: "fd_chunk" (positions per attention chunk of the
 * order-free path), "fd_trace" and "xr_trace" (cycle counters of the order-free chain / of the tensor-core experiment into the "prof"
 * buffer), "xr_pdl" (programmatic dependent launch inside multi-row passes of <= 16 rows, default 1), "xr_variant" (tile-shape experiments
 * of the multi-row GEMMs, default 0) */
int gtb_engine_set_option(gtb_engine_t e, const char* name, int value);

/* Batched decode (BASELINE.json configs[4]; the reference decodes one sequence, tinyllama.cpp:395-440): up to 64
 * sequences advance together and every weight block is loaded once per step for all of them.  Default ("batch_exact" = 1): the
 * ORDER-EXACT multi-row kernels (gtb_xrows.cu) -- every sequence's tokens and logits are bit-identical to the same sequence decoded
 * alone by gtb_engine_decode, i.e. to the reference.  "batch_exact" = 0: the order-free kernels (tolerance contract of "fast_decode",
 * at most 16 sequences).
 *   gtb_engine_batch_create(e, n)        allocate n slots (own K/V cache, tokens, position each); n = 0 frees them
 *   gtb_engine_batch_prefill(e, s, t, n) exact prefill of n prompt ids into slot s (multi-row passes of up to 512 rows), first token appended
 *   gtb_engine_batch_adopt(e, s)         slot s <- the engine's current sequence (after gtb_engine_prefill / _prefill_fast / decode)
 *   gtb_engine_batch_decode(e, k)        k greedy steps of every slot (device-side argmax, tokens stay on the device)
 *   gtb_engine_batch_position / _read_tokens / _read_logits: per-slot state
 * Join / leave: a slot can be (re)filled with gtb_engine_batch_prefill or _adopt between decode calls while the others keep their state; with
 * option "batch_eos" = id (exact mode) a slot that samples id keeps it as its last token and leaves the batch -- its row is no longer
 * computed (tinyllama.cpp:426 `break`), the other slots' bits do not change. */
int gtb_engine_batch_create(gtb_engine_t e, int n_seq);
int gtb_engine_batch_prefill(gtb_engine_t e, int seq, const int32_t* h_tokens, int n_tokens);
int gtb_engine_batch_adopt(gtb_engine_t e, int seq);
int gtb_engine_batch_decode(gtb_engine_t e, int n_steps);
int gtb_engine_batch_position(gtb_engine_t e, int seq, int* pos);
int gtb_engine_batch_read_tokens(gtb_engine_t e, int seq, int32_t* h_tokens, int first, int count);
int gtb_engine_batch_read_logits(gtb_engine_t e, int seq, float* h_logits);
/* "prof": globaltimer stamps (ns) taken by CTA 0 at every phase boundary of the last processed row */
int gtb_engine_read_prof(gtb_engine_t e, long long* h_out, int count);
/* 1 if rows run in the persistent kernel (TinyLlama dimensions, max_ctx <= 2048, grid >= 128), 0 if one kernel per phase */
int gtb_engine_uses_megakernel(gtb_engine_t e, int* yes);
int gtb_engine_weight_bytes(gtb_engine_t e, size_t* nbytes);
/* self-test: the persistent kernel's parallel emulation of the reference's strictly in-order fp32 sum
 * (gten/ops.h:765-767, 982-988) on n non-negative host terms; h_out[0] must equal the sequential sum bit for bit,
 * h_out[1..4] receive the SM cycles of four back-to-back evaluations (cold and warm instruction cache) */
int gtb_selftest_exact_sum(const float* h_terms, int n, float* h_out);
/* self-test: the device restatement of glibc 2.39 expf (gten/ops.h:692, 985 call it) on the float bit patterns
 * [first_bits, first_bits + count), count <= 2^26; tests/test_expf_gpu.py walks all 2^32 inputs against the host libm */
int gtb_selftest_expf(uint32_t first_bits, uint32_t count, float* h_out);

#ifdef __cplusplus
}
#endif
#endif /* GTEN_B200_H */
