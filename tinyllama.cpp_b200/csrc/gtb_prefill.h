// gtb_prefill.h -- host interface of the batched ("fast") prefill path (gtb_prefill.cu).
//
// The reference's prefill is the same row-by-row GEMV loop as decode (gten/ops.h:613-670); the exact path
// reproduces it bit for bit.  This path processes all T prompt rows at once: every Linear becomes one
// [T x K] x [K x N] GEMM on the 5th-generation tensor cores (tcgen05.mma, fp16 operands dequantised from the
// Q8/Q4 blocks, fp32 accumulators in TMEM, operands staged by TMA), with the reference's rounding points
// (the Q8 re-encode after every op) applied in the epilogues and in the row-wise kernels.
// Summation order inside a dot product differs from the reference, so results are tolerance-checked
// (tests/test_prefill_gpu.py), not bit-checked.
#pragma once
#include <stdint.h>

#include "gtb_internal.h"

namespace gtb {

struct PfPlan;

struct PfLayerIO {
    const uint16_t* attn_norm;
    const uint16_t* ffn_norm;
    uint8_t* kq; uint16_t* ks; uint8_t* vq; uint16_t* vs;      // this layer's K/V cache (engine layout)
};

struct PfRun {
    const int32_t* d_tokens;        // [T] token ids (device)
    int T;
    const gtb_weight* embed;        // device-layout embedding table
    const PfLayerIO* layers;        // [n_layers] (host array)
    const float* rope_cos;          // [max_ctx][32]
    const float* rope_sin;
    float* last_res;                // out: dequantised residual stream of row T-1 before the last `down` is added [n_embd]
    float* last_down;               // out: dequantised last-layer `down` output of row T-1 [n_embd]
    float* cap;                     // optional capture: [n_layers][12][T][capw] fp32 (+ embedding at the end), or null
    int capw;
    int n_layers_run;               // <= n_layers (debug: stop early); last_res/last_down are those of the last layer run
};

int pf_create(PfPlan** out, const gtb_model_config& cfg);
void pf_destroy(PfPlan* p);
// which: 0 = q|k|v (rows n_embd + 2*kv_dim), 1 = o, 2 = gate|up (rows 2*n_ffn), 3 = down; data/scales in the DEVICE layout
int pf_set_weight(PfPlan* p, int layer, int which, int wdtype, const void* d_data, const uint16_t* d_scales, int rows, int cols);
bool pf_weights_ready(const PfPlan* p);
int pf_run(PfPlan* p, const PfRun& r);
void pf_set_two_cta(PfPlan* p, bool on);   // CTA-pair (tcgen05 cta_group::2) GEMM for the 256-wide tiles
void pf_set_pdl(bool on);                  // programmatic dependent launch between the kernels of the chain (default on)
void pf_set_attn_two_pass(PfPlan* p, bool on);   // false: single-sweep attention (P-row block scales applied unrounded)
void pf_set_fused(PfPlan* p, bool on);     // false: RoPE/KV append and SiLU*up as separate kernels after plain GEMMs (debug)
size_t pf_bytes(const PfPlan* p);
int64_t pf_launches_last(const PfPlan* p);

// stand-alone GEMM for tests: C[M][N] = A[M][K] . W[N][K]^T, fp16 inputs (device), fp32 output (device)
int pf_gemm_f32(const void* d_A16, const void* d_W16, float* d_C, int M, int N, int K, int bn);

}  // namespace gtb
