// gtb_xrows.h -- host interface of the order-exact MULTI-ROW path (gtb_xrows.cu).
//
// The reference computes every row of TinyLlama::logits independently given its tokens (gten/ops.h:632 loops rows,
// tinyllama.cpp:395-440 loops tokens); a row's bits depend only on its own ordered chains.  This path runs R rows side
// by side -- R consecutive prompt rows of one sequence (exact prefill) or one row of each of R sequences (exact
// batched decode, SURVEY.md 8 f3 / BASELINE.json configs[4]) -- against ONE load of every weight block, and every
// (row, output, lane) still runs the reference's own ordered fp32 chain (ops.h:282-292), so logits and greedy tokens
// are bit-identical to the row-at-a-time kernels and to the reference's -mavx -mf16c build.
#pragma once
#include <stdint.h>

#include "gtb_internal.h"
#include "gtb_kernels.cuh"

namespace gtb {

struct XrPlan;

struct XrLayerW {                      // one layer's weights in the device layout (gtb_internal.h)
    const uint4* w[4];                 // q|k|v, o, gate|up, down
    const uint16_t* s[4];
    const uint16_t* attn_norm;
    const uint16_t* ffn_norm;
};

struct XrModel {
    gtb_model_config cfg;
    const void* emb_w; const uint16_t* emb_s;
    const uint4* head_w; const uint16_t* head_s;
    const uint16_t* final_norm;
    const float* rope_cos; const float* rope_sin;
    const XrLayerW* layers;            // host array [n_layers]
};

struct XrKV {                          // K/V caches of the sequence slots this pass touches
    uint8_t* const* kq; uint16_t* const* ks; uint8_t* const* vq; uint16_t* const* vs;   // host arrays [n_layers] of device pointers
    size_t slot_codes, slot_scales;    // element strides between slots (0 for the engine's own single sequence)
};

struct XrSeq {                         // token rows and positions of the slots
    int32_t* tokens; int tok_stride;   // tokens[slot * tok_stride + pos]
    DevState* st;                      // [n_slots]
};

constexpr int XR_MAX_ROWS = 1024;      // rows per pass (prefill); the launch overheads and chain latencies of a pass are paid once per pass
constexpr int XR_MAX_SLOTS = 64;       // sequences of a batched decode

int xr_create(XrPlan** out, const gtb_model_config& cfg);
void xr_destroy(XrPlan* p);
bool xr_supported(const gtb_model_config& cfg, int gsz);
void xr_set_variant(int v);          // tuning experiments (large-N GEMM configuration)
void xr_set_tensor(bool on);         // experiment: Q4 / Q8 Linears on the tensor cores (gtb_xtensor.cuh); default off = the SIMT dp4a kernel
void xr_set_trace(long long* d_buf);  // debugging: cycle counters of the tensor-core GEMM into a device buffer of >= 32 words (nullptr: off)
void xr_set_pdl(bool on);            // programmatic dependent launch inside a pass (default on)

// Prefill pass: rows = positions [p0, p0 + n_rows) of slot `slot`; n_ctx = the call's row count (P.V lane split, SURVEY
// App. A).  with_head: the LAST row also runs final norm + lm_head + argmax, appends the token and advances the slot.
int xr_prefill_pass(XrPlan* p, const XrModel& m, const XrKV& kv, const XrSeq& sq, int slot, int p0, int n_rows, int n_ctx,
                    bool with_head, int eos_id, float* d_logits);
// Decode pass: one row of each of slots [0, n_slots) at its own position; head + argmax for every row.
// Enqueues the kernels only (capturable into a CUDA graph: positions live in sq.st).
int xr_decode_pass(XrPlan* p, const XrModel& m, const XrKV& kv, const XrSeq& sq, int n_slots, int t_cap, int eos_id,
                   float* d_logits /* [n_slots][n_vocab] */);

}  // namespace gtb
