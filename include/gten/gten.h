// gten/gten.h -- umbrella header of the gten API on B200 (the reference's gten/gten.h:3-8 is a unity include of .cpp files;
// these are ordinary headers over libgten_b200.so).  `#include "gten/gten.h"` + `-lgten_b200` replaces the reference's gten/.
#pragma once
// the reference's unity include drags these in for the application (tensor.cpp, ops.h); keep them so it compiles unchanged
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "log.h"
#include "gten_types.h"
#include "quants.h"
#include "tensor.h"
#include "ops.h"
#include "modules.h"

namespace gten {

/// The whole TinyLlama::logits graph (tinyllama.cpp:45-61) resident on the GPU: one persistent kernel per call instead of
/// one kernel per op.  Same calling protocol as the reference class: logits(tokens, start_pos) takes ALL token ids so far.
class FusedTinyLlama {
public:
    FusedTinyLlama(const int n_ctx, ModuleDtype dtype, int n_vocab = 32003, int n_embd = 2048, int n_ffn = 5632, int n_layers = 22,
                   int n_heads = 32, int n_query_groups = 4)
        : n_ctx_{n_ctx}, logits_{Tensor({n_vocab}, kFloat32)} {
        gtb_model_config c{n_vocab, n_embd, n_ffn, n_layers, n_heads, n_query_groups, n_ctx, (int)dtype.wdtype};
        GTEN_CUDA_OK(gtb_engine_create(&eng_, &c));
    }
    ~FusedTinyLlama() { if (eng_) gtb_engine_destroy(eng_); }
    FusedTinyLlama(const FusedTinyLlama&) = delete;
    FusedTinyLlama& operator=(const FusedTinyLlama&) = delete;
    void load_from_ckpt(const char* path) { GTEN_CUDA_OK(gtb_engine_load_gten(eng_, path)); }
    // Opt in to the batched tensor-core prefill (gtb_engine_prefill_fast) for the prompt call logits(tokens, 0): every Linear
    // of the prompt rows becomes one tcgen05 GEMM.  Its results match the reference within a tolerance, not bit for bit
    // (DESIGN.md 4.4); later calls (start_pos > 0) always take the order-exact path.  Prompts of >= 2 tokens.
    void set_batched_prefill(bool on) { batched_prefill_ = on; }
    // Opt in to the order-free decode kernels (gtb_fastdec.cuh) for every row: same operations and re-encode points as
    // ops.h, free summation order -- 1.5x the token rate of the bit-exact path (0.56 vs 0.85 ms per token), results within the same tolerance as the
    // batched prefill, greedy tokens no longer guaranteed identical.  Q8 / Q4 models.
    void set_fast_decode(bool on) { GTEN_CUDA_OK(gtb_engine_set_option(eng_, "fast_decode", on ? 1 : 0)); }
    Tensor logits(const Tensor& tokens, const int start_pos = 0) {
        if (tokens.numel() > n_ctx_) {
            std::cerr << "Number of prompt tokens (" << tokens.numel() << ") exceed provided maximum ctx size (" << n_ctx_ << ")\n";
            std::exit(EXIT_FAILURE);
        }
        if (batched_prefill_ && start_pos == 0 && tokens.numel() >= 2 && tokens.numel() < n_ctx_) {
            GTEN_CUDA_OK(gtb_engine_prefill_fast(eng_, tokens.data_ptr<int32_t>(), tokens.numel()));
            GTEN_CUDA_OK(gtb_engine_read_logits(eng_, logits_.data_ptr<float>()));
            return logits_;
        }
        GTEN_CUDA_OK(gtb_engine_logits(eng_, tokens.data_ptr<int32_t>(), tokens.numel(), start_pos, logits_.data_ptr<float>()));
        return logits_;
    }
    // candidate set of topk_sample (tinyllama.cpp:466-478) selected on the device: values[k], ids[k], largest first
    void topk(int k, float* values, int32_t* ids) { GTEN_CUDA_OK(gtb_engine_topk(eng_, k, values, ids)); }
    // Batched decode (gtb_engine_batch_*, include/gten_b200.h): up to 64 sequences advance together and share every weight read,
    // through the ORDER-EXACT multi-row kernels: every slot's tokens are the reference's.  batch_prefill(s, ids, n) prefills a prompt
    // straight into slot s (up to 512 rows per pass); batch_adopt(s) moves the sequence of the last logits()/prefill call into slot s;
    // batch_decode(k) runs k greedy steps of all slots.
    void batch_create(int n_seq) { GTEN_CUDA_OK(gtb_engine_batch_create(eng_, n_seq)); }
    void batch_prefill(int slot, const int32_t* ids, int n) { GTEN_CUDA_OK(gtb_engine_batch_prefill(eng_, slot, ids, n)); }
    void batch_adopt(int slot) { GTEN_CUDA_OK(gtb_engine_batch_adopt(eng_, slot)); }
    void batch_decode(int n_steps) { GTEN_CUDA_OK(gtb_engine_batch_decode(eng_, n_steps)); }
    // a slot that samples `eot_token` keeps it as its last token and leaves the batch (tinyllama.cpp:426 `break`); -1: never
    void batch_set_eos(int eot_token) { GTEN_CUDA_OK(gtb_engine_set_option(eng_, "batch_eos", eot_token)); }
    int batch_position(int slot) { int p = 0; GTEN_CUDA_OK(gtb_engine_batch_position(eng_, slot, &p)); return p; }
    void batch_read_tokens(int slot, int32_t* out, int first, int count) { GTEN_CUDA_OK(gtb_engine_batch_read_tokens(eng_, slot, out, first, count)); }
    gtb_engine_t engine() { return eng_; }
private:
    int n_ctx_;
    gtb_engine_t eng_ = nullptr;
    Tensor logits_;
    bool batched_prefill_ = false;
};

}  // namespace gten
