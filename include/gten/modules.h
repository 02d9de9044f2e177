// gten/modules.h -- the gten module classes (reference gten/modules.h:11-192, gten/modules.cpp:11-254) over device tensors.
// Each module owns `weight` and a pre-allocated `acv` of max_ctx rows, resizes acv to the current n_ctx, runs one op and
// returns a shallow alias -- exactly the reference protocol, so TinyLlama (tinyllama.cpp:23-76), load_from_ckpt and
// greedy_sample compile unchanged.  This is the op-at-a-time path; gten::FusedTinyLlama (gten.h) runs the same graph
// resident on the GPU as one persistent kernel per call.
#pragma once
#include <chrono>
#include <cstdlib>
#include <unordered_map>

#include "ops.h"

namespace gten {

struct ModuleDtype { Dtype wdtype; Dtype adtype; };

// Profiling mode: device work is asynchronous, so a host clock around a launch measures launch overhead.  With profiling on
// (gten::set_profiling(true), or GTEN_PROFILE=1 in the environment) a Timer drains the library stream before it reads the
// clock at both ends, so `exec_time` is the device time of the op and print_perf (tinyllama.cpp:515-582) reports real numbers.
inline bool& profiling_flag() {
    static bool on = [] { const char* e = std::getenv("GTEN_PROFILE"); return e && *e && *e != '0'; }();
    return on;
}
inline void set_profiling(bool on) { profiling_flag() = on; }

class Timer {                       // adds whole milliseconds to *time_tracker (reference modules.h:170-192)
public:
    explicit Timer(int64_t* time_tracker) : tracker_{time_tracker} {
        if (profiling_flag()) gtb_sync();
        t0_ = std::chrono::high_resolution_clock::now();
    }
    ~Timer() { stop(); }
    void stop() {
        if (done_) return;
        if (profiling_flag()) {
            // sub-millisecond ops would all truncate to 0: carry the nanoseconds per tracker
            gtb_sync();
            const auto t1 = std::chrono::high_resolution_clock::now();
            int64_t& rem = remainder()[tracker_];
            rem += std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0_).count();
            *tracker_ += rem / 1000000;
            rem %= 1000000;
        } else {
            using ms = std::chrono::milliseconds;
            const auto t1 = std::chrono::high_resolution_clock::now();
            *tracker_ += std::chrono::time_point_cast<ms>(t1).time_since_epoch().count() - std::chrono::time_point_cast<ms>(t0_).time_since_epoch().count();
        }
        done_ = true;
    }
private:
    static std::unordered_map<int64_t*, int64_t>& remainder() { static std::unordered_map<int64_t*, int64_t> m; return m; }
    int64_t* tracker_;
    std::chrono::time_point<std::chrono::high_resolution_clock> t0_;
    bool done_ = false;
};

class Embedding {
public:
    Embedding() = default;
    Embedding(int n_vocab, int d_embed, int max_ctx, ModuleDtype dtype)
        : weight{Tensor({n_vocab, d_embed}, dtype.wdtype)}, emb_acv{Tensor({max_ctx, d_embed}, dtype.adtype)} {}
    Tensor forward(const Tensor& tokens, const int start_pos = 0) {
        Timer t{&exec_time};
        emb_acv.resize({tokens.numel(), weight.dimsize(1)});
        ops::token_embed(weight, tokens, emb_acv, start_pos);
        return emb_acv;
    }
    Tensor weight, emb_acv;
    int64_t exec_time{0};
};

class RMSNorm {
public:
    RMSNorm(int d_in, int max_ctx, ModuleDtype dtype) : weight{Tensor({d_in}, kFloat16)}, acv{Tensor({max_ctx, d_in}, dtype.adtype)} {}
    Tensor forward(const Tensor& inp, const int start_pos = 0) {
        Timer t{&exec_time};
        acv.resize({inp.dimsize(0), inp.dimsize(1)});
        ops::rms_norm(inp, weight, acv, start_pos);
        return acv;
    }
    Tensor weight, acv;
    int64_t exec_time{0};
};

class Residual {
public:
    Residual() = default;
    Residual(int max_ctx, int d_out, Dtype dtype) : acv{Tensor({max_ctx, d_out}, dtype)} {}
    Tensor forward(const Tensor& inp0, const Tensor& inp1, const int start_pos = 0) {
        Timer t{&exec_time};
        acv.resize({inp0.dimsize(0), inp0.dimsize(1)});
        ops::add(inp0, inp1, acv, start_pos);
        return acv;
    }
    Tensor acv;
    int64_t exec_time{0};
};

class Linear {
public:
    Linear() = default;
    Linear(int d_in, int d_out, int max_ctx, ModuleDtype dtype)
        : weight{Tensor({d_out, d_in}, dtype.wdtype)}, acv{Tensor({max_ctx, d_out}, dtype.adtype)}, max_ctx_{max_ctx} {}
    Tensor forward(const Tensor& inp, const int start_pos = 0) {
        Timer t{&exec_time};
        acv.resize({inp.dimsize(0), weight.dimsize(0)});
        ops::matmul_2d(inp, weight, acv, start_pos);
        return acv;
    }
    Tensor weight, acv;
    int64_t exec_time{0};
private:
    int max_ctx_{0};
};

class EmbeddingLinear {            // lm_head on the last row only, fp32 logits (reference modules.cpp:65-81)
public:
    EmbeddingLinear() = default;
    EmbeddingLinear(int n_embd, int n_vocab, int max_ctx, ModuleDtype dtype)
        : weight{Tensor({n_vocab, n_embd}, dtype.wdtype)}, acv{Tensor({n_vocab}, kFloat32)} { (void)max_ctx; }
    Tensor forward(const Tensor& inp) {
        Timer t{&exec_time};
        ops::matmul_2d(inp, weight, acv);
        return acv;
    }
    Tensor weight, acv;
    int64_t exec_time{0};
};

class Multiply {
public:
    Multiply() = default;
    Multiply(int max_ctx, int d_out, Dtype dtype, const bool inplace = false) : inplace_{inplace} {
        if (!inplace) acv = Tensor({max_ctx, d_out}, dtype);
    }
    Tensor forward(Tensor& inp0, const Tensor& inp1, const int start_pos = 0) {
        Timer t{&exec_time};
        if (inplace_) { ops::mul_inplace(inp0, inp1, start_pos); return inp0; }
        acv.resize({inp0.dimsize(0), inp0.dimsize(1)});
        ops::mul(inp0, inp1, acv, start_pos);
        return acv;
    }
    Tensor acv;
    int64_t exec_time{0};
private:
    bool inplace_{false};
};

class SiLU {
public:
    SiLU() = default;
    SiLU(int max_ctx, int d_out, Dtype dtype, const bool inplace = false) : inplace_{inplace} {
        if (!inplace) acv = Tensor({max_ctx, d_out}, dtype);
    }
    Tensor forward(Tensor& inp, const int start_pos = 0) {
        Timer t{&exec_time};
        if (inplace_) { ops::silu_inplace(inp, start_pos); return inp; }
        acv.resize({inp.dimsize(0), inp.dimsize(1)});
        ops::silu(inp, acv, start_pos);
        return acv;
    }
    Tensor acv;
    bool inplace_{false};
    int64_t exec_time{0};
};

class RotaryEmbedding {
public:
    explicit RotaryEmbedding(const int d_head, const bool inplace = true) : d_head_{d_head} { (void)inplace; }
    Tensor forward(Tensor& inp, const int start_pos = 0) {
        Timer t{&exec_time};
        ops::rotary_emb(inp, d_head_, start_pos);
        return inp;
    }
    int64_t exec_time{0};
private:
    int d_head_;
};

class SelfAttention {
public:
    SelfAttention(int n_heads, int n_embed, int n_query_groups, int max_ctx, ModuleDtype dtype)
        : query{Linear(n_embed, n_embed, max_ctx, dtype)},
          key{Linear(n_embed, (n_embed / n_heads) * n_query_groups, max_ctx, dtype)},
          value{Linear(n_embed, (n_embed / n_heads) * n_query_groups, max_ctx, dtype)},
          qkv_proj{Linear(n_embed, n_embed, max_ctx, dtype)},          // this member holds o_proj (tinyllama.cpp:362-364)
          qk_acv{Tensor({1, 1, 1}, dtype.adtype)},                     // the reference's n_heads x max_ctx^2 score buffer is not needed
          qkv_acv{Tensor({max_ctx, n_embed}, dtype.adtype)},
          q_rope{RotaryEmbedding(n_embed / n_heads)}, k_rope{RotaryEmbedding(n_embed / n_heads)},
          n_heads_{n_heads}, max_ctx_{max_ctx} {}
    Tensor forward(const Tensor& inp, const int start_pos) {
        Tensor q = query.forward(inp, start_pos);
        Tensor k = key.forward(inp, start_pos);
        q = q_rope.forward(q, start_pos);
        k = k_rope.forward(k, start_pos);
        Tensor v = value.forward(inp, start_pos);
        const Tensor a = masked_qkv_attn(q, k, v, start_pos);
        return qkv_proj.forward(a, start_pos);
    }
    Linear query, key, value, qkv_proj;
    Tensor qk_acv, qkv_acv;
    RotaryEmbedding q_rope, k_rope;
    int64_t exec_time_attn{0};
private:
    int32_t n_heads_;
    int max_ctx_;
    Tensor masked_qkv_attn(const Tensor& q, const Tensor& k, const Tensor& v, const int start_pos) {
        Timer t{&exec_time_attn};
        qkv_acv.resize({q.dimsize(0), q.dimsize(1)});
        ops::qkv_attn(q, k, v, qk_acv, qkv_acv, max_ctx_, start_pos);
        return qkv_acv;
    }
};

class AttentionBlock {
public:
    AttentionBlock(int n_heads, int d_embed, int n_query_groups, int n_mlp, int max_ctx, ModuleDtype dtype)
        : attn_norm{RMSNorm(d_embed, max_ctx, {kFloat16, dtype.adtype})},
          attn{SelfAttention(n_heads, d_embed, n_query_groups, max_ctx, dtype)},
          inp_res{Residual(max_ctx, d_embed, dtype.adtype)},
          ffn_norm{RMSNorm(d_embed, max_ctx, {kFloat16, dtype.adtype})},
          ffn_gate_proj{Linear(d_embed, n_mlp, max_ctx, dtype)},
          ffn_up_proj{Linear(d_embed, n_mlp, max_ctx, dtype)},
          ffn_down_proj{Linear(n_mlp, d_embed, max_ctx, dtype)},
          attn_res{Residual(max_ctx, d_embed, dtype.adtype)},
          ffn_mul{Multiply(max_ctx, n_mlp, dtype.adtype, /*inplace=*/true)},
          ffn_silu{SiLU(max_ctx, n_mlp, dtype.adtype, /*inplace=*/true)} {}
    Tensor forward(Tensor& inp, const int start_pos) {
        Tensor a = attn_norm.forward(inp, start_pos);
        a = attn.forward(a, start_pos);
        Tensor h = inp_res.forward(inp, a, start_pos);
        Tensor f = ffn_norm.forward(h, start_pos);
        f = ffn_forward(f, start_pos);
        return attn_res.forward(h, f, start_pos);
    }
    Tensor ffn_forward(const Tensor& inp, const int start_pos = 0) {
        Tensor g = ffn_gate_proj.forward(inp, start_pos);
        const Tensor u = ffn_up_proj.forward(inp, start_pos);
        g = ffn_silu.forward(g, start_pos);
        g = ffn_mul.forward(g, u, start_pos);
        return ffn_down_proj.forward(g, start_pos);
    }
    RMSNorm attn_norm;
    SelfAttention attn;
    Residual inp_res;
    RMSNorm ffn_norm;
    Linear ffn_gate_proj, ffn_up_proj, ffn_down_proj;
    Residual attn_res;
    Multiply ffn_mul;
    SiLU ffn_silu;
};

}  // namespace gten
