// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" shell around the UNMODIFIED reference sources, which are
// compiled from where they lie under /root/reference (nothing is copied into
// this repository).  The build recipe is oracle/Makefile; the output goes to
// oracle/_ref/libgten_ref.so (git-ignored, travels to the GPU box).
//
// Build flags are the reference's own "fast" line (README.md:25):
//     g++ -std=c++17 -O3 -fopenmp -mavx -mf16c
// which is the build SURVEY.md App. A names as *the* oracle (the scalar build
// associates sums differently and is a different function).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// How the reference is driven: its `main` shells out to
// a downloader, so it is renamed away; `private` is opened so that weights can
// be injected into TinyLlama's members and per-module activations dumped.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <string_view>
#include <vector>

#define private public
#define main tinyllama_reference_main
#include "/root/reference/tinyllama.cpp"
#undef main
#undef private

namespace {

using gten::Dtype;
using gten::Tensor;

Dtype dt(int code) {
    // codes follow the enum order of gten_types.h:20-26
    switch (code) {
        case 0: return gten::kInt32;
        case 1: return gten::kFloat16;
        case 2: return gten::kFloat32;
        case 3: return gten::kQint8;
        case 4: return gten::kQint4;
    }
    std::fprintf(stderr, "ref_harness: bad dtype code %d\n", code);
    std::abort();
}

// A reduced-size twin of TinyLlama (tinyllama.cpp:23-76) built from the
// reference's own modules, so that parity tests can run a 1-2 layer network
// with a small vocabulary in milliseconds.  The graph is the one in
// TinyLlama::logits (tinyllama.cpp:45-61), nothing else.
struct MiniLlama {
    int n_ctx;
    gten::ModuleDtype dtype;
    gten::Embedding tok_emb;
    gten::RMSNorm norm;
    gten::EmbeddingLinear lm_head;
    std::vector<gten::AttentionBlock> blocks;

    MiniLlama(int n_vocab, int n_embd, int n_ffn, int n_layers, int n_heads, int n_groups,
              int max_ctx, gten::ModuleDtype d)
        : n_ctx{max_ctx}, dtype{d},
          tok_emb{gten::Embedding(n_vocab, n_embd, max_ctx, d)},
          norm{gten::RMSNorm(n_embd, max_ctx, {gten::kFloat16, d.adtype})},
          lm_head{gten::EmbeddingLinear{n_embd, n_vocab, max_ctx, {d.wdtype, gten::kFloat32}}}
    {
        blocks.reserve(n_layers);
        for (int i = 0; i < n_layers; i++)
            blocks.push_back(gten::AttentionBlock(n_heads, n_embd, n_groups, n_ffn, max_ctx, d));
    }

    Tensor logits(const Tensor& tokens, int start_pos) {
        Tensor x = tok_emb.forward(tokens, start_pos);
        for (auto& b : blocks) x = b.forward(x, start_pos);
        x = norm.forward(x, start_pos);
        return lm_head.forward(x);
    }
};

struct RefModel {
    std::unique_ptr<TinyLlama> full;
    std::unique_ptr<MiniLlama> mini;
    int n_vocab = 0, n_layers = 0;

    gten::Embedding& emb() { return full ? full->tok_emb_ : mini->tok_emb; }
    gten::RMSNorm& fnorm() { return full ? full->norm_ : mini->norm; }
    gten::EmbeddingLinear& head() { return full ? full->lm_head_ : mini->lm_head; }
    gten::AttentionBlock& block(int i) { return full ? full->blocks_[i] : mini->blocks[i]; }
    Tensor logits(const Tensor& t, int sp) { return full ? full->logits(t, sp) : mini->logits(t, sp); }
};

// Tensor ids shared with oracle/gten_oracle.c and the product's C-ABI.
enum TensorId {
    T_EMBED = 0, T_FINAL_NORM = 1, T_LM_HEAD = 2,
    T_Q = 10, T_K = 11, T_V = 12, T_O = 13, T_GATE = 14, T_UP = 15, T_DOWN = 16,
    T_ATTN_NORM = 17, T_FFN_NORM = 18,
};

Tensor* weight_of(RefModel* m, int layer, int id) {
    switch (id) {
        case T_EMBED: return &m->emb().weight;
        case T_FINAL_NORM: return &m->fnorm().weight;
        case T_LM_HEAD: return &m->head().weight;
        default: break;
    }
    auto& b = m->block(layer);
    switch (id) {
        case T_Q: return &b.attn.query.weight;
        case T_K: return &b.attn.key.weight;
        case T_V: return &b.attn.value.weight;
        case T_O: return &b.attn.qkv_proj.weight;      // member `qkv_proj` holds o_proj (tinyllama.cpp:362-364)
        case T_GATE: return &b.ffn_gate_proj.weight;
        case T_UP: return &b.ffn_up_proj.weight;
        case T_DOWN: return &b.ffn_down_proj.weight;
        case T_ATTN_NORM: return &b.attn_norm.weight;
        case T_FFN_NORM: return &b.ffn_norm.weight;
    }
    return nullptr;
}

// Activation ids: the 13 rounding points of SURVEY.md App. A, per layer.
enum AcvId {
    A_EMB = 0, A_FINAL_NORM = 1,
    A_ATTN_NORM = 10, A_Q = 11, A_K = 12, A_V = 13, A_ATTN_OUT = 14, A_O = 15, A_INP_RES = 16,
    A_FFN_NORM = 17, A_GATE = 18, A_UP = 19, A_DOWN = 20, A_ATTN_RES = 21,
};

Tensor* acv_of(RefModel* m, int layer, int id) {
    switch (id) {
        case A_EMB: return &m->emb().emb_acv;
        case A_FINAL_NORM: return &m->fnorm().acv;
        default: break;
    }
    auto& b = m->block(layer);
    switch (id) {
        case A_ATTN_NORM: return &b.attn_norm.acv;
        case A_Q: return &b.attn.query.acv;
        case A_K: return &b.attn.key.acv;
        case A_V: return &b.attn.value.acv;
        case A_ATTN_OUT: return &b.attn.qkv_acv;
        case A_O: return &b.attn.qkv_proj.acv;
        case A_INP_RES: return &b.inp_res.acv;
        case A_FFN_NORM: return &b.ffn_norm.acv;
        case A_GATE: return &b.ffn_gate_proj.acv;   // after SiLU and Mul, both in place (modules.cpp:242-243)
        case A_UP: return &b.ffn_up_proj.acv;
        case A_DOWN: return &b.ffn_down_proj.acv;
        case A_ATTN_RES: return &b.attn_res.acv;
    }
    return nullptr;
}

}  // namespace

extern "C" {

const char* ref_build_info() {
    return "reference=/root/reference (unmodified) flags=-std=c++17 -O3 -fopenmp -mavx -mf16c"
#ifdef GTEN_SIMD_AVX
           " simd=avx+f16c"
#else
           " simd=scalar"
#endif
        ;
}

// ---------------------------------------------------------------- model ----
void* ref_model_new(int n_vocab, int n_embd, int n_ffn, int n_layers, int n_heads, int n_groups,
                    int max_ctx, int wdtype) {
    gten::ModuleDtype d;
    d.wdtype = dt(wdtype);
    d.adtype = (d.wdtype == gten::kFloat16) ? gten::kFloat16 : gten::kQint8;  // tinyllama.cpp:258-265
    auto* m = new RefModel();
    m->n_vocab = n_vocab;
    m->n_layers = n_layers;
    const TinyLLamaParams p{};
    if (n_vocab == p.n_vocab && n_embd == p.n_embd && n_ffn == p.n_ffn && n_layers == p.n_layers &&
        n_heads == p.n_heads && n_groups == p.n_query_groups) {
        m->full = std::make_unique<TinyLlama>(max_ctx, d);      // the reference's own class
    } else {
        m->mini = std::make_unique<MiniLlama>(n_vocab, n_embd, n_ffn, n_layers, n_heads, n_groups, max_ctx, d);
    }
    return m;
}

void ref_model_free(void* h) { delete static_cast<RefModel*>(h); }

int ref_model_is_reference_class(void* h) { return static_cast<RefModel*>(h)->full ? 1 : 0; }

// Returns the host pointer of a weight tensor (payload layout = .gten payload) and its byte size.
void* ref_model_weight(void* h, int layer, int id, int64_t* nbytes) {
    Tensor* t = weight_of(static_cast<RefModel*>(h), layer, id);
    if (!t) { *nbytes = 0; return nullptr; }
    *nbytes = (int64_t)t->nbytes();
    return t->data_ptr<void>();
}

// Loads a .gten checkpoint through the reference's own loader (tinyllama.cpp:336-392).
int ref_model_load_ckpt(void* h, const char* path) {
    auto* m = static_cast<RefModel*>(h);
    if (!m->full) return -1;
    std::ifstream f{path, std::ios::binary};
    if (!f.is_open()) return -2;
    m->full->load_from_ckpt(f);
    return 0;
}

// One TinyLlama::logits call (tinyllama.cpp:45-61). tokens = ALL tokens so far.
void ref_model_logits(void* h, const int32_t* tokens, int n_tokens, int start_pos, float* out) {
    auto* m = static_cast<RefModel*>(h);
    Tensor input{tokens, {n_tokens}, gten::kInt32};
    Tensor lg = m->logits(input, start_pos);
    std::memcpy(out, lg.data_ptr<float>(), sizeof(float) * m->n_vocab);
}

// The greedy loop of greedy_sample (tinyllama.cpp:395-440) without the tokenizer:
// first call is the prefill (start_pos 0), then one new row per call; argmax with
// strict '>' (lowest index wins).  No EOS stop so that lengths are deterministic.
// tokens: in/out buffer of capacity n_prompt + n_new.  times[0]=prefill s, times[1]=decode s.
// If logits_out != NULL it receives n_new * n_vocab floats.
void ref_model_generate(void* h, int32_t* tokens, int n_prompt, int n_new, double* times, float* logits_out) {
    auto* m = static_cast<RefModel*>(h);
    int n = n_prompt;
    double t_prefill = 0, t_decode = 0;
    for (int i = 0; i < n_new; i++) {
        Tensor input{tokens, {n}, gten::kInt32};
        const int start_pos = (i == 0) ? 0 : n - 1;
        auto t0 = std::chrono::steady_clock::now();
        Tensor lg = m->logits(input, start_pos);
        auto t1 = std::chrono::steady_clock::now();
        const double dtm = std::chrono::duration<double>(t1 - t0).count();
        if (i == 0) t_prefill += dtm; else t_decode += dtm;
        const float* d = lg.data_ptr<float>();
        if (logits_out) std::memcpy(logits_out + (size_t)i * m->n_vocab, d, sizeof(float) * m->n_vocab);
        float best = -std::numeric_limits<float>::infinity();
        int arg = 0;
        for (int j = 0; j < m->n_vocab; j++) if (d[j] > best) { best = d[j]; arg = j; }
        tokens[n++] = arg;
    }
    if (times) { times[0] = t_prefill; times[1] = t_decode; }
}

// Decode one row of a module activation buffer to fp32 (ops.h:40-70).  Returns row width.
int ref_model_acv(void* h, int layer, int id, int row, float* out) {
    Tensor* t = acv_of(static_cast<RefModel*>(h), layer, id);
    if (!t) return -1;
    const int w = t->dimsize(1);
    gten::ops::read_row_to_float(t->data_ptr<char>() + (size_t)row * t->bstride(0), t->dtype(), out, w);
    return w;
}

// Raw bytes of one row of a module activation buffer.  Returns byte count.
int ref_model_acv_raw(void* h, int layer, int id, int row, void* out) {
    Tensor* t = acv_of(static_cast<RefModel*>(h), layer, id);
    if (!t) return -1;
    const int nb = t->bstride(0);
    std::memcpy(out, t->data_ptr<char>() + (size_t)row * nb, nb);
    return nb;
}

int64_t ref_tensor_mem_allocated() { return gten::G_TensorMemAllocated; }

// ------------------------------------------------------------ scalar/rows --
uint16_t ref_fp32_to_fp16(float f) { return gten::fp32_to_fp16(f); }
float ref_fp16_to_fp32(uint16_t h) { return gten::fp16_to_fp32(h); }

void ref_q8_quantize_row(const float* inp, void* out, int n) {
    gten::ops::q8_quantize_row(inp, static_cast<gten::Q8Block*>(out), n);
}
void ref_q8_dequantize_row(const void* inp, float* out, int n) {
    gten::ops::q8_dequantize_row(static_cast<const gten::Q8Block*>(inp), out, n);
}
void ref_q4_dequantize_row(const void* inp, float* out, int n) {
    gten::ops::q4_dequantize_row(static_cast<const gten::Q4Block*>(inp), out, n);
}
void ref_read_row_to_float(const void* inp, int dtype, float* out, int n) {
    gten::ops::read_row_to_float(static_cast<const char*>(inp), dt(dtype), out, n);
}
void ref_write_row_from_float(float* inp, void* out, int dtype, int n) {
    gten::ops::write_row_from_float(inp, static_cast<char*>(out), dt(dtype), n);
}
float ref_vec_dot_product(const void* a, int adt, const void* b, int bdt, int n) {
    return gten::ops::vec_dot_product(static_cast<const char*>(a), dt(adt), static_cast<const char*>(b), dt(bdt), n);
}

// -------------------------------------------------------------------- ops --
void ref_token_embed(const void* w, int wdt, int n_vocab, int n_embd, const int32_t* tokens, int n_ctx,
                     void* out, int odt, int start_pos) {
    Tensor wt{w, {n_vocab, n_embd}, dt(wdt)};
    Tensor tk{tokens, {n_ctx}, gten::kInt32};
    Tensor o{out, {n_ctx, n_embd}, dt(odt)};
    gten::ops::token_embed(wt, tk, o, start_pos);
}

void ref_matmul_2d(const void* x, int xdt, int n_ctx, int k, const void* w, int wdt, int n_out,
                   void* out, int odt, int out_1d, int start_pos) {
    Tensor xt{x, {n_ctx, k}, dt(xdt)};
    Tensor wt{w, {n_out, k}, dt(wdt)};
    if (out_1d) {
        Tensor o{out, {n_out}, dt(odt)};
        o.set_strides({0});                       // modules.cpp:75 (last-row-only lm_head)
        gten::ops::matmul_2d(xt, wt, o, start_pos);
    } else {
        Tensor o{out, {n_ctx, n_out}, dt(odt)};
        gten::ops::matmul_2d(xt, wt, o, start_pos);
    }
}

void ref_rms_norm(const void* x, int xdt, int n_ctx, int n_embd, const uint16_t* w, void* out, int start_pos) {
    Tensor xt{x, {n_ctx, n_embd}, dt(xdt)};
    Tensor wt{w, {n_embd}, gten::kFloat16};
    Tensor o{out, {n_ctx, n_embd}, dt(xdt)};
    gten::ops::rms_norm(xt, wt, o, start_pos);
}

void ref_rotary_emb(void* x, int xdt, int n_ctx, int n_embd, int d_head, int start_pos) {
    Tensor xt{x, {n_ctx, n_embd}, dt(xdt)};
    gten::ops::rotary_emb(xt, d_head, start_pos);
}

void ref_silu(const void* x, int xdt, int n_ctx, int n_embd, void* out, int start_pos) {
    Tensor xt{x, {n_ctx, n_embd}, dt(xdt)};
    Tensor o{out, {n_ctx, n_embd}, dt(xdt)};
    gten::ops::silu(xt, o, start_pos);
}

void ref_mul(const void* a, const void* b, int xdt, int n_ctx, int n_embd, void* out, int start_pos) {
    Tensor at{a, {n_ctx, n_embd}, dt(xdt)};
    Tensor bt{b, {n_ctx, n_embd}, dt(xdt)};
    Tensor o{out, {n_ctx, n_embd}, dt(xdt)};
    gten::ops::mul(at, bt, o, start_pos);
}

void ref_add(const void* a, const void* b, int xdt, int n_ctx, int n_embd, void* out, int start_pos) {
    Tensor at{a, {n_ctx, n_embd}, dt(xdt)};
    Tensor bt{b, {n_ctx, n_embd}, dt(xdt)};
    Tensor o{out, {n_ctx, n_embd}, dt(xdt)};
    gten::ops::add(at, bt, o, start_pos);
}

// qk: scratch of max_ctx*max_ctx*n_heads elements worth of storage in dtype xdt (caller sizes it like
// SelfAttention does, modules.cpp:180); out: [n_ctx, n_heads*d_head].
void ref_qkv_attn(const void* q, const void* k, const void* v, void* qk, void* out, int xdt,
                  int n_ctx, int n_heads, int n_kv_heads, int d_head, int max_ctx, int start_pos) {
    const int n_embd = n_heads * d_head;
    const int kv_dim = n_kv_heads * d_head;
    Tensor qt{q, {n_ctx, n_embd}, dt(xdt)};
    Tensor kt{k, {n_ctx, kv_dim}, dt(xdt)};
    Tensor vt{v, {n_ctx, kv_dim}, dt(xdt)};
    Tensor qkt{qk, {n_heads, n_ctx, n_ctx}, dt(xdt)};
    Tensor o{out, {n_ctx, n_embd}, dt(xdt)};
    gten::ops::qkv_attn(qt, kt, vt, qkt, o, max_ctx, start_pos);
}

// libm entry points exactly as the reference TU calls them (std::exp/pow/cos/sin on float).
float ref_expf(float x) { return std::exp(x); }
void ref_rope_angles(int pos, int d_head, float* cos_out, float* sin_out) {
    // the expressions of ops.h:728-746
    const float d = static_cast<float>(d_head);
    const float m = static_cast<float>(pos);
    for (int j = 0; j < d_head / 2; ++j) {
        const float m_theta_i = m * std::pow(10000.0f, -(2.0f * j / d));
        cos_out[j] = std::cos(m_theta_i);
        sin_out[j] = std::sin(m_theta_i);
    }
}

}  // extern "C"
