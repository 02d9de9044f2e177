// oracle/dropin_harness.cpp -- TEST INFRASTRUCTURE ONLY (SURVEY.md §8b acceptance test for "drop-in").
//
// Compiles the UNMODIFIED reference application source (tinyllama.cpp: the TinyLlama class, load_from_ckpt, greedy_sample,
// print_perf) against THIS repo's include/gten headers instead of the reference's gten/ directory, and exposes a small C
// entry point that drives TinyLlama::logits with the reference's own per-token protocol.  Built only where /root/reference
// exists (oracle/Makefile copies the two application files to a scratch directory outside the repo so that the quoted
// include "gten/gten.h" resolves to include/gten/gten.h); the resulting library travels to the GPU box in oracle/_ref/.
#include <algorithm>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <random>
#include <string>
#include <string_view>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#define main tinyllama_reference_main
#include "tinyllama.cpp"
#undef main

extern "C" int dropin_generate(const char* gten_path, int wdtype, int max_ctx, const int* prompt, int n_prompt, int n_new,
                               int* out_tokens, float* first_logits) {
    if (gtb_init(0) != 0) { std::fprintf(stderr, "%s\n", gtb_last_error()); return 1; }
    ModuleDtype dtype;
    if (wdtype == (int)kFloat16) dtype = {kFloat16, kFloat16};
    else if (wdtype == (int)kQint8) dtype = {kQint8, kQint8};
    else dtype = {kQint4, kQint8};                                     // tinyllama.cpp:258-265
    std::ifstream ckpt(gten_path, std::ios::binary);
    if (!ckpt.is_open()) return 2;
    TinyLlama model{max_ctx, dtype};
    model.load_from_ckpt(ckpt);
    std::vector<int> tokens(prompt, prompt + n_prompt);
    tokens.reserve(n_prompt + n_new);
    for (int i = 0; i < n_new; i++) {                                  // the loop of greedy_sample (tinyllama.cpp:402-434)
        Tensor input{tokens.data(), {(int)tokens.size()}, kInt32};
        const int start_pos = (i == 0) ? 0 : input.numel() - 1;
        Tensor logits = model.logits(input, start_pos);
        const float* p = logits.data_ptr<float>();
        if (i == 0 && first_logits) std::memcpy(first_logits, p, sizeof(float) * logits.numel());
        float best = -std::numeric_limits<float>::infinity();
        int arg = 0;
        for (int j = 0; j < logits.numel(); j++) if (p[j] > best) { best = p[j]; arg = j; }
        tokens.push_back(arg);
        out_tokens[i] = arg;
    }
    return 0;
}

// The reference's OWN greedy_sample (tinyllama.cpp:395-440) and print_perf (:515-582), unmodified, over this repo's headers:
// Tokenizer built from the reference's tokenizer.bin, text prompt in, generated text (stderr) and the perf table (stdout) out.
// profile != 0 switches gten::Timer to its profiling mode (include/gten/modules.h) so that exec_time is device time.
extern "C" int dropin_greedy_sample(const char* gten_path, int wdtype, int extra_tokens, const char* tokenizer_path, const char* prompt_text,
                                    int profile, char* out_text, int out_cap, int* prompt_ids, int* n_prompt_ids) {
    if (gtb_init(0) != 0) { std::fprintf(stderr, "%s\n", gtb_last_error()); return 1; }
    ModuleDtype dtype;
    if (wdtype == (int)kFloat16) dtype = {kFloat16, kFloat16};
    else if (wdtype == (int)kQint8) dtype = {kQint8, kQint8};
    else dtype = {kQint4, kQint8};
    std::ifstream ckpt(gten_path, std::ios::binary);
    if (!ckpt.is_open()) return 2;
    Tokenizer tokenizer{tokenizer_path, 32000};
    std::string prompt{prompt_text};
    std::string prompt_copy{prompt_text};                              // Tokenizer::encode edits its argument (tokenizer.h:150)
    std::vector<int> ids = tokenizer.encode(prompt_copy);
    *n_prompt_ids = (int)ids.size();
    for (size_t i = 0; i < ids.size(); i++) prompt_ids[i] = ids[i];
    const int n_predict = (int)ids.size() + extra_tokens;             // main(): max_ctx = n_predict (tinyllama.cpp:267)
    TinyLlama model{n_predict, dtype};
    model.load_from_ckpt(ckpt);
    gten::set_profiling(profile != 0);
    char tmpl[] = "/tmp/dropin_out_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd < 0) return 3;
    std::cout.flush(); std::cerr.flush(); fflush(stdout); fflush(stderr);
    const int so = dup(1), se = dup(2);
    dup2(fd, 1); dup2(fd, 2);
    greedy_sample(prompt, model, tokenizer, n_predict);
    std::cout.flush(); std::cerr.flush(); fflush(stdout); fflush(stderr);
    dup2(so, 1); dup2(se, 2); close(so); close(se);
    gten::set_profiling(false);
    const off_t n = lseek(fd, 0, SEEK_END);
    lseek(fd, 0, SEEK_SET);
    const ssize_t got = read(fd, out_text, (size_t)std::min<off_t>(n, out_cap - 1));
    out_text[got > 0 ? got : 0] = 0;
    close(fd); unlink(tmpl);
    return 0;
}

// the text greedy_sample prints for a given token sequence: decode(prev, tok) piece by piece, prev = 1 (BOS) for the first piece
extern "C" int dropin_decode(const char* tokenizer_path, const int* tokens, int n_prompt, int n_total, char* out, int cap) {
    Tokenizer tokenizer{tokenizer_path, 32000};
    std::string s;
    for (int i = n_prompt; i < n_total; i++) {
        if (tokens[i] == tokenizer.eos) break;
        s += tokenizer.decode(i == n_prompt ? 1 : tokens[i - 1], tokens[i]);
    }
    std::strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
    return (int)s.size();
}
