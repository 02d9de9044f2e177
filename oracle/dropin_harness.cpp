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
