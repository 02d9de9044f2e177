// gten/log.h -- the reference's error convention (gten/log.h:6-23): print to stderr, exit(EXIT_FAILURE).
#pragma once
#include <cstdio>
#include <cstdlib>

#define GTEN_ASSERT(condition)                                                                                        \
    do {                                                                                                              \
        if (!(condition)) {                                                                                           \
            std::fprintf(stderr, "\n\x1B[1;31mGTEN ERROR [File `%s` line %d]: Assertion '%s' failed.\n", __FILE__, __LINE__, #condition); \
            std::exit(EXIT_FAILURE);                                                                                  \
        }                                                                                                             \
    } while (0)

#define GTEN_ASSERTM(condition, message, ...)                                                  \
    do {                                                                                       \
        if (!(condition)) {                                                                    \
            std::fprintf(stderr, "\x1B[1;31m\nGTEN ERROR [File `%s` line %d]: ", __FILE__, __LINE__); \
            std::fprintf(stderr, message, ##__VA_ARGS__);                                      \
            std::fprintf(stderr, "\n");                                                        \
            std::exit(EXIT_FAILURE);                                                           \
        }                                                                                      \
    } while (0)

// a non-zero status from libgten_b200.so becomes the reference's assert-and-exit
#define GTEN_CUDA_OK(call) GTEN_ASSERTM((call) == 0, "%s: %s", #call, gtb_last_error())
