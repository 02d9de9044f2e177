// gten/quants.h -- the Q8 / Q4 block formats and their host-side codecs (reference gten/quants.h:17-150).
// These are the host restatements the loader and tools use; the hot path encodes/decodes on the device.
#pragma once
#include <cmath>

#include "gten_types.h"

namespace gten {

namespace globs {
static const int q8_block_size = 32;
static const int q4_block_size = 32;
}  // namespace globs

struct Q8Block { Float16 delta; Qint8 data[32]; };       // value = data * fp32(delta)
struct Q4Block { Float16 delta; Qint4 data[16]; };       // byte i = (elt i + 7) << 4 | (elt i+16 + 7); value = (nibble - 7) * fp32(delta)
static_assert(sizeof(Q8Block) == 34, "Q8Block is 34 bytes");
static_assert(sizeof(Q4Block) == 18, "Q4Block is 18 bytes");

namespace ops {          // the codecs live in gten::ops like the reference's (gten/quants.h:34-152)

[[nodiscard]] inline Qint8 q8_quantize_single(float x, float delta) {
    const float scale = delta ? 1.0f / delta : 0.0f;
    return static_cast<Qint8>(roundf(x * scale));
}
[[nodiscard]] inline float q8_dequantize_single(Qint8 x, float delta) { return x * delta; }

// one block of `block_size` <= 32 values: delta = absmax / 127 (stored fp16), codes from the UNROUNDED delta, roundf
inline void q8_quantize_block(const float* inp, Q8Block* out, const int block_size) {
    float absmax = 0.0f;
    for (int i = 0; i < block_size; i++) absmax = std::fmax(absmax, std::fabs(inp[i]));
    const float delta = absmax / 127.0f;
    out->delta = fp32_to_fp16(delta);
    const float scale = delta ? 1.0f / delta : 0.0f;
    for (int i = 0; i < block_size; i++) out->data[i] = static_cast<Qint8>(roundf(inp[i] * scale));
}
inline void q8_dequantize_block(const Q8Block* inp, float* out, const int block_size) {
    const float delta = fp16_to_fp32(inp->delta);
    for (int i = 0; i < block_size; i++) out[i] = inp->data[i] * delta;
}
inline void q4_dequantize_block(const Q4Block* inp, float* out) {
    const float delta = fp16_to_fp32(inp->delta);
    for (int i = 0; i < 16; i++) {
        out[i] = (static_cast<int>(inp->data[i] >> 4) - 7) * delta;
        out[i + 16] = (static_cast<int>(inp->data[i] & 0x0f) - 7) * delta;
    }
}
inline void q8_quantize_row(const float* inp, Q8Block* out, const int rowsize) {
    const int nfull = rowsize / 32, tail = rowsize % 32;
    for (int b = 0; b < nfull; b++) q8_quantize_block(inp + b * 32, out + b, 32);
    if (tail) q8_quantize_block(inp + nfull * 32, out + nfull, tail);
}
inline void q8_quantize_row_delta(const float* inp, Qint8* out, const float delta, const int rowsize) {
    for (int i = 0; i < rowsize; i++) out[i] = q8_quantize_single(inp[i], delta);
}
inline void q8_dequantize_row(const Q8Block* inp, float* out, int rowsize) {
    const int nfull = rowsize / 32, tail = rowsize % 32;
    for (int b = 0; b < nfull; b++) q8_dequantize_block(inp + b, out + b * 32, 32);
    if (tail) q8_dequantize_block(inp + nfull, out + nfull * 32, tail);
}
inline void q4_dequantize_row(const Q4Block* inp, float* out, int rowsize) {
    GTEN_ASSERT(rowsize % 32 == 0);
    for (int b = 0; b < rowsize / 32; b++) q4_dequantize_block(inp + b, out + b * 32);
}
inline void q8_dequantize_row_delta(const Qint8* x, float* out, float delta, int size) {
    for (int i = 0; i < size; i++) out[i] = x[i] * delta;
}

}  // namespace ops

// round 1 exposed the codecs in gten:: ; both spellings stay valid
using ops::q8_quantize_single; using ops::q8_dequantize_single; using ops::q8_quantize_block; using ops::q8_dequantize_block;
using ops::q4_dequantize_block; using ops::q8_quantize_row; using ops::q8_quantize_row_delta; using ops::q8_dequantize_row;
using ops::q4_dequantize_row; using ops::q8_dequantize_row_delta;

}  // namespace gten
