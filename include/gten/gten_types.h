// gten/gten_types.h -- scalar types and dtype tags of the gten API (reference gten/gten_types.h:15-33, 79-149).
#pragma once
#include <cstdint>
#include <cstring>

#include "log.h"

namespace gten {

typedef int32_t Int32;
typedef uint16_t Float16;     // raw IEEE binary16 bits
typedef int8_t Qint8;
typedef uint8_t Qint4;        // two 4-bit codes per byte

enum class Dtype { Int32, Float16, Float32, Qint8, Qint4 };     // numeric values == the GTB_* codes of gten_b200.h
static const Dtype kInt32 = Dtype::Int32;
static const Dtype kFloat16 = Dtype::Float16;
static const Dtype kFloat32 = Dtype::Float32;
static const Dtype kQint8 = Dtype::Qint8;
static const Dtype kQint4 = Dtype::Qint4;

inline const char* dtype_str(Dtype d) {
    switch (d) {
        case Dtype::Int32: return "Int32";
        case Dtype::Float16: return "Float16";
        case Dtype::Float32: return "Float32";
        case Dtype::Qint8: return "Qint8";
        case Dtype::Qint4: return "Qint4";
    }
    return "";
}

// binary16 -> binary32, exact
inline float fp16_to_fp32(Float16 h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {                                        // subnormal: renormalise
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

// binary32 -> binary16, round to nearest even, overflow -> inf, every NaN -> sign|0x7E00 (reference :99-119)
inline Float16 fp32_to_fp16(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7fffffffu;
    if (ax > 0x7f800000u) return (Float16)(sign | 0x7e00u);
    if (ax >= 0x47800000u) return (Float16)(sign | 0x7c00u);                    // >= 65536: inf (65520..65536 rounds to inf below)
    if (ax < 0x33000001u) return (Float16)sign;                                  // < 2^-25 (or == 2^-25: ties to even -> 0)
    int e = (int)(ax >> 23) - 127;
    uint32_t man = (ax & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }                           // subnormal result
    else { shift = 13; base = (uint32_t)(e + 15) << 10; man &= 0x7fffffu; }
    const uint32_t q = man >> shift, rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    uint32_t r = base + q;
    if (rem > half || (rem == half && (q & 1u))) r++;                            // carries propagate into the exponent (and to inf)
    return (Float16)(sign | r);
}

}  // namespace gten
