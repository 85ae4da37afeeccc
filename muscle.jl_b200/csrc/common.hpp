// Shared host-side definitions for libmuscle_b200 (status codes, dtype helpers, error string).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/muscle_b200.h"

namespace mb200 {

inline bool dtype_valid(int d) { return d >= MB200_F32 && d <= MB200_C128; }
inline bool dtype_is_complex(int d) { return d == MB200_C64 || d == MB200_C128; }
inline bool dtype_is_double(int d) { return d == MB200_F64 || d == MB200_C128; }
inline size_t dtype_size(int d) {
    switch (d) {
        case MB200_F32: return 4;
        case MB200_F64: return 8;
        case MB200_C64: return 8;
        default: return 16;
    }
}
// Base.promote_eltype on {Float32, Float64, ComplexF32, ComplexF64}
inline int dtype_promote(int a, int b) {
    bool c = dtype_is_complex(a) || dtype_is_complex(b);
    bool d = dtype_is_double(a) || dtype_is_double(b);
    return c ? (d ? MB200_C128 : MB200_C64) : (d ? MB200_F64 : MB200_F32);
}
inline const char *dtype_name(int d) {
    switch (d) {
        case MB200_F32: return "Float32";
        case MB200_F64: return "Float64";
        case MB200_C64: return "ComplexF32";
        case MB200_C128: return "ComplexF64";
        default: return "?";
    }
}

// thread-local last error message (mb200_last_error_string)
std::string &last_error();
int fail(int status, const char *fmt, ...);

}  // namespace mb200
