// rg_math.cuh — exact arithmetic building blocks of the device render path.
//
// Precision contract: geometry is IEEE f64, colour is
// IEEE f32, every operation is a single rounding in the reference's operation order.
// This translation unit is compiled with -fmad=false, so `a * b + c` below is a DMUL/FMUL
// followed by a DADD/FADD, never an FMA; f64 and f32 `/` and sqrt are the IEEE-correct
// CUDA defaults (no --use_fast_math).  The only FMAs in the library are the explicit
// fmaf() calls of the conservative FP32 cull (rg_trace.cuh), whose result never reaches
// the image.
//
// cgmath 0.13.0 semantics restated (Cargo.lock:97-98; not vendored in the reference):
//   dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z        normalize(v) = v * (1.0 / sqrt(dot(v,v)))
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rg {

constexpr double kShadowBias = 1e-13;                                   // lib.rs:11
constexpr float kPiF32 = 3.14159265358979323846264338327950288f;        // std::f32::consts::PI

struct D3 { double x, y, z; };   // Point3 / Vector3 (lib.rs:29-30)
struct C3 { float r, g, b; };    // Color (color.rs:7-11)

__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 operator-(D3 a) { return d3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ double dot(D3 a, D3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) {
    return d3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
__device__ __forceinline__ D3 normalize(D3 v) { return v * (1.0 / sqrt(dot(v, v))); }

__device__ __forceinline__ C3 c3(float r, float g, float b) { C3 c; c.r = r; c.g = g; c.b = b; return c; }
__device__ __forceinline__ C3 operator+(C3 a, C3 b) { return c3(a.r + b.r, a.g + b.g, a.b + b.b); }   // color.rs:62-70
__device__ __forceinline__ C3 operator*(C3 a, C3 b) { return c3(a.r * b.r, a.g * b.g, a.b * b.b); }   // color.rs:72-80
__device__ __forceinline__ C3 operator*(C3 a, float s) { return c3(a.r * s, a.g * s, a.b * s); }      // color.rs:98-104
// color.rs:39-43 — f32::min / f32::max return the non-NaN operand, as fminf / fmaxf do.
__device__ __forceinline__ C3 clamp01(C3 a) {
    return c3(fmaxf(fminf(a.r, 1.0f), 0.0f), fmaxf(fminf(a.g, 1.0f), 0.0f), fmaxf(fminf(a.b, 1.0f), 0.0f));
}
// Rust `as u8` from f32: truncate toward zero, saturate, NaN -> 0 (color.rs:32-37).
__device__ __forceinline__ uint32_t f32_as_u8(float v) {
    if (!(v == v) || v <= 0.0f) return 0u;
    if (v >= 255.0f) return 255u;
    return (uint32_t)__float2int_rz(v);
}
__device__ __forceinline__ uchar4 quantise(C3 c) {
    return make_uchar4((unsigned char)f32_as_u8(c.r * 255.0f), (unsigned char)f32_as_u8(c.g * 255.0f),
                       (unsigned char)f32_as_u8(c.b * 255.0f), 255);
}
// Rust `as i32` from f32: truncate toward zero, saturate, NaN -> 0 — cvt.rzi.s32.f32 does exactly that.
__device__ __forceinline__ int32_t f32_as_i32(float v) { return __float2int_rz(v); }

}  // namespace rg
