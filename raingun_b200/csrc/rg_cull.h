// rg_cull.h — constants of the conservative FP32 sphere cull, shared by the host (which
// builds the per-sphere records at scene creation) and the device (rg_trace.cuh).
//
// The reference's sphere test (bodies.rs:92-99) rejects when
//     opp2 = h.h - (h.d)^2  >  r^2,      h = c - o          (all FP64)
// With P a fixed reference point, c' = c - P, o' = o - P:
//     G := opp2 - r^2 = (|c'|^2 - r^2) + c'.(-2 o') - (c'.d - o'.d)^2 + |o'|^2
// The cull evaluates, in FP32 with FMAs,
//     s = c'.d - od            (3 FFMA, od = o'.d rounded once from FP64)
//     q = K + c'.o2            (3 FFMA, K = |c'|^2 - r^2, o2 = -2 o', both rounded once from FP64)
//     g = q - s*s              (1 FFMA)
// and rejects the sphere iff  g > thr,  thr = -|o'|^2 + margin.   8 instructions per test.
//
// Soundness (DESIGN.md "FP32 cull"): with u = 2^-24, C = |c'|, O = |o'|, |d|^2 <= 1 + 1e-6,
//     |g_fp32 + |o'|^2 - G|  <=  u * (18 (C + O)^2 + 5 r^2)  <=  u * (36 C^2 + 36 O^2 + 5 r^2)
// The margin used is kCullSafety (1.75x) larger still, split into a per-sphere part folded
// into K and a per-ray part folded into thr:
//     m_s = u (64 C^2 + 9 r^2),    m_r = u * 64 O^2
// Rays whose direction is not unit length within 1e-6 (reflections off un-normalised
// plane/disk normals, bodies.rs:151-153) get thr = +inf: nothing is culled for them.
// A rejected sphere therefore always fails the reference's own FP64 test, so culling can
// never change an image; RG_OPT_VERIFY_CULL counts violations on the device (must be 0).
#pragma once

namespace rg {
constexpr double kCullU = 5.9604644775390625e-8;   // 2^-24
constexpr double kCullSphereC2 = 64.0;             // m_s = u (64 C^2 + 9 r^2)
constexpr double kCullSphereR2 = 9.0;
constexpr double kCullRayO2 = 64.0;                // m_r = u * 64 O^2
constexpr double kCullUnitTol = 1e-6;              // | |d|^2 - 1 | above this: no culling
constexpr double kCullHuge = 1e30;                 // beyond this magnitude: no culling
}  // namespace rg
