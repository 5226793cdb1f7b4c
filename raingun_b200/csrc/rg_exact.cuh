// rg_exact.cuh — the reference's per-ray arithmetic as device functions (FP64 geometry,
// FP32 colour, reference operation order; compiled with -fmad=false, see rg_math.cuh).
// Each function names the reference lines it reproduces.  Shared by every pipeline.
#pragma once
#include "rg_math.cuh"
#include "rg_scene.cuh"

namespace rg {

struct Ray { D3 o, d; };   // ray.rs:5-20; inverted_direction / signs are derived on use (AABB only)

constexpr uint32_t kNoBody = 0xFFFFFFFFu;

// ---- intersections ----------------------------------------------------------------------
// bodies.rs:76-120.  Assumes nothing about |d| (the reference never renormalises).
// Written as straight-line selects with ONE return value on purpose: with the reference's
// chain of early returns, nvcc 12.9 (NVVM) dropped the `distance0 < 0 && distance1 < 0 ->
// None` case in some inlined copies (the hit flag became a constant 1 and t kept its
// caller-side initial value), which the parity tests caught as spurious t = 0 self-hits.
__device__ __forceinline__ bool sphere_intersect(double cx, double cy, double cz, double radius,
                                                 const Ray &ray, double &t) {
    D3 hyp = d3(cx, cy, cz) - ray.o;
    double adj = dot(hyp, ray.d);
    double opp2 = dot(hyp, hyp) - (adj * adj);
    double r2 = radius * radius;
    if (opp2 > r2) return false;
    double thickness = sqrt(r2 - opp2);
    double d0 = adj - thickness;
    double d1 = adj + thickness;
    const bool n0 = d0 < 0.0, n1 = d1 < 0.0;
    // both < 0 -> None; one < 0 -> the other; else min (f64::min == fmin)
    t = n0 ? d1 : (n1 ? d0 : fmin(d0, d1));
    return !(n0 && n1);
}
// bodies.rs:136-149
__device__ __forceinline__ bool plane_intersect(const double *g, const Ray &ray, double &t) {
    D3 origin = d3(g[0], g[1], g[2]), normal = d3(g[3], g[4], g[5]);
    double den = dot(normal, ray.d);
    if (den > 1e-6) {
        D3 v = origin - ray.o;
        double dist = dot(v, normal) / den;
        if (dist >= 0.0) { t = dist; return true; }
    }
    return false;
}
// bodies.rs:173-192
__device__ __forceinline__ bool disk_intersect(const double *g, const Ray &ray, double &t) {
    D3 origin = d3(g[0], g[1], g[2]), normal = d3(g[3], g[4], g[5]);
    double den = dot(normal, ray.d);
    if (den > 1e-6) {
        D3 v = origin - ray.o;
        double dist = dot(v, normal) / den;
        if (dist >= 0.0) {
            D3 hp = ray.o + ray.d * dist;
            D3 w = hp - origin;
            double d2 = dot(w, w);
            if (sqrt(d2) < g[6]) { t = dist; return true; }
        }
    }
    return false;
}
// bodies.rs:242-282 with Ray::new's inverted_direction / signs (ray.rs:23-35)
__device__ __forceinline__ bool aabb_intersect(const double *g, const Ray &ray, double &t) {
    double ix = 1.0 / ray.d.x, iy = 1.0 / ray.d.y, iz = 1.0 / ray.d.z;
    bool sx = ix < 0.0, sy = iy < 0.0, sz = iz < 0.0;
    double tmin = ((sx ? g[3] : g[0]) - ray.o.x) * ix;
    double tmax = ((sx ? g[0] : g[3]) - ray.o.x) * ix;
    double tymin = ((sy ? g[4] : g[1]) - ray.o.y) * iy;
    double tymax = ((sy ? g[1] : g[4]) - ray.o.y) * iy;
    if (tmin > tymax || tymin > tmax) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    double tzmin = ((sz ? g[5] : g[2]) - ray.o.z) * iz;
    double tzmax = ((sz ? g[2] : g[5]) - ray.o.z) * iz;
    if (tmin > tzmax || tzmin > tmax) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    if (tmin >= 0.0) { t = tmin; return true; }
    if (tmax >= 0.0) { t = tmax; return true; }
    return false;
}
// bodies.rs:336-345 for the non-sphere kinds
__device__ __forceinline__ bool misc_intersect(const DScene &s, uint32_t body, const Ray &ray, double &t) {
    const double *g = s.geom + 8 * (size_t)body;
    switch (s.kind[body]) {
        case RG_BODY_PLANE: return plane_intersect(g, ray, t);
        case RG_BODY_DISK: return disk_intersect(g, ray, t);
        case RG_BODY_AABB: return aabb_intersect(g, ray, t);
        default: return sphere_intersect(g[0], g[1], g[2], g[3], ray, t);
    }
}

// Scene::trace's `min_by` (scene.rs:34-39) keeps the FIRST of equal minima, i.e. the result is
// the lexicographic minimum of (distance, original body index) — an order-independent
// reduction, which is what lets the device test bodies in any order and in parallel.
struct Nearest {
    double t;
    uint32_t body;
    __device__ __forceinline__ void init() { t = 0.0; body = kNoBody; }
    __device__ __forceinline__ bool found() const { return body != kNoBody; }
    __device__ __forceinline__ void offer(double ct, uint32_t cbody) {
        if (body == kNoBody || ct < t || (ct == t && cbody < body)) { t = ct; body = cbody; }
    }
};

// Exact brute-force trace: the reference algorithm verbatim (every body, FP64).  Used by the
// megakernel (validation pipeline) and as the in-kernel ground truth of RG_OPT_VERIFY_CULL.
__device__ __forceinline__ Nearest trace_exact_all(const DScene &s, const Ray &ray, DCounters *ctr) {
    Nearest best;
    best.init();
    unsigned nan_count = 0;
    for (uint32_t i = 0; i < s.n_spheres; ++i) {
        double4 sp = s.sph[i];
        double t;
        if (sphere_intersect(sp.x, sp.y, sp.z, sp.w, ray, t)) {
            if (t != t) { ++nan_count; continue; }
            best.offer(t, s.sph_body[i]);
        }
    }
    for (uint32_t i = 0; i < s.n_misc; ++i) {
        uint32_t b = s.misc_body[i];
        double t;
        if (misc_intersect(s, b, ray, t)) {
            if (t != t) { ++nan_count; continue; }
            best.offer(t, b);
        }
    }
    if (nan_count) atomicAdd(&ctr->err_nan, (unsigned long long)nan_count);  // scene.rs:38 would panic
    return best;
}

// ---- surface data -----------------------------------------------------------------------
__device__ __forceinline__ bool is_close(double a, double b) { return fabs(a - b) < 1e-8; }  // bodies.rs:9-11

// bodies.rs:122-124, 151-153, 194-196, 284-328
__device__ __forceinline__ D3 surface_normal(const DScene &s, uint32_t body, D3 hp, DCounters *ctr) {
    const double *g = s.geom + 8 * (size_t)body;
    switch (s.kind[body]) {
        case RG_BODY_SPHERE: return normalize(hp - d3(g[0], g[1], g[2]));
        case RG_BODY_PLANE:
        case RG_BODY_DISK: return -d3(g[3], g[4], g[5]);
        default:
            if (is_close(hp.x, g[0])) return d3(-1.0, -0.0, -0.0);
            if (is_close(hp.x, g[3])) return d3(1.0, 0.0, 0.0);
            if (is_close(hp.y, g[1])) return d3(-0.0, -1.0, -0.0);
            if (is_close(hp.y, g[4])) return d3(0.0, 1.0, 0.0);
            if (is_close(hp.z, g[2])) return d3(-0.0, -0.0, -1.0);
            if (is_close(hp.z, g[5])) return d3(0.0, 0.0, 1.0);
            atomicAdd(&ctr->err_aabb, 1ull);   // bodies.rs:324 assert!(false)
            return d3(1.0, 0.0, 0.0);
    }
}

// bodies.rs:126-132, 155-169, 198-212, 330-333
__device__ __forceinline__ void texture_coords(const DScene &s, uint32_t body, D3 hp, float &u, float &v) {
    const double *g = s.geom + 8 * (size_t)body;
    switch (s.kind[body]) {
        case RG_BODY_SPHERE: {
            D3 hv = hp - d3(g[0], g[1], g[2]);
            u = (1.0f + ((float)atan2(hv.z, hv.x)) / kPiF32) * 0.5f;
            v = ((float)acos(hv.y / g[3])) / kPiF32;
            return;
        }
        case RG_BODY_PLANE:
        case RG_BODY_DISK: {
            D3 n = d3(g[3], g[4], g[5]);
            D3 xa = cross(n, d3(0.0, 0.0, 1.0));
            if (dot(xa, xa) == 0.0) xa = cross(n, d3(0.0, 1.0, 0.0));
            D3 ya = cross(n, xa);
            D3 hv = hp - d3(g[0], g[1], g[2]);
            u = (float)dot(hv, xa);
            v = (float)dot(hv, ya);
            return;
        }
        default: u = 0.0f; v = 0.0f; return;
    }
}

// material.rs:70-79 — truncation toward zero, then Rust's `%`, then +max if negative.
__device__ __forceinline__ uint32_t tex_wrap(float val, uint32_t max) {
    int32_t smax = (int32_t)max;
    float fc = val * (float)max;
    int32_t w = f32_as_i32(fc) % smax;
    return (uint32_t)(w < 0 ? w + smax : w);
}

// material.rs:56-68,82-89 + color.rs:26-30.  The texel is fetched through a CUDA texture
// object with point filtering at the integer texel the reference's `wrap` selects, so the
// "filtering" is the reference's: nearest texel, no interpolation, no mip levels.
__device__ __forceinline__ C3 body_color(const DScene &s, uint32_t body, float u, float v) {
    const BodyMat &m = s.mat[body];
    if (m.coloration == RG_COLORATION_COLOR) return c3(m.color[0], m.color[1], m.color[2]);
    DTex t = s.tex[m.tex];
    uint32_t x = tex_wrap(u + m.tex_off[0], t.w);
    uint32_t y = tex_wrap(v + m.tex_off[1], t.h);
    uchar4 p = tex2D<uchar4>(t.obj, (float)x + 0.5f, (float)y + 0.5f);
    return c3((float)p.x / 255.0f, (float)p.y / 255.0f, (float)p.z / 255.0f);
}

// body.color(&body.texture_coords(&hit_point)) (rendering.rs:137-138, :98).  Coloration::Color
// ignores the coordinates (material.rs:84-85), so they — and the atan2/acos behind the sphere's —
// are only evaluated for textured bodies; the colour returned is the same either way.
__device__ __forceinline__ C3 body_color_at(const DScene &s, uint32_t body, D3 hp) {
    const BodyMat &m = s.mat[body];
    if (m.coloration == RG_COLORATION_COLOR) return c3(m.color[0], m.color[1], m.color[2]);
    float u, v;
    texture_coords(s, body, hp, u, v);
    return body_color(s, body, u, v);
}

// ---- lights (lights.rs:28-58) -------------------------------------------------------------
__device__ __forceinline__ D3 light_direction_from(const DLight &l, D3 p) {
    if (l.kind == RG_LIGHT_DIRECTIONAL) return d3(l.v[0], l.v[1], l.v[2]);   // hoisted normalize(-direction)
    return normalize(d3(l.v[0], l.v[1], l.v[2]) - p);
}
__device__ __forceinline__ double light_distance(const DLight &l, D3 p) {
    if (l.kind == RG_LIGHT_DIRECTIONAL) return __longlong_as_double(0x7FF0000000000000ll);
    D3 w = d3(l.v[0], l.v[1], l.v[2]) - p;
    return sqrt(dot(w, w));
}
__device__ __forceinline__ float light_intensity(const DLight &l, D3 p) {
    if (l.kind == RG_LIGHT_DIRECTIONAL) return l.intensity;
    D3 w = d3(l.v[0], l.v[1], l.v[2]) - p;
    float r2 = (float)dot(w, w);
    return l.intensity / ((4.0f * kPiF32) * r2);
}

// ---- secondary rays -----------------------------------------------------------------------
// rendering.rs:174-200
__device__ __forceinline__ double fresnel(D3 incident, D3 normal, float index) {
    double idn = dot(incident, normal);
    double eta_i, eta_t;
    if (idn > 0.0) { eta_i = (double)index; eta_t = 1.0; }
    else { eta_i = 1.0; eta_t = (double)index; }
    double sin_t = eta_i / eta_t * sqrt(fmax(1.0 - idn * idn, 0.0));
    if (sin_t > 1.0) return 1.0;
    double cos_t = sqrt(fmax(1.0 - sin_t * sin_t, 0.0));
    double cos_i = fabs(cos_t);
    double r_s = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    double r_p = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (r_s * r_s + r_p * r_p) / 2.0;
}
// ray.rs:56-60
__device__ __forceinline__ Ray create_reflection(D3 normal, D3 incident, D3 p) {
    Ray r;
    r.o = p + normal * kShadowBias;
    double k = 2.0 * dot(incident, normal);
    r.d = incident - normal * k;
    return r;
}
// ray.rs:62-94
__device__ __forceinline__ bool create_transmission(D3 normal, D3 incident, D3 p, double bias, float index, Ray &out) {
    D3 ref_n = normal;
    double eta_t = (double)index, eta_i = 1.0;
    double idn = dot(incident, normal);
    if (idn < 0.0) idn = -idn;
    else { ref_n = -normal; eta_t = 1.0; eta_i = (double)index; }
    double eta = eta_i / eta_t;
    double k = 1.0 - (eta * eta) * (1.0 - idn * idn);
    if (k < 0.0) return false;
    out.o = p + ref_n * (-bias);
    out.d = (incident + ref_n * idn) * eta - ref_n * sqrt(k);
    return true;
}
// ray.rs:37-54 (origin at the camera; `fov_adj` hoisted)
__device__ __forceinline__ Ray create_prime(const DScene &s, uint32_t x, uint32_t y, uint32_t width, uint32_t height) {
    double aspect = (double)width / (double)height;
    double sx = ((((double)x + 0.5) / (double)width) * 2.0 - 1.0) * aspect * s.fov_adj;
    double sy = (1.0 - (((double)y + 0.5) / (double)height) * 2.0) * s.fov_adj;
    Ray r;
    r.o = d3(0.0, 0.0, 0.0);
    r.d = normalize(d3(sx, sy, -1.0));
    return r;
}

// One light's term of shade_diffuse (rendering.rs:157-169).  `a` = max(n.L as f32, 0),
// `intensity` = light.intensity(hit_point) — both computed where the hit is shaded.
__device__ __forceinline__ C3 light_term(C3 body_col, const DLight &l, float albedo, float a, float intensity, bool in_light) {
    float li = in_light ? intensity : 0.0f;
    float power = a * li;
    float reflected = albedo / kPiF32;
    C3 lc = (c3(l.color[0], l.color[1], l.color[2]) * power) * reflected;
    return body_col * lc;
}

}  // namespace rg
