// raingun_oracle.cpp — CPU restatement of raingun's per-pixel render path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under raingun_b200/ may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / `--impl reference` legs use it, as the checker and as the
// timed CPU baseline.  The product path is the CUDA library.
//
// The reference (Rust, /root/reference) cannot be built here (no cargo/rustc,
// 139 un-vendored crates), so this is a literal restatement of its arithmetic,
// in its operation order, compiled with -ffp-contract=off so that every f64 and
// f32 operation is a single IEEE-754 rounding exactly as in the Rust binary.
// Un-vendored third-party arithmetic restated from its published source:
//   cgmath 0.13.0 (Cargo.lock:97-98)  dot = (x*x' + y*y') + z*z'
//                                     normalize(v) = v * (1.0 / sqrt(dot(v,v)))
//                                     cross, +, -, scalar *, 1.0 / Vector3
// PINNED by the reference's only artefacts for this path, examples/test{1,2,3}.png: all three
// renders are reproduced bit for bit (tests/test_oracle_golden.py).  test2.png has no textures
// and pins the arithmetic; test1.png / test3.png additionally pin the texels, which the hosts
// decode with raingun_b200/host/rgh_jpeg.cpp — a restatement of jpeg-decoder 0.1.11 (image
// 0.12.3's JPEG back end, Cargo.lock; not vendored) whose output this oracle only consumes.
//
// Every function cites the reference file:line it follows.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/raingun_b200.h"

namespace {

const double SHADOW_BIAS = 1e-13;          // lib.rs:11
const float PI_F32 = 3.14159265358979323846264338327950288f;  // std::f32::consts::PI

struct V3 { double x, y, z; };             // lib.rs:29-30 (cgmath f64)
struct Color { float r, g, b; };           // color.rs:7-11

// ---- cgmath 0.13 restated ---------------------------------------------------
inline V3 v_sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 v_add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 v_mul(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 v_neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline double v_dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 v_cross(V3 a, V3 b) {
    return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}
inline V3 v_normalize(V3 v) { return v_mul(v, 1.0 / std::sqrt(v_dot(v, v))); }

// ---- color.rs:39-43,62-112 ---------------------------------------------------
inline Color c_add(Color a, Color b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }
inline Color c_mul(Color a, Color b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }
inline Color c_scale(Color a, float s) { return {a.r * s, a.g * s, a.b * s}; }
inline Color c_clamp(Color a) {            // color.rs:39-43 (f32::min/max ignore NaN)
    return {fmaxf(fminf(a.r, 1.0f), 0.0f), fmaxf(fminf(a.g, 1.0f), 0.0f),
            fmaxf(fminf(a.b, 1.0f), 0.0f)};
}
// Rust `as u8` from f32: truncate toward zero, saturate, NaN -> 0.
inline uint8_t f32_as_u8(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
// Rust `as i32` from f32: truncate toward zero, saturate, NaN -> 0.
inline int32_t f32_as_i32(float v) {
    if (!(v == v)) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}

struct Ray {                                // ray.rs:5-20
    V3 origin, direction, inverted_direction;
    int signs[3];
};
inline Ray ray_new(V3 origin, V3 direction) {  // ray.rs:23-35
    Ray r;
    r.origin = origin;
    r.direction = direction;
    r.inverted_direction = {1.0 / direction.x, 1.0 / direction.y, 1.0 / direction.z};
    r.signs[0] = r.inverted_direction.x < 0.0 ? 1 : 0;
    r.signs[1] = r.inverted_direction.y < 0.0 ? 1 : 0;
    r.signs[2] = r.inverted_direction.z < 0.0 ? 1 : 0;
    return r;
}

struct Counters {
    uint64_t primary = 0, shadow = 0, reflection = 0, transmission = 0;
    uint64_t exact_tests = 0;
    uint64_t err_nan = 0, err_trans = 0, err_aabb = 0;
};

struct Scene {
    const rg_scene_desc *d;
    double fov_adjustment;                  // ray.rs:46 (pure function of the scene)
    uint32_t max_depth;
    Color default_color;
};

inline const double *geom(const Scene &s, uint32_t i) { return s.d->body_geom + 8 * (size_t)i; }

// bodies.rs:76-120
inline bool sphere_intersect(const double *g, const Ray &ray, double *out) {
    V3 center{g[0], g[1], g[2]};
    double radius = g[3];
    V3 hyp = v_sub(center, ray.origin);
    double adj = v_dot(hyp, ray.direction);
    double opp2 = v_dot(hyp, hyp) - (adj * adj);
    double r2 = radius * radius;
    if (opp2 > r2) return false;
    double thickness = std::sqrt(r2 - opp2);
    double d0 = adj - thickness;
    double d1 = adj + thickness;
    if (d0 < 0.0 && d1 < 0.0) return false;
    if (d0 < 0.0) { *out = d1; return true; }
    if (d1 < 0.0) { *out = d0; return true; }
    *out = std::fmin(d0, d1);
    return true;
}
// bodies.rs:136-149
inline bool plane_intersect(const double *g, const Ray &ray, double *out) {
    V3 origin{g[0], g[1], g[2]}, normal{g[3], g[4], g[5]};
    double den = v_dot(normal, ray.direction);
    if (den > 1e-6) {
        V3 v = v_sub(origin, ray.origin);
        double dist = v_dot(v, normal) / den;
        if (dist >= 0.0) { *out = dist; return true; }
    }
    return false;
}
// bodies.rs:173-192
inline bool disk_intersect(const double *g, const Ray &ray, double *out) {
    V3 origin{g[0], g[1], g[2]}, normal{g[3], g[4], g[5]};
    double radius = g[6];
    double den = v_dot(normal, ray.direction);
    if (den > 1e-6) {
        V3 v = v_sub(origin, ray.origin);
        double dist = v_dot(v, normal) / den;
        if (dist >= 0.0) {
            V3 hp = v_add(ray.origin, v_mul(ray.direction, dist));
            V3 w = v_sub(hp, origin);
            double d2 = v_dot(w, w);
            if (std::sqrt(d2) < radius) { *out = dist; return true; }
        }
    }
    return false;
}
// bodies.rs:242-282
inline bool aabb_intersect(const double *g, const Ray &ray, double *out) {
    const double *b0 = g, *b1 = g + 3;  // bounds[0], bounds[1]
    auto bx = [&](int s) { return s ? b1[0] : b0[0]; };
    auto by = [&](int s) { return s ? b1[1] : b0[1]; };
    auto bz = [&](int s) { return s ? b1[2] : b0[2]; };
    double tmin = (bx(ray.signs[0]) - ray.origin.x) * ray.inverted_direction.x;
    double tmax = (bx(1 - ray.signs[0]) - ray.origin.x) * ray.inverted_direction.x;
    double tymin = (by(ray.signs[1]) - ray.origin.y) * ray.inverted_direction.y;
    double tymax = (by(1 - ray.signs[1]) - ray.origin.y) * ray.inverted_direction.y;
    if (tmin > tymax || tymin > tmax) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    double tzmin = (bz(ray.signs[2]) - ray.origin.z) * ray.inverted_direction.z;
    double tzmax = (bz(1 - ray.signs[2]) - ray.origin.z) * ray.inverted_direction.z;
    if (tmin > tzmax || tzmin > tmax) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    if (tmin >= 0.0) { *out = tmin; return true; }
    if (tmax >= 0.0) { *out = tmax; return true; }
    return false;
}
// bodies.rs:336-345
inline bool body_intersect(const Scene &s, uint32_t i, const Ray &ray, double *out) {
    const double *g = geom(s, i);
    switch (s.d->body_kind[i]) {
        case RG_BODY_SPHERE: return sphere_intersect(g, ray, out);
        case RG_BODY_PLANE: return plane_intersect(g, ray, out);
        case RG_BODY_DISK: return disk_intersect(g, ray, out);
        default: return aabb_intersect(g, ray, out);
    }
}

// scene.rs:34-39 — min_by keeps the FIRST of equal minima; a NaN distance would
// panic in partial_cmp().unwrap(): counted, and the candidate is dropped.
inline bool trace(const Scene &s, const Ray &ray, double *t_out, uint32_t *body_out, Counters &c) {
    bool found = false;
    double best = 0.0;
    uint32_t best_i = 0;
    const uint32_t n = s.d->n_bodies;
    for (uint32_t i = 0; i < n; ++i) {
        double t;
        if (!body_intersect(s, i, ray, &t)) continue;
        if (t != t) { c.err_nan++; continue; }
        if (!found || t < best) { found = true; best = t; best_i = i; }
    }
    c.exact_tests += n;
    *t_out = best;
    *body_out = best_i;
    return found;
}

inline bool is_close(double a, double b) { return std::fabs(a - b) < 1e-8; }  // bodies.rs:9-11

// bodies.rs:122-124, 151-153, 194-196, 284-328
inline V3 surface_normal(const Scene &s, uint32_t i, V3 hp, Counters &c) {
    const double *g = geom(s, i);
    switch (s.d->body_kind[i]) {
        case RG_BODY_SPHERE: return v_normalize(v_sub(hp, V3{g[0], g[1], g[2]}));
        case RG_BODY_PLANE:
        case RG_BODY_DISK: return v_neg(V3{g[3], g[4], g[5]});
        default:
            if (is_close(hp.x, g[0])) return {-1.0, -0.0, -0.0};
            if (is_close(hp.x, g[3])) return {1.0, 0.0, 0.0};
            if (is_close(hp.y, g[1])) return {-0.0, -1.0, -0.0};
            if (is_close(hp.y, g[4])) return {0.0, 1.0, 0.0};
            if (is_close(hp.z, g[2])) return {-0.0, -0.0, -1.0};
            if (is_close(hp.z, g[5])) return {0.0, 0.0, 1.0};
            c.err_aabb++;
            return {1.0, 0.0, 0.0};
    }
}

// bodies.rs:126-132, 155-169, 198-212, 330-333
inline void texture_coords(const Scene &s, uint32_t i, V3 hp, float *u, float *v) {
    const double *g = geom(s, i);
    switch (s.d->body_kind[i]) {
        case RG_BODY_SPHERE: {
            V3 hv = v_sub(hp, V3{g[0], g[1], g[2]});
            *u = (1.0f + ((float)std::atan2(hv.z, hv.x)) / PI_F32) * 0.5f;
            *v = ((float)std::acos(hv.y / g[3])) / PI_F32;
            return;
        }
        case RG_BODY_PLANE:
        case RG_BODY_DISK: {
            V3 n{g[3], g[4], g[5]};
            V3 xa = v_cross(n, V3{0.0, 0.0, 1.0});
            if (v_dot(xa, xa) == 0.0) xa = v_cross(n, V3{0.0, 1.0, 0.0});
            V3 ya = v_cross(n, xa);
            V3 hv = v_sub(hp, V3{g[0], g[1], g[2]});
            *u = (float)v_dot(hv, xa);
            *v = (float)v_dot(hv, ya);
            return;
        }
        default: *u = 0.0f; *v = 0.0f; return;
    }
}

// material.rs:70-79
inline uint32_t tex_wrap(float val, uint32_t max) {
    int32_t smax = (int32_t)max;
    float fc = val * (float)max;
    int32_t w = f32_as_i32(fc) % smax;
    return (uint32_t)(w < 0 ? w + smax : w);
}
// material.rs:56-68,82-89 + color.rs:26-30
inline Color body_color(const Scene &s, uint32_t i, float u, float v) {
    const rg_scene_desc *d = s.d;
    if (d->coloration_kind[i] == RG_COLORATION_COLOR)
        return {d->color[3 * i], d->color[3 * i + 1], d->color[3 * i + 2]};
    const rg_texture_desc &t = d->textures[d->texture_id[i]];
    uint32_t x = tex_wrap(u + d->texture_offset[2 * i], t.width);
    uint32_t y = tex_wrap(v + d->texture_offset[2 * i + 1], t.height);
    const uint8_t *p = t.pixels + ((size_t)y * t.width + x) * t.channels;
    return {(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f};
}

// lights.rs:46-51
inline V3 light_direction_from(const Scene &s, uint32_t l, V3 p) {
    const double *lv = s.d->light_vec + 3 * l;
    V3 v{lv[0], lv[1], lv[2]};
    if (s.d->light_kind[l] == RG_LIGHT_DIRECTIONAL) return v_normalize(v_neg(v));
    return v_normalize(v_sub(v, p));
}
// lights.rs:53-58
inline double light_distance(const Scene &s, uint32_t l, V3 p) {
    if (s.d->light_kind[l] == RG_LIGHT_DIRECTIONAL) return std::numeric_limits<double>::infinity();
    const double *lv = s.d->light_vec + 3 * l;
    V3 w = v_sub(V3{lv[0], lv[1], lv[2]}, p);
    return std::sqrt(v_dot(w, w));
}
// lights.rs:36-44
inline float light_intensity(const Scene &s, uint32_t l, V3 p) {
    float intensity = s.d->light_intensity[l];
    if (s.d->light_kind[l] == RG_LIGHT_DIRECTIONAL) return intensity;
    const double *lv = s.d->light_vec + 3 * l;
    V3 w = v_sub(V3{lv[0], lv[1], lv[2]}, p);
    float r2 = (float)v_dot(w, w);
    return intensity / ((4.0f * PI_F32) * r2);
}

// rendering.rs:132-172
Color shade_diffuse(const Scene &s, uint32_t body, V3 hp, V3 n, Counters &c) {
    float u, v;
    texture_coords(s, body, hp, &u, &v);
    Color bc = body_color(s, body, u, v);
    Color fin{0.0f, 0.0f, 0.0f};
    for (uint32_t l = 0; l < s.d->n_lights; ++l) {
        V3 dir = light_direction_from(s, l, hp);
        Ray shadow = ray_new(v_add(hp, v_mul(n, SHADOW_BIAS)), dir);
        double t;
        uint32_t b;
        c.shadow++;
        bool hit = trace(s, shadow, &t, &b, c);
        bool in_light = !hit || t > light_distance(s, l, hp);
        float li = in_light ? light_intensity(s, l, hp) : 0.0f;
        float power = fmaxf((float)v_dot(n, dir), 0.0f) * li;
        float reflected = s.d->albedo[body] / PI_F32;
        const float *lc = s.d->light_color + 3 * l;
        Color light_color = c_scale(c_scale(Color{lc[0], lc[1], lc[2]}, power), reflected);
        fin = c_add(fin, c_mul(bc, light_color));
    }
    return c_clamp(fin);
}

// rendering.rs:174-200
double fresnel(V3 incident, V3 normal, float index) {
    double idn = v_dot(incident, normal);
    double eta_i, eta_t;
    if (idn > 0.0) { eta_i = (double)index; eta_t = 1.0; }
    else { eta_i = 1.0; eta_t = (double)index; }
    double sin_t = eta_i / eta_t * std::sqrt(std::fmax(1.0 - idn * idn, 0.0));
    if (sin_t > 1.0) return 1.0;
    double cos_t = std::sqrt(std::fmax(1.0 - sin_t * sin_t, 0.0));
    double cos_i = std::fabs(cos_t);
    double r_s = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    double r_p = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (r_s * r_s + r_p * r_p) / 2.0;
}

// ray.rs:56-60
inline Ray create_reflection(V3 normal, V3 incident, V3 p) {
    V3 origin = v_add(p, v_mul(normal, SHADOW_BIAS));
    double k = 2.0 * v_dot(incident, normal);
    V3 direction = v_sub(incident, v_mul(normal, k));
    return ray_new(origin, direction);
}
// ray.rs:62-94
inline bool create_transmission(V3 normal, V3 incident, V3 p, double bias, float index, Ray *out) {
    V3 ref_n = normal;
    double eta_t = (double)index, eta_i = 1.0;
    double idn = v_dot(incident, normal);
    if (idn < 0.0) idn = -idn;
    else { ref_n = v_neg(normal); eta_t = 1.0; eta_i = (double)index; }
    double eta = eta_i / eta_t;
    double k = 1.0 - (eta * eta) * (1.0 - idn * idn);
    if (k < 0.0) return false;
    V3 origin = v_add(p, v_mul(ref_n, -bias));
    V3 direction = v_sub(v_mul(v_add(incident, v_mul(ref_n, idn)), eta), v_mul(ref_n, std::sqrt(k)));
    *out = ray_new(origin, direction);
    return true;
}

Color cast_ray(const Scene &s, const Ray &ray, uint32_t depth, int kind, Counters &c);

// rendering.rs:80-120
Color get_color(const Scene &s, const Ray &ray, double distance, uint32_t body, uint32_t depth,
                Counters &c) {
    V3 hp = v_add(ray.origin, v_mul(ray.direction, distance));
    V3 n = surface_normal(s, body, hp, c);
    const rg_scene_desc *d = s.d;
    switch (d->surface_kind[body]) {
        case RG_SURFACE_DIFFUSE: return shade_diffuse(s, body, hp, n, c);
        case RG_SURFACE_REFLECTING: {
            float reflectivity = d->surface_param[2 * body];
            Color diffuse = shade_diffuse(s, body, hp, n, c);
            Ray rr = create_reflection(n, ray.direction, hp);
            Color refl = cast_ray(s, rr, depth + 1, 0, c);
            return c_add(c_scale(diffuse, 1.0f - reflectivity), c_scale(refl, reflectivity));
        }
        default: {
            float index = d->surface_param[2 * body];
            float transparency = d->surface_param[2 * body + 1];
            float kr = (float)fresnel(ray.direction, n, index);
            float u, v;
            texture_coords(s, body, hp, &u, &v);
            Color surface = body_color(s, body, u, v);
            Color refraction;
            if (kr < 1.0f) {
                Ray tr;
                if (create_transmission(n, ray.direction, hp, SHADOW_BIAS, index, &tr)) {
                    refraction = cast_ray(s, tr, depth + 1, 1, c);
                } else {
                    c.err_trans++;  // rendering.rs:106 unwrap() on None
                    refraction = s.default_color;
                }
            } else {
                refraction = s.default_color;
            }
            Ray rr = create_reflection(n, ray.direction, hp);
            Color reflection = cast_ray(s, rr, depth + 1, 0, c);
            Color col = c_add(c_scale(reflection, kr), c_scale(refraction, 1.0f - kr));
            col = c_mul(c_scale(col, transparency), surface);
            return col;
        }
    }
}

// rendering.rs:122-130
Color cast_ray(const Scene &s, const Ray &ray, uint32_t depth, int kind, Counters &c) {
    if (depth >= s.max_depth) return s.default_color;
    if (kind == 0) c.reflection++; else c.transmission++;
    double t;
    uint32_t b;
    if (trace(s, ray, &t, &b, c)) return get_color(s, ray, t, b, depth, c);
    return s.default_color;
}

// ray.rs:37-54
inline Ray create_prime(const Scene &s, uint32_t x, uint32_t y, uint32_t width, uint32_t height) {
    double aspect = (double)width / (double)height;
    double fa = s.fov_adjustment;
    double sx = ((((double)x + 0.5) / (double)width) * 2.0 - 1.0) * aspect * fa;
    double sy = (1.0 - (((double)y + 0.5) / (double)height) * 2.0) * fa;
    return ray_new(V3{0.0, 0.0, 0.0}, v_normalize(V3{sx, sy, -1.0}));
}

// rendering.rs:71-78
inline Color render_pixel(const Scene &s, uint32_t x, uint32_t y, uint32_t w, uint32_t h, Counters &c) {
    Ray ray = create_prime(s, x, y, w, h);
    double t;
    uint32_t b;
    c.primary++;
    if (trace(s, ray, &t, &b, c)) return get_color(s, ray, t, b, 0, c);
    return s.default_color;
}

}  // namespace

extern "C" {

// f64::to_radians = self * (PI / 180.0); ray.rs:46
double rgo_fov_adjustment(double fov_degrees) {
    const double pi = 3.14159265358979323846264338327950288;
    return std::tan((fov_degrees * (pi / 180.0)) / 2.0);
}

// rendering.rs:24-38 restricted to rows [y0, y1); `threads` worker threads pull
// rows from a shared counter (the analogue of rayon's work-stealing par_iter).
// If `rgb_f32` is non-NULL it also receives the unquantised colours (3 floats
// per pixel) — what render_image_stream sends (rendering.rs:59-65).
int rgo_render_rows(const rg_scene_desc *desc, uint32_t width, uint32_t height, uint32_t y0,
                    uint32_t y1, uint32_t max_depth_override, int threads, uint8_t *rgba_out,
                    float *rgb_f32, rg_stats *stats) {
    if (!desc || !rgba_out) return RG_E_INVALID;
    if (width < height) return RG_E_PORTRAIT;  // ray.rs:42
    if ((uint64_t)width * height > 0xFFFFFFFFull) return RG_E_TOO_LARGE;
    if (y1 > height || y0 > y1) return RG_E_INVALID;
    Scene s;
    s.d = desc;
    s.fov_adjustment = rgo_fov_adjustment(desc->fov);
    s.max_depth = desc->max_recursion_depth;
    if (max_depth_override != 0xFFFFFFFFu && max_depth_override < s.max_depth)
        s.max_depth = max_depth_override;  // main.rs:119-123
    s.default_color = {desc->default_color[0], desc->default_color[1], desc->default_color[2]};
    if (threads < 1) threads = 1;
    // rendering.rs:27-35 is a rayon par_iter over PIXELS (adaptive splitting): hand pixels out in
    // small chunks from a shared counter, so that a short row band still keeps every thread busy
    const uint64_t total = (uint64_t)(y1 > y0 ? y1 - y0 : 0) * width;
    const uint64_t kChunk = 64;
    std::atomic<uint64_t> next_px{0};
    std::vector<Counters> counters((size_t)threads);
    auto worker = [&](int tid) {
        Counters &c = counters[(size_t)tid];
        for (;;) {
            const uint64_t p0 = next_px.fetch_add(kChunk);
            if (p0 >= total) break;
            const uint64_t p1 = p0 + kChunk < total ? p0 + kChunk : total;
            for (uint64_t p = p0; p < p1; ++p) {
                const uint32_t x = (uint32_t)(p % width), y = y0 + (uint32_t)(p / width);
                Color col = render_pixel(s, x, y, width, height, c);
                size_t o = ((size_t)(y - y0) * width + x);
                rgba_out[4 * o + 0] = f32_as_u8(col.r * 255.0f);  // color.rs:32-37
                rgba_out[4 * o + 1] = f32_as_u8(col.g * 255.0f);
                rgba_out[4 * o + 2] = f32_as_u8(col.b * 255.0f);
                rgba_out[4 * o + 3] = 255;
                if (rgb_f32) { rgb_f32[3 * o] = col.r; rgb_f32[3 * o + 1] = col.g; rgb_f32[3 * o + 2] = col.b; }
            }
        }
    };
    if (threads == 1) worker(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
        for (auto &t : pool) t.join();
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (const Counters &c : counters) {
            stats->rays_primary += c.primary;
            stats->rays_shadow += c.shadow;
            stats->rays_reflection += c.reflection;
            stats->rays_transmission += c.transmission;
            stats->exact_tests += c.exact_tests;
            stats->err_nan_distance += c.err_nan;
            stats->err_transmission_none += c.err_trans;
            stats->err_aabb_normal += c.err_aabb;
        }
        stats->body_tests = stats->exact_tests;
    }
    return RG_OK;
}

// Known-answer probes for unit tests of individual reference functions.
uint32_t rgo_texture_wrap(float val, uint32_t max) { return tex_wrap(val, max); }       // material.rs:70-79
double rgo_fresnel(const double *incident, const double *normal, float index) {        // rendering.rs:174-200
    return fresnel(V3{incident[0], incident[1], incident[2]}, V3{normal[0], normal[1], normal[2]}, index);
}
uint8_t rgo_quantise(float c) { return f32_as_u8(c * 255.0f); }                        // color.rs:32-37
// Body::intersect through Ray::new (bodies.rs:336-345, ray.rs:23-35): returns 1 and *t on a hit.
int rgo_intersect(uint8_t kind, const double *geom8, const double *origin, const double *direction, double *t) {
    Ray r = ray_new(V3{origin[0], origin[1], origin[2]}, V3{direction[0], direction[1], direction[2]});
    switch (kind) {
        case RG_BODY_SPHERE: return sphere_intersect(geom8, r, t);
        case RG_BODY_PLANE: return plane_intersect(geom8, r, t);
        case RG_BODY_DISK: return disk_intersect(geom8, r, t);
        default: return aabb_intersect(geom8, r, t);
    }
}
int rgo_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
