// rg_grid.cuh — Scene::trace with EXACT spatial culling.
//
// Scene::trace is a pure minimum over bodies (scene.rs:34-39), so any body that provably
// returns None — or a distance no smaller than one already found — may be skipped without
// changing the result.  Spheres are binned into a uniform grid; a ray walks the cells it
// pierces (3-D DDA in FP32 grid coordinates) and runs, per listed sphere, the FP32
// conservative cull and then the reference's exact FP64 test.  The result is the same
// lexicographic (distance, body index) minimum as the brute-force scan.
//
// Why no hit can be missed (DESIGN.md "grid"):
//   * a sphere is listed in every cell its bounding box, inflated by kGridInflate cells,
//     overlaps; FP32 traversal error is bounded by ~3e-4 cell (|grid coords| <= kGridMaxCoord),
//     far below the inflation, so the cell containing a hit point — or a neighbour the
//     inflated box also covers — is always visited;
//   * the walk stops only when the best exact distance lies before the current cell's exit
//     (with slack), so every sphere that could be nearer has been listed in a visited cell;
//   * rays the argument does not cover (direction not unit within 1e-9 — reflections off
//     un-normalised plane normals — non-finite or far-away origins) skip the grid and scan
//     every sphere exactly;
//   * spheres too large for the grid ("loose") and all non-sphere bodies are tested per ray.
#pragma once
#include "rg_trace.cuh"

namespace rg {

constexpr int kGridTraceThreads = 128;
constexpr float kGridInflate = 2e-3f;       // in cells; see above
constexpr float kGridMaxCoord = 2048.0f;    // |grid coordinate| limit for FP32 traversal
constexpr double kGridUnitTol = 1e-9;

template <bool ANY>
struct GridHit {
    Nearest best;
    bool occluded;
    double tmax;
    __device__ __forceinline__ void offer(double t, uint32_t body) {
        if (ANY) occluded = occluded || (t <= tmax);
        else best.offer(t, body);
    }
    // upper bound on distances still worth looking at (FP32, rounded up generously)
    __device__ __forceinline__ float bound() const {
        if (ANY) return tmax < 3.0e38 ? (float)tmax * 1.000001f + 1e-30f : 3.4e38f;
        return best.found() ? (float)best.t * 1.000001f + 1e-30f : 3.4e38f;
    }
};

template <bool ANY>
__device__ __forceinline__ void test_sphere(const DScene &s, const Ray &ray, const CullRay &cr, uint32_t sph,
                                            float4 rec, GridHit<ANY> &h, unsigned &n_exact, unsigned &nan_count) {
    if (cull_reject(cr, rec)) return;
    const double4 sp = s.sph[sph];
    double t;
    ++n_exact;
    if (sphere_intersect(sp.x, sp.y, sp.z, sp.w, ray, t)) {
        if (t != t) { ++nan_count; return; }
        h.offer(t, s.sph_body[sph]);
    }
}

template <bool ANY>
__global__ void __launch_bounds__(kGridTraceThreads) k_trace_grid(const DScene s, const TraceArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n_exact = 0, nan_count = 0;
    if (i < a.n) {
        const uint32_t pi = phys_index(a, i);
        const Ray ray = load_ray(a.q, pi);
        const GridDev &g = s.grid;
        GridHit<ANY> h;
        h.best.init();
        h.occluded = false;
        h.tmax = ANY ? a.tmax[pi] : 0.0;

        for (uint32_t m = 0; m < s.n_misc; ++m) {          // planes, disks, boxes: exact, per ray
            const uint32_t b = s.misc_body[m];
            double t;
            if (misc_intersect(s, b, ray, t)) {
                if (t != t) { ++nan_count; continue; }
                h.offer(t, b);
            }
        }
        n_exact += s.n_misc;

        const CullRay cr = make_cull_ray(s, ray, true);
        const D3 op = ray.o - d3(s.cull_ref[0], s.cull_ref[1], s.cull_ref[2]);
        const double D2 = dot(ray.d, ray.d);
        // grid-space ray (FP32): cells are unit cubes
        const float ogx = ((float)op.x - g.lo[0]) * g.inv_cell[0], ogy = ((float)op.y - g.lo[1]) * g.inv_cell[1],
                    ogz = ((float)op.z - g.lo[2]) * g.inv_cell[2];
        const float dgx = (float)ray.d.x * g.inv_cell[0], dgy = (float)ray.d.y * g.inv_cell[1],
                    dgz = (float)ray.d.z * g.inv_cell[2];
        const bool walkable = fabs(D2 - 1.0) <= kGridUnitTol && fabsf(ogx) <= kGridMaxCoord &&
                              fabsf(ogy) <= kGridMaxCoord && fabsf(ogz) <= kGridMaxCoord;   // false on NaN
        if (!walkable) {
            if (!(ANY && h.occluded))
                for (uint32_t k = 0; k < s.n_spheres; ++k) test_sphere<ANY>(s, ray, cr, k, s.cull4[k], h, n_exact, nan_count);
        } else if (!(ANY && h.occluded)) {
            for (uint32_t k = 0; k < g.n_loose; ++k) { const uint32_t sp = g.loose[k]; test_sphere<ANY>(s, ray, cr, sp, s.cull4[sp], h, n_exact, nan_count); }
            // clip the ray to the grid box [0, dim] (slabs; a zero component gives +-inf, handled by min/max)
            const float ix = 1.0f / dgx, iy = 1.0f / dgy, iz = 1.0f / dgz;
            const float dimx = (float)g.dim[0], dimy = (float)g.dim[1], dimz = (float)g.dim[2];
            float t0x = (0.0f - ogx) * ix, t1x = (dimx - ogx) * ix;
            float t0y = (0.0f - ogy) * iy, t1y = (dimy - ogy) * iy;
            float t0z = (0.0f - ogz) * iz, t1z = (dimz - ogz) * iz;
            // a ray parallel to a slab (d = 0): inside -> (-inf, +inf), outside -> empty
            if (dgx == 0.0f) { bool in = ogx >= 0.0f && ogx <= dimx; t0x = in ? -3.4e38f : 3.4e38f; t1x = in ? 3.4e38f : -3.4e38f; }
            if (dgy == 0.0f) { bool in = ogy >= 0.0f && ogy <= dimy; t0y = in ? -3.4e38f : 3.4e38f; t1y = in ? 3.4e38f : -3.4e38f; }
            if (dgz == 0.0f) { bool in = ogz >= 0.0f && ogz <= dimz; t0z = in ? -3.4e38f : 3.4e38f; t1z = in ? 3.4e38f : -3.4e38f; }
            float tenter = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
            float texit = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
            // slack: the box walls are kGridInflate-padded away from every sphere, so nudging inwards is safe
            if (tenter <= texit * 1.000001f + 1e-6f && tenter <= h.bound()) {
                const float ts = tenter;
                int cx = min(max((int)floorf(ogx + ts * dgx), 0), g.dim[0] - 1);
                int cy = min(max((int)floorf(ogy + ts * dgy), 0), g.dim[1] - 1);
                int cz = min(max((int)floorf(ogz + ts * dgz), 0), g.dim[2] - 1);
                const int sx = dgx > 0.0f ? 1 : -1, sy = dgy > 0.0f ? 1 : -1, sz = dgz > 0.0f ? 1 : -1;
                for (;;) {
                    // parameter at which the ray leaves this cell, per axis (recomputed, not accumulated)
                    const float nx = dgx != 0.0f ? ((float)(cx + (sx > 0)) - ogx) * ix : 3.4e38f;
                    const float ny = dgy != 0.0f ? ((float)(cy + (sy > 0)) - ogy) * iy : 3.4e38f;
                    const float nz = dgz != 0.0f ? ((float)(cz + (sz > 0)) - ogz) * iz : 3.4e38f;
                    const float tnext = fminf(nx, fminf(ny, nz));
                    const uint32_t cell = ((uint32_t)cz * (uint32_t)g.dim[1] + (uint32_t)cy) * (uint32_t)g.dim[0] + (uint32_t)cx;
                    const uint32_t b0 = g.cell_start[cell], b1 = g.cell_start[cell + 1];
                    for (uint32_t k = b0; k < b1; ++k) test_sphere<ANY>(s, ray, cr, g.cell_items[k], g.cell_cull4[k], h, n_exact, nan_count);
                    if (ANY && h.occluded) break;
                    // everything nearer than the cell exit has been seen: done (slack for FP32 t error)
                    if (h.bound() < tnext * 0.99999f - 1e-5f) break;
                    if (nx <= ny && nx <= nz) { cx += sx; if ((unsigned)cx >= (unsigned)g.dim[0]) break; }
                    else if (ny <= nz) { cy += sy; if ((unsigned)cy >= (unsigned)g.dim[1]) break; }
                    else { cz += sz; if ((unsigned)cz >= (unsigned)g.dim[2]) break; }
                }
            }
        }
        if (ANY) a.out_lit[pi] = h.occluded ? 0 : 1;
        else { a.out_t[pi] = h.best.t; a.out_body[pi] = h.best.body; }
    }
    unsigned long long ne = n_exact;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ne += __shfl_xor_sync(0xffffffffu, ne, o);
        nan_count += __shfl_xor_sync(0xffffffffu, nan_count, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (ne) atomicAdd(&a.ctr->exact_tests, ne);
        if (nan_count) atomicAdd(&a.ctr->err_nan, (unsigned long long)nan_count);
    }
}

}  // namespace rg
