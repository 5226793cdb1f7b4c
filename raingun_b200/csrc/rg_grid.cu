// rg_grid.cu — host-side construction of the exact-culling grid (rg_grid.cuh) at scene upload.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "rg_grid.cuh"
#include "rg_host.h"

namespace rg {

template <typename T>
static int upload_vec(rg_scene *sc, const std::vector<T> &v, const T **out) {
    *out = nullptr;
    if (v.empty()) return RG_OK;
    void *p = sc->arena.alloc(v.size() * sizeof(T));
    if (!p) return RG_E_NOMEM;
    RG_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const T *>(p);
    return RG_OK;
}

int grid_build(rg_scene *sc, const std::vector<double> &sph, const std::vector<float4> &cull) {
    GridDev &g = sc->ds.grid;
    g = GridDev{};
    const uint32_t n = sc->ds.n_spheres;
    if (n < 8) return RG_OK;   // brute force is the right tool for a handful of bodies

    // bounds of the finite spheres, relative to the cull reference point P
    const double *P = sc->ds.cull_ref;
    std::vector<uint32_t> loose, binned;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    bool have = false;
    std::vector<double> rel((size_t)n * 4);
    for (uint32_t i = 0; i < n; ++i) {
        double c[3] = {sph[4 * (size_t)i] - P[0], sph[4 * (size_t)i + 1] - P[1], sph[4 * (size_t)i + 2] - P[2]};
        double r = std::fabs(sph[4 * (size_t)i + 3]);
        bool ok = std::isfinite(c[0]) && std::isfinite(c[1]) && std::isfinite(c[2]) && std::isfinite(r) &&
                  std::fabs(c[0]) < 1e6 && std::fabs(c[1]) < 1e6 && std::fabs(c[2]) < 1e6 && r < 1e6;
        rel[4 * (size_t)i] = c[0]; rel[4 * (size_t)i + 1] = c[1]; rel[4 * (size_t)i + 2] = c[2]; rel[4 * (size_t)i + 3] = r;
        if (!ok) { loose.push_back(i); continue; }
        binned.push_back(i);
        for (int k = 0; k < 3; ++k) {
            if (!have) { lo[k] = c[k] - r; hi[k] = c[k] + r; }
            else { lo[k] = std::fmin(lo[k], c[k] - r); hi[k] = std::fmax(hi[k], c[k] + r); }
        }
        have = true;
    }
    if (binned.size() < 8) return RG_OK;

    // resolution: ~kDensity cells per sphere, cubic cells, at most 256 per axis
    const double kDensity = 4.0;
    double ext[3], vol = 1.0;
    for (int k = 0; k < 3; ++k) { ext[k] = std::fmax(hi[k] - lo[k], 1e-9); vol *= ext[k]; }
    double cell = std::cbrt(vol / (kDensity * (double)binned.size()));
    if (!(cell > 0.0) || !std::isfinite(cell)) return RG_OK;
    int dim[3];
    for (int k = 0; k < 3; ++k) dim[k] = (int)std::min(256.0, std::max(1.0, std::ceil(ext[k] / cell)));
    // pad the box so that no sphere comes within the inflation distance of a wall
    double csz[3];
    for (int k = 0; k < 3; ++k) {
        double c0 = ext[k] / dim[k];
        lo[k] -= 0.02 * c0;
        hi[k] += 0.02 * c0;
        csz[k] = (hi[k] - lo[k]) / dim[k];
    }
    const size_t ncells = (size_t)dim[0] * dim[1] * dim[2];

    // a sphere is listed in every cell its bounding box inflated by kGridInflate cells overlaps
    auto cell_range = [&](uint32_t i, int k, int &a, int &b) {
        double c = rel[4 * (size_t)i + k], r = rel[4 * (size_t)i + 3];
        double fa = (c - r - lo[k]) / csz[k] - 2.0 * kGridInflate;
        double fb = (c + r - lo[k]) / csz[k] + 2.0 * kGridInflate;
        a = std::max(0, std::min(dim[k] - 1, (int)std::floor(fa)));
        b = std::max(0, std::min(dim[k] - 1, (int)std::floor(fb)));
    };
    const uint64_t kMaxCellsPerSphere = 512;
    std::vector<uint32_t> count(ncells + 1, 0);
    std::vector<uint32_t> kept;
    for (uint32_t i : binned) {
        int a[3], b[3];
        for (int k = 0; k < 3; ++k) cell_range(i, k, a[k], b[k]);
        uint64_t span = (uint64_t)(b[0] - a[0] + 1) * (b[1] - a[1] + 1) * (b[2] - a[2] + 1);
        if (span > kMaxCellsPerSphere) { loose.push_back(i); continue; }
        kept.push_back(i);
        for (int z = a[2]; z <= b[2]; ++z)
            for (int y = a[1]; y <= b[1]; ++y)
                for (int x = a[0]; x <= b[0]; ++x) count[((size_t)z * dim[1] + y) * dim[0] + x + 1]++;
    }
    if (loose.size() > 256 || kept.size() < 8) return RG_OK;   // not a scene a uniform grid suits
    for (size_t c = 0; c < ncells; ++c) count[c + 1] += count[c];
    std::vector<uint32_t> start(count.begin(), count.end());
    std::vector<uint32_t> items(start[ncells]);
    std::vector<float4> items_cull(start[ncells]);
    std::vector<uint32_t> fill(start.begin(), start.end() - 1);
    for (uint32_t i : kept) {   // ascending sphere index, so every cell list is sorted
        int a[3], b[3];
        for (int k = 0; k < 3; ++k) cell_range(i, k, a[k], b[k]);
        for (int z = a[2]; z <= b[2]; ++z)
            for (int y = a[1]; y <= b[1]; ++y)
                for (int x = a[0]; x <= b[0]; ++x) {
                    uint32_t pos = fill[((size_t)z * dim[1] + y) * dim[0] + x]++;
                    items[pos] = i;
                    items_cull[pos] = cull[i];
                }
    }
    std::sort(loose.begin(), loose.end());

    // Inline cell records: the first two items of every cell live IN the cell's 48-byte record
    // (their cull records + indices), so a cell visit is one fixed-size fetch and two culls for
    // every lane; only cells with more than two items (~4 % at the default density) touch the
    // overflow lists above.  Empty slots hold a record that every cullable ray rejects.
    const float kInf = std::numeric_limits<float>::infinity();
    std::vector<float4> recs(ncells * 3);
    for (size_t c = 0; c < ncells; ++c) {
        const uint32_t b0 = start[c], b1 = start[c + 1], n = b1 - b0;
        recs[3 * c] = n > 0 ? items_cull[b0] : make_float4(0.f, 0.f, 0.f, kInf);
        recs[3 * c + 1] = n > 1 ? items_cull[b0 + 1] : make_float4(0.f, 0.f, 0.f, kInf);
        uint4 meta = make_uint4(n > 0 ? items[b0] : 0xFFFFFFFFu, n > 1 ? items[b0 + 1] : 0xFFFFFFFFu,
                                n > 2 ? b0 + 2 : 0u, n > 2 ? b1 : 0u);
        std::memcpy(&recs[3 * c + 2], &meta, sizeof(meta));
    }

    int rc;
    if ((rc = upload_vec(sc, recs, &g.cell_rec))) return rc;
    if ((rc = upload_vec(sc, start, &g.cell_start))) return rc;
    if ((rc = upload_vec(sc, items, &g.cell_items))) return rc;
    if ((rc = upload_vec(sc, items_cull, &g.cell_cull4))) return rc;
    if ((rc = upload_vec(sc, loose, &g.loose))) return rc;
    g.n_loose = (uint32_t)loose.size();
    for (int k = 0; k < 3; ++k) {
        g.lo[k] = (float)lo[k];
        g.cell[k] = (float)csz[k];
        g.inv_cell[k] = (float)(1.0 / csz[k]);
        g.dim[k] = dim[k];
    }
    g.enabled = 1;
    return RG_OK;
}

}  // namespace rg
