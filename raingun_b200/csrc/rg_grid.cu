#include "rg_host.h"
namespace rg {
int grid_build(rg_scene *sc, const std::vector<double> &) {
    sc->ds.grid.enabled = 0;
    return RG_OK;
}
}
