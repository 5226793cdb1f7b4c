#include "rg_host.h"
namespace rg {
int wavefront_render(rg_scene *, uint32_t, uint32_t, uint32_t, uint32_t, uchar4 *, cudaStream_t, rg_stats *) {
    set_error("wavefront pipeline not built yet");
    return RG_E_INVALID;
}
}
