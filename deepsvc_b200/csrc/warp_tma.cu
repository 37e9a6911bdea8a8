// placeholder: staged (shared-memory / TMA) forward warp -- filled in next.
#include "warp_common.cuh"
using namespace dsvc;
int dsvc_warp_fwd_tma_launch(const float*, const float*, float*, const float*, const float*,
                             const WarpParams&, bool, cudaStream_t) {
    return -1;
}
