// placeholder until the tcgen05 path lands (next commit)
#include "dx_internal.cuh"
int dx_launch_mlp_bf16(cudaStream_t, const DomainDev&, const NetDev&, const MlpBufs&, const Ctl*, const uint8_t*, const uint8_t*,
                       float*, float*, int, long long) {
    dx_set_error("NET_BF16 not built");
    return 0;
}
