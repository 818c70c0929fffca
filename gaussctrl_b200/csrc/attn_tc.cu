// tcgen05 multi-source attention (placeholder until the TMEM kernel lands: reports "unsupported" for every shape so
// GCB_ATTN_AUTO resolves to the mma.sync kernel).
#include "../../include/gaussctrl_b200.h"
#include "common.cuh"

int gcb_attn_tc_supported(int Nq, int Nk, int heads, int d) {
    (void)Nq; (void)Nk; (void)heads; (void)d;
    return 0;
}
int gcb_attn_tc_launch(const void*, int, const void*, const void*, int, const void*, const void*, int, void*, int, int,
                       int, int, int, int, int, const int32_t*, const float*, float, cudaStream_t) {
    gcb_set_error("tcgen05 attention kernel not built");
    return GCB_ERR_UNSUPPORTED;
}
