// ck_net_tc.cu -- tcgen05 tower (placeholder until the kernel lands in this round)
#include "ck_net.cuh"
namespace ck {
int net_tc_prepare(ck_net *) { return CK_OK; }
int net_tc_tower(ck_net *, const ck_leaf *, int64_t, const int32_t *, float *, float *, cudaStream_t, int *) {
    return fail(CK_ERR_STATE, "tcgen05 tower not built yet");
}
}  // namespace ck
