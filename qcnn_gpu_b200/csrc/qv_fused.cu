// placeholder until the tcgen05 kernel lands
#include "qv_fused.h"
namespace qv {
FusedModel *fused_upload(const ModelHost &, cudaStream_t) { return nullptr; }
void fused_free(FusedModel *) {}
cudaError_t fused_forward(const FusedModel *, const uint8_t *, uint8_t *, int, int, int, cudaStream_t, long long *) { return cudaErrorNotSupported; }
}
