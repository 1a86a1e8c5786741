// placeholder until the tcgen05 path lands
#include "encoder_tc.h"
namespace stito {
bool tc_available() { return false; }
const char *tc_last_error() { return "tensor-core encoder not built"; }
int tc_prepare_layer(const float *, int, int, ConvLayer *cl, std::vector<void *> *) { cl->w_hi = cl->w_lo = nullptr; cl->w_unscale = 1.0f; return 0; }
int tc_encoder_forward(cudaStream_t, const EncoderDev &, TcWorkspace &, const float *, int, int, int, float *, int *, cudaEvent_t *) { return -1; }
void tc_workspace_release(TcWorkspace *) {}
}  // namespace stito
