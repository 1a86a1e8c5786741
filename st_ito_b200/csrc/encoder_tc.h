// Tensor-core (tcgen05 / TMEM / TMA) encoder path: declarations used by the C-ABI layer.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "stito_internal.h"

namespace stito {

struct TcWorkspace {
    static constexpr int kBufs = 7;  // 0-3: activations (see tc_encoder_forward), 4-6: Winograd V_hi, V_lo, M
    void *buf[kBufs] = {};
    size_t cap[kBufs] = {};
};

// true when the tcgen05 encoder is compiled in and enabled
bool tc_available();
const char *tc_last_error();
// Split the BN-folded fp32 weights wf [9][cin][cout] into the fp16 hi/lo operands of layer `cl`.
int tc_prepare_layer(const float *wf, int cin, int cout, ConvLayer *cl, std::vector<void *> *owned);
// feat [N][T][mel] fp32 (normalised log-mel) -> pooled [N][2048] fp32
int tc_encoder_forward(cudaStream_t st, const EncoderDev &enc, TcWorkspace &ws, const float *feat, int N, int T,
                       int mel, float *pooled, int *launches, cudaEvent_t *layer_events);
void tc_workspace_release(TcWorkspace *ws);

}  // namespace stito
