// Tensor-core (tcgen05 / TMEM / TMA) encoder path: declarations used by the C-ABI layer.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "stito_internal.h"

namespace stito {

// CUtensorMap descriptors are pure functions of (base pointer, geometry, box): the workspace buffers and the
// weights do not move between generations, so every descriptor is encoded once and looked up afterwards
// (cuTensorMapEncodeTiled is a driver call; a generation used to make ~76 of them).
struct TcMapKey {
    const void *base;
    int dims[4];   // activation map: N, H, W, C; weight map: Cout, K, 0, 0
    int box[4];    // activation map: slabk, BW, BH, IPT; weight map: slabk, BN, 0, 0
    bool operator==(const TcMapKey &o) const {
        if (base != o.base) return false;
        for (int i = 0; i < 4; ++i)
            if (dims[i] != o.dims[i] || box[i] != o.box[i]) return false;
        return true;
    }
};
struct alignas(64) TcMapEntry {
    unsigned char map[128];  // CUtensorMap (128 bytes, 64-byte aligned)
    TcMapKey key;
};

struct TcWorkspace {
    static constexpr int kBufs = 7;  // 0-3: activations (see tc_encoder_forward), 4-6: Winograd V_hi, V_lo, M
    void *buf[kBufs] = {};
    size_t cap[kBufs] = {};
    std::vector<TcMapEntry *> maps;  // descriptor cache (cleared whenever a buffer is reallocated)
    int *overflow_flag = nullptr;    // device int: set when an activation exceeded the fp16 range (see store_tile)
    // Per-layer storage scale of the fp16 hi/lo activation pairs: layer l's output is stored as value * 2^act_shift[l]
    // (layer 11 writes fp32).  Calibrated by the C-ABI layer from `amax` (device, 12 unsigned = float bits of the largest
    // stored value per layer, nullable).
    int act_shift[12] = {6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6};
    unsigned *amax = nullptr;
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
