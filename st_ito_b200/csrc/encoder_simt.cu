// AFx-Rep encoder body on CUDA cores, fp32 (precision mode 0): NHWC implicit-GEMM 3x3
// convolutions with folded BatchNorm + ReLU, 2x2 average pooling, global pooling and the two
// linear heads.  Replaces Cnn14.forward's body (st_ito/models/panns.py:250-279) and ConvBlock
// (panns.py:25-80).  This is the exact-arithmetic reference mode of the library (same fp32
// semantics as the oracle up to summation order); the throughput mode is encoder_tc.cu.
#include "stito_internal.h"

namespace stito {

namespace {

// ------------------------------------------------------------ first conv (Cin == 1)
// x [N][H][W], w [9][1][64] -> y [N][H][W][64].  Memory-bound: 9 taps, 64 outputs per pixel.
__global__ void __launch_bounds__(256) conv_first_kernel(const float *__restrict__ x,
                                                         const float *__restrict__ w,
                                                         const float *__restrict__ bias,
                                                         float *__restrict__ y, int N, int H, int W) {
    __shared__ float ws[9][64];
    __shared__ float bs[64];
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) ws[i / 64][i % 64] = w[i];
    if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int64_t total = (int64_t)N * H * W * 8;  // 8 groups of 8 output channels per pixel
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(idx & 7);
        const int64_t pix = idx >> 3;
        const int wq = (int)(pix % W);
        const int hq = (int)((pix / W) % H);
        const int64_t n = pix / ((int64_t)W * H);
        float v[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int hh = hq + kh - 1, ww = wq + kw - 1;
                v[kh * 3 + kw] = (hh >= 0 && hh < H && ww >= 0 && ww < W)
                                     ? __ldg(x + (n * H + hh) * W + ww) : 0.0f;
            }
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fmaf(v[t], ws[t][g * 8 + c], acc[c]);
        float4 o0, o1;
        o0.x = fmaxf(acc[0] + bs[g * 8 + 0], 0.0f); o0.y = fmaxf(acc[1] + bs[g * 8 + 1], 0.0f);
        o0.z = fmaxf(acc[2] + bs[g * 8 + 2], 0.0f); o0.w = fmaxf(acc[3] + bs[g * 8 + 3], 0.0f);
        o1.x = fmaxf(acc[4] + bs[g * 8 + 4], 0.0f); o1.y = fmaxf(acc[5] + bs[g * 8 + 5], 0.0f);
        o1.z = fmaxf(acc[6] + bs[g * 8 + 6], 0.0f); o1.w = fmaxf(acc[7] + bs[g * 8 + 7], 0.0f);
        float4 *dst = reinterpret_cast<float4 *>(y + pix * 64 + g * 8);
        dst[0] = o0;
        dst[1] = o1;
    }
}

// ------------------------------------------------------------- generic 3x3 conv, SIMT
// GEMM view: M = N*H*W output pixels, N = Cout, K = 9*Cin (tap-major, channel-minor).
// 128 x BN tile, BK = 16, 8x8 outputs per thread, cin % 16 == 0, cout % BN == 0.
constexpr int BM = 128, BK = 16, APAD = 4;

template <int BN>
__global__ void __launch_bounds__((BM / 8) * (BN / 8)) conv3x3_simt_kernel(
    const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
    float *__restrict__ y, int Nimg, int H, int W, int Cin, int Cout) {
    constexpr int NT = (BM / 8) * (BN / 8);
    constexpr int A_LD = (BM * BK / 4) / NT;  // float4 loads of A per thread
    constexpr int B_LD = (BK * BN / 4) / NT;
    __shared__ __align__(16) float As[BK][BM + APAD];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int64_t M = (int64_t)Nimg * H * W;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-loader role: A_LD (pixel, channel-quad) pairs
    int a_pix[A_LD], a_q[A_LD], a_h[A_LD], a_w[A_LD];
    int64_t a_base[A_LD];
    bool a_ok[A_LD];
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
        const int idx = tid + i * NT;
        a_pix[i] = idx >> 2;
        a_q[i] = idx & 3;
        const int64_t m = m0 + a_pix[i];
        a_ok[i] = m < M;
        const int64_t mm = a_ok[i] ? m : 0;
        a_w[i] = (int)(mm % W);
        a_h[i] = (int)((mm / W) % H);
        a_base[i] = mm * Cin;  // NHWC offset of the pixel's channel 0
    }
    const int ty = tid / (BN / 8), tx = tid % (BN / 8);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    for (int tap = 0; tap < 9; ++tap) {
        const int dh = tap / 3 - 1, dw = tap % 3 - 1;
        const int64_t shift = ((int64_t)dh * W + dw) * Cin;
        bool ok[A_LD];
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            const int hh = a_h[i] + dh, ww = a_w[i] + dw;
            ok[i] = a_ok[i] && hh >= 0 && hh < H && ww >= 0 && ww < W;
        }
        for (int c0 = 0; c0 < Cin; c0 += BK) {
            float4 av[A_LD], bv[B_LD];
#pragma unroll
            for (int i = 0; i < A_LD; ++i)
                av[i] = ok[i] ? __ldg(reinterpret_cast<const float4 *>(x + a_base[i] + shift + c0 + a_q[i] * 4))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                const int idx = tid + i * NT;
                const int row = idx / (BN / 4), col = idx % (BN / 4);
                bv[i] = __ldg(reinterpret_cast<const float4 *>(
                    w + ((int64_t)tap * Cin + c0 + row) * Cout + n0 + col * 4));
            }
            __syncthreads();  // previous tile fully consumed
#pragma unroll
            for (int i = 0; i < A_LD; ++i) {
                As[a_q[i] * 4 + 0][a_pix[i]] = av[i].x;
                As[a_q[i] * 4 + 1][a_pix[i]] = av[i].y;
                As[a_q[i] * 4 + 2][a_pix[i]] = av[i].z;
                As[a_q[i] * 4 + 3][a_pix[i]] = av[i].w;
            }
#pragma unroll
            for (int i = 0; i < B_LD; ++i) {
                const int idx = tid + i * NT;
                const int row = idx / (BN / 4), col = idx % (BN / 4);
                *reinterpret_cast<float4 *>(&Bs[row][col * 4]) = bv[i];
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][BN / 2 + tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }
    // epilogue: + bias, ReLU, NHWC store
    float bsv[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bsv[j] = __ldg(bias + n0 + tx * 4 + j);
        bsv[4 + j] = __ldg(bias + n0 + BN / 2 + tx * 4 + j);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
        float4 o0, o1;
        o0.x = fmaxf(acc[i][0] + bsv[0], 0.f); o0.y = fmaxf(acc[i][1] + bsv[1], 0.f);
        o0.z = fmaxf(acc[i][2] + bsv[2], 0.f); o0.w = fmaxf(acc[i][3] + bsv[3], 0.f);
        o1.x = fmaxf(acc[i][4] + bsv[4], 0.f); o1.y = fmaxf(acc[i][5] + bsv[5], 0.f);
        o1.z = fmaxf(acc[i][6] + bsv[6], 0.f); o1.w = fmaxf(acc[i][7] + bsv[7], 0.f);
        *reinterpret_cast<float4 *>(y + m * Cout + n0 + tx * 4) = o0;
        *reinterpret_cast<float4 *>(y + m * Cout + n0 + BN / 2 + tx * 4) = o1;
    }
}

// ---------------------------------------------------------------------- pooling
__global__ void __launch_bounds__(256) avgpool_kernel(const float *__restrict__ x, float *__restrict__ y,
                                                      int N, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
    const int64_t total = (int64_t)N * Ho * Wo * C4;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % C4);
        const int64_t pix = idx / C4;
        const int wo = (int)(pix % Wo);
        const int ho = (int)((pix / Wo) % Ho);
        const int64_t n = pix / ((int64_t)Wo * Ho);
        const float4 *src = reinterpret_cast<const float4 *>(x + ((n * H + 2 * ho) * W + 2 * wo) * C) + c4;
        const float4 a = __ldg(src), b = __ldg(src + C4);
        const float4 c = __ldg(src + (int64_t)W * C4), d = __ldg(src + (int64_t)W * C4 + C4);
        float4 o;
        o.x = ((a.x + b.x) + (c.x + d.x)) * 0.25f;
        o.y = ((a.y + b.y) + (c.y + d.y)) * 0.25f;
        o.z = ((a.z + b.z) + (c.z + d.z)) * 0.25f;
        o.w = ((a.w + b.w) + (c.w + d.w)) * 0.25f;
        reinterpret_cast<float4 *>(y)[idx] = o;
    }
}

// x [N][H][W][C] -> y [N][C]:  m_h = mean_w x;  y = max_h m_h + mean_h m_h   (panns.py:262-266)
__global__ void __launch_bounds__(256) global_pool_kernel(const float *__restrict__ x, float *__restrict__ y,
                                                          int N, int H, int W, int C) {
    const int64_t total = (int64_t)N * C;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % C);
    const int64_t n = idx / C;
    float mx = -INFINITY, sum = 0.0f;
    for (int h = 0; h < H; ++h) {
        float s = 0.0f;
        for (int w = 0; w < W; ++w) s += __ldg(x + ((n * H + h) * W + w) * C + c);
        const float m = s / (float)W;
        mx = fmaxf(mx, m);
        sum += m;
    }
    y[idx] = mx + sum / (float)H;
}

// one warp per (item, head, output): out[b][e] = bias[e] + sum_k pooled[row][k] * w[e][k]
__global__ void __launch_bounds__(256) heads_kernel(const float *__restrict__ pooled,
                                                    const float *__restrict__ w_mid,
                                                    const float *__restrict__ b_mid,
                                                    const float *__restrict__ w_side,
                                                    const float *__restrict__ b_side, int B, int chs,
                                                    int E, int Kdim, float *__restrict__ mid,
                                                    float *__restrict__ side) {
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (int64_t)B * 2 * E) return;
    const int e = (int)(gw % E);
    const int head = (int)((gw / E) % 2);
    const int64_t b = gw / (2 * (int64_t)E);
    if (chs == 1 && head == 1) return;  // mono: side = mid, written by the head-0 warp
    const float4 *row = reinterpret_cast<const float4 *>(pooled + (b * chs + head) * Kdim);
    const float4 *w = reinterpret_cast<const float4 *>((head ? w_side : w_mid) + (int64_t)e * Kdim);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k = lane; k < Kdim / 4; k += 32) {
        const float4 x = __ldg(row + k), y = __ldg(w + k);
        a0 = fmaf(x.x, y.x, a0); a1 = fmaf(x.y, y.y, a1); a2 = fmaf(x.z, y.z, a2); a3 = fmaf(x.w, y.w, a3);
    }
    float v = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
        v += __ldg((head ? b_side : b_mid) + e);
        (head ? side : mid)[b * E + e] = v;
        if (chs == 1) side[b * E + e] = v;  // mono: side_embed = mid_embed (panns.py:273-274)
    }
}

inline int blocks_for(int64_t total, int threads) {
    int64_t b = (total + threads - 1) / threads;
    const int64_t cap = 148 * 32;
    return (int)(b < cap ? b : cap);
}

}  // namespace

cudaError_t launch_conv_first(cudaStream_t st, const float *x, const ConvLayer &l, float *y, int N, int H,
                              int W, int *launches) {
    if (l.cin != 1 || l.cout != 64) return cudaErrorInvalidValue;
    conv_first_kernel<<<blocks_for((int64_t)N * H * W * 8, 256), 256, 0, st>>>(x, l.w, l.bias, y, N, H, W);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_conv_simt(cudaStream_t st, const float *x, const ConvLayer &l, float *y, int N, int H,
                             int W, int *launches) {
    if (l.cin % BK != 0 || l.cout % 64 != 0) return cudaErrorInvalidValue;
    const int64_t M = (int64_t)N * H * W;
    if (l.cout % 128 == 0) {
        dim3 grid((unsigned)((M + BM - 1) / BM), l.cout / 128);
        conv3x3_simt_kernel<128><<<grid, 256, 0, st>>>(x, l.w, l.bias, y, N, H, W, l.cin, l.cout);
    } else {
        dim3 grid((unsigned)((M + BM - 1) / BM), l.cout / 64);
        conv3x3_simt_kernel<64><<<grid, 128, 0, st>>>(x, l.w, l.bias, y, N, H, W, l.cin, l.cout);
    }
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_avgpool(cudaStream_t st, const float *x, float *y, int N, int H, int W, int C,
                           int *launches) {
    if (C % 4 != 0) return cudaErrorInvalidValue;
    const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 4);
    avgpool_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, y, N, H, W, C);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_global_pool(cudaStream_t st, const float *x, float *y, int N, int H, int W, int C,
                               int *launches) {
    const int64_t total = (int64_t)N * C;
    global_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, y, N, H, W, C);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_heads(cudaStream_t st, const float *pooled, const EncoderDev &enc, int B, int chs,
                         float *mid, float *side, int *launches) {
    const int64_t total = (int64_t)B * 2 * enc.embed_dim * 32;  // one warp per output
    heads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pooled, enc.fc_w[0], enc.fc_b[0],
                                                                 enc.fc_w[1], enc.fc_b[1], B, chs,
                                                                 enc.embed_dim, 2048, mid, side);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace stito
