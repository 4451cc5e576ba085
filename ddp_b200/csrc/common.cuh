// Shared constants and small device helpers of the DDP decode-head kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddp {

constexpr int kE = 256;        // embed dims
constexpr int kHeads = 8;
constexpr int kPoints = 4;
constexpr int kHeadDim = 32;
constexpr int kFFN = 1024;
constexpr int kTimeDim = 1024;
constexpr int kSampW = 96;     // 64 sampling offsets (head, point, xy) + 32 attention weights (head, point)
constexpr int kMaxLayers = 8;
constexpr int kMaxSteps = 1024;

__device__ __forceinline__ float gelu_erf(float x) {
    // nn.GELU() (exact erf form), vmmcv/cnn/bricks/transformer.py:253-263 with act_cfg=GELU
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace ddp
