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

// One sampling point of the deformable attention, resolved once per (token, head, point) in the epilogue of the
// sampling projection.  The reference's coordinate chain is evaluated op by op in fp32 so that the pixel position
// rounds exactly as the reference's does:
//   ref = (j + .5) / W                    deformable_head_with_time.py:76-84
//   loc = ref + off / W                   mmcv/ops/multi_scale_deform_attn.py:329-334
//   grid = 2 loc - 1                      :121
//   x = (grid + 1) * (W / 2) - .5         F.grid_sample(align_corners=False) unnormalisation (zero padding)
// Record: word 0 = clamped top-left token index (26 bits) | x-step bit 26 | y-step bit 27 | corner-valid bits 28..31
// (00, 01, 10, 11 = (y0,x0), (y0,x1), (y1,x0), (y1,x1)); fx, fy = bilinear fractions.
constexpr int kRecW = 128;     // words per token: [0,32) index words, [32,64) fx, [64,96) fy, [96,128) attention weights
// refx = (j + .5) / W and refy are computed by the caller with a true division (once per token); off / W uses the
// reciprocal (exact when W is a power of two, within 1 ulp otherwise: < 1e-7 px).
__device__ __forceinline__ void msda_resolve(float offx, float offy, float refx, float refy, float rW, float rH, int H, int W,
                                             uint32_t& word, float& fx, float& fy) {
    const float fW = (float)W, fH = (float)H;
    const float lx = __fadd_rn(refx, __fmul_rn(offx, rW));
    const float ly = __fadd_rn(refy, __fmul_rn(offy, rH));
    const float gx = __fadd_rn(__fmul_rn(2.0f, lx), -1.0f);
    const float gy = __fadd_rn(__fmul_rn(2.0f, ly), -1.0f);
    const float x = __fadd_rn(__fmul_rn(__fadd_rn(gx, 1.0f), fW * 0.5f), -0.5f);
    const float y = __fadd_rn(__fmul_rn(__fadd_rn(gy, 1.0f), fH * 0.5f), -0.5f);
    const float xf = floorf(x), yf = floorf(y);
    fx = x - xf;
    fy = y - yf;
    // NaN / huge offsets: every corner invalid
    const int x0 = (xf >= -2.0f && xf <= fW) ? (int)xf : -2;
    const int y0 = (yf >= -2.0f && yf <= fH) ? (int)yf : -2;
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
    const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
    const int x0c = min(max(x0, 0), W - 1), x1c = min(max(x0 + 1, 0), W - 1);
    const int y0c = min(max(y0, 0), H - 1), y1c = min(max(y0 + 1, 0), H - 1);
    word = (uint32_t)(y0c * W + x0c) | ((uint32_t)(x1c - x0c) << 26) | ((uint32_t)(y1c - y0c) << 27) |
           ((uint32_t)(vy0 && vx0) << 28) | ((uint32_t)(vy0 && vx1) << 29) | ((uint32_t)(vy1 && vx0) << 30) |
           ((uint32_t)(vy1 && vx1) << 31);
}

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
