// Non-GEMM kernels of the DDP decode head: deformable gather, step epilogues (argmax -> embedding
// LUT -> DDIM update, softmax accumulation), positional encoding, time embeddings, layout changes.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace ddp {

// fp16 planes of an fp32 activation for the tensor-core GEMMs: hi = fp16(16 x), lo = fp16(16 x - hi)
// (see gemm_tc.cuh); lo == nullptr stores the hi plane only.
constexpr float kSplitScale = 16.0f;
__device__ __forceinline__ void split8_store(const float (&v)[8], __half* hi, __half* lo) {
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float s = fminf(fmaxf(v[i] * kSplitScale, -65504.0f), 65504.0f);
        h[i] = __float2half_rn(s);
        l[i] = __float2half_rn(s - __half2float(h[i]));
    }
    *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(l);
}

// elementwise fp32 -> fp16 planes, n a multiple of 8
__global__ void k_split_planes(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo, size_t n8) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n8) return;
    float4 a = *reinterpret_cast<const float4*>(src + idx * 8);
    float4 b = *reinterpret_cast<const float4*>(src + idx * 8 + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    split8_store(v, hi + idx * 8, lo ? lo + idx * 8 : nullptr);
}

// weights: dst planes [rows_pad][K] fp16 of scale * src[r*row_stride + k*k_stride + off]; rows >= rows are zero
__global__ void k_split_weight(const float* __restrict__ src, int rows, int K, int row_stride, int k_stride, int off,
                               float scale, __half* __restrict__ hi, __half* __restrict__ lo, int row0, int ld) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * K) return;
    int k = idx % K, r = idx / K;
    float s = src[(size_t)r * row_stride + (size_t)k * k_stride + off] * scale;
    __half hh = __float2half_rn(s);
    hi[(size_t)(row0 + r) * ld + k] = hh;
    lo[(size_t)(row0 + r) * ld + k] = __float2half_rn(s - __half2float(hh));
}
// scaled identity block: hi[r][col0 + r] = c (a power of two), used to carry a residual through the GEMM
__global__ void k_identity_block(__half* __restrict__ hi, int n, int ld, int col0, float c) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) hi[(size_t)r * ld + col0 + r] = __float2half_rn(c);
}

// ------------------------------------------------------------------------------------------------
// layout changes
// ------------------------------------------------------------------------------------------------
// src [imgs][C][N] (NCHW) -> dst [imgs][N][C] (token-major).  32x32 smem tiles.
__global__ void k_nchw_to_tokens(const float* __restrict__ src, float* __restrict__ dst, int C, int N) {
    __shared__ float tile[32][33];
    int img = blockIdx.z;
    int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* s = src + (size_t)img * C * N;
    float* d = dst + (size_t)img * C * N;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, n = n0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && n < N) ? s[(size_t)c * N + n] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int n = n0 + i, c = c0 + threadIdx.x;
        if (n < N && c < C) d[(size_t)n * C + c] = tile[threadIdx.x][i];
    }
}

// src [rows_in][K] -> dst [K][ld] with dst[k][col0 + r] = src[r][k*src_kstride + k_off] (weights repack);
// generic strided gather so that conv weights (O, I, kh, kw) can be sliced.
__global__ void k_repack_transposed(const float* __restrict__ src, int rows, int K, int src_row_stride,
                                    int src_k_stride, int src_off, float* __restrict__ dst, int ld, int col0) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * K) return;
    int r = idx % rows, k = idx / rows;
    dst[(size_t)k * ld + col0 + r] = src[(size_t)r * src_row_stride + (size_t)k * src_k_stride + src_off];
}

// ------------------------------------------------------------------------------------------------
// shape-only constants
// ------------------------------------------------------------------------------------------------
// SinePositionalEncoding(num_feats=128, normalize=True, offset=-0.5, scale=2*pi, temperature=10000, eps=1e-6)
// segmentation/mmseg/models/utils/transformer.py:78-113, evaluated with the same fp32 op order.
// pe [N][256], channel = [pos_y(128) | pos_x(128)], pos[2k] = sin, pos[2k+1] = cos.
__global__ void k_sine_pe(float* __restrict__ pe, int H, int W) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int N = H * W;
    if (idx >= N * kE) return;
    int c = idx % kE, n = idx / kE;
    int i = n / W, j = n % W;
    const float scale = 6.283185307179586f;
    float embed;
    int k;
    if (c < 128) {
        embed = __fmul_rn(__fdiv_rn((float)(i + 1) - 0.5f, __fadd_rn((float)H, 1e-6f)), scale);
        k = c;
    } else {
        embed = __fmul_rn(__fdiv_rn((float)(j + 1) - 0.5f, __fadd_rn((float)W, 1e-6f)), scale);
        k = c - 128;
    }
    float expo = (float)(2 * (k / 2)) / 128.0f;
    float dim_t = powf(10000.0f, expo);
    float v = __fdiv_rn(embed, dim_t);
    pe[idx] = (k & 1) ? cosf(v) : sinf(v);
}

// ------------------------------------------------------------------------------------------------
// time embeddings (data independent: computed once per plan for all T steps)
// ------------------------------------------------------------------------------------------------
// LearnedSinusoidalPosEmb, segmentation/mmseg/models/segmentors/ddp.py:31-46: [l, sin(l w 2pi), cos(l w 2pi)]
__global__ void k_fourier(const float* __restrict__ time_in, const float* __restrict__ w, int half,
                          float* __restrict__ out, int T) {
    int t = blockIdx.x;
    int i = threadIdx.x;
    int width = 2 * half + 1;
    if (t >= T || i >= width) return;
    float l = time_in[t];
    float v;
    if (i == 0) v = l;
    else {
        int k = (i - 1) % half;
        float f = __fmul_rn(__fmul_rn(__fmul_rn(l, w[k]), 2.0f), 3.14159265358979323846f);
        v = (i - 1 < half) ? sinf(f) : cosf(f);
    }
    out[t * width + i] = v;
}

// y[t][i] = act_out( sum_k W[i][k] * act_in(x[t][k]) + b[i] );  one warp per (i, t)
template <int ACT_IN /*0 none, 1 silu*/, int ACT_OUT /*0 none, 1 gelu*/>
__global__ void k_gemv(const float* __restrict__ Wm, const float* __restrict__ b, const float* __restrict__ x,
                       float* __restrict__ y, int rows, int K, int x_stride, int y_stride) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int t = blockIdx.y;
    if (warp >= rows) return;
    const float* wr = Wm + (size_t)warp * K;
    const float* xr = x + (size_t)t * x_stride;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) {
        float xv = xr[k];
        if (ACT_IN == 1) xv = silu(xv);
        s = fmaf(wr[k], xv, s);
    }
    s = warp_sum(s);
    if (lane == 0) {
        s += b[warp];
        if (ACT_OUT == 1) s = gelu_erf(s);
        y[(size_t)t * y_stride + warp] = s;
    }
}

// LN2 followed by FiLM, y = (n * gamma + beta) * (scale + 1) + shift, as one affine map of the normalised value n:
// g = gamma * (scale + 1), b = beta * (scale + 1) + shift   (transformer.py:390-392 then 413-417)
__global__ void k_fold_film(const float* __restrict__ film, int film_stride, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float* __restrict__ g, float* __restrict__ b, int out_stride, int T) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * kE) return;
    int t = idx / kE, c = idx % kE;
    float sc = film[(size_t)t * film_stride + c] + 1.0f;
    float sh = film[(size_t)t * film_stride + kE + c];
    g[(size_t)t * out_stride + c] = gamma[c] * sc;
    b[(size_t)t * out_stride + c] = fmaf(beta[c], sc, sh);
}

// (sigmoid(E[c]) * 2 - 1) * bit_scale    segmentation/mmseg/models/segmentors/ddp.py:236-237
__global__ void k_embed_lut(const float* __restrict__ emb, float* __restrict__ lut, int n, float bit_scale) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float s = sigmoidf_(emb[idx]);
    lut[idx] = __fmul_rn(__fadd_rn(__fmul_rn(s, 2.0f), -1.0f), bit_scale);
}

// ------------------------------------------------------------------------------------------------
// multi-scale deformable attention gather (1 level, 8 heads x 4 points, head dim 32)
// ------------------------------------------------------------------------------------------------
// out[row][n][32m + d] = sum_p a[n,m,p] * sum_corners w_c * V[row][corner_c][32m + d]
// The sampling positions were resolved by the sampling projection's epilogue (msda_resolve, common.cuh): this kernel
// is loads + FMAs only, no branches.  One warp per token; lane = (head-in-group = lane / 8, 4 channels); the warp walks
// two groups of 4 heads so that every warp-wide float4 load covers four whole 128-byte lines (one per head).
// launch_bounds(256, 6): 40 registers, 48 resident warps — 3 % faster than (256, 4) / 64 registers in tools/ubench_gather.cu
// (v7: 3 / 4 / 5 / 6 / 8 CTAs per SM = 0.310 / 0.262 / 0.261 / 0.254 / 0.292 ms)
__global__ void __launch_bounds__(256, 6)
k_msda_gather(const float* __restrict__ V, const uint32_t* __restrict__ rec, float* __restrict__ out,
              __half* __restrict__ out_hi, __half* __restrict__ out_lo, int N, int W, int total_tokens) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= total_tokens) return;
    int lane = threadIdx.x & 31;
    int row = warp / N;
    const uint32_t* rp = rec + (size_t)warp * kRecW;
#pragma unroll
    for (int hg = 0; hg < 2; ++hg) {
        const int m = hg * 4 + (lane >> 3);
        const int ch = m * kHeadDim + (lane & 7) * 4;
        const float* Vr = V + (size_t)row * N * kE + ch;
        const uint4 wd = *reinterpret_cast<const uint4*>(rp + m * 4);
        const float4 fx = *reinterpret_cast<const float4*>(rp + 32 + m * 4);
        const float4 fy = *reinterpret_cast<const float4*>(rp + 64 + m * 4);
        const float4 aw = *reinterpret_cast<const float4*>(rp + 96 + m * 4);
        const uint32_t w4[4] = {wd.x, wd.y, wd.z, wd.w};
        const float fx4[4] = {fx.x, fx.y, fx.z, fx.w}, fy4[4] = {fy.x, fy.y, fy.z, fy.w}, a4[4] = {aw.x, aw.y, aw.z, aw.w};
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        // all 16 corner loads of this head group are issued before any is used (memory-level parallelism)
        float4 vv[kPoints][4];
#pragma unroll
        for (int p = 0; p < kPoints; ++p) {
            const uint32_t wv = w4[p];
            const int base = (int)(wv & 0x03FFFFFFu);
            const int dx = (int)((wv >> 26) & 1u);
            const int dy = ((wv >> 27) & 1u) ? W : 0;
            vv[p][0] = __ldg(reinterpret_cast<const float4*>(Vr + (size_t)base * kE));
            vv[p][1] = __ldg(reinterpret_cast<const float4*>(Vr + (size_t)(base + dx) * kE));
            vv[p][2] = __ldg(reinterpret_cast<const float4*>(Vr + (size_t)(base + dy) * kE));
            vv[p][3] = __ldg(reinterpret_cast<const float4*>(Vr + (size_t)(base + dy + dx) * kE));
        }
#pragma unroll
        for (int p = 0; p < kPoints; ++p) {
            const uint32_t wv = w4[p];
            const float wx1 = fx4[p], wy1 = fy4[p], wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
            const float c00 = (wv & (1u << 28)) ? wy0 * wx0 : 0.f;
            const float c01 = (wv & (1u << 29)) ? wy0 * wx1 : 0.f;
            const float c10 = (wv & (1u << 30)) ? wy1 * wx0 : 0.f;
            const float c11 = (wv & (1u << 31)) ? wy1 * wx1 : 0.f;
            const float4 v00 = vv[p][0], v01 = vv[p][1], v10 = vv[p][2], v11 = vv[p][3];
            float s0 = c00 * v00.x, s1 = c00 * v00.y, s2 = c00 * v00.z, s3 = c00 * v00.w;
            s0 = fmaf(c01, v01.x, s0); s1 = fmaf(c01, v01.y, s1); s2 = fmaf(c01, v01.z, s2); s3 = fmaf(c01, v01.w, s3);
            s0 = fmaf(c10, v10.x, s0); s1 = fmaf(c10, v10.y, s1); s2 = fmaf(c10, v10.z, s2); s3 = fmaf(c10, v10.w, s3);
            s0 = fmaf(c11, v11.x, s0); s1 = fmaf(c11, v11.y, s1); s2 = fmaf(c11, v11.z, s2); s3 = fmaf(c11, v11.w, s3);
            acc[0] = fmaf(a4[p], s0, acc[0]); acc[1] = fmaf(a4[p], s1, acc[1]);
            acc[2] = fmaf(a4[p], s2, acc[2]); acc[3] = fmaf(a4[p], s3, acc[3]);
        }
        const size_t o = (size_t)warp * kE + ch;
        if (out) *reinterpret_cast<float4*>(out + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (out_hi) {
            const float a0 = acc[0] * kSplitScale, a1 = acc[1] * kSplitScale, a2 = acc[2] * kSplitScale, a3 = acc[3] * kSplitScale;
            __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
            *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
            if (out_lo) {
                const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
                __half2 l01 = __floats2half2_rn(a0 - b01.x, a1 - b01.y), l23 = __floats2half2_rn(a2 - b23.x, a3 - b23.y);
                *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// step epilogues
// ------------------------------------------------------------------------------------------------
// depth: q[row][n][c] = cond[b][n][c] + w_m[c] * d_t[row][n]      (down = ConvModule(257 -> 256), rank-1 half)
__global__ void k_depth_head_in(const float* __restrict__ cond, const float* __restrict__ wm,
                                const float* __restrict__ state, float* __restrict__ q, __half* __restrict__ q_hi,
                                __half* __restrict__ q_lo, int N, int R, int total_tokens) {
    // two float4 per thread so that the fp16 planes can be stored 16 bytes at a time
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // 8-float index
    size_t tok = idx / (kE / 8);
    if (tok >= (size_t)total_tokens) return;
    int c = (int)(idx % (kE / 8)) * 8;
    int row = (int)(tok / N), n = (int)(tok % N);
    int b = row / R;
    float d = state[tok];
    const float* cp = cond + ((size_t)b * N + n) * kE + c;
    float4 c0 = *reinterpret_cast<const float4*>(cp), c1 = *reinterpret_cast<const float4*>(cp + 4);
    float4 w0 = *reinterpret_cast<const float4*>(wm + c), w1 = *reinterpret_cast<const float4*>(wm + c + 4);
    float o[8] = {fmaf(w0.x, d, c0.x), fmaf(w0.y, d, c0.y), fmaf(w0.z, d, c0.z), fmaf(w0.w, d, c0.w),
                  fmaf(w1.x, d, c1.x), fmaf(w1.y, d, c1.y), fmaf(w1.z, d, c1.z), fmaf(w1.w, d, c1.w)};
    *reinterpret_cast<float4*>(q + tok * kE + c) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(q + tok * kE + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
    if (q_hi) split8_store(o, q_hi + tok * kE + c, q_lo ? q_lo + tok * kE + c : nullptr);
}

struct SegStepParams {
    const float* logits;   // [rows][N][C]
    float* state;          // [rows][N][256] in/out
    __half* state_hi;      // optional fp16 planes of the new state (A operand of the next head-in GEMM)
    __half* state_lo;
    float* accum;          // [B][N][C] running sum of softmax prob (accumulation) or of last-step logits
    const float* lut;      // [(C+1)][256]
    int N, R, C, B;
    float alpha, sigma, alpha_next, sigma_next;
    int accumulate_prob;   // accumulation=True: add softmax(logit) every step
    int add_logits;        // accumulation=False and last step: add raw logits
    // ddpm_sample (ddp.py:248-290): m <- alpha' * (m * (1 - c) / alpha + c * m_hat) + std * noise
    int ddpm;
    float one_minus_c, c, std;
    const float* step_noise;   // this step's noise, (rows, 256, h, w) NCHW, or null (t_next == 0: no noise)
    // per-pixel uncertainty (ddp_set_uncertainty_outputs): class of every (sample, token) at this step, and the running
    // count of class changes between consecutive steps summed over the R samples
    uint8_t* row_cls;          // [rows][N] or null
    int32_t* changes;          // [B][N] or null
    int first_step;
};

// One warp per image token (b, n), looping over the R stochastic samples so that the accumulation
// order is fixed.  ddp.py:235-243: argmax -> embedding -> squash -> DDIM update; softmax accumulate.
// KMAX = ceil(C / 32) rounded up to 1, 2, 4 or 8: the class loops are unrolled over lane + 32 k, and with 19 classes seven of
// eight iterations (exp, IEEE division, predicated loads / stores) would be issued for nothing: the kernel is issue-bound.
template <int KMAX>
__global__ void __launch_bounds__(256) k_seg_step(SegStepParams p) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= p.B * p.N) return;
    int b = warp / p.N, n = warp - b * p.N;
    const int C = p.C;
    const float sig = fmaxf(p.sigma, 1e-8f);
    for (int r = 0; r < p.R; ++r) {
        size_t tok = ((size_t)(b * p.R + r)) * p.N + n;
        const float* lg = p.logits + tok * C;
        // the old state does not depend on the argmax: request it together with the logits (one DRAM latency, not two)
        float* st = p.state + tok * kE + lane * 8;
        const float4 mt2[2] = {*reinterpret_cast<const float4*>(st), *reinterpret_cast<const float4*>(st + 4)};
        float v[KMAX];                   // C <= 32 KMAX
        float best = -INFINITY;
        int besti = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            int c = lane + k * 32;
            v[k] = (c < C) ? lg[c] : -INFINITY;
            if (c < C && (v[k] > best)) { best = v[k]; besti = c; }   // increasing c: first max wins
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (p.accumulate_prob) {
            float e[KMAX], s = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                int c = lane + k * 32;
                e[k] = (c < C) ? expf(v[k] - best) : 0.f;
                s += e[k];
            }
            s = warp_sum(s);
            float* ac = p.accum + ((size_t)b * p.N + n) * C;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                int c = lane + k * 32;
                if (c < C) ac[c] += e[k] / s;
            }
        } else if (p.add_logits) {
            float* ac = p.accum + ((size_t)b * p.N + n) * C;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                int c = lane + k * 32;
                if (c < C) ac[c] += v[k];
            }
        }
        if (p.row_cls && lane == 0) {
            if (p.changes && !p.first_step && p.row_cls[tok] != (uint8_t)besti) p.changes[(size_t)b * p.N + n] += 1;
            p.row_cls[tok] = (uint8_t)besti;
        }
        // m <- m_hat * alpha' + ((m - alpha * m_hat) / max(sigma, 1e-8)) * sigma'
        const float* lr = p.lut + (size_t)besti * kE + lane * 8;
        float nv[8];
#pragma unroll
        for (int h4 = 0; h4 < 2; ++h4) {
            float4 mh = *reinterpret_cast<const float4*>(lr + h4 * 4);
            const float4 mt = mt2[h4];
            float4 o;
            o.x = __fadd_rn(__fmul_rn(mh.x, p.alpha_next), __fmul_rn(__fdiv_rn(__fadd_rn(mt.x, -__fmul_rn(p.alpha, mh.x)), sig), p.sigma_next));
            o.y = __fadd_rn(__fmul_rn(mh.y, p.alpha_next), __fmul_rn(__fdiv_rn(__fadd_rn(mt.y, -__fmul_rn(p.alpha, mh.y)), sig), p.sigma_next));
            o.z = __fadd_rn(__fmul_rn(mh.z, p.alpha_next), __fmul_rn(__fdiv_rn(__fadd_rn(mt.z, -__fmul_rn(p.alpha, mh.z)), sig), p.sigma_next));
            o.w = __fadd_rn(__fmul_rn(mh.w, p.alpha_next), __fmul_rn(__fdiv_rn(__fadd_rn(mt.w, -__fmul_rn(p.alpha, mh.w)), sig), p.sigma_next));
            if (p.ddpm) {
                const float mtv[4] = {mt.x, mt.y, mt.z, mt.w}, mhv[4] = {mh.x, mh.y, mh.z, mh.w};
                float ov[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int ch = lane * 8 + h4 * 4 + e;
                    const float nz = p.step_noise ? p.step_noise[((size_t)(b * p.R + r) * kE + ch) * p.N + n] : 0.f;
                    const float mean = __fmul_rn(p.alpha_next, __fadd_rn(__fdiv_rn(__fmul_rn(mtv[e], p.one_minus_c), p.alpha),
                                                                         __fmul_rn(p.c, mhv[e])));
                    ov[e] = __fadd_rn(mean, __fmul_rn(p.std, nz));
                }
                o = make_float4(ov[0], ov[1], ov[2], ov[3]);
            }
            *reinterpret_cast<float4*>(st + h4 * 4) = o;
            nv[h4 * 4 + 0] = o.x; nv[h4 * 4 + 1] = o.y; nv[h4 * 4 + 2] = o.z; nv[h4 * 4 + 3] = o.w;
        }
        if (p.state_hi) {
            size_t o8 = tok * kE + lane * 8;
            split8_store(nv, p.state_hi + o8, p.state_lo ? p.state_lo + o8 : nullptr);
        }
    }
}

// out[b][c][n] = accum[b][n][c] / count ; cls[b][n] = argmax_c out      (ddp.py:244-245 mean over dim 0)
// spread[b][n] (optional) = 1 - (#samples r whose LAST-step class equals cls[b][n]) / R      (ddp_set_uncertainty_outputs)
__global__ void k_seg_finalize(const float* __restrict__ accum, float* __restrict__ out, int32_t* __restrict__ cls,
                               int N, int C, float count, const uint8_t* __restrict__ row_cls = nullptr,
                               float* __restrict__ spread = nullptr, int R = 1) {
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* s = accum + (size_t)b * N * C;
    float* d = out + (size_t)b * C * N;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int n = n0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (n < N && c < C) ? __fdiv_rn(s[(size_t)n * C + c], count) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, n = n0 + threadIdx.x;
        if (n < N && c < C) d[(size_t)c * N + n] = tile[threadIdx.x][i];
    }
    if ((cls != nullptr || spread != nullptr) && blockIdx.y == 0) {
        // one thread per token of this tile: argmax over all classes
        int t = threadIdx.y * 32 + threadIdx.x;
        if (t < 32) {
            int n = n0 + t;
            if (n < N) {
                float best = -INFINITY; int bi = 0;
                for (int c = 0; c < C; ++c) {
                    float v = __fdiv_rn(s[(size_t)n * C + c], count);
                    if (v > best) { best = v; bi = c; }
                }
                if (cls) cls[(size_t)b * N + n] = bi;
                if (spread) {
                    int agree = 0;
                    for (int r = 0; r < R; ++r) agree += row_cls[((size_t)b * R + r) * N + n] == (uint8_t)bi;
                    spread[(size_t)b * N + n] = __fadd_rn(1.0f, -__fdiv_rn((float)agree, (float)R));
                }
            }
        }
    }
}

// Post-loop tail: x4 bilinear resize (align_corners=False) + softmax + argmax of
// EncoderDecoder.whole_inference / inference / simple_test (encoder_decoder.py:229-304, ddp.py:124-128) in one pass:
// softmax is monotonic, so the class map is argmax_C of the resized logits; the (B,C,H,W) tensor is never materialised.
// Interpolation follows ATen's upsample_bilinear2d: src = (dst + .5) * (in / out) - .5 clamped at 0, lambda = src - floor.
__global__ void k_resize_argmax(const float* __restrict__ logits, uint8_t* __restrict__ cls, int C, int h, int w, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    if (x >= W) return;
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    float fy = __fadd_rn(__fmul_rn(sy, (float)y + 0.5f), -0.5f); fy = fy < 0.f ? 0.f : fy;
    float fx = __fadd_rn(__fmul_rn(sx, (float)x + 0.5f), -0.5f); fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.0f - ly, hx = 1.0f - lx;
    const float* base = logits + (size_t)b * C * h * w;
    float best = -INFINITY;
    int bi = 0;
    for (int c = 0; c < C; ++c) {
        const float* p = base + (size_t)c * h * w;
        const float v = __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, p[y0 * w + x0]), __fmul_rn(lx, p[y0 * w + x1]))),
                                  __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, p[y1 * w + x0]), __fmul_rn(lx, p[y1 * w + x1]))));
        if (v > best) { best = v; bi = c; }
    }
    cls[((size_t)b * H + y) * W + x] = (uint8_t)bi;
}

// Whole inference() tail of one (possibly augmented) view (encoder_decoder.py:229-283, ddp.py:124-128) in one pass:
//   logits (B,C,h,w) --bilinear--> (img_h,img_w) [encode_decode] --crop to (crop_h,crop_w), bilinear--> (H,W) [whole_inference,
//   rescale=True; skipped when rescale == 0] --softmax over C--> flip back --> probs (B,C,H,W), overwritten or accumulated
//   (aug_test sums the views).  The two resizes are evaluated nested, each with ATen's upsample_bilinear2d arithmetic
//   (align_corners=False), so the intermediate image is never materialised but every intermediate value is rounded as
//   the reference rounds it.
struct TailParams {
    const float* logits; float* probs;
    int C, h, w, img_h, img_w, crop_h, crop_w, H, W;
    int rescale, flip, accumulate;       // flip: 0 none, 1 horizontal, 2 vertical
};
__device__ __forceinline__ void tail_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float f = __fadd_rn(__fmul_rn(scale, (float)dst + 0.5f), -0.5f);
    f = f < 0.f ? 0.f : f;
    i0 = (int)f;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = f - (float)i0;
}
__device__ __forceinline__ float tail_lerp4(float v00, float v01, float v10, float v11, float ly, float lx) {
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    return __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01))),
                     __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11))));
}
// value of the img-size map (first resize) at (Y, X) for one class plane
__device__ __forceinline__ float tail_stage1(const float* __restrict__ p, int Y, int X, const TailParams& t) {
    int y0, y1, x0, x1; float ly, lx;
    tail_src(Y, (float)t.h / (float)t.img_h, t.h, y0, y1, ly);
    tail_src(X, (float)t.w / (float)t.img_w, t.w, x0, x1, lx);
    return tail_lerp4(p[y0 * t.w + x0], p[y0 * t.w + x1], p[y1 * t.w + x0], p[y1 * t.w + x1], ly, lx);
}
__global__ void __launch_bounds__(128) k_tail_probs(TailParams t) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= t.W) return;
    int Y0 = y, Y1 = y, X0 = x, X1 = x; float LY = 0.f, LX = 0.f;
    if (t.rescale) {
        tail_src(y, (float)t.crop_h / (float)t.H, t.crop_h, Y0, Y1, LY);
        tail_src(x, (float)t.crop_w / (float)t.W, t.crop_w, X0, X1, LX);
    }
    const float* base = t.logits + (size_t)b * t.C * t.h * t.w;
    auto value = [&](int c) -> float {
        const float* p = base + (size_t)c * t.h * t.w;
        if (!t.rescale) return tail_stage1(p, y, x, t);
        return tail_lerp4(tail_stage1(p, Y0, X0, t), tail_stage1(p, Y0, X1, t), tail_stage1(p, Y1, X0, t), tail_stage1(p, Y1, X1, t), LY, LX);
    };
    float mx = -INFINITY;
    for (int c = 0; c < t.C; ++c) mx = fmaxf(mx, value(c));
    float sum = 0.f;
    for (int c = 0; c < t.C; ++c) sum += expf(value(c) - mx);
    const int xo = t.flip == 1 ? t.W - 1 - x : x, yo = t.flip == 2 ? t.H - 1 - y : y;
    float* o = t.probs + ((size_t)b * t.C * t.H + yo) * t.W + xo;
    for (int c = 0; c < t.C; ++c) {
        const float pr = __fdiv_rn(expf(value(c) - mx), sum);
        float* oc = o + (size_t)c * t.H * t.W;
        *oc = t.accumulate ? *oc + pr : pr;
    }
}
// class map of accumulated probabilities: argmax over C (first maximum wins, as torch.argmax)
__global__ void __launch_bounds__(128) k_probs_argmax(const float* __restrict__ probs, uint8_t* __restrict__ cls, int C, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const float* p = probs + ((size_t)b * C * H + y) * W + x;
    float best = -INFINITY; int bi = 0;
    for (int c = 0; c < C; ++c) {
        const float v = p[(size_t)c * H * W];
        if (v > best) { best = v; bi = c; }
    }
    cls[((size_t)b * H + y) * W + x] = (uint8_t)bi;
}

struct DepthStepParams {
    const float* taps;     // [rows][N][16]: per-token dot products with the 9 conv3x3 taps (cols 0..8)
    float* state;          // [rows][N] depth_t in/out
    float* pred;           // [rows][N] relu(conv)+min_depth (tap / last-step output), may be null
    float* out;            // [B][N] final mean over r, clamped (written when last != 0)
    int H, W, R, B;
    float conv_bias, min_depth, max_depth, bit_scale;
    float gamma_now, gamma_next;
    int last;
    float* spread;         // [B][N] optional: population standard deviation over the R samples of the last-step prediction
};

// depth head tail + ddim_step.  depth/depth/models/decode_heads/decode_head.py:233-270 (relu(conv3x3)+min_depth),
// depth/depth/models/depther/ddp.py:220-227, 240-246.  One thread per image token, looping over r.
__global__ void k_depth_step(DepthStepParams p) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int N = p.H * p.W;
    if (idx >= p.B * N) return;
    int b = idx / N, n = idx - b * N;
    int i = n / p.W, j = n - i * p.W;
    float sum = 0.f;
    const float sa_now = sqrtf(p.gamma_now), sa_next = sqrtf(p.gamma_next);
    const float inv_sn = __fdiv_rn(1.0f, sqrtf(__fadd_rn(1.0f, -p.gamma_now)));
    const float sn_next = sqrtf(__fadd_rn(1.0f, -p.gamma_next));
    for (int r = 0; r < p.R; ++r) {
        size_t base = ((size_t)(b * p.R + r)) * N;
        float acc = 0.f;
#pragma unroll
        for (int di = -1; di <= 1; ++di)
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                int ii = i + di, jj = j + dj;
                if (ii < 0 || ii >= p.H || jj < 0 || jj >= p.W) continue;
                acc += p.taps[(base + (size_t)ii * p.W + jj) * 16 + (di + 1) * 3 + (dj + 1)];
            }
        float d = fmaxf(acc + p.conv_bias, 0.f) + p.min_depth;
        if (p.pred) p.pred[base + n] = d;
        sum += d;
        float dn = __fdiv_rn(__fadd_rn(d, -p.min_depth), __fadd_rn(p.max_depth, -p.min_depth));
        dn = __fmul_rn(__fadd_rn(__fmul_rn(dn, 2.0f), -1.0f), p.bit_scale);
        dn = fminf(fmaxf(dn, -p.bit_scale), p.bit_scale);
        float xt = p.state[base + n];
        float eps = __fmul_rn(inv_sn, __fadd_rn(xt, -__fmul_rn(sa_now, dn)));
        p.state[base + n] = __fadd_rn(__fmul_rn(sa_next, dn), __fmul_rn(sn_next, eps));
    }
    if (p.last) {
        float o = __fdiv_rn(sum, (float)p.R);
        p.out[idx] = fminf(fmaxf(o, p.min_depth), p.max_depth);
        if (p.spread) {
            float ss = 0.f;
            for (int r = 0; r < p.R; ++r) {
                const float e = __fadd_rn(p.pred[((size_t)(b * p.R + r)) * N + n], -o);
                ss = __fadd_rn(ss, __fmul_rn(e, e));
            }
            p.spread[idx] = sqrtf(__fdiv_rn(ss, (float)p.R));
        }
    }
}

}  // namespace ddp
