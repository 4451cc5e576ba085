// fp32 CUDA-core GEMM with fused epilogues: C[M, Nout] = A[M, K] * Wt[K, ldw] (+ epilogue).
//
// This is the DDP_GEMM_FP32 arithmetic mode: plain fp32 FMA, the same precision class as the
// reference's fp32 path (fp16_enabled=False, segmentation/mmseg/models/decode_heads/decode_head.py:138).
// Tokens are rows; a CTA owns BM complete rows when BN == 256, which lets LayerNorm / FiLM /
// softmax-of-4 run in the epilogue on registers (one warp holds 8 full rows).
//
// Tile: BM x BN x 16, 256 threads = 8 warps; warp `ty` owns rows ty*TM..+TM-1, lane `tx` owns
// columns {tx*4..tx*4+3} (+128 for BN == 256).  Weights are stored transposed ([K][ldw], ldw a
// multiple of BN) so B-tile loads are coalesced float4 and need no bounds checks.
#pragma once
#include "common.cuh"
#include "neck_plan.h"

namespace ddp {

__device__ __forceinline__ int frag_col(int c, int tx) {
    return tx * 4 + (c & 3) + (c >> 2) * 128;
}

// ---- epilogues -------------------------------------------------------------------------------
// operator()(acc, grow0, tx, n0): acc[r][c] is row grow0 + r, column n0 + frag_col(c, tx).

// out[row*ldc + col] = acc + bias[col]   (col < ncols)
struct EpiBias {
    float* out; const float* bias; int ldc; int ncols; int M;
    template <int TM, int TN>
    __device__ __forceinline__ void operator()(float (&acc)[TM][TN], int grow0, int tx, int n0) const {
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            int row = grow0 + r;
            if (row >= M) continue;
#pragma unroll
            for (int c = 0; c < TN; ++c) {
                int col = n0 + frag_col(c, tx);
                if (col < ncols) out[(size_t)row * ldc + col] = acc[r][c] + (bias ? bias[col] : 0.f);
            }
        }
    }
};

// q[row][col] = acc + cond[image(row)][n(row)][col]      (transform: W_x x + b is step-invariant)
struct EpiAddCond {
    float* out; const float* cond; int N; int R; int M;
    template <int TM, int TN>
    __device__ __forceinline__ void operator()(float (&acc)[TM][TN], int grow0, int tx, int n0) const {
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            int row = grow0 + r;
            if (row >= M) continue;
            int n = row % N;
            int b = (row / N) / R;
            const float* cp = cond + ((size_t)b * N + n) * kE;
#pragma unroll
            for (int c4 = 0; c4 < TN / 4; ++c4) {
                int col = n0 + tx * 4 + c4 * 128;
                float4 cv = *reinterpret_cast<const float4*>(cp + col);
                float4 o = make_float4(acc[r][c4 * 4 + 0] + cv.x, acc[r][c4 * 4 + 1] + cv.y,
                                       acc[r][c4 * 4 + 2] + cv.z, acc[r][c4 * 4 + 3] + cv.w);
                *reinterpret_cast<float4*>(out + (size_t)row * kE + col) = o;
            }
        }
    }
};

// sampling offsets + attention weights: s = acc + pew[n][col] (pew = PE * W^T + bias, shape-only),
// softmax over each head's 4 points for cols 64..95 (vmmcv/ops/multi_scale_deform_attn.py:319-328).
// BN == 128: lane tx holds exactly one group of 4 consecutive columns.
struct EpiSampling {
    float* out;            // optional [M][96] offsets | softmaxed weights (test tap), may be null
    uint32_t* rec;         // [M][kRecW] sampling records consumed by k_msda_gather
    const float* pew; int N; int H; int W; int M;
    template <int TM, int TN>
    __device__ __forceinline__ void operator()(float (&acc)[TM][TN], int grow0, int tx, int n0) const {
        static_assert(TN == 4, "EpiSampling needs BN == 128");
        int col = tx * 4;
        if (col >= kSampW) return;
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            int row = grow0 + r;
            if (row >= M) continue;
            int n = row % N;
            float4 pv = *reinterpret_cast<const float4*>(pew + (size_t)n * kSampW + col);
            float v0 = acc[r][0] + pv.x, v1 = acc[r][1] + pv.y, v2 = acc[r][2] + pv.z, v3 = acc[r][3] + pv.w;
            uint32_t* rp = rec + (size_t)row * kRecW;
            if (col >= 64) {
                float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
                v0 = expf(v0 - mx); v1 = expf(v1 - mx); v2 = expf(v2 - mx); v3 = expf(v3 - mx);
                float s = (v0 + v1) + (v2 + v3);
                v0 /= s; v1 /= s; v2 /= s; v3 /= s;
                *reinterpret_cast<float4*>(rp + 96 + (col - 64)) = make_float4(v0, v1, v2, v3);
            } else {
                // columns col..col+3 = (x, y) offsets of sampling points k = col/2 and k+1
                const int i = n / W, j = n - i * W, k = col >> 1;
                const float refx = __fdiv_rn((float)j + 0.5f, (float)W), refy = __fdiv_rn((float)i + 0.5f, (float)H);
                const float rW = __frcp_rn((float)W), rH = __frcp_rn((float)H);
                uint32_t w0, w1; float fx0, fy0, fx1, fy1;
                msda_resolve(v0, v1, refx, refy, rW, rH, H, W, w0, fx0, fy0);
                msda_resolve(v2, v3, refx, refy, rW, rH, H, W, w1, fx1, fy1);
                *reinterpret_cast<uint2*>(rp + k) = make_uint2(w0, w1);
                *reinterpret_cast<float2*>(rp + 32 + k) = make_float2(fx0, fx1);
                *reinterpret_cast<float2*>(rp + 64 + k) = make_float2(fy0, fy1);
            }
            if (out) *reinterpret_cast<float4*>(out + (size_t)row * kSampW + col) = make_float4(v0, v1, v2, v3);
        }
    }
};

// hid = gelu(acc + b1)
struct EpiGelu {
    float* out; const float* bias; int ldc; int M;
    template <int TM, int TN>
    __device__ __forceinline__ void operator()(float (&acc)[TM][TN], int grow0, int tx, int n0) const {
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            int row = grow0 + r;
            if (row >= M) continue;
#pragma unroll
            for (int c4 = 0; c4 < TN / 4; ++c4) {
                int col = n0 + tx * 4 + c4 * 128;
                float4 bv = *reinterpret_cast<const float4*>(bias + col);
                float4 o = make_float4(gelu_erf(acc[r][c4 * 4 + 0] + bv.x), gelu_erf(acc[r][c4 * 4 + 1] + bv.y),
                                       gelu_erf(acc[r][c4 * 4 + 2] + bv.z), gelu_erf(acc[r][c4 * 4 + 3] + bv.w));
                *reinterpret_cast<float4*>(out + (size_t)row * ldc + col) = o;
            }
        }
    }
};

// y = LayerNorm(acc + bias + resid) * gamma + beta ; optional FiLM y*(scale+1)+shift
// (post-norm residual blocks + time modulation, segmentation/mmseg/models/utils/transformer.py:390-417).
// BN == 256: the 32 lanes of a warp hold complete rows.  `out` may alias `resid`.
struct EpiResidualLN {
    float* out; const float* resid; const float* bias; const float* gamma; const float* beta;
    const float* film;   // [512] = scale | shift, or nullptr
    int M;
    template <int TM, int TN>
    __device__ __forceinline__ void operator()(float (&acc)[TM][TN], int grow0, int tx, int n0) const {
        static_assert(TN == 8, "EpiResidualLN needs BN == 256");
        float4 b0 = *reinterpret_cast<const float4*>(bias + tx * 4);
        float4 b1 = *reinterpret_cast<const float4*>(bias + 128 + tx * 4);
        float4 g0 = *reinterpret_cast<const float4*>(gamma + tx * 4);
        float4 g1 = *reinterpret_cast<const float4*>(gamma + 128 + tx * 4);
        float4 e0 = *reinterpret_cast<const float4*>(beta + tx * 4);
        float4 e1 = *reinterpret_cast<const float4*>(beta + 128 + tx * 4);
        float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
        float sc[8], sh[8];
        if (film) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                int col = frag_col(c, tx);
                sc[c] = film[col] + 1.0f;
                sh[c] = film[kE + col];
            }
        }
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            int row = grow0 + r;          // warp-uniform
            int lrow = row < M ? row : M - 1;
            float4 r0 = *reinterpret_cast<const float4*>(resid + (size_t)lrow * kE + tx * 4);
            float4 r1 = *reinterpret_cast<const float4*>(resid + (size_t)lrow * kE + 128 + tx * 4);
            float x[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) { x[c] = (acc[r][c] + bb[c]) + x[c]; s += x[c]; }
            float mean = warp_sum(s) * (1.0f / kE);
            float v = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) { float d = x[c] - mean; v += d * d; }
            float var = warp_sum(v) * (1.0f / kE);
            float rstd = 1.0f / sqrtf(var + 1e-5f);
            float y[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                y[c] = (x[c] - mean) * rstd * gg[c] + ee[c];
                if (film) y[c] = y[c] * sc[c] + sh[c];
            }
            if (row < M) {
                *reinterpret_cast<float4*>(out + (size_t)row * kE + tx * 4) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4*>(out + (size_t)row * kE + 128 + tx * 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
        }
    }
};

// ---- kernel ----------------------------------------------------------------------------------
// A_MODE == 0 (false): A is row-major [M][lda].
// A_MODE == 1 (true) : A is a stack of images in NCHW, element (m, k) = A[(m / n_img) * K * n_img + k * n_img + m % n_img]
//                      (reads the neck feature x / the backbone maps without a transposed copy).
// A_MODE == 2        : 3x3 convolution, padding 1, over token-major images [imgs][n_img][K / 9] of width `lda`:
//                      element (m, k) = channel k % (K/9) of the token shifted by tap k / (K/9), zero outside
//                      (neck::conv3_src_offset; the FPN output convs, segmentation/mmseg/models/necks/fpn.py:130-139).
template <int BM, int BN, int A_MODE, class Epi>
__global__ void __launch_bounds__(256, 2)
gemm_simt_kernel(const float* __restrict__ A, int lda, int n_img, const float* __restrict__ Wt, int ldw,
                 int M, int K, Epi epi) {
    constexpr int BK = 16;
    constexpr int TM = BM / 8;
    constexpr int TN = BN / 32;
    static_assert(BM == 64, "A-tile loader assumes BM == 64");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 31;
    const int ty = tid >> 5;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    float acc[TM][TN];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = 0.f;

    // per-thread A load coordinates
    float a_reg[4];
    float4 b_reg[BN / 64];
    const int a_row = tid >> 2, a_kq = (tid & 3) * 4;          // row-major loader
    const int k_k = tid >> 4, k_rq = (tid & 15) * 4;           // KN loader

    auto load_tiles = [&](int k0) {
        if (A_MODE == 0) {
            int row = m0 + a_row;
            if (row >= M) row = M - 1;
            float4 v = *reinterpret_cast<const float4*>(A + (size_t)row * lda + k0 + a_kq);
            a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w;
        } else if (A_MODE == 2) {
            // the 4 consecutive k of this thread share one tap (K / 9 is a multiple of 16) -> one float4 or zeros
            int row = m0 + a_row;
            if (row >= M) row = M - 1;
            const long long off = neck::conv3_src_offset(row, k0 + a_kq, n_img, lda, K / 9);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off >= 0) v = *reinterpret_cast<const float4*>(A + off);
            a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int row = m0 + k_rq + i;
                if (row >= M) row = M - 1;
                int img = row / n_img, n = row - img * n_img;
                a_reg[i] = A[((size_t)img * K + (k0 + k_k)) * n_img + n];
            }
        }
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
            int idx = tid + i * 256;
            int r = idx / (BN / 4), c4 = idx % (BN / 4);
            b_reg[i] = *reinterpret_cast<const float4*>(Wt + (size_t)(k0 + r) * ldw + n0 + c4 * 4);
        }
    };
    auto store_tiles = [&](int buf) {
        if (A_MODE != 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) As[buf][a_kq + i][a_row] = a_reg[i];
        } else {
            *reinterpret_cast<float4*>(&As[buf][k_k][k_rq]) = make_float4(a_reg[0], a_reg[1], a_reg[2], a_reg[3]);
        }
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
            int idx = tid + i * 256;
            int r = idx / (BN / 4), c4 = idx % (BN / 4);
            *reinterpret_cast<float4*>(&Bs[buf][r][c4 * 4]) = b_reg[i];
        }
    };

    const int nk = K / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int r4 = 0; r4 < TM / 4; ++r4) {
                float4 v = *reinterpret_cast<const float4*>(&As[cur][k][ty * TM + r4 * 4]);
                a[r4 * 4 + 0] = v.x; a[r4 * 4 + 1] = v.y; a[r4 * 4 + 2] = v.z; a[r4 * 4 + 3] = v.w;
            }
#pragma unroll
            for (int c4 = 0; c4 < TN / 4; ++c4) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4 + c4 * 128]);
                b[c4 * 4 + 0] = v.x; b[c4 * 4 + 1] = v.y; b[c4 * 4 + 2] = v.z; b[c4 * 4 + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        if (kt + 1 < nk) {
            store_tiles(cur ^ 1);
            __syncthreads();
        }
    }
    epi(acc, m0 + ty * TM, tx, n0);
}

template <int BN, int A_MODE, class Epi>
inline void launch_gemm_simt(const float* A, int lda, int n_img, const float* Wt, int ldw, int M, int K,
                                    int ncols_padded, const Epi& epi, cudaStream_t st) {
    dim3 grid((M + 63) / 64, ncols_padded / BN);
    gemm_simt_kernel<64, BN, A_MODE, Epi><<<grid, 256, 0, st>>>(A, lda, n_img, Wt, ldw, M, K, epi);
}

}  // namespace ddp
