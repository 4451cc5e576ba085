// Micro-benchmark for the deformable gather (k_msda_gather, 20 % of the decode step): is it bound by how its tokens are
// ORDERED over the SMs?  ncu of round 1: L1 wavefronts 60 %, L1 hit rate 57 % — the 43 % of 16 KB per token that miss L1
// come from L2 at ~6 TB/s, so locality in L1 may be the lever.  Variants, all with the kernel's exact arithmetic
// (outputs are compared bit for bit with the product kernel):
//   v0  the product kernel: warp = token in row-major order, 8 consecutive tokens per CTA, CTAs scheduled by the hardware
//   v1  persistent CTAs (k per SM), each walking a CONTIGUOUS range of 8 x 8 token tiles in row-major tile order
//   v2  as v1 with the tiles in Morton (Z) order, so that consecutive tiles of a CTA are 2-D neighbours
//   v3  as v2, head group 0 for the whole tile, then head group 1 (halves the L1 working set of a tile)
// Geometry = the headline shape (8 rows x 128 x 256 tokens); sampling records synthesised like the synthetic weights
// produce them: the reference's ring bias (direction = head, radius = point + 1) plus N(0, sigma) pixels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_gather tools/ubench_gather.cu
//   v4  TMA-staged: a CTA owns a TX x TY token tile; per head, the (TX + halo) x (TY + halo) window of that head's 32
//       channels of V (positioned at the minimum corner the tile's records reach for that head) is pulled into shared
//       memory by ONE 4-D TMA box load (out-of-range rows / columns are zero-filled), double / triple buffered over the 8
//       heads; the gather then reads shared memory (latency ~30 clk instead of L1-miss -> L2 ~600 clk); corners outside
//       the window fall back to the global load.  Same arithmetic, bit-identical output.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include "../ddp_b200/csrc/common.cuh"
#include "../ddp_b200/csrc/kernels.cuh"
#include "../ddp_b200/csrc/gemm_tc.cuh"
using namespace ddp;
using namespace ddp::tc;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

template <int TX, int TY, int PW, int PH, int NBUF, int THREADS, int PPB = 4>
__global__ void __launch_bounds__(THREADS, 1)
k_gather_smem(const __grid_constant__ CUtensorMap mapV, const float* __restrict__ V, const uint32_t* __restrict__ rec,
              __half* out_hi, __half* out_lo, int H, int W, int rows) {
    constexpr int kBufFloats = PW * PH * 32;
    constexpr int kTok = TX * TY;
    constexpr int kWarps = THREADS / 32;
    constexpr int kTokPerWarp = kTok / kWarps;
    static_assert(kTok % kWarps == 0 && kTokPerWarp % 4 == 0, "tile / warp split");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    float* bufs = reinterpret_cast<float*>(smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NBUF * kBufFloats * 4);
    int* org = reinterpret_cast<int*>(full + NBUF);          // [8][2] window origin (x, y) per head
    const int N = H * W;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY, n_tiles = rows * tiles_x * tiles_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 3, l8 = lane & 7;
    const float rW = 1.0f / (float)W;
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&full[b], 1);
        fence_barrier_init();
        fence_proxy_async();
        tma_prefetch_desc(&mapV);
    }
    __syncthreads();
    uint32_t phase[NBUF];
#pragma unroll
    for (int b = 0; b < NBUF; ++b) phase[b] = 0;
    int it = 0;           // running head-pass counter -> buffer index
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int row = t / (tiles_x * tiles_y), r = t - row * tiles_x * tiles_y;
        const int ty = r / tiles_x, tx = r - ty * tiles_x;
        // ---- window origin per head: minimum top-left corner over the tile's tokens and the head's 4 points ----
        if (threadIdx.x < 16) org[threadIdx.x] = 0x7fffffff;
        __syncthreads();
        for (int e = threadIdx.x; e < kTok * 8; e += THREADS) {
            const int tl = e >> 3, m = e & 7;
            const int i = ty * TY + tl / TX, j = tx * TX + tl % TX;
            if (i < H && j < W) {
                const uint4 wd = *reinterpret_cast<const uint4*>(rec + ((size_t)row * N + (size_t)i * W + j) * kRecW + m * 4);
                const uint32_t w4[4] = {wd.x, wd.y, wd.z, wd.w};
                int mx = 0x7fffffff, my = 0x7fffffff;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int base = (int)(w4[p] & 0x03FFFFFFu);
                    int y = (int)((float)base * rW);
                    int x = base - y * W;
                    if (x >= W) { x -= W; ++y; } else if (x < 0) { x += W; --y; }
                    mx = min(mx, x); my = min(my, y);
                }
                atomicMin(&org[m * 2], mx);
                atomicMin(&org[m * 2 + 1], my);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int m = 0; m < NBUF && m < 8; ++m) {
                const int b = (it + m) % NBUF;
                mbar_expect_tx(&full[b], kBufFloats * 4);
                tma_load_4d(bufs + (size_t)b * kBufFloats, &mapV, &full[b], m * 32, org[m * 2], org[m * 2 + 1], row);
            }
        }
        for (int m = 0; m < 8; ++m, ++it) {
            const int b = it % NBUF;
            const int ox = org[m * 2], oy = org[m * 2 + 1];
            mbar_wait(&full[b], phase[b]); phase[b] ^= 1;
            const float* sb = bufs + (size_t)b * kBufFloats + l8 * 4;
#pragma unroll 1
            for (int q = 0; q < kTokPerWarp / 4; ++q) {
                const int tl = warp * kTokPerWarp + q * 4 + grp;
                const int i = ty * TY + tl / TX, j = tx * TX + tl % TX;
                if (i >= H || j >= W) continue;
                const int token = row * N + i * W + j;
                const uint32_t* rp = rec + (size_t)token * kRecW;
                const int ch = m * kHeadDim + l8 * 4;
                const float* Vr = V + (size_t)row * N * kE + ch;
                const uint4 wd = *reinterpret_cast<const uint4*>(rp + m * 4);
                const float4 fx = *reinterpret_cast<const float4*>(rp + 32 + m * 4);
                const float4 fy = *reinterpret_cast<const float4*>(rp + 64 + m * 4);
                const float4 aw = *reinterpret_cast<const float4*>(rp + 96 + m * 4);
                const uint32_t w4[4] = {wd.x, wd.y, wd.z, wd.w};
                const float fx4[4] = {fx.x, fx.y, fx.z, fx.w}, fy4[4] = {fy.x, fy.y, fy.z, fy.w}, a4[4] = {aw.x, aw.y, aw.z, aw.w};
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                float4 vv[kPoints][4];
#pragma unroll
                for (int pb = 0; pb < kPoints; pb += PPB) {      // PPB points' corner loads in flight at a time
#pragma unroll
                for (int p = pb; p < pb + PPB; ++p) {
                    const uint32_t wv = w4[p];
                    const int base = (int)(wv & 0x03FFFFFFu);
                    const int dx = (int)((wv >> 26) & 1u);
                    const int dyb = (int)((wv >> 27) & 1u);
                    int y = (int)((float)base * rW);
                    int x = base - y * W;
                    if (x >= W) { x -= W; ++y; } else if (x < 0) { x += W; --y; }
                    const int lx = x - ox, ly = y - oy;
                    if (lx + dx < PW && ly + dyb < PH) {          // lx, ly >= 0 by construction of the origin
                        const float* s0 = sb + (ly * PW + lx) * 32;
                        vv[p][0] = *reinterpret_cast<const float4*>(s0);
                        vv[p][1] = *reinterpret_cast<const float4*>(s0 + dx * 32);
                        vv[p][2] = *reinterpret_cast<const float4*>(s0 + dyb * PW * 32);
                        vv[p][3] = *reinterpret_cast<const float4*>(s0 + (dyb * PW + dx) * 32);
                    } else {
                        const int dy = dyb ? W : 0;
                        vv[p][0] = __ldg(reinterpret_cast<const float4*>(Vr + (size_t)base * kE));
                        vv[p][1] = __ldg(reinterpret_cast<const float4*>(Vr + (size_t)(base + dx) * kE));
                        vv[p][2] = __ldg(reinterpret_cast<const float4*>(Vr + (size_t)(base + dy) * kE));
                        vv[p][3] = __ldg(reinterpret_cast<const float4*>(Vr + (size_t)(base + dy + dx) * kE));
                    }
                }
#pragma unroll
                for (int p = pb; p < pb + PPB; ++p) {
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
                }
                const size_t o = (size_t)token * kE + ch;
                const float a0 = acc[0] * kSplitScale, a1 = acc[1] * kSplitScale, a2 = acc[2] * kSplitScale, a3 = acc[3] * kSplitScale;
                __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
                *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
                const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
                __half2 l01 = __floats2half2_rn(a0 - b01.x, a1 - b01.y), l23 = __floats2half2_rn(a2 - b23.x, a3 - b23.y);
                *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
            }
            __syncthreads();                 // every warp is done with buffer b
            if (threadIdx.x == 0 && m + NBUF < 8) {
                mbar_expect_tx(&full[b], kBufFloats * 4);
                tma_load_4d(bufs + (size_t)b * kBufFloats, &mapV, &full[b], (m + NBUF) * 32, org[(m + NBUF) * 2], org[(m + NBUF) * 2 + 1], row);
            }
        }
    }
}

// the product kernel as it was before the unsigned-index / IMAD.WIDE addressing (A/B on the same box)
__global__ void __launch_bounds__(256, 4)
k_msda_gather_old(const float* __restrict__ V, const uint32_t* __restrict__ rec, float* __restrict__ out,
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


// v7: the product kernel with other occupancy targets (register caps 51 / 42 / 32 -> 40 / 48 / 64 warps per SM)
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k_msda_gather_lb(const float* __restrict__ V, const uint32_t* __restrict__ rec, float* __restrict__ out,
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


// one head group (4 heads) of one token: the body of k_msda_gather, verbatim arithmetic
__device__ __forceinline__ void gather_group(const float* __restrict__ V, const uint32_t* __restrict__ rec, __half* out_hi,
                                             __half* out_lo, int N, int W, int token, int hg, int lane) {
    const int row = token / N;
    const uint32_t* rp = rec + (size_t)token * kRecW;
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
    const size_t o = (size_t)token * kE + ch;
    const float a0 = acc[0] * kSplitScale, a1 = acc[1] * kSplitScale, a2 = acc[2] * kSplitScale, a3 = acc[3] * kSplitScale;
    __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
    *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
    const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
    __half2 l01 = __floats2half2_rn(a0 - b01.x, a1 - b01.y), l23 = __floats2half2_rn(a2 - b23.x, a3 - b23.y);
    *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
}

__host__ __device__ inline uint32_t morton_decode_x(uint32_t z) {          // even bits
    z &= 0x55555555u; z = (z | (z >> 1)) & 0x33333333u; z = (z | (z >> 2)) & 0x0F0F0F0Fu;
    z = (z | (z >> 4)) & 0x00FF00FFu; z = (z | (z >> 8)) & 0x0000FFFFu;
    return z;
}

// tile t (8 x 8 tokens) of image `row` -> its top-left token; tiles_x x tiles_y tiles per image (powers of two for Z order)
template <bool ZORDER>
__host__ __device__ inline void tile_origin(int t, int tiles_x, int tiles_y, int& row, int& ti, int& tj) {
    const int per_img = tiles_x * tiles_y;
    row = t / per_img;
    const int r = t - row * per_img;
    if (ZORDER) {
        // tiles_x = 2 tiles_y here (32 x 16): Z-order inside each 16 x 16 half, halves side by side
        const int half = r / (tiles_y * tiles_y), z = r - half * tiles_y * tiles_y;
        tj = (int)morton_decode_x((uint32_t)z) + half * tiles_y;
        ti = (int)morton_decode_x((uint32_t)z >> 1);
    } else {
        ti = r / tiles_x;
        tj = r - ti * tiles_x;
    }
}

// persistent CTAs: CTA c walks tiles [c * n / G, (c + 1) * n / G); 8 warps, warp w takes tile rows w (8 tokens each)
template <bool ZORDER, bool GROUP_MAJOR>
__global__ void __launch_bounds__(256, 4)
k_gather_tiled(const float* __restrict__ V, const uint32_t* __restrict__ rec, __half* out_hi, __half* out_lo, int H, int W,
               int n_tiles) {
    const int N = H * W, tiles_x = W / 8, tiles_y = H / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = (long long)blockIdx.x * n_tiles / gridDim.x, t1 = (long long)(blockIdx.x + 1) * n_tiles / gridDim.x;
    for (int t = (int)t0; t < (int)t1; ++t) {
        int row, ti, tj;
        tile_origin<ZORDER>(t, tiles_x, tiles_y, row, ti, tj);
        const int tok0 = row * N + (ti * 8 + warp) * W + tj * 8;
        if (GROUP_MAJOR) {
            for (int hg = 0; hg < 2; ++hg) {
                for (int x = 0; x < 8; ++x) gather_group(V, rec, out_hi, out_lo, N, W, tok0 + x, hg, lane);
                __syncthreads();           // the whole tile finishes a head group before the next one starts
            }
        } else {
            for (int x = 0; x < 8; ++x) {
                gather_group(V, rec, out_hi, out_lo, N, W, tok0 + x, 0, lane);
                gather_group(V, rec, out_hi, out_lo, N, W, tok0 + x, 1, lane);
            }
        }
    }
}

// v5: ONE 1024-thread CTA per SM (32 warps = the residency of the product kernel, 64 registers), walking TX x TY = 32-token
// tiles in row-major tile order, one token per warp, head group 0 for the whole tile, barrier, head group 1.  The four
// resident CTAs of the product kernel sit on four unrelated image rows (CTAs are dealt round-robin over the SMs), so an
// SM's L1 sees four 11-row neighbourhoods at once (57 % hits); here it sees ONE (TY + 10) x (TX + 10) neighbourhood of
// half the heads at a time (~126 KB for 8 x 4).
template <int TX, int TY, bool GROUP_MAJOR>
__global__ void __launch_bounds__(1024, 1)
k_gather_patch(const float* __restrict__ V, const uint32_t* __restrict__ rec, __half* out_hi, __half* out_lo, int H, int W, int rows) {
    static_assert(TX * TY == 32, "one token per warp");
    const int N = H * W, tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY, n_tiles = rows * tiles_x * tiles_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = (long long)blockIdx.x * n_tiles / gridDim.x, t1 = (long long)(blockIdx.x + 1) * n_tiles / gridDim.x;
    for (int t = (int)t0; t < (int)t1; ++t) {
        const int row = t / (tiles_x * tiles_y), r = t - row * tiles_x * tiles_y;
        const int ty = r / tiles_x, tx = r - ty * tiles_x;
        const int i = ty * TY + warp / TX, j = tx * TX + warp % TX;
        const bool ok = i < H && j < W;
        const int token = row * N + i * W + j;
        if (GROUP_MAJOR) {
            if (ok) gather_group(V, rec, out_hi, out_lo, N, W, token, 0, lane);
            __syncthreads();
            if (ok) gather_group(V, rec, out_hi, out_lo, N, W, token, 1, lane);
            __syncthreads();
        } else {
            if (ok) { gather_group(V, rec, out_hi, out_lo, N, W, token, 0, lane); gather_group(V, rec, out_hi, out_lo, N, W, token, 1, lane); }
        }
    }
}

// v6: the product kernel's geometry (warp = token, 256 threads) with register-free L1 prefetches: the SASS of the product
// kernel runs six dependent memory round trips per token (records, two batches of 8 corner loads, store; twice) because
// 64 registers hold only 8 corner loads at a time.  Here the corner lines of head group 1 (MODE & 1) and of the second
// batch of head group 0 (MODE & 2) are prefetched into L1 (prefetch.global.L1, no destination register) while head group
// 0's first batch is in flight.
__device__ __forceinline__ void pf_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int MODE>
__global__ void __launch_bounds__(256, 4)
k_gather_pf(const float* __restrict__ V, const uint32_t* __restrict__ rec, __half* out_hi, __half* out_lo, int N, int W, int total_tokens) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= total_tokens) return;
    int lane = threadIdx.x & 31;
    int row = warp / N;
    const uint32_t* rp = rec + (size_t)warp * kRecW;
    auto prefetch_group = [&](int hg, int p0, int p1) {
        const int m = hg * 4 + (lane >> 3);
        const float* Vr = V + (size_t)row * N * kE + m * kHeadDim + (lane & 7) * 4;
        const uint4 wd = *reinterpret_cast<const uint4*>(rp + m * 4);
        const uint32_t w4[4] = {wd.x, wd.y, wd.z, wd.w};
#pragma unroll
        for (int p = p0; p < p1; ++p) {
            const uint32_t wv = w4[p];
            const int base = (int)(wv & 0x03FFFFFFu);
            const int dx = (int)((wv >> 26) & 1u);
            const int dy = ((wv >> 27) & 1u) ? W : 0;
            if ((lane & 7) == 0) {      // one lane per 128-byte line
                pf_l1(Vr + (size_t)base * kE); pf_l1(Vr + (size_t)(base + dx) * kE);
                pf_l1(Vr + (size_t)(base + dy) * kE); pf_l1(Vr + (size_t)(base + dy + dx) * kE);
            }
        }
    };
    if (MODE & 2) prefetch_group(0, 2, 4);
    if (MODE & 1) prefetch_group(1, 0, 4);
    gather_group(V, rec, out_hi, out_lo, N, W, warp, 0, lane);
    gather_group(V, rec, out_hi, out_lo, N, W, warp, 1, lane);
}

// sampling records as the sampling projection's epilogue writes them, from synthetic offsets
__global__ void k_make_records(uint32_t* rec, int H, int W, int total, float sigma, uint32_t seed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // (token, head * 4 + point)
    if (idx >= total * 32) return;
    const int token = idx >> 5, k = idx & 31, m = k >> 2, p = k & 3;
    const int N = H * W, n = token % N, i = n / W, j = n - i * W;
    uint32_t s = seed ^ (uint32_t)idx * 2654435761u;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return (s >> 8) * (1.0f / 16777216.0f); };
    const float u1 = fmaxf(rnd(), 1e-7f), u2 = rnd();
    const float g0 = sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2), g1 = sqrtf(-2.f * logf(u1)) * sinf(6.2831853f * u2);
    const float th = 6.2831853f * m / 8.f;
    float cx = cosf(th), cy = sinf(th);
    const float mx = fmaxf(fabsf(cx), fabsf(cy));
    cx /= mx; cy /= mx;
    const float offx = cx * (p + 1) + sigma * g0, offy = cy * (p + 1) + sigma * g1;
    const float refx = __fdiv_rn((float)j + 0.5f, (float)W), refy = __fdiv_rn((float)i + 0.5f, (float)H);
    uint32_t word; float fx, fy;
    msda_resolve(offx, offy, refx, refy, __frcp_rn((float)W), __frcp_rn((float)H), H, W, word, fx, fy);
    uint32_t* rp = rec + (size_t)token * kRecW;
    rp[k] = word;
    reinterpret_cast<float*>(rp)[32 + k] = fx;
    reinterpret_cast<float*>(rp)[64 + k] = fy;
    reinterpret_cast<float*>(rp)[96 + k] = 0.25f + 0.1f * (rnd() - 0.5f);
}

template <class F>
float time_ms(F launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char** argv) {
    const int rows = 8, H = 128, W = 256, N = H * W, total = rows * N;
    const float sigma = argc > 1 ? (float)atof(argv[1]) : 0.5f;
    float* V; uint32_t* rec; __half *hi0, *lo0, *hi1, *lo1;
    cudaMalloc(&V, (size_t)total * kE * 4); cudaMalloc(&rec, (size_t)total * kRecW * 4);
    cudaMalloc(&hi0, (size_t)total * kE * 2); cudaMalloc(&lo0, (size_t)total * kE * 2);
    cudaMalloc(&hi1, (size_t)total * kE * 2); cudaMalloc(&lo1, (size_t)total * kE * 2);
    {
        std::vector<float> h((size_t)total * kE);
        srand(3);
        for (auto& v : h) v = (rand() % 2001 - 1000) / 500.0f;
        cudaMemcpy(V, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    }
    k_make_records<<<(total * 32 + 255) / 256, 256>>>(rec, H, W, total, sigma, 12345u);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("record setup failed\n"); return 1; }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int n_tiles = rows * (H / 8) * (W / 8);
    for (int z = 0; z < 2; ++z) {         // both tile orders must visit every tile exactly once
        std::vector<char> seen(n_tiles, 0);
        for (int t = 0; t < n_tiles; ++t) {
            int row, ti, tj;
            if (z) tile_origin<true>(t, W / 8, H / 8, row, ti, tj); else tile_origin<false>(t, W / 8, H / 8, row, ti, tj);
            const int id = (row * (H / 8) + ti) * (W / 8) + tj;
            if (ti < 0 || ti >= H / 8 || tj < 0 || tj >= W / 8 || seen[id]) { printf("tile order %d is not a bijection at t = %d\n", z, t); return 1; }
            seen[id] = 1;
        }
    }
    const double gb = (double)total * (kE * 4 + kRecW * 4 + 2 * kE * 2) / 1e9;       // algorithmic HBM bytes

    auto v0 = [&]() { k_msda_gather<<<(unsigned)(((size_t)total * 32 + 255) / 256), 256>>>(V, rec, nullptr, hi0, lo0, N, W, total); };
    const float ms0 = time_ms(v0, 20);
    printf("sigma %.2f px   v0 product kernel                       %.3f ms  (%.0f GB/s algorithmic)\n", sigma, ms0, gb / (ms0 / 1e3));
    {
        auto v0o = [&]() { k_msda_gather_old<<<(unsigned)(((size_t)total * 32 + 255) / 256), 256>>>(V, rec, nullptr, hi1, lo1, N, W, total); };
        const float msa = time_ms(v0o, 20), msb = time_ms(v0, 20), msc = time_ms(v0o, 20), msd = time_ms(v0, 20);
        printf("               A/B same box: old addressing %.4f / %.4f ms, new (IMAD.WIDE.U32) %.4f / %.4f ms\n", msa, msc, msb, msd);
    }
    std::vector<uint16_t> ref((size_t)total * kE), got((size_t)total * kE);
    cudaMemcpy(ref.data(), hi0, ref.size() * 2, cudaMemcpyDeviceToHost);
    int rc = 0;
    auto run = [&](const char* name, auto kern) {
        for (int k : {1, 2, 4}) {
            const int grid = sms * k;
            cudaMemset(hi1, 0, (size_t)total * kE * 2);
            auto l = [&]() { kern<<<grid, 256>>>(V, rec, hi1, lo1, H, W, n_tiles); };
            const float ms = time_ms(l, 20);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(got.data(), hi1, got.size() * 2, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
            printf("               %-32s %d CTA/SM  %.3f ms  (%.2fx of v0)  %s%s\n", name, k, ms, ms0 / ms,
                   bad ? "OUTPUT DIFFERS " : "bit-identical ", e == cudaSuccess ? "" : cudaGetErrorString(e));
            if (bad || e != cudaSuccess) rc = 2;
        }
    };
    {   // v4: TMA-staged windows
        PFN_encodeTiled enc = get_encode_fn();
        auto run4 = [&](const char* name, auto kern, int PW, int PH, int NBUF, int threads) {
            CUtensorMap map;
            cuuint64_t dims[4] = {(cuuint64_t)kE, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)rows};
            cuuint64_t strides[3] = {(cuuint64_t)kE * 4, (cuuint64_t)W * kE * 4, (cuuint64_t)H * W * kE * 4};
            cuuint32_t box[4] = {32, (cuuint32_t)PW, (cuuint32_t)PH, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            if (!enc || enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, V, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("               %-32s tensor map encode failed\n", name);
                rc = 2;
                return;
            }
            const int smem = NBUF * PW * PH * 128 + NBUF * 8 + 64 + 128;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaMemset(hi1, 0, (size_t)total * kE * 2);
            auto l = [&]() { kern<<<sms, threads, smem>>>(map, V, rec, hi1, lo1, H, W, rows); };
            const float ms = time_ms(l, 20);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(got.data(), hi1, got.size() * 2, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
            printf("               %-32s smem %3d KB  %.3f ms  (%.2fx of v0)  %s%s\n", name, smem >> 10, ms, ms0 / ms,
                   bad ? "OUTPUT DIFFERS " : "bit-identical ", e == cudaSuccess ? "" : cudaGetErrorString(e));
            if (bad || e != cudaSuccess) rc = 2;
        };
        run4("v4 TMA 16x16 win 24x24 x2 512t", k_gather_smem<16, 16, 24, 24, 2, 512>, 24, 24, 2, 512);
        run4("v4 TMA 16x16 win 24x24 x3 512t", k_gather_smem<16, 16, 24, 24, 3, 512>, 24, 24, 3, 512);
        run4("v4 TMA 16x16 win 24x24 x3 1024t", k_gather_smem<16, 16, 24, 24, 3, 1024>, 24, 24, 3, 1024);
        run4("v4 TMA 16x8 win 24x16 x4 512t", k_gather_smem<16, 8, 24, 16, 4, 512>, 24, 16, 4, 512);
        run4("v4 TMA 32x8 win 40x16 x2 1024t", k_gather_smem<32, 8, 40, 16, 2, 1024>, 40, 16, 2, 1024);
        run4("v4 TMA 16x16 win 28x28 x2 1024t", k_gather_smem<16, 16, 28, 28, 2, 1024>, 28, 28, 2, 1024);
        run4("v4b 16x16 win 24x24 x2 1024t 2pt", k_gather_smem<16, 16, 24, 24, 2, 1024, 2>, 24, 24, 2, 1024);
        run4("v4b 16x16 win 24x24 x3 1024t 2pt", k_gather_smem<16, 16, 24, 24, 3, 1024, 2>, 24, 24, 3, 1024);
        run4("v4b 16x16 win 24x24 x2 1024t 1pt", k_gather_smem<16, 16, 24, 24, 2, 1024, 1>, 24, 24, 2, 1024);
        run4("v4b 16x16 win 24x24 x2 512t 2pt", k_gather_smem<16, 16, 24, 24, 2, 512, 2>, 24, 24, 2, 512);
    }
    {
        auto run5 = [&](const char* name, auto kern) {
            cudaMemset(hi1, 0, (size_t)total * kE * 2);
            auto l = [&]() { kern<<<sms, 1024>>>(V, rec, hi1, lo1, H, W, rows); };
            const float ms = time_ms(l, 20);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(got.data(), hi1, got.size() * 2, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
            printf("               %-32s 1 CTA/SM   %.3f ms  (%.2fx of v0)  %s%s\n", name, ms, ms0 / ms,
                   bad ? "OUTPUT DIFFERS " : "bit-identical ", e == cudaSuccess ? "" : cudaGetErrorString(e));
            if (bad || e != cudaSuccess) rc = 2;
        };
        run5("v5 patch 8x4, head-group major", k_gather_patch<8, 4, true>);
        run5("v5 patch 8x4, token major", k_gather_patch<8, 4, false>);
        run5("v5 patch 16x2, head-group major", k_gather_patch<16, 2, true>);
        run5("v5 patch 4x8, head-group major", k_gather_patch<4, 8, true>);
        run5("v5 patch 32x1, head-group major", k_gather_patch<32, 1, true>);
    }
    {
        auto run6 = [&](const char* name, auto kern) {
            cudaMemset(hi1, 0, (size_t)total * kE * 2);
            auto l = [&]() { kern<<<(unsigned)(((size_t)total * 32 + 255) / 256), 256>>>(V, rec, hi1, lo1, N, W, total); };
            const float ms = time_ms(l, 20);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(got.data(), hi1, got.size() * 2, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
            printf("               %-32s            %.3f ms  (%.2fx of v0)  %s%s\n", name, ms, ms0 / ms,
                   bad ? "OUTPUT DIFFERS " : "bit-identical ", e == cudaSuccess ? "" : cudaGetErrorString(e));
            if (bad || e != cudaSuccess) rc = 2;
        };
        auto run7 = [&](const char* name, auto kern) {
            cudaMemset(hi1, 0, (size_t)total * kE * 2);
            auto l = [&]() { kern<<<(unsigned)(((size_t)total * 32 + 255) / 256), 256>>>(V, rec, nullptr, hi1, lo1, N, W, total); };
            const float ms = time_ms(l, 20);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(got.data(), hi1, got.size() * 2, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
            printf("               %-32s            %.3f ms  (%.2fx of v0)  %s%s\n", name, ms, ms0 / ms,
                   bad ? "OUTPUT DIFFERS " : "bit-identical ", e == cudaSuccess ? "" : cudaGetErrorString(e));
        };
        run7("v7 launch_bounds(256, 3)", k_msda_gather_lb<3>);
        run7("v7 launch_bounds(256, 5)", k_msda_gather_lb<5>);
        run7("v7 launch_bounds(256, 6)", k_msda_gather_lb<6>);
        run7("v7 launch_bounds(256, 8)", k_msda_gather_lb<8>);
        run6("v6 no prefetch (gather_group x2)", k_gather_pf<0>);
        run6("v6 L1 prefetch of head group 1", k_gather_pf<1>);
        run6("v6 L1 prefetch hg0 batch 2", k_gather_pf<2>);
        run6("v6 L1 prefetch both", k_gather_pf<3>);
    }
    run("v1 persistent, row-major tiles", k_gather_tiled<false, false>);
    run("v2 persistent, Z-order tiles", k_gather_tiled<true, false>);
    run("v3 Z-order, head-group major", k_gather_tiled<true, true>);
    return rc;
}
