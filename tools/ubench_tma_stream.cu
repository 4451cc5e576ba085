// Micro-benchmark: how fast can 148 persistent CTAs stream fp16 activation planes from HBM through TMA, by layout?
//   (a) row-major planes [M][256]: a 128-row x 64-column box = 128 segments of 128 B at a 512 B stride (today's layout)
//   (b) K-block-major planes [4][M][64]: the same box is one contiguous 16 KB block
// No compute: the consumer releases every stage at once.  Answers whether the A stream bounds the K = 256 projections.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_tma_stream tools/ubench_tma_stream.cu -lcuda
#include <cstdio>
#include <vector>
#include "../ddp_b200/csrc/common.cuh"
#include "../ddp_b200/csrc/gemm_tc.cuh"
using namespace ddp;
using namespace ddp::tc;

constexpr int kStagesS = 8;
constexpr int kBox = 16384;

__global__ void __launch_bounds__(64, 1)
k_stream(const __grid_constant__ CUtensorMap mapHi, const __grid_constant__ CUtensorMap mapLo, int M, int kb_major) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[kStagesS], empty[kStagesS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    const int n_tiles = M / 128;
    if (warp == 0 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kb = 0; kb < 4; ++kb)
                for (int pl = 0; pl < 2; ++pl) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], kBox);
                    const int c0 = kb_major ? 0 : kb * 64, c1 = kb_major ? kb * M + tile * 128 : tile * 128;
                    tma_load_2d(smem + stage * kBox, pl ? &mapLo : &mapHi, &full[stage], c0, c1);
                    if (++stage == kStagesS) { stage = 0; phase ^= 1; }
                }
    } else if (warp == 1 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int i = 0; i < 8; ++i) {
                mbar_wait(&full[stage], phase);
                mbar_arrive(&empty[stage]);
                if (++stage == kStagesS) { stage = 0; phase ^= 1; }
            }
    }
}

int main() {
    const int M = 262144;
    const size_t plane = (size_t)M * 256 * 2;
    __half *hi, *lo; uint8_t* flush;
    cudaMalloc(&hi, plane); cudaMalloc(&lo, plane); cudaMalloc(&flush, 512u << 20);
    cudaMemset(hi, 1, plane); cudaMemset(lo, 2, plane);
    const int smem = kStagesS * kBox + 1024;
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int layout = 0; layout < 2; ++layout) {
        CUtensorMap mh, ml;
        bool ok = layout ? (make_map_f16(&mh, hi, 4 * (uint64_t)M, 64, 128) && make_map_f16(&ml, lo, 4 * (uint64_t)M, 64, 128))
                         : (make_map_f16(&mh, hi, M, 256, 128) && make_map_f16(&ml, lo, M, 256, 128));
        if (!ok) { printf("map failed\n"); return 1; }
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemset(flush, rep, 512u << 20);                 // evict the planes from L2
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k_stream<<<148, 64, smem>>>(mh, ml, M, layout);
            cudaEventRecord(e1);
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 2; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        printf("%s: %.3f ms for %.0f MB  => %.2f TB/s\n", layout ? "K-block-major [4][M][64]" : "row-major [M][256]      ", best,
               2.0 * plane / 1e6, 2.0 * plane / best / 1e9);
    }
    return 0;
}
