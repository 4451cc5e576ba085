// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (M=128) by operand source and N, on all SMs at once.
// Answers "what bounds the fused FFN's MMA stream": operands from shared memory (SS) or A from TMEM (TS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_mma tools/ubench_mma.cu -lcuda
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../ddp_b200/csrc/common.cuh"
#include "../ddp_b200/csrc/gemm_tc.cuh"
#include "../ddp_b200/csrc/ffn_fused.cuh"
using namespace ddp;
using namespace ddp::tc;

template <int N, bool TS, int NMMA>
__global__ void __launch_bounds__(128, 1) k_ubench(long long* out, int rounds) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(128, N);
        const uint64_t adesc = make_smem_desc(smem_u32(smem));
        const uint64_t bdesc = make_smem_desc(smem_u32(smem + 16384));
        uint32_t phase = 0;
        long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < NMMA; ++i) {
                    const int k = i & 3;
                    if (TS) umma_f16_ts(tb, tb + 384 + k * 8, bdesc + 2 * k, idesc, 1u);
                    else umma_f16(tb, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, phase);
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, bool TS>
void run(const char* name, int grid) {
    long long* d; cudaMalloc(&d, 256 * sizeof(long long));
    const int rounds = 256, nm = 12;
    auto kern = k_ubench<N, TS, nm>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
    for (int rep = 0; rep < 2; ++rep) kern<<<grid, 128, 16384 + 32768 + 1024>>>(d, rounds);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double per = 1.0 / (rounds * nm);
    const double ideal = 128.0 * N * 16 / 8192.0 * 2;   // cycles at 8192 dense fp16 FLOP/clk/SM
    printf("%-10s grid %3d  cycles/MMA(K=16) min %.1f med %.1f max %.1f   (ideal %.0f)  %s\n", name, grid, h[0] * per,
           h[grid / 2] * per, h[grid - 1] * per, ideal, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<64, false>("SS N=64", grid);
        run<128, false>("SS N=128", grid);
        run<256, false>("SS N=256", grid);
        run<64, true>("TS N=64", grid);
        run<128, true>("TS N=128", grid);
        run<256, true>("TS N=256", grid);
    }
    return 0;
}
