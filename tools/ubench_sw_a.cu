// Functional check + timing for the planned gather -> out_proj fusion ("gather as the A-operand
// producer"): can the A tile of a tcgen05.mma be WRITTEN BY THE SM's OWN THREADS (generic-proxy st.shared in the
// 128-byte-swizzle K-major layout TMA would have produced) instead of arriving through TMA?
//   D1 = A * B^T with A loaded by TMA                    (the layout of record)
//   D2 = A * B^T with A written by the 4 warps with st.shared.v4 at swizzled addresses, fence.proxy.async, barrier
// Both are compared with a CPU product; then the software-producer path is timed (cycles per 128 x 64 fp16 tile written
// by 128 threads) to see what a warp-per-token gather would pay for staging its results.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_sw_a tools/ubench_sw_a.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../ddp_b200/csrc/common.cuh"
#include "../ddp_b200/csrc/gemm_tc.cuh"
using namespace ddp;
using namespace ddp::tc;

// Byte offset of the 16-byte chunk `c16` (0..7) of row `r` inside a K-major tile with 128-byte rows under
// CU_TENSOR_MAP_SWIZZLE_128B: 8-row atoms of 1024 B, chunk index XOR-ed with the row index inside the atom.
__host__ __device__ inline uint32_t swz128(int r, int c16) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(128, 1)
k_sw_a(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __half* A, float* D1, float* D2,
       long long* cycles, int rounds) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 128 rows x 128 B, written by TMA
    uint8_t* sA2 = smem + 16384;        // same tile, written by the threads
    uint8_t* sB = smem + 32768;         // 128 rows x 128 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152);
    uint64_t* full = bars + 0;
    uint64_t* done = bars + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); fence_barrier_init(); }
    __syncthreads();
    if (warp == 0) tmem_alloc(slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(full, 2 * 16384);
        tma_load_2d(sA, &mapA, full, 0, 0);
        tma_load_2d(sB, &mapB, full, 0, 0);
    }
    // software producer: warp w writes rows 32 w .. 32 w + 31 (one row per 8 lanes: lane & 7 = chunk, lane >> 3 = row in group of 4)
    long long t0 = clock64();
    for (int it = 0; it < rounds; ++it) {
        for (int rr = 0; rr < 32; rr += 4) {
            const int r = warp * 32 + rr + (lane >> 3), c16 = lane & 7;
            const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * 64 + c16 * 8);
            *reinterpret_cast<uint4*>(sA2 + swz128(r, c16)) = v;
        }
    }
    fence_proxy_async();                // generic-proxy writes -> visible to the async proxy (tcgen05.mma operand fetch)
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    constexpr uint32_t idesc = make_idesc(128, 128);
    if (warp == 1) {
        mbar_wait(full, 0);
        tc_fence_after();
        const uint64_t ad = make_smem_desc(smem_u32(sA)), ad2 = make_smem_desc(smem_u32(sA2)), bd = make_smem_desc(smem_u32(sB));
        if (elect_one()) {
            for (int k = 0; k < 4; ++k) umma_f16(tb, ad + 2 * k, bd + 2 * k, idesc, k > 0);
            for (int k = 0; k < 4; ++k) umma_f16(tb + 128, ad2 + 2 * k, bd + 2 * k, idesc, k > 0);
            umma_commit(done);
        }
        __syncwarp();
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t t_row = tb + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld32(t_row + c, v);
        for (int i = 0; i < 32; ++i) D1[(size_t)row * 128 + c + i] = v[i];
        tmem_ld32(t_row + 128 + c, v);
        for (int i = 0; i < 32; ++i) D2[(size_t)row * 128 + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

int main() {
    const int M = 128, N = 128, K = 64;
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K);
    srand(2);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2half((rand() % 17 - 8) / 8.0f); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2half((rand() % 13 - 6) / 4.0f); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float *dD1, *dD2; long long* dC;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD1, M * N * 4); cudaMalloc(&dD2, M * N * 4);
    cudaMalloc(&dC, 256 * sizeof(long long));
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    CUtensorMap mA, mB;
    if (!make_map_f16(&mA, dA, M, K, 128) || !make_map_f16(&mB, dB, N, K, 128)) { printf("tensor map failed\n"); return 1; }
    const int smem = 49152 + 1024 + 256;
    cudaFuncSetAttribute(k_sw_a, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 148}) {
        const int rounds = grid == 1 ? 1 : 64;
        cudaMemset(dD1, 0xff, M * N * 4); cudaMemset(dD2, 0xff, M * N * 4);
        k_sw_a<<<grid, 128, smem>>>(mA, mB, dA, dD1, dD2, dC, rounds);
        cudaError_t e = cudaDeviceSynchronize();
        printf("grid %3d launch: %s\n", grid, cudaGetErrorString(e));
        if (e != cudaSuccess) return 2;
        std::vector<float> h1(M * N), h2(M * N);
        std::vector<long long> hc(grid);
        cudaMemcpy(h1.data(), dD1, M * N * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h2.data(), dD2, M * N * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hc.data(), dC, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        int bad1 = 0, bad2 = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0;
                for (int k = 0; k < K; ++k) r += (double)fA[m * K + k] * fB[n * K + k];
                if (!(fabs(h1[m * N + n] - r) <= 1e-3)) { if (bad1 < 3) printf("TMA-A mismatch m=%d n=%d got %f want %f\n", m, n, h1[m * N + n], r); ++bad1; }
                if (!(fabs(h2[m * N + n] - r) <= 1e-3)) { if (bad2 < 3) printf("SW-A  mismatch m=%d n=%d got %f want %f\n", m, n, h2[m * N + n], r); ++bad2; }
            }
        printf("grid %3d: TMA-written A: %d bad of %d; thread-written swizzled A: %d bad; %.0f cycles per 16 KB tile written by 128 threads\n",
               grid, bad1, M * N, bad2, (double)hc[0] / rounds);
        if (bad1 || bad2) return 3;
    }
    return 0;
}
