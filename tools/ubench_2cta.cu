// Functional check of the CTA-pair primitives in ddp_b200/csrc/gemm_tc.cuh on one cluster:
//   D  = A * B^T   (A [256 x 64], B [128 x 64], fp16, M = 256 across two CTAs), operands from shared memory (SS)
//   D2 = A * B^T   with A re-staged in tensor memory by the epilogue warps of both CTAs (TS) after a remote arrive
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_2cta tools/ubench_2cta.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../ddp_b200/csrc/common.cuh"
#include "../ddp_b200/csrc/gemm_tc.cuh"
#include "../ddp_b200/csrc/ffn_fused.cuh"
using namespace ddp;
using namespace ddp::tc;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k_pair(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __half* A, float* D, float* D2) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 128 rows x 128 B
    uint8_t* sB = smem + 16384;         // 64 rows x 128 B (this CTA's half of N)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
    uint64_t* full = bars + 0;          // leader: TMA bytes of both CTAs
    uint64_t* done = bars + 1;          // both: MMA commit (multicast)
    uint64_t* a_ready = bars + 2;       // leader: 8 epilogue warps (4 local + 4 remote)
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        mbar_init(full, 1); mbar_init(done, 1); mbar_init(a_ready, 8);
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_alloc_pair(slot, 256);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tb = *slot;
    if (threadIdx.x == 0) {
        if (rank == 0) mbar_expect_tx(full, 2 * (16384 + 8192));
        tma_load_2d_pair(sA, &mapA, full, 0, rank * 128);
        tma_load_2d_pair(sB, &mapB, full, 0, rank * 64);
    }
    constexpr uint32_t idesc = make_idesc(256, 128);
    if (warp == 1 && rank == 0) {
        mbar_wait_cluster(full, 0);
        tc_fence_after();
        const uint64_t ad = make_smem_desc(smem_u32(sA)), bd = make_smem_desc(smem_u32(sB));
        if (elect_one()) {
            for (int k = 0; k < 4; ++k) umma_f16_pair(tb, ad + 2 * k, bd + 2 * k, idesc, k > 0);
            umma_commit_pair(done, 3);
        }
        __syncwarp();
    }
    // all warps of both CTAs: read D (own 128 rows), then stage A in TMEM for the TS pass
    mbar_wait_cluster(done, 0);
    tc_fence_after();
    const uint32_t t_row = tb + ((uint32_t)(warp * 32) << 16);
    const int row = rank * 128 + warp * 32 + lane;
    for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld32(t_row + c, v);
        for (int i = 0; i < 32; ++i) D[(size_t)row * 128 + c + i] = v[i];
    }
    {
        float v[32];
        const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)row * 64);
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(arow[i]);
        tmem_st32(t_row + 128 + 64, v);                  // A planes at columns [192, 224)
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(a_ready);
    }
    if (warp == 1 && rank == 0) {
        mbar_wait_cluster(a_ready, 0);
        tc_fence_after();
        const uint64_t bd = make_smem_desc(smem_u32(sB));
        if (elect_one()) {
            for (int k = 0; k < 4; ++k) umma_f16_ts_pair(tb, tb + 192 + k * 8, bd + 2 * k, idesc, 0u + (k > 0));
            umma_commit_pair(done, 3);
        }
        __syncwarp();
    }
    mbar_wait_cluster(done, 1);
    tc_fence_after();
    for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld32(t_row + c, v);
        for (int i = 0; i < 32; ++i) D2[(size_t)row * 128 + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair(tb, 256);
}

int main() {
    const int M = 256, N = 128, K = 64;
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K);
    srand(1);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2half((rand() % 17 - 8) / 8.0f); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2half((rand() % 13 - 6) / 4.0f); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float *dD, *dD2;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dD2, M * N * 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * N * 4); cudaMemset(dD2, 0xff, M * N * 4);
    CUtensorMap mA, mB;
    if (!make_map_f16(&mA, dA, M, K, 128) || !make_map_f16(&mB, dB, N, K, 64)) { printf("tensor map failed\n"); return 1; }
    const int smem = 16384 + 8192 + 1024 + 256;
    cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_pair<<<2, 128, smem>>>(mA, mB, dA, dD, dD2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> hD(M * N), hD2(M * N);
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD2.data(), dD2, M * N * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0; int bad1 = 0, bad2 = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)fA[m * K + k] * fB[n * K + k];
            const double a = fabs(hD[m * N + n] - r), b = fabs(hD2[m * N + n] - r);
            if (!(a <= 1e-3)) { if (bad1 < 4) printf("SS mismatch m=%d n=%d got %f want %f\n", m, n, hD[m * N + n], r); ++bad1; }
            if (!(b <= 1e-3)) { if (bad2 < 4) printf("TS mismatch m=%d n=%d got %f want %f\n", m, n, hD2[m * N + n], r); ++bad2; }
            if (a == a) e1 = fmax(e1, a);
            if (b == b) e2 = fmax(e2, b);
        }
    printf("SS: max err %.3g, %d bad of %d;  TS: max err %.3g, %d bad\n", e1, bad1, M * N, e2, bad2);
    return (bad1 || bad2) ? 3 : 0;
}
