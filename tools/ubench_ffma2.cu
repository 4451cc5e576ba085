// Does sm_100's packed fp32 FMA (fma.rn.f32x2 -> FFMA2) double the fp32 FMA rate per issue slot, or is it half-rate?
// Decides whether rewriting the FFN's GELU epilogue / the gather's bilinear FMAs with packed ops can pay.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_ffma2 tools/ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.001f, -0.001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) {
                unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a[i]);
                const unsigned long long mm = *reinterpret_cast<const unsigned long long*>(&m), cc = *reinterpret_cast<const unsigned long long*>(&c);
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(mm), "l"(cc));
                a[i] = *reinterpret_cast<float2*>(&r);
            } else {
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i].x) : "f"(a[i].x), "f"(m.x), "f"(c.x));
                asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i].y) : "f"(a[i].y), "f"(m.y), "f"(c.y));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    for (int packed = 0; packed < 2; ++packed) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto launch = [&]() { if (packed) k<1><<<148 * 8, 256>>>(out, iters); else k<0><<<148 * 8, 256>>>(out, iters); };
        launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * 8 * 256 * (double)iters * 16;      // scalar FMAs
        printf("%s: %.3f ms, %.1f TFLOP/s fp32 (2 flop per FMA), %s\n", packed ? "fma.rn.f32x2 (FFMA2)" : "fma.rn.f32   (FFMA) ", ms,
               2 * fma / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
