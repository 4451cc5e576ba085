// tcgen05 GEMM for sm_100a: C[M, N] = A[M, K] * W[N, K]^T with fused epilogues.
//
// Arithmetic.  The reference runs fp32.  Here every fp32 operand is carried as TWO fp16 planes,
// hi = fp16(s*x) and lo = fp16(s*x - hi) (s a power of two that keeps `lo` out of the fp16
// subnormals), and each product is three fp16 tensor-core MMAs accumulated in fp32 in TMEM:
//     A*W ~= A_hi*W_hi + A_lo*W_hi + A_hi*W_lo          (dropped term A_lo*W_lo ~ 2^-22)
// fp16 x fp16 products are exact in fp32, so the result is fp32-faithful (DDP_GEMM_TC_3XF16).
// NSPLIT == 1 issues only the hi*hi MMA (DDP_GEMM_TC_F16, fast, not parity-grade).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles (128B swizzle) of A_hi, A_lo, W_hi, W_lo
//            into a ring of shared-memory stages, completion on mbarriers.
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN, K=16, kind::f16) on
//            shared-memory descriptors; tcgen05.commit releases stages / publishes accumulators.
//   warps 2-5 epilogue: tcgen05.ld the 128 x BN fp32 accumulator from TMEM (thread = row), apply the
//            fused epilogue, store.  Two accumulator buffers in TMEM let the epilogue of tile i
//            overlap the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace ddp {
namespace tc {

constexpr int BM = 128;        // rows per tile (UMMA M)
constexpr int BK = 64;         // K elements per stage = one 128-byte swizzle atom of fp16
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr float kActScale = 16.0f;       // activations are stored as fp16 planes of 16*x
constexpr float kInvActScale = 1.0f / 16.0f;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; one thread issues for the CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (lane_base + t), columns col..col+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, 128-byte swizzle (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 ([4,6) = 1), A/B fp16 (0), K-major both,
// N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// fp32 -> (hi, lo) fp16 planes of kActScale * x
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    float s = fminf(fmaxf(x * kActScale, -65504.0f), 65504.0f);
    hi = __float2half_rn(s);
    lo = __float2half_rn(s - __half2float(hi));
}

// ---------------------------------------------------------------------------------------------
// epilogues: thread owns one row (global row index `row`), columns arrive in chunks of 32
// ---------------------------------------------------------------------------------------------
struct SplitOut {            // optional fp16 planes of the output, row-major [M][ld]
    __half* hi; __half* lo; int ld;
};

__device__ __forceinline__ void store_row_f32(float* dst, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(dst + i * 4) = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
}
template <int NSPLIT>
__device__ __forceinline__ void store_row_split(const SplitOut& o, size_t row, int col, const float (&v)[32]) {
    __align__(16) __half h[32];
    __align__(16) __half l[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) split_f16(v[i], h[i], l[i]);
    uint4* dh = reinterpret_cast<uint4*>(o.hi + row * o.ld + col);
#pragma unroll
    for (int i = 0; i < 4; ++i) dh[i] = reinterpret_cast<const uint4*>(h)[i];
    if (NSPLIT > 1) {
        uint4* dl = reinterpret_cast<uint4*>(o.lo + row * o.ld + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) dl[i] = reinterpret_cast<const uint4*>(l)[i];
    }
}

enum { EPI_BIAS = 0, EPI_ADD_COND = 1, EPI_SAMPLING = 2, EPI_GELU = 3, EPI_RES_LN = 4 };

struct EpiParams {
    float scale;             // undoes the operand scales: acc * scale = A*W^T
    const float* bias;       // [N] or null
    float* out;              // fp32 output [M][ldc] or null
    int ldc;
    int ncols;               // valid columns (<= N)
    SplitOut split;          // fp16 planes of the output (hi null = none)
    // EPI_ADD_COND
    const float* cond; int N_tok; int R;
    // EPI_SAMPLING
    const float* pew;
    // EPI_RES_LN
    const float* resid; const float* gamma; const float* beta; const float* film;
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int BN, int NSPLIT>
struct Cfg {
    static constexpr int kABytes = BM * BK * 2;               // one fp16 plane of an A stage
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = NSPLIT == 1 ? (kABytes + kBBytes) : 2 * (kABytes + kBBytes);
    static constexpr int kStages = (200 * 1024 / kStageBytes) > 6 ? 6 : (200 * 1024 / kStageBytes);
    static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(kStages >= 2, "need at least two stages");
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
};

template <int BN, int NSPLIT, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
               int M, int K, int n_tiles_n, EpiParams ep) {
    using C = Cfg<BN, NSPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
    uint64_t* empty_bar = full_bar + C::kStages;
    uint64_t* tfull_bar = empty_bar + C::kStages;     // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles_m = (M + BM - 1) / BM;
    const int n_tiles = n_tiles_m * n_tiles_n;
    const int n_kb = K / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapAhi); tma_prefetch_desc(&mapBhi);
        if (NSPLIT > 1) { tma_prefetch_desc(&mapAlo); tma_prefetch_desc(&mapBlo); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles_n) * BM;
                const int n0 = (tile % n_tiles_n) * BN;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * C::kStageBytes;
                    mbar_expect_tx(&full_bar[stage], C::kStageBytes);
                    tma_load_2d(st, &mapAhi, &full_bar[stage], kb * BK, m0);
                    tma_load_2d(st + C::kABytes, &mapBhi, &full_bar[stage], kb * BK, n0);
                    if (NSPLIT > 1) {
                        tma_load_2d(st + C::kABytes + C::kBBytes, &mapAlo, &full_bar[stage], kb * BK, m0);
                        tma_load_2d(st + 2 * C::kABytes + C::kBBytes, &mapBlo, &full_bar[stage], kb * BK, n0);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * C::kStageBytes);
                    const uint64_t a_hi = make_smem_desc(st);
                    const uint64_t b_hi = make_smem_desc(st + C::kABytes);
                    const uint64_t a_lo = make_smem_desc(st + C::kABytes + C::kBBytes);
                    const uint64_t b_lo = make_smem_desc(st + 2 * C::kABytes + C::kBBytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);       // +32 B per K=16 step inside the swizzle atom
                        const uint32_t first = (kb | k) != 0;
                        if (NSPLIT > 1) {
                            umma_f16(d_tmem, a_lo + koff, b_hi + koff, idesc, first);    // small terms first
                            umma_f16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1u);
                            umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, 1u);
                        } else {
                            umma_f16(d_tmem, a_hi + koff, b_hi + koff, idesc, first);
                        }
                    }
                    umma_commit(&empty_bar[stage]);                 // stage free once these MMAs retire
                    if (kb == n_kb - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..5): thread = row =====================
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles_n) * BM;
            const int n0 = (tile % n_tiles_n) * BN;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < M;
            const size_t srow = row_ok ? (size_t)row : (size_t)(M - 1);

            if (EPI == EPI_RES_LN) {
                // pass 1: x = acc*scale + bias + resid -> back into TMEM; robust mean / M2 (Chan merge of 32-chunks)
                float mean = 0.f, m2 = 0.f;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    float v[32];
                    tmem_ld32(t_row + c, v);
                    const float* rp = ep.resid + srow * kE + c;
                    float cs = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 r4 = *reinterpret_cast<const float4*>(rp + i * 4);
                        float4 b4 = *reinterpret_cast<const float4*>(ep.bias + c + i * 4);
                        v[i * 4 + 0] = (v[i * 4 + 0] * ep.scale + b4.x) + r4.x;
                        v[i * 4 + 1] = (v[i * 4 + 1] * ep.scale + b4.y) + r4.y;
                        v[i * 4 + 2] = (v[i * 4 + 2] * ep.scale + b4.z) + r4.z;
                        v[i * 4 + 3] = (v[i * 4 + 3] * ep.scale + b4.w) + r4.w;
                        cs += (v[i * 4 + 0] + v[i * 4 + 1]) + (v[i * 4 + 2] + v[i * 4 + 3]);
                    }
                    tmem_st32(t_row + c, v);
                    const float cm = cs * (1.0f / 32.0f);
                    float cm2 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) { float d = v[i] - cm; cm2 = fmaf(d, d, cm2); }
                    const float na = (float)c, nb = 32.0f, nab = na + nb;
                    const float delta = cm - mean;
                    mean += delta * (nb / nab);
                    m2 += cm2 + delta * delta * (na * nb / nab);
                }
                const float rstd = 1.0f / sqrtf(m2 * (1.0f / BN) + 1e-5f);
                // pass 2: normalise, FiLM, store
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    float v[32];
                    tmem_ld32(t_row + c, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float y = (v[i] - mean) * rstd * ep.gamma[c + i] + ep.beta[c + i];
                        if (ep.film) y = y * (ep.film[c + i] + 1.0f) + ep.film[kE + c + i];
                        v[i] = y;
                    }
                    if (row_ok) {
                        store_row_f32(ep.out + (size_t)row * ep.ldc + c, v);
                        if (ep.split.hi) store_row_split<NSPLIT>(ep.split, (size_t)row, c, v);
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    float v[32];
                    tmem_ld32(t_row + c, v);
                    const int col0 = n0 + c;
                    if (EPI == EPI_ADD_COND) {
                        const int n = (int)(srow % ep.N_tok);
                        const int b = (int)(srow / ep.N_tok) / ep.R;
                        const float* cp = ep.cond + ((size_t)b * ep.N_tok + n) * kE + col0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 c4 = *reinterpret_cast<const float4*>(cp + i * 4);
                            v[i * 4 + 0] = v[i * 4 + 0] * ep.scale + c4.x; v[i * 4 + 1] = v[i * 4 + 1] * ep.scale + c4.y;
                            v[i * 4 + 2] = v[i * 4 + 2] * ep.scale + c4.z; v[i * 4 + 3] = v[i * 4 + 3] * ep.scale + c4.w;
                        }
                    } else if (EPI == EPI_SAMPLING) {
                        const int n = (int)(srow % ep.N_tok);
                        const float* pp = ep.pew + (size_t)n * kSampW + col0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = (col0 + i < kSampW) ? v[i] * ep.scale + pp[i] : 0.f;
                        if (col0 >= 64) {           // attention weights: softmax over each head's 4 points
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                float mx = fmaxf(fmaxf(v[g * 4], v[g * 4 + 1]), fmaxf(v[g * 4 + 2], v[g * 4 + 3]));
                                float e0 = expf(v[g * 4] - mx), e1 = expf(v[g * 4 + 1] - mx), e2 = expf(v[g * 4 + 2] - mx), e3 = expf(v[g * 4 + 3] - mx);
                                float s = (e0 + e1) + (e2 + e3);
                                v[g * 4] = e0 / s; v[g * 4 + 1] = e1 / s; v[g * 4 + 2] = e2 / s; v[g * 4 + 3] = e3 / s;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float x = v[i] * ep.scale + (ep.bias ? ep.bias[col0 + i] : 0.f);
                            if (EPI == EPI_GELU) x = gelu_erf(x);
                            v[i] = x;
                        }
                    }
                    if (row_ok) {
                        if (ep.out) {
                            if (col0 + 32 <= ep.ncols && (ep.ldc & 3) == 0) {
                                store_row_f32(ep.out + (size_t)row * ep.ldc + col0, v);
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (col0 + i < ep.ncols) ep.out[(size_t)row * ep.ldc + col0 + i] = v[i];
                            }
                        }
                        if (ep.split.hi) store_row_split<NSPLIT>(ep.split, (size_t)row, col0, v);
                    }
                }
            }
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D fp16 row-major tensor [rows][cols] (cols contiguous), box = [box_rows][64], 128-byte swizzle, zero OOB fill
inline bool make_map_f16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int NSPLIT, int EPI>
inline cudaError_t launch_gemm_tc(const CUtensorMap& aHi, const CUtensorMap& aLo, const CUtensorMap& bHi,
                                  const CUtensorMap& bLo, int M, int K, int n_cols_padded, const EpiParams& ep,
                                  int num_sms, cudaStream_t st) {
    using C = Cfg<BN, NSPLIT>;
    static bool attr_set = false;
    auto kern = gemm_tc_kernel<BN, NSPLIT, EPI>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int n_tiles_n = n_cols_padded / BN;
    const int n_tiles = ((M + BM - 1) / BM) * n_tiles_n;
    const int grid = n_tiles < num_sms ? n_tiles : num_sms;
    kern<<<grid, kThreads, C::kSmemBytes, st>>>(aHi, aLo, bHi, bLo, M, K, n_tiles_n, ep);
    return cudaSuccess;
}

}  // namespace tc
}  // namespace ddp
