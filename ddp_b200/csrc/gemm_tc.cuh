// tcgen05 GEMM for sm_100a: C[M, N] = A[M, K] * W[N, K]^T with fused epilogues.
//
// Arithmetic.  The reference runs fp32.  Here every fp32 operand is carried as TWO fp16 planes,
// hi = fp16(s*x) and lo = fp16(s*x - hi) (s a power of two that keeps `lo` out of the fp16
// subnormals), and each product is three fp16 tensor-core MMAs accumulated in fp32 in TMEM:
//     A*W ~= A_hi*W_hi + A_lo*W_hi + A_hi*W_lo          (dropped term A_lo*W_lo ~ 2^-22)
// fp16 x fp16 products are exact in fp32, so the result is fp32-faithful (DDP_GEMM_TC_3XF16).
// NSPLIT == 1 issues only the hi*hi MMA (DDP_GEMM_TC_F16, fast, not parity-grade).
//
// Structure (one persistent CTA per SM, 320 threads; 576 with the 16-warp sampling / GELU epilogues; template PAIR
// runs the same kernel on CTA pairs with cta_group::2 MMAs, see Cfg):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles (128B swizzle) of A_hi, A_lo, W_hi, W_lo
//            into a ring of shared-memory stages, completion on mbarriers.
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN, K=16, kind::f16) on
//            shared-memory descriptors; tcgen05.commit releases stages / publishes accumulators.
//   warps 2-9 epilogue: tcgen05.ld the 128 x BN fp32 accumulator from TMEM (thread = row, two warps per
//            32-row lane quarter, each half of the columns), apply the fused epilogue, transpose through
//            a per-warp shared-memory tile and store whole 128-byte lines.  Two accumulator buffers in
//            TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace ddp {
namespace tc {

constexpr int BM = 128;        // rows per tile (UMMA M)
constexpr int BK = 64;         // K elements per stage = one 128-byte swizzle atom of fp16
constexpr int UMMA_K = 16;
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
// pull one box of a tensor into L2 (no shared-memory destination): hides the HBM latency of a later tma_load_2d
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// TMA box STORE shared -> global (bulk async group of the issuing thread).  The shared tile must stay untouched until
// tma_store_wait_read() (same thread) returns; rows / columns outside the tensor are clipped by the hardware.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// One lane of a converged warp (elect.sync picks the same lane every time).  The MMA-issuing WARP runs its loop
// with all 32 lanes (warp-uniform control flow and operands, so descriptors live in uniform registers) and only
// the tcgen05 instructions themselves are predicated on the elected lane.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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

// ---------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) wrappers: two CTAs of a cluster on one TPC run one M = 256 MMA, each holding its 128 rows
// of A / D and half of the B tile.  The leader (cluster rank 0) issues the MMAs; TMA loads of both CTAs report to
// the leader's mbarrier; tcgen05.commit multicasts its arrival to the same barrier in both CTAs.
// (functional check: tools/ubench_2cta.cu)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p`'s counterpart in the leader CTA (bit 24 of a shared-window address is the CTA's
// rank inside its pair; clearing it selects rank 0)
__device__ __forceinline__ uint32_t leader_smem_u32(const void* p) { return smem_u32(p) & 0xFEFFFFFFu; }

// TMA load into this CTA's shared memory; the bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// arrive on the leader's copy of `bar`.  Default semantics (release at CTA scope) on purpose: what is handed over is
// tensor memory, ordered by tcgen05.fence::before_thread_sync / after_thread_sync around the barrier (the same
// hand-shake CUTLASS's 2-SM kernels use); `.release.cluster` compiles to MEMBAR + ERRBAR + CCTL.IVALL (an L1 flush)
// and cost ~15 % of the fused q-projection's warp samples (ncu source view, profiles/r01_notes.md).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_smem_u32(bar)) : "memory");
}
// wait on this CTA's own barrier for arrivals that may come from the peer CTA (see mbar_arrive_leader for the semantics)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
    }
}

// executed by the same warp index in BOTH CTAs
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[2 x 128 lanes] (+)= A * B^T with M = 256: descriptors / TMEM addresses are the leader's, the peer uses the same offsets
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on `bar` (same offset) in every CTA of `mask` once all earlier MMAs of this thread have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask = 3) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}


// 32 lanes x 64 columns of fp32: thread t of the warp gets lane (lane_base + t), columns col..col+63
#define TMEM_R16(r, o) "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7]), \
                       "=r"(r[o+8]), "=r"(r[o+9]), "=r"(r[o+10]), "=r"(r[o+11]), "=r"(r[o+12]), "=r"(r[o+13]), "=r"(r[o+14]), "=r"(r[o+15])
#define TMEM_W16(v, o) "r"(__float_as_uint(v[o+0])), "r"(__float_as_uint(v[o+1])), "r"(__float_as_uint(v[o+2])), "r"(__float_as_uint(v[o+3])), \
                       "r"(__float_as_uint(v[o+4])), "r"(__float_as_uint(v[o+5])), "r"(__float_as_uint(v[o+6])), "r"(__float_as_uint(v[o+7])), \
                       "r"(__float_as_uint(v[o+8])), "r"(__float_as_uint(v[o+9])), "r"(__float_as_uint(v[o+10])), "r"(__float_as_uint(v[o+11])), \
                       "r"(__float_as_uint(v[o+12])), "r"(__float_as_uint(v[o+13])), "r"(__float_as_uint(v[o+14])), "r"(__float_as_uint(v[o+15]))
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : TMEM_R16(r, 0), TMEM_R16(r, 16)
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    tmem_ld32(taddr, &v[0]);
    tmem_ld32(taddr + 32, &v[32]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), TMEM_W16(v, 0), TMEM_W16(v, 16)
        : "memory");
}
__device__ __forceinline__ void tmem_st64(uint32_t taddr, const float (&v)[64]) {
    tmem_st32(taddr, &v[0]);
    tmem_st32(taddr + 32, &v[32]);
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

// ---------------------------------------------------------------------------------------------
// epilogue helpers.  A warp owns 32 rows (thread = row) and 64-column chunks; results go through a
// 4 KB per-warp shared-memory tile so that global stores are whole 128-byte lines.
// ---------------------------------------------------------------------------------------------
struct SplitOut {            // optional fp16 planes of the output, row-major [M][ld]
    __half* hi; __half* lo; int ld;
};

// gelu(x) = x * Phi(x), Phi via erf (Abramowitz-Stegun 7.1.26, |err(erf)| <= 1.5e-7), evaluated on |x| so that
// the x < 0 branch has no cancellation.  nn.GELU() of the reference is the exact-erf form.
__device__ __forceinline__ float gelu_fast(float x) {
    const float u = fabsf(x) * 0.70710678118654752440f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.0f)));          // MUFU.RCP, 1 ulp
    float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    p *= t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * u * -1.44269504088896340736f));    // MUFU.EX2, 2 ulp
    const float h = p * e;                                                                  // 0.5 * erfc(u)
    return x * (x < 0.f ? h : 1.0f - h);
}

// 32 rows x 32 fp32 columns: thread `lane` holds its row's 32 values; writes rows [0, rows_valid) of the
// warp's block to dst (row stride ld floats) as full 128-byte lines.
__device__ __forceinline__ void stage_store_f32(float* stg, const float* v, float* dst, int ld, int rows_valid, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = (lane >> 3) + 4 * i;
        float4 x = *reinterpret_cast<const float4*>(stg + r * 32 + ((c ^ (r & 7)) << 2));
        if (r < rows_valid) *reinterpret_cast<float4*>(dst + (size_t)r * ld + c * 4) = x;
    }
    __syncwarp();
}
// 32 rows x 64 fp16 columns (one plane): `h` = the row's 64 halves packed as 8 uint4
__device__ __forceinline__ void stage_store_f16(float* stg, const uint4* h, __half* dst, int ld, int rows_valid, int lane) {
    uint4* s16 = reinterpret_cast<uint4*>(stg);
#pragma unroll
    for (int j = 0; j < 8; ++j) s16[lane * 8 + (j ^ (lane & 7))] = h[j];
    __syncwarp();
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = (lane >> 3) + 4 * i;
        uint4 x = s16[r * 8 + (c ^ (r & 7))];
        if (r < rows_valid) *reinterpret_cast<uint4*>(dst + (size_t)r * ld + c * 8) = x;
    }
    __syncwarp();
}
// 32 rows x 32 fp16 columns (64-byte rows) through a 2 KB tile: `h` = the row's 32 halves packed as 4 uint4
__device__ __forceinline__ void stage_store_f16_32(float* stg, const uint4* h, __half* dst, int ld, int rows_valid, int lane) {
    uint4* s16 = reinterpret_cast<uint4*>(stg);
#pragma unroll
    for (int j = 0; j < 4; ++j) s16[lane * 4 + (j ^ ((lane >> 1) & 3))] = h[j];
    __syncwarp();
    const int c = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (lane >> 2) + 8 * i;
        uint4 x = s16[r * 4 + (c ^ ((r >> 1) & 3))];
        if (r < rows_valid) *reinterpret_cast<uint4*>(dst + (size_t)r * ld + c * 8) = x;
    }
    __syncwarp();
}
// fp16 planes of kActScale * v[0..63]; lo plane skipped when NSPLIT == 1
// 32 rows x 128 bytes (one 64-column fp16 plane block) through the warp's 4 KB tile and ONE TMA box store: the XOR pattern
// of stage_store_f16 is the SWIZZLE_128B layout; map: box {64 halves, 32 rows}
__device__ __forceinline__ void stage_tma_store_128(float* stg, const uint4* h, const CUtensorMap* map, int c0, int row0, int lane) {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
    uint4* s16 = reinterpret_cast<uint4*>(stg);
#pragma unroll
    for (int j = 0; j < 8; ++j) s16[lane * 8 + (j ^ (lane & 7))] = h[j];
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) tma_store_2d(map, stg, c0, row0);
}
template <int NSPLIT>
__device__ __forceinline__ void split_store64(float* stg, const float (&v)[64], const SplitOut& o, size_t row0, int col,
                                              int rows_valid, int lane, const CUtensorMap* mhi = nullptr,
                                              const CUtensorMap* mlo = nullptr) {
    uint4 hi[8], lo[8];
    __half2* h2 = reinterpret_cast<__half2*>(hi);
    __half2* l2 = reinterpret_cast<__half2*>(lo);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float a = v[2 * i] * kActScale, b = v[2 * i + 1] * kActScale;
        const __half2 hh = __floats2half2_rn(a, b);
        h2[i] = hh;
        if (NSPLIT > 1) {
            const float2 back = __half22float2(hh);
            l2[i] = __floats2half2_rn(a - back.x, b - back.y);
        }
    }
    if (mhi) {
        stage_tma_store_128(stg, hi, mhi, col, (int)row0, lane);
        if (NSPLIT > 1) stage_tma_store_128(stg, lo, mlo, col, (int)row0, lane);
        return;
    }
    stage_store_f16(stg, hi, o.hi + row0 * o.ld + col, o.ld, rows_valid, lane);
    if (NSPLIT > 1) stage_store_f16(stg, lo, o.lo + row0 * o.ld + col, o.ld, rows_valid, lane);
}

enum { EPI_BIAS = 0, EPI_ADD_COND = 1, EPI_SAMPLING = 2, EPI_GELU = 3, EPI_RES_LN = 4 };

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : TMEM_R16(r, 0)
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 rows x 32 bytes through a 1 KB tile: `h` = the row's 32 bytes as 2 uint4; dst / ld in halves
__device__ __forceinline__ void stage_store_32b(float* stg, const uint4* h, __half* dst, int ld, int rows_valid, int lane) {
    uint4* s16 = reinterpret_cast<uint4*>(stg);
    s16[lane * 2 + (0 ^ ((lane >> 2) & 1))] = h[0];
    s16[lane * 2 + (1 ^ ((lane >> 2) & 1))] = h[1];
    __syncwarp();
    const int c = lane & 1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = (lane >> 1) + 16 * i;
        uint4 x = s16[r * 2 + (c ^ ((r >> 2) & 1))];
        if (r < rows_valid) *reinterpret_cast<uint4*>(dst + (size_t)r * ld + c * 8) = x;
    }
    __syncwarp();
}

// The two staged stores above with the second half (LDS + STG by every thread) replaced by ONE TMA box store issued by
// lane 0: the XOR patterns of stage_store_f16_32 / stage_store_32b ARE the SWIZZLE_64B / SWIZZLE_32B shared-memory
// layouts, so the tile can be handed to the TMA unit as it is.  Halves the LSU work of a store-bound epilogue.  The tile is
// reused by the next call: lane 0 first waits until the previous box store has READ it (stage_tma_sync).
__device__ __forceinline__ void stage_tma_sync(int lane) {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
}
// 32 rows x 64 bytes; map: box {64 bytes, 32 rows}, SWIZZLE_64B; c0 in ELEMENTS of the map's data type
__device__ __forceinline__ void stage_tma_store_64(float* stg, const uint4* h, const CUtensorMap* map, int c0, int row0, int lane) {
    stage_tma_sync(lane);
    uint4* s16 = reinterpret_cast<uint4*>(stg);
#pragma unroll
    for (int j = 0; j < 4; ++j) s16[lane * 4 + (j ^ ((lane >> 1) & 3))] = h[j];
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) tma_store_2d(map, stg, c0, row0);
}
// 32 rows x 32 bytes; map: box {32 bytes, 32 rows}, SWIZZLE_32B
__device__ __forceinline__ void stage_tma_store_32(float* stg, const uint4* h, const CUtensorMap* map, int c0, int row0, int lane) {
    stage_tma_sync(lane);
    uint4* s16 = reinterpret_cast<uint4*>(stg);
    s16[lane * 2 + (0 ^ ((lane >> 2) & 1))] = h[0];
    s16[lane * 2 + (1 ^ ((lane >> 2) & 1))] = h[1];
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) tma_store_2d(map, stg, c0, row0);
}

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
    const float* pew; uint32_t* rec; int H; int W;
    // EPI_RES_LN: y = LN(acc*scale + bias) * g + b; the residual is part of the accumulator (identity block of W),
    // g / b already carry the FiLM (scale+1), shift
    const float* ln_g; const float* ln_b;
    // 3x3 convolution as an implicit GEMM (the neck, fpn.py:130-139): the A planes hold a ZERO-BORDERED token grid
    // [imgs][(H + 2) * (W + 2)][256] with row pitch conv_pitch = W + 2, K = 9 * 256 with k = tap * 256 + c.  K block kb
    // (tap = kb / 4) loads the A tile shifted by (tap / 3 - 1) rows and (tap % 3 - 1) columns = a plain offset of the TMA row
    // coordinate: the zero border supplies the padding and keeps shifts from wrapping into the neighbouring row / image,
    // rows outside the tensor are zero-filled by TMA.  0 = ordinary GEMM.
    int conv_pitch;
    // 1: the fp16 output planes (split) leave through TMA box stores (kernel parameters mapOhi / mapOlo: the planes with
    // {64 halves, 32 rows} boxes, SWIZZLE_128B) instead of LDS + STG by every thread.  EPI_RES_LN / EPI_ADD_COND / EPI_BIAS.
    int tma_stores;
};

// Sampling-projection epilogue of ONE warp.  Accumulator columns at t_row: 0..63 = (x, y) offsets of 32 sampling points
// (head, point), 64..95 attention logits, 96..127 padding.  Four warps per TMEM lane quarter: warp `half` (0..3) owns
// heads 2*half and 2*half + 1, i.e. offset columns [16 half, +16) and logits [64 + 8 half, +8); it writes the matching
// 32-byte pieces of the four record sections (index words, fx, fy, attention weights).  `stg` = this warp's >= 1 KB tile.
// This row's share of pew = PE * W_s^T + b_s (16 offsets + 8 logits of the warp's two heads): six 16-byte loads at a
// 384-byte row stride, i.e. L2-latency loads.  A caller whose epilogue is on the critical path issues them EARLY (before
// waiting for the accumulator) and hands them to sampling_epilogue.
struct PewRow { float4 o[4]; float4 a[2]; };
__device__ __forceinline__ PewRow sampling_prefetch(const EpiParams& ep, int half, size_t srow) {
    const int n = (int)(srow % ep.N_tok);
    const float* pp = ep.pew + (size_t)n * kSampW;
    PewRow r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.o[i] = __ldg(reinterpret_cast<const float4*>(pp + 16 * half + i * 4));
#pragma unroll
    for (int i = 0; i < 2; ++i) r.a[i] = __ldg(reinterpret_cast<const float4*>(pp + 64 + 8 * half + i * 4));
    return r;
}

// rec_map != nullptr: the record pieces leave through TMA box stores (map over rec as [M][128] 32-bit words, box {8, 32},
// SWIZZLE_32B) instead of per-thread stores; the caller owns the bulk-group discipline of `stg` (stage_tma_sync).
__device__ __forceinline__ void sampling_epilogue(const EpiParams& ep, uint32_t t_row, int half, int wrow0, size_t srow,
                                                  int rows_valid, int lane, float* stg, const CUtensorMap* rec_map = nullptr,
                                                  const PewRow* pre = nullptr) {
    const int n = (int)(srow % ep.N_tok);
    __half* rech = reinterpret_cast<__half*>(ep.rec + (size_t)wrow0 * kRecW);       // 2 halves per record word
    const float* pp = ep.pew + (size_t)n * kSampW;
    float o[16], a[8];
    tmem_ld16(t_row + 16 * half, o);
    tmem_ld8(t_row + 64 + 8 * half, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 p4 = pre ? pre->o[i] : *reinterpret_cast<const float4*>(pp + 16 * half + i * 4);
        o[i * 4 + 0] = fmaf(o[i * 4 + 0], ep.scale, p4.x); o[i * 4 + 1] = fmaf(o[i * 4 + 1], ep.scale, p4.y);
        o[i * 4 + 2] = fmaf(o[i * 4 + 2], ep.scale, p4.z); o[i * 4 + 3] = fmaf(o[i * 4 + 3], ep.scale, p4.w);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float4 p4 = pre ? pre->a[i] : *reinterpret_cast<const float4*>(pp + 64 + 8 * half + i * 4);
        a[i * 4 + 0] = fmaf(a[i * 4 + 0], ep.scale, p4.x); a[i * 4 + 1] = fmaf(a[i * 4 + 1], ep.scale, p4.y);
        a[i * 4 + 2] = fmaf(a[i * 4 + 2], ep.scale, p4.z); a[i * 4 + 3] = fmaf(a[i * 4 + 3], ep.scale, p4.w);
    }
    if (ep.out) {                       // raw offsets for the test tap (fp32, 2 x 8 columns)
        if (rec_map) stage_tma_sync(lane);
        __half* oh = reinterpret_cast<__half*>(ep.out + (size_t)wrow0 * ep.ldc + 16 * half);
        stage_store_32b(stg, reinterpret_cast<const uint4*>(&o[0]), oh, 2 * ep.ldc, rows_valid, lane);
        stage_store_32b(stg, reinterpret_cast<const uint4*>(&o[8]), oh + 16, 2 * ep.ldc, rows_valid, lane);
    }
    {
        const int ti = n / ep.W, tj = n - ti * ep.W;
        const float refx = __fdiv_rn((float)tj + 0.5f, (float)ep.W), refy = __fdiv_rn((float)ti + 0.5f, (float)ep.H);
        const float rW = __frcp_rn((float)ep.W), rH = __frcp_rn((float)ep.H);
        uint4 widx[2], wfx[2], wfy[2];
        uint32_t* wi = reinterpret_cast<uint32_t*>(widx);
        float* fxp = reinterpret_cast<float*>(wfx);
        float* fyp = reinterpret_cast<float*>(wfy);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            msda_resolve(o[2 * k], o[2 * k + 1], refx, refy, rW, rH, ep.H, ep.W, wi[k], fxp[k], fyp[k]);
        if (rec_map) {
            stage_tma_store_32(stg, widx, rec_map, 8 * half, wrow0, lane);
            stage_tma_store_32(stg, wfx, rec_map, 32 + 8 * half, wrow0, lane);
            stage_tma_store_32(stg, wfy, rec_map, 64 + 8 * half, wrow0, lane);
        } else {
            stage_store_32b(stg, widx, rech + 2 * (8 * half), 2 * kRecW, rows_valid, lane);
            stage_store_32b(stg, wfx, rech + 2 * (32 + 8 * half), 2 * kRecW, rows_valid, lane);
            stage_store_32b(stg, wfy, rech + 2 * (64 + 8 * half), 2 * kRecW, rows_valid, lane);
        }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {       // softmax over each head's 4 points
        const float mx = fmaxf(fmaxf(a[g * 4], a[g * 4 + 1]), fmaxf(a[g * 4 + 2], a[g * 4 + 3]));
        const float e0 = expf(a[g * 4] - mx), e1 = expf(a[g * 4 + 1] - mx), e2 = expf(a[g * 4 + 2] - mx), e3 = expf(a[g * 4 + 3] - mx);
        const float sden = (e0 + e1) + (e2 + e3);
        a[g * 4] = e0 / sden; a[g * 4 + 1] = e1 / sden; a[g * 4 + 2] = e2 / sden; a[g * 4 + 3] = e3 / sden;
    }
    if (rec_map) stage_tma_store_32(stg, reinterpret_cast<const uint4*>(&a[0]), rec_map, 96 + 8 * half, wrow0, lane);
    else stage_store_32b(stg, reinterpret_cast<const uint4*>(&a[0]), rech + 2 * (96 + 8 * half), 2 * kRecW, rows_valid, lane);
    if (ep.out && rec_map) stage_tma_sync(lane);
    if (ep.out)
        stage_store_32b(stg, reinterpret_cast<const uint4*>(&a[0]),
                        reinterpret_cast<__half*>(ep.out + (size_t)wrow0 * ep.ldc + 64 + 8 * half), 2 * ep.ldc, rows_valid, lane);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
constexpr int kStageAreaBytes = 32768;       // store staging tiles of all epilogue warps
__host__ __device__ constexpr int epi_warps(int epi) { return (epi == EPI_GELU || epi == EPI_SAMPLING) ? 16 : 8; }    // issue-bound epilogues: 16 warps
__host__ __device__ constexpr int tc_threads(int epi) { return 32 * (2 + epi_warps(epi)); }

// PAIR: the kernel runs on CTA pairs (cluster of 2): M = 256 per MMA, each CTA stages its 128 rows of A and HALF of the
// B tile (its BN / 2 rows), so the L2 -> SM stream of weight tiles per token halves.
template <int BN, int NSPLIT, int EPI = EPI_BIAS, bool PAIR = false>
struct Cfg {
    static constexpr int kEpiWarps = epi_warps(EPI);
    static constexpr int kStageTileBytes = kStageAreaBytes / kEpiWarps;
    static constexpr int kABytes = BM * BK * 2;               // one fp16 plane of an A stage
    static constexpr int kBRows = PAIR ? BN / 2 : BN;         // B rows staged by THIS CTA
    static constexpr int kBBytes = kBRows * BK * 2;
    static constexpr int kStageBytes = NSPLIT == 1 ? (kABytes + kBBytes) : 2 * (kABytes + kBBytes);
    static constexpr int kStages = (192 * 1024 / kStageBytes) > 6 ? 6 : (192 * 1024 / kStageBytes);
    static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
    static constexpr int kEpiActive = (EPI == EPI_GELU || EPI == EPI_SAMPLING) ? 16 : (BN >= 128 ? 8 : 4);      // epilogue warps that take part
    static constexpr int kColsPerWarp = (EPI == EPI_GELU || EPI == EPI_SAMPLING) ? BN / 4 : (BN >= 128 ? BN / 2 : BN);
    static constexpr int kSmemBytes = kStages * kStageBytes + kStageAreaBytes + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(EPI != EPI_GELU || BN == 256, "the 16-warp GELU epilogue assumes 64 columns per warp");
    static_assert(kStages >= 2, "need at least two stages");
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "UMMA N / epilogue chunking");
    static_assert(!PAIR || kBBytes % 1024 == 0, "a pair member's B half must keep the 1024-byte swizzle alignment");
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB shared memory of sm_100");
};

template <int BN, int NSPLIT, int EPI, bool PAIR>
__global__ void __launch_bounds__(tc_threads(EPI), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapA2hi, const __grid_constant__ CUtensorMap mapA2lo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
               const __grid_constant__ CUtensorMap mapOhi, const __grid_constant__ CUtensorMap mapOlo,
               int M, int K, int K1, int n_tiles_n, EpiParams ep) {
    // A is the K-concatenation [A (K1 columns) | A2 (K - K1 columns)]: the residual of a post-norm block rides
    // along as extra K against a scaled identity block of W, so the epilogue never reads it from global memory.
    using C = Cfg<BN, NSPLIT, EPI, PAIR>;
    constexpr int kStageTileBytes = C::kStageTileBytes;
    constexpr int kNCta = PAIR ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (128B swizzle) by pointer arithmetic on the shared array, so that the compiler keeps
    // the shared address space (LDS/STS instead of generic accesses) for the staging tiles
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_tiles = smem + C::kStages * C::kStageBytes;
    uint8_t* misc = stage_tiles + kStageAreaBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(misc);
    uint64_t* empty_bar = full_bar + C::kStages;
    uint64_t* tfull_bar = empty_bar + C::kStages;     // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles_m = (M + BM - 1) / BM;
    // a "tile" below is one round of work of this CTA: BM rows x BN columns; a pair takes two consecutive row tiles
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int n_tiles = (PAIR ? (n_tiles_m + 1) / 2 : n_tiles_m) * n_tiles_n;
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto row_of = [&](int tile) { return PAIR ? (2 * (tile / n_tiles_n) + (int)rank) * BM : (tile / n_tiles_n) * BM; };
    const int n_kb = K / BK;
    const int n_kb1 = K1 / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapAhi); tma_prefetch_desc(&mapBhi);
        if (NSPLIT > 1) { tma_prefetch_desc(&mapAlo); tma_prefetch_desc(&mapBlo); }
        if (n_kb1 < n_kb) { tma_prefetch_desc(&mapA2hi); if (NSPLIT > 1) tma_prefetch_desc(&mapA2lo); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], C::kEpiActive * kNCta); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (PAIR) {                                          // barriers of both CTAs exist before anyone signals them
        __syncthreads();
        cluster_sync_all();
        if (warp == 2) tmem_alloc_pair(tmem_slot, C::kTmemCols);
    } else {
        if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            auto load = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
                if (PAIR) tma_load_2d_pair(dst, map, bar, c0, c1); else tma_load_2d(dst, map, bar, c0, c1);
            };
            for (int tile = tile0; tile < n_tiles; tile += tile_step) {
                const int m0 = row_of(tile);
                const int n0 = (tile % n_tiles_n) * BN + (int)rank * C::kBRows;      // a pair member stages its half of the B rows
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * C::kStageBytes;
                    if (leader) mbar_expect_tx(&full_bar[stage], kNCta * C::kStageBytes);   // bytes of both CTAs land on the leader's barrier
                    const bool second = kb >= n_kb1;
                    int ka = (second ? kb - n_kb1 : kb) * BK;
                    int ma = m0;
                    if (ep.conv_pitch) {                       // implicit 3x3 convolution: tap-shifted rows of the bordered grid
                        const int tap = kb / (kE / BK);
                        ka = (kb - tap * (kE / BK)) * BK;
                        ma = m0 + (tap / 3 - 1) * ep.conv_pitch + (tap % 3 - 1);
                    }
                    load(st, second ? &mapA2hi : &mapAhi, &full_bar[stage], ka, ma);
                    load(st + C::kABytes, &mapBhi, &full_bar[stage], kb * BK, n0);
                    if (NSPLIT > 1) {
                        load(st + C::kABytes + C::kBBytes, second ? &mapA2lo : &mapAlo, &full_bar[stage], ka, ma);
                        load(st + 2 * C::kABytes + C::kBBytes, &mapBlo, &full_bar[stage], kb * BK, n0);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (the leader's, for a pair) =====================
        if (leader) {
            constexpr uint32_t idesc = make_idesc(kNCta * BM, BN);
            auto umma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accum) {
                if (PAIR) umma_f16_pair(d, a, b, idesc, accum); else umma_f16(d, a, b, idesc, accum);
            };
            auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_pair(bar, 3); else umma_commit(bar); };
            auto wait = [&](uint64_t* bar, uint32_t parity) { if (PAIR) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity); };
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = tile0; tile < n_tiles; tile += tile_step) {
                wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < n_kb; ++kb) {
                    wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * C::kStageBytes);
                    const uint64_t a_hi = make_smem_desc(st);
                    const uint64_t b_hi = make_smem_desc(st + C::kABytes);
                    const uint64_t a_lo = make_smem_desc(st + C::kABytes + C::kBBytes);
                    const uint64_t b_lo = make_smem_desc(st + 2 * C::kABytes + C::kBBytes);
                    // K blocks past K1 multiply the residual operand by the 2^shift * I block of the augmented weight: a power
                    // of two is exact in fp16, its lo plane is all zeros, so the a_hi * b_lo MMA adds exactly nothing
                    const bool ident = EPI == EPI_RES_LN && kb >= n_kb1;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);       // +32 B per K=16 step inside the swizzle atom
                        const uint32_t accum = (kb | k) != 0;
                        if (elect_one()) {
                            if (NSPLIT > 1) {
                                umma(d_tmem, a_lo + koff, b_hi + koff, accum);    // small terms first
                                if (!ident) umma(d_tmem, a_hi + koff, b_lo + koff, 1u);
                                umma(d_tmem, a_hi + koff, b_hi + koff, 1u);
                            } else {
                                umma(d_tmem, a_hi + koff, b_hi + koff, accum);
                            }
                        }
                    }
                    if (elect_one()) {
                        commit(&empty_bar[stage]);                 // stage free once these MMAs retire
                        if (kb == n_kb - 1) commit(&tfull_bar[acc]);   // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp - 2 < C::kEpiActive) {
        // ===================== epilogue: thread = row, 64-column chunks =====================
        const int ew = warp - 2;
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = ew >> 2;                                // which part (half, or quarter for GELU) of the tile's columns
        const int cbeg = half * C::kColsPerWarp;
        float* stg = reinterpret_cast<float*>(stage_tiles + ew * kStageTileBytes);
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step) {
            const int m0 = row_of(tile);
            const int n0 = (tile % n_tiles_n) * BN;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
            const int wrow0 = m0 + q * 32;                       // first row of this warp's 32-row block
            const int row = wrow0 + lane;
            const int rows_valid = M - wrow0 < 0 ? 0 : (M - wrow0 > 32 ? 32 : M - wrow0);
            const size_t srow = row < M ? (size_t)row : (size_t)(M - 1);

            if (EPI == EPI_RES_LN) {
                // pass 1: x = acc*scale + bias (residual included via the identity block) -> back into TMEM;
                // robust mean / M2 (Chan merge of 64-chunks)
                float mean = 0.f, m2 = 0.f;
#pragma unroll 1
                for (int c = cbeg; c < cbeg + C::kColsPerWarp; c += 64) {
                    float v[64];
                    tmem_ld64(t_row + c, v);
                    float cs = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + c + i * 4));
                        v[i * 4 + 0] = fmaf(v[i * 4 + 0], ep.scale, b4.x);      // the accumulator already holds W g + residual
                        v[i * 4 + 1] = fmaf(v[i * 4 + 1], ep.scale, b4.y);
                        v[i * 4 + 2] = fmaf(v[i * 4 + 2], ep.scale, b4.z);
                        v[i * 4 + 3] = fmaf(v[i * 4 + 3], ep.scale, b4.w);
                        cs += (v[i * 4 + 0] + v[i * 4 + 1]) + (v[i * 4 + 2] + v[i * 4 + 3]);
                    }
                    tmem_st64(t_row + c, v);
                    const float cm = cs * (1.0f / 64.0f);
                    float cm2 = 0.f;
#pragma unroll
                    for (int i = 0; i < 64; ++i) { const float d = v[i] - cm; cm2 = fmaf(d, d, cm2); }
                    const float na = (float)(c - cbeg), nb = 64.0f, nab = na + nb;
                    const float delta = cm - mean;
                    mean += delta * (nb / nab);
                    m2 += cm2 + delta * delta * (na * nb / nab);
                }
                // combine with the warp that owns the other half of the columns of the same rows: each warp
                // drops its partial into the PARTNER's staging tile (idle between the two passes) and reads its own
                {
                    float2* mine = reinterpret_cast<float2*>(stg);
                    float2* theirs = reinterpret_cast<float2*>(stage_tiles + (ew ^ 4) * kStageTileBytes);
                    theirs[lane] = make_float2(mean, m2);
                    named_bar_sync(1 + q, 64);
                    const float2 o = mine[lane];
                    const float2 lo_half = half == 0 ? make_float2(mean, m2) : o;     // same operand order in both warps
                    const float2 hi_half = half == 0 ? o : make_float2(mean, m2);
                    const float delta = hi_half.x - lo_half.x;
                    mean = lo_half.x + delta * 0.5f;
                    m2 = (lo_half.y + hi_half.y) + delta * delta * ((float)(C::kColsPerWarp) * 0.5f);
                    __syncwarp();
                }
                const float rstd = 1.0f / sqrtf(m2 * (1.0f / BN) + 1e-5f);
                // pass 2: normalise (+FiLM, folded into ln_g / ln_b), store
#pragma unroll 1
                for (int c = cbeg; c < cbeg + C::kColsPerWarp; c += 64) {
                    float v[64];
                    tmem_ld64(t_row + c, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(ep.ln_g + c + i * 4));
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.ln_b + c + i * 4));
                        v[i * 4 + 0] = fmaf((v[i * 4 + 0] - mean) * rstd, g4.x, b4.x);
                        v[i * 4 + 1] = fmaf((v[i * 4 + 1] - mean) * rstd, g4.y, b4.y);
                        v[i * 4 + 2] = fmaf((v[i * 4 + 2] - mean) * rstd, g4.z, b4.z);
                        v[i * 4 + 3] = fmaf((v[i * 4 + 3] - mean) * rstd, g4.w, b4.w);
                    }
                    if (ep.out) {                     // fp32 copy only when someone reads it (tests tap it); the planes carry the data
                        if (ep.tma_stores) { if (lane == 0) tma_store_wait_read(); __syncwarp(); }
                        float* dst = ep.out + (size_t)wrow0 * ep.ldc + c;
                        stage_store_f32(stg, &v[0], dst, ep.ldc, rows_valid, lane);
                        stage_store_f32(stg, &v[32], dst + 32, ep.ldc, rows_valid, lane);
                    }
                    if (ep.split.hi) split_store64<NSPLIT>(stg, v, ep.split, (size_t)wrow0, c, rows_valid, lane,
                                                           ep.tma_stores ? &mapOhi : nullptr, ep.tma_stores ? &mapOlo : nullptr);
                }
                if (ep.tma_stores) { if (lane == 0) tma_store_wait_read(); __syncwarp(); }     // the TMA unit has read this tile ...
                named_bar_sync(1 + q, 64);        // ... and the partner may write the next tile's partials into it only now
            } else if (EPI == EPI_GELU) {
                // 16 warps x 64 columns, in two 32-column rounds; everything in the 16x-scaled domain of the fp16 planes:
                // z16 = 16 (acc*scale + bias); planes of gelu(z) * 16 = z16 * Phi(z16 / 16)
                const float s16 = ep.scale * kActScale;
#pragma unroll 1
                for (int c = cbeg; c < cbeg + C::kColsPerWarp; c += 32) {
                    float v[32];
                    tmem_ld32(t_row + c, v);
                    const int col0 = n0 + c;
                    uint4 hi[4], lo[4];
                    __half2* h2 = reinterpret_cast<__half2*>(hi);
                    __half2* l2 = reinterpret_cast<__half2*>(lo);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i * 4));
                        float g[4];
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float z16 = fmaf(v[i * 4 + e], s16, bb[e] * kActScale);
                            const float u = fabsf(z16) * (0.70710678118654752440f * kInvActScale);
                            float t, ex;
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.0f)));
                            float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
                            p = fmaf(p, t, 0.5f * 1.421413741f);
                            p = fmaf(p, t, 0.5f * -0.284496736f);
                            p = fmaf(p, t, 0.5f * 0.254829592f);
                            p *= t;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(u * u * -1.44269504088896340736f));
                            const float hh = p * ex;
                            g[e] = z16 * (z16 < 0.f ? hh : 1.0f - hh);
                        }
                        const __half2 h01 = __floats2half2_rn(g[0], g[1]), h23 = __floats2half2_rn(g[2], g[3]);
                        h2[i * 2] = h01; h2[i * 2 + 1] = h23;
                        if (NSPLIT > 1) {
                            const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
                            l2[i * 2] = __floats2half2_rn(g[0] - b01.x, g[1] - b01.y);
                            l2[i * 2 + 1] = __floats2half2_rn(g[2] - b23.x, g[3] - b23.y);
                        }
                    }
                    stage_store_f16_32(stg, hi, ep.split.hi + (size_t)wrow0 * ep.split.ld + col0, ep.split.ld, rows_valid, lane);
                    if (NSPLIT > 1)
                        stage_store_f16_32(stg, lo, ep.split.lo + (size_t)wrow0 * ep.split.ld + col0, ep.split.ld, rows_valid, lane);
                }
            } else if (EPI == EPI_SAMPLING) {
                sampling_epilogue(ep, t_row, half, wrow0, srow, rows_valid, lane, stg);
            } else {
#pragma unroll 1
                for (int c = cbeg; c < cbeg + C::kColsPerWarp; c += 64) {
                    constexpr int W = C::kColsPerWarp < 64 ? C::kColsPerWarp : 64;     // 32 or 64 columns this round
                    float v[64];
                    tmem_ld32(t_row + c, &v[0]);
                    if (W == 64) tmem_ld32(t_row + c + 32, &v[32]);
                    const int col0 = n0 + c;
                    if (EPI == EPI_ADD_COND) {
                        const int n = (int)(srow % ep.N_tok);
                        const int b = (int)(srow / ep.N_tok) / ep.R;
                        const float* cp = ep.cond + ((size_t)b * ep.N_tok + n) * kE + col0;
#pragma unroll
                        for (int i = 0; i < W / 4; ++i) {
                            const float4 c4 = *reinterpret_cast<const float4*>(cp + i * 4);
                            v[i * 4 + 0] = fmaf(v[i * 4 + 0], ep.scale, c4.x); v[i * 4 + 1] = fmaf(v[i * 4 + 1], ep.scale, c4.y);
                            v[i * 4 + 2] = fmaf(v[i * 4 + 2], ep.scale, c4.z); v[i * 4 + 3] = fmaf(v[i * 4 + 3], ep.scale, c4.w);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < W / 4; ++i) {
                            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ep.bias) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i * 4));
                            v[i * 4 + 0] = fmaf(v[i * 4 + 0], ep.scale, b4.x); v[i * 4 + 1] = fmaf(v[i * 4 + 1], ep.scale, b4.y);
                            v[i * 4 + 2] = fmaf(v[i * 4 + 2], ep.scale, b4.z); v[i * 4 + 3] = fmaf(v[i * 4 + 3], ep.scale, b4.w);
                            if (EPI == EPI_GELU) {
                                v[i * 4 + 0] = gelu_fast(v[i * 4 + 0]); v[i * 4 + 1] = gelu_fast(v[i * 4 + 1]);
                                v[i * 4 + 2] = gelu_fast(v[i * 4 + 2]); v[i * 4 + 3] = gelu_fast(v[i * 4 + 3]);
                            }
                        }
                    }
                    if (ep.out) {
                        if (ep.tma_stores) { if (lane == 0) tma_store_wait_read(); __syncwarp(); }
                        if ((ep.ldc & 3) == 0 && col0 + 32 <= ep.ncols) {
                            float* dst = ep.out + (size_t)wrow0 * ep.ldc + col0;
                            stage_store_f32(stg, &v[0], dst, ep.ldc, rows_valid, lane);
                            if (W == 64 && col0 + 64 <= ep.ncols) stage_store_f32(stg, &v[32], dst + 32, ep.ldc, rows_valid, lane);
                        } else if (row < M) {
#pragma unroll
                            for (int i = 0; i < W; ++i)
                                if (col0 + i < ep.ncols) ep.out[(size_t)row * ep.ldc + col0 + i] = v[i];
                        }
                    }
                    if (W == 64 && ep.split.hi) split_store64<NSPLIT>(stg, v, ep.split, (size_t)wrow0, col0, rows_valid, lane,
                                                                      ep.tma_stores ? &mapOhi : nullptr, ep.tma_stores ? &mapOlo : nullptr);
                }
            }
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]); }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (ep.tma_stores && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();                        // the peer may still read this CTA's shared / tensor memory
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
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

// 2-D row-major tensor [rows][cols] of 32-bit elements for TMA box STORES from a warp's staging tile:
// box = [32 rows][box_cols], swizzle 64B (box_cols = 16) or 32B (box_cols = 8)
inline bool make_store_map_32bit(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, bool is_float) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 4};
    cuuint32_t box[2] = {box_cols, 32};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp16 plane [rows][cols] for TMA box STORES of 32 rows x 128 bytes (64 halves), SWIZZLE_128B
inline bool make_store_map_f16_128B(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// fp16 plane [rows][cols] for TMA box STORES of 32 rows x 32 bytes (16 halves), SWIZZLE_32B
inline bool make_store_map_f16_32B(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {16, 32};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int NSPLIT, int EPI, bool PAIR = false>
inline cudaError_t launch_gemm_tc(const CUtensorMap& aHi, const CUtensorMap& aLo, const CUtensorMap& a2Hi,
                                  const CUtensorMap& a2Lo, const CUtensorMap& bHi, const CUtensorMap& bLo, int M, int K,
                                  int K1, int n_cols_padded, const EpiParams& ep, int num_sms, cudaStream_t st,
                                  const CUtensorMap* oHi = nullptr, const CUtensorMap* oLo = nullptr) {
    // PAIR: bHi / bLo must be maps with BN / 2-row boxes; oHi / oLo (with ep.tma_stores): store maps of the output planes
    const CUtensorMap& mOh = oHi ? *oHi : bHi;
    const CUtensorMap& mOl = oLo ? *oLo : bHi;
    using C = Cfg<BN, NSPLIT, EPI, PAIR>;
    auto kern = gemm_tc_kernel<BN, NSPLIT, EPI, PAIR>;
    {   // the attribute is per device: remember it per device ordinal (one handle per GPU, but several GPUs per process are legal)
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    const int n_tiles_n = n_cols_padded / BN;
    const int n_tiles_m = (M + BM - 1) / BM;
    if (!PAIR) {
        const int n_tiles = n_tiles_m * n_tiles_n;
        const int grid = n_tiles < num_sms ? n_tiles : num_sms;
        kern<<<grid, tc_threads(EPI), C::kSmemBytes, st>>>(aHi, aLo, a2Hi, a2Lo, bHi, bLo, mOh, mOl, M, K, K1, n_tiles_n, ep);
        return cudaSuccess;
    }
    const int n_rounds = ((n_tiles_m + 1) / 2) * n_tiles_n;
    const int pairs = n_rounds < num_sms / 2 ? n_rounds : num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(tc_threads(EPI));
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, aHi, aLo, a2Hi, a2Lo, bHi, bLo, mOh, mOl, M, K, K1, n_tiles_n, ep);
}

}  // namespace tc
}  // namespace ddp
