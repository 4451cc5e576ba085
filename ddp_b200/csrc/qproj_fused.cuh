// Value + sampling projections of one deformable-attention layer in ONE kernel (sm_100a, CTA pairs).
//
// Both projections multiply the same tokens q (reference: value = value_proj(q), offsets / attention weights =
// Linear(q + pos), mmcv/ops/multi_scale_deform_attn.py:299-328), and both GEMMs are bound by streaming q's fp16 planes
// from HBM (profiles/r01_notes.md).  Here every 128 x 64 tile of q is staged ONCE and used by three MMA groups:
//   S  = q * W_s^T   (N = 128: 64 offsets, 32 attention logits, 32 padding)  -> TMEM columns [256, 384) / [384, 512)
//   V0 = q * W_v[0:128]^T, V1 = q * W_v[128:256]^T                            -> TMEM columns [0, 256)
// which removes one full read of the planes (268 MB of 938 MB per layer at the headline shape) and one launch.
//
// Structure (608 threads, cluster of 2 CTAs = 256 tokens per round, cta_group::2 MMAs issued by the leader):
//   warp 0      TMA producer of the A ring (4 x [hi | lo] 128 x 64 tiles of this CTA's rows: a whole tile of lookahead,
//               the planes come from HBM)
//   warp 18     TMA producer of the weight ring (4 x [hi | lo] halves: this CTA's 64 of the 128 rows of one weight
//               tile, L2 resident), in consumption order S(kb), V0(kb), V1(kb).  Two producers because one in-order
//               thread would tie the A lookahead to the depth of the weight ring (measured: 0.222 ms/launch)
//   warp 1      MMA issuer (leader CTA): per K block 3 groups of 4 x (lo.hi, hi.lo, hi.hi) MMAs, M = 256, N = 128
//   warps 2-17  epilogue, thread = row, four warps per TMEM lane quarter: first the value tile (bias, fp32 store;
//               releases V's columns early so the next round's V MMAs can start), then the sampling epilogue
//               (sampling_epilogue() of gemm_tc.cuh: softmax of 4, resolved sampling records)
// The S accumulator is double buffered (TMEM has 128 spare columns), V is single buffered: the next round's S group
// of K block 0 is issued before the first V group, which hides most of the value epilogue.
#pragma once
#include "gemm_tc.cuh"

namespace ddp {
namespace tc {

struct QprojParams {
    float scale_v;           // 1 / (2^shift_v * 16)
    const float* bias_v;     // [256]
    float* V;                // [M][256] fp32
    EpiParams samp;          // scale, pew, N_tok, H, W, rec, out (optional raw tap), ldc: as for EPI_SAMPLING
    int tma_stores;          // 1: the value tile and the records leave through TMA box stores (mapVout / mapRec)
    int pew_early;           // 1: this row's pew values are loaded before the value tile (their L2 latency hides under it)
};

constexpr int kQpThreads = 32 * 19;                   // A producer, MMA issuer, 16 epilogue warps, weight producer
constexpr int kQpAStages = 4;
constexpr int kQpBUnits = 4;
constexpr int kQpAPlane = BM * BK * 2;                  // 16 KB
constexpr int kQpAStage = 2 * kQpAPlane;                // [hi | lo]
constexpr int kQpBPlane = 64 * BK * 2;                  // a pair member's 64 rows: 8 KB
constexpr int kQpBUnit = 2 * kQpBPlane;                 // [hi | lo]
constexpr int kQpStageTile = 2048;                      // per-warp store staging tile
constexpr int kQpStageArea = 16 * kQpStageTile;
constexpr int kQpSmem = kQpAStages * kQpAStage + kQpBUnits * kQpBUnit + kQpStageArea + 1024 + 256;
static_assert(kQpSmem <= 232448, "fused q-projection kernel exceeds shared memory");

template <int NSPLIT>
__global__ void __launch_bounds__(kQpThreads, 1)
qproj_fused_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapVhi, const __grid_constant__ CUtensorMap mapVlo,
                   const __grid_constant__ CUtensorMap mapShi, const __grid_constant__ CUtensorMap mapSlo,
                   const __grid_constant__ CUtensorMap mapVout, const __grid_constant__ CUtensorMap mapRec,
                   int M, int K, QprojParams p) {
    constexpr int kPl = NSPLIT > 1 ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_ring = smem;
    uint8_t* b_ring = a_ring + kQpAStages * kQpAStage;
    uint8_t* stage_tiles = b_ring + kQpBUnits * kQpBUnit;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_tiles + kQpStageArea);
    uint64_t* a_full = bars + 0;                         // [4] leader: bytes of both CTAs
    uint64_t* a_empty = bars + 4;                        // [4] both: commit multicast
    uint64_t* b_full = bars + 8;                         // [4]
    uint64_t* b_empty = bars + 12;                       // [4]
    uint64_t* v_full = bars + 16;                        // both: V accumulator complete
    uint64_t* v_empty = bars + 17;                       // leader: 32 epilogue warps drained V
    uint64_t* s_full = bars + 18;                        // [2]
    uint64_t* s_empty = bars + 20;                       // [2] leader
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_rounds = (M + 2 * BM - 1) / (2 * BM);
    const int round0 = (int)(blockIdx.x >> 1);
    const int round_step = (int)(gridDim.x >> 1);
    const int n_kb = K / BK;

    if (warp == 0 && lane == 0) {
        if (p.tma_stores) { tma_prefetch_desc(&mapVout); tma_prefetch_desc(&mapRec); }
        tma_prefetch_desc(&mapAhi); tma_prefetch_desc(&mapVhi); tma_prefetch_desc(&mapShi);
        if (NSPLIT > 1) { tma_prefetch_desc(&mapAlo); tma_prefetch_desc(&mapVlo); tma_prefetch_desc(&mapSlo); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kQpAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kQpBUnits; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(v_full, 1); mbar_init(v_empty, 32);
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 32); }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();
    cluster_sync_all();                                  // barriers of both CTAs exist before anyone signals them
    if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tV = tmem_base, tS = tmem_base + 256;

    if (warp == 0) {
        // ===================== TMA producer: activation planes =====================
        if (lane == 0) {
            int as = 0; uint32_t aph = 0;
            for (int r = round0; r < n_rounds; r += round_step) {
                const int m0 = (2 * r + (int)rank) * BM;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&a_empty[as], aph ^ 1);
                    uint8_t* st = a_ring + as * kQpAStage;
                    if (leader) mbar_expect_tx(&a_full[as], 2 * kPl * kQpAPlane);
                    tma_load_2d_pair(st, &mapAhi, &a_full[as], kb * BK, m0);
                    if (NSPLIT > 1) tma_load_2d_pair(st + kQpAPlane, &mapAlo, &a_full[as], kb * BK, m0);
                    if (++as == kQpAStages) { as = 0; aph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 18) {
        // ===================== TMA producer: weight units =====================
        if (lane == 0) {
            int bs = 0; uint32_t bph = 0;
            for (int r = round0; r < n_rounds; r += round_step)
                for (int kb = 0; kb < n_kb; ++kb)
                    for (int u = 0; u < 3; ++u) {        // S, V0, V1: this CTA's 64 rows of the 128-row weight tile
                        mbar_wait(&b_empty[bs], bph ^ 1);
                        uint8_t* ub = b_ring + bs * kQpBUnit;
                        if (leader) mbar_expect_tx(&b_full[bs], 2 * kPl * kQpBPlane);
                        const int row = (u == 0 ? 0 : (u - 1) * 128) + (int)rank * 64;
                        tma_load_2d_pair(ub, u == 0 ? &mapShi : &mapVhi, &b_full[bs], kb * BK, row);
                        if (NSPLIT > 1) tma_load_2d_pair(ub + kQpBPlane, u == 0 ? &mapSlo : &mapVlo, &b_full[bs], kb * BK, row);
                        if (++bs == kQpBUnits) { bs = 0; bph ^= 1; }
                    }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA; whole warp runs the loop, tcgen05 on the elected lane) =====================
        if (leader) {
            constexpr uint32_t idesc = make_idesc(2 * BM, 128);
            int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
            // S buffer of this CTA's i-th round: i & 1, its barrier phase (i >> 1) & 1.  Kept as bits of one register: a
            // two-element array indexed by sbuf lands in LOCAL memory (ncu source view: 8 % of the kernel's stall samples
            // sat behind the LDL of the phase in front of the s_full wait)
            uint32_t vph = 0, sph = 0;
            int sbuf = 0;
            for (int r = round0; r < n_rounds; r += round_step) {
                mbar_wait_cluster(&s_empty[sbuf], ((sph >> sbuf) & 1u) ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait_cluster(&a_full[as], aph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(a_ring + as * kQpAStage);
                    const uint64_t ahi = make_smem_desc(a_addr), alo = make_smem_desc(a_addr + kQpAPlane);
                    for (int u = 0; u < 3; ++u) {
                        if (kb == 0 && u == 1) {         // V of the previous round must have been drained
                            mbar_wait_cluster(v_empty, vph ^ 1);
                            tc_fence_after();
                        }
                        mbar_wait_cluster(&b_full[bs], bph);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(b_ring + bs * kQpBUnit);
                        const uint64_t bhi = make_smem_desc(b_addr), blo = make_smem_desc(b_addr + kQpBPlane);
                        const uint32_t d = u == 0 ? tS + (uint32_t)sbuf * 128u : tV + (uint32_t)(u - 1) * 128u;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                const uint32_t acc = (kb | k) != 0;
                                if (NSPLIT > 1) {
                                    umma_f16_pair(d, alo + 2 * k, bhi + 2 * k, idesc, acc);      // small terms first
                                    umma_f16_pair(d, ahi + 2 * k, blo + 2 * k, idesc, 1u);
                                    umma_f16_pair(d, ahi + 2 * k, bhi + 2 * k, idesc, 1u);
                                } else {
                                    umma_f16_pair(d, ahi + 2 * k, bhi + 2 * k, idesc, acc);
                                }
                            }
                            umma_commit_pair(&b_empty[bs], 3);
                        }
                        __syncwarp();
                        if (++bs == kQpBUnits) { bs = 0; bph ^= 1; }
                    }
                    if (elect_one()) umma_commit_pair(&a_empty[as], 3);
                    __syncwarp();
                    if (++as == kQpAStages) { as = 0; aph ^= 1; }
                }
                if (elect_one()) {
                    umma_commit_pair(v_full, 3);
                    umma_commit_pair(&s_full[sbuf], 3);
                }
                __syncwarp();
                vph ^= 1;
                sph ^= 1u << sbuf;
                sbuf ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps 2..17: thread = row, four warps per lane quarter =====================
        const int ew = warp - 2;
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int part = ew >> 2;                        // value: columns [64 part, +64); sampling: heads 2 part, 2 part + 1
        float* stg = reinterpret_cast<float*>(stage_tiles + ew * kQpStageTile);
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        uint32_t vfph = 0, sfph = 0;       // bit sbuf = phase of s_full[sbuf] (a register, not a local array)
        int sbuf = 0;
        for (int r = round0; r < n_rounds; r += round_step) {
            const int m0 = (2 * r + (int)rank) * BM;
            const int wrow0 = m0 + q * 32;
            const int row = wrow0 + lane;
            const int rows_valid = M - wrow0 < 0 ? 0 : (M - wrow0 > 32 ? 32 : M - wrow0);
            const size_t srow = row < M ? (size_t)row : (size_t)(M - 1);
            PewRow pw;
            if (p.pew_early) pw = sampling_prefetch(p.samp, part, srow);
            // ---- value tile: V = acc * scale + bias ----
            // Both 32-column blocks are pulled into registers FIRST and V's TMEM columns are released at once: V is single
            // buffered, so the next round's V MMAs wait for this release — with the release after the stores (round 1) the
            // tensor pipe idled through the whole LSU-bound store phase of every round (ncu source view: 12 % of the
            // kernel's stall samples sat on the v_full wait, 37 % in the store phase).
            mbar_wait(v_full, vfph); vfph ^= 1;
            tc_fence_after();
            float va[32], vb[32];
            tmem_ld32(tV + part * 64 + lane_sel, va);
            tmem_ld32(tV + part * 64 + 32 + lane_sel, vb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(v_empty);              // V's columns may be overwritten by the next round
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                float* v = c == 0 ? va : vb;
                const int col = part * 64 + c;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias_v + col + i * 4));
                    v[i * 4 + 0] = fmaf(v[i * 4 + 0], p.scale_v, b4.x); v[i * 4 + 1] = fmaf(v[i * 4 + 1], p.scale_v, b4.y);
                    v[i * 4 + 2] = fmaf(v[i * 4 + 2], p.scale_v, b4.z); v[i * 4 + 3] = fmaf(v[i * 4 + 3], p.scale_v, b4.w);
                }
                // 16 floats = 64-byte rows: the tile geometry of 32 fp16 columns
                if (p.tma_stores) {
                    stage_tma_store_64(stg, reinterpret_cast<const uint4*>(&v[0]), &mapVout, col, wrow0, lane);
                    stage_tma_store_64(stg, reinterpret_cast<const uint4*>(&v[16]), &mapVout, col + 16, wrow0, lane);
                } else {
                    __half* dst = reinterpret_cast<__half*>(p.V + (size_t)wrow0 * kE + col);
                    stage_store_f16_32(stg, reinterpret_cast<const uint4*>(&v[0]), dst, 2 * kE, rows_valid, lane);
                    stage_store_f16_32(stg, reinterpret_cast<const uint4*>(&v[16]), dst + 32, 2 * kE, rows_valid, lane);
                }
            }
            // ---- sampling tile ----
            mbar_wait(&s_full[sbuf], (sfph >> sbuf) & 1u); sfph ^= 1u << sbuf;
            tc_fence_after();
            sampling_epilogue(p.samp, tS + (uint32_t)sbuf * 128u + lane_sel, part, wrow0, srow, rows_valid, lane, stg,
                              p.tma_stores ? &mapRec : nullptr, p.pew_early ? &pw : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&s_empty[sbuf]);
            sbuf ^= 1;
        }
        if (p.tma_stores && lane == 0) tma_store_wait_all();         // this thread's box stores are complete before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // the peer may still read this CTA's shared / tensor memory
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// mapV*: value_proj weight planes with 64-row boxes; mapS*: sampling weight planes (128 padded rows) with 64-row boxes
template <int NSPLIT>
inline cudaError_t launch_qproj_fused(const CUtensorMap& aHi, const CUtensorMap& aLo, const CUtensorMap& vHi,
                                      const CUtensorMap& vLo, const CUtensorMap& sHi, const CUtensorMap& sLo,
                                      const CUtensorMap& vOut, const CUtensorMap& recOut, int M, int K,
                                      const QprojParams& p, int num_sms, cudaStream_t st) {
    auto kern = qproj_fused_kernel<NSPLIT>;
    {   // per-device attribute
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kQpSmem);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    const int n_rounds = (M + 2 * BM - 1) / (2 * BM);
    const int pairs = n_rounds < num_sms / 2 ? n_rounds : num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kQpThreads);
    cfg.dynamicSmemBytes = kQpSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, aHi, aLo, vHi, vLo, sHi, sLo, vOut, recOut, M, K, p);
}

}  // namespace tc
}  // namespace ddp
