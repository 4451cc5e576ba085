// Fused FFN block for sm_100a:  q <- FiLM(LN2(q + W2 gelu(W1 q + b1) + b2))   for a 128-token tile per CTA,
// with the 1024-wide hidden activation never leaving the SM (the unfused pair writes + reads 8 KB/token of it).
//
//   A1 = q planes (128 x 256, fp16 hi/lo)        shared memory, loaded once per tile by TMA
//   for each chunk c of 128 hidden columns (8 chunks):
//       D1 (TMEM, 128 x 128 fp32) = A1 * W1[c]^T                       tcgen05.mma, A and B from shared memory
//       hid = gelu(D1 * s + b1)  -> fp16 hi/lo planes written back into TMEM (tcgen05.st, 2 halves per column)
//       D2 (TMEM, 128 x 256 fp32) += hid * W2[:, c]^T                  tcgen05.mma with the A operand FROM TMEM
//   D2 += A1 * (2^shift I)^T        (the residual, through the identity block of the augmented W2)
//   epilogue: LayerNorm + folded FiLM, fp16 planes (and optionally fp32) of the new q
//
// TMEM (512 columns): D2 [0,256) | D1 [256,384) | hid planes [384,512) (hi 64 + lo 64 columns).
// An SS-mode N=64 MMA is operand-fetch bound (48 instead of 32 cycles, tools/ubench_mma.cu) while N=128 runs at the math
// rate: hence 128-wide hidden chunks, at the price of single-buffered D1 / hid (TMEM is full).
// Warps: 0 TMA producer (A1 + a ring of [hi | lo] weight units), 1 MMA issuer, 2..17 epilogue (thread = row, four
// warps per TMEM lane quarter).  MMA1 of chunk c+1 overlaps the GELU epilogue and MMA2 of chunk c.
// By default the kernel runs on CTA pairs (template PAIR, see below): 256 tokens per pair, cta_group::2 MMAs.
#pragma once
#include "gemm_tc.cuh"

namespace ddp {
namespace tc {

struct FfnParams {
    float s1_16;             // 16 / (2^shift1 * 16): turns the D1 accumulator into 16 * (W1 q)
    float s2;                // 1 / (2^shift2 * 16)
    const float* b1;         // [1024]
    const float* b2;         // [256]
    const float* ln_g;       // [256] LN2 gamma * (film scale + 1)
    const float* ln_b;       // [256] LN2 beta * (film scale + 1) + film shift
    SplitOut split;          // planes of the new q
    float* out;              // optional fp32 copy [M][256]
    unsigned long long* dbg; // optional [8] cycle counters of the MMA issuer (profiling aid)
    int tma_stores;          // 1: the new q planes leave through TMA box stores (mapOhi / mapOlo) from the staging tiles
    // 1: chunk 0 of the NEXT tile is pulled ahead of the current tile's last MMA2 and LayerNorm: its MMA1 is issued between
    // MMA2(6) and MMA2(7), its GELU runs while MMA2(7) executes, and only then do the epilogue warps turn to the LayerNorm.
    // Without it the tensor pipe waits at every tile boundary for LN(t) AND GELU(t+1, 0), which share the same 16 warps.
    int pull_ahead;
};

constexpr int kFfnThreads = 32 * 18;
constexpr int kFfnUnit = 16384;
constexpr int kFfnRingBytes = 5 * kFfnUnit;   // weight tiles in flight (a unit = hi + lo plane of one [128 x 64] tile, or a pair member's halves)
constexpr int kFfnStageTile = 1024;       // per-warp store staging tile (32 rows x 32 bytes)
constexpr int kFfnStageArea = 16 * kFfnStageTile;
constexpr int kFfnSmem = 8 * kFfnUnit + kFfnRingBytes + kFfnStageArea + 1024 + 256;
static_assert(kFfnSmem <= 232448, "fused FFN kernel exceeds shared memory");

// tcgen05.mma with the A operand in tensor memory (lane = row, two fp16 K-elements per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// PAIR: two CTAs of a cluster (one TPC) work on 256 tokens with cta_group::2 MMAs (M = 256).  Each CTA keeps its own
// 128 rows of q / D1 / hidden planes / D2 and loads only HALF of every weight tile (its 64 of the 128 N rows), which
// halves the L2 -> SM weight stream that bounds the single-CTA kernel.  The leader (cluster rank 0) issues all MMAs;
// TMA loads of both CTAs are credited to the leader's barriers, commits are multicast to both CTAs, and the epilogue
// warps of both CTAs arrive on the leader's barriers (CTA-pair wrappers in gemm_tc.cuh).
template <int NSPLIT, bool PAIR, bool DBG>
__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap mapA1hi, const __grid_constant__ CUtensorMap mapA1lo,
                 const __grid_constant__ CUtensorMap mapW1hi, const __grid_constant__ CUtensorMap mapW1lo,
                 const __grid_constant__ CUtensorMap mapW2hi, const __grid_constant__ CUtensorMap mapW2lo,
                 const __grid_constant__ CUtensorMap mapOhi, const __grid_constant__ CUtensorMap mapOlo,
                 int M, FfnParams p) {
    constexpr int kChunks = kFFN / 128;                 // 8
    constexpr int kPl = NSPLIT > 1 ? 2 : 1;             // planes per operand
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a1 = smem;                                  // [plane][kb] 16 KB tiles
    uint8_t* ring = smem + 8 * kFfnUnit;
    constexpr int kPlaneB = PAIR ? kFfnUnit / 2 : kFfnUnit;  // bytes of one plane of a weight tile in THIS CTA's shared memory
    constexpr int kUnitB = 2 * kPlaneB;                      // ring unit: [hi plane | lo plane] (or the two N halves of the identity)
    constexpr int kFfnRing = kFfnRingBytes / kUnitB;         // 5 for a pair, 2 for a single CTA
    constexpr int kNCta = PAIR ? 2 : 1;
    uint8_t* stage_tiles = ring + kFfnRingBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_tiles + kFfnStageArea);
    uint64_t* a1_full = bars + 0;
    uint64_t* a1_empty = bars + 1;
    uint64_t* ring_full = bars + 2;                      // [kFfnRing <= 10]
    uint64_t* ring_empty = bars + 12;                    // [kFfnRing <= 10]
    uint64_t* d1_full = bars + 22;
    uint64_t* d1_empty = bars + 23;
    uint64_t* a2_full = bars + 24;
    uint64_t* a2_empty = bars + 25;
    uint64_t* d2_full = bars + 26;
    uint64_t* d2_empty = bars + 27;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
    static_assert(kFfnRing <= 10, "barrier slots");
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // tile = BM rows per CTA; a pair takes two consecutive tiles (256 rows) per round
    const int n_rounds = PAIR ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM;
    const int round0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int round_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto row_of = [&](int r) { return PAIR ? (2 * r + (int)rank) * BM : r * BM; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA1hi); tma_prefetch_desc(&mapW1hi); tma_prefetch_desc(&mapW2hi);
        if (NSPLIT > 1) { tma_prefetch_desc(&mapA1lo); tma_prefetch_desc(&mapW1lo); tma_prefetch_desc(&mapW2lo); }
    }
    if (warp == 1 && lane == 0) {
        mbar_init(a1_full, 1); mbar_init(a1_empty, 1);
        for (int s = 0; s < kFfnRing; ++s) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], 1); }
        mbar_init(d1_full, 1); mbar_init(d1_empty, 16 * kNCta);
        mbar_init(a2_full, 16 * kNCta); mbar_init(a2_empty, 1);
        mbar_init(d2_full, 1); mbar_init(d2_empty, 16 * kNCta);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (PAIR) {                                          // barriers of both CTAs exist before anyone signals them
        __syncthreads();
        cluster_sync_all();
        if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
    } else {
        if (warp == 2) tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tD2 = tmem_base, tD1 = tmem_base + 256, tA2 = tmem_base + 384;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t rphase = 0; uint32_t tphase = 0;
            auto put = [&](const CUtensorMap* mapa, int ca, const CUtensorMap* mapb, int cb, int c0) {
                // one ring unit = up to two [128 N rows][64 k] tiles, back to back (a pair member loads its 64 rows of each)
                mbar_wait(&ring_empty[stage], rphase ^ 1);
                uint8_t* st = ring + stage * kUnitB;
                const int bytes = (mapb ? 2 : 1) * kPlaneB;
                if (PAIR) {
                    if (leader) mbar_expect_tx(&ring_full[stage], 2 * bytes);
                    tma_load_2d_pair(st, mapa, &ring_full[stage], c0, ca + (int)rank * 64);
                    if (mapb) tma_load_2d_pair(st + kPlaneB, mapb, &ring_full[stage], c0, cb + (int)rank * 64);
                } else {
                    mbar_expect_tx(&ring_full[stage], bytes);
                    tma_load_2d(st, mapa, &ring_full[stage], c0, ca);
                    if (mapb) tma_load_2d(st + kPlaneB, mapb, &ring_full[stage], c0, cb);
                }
                if (++stage == kFfnRing) { stage = 0; rphase ^= 1; }
            };
            auto put_w2 = [&](int cc) {       // W2 columns of hidden chunk cc, per 64-wide K block and output half: hi, lo
                for (int kb2 = 0; kb2 < 2; ++kb2)
                    for (int nh = 0; nh < 2; ++nh)
                        put(&mapW2hi, nh * 128, NSPLIT > 1 ? &mapW2lo : nullptr, nh * 128, cc * 128 + kb2 * 64);
            };
            auto load_a1 = [&](int m0) {                 // q planes of one tile (waits until the previous tile's MMAs released them)
                mbar_wait(a1_empty, tphase ^ 1);
                tphase ^= 1;
                if (!PAIR || leader) mbar_expect_tx(a1_full, kNCta * kPl * 4 * kFfnUnit);
                for (int kb = 0; kb < 4; ++kb) {
                    if (PAIR) {
                        tma_load_2d_pair(a1 + kb * kFfnUnit, &mapA1hi, a1_full, kb * BK, m0);
                        if (NSPLIT > 1) tma_load_2d_pair(a1 + (4 + kb) * kFfnUnit, &mapA1lo, a1_full, kb * BK, m0);
                    } else {
                        tma_load_2d(a1 + kb * kFfnUnit, &mapA1hi, a1_full, kb * BK, m0);
                        if (NSPLIT > 1) tma_load_2d(a1 + (4 + kb) * kFfnUnit, &mapA1lo, a1_full, kb * BK, m0);
                    }
                }
            };
            auto put_w1 = [&](int c) {                   // W1 rows of hidden chunk c (128 rows), per K block: hi | lo
                for (int kb = 0; kb < 4; ++kb) put(&mapW1hi, c * 128, NSPLIT > 1 ? &mapW1lo : nullptr, c * 128, kb * BK);
            };
            bool pulled = false;                         // this tile's planes + W1 chunk 0 were queued at the end of the previous tile
            for (int r = round0; r < n_rounds; r += round_step) {
                if (!pulled) load_a1(row_of(r));
                for (int c = 0; c < kChunks; ++c) {
                    if (!(c == 0 && pulled)) put_w1(c);
                    if (c == kChunks - 1)                // identity block of the augmented W2 (hi plane only): the residual
                        for (int kb = 0; kb < 4; ++kb)   // is issued right after the last MMA1 (see the MMA warp); unit = both N halves
                            put(&mapW2hi, 0, &mapW2hi, 128, kFFN + kb * BK);
                    if (c >= 1) put_w2(c - 1);
                }
                pulled = p.pull_ahead && r + round_step < n_rounds;
                if (pulled) {                            // the ring is FIFO: same order as the MMA warp consumes (… M2(6), M1'(0), M2(7))
                    load_a1(row_of(r + round_step));
                    put_w1(0);
                }
                put_w2(kChunks - 1);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp runs the loop; tcgen05 instructions on the elected lane) =====================
        if (leader) {
            constexpr uint32_t idesc128 = make_idesc(kNCta * BM, 128);
            auto mma_ss = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
                if (PAIR) umma_f16_pair(d, a, b, idesc128, acc); else umma_f16(d, a, b, idesc128, acc);
            };
            auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t acc) {
                if (PAIR) umma_f16_ts_pair(d, a, b, idesc128, acc); else umma_f16_ts(d, a, b, idesc128, acc);
            };
            auto commit = [&](uint64_t* bar) { if (PAIR) umma_commit_pair(bar, 3); else umma_commit(bar); };
            auto wait = [&](uint64_t* bar, uint32_t parity) { if (PAIR) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity); };
            int stage = 0; uint32_t rphase = 0, tphase = 0;
            uint32_t d1e_phase = 0, a2f_phase = 0, d2e_phase = 0;
            const uint32_t a1_addr = smem_u32(a1);
            auto now = [&]() -> long long { return DBG ? clock64() : 0ll; };
            long long tw_ring = 0, tw_d1e = 0, tw_a2f = 0, tw_d2e = 0, tw_a1 = 0, tw_i1 = 0, tw_i2 = 0, t_all = now();
            // wait for the next ring unit; returns the address of its first tile, the second follows at +kPlaneB
            auto ring_wait = [&]() -> uint32_t {
                long long t0 = now();
                wait(&ring_full[stage], rphase);
                tc_fence_after();
                tw_ring += now() - t0;
                return smem_u32(ring + stage * kUnitB);
            };
            auto ring_release = [&]() {
                if (elect_one()) commit(&ring_empty[stage]);
                __syncwarp();
                if (++stage == kFfnRing) { stage = 0; rphase ^= 1; }
            };
            auto mma2 = [&](int cc) {                    // D2 += hid(cc) * W2[:, cc]^T, hid planes in TMEM
                { long long t0 = now(); wait(a2_full, a2f_phase); a2f_phase ^= 1; tw_a2f += now() - t0; }
                tc_fence_after();
                const uint32_t ahi = tA2, alo = tA2 + 64;
                for (int kb2 = 0; kb2 < 2; ++kb2)
                    for (int nh = 0; nh < 2; ++nh) {
                        const uint32_t uhi = ring_wait();
                        const uint64_t bhi = make_smem_desc(uhi), blo = make_smem_desc(uhi + kPlaneB);
                        const uint32_t d = tD2 + nh * 128;
                        const long long ti0 = now();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t acc = (cc | kb2 | k) != 0;
                                const uint32_t ka = 8 * (kb2 * 4 + k);
                                if (NSPLIT > 1) {
                                    mma_ts(d, alo + ka, bhi + 2 * k, acc);
                                    mma_ts(d, ahi + ka, blo + 2 * k, 1u);
                                    mma_ts(d, ahi + ka, bhi + 2 * k, 1u);
                                } else {
                                    mma_ts(d, ahi + ka, bhi + 2 * k, acc);
                                }
                            }
                        }
                        __syncwarp();
                        tw_i2 += now() - ti0;
                        ring_release();
                    }
                if (elect_one()) commit(a2_empty);
                __syncwarp();
            };
            auto mma1 = [&](int c) {                     // D1 = A1 * W1[c]^T   (N = 128)
                (void)c;
                { long long t0 = now(); wait(d1_empty, d1e_phase ^ 1); d1e_phase ^= 1; tw_d1e += now() - t0; }
                tc_fence_after();
                for (int kb = 0; kb < 4; ++kb) {
                    const uint32_t uhi = ring_wait();
                    const uint64_t ahi = make_smem_desc(a1_addr + kb * kFfnUnit);
                    const uint64_t alo = make_smem_desc(a1_addr + (4 + kb) * kFfnUnit);
                    const uint64_t bhi = make_smem_desc(uhi), blo = make_smem_desc(uhi + kPlaneB);
                    const long long ti0 = now();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t acc = (kb | k) != 0;
                            if (NSPLIT > 1) {
                                mma_ss(tD1, alo + 2 * k, bhi + 2 * k, acc);
                                mma_ss(tD1, ahi + 2 * k, blo + 2 * k, 1u);
                                mma_ss(tD1, ahi + 2 * k, bhi + 2 * k, 1u);
                            } else {
                                mma_ss(tD1, ahi + 2 * k, bhi + 2 * k, acc);
                            }
                        }
                    }
                    __syncwarp();
                    tw_i1 += now() - ti0;
                    ring_release();
                }
                if (elect_one()) commit(d1_full);
                __syncwarp();
            };
            auto wait_a1 = [&]() {
                long long t0 = now();
                wait(a1_full, tphase);
                tphase ^= 1;
                tw_a1 += now() - t0;
                tc_fence_after();
            };
            bool pulled = false;                         // MMA1(0) of this tile was already issued at the end of the previous one
            for (int r = round0; r < n_rounds; r += round_step) {
                if (!pulled) wait_a1();
                const bool pull_next = p.pull_ahead && r + round_step < n_rounds;
                for (int c = 0; c <= kChunks; ++c) {
                    if (c < kChunks && !(c == 0 && pulled)) mma1(c);
                    if (c == kChunks - 1) {              // last MMA1 issued: add the residual now and release q's planes early
                    // residual: D2 += q * (2^shift I)^T with q's planes still in shared memory
                    for (int kb = 0; kb < 4; ++kb) {
                        const uint64_t ahi = make_smem_desc(a1_addr + kb * kFfnUnit);
                        const uint64_t alo = make_smem_desc(a1_addr + (4 + kb) * kFfnUnit);
                        const uint32_t ui = ring_wait();
                        if (elect_one()) {
                            // K block kb of the identity holds 2^shift on rows (= output channels) [64 kb, 64 kb + 64) and
                            // zeros elsewhere: only output half nh = kb / 2 receives anything, the other half's MMAs
                            // would add exact zeros and are skipped (half of the residual's tensor work)
                            const int nh = kb >> 1;
                            const uint64_t bi = make_smem_desc(ui + nh * kPlaneB);
                            const uint32_t d = tD2 + nh * 128;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (NSPLIT > 1) mma_ss(d, alo + 2 * k, bi + 2 * k, 1u);
                                mma_ss(d, ahi + 2 * k, bi + 2 * k, 1u);
                            }
                        }
                        __syncwarp();
                        ring_release();
                    }
                        if (elect_one()) commit(a1_empty);   // q planes may be overwritten by the next tile's load
                        __syncwarp();
                    }
                    if (c == kChunks && pull_next) {     // between MMA2(6) and MMA2(7): chunk 0 of the next tile
                        wait_a1();
                        mma1(0);
                    }
                    if (c >= 1) {
                        if (c == 1) {                    // D2 of the previous tile must have been drained
                            { long long t0 = now(); wait(d2_empty, d2e_phase ^ 1); d2e_phase ^= 1; tw_d2e += now() - t0; }
                            tc_fence_after();
                        }
                        mma2(c - 1);
                    }
                }
                if (elect_one()) commit(d2_full);
                __syncwarp();
                pulled = pull_next;
            }
            if (DBG && p.dbg && lane == 0) {
                atomicAdd(&p.dbg[0], (unsigned long long)(now() - t_all));
                atomicAdd(&p.dbg[1], (unsigned long long)tw_ring);
                atomicAdd(&p.dbg[2], (unsigned long long)tw_d1e);
                atomicAdd(&p.dbg[3], (unsigned long long)tw_a2f);
                atomicAdd(&p.dbg[4], (unsigned long long)tw_d2e);
                atomicAdd(&p.dbg[5], (unsigned long long)tw_a1);
                atomicAdd(&p.dbg[6], (unsigned long long)tw_i1);
                atomicAdd(&p.dbg[7], (unsigned long long)tw_i2);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps 2..17: thread = row =====================
        const int ew = warp - 2;
        const int q = warp & 3;
        const int part = ew >> 2;                        // which quarter of the columns
        float* stg = reinterpret_cast<float*>(stage_tiles + ew * kFfnStageTile);
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        uint32_t d1f_phase = 0, a2e_phase = 0;
        uint32_t d2f_phase = 0;
        auto arrive = [&](uint64_t* bar) { if (PAIR) mbar_arrive_leader(bar); else mbar_arrive(bar); };
        // GELU of this warp's 32 columns of hidden chunk c: D1 -> registers (D1 released at once) -> fp16 hi/lo planes back into TMEM
        auto gelu_chunk = [&](int c) {
                mbar_wait(d1_full, d1f_phase); d1f_phase ^= 1;
                tc_fence_after();
                float v[32];
                tmem_ld32(tD1 + part * 32 + lane_sel, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive(d1_empty);                // D1 may be overwritten by MMA1 of chunk c + 1
                uint32_t hi[16], lo[16];
                const float* bp = p.b1 + c * 128 + part * 32;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + i * 4));
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                    float g[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float z16 = fmaf(v[i * 4 + e], p.s1_16, bb[e] * kActScale);
                        const float u = fabsf(z16) * (0.70710678118654752440f * kInvActScale);
                        float t, ex;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.0f)));
                        float pl = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
                        pl = fmaf(pl, t, 0.5f * 1.421413741f);
                        pl = fmaf(pl, t, 0.5f * -0.284496736f);
                        pl = fmaf(pl, t, 0.5f * 0.254829592f);
                        pl *= t;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(u * u * -1.44269504088896340736f));
                        const float hh = pl * ex;
                        g[e] = z16 * (z16 < 0.f ? hh : 1.0f - hh);
                    }
                    const __half2 h01 = __floats2half2_rn(g[0], g[1]), h23 = __floats2half2_rn(g[2], g[3]);
                    hi[i * 2] = *reinterpret_cast<const uint32_t*>(&h01);
                    hi[i * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                    if (NSPLIT > 1) {
                        const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn(g[0] - b01.x, g[1] - b01.y);
                        const __half2 l23 = __floats2half2_rn(g[2] - b23.x, g[3] - b23.y);
                        lo[i * 2] = *reinterpret_cast<const uint32_t*>(&l01);
                        lo[i * 2 + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                    }
                }
                // the hid planes must have been consumed by MMA2 of chunk c - 1
                mbar_wait(a2_empty, a2e_phase ^ 1); a2e_phase ^= 1;
                tc_fence_after();
                tmem_st8(tA2 + part * 16 + lane_sel, hi);
                tmem_st8(tA2 + part * 16 + 8 + lane_sel, hi + 8);
                if (NSPLIT > 1) {
                    tmem_st8(tA2 + 64 + part * 16 + lane_sel, lo);
                    tmem_st8(tA2 + 64 + part * 16 + 8 + lane_sel, lo + 8);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive(a2_full);
                    };
        bool pulled = false;                             // chunk 0 of this tile was processed at the end of the previous tile
        for (int r = round0; r < n_rounds; r += round_step) {
            const int m0 = row_of(r);
            const int wrow0 = m0 + q * 32;
            const int rows_valid = M - wrow0 < 0 ? 0 : (M - wrow0 > 32 ? 32 : M - wrow0);
            // ---- per hidden chunk: GELU of this warp's 32 columns, planes back into TMEM ----
#pragma unroll 1
            for (int c = pulled ? 1 : 0; c < kChunks; ++c) gelu_chunk(c);
            // chunk 0 of the NEXT tile before this tile's LayerNorm: its math runs while MMA2(7) executes (FfnParams::pull_ahead)
            pulled = p.pull_ahead && r + round_step < n_rounds;
            if (pulled) gelu_chunk(0);
            // ---- final: LayerNorm (+ folded FiLM) of this warp's 64 columns of D2 ----
            mbar_wait(d2_full, d2f_phase); d2f_phase ^= 1;
            tc_fence_after();
            const uint32_t t_row = tD2 + part * 64 + lane_sel;
            float mean = 0.f, m2 = 0.f;
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                float v[32];
                tmem_ld32(t_row + c, v);
                float cs = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + part * 64 + c + i * 4));
                    v[i * 4 + 0] = fmaf(v[i * 4 + 0], p.s2, b4.x); v[i * 4 + 1] = fmaf(v[i * 4 + 1], p.s2, b4.y);
                    v[i * 4 + 2] = fmaf(v[i * 4 + 2], p.s2, b4.z); v[i * 4 + 3] = fmaf(v[i * 4 + 3], p.s2, b4.w);
                    cs += (v[i * 4 + 0] + v[i * 4 + 1]) + (v[i * 4 + 2] + v[i * 4 + 3]);
                }
                tmem_st32(t_row + c, v);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                const float cm = cs * (1.0f / 32.0f);
                float cm2 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) { const float d = v[i] - cm; cm2 = fmaf(d, d, cm2); }
                const float na = (float)c, nb = 32.0f, nab = na + nb;
                const float delta = cm - mean;
                mean += delta * (nb / nab);
                m2 += cm2 + delta * delta * (na * nb / nab);
            }
            {   // combine the four column quarters of each row: every warp of the lane quarter gets all four partials
                for (int w = 0; w < 4; ++w) {
                    float2* dst = reinterpret_cast<float2*>(stage_tiles + (w * 4 + (ew & 3)) * kFfnStageTile);
                    dst[part * 32 + lane] = make_float2(mean, m2);
                }
                named_bar_sync(1 + q, 128);
                const float2* mine = reinterpret_cast<const float2*>(stg);
                float mu = 0.f, s2v = 0.f, n = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {            // fixed order: identical result in all four warps
                    const float2 o = mine[w * 32 + lane];
                    const float nb = 64.0f, nab = n + nb;
                    const float delta = o.x - mu;
                    mu += delta * (nb / nab);
                    s2v += o.y + delta * delta * (n * nb / nab);
                    n = nab;
                }
                mean = mu; m2 = s2v;
                __syncwarp();
            }
            const float rstd = 1.0f / sqrtf(m2 * (1.0f / kE) + 1e-5f);
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                float v[32];
                tmem_ld32(t_row + c, v);
                const int col = part * 64 + c;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col + i * 4));
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col + i * 4));
                    v[i * 4 + 0] = fmaf((v[i * 4 + 0] - mean) * rstd, g4.x, b4.x);
                    v[i * 4 + 1] = fmaf((v[i * 4 + 1] - mean) * rstd, g4.y, b4.y);
                    v[i * 4 + 2] = fmaf((v[i * 4 + 2] - mean) * rstd, g4.z, b4.z);
                    v[i * 4 + 3] = fmaf((v[i * 4 + 3] - mean) * rstd, g4.w, b4.w);
                }
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {          // 16 columns (32-byte plane rows) per staging round
                    uint4 hi[2], lo[2];
                    __half2* h2 = reinterpret_cast<__half2*>(hi);
                    __half2* l2 = reinterpret_cast<__half2*>(lo);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float a = v[sub * 16 + 2 * i] * kActScale, bq = v[sub * 16 + 2 * i + 1] * kActScale;
                        const __half2 hh = __floats2half2_rn(a, bq);
                        h2[i] = hh;
                        if (NSPLIT > 1) {
                            const float2 back = __half22float2(hh);
                            l2[i] = __floats2half2_rn(a - back.x, bq - back.y);
                        }
                    }
                    if (p.tma_stores) {      // one TMA box store per 32 x 32-byte piece instead of LDS + STG by every thread
                        stage_tma_store_32(stg, hi, &mapOhi, col + sub * 16, wrow0, lane);
                        if (NSPLIT > 1) stage_tma_store_32(stg, lo, &mapOlo, col + sub * 16, wrow0, lane);
                    } else {
                        const size_t o = (size_t)wrow0 * p.split.ld + col + sub * 16;
                        stage_store_32b(stg, hi, p.split.hi + o, p.split.ld, rows_valid, lane);
                        if (NSPLIT > 1) stage_store_32b(stg, lo, p.split.lo + o, p.split.ld, rows_valid, lane);
                    }
                }
                if (p.out) {                                 // fp32 copy for test taps: 8 columns (32-byte rows) at a time
                    if (p.tma_stores) stage_tma_sync(lane);
#pragma unroll
                    for (int s8 = 0; s8 < 4; ++s8)
                        stage_store_32b(stg, reinterpret_cast<const uint4*>(&v[s8 * 8]),
                                        reinterpret_cast<__half*>(p.out + (size_t)wrow0 * kE + col + s8 * 8), 2 * kE, rows_valid, lane);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive(d2_empty);
            if (p.tma_stores) stage_tma_sync(lane);      // the TMA unit has read this warp's tile ...
            named_bar_sync(1 + q, 128);                  // ... before partners reuse it for the next tile's partials
        }
        if (p.tma_stores && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();                        // the peer may still read this CTA's shared / tensor memory
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

template <int NSPLIT, bool PAIR, bool DBG>
inline cudaError_t launch_ffn_fused_(const CUtensorMap& a1Hi, const CUtensorMap& a1Lo, const CUtensorMap& w1Hi,
                                    const CUtensorMap& w1Lo, const CUtensorMap& w2Hi, const CUtensorMap& w2Lo,
                                    const CUtensorMap& oHi, const CUtensorMap& oLo, int M,
                                    const FfnParams& p, int num_sms, cudaStream_t st) {
    auto kern = ffn_fused_kernel<NSPLIT, PAIR, DBG>;
    {   // per-device attribute
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    if (!PAIR) {
        const int n_tiles = (M + BM - 1) / BM;
        const int grid = n_tiles < num_sms ? n_tiles : num_sms;
        kern<<<grid, kFfnThreads, kFfnSmem, st>>>(a1Hi, a1Lo, w1Hi, w1Lo, w2Hi, w2Lo, oHi, oLo, M, p);
        return cudaGetLastError();
    }
    const int n_rounds = (M + 2 * BM - 1) / (2 * BM);
    const int pairs = n_rounds < num_sms / 2 ? n_rounds : num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kFfnThreads);
    cfg.dynamicSmemBytes = kFfnSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a1Hi, a1Lo, w1Hi, w1Lo, w2Hi, w2Lo, oHi, oLo, M, p);
}

template <int NSPLIT, bool PAIR>
inline cudaError_t launch_ffn_fused(const CUtensorMap& a1Hi, const CUtensorMap& a1Lo, const CUtensorMap& w1Hi,
                                    const CUtensorMap& w1Lo, const CUtensorMap& w2Hi, const CUtensorMap& w2Lo,
                                    const CUtensorMap& oHi, const CUtensorMap& oLo, int M,
                                    const FfnParams& p, int num_sms, cudaStream_t st) {
    return p.dbg ? launch_ffn_fused_<NSPLIT, PAIR, true>(a1Hi, a1Lo, w1Hi, w1Lo, w2Hi, w2Lo, oHi, oLo, M, p, num_sms, st)
                 : launch_ffn_fused_<NSPLIT, PAIR, false>(a1Hi, a1Lo, w1Hi, w1Lo, w2Hi, w2Lo, oHi, oLo, M, p, num_sms, st);
}

}  // namespace tc
}  // namespace ddp
