// CUDA backend + C ABI of the neck (FPN + MultiStageMerging) in front of the decode loop — SURVEY 8f #2.
// The launch sequence, kernel bodies, weight repack and workspace carve-up live in neck_plan.h (shared with the
// host emulation the CPU tests use); this file supplies the launches: generic one-thread-per-element kernels for the
// functors, the fp32 CUDA-core GEMM of gemm_simt.cuh for the convolutions (first correct path: fp32 like the
// reference's neck; a tcgen05 implicit-GEMM for the 3x3 convolutions is the open kernel work) and
// the tiled transposing copy of kernels.cuh for the NCHW outputs.
//
// Included at the end of ddp_b200.cu (one translation unit: the kernels of kernels.cuh are not inline).
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ddp_b200.h"
#include "gemm_simt.cuh"
#include "kernels.cuh"
#include "neck_plan.h"

namespace ddp {
namespace neck {

template <class F>
__global__ void __launch_bounds__(256) k_for_each(F f, size_t n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) f(idx);
}

// ---- tensor-core path of the 3x3 convolutions (85 % of the neck's FLOPs): tc_3xf16 arithmetic of the decode loop ----
// The convolution is an implicit GEMM of the loop's tcgen05 kernel (gemm_tc.cuh, EpiParams::conv_pitch): the normalised
// level map is written once as fp16 hi/lo planes on a ZERO-BORDERED token grid [(H + 2) * (W + 2)], every (tap, 64-channel)
// K block is one TMA box load of the same planes at a shifted row coordinate (no im2col buffer), the accumulator lives in
// TMEM, and the bordered fp32 result is compacted back to the token-major map the GroupNorm kernels read.
struct PadSplit {            // idx = bordered token * 32 + 8-channel group
    const float* src; __half* hi; __half* lo; int H, W;
    __device__ __forceinline__ void operator()(size_t idx) const {
        const int c8 = (int)(idx & 31);
        const size_t t = idx >> 5;
        const int Wp = W + 2, Np = (H + 2) * Wp;
        const size_t b = t / Np;
        const int n = (int)(t - b * Np);
        const int ip = n / Wp, jp = n - ip * Wp;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (ip >= 1 && ip <= H && jp >= 1 && jp <= W) {
            const float* sp = src + ((b * H + (ip - 1)) * (size_t)W + (jp - 1)) * kC + c8 * 8;
            const float4 a = *reinterpret_cast<const float4*>(sp), c = *reinterpret_cast<const float4*>(sp + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        }
        split8_store(v, hi + t * kC + c8 * 8, lo + t * kC + c8 * 8);
    }
};
struct Unpad {               // idx = token * 64 + float4 group
    const float* src; float* dst; int H, W;
    __device__ __forceinline__ void operator()(size_t idx) const {
        const int c4 = (int)(idx & 63);
        const size_t t = idx >> 6;
        const int N = H * W, Wp = W + 2;
        const size_t b = t / N;
        const int n = (int)(t - b * N);
        const int i = n / W, j = n - i * W;
        const size_t tp = b * (size_t)(H + 2) * Wp + (size_t)(i + 1) * Wp + (j + 1);
        *reinterpret_cast<float4*>(dst + t * kC + c4 * 4) = *reinterpret_cast<const float4*>(src + tp * kC + c4 * 4);
    }
};

// GroupNorm statistics, second stage: one WARP per (image, group) instead of one thread (the one-thread form walked
// 512 chunks x 8 channels of fp64 partials serially: 0.45 ms per launch, 21 % of the neck, for 256 threads of work).
__global__ void __launch_bounds__(256) k_gn_finalize_warp(GnFinalize f, size_t n) {
    const size_t idx = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= n) return;
    const int g = (int)(idx % f.G);
    const size_t b = idx / f.G;
    const int cpg = f.C / f.G;
    double s = 0.0, ss = 0.0;
    for (int e = lane; e < f.chunks * cpg; e += 32) {
        const int chunk = e / cpg, k = e - chunk * cpg;
        const size_t p = ((b * f.chunks + chunk) * f.C + (size_t)g * cpg + k) * 2;
        s += f.part[p];
        ss += f.part[p + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) {
        const double cnt = (double)f.N * cpg;
        const double mean = s / cnt;
        double var = ss / cnt - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        f.stats[idx * 2] = (float)mean;
        f.stats[idx * 2 + 1] = (float)(1.0 / sqrt(var + (double)f.eps));
    }
}

// GnApply / MsmMerge, four channels per thread (float4 loads and stores; per-element arithmetic identical to the functors
// of neck_plan.h, which the host emulation keeps using).  The scalar forms were 25 % + 16 % of the neck after the
// convolutions moved to the tensor cores.
__global__ void __launch_bounds__(256) k_gn_apply4(GnApply f, size_t n4) {
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const int C4 = f.C / 4;
    const int c = (int)(i4 % C4) * 4;
    const size_t t = i4 / C4;
    const int n = (int)(t % f.N);
    const size_t b = t / f.N;
    const int g = c / (f.C / f.G);                       // the four channels share a group (C / G is a multiple of 4)
    const float mean = f.stats[(b * f.G + g) * 2], rstd = f.stats[(b * f.G + g) * 2 + 1];
    const float4 x = *reinterpret_cast<const float4*>(f.src + i4 * 4);
    const float4 ga = *reinterpret_cast<const float4*>(f.gamma + c), be = *reinterpret_cast<const float4*>(f.beta + c);
    float4 y;
    y.x = (x.x - mean) * rstd * ga.x + be.x; y.y = (x.y - mean) * rstd * ga.y + be.y;
    y.z = (x.z - mean) * rstd * ga.z + be.z; y.w = (x.w - mean) * rstd * ga.w + be.w;
    if (f.up) {
        const int i = n / f.W, j = n - i * f.W;
        const int iu = nearest_src(i, f.sh, f.Hu), ju = nearest_src(j, f.sw, f.Wu);
        const float4 u = *reinterpret_cast<const float4*>(f.up + ((b * f.Hu + iu) * (size_t)f.Wu + ju) * f.C + c);
        y.x += u.x; y.y += u.y; y.z += u.z; y.w += u.w;
    }
    *reinterpret_cast<float4*>(f.dst + i4 * 4) = y;
}
__global__ void __launch_bounds__(256) k_msm_merge4(MsmMerge m, size_t n4) {
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const int C4 = m.C / 4;
    const int c = (int)(i4 % C4) * 4;
    const size_t t = i4 / C4;
    const int N0 = m.H[0] * m.W[0];
    const int n = (int)(t % N0);
    const size_t b = t / N0;
    const int i = n / m.W[0], j = n - i * m.W[0];
    float4 acc = *reinterpret_cast<const float4*>(m.z[0] + i4 * 4);
#pragma unroll
    for (int l = 1; l < kMaxLevels; ++l) {
        if (l >= m.L) break;
        const int h = m.H[l], w = m.W[l];
        int y0, y1, x0, x1; float ly, lx;
        bilinear_src(i, (float)h / (float)m.H[0], h, y0, y1, ly);
        bilinear_src(j, (float)w / (float)m.W[0], w, x0, x1, lx);
        const float hy = 1.0f - ly, hx = 1.0f - lx;
        const float* p = m.z[l] + b * (size_t)h * w * m.C + c;
        const float4 p00 = *reinterpret_cast<const float4*>(p + ((size_t)y0 * w + x0) * m.C), p01 = *reinterpret_cast<const float4*>(p + ((size_t)y0 * w + x1) * m.C);
        const float4 p10 = *reinterpret_cast<const float4*>(p + ((size_t)y1 * w + x0) * m.C), p11 = *reinterpret_cast<const float4*>(p + ((size_t)y1 * w + x1) * m.C);
        acc.x += hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x);
        acc.y += hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y);
        acc.z += hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z);
        acc.w += hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w);
    }
    *reinterpret_cast<float4*>(m.dst + i4 * 4) = acc;
}

struct NeckTc {
    bool on = false;
    int num_sms = 148;
    __half* w_arena = nullptr;                      // per level: hi [256][9 * 256], lo [256][9 * 256], k = tap * 256 + ci
    CUtensorMap w_hi[kMaxLevels], w_lo[kMaxLevels];
    float inv_scale[kMaxLevels] = {1.f, 1.f, 1.f, 1.f};
};

struct CudaBackend {
    cudaStream_t st;
    const NeckTc* tc = nullptr;
    const Weights* w = nullptr;
    const Dims* d = nullptr;
    char* conv_scratch = nullptr;
    int64_t launches = 0;
    cudaError_t err = cudaSuccess;

    template <class F>
    void for_each(size_t n, const F& f) {
        if (n == 0) return;
        k_for_each<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f, n);
        ++launches;
    }
    void for_each(size_t n, const GnApply& f) {             // overloads: four channels per thread
        if (n == 0) return;
        if (f.C % 4 || (f.C / f.G) % 4) { k_for_each<GnApply><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f, n); ++launches; return; }
        k_gn_apply4<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(f, n / 4);
        ++launches;
    }
    void for_each(size_t n, const MsmMerge& m) {
        if (n == 0) return;
        if (m.C % 4) { k_for_each<MsmMerge><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, n); ++launches; return; }
        k_msm_merge4<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(m, n / 4);
        ++launches;
    }
    void for_each(size_t n, const GnFinalize& f) {          // overload: one warp per (image, group)
        if (n == 0) return;
        k_gn_finalize_warp<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(f, n);
        ++launches;
    }
    // 3x3 convolution of level l on the tensor cores; false = not available for this call (fp32 path takes over)
    bool conv3_tc(const float* A, const float* Wt, float* out) {
        if (!tc || !tc->on || !conv_scratch) return false;
        int l = -1;
        for (int i = 0; i < d->L; ++i) if (w->fpn_t[i] == Wt) l = i;
        if (l < 0) return false;
        const int H = d->H[l], W = d->W[l];
        const size_t Mp = (size_t)d->B * (H + 2) * (W + 2);
        if (Mp >= (1ull << 31) / 2) return false;
        __half* hi = reinterpret_cast<__half*>(conv_scratch);
        __half* lo = hi + Mp * kC;
        float* out_pad = reinterpret_cast<float*>(lo + Mp * kC);
        CUtensorMap a_hi, a_lo;
        if (!ddp::tc::make_map_f16(&a_hi, hi, Mp, kC, ddp::tc::BM) || !ddp::tc::make_map_f16(&a_lo, lo, Mp, kC, ddp::tc::BM)) return false;
        for_each(Mp * 32, PadSplit{A, hi, lo, H, W});
        ddp::tc::EpiParams ep{};
        ep.scale = tc->inv_scale[l]; ep.bias = nullptr; ep.out = out_pad; ep.ldc = kC; ep.ncols = kC;
        ep.conv_pitch = W + 2;
        cudaError_t e = ddp::tc::launch_gemm_tc<256, 3, ddp::tc::EPI_BIAS>(a_hi, a_lo, a_hi, a_lo, tc->w_hi[l], tc->w_lo[l], (int)Mp,
                                                                         9 * kC, 9 * kC, kC, ep, tc->num_sms, st);
        if (e != cudaSuccess) { err = e; return false; }
        ++launches;
        for_each((size_t)d->B * H * W * 64, Unpad{out_pad, out, H, W});
        return true;
    }
    void gemm(int mode, const float* A, int lda, int n_img, const float* Wt, long long M, int K, float* out) {
        if (mode == A_CONV3 && conv3_tc(A, Wt, out)) return;
        EpiBias epi{out, nullptr, kC, kC, (int)M};
        if (mode == A_ROW_MAJOR) launch_gemm_simt<256, A_ROW_MAJOR>(A, lda, n_img, Wt, kC, (int)M, K, kC, epi, st);
        else if (mode == A_NCHW) launch_gemm_simt<256, A_NCHW>(A, lda, n_img, Wt, kC, (int)M, K, kC, epi, st);
        else launch_gemm_simt<256, A_CONV3>(A, lda, n_img, Wt, kC, (int)M, K, kC, epi, st);
        ++launches;
    }
    // src [B][N][C] token-major -> dst [B][C][N]: the tiled transpose of kernels.cuh with the roles of C and N swapped
    void tokens_to_nchw(const float* src, float* dst, int B, int N, int C) {
        dim3 grid((C + 31) / 32, (N + 31) / 32, B);
        k_nchw_to_tokens<<<grid, dim3(32, 8), 0, st>>>(src, dst, N, C);
        ++launches;
    }
};

}  // namespace neck
}  // namespace ddp

struct ddp_neck {
    ddp_neck_config cfg;
    int device = 0;
    std::string err;
    struct Spec { std::string name; int64_t numel; std::vector<float> host; bool set = false; };
    std::vector<Spec> specs;
    bool committed = false, planned = false;
    float* w_arena = nullptr;
    ddp::neck::Weights w{};
    ddp::neck::Dims dims{};
    ddp::neck::NeckTc tc;        // tensor-core 3x3 convolutions (DDP_B200_NECK_TC, default on)
    size_t ws_bytes = 0;
    int64_t launches = 0;
};

namespace {

std::string g_neck_create_err;

int nfail(ddp_neck* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_neck_create_err = buf;
    return code;
}

#define NECK_CUDA_TRY(h, expr)                                                                    \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return nfail(h, DDP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                         __FILE__, __LINE__);                                                     \
    } while (0)

// "neck.0.lateral_convs.1.gn.weight", "0.lateral_convs.1.gn.weight" and "lateral_convs.1.gn.weight" name the same
// tensor: the key relative to its module is what identifies it (the three stems are unique across FPN and MSM).
const char* neck_key_stem(const char* name) {
    for (const char* stem : {"lateral_convs.", "fpn_convs.", "down."}) {
        const char* p = strstr(name, stem);
        if (p) return p;
    }
    return name;
}

ddp_neck::Spec* neck_find(ddp_neck* h, const char* name) {
    const std::string key = neck_key_stem(name);
    for (auto& s : h->specs)
        if (s.name == key) return &s;
    return nullptr;
}

const std::vector<float>& neck_host(ddp_neck* h, const std::string& key) { return neck_find(h, key.c_str())->host; }

}  // namespace

extern "C" {

const char* ddp_neck_last_error(const ddp_neck* h) { return h ? h->err.c_str() : g_neck_create_err.c_str(); }

int ddp_neck_create(const ddp_neck_config* cfg, ddp_neck** out) {
    using namespace ddp::neck;
    if (!cfg || !out) return nfail(nullptr, DDP_ERR_INVALID, "ddp_neck_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != DDP_ABI_VERSION)
        return nfail(nullptr, DDP_ERR_INVALID, "ddp_neck_create: abi_version %d != %d", cfg->abi_version, DDP_ABI_VERSION);
    if (cfg->stages < 1 || cfg->stages > 3) return nfail(nullptr, DDP_ERR_INVALID, "ddp_neck_create: stages must be 1, 2 or 3");
    if (cfg->num_levels < 1 || cfg->num_levels > kMaxLevels)
        return nfail(nullptr, DDP_ERR_INVALID, "ddp_neck_create: num_levels %d not in [1, %d]", cfg->num_levels, kMaxLevels);
    if (cfg->out_channels != kC)
        return nfail(nullptr, DDP_ERR_UNSUPPORTED, "ddp_neck_create: out_channels %d (every DDP config uses 256)", cfg->out_channels);
    if (cfg->num_groups < 1 || kC % cfg->num_groups)
        return nfail(nullptr, DDP_ERR_INVALID, "ddp_neck_create: num_groups %d does not divide 256", cfg->num_groups);
    for (int l = 0; l < cfg->num_levels; ++l) {
        const int c = (cfg->stages & STAGE_FPN) ? cfg->in_channels[l] : kC;
        if (c < 16 || c % 16)
            return nfail(nullptr, DDP_ERR_UNSUPPORTED, "ddp_neck_create: in_channels[%d] = %d must be a positive multiple of 16", l, c);
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
        return nfail(nullptr, DDP_ERR_CUDA, "ddp_neck_create: no CUDA device (this library has no CPU path)");
    if (major != 10)
        return nfail(nullptr, DDP_ERR_CUDA, "ddp_neck_create: device compute capability %d.x, this build is sm_100a only", major);
    ddp_neck* h = new ddp_neck();
    h->cfg = *cfg;
    h->device = dev;
    {
        const char* e = getenv("DDP_B200_NECK_TC");                    // 0 = fp32 CUDA-core 3x3 convolutions (the first path)
        h->tc.on = (e == nullptr || atoi(e) != 0) && ddp::tc::get_encode_fn() != nullptr;
        int sms = 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) h->tc.num_sms = sms;
    }
    const int L = cfg->num_levels;
    auto add = [&](const std::string& name, int64_t numel) {
        ddp_neck::Spec s;
        s.name = name;
        s.numel = numel;
        h->specs.push_back(std::move(s));
    };
    if (cfg->stages & STAGE_FPN) {
        for (int l = 0; l < L; ++l) {
            const std::string p = "lateral_convs." + std::to_string(l) + ".";
            add(p + "conv.weight", (int64_t)kC * cfg->in_channels[l]);
            add(p + "gn.weight", kC);
            add(p + "gn.bias", kC);
        }
        for (int l = 0; l < L; ++l) {
            const std::string p = "fpn_convs." + std::to_string(l) + ".";
            add(p + "conv.weight", (int64_t)kC * kC * 9);
            add(p + "gn.weight", kC);
            add(p + "gn.bias", kC);
        }
    }
    if (cfg->stages & STAGE_MERGE) {
        add("down.conv.weight", (int64_t)kC * kC * L);
        add("down.gn.weight", kC);
        add("down.gn.bias", kC);
    }
    *out = h;
    return DDP_OK;
}

void ddp_neck_destroy(ddp_neck* h) {
    if (h && h->tc.w_arena) cudaFree(h->tc.w_arena);
    if (!h) return;
    if (h->w_arena) cudaFree(h->w_arena);
    delete h;
}

int ddp_neck_weight_count(const ddp_neck* h) { return h ? (int)h->specs.size() : 0; }

const char* ddp_neck_weight_name(const ddp_neck* h, int index, int64_t* numel) {
    if (!h || index < 0 || index >= (int)h->specs.size()) return nullptr;
    if (numel) *numel = h->specs[index].numel;
    return h->specs[index].name.c_str();
}

int ddp_neck_set_weight(ddp_neck* h, const char* name, const float* host_data, int64_t numel) {
    if (!h) return DDP_ERR_INVALID;
    if (!name || !host_data) return nfail(h, DDP_ERR_INVALID, "ddp_neck_set_weight: null argument");
    ddp_neck::Spec* s = neck_find(h, name);
    if (!s) return nfail(h, DDP_ERR_WEIGHT, "ddp_neck_set_weight: '%s' is not a weight of this neck", name);
    if (numel != s->numel)
        return nfail(h, DDP_ERR_WEIGHT, "ddp_neck_set_weight: '%s' has %lld elements, expected %lld", name, (long long)numel,
                     (long long)s->numel);
    s->host.assign(host_data, host_data + numel);
    s->set = true;
    h->committed = false;
    return DDP_OK;
}

int ddp_neck_commit_weights(ddp_neck* h) {
    using namespace ddp::neck;
    if (!h) return DDP_ERR_INVALID;
    for (auto& s : h->specs)
        if (!s.set) return nfail(h, DDP_ERR_STATE, "ddp_neck_commit_weights: weight '%s' was never set", s.name.c_str());
    const int L = h->cfg.num_levels;
    // repack on the host (neck_plan.h), one upload
    std::vector<float> arena;
    std::vector<size_t> offs;
    auto put = [&](const std::vector<float>& v) {
        offs.push_back(arena.size());
        arena.insert(arena.end(), v.begin(), v.end());
        arena.resize((arena.size() + 63) / 64 * 64);           // 256-byte aligned pieces
    };
    if (h->cfg.stages & STAGE_FPN) {
        for (int l = 0; l < L; ++l) {
            const std::string p = "lateral_convs." + std::to_string(l) + ".", q = "fpn_convs." + std::to_string(l) + ".";
            put(repack_1x1(neck_host(h, p + "conv.weight").data(), kC, h->cfg.in_channels[l], 0, h->cfg.in_channels[l]));
            put(neck_host(h, p + "gn.weight"));
            put(neck_host(h, p + "gn.bias"));
            put(repack_3x3(neck_host(h, q + "conv.weight").data(), kC, kC));
            put(neck_host(h, q + "gn.weight"));
            put(neck_host(h, q + "gn.bias"));
        }
    }
    if (h->cfg.stages & STAGE_MERGE) {
        for (int l = 0; l < L; ++l) put(repack_1x1(neck_host(h, "down.conv.weight").data(), kC, kC * L, kC * l, kC));
        put(neck_host(h, "down.gn.weight"));
        put(neck_host(h, "down.gn.bias"));
    }
    NECK_CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->w_arena) { cudaFree(h->w_arena); h->w_arena = nullptr; }
    NECK_CUDA_TRY(h, cudaMalloc(&h->w_arena, arena.size() * sizeof(float)));
    NECK_CUDA_TRY(h, cudaMemcpy(h->w_arena, arena.data(), arena.size() * sizeof(float), cudaMemcpyHostToDevice));
    size_t i = 0;
    Weights w{};
    if (h->cfg.stages & STAGE_FPN) {
        for (int l = 0; l < L; ++l) {
            w.lat_t[l] = h->w_arena + offs[i++]; w.lat_g[l] = h->w_arena + offs[i++]; w.lat_b[l] = h->w_arena + offs[i++];
            w.fpn_t[l] = h->w_arena + offs[i++]; w.fpn_g[l] = h->w_arena + offs[i++]; w.fpn_b[l] = h->w_arena + offs[i++];
        }
    }
    if (h->cfg.stages & STAGE_MERGE) {
        for (int l = 0; l < L; ++l) w.down_t[l] = h->w_arena + offs[i++];
        w.down_g = h->w_arena + offs[i++];
        w.down_b = h->w_arena + offs[i++];
    }
    h->w = w;
    if (h->tc.on && (h->cfg.stages & STAGE_FPN)) {
        // 3x3 weights as fp16 hi/lo planes [256 out][k = tap * 256 + ci] of 2^shift * W (shift per level: the largest
        // |w| lands in [128, 256), so both planes are normal fp16 numbers), split on the host, one upload
        const size_t per = (size_t)kC * 9 * kC;
        std::vector<__half> planes(2 * per * L);
        for (int l = 0; l < L; ++l) {
            const std::vector<float>& src = neck_host(h, "fpn_convs." + std::to_string(l) + ".conv.weight");   // (256,256,3,3)
            float mx = 0.f;
            for (float v : src) mx = fmaxf(mx, fabsf(v));
            int e = 0, shift = 0;
            if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); shift = 8 - e; shift = shift > 15 ? 15 : (shift < -8 ? -8 : shift); }
            const float scale = ldexpf(1.0f, shift);
            h->tc.inv_scale[l] = 1.0f / (scale * ddp::tc::kActScale);
            __half* hi = planes.data() + (size_t)l * 2 * per;
            __half* lo = hi + per;
            for (int co = 0; co < kC; ++co)
                for (int ci = 0; ci < kC; ++ci)
                    for (int tap = 0; tap < 9; ++tap) {
                        const float sv = src[((size_t)co * kC + ci) * 9 + tap] * scale;
                        const __half hh = __float2half_rn(sv);
                        const size_t k = (size_t)co * 9 * kC + (size_t)tap * kC + ci;
                        hi[k] = hh;
                        lo[k] = __float2half_rn(sv - __half2float(hh));
                    }
        }
        if (h->tc.w_arena) { cudaFree(h->tc.w_arena); h->tc.w_arena = nullptr; }
        NECK_CUDA_TRY(h, cudaMalloc(&h->tc.w_arena, planes.size() * sizeof(__half)));
        NECK_CUDA_TRY(h, cudaMemcpy(h->tc.w_arena, planes.data(), planes.size() * sizeof(__half), cudaMemcpyHostToDevice));
        for (int l = 0; l < L; ++l) {
            __half* hi = h->tc.w_arena + (size_t)l * 2 * per;
            if (!ddp::tc::make_map_f16(&h->tc.w_hi[l], hi, kC, 9 * kC, 256) || !ddp::tc::make_map_f16(&h->tc.w_lo[l], hi + per, kC, 9 * kC, 256))
                return nfail(h, DDP_ERR_CUDA, "ddp_neck_commit_weights: cuTensorMapEncodeTiled failed for the 3x3 weight planes of level %d", l);
        }
    }
    h->committed = true;
    return DDP_OK;
}

int ddp_neck_plan(ddp_neck* h, int B, const int32_t* heights, const int32_t* widths, size_t* workspace_bytes) {
    using namespace ddp::neck;
    if (!h) return DDP_ERR_INVALID;
    if (!heights || !widths) return nfail(h, DDP_ERR_INVALID, "ddp_neck_plan: null argument");
    if (B < 1 || B > 65535) return nfail(h, DDP_ERR_INVALID, "ddp_neck_plan: B = %d must be in [1, 65535]", B);
    Dims d{};
    d.L = h->cfg.num_levels; d.B = B; d.stages = h->cfg.stages; d.groups = h->cfg.num_groups; d.eps = h->cfg.eps;
    for (int l = 0; l < d.L; ++l) {
        if (heights[l] < 1 || widths[l] < 1) return nfail(h, DDP_ERR_INVALID, "ddp_neck_plan: level %d is %d x %d", l, heights[l], widths[l]);
        d.C[l] = (d.stages & STAGE_FPN) ? h->cfg.in_channels[l] : kC;
        d.H[l] = heights[l];
        d.W[l] = widths[l];
        // rows of the GEMMs and element counts of the per-element kernels are 32-bit in the GEMM kernel / grid.x
        if ((long long)B * d.tokens(l) > (1LL << 24) || d.tokens(l) > 65535LL * 32)
            return nfail(h, DDP_ERR_UNSUPPORTED, "ddp_neck_plan: B * h * w = %lld tokens at level %d exceeds 2^24 (or h * w > 2^21)",
                         (long long)B * d.tokens(l), l);
    }
    h->dims = d;
    h->ws_bytes = carve(d, nullptr, nullptr);
    h->planned = true;
    if (workspace_bytes) *workspace_bytes = h->ws_bytes;
    return DDP_OK;
}

int ddp_neck_forward(ddp_neck* h, const float* const* inputs, float* x_out, float* const* fpn_outs, void* workspace,
                     size_t workspace_bytes, void* stream) {
    using namespace ddp::neck;
    if (!h) return DDP_ERR_INVALID;
    if (!h->committed) return nfail(h, DDP_ERR_STATE, "ddp_neck_forward: call ddp_neck_commit_weights first");
    if (!h->planned) return nfail(h, DDP_ERR_STATE, "ddp_neck_forward: call ddp_neck_plan first");
    if (!inputs || !workspace) return nfail(h, DDP_ERR_INVALID, "ddp_neck_forward: null pointer");
    const Dims& d = h->dims;
    for (int l = 0; l < d.L; ++l)
        if (!inputs[l]) return nfail(h, DDP_ERR_INVALID, "ddp_neck_forward: inputs[%d] is null", l);
    if ((d.stages & STAGE_MERGE) && !x_out) return nfail(h, DDP_ERR_INVALID, "ddp_neck_forward: x_out is null");
    if (d.stages == STAGE_FPN) {
        if (!fpn_outs) return nfail(h, DDP_ERR_INVALID, "ddp_neck_forward: an FPN-only neck needs fpn_outs");
        for (int l = 0; l < d.L; ++l)
            if (!fpn_outs[l]) return nfail(h, DDP_ERR_INVALID, "ddp_neck_forward: fpn_outs[%d] is null", l);
    }
    if (workspace_bytes < h->ws_bytes)
        return nfail(h, DDP_ERR_WORKSPACE, "ddp_neck_forward: workspace %zu < required %zu", workspace_bytes, h->ws_bytes);
    if (reinterpret_cast<uintptr_t>(workspace) % 256)
        return nfail(h, DDP_ERR_WORKSPACE, "ddp_neck_forward: workspace must be 256-byte aligned");
    Buffers buf;
    carve(d, static_cast<char*>(workspace), &buf);
    CudaBackend be{static_cast<cudaStream_t>(stream)};
    be.tc = &h->tc; be.w = &h->w; be.d = &h->dims; be.conv_scratch = buf.conv_scratch;
    neck_run(be, d, h->w, buf, inputs, x_out, fpn_outs);
    h->launches = be.launches;
    NECK_CUDA_TRY(h, be.err);
    NECK_CUDA_TRY(h, cudaGetLastError());
    return DDP_OK;
}

int64_t ddp_neck_last_launch_count(const ddp_neck* h) { return h ? h->launches : 0; }

}  // extern "C"
