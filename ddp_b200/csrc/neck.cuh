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

struct CudaBackend {
    cudaStream_t st;
    int64_t launches = 0;

    template <class F>
    void for_each(size_t n, const F& f) {
        if (n == 0) return;
        k_for_each<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f, n);
        ++launches;
    }
    void gemm(int mode, const float* A, int lda, int n_img, const float* Wt, long long M, int K, float* out) {
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
    neck_run(be, d, h->w, buf, inputs, x_out, fpn_outs);
    h->launches = be.launches;
    NECK_CUDA_TRY(h, cudaGetLastError());
    return DDP_OK;
}

int64_t ddp_neck_last_launch_count(const ddp_neck* h) { return h ? h->launches : 0; }

}  // extern "C"
