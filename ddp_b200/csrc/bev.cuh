// CUDA backend + C ABI of the BEV map-segmentation variant of the decode loop — SURVEY 8f #4.
// The launch sequence and the per-element kernel bodies live in bev_plan.h (shared with the host emulation the CPU
// tests use).  The denoiser is NOT re-implemented: a ddp_bev owns an inner segmentation ddp_handle (6 classes,
// num_layers = 5, planned on the OUTPUT grid) and calls the hardware-verified denoiser (head_forward_impl, the body of
// ddp_head_forward) once per step on the tokens its grid sample wrote, with the step's FiLM vectors taken from the inner
// handle's precomputed table.
//
// Included at the end of ddp_b200.cu (same translation unit: it reaches into ddp_handle).
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ddp_b200.h"
#include "bev_plan.h"
#include "gemm_simt.cuh"
#include "kernels.cuh"
#include "neck.cuh"          // k_for_each

struct ddp_bev {
    ddp_bev_config cfg;
    int device = 0;
    std::string err;
    ddp_handle* inner = nullptr;
    std::vector<float> tr_w, tr_b, emb;      // host copies of transform.conv.{weight,bias}, embedding_table.weight
    bool have_tr_w = false, have_tr_b = false, have_emb = false, committed = false, planned = false;
    std::vector<std::string> names;          // reference state-dict keys, index-aligned with `numels`
    std::vector<int64_t> numels;
    float* w_arena = nullptr;                // wx_t | wm_t | b_tr | emb
    float* g_arena = nullptr;                // grid_y | grid_x
    // tensor-core head-in (q = cond + W_m m_t on the state grid), when the inner handle runs a tc_* gemm mode
    float* tr_w_dev = nullptr;               // transform.conv.weight as the reference stores it, (256, feat + 256)
    __half* tc_arena = nullptr;              // W_m as scaled fp16 planes
    TcWeight tc_in;
    CUtensorMap mA_state[2];                 // m_t planes of the workspace the maps were last made for
    const void* maps_ws = nullptr;
    ddp::bev::Weights w{};
    ddp::bev::Dims dims{};
    size_t own_bytes = 0, ws_bytes = 0;
    int64_t launches = 0;
};

namespace {

std::string g_bev_create_err;

int bfail(ddp_bev* h, int code, const char* fmt, ...) {
    char buf[640];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_bev_create_err = buf;
    return code;
}

#define BEV_CUDA_TRY(h, expr)                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return bfail(h, DDP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                         __FILE__, __LINE__);                                                     \
    } while (0)

// reference key of the BEV model -> key of the inner segmentation handle ("" = not forwarded)
std::string bev_inner_key(const std::string& k) {
    const std::string hp = "heads.map.";
    if (k.compare(0, hp.size(), hp) == 0) return "decode_head." + k.substr(hp.size());
    if (k == "transform.conv.weight" || k == "transform.conv.bias") return "";
    return k;                                   // embedding_table.weight, time_mlp.*
}

// GridSample / StepUpdate with four channels per thread (float4 loads / stores; the geometry, the six sigmoids of the source
// pixel and the schedule arithmetic are computed once per thread instead of once per channel; per-element arithmetic
// identical to the functors of bev_plan.h, which the host emulation keeps using)
__global__ void __launch_bounds__(256) k_bev_grid_sample4(ddp::bev::GridSample f, size_t n4) {
    using namespace ddp::bev;
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const int c = (int)(i4 % (kEmbed / 4)) * 4;
    const size_t t = i4 / (kEmbed / 4);
    const size_t n_out = (size_t)f.Ho * f.Wo;
    const int n = (int)(t % n_out);
    const size_t row = t / n_out;
    const int Y = n / f.Wo, X = n - Y * f.Wo;
    const float x = (f.grid_x[X] + 1.0f) * ((float)f.w * 0.5f) - 0.5f;
    const float y = (f.grid_y[Y] + 1.0f) * ((float)f.h * 0.5f) - 0.5f;
    const float xf = floorf(x), yf = floorf(y);
    const float wx1 = x - xf, wy1 = y - yf, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
    const int x0 = (xf >= -2.0f && xf <= (float)f.w) ? (int)xf : -2;
    const int y0 = (yf >= -2.0f && yf <= (float)f.h) ? (int)yf : -2;
    const float* p = f.src + row * (size_t)f.h * f.w * kEmbed + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto add = [&](int yy, int xx, float wgt) {
        const float4 v = *reinterpret_cast<const float4*>(p + ((size_t)yy * f.w + xx) * kEmbed);
        acc.x += v.x * wgt; acc.y += v.y * wgt; acc.z += v.z * wgt; acc.w += v.w * wgt;
    };
    if (y0 >= 0 && y0 < f.h) {
        if (x0 >= 0 && x0 < f.w) add(y0, x0, wy0 * wx0);
        if (x0 + 1 >= 0 && x0 + 1 < f.w) add(y0, x0 + 1, wy0 * wx1);
    }
    if (y0 + 1 >= 0 && y0 + 1 < f.h) {
        if (x0 >= 0 && x0 < f.w) add(y0 + 1, x0, wy1 * wx0);
        if (x0 + 1 >= 0 && x0 + 1 < f.w) add(y0 + 1, x0 + 1, wy1 * wx1);
    }
    *reinterpret_cast<float4*>(f.dst + i4 * 4) = acc;
}
// hi / lo: optional fp16 planes of the new state (16 x, hi + lo as split8_store writes them) for the tensor-core head-in
__global__ void __launch_bounds__(256) k_bev_step_update4(ddp::bev::StepUpdate u, size_t n4, __half* hi, __half* lo) {
    using namespace ddp::bev;
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const int c = (int)(i4 % (kEmbed / 4)) * 4;
    const size_t t = i4 / (kEmbed / 4);
    const size_t n_state = (size_t)u.h * u.w, n_out = (size_t)u.Ho * u.Wo;
    const int n = (int)(t % n_state);
    const size_t row = t / n_state;
    const int i = n / u.w, j = n - i * u.w;
    const int Y = ddp::neck::nearest_src(i, u.sh, u.Ho), X = ddp::neck::nearest_src(j, u.sw, u.Wo);
    const float* lp = u.logits + row * kClasses * n_out + (size_t)Y * u.Wo + X;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < kClasses; ++s) {
        const bool on = sigmoid_f(lp[(size_t)s * n_out]) > u.threshold;
        const float4 e = *reinterpret_cast<const float4*>(u.emb + (size_t)(on ? s + 1 : 0) * kEmbed + c);
        sum.x += e.x; sum.y += e.y; sum.z += e.z; sum.w += e.w;
    }
    const float sg = u.sigma > 1e-8f ? u.sigma : 1e-8f;
    float4 m = *reinterpret_cast<const float4*>(u.state + i4 * 4);
    auto upd = [&](float sm, float mv) {
        const float mean = sm / (float)kClasses;
        const float pred = (sigmoid_f(mean) * 2.0f - 1.0f) * u.bit_scale;
        const float eps = (mv - u.alpha * pred) / sg;
        return pred * u.alpha_next + eps * u.sigma_next;
    };
    m.x = upd(sum.x, m.x); m.y = upd(sum.y, m.y); m.z = upd(sum.z, m.z); m.w = upd(sum.w, m.w);
    *reinterpret_cast<float4*>(u.state + i4 * 4) = m;
    if (hi) {
        const float v[4] = {m.x, m.y, m.z, m.w};
        __align__(8) __half hh[4];
        __align__(8) __half ll[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float sc = fminf(fmaxf(v[e] * ddp::kSplitScale, -65504.0f), 65504.0f);
            hh[e] = __float2half_rn(sc);
            ll[e] = __float2half_rn(sc - __half2float(hh[e]));
        }
        *reinterpret_cast<uint2*>(hi + i4 * 4) = *reinterpret_cast<const uint2*>(hh);
        if (lo) *reinterpret_cast<uint2*>(lo + i4 * 4) = *reinterpret_cast<const uint2*>(ll);
    }
}

// q = cond + W_m m_t on tcgen05 (the segmentation loop's head-in GEMM, EPI_ADD_COND, fp32 output for the grid sample)
static int bev_head_in_tc(ddp_bev* b, const CUtensorMap* maps, const float* cond, int N, int R, int M, float* q, cudaStream_t st) {
    ddp_handle* h = b->inner;
    tc::EpiParams ep{};
    ep.scale = b->tc_in.inv_scale; ep.out = q; ep.ldc = kE; ep.ncols = kE;
    ep.cond = cond; ep.N_tok = N; ep.R = R;
    TC_GEMM(h, DDP_K_HEAD_IN, st, 256, tc::EPI_ADD_COND, maps, b->tc_in, M, kE, ep);
    return DDP_OK;
}

struct BevCudaBackend {
    ddp_bev* h;
    void* inner_ws;
    cudaStream_t st;
    __half* st_hi = nullptr;     // fp16 planes of m_t (tensor-core head-in); null = fp32 head-in
    __half* st_lo = nullptr;
    bool planes_valid = false;   // the planes hold the current m_t (StepUpdate writes them; the first step splits the noise)
    int err_rc = 0;
    int64_t launches = 0;

    template <class F>
    void for_each(size_t n, const F& f) {
        if (n == 0) return;
        ddp::neck::k_for_each<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f, n);
        ++launches;
    }
    void for_each(size_t n, const ddp::bev::GridSample& f) {        // overloads: four channels per thread
        if (n == 0) return;
        k_bev_grid_sample4<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(f, n / 4);
        ++launches;
    }
    void for_each(size_t n, const ddp::bev::StepUpdate& u) {
        if (n == 0) return;
        k_bev_step_update4<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(u, n / 4, st_hi, st_lo);
        planes_valid = st_hi != nullptr;
        ++launches;
    }
    void gemm_cond(const float* x, int feat, int N, int B, const float* wx_t, const float* bias, float* cond) {
        ddp::EpiBias epi{cond, bias, ddp::kE, ddp::kE, B * N};
        ddp::launch_gemm_simt<256, 1>(x, 0, N, wx_t, ddp::kE, B * N, feat, ddp::kE, epi, st);
        ++launches;
    }
    void gemm_head_in(const float* state, const float* wm_t, const float* cond, int N, int R, int rows, float* q) {
        if (st_hi) {
            if (!planes_valid) {
                const size_t n8 = (size_t)rows * N * ddp::kE / 8;
                ddp::k_split_planes<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(state, st_hi, st_lo, n8);
                planes_valid = true;
                ++launches;
            }
            const int rc = bev_head_in_tc(h, h->mA_state, cond, N, R, rows * N, q, st);
            if (rc && !err_rc) err_rc = bfail(h, rc, "tensor-core head-in failed: %s", ddp_last_error(h->inner));
            ++launches;
            return;
        }
        ddp::EpiAddCond epi{q, cond, N, R, rows * N};
        ddp::launch_gemm_simt<256, 0>(state, ddp::kE, 0, wm_t, ddp::kE, rows * N, ddp::kE, ddp::kE, epi, st);
        ++launches;
    }
    void nchw_to_tokens(const float* src, float* dst, int imgs, int C, int N) {
        dim3 grid((N + 31) / 32, (C + 31) / 32, imgs);
        ddp::k_nchw_to_tokens<<<grid, dim3(32, 8), 0, st>>>(src, dst, C, N);
        ++launches;
    }
    int denoise(int k, const float* feat_tokens, float* /*scratch_nchw*/, float* logits) {
        ddp_handle* in = h->inner;
        if (err_rc) return err_rc;
        const int rc = head_forward_impl(in, feat_tokens, true, nullptr, k, logits, inner_ws, in->ws_compute_bytes, st);
        if (rc) return bfail(h, rc, "denoiser (ddp_head_forward) failed at step %d: %s", k, ddp_last_error(in));
        launches += in->launches;
        return 0;
    }
};

}  // namespace

extern "C" {

const char* ddp_bev_last_error(const ddp_bev* h) { return h ? h->err.c_str() : g_bev_create_err.c_str(); }

int ddp_bev_create(const ddp_bev_config* cfg, ddp_bev** out) {
    if (!cfg || !out) return bfail(nullptr, DDP_ERR_INVALID, "ddp_bev_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != DDP_ABI_VERSION)
        return bfail(nullptr, DDP_ERR_INVALID, "ddp_bev_create: abi_version %d != %d", cfg->abi_version, DDP_ABI_VERSION);
    if (cfg->feat_channels < 16 || cfg->feat_channels % 16)
        return bfail(nullptr, DDP_ERR_UNSUPPORTED, "ddp_bev_create: feat_channels %d must be a positive multiple of 16", cfg->feat_channels);
    if (cfg->noise_schedule != DDP_SCHEDULE_COSINE && cfg->noise_schedule != DDP_SCHEDULE_LINEAR)
        return bfail(nullptr, DDP_ERR_INVALID, "invalid noise schedule %d", cfg->noise_schedule);        // fusion_models/ddp.py:102 ValueError
    if (cfg->diffusion != DDP_DIFFUSION_DDIM)   // the reference's BEV ddpm_sample indexes a (r b) tensor with [0] and embeds floats: it cannot run
        return bfail(nullptr, DDP_ERR_UNSUPPORTED, "ddp_bev_create: only diffusion='ddim' is built for the BEV variant");
    ddp_config ic{};
    ic.abi_version = DDP_ABI_VERSION; ic.task = DDP_TASK_SEG; ic.num_classes = ddp::bev::kClasses;
    ic.timesteps = cfg->timesteps; ic.time_difference = cfg->time_difference; ic.noise_schedule = cfg->noise_schedule;
    ic.diffusion = DDP_DIFFUSION_DDIM; ic.accumulation = 0; ic.learned_sinusoidal_dim = cfg->learned_sinusoidal_dim;
    ic.num_layers = cfg->num_layers; ic.gemm_mode = cfg->gemm_mode; ic.sample_range_lo = 0.f; ic.bit_scale = cfg->bit_scale;
    ddp_handle* inner = nullptr;
    const int rc = ddp_create(&ic, &inner);
    if (rc) return bfail(nullptr, rc, "ddp_bev_create: %s", ddp_last_error(nullptr));
    ddp_bev* h = new ddp_bev();
    h->cfg = *cfg;
    h->device = inner->device;
    h->inner = inner;
    // reference keys: the inner handle's list with decode_head.* renamed and the transform resized to feat_channels
    for (const auto& s : inner->specs) {
        std::string k = s.name;
        int64_t n = s.numel;
        const std::string dh = "decode_head.";
        if (k.compare(0, dh.size(), dh) == 0) k = "heads.map." + k.substr(dh.size());
        if (k == "transform.conv.weight") n = (int64_t)ddp::kE * (cfg->feat_channels + ddp::kE);
        h->names.push_back(k);
        h->numels.push_back(n);
    }
    *out = h;
    return DDP_OK;
}

void ddp_bev_destroy(ddp_bev* h) {
    if (!h) return;
    if (h->inner) ddp_destroy(h->inner);
    if (h->w_arena) cudaFree(h->w_arena);
    if (h->g_arena) cudaFree(h->g_arena);
    if (h->tr_w_dev) cudaFree(h->tr_w_dev);
    if (h->tc_arena) cudaFree(h->tc_arena);
    delete h;
}

int ddp_bev_weight_count(const ddp_bev* h) { return h ? (int)h->names.size() : 0; }

const char* ddp_bev_weight_name(const ddp_bev* h, int index, int64_t* numel) {
    if (!h || index < 0 || index >= (int)h->names.size()) return nullptr;
    if (numel) *numel = h->numels[index];
    return h->names[index].c_str();
}

int ddp_bev_set_weight(ddp_bev* h, const char* name, const float* host_data, int64_t numel) {
    if (!h) return DDP_ERR_INVALID;
    if (!name || !host_data) return bfail(h, DDP_ERR_INVALID, "ddp_bev_set_weight: null argument");
    const std::string k = name;
    size_t i = 0;
    while (i < h->names.size() && h->names[i] != k) ++i;
    if (i == h->names.size()) return bfail(h, DDP_ERR_WEIGHT, "ddp_bev_set_weight: '%s' is not a weight of the BEV decode path", name);
    if (numel != h->numels[i])
        return bfail(h, DDP_ERR_WEIGHT, "ddp_bev_set_weight: '%s' has %lld elements, expected %lld", name, (long long)numel,
                     (long long)h->numels[i]);
    h->committed = false;
    if (k == "transform.conv.weight") { h->tr_w.assign(host_data, host_data + numel); h->have_tr_w = true; return DDP_OK; }
    if (k == "transform.conv.bias") { h->tr_b.assign(host_data, host_data + numel); h->have_tr_b = true; return DDP_OK; }
    if (k == "embedding_table.weight") { h->emb.assign(host_data, host_data + numel); h->have_emb = true; }
    const int rc = ddp_set_weight(h->inner, bev_inner_key(k).c_str(), host_data, numel);
    if (rc) return bfail(h, rc, "ddp_bev_set_weight: %s", ddp_last_error(h->inner));
    return DDP_OK;
}

int ddp_bev_commit_weights(ddp_bev* h) {
    using namespace ddp::bev;
    if (!h) return DDP_ERR_INVALID;
    if (!h->have_tr_w || !h->have_tr_b || !h->have_emb)
        return bfail(h, DDP_ERR_STATE, "ddp_bev_commit_weights: transform.conv.{weight,bias} / embedding_table.weight were never set");
    // the inner handle's own transform (256 x 512) is never used by ddp_head_forward: the BEV transform runs here
    {
        std::vector<float> zeros((size_t)ddp::kE * 2 * ddp::kE, 0.f);
        int rc = ddp_set_weight(h->inner, "transform.conv.weight", zeros.data(), (int64_t)zeros.size());
        if (!rc) rc = ddp_set_weight(h->inner, "transform.conv.bias", zeros.data(), ddp::kE);
        if (!rc) rc = ddp_commit_weights(h->inner);
        if (rc) return bfail(h, rc, "ddp_bev_commit_weights: %s", ddp_last_error(h->inner));
    }
    const int feat = h->cfg.feat_channels;
    std::vector<float> wx = ddp::neck::repack_1x1(h->tr_w.data(), ddp::kE, feat + ddp::kE, 0, feat);
    std::vector<float> wm = ddp::neck::repack_1x1(h->tr_w.data(), ddp::kE, feat + ddp::kE, feat, ddp::kE);
    std::vector<float> arena;
    std::vector<size_t> offs;
    for (const std::vector<float>* v : {&wx, &wm, &h->tr_b, &h->emb}) {
        offs.push_back(arena.size());
        arena.insert(arena.end(), v->begin(), v->end());
        arena.resize((arena.size() + 63) / 64 * 64);
    }
    BEV_CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->w_arena) { cudaFree(h->w_arena); h->w_arena = nullptr; }
    BEV_CUDA_TRY(h, cudaMalloc(&h->w_arena, arena.size() * sizeof(float)));
    BEV_CUDA_TRY(h, cudaMemcpy(h->w_arena, arena.data(), arena.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->w.wx_t = h->w_arena + offs[0];
    h->w.wm_t = h->w_arena + offs[1];
    h->w.b_tr = h->w_arena + offs[2];
    h->w.emb = h->w_arena + offs[3];
    const char* sw = getenv("DDP_B200_BEV_HEAD_IN_TC");          // 0: keep the fp32 head-in GEMM in a tc_* gemm mode (A/B switch)
    if (h->inner->tc && !(sw && atoi(sw) == 0)) {
        // W_m (columns feat .. feat + 255 of the transform) as scaled fp16 planes + TMA maps, like the inner handle's tc_in
        if (h->tr_w_dev) { cudaFree(h->tr_w_dev); h->tr_w_dev = nullptr; }
        if (h->tc_arena) { cudaFree(h->tc_arena); h->tc_arena = nullptr; }
        BEV_CUDA_TRY(h, cudaMalloc(&h->tr_w_dev, h->tr_w.size() * sizeof(float)));
        BEV_CUDA_TRY(h, cudaMemcpy(h->tr_w_dev, h->tr_w.data(), h->tr_w.size() * sizeof(float), cudaMemcpyHostToDevice));
        const size_t halves = 2 * (size_t)ddp::kE * ddp::kE;
        BEV_CUDA_TRY(h, cudaMalloc(&h->tc_arena, halves * sizeof(__half)));
        BEV_CUDA_TRY(h, cudaMemset(h->tc_arena, 0, halves * sizeof(__half)));
        __half* cur = h->tc_arena;
        const int rc = make_tc_weight(h->inner, h->tc_in, cur, ddp::kE, ddp::kE, 256,
                                      {{h->tr_w_dev, &h->tr_w, ddp::kE, feat + ddp::kE, 1, feat, 0}}, nullptr);
        if (rc) return bfail(h, rc, "ddp_bev_commit_weights: %s", ddp_last_error(h->inner));
        BEV_CUDA_TRY(h, cudaDeviceSynchronize());
    } else if (h->tc_arena) {
        cudaFree(h->tc_arena); h->tc_arena = nullptr;
    }
    h->maps_ws = nullptr;
    h->committed = true;
    h->planned = false;
    return DDP_OK;
}

int ddp_bev_set_schedule(ddp_bev* h, int timesteps, const float* time_in, const float* a_now, const float* s_now,
                         const float* a_next, const float* s_next) {
    if (!h) return DDP_ERR_INVALID;
    const int rc = ddp_set_schedule(h->inner, timesteps, time_in, a_now, s_now, a_next, s_next);
    if (rc) return bfail(h, rc, "ddp_bev_set_schedule: %s", ddp_last_error(h->inner));
    return DDP_OK;
}

int ddp_bev_plan(ddp_bev* h, int B, int R, int in_h, int in_w, int out_h, int out_w, const float* grid_y, const float* grid_x,
                 size_t* workspace_bytes) {
    using namespace ddp::bev;
    if (!h) return DDP_ERR_INVALID;
    if (!h->committed) return bfail(h, DDP_ERR_STATE, "ddp_bev_plan: call ddp_bev_commit_weights first");
    if (!grid_y || !grid_x) return bfail(h, DDP_ERR_INVALID, "ddp_bev_plan: null grid");
    if (B < 1 || R < 1 || in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1 || B * (long long)R > 65535)
        return bfail(h, DDP_ERR_INVALID, "ddp_bev_plan: B, R, grid sizes must be >= 1 (B * R <= 65535)");
    if ((long long)B * R * in_h * in_w > (1LL << 23) || (long long)B * R * out_h * out_w > (1LL << 23))
        return bfail(h, DDP_ERR_UNSUPPORTED, "ddp_bev_plan: more than 2^23 tokens in flight");   // 32-bit element counts x 256 channels
    size_t inner_bytes = 0;
    const int rc = ddp_plan(h->inner, B, R, out_h, out_w, &inner_bytes);       // the denoiser runs on the output grid
    if (rc) return bfail(h, rc, "ddp_bev_plan: %s", ddp_last_error(h->inner));
    BEV_CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->g_arena) { cudaFree(h->g_arena); h->g_arena = nullptr; }
    BEV_CUDA_TRY(h, cudaMalloc(&h->g_arena, (size_t)(out_h + out_w) * sizeof(float)));
    BEV_CUDA_TRY(h, cudaMemcpy(h->g_arena, grid_y, out_h * sizeof(float), cudaMemcpyHostToDevice));
    BEV_CUDA_TRY(h, cudaMemcpy(h->g_arena + out_h, grid_x, out_w * sizeof(float), cudaMemcpyHostToDevice));
    h->w.grid_y = h->g_arena;
    h->w.grid_x = h->g_arena + out_h;
    Dims d{B, R, h->cfg.timesteps, h->cfg.feat_channels, in_h, in_w, out_h, out_w, h->cfg.bit_scale, h->cfg.threshold};
    h->dims = d;
    h->own_bytes = carve(d, nullptr, nullptr);
    h->ws_bytes = h->own_bytes + (h->inner->ws_compute_bytes + 255) / 256 * 256;
    h->maps_ws = nullptr;
    h->planned = true;
    if (workspace_bytes) *workspace_bytes = h->ws_bytes;
    return DDP_OK;
}

int ddp_bev_sample(ddp_bev* h, const float* x, const float* noise, float* out, void* workspace, size_t workspace_bytes,
                   void* stream) {
    using namespace ddp::bev;
    if (!h) return DDP_ERR_INVALID;
    if (!h->planned) return bfail(h, DDP_ERR_STATE, "ddp_bev_sample: call ddp_bev_plan first");
    if (!x || !noise || !out || !workspace) return bfail(h, DDP_ERR_INVALID, "ddp_bev_sample: null pointer");
    if (workspace_bytes < h->ws_bytes)
        return bfail(h, DDP_ERR_WORKSPACE, "ddp_bev_sample: workspace %zu < required %zu", workspace_bytes, h->ws_bytes);
    if (reinterpret_cast<uintptr_t>(workspace) % 256)
        return bfail(h, DDP_ERR_WORKSPACE, "ddp_bev_sample: workspace must be 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ddp_handle* in = h->inner;
    if (in->time_dirty) {          // schedule changed after the plan: refresh the per-step time embeddings
        const int rc = compute_time_constants(in, st);
        if (rc) return bfail(h, rc, "ddp_bev_sample: %s", ddp_last_error(in));
    }
    Buffers buf;
    carve(h->dims, static_cast<char*>(workspace), &buf);
    BevCudaBackend be{h, static_cast<char*>(workspace) + h->own_bytes, st};
    if (h->tc_arena) {
        const size_t M = (size_t)h->dims.rows() * h->dims.n_state();
        be.st_hi = static_cast<__half*>(buf.state_planes);
        be.st_lo = in->nsplit == 3 ? be.st_hi + M * ddp::kE : nullptr;
        if (h->maps_ws != workspace) {
            if (!tc::make_map_f16(&h->mA_state[0], be.st_hi, M, ddp::kE, tc::BM) ||
                !tc::make_map_f16(&h->mA_state[1], be.st_hi + M * ddp::kE, M, ddp::kE, tc::BM))
                return bfail(h, DDP_ERR_CUDA, "ddp_bev_sample: cuTensorMapEncodeTiled failed for the state planes");
            h->maps_ws = workspace;
        }
    }
    Schedule sch{in->a_now.data(), in->s_now.data(), in->a_next.data(), in->s_next.data()};
    const int rc = bev_run(be, h->dims, h->w, sch, buf, x, noise, out);
    h->launches = be.launches;
    if (rc) return rc;
    if (be.err_rc) return be.err_rc;
    BEV_CUDA_TRY(h, cudaGetLastError());
    return DDP_OK;
}

int64_t ddp_bev_last_launch_count(const ddp_bev* h) { return h ? h->launches : 0; }

}  // extern "C"
