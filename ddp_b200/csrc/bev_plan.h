// BEV map-segmentation variant of the decode loop as a launch sequence around the SAME denoiser.
//
// Reference: bev/mmdet3d/models/fusion_models/ddp.py:268-301 (DDP.ddim_sample) and
// bev/mmdet3d/models/heads/segm/deformable_head_with_time.py:58-98 (BEVGridTransform), :178-241 (head forward).
// Differences from the segmentation loop, all outside the transformer layers:
//   * two token grids: the diffusion state lives on the fused-BEV feature grid (h x w, 128 x 128 in the shipped
//     configs), the denoiser runs on the output map grid (H' x W', 200 x 200) after a bilinear grid_sample of its input;
//   * x has feat_channels (256 camera-only, 512 fusion) channels;
//   * the head ends in sigmoid over 6 independent classes; the loop thresholds at 0.5, nearest-resizes the multi-hot
//     map back to the state grid, embeds every class slot (index s + 1 if on, 0 if off) and takes the MEAN of the six
//     embeddings before the squash and the DDIM update;
//   * the result is the mean of the sigmoid maps of ALL T * R (step, sample) pairs.
//
// Like neck_plan.h this header holds the per-element kernel bodies as functors and the launch sequence `bev_run`,
// templated over a backend.  The product instantiates it with the CUDA backend of bev.cuh, where the denoiser is the
// hardware-verified ddp_head_forward of an inner segmentation handle (6 classes, 5 layers) planned on the output
// grid.  tests/emu/bev_emu.cpp instantiates the same sequence with a sequential host backend whose "denoiser" replays
// logits recorded from the oracle (teacher forcing), which checks everything this file adds.  Test infrastructure only:
// the library has no CPU path.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "neck_plan.h"      // DDP_HD, nearest_src

namespace ddp {
namespace bev {

constexpr int kEmbed = 256;          // tmp_channels = embed dims (fusion_models/ddp.py:78,104)
constexpr int kClasses = 6;      // self.num_classes = 6 (fusion_models/ddp.py:89)

struct Dims {
    int B, R, T;
    int feat;                    // channels of x
    int h, w;                    // state grid
    int Ho, Wo;                  // output (denoiser) grid
    float bit_scale, threshold;
    long long n_state() const { return (long long)h * w; }
    long long n_out() const { return (long long)Ho * Wo; }
    int rows() const { return B * R; }
};

struct Schedule { const float *a_now, *s_now, *a_next, *s_next; };      // host arrays [T] (ddp.py:278-285)

struct Weights {                 // device (or, in the emulation, host) pointers
    const float* wx_t;           // [feat][256]  transform.conv.weight[:, :feat]^T
    const float* wm_t;           // [256][256]   transform.conv.weight[:, feat:]^T
    const float* b_tr;           // [256]
    const float* emb;            // [7][256] embedding_table.weight
    const float* grid_y;         // [Ho] normalised sampling coordinates of the output rows (BEVGridTransform coords[0])
    const float* grid_x;         // [Wo]
};

struct Buffers {
    float* cond;                 // [B][n_state][256]
    float* state;                // [rows][n_state][256] m_t, token-major
    void* state_planes;          // same bytes as `state`: room for m_t as two fp16 planes (the CUDA backend's tensor-core head-in)
    float* q_state;              // [rows][n_state][256] transform(cat[x, m_t]) on the state grid
    float* q_out;                // [rows][n_out][256]   ... resampled onto the output grid, token-major
    float* feat_nchw;            // [rows][256][n_out]   scratch for a denoiser that wants decode_head.forward's NCHW input
    float* logits;               // [rows][6][n_out]     conv_seg output before the sigmoid, NCHW
    float* accum;                // [B][6][n_out]        running sum of sigmoid maps
};

inline size_t carve(const Dims& d, char* base, Buffers* out) {
    size_t off = 0;
    auto take = [&](size_t nfloats) {
        float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
        off += (nfloats * sizeof(float) + 255) / 256 * 256;
        return p;
    };
    Buffers b{};
    b.cond = take((size_t)d.B * d.n_state() * kEmbed);
    b.state = take((size_t)d.rows() * d.n_state() * kEmbed);
    b.state_planes = take((size_t)d.rows() * d.n_state() * kEmbed);
    b.q_state = take((size_t)d.rows() * d.n_state() * kEmbed);
    b.q_out = take((size_t)d.rows() * d.n_out() * kEmbed);
    b.feat_nchw = take((size_t)d.rows() * d.n_out() * kEmbed);
    b.logits = take((size_t)d.rows() * d.n_out() * kClasses);
    b.accum = take((size_t)d.B * d.n_out() * kClasses);
    if (out) *out = b;
    return off;
}

// F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False) of a token-major map at the separable grid
// of BEVGridTransform: output token (Y, X) samples the input at pixel ((gx + 1) * w / 2 - .5, (gy + 1) * h / 2 - .5)
// (ATen GridSamplerKernel ComputeLocation / ApplyGridSample<bilinear, zeros>).
struct GridSample {              // idx = (row * n_out + Y * Wo + X) * 256 + c
    const float* src; float* dst; const float* grid_y; const float* grid_x; int h, w, Ho, Wo;
    DDP_HD void operator()(size_t idx) const {
        const int c = (int)(idx % kEmbed);
        const size_t t = idx / kEmbed;
        const size_t n_out = (size_t)Ho * Wo;
        const int n = (int)(t % n_out);
        const size_t row = t / n_out;
        const int Y = n / Wo, X = n - Y * Wo;
        const float x = (grid_x[X] + 1.0f) * ((float)w * 0.5f) - 0.5f;
        const float y = (grid_y[Y] + 1.0f) * ((float)h * 0.5f) - 0.5f;
        const float xf = floorf(x), yf = floorf(y);
        const float wx1 = x - xf, wy1 = y - yf, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
        // coordinates far outside (or NaN) contribute nothing
        const int x0 = (xf >= -2.0f && xf <= (float)w) ? (int)xf : -2;
        const int y0 = (yf >= -2.0f && yf <= (float)h) ? (int)yf : -2;
        const float* p = src + row * (size_t)h * w * kEmbed + c;
        float acc = 0.f;
        if (y0 >= 0 && y0 < h) {
            if (x0 >= 0 && x0 < w) acc += p[((size_t)y0 * w + x0) * kEmbed] * (wy0 * wx0);
            if (x0 + 1 >= 0 && x0 + 1 < w) acc += p[((size_t)y0 * w + x0 + 1) * kEmbed] * (wy0 * wx1);
        }
        if (y0 + 1 >= 0 && y0 + 1 < h) {
            if (x0 >= 0 && x0 < w) acc += p[((size_t)(y0 + 1) * w + x0) * kEmbed] * (wy1 * wx0);
            if (x0 + 1 >= 0 && x0 + 1 < w) acc += p[((size_t)(y0 + 1) * w + x0 + 1) * kEmbed] * (wy1 * wx1);
        }
        dst[idx] = acc;
    }
};

DDP_HD float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// accum[b][s][n] += sum_r sigmoid(logits[b * R + r][s][n]), samples in fixed order (ddp.py:297-300: cat + mean).
struct Accumulate {              // idx = (b * 6 + s) * n_out + n
    const float* logits; float* accum; int R; size_t n_out;
    DDP_HD void operator()(size_t idx) const {
        const size_t n = idx % n_out;
        const size_t t = idx / n_out;
        const int s = (int)(t % kClasses);
        const size_t b = t / kClasses;
        float acc = accum[idx];
        for (int r = 0; r < R; ++r) acc += sigmoid_f(logits[((b * R + r) * kClasses + s) * n_out + n]);
        accum[idx] = acc;
    }
};

// Threshold -> nearest resize to the state grid -> mean of the six class-slot embeddings -> squash -> DDIM update
// (ddp.py:289-296).  One thread per state element; the six logits of its source pixel are shared by the 256 channels.
struct StepUpdate {              // idx = (row * n_state + i * w + j) * 256 + c
    const float* logits; const float* emb; float* state;
    int h, w, Ho, Wo; float sh, sw;          // sh = Ho / h, sw = Wo / w: F.interpolate(mode='nearest') scales
    float threshold, bit_scale, alpha, sigma, alpha_next, sigma_next;
    DDP_HD void operator()(size_t idx) const {
        const int c = (int)(idx % kEmbed);
        const size_t t = idx / kEmbed;
        const size_t n_state = (size_t)h * w, n_out = (size_t)Ho * Wo;
        const int n = (int)(t % n_state);
        const size_t row = t / n_state;
        const int i = n / w, j = n - i * w;
        const int Y = neck::nearest_src(i, sh, Ho), X = neck::nearest_src(j, sw, Wo);
        const float* lp = logits + row * kClasses * n_out + (size_t)Y * Wo + X;
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < kClasses; ++s) {
            const bool on = sigmoid_f(lp[(size_t)s * n_out]) > threshold;
            sum += emb[(size_t)(on ? s + 1 : 0) * kEmbed + c];
        }
        const float mean = sum / (float)kClasses;
        const float pred = (sigmoid_f(mean) * 2.0f - 1.0f) * bit_scale;
        const float m = state[idx];
        const float sg = sigma > 1e-8f ? sigma : 1e-8f;
        const float eps = (m - alpha * pred) / sg;
        state[idx] = pred * alpha_next + eps * sigma_next;
    }
};

struct Finalize {                // out = accum / (T * R)
    const float* accum; float* out; float count;
    DDP_HD void operator()(size_t idx) const { out[idx] = accum[idx] / count; }
};

struct Zero {
    float* p;
    DDP_HD void operator()(size_t idx) const { p[idx] = 0.f; }
};

// The launch sequence.  x: NCHW (B, feat, h, w); noise: NCHW (B, R, 256, h, w); out: NCHW (B, 6, Ho, Wo).
// Backend: for_each(n, functor); gemm_cond(x, feat, n_state, B, wx_t, bias, cond); gemm_head_in(state, wm_t, cond,
// n_state, R, rows, q); nchw_to_tokens(src, dst, imgs, C, N);
// denoise(step, feat_tokens, scratch_nchw, logits_nchw) -> 0 or an error code: the denoiser's input arrives as tokens
// ((rows, n_out, 256), what GridSample writes); a backend whose denoiser wants decode_head.forward's NCHW layout transposes
// into scratch_nchw itself (the host emulation does, the CUDA denoiser consumes tokens directly).
template <class Backend>
int bev_run(Backend& be, const Dims& d, const Weights& w, const Schedule& sch, const Buffers& buf, const float* x,
            const float* noise, float* out) {
    const size_t ns = (size_t)d.n_state(), no = (size_t)d.n_out();
    const int rows = d.rows();
    be.gemm_cond(x, d.feat, (int)ns, d.B, w.wx_t, w.b_tr, buf.cond);                       // step-invariant half of transform
    be.nchw_to_tokens(noise, buf.state, rows, kEmbed, (int)ns);                                 // mask_t = randn (ddp.py:275)
    be.for_each((size_t)d.B * kClasses * no, Zero{buf.accum});
    for (int k = 0; k < d.T; ++k) {
        be.gemm_head_in(buf.state, w.wm_t, buf.cond, (int)ns, d.R, rows, buf.q_state);      // transform(cat[x, mask_t])
        be.for_each((size_t)rows * no * kEmbed, GridSample{buf.q_state, buf.q_out, w.grid_y, w.grid_x, d.h, d.w, d.Ho, d.Wo});
        const int rc = be.denoise(k, buf.q_out, buf.feat_nchw, buf.logits);                 // 5 x (MSDA, LN, FFN, LN, FiLM) + conv_seg
        if (rc) return rc;
        be.for_each((size_t)d.B * kClasses * no, Accumulate{buf.logits, buf.accum, d.R, no});
        StepUpdate u{buf.logits, w.emb, buf.state, d.h, d.w, d.Ho, d.Wo, (float)d.Ho / (float)d.h, (float)d.Wo / (float)d.w,
                     d.threshold, d.bit_scale, sch.a_now[k], sch.s_now[k], sch.a_next[k], sch.s_next[k]};
        be.for_each((size_t)rows * ns * kEmbed, u);
    }
    be.for_each((size_t)d.B * kClasses * no, Finalize{buf.accum, out, (float)(d.T * d.R)});
    return 0;
}

}  // namespace bev
}  // namespace ddp
