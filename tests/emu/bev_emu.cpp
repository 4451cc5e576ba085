// Host emulation of the BEV loop's launch sequence.  TEST INFRASTRUCTURE ONLY — compiled by tests/test_bev_emu_cpu.py
// with g++ into a temporary directory, never part of libddp_b200.so.
//
// Instantiates ddp::bev::bev_run (ddp_b200/csrc/bev_plan.h: the launch sequence and per-element kernel bodies of the
// CUDA build) with a sequential backend.  The denoiser (5 transformer layers + conv_seg: ddp_head_forward in the CUDA
// build) is NOT emulated: `denoise` replays the logits the oracle recorded for that step (teacher forcing) and dumps
// its input, so the test can compare what this file's code feeds the denoiser and what it does with the result.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "bev_plan.h"

using namespace ddp::bev;

namespace {
struct HostBackend {
    const Dims& d;
    const float* replay;      // [T][rows][6][n_out]
    float* feat_dump;         // [T][rows][256][n_out]
    float* state_dump;        // [T][rows][n_state][256]
    int step = 0;

    template <class F>
    void for_each(size_t n, const F& f) {
        for (size_t i = 0; i < n; ++i) f(i);
    }
    void for_each(size_t n, const StepUpdate& f) {
        for (size_t i = 0; i < n; ++i) f(i);
        memcpy(state_dump + (size_t)step * n, f.state, n * sizeof(float));
        ++step;
    }
    void gemm_cond(const float* x, int feat, int N, int B, const float* wx_t, const float* bias, float* cond) {
        for (long long row = 0; row < (long long)B * N; ++row) {
            const long long img = row / N, n = row - img * N;
            for (int co = 0; co < kEmbed; ++co) cond[row * kEmbed + co] = 0.f;
            for (int k = 0; k < feat; ++k) {
                const float a = x[(img * feat + k) * N + n];
                for (int co = 0; co < kEmbed; ++co) cond[row * kEmbed + co] += a * wx_t[(size_t)k * kEmbed + co];
            }
            for (int co = 0; co < kEmbed; ++co) cond[row * kEmbed + co] += bias[co];
        }
    }
    void gemm_head_in(const float* state, const float* wm_t, const float* cond, int N, int R, int rows, float* q) {
        std::vector<float> acc(kEmbed);
        for (long long row = 0; row < (long long)rows * N; ++row) {
            for (int co = 0; co < kEmbed; ++co) acc[co] = 0.f;
            for (int k = 0; k < kEmbed; ++k) {
                const float a = state[row * kEmbed + k];
                for (int co = 0; co < kEmbed; ++co) acc[co] += a * wm_t[(size_t)k * kEmbed + co];
            }
            const long long b = (row / N) / R, n = row % N;
            for (int co = 0; co < kEmbed; ++co) q[row * kEmbed + co] = acc[co] + cond[(b * N + n) * kEmbed + co];
        }
    }
    void nchw_to_tokens(const float* src, float* dst, int imgs, int C, int N) {
        for (int b = 0; b < imgs; ++b)
            for (int c = 0; c < C; ++c)
                for (int n = 0; n < N; ++n) dst[((size_t)b * N + n) * C + c] = src[((size_t)b * C + c) * N + n];
    }
    void tokens_to_nchw(const float* src, float* dst, int imgs, int N, int C) {
        for (int b = 0; b < imgs; ++b)
            for (int n = 0; n < N; ++n)
                for (int c = 0; c < C; ++c) dst[((size_t)b * C + c) * N + n] = src[((size_t)b * N + n) * C + c];
    }
    int denoise(int k, const float* feat_tokens, float* feat_nchw, float* logits) {
        tokens_to_nchw(feat_tokens, feat_nchw, d.rows(), (int)d.n_out(), kEmbed);     // decode_head.forward takes NCHW
        const size_t nf = (size_t)d.rows() * kEmbed * d.n_out(), nl = (size_t)d.rows() * kClasses * d.n_out();
        memcpy(feat_dump + (size_t)k * nf, feat_nchw, nf * sizeof(float));
        memcpy(logits, replay + (size_t)k * nl, nl * sizeof(float));
        return 0;
    }
};
}  // namespace

// transform_w: (256, feat + 256) as in the reference; sched: [4][T] = a_now, s_now, a_next, s_next.
extern "C" int bev_emu_run(int B, int R, int T, int feat, int h, int w, int Ho, int Wo, float bit_scale, float threshold,
                           const float* transform_w, const float* transform_b, const float* emb, const float* grid_y,
                           const float* grid_x, const float* sched, const float* x, const float* noise, const float* replay,
                           float* out, float* feat_dump, float* state_dump) {
    Dims d{B, R, T, feat, h, w, Ho, Wo, bit_scale, threshold};
    std::vector<float> wx = ddp::neck::repack_1x1(transform_w, kEmbed, feat + kEmbed, 0, feat);
    std::vector<float> wm = ddp::neck::repack_1x1(transform_w, kEmbed, feat + kEmbed, feat, kEmbed);
    Weights wt{wx.data(), wm.data(), transform_b, emb, grid_y, grid_x};
    Schedule sch{sched, sched + T, sched + 2 * T, sched + 3 * T};
    const size_t bytes = carve(d, nullptr, nullptr);
    char* ws = static_cast<char*>(malloc(bytes));
    if (!ws) return -1;
    Buffers buf;
    carve(d, ws, &buf);
    HostBackend be{d, replay, feat_dump, state_dump};
    const int rc = bev_run(be, d, wt, sch, buf, x, noise, out);
    free(ws);
    return rc;
}
