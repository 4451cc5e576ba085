// Host emulation of the neck's launch sequence.  TEST INFRASTRUCTURE ONLY — compiled by tests/test_neck_emu_cpu.py
// with g++ into a temporary directory, never part of libddp_b200.so.
//
// It instantiates ddp::neck::neck_run (ddp_b200/csrc/neck_plan.h: the very launch sequence, per-element kernel
// bodies, weight repack and workspace carve-up the CUDA build uses) with a sequential backend: for_each is a loop,
// gemm is a naive triple loop over the same A-operand addressing (row-major / NCHW / 3x3 taps).  What it checks
// in a GPU-less container: indexing, interpolation, GroupNorm statistics and buffer wiring against the oracle.
// What it cannot check: the CUDA GEMM kernel's tile loaders and the transposing copy (GPU tests do).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "neck_plan.h"

using namespace ddp::neck;

namespace {
struct HostBackend {
    template <class F>
    void for_each(size_t n, const F& f) {
        for (size_t i = 0; i < n; ++i) f(i);
    }
    void gemm(int mode, const float* A, int lda, int n_img, const float* Wt, long long M, int K, float* out) {
        std::vector<float> acc(kC);
        for (long long row = 0; row < M; ++row) {
            for (int co = 0; co < kC; ++co) acc[co] = 0.f;
            for (int k = 0; k < K; ++k) {
                float a;
                if (mode == A_ROW_MAJOR) {
                    a = A[row * lda + k];
                } else if (mode == A_NCHW) {
                    const long long img = row / n_img, n = row - img * n_img;
                    a = A[(img * K + k) * n_img + n];
                } else {
                    const long long off = conv3_src_offset(row, k, n_img, lda, K / 9);
                    a = off < 0 ? 0.f : A[off];
                }
                const float* wr = Wt + (size_t)k * kC;
                for (int co = 0; co < kC; ++co) acc[co] += a * wr[co];
            }
            memcpy(out + row * kC, acc.data(), kC * sizeof(float));
        }
    }
    void tokens_to_nchw(const float* src, float* dst, int B, int N, int C) {
        for (int b = 0; b < B; ++b)
            for (int n = 0; n < N; ++n)
                for (int c = 0; c < C; ++c) dst[((size_t)b * C + c) * N + n] = src[((size_t)b * N + n) * C + c];
    }
};
}  // namespace

// weights in the reference's own layouts: lat_w[l] (256, C_l), fpn_w[l] (256, 256, 3, 3), down_w (256, 256 L);
// gn arrays: [lat_g0, lat_b0, ..., fpn_g0, fpn_b0, ..., down_g, down_b].
extern "C" int neck_emu_forward(int stages, int L, int B, const int* C, const int* H, const int* W, int groups, float eps,
                                const float* const* lat_w, const float* const* fpn_w, const float* down_w,
                                const float* const* gn, const float* const* inputs, float* x_out, float* const* fpn_outs) {
    Dims d{};
    d.L = L; d.B = B; d.stages = stages; d.groups = groups; d.eps = eps;
    for (int l = 0; l < L; ++l) { d.C[l] = C[l]; d.H[l] = H[l]; d.W[l] = W[l]; }
    std::vector<std::vector<float>> keep;
    Weights w{};
    for (int l = 0; l < L; ++l) {
        if (stages & STAGE_FPN) {
            keep.push_back(repack_1x1(lat_w[l], kC, C[l], 0, C[l])); w.lat_t[l] = keep.back().data();
            keep.push_back(repack_3x3(fpn_w[l], kC, kC)); w.fpn_t[l] = keep.back().data();
            w.lat_g[l] = gn[2 * l]; w.lat_b[l] = gn[2 * l + 1];
            w.fpn_g[l] = gn[2 * L + 2 * l]; w.fpn_b[l] = gn[2 * L + 2 * l + 1];
        }
        if (stages & STAGE_MERGE) {
            keep.push_back(repack_1x1(down_w, kC, kC * L, kC * l, kC)); w.down_t[l] = keep.back().data();
        }
    }
    w.down_g = gn[4 * L]; w.down_b = gn[4 * L + 1];
    const size_t bytes = carve(d, nullptr, nullptr);
    char* ws = static_cast<char*>(malloc(bytes));
    if (!ws) return -1;
    Buffers buf;
    carve(d, ws, &buf);
    HostBackend be;
    neck_run(be, d, w, buf, inputs, x_out, fpn_outs);
    free(ws);
    return 0;
}
