// Neck in front of the decode loop: FPN + MultiStageMerging as one launch sequence.
//
// Reference: segmentation/mmseg/models/necks/fpn.py:119-140, 162-213 and multi_stage_merging.py:40-52 with the
// arguments every DDP config passes (GN(32) after every conv, no conv bias, no activation, nearest top-down
// upsampling, bilinear align_corners=False merging; e.g. configs/cityscapes/ddp_swin_t_4x4_512x1024_160k_cityscapes.py:38-55).
//
// This header holds (1) the per-element bodies of the neck's elementwise kernels as functors, (2) the host-side weight
// repack and (3) the launch sequence `neck_run`, templated over a backend that supplies `for_each`, `gemm` and
// `tokens_to_nchw`.  The product instantiates it with the CUDA backend in neck.cuh.  tests/emu/neck_emu.cpp instantiates
// the SAME sequence and bodies with a sequential host backend so that indexing, interpolation and buffer wiring can be
// checked against the oracle in the GPU-less build container; that emulation is test infrastructure and is never
// linked into libddp_b200.so (which has no CPU path).
//
// Layout: every intermediate is token-major fp32 [B][N_l][256] (token n = i * W_l + j, channel fastest), the layout
// the decode loop uses, so the 1x1 convolutions are plain GEMMs and the 3x3 convolution is a GEMM whose A loader
// shifts tokens (K = 9 * 256, zero padding).  B200-first restructuring of MultiStageMerging: the reference resizes
// the four 256-channel maps to level 0's size, concatenates 1024 channels and applies the 1x1 `down` conv there;
// here each level goes through ITS 256x256 slice of `down` at its own resolution and the four results are
// bilinearly merged (conv on channels and resize on space commute), which cuts the merge GEMM's FLOPs from
// 4 * N_0 to 1.33 * N_0 tokens and never materialises the 1024-channel tensor.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <vector>

#if defined(__CUDACC__)
#define DDP_HD __host__ __device__ __forceinline__
#else
#define DDP_HD inline
#endif

namespace ddp {
namespace neck {

constexpr int kMaxLevels = 4;
constexpr int kC = 256;          // out_channels of FPN and MultiStageMerging in every DDP config
constexpr int kGnChunk = 64;     // tokens per partial-sum chunk of the GroupNorm statistics

enum { A_ROW_MAJOR = 0, A_NCHW = 1, A_CONV3 = 2 };
enum { STAGE_FPN = 1, STAGE_MERGE = 2 };

// A operand of the 3x3 convolution as a GEMM: element (row, k) with k = tap * C + c, tap = ky * 3 + kx, is channel c
// of the token at (i + ky - 1, j + kx - 1) of the same image, zero outside (F.conv2d(padding=1), fpn.py:130-139).
// Returns the element offset into the token-major map or -1 for padding.
DDP_HD long long conv3_src_offset(long long row, int k, int N, int W, int C) {
    const int tap = k / C, c = k - tap * C;
    const long long img = row / N;
    const int n = (int)(row - img * N);
    const int i = n / W, j = n - i * W;
    const int ii = i + tap / 3 - 1, jj = j + tap % 3 - 1;
    if (ii < 0 || jj < 0 || jj >= W || ii >= N / W) return -1;
    return ((img * N) + (long long)ii * W + jj) * C + c;
}

// F.interpolate(mode='nearest') source index (ATen nearest_neighbor_compute_source_index; scale = in / out in fp32).
DDP_HD int nearest_src(int dst, float scale, int in_size) {
    const int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

// F.interpolate(mode='bilinear', align_corners=False) source coordinate (ATen area_pixel_compute_source_index).
DDP_HD void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float f = scale * ((float)dst + 0.5f) - 0.5f;
    f = f < 0.f ? 0.f : f;
    i0 = (int)f;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = f - (float)i0;
}

// ---- GroupNorm statistics: per (image, channel, chunk of tokens) partial sums, then per (image, group) mean / rstd --
struct GnPartial {           // idx = (b * chunks + chunk) * C + c
    const float* src; double* part; int N, C, chunks;
    DDP_HD void operator()(size_t idx) const {
        const int c = (int)(idx % C);
        const size_t t = idx / C;
        const int chunk = (int)(t % chunks);
        const size_t b = t / chunks;
        const int n0 = chunk * kGnChunk;
        const int n1 = n0 + kGnChunk < N ? n0 + kGnChunk : N;
        double s = 0.0, ss = 0.0;
        for (int n = n0; n < n1; ++n) {
            const double v = (double)src[(b * N + n) * C + c];
            s += v;
            ss += v * v;
        }
        part[idx * 2] = s;
        part[idx * 2 + 1] = ss;
    }
};

struct GnFinalize {          // idx = b * G + g  ->  stats[idx] = (mean, 1 / sqrt(var + eps)), biased variance
    const double* part; float* stats; int N, C, G, chunks; float eps;
    DDP_HD void operator()(size_t idx) const {
        const int g = (int)(idx % G);
        const size_t b = idx / G;
        const int cpg = C / G;
        double s = 0.0, ss = 0.0;
        for (int chunk = 0; chunk < chunks; ++chunk)
            for (int k = 0; k < cpg; ++k) {
                const size_t p = ((b * chunks + chunk) * C + (size_t)g * cpg + k) * 2;
                s += part[p];
                ss += part[p + 1];
            }
        const double cnt = (double)N * cpg;
        const double mean = s / cnt;
        double var = ss / cnt - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        stats[idx * 2] = (float)mean;
        stats[idx * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
};

// y = GN(x) (+ nearest-upsampled upper level: the FPN top-down add, fpn.py:173-183).  dst may alias src.
struct GnApply {             // idx = (b * N + n) * C + c
    const float* src; float* dst; const float* stats; const float* gamma; const float* beta;
    const float* up;         // token-major upper level [B][Hu * Wu][C] or nullptr
    int N, W, C, G, Hu, Wu; float sh, sw;     // sh = Hu / H, sw = Wu / W (fp32, as ATen computes the scale)
    DDP_HD void operator()(size_t idx) const {
        const int c = (int)(idx % C);
        const size_t t = idx / C;
        const int n = (int)(t % N);
        const size_t b = t / N;
        const int g = c / (C / G);
        const float mean = stats[(b * G + g) * 2], rstd = stats[(b * G + g) * 2 + 1];
        float y = (src[idx] - mean) * rstd * gamma[c] + beta[c];
        if (up) {
            const int i = n / W, j = n - i * W;
            const int iu = nearest_src(i, sh, Hu), ju = nearest_src(j, sw, Wu);
            y += up[((b * Hu + iu) * (size_t)Wu + ju) * C + c];
        }
        dst[idx] = y;
    }
};

// x_pre = z_0 + sum_{l >= 1} bilinear_resize(z_l -> (H0, W0)), z_l = O_l * down[:, 256 l : 256 (l + 1)]^T
// (multi_stage_merging.py:44-51 with the 1x1 conv moved in front of the resize).
struct MsmMerge {            // idx = (b * N0 + n) * C + c
    const float* z[kMaxLevels]; int H[kMaxLevels], W[kMaxLevels]; int L, C; float* dst;
    DDP_HD void operator()(size_t idx) const {
        const int c = (int)(idx % C);
        const size_t t = idx / C;
        const int N0 = H[0] * W[0];
        const int n = (int)(t % N0);
        const size_t b = t / N0;
        const int i = n / W[0], j = n - i * W[0];
        float acc = z[0][idx];
#pragma unroll
        for (int l = 1; l < kMaxLevels; ++l) {       // fixed trip count: the parameter arrays stay in constant memory
            if (l >= L) break;
            const int h = H[l], w = W[l];
            int y0, y1, x0, x1; float ly, lx;
            bilinear_src(i, (float)h / (float)H[0], h, y0, y1, ly);
            bilinear_src(j, (float)w / (float)W[0], w, x0, x1, lx);
            const float hy = 1.0f - ly, hx = 1.0f - lx;
            const float* p = z[l] + b * (size_t)h * w * C + c;
            const float top = hx * p[((size_t)y0 * w + x0) * C] + lx * p[((size_t)y0 * w + x1) * C];
            const float bot = hx * p[((size_t)y1 * w + x0) * C] + lx * p[((size_t)y1 * w + x1) * C];
            acc += hy * top + ly * bot;
        }
        dst[idx] = acc;
    }
};

// ---- weights: host-side repack into the GEMMs' [K][256] (K-major rows, output channel fastest) layout -------------
// lateral (256, C_l, 1, 1)      -> [C_l][256]
// fpn     (256, 256, 3, 3)      -> [(ky * 3 + kx) * 256 + ci][256]
// down    (256, 256 * L, 1, 1)  -> per level l: [ci][256] from columns 256 l .. 256 l + 255
inline std::vector<float> repack_1x1(const float* w, int cout, int cin_total, int cin0, int cin) {
    std::vector<float> t((size_t)cin * cout);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci) t[(size_t)ci * cout + co] = w[(size_t)co * cin_total + cin0 + ci];
    return t;
}

inline std::vector<float> repack_3x3(const float* w, int cout, int cin) {
    std::vector<float> t((size_t)9 * cin * cout);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int tap = 0; tap < 9; ++tap)
                t[((size_t)tap * cin + ci) * cout + co] = w[((size_t)co * cin + ci) * 9 + tap];
    return t;
}

struct Dims {
    int L, B, stages;
    int C[kMaxLevels];       // input channels per level (STAGE_MERGE only: 256 each)
    int H[kMaxLevels], W[kMaxLevels];
    int groups; float eps;
    long long tokens(int l) const { return (long long)H[l] * W[l]; }
};

struct Weights {             // device pointers (CUDA backend) or host pointers (emulation)
    const float* lat_t[kMaxLevels]; const float* lat_g[kMaxLevels]; const float* lat_b[kMaxLevels];
    const float* fpn_t[kMaxLevels]; const float* fpn_g[kMaxLevels]; const float* fpn_b[kMaxLevels];
    const float* down_t[kMaxLevels]; const float* down_g; const float* down_b;
};

struct Buffers {
    float* lat[kMaxLevels];  // [B][N_l][256] lateral -> top-down sum (in place); lat[0] is reused for x_pre
    float* fo[kMaxLevels];   // [B][N_l][256] FPN outputs
    float* z[kMaxLevels];    // [B][N_l][256] per-level slices of the `down` conv
    double* part;            // GroupNorm partial sums
    float* stats;            // [B][groups][2]
    char* conv_scratch;      // tensor-core 3x3 convolution (CUDA build): zero-bordered fp16 planes + bordered fp32 output
};

// bytes of the bordered-grid scratch of the tensor-core 3x3 convolution at the largest level: per bordered token
// (H + 2) * (W + 2): 256 channels x (fp16 hi + fp16 lo + fp32 out)
inline size_t conv_scratch_bytes(const Dims& d) {
    size_t mx = 0;
    for (int l = 0; l < d.L; ++l) {
        const size_t mp = (size_t)d.B * (d.H[l] + 2) * (d.W[l] + 2);
        mx = mp > mx ? mp : mx;
    }
    return mx * kC * 8 + 1024;
}

inline size_t gn_chunks(long long N) { return (size_t)((N + kGnChunk - 1) / kGnChunk); }

// Workspace carve-up shared by the CUDA build and the emulation (offsets in bytes, 256-byte aligned).
inline size_t carve(const Dims& d, char* base, Buffers* out) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return p;
    };
    Buffers b{};
    for (int l = 0; l < d.L; ++l) {
        const size_t n = (size_t)d.B * d.tokens(l) * kC * sizeof(float);
        b.lat[l] = reinterpret_cast<float*>(take(n));
        b.fo[l] = reinterpret_cast<float*>(take(n));
        b.z[l] = reinterpret_cast<float*>(take(n));
    }
    long long max_tokens = 0;
    for (int l = 0; l < d.L; ++l) max_tokens = d.tokens(l) > max_tokens ? d.tokens(l) : max_tokens;
    b.part = reinterpret_cast<double*>(take((size_t)d.B * gn_chunks(max_tokens) * kC * 2 * sizeof(double)));
    b.stats = reinterpret_cast<float*>(take((size_t)d.B * d.groups * 2 * sizeof(float)));
    b.conv_scratch = (d.stages & STAGE_FPN) ? take(conv_scratch_bytes(d)) : nullptr;
    if (out) *out = b;
    return off;
}

// GroupNorm of a token-major map, in place, optionally adding the nearest-upsampled upper level.
template <class Backend>
void group_norm(Backend& be, const Dims& d, const Buffers& buf, float* x, int l, const float* gamma, const float* beta,
                const float* up, int lu) {
    const int N = (int)d.tokens(l);
    const int chunks = (int)gn_chunks(N);
    be.for_each((size_t)d.B * chunks * kC, GnPartial{x, buf.part, N, kC, chunks});
    be.for_each((size_t)d.B * d.groups, GnFinalize{buf.part, buf.stats, N, kC, d.groups, chunks, d.eps});
    GnApply a{x, x, buf.stats, gamma, beta, up, N, d.W[l], kC, d.groups, 1, 1, 1.f, 1.f};
    if (up) {
        a.Hu = d.H[lu]; a.Wu = d.W[lu];
        a.sh = (float)d.H[lu] / (float)d.H[l];
        a.sw = (float)d.W[lu] / (float)d.W[l];
    }
    be.for_each((size_t)d.B * N * kC, a);
}

// The launch sequence.  inputs[l]: NCHW (B, C_l, H_l, W_l).  x_out: NCHW (B, 256, H_0, W_0) (STAGE_MERGE).
// fpn_outs[l]: optional NCHW copies of the FPN outputs (required when stages == STAGE_FPN).
template <class Backend>
void neck_run(Backend& be, const Dims& d, const Weights& w, const Buffers& buf, const float* const* inputs, float* x_out,
              float* const* fpn_outs) {
    const bool do_fpn = (d.stages & STAGE_FPN) != 0, do_merge = (d.stages & STAGE_MERGE) != 0;
    if (do_fpn) {
        // laterals: 1x1 conv straight from the NCHW backbone maps (fpn.py:165-168)
        for (int l = 0; l < d.L; ++l)
            be.gemm(A_NCHW, inputs[l], 0, (int)d.tokens(l), w.lat_t[l], (long long)d.B * d.tokens(l), d.C[l], buf.lat[l]);
        // GN, then the top-down path from the coarsest level (fpn.py:171-183): level l adds the FINISHED level l + 1
        for (int l = d.L - 1; l >= 0; --l)
            group_norm(be, d, buf, buf.lat[l], l, w.lat_g[l], w.lat_b[l], l + 1 < d.L ? buf.lat[l + 1] : nullptr, l + 1);
        // 3x3 output convs + GN (fpn.py:186-188)
        for (int l = 0; l < d.L; ++l) {
            be.gemm(A_CONV3, buf.lat[l], d.W[l], (int)d.tokens(l), w.fpn_t[l], (long long)d.B * d.tokens(l), 9 * kC, buf.fo[l]);
            group_norm(be, d, buf, buf.fo[l], l, w.fpn_g[l], w.fpn_b[l], nullptr, 0);
            if (fpn_outs && fpn_outs[l]) be.tokens_to_nchw(buf.fo[l], fpn_outs[l], d.B, (int)d.tokens(l), kC);
        }
    }
    if (do_merge) {
        for (int l = 0; l < d.L; ++l) {
            if (do_fpn)
                be.gemm(A_ROW_MAJOR, buf.fo[l], kC, 0, w.down_t[l], (long long)d.B * d.tokens(l), kC, buf.z[l]);
            else
                be.gemm(A_NCHW, inputs[l], 0, (int)d.tokens(l), w.down_t[l], (long long)d.B * d.tokens(l), kC, buf.z[l]);
        }
        MsmMerge m{};
        for (int l = 0; l < d.L; ++l) { m.z[l] = buf.z[l]; m.H[l] = d.H[l]; m.W[l] = d.W[l]; }
        m.L = d.L; m.C = kC; m.dst = buf.lat[0];
        be.for_each((size_t)d.B * d.tokens(0) * kC, m);
        group_norm(be, d, buf, buf.lat[0], 0, w.down_g, w.down_b, nullptr, 0);
        be.tokens_to_nchw(buf.lat[0], x_out, d.B, (int)d.tokens(0), kC);
    }
}

}  // namespace neck
}  // namespace ddp
