// AddressSanitizer / UBSan driver for the two host emulations (odd pyramid shapes, a BEV grid partly outside the input):
// the index arithmetic of every kernel body and the workspace carve-up run under the sanitizers.  Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" int neck_emu_forward(int stages, int L, int B, const int* C, const int* H, const int* W, int groups, float eps,
                                const float* const* lat_w, const float* const* fpn_w, const float* down_w,
                                const float* const* gn, const float* const* inputs, float* x_out, float* const* fpn_outs);
extern "C" int bev_emu_run(int B, int R, int T, int feat, int h, int w, int Ho, int Wo, float bit_scale, float threshold,
                           const float* transform_w, const float* transform_b, const float* emb, const float* grid_y,
                           const float* grid_x, const float* sched, const float* x, const float* noise, const float* replay,
                           float* out, float* feat_dump, float* state_dump);
static std::vector<float> rnd(size_t n) { std::vector<float> v(n); for (auto& x : v) x = (rand() % 2001 - 1000) / 1000.0f; return v; }
int main() {
    // neck: odd pyramid 7x5 -> 4x3 -> 2x2 -> 1x1, B = 2, every stage mask
    {
        const int L = 4, B = 2; int C[4] = {32, 48, 64, 16}, H[4] = {7, 4, 2, 1}, W[4] = {5, 3, 2, 1};
        std::vector<std::vector<float>> lat, fpn, gn, in, fo;
        std::vector<const float*> latp, fpnp, gnp, inp; std::vector<float*> fop;
        for (int l = 0; l < L; ++l) { lat.push_back(rnd(256 * C[l])); fpn.push_back(rnd(256 * 256 * 9)); in.push_back(rnd((size_t)B * C[l] * H[l] * W[l])); fo.push_back(std::vector<float>((size_t)B * 256 * H[l] * W[l])); }
        for (int i = 0; i < 4 * L + 2; ++i) gn.push_back(rnd(256));
        for (int l = 0; l < L; ++l) { latp.push_back(lat[l].data()); fpnp.push_back(fpn[l].data()); inp.push_back(in[l].data()); fop.push_back(fo[l].data()); }
        for (auto& g : gn) gnp.push_back(g.data());
        auto down = rnd(256 * 256 * L);
        std::vector<float> x((size_t)B * 256 * H[0] * W[0]);
        for (int stages : {3, 1}) {
            int rc = neck_emu_forward(stages, L, B, C, H, W, 32, 1e-5f, latp.data(), fpnp.data(), down.data(), gnp.data(), inp.data(), x.data(), fop.data());
            printf("neck stages %d rc %d x[0] %f\n", stages, rc, x[0]);
        }
        int C2[4] = {256, 256, 256, 256};
        std::vector<const float*> fin; for (int l = 0; l < L; ++l) fin.push_back(fo[l].data());
        int rc = neck_emu_forward(2, L, B, C2, H, W, 32, 1e-5f, latp.data(), fpnp.data(), down.data(), gnp.data(), fin.data(), x.data(), nullptr);
        printf("neck stages 2 rc %d x[0] %f\n", rc, x[0]);
    }
    // bev: grid partly outside the input, B = 2, R = 3, T = 2
    {
        const int B = 2, R = 3, T = 2, feat = 32, h = 5, w = 7, Ho = 9, Wo = 6;
        auto tw = rnd(256 * (feat + 256)), tb = rnd(256), emb = rnd(7 * 256), x = rnd((size_t)B * feat * h * w), noise = rnd((size_t)B * R * 256 * h * w);
        std::vector<float> gy(Ho), gx(Wo);
        for (int i = 0; i < Ho; ++i) gy[i] = -1.3f + 2.6f * i / (Ho - 1);
        for (int i = 0; i < Wo; ++i) gx[i] = -1.1f + 2.4f * i / (Wo - 1);
        std::vector<float> sched = {0.1f, 0.8f, 0.99f, 0.6f, 0.8f, 1.0f, 0.14f, 0.003f};
        auto replay = rnd((size_t)T * B * R * 6 * Ho * Wo);
        std::vector<float> out((size_t)B * 6 * Ho * Wo), fd((size_t)T * B * R * 256 * Ho * Wo), sd((size_t)T * B * R * h * w * 256);
        int rc = bev_emu_run(B, R, T, feat, h, w, Ho, Wo, 0.01f, 0.5f, tw.data(), tb.data(), emb.data(), gy.data(), gx.data(), sched.data(), x.data(), noise.data(), replay.data(), out.data(), fd.data(), sd.data());
        printf("bev rc %d out[0] %f\n", rc, out[0]);
    }
    return 0;
}
