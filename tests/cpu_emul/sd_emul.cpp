// TEST INFRASTRUCTURE ONLY.  Runs csrc/sd_kernel.cuh (the DiMP / PrDiMP / DiMP-L2 / GNSteepestDescent+LinearFilterHinge online optimisers as
// one persistent cooperative kernel on the cp.async sweeps of csrc/corr2.cuh -- the same sources the CUDA build compiles) on the CPU under
// cuda_shim.h, with the parameter blocks of the four C entry points of csrc/sd_optimizer.cu.  Two builds: plain (OS thread per CUDA thread,
// what ThreadSanitizer needs) runs ONE CTA (static __shared__ arrays) -- the decomposition launch_sd picks on a 1-SM device; with
// -DB200_EMUL_COOP_FIBERS (block = OS thread, threads = fibers, static __shared__ = thread_local) the grid launch_sd picks for `sms` SMs.  Built and called by
// tests/test_sd_kernels_cpu.py.
#include "cuda_shim.h"

#include "../../pytracking_b200/csrc/sd_kernel.cuh"

using namespace b200trk;

static int g_sms = 1;                                        // set by the entry points (plain build: always one CTA)

template <int FS, int MODE>
static int run_sd(SdParams P) {
    using K = Corr2<FS>;
    constexpr int SLOTS = K::SLOTS;
    // launch_sd (csrc/sd_optimizer.cu): passes x 16 channels per CTA (NCH chunks) x NG sample groups, scored by the busiest CTA
#ifdef B200_EMUL_COOP_FIBERS
    const int sms = g_sms;
#else
    const int sms = 1;                                       // one live block only (static __shared__ arrays)
#endif
    const size_t limit = 227 * 1024 - 512, item = (size_t)K::ITEM_FLOATS * sizeof(float);
    int passes = 0, NCH = 0, NG = 0, spc = 0, best_cost = 1 << 30;
    for (int p = 64; p >= 1; p >>= 1) {
        if (P.C % (SLOTS * p) != 0) continue;
        const int nch = P.C / (SLOTS * p);
        if (nch > sms) continue;
        int ng = sms / nch; if (ng > P.n) ng = P.n; if (ng < 1) ng = 1;
        const int sp = (P.n + ng - 1) / ng;
        if (sp > SD_SPC_MAX) continue;
        const size_t fx = (size_t)(K::NT * SLOTS * K::RED_STRIDE + 2 * p * SLOTS * K::VEC_STRIDE + sp * (5 * K::NPOS + K::PMAP)) * sizeof(float);
        if (fx + 2 * item > limit) continue;
        int cost = sp * p * 4;
        if (fx + 3 * item > limit) cost += cost / 2;
        if (p > 4) cost += 1;
        if (cost < best_cost) { best_cost = cost; passes = p; NCH = nch; NG = ng; spc = sp; }
    }
    if (!passes || P.num_iter + 1 > K::NCONS) return 2;
    P.passes = passes; P.NCH = NCH; P.NG = NG; P.spc_max = spc; P.dbg_mode = 0; P.trace = nullptr;
    const size_t n_gpart = (size_t)NG * P.C * 16, n_qpart = (size_t)P.n * NCH * K::NPOS, n_loss = (size_t)(P.num_iter + 1) * (NG + NCH);
    std::vector<float> ws(n_gpart + n_qpart + NG + NCH + n_loss + 64, -1e30f);
    std::vector<unsigned> bar(256, 0u);
    P.barrier = bar.data();
    float* f = ws.data();
    P.gpart = f; f += n_gpart;
    P.qpart = f; f += n_qpart;
    P.hpart = f; f += NG;
    P.gnorm = f; f += NCH;
    P.lossr = f; f += (size_t)(P.num_iter + 1) * NG;
    P.lossw = f;
    const int cchunk = passes * SLOTS;
    const size_t fixed = (size_t)(K::NT * SLOTS * K::RED_STRIDE + 2 * cchunk * K::VEC_STRIDE + spc * (5 * K::NPOS + K::PMAP)) * sizeof(float);
    const unsigned grid = (unsigned)(NCH * NG);
#ifdef B200_EMUL_COOP_FIBERS
#define SD_LAUNCH(NST) cpu_emul::launch_coop(sd_kernel<FS, NST, MODE>, grid, (unsigned)K::NCONS, fixed + NST * item, P)
#else
#define SD_LAUNCH(NST) cpu_emul::launch(sd_kernel<FS, NST, MODE>, grid, (unsigned)K::NCONS, fixed + NST * item, P)
#endif
    if (fixed + 4 * item <= limit) SD_LAUNCH(4);
    else if (fixed + 3 * item <= limit) SD_LAUNCH(3);
    else if (fixed + 2 * item <= limit) SD_LAUNCH(2);
    else return 2;
#undef SD_LAUNCH
    return 0;
}

template <int MODE>
static int dispatch(const SdParams& P, int H, int W) {
    if (H == 18 && W == 18) return run_sd<18, MODE>(P);
    if (H == 22 && W == 22) return run_sd<22, MODE>(P);
    return 2;
}

static void common(SdParams& P, const float* w, float* wo, const float* feat, const float* bb, const float* sw, int n, int C, int num_iter, float* its,
                   float* losses) {
    P.w_in = w; P.w_out = wo; P.feat = feat; P.bb = bb; P.sample_weight = sw; P.n = n; P.C = C; P.num_iter = num_iter;
    P.iterates_out = its; P.losses_out = losses;
    if (its) std::memcpy(its, w, (size_t)C * 16 * sizeof(float));       // the entry points copy the initial filter into iterates[0]
}

extern "C" int sd_emul_dimp_sd_gn(const float* w, float* wo, const float* feat, const float* bb, const float* sw, int n, int C, int H, int W, int num_iter,
                                  const float* label_lut, const float* mask_lut, const float* spatial_lut, int num_bins, float bin_displacement,
                                  float feat_stride, float step_length, float reg_weight, float alpha_eps, float* its, float* losses, int sms) {
    g_sms = sms;
    SdParams P{};
    common(P, w, wo, feat, bb, sw, n, C, num_iter, its, losses);
    P.label_lut = label_lut; P.mask_lut = mask_lut; P.spatial_lut = spatial_lut; P.num_bins = num_bins;
    P.inv_bin_disp = 1.0f / bin_displacement; P.inv_feat_stride = 1.0f / feat_stride;
    P.step_length = step_length; P.reg_weight = reg_weight; P.alpha_eps = alpha_eps;
    return dispatch<0>(P, H, W);
}

extern "C" int sd_emul_prdimp_sd_newton(const float* w, float* wo, const float* feat, const float* bb, const float* sw, int n, int C, int H, int W,
                                        int num_iter, float gauss_sigma, float feat_stride, float step_length, float reg_weight, float alpha_eps,
                                        int has_softmax_reg, float softmax_reg, float label_threshold, int normalize_label, float label_shrink,
                                        float uni_weight, float* its, float* losses, int sms) {
    g_sms = sms;
    SdParams P{};
    common(P, w, wo, feat, bb, sw, n, C, num_iter, its, losses);
    P.gauss_sigma = gauss_sigma; P.has_softmax_reg = has_softmax_reg; P.softmax_reg = softmax_reg; P.label_threshold = label_threshold;
    P.normalize_label = normalize_label; P.label_shrink = label_shrink; P.uni_weight = uni_weight;
    P.inv_feat_stride = 1.0f / feat_stride; P.step_length = step_length; P.reg_weight = reg_weight; P.alpha_eps = alpha_eps;
    return dispatch<1>(P, H, W);
}

extern "C" int sd_emul_dimp_l2_sd_gn(const float* w, float* wo, const float* feat, const float* bb, const float* sw, int n, int C, int H, int W, int num_iter,
                                     float gauss_sigma, float hinge_threshold, float feat_stride, float step_length, float reg_weight, float alpha_eps,
                                     float* its, float* losses, int sms) {
    g_sms = sms;
    SdParams P{};
    common(P, w, wo, feat, bb, sw, n, C, num_iter, its, losses);
    P.gauss_sigma = gauss_sigma; P.label_threshold = hinge_threshold;
    P.inv_feat_stride = 1.0f / feat_stride; P.step_length = step_length; P.reg_weight = reg_weight; P.alpha_eps = alpha_eps;
    return dispatch<2>(P, H, W);
}

extern "C" int sd_emul_gn_sd_hinge(const float* w, float* wo, const float* feat, const float* train_label, const float* sw, int n, int C, int H, int W,
                                   int num_iter, float filter_reg, float hinge_threshold, float activation_leak, int score_act, float act_param,
                                   float steplength_reg, float* its, float* losses, int sms) {
    g_sms = sms;
    SdParams P{};
    common(P, w, wo, feat, train_label, sw, n, C, num_iter, its, losses);
    P.label_in = train_label; P.label_threshold = hinge_threshold; P.act_leak = activation_leak; P.act_kind = score_act; P.act_b = act_param;
    P.inv_feat_stride = 1.f / 16.f;
    P.step_length = 1.f; P.reg_weight = filter_reg * filter_reg; P.alpha_eps = steplength_reg;
    P.loss_scale = 1.f / ((float)n * (float)((H + 1) * (W + 1)) + (float)C * 16.f);
    return dispatch<3>(P, H, W);
}
