// Stage 3 (ATOM initialisation): GaussNewtonCG.run(num_cg_iter, num_gn_iter) on FactorizedConvProblem -- the joint
// optimisation of the filter w [1,Cc,k,k] and the projection matrix P [Cc,Cin,1,1] over the augmented first-frame samples.
//   reference: pytracking/libs/optimization.py:328-421 (run / run_GN_iter / A), :72-163 (run_CG),
//              pytracking/tracker/atom/optim.py:6-68 (residuals, ip_input, M1), wiring pytracking/tracker/atom/atom.py:157-178.
// The reference differentiates through operation.conv1x1 / conv2d twice per CG iteration; here J and J^T are explicit
// (SURVEY.md 9.5), projection_activation = identity (the ATOM default, atom/default.py):
//   comp_i = P X_i ;  s_i = conv_same(comp_i, w) ;  f0 = [sqrt(sw_i)(phi(s_i) - y_i), sqrt(l_w) w, sqrt(l_P) P]
//   J (dw, dP)_i = sqrt(sw_i) phi'(s_i) [conv_same(comp_i, dw) + conv_same(dP X_i, w)]
//   J^T u        = [A_comp^T(sqrt(sw) phi' u) + sqrt(l_w) u_w ;  sum_i B_w^T(sqrt(sw_i) phi'_i u_i) X_i^T + sqrt(l_P) u_P]
// where B_w^T spreads a score-sized map back onto the Cc compressed channels through the filter taps.
// This runs once per sequence, so it is a stream of small kernels with DEVICE-resident CG scalars (no host round trip
// inside the 60 CG iterations) built from the stage-2 kernels (conv1x1, conv2d_same, apply_feat_transpose).
#include "common.cuh"
#include <cstdlib>

extern "C" {
int b200trk_conv1x1(const float*, const float*, float*, int, int, int, int, int, b200trk_stream_t);
int b200trk_conv2d_same(const float*, const float*, float*, int, int, int, int, int, b200trk_stream_t);
int b200trk_apply_feat_transpose(const float*, const float*, float*, int, int, int, int, int, b200trk_stream_t);
}

#include "atom_gn_kernels.cuh"      // the eight gn_* kernels, GnVec

using namespace b200trk;

extern "C" int b200trk_atom_gn_joint(float* filter, float* proj, const float* samples, const float* y, const float* sample_weight,
                                     int n, int Cin, int Cc, int H, int W, int k, int num_cg_iter, int num_gn_iter, float filter_reg,
                                     float projection_reg, int fletcher_reeves, int activation, float act_param,
                                     b200trk_stream_t stream) {
    B200_REQUIRE(filter && proj && samples && y && sample_weight, "atom_gn_joint: null pointer");
    B200_REQUIRE(n > 0 && Cin > 0 && Cc > 0 && Cc % 16 == 0, "atom_gn_joint: bad shape (n=%d, Cin=%d, Cc=%d; Cc must be a multiple of 16)", n, Cin, Cc);
    B200_REQUIRE(k == 4 && H == W && (H == 18 || H == 22), "atom_gn_joint: only a 4x4 filter on 18x18 / 22x22 features is supported");
    B200_REQUIRE(num_cg_iter >= 0 && num_gn_iter >= 0 && activation >= 0 && activation <= 3, "atom_gn_joint: bad iteration counts / activation");
    B200_REQUIRE(filter_reg > 0.f && projection_reg > 0.f, "atom_gn_joint: the M1 preconditioner divides by the regularisation weights");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_cg_iter == 0 || num_gn_iter == 0) return 0;
    const int HW = H * W, HWp = (H + 1) * (W + 1), nw = Cc * 16, nP = Cc * Cin, N = nw + nP;
    // scratch carving (slot 5)
    size_t fl = 0;
    auto take = [&](size_t cnt) { const size_t o = fl; fl += (cnt + 63) / 64 * 64; return o; };
    const size_t o_comp = take((size_t)n * Cc * HW), o_compp = take((size_t)n * Cc * HW), o_T = take((size_t)n * Cc * HW);
    const size_t o_s = take((size_t)n * HW), o_t1 = take((size_t)n * HW), o_t2 = take((size_t)n * HW), o_D = take((size_t)n * HW), o_u = take((size_t)n * HW);
    const size_t o_upad = take((size_t)n * HWp), o_gw = take(nw), o_part = take((size_t)n * nP);
    const size_t o_r = take(N), o_rp = take(N), o_p = take(N), o_x = take(N), o_q = take(N), o_sc = take(64);
    float* ws = (float*)workspace(fl * sizeof(float), 5);
    if (!ws) return 3;
    float *comp = ws + o_comp, *compp = ws + o_compp, *T = ws + o_T, *s = ws + o_s, *t1 = ws + o_t1, *t2 = ws + o_t2, *D = ws + o_D;
    float *u = ws + o_u, *upad = ws + o_upad, *gw = ws + o_gw, *part = ws + o_part, *sc = ws + o_sc;
    GnVec V{ws + o_r, ws + o_rp, ws + o_p, ws + o_x, ws + o_q, nw, nP, filter_reg, projection_reg};
    const int mapT = (n * HWp + 255) / 256, expT = (n * Cc * HW + 255) / 256;
    const dim3 txt_grid((Cin + 63) / 64, (Cc + 63) / 64, n);
    auto launched = [&]() { g_launch_count.fetch_add(1, std::memory_order_relaxed); return cudaGetLastError() == cudaSuccess; };
    for (int gn = 0; gn < num_gn_iter; ++gn) {
        // ---- linearise at (w, P) ----
        if (int e = b200trk_conv1x1(samples, proj, comp, n, Cin, Cc, H, W, stream)) return e;
        if (int e = b200trk_conv2d_same(comp, filter, s, n, Cc, H, W, k, stream)) return e;
        gn_linearise_kernel<<<mapT, 256, 0, st>>>(s, y, sample_weight, u, upad, D, n, H, W, activation, act_param);
        if (!launched()) { set_error("atom_gn_joint: linearise launch failed"); return 1; }
        if (int e = b200trk_apply_feat_transpose(comp, upad, gw, n, Cc, H, W, k, stream)) return e;
        gn_expand_kernel<<<expT, 256, 0, st>>>(u, filter, T, n, Cc, H, W);
        if (!launched()) { set_error("atom_gn_joint: expand launch failed"); return 1; }
        gn_txt_kernel<<<txt_grid, 256, 0, st>>>(T, samples, part, Cc, Cin, HW);
        if (!launched()) { set_error("atom_gn_joint: gemm launch failed"); return 1; }
        gn_init_kernel<<<1, 1024, 0, st>>>(V, gw, part, n, filter, proj, sc);
        if (!launched()) { set_error("atom_gn_joint: init launch failed"); return 1; }
        // ---- CG on J^T J dx = -J^T f0 ----
        for (int ii = 0; ii < num_cg_iter; ++ii) {
            gn_dir_kernel<<<1, 1024, 0, st>>>(V, sc, fletcher_reeves);
            if (!launched()) { set_error("atom_gn_joint: direction launch failed"); return 1; }
            if (int e = b200trk_conv1x1(samples, V.p + nw, compp, n, Cin, Cc, H, W, stream)) return e;        // dP X
            if (int e = b200trk_conv2d_same(comp, V.p, t1, n, Cc, H, W, k, stream)) return e;                 // conv_same(P X, dw)
            if (int e = b200trk_conv2d_same(compp, filter, t2, n, Cc, H, W, k, stream)) return e;             // conv_same(dP X, w)
            gn_mapu_kernel<<<mapT, 256, 0, st>>>(t1, t2, D, u, upad, n, H, W);
            if (!launched()) { set_error("atom_gn_joint: map launch failed"); return 1; }
            if (int e = b200trk_apply_feat_transpose(comp, upad, gw, n, Cc, H, W, k, stream)) return e;
            gn_expand_kernel<<<expT, 256, 0, st>>>(u, filter, T, n, Cc, H, W);
            if (!launched()) { set_error("atom_gn_joint: expand launch failed"); return 1; }
            gn_txt_kernel<<<txt_grid, 256, 0, st>>>(T, samples, part, Cc, Cin, HW);
            if (!launched()) { set_error("atom_gn_joint: gemm launch failed"); return 1; }
            gn_step_kernel<<<1, 1024, 0, st>>>(V, gw, part, n, sc, fletcher_reeves, ii == num_cg_iter - 1 ? 1 : 0);
            if (!launched()) { set_error("atom_gn_joint: step launch failed"); return 1; }
        }
        gn_apply_kernel<<<(N + 255) / 256, 256, 0, st>>>(V, filter, proj);
        if (!launched()) { set_error("atom_gn_joint: apply launch failed"); return 1; }
    }
    return 0;
}
