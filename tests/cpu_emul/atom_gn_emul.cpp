// TEST INFRASTRUCTURE ONLY.  Runs ATOM's first-frame joint optimisation -- csrc/atom_gn_kernels.cuh together with the stage-2 kernels it is
// built from (csrc/atom_ops_kernels.cuh conv1x1, csrc/corr_kernels.cuh apply_filter with the 'same' crop / apply_feat_transpose), the same
// sources the CUDA build compiles -- on the CPU under cuda_shim.h, launch for launch as b200trk_atom_gn_joint (csrc/atom_gn.cu) issues them.
// Built and called by tests/test_atom_gn_kernels_cpu.py.
#include "cuda_shim.h"

#include "../../pytracking_b200/csrc/atom_ops_kernels.cuh"
#include "../../pytracking_b200/csrc/corr_kernels.cuh"
#include "../../pytracking_b200/csrc/atom_gn_kernels.cuh"

using namespace b200trk;

namespace {

int pick_passes(int C, int slots, int max_passes) {
    for (int p = max_passes; p >= 1; p >>= 1)
        if (C % (slots * p) == 0) return p;
    return 0;
}

void conv1x1(const float* x, const float* P, float* out, int S, int Cin, int Cout, int H, int W) {          // b200trk_conv1x1
    const int HW = H * W;
    cpu_emul::launch_blocks(conv1x1_kernel, (unsigned)((HW + 63) / 64), (unsigned)((Cout + 63) / 64), (unsigned)S, 256u, (size_t)0, x, P, out, Cin, Cout, HW);
}

template <int FS>
int conv_same(const float* feat, const float* filt, float* scores, int n, int C) {                          // b200trk_conv2d_same
    constexpr int SLOTS = CorrSlots<FS>::value;
    using K = CorrCta<FS, SLOTS>;
    using G = CorrGeom<FS>;
    const int passes = pick_passes(C, SLOTS, n >= 16 ? 4 : 1);
    if (passes <= 0) return 2;
    const int NCH = C / (SLOTS * passes);
    std::vector<unsigned> counters(1024, 0u);
    std::vector<float> part((size_t)n * NCH * G::NPOS, -1e30f);
    const size_t smem = (size_t)(K::PLANES_FLOATS + K::RED_FLOATS + passes * SLOTS * 16) * sizeof(float);
    cpu_emul::launch_blocks(apply_filter_kernel<FS, SLOTS>, (unsigned)NCH, (unsigned)n, 1u, (unsigned)K::NTHREADS, smem, feat, filt, scores, part.data(),
                            counters.data(), C, n, passes, (float*)nullptr, (int64_t*)nullptr, 1);
    return 0;
}

template <int FS>
int feat_transpose(const float* feat, const float* resid, float* grad, int n, int C, int sms) {             // b200trk_apply_feat_transpose
    constexpr int SLOTS = CorrSlots<FS>::value;
    using K = CorrCta<FS, SLOTS>;
    using G = CorrGeom<FS>;
    const int passes = pick_passes(C, SLOTS, 4);
    if (passes <= 0) return 2;
    const int NCH = C / (SLOTS * passes);
    int NG = sms / NCH; if (NG < 1) NG = 1; if (NG > n) NG = n;
    const int SPC_CAP = 8;
    if ((n + NG - 1) / NG > SPC_CAP) NG = (n + SPC_CAP - 1) / SPC_CAP;
    const int spc_max = (n + NG - 1) / NG;
    std::vector<unsigned> counters(1024, 0u);
    std::vector<float> gpart((size_t)NG * C * 16, -1e30f);
    const size_t smem = (size_t)(K::PLANES_FLOATS + K::RED_FLOATS + spc_max * G::NPOS) * sizeof(float);
    cpu_emul::launch_blocks(feat_transpose_kernel<FS, SLOTS>, (unsigned)NCH, (unsigned)NG, 1u, (unsigned)K::NTHREADS, smem, feat, resid, grad, gpart.data(),
                            counters.data(), C, n, passes, spc_max);
    return 0;
}

template <int FS>
int gn_joint(float* filter, float* proj, const float* samples, const float* y, const float* sample_weight, int n, int Cin, int Cc, int num_cg_iter,
             int num_gn_iter, float filter_reg, float projection_reg, int fletcher_reeves, int activation, float act_param, int sms) {
    const int H = FS, W = FS, HW = H * W, HWp = (H + 1) * (W + 1), nw = Cc * 16, nP = Cc * Cin, N = nw + nP;
    const float poison = -1e30f;
    std::vector<float> comp((size_t)n * Cc * HW, poison), compp((size_t)n * Cc * HW, poison), T((size_t)n * Cc * HW, poison), s((size_t)n * HW, poison),
        t1((size_t)n * HW, poison), t2((size_t)n * HW, poison), D((size_t)n * HW, poison), u((size_t)n * HW, poison), upad((size_t)n * HWp, poison),
        gw(nw, poison), part((size_t)n * nP, poison), r(N, poison), rp(N, poison), p(N, poison), x(N, poison), q(N, poison), sc(64, poison);
    GnVec V{r.data(), rp.data(), p.data(), x.data(), q.data(), nw, nP, filter_reg, projection_reg};
    const unsigned mapT = (unsigned)((n * HWp + 255) / 256), expT = (unsigned)((n * Cc * HW + 255) / 256);
    const unsigned tgx = (unsigned)((Cin + 63) / 64), tgy = (unsigned)((Cc + 63) / 64);
    for (int gn = 0; gn < num_gn_iter; ++gn) {
        conv1x1(samples, proj, comp.data(), n, Cin, Cc, H, W);
        if (int e = conv_same<FS>(comp.data(), filter, s.data(), n, Cc)) return e;
        cpu_emul::launch_blocks(gn_linearise_kernel, mapT, 1u, 1u, 256u, (size_t)0, (const float*)s.data(), y, sample_weight, u.data(), upad.data(), D.data(), n, H,
                                W, activation, act_param);
        if (int e = feat_transpose<FS>(comp.data(), upad.data(), gw.data(), n, Cc, sms)) return e;
        cpu_emul::launch_blocks(gn_expand_kernel, expT, 1u, 1u, 256u, (size_t)0, (const float*)u.data(), (const float*)filter, T.data(), n, Cc, H, W);
        cpu_emul::launch_blocks(gn_txt_kernel, tgx, tgy, (unsigned)n, 256u, (size_t)0, (const float*)T.data(), samples, part.data(), Cc, Cin, HW);
        cpu_emul::launch_blocks(gn_init_kernel, 1u, 1u, 1u, 1024u, (size_t)0, V, (const float*)gw.data(), (const float*)part.data(), n, (const float*)filter,
                                (const float*)proj, sc.data());
        for (int ii = 0; ii < num_cg_iter; ++ii) {
            cpu_emul::launch_blocks(gn_dir_kernel, 1u, 1u, 1u, 1024u, (size_t)0, V, sc.data(), fletcher_reeves);
            conv1x1(samples, V.p + nw, compp.data(), n, Cin, Cc, H, W);
            if (int e = conv_same<FS>(comp.data(), V.p, t1.data(), n, Cc)) return e;
            if (int e = conv_same<FS>(compp.data(), filter, t2.data(), n, Cc)) return e;
            cpu_emul::launch_blocks(gn_mapu_kernel, mapT, 1u, 1u, 256u, (size_t)0, (const float*)t1.data(), (const float*)t2.data(), (const float*)D.data(), u.data(),
                                    upad.data(), n, H, W);
            if (int e = feat_transpose<FS>(comp.data(), upad.data(), gw.data(), n, Cc, sms)) return e;
            cpu_emul::launch_blocks(gn_expand_kernel, expT, 1u, 1u, 256u, (size_t)0, (const float*)u.data(), (const float*)filter, T.data(), n, Cc, H, W);
            cpu_emul::launch_blocks(gn_txt_kernel, tgx, tgy, (unsigned)n, 256u, (size_t)0, (const float*)T.data(), samples, part.data(), Cc, Cin, HW);
            cpu_emul::launch_blocks(gn_step_kernel, 1u, 1u, 1u, 1024u, (size_t)0, V, (const float*)gw.data(), (const float*)part.data(), n, sc.data(), fletcher_reeves,
                                    ii == num_cg_iter - 1 ? 1 : 0);
        }
        cpu_emul::launch_blocks(gn_apply_kernel, (unsigned)((N + 255) / 256), 1u, 1u, 256u, (size_t)0, V, filter, proj);
    }
    return 0;
}

}  // namespace

extern "C" int atom_gn_emul_joint(float* filter, float* proj, const float* samples, const float* y, const float* sample_weight, int n, int Cin, int Cc, int H,
                                  int W, int num_cg_iter, int num_gn_iter, float filter_reg, float projection_reg, int fletcher_reeves, int activation,
                                  float act_param, int sms) {
    if (Cc % 16 != 0 || H != W) return 2;
    if (H == 18) return gn_joint<18>(filter, proj, samples, y, sample_weight, n, Cin, Cc, num_cg_iter, num_gn_iter, filter_reg, projection_reg, fletcher_reeves,
                                     activation, act_param, sms);
    if (H == 22) return gn_joint<22>(filter, proj, samples, y, sample_weight, n, Cin, Cc, num_cg_iter, num_gn_iter, filter_reg, projection_reg, fletcher_reeves,
                                     activation, act_param, sms);
    return 2;
}
