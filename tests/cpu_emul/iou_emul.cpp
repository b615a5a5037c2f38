// TEST INFRASTRUCTURE ONLY.  Runs csrc/iou_refine_kernels.cuh together with csrc/prroi_kernels.cuh (IoUNet predict_iou, its analytic box
// gradient and the refinement loops -- the same sources the CUDA build compiles) on the CPU under cuda_shim.h, in the launch sequence of
// iou_eval / b200trk_iou_predict / b200trk_iou_refine (csrc/iou_refine.cu).  Built and called by tests/test_iou_kernels_cpu.py.
#include "cuda_shim.h"

#include "../../include/b200trk.h"

#include "../../pytracking_b200/csrc/prroi_kernels.cuh"

using namespace b200trk;

#include "../../pytracking_b200/csrc/iou_refine_kernels.cuh"

namespace {

struct Pred {
    int C3, P3, C4, P4, D3, D4;
    std::vector<float> w3, b3, w4, b4, wp, rois, pool3, pool4, gpool3, gpool4, act, part, grois3, grois4, rel, sznorm, step;
    float bp;
};

void prroi_forward(const float* feat, const float* rois, float* out, int C, int H, int W, int R, int ph, int pw, float scale) {
    cpu_emul::launch_blocks(prroi_forward_kernel, (unsigned)(ph * pw), (unsigned)R, 1u, 128u, (size_t)0, feat, rois, out, C, H, W, ph, pw, scale);
}

void prroi_coor_backward(const float* feat, const float* rois, const float* out, const float* ograd, float* rgrad, int C, int H, int W, int R, int ph, int pw,
                         float scale) {
    std::vector<float> part((size_t)R * ph * pw * 4, -1e30f);
    cpu_emul::launch_blocks(prroi_coor_backward_kernel, (unsigned)(ph * pw), (unsigned)R, 1u, 128u, (size_t)0, feat, rois, out, ograd, part.data(), C, H, W, ph, pw,
                            scale);
    cpu_emul::launch_blocks(prroi_coor_reduce_kernel, (unsigned)((R + 63) / 64), 1u, 1u, 64u, (size_t)0, (const float*)part.data(), rgrad, R, ph * pw);
}

// iou_eval of csrc/iou_refine.cu, launch for launch
void iou_eval(Pred& p, const float* mod3, const float* mod4, const float* feat3, int H3, int W3, const float* feat4, int H4, int W4, float* boxes, int R,
              float* iou_out, float* grad_out, int step_mode, float decay, bool need_grad) {
    const int K3 = p.C3 * p.P3 * p.P3, K4 = p.C4 * p.P4 * p.P4;
    cpu_emul::launch_blocks(make_rois_kernel, 1u, 1u, 1u, 32u, (size_t)0, (const float*)boxes, p.rois.data(), R);
    prroi_forward(feat3, p.rois.data(), p.pool3.data(), p.C3, H3, W3, R, p.P3, p.P3, 1.f / 8.f);
    prroi_forward(feat4, p.rois.data(), p.pool4.data(), p.C4, H4, W4, R, p.P4, p.P4, 1.f / 16.f);
    const int ns3 = (K3 + FC_KS - 1) / FC_KS, ns4 = (K4 + FC_KS - 1) / FC_KS, nb3 = (p.D3 + 7) / 8, nb4 = (p.D4 + 7) / 8;
    cpu_emul::launch_blocks(fc_forward_kernel<IOU_RMAX>, (unsigned)(nb3 * ns3 + nb4 * ns4), 1u, 1u, 256u, (size_t)0, (const float*)p.w3.data(), (const float*)p.w4.data(),
                            (const float*)p.pool3.data(), (const float*)p.pool4.data(), mod3, mod4, p.part.data(), R, K3, p.P3 * p.P3, K4, p.P4 * p.P4, p.D3, p.D4, nb3,
                            ns3, ns4);
    cpu_emul::launch_blocks(iou_head_kernel, 1u, 1u, 1u, (unsigned)(32 * IOU_RMAX), (size_t)0, (const float*)p.part.data(), (const float*)p.b3.data(),
                            (const float*)p.b4.data(), p.act.data(), (const float*)p.wp.data(), p.bp, iou_out, R, p.D3, p.D4, ns3, ns4, IOU_RMAX);
    if (!need_grad) return;
    const int nb = (K3 + 255) / 256 + (K4 + 255) / 256;
    const size_t smem = (size_t)(p.D3 > p.D4 ? p.D3 : p.D4) * IOU_RMAX * sizeof(float);
    cpu_emul::launch_blocks(fc_backward_kernel<IOU_RMAX>, (unsigned)nb, 1u, 1u, 256u, smem, (const float*)p.w3.data(), (const float*)p.w4.data(),
                            (const float*)p.act.data(), (const float*)p.wp.data(), mod3, mod4, p.gpool3.data(), p.gpool4.data(), R, K3, p.P3 * p.P3, K4, p.P4 * p.P4,
                            p.D3, p.D4);
    prroi_coor_backward(feat3, p.rois.data(), p.pool3.data(), p.gpool3.data(), p.grois3.data(), p.C3, H3, W3, R, p.P3, p.P3, 1.f / 8.f);
    prroi_coor_backward(feat4, p.rois.data(), p.pool4.data(), p.gpool4.data(), p.grois4.data(), p.C4, H4, W4, R, p.P4, p.P4, 1.f / 16.f);
    cpu_emul::launch_blocks(box_step_kernel, 1u, 1u, 1u, 32u, (size_t)0, (const float*)p.grois3.data(), (const float*)p.grois4.data(), boxes, p.rel.data(),
                            (const float*)p.sznorm.data(), grad_out, p.step.data(), decay, R, step_mode);
}

}  // namespace

// num_iter == 0: b200trk_iou_predict (iou_out, optional grad_out); num_iter >= 1: b200trk_iou_refine (boxes updated in place)
extern "C" int iou_emul_run(const b200trk_linear_block_t* fc3, const b200trk_linear_block_t* fc4, const float* wp, const float* bp, int C3, int P3, int C4,
                            int P4, int D3, int D4, const float* mod3, const float* mod4, const float* feat3, int H3, int W3, const float* feat4, int H4,
                            int W4, float* boxes, int R, int num_iter, float step_length, float step_decay, int relative, float* iou_out, float* grad_out) {
    if (R < 1 || R > IOU_RMAX) return 2;
    Pred p;
    p.C3 = C3; p.P3 = P3; p.C4 = C4; p.P4 = P4; p.D3 = D3; p.D4 = D4;
    const int K3 = C3 * P3 * P3, K4 = C4 * P4 * P4;
    fold_linear(*fc3, D3, K3, p.w3, p.b3);
    fold_linear(*fc4, D4, K4, p.w4, p.b4);
    p.wp.assign(wp, wp + D3 + D4);
    p.bp = bp ? bp[0] : 0.f;
    const float poison = -1e30f;
    p.rois.assign(IOU_RMAX * 5, poison); p.pool3.assign((size_t)IOU_RMAX * K3, poison); p.pool4.assign((size_t)IOU_RMAX * K4, poison);
    p.gpool3.assign((size_t)IOU_RMAX * K3, poison); p.gpool4.assign((size_t)IOU_RMAX * K4, poison); p.act.assign((size_t)IOU_RMAX * (D3 + D4), poison);
    p.part.assign((size_t)((K3 > K4 ? K3 : K4) / FC_KS + 1) * IOU_RMAX * (D3 + D4), poison);
    p.grois3.assign(IOU_RMAX * 5, poison); p.grois4.assign(IOU_RMAX * 5, poison); p.rel.assign(IOU_RMAX * 4, poison); p.sznorm.assign(4, poison);
    p.step.assign(4, poison);
    if (num_iter == 0) {
        iou_eval(p, mod3, mod4, feat3, H3, W3, feat4, H4, W4, boxes, R, iou_out, grad_out, 0, 1.f, grad_out != nullptr);
        return 0;
    }
    cpu_emul::launch_blocks(set_scalar_kernel, 1u, 1u, 1u, 1u, (size_t)0, p.step.data(), step_length);
    if (relative) cpu_emul::launch_blocks(to_rel_kernel, 1u, 1u, 1u, 32u, (size_t)0, (const float*)boxes, p.rel.data(), p.sznorm.data(), R);
    for (int it = 0; it < num_iter; ++it)
        iou_eval(p, mod3, mod4, feat3, H3, W3, feat4, H4, W4, boxes, R, it == num_iter - 1 ? iou_out : nullptr, nullptr, relative ? 2 : 1, step_decay, true);
    return 0;
}
