// Stage 3 (ATOM): ConjugateGradient.run(num_iter) on ConvProblem as ONE persistent cooperative kernel.
//   reference: pytracking/libs/optimization.py:227-275 (run), :72-163 (run_CG), problem pytracking/tracker/atom/optim.py:71-99,
//   wiring pytracking/tracker/atom/atom.py:189-217 (direction_forget_factor = 0: the CG state is reset every run).
//
// The reference gets J p and J^T u by double backward through operation.conv2d; here they are explicit:
//   s_i = conv_same(X_i, w) = rows/cols 0..FS-1 of the (FS+1)^2 correlation map A w
//   f0  = [sqrt(sw_i) (phi(s_i) - y_i), sqrt(reg) w]
//   b   = -J^T f0 = -(A^T(sw_i phi'(s_i) (phi(s_i) - y_i)) + reg w)
//   A_cg p = J^T J p = A^T(sw_i phi'(s_i)^2 (A p)_i) + reg p          (maps masked to the FS x FS window)
// CG loop exactly as run_CG (M1 = M2 = identity for ConvProblem, Polak-Ribiere or Fletcher-Reeves beta clamped at 0,
// alpha = rho / <p,q>, residual not updated on the last iteration).
//
// Decomposition and data movement are those of the DiMP optimiser (sd_optimizer.cu): CTA = (channel chunk, sample
// group), corr2.cuh sweeps, per-sample maps and the chunk slices of w, r, p, x, r_prev, q stay in shared memory for the
// whole call. Cross-CTA exchange per CG iteration: qpart (A p partial maps), gpart (A^T partial gradients) and, when the
// filter spans several channel chunks, three scalars (<r,r>, <r_prev,r>, <p,q>) -- all summed in a fixed order.
#include "atom_cg_kernel.cuh"      // corr2.cuh + CgParams, atom_cg_kernel
#include <cstdlib>

namespace b200trk {

template <int FS, int NST>
static int launch_cg_nst(const CgParams& P, size_t smem, cudaStream_t st) {
    using K = Corr2<FS>;
    auto kern = atom_cg_kernel<FS, NST>;
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&P};
    B200_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(P.NCH * P.NG), dim3(K::NCONS), args, smem, st));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

template <int FS>
static int launch_cg(CgParams P, cudaStream_t st) {
    using K = Corr2<FS>;
    const int sms = device_sm_count();
    // fewest channel passes per CTA that still keeps the samples of a CTA within shared memory: more chunks means
    // fewer sample groups, i.e. a cheaper cross-group gradient reduction
    int passes = 0, NCH = 0, NG = 0, spc = 0;
    for (int p = 1; p <= 4; p <<= 1) {
        if (P.C % (K::SLOTS * p) != 0) continue;
        const int nch = P.C / (K::SLOTS * p);
        if (nch > sms) continue;
        int ng = sms / nch; if (ng > P.n) ng = P.n; if (ng < 1) ng = 1;
        const int s = (P.n + ng - 1) / ng;
        if (s <= CG_SPC_MAX) { passes = p; NCH = nch; NG = ng; spc = s; break; }
    }
    B200_REQUIRE(passes > 0, "atom_cg_filter: n=%d samples x C=%d channels do not fit the persistent kernel", P.n, P.C);
    P.passes = passes; P.NCH = NCH; P.NG = NG; P.spc_max = spc;
    const size_t n_gpart = (size_t)NG * P.C * 16, n_qpart = (size_t)P.n * NCH * K::NPOS;
    const size_t n_dots = (size_t)(2 * P.num_iter + 4) * NCH * 2;
    const size_t total = (n_gpart + n_qpart + n_dots + 64) * sizeof(float) + 256;
    char* ws = (char*)workspace(total, 4);
    if (!ws) return 3;
    P.barrier = (unsigned*)ws;
    float* f = (float*)(ws + 256);
    P.gpart = f; f += n_gpart;
    P.qpart = f; f += n_qpart;
    P.dots = f;
    B200_CHECK_CUDA(cudaMemsetAsync(P.barrier, 0, 256, st));
    const int cchunk = passes * K::SLOTS;
    const size_t fixed = (size_t)(K::NT * K::SLOTS * K::RED_STRIDE + 6 * cchunk * K::VEC_STRIDE + spc * (2 * K::NPOS + K::PMAP) + K::NCONS * 4) * sizeof(float);
    const size_t limit = 227 * 1024 - 512;
    const size_t item = (size_t)K::ITEM_FLOATS * sizeof(float);
    if (fixed + 4 * item <= limit) return launch_cg_nst<FS, 4>(P, fixed + 4 * item, st);
    if (fixed + 3 * item <= limit) return launch_cg_nst<FS, 3>(P, fixed + 3 * item, st);
    B200_REQUIRE(fixed + 2 * item <= limit, "atom_cg_filter: %d samples per CTA do not fit in shared memory", spc);
    return launch_cg_nst<FS, 2>(P, fixed + 2 * item, st);
}

}  // namespace b200trk

using namespace b200trk;

extern "C" int b200trk_atom_cg_filter(const float* filter, float* filter_out, const float* feat, const float* y,
                                      const float* sample_weight, int n, int C, int H, int W, int k, int num_iter,
                                      float filter_reg, int fletcher_reeves, int activation, float act_param,
                                      b200trk_stream_t stream) {
    B200_REQUIRE(filter && filter_out && feat && y && sample_weight, "atom_cg_filter: null pointer");
    B200_REQUIRE(n > 0 && C > 0, "atom_cg_filter: empty sample memory (n=%d, C=%d)", n, C);
    B200_REQUIRE(k == 4, "atom_cg_filter: filter size %d not supported by the CUDA path (only 4)", k);
    B200_REQUIRE(H == W && (H == 18 || H == 22), "atom_cg_filter: feature size %dx%d not supported (18x18, 22x22)", H, W);
    B200_REQUIRE(num_iter >= 0 && num_iter <= 256, "atom_cg_filter: num_iter=%d", num_iter);
    B200_REQUIRE(activation >= 0 && activation <= 3, "atom_cg_filter: activation %d (0 none, 1 relu, 2 elu, 3 mlu)", activation);
    B200_REQUIRE(C % 16 == 0, "atom_cg_filter: C=%d must be a multiple of 16", C);
    cudaStream_t st = (cudaStream_t)stream;
    if (num_iter == 0) {
        if (filter_out != filter)
            B200_CHECK_CUDA(cudaMemcpyAsync(filter_out, filter, (size_t)C * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    CgParams P{};
    P.w_in = filter; P.w_out = filter_out; P.feat = feat; P.y = y; P.sample_weight = sample_weight;
    P.n = n; P.C = C; P.num_iter = num_iter; P.fletcher_reeves = fletcher_reeves ? 1 : 0; P.act = activation;
    P.act_param = act_param; P.reg = filter_reg;
    if (H == 18) return launch_cg<18>(P, st);
    return launch_cg<22>(P, st);
}
