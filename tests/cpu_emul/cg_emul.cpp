// TEST INFRASTRUCTURE ONLY.  Runs csrc/atom_cg_kernel.cuh (ATOM's per-frame ConjugateGradient.run on ConvProblem as one persistent
// cooperative kernel on the cp.async sweeps of csrc/corr2.cuh -- the same sources the CUDA build compiles) on the CPU under cuda_shim.h.
// ONE CTA only: the kernel keeps small arrays in static __shared__ storage, which the shim can give to a single live block, so the
// decomposition is the one launch_cg (csrc/atom_cg.cu) picks on a 1-SM device -- all channels and all (<= 8) samples in one CTA; the
// cross-CTA exchange is left to the -m gpu tests.  Built and called by tests/test_sd_kernels_cpu.py.
#include "cuda_shim.h"

#include "../../pytracking_b200/csrc/atom_cg_kernel.cuh"

using namespace b200trk;

template <int FS>
static int run_cg(CgParams P) {
    using K = Corr2<FS>;
    int passes = 0;
    for (int p = 1; p <= 4; p <<= 1)
        if (P.C == K::SLOTS * p) passes = p;
    if (!passes || P.n > CG_SPC_MAX) return 2;
    P.passes = passes; P.NCH = 1; P.NG = 1; P.spc_max = P.n;
    const size_t n_gpart = (size_t)P.C * 16, n_qpart = (size_t)P.n * K::NPOS, n_dots = (size_t)(2 * P.num_iter + 4) * 2;
    std::vector<float> ws(n_gpart + n_qpart + n_dots + 128, -1e30f);
    std::vector<unsigned> bar(64, 0u);
    P.barrier = bar.data();
    P.gpart = ws.data(); P.qpart = P.gpart + n_gpart; P.dots = P.qpart + n_qpart;
    const int cchunk = passes * K::SLOTS;
    const size_t fixed = (size_t)(K::NT * K::SLOTS * K::RED_STRIDE + 6 * cchunk * K::VEC_STRIDE + P.n * (2 * K::NPOS + K::PMAP) + K::NCONS * 4) * sizeof(float);
    const size_t item = (size_t)K::ITEM_FLOATS * sizeof(float);
    cpu_emul::launch(atom_cg_kernel<FS, 2>, 1u, (unsigned)K::NCONS, fixed + 2 * item, P);
    return 0;
}

extern "C" int cg_emul_atom_cg_filter(const float* filter, float* filter_out, const float* feat, const float* y, const float* sample_weight, int n, int C,
                                      int H, int W, int num_iter, float filter_reg, int fletcher_reeves, int activation, float act_param) {
    CgParams P{};
    P.w_in = filter; P.w_out = filter_out; P.feat = feat; P.y = y; P.sample_weight = sample_weight;
    P.n = n; P.C = C; P.num_iter = num_iter; P.fletcher_reeves = fletcher_reeves ? 1 : 0; P.act = activation; P.act_param = act_param; P.reg = filter_reg;
    if (H == 18 && W == 18) return run_cg<18>(P);
    if (H == 22 && W == 22) return run_cg<22>(P);
    return 2;
}
