// ECO optimisers -- device build and C ABI of the kernels in eco_cg_kernel.cuh (online FilterOptim.run) and
// eco_joint_kernel.cuh (first-frame GaussNewtonCG on FactorizedConvProblem); design notes in the headers.
//   reference: pytracking/tracker/eco/optim.py:8-208, pytracking/libs/optimization.py:72-163, 328-421.
#include "common.cuh"
#include "launch.cuh"

#ifndef B200_CPU_EMUL      // the CPU test tier has its own (pthread) barrier under the same name
namespace b200trk {

// grid_barrier (common.cuh) with a bounded wait: if the other CTAs never arrive (which a cooperative launch rules out) the poll gives
// up after ~2^20 round trips, raises word 16 of the counter block and every later barrier of this CTA falls through, so that a defect
// can only ever produce a wrong result, never a kernel that does not terminate.
__device__ __forceinline__ void eco_grid_barrier(unsigned* counter, unsigned& epoch, unsigned& dead) {
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        const unsigned target = epoch * gridDim.x;
        unsigned spins = 0;
        while (!dead && ld_acquire_u32(counter) < target) {
            if (++spins > (1u << 20)) { dead = 1u; counter[16] = 1u; }
        }
    }
    __syncthreads();
}

}  // namespace b200trk
#endif

#include "eco_cg_kernel.cuh"
#include "eco_joint_kernel.cuh"

namespace b200trk {

template <int G, int CPL>
static int launch_eco(const EcoPlan& pl, EcoParams& P, cudaStream_t st) {
    if (int e = b200_launch_cooperative(eco_cg_kernel<G, CPL>, pl.grid, pl.block, pl.smem_bytes, st, P)) return e;
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

}  // namespace b200trk

using namespace b200trk;

extern "C" int b200trk_eco_filter_cg(float* filter, const float* samples, const float* yf, const float* sample_weights,
                                     const float* reg_filter, int reg_h, int reg_w, float* sample_energy, int has_energy,
                                     const float* new_xf, float* p, float* r_prev, float* rho, int has_state,
                                     int H, int Wh, int N, int C, int num_iter, int fletcher_reeves, int standard_alpha,
                                     float direction_forget_factor, float precond_learning_rate, float precond_data_param,
                                     float precond_reg_param, b200trk_stream_t stream) {
    B200_REQUIRE(filter && samples && yf && sample_weights && reg_filter && sample_energy && p && rho,
                 "eco_filter_cg: null pointer");
    B200_REQUIRE(fletcher_reeves || r_prev, "eco_filter_cg: the Polak-Ribiere formula needs the r_prev buffer");
    B200_REQUIRE(has_energy || new_xf, "eco_filter_cg: no sample energy yet and no new sample to initialise it from");
    B200_REQUIRE(H > 0 && Wh > 0 && N > 0 && N <= 4096, "eco_filter_cg: H=%d Wh=%d N=%d", H, Wh, N);
    B200_REQUIRE(C == 16 || C == 32 || C == 64 || C == 128, "eco_filter_cg: C=%d (compressed_dim) must be 16, 32, 64 or 128", C);
    B200_REQUIRE(reg_h >= 1 && reg_w >= 1 && reg_h <= 8 && reg_w <= 8, "eco_filter_cg: regularisation filter %dx%d (at most 8x8)", reg_h, reg_w);
    // optim.py:178-183 rebuilds reg_w - 1 negative-kx columns from the half spectrum and pads reg_h - 1 rows: both must exist
    B200_REQUIRE(reg_w <= Wh && reg_h <= H, "eco_filter_cg: regularisation filter %dx%d larger than the %dx%d half spectrum", reg_h, reg_w, H, Wh);
    B200_REQUIRE(num_iter >= 0 && num_iter <= 1024, "eco_filter_cg: num_iter=%d", num_iter);
    B200_REQUIRE(((uintptr_t)samples & 15) == 0 && ((uintptr_t)filter & 7) == 0 && ((uintptr_t)p & 7) == 0 && (!r_prev || ((uintptr_t)r_prev & 7) == 0) &&
                 (!new_xf || ((uintptr_t)new_xf & 7) == 0), "eco_filter_cg: samples must be 16-byte aligned, complex tensors 8-byte aligned");
    if (num_iter == 0) return 0;                                         // optim.py:141-142: nothing happens, not even the energy update
    cudaStream_t st = (cudaStream_t)stream;
    EcoPlan pl = eco_plan(H, Wh, N, C, num_iter, device_sm_count(), 256);
    B200_REQUIRE(pl.smem_bytes <= 227 * 1024, "eco_filter_cg: %zu bytes of shared memory", pl.smem_bytes);
    char* ws = (char*)workspace(pl.ws_bytes, 6);
    if (!ws) return 3;
    EcoParams P{};
    P.hf = filter; P.samples = samples; P.yf = yf; P.sw = sample_weights; P.reg_filter = reg_filter;
    P.sample_energy = sample_energy; P.new_xf = new_xf; P.p_state = p; P.r_prev_state = r_prev; P.rho_state = rho;
    P.has_state = (has_state && direction_forget_factor != 0.f) ? 1 : 0;   // optimization.py:82-85: a zero factor resets the state
    P.has_energy = has_energy ? 1 : 0;
    P.H = H; P.Wh = Wh; P.N = N; P.C = C; P.rh = reg_h; P.rw = reg_w; P.num_iter = num_iter;
    P.fletcher_reeves = fletcher_reeves ? 1 : 0; P.standard_alpha = standard_alpha ? 1 : 0;
    P.dff = direction_forget_factor; P.lr = precond_learning_rate; P.pdp = precond_data_param; P.prp = precond_reg_param;
    P.barrier = (unsigned*)ws;
    P.xw = (float2*)(ws + pl.off_xw); P.pw = (float2*)(ws + pl.off_pw); P.resw = (float2*)(ws + pl.off_resw);
    P.rpw = (float2*)(ws + pl.off_rpw); P.qw = (float2*)(ws + pl.off_qw); P.dM = (float*)(ws + pl.off_dM);
    P.dots = (float*)(ws + pl.off_dots);
    P.GPP = pl.GPP; P.res_slabs = pl.res_slabs; P.npx_max = pl.npx_max;
    B200_CHECK_CUDA(cudaMemsetAsync(P.barrier, 0, 256, st));
    if (pl.G == 16) return launch_eco<16, 1>(pl, P, st);
    if (pl.CPL == 1) return launch_eco<32, 1>(pl, P, st);
    if (pl.CPL == 2) return launch_eco<32, 2>(pl, P, st);
    return launch_eco<32, 4>(pl, P, st);
}

extern "C" int b200trk_eco_joint_gn(float* filter, float* proj, const float* samples, const float* yf, const float* sample_weights_sqrt,
                                    const float* reg_filter, int reg_h, int reg_w, const float* diag_M_filter, float diag_M_proj,
                                    float projection_reg, int H, int Wh, int N, int Cin, int C, int num_cg_iter, int num_gn_iter,
                                    b200trk_stream_t stream) {
    B200_REQUIRE(filter && proj && samples && yf && sample_weights_sqrt && reg_filter && diag_M_filter, "eco_joint_gn: null pointer");
    B200_REQUIRE(H > 0 && Wh > 0 && N > 0 && N <= 1024 && Cin > 0 && C > 0 && C <= 512 && Cin <= 4096, "eco_joint_gn: H=%d Wh=%d N=%d Cin=%d C=%d", H, Wh, N, Cin, C);
    B200_REQUIRE(reg_h >= 1 && reg_w >= 1 && reg_h <= 8 && reg_w <= 8 && reg_w <= Wh && reg_h <= H,
                 "eco_joint_gn: regularisation filter %dx%d (at most 8x8 and not larger than the %dx%d half spectrum)", reg_h, reg_w, H, Wh);
    B200_REQUIRE(num_cg_iter >= 0 && num_gn_iter >= 0 && (long long)num_cg_iter * num_gn_iter <= 100000, "eco_joint_gn: %d x %d iterations", num_cg_iter, num_gn_iter);
    B200_REQUIRE(diag_M_proj > 0.f, "eco_joint_gn: diag_M_proj=%g", diag_M_proj);
    B200_REQUIRE(((uintptr_t)samples & 7) == 0 && ((uintptr_t)filter & 7) == 0, "eco_joint_gn: complex tensors must be 8-byte aligned");
    if (num_gn_iter == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const EcoJointPlan pl = eco_joint_plan(H, Wh, N, Cin, C, num_cg_iter, num_gn_iter, device_sm_count(), 256);
    B200_REQUIRE(pl.smem_bytes <= 227 * 1024, "eco_joint_gn: %zu bytes of shared memory (N=%d, Cin=%d, C=%d)", pl.smem_bytes, N, Cin, C);
    char* ws = (char*)workspace(pl.ws_bytes, 7);
    if (!ws) return 3;
    const size_t field = (size_t)H * Wh * C * 2 * sizeof(float), nelem = (size_t)Cin * C;
    EcoJointParams P{};
    P.hf = filter; P.proj = proj; P.samples = samples; P.yf = yf; P.sw_sqrt = sample_weights_sqrt; P.reg_filter = reg_filter;
    P.dMh_in = diag_M_filter; P.dMP = diag_M_proj; P.lambda = projection_reg;
    P.H = H; P.Wh = Wh; P.N = N; P.Cin = Cin; P.C = C; P.rh = reg_h; P.rw = reg_w; P.num_cg = num_cg_iter; P.num_gn = num_gn_iter;
    P.barrier = (unsigned*)ws;
    P.h0w = (float2*)(ws + pl.off_fields); P.phw = (float2*)(ws + pl.off_fields + field); P.xhw = (float2*)(ws + pl.off_fields + 2 * field);
    P.rhw = (float2*)(ws + pl.off_fields + 3 * field); P.qhw = (float2*)(ws + pl.off_fields + 4 * field);
    P.dMh = (float*)(ws + pl.off_dMh); P.c0w = (float2*)(ws + pl.off_c0); P.wv = (float2*)(ws + pl.off_wv);
    P.pP = (float*)(ws + pl.off_P); P.xP = P.pP + nelem; P.rP = P.xP + nelem; P.qP = P.rP + nelem;
    P.dots = (float*)(ws + pl.off_dots);
    P.res_slabs = pl.res_slabs; P.npx_max = pl.npx_max; P.EPB = pl.EPB; P.SPL = pl.SPL; P.stage_pm = pl.stage_pm; P.wide = pl.wide;
    B200_CHECK_CUDA(cudaMemsetAsync(P.barrier, 0, 256, st));
    if (int e = b200_launch_cooperative(eco_joint_kernel, pl.grid, pl.block, pl.smem_bytes, st, P)) return e;
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}
