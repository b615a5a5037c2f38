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
#include "corr2.cuh"
#include <cstdlib>

namespace b200trk {

constexpr int CG_SPC_MAX = 8;

struct CgParams {
    const float* w_in; float* w_out; const float* feat; const float* y; const float* sample_weight;
    int n, C, passes, NCH, NG, num_iter, spc_max, fletcher_reeves, act;
    float act_param, reg;
    float* gpart; float* qpart; float* dots; unsigned* barrier;
};

__device__ __forceinline__ float cg_act(float s, int kind, float a) {
    if (kind == 1) return fmaxf(s, 0.f);
    if (kind == 2) return s > 0.f ? s : (expf(s) - 1.f);                   // F.elu, alpha = 1
    if (kind == 3) return s >= 0.f ? s : a * (expf(s / a) - 1.f);          // F.elu(F.leaky_relu(s, 1/a), a)
    return s;
}
__device__ __forceinline__ float cg_act_deriv(float s, int kind, float a) {
    if (kind == 1) return s > 0.f ? 1.f : 0.f;
    if (kind == 2) return s > 0.f ? 1.f : expf(s);
    if (kind == 3) return s >= 0.f ? 1.f : expf(s / a);
    return 1.f;
}

template <int FS, int NST>
__global__ void __launch_bounds__(Corr2<FS>::NCONS, 1)
atom_cg_kernel(CgParams P) {
    using K = Corr2<FS>;
    constexpr int NPOS = K::NPOS, OS = K::OS, NTH = K::NCONS, SLOTS = K::SLOTS, VS = K::VEC_STRIDE, PMAP = K::PMAP, PW = K::PW;
    extern __shared__ __align__(16) float smem[];
    float* stages = smem;
    float* red = stages + NST * K::ITEM_FLOATS;
    const int cchunk = P.passes * SLOTS;
    const int VF = cchunk * VS;
    float* wv = red + K::NT * SLOTS * K::RED_STRIDE;   // chunk slices (tap vectors, stride VS): w, r, p, x, r_prev, q
    float* rv = wv + VF;
    float* pv = rv + VF;
    float* xv = pv + VF;
    float* rpv = xv + VF;
    float* qv = rpv + VF;
    float* sT = qv + VF;                               // [spc][PMAP] tile-padded map fed to the transpose sweep
    float* part = sT + P.spc_max * PMAP;               // [<=NTH] float4 scratch of the group reduction
    float* sS = part + NTH * 4;                        // [spc][NPOS] A w, later A p
    float* sD = sS + P.spc_max * NPOS;                 // [spc][NPOS] sw * phi'(s)^2, zero outside the FS x FS window
    __shared__ float s_red[32];
    __shared__ float s_sw[CG_SPC_MAX];
    __shared__ float s_scal[4];

    const int tid = threadIdx.x;
    const int chunk = blockIdx.x % P.NCH, group = blockIdx.x / P.NCH;
    typename K::Ctx cx{P.feat, P.C, P.n, chunk * cchunk, P.passes, group, P.NG, 0};
    const int spc = cx.spc();
    unsigned epoch = 0;
    const size_t qstride = (size_t)P.NCH * NPOS;
    const int E = cchunk * 16;

    auto vidx = [&](int o) { return (o >> 4) * VS + (o & 15); };
    // own chunk of a gradient-type vector: sum of the NG group partials in a fixed order ((g mod GS) subsets, then subsets)
    auto reduce_groups = [&](float* dst, float scale, const float* addv, float addscale) {
        const int E4 = E / 4;
        const int GS = max(1, min(NTH / E4, 8));
        if (tid < E4 * GS) {
            const int e4 = tid % E4, gs = tid / E4;
            const float4* gp = reinterpret_cast<const float4*>(P.gpart + (size_t)chunk * E) + e4;
            const size_t gstride4 = (size_t)P.C * 4;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int g0 = gs; g0 < P.NG; g0 += GS * 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    v[u] = (g0 + u * GS < P.NG) ? __ldcg(gp + (size_t)(g0 + u * GS) * gstride4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 8; ++u) { s4.x += v[u].x; s4.y += v[u].y; s4.z += v[u].z; s4.w += v[u].w; }
            }
            reinterpret_cast<float4*>(part)[gs * E4 + e4] = s4;
        }
        __syncthreads();
        if (tid < E4) {
            float4 s4 = reinterpret_cast<float4*>(part)[tid];
            for (int gs = 1; gs < GS; ++gs) {
                const float4 v = reinterpret_cast<float4*>(part)[gs * E4 + tid];
                s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
            }
            const int vi = (tid >> 2) * VS + (tid & 3) * 4;
            const float4 a4 = *reinterpret_cast<const float4*>(addv + vi);
            s4.x = scale * (s4.x + addscale * a4.x); s4.y = scale * (s4.y + addscale * a4.y);
            s4.z = scale * (s4.z + addscale * a4.z); s4.w = scale * (s4.w + addscale * a4.w);
            *reinterpret_cast<float4*>(dst + vi) = s4;
        }
        __syncthreads();
    };
    // full-vector inner products: chunk-local partials, exchanged through `dots` when the filter spans several chunks
    int dot_slot = 0;
    auto dot2 = [&](const float* a0, const float* b0, const float* a1, const float* b1, float& d0, float& d1) {
        float l0 = 0.f, l1 = 0.f;
        for (int o = tid; o < E; o += NTH) {
            const int vi = vidx(o);
            l0 += a0[vi] * b0[vi];
            if (a1) l1 += a1[vi] * b1[vi];
        }
        l0 = block_sum(l0, s_red);
        l1 = block_sum(l1, s_red);
        if (P.NCH > 1) {
            float* slot = P.dots + (size_t)dot_slot * P.NCH * 2;
            if (group == 0 && tid == 0) { slot[chunk * 2] = l0; slot[chunk * 2 + 1] = l1; }
            grid_barrier(P.barrier, epoch);
            if (tid == 0) {
                s_scal[0] = ordered_sum_ldcg(slot, 2, P.NCH);
                s_scal[1] = ordered_sum_ldcg(slot + 1, 2, P.NCH);
            }
            __syncthreads();
            l0 = s_scal[0]; l1 = s_scal[1];
            __syncthreads();
        }
        ++dot_slot;
        d0 = l0; d1 = l1;
    };

    // ---- prologue -------------------------------------------------------------------------------------------
    K::zero_stages(stages, NST);
    for (int o = tid; o < P.spc_max * PMAP; o += NTH) sT[o] = 0.f;
    for (int o = tid; o < E; o += NTH) {
        wv[vidx(o)] = P.w_in[(size_t)chunk * E + o];
        xv[vidx(o)] = 0.f; pv[vidx(o)] = 0.f; rpv[vidx(o)] = 0.f;
    }
    if (tid < spc) s_sw[tid] = P.sample_weight[cx.sample(tid)];
    __syncthreads();

    // ---- s = A w --------------------------------------------------------------------------------------------
    K::template sweep_prologue<true, NST>(cx, stages);
    K::template sweep_apply<NST>(cx, stages, wv, P.qpart + (size_t)chunk * NPOS, qstride);
    K::template sweep_prologue<false, NST>(cx, stages);
    grid_barrier(P.barrier, epoch);
    for (int o = tid; o < spc * NPOS; o += NTH) {
        const int j = o / NPOS, pos = o - j * NPOS;
        const int yy = pos / OS, xx = pos - yy * OS;
        const float s = ordered_sum_ldcg(P.qpart + (size_t)cx.sample(j) * qstride + pos, NPOS, P.NCH);
        float r0 = 0.f, dd = 0.f;
        if (yy < FS && xx < FS) {                       // conv 'same': the last row / column of the even-filter map is cropped
            const float yl = P.y[((size_t)cx.sample(j) * FS + yy) * FS + xx];
            const float a = cg_act(s, P.act, P.act_param), d = cg_act_deriv(s, P.act, P.act_param);
            r0 = s_sw[j] * d * (a - yl);
            dd = s_sw[j] * d * d;
        }
        sD[o] = dd;
        sT[j * PMAP + yy * PW + xx] = r0;
    }
    __syncthreads();

    // ---- r = b = -(A^T r0 + reg w) ----------------------------------------------------------------------------
    K::template sweep_transpose<NST>(cx, stages, red, sT, P.gpart + ((size_t)group * P.C + chunk * cchunk) * 16);
    if (P.num_iter > 0) K::template sweep_prologue<true, NST>(cx, stages);
    grid_barrier(P.barrier, epoch);
    reduce_groups(rv, -1.f, wv, P.reg);

    float rho = 1.f;
    for (int ii = 0; ii < P.num_iter; ++ii) {
        const float rho1 = rho;
        float rho2 = 0.f;
        dot2(rv, rv, (ii > 0 && !P.fletcher_reeves) ? rpv : nullptr, rv, rho, rho2);
        if (rho == 0.f) break;                          // check_zero(rho): return the current iterate (uniform across the grid)
        float beta = 0.f;
        if (ii > 0) {
            beta = P.fletcher_reeves ? rho / rho1 : (rho - rho2) / rho1;
            beta = fmaxf(beta, 0.f);
        }
        for (int o = tid; o < E; o += NTH) { const int vi = vidx(o); pv[vi] = (ii == 0) ? rv[vi] : rv[vi] + pv[vi] * beta; }
        __syncthreads();

        // ---- q = A^T(D (A p)) + reg p -------------------------------------------------------------------------
        K::template sweep_apply<NST>(cx, stages, pv, P.qpart + (size_t)chunk * NPOS, qstride);
        K::template sweep_prologue<false, NST>(cx, stages);
        grid_barrier(P.barrier, epoch);
        for (int o = tid; o < spc * NPOS; o += NTH) {
            const int j = o / NPOS, pos = o - j * NPOS;
            const int yy = pos / OS, xx = pos - yy * OS;
            const float t = ordered_sum_ldcg(P.qpart + (size_t)cx.sample(j) * qstride + pos, NPOS, P.NCH);
            sT[j * PMAP + yy * PW + xx] = sD[o] * t;
        }
        __syncthreads();
        K::template sweep_transpose<NST>(cx, stages, red, sT, P.gpart + ((size_t)group * P.C + chunk * cchunk) * 16);
        if (ii + 1 < P.num_iter) K::template sweep_prologue<true, NST>(cx, stages);
        grid_barrier(P.barrier, epoch);
        reduce_groups(qv, 1.f, pv, P.reg);

        float pq, unused;
        dot2(pv, qv, nullptr, nullptr, pq, unused);
        const float alpha = rho / pq;
        for (int o = tid; o < E; o += NTH) {
            const int vi = vidx(o);
            if (!P.fletcher_reeves) rpv[vi] = rv[vi];
            xv[vi] += pv[vi] * alpha;
            if (ii < P.num_iter - 1) rv[vi] -= qv[vi] * alpha;
        }
        __syncthreads();
    }
    K::template wait_group<0>();
    if (group == 0)
        for (int o = tid; o < E; o += NTH) P.w_out[(size_t)chunk * E + o] = wv[vidx(o)] + xv[vidx(o)];
}

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
