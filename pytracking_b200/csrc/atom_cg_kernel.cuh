// The persistent cooperative kernel of atom_cg.cu (ATOM's per-frame ConjugateGradient.run on ConvProblem; derivation and references in
// atom_cg.cu's header comment).  SIMT CUDA C on the sweeps of corr2.cuh (whose cp.async copies have a host form) in a header of its own so that
// the SAME source also compiles as host code under tests/cpu_emul/cuda_shim.h (tests/test_sd_kernels_cpu.py).  Included by atom_cg.cu only.
#pragma once
#include "corr2.cuh"

#ifndef B200_DYN_SMEM_F16
#ifdef B200_CPU_EMUL
#define B200_DYN_SMEM_F16(name) float* name = reinterpret_cast<float*>(::cpu_emul::dyn_smem())
#else
#define B200_DYN_SMEM_F16(name) extern __shared__ __align__(16) float name[]
#endif
#endif

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
    B200_DYN_SMEM_F16(smem);
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

}  // namespace b200trk
