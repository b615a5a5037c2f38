// The persistent cooperative kernel of sd_optimizer.cu (DiMP steepest-descent Gauss-Newton, PrDiMP steepest-descent Newton, DiMP-L2,
// GNSteepestDescent + LinearFilterHinge on fp32 CUDA-core sweeps; derivation and references in sd_optimizer.cu's header comment).  SIMT CUDA C
// on the sweeps of corr2.cuh (whose cp.async copies have a host form) in a header of its own so that the SAME source also compiles as host
// code under tests/cpu_emul/cuda_shim.h (tests/test_sd_kernels_cpu.py).  Included by sd_optimizer.cu only.
#pragma once
#include "corr2.cuh"
#include "sd_common.cuh"

#ifndef B200_DYN_SMEM_F16
#ifdef B200_CPU_EMUL
#define B200_DYN_SMEM_F16(name) float* name = reinterpret_cast<float*>(::cpu_emul::dyn_smem())
#else
#define B200_DYN_SMEM_F16(name) extern __shared__ __align__(16) float name[]
#endif
#endif

namespace b200trk {

template <int FS, int NST, int MODE>
__global__ void __launch_bounds__(Corr2<FS>::NCONS, 1)
sd_kernel(SdParams P) {
    using K = Corr2<FS>;
    constexpr int NPOS = K::NPOS, OS = K::OS, NTH = K::NCONS, SLOTS = K::SLOTS, VS = K::VEC_STRIDE, PMAP = K::PMAP, PW = K::PW;
    B200_DYN_SMEM_F16(smem);
    float* stages = smem;                                     // [NST][ITEM_FLOATS]  zero-bordered sample planes
    float* red = stages + NST * K::ITEM_FLOATS;               // [NT*16*17]          tile -> channel gradient reduction
    const int cchunk = P.passes * SLOTS;
    float* wv = red + K::NT * SLOTS * K::RED_STRIDE;          // [cchunk][VS] current filter taps of the chunk
    float* gv = wv + cchunk * VS;                             // [cchunk][VS] gradient taps of the chunk
    float* sT = gv + cchunk * VS;                             // [spc][PMAP] mapped residual, tile-padded (zero outside the map); 16-byte aligned
    float* sS = sT + P.spc_max * PMAP;                        // [spc][NPOS] scores
    float* sY = sS + P.spc_max * NPOS;                        // DiMP: label y        | PrDiMP: label density p
    float* sM = sY + P.spc_max * NPOS;                        // DiMP: target mask m  | PrDiMP: softmax(s)
    float* sV = sM + P.spc_max * NPOS;                        // DiMP: sqrt(sw)*v     | PrDiMP: unused
    float* sQ = sV + P.spc_max * NPOS;                        // q = A g
    __shared__ float s_red[32];
    __shared__ float s_sw[SD_SPC_MAX];
    __shared__ float s_scal[4];

    const int tid = threadIdx.x;
    const int chunk = blockIdx.x % P.NCH, group = blockIdx.x / P.NCH;
    typename K::Ctx cx{P.feat, P.C, P.n, chunk * cchunk, P.passes, group, P.NG, P.dbg_mode, P.feat_pitch};
    const int spc = cx.spc();
    unsigned epoch = 0;
    const size_t qstride = (size_t)P.NCH * NPOS;
    const float reg = P.reg_weight;

    SD_STAMP(0);
    // ---- prologue: zero staging planes + padded residual maps, load filter chunk, build per-sample label maps ----
    K::zero_stages(stages, NST);
    for (int o = tid; o < P.spc_max * PMAP; o += NTH) sT[o] = 0.f;
    for (int o = tid; o < cchunk * 16; o += NTH) wv[(o >> 4) * VS + (o & 15)] = P.w_in[(size_t)chunk * cchunk * 16 + o];
    if (tid < spc) {
        const int i = cx.sample(tid);
        s_sw[tid] = P.sample_weight ? P.sample_weight[i] : 1.0f / (float)P.n;
    }
    __syncthreads();
    K::template sweep_prologue<true, NST>(cx, stages);         // first planes are in flight while the label maps are built
    for (int j = 0; j < spc; ++j) {
        const int i = cx.sample(j);
        const float bx = P.bb[4 * i], by = P.bb[4 * i + 1], bw = P.bb[4 * i + 2], bh = P.bb[4 * i + 3];
        // centre (row, col) in score cells; even filter -> no half-cell offset (optimizer.py:112-113)
        const float crow = (by + bh / 2.f) * P.inv_feat_stride;
        const float ccol = (bx + bw / 2.f) * P.inv_feat_stride;
        if (MODE == 0) {
            const float sqsw = sqrtf(s_sw[j]);
            for (int pos = tid; pos < NPOS; pos += NTH) {
                const float d0 = (float)(pos / OS) - crow, d1 = (float)(pos % OS) - ccol;
                const float rho = sqrtf(d0 * d0 + d1 * d1) * P.inv_bin_disp;
                sY[j * NPOS + pos] = lut_lerp(P.label_lut, P.num_bins, rho);
                sM[j * NPOS + pos] = 1.f / (1.f + expf(-lut_lerp(P.mask_lut, P.num_bins, rho)));
                sV[j * NPOS + pos] = sqsw * lut_lerp(P.spatial_lut, P.num_bins, rho);
            }
        } else if (MODE == 3) {
            // LinearFilterHinge.forward (residual_modules.py:112-135): the label maps are an input of the residual module
            const float sqsw = sqrtf(s_sw[j]);
            for (int pos = tid; pos < NPOS; pos += NTH) {
                const float lab = P.label_in[(size_t)i * NPOS + pos];
                const float m = fminf(((lab > P.label_threshold) ? 1.f : 0.f) + P.act_leak, 1.f);
                sY[j * NPOS + pos] = m * lab;
                sM[j * NPOS + pos] = m;
                sV[j * NPOS + pos] = sqsw;
            }
        } else if (MODE == 2) {
            // DiMPL2SteepestDescentGN (optimizer.py:201-208,236-241): Gaussian label, hard hinge mask, weight sqrt(sw)
            const float c = -1.0f / (2.f * P.gauss_sigma * P.gauss_sigma);
            const float sqsw = sqrtf(s_sw[j]);
            for (int pos = tid; pos < NPOS; pos += NTH) {
                const float d0 = (float)(pos / OS) - crow, d1 = (float)(pos % OS) - ccol;
                const float gss = expf(c * d0 * d0) * expf(c * d1 * d1);
                const float m = (gss > P.label_threshold) ? 1.f : 0.f;
                sY[j * NPOS + pos] = gss * m;
                sM[j * NPOS + pos] = m;
                sV[j * NPOS + pos] = sqsw;
            }
        } else {
            const float c = -1.0f / (2.f * P.gauss_sigma * P.gauss_sigma);
            const float nrm = 1.f / (2.f * 3.14159265358979323846f * P.gauss_sigma * P.gauss_sigma);
            float loc = 0.f;
            for (int pos = tid; pos < NPOS; pos += NTH) {
                const float d0 = (float)(pos / OS) - crow, d1 = (float)(pos % OS) - ccol;
                float gss = (expf(c * d0 * d0) * nrm) * expf(c * d1 * d1);
                gss = (gss > P.label_threshold) ? gss : 0.f;
                sY[j * NPOS + pos] = gss;
                loc += gss;
            }
            const float tot = block_sum(loc, s_red);
            const float inv = P.normalize_label ? 1.f / (tot + 1e-8f) : 1.f;
            for (int pos = tid; pos < NPOS; pos += NTH)
                sY[j * NPOS + pos] = (1.f - P.label_shrink) *
                                     ((1.f - P.uni_weight) * (sY[j * NPOS + pos] * inv) + P.uni_weight / (float)NPOS);
        }
    }
    __syncthreads();

    SD_STAMP(1);
    // ---- s0 = A w0 -----------------------------------------------------------------------------------------
    K::template sweep_apply<NST>(cx, stages, wv, P.qpart + (size_t)chunk * NPOS, qstride);
    SD_STAMP(2);
    if (P.num_iter > 0) K::template sweep_prologue<false, NST>(cx, stages);   // planes of the first gradient sweep fly across the barrier
    grid_barrier(P.barrier, epoch);
    SD_STAMP(3);
    for (int o = tid; o < spc * NPOS; o += NTH) {
        const int j = o / NPOS, pos = o - j * NPOS;
        sS[o] = ordered_sum_ldcg(P.qpart + (size_t)cx.sample(j) * qstride + pos, NPOS, P.NCH);
    }
    __syncthreads();

    SD_STAMP(4);
    for (int it = 0; it <= P.num_iter; ++it) {
        const int tb = 8 + it * 10;
        SD_STAMP(tb + 0);
        // ---- residuals from the current scores (also the loss terms of iterate `it`) -----------------------
        float lloc = 0.f;
        if (MODE != 1) {
            for (int o = tid; o < spc * NPOS; o += NTH) {
                const int j = o / NPOS, pos = o - j * NPOS;
                const float s = sS[o], m = sM[o], vh = sV[o];
                float act, dact;
                if (MODE == 3 && P.act_kind == 1) {      // BentIdentPar (activation.py:53-74)
                    const float rt = sqrtf(s * s + 4.f * P.act_b * P.act_b);
                    act = 0.5f * (1.f - m) * (rt - 2.f * P.act_b) + 0.5f * (1.f + m) * s;
                    dact = 0.5f * (1.f - m) * (s / rt) + 0.5f * (1.f + m);
                } else if (MODE == 0 || MODE == 3) {
                    act = 0.5f * (1.f - m) * fabsf(s) + 0.5f * (1.f + m) * s;
                    const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
                    dact = 0.5f * (1.f - m) * sg + 0.5f * (1.f + m);
                } else {            // optimizer.py:258-259: mask*s + (1-mask)*relu(s), derivative mask + (1-mask)*(s > 0)
                    act = m * s + (1.f - m) * fmaxf(s, 0.f);
                    dact = m + (1.f - m) * ((s > 0.f) ? 1.f : 0.f);
                }
                const float r = vh * (act - sY[o]);
                lloc += r * r;
                sT[j * PMAP + (pos / OS) * PW + (pos % OS)] = dact * (vh * r);
            }
        } else {
            for (int j = 0; j < spc; ++j) {
                // softmax over the map with one extra constant logit (activation.py:7-16)
                float mx = P.has_softmax_reg ? P.softmax_reg : -INFINITY;
                for (int pos = tid; pos < NPOS; pos += NTH) mx = fmaxf(mx, sS[j * NPOS + pos]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                __syncthreads();
                if ((tid & 31) == 0) s_red[tid >> 5] = mx;
                __syncthreads();
                mx = s_red[0];
                for (int wq = 1; wq < (NTH + 31) / 32; ++wq) mx = fmaxf(mx, s_red[wq]);
                float se = 0.f, ps = 0.f;
                for (int pos = tid; pos < NPOS; pos += NTH) {
                    const float e = expf(sS[j * NPOS + pos] - mx);
                    sM[j * NPOS + pos] = e;
                    se += e;
                    ps += sY[j * NPOS + pos] * sS[j * NPOS + pos];
                }
                se = block_sum(se, s_red);
                ps = block_sum(ps, s_red);
                const float den = se + (P.has_softmax_reg ? expf(P.softmax_reg - mx) : 0.f);
                const float inv = 1.f / den;
                const float sw = s_sw[j];
                for (int pos = tid; pos < NPOS; pos += NTH) {
                    const float sm = sM[j * NPOS + pos] * inv;
                    sM[j * NPOS + pos] = sm;
                    sT[j * PMAP + (pos / OS) * PW + (pos % OS)] = sw * (sm - sY[j * NPOS + pos]);
                }
                // loss_i = sw * (log(sum exp(s) + exp(reg)) - sum p*s)  (optimizer.py:393-396)
                if (tid == 0) lloc += sw * ((logf(den) + mx) - ps);
            }
        }
        if (P.losses_out) {
            const float lr = block_sum(lloc, s_red);
            float lw = 0.f;
            for (int o = tid; o < cchunk * 16; o += NTH) { const float w = wv[(o >> 4) * VS + (o & 15)]; lw += w * w; }
            lw = block_sum(lw, s_red);
            if (tid == 0) {
                if (chunk == 0) P.lossr[it * P.NG + group] = lr;
                if (group == 0) P.lossw[it * P.NCH + chunk] = lw;
            }
        }
        if (it == P.num_iter) break;
        __syncthreads();

        SD_STAMP(tb + 1);
        // ---- phase 1: partial gradient of the chunk over the CTA's samples -----------------------------------
        K::template sweep_transpose<NST>(cx, stages, red, sT, P.gpart + ((size_t)group * P.C + chunk * cchunk) * 16);
        SD_STAMP(tb + 2);
        K::template sweep_prologue<true, NST>(cx, stages);
        grid_barrier(P.barrier, epoch);
        SD_STAMP(tb + 3);

        // ---- phase 2: g = sum_groups gpart + reg*w ; ||g_chunk||^2 ; partial q = A g ---------------------------
        float gl = 0.f;
        for (int o4 = tid; o4 < cchunk * 4; o4 += NTH) {       // 4 consecutive taps per thread, all group loads in flight at once
            const float4* gp = reinterpret_cast<const float4*>(P.gpart + (size_t)chunk * cchunk * 16) + o4;
            const size_t gstride4 = (size_t)P.C * 4;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int g0 = 0; g0 < P.NG; g0 += 10) {
                float4 v[10];
#pragma unroll
                for (int u = 0; u < 10; ++u)
                    v[u] = (g0 + u < P.NG) ? __ldcg(gp + (size_t)(g0 + u) * gstride4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 10; ++u) { s4.x += v[u].x; s4.y += v[u].y; s4.z += v[u].z; s4.w += v[u].w; }
            }
            const int vi = (o4 >> 2) * VS + (o4 & 3) * 4;
            const float4 w4 = *reinterpret_cast<const float4*>(wv + vi);
            s4.x += reg * w4.x; s4.y += reg * w4.y; s4.z += reg * w4.z; s4.w += reg * w4.w;
            *reinterpret_cast<float4*>(gv + vi) = s4;
            gl += s4.x * s4.x + s4.y * s4.y + s4.z * s4.z + s4.w * s4.w;
        }
        gl = block_sum(gl, s_red);
        if (group == 0 && tid == 0) P.gnorm[chunk] = gl;
        __syncthreads();
        SD_STAMP(tb + 4);
        K::template sweep_apply<NST>(cx, stages, gv, P.qpart + (size_t)chunk * NPOS, qstride);
        SD_STAMP(tb + 5);
        if (it + 1 < P.num_iter) K::template sweep_prologue<false, NST>(cx, stages);
        grid_barrier(P.barrier, epoch);
        SD_STAMP(tb + 6);

        // ---- phase 3: q_i over all chunks, curvature term --------------------------------------------------------
        float hl = 0.f;
        if (MODE != 1) {
            for (int o = tid; o < spc * NPOS; o += NTH) {
                const int j = o / NPOS, pos = o - j * NPOS;
                const float q = ordered_sum_ldcg(P.qpart + (size_t)cx.sample(j) * qstride + pos, NPOS, P.NCH);
                sQ[o] = q;
                const float s = sS[o], m = sM[o];
                float dact;
                if (MODE == 3 && P.act_kind == 1) {
                    dact = 0.5f * (1.f - m) * (s / sqrtf(s * s + 4.f * P.act_b * P.act_b)) + 0.5f * (1.f + m);
                } else if (MODE == 0 || MODE == 3) {
                    const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
                    dact = 0.5f * (1.f - m) * sg + 0.5f * (1.f + m);
                } else {
                    dact = m + (1.f - m) * ((s > 0.f) ? 1.f : 0.f);
                }
                const float h = sV[o] * (dact * q);
                hl += h * h;
            }
            hl = block_sum(hl, s_red);
        } else {
            for (int j = 0; j < spc; ++j) {
                float dotl = 0.f;
                for (int pos = tid; pos < NPOS; pos += NTH) {
                    const float q = ordered_sum_ldcg(P.qpart + (size_t)cx.sample(j) * qstride + pos, NPOS, P.NCH);
                    sQ[j * NPOS + pos] = q;
                    dotl += sM[j * NPOS + pos] * q;
                }
                const float dot = block_sum(dotl, s_red);
                float gh = 0.f;
                for (int pos = tid; pos < NPOS; pos += NTH) {
                    const float q = sQ[j * NPOS + pos], sm = sM[j * NPOS + pos];
                    gh += q * (sm * q - sm * dot);
                }
                gh = block_sum(gh, s_red);
                hl += s_sw[j] * fmaxf(gh, 0.f);   // identical on all threads
            }
        }
        if (chunk == 0 && tid == 0) P.hpart[group] = hl;
        SD_STAMP(tb + 7);
        grid_barrier(P.barrier, epoch);
        SD_STAMP(tb + 8);

        // ---- step length and update --------------------------------------------------------------------------------
        if (tid == 0) {
            const float gn = ordered_sum_ldcg(P.gnorm, 1, P.NCH);
            const float hn = ordered_sum_ldcg(P.hpart, 1, P.NG);
            const float den = fmaxf(hn + (reg + P.alpha_eps) * gn, 1e-8f);
            s_scal[0] = P.step_length * (gn / den);
        }
        __syncthreads();
        const float sa = s_scal[0];
        for (int o = tid; o < spc * NPOS; o += NTH) sS[o] -= sa * sQ[o];
        for (int o = tid; o < cchunk * 16; o += NTH) {
            const int vi = (o >> 4) * VS + (o & 15);
            const float w = wv[vi] - sa * gv[vi];
            wv[vi] = w;
            if (group == 0 && P.iterates_out)
                P.iterates_out[((size_t)(it + 1) * P.C + chunk * cchunk) * 16 + o] = w;
        }
        __syncthreads();
    }

    // ---- epilogue -------------------------------------------------------------------------------------------------------
    K::template wait_group<0>();
    if (group == 0)
        for (int o = tid; o < cchunk * 16; o += NTH) P.w_out[(size_t)chunk * cchunk * 16 + o] = wv[(o >> 4) * VS + (o & 15)];
    if (P.losses_out) {
        grid_barrier(P.barrier, epoch);
        if (blockIdx.x == 0 && tid <= P.num_iter) {
            const float l = ordered_sum_ldcg(P.lossr + tid * P.NG, 1, P.NG);
            const float lw = ordered_sum_ldcg(P.lossw + tid * P.NCH, 1, P.NCH);
            P.losses_out[tid] = (MODE == 3) ? (l + reg * lw) * P.loss_scale : l + reg * lw;
        }
    }
}

}  // namespace b200trk
