// Stage 3: the online filter optimisers as ONE persistent cooperative kernel per call.
//   MODE 0: DiMPSteepestDescentGN      (ltr/models/target_classifier/optimizer.py:85-170)
//   MODE 1: PrDiMPSteepestDescentNewton (ltr/models/target_classifier/optimizer.py:355-439)
//
// Algorithm with one algebraic restructuring: the score maps are carried across
// iterations by linearity, s_{k+1} = A w_{k+1} = s_k - step*alpha_k * (A g_k), so an iteration costs two
// sweeps over the sample memory (A^T r and A g) instead of the reference's three.
//
// Decomposition: CTA = (channel chunk, sample group); grid = NCH x NG <= #SMs, launched cooperatively. The sample
// planes are streamed by the cp.async multistage sweeps of corr2.cuh (next sweep's first planes are prefetched
// across each grid barrier).
// Residual maps, labels, the chunk's filter taps and gradient stay in shared memory for the whole call;
// the sample memory is streamed from L2 (it is re-read 2x per iteration; 33 MB at n=50 is L2 resident).
// Cross-CTA exchange per iteration (all via L2, fixed summation order => bitwise deterministic):
//   gpart [NG][C*16]   partial gradients     -> barrier 1 -> each CTA sums its own chunk over the groups
//   qpart [n][NCH][NPOS] partial A g maps    -> barrier 2 -> each CTA sums its own samples over the chunks
//   hpart [NG], gnorm [NCH] scalars          -> barrier 3 -> step length alpha
#include "corr2.cuh"
#include "sd_common.cuh"
#include <cstdlib>

namespace b200trk {

template <int FS, int NST, int MODE>
__global__ void __launch_bounds__(Corr2<FS>::NCONS, 1)
sd_kernel(SdParams P) {
    using K = Corr2<FS>;
    constexpr int NPOS = K::NPOS, OS = K::OS, NTH = K::NCONS, SLOTS = K::SLOTS, VS = K::VEC_STRIDE, PMAP = K::PMAP, PW = K::PW;
    extern __shared__ __align__(16) float smem[];
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

template <int FS, int NST, int MODE>
static int launch_sd_nst(const SdParams& P, size_t smem, cudaStream_t st) {
    using K = Corr2<FS>;
    auto kern = sd_kernel<FS, NST, MODE>;
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&P};
    B200_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(P.NCH * P.NG), dim3(K::NCONS), args, smem, st));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

template <int FS, int MODE>
static int launch_sd(SdParams P, cudaStream_t st) {
    using K = Corr2<FS>;
    constexpr int SLOTS = K::SLOTS;
    {
        // tensor-core sweeps first (sd_tc.cu); shapes it does not claim fall through to the CUDA-core kernel below
        int handled = 0;
        if (int e = launch_sd_tc<FS, MODE>(P, st, &handled)) return e;
        if (handled) return 0;
    }
    // Decomposition: passes x 16 channels per CTA (NCH chunks) x NG sample groups. The sweep time is set by the busiest CTA,
    // ceil(n / NG) * passes items, so every admissible `passes` is scored (n = 50 on 148 SMs: 4 passes -> 8 items, 2 -> 6).
    const int sms = device_sm_count();
    const size_t limit = 227 * 1024 - 512;
    const size_t item = (size_t)K::ITEM_FLOATS * sizeof(float);
    const int forced = [] { const char* v = getenv("B200TRK_SD_PASSES"); return v ? atoi(v) : 0; }();
    int passes = 0, NCH = 0, NG = 0, spc = 0, best_cost = 1 << 30;
    for (int p = 64; p >= 1; p >>= 1) {
        if (P.C % (SLOTS * p) != 0) continue;
        if (forced && p != forced) continue;
        const int nch = P.C / (SLOTS * p);
        if (nch > sms) continue;
        int ng = sms / nch; if (ng > P.n) ng = P.n; if (ng < 1) ng = 1;
        const int sp = (P.n + ng - 1) / ng;
        if (sp > SD_SPC_MAX) continue;
        const size_t fx = (size_t)(K::NT * SLOTS * K::RED_STRIDE + 2 * p * SLOTS * K::VEC_STRIDE + sp * (5 * K::NPOS + K::PMAP)) * sizeof(float);
        if (fx + 2 * item > limit) continue;
        int cost = sp * p * 4;
        if (fx + 3 * item > limit) cost += cost / 2;        // a 2-stage pipeline exposes the copy latency
        if (p > 4) cost += 1;                               // prefer <= 4 passes at equal balance (fewer partial-gradient rows)
        if (cost < best_cost) { best_cost = cost; passes = p; NCH = nch; NG = ng; spc = sp; }
    }
    B200_REQUIRE(passes > 0, "sd optimizer: no decomposition for C=%d, n=%d (C must be a multiple of %d, at most %d samples per CTA)",
                 P.C, P.n, SLOTS, SD_SPC_MAX);
    B200_REQUIRE(P.num_iter + 1 <= K::NCONS, "sd optimizer: num_iter=%d too large", P.num_iter);
    P.passes = passes; P.NCH = NCH; P.NG = NG; P.spc_max = spc;
    { const char* v = getenv("B200TRK_SD_DBG"); P.dbg_mode = v ? atoi(v) : 0; }
    P.trace = getenv("B200TRK_SD_TRACE") ? (unsigned long long*)workspace(1024, 3) : nullptr;

    const size_t n_gpart = (size_t)NG * P.C * 16, n_qpart = (size_t)P.n * NCH * K::NPOS;
    const size_t n_loss = (size_t)(P.num_iter + 1) * (NG + NCH);
    const size_t total = (n_gpart + n_qpart + NG + NCH + n_loss + 64) * sizeof(float) + 1024;
    char* ws = (char*)workspace(total, 2);
    if (!ws) return 3;
    P.barrier = (unsigned*)ws;
    float* f = (float*)(ws + 1024);
    P.gpart = f; f += n_gpart;
    P.qpart = f; f += n_qpart;
    P.hpart = f; f += NG;
    P.gnorm = f; f += NCH;
    P.lossr = f; f += (size_t)(P.num_iter + 1) * NG;
    P.lossw = f;
    B200_CHECK_CUDA(cudaMemsetAsync(P.barrier, 0, 1024, st));

    const int cchunk = passes * SLOTS;
    const size_t fixed = (size_t)(K::NT * SLOTS * K::RED_STRIDE + 2 * cchunk * K::VEC_STRIDE + spc * (5 * K::NPOS + K::PMAP)) * sizeof(float);
    if (fixed + 4 * item <= limit) return launch_sd_nst<FS, 4, MODE>(P, fixed + 4 * item, st);
    if (fixed + 3 * item <= limit) return launch_sd_nst<FS, 3, MODE>(P, fixed + 3 * item, st);
    B200_REQUIRE(fixed + 2 * item <= limit, "sd optimizer: %d samples per CTA do not fit in shared memory", spc);
    return launch_sd_nst<FS, 2, MODE>(P, fixed + 2 * item, st);
}

static int check_common(const char* who, const float* w, float* wo, const float* feat, const float* bb, int n, int C,
                        int H, int W, int k, int num_iter) {
    B200_REQUIRE(w && wo && feat && bb, "%s: null pointer", who);
    B200_REQUIRE(n > 0 && C > 0, "%s: empty sample memory (n=%d, C=%d)", who, n, C);
    B200_REQUIRE(k == 4, "%s: filter size %d not supported by the CUDA path (only 4)", who, k);
    B200_REQUIRE(H == W && (H == 18 || H == 22), "%s: feature size %dx%d not supported (18x18, 22x22)", who, H, W);
    B200_REQUIRE(num_iter >= 0, "%s: num_iter=%d", who, num_iter);
    return 0;
}

}  // namespace b200trk

using namespace b200trk;

int b200trk::dimp_sd_gn_pitched(const float* weights, float* weights_out, const float* feat, int feat_pitch, const float* bb,
                                const float* sample_weight, int n, int C, int H, int W, int k, int num_iter,
                                const float* label_lut, const float* mask_lut, const float* spatial_lut,
                                int num_bins, float bin_displacement, float feat_stride, float step_length,
                                float reg_weight, float alpha_eps, float* iterates_out, float* losses_out, cudaStream_t st) {
    if (int e = check_common("dimp_sd_gn", weights, weights_out, feat, bb, n, C, H, W, k, num_iter)) return e;
    B200_REQUIRE(label_lut && mask_lut && spatial_lut && num_bins >= 2, "dimp_sd_gn: LUTs missing");
    B200_REQUIRE(bin_displacement > 0.f && feat_stride > 0.f, "dimp_sd_gn: bad bin_displacement / feat_stride");
    B200_REQUIRE(feat_pitch == 0 || (feat_pitch >= H * W && feat_pitch % 2 == 0), "dimp_sd_gn: bad channel-plane pitch %d", feat_pitch);
    if (iterates_out)
        B200_CHECK_CUDA(cudaMemcpyAsync(iterates_out, weights, (size_t)C * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SdParams P{};
    P.w_in = weights; P.w_out = weights_out; P.feat = feat; P.bb = bb; P.sample_weight = sample_weight;
    P.feat_pitch = (feat_pitch == H * W) ? 0 : feat_pitch;
    P.n = n; P.C = C; P.num_iter = num_iter;
    P.label_lut = label_lut; P.mask_lut = mask_lut; P.spatial_lut = spatial_lut; P.num_bins = num_bins;
    P.inv_bin_disp = 1.0f / bin_displacement; P.inv_feat_stride = 1.0f / feat_stride;
    P.step_length = step_length; P.reg_weight = reg_weight; P.alpha_eps = alpha_eps;
    P.iterates_out = iterates_out; P.losses_out = losses_out;
    if (H == 18) return launch_sd<18, 0>(P, st);
    return launch_sd<22, 0>(P, st);
}

extern "C" int b200trk_dimp_sd_gn(const float* weights, float* weights_out, const float* feat, const float* bb,
                                  const float* sample_weight, int n, int C, int H, int W, int k, int num_iter,
                                  const float* label_lut, const float* mask_lut, const float* spatial_lut,
                                  int num_bins, float bin_displacement, float feat_stride, float step_length,
                                  float reg_weight, float alpha_eps, float* iterates_out, float* losses_out,
                                  b200trk_stream_t stream) {
    return dimp_sd_gn_pitched(weights, weights_out, feat, 0, bb, sample_weight, n, C, H, W, k, num_iter, label_lut, mask_lut, spatial_lut,
                              num_bins, bin_displacement, feat_stride, step_length, reg_weight, alpha_eps, iterates_out, losses_out,
                              (cudaStream_t)stream);
}

extern "C" int b200trk_prdimp_sd_newton(const float* weights, float* weights_out, const float* feat, const float* bb,
                                        const float* sample_weight, int n, int C, int H, int W, int k, int num_iter,
                                        float gauss_sigma, float feat_stride, float step_length, float reg_weight,
                                        float alpha_eps, int has_softmax_reg, float softmax_reg, float label_threshold,
                                        int normalize_label, float label_shrink, float uni_weight,
                                        float* iterates_out, float* losses_out, b200trk_stream_t stream) {
    if (int e = check_common("prdimp_sd_newton", weights, weights_out, feat, bb, n, C, H, W, k, num_iter)) return e;
    B200_REQUIRE(gauss_sigma > 0.f, "prdimp_sd_newton: gauss_sigma must be > 0 (one-hot labels not implemented)");
    cudaStream_t st = (cudaStream_t)stream;
    if (iterates_out)
        B200_CHECK_CUDA(cudaMemcpyAsync(iterates_out, weights, (size_t)C * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SdParams P{};
    P.w_in = weights; P.w_out = weights_out; P.feat = feat; P.bb = bb; P.sample_weight = sample_weight;
    P.n = n; P.C = C; P.num_iter = num_iter;
    P.gauss_sigma = gauss_sigma; P.has_softmax_reg = has_softmax_reg; P.softmax_reg = softmax_reg;
    P.label_threshold = label_threshold; P.normalize_label = normalize_label; P.label_shrink = label_shrink;
    P.uni_weight = uni_weight;
    P.inv_feat_stride = 1.0f / feat_stride;
    P.step_length = step_length; P.reg_weight = reg_weight; P.alpha_eps = alpha_eps;
    P.iterates_out = iterates_out; P.losses_out = losses_out;
    if (H == 18) return launch_sd<18, 1>(P, st);
    return launch_sd<22, 1>(P, st);
}

extern "C" int b200trk_dimp_l2_sd_gn(const float* weights, float* weights_out, const float* feat, const float* bb,
                                     const float* sample_weight, int n, int C, int H, int W, int k, int num_iter,
                                     float gauss_sigma, float hinge_threshold, float feat_stride, float step_length,
                                     float reg_weight, float alpha_eps, float* iterates_out, float* losses_out,
                                     b200trk_stream_t stream) {
    if (int e = check_common("dimp_l2_sd_gn", weights, weights_out, feat, bb, n, C, H, W, k, num_iter)) return e;
    B200_REQUIRE(gauss_sigma > 0.f && feat_stride > 0.f, "dimp_l2_sd_gn: gauss_sigma and feat_stride must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if (iterates_out)
        B200_CHECK_CUDA(cudaMemcpyAsync(iterates_out, weights, (size_t)C * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SdParams P{};
    P.w_in = weights; P.w_out = weights_out; P.feat = feat; P.bb = bb; P.sample_weight = sample_weight;
    P.n = n; P.C = C; P.num_iter = num_iter;
    P.gauss_sigma = gauss_sigma; P.label_threshold = hinge_threshold;
    P.inv_feat_stride = 1.0f / feat_stride;
    P.step_length = step_length; P.reg_weight = reg_weight; P.alpha_eps = alpha_eps;
    P.iterates_out = iterates_out; P.losses_out = losses_out;
    if (H == 18) return launch_sd<18, 2>(P, st);
    return launch_sd<22, 2>(P, st);
}

extern "C" int b200trk_gn_sd_hinge(const float* weights, float* weights_out, const float* feat, const float* train_label,
                                   const float* sample_weight, int n, int C, int H, int W, int k, int num_iter,
                                   float filter_reg, float hinge_threshold, float activation_leak, int score_act, float act_param,
                                   float steplength_reg, float* iterates_out, float* losses_out, b200trk_stream_t stream) {
    if (int e = check_common("gn_sd_hinge", weights, weights_out, feat, train_label, n, C, H, W, k, num_iter)) return e;
    B200_REQUIRE(score_act == 0 || score_act == 1, "gn_sd_hinge: score_act must be 0 (relu) or 1 (bentpar)");
    cudaStream_t st = (cudaStream_t)stream;
    if (iterates_out)
        B200_CHECK_CUDA(cudaMemcpyAsync(iterates_out, weights, (size_t)C * 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SdParams P{};
    P.w_in = weights; P.w_out = weights_out; P.feat = feat; P.bb = train_label /* unused boxes: any valid pointer */;
    P.sample_weight = sample_weight; P.n = n; P.C = C; P.num_iter = num_iter;
    P.label_in = train_label; P.label_threshold = hinge_threshold; P.act_leak = activation_leak; P.act_kind = score_act;
    P.act_b = act_param;
    P.inv_feat_stride = 1.f / 16.f;
    P.step_length = 1.f; P.reg_weight = filter_reg * filter_reg; P.alpha_eps = steplength_reg;
    P.loss_scale = 1.f / ((float)n * (float)((H + 1) * (W + 1)) + (float)C * 16.f);   // GNSteepestDescent._compute_loss: mean over all residual entries
    P.iterates_out = iterates_out; P.losses_out = losses_out;
    if (H == 18) return launch_sd<18, 3>(P, st);
    return launch_sd<22, 3>(P, st);
}
