// The two device kernels of the whole-frame DiMP call (dimp_tracker.cu): the crop sampler of pytracking/features/preprocessing.py:55-148
// and localize_target / localize_advanced (pytracking/tracker/dimp/dimp.py:196-303).  Plain SIMT CUDA C with explicitly rounded
// arithmetic, kept in a header of their own so that the SAME source also compiles as host code under tests/cpu_emul/cuda_shim.h:
// tests/test_dimp_kernels_cpu.py runs the sampler against torch-CPU bilinear interpolation (bit-exact) and the localisation kernel
// against the decisions recorded from the unmodified reference tracker, on the CPU.  Included by dimp_tracker.cu only.
#pragma once

// ---------------------------------------------------------------------------------------------------------------------------
// sample_patch on the device
// ---------------------------------------------------------------------------------------------------------------------------
namespace {

// ATen's linear source index / weights (aten/src/ATen/native/UpSample.h area_pixel_compute_source_index, guard_index_and_lambda;
// cpu/UpSampleKernel.cpp HelperInterpLinear), in the fused form gcc emits for the AVX2 / AVX512 builds of torch:
//   src = fma(scale, dst + 0.5, -0.5), clamped at 0;   value = fma(t0, w0, t1 * w1)
__device__ __forceinline__ void linear_src(int o, int in_size, int out_size, float scale, int& i0, int& i1, float& w0, float& w1) {
    if (in_size == out_size) { i0 = i1 = o; w0 = 1.f; w1 = 0.f; return; }
    float src = __fmaf_rn(scale, __fadd_rn((float)o, 0.5f), -0.5f);
    src = src < 0.f ? 0.f : src;
    int f = (int)floorf(src);
    i0 = f < in_size - 1 ? f : in_size - 1;
    float l1 = __fsub_rn(src, (float)i0);
    l1 = fminf(fmaxf(l1, 0.f), 1.f);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    w0 = __fsub_rn(1.f, l1);
    w1 = l1;
}

__global__ void __launch_bounds__(256) sample_patch_kernel(const uint8_t* __restrict__ im, int H, int W, b200trk_crop_geom_t g,
                                                           int win_h, int win_w, float* __restrict__ out) {
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
    const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ox >= win_w || oy >= win_h) return;
    const int H2 = (H - g.os_r + g.df - 1) / g.df, W2 = (W - g.os_c + g.df - 1) / g.df;      // im[..., os::df, os::df]
    const float sh = __fdiv_rn((float)g.in_h, (float)g.out_h), sw = __fdiv_rn((float)g.in_w, (float)g.out_w);
    int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
    linear_src(oy + g.win_r, g.in_h, g.out_h, sh, y0, y1, wy0, wy1);
    linear_src(ox + g.win_c, g.in_w, g.out_w, sw, x0, x1, wx0, wx1);
    // patch row p -> decimated image row clamp(tl + p) (F.pad 'replicate' / negative pad = crop) -> image row os + df * r
    auto row = [&](int p) { int r = g.tl_r + p; r = r < 0 ? 0 : (r > H2 - 1 ? H2 - 1 : r); return g.os_r + g.df * r; };
    auto col = [&](int p) { int c = g.tl_c + p; c = c < 0 ? 0 : (c > W2 - 1 ? W2 - 1 : c); return g.os_c + g.df * c; };
    const uint8_t* r0 = im + (size_t)row(y0) * W * 3;
    const uint8_t* r1 = im + (size_t)row(y1) * W * 3;
    const int c0 = col(x0) * 3, c1 = col(x1) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float p00 = (float)r0[c0 + ch], p01 = (float)r0[c1 + ch], p10 = (float)r1[c0 + ch], p11 = (float)r1[c1 + ch];
        float v;
        if (g.in_h == g.out_h && g.in_w == g.out_w) {
            v = p00;                                                   // preprocessing.py:142-143: no resampling
        } else {
            const float t0 = __fmaf_rn(p00, wx0, __fmul_rn(p01, wx1));
            const float t1 = __fmaf_rn(p10, wx0, __fmul_rn(p11, wx1));
            v = __fmaf_rn(t0, wy0, __fmul_rn(t1, wy1));
        }
        out[((size_t)ch * win_h + oy) * win_w + ox] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// localize_target / localize_advanced on the device (one CTA)
// ---------------------------------------------------------------------------------------------------------------------------
struct LocArgs {
    int S, Ho, Wo, advanced;
    double not_found_thr, uncertain_thr, hard_sample_thr;
    float distractor_thr, hard_negative_thr, not_found_thr_f, disp_thr;
    float neigh[8][2], prev_vec[8][2];
};

struct AM { float v; int r, c; };
// dcf.max2d order (pytracking/libs/dcf.py:156-164): the column of the maximum first (smallest on ties), then the smallest row
__device__ __forceinline__ bool am_better(const AM& a, const AM& b) {
    if (a.v != b.v) return a.v > b.v;
    if (a.c != b.c) return a.c < b.c;
    return a.r < b.r;
}
__device__ AM block_argmax(AM m, AM* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        AM t; t.v = __shfl_xor_sync(0xffffffffu, m.v, o); t.r = __shfl_xor_sync(0xffffffffu, m.r, o); t.c = __shfl_xor_sync(0xffffffffu, m.c, o);
        if (am_better(t, m)) m = t;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    AM r = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) if (am_better(red[w], r)) r = red[w];
    return r;
}

__global__ void __launch_bounds__(256) localize_kernel(const float* __restrict__ scores, LocArgs a, b200trk_loc_result_t* __restrict__ res) {
    __shared__ AM red[8];
    const int n = a.Ho * a.Wo;
    const AM none = {__int_as_float(0xff800000), 1 << 30, 1 << 30};
    // max_score1, max_disp1 per scale; scale_ind = first scale with the largest maximum (torch.max(max_score1, dim=0))
    AM best1 = none; int scale_ind = 0;
    for (int s = 0; s < a.S; ++s) {
        AM m = none;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            AM t = {scores[(size_t)s * n + i], i / a.Wo, i % a.Wo};
            if (am_better(t, m)) m = t;
        }
        m = block_argmax(m, red);
        if (s == 0 || m.v > best1.v) { best1 = m; scale_ind = s; }
    }
    b200trk_loc_result_t out;
    out.flag = 0; out.scale_ind = scale_ind; out.r1 = best1.r; out.c1 = best1.c; out.r2 = -1; out.c2 = -1; out.use_second = 0;
    out.score1 = best1.v; out.score2 = 0.f; out.max_score = best1.v;
    for (int i = 0; i < 6; ++i) out.pad_[i] = 0;
    bool done = !a.advanced;
    if (!done) {
        const double s1 = (double)best1.v;
        if (s1 < a.not_found_thr) { out.flag = 4; done = true; }
        else if (s1 < a.uncertain_thr) { out.flag = 3; done = true; }
        else if (s1 < a.hard_sample_thr) { out.flag = 2; done = true; }
    }
    if (!done) {
        // mask out the target neighbourhood (dimp.py:267-274): Python round() of doubles = round-half-even = rint()
        const double n0 = (double)a.neigh[scale_ind][0], n1 = (double)a.neigh[scale_ind][1];
        const int top = max((int)rint((double)best1.r - n0 / 2), 0), bottom = min((int)rint((double)best1.r + n0 / 2 + 1), a.Ho);
        const int left = max((int)rint((double)best1.c - n1 / 2), 0), right = min((int)rint((double)best1.c + n1 / 2 + 1), a.Wo);
        AM m = none;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int r = i / a.Wo, c = i % a.Wo;
            const bool inside = r >= top && r < bottom && c >= left && c < right;
            AM t = {inside ? 0.f : scores[(size_t)scale_ind * n + i], r, c};
            if (am_better(t, m)) m = t;
        }
        m = block_argmax(m, red);
        out.r2 = m.r; out.c2 = m.c; out.score2 = m.v;
        const float cy = __fdiv_rn((float)(a.Ho - 1), 2.f), cx = __fdiv_rn((float)(a.Wo - 1), 2.f);     // score_center = (score_sz - 1)/2
        const float d1y = __fsub_rn((float)best1.r, cy), d1x = __fsub_rn((float)best1.c, cx);
        const float d2y = __fsub_rn((float)m.r, cy), d2x = __fsub_rn((float)m.c, cx);
        const float pvy = a.prev_vec[scale_ind][0], pvx = a.prev_vec[scale_ind][1];
        if (m.v > __fmul_rn(a.distractor_thr, best1.v)) {
            const float e1y = __fsub_rn(d1y, pvy), e1x = __fsub_rn(d1x, pvx), e2y = __fsub_rn(d2y, pvy), e2x = __fsub_rn(d2x, pvx);
            const float n1f = __fsqrt_rn(__fadd_rn(__fmul_rn(e1y, e1y), __fmul_rn(e1x, e1x)));
            const float n2f = __fsqrt_rn(__fadd_rn(__fmul_rn(e2y, e2y), __fmul_rn(e2x, e2x)));
            if (n2f > a.disp_thr && n1f < a.disp_thr) out.flag = 2;
            else if (n2f < a.disp_thr && n1f > a.disp_thr) { out.flag = 2; out.use_second = 1; }
            else out.flag = 3;
        } else if (m.v > __fmul_rn(a.hard_negative_thr, best1.v) && m.v > a.not_found_thr_f) {
            out.flag = 2;
        } else {
            out.flag = 1;
        }
    }
    if (threadIdx.x == 0) *res = out;
}

// LocArgs from the tracker parameters, as b200trk_dimp_localize passes them to localize_kernel (host)
inline LocArgs make_loc_args(int S, int Ho, int Wo, const b200trk_dimp_params_t* p, const float* neigh, const float* prev_vec) {
    LocArgs a;
    a.S = S; a.Ho = Ho; a.Wo = Wo; a.advanced = p->advanced_localization;
    a.not_found_thr = p->target_not_found_threshold; a.uncertain_thr = p->uncertain_threshold; a.hard_sample_thr = p->hard_sample_threshold;
    a.distractor_thr = (float)p->distractor_threshold; a.hard_negative_thr = (float)p->hard_negative_threshold;
    a.not_found_thr_f = (float)p->target_not_found_threshold;
    a.disp_thr = (float)(p->dispalcement_scale * std::sqrt((double)(Ho * Wo)) / 2);               // dimp.py:287
    for (int s = 0; s < 8; ++s)
        for (int i = 0; i < 2; ++i) {
            a.neigh[s][i] = (s < S && neigh) ? neigh[2 * s + i] : 0.f;
            a.prev_vec[s][i] = (s < S && prev_vec) ? prev_vec[2 * s + i] : 0.f;
        }
    return a;
}

}  // namespace
